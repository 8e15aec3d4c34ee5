set -x
timeout 600 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest28.log 2>&1; tail -3 gpurun_out/pytest28.log
timeout 900 python tools/ab_r02.py configs nospec,base,spec_l8,spec_l12 c3,c3_tree,c4 > gpurun_out/ab28_mesh_spec.log 2>&1; cat gpurun_out/ab28_mesh_spec.log | cut -c1-150
