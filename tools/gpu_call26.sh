set -x
timeout 900 python tools/ab_r02.py configs prev,base c3,c3_tree,c4 > gpurun_out/ab26_mesh_votes.log 2>&1; cat gpurun_out/ab26_mesh_votes.log | cut -c1-100
timeout 600 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest26.log 2>&1; tail -3 gpurun_out/pytest26.log
