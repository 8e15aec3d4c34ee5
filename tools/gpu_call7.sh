set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_fullsize.py -m gpu -q -s -x --durations=5 > gpurun_out/pytest7.log 2>&1; tail -8 gpurun_out/pytest7.log
grep -hE "IMAGE_STATS|C3_CRN|FAILED|^E  " gpurun_out/pytest7.log | cut -c1-300 | head -20
timeout 600 python tools/ab_r02.py configs meshv1,base c3,c3_tree,c4 > gpurun_out/ab7_mesh.log 2>&1; cat gpurun_out/ab7_mesh.log
timeout 300 python tools/ab_r02.py configs base c5_100,c5 > gpurun_out/ab7_lights.log 2>&1; cat gpurun_out/ab7_lights.log
timeout 300 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py -m gpu -q -k "lightgrid or mixedlights or light" > gpurun_out/pytest7b.log 2>&1; tail -4 gpurun_out/pytest7b.log
timeout 300 python tools/ab_r02.py configs bounds c3,c5 > gpurun_out/ab7_bounds.log 2>&1; cat gpurun_out/ab7_bounds.log
