set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -m gpu -q -k "not mesh and not million" > gpurun_out/pytest12.log 2>&1; tail -4 gpurun_out/pytest12.log
timeout 300 python tools/ab_r02.py run walls,base c2,c1 > gpurun_out/ab12_walls.log 2>&1; cat gpurun_out/ab12_walls.log
