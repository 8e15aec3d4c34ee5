set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "lightgrid or mixedlights or light or c5 or emitters" > gpurun_out/pytest20.log 2>&1; tail -4 gpurun_out/pytest20.log
timeout 300 python tools/ab_r02.py configs lightbin,base c5_100,c5 > gpurun_out/ab20_light_wide.log 2>&1; cat gpurun_out/ab20_light_wide.log
