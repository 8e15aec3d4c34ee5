set -x
mkdir -p gpurun_out
timeout 300 python tools/diag_smallpt.py smallpt 24 > gpurun_out/diag_smallpt.log 2>&1; cat gpurun_out/diag_smallpt.log
timeout 600 python -m pytest tests/test_gpu_mesh.py tests/test_host_cpp.py -m gpu -q -s --durations=5 > gpurun_out/pytest4.log 2>&1; tail -12 gpurun_out/pytest4.log
grep -hE "IMAGE_STATS|C3_CRN|FAILED|^E  " gpurun_out/pytest4.log | cut -c1-420 | head -30
