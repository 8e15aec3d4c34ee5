set -x
mkdir -p gpurun_out
timeout 300 python tools/diag_smallpt.py smallpt 24 > gpurun_out/diag_smallpt2.log 2>&1; cat gpurun_out/diag_smallpt2.log
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py -m gpu -q -s -k "smallpt" > gpurun_out/pytest5.log 2>&1; tail -12 gpurun_out/pytest5.log
grep -hE "IMAGE_STATS|FAILED|^E  " gpurun_out/pytest5.log | cut -c1-420 | head -30
