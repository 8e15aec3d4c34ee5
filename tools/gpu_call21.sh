set -x
timeout 600 python -m pytest tests/test_host_cpp.py -m gpu -q -s > gpurun_out/pytest21.log 2>&1; tail -12 gpurun_out/pytest21.log
