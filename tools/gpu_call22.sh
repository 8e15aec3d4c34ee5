# the whole GPU parity suite on the -DIPT_DEBUG_BOUNDS build: any queue append beyond its capacity makes ipt_render fail
set -x
IPT_B200_LIB=ipt_b200/lib/variants/bounds.so timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_host_cpp.py > gpurun_out/pytest22_bounds.log 2>&1; tail -4 gpurun_out/pytest22_bounds.log
