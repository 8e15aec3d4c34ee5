set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest15.log 2>&1; tail -3 gpurun_out/pytest15.log
timeout 600 python tools/run_configs.py c3,c3_tree,c4 > gpurun_out/configs15.jsonl 2>&1; cut -c1-230 gpurun_out/configs15.jsonl
