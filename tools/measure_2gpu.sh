# Two GPUs of one box: the C++ multi-device entry on two real devices (host_demo / ab_dropin through the tests), the
# product's NCCL collective through the C ABI, and the bench line at N = 2 (with the BASELINE configs[3] sub-record).
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_host_cpp.py -m gpu -q -s > gpurun_out/pytest9_hostcpp.log 2>&1; tail -5 gpurun_out/pytest9_hostcpp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_nccl_abi.py > gpurun_out/nccl_abi_2gpu.log 2>&1; tail -4 gpurun_out/nccl_abi_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 32 --warmup 3 > gpurun_out/scale_r02_2gpu.json 2> gpurun_out/scale_r02_2gpu.err; tail -c 400 gpurun_out/scale_r02_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/scale_r02_2gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','device_ms_per_step_ranks','collective_ms','host_overhead_ms_per_step')}, d['e2e']['value'])
c=d.get('c4'); print('c4', c and {k:c.get(k) for k in ('n_gpus','value','mrays_per_s','ms_per_step','single_gpu_value','collective_ms')}, c and c['e2e'])
PY
