# N GPUs of one box (bash tools/measure_ngpu.sh N): the bench line at N (BASELINE configs[1], weak scaling) with the BASELINE
# configs[3] sub-record (10 M triangles, 3840x2160, pass-sharded).
N=${1:-4}
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 64 --warmup 3 > gpurun_out/scale_r02_${N}gpu.json 2> gpurun_out/scale_r02_${N}gpu.err; tail -c 300 gpurun_out/scale_r02_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/scale_r02_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','device_ms_per_step_ranks','collective_ms','host_overhead_ms_per_step')}, d['e2e']['value'])
c=d.get('c4'); print('c4', c and {k:c.get(k) for k in ('n_gpus','value','mrays_per_s','ms_per_step','collective_ms')}, c and c['e2e'])
PY
