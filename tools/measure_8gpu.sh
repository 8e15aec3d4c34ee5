# Eight GPUs of one box: the bench line at N = 8 (BASELINE configs[1], weak scaling) with the BASELINE configs[3] sub-record
# (10 M triangles, 3840x2160, pass-sharded), and N = 1 on the same box for the denominators.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 3 --no-cpu-baseline --with-c4 > gpurun_out/scale_r02_1gpu_same_box.json 2> gpurun_out/scale_r02_1gpu.err; tail -c 300 gpurun_out/scale_r02_1gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 64 --warmup 3 > gpurun_out/scale_r02_8gpu.json 2> gpurun_out/scale_r02_8gpu.err; tail -c 400 gpurun_out/scale_r02_8gpu.err
python - <<'PY'
import json
for f in ('gpurun_out/scale_r02_1gpu_same_box.json','gpurun_out/scale_r02_8gpu.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','device_ms_per_step_ranks','collective_ms','host_overhead_ms_per_step')}, d['e2e']['value'])
    c=d.get('c4'); print('c4', c and {k:c.get(k) for k in ('n_gpus','value','mrays_per_s','ms_per_step','collective_ms','device_ms_per_step_ranks')}, c and c['e2e'])
PY
