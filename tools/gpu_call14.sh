set -x
mkdir -p gpurun_out
for bp in 8388608 16777216 33554432; do
timeout 300 python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-c4 --batch-paths $bp > gpurun_out/c4_bp_$bp.json 2>> gpurun_out/bench_err14.log
timeout 300 python bench.py --workload c3 --passes-per-step 16 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-c4 --batch-paths $bp > gpurun_out/c3_bp_$bp.json 2>> gpurun_out/bench_err14.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c[34]_bp_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), 'Mpaths/s', round(d['mrays_per_s'],1), 'Mrays/s', 'ms/step', round(d['ms_per_step'],2), 'launches', d['gpu_launches'])
    except Exception as e: print(f, 'failed', e)
PY
tail -3 gpurun_out/bench_err14.log
