set -x
mkdir -p gpurun_out
for v in onelane base; do
IPT_B200_LIB=ipt_b200/lib/variants/$v.so timeout 300 python bench.py --workload c4 --passes-per-step 8 --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-c4 > gpurun_out/c4x8_$v.json 2>> gpurun_out/bench_err17.log
IPT_B200_LIB=ipt_b200/lib/variants/$v.so timeout 300 python bench.py --workload c3 --passes-per-step 32 --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-c4 > gpurun_out/c3x32_$v.json 2>> gpurun_out/bench_err17.log
done
python - <<'PY'
import json
for w in ('c4x8','c3x32'):
  for v in ('onelane','base'):
    d=json.loads(open(f'gpurun_out/{w}_{v}.json').read().strip().splitlines()[-1]); print(w, v, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2))
PY
