set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest16.log 2>&1; tail -3 gpurun_out/pytest16.log
timeout 600 python tools/ab_r02.py configs onelane,base c3,c3_tree,c4 > gpurun_out/ab16_lanes.log 2>&1; cat gpurun_out/ab16_lanes.log
for v in onelane base; do IPT_B200_LIB=ipt_b200/lib/variants/$v.so timeout 300 python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 4 --no-c4 > gpurun_out/c4_$v.json 2>> gpurun_out/bench_err16.log; done
python - <<'PY'
import json
for v in ('onelane','base'):
    d=json.loads(open(f'gpurun_out/c4_{v}.json').read().strip().splitlines()[-1]); print(v, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2), 'dev ms', round(d['device_ms_per_step'],2), 'kernel_ms', d['roofline']['kernel_ms'])
PY
