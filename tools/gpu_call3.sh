# Round-2 GPU call 3: smallpt image per arithmetic variant (diagnosis of the bias the z-test saw), bench line (host overhead),
# mesh A/B after the zero-component fix, queue bounds-check build on every workload, then the parity suite.
set -x
mkdir -p gpurun_out
for v in base nofast; do IPT_B200_LIB=ipt_b200/lib/variants/$v.so timeout 120 python tools/dump_image.py smallpt 64 64 16384 gpurun_out/smallpt_$v.npz; done
timeout 300 python bench.py --steps 24 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -c 600 gpurun_out/bench3.err; python -c "
import json; d=json.loads(open('gpurun_out/bench3.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','device_ms_per_step','host_overhead_ms_per_step')}, d['e2e'], d['roofline']['frac'])"
timeout 300 python tools/ab_r02.py configs base,wide c3,c3_tree > gpurun_out/ab3_mesh.log 2>&1; cat gpurun_out/ab3_mesh.log
timeout 300 python tools/ab_r02.py configs bounds c1,c2,c3,c5_100,c5 > gpurun_out/ab3_bounds.log 2>&1; cat gpurun_out/ab3_bounds.log
timeout 600 python -m pytest tests/test_gpu_mesh.py -m gpu -q -s --durations=8 > gpurun_out/pytest3m.log 2>&1; tail -15 gpurun_out/pytest3m.log
timeout 900 python -m pytest tests -m gpu -q -s --durations=12 --deselect tests/test_gpu_mesh.py > gpurun_out/pytest3a.log 2>&1; tail -18 gpurun_out/pytest3a.log
grep -hE "IMAGE_STATS|C3_CRN|FAILED|^E  " gpurun_out/pytest3a.log gpurun_out/pytest3m.log | cut -c1-420 | head -60
