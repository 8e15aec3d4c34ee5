# Round-2 GPU call 1: parity suite with statistics, A/B of the fused-kernel variants, config sweep.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/smi1.txt
python -m pytest tests -m gpu -q -s > gpurun_out/pytest1.log 2>&1; tail -3 gpurun_out/pytest1.log
grep -E "IMAGE_STATS|FAILED|^E  " gpurun_out/pytest1.log | cut -c1-260 | head -60
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tools/ab_r02.py run all c2 > gpurun_out/ab1.log 2>&1; cat gpurun_out/ab1.log
python tools/ab_r02.py run base,r7,nearr1 c1 > gpurun_out/ab1_c1.log 2>&1; cat gpurun_out/ab1_c1.log
for b in 1048576 4194304; do IPT_B200_LIB=ipt_b200/lib/variants/base.so python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 --batch-paths $b 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch', $b, d['value'])"; done
python tools/run_configs.py c1,c2,c3,c3_tree,c5_100,c5 > gpurun_out/configs1.jsonl 2>&1; cut -c1-400 gpurun_out/configs1.jsonl
