import json, os, subprocess, sys
sys.path.insert(0, '.')
from ipt_b200 import build
variants = {
    "lb1": ["IPT_LEAF_BATCH=1"], "lb4": ["IPT_LEAF_BATCH=4"], "lb10": [], "lb16": ["IPT_LEAF_BATCH=16"], "lb24": ["IPT_LEAF_BATCH=24"],
    "lb4_s12": ["IPT_LEAF_BATCH=4", "IPT_STACK_SHORT=12"], "lb1_s12_st4": ["IPT_LEAF_BATCH=1", "IPT_STACK_SHORT=12", "IPT_TRAV_STEPS=4"],
    "lb10_st12": ["IPT_TRAV_STEPS=12"], "lb16_st16": ["IPT_LEAF_BATCH=16", "IPT_TRAV_STEPS=16"], "lb10_rf4": ["IPT_REFILL_MIN=4"], "lb10_rf16": ["IPT_REFILL_MIN=16"],
}
sel = sys.argv[1].split(",") if len(sys.argv) > 1 else list(variants)
for name in sel:
    so = build.build_variant("mesh_" + name, variants[name])
    env = dict(os.environ, IPT_B200_LIB=str(so))
    r = subprocess.run([sys.executable, "tests/run_configs.py", "c3_tree"], env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{name:14s} {d['mrays_per_s']:8.1f} Mrays/s  ext {d['ms_extend']:7.1f} ms", flush=True)
    except Exception as e:
        print(name, "failed", r.stderr[-300:])
