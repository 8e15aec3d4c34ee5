import json, os, subprocess, sys
sys.path.insert(0, '.')
from ipt_b200 import build
variants = {
    "lb1": ["IPT_LEAF_BATCH=1"], "lb2": ["IPT_LEAF_BATCH=2"], "lb4": ["IPT_LEAF_BATCH=4"], "lb6": ["IPT_LEAF_BATCH=6"], "lb8": ["IPT_LEAF_BATCH=8"],
    "lb4_st8": ["IPT_LEAF_BATCH=4", "IPT_TRAV_STEPS=8"], "lb4_st16": ["IPT_LEAF_BATCH=4", "IPT_TRAV_STEPS=16"], "lb4_rf4": ["IPT_LEAF_BATCH=4", "IPT_REFILL_MIN=4"],
    "lb4_rf12": ["IPT_LEAF_BATCH=4", "IPT_REFILL_MIN=12"],
    "a": ["IPT_LEAF_BATCH=6", "IPT_TRAV_STEPS=16", "IPT_REFILL_MIN=4"], "b": ["IPT_LEAF_BATCH=4", "IPT_TRAV_STEPS=24", "IPT_REFILL_MIN=4"],
    "c": ["IPT_LEAF_BATCH=6", "IPT_TRAV_STEPS=24", "IPT_REFILL_MIN=2"], "d": ["IPT_LEAF_BATCH=4", "IPT_TRAV_STEPS=16", "IPT_REFILL_MIN=2"],
    "e": ["IPT_LEAF_BATCH=6", "IPT_TRAV_STEPS=32", "IPT_REFILL_MIN=4"],
    "mb4": [], "mb5": ["IPT_MESH_MIN_BLOCKS=5"], "mb6": ["IPT_MESH_MIN_BLOCKS=6"], "mb5_st24": ["IPT_MESH_MIN_BLOCKS=5", "IPT_TRAV_STEPS=24"],
}
sel = sys.argv[1].split(",") if len(sys.argv) > 1 else list(variants)
for name in sel:
    so = build.build_variant("mesh_" + name, variants[name])
    env = dict(os.environ, IPT_B200_LIB=str(so))
    r = subprocess.run([sys.executable, "tools/run_configs.py", "c3_tree,c3"], env=env, capture_output=True, text=True)
    try:
        for line in r.stdout.strip().splitlines()[-2:]:
            d = json.loads(line)
            print(f"{name:14s} {d['config']:8s} {d['mrays_per_s']:8.1f} Mrays/s  ext {d['ms_extend']:7.1f} ms", flush=True)
    except Exception as e:
        print(name, "failed", r.stderr[-300:])
