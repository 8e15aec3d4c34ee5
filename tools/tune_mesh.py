import json, os, subprocess, sys
sys.path.insert(0, '.')
from ipt_b200 import build
variants = {
    "mb4": [], "mb5": ["IPT_MESH_MIN_BLOCKS=5"], "mb6": ["IPT_MESH_MIN_BLOCKS=6"], "mb5_st24": ["IPT_MESH_MIN_BLOCKS=5", "IPT_TRAV_STEPS=24"],
}
sel = sys.argv[1].split(",") if len(sys.argv) > 1 else list(variants)
for name in sel:
    so = build.build_variant("mesh_" + name, variants[name])
    env = dict(os.environ, IPT_B200_LIB=str(so))
    r = subprocess.run([sys.executable, "tools/run_configs.py", "c3_tree"], env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{name:14s} {d['mrays_per_s']:8.1f} Mrays/s  ext {d['ms_extend']:7.1f} ms", flush=True)
    except Exception as e:
        print(name, "failed", r.stderr[-300:])
