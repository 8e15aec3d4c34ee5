import json, os, subprocess, sys
names = sys.argv[1].split(",")
for rep in range(2):
    for name in names:
        env = dict(os.environ, IPT_B200_LIB=f"ipt_b200/lib/variants/{name}.so")
        r = subprocess.run([sys.executable, "tools/run_configs.py", "c1,c2"], env=env, capture_output=True, text=True)
        out = [json.loads(l) for l in r.stdout.strip().splitlines()]
        print(name, " ".join(f"{d['config']} {d['mpaths_per_s']:.1f} (ext {d['ms_extend']:.1f} sh {d['ms_shade']:.1f})" for d in out), flush=True)
