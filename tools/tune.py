"""Kernel tuning sweep: builds variants of the same sources and benches each (GPU box only)."""
import json, os, subprocess, sys
sys.path.insert(0, '.')
from ipt_b200 import build
variants = {  # -D flags of ipt_b200/csrc; the adopted values are the defaults in the sources (profiles/tuning_r01.md)
    "base": [],
    "fused2": ["IPT_SHADE_FUSED_MIN_BLOCKS=2"], "fused4": ["IPT_SHADE_FUSED_MIN_BLOCKS=4"],
    "next2": ["IPT_SHADE_NEXT_MIN_BLOCKS=2"], "next4": ["IPT_SHADE_NEXT_MIN_BLOCKS=4"],
    "n4f4": ["IPT_SHADE_FUSED_MIN_BLOCKS=4", "IPT_SHADE_NEXT_MIN_BLOCKS=4"], "n2f2": ["IPT_SHADE_FUSED_MIN_BLOCKS=2", "IPT_SHADE_NEXT_MIN_BLOCKS=2"],
    "shade3": ["IPT_SHADE_MIN_BLOCKS=3"], "ext2": ["IPT_EXTEND_MIN_BLOCKS=2"], "ext4": ["IPT_EXTEND_MIN_BLOCKS=4"],
    "lights4": ["IPT_INLINE_LIGHTS=4"], "accurate_log2": ["IPT_LOBE_LOG2=log2f"],
}
sel = sys.argv[1].split(",") if len(sys.argv) > 1 else list(variants)
batches = [int(b) for b in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0"])]
for name in sel:
    so = build.build_variant(name, variants[name])
    for b in batches:
        env = dict(os.environ, IPT_B200_LIB=str(so))
        r = subprocess.run([sys.executable, "bench.py", "--steps", "12", "--warmup", "3", "--no-cpu-baseline", "--batch-paths", str(b)], env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            k = d["roofline"]["kernel_ms"]
            print(f"{name:18s} batch {b:8d}  {d['value']:7.1f} Mpaths/s  extend {k['extend']:7.1f}  shade {k['shade']:7.1f}  sm {d['clocks']['sm_mhz']}", flush=True)
        except Exception as e:
            print(name, b, "failed", r.stderr[-400:])
