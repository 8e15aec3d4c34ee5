"""Renders a sample scene on the GPU and stores sum / sumsq / count (diagnostics of the image z-tests):
    python tools/dump_image.py <scene> <width> <height> <passes> <out.npz> [seed]
IPT_B200_LIB selects a tuning variant of the library (tools/ab_r02.py)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from ipt_b200 import capi

scene, W, H, passes, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
seed = int(sys.argv[6]) if len(sys.argv) > 6 else 2024
capi.load()
sd = capi.SceneDescription(scene)
sc = capi.Scene(sd)
s, q, c, st = sc.render_host(capi.default_params(width=W, height=H, pass_count=passes, seed=seed))
np.savez(out, sum=s, sumsq=q, count=c, passes=passes, rays=st.rays, paths=st.paths)
print(scene, "mean", s.sum() / max(c.sum(), 1), "rays/path", st.rays / st.paths, "ms", st.ms_total)
