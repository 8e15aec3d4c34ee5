"""Short run of the hot path for ncu: `python tools/profile_run.py [scene] [W] [H] [passes]`."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ipt_b200 import capi
name = sys.argv[1] if len(sys.argv) > 1 else 'cornell'
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
H = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
passes = int(sys.argv[4]) if len(sys.argv) > 4 else 1
sd = capi.SceneDescription(name); sc = capi.Scene(sd); pl = capi.Plane(sc, W, H)
st = pl.render(capi.default_params(width=W, height=H, pass_count=passes))
print(name, W, H, passes, 'ms', st.ms_total, 'paths', st.paths, 'rays', st.rays, 'launches', st.kernel_launches)
