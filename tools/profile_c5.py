"""Short many-light run for ncu: `python tools/profile_c5.py [rows] [crop]`."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ipt_b200 import capi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100
crop = int(sys.argv[2]) if len(sys.argv) > 2 else 512
sd = capi.SceneDescription(f"lightgrid:{rows}x{rows}"); sc = capi.Scene(sd); pl = capi.Plane(sc, 2048, 2048)
kw = dict(width=2048, height=2048, pass_count=1, tile_x0=768, tile_y0=768, tile_w=crop, tile_h=crop)
pl.render(capi.default_params(**kw))
st = pl.render(capi.default_params(flags=capi.FLAG_TIME_KERNELS, pass_begin=1, **kw))
print('lightgrid', rows, crop, 'ms', st.ms_total, 'shade', st.ms_shade, 'Mpaths/s', st.paths / st.ms_total / 1e3, 'rays/path', st.rays / st.paths)
