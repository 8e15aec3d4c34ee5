"""Short run of one bench workload for ncu: `python tools/profile_run.py <workload> [passes]` with workload = c1 .. c5 of
bench.py (scene, frame, depth and split schedule of that BASELINE config; c5 renders its 512x512 crop), or the old form
`python tools/profile_run.py <scene> <W> <H> <passes>`. One warm-up pass (workspace allocation), then `passes` passes."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench
from ipt_b200 import capi

if sys.argv[1] in bench.WORKLOADS:
    w = bench.WORKLOADS[sys.argv[1]]
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    name, W, H = w["scene"], w["width"], w["height"]
    kw = dict(width=W, height=H, depth_max=w["depth_max"], schedule=w["schedule"])
    if w.get("tile"):
        t = w["tile"]; kw.update(tile_x0=t[0], tile_y0=t[1], tile_w=t[2], tile_h=t[3])
else:
    name = sys.argv[1]; W = int(sys.argv[2]); H = int(sys.argv[3]); passes = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    kw = dict(width=W, height=H)
sd = capi.SceneDescription(name); sc = capi.Scene(sd); pl = capi.Plane(sc, W, H)
pl.render(capi.default_params(pass_count=1, **kw))
st = pl.render(capi.default_params(pass_begin=1, pass_count=passes, **kw))
print(name, W, H, passes, 'ms', st.ms_total, 'paths', st.paths, 'rays', st.rays, 'launches', st.kernel_launches, 'Mpaths/s', st.paths / st.ms_total / 1e3)
