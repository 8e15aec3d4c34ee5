"""Small end-to-end run for compute-sanitizer (memcheck): every kernel family on tiny inputs."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from ipt_b200 import capi
for name, kw in [("box", dict(width=32, height=24, pass_count=2)), ("cornell", dict(width=24, height=24, pass_count=1)),
                 ("smallpt", dict(width=16, height=16, pass_count=1, schedule=[4, 2, 1, 1])), ("mesh:3000", dict(width=24, height=24, pass_count=1, schedule=[4, 2, 2, 1])),
                 ("lightgrid:4x4", dict(width=16, height=16, pass_count=1, schedule=[4, 2, 1, 1])), ("mesh:1", dict(width=8, height=8, pass_count=1))]:
    sd = capi.SceneDescription(name); sc = capi.Scene(sd)
    s, q, c, st = sc.render_host(capi.default_params(**kw))
    rng = np.random.default_rng(0)
    o = rng.uniform(-0.9, 0.9, (500, 3)).astype(np.float32); d = rng.normal(size=(500, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    sc.trace_batch(o, d); sc.preview_batch(o, d); sc.camera_rays(rng.random((100, 2)).astype(np.float32))
    sc.ddf_sample(2, 100, to=[0, 0, 1]); sc.ddf_value(40, d, to=[0.6, 0, 0.8]); sc.light_ddf_value([0, 0, -1], d); sc.light_ddf_sample([0, 0, -1], 100)
    co, cd = sc.camera_rays(np.array([[0.5, 0.35]], np.float32))
    try:
        sc.mix_sample(co[0], cd[0], 200)
    except capi.IptError:
        pass
    pl = capi.Plane(sc, 16, 16); pl.add_rays(rng.random(50), rng.random(50), rng.random(50)); pl.resolve(); pl.close()
    if name.startswith("mesh"):
        sc.bvh_export()
    print(name, "ok", st.rays, flush=True)
    sc.close()
