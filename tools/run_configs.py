"""Runs every BASELINE.json config briefly on one GPU and prints one JSON line per config (rates are pass-count
independent). `python tools/run_configs.py [c1,c2,c3,c4,c5]`"""
import json, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ipt_b200 import capi

CONFIGS = {
    "c1": dict(scene="box", width=640, height=640, depth_max=4, schedule=[16, 8, 4, 2], passes=64),
    "c2": dict(scene="cornell", width=1024, height=1024, depth_max=4, schedule=[16, 8, 4, 2], passes=32),
    "c3": dict(scene="mesh:1000000", width=1920, height=1080, depth_max=8, schedule=[1] * 8, passes=16),
    "c3_tree": dict(scene="mesh:1000000", width=1920, height=1080, depth_max=4, schedule=[16, 8, 4, 2], passes=2),
    "c4": dict(scene="mesh:10000000", width=3840, height=2160, depth_max=8, schedule=[1] * 8, passes=4),
    "c5_100": dict(scene="lightgrid:10x10", width=2048, height=2048, depth_max=4, schedule=[16, 8, 4, 2], passes=16, tile=(768, 768, 512, 512)),
    "c5": dict(scene="lightgrid:100x100", width=2048, height=2048, depth_max=4, schedule=[16, 8, 4, 2], passes=16, tile=(768, 768, 512, 512)),
}
sel = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c1", "c2", "c3", "c3_tree", "c5_100", "c5"]
for name in sel:
    c = CONFIGS[name]
    t0 = time.perf_counter()
    sd = capi.SceneDescription(c["scene"])
    t1 = time.perf_counter()
    sc = capi.Scene(sd)
    t2 = time.perf_counter()
    pl = capi.Plane(sc, c["width"], c["height"])
    kw = dict(width=c["width"], height=c["height"], depth_max=c["depth_max"], schedule=c["schedule"], pass_count=1)
    if "tile" in c:
        kw.update(tile_x0=c["tile"][0], tile_y0=c["tile"][1], tile_w=c["tile"][2], tile_h=c["tile"][3])
    pl.render(capi.default_params(**kw))  # warm-up (workspace allocation)
    kw["pass_count"] = c["passes"]; kw["pass_begin"] = 1; kw["flags"] = capi.FLAG_TIME_KERNELS
    st = pl.render(capi.default_params(**kw))
    s, q, cnt = pl.download()
    out = dict(config=name, scene=c["scene"], frame=[c["width"], c["height"]], tile=c.get("tile"), schedule=c["schedule"], passes=c["passes"],
               mpaths_per_s=st.paths / st.ms_total / 1e3, mrays_per_s=st.rays / st.ms_total / 1e3, rays_per_path=st.rays / st.paths,
               ms_total=st.ms_total, ms_extend=st.ms_extend, ms_shade=st.ms_shade, bvh_nodes_per_ray=st.bvh_nodes_visited / max(st.rays, 1),
               tris_per_ray=st.triangles_tested / max(st.rays, 1),
               light_nodes_per_ray=st.light_bvh_nodes_visited / max(st.rays, 1), lights_per_ray=st.lights_tested / max(st.rays, 1), image_mean=float(s.sum() / max(cnt.sum(), 1)),
               scene_desc_s=t1 - t0, scene_create_s=t2 - t1, launches=st.kernel_launches)
    print(json.dumps(out), flush=True)
    pl.close(); sc.close()
