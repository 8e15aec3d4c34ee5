"""BASELINE.json's configurations at their FULL sizes, through properties that do not need a CPU render of that size
(pass additivity, tile union, linearity in the emitted power, fused == queued last level, frame-size independence of the
mean), plus oracle comparisons where the oracle still finishes in seconds (two-level trees at 640x640, bit-exact traces
against the 1 M-triangle LBVH and the 10 000-emitter collection)."""
import numpy as np
import pytest

import oracle_lib
from helpers import bits, ray_batch
from ipt_b200 import capi

pytestmark = pytest.mark.gpu
C2 = dict(width=1024, height=1024, depth_max=4, schedule=[16, 8, 4, 2], plane_mode=capi.PLANE_LINEAR)


@pytest.fixture(scope="module")
def cornell(lib):
    sd = capi.SceneDescription("cornell")
    sc = capi.Scene(sd)
    yield sd, sc
    sc.close()


def close(a, b, rtol=2e-5):
    return np.allclose(a, b, rtol=rtol, atol=1e-7)


def test_c2_pass_ranges_and_tiles_add_up(cornell):
    """configs[1] frame, full tree: passes [0,2) in one call == [0,1) + [1,2); four 512x512 tiles == the frame."""
    sd, sc = cornell
    whole = capi.Plane(sc, 1024, 1024)
    st = whole.render(capi.default_params(pass_begin=0, pass_count=2, **C2))
    s, q, c = whole.download()
    assert st.paths == 2 * 1024 * 1024 and (c == 2).all() and 95 < st.rays / st.paths < 115
    parts = capi.Plane(sc, 1024, 1024)
    parts.render(capi.default_params(pass_begin=0, pass_count=1, **C2))
    for (x0, y0) in [(0, 0), (512, 0), (0, 512), (512, 512)]:
        parts.render(capi.default_params(pass_begin=1, pass_count=1, tile_x0=x0, tile_y0=y0, tile_w=512, tile_h=512, **C2))
    s2, q2, c2 = parts.download()
    assert np.array_equal(c, c2) and close(s, s2) and close(q, q2, 5e-5)
    whole.close(); parts.close()


def test_c2_fused_equals_queued_at_full_size(cornell):
    sd, sc = cornell
    a = sc.render_host(capi.default_params(pass_count=1, **C2))
    b = sc.render_host(capi.default_params(pass_count=1, flags=capi.FLAG_NO_FUSED_LAST_LEVEL, **C2))
    assert close(a[0], b[0]) and np.array_equal(a[2], b[2])
    assert a[3].rays == b[3].rays and a[3].light_hits == b[3].light_hits and list(a[3].rays_at_depth) == list(b[3].rays_at_depth)
    assert a[3].rays_resolved_in_shade == a[3].rays - a[3].paths > 0 and b[3].rays_resolved_in_shade == 0


def test_c2_image_is_linear_in_the_emitted_power(cornell):
    """The estimator is linear in leaf emissions (main.cpp:123,172-181): 4x the light's power is exactly 4x every term."""
    sd, sc = cornell
    base = sc.render_host(capi.default_params(pass_count=1, **C2))
    sd4 = capi.SceneDescription("cornell")
    for i in range(sd4.desc.n_lights):
        sd4.desc.lights[i].power *= 4.0
    sc4 = capi.Scene(sd4)
    four = sc4.render_host(capi.default_params(pass_count=1, **C2))
    assert four[3].rays == base[3].rays  # the sampling does not depend on the power of a single light
    assert close(four[0], 4.0 * base[0]) and close(four[1], 16.0 * base[1], 5e-5)
    sc4.close()


def test_c2_mean_radiance_does_not_depend_on_the_frame_size(cornell, oracle):
    """Same view, same estimator: the frame mean at 1024x1024 (device) and at 128x128 (oracle, Philox) agree statistically."""
    sd, sc = cornell
    s, q, c, st = sc.render_host(capi.default_params(pass_count=1, flags=capi.FLAG_KEEP_ZERO_WEIGHT, **C2))  # count rays like the reference
    small = capi.default_params(width=128, height=128, pass_count=8, depth_max=4, schedule=[16, 8, 4, 2], plane_mode=capi.PLANE_LINEAR)
    o = oracle.render(sd.ptr, small, oracle_lib.RNG_PHILOX, 0)
    m_gpu, m_cpu = s.sum() / c.sum(), o["sum"].sum() / o["counters"].sum()
    var_cpu = (o["sumsq"].sum() / o["counters"].sum() - m_cpu ** 2) / o["counters"].sum()
    var_gpu = (q.sum() / c.sum() - m_gpu ** 2) / c.sum()
    assert abs(m_gpu - m_cpu) < 4 * np.sqrt(var_cpu + var_gpu) + 2e-3 * m_cpu  # + pixel-footprint term of the coarser grid
    assert abs(st.rays / st.paths - o["rays"] / o["counters"].sum()) < 0.02 * st.rays / st.paths


def test_c1_two_level_tree_per_pixel_at_640(lib, oracle):
    """configs[0] frame (640x640, the reference's hard-coded size): camera ray + 16 mixture samples, per pixel against the oracle."""
    sd = capi.SceneDescription("box")
    sc = capi.Scene(sd)
    p = capi.default_params(width=640, height=640, pass_count=1, depth_max=2, schedule=[16, 8], flags=capi.FLAG_KEEP_ZERO_WEIGHT)
    s, q, c, st = sc.render_host(p)
    o = oracle.render(sd.ptr, p, oracle_lib.RNG_PHILOX, 0)
    assert np.array_equal(c, o["counters"].astype(np.uint32)) and st.rays == o["rays"]
    scale = max(o["sum"].max(), 1e-12)
    assert (np.abs(s - o["sum"]) / scale > 1e-5).mean() < 2e-3
    assert abs(s.sum() - o["sum"].sum()) <= 2e-4 * o["sum"].sum()
    sc.close()


def test_c3_one_million_triangles_bit_exact(lib, oracle):
    """configs[2] geometry: LBVH of the 1 M-triangle mesh equal to the oracle's node for node, closest hits bit-exact."""
    sd = capi.SceneDescription("mesh:1000000")
    sc = capi.Scene(sd)
    nodes, order, keys = sc.bvh_export()
    cn, cids, ckeys = oracle.bvh_build(sd.triangles())
    assert np.array_equal(keys, ckeys) and np.array_equal(order, cids)
    assert nodes.tobytes() == cn.tobytes()
    o, d, _ = ray_batch("box", lambda xy: oracle.camera_rays(sd.ptr, xy), n_cam_side=96, n_random=30000)
    g = sc.trace_batch(o, d)
    c = oracle.trace_batch(sd.ptr, o, d, use_bvh=1)
    assert np.array_equal(g["prim"], c["prim"]) and np.array_equal(bits(g["t"]), bits(c["t"]))
    assert np.array_equal(g["outcome"], c["outcome"]) and (c["prim"] >= sd.desc.n_prims).sum() > 10000
    sc.close()


def test_c5_ten_thousand_emitters_bit_exact(lib, oracle):
    """configs[4] lighting: nearest-light ids / positions through the light LBVH against the oracle's O(L) scan, and the
    mixture density (all-hits query) against the oracle's sum over 10 000 lights."""
    sd = capi.SceneDescription("lightgrid:100x100")
    sc = capi.Scene(sd)
    rng = np.random.default_rng(8)
    n = 6000
    o = np.stack([rng.uniform(-0.95, 0.95, n), rng.uniform(-0.95, 0.95, n), rng.uniform(-0.99, 0.5, n)], 1).astype(np.float32)
    tgt = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), np.full(n, 0.99)], 1).astype(np.float32)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    g = sc.trace_batch(o, d)
    c = oracle.trace_batch(sd.ptr, o, d)
    assert (c["light"] != capi.IPT_NO_HIT).sum() > 500
    assert np.array_equal(g["light"], c["light"]) and np.array_equal(bits(g["light_pos"]), bits(c["light_pos"]))
    assert np.array_equal(g["outcome"], c["outcome"])
    pos = np.array([0.1, -0.3, -1.0], np.float32)
    assert np.allclose(sc.light_ddf_value(pos, d[:1500]), oracle.light_ddf_value(sd.ptr, pos, d[:1500]), rtol=2e-4, atol=1e-6)
    sc.close()


def test_c4_ten_million_triangles_bit_exact(lib, oracle):
    """configs[3] geometry at its full size: the device-built LBVH of the 10 M-triangle mesh (Morton keys, radix sort,
    Karras hierarchy, refit) is byte-identical to the CPU restatement, and 30 000 + 96x96 rays find the same closest
    triangle with the same hit distance, bit for bit."""
    sd = capi.SceneDescription("mesh:10000000")
    sc = capi.Scene(sd)
    nodes, order, keys = sc.bvh_export()
    cn, cids, ckeys = oracle.bvh_build(sd.triangles())
    assert np.array_equal(keys, ckeys) and np.array_equal(order, cids)
    assert nodes.tobytes() == cn.tobytes()
    qn, grid = sc.bvh_export_compact()
    cq, cgrid = oracle.bvh_compact(cn)
    assert np.array_equal(bits(grid), bits(cgrid)) and qn.tobytes() == cq.tobytes()
    del cn, cids, ckeys, nodes, order, keys, qn, cq
    o, d, _ = ray_batch("box", lambda xy: oracle.camera_rays(sd.ptr, xy), n_cam_side=96, n_random=30000)
    g = sc.trace_batch(o, d)
    c = oracle.trace_batch(sd.ptr, o, d, use_bvh=1)
    assert np.array_equal(g["prim"], c["prim"]) and np.array_equal(bits(g["t"]), bits(c["t"]))
    assert np.array_equal(g["outcome"], c["outcome"]) and np.array_equal(g["light"], c["light"])
    assert (c["prim"] >= sd.desc.n_prims).sum() > 20000
    sc.close()


def test_c4_frame_passes_and_tiles_add_up(lib):
    """configs[3] frame (3840x2160, depth 8, one child per hit): a pass rendered whole == the same pass rendered as four
    tiles, and the per-depth ray counts agree — the sharding the 2/4/8-GPU runs rely on."""
    sd = capi.SceneDescription("mesh:200000")
    sc = capi.Scene(sd)
    kw = dict(width=3840, height=2160, depth_max=8, schedule=[1] * 8, plane_mode=capi.PLANE_LINEAR)
    whole = capi.Plane(sc, 3840, 2160)
    st = whole.render(capi.default_params(pass_begin=3, pass_count=1, **kw))
    s, q, c = whole.download()
    parts = capi.Plane(sc, 3840, 2160)
    rays = 0
    for (x0, y0) in [(0, 0), (1920, 0), (0, 1080), (1920, 1080)]:
        rays += parts.render(capi.default_params(pass_begin=3, pass_count=1, tile_x0=x0, tile_y0=y0, tile_w=1920, tile_h=1080, **kw)).rays
    s2, q2, c2 = parts.download()
    assert st.paths == 3840 * 2160 and rays == st.rays and (c == 1).all() and np.array_equal(c, c2)
    assert close(s, s2) and close(q, q2, 5e-5)
    whole.close(); parts.close(); sc.close()
