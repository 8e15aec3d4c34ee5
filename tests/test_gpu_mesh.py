"""GPU checks of the mesh extension: the device-built LBVH is byte-identical to the CPU restatement (integer work:
Morton codes, radix sort, Karras hierarchy; float boxes by a fixed operation order), traversal results are bit-exact
against the oracle, renders agree per pixel."""
import numpy as np
import pytest

import oracle_lib
from helpers import bits, ray_batch
from ipt_b200 import capi
from test_oracle_mesh import custom_scene

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [2, 3, 100, 4097, 300000])
def test_lbvh_build_bit_exact(n, lib, oracle):
    sd = capi.SceneDescription(f"mesh:{n}")
    sc = capi.Scene(sd)
    nodes, order, keys = sc.bvh_export()
    cn, cids, ckeys = oracle.bvh_build(sd.triangles())
    assert np.array_equal(keys, ckeys), "Morton keys / radix sort"
    assert np.array_equal(order, cids), "sorted primitive order"
    assert nodes.tobytes() == cn.tobytes(), "hierarchy + boxes"
    qn, grid = sc.bvh_export_compact()   # the 32-byte nodes the traversal kernels read
    cq, cgrid = oracle.bvh_compact(cn)
    assert np.array_equal(bits(grid), bits(cgrid)) and qn.tobytes() == cq.tobytes(), "quantised traversal nodes"
    sc.close()


def test_lbvh_with_duplicate_keys(lib, oracle):
    rng = np.random.default_rng(5)
    tris = np.tile(np.array([[0.1, 0.1, 0.1, 0.02, 0, 0, 0, 0.02, 0]], np.float32), (500, 1))
    tris[250:, :3] = rng.uniform(-0.5, 0.5, (250, 3)).astype(np.float32)
    tris[400:, :] = tris[399]
    sd, keep = custom_scene(tris)
    sc = capi.Scene(sd)
    nodes, order, keys = sc.bvh_export()
    cn, cids, ckeys = oracle.bvh_build(tris)
    assert np.array_equal(keys, ckeys) and np.array_equal(order, cids) and nodes.tobytes() == cn.tobytes()
    o, d, _ = ray_batch("box", lambda xy: oracle.camera_rays(sd.ptr, xy), n_cam_side=32, n_random=4000)
    g = sc.trace_batch(o, d); c = oracle.trace_batch(sd.ptr, o, d, use_bvh=0)
    assert np.array_equal(g["prim"], c["prim"]) and np.array_equal(bits(g["t"]), bits(c["t"]))
    sc.close()


@pytest.mark.parametrize("n", [1, 2, 1000, 100000])
def test_mesh_trace_batch_bit_exact(n, lib, oracle):
    sd = capi.SceneDescription(f"mesh:{n}")
    sc = capi.Scene(sd)
    o, d, _ = ray_batch("box", lambda xy: oracle.camera_rays(sd.ptr, xy), n_cam_side=64, n_random=20000)
    g = sc.trace_batch(o, d)
    c = oracle.trace_batch(sd.ptr, o, d, use_bvh=1 if n > 2000 else 0)
    assert np.array_equal(g["prim"], c["prim"])
    assert np.array_equal(bits(g["t"]), bits(c["t"]))
    assert np.array_equal(g["outcome"], c["outcome"]) and np.array_equal(g["light"], c["light"])
    if n >= 1000:
        assert (c["prim"] >= sd.desc.n_prims).sum() > 500
    sc.close()


def test_mesh_render_matches_oracle(lib, oracle):
    sd = capi.SceneDescription("mesh:20000")
    sc = capi.Scene(sd)
    p = capi.default_params(width=64, height=64, pass_count=2, depth_max=2, schedule=[16, 8], flags=capi.FLAG_KEEP_ZERO_WEIGHT)
    s, q, cnt, st = sc.render_host(p)
    o = oracle.render(sd.ptr, p, oracle_lib.RNG_PHILOX, 1)
    scale = max(o["sum"].max(), 1e-12)
    assert (np.abs(s - o["sum"]) / scale > 1e-5).mean() < 4e-3
    assert abs(int(st.rays) - int(o["rays"])) <= 2e-4 * o["rays"] + 2
    assert st.bvh_nodes_visited > 0 and st.triangles_tested > 0
    # depth 8, one child per hit (BASELINE configs[2] schedule): statistical agreement with common random numbers
    p8 = capi.default_params(width=64, height=64, pass_count=4, depth_max=8, schedule=[1] * 8, flags=capi.FLAG_KEEP_ZERO_WEIGHT)
    s8, q8, c8, st8 = sc.render_host(p8)
    o8 = oracle.render(sd.ptr, p8, oracle_lib.RNG_PHILOX, 1)
    assert abs(s8.sum() - o8["sum"].sum()) < 0.05 * o8["sum"].sum() + 1e-6
    assert abs(int(st8.rays) - int(o8["rays"])) < 2e-2 * o8["rays"]
    sc.close()


@pytest.mark.parametrize("n", [1, 2, 3, 50])
def test_tiny_meshes_render_like_the_oracle(n, lib, oracle):
    """The persistent mesh kernel on degenerate trees (a lone triangle has no node at all, two triangles a single node): the
    two-level tree per pixel against the oracle, and the last traced depth resolved in full (IPT_FLAG_RESOLVE_LAST_LEVEL:
    every ray traced, not only the ones that reach a light) gives the same image as the shadow-ray shortcut."""
    sd = capi.SceneDescription(f"mesh:{n}")
    sc = capi.Scene(sd)
    p = capi.default_params(width=48, height=48, pass_count=2, depth_max=2, schedule=[16, 8], flags=capi.FLAG_KEEP_ZERO_WEIGHT)
    s, q, cnt, st = sc.render_host(p)
    o = oracle.render(sd.ptr, p, oracle_lib.RNG_PHILOX, 0)
    scale = max(o["sum"].max(), 1e-12)
    assert np.array_equal(cnt.astype(np.uint64), o["counters"])
    assert (np.abs(s - o["sum"]) / scale > 1e-5).mean() < 4e-3
    assert abs(int(st.rays) - int(o["rays"])) <= 2e-4 * o["rays"] + 2
    full = sc.render_host(capi.default_params(width=48, height=48, pass_count=2, depth_max=3, schedule=[8, 4, 2]))
    resolved = sc.render_host(capi.default_params(width=48, height=48, pass_count=2, depth_max=3, schedule=[8, 4, 2], flags=capi.FLAG_RESOLVE_LAST_LEVEL))
    assert np.allclose(full[0], resolved[0], rtol=2e-5, atol=1e-7) and full[3].rays == resolved[3].rays
    assert full[3].light_hits == resolved[3].light_hits
    sc.close()


def test_c3_one_million_triangles_depth_8_with_common_random_numbers(lib, oracle):
    """BASELINE configs[2] (1 M triangles, depth 8, one child per hit) with the SAME Philox numbers on both sides: the mesh
    path traces every depth with the reference's exact arithmetic, so a pixel differs from the oracle only where an ulp of
    the sampled direction (CUDA vs glibc sin/cos) flips a hit decision somewhere along one of its paths. The large majority
    of the pixels is identical to float rounding, the rest is unbiased, and the ray counts agree per depth."""
    sd = capi.SceneDescription("mesh:1000000")
    sc = capi.Scene(sd)
    p = capi.default_params(width=96, height=96, pass_count=16, depth_max=8, schedule=[1] * 8, flags=capi.FLAG_KEEP_ZERO_WEIGHT)
    s, q, cnt, st = sc.render_host(p)
    o = oracle.render(sd.ptr, p, oracle_lib.RNG_PHILOX, 1)
    assert np.array_equal(cnt.astype(np.uint64), o["counters"]) and st.paths == 96 * 96 * 16
    scale = max(o["sum"].max(), 1e-12)
    diff = s - o["sum"]
    same = np.abs(diff) / scale <= 1e-5
    lit = o["sum"] > 0
    print(f"C3_CRN identical={same.mean():.5f} identical_lit={same[lit].mean():.5f} lit={lit.mean():.4f} rays {st.rays} vs {o['rays']}")
    assert same.mean() > 0.97 and same[lit].mean() > 0.9
    d = diff[~same]
    if d.size > 20:
        assert abs(d.mean()) < 5 * d.std() / np.sqrt(d.size)
    assert st.rays_at_depth[0] == o["rays_at_depth"][0] == 96 * 96 * 16
    for k in range(8):
        assert abs(int(st.rays_at_depth[k]) - int(o["rays_at_depth"][k])) <= 5e-3 * o["rays_at_depth"][k] + 8
    sc.close()


@pytest.mark.parametrize("scene,size,depth,schedule,cpu_passes,gpu_passes,block", [
    # BASELINE configs[2]: depth 8, one child per hit. A path reaches the light with probability ~1e-3 and with weights that
    # span orders of magnitude, so neither a pixel nor a 16x16 block of pixels has a Gaussian mean at any pass count the
    # oracle can afford (0.2 Mrays/s per core on this mesh; measured: block scores with |z| up to 7 between two ORACLE runs'
    # worth of samples). block = 0: only the frame mean (thousands of non-zero samples) and the ray counts are compared;
    # the sharp per-pixel comparison of this config is the common-random-number test above
    ("mesh:1000000", 64, 8, [1] * 8, 1024, 65536, 0),
    # the reference's own tree on a mesh: splitting 16/8/4/2 makes every pixel well behaved
    ("mesh:100000", 32, 4, [16, 8, 4, 2], 256, 4096, 1),
])
def test_mesh_image_z_test_against_oracle(scene, size, depth, schedule, cpu_passes, gpu_passes, block, lib, oracle):
    """Converged-image parity of the mesh path (the reference has no mesh geometry, SURVEY S1: the oracle's GeometryMesh over
    the CPU LBVH is the checker). INDEPENDENT random numbers on the two sides (different Philox keys): z-test of the means
    per cell (pixel, or block of pixels where single pixels are too sparse), bias detector, image mean. The oracle runs as
    forked processes over disjoint pass ranges. Tolerances: >= 99.7 % of the cells |z| < 3 up to three standard errors of
    that fraction, |mean z| below 3 / sqrt(cells) (three standard errors of the mean of N(0,1) scores) and never above
    0.1 where there are enough cells for that, image mean within 3 sigma of the two estimates + 0.2 %, and the relative
    RMSE of 8x8 block means below 1 % + 1.5x the RMSE the two sides' own variances predict (the CPU side cannot afford the
    passes that would push its own noise below 1 %)."""
    from helpers import block_sums, oracle_render_parallel
    from test_gpu_golden import image_stats
    from test_oracle_golden import same_coverage

    sd = capi.SceneDescription(scene)
    sc = capi.Scene(sd)
    kw = dict(width=size, height=size, depth_max=depth, schedule=schedule)
    o = oracle_render_parallel(scene, cpu_passes, seed=1234, use_bvh=1, **kw)
    s, q, cnt, st = sc.render_host(capi.default_params(pass_count=gpu_passes, seed=98765, flags=capi.FLAG_KEEP_ZERO_WEIGHT, **kw))  # count rays like the oracle
    assert same_coverage(cnt, gpu_passes, o["count"], cpu_passes)
    assert abs(st.rays / st.paths - o["rays"] / o["count"].sum()) < 0.01 * st.rays / st.paths
    if block == 0:
        n1, n2 = float(cnt.sum()), float(o["count"].sum())
        m1, m2 = s.sum(dtype=np.float64) / n1, o["sum"].sum() / n2
        sigma = np.sqrt((q.sum(dtype=np.float64) / n1 - m1 * m1) / n1 + (o["sumsq"].sum() / n2 - m2 * m2) / n2)
        print(f"IMAGE_STATS {scene} depth {depth} frame mean gpu={m1:.6g} cpu={m2:.6g} sigma={sigma:.3g} z={(m1 - m2) / sigma:.3f}")
        assert abs(m1 - m2) < 3 * sigma + 0.002 * m2
        sc.close()
        return
    g = dict(sum=block_sums(o["sum"], block), sumsq=block_sums(o["sumsq"], block), count=block_sums(o["count"], block))
    r = image_stats(block_sums(s.astype(np.float64), block), block_sums(q.astype(np.float64), block), block_sums(cnt.astype(np.uint64), block), g,
                    block=max(1, 8 // block))
    print(f"IMAGE_STATS {scene} depth {depth} cells of {block}x{block} " + " ".join(f"{k}={v:.5g}" for k, v in r.items()))
    assert r["frac3"] >= r["frac3_floor"], r
    assert abs(r["mean_z"]) < max(0.1, 3.0 / np.sqrt(r["n_lit"])), r
    assert abs(r["mean_rel"]) < 3 * r["mean_rel_sigma"] + 0.002, r
    assert r["block_rel_rmse"] < 0.01 + 1.5 * r["block_noise_rel"], r
    assert abs(st.rays / st.paths - o["rays"] / o["count"].sum()) < 0.01 * st.rays / st.paths
    sc.close()
