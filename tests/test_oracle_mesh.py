"""CPU checks of the mesh extension (BASELINE configs[2..3]): the oracle's LBVH traversal returns exactly what the
brute-force scan (the semantic definition) returns, and the tree is well formed. No GPU."""
import ctypes as C

import numpy as np
import pytest

from helpers import bits, ray_batch
from ipt_b200 import capi


def custom_scene(tris):
    """The 'mesh:<n>' scene with its triangle array replaced (kept alive by the returned tuple)."""
    tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
    sd = capi.SceneDescription("mesh:1")
    sd.desc.n_triangles = tris.shape[0]
    sd.desc.triangles = tris.ctypes.data_as(capi.f32p)
    return sd, tris


@pytest.mark.parametrize("n", [1, 2, 3, 257, 5000])
def test_bvh_traversal_equals_brute_force(n, lib, oracle):
    sd = capi.SceneDescription(f"mesh:{n}")
    o, d, _ = ray_batch("box", lambda xy: oracle.camera_rays(sd.ptr, xy), n_cam_side=48, n_random=6000)
    a = oracle.trace_batch(sd.ptr, o, d, use_bvh=0)
    b = oracle.trace_batch(sd.ptr, o, d, use_bvh=1)
    assert np.array_equal(a["prim"], b["prim"]) and np.array_equal(bits(a["t"]), bits(b["t"]))
    assert np.array_equal(a["outcome"], b["outcome"])
    if n >= 257:
        assert (a["prim"] >= sd.desc.n_prims).sum() > 100  # triangles are actually hit


def test_duplicate_and_coplanar_triangles_tie_break_to_lowest_index(lib, oracle):
    """Identical triangles (equal Morton keys AND equal hit distances): the linear scan keeps the first; so must the tree."""
    base = np.array([[-0.3, -0.3, 0.2, 0.6, 0, 0, 0, 0.6, 0]], np.float32)  # normal +z... faces +z: seen from above
    flipped = np.array([[-0.3, -0.3, 0.2, 0, 0.6, 0, 0.6, 0, 0]], np.float32)  # normal -z: seen from below
    tris = np.concatenate([flipped, flipped, base, flipped, base, base, flipped] * 9)
    sd, keep = custom_scene(tris)
    rng = np.random.default_rng(0)
    o = np.stack([rng.uniform(-0.4, 0.4, 3000), rng.uniform(-0.4, 0.4, 3000), np.where(rng.random(3000) < 0.5, -0.5, 0.7)], 1).astype(np.float32)
    d = np.zeros_like(o); d[:, 2] = np.where(o[:, 2] < 0, 1, -1)
    a = oracle.trace_batch(sd.ptr, o, d, use_bvh=0)
    b = oracle.trace_batch(sd.ptr, o, d, use_bvh=1)
    assert np.array_equal(a["prim"], b["prim"]) and np.array_equal(bits(a["t"]), bits(b["t"]))
    tri_hits = a["prim"][a["prim"] >= sd.desc.n_prims] - sd.desc.n_prims
    assert set(np.unique(tri_hits)) <= {0, 2}  # first flipped (index 0) from below, first base (index 2) from above


def test_tree_is_well_formed(lib, oracle):
    sd = capi.SceneDescription("mesh:3000")
    tris = sd.triangles()
    nodes, ids, keys = oracle.bvh_build(tris)
    n = len(tris)
    assert sorted(ids.tolist()) == list(range(n))
    assert (np.diff(keys.astype(np.int64)) >= 0).all() and keys.max() < (1 << 63)
    leaves = []
    for k, nd in enumerate(nodes):
        for child, lo, hi in ((nd["left"], nd["lo0"], nd["hi0"]), (nd["right"], nd["lo1"], nd["hi1"])):
            if child & 0x80000000:
                leaves.append(int(child & 0x7FFFFFFF))
                t = tris[ids[child & 0x7FFFFFFF]]
                v = np.stack([t[:3], t[:3] + t[3:6], t[:3] + t[6:9]])
                assert (v >= lo).all() and (v <= hi).all()
            else:
                c = nodes[child]
                assert c["parent"] == k
                assert (np.minimum(c["lo0"], c["lo1"]) == lo).all() and (np.maximum(c["hi0"], c["hi1"]) == hi).all()
    assert sorted(leaves) == list(range(n))
    assert nodes[0]["parent"] == 0xFFFFFFFF


def test_mesh_render_bvh_equals_brute_force(lib, oracle):
    import oracle_lib

    sd = capi.SceneDescription("mesh:400")
    p = capi.default_params(width=24, height=24, pass_count=1, schedule=[4, 2, 1, 1])
    a = oracle.render(sd.ptr, p, oracle_lib.RNG_PHILOX, 0)
    b = oracle.render(sd.ptr, p, oracle_lib.RNG_PHILOX, 1)
    assert np.array_equal(a["sum"], b["sum"]) and a["rays"] == b["rays"]


@pytest.mark.parametrize("n", [2, 37, 5000, 60000])
def test_quantised_traversal_nodes_contain_the_float_boxes(n, lib, oracle):
    """The 32-byte nodes the traversal kernels read (16-bit grid over the root box): every quantised child box contains
    the float box of the 64-byte node with at least half a grid cell to spare on every side, is at most 3 cells larger, and carries the same child ids — so the traversal can visit a
    node too many, never one too few."""
    sd = capi.SceneDescription(f"mesh:{n}")
    nodes, ids, keys = oracle.bvh_build(sd.triangles())
    q, grid = oracle.bvh_compact(nodes)
    lo, scale = grid[:3].astype(np.float64), grid[3:].astype(np.float64)
    assert np.array_equal(q[:, 6], nodes["left"]) and np.array_equal(q[:, 7], nodes["right"])

    def halves(w):
        return (w & 0xFFFF).astype(np.float64), (w >> 16).astype(np.float64)

    w = [halves(q[:, k]) for k in range(6)]
    for child, (name_lo, name_hi) in enumerate([("lo0", "hi0"), ("lo1", "hi1")]):
        b = 3 * child
        qlo = np.stack([w[b][0], w[b][1], w[b + 1][0]], 1)
        qhi = np.stack([w[b + 1][1], w[b + 2][0], w[b + 2][1]], 1)
        glo = (nodes[name_lo].astype(np.float64) - lo) * scale * 65536.0 + 4.0   # float boxes in units of grid cells
        ghi = (nodes[name_hi].astype(np.float64) - lo) * scale * 65536.0 + 4.0
        assert (qlo <= glo - 0.5).all() and (glo - qlo <= 3).all()
        assert (qhi >= ghi + 0.5).all() and (qhi - ghi <= 3).all()
        assert (qlo >= 1).all() and (qhi <= 65534).all()                         # the root box leaves room for the rounding


def test_quantised_traversal_finds_the_same_hits_with_barely_more_visits(oracle):
    """The 32-byte traversal nodes (16-bit grid, rounded outwards) through the restated grid-space slab test: same closest
    triangle and distance as the walk over the float nodes for every ray, a fraction of a per cent more node visits — and a
    direction component that is exactly 0 (sampled directions hit one at a 2^-23 rate) still culls: with an infinite
    reciprocal such a ray used to walk the whole tree."""
    sd = capi.SceneDescription("mesh:20000")
    tris = sd.triangles()
    rng = np.random.default_rng(11)
    n = 6000
    o = ((rng.random((n, 3)) * 2 - 1) * 0.95).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    d = np.ascontiguousarray(d, np.float32)
    for a in range(3):  # axis-parallel components, both signs of zero
        d[a * 100:(a + 1) * 100, a] = 0.0
        d[300 + a * 100:300 + (a + 1) * 100, a] = -0.0
    f = oracle.bvh_trace_counts(tris, o, d, 0)
    q = oracle.bvh_trace_counts(tris, o, d, 1)
    assert np.array_equal(f["prim"], q["prim"]) and np.array_equal(f["t"].view(np.uint32), q["t"].view(np.uint32))
    assert (f["prim"] != 0xFFFFFFFF).mean() > 0.05
    assert q["nodes"].sum() >= f["nodes"].sum() * 0.999 and q["nodes"].sum() <= f["nodes"].sum() * 1.02
    zero = slice(0, 600)
    assert q["nodes"][zero].max() <= 4 * max(int(f["nodes"].max()), 1)  # nowhere near the 19 999 nodes of the whole tree
