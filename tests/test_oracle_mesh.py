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
