"""Live pin of the oracle against the compiled reference (oracle/_ref/libipt_ref.so): with libc drand48 seeded
identically, a whole render_sample pass of the restatement is BIT-IDENTICAL to the reference's, on every reference
scene and on the C2 extension scene. Skipped where the reference library does not exist."""
import numpy as np
import pytest

import oracle_lib
from helpers import SCENES_ANALYTIC, bits
from ipt_b200 import capi


def _schedule(n, depth_max):
    out = []
    for _ in range(depth_max):
        out.append(n)
        n //= 2  # main.cpp:177 passes n_rays/2 down
    return out


@pytest.mark.parametrize("scene,n_rays,depth_max", [("box", 4, 3), ("cornell", 4, 3), ("corner", 4, 3), ("square", 8, 4),
                                                    ("smallpt", 2, 3), ("fractal", 4, 4), ("openspheres", 4, 3), ("box", 3, 4),
                                                    ("mixedlights", 4, 3)])
def test_whole_pass_bit_identical_to_reference(scene, n_rays, depth_max, lib, oracle, ref):
    """Verbatim src/main.cpp render_sample (640x640) vs the oracle, same srand48 seed: pixels, counters and ray count."""
    sd = capi.SceneDescription(scene)
    h = ref.scene(scene)
    ref.set_tree(n_rays, depth_max)
    ref.seed(4242)
    r = ref.render(h, 1)
    p = capi.default_params(depth_max=depth_max, schedule=_schedule(n_rays, depth_max))
    oracle.seed(4242)
    o = oracle.render(sd.ptr, p, oracle_lib.RNG_DRAND48, 0)
    assert np.array_equal(bits(r["pixels"]), bits(o["pixels"]))
    assert np.array_equal(r["counters"], o["counters"])
    assert r["rays"] == o["rays"]
    assert r["pixels"].max() > 0


def test_parametrised_loop_equals_verbatim_loop(ref):
    """oracle/ref_driver.cpp's W,H loop (used for non-640 goldens) is the reference loop at 640x640."""
    h = ref.scene("box")
    ref.set_tree(2, 3)
    ref.seed(7); a = ref.render(h, 1)
    ref.seed(7); b = ref.render(h, 1, 640, 640, verbatim=False)
    assert np.array_equal(bits(a["pixels"]), bits(b["pixels"])) and a["rays"] == b["rays"]


@pytest.mark.parametrize("scene", ["box", "cornell", "lightgrid:3x3", "fractal", "mixedlights"])
def test_mixture_samples_bit_identical(scene, lib, oracle, ref):
    """unite(light_ddf,1,sdf,1) at a surface hit (main.cpp:142-143): sample(), value() and sdf value() sequences."""
    sd = capi.SceneDescription(scene)
    h = ref.scene(scene)
    o0, d0 = oracle.camera_rays(sd.ptr, np.array([[0.5, 0.3], [0.62, 0.45], [0.5, 0.5]], np.float32))
    done = 0
    for o, d in zip(o0, d0):
        if oracle.trace_batch(sd.ptr, [o], [d])["prim"][0] == capi.IPT_NO_HIT:
            continue
        ref.seed(11); w_r, m_r, s_r = ref.mix_sample(h, o, d, 400)
        oracle.seed(11); w_o, m_o, s_o = oracle.mix_sample(sd.ptr, o, d, 400)
        assert np.array_equal(bits(w_r), bits(w_o))
        nz = np.any(w_r != 0, axis=1)  # value(vec3()) reads the racy Lighting::last_sample; the loop discards it
        assert np.array_equal(bits(m_r[nz]), bits(m_o[nz])) and np.array_equal(bits(s_r[nz]), bits(s_o[nz]))
        done += 1
    assert done > 0


@pytest.mark.parametrize("kind", [0, 1, 2, 40])
def test_base_ddf_samples_bit_identical(kind, oracle, ref):
    for to in (None, [0.6, 0.0, 0.8], [0, 0, -1]):
        ref.seed(5); a = ref.ddf_sample(kind, 300, to=to)
        oracle.seed(5); b = oracle.ddf_sample(kind, 300, to=to)
        assert np.array_equal(bits(a), bits(b))


def test_light_ddf_value_and_sample(lib, oracle, ref):
    """Lighting::distributionInPoint(pos) (CollectionLighting.cpp:12-21): the iteratively renormalised weights."""
    rng = np.random.default_rng(2)
    w = rng.normal(size=(300, 3)).astype(np.float32)
    w /= np.linalg.norm(w, axis=1, keepdims=True).astype(np.float32)
    w[:, 2] = np.abs(w[:, 2])
    for scene in ["box", "lightgrid:3x3", "corner", "fractal", "mixedlights"]:
        sd = capi.SceneDescription(scene)
        h = ref.scene(scene)
        pos = np.array([0.1, -0.5, -1.0], np.float32)
        assert np.array_equal(bits(ref.light_ddf_value(h, pos, w)), bits(oracle.light_ddf_value(sd.ptr, pos, w))), scene
        ref.seed(3); a = ref.light_ddf_sample(h, pos, 200)
        oracle.seed(3); b = oracle.light_ddf_sample(sd.ptr, pos, 200)
        assert np.array_equal(bits(a), bits(b)), scene
