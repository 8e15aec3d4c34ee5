"""Output stage on the device (ipt_b200/csrc/ipt_output.cuh) against the oracle restatement of src/gui.cpp and the
reference's golden vectors: glare bit-exact, Gui::save bytes bit-exact, normalize within 1 ulp."""
from pathlib import Path

import numpy as np
import pytest

from helpers import bits
from ipt_b200 import capi
from test_output_host import decode_png_gray8

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def hdr_image(rng, h, w, scale=6.0, power=5, zero_frac=0.3):
    img = (rng.random((h, w)).astype(np.float32) ** power * np.float32(scale)).astype(np.float32)
    img[rng.random((h, w)) < zero_frac] = 0
    img.flat[0] = np.float32(scale) * np.float32(0.75)
    return img


def ulp_diff(a, b):
    return np.abs(bits(a).astype(np.int64) - bits(b).astype(np.int64))


@pytest.mark.parametrize("name", ["synthetic", "box"])
def test_reference_golden_output_stage(name, lib):
    kat = np.load(GOLD / "output_kat.npz")
    img = kat[f"{name}_image"]
    for c, g in zip(kat[f"{name}_cutoffs"], kat[f"{name}_glare"]):
        out, nb = capi.image_glare(img, float(c))
        assert nb == int((img > c).sum())
        assert np.array_equal(bits(out), bits(g)), c
    assert np.array_equal(capi.image_save_bytes(img), kat[f"{name}_bytes"])
    assert ulp_diff(capi.image_normalize(img), kat[f"{name}_normalize"]).max() <= 1


@pytest.mark.parametrize("shape,cutoff", [((1, 1), 0.5), ((3, 5), 0.1), ((33, 7), 1.01), ((64, 129), 0.4), ((200, 333), 2.0),
                                          ((640, 640), 5.9), ((16, 16), 1e-3), ((50, 50), 100.0)])
def test_glare_bit_exact(shape, cutoff, lib, oracle):
    """Ragged sizes, no / few / all pixels above the cutoff; the default Gui resolution with ~0.2 % halo sources."""
    img = hdr_image(np.random.default_rng(shape[0] * 1000 + shape[1]), *shape)
    out, nb = capi.image_glare(img, cutoff)
    assert nb == int((img > np.float32(cutoff)).sum())
    assert np.array_equal(bits(out), bits(oracle.image_glare(img, cutoff)))


def test_glare_guarded_paths_bit_exact(lib, oracle):
    """Tiny cutoffs leave the range of the inline division (IEEE fallback); a 4200-wide strip takes the double sqrt path."""
    rng = np.random.default_rng(9)
    img = (hdr_image(rng, 24, 31) * np.float32(1e-22)).astype(np.float32)
    out, nb = capi.image_glare(img, 1e-24)
    assert nb > 100 and np.array_equal(bits(out), bits(oracle.image_glare(img, 1e-24)))
    strip = hdr_image(rng, 2, 4200, zero_frac=0.9, power=12)
    out, nb = capi.image_glare(strip, 2.0)
    assert 0 < nb < 400 and np.array_equal(bits(out), bits(oracle.image_glare(strip, 2.0)))


@pytest.mark.parametrize("shape", [(1, 2), (5, 3), (97, 33), (640, 640), (1024, 1024)])
def test_save_bytes_bit_exact(shape, lib, oracle):
    rng = np.random.default_rng(shape[0])
    img = hdr_image(rng, *shape)
    assert np.array_equal(capi.image_save_bytes(img), oracle.image_save_bytes(img))
    lifted = (img + np.float32(0.37)).astype(np.float32)  # minimum above zero: normalize(0,255) stretches from it
    got = capi.image_save_bytes(lifted)
    assert np.array_equal(got, oracle.image_save_bytes(lifted)) and got.min() == 0 and got.max() == 255


def test_save_bytes_constant_and_black(lib, oracle):
    flat = np.full((9, 4), 0.7, np.float32)
    assert np.array_equal(capi.image_save_bytes(flat), oracle.image_save_bytes(flat))  # constant image -> all 0 (CImg.h:33175-33178)
    for bad in (np.zeros((4, 4), np.float32), np.array([[1.0, -0.5]], np.float32), np.array([[1.0, np.nan]], np.float32)):
        with pytest.raises(capi.IptError) as e:
            capi.image_save_bytes(bad)
        assert e.value.code == capi.IPT_ERR_INVALID


def test_normalize_within_one_ulp(lib, oracle):
    img = hdr_image(np.random.default_rng(4), 300, 200)
    got, want = capi.image_normalize(img), oracle.image_normalize(img)
    d = ulp_diff(got, want)
    assert d.max() <= 1 and (d > 0).mean() < 0.2  # powf is 0.82 ulp, the device evaluates pow in double
    assert got.max() == 1.0 and got.min() == 0.0


def test_plane_display_and_save(lib, oracle, tmp_path):
    """Gui's path end to end on the device: render into Gui's cell mapping, then updateDisplay's filter chain and save()."""
    sd = capi.SceneDescription("box")
    sc = capi.Scene(sd)
    W = H = 96
    plane = capi.Plane(sc, W, H)
    p = capi.default_params(width=W, height=H, pass_count=8, plane_mode=capi.PLANE_GUI)
    plane.render(p)
    s, q, c = plane.download()
    mean = np.where(c > 0, s / np.maximum(c, 1).astype(np.float32), np.float32(0)).astype(np.float32)
    cutoff = float(np.float32(mean.max()) * np.float32(0.25))
    shown, ms = plane.display(cutoff)
    want = oracle.image_normalize(oracle.image_glare(mean, cutoff))
    assert (mean > cutoff).sum() > 10 and ms > 0
    assert ulp_diff(shown, want).max() <= 1
    by = plane.save_bytes()
    assert np.array_equal(by, oracle.image_save_bytes(mean))
    plane.save_png(tmp_path / "result.png")
    assert np.array_equal(decode_png_gray8((tmp_path / "result.png").read_bytes()), by)
    plane.close(); sc.close()


def test_progressive_session_keys_checkpoint_and_resume(lib, oracle, tmp_path):
    """The reference's interactive loop, headless (ipt_b200/progressive.py): progressive steps, an arrow key that moves the
    camera and restarts the image (gui.cpp:105-137,152-160), checkpoint + resume equal to the uninterrupted session."""
    from ipt_b200.progressive import KEY_LEFT, KEY_UP, ProgressiveSession

    kw = dict(scene_name="box", width=64, height=64, passes_per_call=2, seed=5)
    a = ProgressiveSession(**kw)
    assert a.step(2) == 4
    before = a.plane.download()[0].copy()
    a.key(KEY_LEFT); a.key(KEY_UP)
    assert a.samples_per_pixel == 0 and a.plane.download()[2].sum() == 0        # resetImage
    want = oracle.camera_orbit(oracle.camera_orbit([0.0, -3.0, 0.1], list(a.description.desc.camera.direction), 0)[0],
                               oracle.camera_orbit([0.0, -3.0, 0.1], list(a.description.desc.camera.direction), 0)[1], 3)
    got = np.array([list(a.camera.position), list(a.camera.direction), list(a.camera.right), list(a.camera.up)], np.float32)
    assert np.array_equal(bits(got), bits(want))
    a.step(1)
    ck = tmp_path / "session"          # no suffix: the file written and the file looked for must be the same one
    written = a.checkpoint(ck)
    assert written == tmp_path / "session.npz" and written.exists() and not list(tmp_path.glob("*.tmp*"))
    with pytest.raises(ValueError):     # another estimator must not resume these accumulators
        ProgressiveSession(checkpoint_path=ck, **dict(kw, depth_max=3))
    with pytest.raises(ValueError):
        ProgressiveSession(checkpoint_path=ck, **dict(kw, scene_name="corner"))
    a.step(2)
    s_a, q_a, c_a = a.plane.download()
    assert a.samples_per_pixel == 6 and (c_a.sum() == 6 * 64 * 64) and not np.allclose(s_a / 6, before / 4, atol=1e-3)  # another view
    b = ProgressiveSession(checkpoint_path=ck, **kw)                                # a new process would do exactly this
    assert b.samples_per_pixel == 2 and b.next_pass == a.next_pass - 4
    b.step(2)
    s_b, q_b, c_b = b.plane.download()
    assert np.array_equal(c_a, c_b) and np.allclose(s_a, s_b, rtol=2e-5, atol=1e-7)
    shown = b.display()
    assert shown.shape == (64, 64) and shown.max() == 1.0
    b.save(tmp_path / "result.png")
    assert np.array_equal(decode_png_gray8((tmp_path / "result.png").read_bytes()), b.plane.save_bytes())
    a.close(); b.close()
