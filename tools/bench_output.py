"""Output stage measurement: Gui::updateDisplay's filter chain normalize(glare(image, cutoff)) and
Gui::save's bytes at the reference's 640x640, device time on the B200 next to the compiled reference on one host core.

    python tools/bench_output.py [--size 640] [--sources 2000] [--no-cpu]

One JSON line. A "pair" is one (output pixel, halo source) evaluation of draw_halo (gui.cpp:28-36): hypot, two float
divisions, one add. The chain is arithmetic bound (W*H*sources pairs, 12 bytes of image traffic per pixel)."""
import argparse
import json
import sys
import time

import numpy as np

sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ipt_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--sources", type=int, default=2000)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    W = H = a.size
    sc = capi.Scene(capi.SceneDescription("box"))
    plane = capi.Plane(sc, W, H)
    plane.render(capi.default_params(width=W, height=H, pass_count=16, plane_mode=capi.PLANE_GUI))
    s, q, c = plane.download()
    mean = (s / np.maximum(c, 1)).astype(np.float32)
    cutoff = float(np.sort(mean.ravel())[-a.sources - 1])  # exactly `sources` pixels above it (up to ties)
    nb = int((mean > np.float32(cutoff)).sum())
    for _ in range(3):
        plane.display(cutoff)
    ms, wall = [], []
    for _ in range(a.reps):
        t = time.perf_counter()
        _, m = plane.display(cutoff)
        wall.append((time.perf_counter() - t) * 1e3); ms.append(m)
    t = time.perf_counter()
    for _ in range(a.reps):
        by = plane.save_bytes()
    save_ms = (time.perf_counter() - t) * 1e3 / a.reps
    pairs = W * H * nb
    out = {"stage": "normalize(glare(image, cutoff)) + Gui::save bytes", "size": [W, H], "halo_sources": nb, "pairs": pairs,
           "gpu_display_device_ms": float(np.median(ms)), "gpu_display_e2e_ms": float(np.median(wall)),
           "gpu_gpairs_per_s": pairs / (float(np.median(ms)) * 1e-3) / 1e9, "gpu_save_bytes_e2e_ms": save_ms}
    if not a.no_cpu:
        import oracle_lib
        ref = oracle_lib.load_ref()
        chk, kind = (ref, "reference") if ref is not None else (oracle_lib.load_oracle(), "port")
        t = time.perf_counter(); g = chk.image_glare(mean, cutoff); shown = chk.image_normalize(g); cpu_ms = (time.perf_counter() - t) * 1e3
        t = time.perf_counter(); cb = chk.image_save_bytes(mean); cpu_save_ms = (time.perf_counter() - t) * 1e3
        got, _ = plane.display(cutoff)
        ulp = np.abs(got.view(np.uint32).astype(np.int64) - shown.view(np.uint32).astype(np.int64)).max()
        out.update(cpu_kind=kind, cpu_cores=1, cpu_display_ms=cpu_ms, cpu_gpairs_per_s=pairs / (cpu_ms * 1e-3) / 1e9, cpu_save_bytes_ms=cpu_save_ms,
                   display_max_ulp_vs_cpu=int(ulp), save_bytes_equal=bool(np.array_equal(by, cb)),
                   glare_bit_equal=bool(np.array_equal(capi.image_glare(mean, cutoff)[0].view(np.uint32), g.view(np.uint32))))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
