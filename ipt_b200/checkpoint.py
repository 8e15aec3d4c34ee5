"""Accumulator checkpoints (SURVEY.md §5 "checkpoint / resume", §8f item 4) — harness glue over ipt_plane_download /
ipt_plane_upload. The reference has no checkpointing (only result.png at exit, main.cpp:288-289); its progressive state
is exactly GridRenderPlane's `pixels` + `pixel_counters`. Because the Philox counter carries the pass index, a render is
resumed by uploading (sum, sumsq, count) and continuing at `next_pass`: the result equals the uninterrupted render.

A checkpoint names the estimator it belongs to (`identity`: scene, depth_max, split schedule, plane mode, ...): resuming it
under other render parameters would silently mix the accumulators of two different estimators, so `load` refuses."""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np


def normalise(path) -> Path:
    """np.savez appends '.npz' to a name without it; the file a later `exists()` / `load()` looks for is this one."""
    p = Path(path)
    return p if p.suffix == ".npz" else p.with_name(p.name + ".npz")


def save(path, plane, next_pass: int, seed: int, identity: dict | None = None, **meta) -> Path:
    """Atomic: written to a temporary file in the same directory and renamed over `path`, so a crash mid-save leaves the
    previous checkpoint intact."""
    path = normalise(path)
    s, q, c = plane.download()
    tmp = path.with_name(path.name + ".tmp.npz")
    ident = {f"id_{k}": np.asarray(v) for k, v in (identity or {}).items()}
    np.savez_compressed(tmp, sum=s, sumsq=q, count=c, next_pass=next_pass, seed=seed, width=plane.width, height=plane.height,
                        **ident, **{k: np.asarray(v) for k, v in meta.items()})
    os.replace(tmp, path)
    return path


def load(path, plane, identity: dict | None = None) -> dict:
    d = np.load(normalise(path))
    if (int(d["width"]), int(d["height"])) != (plane.width, plane.height):
        raise ValueError("checkpoint frame size does not match the plane")
    for k, v in (identity or {}).items():
        key = f"id_{k}"
        if key not in d.files:
            raise ValueError(f"checkpoint does not record '{k}': it cannot be matched to these render parameters")
        if not np.array_equal(d[key], np.asarray(v)):
            raise ValueError(f"checkpoint was written with {k} = {d[key].tolist()!r}, the session uses {np.asarray(v).tolist()!r}")
    plane.upload(d["sum"], d["sumsq"], d["count"])
    return {k: d[k] for k in d.files if k not in ("sum", "sumsq", "count")}
