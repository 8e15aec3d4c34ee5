"""Accumulator checkpoints (SURVEY.md §5 "checkpoint / resume", §8f item 4) — harness glue over ipt_plane_download /
ipt_plane_upload. The reference has no checkpointing (only result.png at exit, main.cpp:288-289); its progressive state
is exactly GridRenderPlane's `pixels` + `pixel_counters`. Because the Philox counter carries the pass index, a render is
resumed by uploading (sum, sumsq, count) and continuing at `next_pass`: the result equals the uninterrupted render."""
from __future__ import annotations

import numpy as np


def save(path, plane, next_pass: int, seed: int, **meta):
    s, q, c = plane.download()
    np.savez_compressed(path, sum=s, sumsq=q, count=c, next_pass=next_pass, seed=seed, width=plane.width, height=plane.height,
                        **{k: np.asarray(v) for k, v in meta.items()})


def load(path, plane) -> dict:
    d = np.load(path)
    if (int(d["width"]), int(d["height"])) != (plane.width, plane.height):
        raise ValueError("checkpoint frame size does not match the plane")
    plane.upload(d["sum"], d["sumsq"], d["count"])
    return {k: d[k] for k in d.files if k not in ("sum", "sumsq", "count")}
