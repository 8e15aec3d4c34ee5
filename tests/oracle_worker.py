"""One worker of helpers.oracle_render_parallel: renders a pass range with the CPU oracle (Philox mode) in a process of its
own and stores the accumulators. Test infrastructure only.

    python tests/oracle_worker.py <scene> <json params> <seed> <pass_begin> <pass_count> <use_bvh> <out.npz>
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

if __name__ == "__main__":
    import oracle_lib
    from ipt_b200 import capi

    scene, kw, seed, begin, count, use_bvh, out = sys.argv[1], json.loads(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), sys.argv[7]
    sd = capi.SceneDescription(scene)
    p = capi.default_params(pass_begin=begin, pass_count=count, seed=seed, **kw)
    o = oracle_lib.load_oracle().render(sd.ptr, p, oracle_lib.RNG_PHILOX, use_bvh)
    np.savez(out, sum=o["sum"], sumsq=o["sumsq"], count=o["counters"], rays=np.uint64(o["rays"]))
