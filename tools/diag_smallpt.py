"""Diagnosis of the smallpt image bias (GPU box): (1) rays that START on surfaces (the self-intersection regime of
GeometrySmallPt's eps = 1e-4 on radius-1000 spheres), device vs oracle bit for bit; (2) common-random-number renders at
growing depth: per-depth ray counts and image sums, device vs oracle."""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import oracle_lib
from ipt_b200 import capi

scene = sys.argv[1] if len(sys.argv) > 1 else "smallpt"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 16
capi.load()
oracle = oracle_lib.load_oracle()
sd = capi.SceneDescription(scene)
sc = capi.Scene(sd)
rng = np.random.default_rng(3)
xy = rng.random((20000, 2)).astype(np.float32)
o, d = oracle.camera_rays(sd.ptr, xy)
first = oracle.trace_batch(sd.ptr, o, d)
hit = first["outcome"] == 1
pos, nrm = first["pos"][hit], first["normal"][hit]
for rep in range(3):
    v = rng.normal(size=pos.shape).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True).astype(np.float32)
    flip = (v * nrm).sum(1) < 0
    v[flip] = -v[flip]
    v = v.astype(np.float32)
    g = sc.trace_batch(pos, v); c = oracle.trace_batch(sd.ptr, pos, v)
    same_prim = g["prim"] == c["prim"]; same_t = g["t"].view(np.uint32) == c["t"].view(np.uint32); same_out = g["outcome"] == c["outcome"]
    tiny = (c["t"] < 1e-2) & (c["outcome"] == 1)
    print(f"on-surface rays {len(pos)}: prim equal {same_prim.mean():.6f} t bits equal {same_t.mean():.6f} outcome equal {same_out.mean():.6f} self-hits(t<1e-2) oracle {tiny.mean():.4f} device {((g['t'] < 1e-2) & (g['outcome'] == 1)).mean():.4f}")
    # second generation: from the hits of these rays
    ok = c["outcome"] == 1
    pos, nrm = c["pos"][ok], c["normal"][ok]

for depth, sched in ((2, [16, 8]), (3, [16, 8, 4]), (4, [16, 8, 4, 2])):
    p = capi.default_params(width=48, height=48, pass_count=passes, depth_max=depth, schedule=sched, flags=capi.FLAG_KEEP_ZERO_WEIGHT, seed=11)
    t0 = time.time()
    s, q, cnt, st = sc.render_host(p)
    oo = oracle.render(sd.ptr, p, oracle_lib.RNG_PHILOX, 0)
    diff = s.astype(np.float64) - oo["sum"]
    scale = max(oo["sum"].max(), 1e-12)
    same = np.abs(diff) / scale <= 1e-5
    print(f"depth {depth}: identical {same.mean():.4f} sum gpu/cpu {s.sum() / oo['sum'].sum():.6f} diff mean {diff.mean():.4g} +- {diff.std() / np.sqrt(diff.size):.3g} (image mean {oo['sum'].mean():.4g})")
    print("   rays_at_depth gpu", list(st.rays_at_depth[:depth]), "cpu", list(oo["rays_at_depth"][:depth]), "ratio", [round(a / max(b, 1), 5) for a, b in zip(st.rays_at_depth[:depth], oo["rays_at_depth"][:depth])])
    print(f"   gpu light_hits {st.light_hits} surface {st.surface_hits} miss {st.misses} failed {st.failed_samples} dropped {st.nonfinite_dropped}  ({time.time() - t0:.1f} s)")
sc.close()
