"""Round-2 A/B of the fused shade kernels: builds -D variants of the same sources (here, on CPU) and benches each on the GPU box.

    python tools/ab_r02.py build [names]     # in the container (cross-compiles)
    python tools/ab_r02.py run [names] [workloads]   # on the GPU box: two interleaved repetitions per variant
"""
import json, os, subprocess, sys
sys.path.insert(0, '.')
VARIANTS = {
    "base": [],
    "nofast": ["IPT_FAST_SECONDARY=0"],
    "u24": ["IPT_U01_BITS=24"],
    "r7": ["IPT_PHILOX_ROUNDS=7"],
    "r7b4": ["IPT_PHILOX_ROUNDS=7", "IPT_SHADE_FUSED_MIN_BLOCKS=4", "IPT_SHADE_NEXT_MIN_BLOCKS=4"],
    "nearr1": ["IPT_FAST_SECONDARY=0", "IPT_U01_BITS=24"],
}
if __name__ == "__main__":
    what = sys.argv[1]
    names = sys.argv[2].split(",") if len(sys.argv) > 2 and sys.argv[2] != "all" else list(VARIANTS)
    if what == "build":
        from ipt_b200 import build
        for n in names:
            print(n, build.build_variant(n, VARIANTS[n]), flush=True)
    else:
        workloads = sys.argv[3].split(",") if len(sys.argv) > 3 else ["c2"]
        for rep in range(2):
            for n in names:
                for w in workloads:
                    env = dict(os.environ, IPT_B200_LIB=f"ipt_b200/lib/variants/{n}.so")
                    r = subprocess.run([sys.executable, "bench.py", "--workload", w, "--steps", "12", "--warmup", "3", "--no-cpu-baseline", "--e2e-steps", "1"],
                                       env=env, capture_output=True, text=True)
                    try:
                        d = json.loads(r.stdout.strip().splitlines()[-1])
                        k = d["roofline"]["kernel_ms"]
                        print(f"{n:10s} {w}  {d['value']:7.1f} Mpaths/s  extend {k['extend']:7.1f}  shade {k['shade']:7.1f}  rays/path {d['rays_per_path']:.3f} mean {d['image_mean']:.6f} sm {d['clocks']['sm_mhz']}", flush=True)
                    except Exception as e:
                        print(n, w, "failed", r.stderr[-600:], flush=True)
