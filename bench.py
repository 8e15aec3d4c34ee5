#!/usr/bin/env python
"""bench.py — the headline benchmark of BASELINE.json: path-tracing throughput (Mpaths/s, with Mrays/s beside it) of
the trace loop on BASELINE configs[1]: the Cornell-box-style scene ("cornell"), 1024x1024, diffuse+glossy DDFs,
split schedule 16/8/4/2, depth 4, 1024 spp for the whole job.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one call of the hot path over one batch: `--passes-per-step` passes (default 8) of the full 1024x1024
frame = 8.4 M camera paths. The default K = 128 steps is the whole 1024-spp job. N GPUs: one process per GPU
(torchrun), every rank renders its own pass range (Philox counters make them disjoint), weak scaling; the only
collective is the all-reduce of the accumulators (sum, sumsq, count) at the end of the timed region.

Printed JSON keys are described in DESIGN.md §Measurement. `--impl reference` times the reference's own CPU
implementation (oracle/_ref, the unmodified dimalit/ipt sources + the C2 plug-in; else the oracle port) on all host
cores with forked single-threaded workers (drand48 is process-global).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

# The bench line is BASELINE configs[1] ("c2"); the other configs are parity-test cases, selectable for the record.
WORKLOADS = {
    "c1": dict(config="configs[0]", scene="box", width=640, height=640, depth_max=4, schedule=[16, 8, 4, 2], spp_total=1024, passes_per_step=16,
               text="default scene of src/sample_scenes.cpp (make_scene_box: open unit box + sphere + one square light)"),
    "c2": dict(config="configs[1]", scene="cornell", width=1024, height=1024, depth_max=4, schedule=[16, 8, 4, 2], spp_total=1024, passes_per_step=8,
               text="Cornell-box-style scene 'cornell' (5 box planes, Lambert sphere, glossy sphere, ceiling area light)"),
    "c3": dict(config="configs[2]", scene="mesh:1000000", width=1920, height=1080, depth_max=8, schedule=[1] * 8, spp_total=256, passes_per_step=4,
               text="procedural 1M-triangle random mesh inside the open box, max depth 8"),
    "c4": dict(config="configs[3]", scene="mesh:10000000", width=3840, height=2160, depth_max=8, schedule=[1] * 8, spp_total=4096, passes_per_step=1,
               text="10M-triangle synthetic mesh, sample-sharded across the GPUs"),
    "c5": dict(config="configs[4]", scene="lightgrid:100x100", width=2048, height=2048, depth_max=4, schedule=[16, 8, 4, 2], spp_total=16, passes_per_step=1,
               tile=(768, 768, 512, 512), text="many-light CollectionLighting scene (10k emitters), 512x512 crop of the 2048x2048 frame per step"),
}
WORKLOAD = dict(WORKLOADS["c2"])
CPU_SAMPLE = dict(width=128, height=128)  # the CPU legs render the same view/estimator at this frame size


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


# ---------------------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes anything under oracle/)
# ---------------------------------------------------------------------------------------------------------------
_worker_state = {}


def _cpu_worker_init(kind, counter):
    import oracle_lib

    with counter.get_lock():
        rank = counter.value
        counter.value += 1
    W, H = CPU_SAMPLE["width"], CPU_SAMPLE["height"]
    if kind == "reference":
        ref = oracle_lib.load_ref()
        h = ref.scene(WORKLOAD["scene"])
        ref.set_tree(WORKLOAD["schedule"][0], WORKLOAD["depth_max"])
        ref.seed(1000 + rank)
        _worker_state.update(kind=kind, ref=ref, scene=h, W=W, H=H)
    else:
        from ipt_b200 import capi

        orc = oracle_lib.load_oracle()
        sd = capi.SceneDescription(WORKLOAD["scene"])
        orc.seed(1000 + rank)
        p = capi.default_params(width=W, height=H, depth_max=WORKLOAD["depth_max"], schedule=WORKLOAD["schedule"], pass_count=1)
        _worker_state.update(kind=kind, orc=orc, sd=sd, p=p, W=W, H=H)


def _cpu_worker_pass(_):
    st = _worker_state
    if st["kind"] == "reference":
        r = st["ref"].render(st["scene"], 1, st["W"], st["H"], verbatim=False)
        return st["W"] * st["H"], int(r["rays"])
    import oracle_lib

    r = st["orc"].render(st["sd"].ptr, st["p"], oracle_lib.RNG_DRAND48, 0)
    return st["W"] * st["H"], int(r["rays"])


class CpuPool:
    """nproc forked single-threaded workers, each with its own drand48 stream (srand48(1000+rank))."""

    def __init__(self):
        import oracle_lib

        self.kind = "reference" if oracle_lib.build_ref() is not None else "port"
        if self.kind == "port":
            oracle_lib.build_oracle()
        self.cores = os.cpu_count() or 1
        try:
            self.cores = len(os.sched_getaffinity(0))
        except Exception:
            pass
        ctx = mp.get_context("fork")
        counter = ctx.Value("i", 0)
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)  # the reference prints "NEW POOL ..." from its allocator (ddf.cpp:28)
        try:
            self.pool = ctx.Pool(self.cores, initializer=_cpu_worker_init, initargs=(self.kind, counter))
            self.pool.map(_cpu_worker_pass, range(self.cores))  # warm every worker (allocator pools, page-in)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(devnull)
            os.close(saved)

    def step(self, passes_per_worker=1):
        """One bounded sample: every worker renders `passes_per_worker` passes of the CPU_SAMPLE frame."""
        t0 = time.perf_counter()
        out = self.pool.map(_cpu_worker_pass, range(self.cores * passes_per_worker), chunksize=passes_per_worker)
        dt = time.perf_counter() - t0
        return sum(o[0] for o in out), sum(o[1] for o in out), dt

    def close(self):
        self.pool.close()
        self.pool.join()

    def sample_text(self, steps, passes_per_worker=1):
        return (f"{WORKLOAD['scene']} view at {CPU_SAMPLE['width']}x{CPU_SAMPLE['height']}, same estimator (16/8/4/2, depth 4), "
                f"{steps} x {passes_per_worker} pass(es) on each of {self.cores} forked single-threaded processes, libc drand48")


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each step is a bounded sample (one pass of the CPU_SAMPLE frame per host core); shrink the frame for long runs so
    # that `--steps K --warmup W` still ends within a few minutes (the rate does not depend on the frame size)
    total = args.steps + args.warmup
    side = 128 if total <= 20 else (96 if total <= 48 else 64)
    CPU_SAMPLE.update(width=side, height=side)
    pool = CpuPool()
    for _ in range(args.warmup):
        pool.step()
    paths = rays = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        p, r, _ = pool.step()
        paths += p
        rays += r
    dt = time.perf_counter() - t0
    value = paths / dt / 1e6
    line = {
        "impl": "reference", "metric": "path-tracing throughput", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(), "parallelism": f"{pool.cores} host processes"},
        "mrays_per_s": rays / dt / 1e6, "rays_per_path": rays / max(paths, 1),
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": pool.cores, "kind": pool.kind, "sample": pool.sample_text(args.steps),
                         "cpu": cpu_model()},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    pool.close()
    print(json.dumps(line), flush=True)


def workload_text():
    w = WORKLOAD
    return (f"BASELINE {w['config']}: {w['text']}, {w['width']}x{w['height']}, {w['spp_total']} spp job, "
            f"split schedule {'/'.join(map(str, w['schedule']))}, depth {w['depth_max']}")


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), threading.Event(), None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        self.stop_flag.set()
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from ipt_b200 import build, capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    build.build()
    lib = capi.load()
    w = WORKLOAD
    W, H = w["width"], w["height"]
    pps = args.passes_per_step
    sd = capi.SceneDescription(w["scene"])
    sc = capi.Scene(sd, local)
    # accumulators are torch tensors so that NCCL can reduce them in place; the library accumulates into them
    acc_sum = torch.zeros(H * W, dtype=torch.float32, device="cuda")
    acc_sq = torch.zeros(H * W, dtype=torch.float32, device="cuda")
    acc_cnt = torch.zeros(H * W, dtype=torch.int32, device="cuda")
    plane = capi.Plane(sc, W, H, wrap=(acc_sum.data_ptr(), acc_sq.data_ptr(), acc_cnt.data_ptr()))
    total_steps = args.warmup + args.steps
    first_pass = rank * total_steps * pps  # disjoint pass ranges per rank

    tile = w.get("tile")
    paths_per_pass = (tile[2] * tile[3]) if tile else W * H

    def params(step, flags=capi.FLAG_TIME_KERNELS):
        kw = dict(tile_x0=tile[0], tile_y0=tile[1], tile_w=tile[2], tile_h=tile[3]) if tile else {}
        return capi.default_params(width=W, height=H, depth_max=w["depth_max"], schedule=w["schedule"], seed=args.seed,
                                   pass_begin=first_pass + step * pps, pass_count=pps, flags=flags, batch_paths=args.batch_paths, **kw)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput: the scene is in HBM, accumulators stay in HBM ----
    for s in range(args.warmup):
        plane.render(params(s))
    acc_sum.zero_(); acc_sq.zero_(); acc_cnt.zero_()
    sampler = ClockSampler(local)
    sampler.start()
    agg = dict(paths=0, rays=0, launches=0, ms_dev=0.0, ms_ext=0.0, ms_shade=0.0, ms_gen=0.0, ms_acc=0.0, n_ext=0, n_shade=0, queued=0,
               surface=0, light=0, queue_bytes=0, nodes=0, tris=0, fused=0)
    sync()
    t0 = time.perf_counter()
    for s in range(args.steps):
        st = plane.render(params(args.warmup + s))
        agg["paths"] += st.paths; agg["rays"] += st.rays; agg["launches"] += st.kernel_launches; agg["ms_dev"] += st.ms_total
        agg["ms_ext"] += st.ms_extend; agg["ms_shade"] += st.ms_shade; agg["ms_gen"] += st.ms_generate; agg["ms_acc"] += st.ms_accumulate
        agg["n_ext"] += st.n_extend; agg["n_shade"] += st.n_shade; agg["surface"] += st.surface_hits; agg["light"] += st.light_hits
        agg["queue_bytes"] += st.queue_bytes; agg["fused"] += st.rays_resolved_in_shade; agg["nodes"] += st.bvh_nodes_visited; agg["tris"] += st.triangles_tested
    if world > 1:  # the only collective of the path: reduce the accumulators over NVLink
        dist.all_reduce(acc_sum); dist.all_reduce(acc_sq); dist.all_reduce(acc_cnt)
    sync()
    elapsed = time.perf_counter() - t0
    clocks = sampler.result()
    t = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
    tot = torch.tensor([agg["paths"], agg["rays"], agg["launches"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    elapsed = float(t.item())
    paths_all, rays_all, launches_all = (float(x) for x in tot.tolist())
    image_mean = float(acc_sum.sum().item() / max(int(acc_cnt.sum().item()), 1))

    # ---- end to end through the C ABI with HOST buffers: per step camera in, sum/sumsq/count out ----
    cam = sd.desc.camera
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    # the caller's result buffers: pinned host memory, reused every step
    host_out = (torch.empty(H * W, dtype=torch.float32).pin_memory().numpy(), torch.empty(H * W, dtype=torch.float32).pin_memory().numpy(),
                torch.empty(H * W, dtype=torch.int32).pin_memory().numpy().view(np.uint32))
    sc.render_host(params(args.warmup, flags=0), out=host_out)  # warm the host plane / staging path
    sync()
    t0 = time.perf_counter()
    e2e_paths = 0
    for s in range(e2e_steps):
        capi.check(lib.ipt_scene_set_camera(sc.handle, C.byref(cam)))
        hs, hq, hc, st = sc.render_host(params(args.warmup + s, flags=0), out=host_out)
        e2e_paths += st.paths
        if world > 1:  # N GPUs: the per-rank results are merged where the user reads them
            part = torch.from_numpy(hs).cuda()
            dist.all_reduce(part)
            hs = part.cpu().numpy()
        _ = float(hs[0, 0])
    sync()
    e2e_elapsed = time.perf_counter() - t0
    t = torch.tensor([e2e_elapsed], dtype=torch.float64, device="cuda")
    ep = torch.tensor([e2e_paths], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ep)
    e2e_value = float(ep.item()) / float(t.item()) / 1e6

    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant kernel and its algorithmic bytes: the wavefront MODEL of SURVEY.md §8d — every ray
        # is a 36 B record written once and read once, every queued hit a 32 B record written once and read once.
        # The fused shade kernels trace the children they spawn, so those rays are never queued: a fused launch is
        # credited with the model bytes of the two kernels it replaces (36 B write + 36 B read per such ray, the hit
        # record it appends, 8 B per light hit), and `implemented_bytes_per_launch` says what the implementation itself
        # has to move (hit records in and out, path values).
        children = agg["rays"] - agg["paths"]
        queued = (agg["queue_bytes"] - 72 * agg["rays"] - 8 * agg["paths"]) // 64
        fused = agg["fused"]
        hits_by_shade = max(queued - agg["paths"], 0) if fused == children and fused else 0  # all but the camera rays' hits
        bytes_ext = 36 * (agg["rays"] - fused) + 32 * (queued - hits_by_shade) + 8 * agg["light"] * (0 if fused else 1) + 64 * (agg["nodes"] + agg["tris"])
        bytes_shade = 32 * queued + 36 * children + 36 * fused + 32 * hits_by_shade + 8 * agg["light"] * (1 if fused else 0)
        bytes_shade_impl = 32 * queued + 36 * (children - fused) + 32 * hits_by_shade + 8 * agg["light"] * (1 if fused else 0)
        dom = "shade" if agg["ms_shade"] >= agg["ms_ext"] else "extend"
        dom_ms = agg["ms_shade"] if dom == "shade" else agg["ms_ext"]
        dom_n = agg["n_shade"] if dom == "shade" else agg["n_ext"]
        dom_bytes = bytes_shade if dom == "shade" else bytes_ext
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        traffic = None
        prof = ROOT / "profiles" / "roofline_latest.json"
        if prof.exists() and args.workload == "c2" and not args.batch_paths:  # the ncu capture is of this workload at the default batch
            try:
                traffic = json.loads(prof.read_text()).get(dom, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "path-tracing throughput", "value": paths_all / elapsed / 1e6, "unit": "Mpaths/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_text(), "passes_per_step": pps, "paths_per_step_per_gpu": paths_per_pass * pps,
                       "parallelism": f"pass-sharded x{world}" if world > 1 else "single GPU",
                       "l2": "per-step ray/hit queue working set (>1 GB) exceeds the 126 MB L2; no flush needed",
                       "batch_paths": args.batch_paths or "library default (2^28 / widest queued tree level = 2^21 paths)"},
            "mrays_per_s": rays_all / elapsed / 1e6, "rays_per_path": rays_all / max(paths_all, 1), "image_mean": image_mean,
            "device_ms_per_step": agg["ms_dev"] / args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": C.sizeof(capi.Camera) + C.sizeof(capi.RenderParams),
                    "d2h_bytes_per_step": 12 * W * H, "steps": e2e_steps,
                    "call": "ipt_scene_set_camera + ipt_render_host (host sum/sumsq/count buffers)"},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "kernel": ("k_extend_mesh" if dom == "extend" and w["scene"].startswith("mesh") else f"k_{dom}"), "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "launches": dom_n, "avg_launch_ms": dom_ms / max(dom_n, 1),
                         "algorithmic_bytes_per_launch": dom_bytes / max(dom_n, 1),
                         "implemented_bytes_per_launch": (bytes_shade_impl if dom == "shade" else bytes_ext) / max(dom_n, 1),
                         "model": "SURVEY 8d wavefront queue model: 36 B ray record written + read per ray, 32 B hit record written + read per queued hit",
                         "whole_pipeline": {"bytes": agg["queue_bytes"] + 64 * (agg["nodes"] + agg["tris"]),
                                            "achieved_gbs": (agg["queue_bytes"] + 64 * (agg["nodes"] + agg["tris"])) / (agg["ms_dev"] * 1e-3) / 1e9,
                                            "frac": (agg["queue_bytes"] + 64 * (agg["nodes"] + agg["tris"])) / (agg["ms_dev"] * 1e-3) / 1e9 / peak},
                         "kernel_ms": {"generate": agg["ms_gen"], "extend": agg["ms_ext"], "shade": agg["ms_shade"], "accumulate": agg["ms_acc"]}},
        }
        if world == 1 and not args.no_cpu_baseline and w["scene"] not in ("box", "cornell"):
            line["cpu_baseline"] = {"value": None, "unit": "Mpaths/s", "cores": 0, "kind": "port",
                                    "sample": "not run: the reference has no mesh / 10k-light scene; see the c1/c2 workloads"}
        elif world == 1 and not args.no_cpu_baseline:
            pool = CpuPool()
            paths = rays = 0
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < args.cpu_seconds or reps < 2:
                p, r, _ = pool.step()
                paths += p; rays += r; reps += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": paths / dt / 1e6, "unit": "Mpaths/s", "cores": pool.cores, "kind": pool.kind,
                                    "sample": pool.sample_text(reps), "mrays_per_s": rays / dt / 1e6, "cpu": cpu_model()}
            pool.close()
        print(json.dumps(line), flush=True)
    plane.close()
    sc.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="c2 = BASELINE configs[1], the bench line")
    ap.add_argument("--passes-per-step", type=int, default=0)
    ap.add_argument("--batch-paths", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=16)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    WORKLOAD.clear()
    WORKLOAD.update(WORKLOADS[args.workload])
    if args.passes_per_step <= 0:
        args.passes_per_step = WORKLOAD["passes_per_step"]
    if args.impl == "reference":
        if args.steps == 128:
            args.steps = 8  # each step is a bounded sample: keep the default run within minutes
        run_reference_arm(args)
    else:
        if args.warmup < 3:
            args.warmup = 3  # timing rule: at least 3 warm-up steps
        run_ours(args)


if __name__ == "__main__":
    main()
