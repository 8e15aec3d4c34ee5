"""The N>1 path on CPU: world_size 2 over gloo. Each rank produces the accumulators of its pass range (here with the
CPU oracle standing in for the device kernels, as the checker) and the all-reduce must reproduce the single-process
result; pass ranges must be disjoint and cover the job."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

from ipt_b200 import sharding

ROOT = Path(__file__).resolve().parent.parent


def test_shard_passes_partition():
    for total in [0, 1, 7, 8, 1024, 4096 + 3]:
        for world in [1, 2, 3, 4, 8]:
            seen = []
            for r in range(world):
                b, c = sharding.shard_passes(total, world, r, first_pass=5)
                seen += list(range(b, b + c))
            assert seen == list(range(5, 5 + total))
            counts = [sharding.shard_passes(total, world, r)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        sharding.shard_passes(8, 2, 2)
    assert sharding.shard_tiles(640, 641, 2, 1) == (0, 321, 640, 320)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist

    import oracle_lib
    from ipt_b200 import capi, sharding as sh

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    capi.load()
    orc = oracle_lib.load_oracle()
    sd = capi.SceneDescription("box")
    W = H = 24
    total = 5
    begin, count = sh.shard_passes(total, world, rank)
    p = capi.default_params(width=W, height=H, pass_begin=begin, pass_count=count, schedule=[4, 2, 1, 1], seed=3)
    r = orc.render(sd.ptr, p, oracle_lib.RNG_PHILOX, 0)
    s = torch.from_numpy(r["sum"].astype(np.float32)); q = torch.from_numpy(r["sumsq"].astype(np.float32))
    c = torch.from_numpy(r["counters"].astype(np.int32))
    sh.allreduce_accumulators(dist, s, q, c)
    if rank == 0:
        np.savez(Path(out_dir) / "merged.npz", sum=s.numpy(), sumsq=q.numpy(), count=c.numpy())
    dist.destroy_process_group()


def test_two_rank_allreduce_reproduces_single_process(tmp_path, lib, oracle):
    import torch.multiprocessing as tmp_mp

    import oracle_lib
    from ipt_b200 import capi

    port = 29500 + os.getpid() % 2000
    tmp_mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    merged = np.load(tmp_path / "merged.npz")
    sd = capi.SceneDescription("box")
    p = capi.default_params(width=24, height=24, pass_begin=0, pass_count=5, schedule=[4, 2, 1, 1], seed=3)
    one = oracle.render(sd.ptr, p, oracle_lib.RNG_PHILOX, 0)
    assert np.array_equal(merged["count"].astype(np.uint64), one["counters"])
    assert np.allclose(merged["sum"], one["sum"], rtol=1e-6, atol=1e-7)
    assert np.allclose(merged["sumsq"], one["sumsq"], rtol=1e-6, atol=1e-7)
