"""Multi-GPU sharding of the trace loop (harness glue; one process per GPU, torch.distributed for the plumbing).

The path shards by construction: passes are independent (the reference already runs whole-frame passes in parallel
threads and merges them by a running mean, main.cpp:258-277, gui.cpp:165-182), and the Philox counter carries
(pixel, pass, node), so rank r simply renders passes [begin_r, begin_r + count_r) of the same frame. The only exchange
is the final reduction of the accumulators (sum f32, sumsq f32, count u32: 12 B per pixel).
"""
from __future__ import annotations


def shard_passes(total_passes: int, world: int, rank: int, first_pass: int = 0):
    """Contiguous, disjoint, balanced pass ranges: returns (pass_begin, pass_count) of `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank outside the world")
    base, extra = divmod(total_passes, world)
    count = base + (1 if rank < extra else 0)
    begin = first_pass + rank * base + min(rank, extra)
    return begin, count


def shard_tiles(width: int, height: int, world: int, rank: int):
    """Alternative: horizontal bands of loop rows (tile_x0, tile_y0, tile_w, tile_h); bounds queue memory per GPU."""
    begin, count = shard_passes(height, world, rank)
    return 0, begin, width, count


def allreduce_accumulators(dist, acc_sum, acc_sumsq, acc_count):
    """The path's only collective: element-wise SUM of the three accumulators over all ranks (NCCL over NVLink on the
    GPU box, gloo in the CPU tests). Tensors are reduced in place."""
    dist.all_reduce(acc_sum)
    dist.all_reduce(acc_sumsq)
    dist.all_reduce(acc_count)
