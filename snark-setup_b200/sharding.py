"""Index-range sharding of the hot path across ranks / GPUs.

The path partitions with no data-path collective: contribute elements are independent; a ratio check
splits into per-shard (s, sx) partial sums that the host adds, shards overlapping by ONE element
because power_pairs pairs v[i] with v[i+1] (setup-utils/src/helpers.rs:388-390, the same reason
iter_chunk overlaps its windows, phase1/src/helpers/buffers.rs:54).  Every vector is split separately
into equal contiguous parts: indices < 2^k carry 4 G1 + 1 G2 multiplications and indices >= 2^k only one,
so the reference's equal-index chunks would be 5:1 imbalanced.
"""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """[start, end) of `rank`'s contiguous, balanced share of n elements."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def ratio_shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Element range a rank must READ to produce its partial power_pairs: its pairs (i, i+1) for i in
    shard_range(n-1), i.e. one element of overlap with the next shard.  Empty when it owns no pair."""
    if n < 2:
        return 0, 0
    s, e = shard_range(n - 1, rank, world)
    return (s, e + 1) if e > s else (s, s)


def contribute_plan(g1_count: int, other_count: int, first_power: int, rank: int, world: int):
    """Per-vector (element_start, element_end, first_power) for one rank: tau_g1, tau_g2, alpha_g1, beta_g1.
    beta_g2 (one element) belongs to rank 0."""
    out = []
    for n in (g1_count, other_count, other_count, other_count):
        s, e = shard_range(n, rank, world)
        out.append((s, e, first_power + s))
    return out


def max_over_ranks(x: float, dist=None, device=None) -> float:
    """Timing reduction used by bench.py: the slowest rank defines the step."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
