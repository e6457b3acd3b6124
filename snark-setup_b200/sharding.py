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
    """Element range a rank must READ for its partial power_pairs (api_ratio.inl does the same arithmetic): its own
    elements shard_range(n) plus ONE element of overlap with the next shard, because the pair (i, i+1) belongs to the
    owner of element i.  The last non-empty shard reads no overlap; a shard that ends up with one element and no
    successor owns no pair."""
    s, e = shard_range(n, rank, world)
    if e == s:
        return s, s
    return (s, e + 1) if e < n else (s, e)


def vector_offsets(g1_count: int, other_count: int, compressed: bool, g1_sizes=(96, 48), g2_sizes=(192, 96)):
    """[(byte offset, element count, element size)] of tau_g1, tau_g2, alpha_g1, beta_g1, beta_g2 in a Groth16
    accumulator buffer (phase1/src/helpers/buffers.rs:293-341); sizes default to BLS12-377 (uncompressed, compressed)."""
    s1, s2 = g1_sizes[1 if compressed else 0], g2_sizes[1 if compressed else 0]
    out, off = [], 64
    for cnt, sz in ((g1_count, s1), (other_count, s2), (other_count, s1), (other_count, s1), (1, s2)):
        out.append((off, cnt, sz))
        off += cnt * sz
    return out


def shard_bytes(n: int, rank: int, world: int, element_size: int) -> int:
    """Bytes of one vector that belong to `rank`'s shard."""
    s, e = shard_range(n, rank, world)
    return (e - s) * element_size


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
