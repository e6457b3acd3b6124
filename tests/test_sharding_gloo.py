"""N > 1 path on CPU: two gloo ranks shard one small ceremony by index range (no data-path collective),
each computes its slices with the CPU oracle (the checker stands in for the GPU here), rank 0 reassembles
and compares with the unsharded result; the (s, sx) partial sums of a sharded ratio check add up to the
whole-vector power_pairs; max-over-ranks timing reduction works over gloo."""
import os
import socket
import sys

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import random
    import coracle as O
    import pyref as R
    from snark_setup_b200 import sharding as sh

    cv = R.BLS12_377
    k = 4
    prm = R.Phase1Parameters(cv, k, 8)
    rng = random.Random(1234)  # same on every rank
    k0 = [rng.randrange(2, cv.r) for _ in range(3)]
    k1 = [rng.randrange(2, cv.r) for _ in range(3)]
    acc = O.phase1_computation(0, bytes(R.phase1_initialization(prm, False)), prm.get_length(False), False, False, 3,
                               prm.g1_chunk_size, prm.other_chunk_size, 0, *k0)
    offs_in, offs_out = prm.split_offsets(False), prm.split_offsets(True)
    plan = sh.contribute_plan(prm.g1_chunk_size, prm.other_chunk_size, 0, rank, world)
    pieces = []
    for v, (s, e, fp) in enumerate(plan):
        grp = 1 if v == 1 else 0
        o, _, sz = offs_in[v]
        coeff = [None, None, k1[1], k1[2]][v]
        pieces.append(O.apply_powers(0, grp, acc[o + s * sz:o + e * sz], False, 3, True, e - s, tau=k1[0], first_power=fp,
                                     coeff=coeff))
    gathered = [None] * world
    dist.all_gather_object(gathered, pieces)  # test-side reassembly only; the product never gathers points
    # sharded ratio check on tau_g1 of the (uncompressed) accumulator
    g = cv.g1
    o, n, sz = offs_in[0]
    s, e = sh.ratio_shard_range(n, rank, world)
    rho = [rng.randrange(cv.r) for _ in range(n - 1)]
    part = (O.msm(0, 0, acc[o + s * sz:o + (e - 1) * sz], False, e - 1 - s, rho[s:e - 1]),
            O.msm(0, 0, acc[o + (s + 1) * sz:o + e * sz], False, e - 1 - s, rho[s:e - 1]))
    parts = [None] * world
    dist.all_gather_object(parts, part)
    slowest = sh.max_over_ranks(1.0 + rank, dist)
    if rank == 0:
        whole = O.phase1_computation(0, acc, prm.get_length(True), False, True, 3, prm.g1_chunk_size, prm.other_chunk_size,
                                     0, *k1)
        ok = True
        for v in range(4):
            o2, cnt, sz2 = offs_out[v]
            ok = ok and b"".join(gathered[r][v] for r in range(world)) == whole[o2:o2 + cnt * sz2]
        S = SX = None
        for ps, psx in parts:
            S = g.add(S, g.decode(ps, False))
            SX = g.add(SX, g.decode(psx, False))
        ok = ok and g.encode(S, False) == O.msm(0, 0, acc[o:o + (n - 1) * sz], False, n - 1, rho)
        ok = ok and g.encode(SX, False) == O.msm(0, 0, acc[o + sz:o + n * sz], False, n - 1, rho)
        q.put((ok, slowest))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, slowest = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and slowest == 2.0


def test_shard_ranges_cover_exactly():
    sys.path.insert(0, ROOT)
    from snark_setup_b200 import sharding as sh
    for n in (0, 1, 2, 7, 31, 1024, (1 << 21) - 1):
        for world in (1, 2, 3, 4, 8):
            r = [sh.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(e - s for s, e in r) - min(e - s for s, e in r) <= 1
            pairs = sum(max(0, e - s - 1) for s, e in (sh.ratio_shard_range(n, k, world) for k in range(world)))
            assert pairs == max(0, n - 1)
