"""GPU parity: the CUDA path (through the C ABI) against the pure-Python restatement oracle/pyref.py
on seeded inputs.  Bit-exact: these are integer/byte computations."""
import random

import pytest

import pyref as R
import snark_setup_b200 as S

pytestmark = pytest.mark.gpu

CURVES = [(S.BLS12_377, R.BLS12_377), (S.BW6_761, R.BW6_761)]
GROUPS = [(cid, cv, gid, g) for cid, cv in CURVES for gid, g in ((S.G1, cv.g1), (S.G2, cv.g2))]
IDS = [g.name for _, _, _, g in GROUPS]


def rand_points(g, n, rng):
    acc = g.mul(g.gen, rng.randrange(1, g.r))
    step = g.mul(g.gen, rng.randrange(1, g.r))
    out = []
    for _ in range(n):
        out.append(acc)
        acc = g.add(acc, step)
    return out


@pytest.mark.parametrize("cid,cv,gid,g", GROUPS, ids=IDS)
def test_transcode_roundtrip(cid, cv, gid, g):
    """setup-utils/src/io/mod.rs:33-120: write -> read round trips, compressed and not, + infinity."""
    rng = random.Random(11)
    pts = rand_points(g, 70, rng) + [None]
    for ci in (False, True):
        buf = g.write_batch(pts, ci)
        for co in (False, True):
            got = S.transcode(cid, gid, buf, ci, S.CHECK_NO, co)
            assert got == g.write_batch(pts, co), (ci, co)
    with pytest.raises(S.PointAtInfinity):
        S.transcode(cid, gid, g.write_batch(pts, True), True, S.CHECK_ONLY_NON_ZERO, False)
    ok = g.write_batch(pts[:-1], True)
    assert S.transcode(cid, gid, ok, True, S.CHECK_FULL, False) == g.write_batch(pts[:-1], False)


@pytest.mark.parametrize("cid,cv,gid,g", GROUPS, ids=IDS)
def test_decode_errors(cid, cv, gid, g):
    rng = random.Random(5)
    pts = rand_points(g, 40, rng)
    buf = bytearray(g.write_batch(pts, True))
    bad = bytearray(buf)
    bad[17 * g.csize + g.csize - 1] |= 0xC0
    with pytest.raises(S.UnexpectedFlags) as ei:
        S.transcode(cid, gid, bytes(bad), True, S.CHECK_NO, False)
    assert ei.value.index == 17
    bad = bytearray(buf)
    bad[9 * g.csize:10 * g.csize] = b"\xff" * (g.csize - 1) + b"\x3f"
    with pytest.raises(S.InvalidData) as ei:
        S.transcode(cid, gid, bytes(bad), True, S.CHECK_NO, False)
    assert ei.value.index == 9
    # an x whose x^3+b has no square root
    x = 2
    while True:
        xx = x if g.F.degree == 1 else (x, 1)
        try:
            g.decode(g.F.to_bytes(xx, 0, 2), True, R.NO)
            x += 1
        except R.InvalidData:
            break
    bad = bytearray(buf)
    bad[3 * g.csize:4 * g.csize] = g.F.to_bytes(xx, 0, 2)
    with pytest.raises(S.InvalidData) as ei:
        S.transcode(cid, gid, bytes(bad), True, S.CHECK_NO, False)
    assert ei.value.index == 3


@pytest.mark.parametrize("cid,cv,gid,g", GROUPS, ids=IDS)
def test_apply_powers_matches_oracle(cid, cv, gid, g):
    """phase1/src/computation.rs:323-445: output == batch_exp(generate_powers_of_tau) per element."""
    rng = random.Random(2024)
    n = 96
    pts = rand_points(g, n, rng)
    pts[5] = None
    tau, coeff = rng.randrange(cv.r), rng.randrange(cv.r)
    start = 1000003
    powers = R.generate_powers_of_tau(cv, tau, start, start + n)
    assert S.generate_powers_of_tau(cid, tau, start, start + n) == powers
    for cf in (None, coeff):
        want_pts = R.batch_exp(g, pts, powers, cf)
        for ci in (False, True):
            inp = g.write_batch(pts, ci)
            for co in (False, True):
                got = S.apply_powers(cid, gid, inp, ci, S.CHECK_NO, co, n, tau=tau, first_power=start, coeff=cf)
                assert got == g.write_batch(want_pts, co), (cf is not None, ci, co)
    # explicit scalars (batch_exp), in place, incl. the edge scalars 0, 1, r-1
    exps = [0, 1, cv.r - 1] + [rng.randrange(cv.r) for _ in range(n - 3)]
    bases = bytearray(g.write_batch(pts, False))
    S.batch_exp(cid, gid, bases, exps, coeff)
    assert bytes(bases) == g.write_batch(R.batch_exp(g, pts, exps, coeff), False)
    with pytest.raises(S.InvalidLength):
        S.batch_exp(cid, gid, bases, exps[:-1])
    # batch_mul (phase2 delta^-1)
    bases = bytearray(g.write_batch(pts, False))
    S.batch_mul(cid, gid, bases, coeff)
    assert bytes(bases) == g.write_batch(R.batch_mul(g, pts, coeff), False)


@pytest.mark.parametrize("cid,cv,gid,g", GROUPS, ids=IDS)
def test_check_subgroup(cid, cv, gid, g):
    rng = random.Random(77)
    pts = rand_points(g, 33, rng)
    S.check_subgroup(cid, gid, g.write_batch(pts, True), True)
    S.check_subgroup(cid, gid, g.write_batch(pts, False), False)
    # a curve point outside the r-torsion: random x on the curve (cofactor > 1 for all four groups)
    x = 5
    while True:
        xx = x if g.F.degree == 1 else (x, 3)
        try:
            P = g.decode(g.F.to_bytes(xx, 0, 2), True, R.NO)
            if not g.in_subgroup(P):
                break
        except R.InvalidData:
            pass
        x += 1
    pts[20] = P
    with pytest.raises(S.IncorrectSubgroup) as ei:
        S.check_subgroup(cid, gid, g.write_batch(pts, True), True)
    assert ei.value.index == 20
    with pytest.raises(S.InvalidData):
        S.transcode(cid, gid, g.write_batch(pts, True), True, S.CHECK_FULL, False)


@pytest.mark.parametrize("power,batch", [(3, 4), (5, 16)])
def test_phase1_computation_bls12_377(power, batch):
    """Phase1::computation vs the oracle's restatement, all compression combinations."""
    cv = R.BLS12_377
    rng = random.Random(power)
    rp = R.Phase1Parameters(cv, power, batch)
    sp = S.Phase1Parameters(S.BLS12_377, power, batch)
    assert (sp.accumulator_size, sp.contribution_size) == (rp.accumulator_size, rp.contribution_size)
    t0, a0, b0 = (rng.randrange(1, cv.r) for _ in range(3))
    tau, alpha, beta = (rng.randrange(1, cv.r) for _ in range(3))
    for cin in (False, True):
        acc = bytes(R.phase1_computation(rp, bytes(R.phase1_initialization(rp, False)), False, cin, R.NO, t0, a0, b0))
        for cout in (False, True):
            want = R.phase1_computation(rp, acc, cin, cout, R.NO, tau, alpha, beta)
            out = bytearray(sp.get_length(cout))
            S.phase1_computation(sp, acc, out, cin, cout, S.CHECK_NO, tau, alpha, beta)
            assert bytes(out) == bytes(want), (cin, cout)


@pytest.mark.parametrize("chunk_index", [0, 1, 2, 3])
def test_phase1_computation_chunked_mode(chunk_index):
    """ContributionMode::Chunked (phase1/src/objects/parameters.rs:248-294; computation.rs:60-66): chunk c of
    size 8 of a 2^4-power ceremony; element j of the chunk gets tau^(c*8 + j); chunks >= 2^k hold tau_g1 only."""
    cv = R.BLS12_377
    power, batch, csz = 4, 5, 8
    rng = random.Random(40 + chunk_index)
    rp = R.Phase1Parameters(cv, power, batch, R.CHUNKED_MODE, chunk_index, csz)
    sp = S.Phase1Parameters(S.BLS12_377, power, batch, 1, chunk_index, csz)
    assert (sp.g1_chunk_size, sp.other_chunk_size) == (rp.g1_chunk_size, rp.other_chunk_size)
    assert sp.accumulator_size == rp.accumulator_size and sp.contribution_size == rp.contribution_size
    tau, alpha, beta = (rng.randrange(2, cv.r) for _ in range(3))
    # a chunk-shaped input of distinct points
    acc = bytearray(rp.get_length(False))
    for vec, (o, c, s) in enumerate(rp.split_offsets(False)):
        g = cv.g2 if vec in (1, 4) else cv.g1
        pts = [g.mul(g.gen, rng.randrange(1, cv.r)) for _ in range(c)]
        acc[o:o + c * s] = g.write_batch(pts, False)
    want = R.phase1_computation(rp, bytes(acc), False, True, R.NO, tau, alpha, beta)
    out = bytearray(sp.get_length(True))
    S.phase1_computation(sp, bytes(acc), out, False, True, S.CHECK_NO, tau, alpha, beta)
    assert bytes(out) == bytes(want)


@pytest.mark.parametrize("mode,chunk_index,chunk_size", [(0, 0, 0), (1, 0, 4), (1, 1, 4)])
def test_phase1_marlin(mode, chunk_index, chunk_size):
    """ProvingSystem::Marlin (phase1/src/computation.rs:195-302): tau_g1 over 2^k powers, and on chunk 0 the
    k+2 tau_g2 elements with inverse degree-bound powers and the 3+3k alpha_g1 elements."""
    cv = R.BLS12_377
    power, batch = 3, 8
    rng = random.Random(900 + chunk_index)
    rp = R.Phase1Parameters(cv, power, batch, mode, chunk_index, chunk_size, R.MARLIN)
    sp = S.Phase1Parameters(S.BLS12_377, power, batch, mode, chunk_index, chunk_size, 1)
    assert (sp.accumulator_size, sp.contribution_size) == (rp.accumulator_size, rp.contribution_size)
    tau, alpha = rng.randrange(2, cv.r), rng.randrange(2, cv.r)
    acc = bytearray(rp.get_length(False))
    for vec, (o, c, s) in enumerate(rp.split_offsets(False)):
        g = cv.g2 if vec in (1, 4) else cv.g1
        acc[o:o + c * s] = g.write_batch([g.mul(g.gen, rng.randrange(1, cv.r)) for _ in range(c)], False)
    for cout in (False, True):
        want = R.phase1_computation(rp, bytes(acc), False, cout, R.NO, tau, alpha, 0)
        out = bytearray(sp.get_length(cout))
        S.phase1_computation(sp, bytes(acc), out, False, cout, S.CHECK_NO, tau, alpha, 1)
        assert bytes(out) == bytes(want), cout
    # verification loop: nonzero + subgroup checks and re-emit of every Marlin vector; tau_g1 ratio pair
    resp = bytes(R.phase1_computation(rp, bytes(acc), False, True, R.NO, tau, alpha, 0))
    newc = bytearray(sp.get_length(False))
    S.phase1_verification_vectors(sp, resp, True, newc, False, ratio_check=False)
    assert bytes(newc[64:]) == bytes(R.phase1_computation(rp, bytes(acc), False, False, R.NO, tau, alpha, 0))[64:]


def test_config_c1_complete_transcript_with_public_key():
    """BASELINE.json configs[0] as a TRANSCRIPT: keys from seeds (oracle/pyref_host.py restates derive_rng_from_seed /
    key_generation / hash_to_g2 — recalled, see its header), new -> contribute -> verify at power 10 with the compressed
    PublicKey appended to the response (public_key.rs:40-55, contribute.rs:135-137); the COMPLETE response file hashes
    to the oracle's, the proofs of knowledge verify (verification.rs:83-133) and the four ratio checks pass on the
    device."""
    import hashlib
    import coracle as O
    import pyref_host as H
    cv, cid = R.BLS12_377, S.BLS12_377
    rp, sp = R.Phase1Parameters(cv, 10, 256), S.Phase1Parameters(cid, 10, 256)
    blank = bytearray(S.phase1_initialization(sp, False))
    blank[:64] = hashlib.blake2b(b"").digest()                      # blank_hash (helpers.rs:392-394)
    assert bytes(blank[64:]) == bytes(R.phase1_initialization(rp, False))[64:]
    digest = hashlib.blake2b(bytes(blank)).digest()                 # calculate_hash of the challenge
    pk, keys = H.key_generation(cv, H.derive_rng_from_seed(b"c1-contributor-1"), digest)
    resp = bytearray(sp.contribution_size)
    resp[:64] = digest                                              # the response starts with the challenge hash
    S.phase1_computation(sp, bytes(blank), memoryview(resp)[:sp.get_length(True)], False, True, S.CHECK_NO, *keys)
    resp[sp.contribution_size - sp.public_key_size:] = H.public_key_bytes(cv, pk)
    want = bytearray(rp.contribution_size)
    body = O.phase1_computation(0, bytes(blank), rp.get_length(True), False, True, 3, rp.g1_chunk_size, rp.other_chunk_size,
                                0, *keys)
    want[:len(body)] = body
    want[:64] = digest
    want[rp.contribution_size - rp.public_key_size:] = H.public_key_bytes(cv, pk)
    assert hashlib.blake2b(bytes(resp)).digest() == hashlib.blake2b(bytes(want)).digest()   # the .hash file of the response
    # verification: proofs of knowledge (host, oracle pairing) + the vectors and their four ratio verdicts (device)
    assert H.verify_proofs_of_knowledge(cv, pk, digest)
    newc = bytearray(sp.get_length(False))
    S.phase1_verification_ratios(sp, bytes(resp[:sp.get_length(True)]), True, newc, False, seed=bytes(range(32)))
    assert bytes(newc[64:]) == O.phase1_computation(0, bytes(blank), rp.get_length(False), False, False, 3, rp.g1_chunk_size,
                                                    rp.other_chunk_size, 0, *keys)[64:]
    # first-element checks of verification.rs:136-213: tau_g1[1] = tau * G1 etc. against the public key
    offs = rp.split_offsets(False)
    tau_g1_1 = cv.g1.decode(bytes(newc[offs[0][0] + 96:offs[0][0] + 192]), False)
    assert R.same_ratio(cv, (cv.g1.gen, tau_g1_1), (H.compute_g2_s(cv, H.bls12_377_g2_cofactor(), digest, *pk["tau_g1"], 0), pk["tau_g2"]))


def test_config_c1_bls12_377_power10_batch256():
    """BASELINE.json configs[0]: new + contribute + verify at power 10, batch 256 — the whole response and the
    whole new challenge byte for byte against the C++ oracle, plus the BLAKE2b-512 digests a CLI run would
    write to the .hash files (phase1-cli/src/contribute.rs:143-151)."""
    import hashlib
    import coracle as O
    cv, cid = R.BLS12_377, S.BLS12_377
    rp = R.Phase1Parameters(cv, 10, 256)
    sp = S.Phase1Parameters(cid, 10, 256)
    assert (sp.accumulator_size, sp.contribution_size) == (589984, 295600)
    assert len(sp.iter_chunk()) == 9
    rng = random.Random(1010)
    k0 = [rng.randrange(2, cv.r) for _ in range(3)]
    k1 = [rng.randrange(2, cv.r) for _ in range(3)]
    blank = bytes(R.phase1_initialization(rp, False))
    args = (rp.g1_chunk_size, rp.other_chunk_size, 0)
    chal_o = O.phase1_computation(0, blank, rp.get_length(False), False, False, 3, *args, *k0)
    chal = bytearray(sp.get_length(False))
    S.phase1_computation(sp, blank, chal, False, False, S.CHECK_NO, *k0)
    assert bytes(chal) == chal_o
    resp_o = O.phase1_computation(0, chal_o, rp.get_length(True), False, True, 3, *args, *k1)
    resp = bytearray(sp.get_length(True))
    S.phase1_computation(sp, bytes(chal), resp, False, True, S.CHECK_NO, *k1)
    assert hashlib.blake2b(bytes(resp)).digest() == hashlib.blake2b(resp_o).digest()
    newc_o = O.phase1_computation(0, chal_o, rp.get_length(False), False, False, 3, *args, *k1)
    newc = bytearray(sp.get_length(False))
    pairs = S.phase1_verification_vectors(sp, bytes(resp), True, newc, False, seed=bytes(range(32)))
    assert hashlib.blake2b(bytes(newc)).digest() == hashlib.blake2b(newc_o).digest()
    tau = k0[0] * k1[0] % cv.r
    for (s, sx), grp in zip(pairs, (0, 1, 0, 0)):
        assert O.apply_powers(0, grp, s, False, 3, False, 1, powers=[tau]) == sx


@pytest.mark.parametrize("curve", ["bls12_377", "bw6_761"])
def test_phase1_initialization_is_all_generators(curve):
    """phase1/src/initialization.rs:69-111 (the blank accumulator holds the generators), full and chunked mode,
    compressed and uncompressed, byte for byte against the oracle."""
    cv = R.CURVES[curve]
    cid = S.BLS12_377 if curve == "bls12_377" else S.BW6_761
    for args in ((4, 8, 0, 0, 0), (4, 4, 1, 0, 8), (4, 4, 1, 1, 8), (4, 4, 1, 3, 8)):
        rp = R.Phase1Parameters(cv, *args[:2], *args[2:])
        sp = S.Phase1Parameters(cid, *args[:2], *args[2:])
        for compressed in (False, True):
            assert S.phase1_initialization(sp, compressed) == bytes(R.phase1_initialization(rp, compressed)), (args, compressed)
    # config C1 through the C ABI only: new -> contribute -> verify with the verdict
    sp = S.Phase1Parameters(cid, 5, 16)
    acc0 = S.phase1_initialization(sp, False)
    resp = bytearray(sp.get_length(True))
    S.phase1_computation(sp, acc0, resp, False, True, S.CHECK_NO, 11, 22, 33)
    S.phase1_verification_ratios(sp, bytes(resp), True, bytearray(sp.get_length(False)), False, seed=bytes(32))
