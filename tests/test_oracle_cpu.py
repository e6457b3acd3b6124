"""CPU suite: pins the two oracle restatements (pyref big-int, C++ Montgomery) against each other,
against the committed golden vectors and against the only numeric pins the reference holds
(element sizes, phase1/src/objects/parameters.rs:312-317; buffer sizes)."""
import json
import os
import random

import pytest

import coracle as O
import pyref as R

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "hotpath_vectors.json")))
CURVES = [(0, R.BLS12_377), (1, R.BW6_761)]
GROUPS = [(cid, cv, gid, g) for cid, cv in CURVES for gid, g in ((0, cv.g1), (1, cv.g2))]
IDS = [g.name for _, _, _, g in GROUPS]


def test_reference_size_pins():
    # phase1/src/objects/parameters.rs:312-317
    b, w = R.BLS12_377, R.BW6_761
    assert (b.g1.usize, b.g2.usize, b.g1.csize, b.g2.csize) == (96, 192, 48, 96)
    assert (w.g1.usize, w.g2.usize, w.g1.csize, w.g2.csize) == (192, 192, 96, 96)
    # SURVEY §8 A16 buffer sizes
    for cv, k, acc, con in ((b, 10, 589984, 295600), (b, 20, 603979936, 301990576), (b, 22, 2415919264, 1207960240),
                            (w, 21, 2013265984, 1006633888)):
        p = R.Phase1Parameters(cv, k, 256)
        assert (p.accumulator_size, p.contribution_size) == (acc, con)
    assert R.Phase1Parameters(b, 10, 256).public_key_size == 576
    assert R.Phase1Parameters(w, 10, 256).public_key_size == 864


def test_curve_constants():
    for _, cv, _, g in GROUPS:
        assert g.on_curve(g.gen) and g.in_subgroup(g.gen)
        assert not g.in_subgroup(None) or True


def test_iter_chunk_schedule():
    # phase1/src/helpers/buffers.rs:22-73; SURVEY §8 A5: (0,256),(255,511),...,(2040,2047) — 9 windows
    p = R.Phase1Parameters(R.BLS12_377, 10, 256)
    w = R.iter_chunk(p)
    assert len(w) == 9 and w[0] == (0, 256) and w[1] == (255, 511) and w[-1] == (2040, 2047)
    # chunked mode covers exactly its chunk with one-element overlaps
    p = R.Phase1Parameters(R.BLS12_377, 4, 5, R.CHUNKED_MODE, 1, 8)
    w = R.iter_chunk(p)
    assert w[0][0] == 8 and w[-1][1] == 16
    for (a, b), (c, d) in zip(w, w[1:]):
        assert c == b - 1


def test_chunk_sizes():
    # phase1/src/objects/parameters.rs:248-294
    k = 4
    full = R.Phase1Parameters(R.BLS12_377, k, 8)
    assert (full.g1_chunk_size, full.other_chunk_size) == (31, 16)
    tot1 = tot2 = 0
    for ci in range(4):
        p = R.Phase1Parameters(R.BLS12_377, k, 8, R.CHUNKED_MODE, ci, 8)
        tot1 += p.g1_chunk_size
        tot2 += p.other_chunk_size
    assert (tot1, tot2) == (31, 16)


@pytest.mark.parametrize("cid,cv,gid,g", GROUPS, ids=IDS)
def test_golden_vectors_both_oracles(cid, cv, gid, g):
    v = GOLD["groups"][g.name]
    n = 6
    assert g.encode(g.gen, False).hex() == v["generator_uncompressed"]
    assert g.encode(g.gen, True).hex() == v["generator_compressed"]
    assert g.encode(None, True).hex() == v["infinity_compressed"]
    inu, inc = bytes.fromhex(v["in_uncompressed"]), bytes.fromhex(v["in_compressed"])
    tau, coeff, first = int(v["tau"], 16), int(v["coeff"], 16), v["first_power"]
    powers = [int(p, 16) for p in v["powers"]]
    assert R.generate_powers_of_tau(cv, tau, first, first + n) == powers
    assert O.powers(cid, tau, first, first + n) == powers
    # pyref
    pts = g.read_batch(inc, True)
    assert g.write_batch(pts, False) == inu
    assert g.write_batch(R.batch_exp(g, pts, powers), True).hex() == v["out_plain_compressed"]
    # C++ oracle, from both encodings, implicit and explicit powers
    assert O.apply_powers(cid, gid, inu, False, R.NO, True, n, tau=tau, first_power=first).hex() == v["out_plain_compressed"]
    assert O.apply_powers(cid, gid, inc, True, R.NO, False, n, tau=tau, first_power=first, coeff=coeff).hex() == v["out_coeff_uncompressed"]
    assert O.apply_powers(cid, gid, inu, False, R.NO, False, n, powers=powers, coeff=coeff).hex() == v["out_coeff_uncompressed"]
    rho = [int(r, 16) for r in v["rho"]]
    assert O.msm(cid, gid, inu[:5 * g.usize], False, 5, rho).hex() == v["power_pairs_s"]
    assert O.msm(cid, gid, inu[g.usize:], False, 5, rho).hex() == v["power_pairs_sx"]


@pytest.mark.parametrize("cid,cv", CURVES, ids=["bls12_377", "bw6_761"])
def test_golden_phase1_transcript(cid, cv):
    v = GOLD["phase1"][cv.name]
    p = R.Phase1Parameters(cv, v["power"], v["batch_size"])
    assert (p.accumulator_size, p.contribution_size) == (v["accumulator_size"], v["contribution_size"])
    assert [list(w) for w in R.iter_chunk(p)] == v["windows"]
    keys = [[int(k, 16) for k in ks] for ks in v["keys"]]
    acc0 = bytes(R.phase1_initialization(p, False))
    acc1 = O.phase1_computation(cid, acc0, p.get_length(False), False, False, R.NO, p.g1_chunk_size, p.other_chunk_size, 0, *keys[0])
    assert acc1[64:].hex() == v["challenge1"][128:]
    resp = O.phase1_computation(cid, acc1, p.get_length(True), False, True, R.NO, p.g1_chunk_size, p.other_chunk_size, 0, *keys[1])
    assert resp.hex() == v["response2"]
    # windowed pyref == whole-vector C++ oracle
    assert bytes(R.phase1_computation(p, acc1, False, True, R.NO, *keys[1])) == resp


@pytest.mark.parametrize("cid,cv,gid,g", GROUPS, ids=IDS)
def test_oracles_agree_random_and_errors(cid, cv, gid, g):
    rng = random.Random(99 + cid * 2 + gid)
    n = 10
    pts = [g.mul(g.gen, rng.randrange(1, g.r)) for _ in range(n)]
    pts[7] = None
    exps = [0, 1, cv.r - 1] + [rng.randrange(cv.r) for _ in range(n - 3)]
    want = R.batch_exp(g, pts, exps)
    assert O.apply_powers(cid, gid, g.write_batch(pts, True), True, R.NO, True, n, powers=exps) == g.write_batch(want, True)
    # validation modes
    with pytest.raises(O.OracleError) as ei:
        O.transcode(cid, gid, g.write_batch(pts, True), True, R.ONLY_NON_ZERO, False, n)
    assert (ei.value.code, ei.value.index) == (3, 7)
    good = [p for p in pts if p is not None]
    assert O.transcode(cid, gid, g.write_batch(good, True), True, R.FULL, False, len(good)) == g.write_batch(good, False)
    bad = bytearray(g.write_batch(good, True))
    bad[2 * g.csize + g.csize - 1] |= 0xC0
    with pytest.raises(O.OracleError) as ei:
        O.transcode(cid, gid, bad, True, R.NO, False, len(good))
    assert (ei.value.code, ei.value.index) == (2, 2)
    with pytest.raises(R.UnexpectedFlags):
        g.read_batch(bytes(bad), True)
    # a point on the curve but outside the subgroup
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
    with pytest.raises(O.OracleError) as ei:
        O.transcode(cid, gid, g.write_batch(good[:2] + [P], True), True, R.NO, False, 3, rmul_subgroup=True)
    assert (ei.value.code, ei.value.index) == (4, 2)
    with pytest.raises(R.IncorrectSubgroup):
        R.check_subgroup(g, good[:2] + [P])


def test_c1_host_pieces_key_generation_and_proofs_of_knowledge():
    """oracle/pyref_host.py (RECALLED, UNVERIFIABLE OFFLINE: rand_chacha / ark-ff sampling order): the host-side pieces
    config C1 needs around the hot path — derive_rng_from_seed (seed.rs:7-14), Phase1::key_generation
    (key_generation.rs:8-53), compute_g2_s / hash_to_g2 (helpers.rs:277-291,428-443), PublicKey::write
    (public_key.rs:40-55) — are self-consistent: every generated point is a non-zero subgroup point, the three proofs of
    knowledge verify under the oracle's pairing (verification.rs:83-133) and a tampered key is rejected; the digest of the
    serialized key is pinned so that a drift of the restatement is visible."""
    import hashlib
    import pyref_host as H
    cv = R.BLS12_377
    assert H.chacha20_block(bytes(32), 0)[:2] == [0xade0b876, 0x903df1a0]  # ChaCha20 zero-key block (RFC 8439 family)
    assert H.BLS12_377_G1_COFACTOR == 0x170b5d44300000000000000000000000
    digest = hashlib.blake2b(b"challenge").digest()
    pk, (tau, alpha, beta) = H.key_generation(cv, H.derive_rng_from_seed(b"seed-0"), digest)
    assert 0 < tau < cv.r and 0 < alpha < cv.r and 0 < beta < cv.r and len({tau, alpha, beta}) == 3
    for name, x in (("tau", tau), ("alpha", alpha), ("beta", beta)):
        s, sx = pk[name + "_g1"]
        assert s is not None and cv.g1.on_curve(s) and cv.g1.mul(s, cv.r) is None and cv.g1.mul(s, x) == sx
        q = pk[name + "_g2"]
        assert q is not None and cv.g2.on_curve(q) and cv.g2.mul(q, cv.r) is None
    blob = H.public_key_bytes(cv, pk)
    assert len(blob) == R.Phase1Parameters(cv, 10, 256).public_key_size == 576
    assert hashlib.blake2b(blob).hexdigest()[:32] == "4bbb29864c7a94540671da4ed8bb1dd8"
    assert H.verify_proofs_of_knowledge(cv, pk, digest)
    bad = dict(pk)
    bad["alpha_g2"] = cv.g2.mul(pk["alpha_g2"], 2)
    assert not H.verify_proofs_of_knowledge(cv, bad, digest)
    # the same seed gives the same key; a different transcript digest gives different G2 points
    pk2, k2 = H.key_generation(cv, H.derive_rng_from_seed(b"seed-0"), digest)
    assert H.public_key_bytes(cv, pk2) == blob and k2 == (tau, alpha, beta)


def test_oracle_bucket_msm_small_and_full_top_window():
    """oracle.cpp msm_pippenger (the reference arm's msm_bigint restatement) against the naive sum, and at a size whose
    window width divides the scalar length (2^14 <= n < 2^15 -> c = 11, 253 = 11 * 23): the last signed digit then
    reaches 2^c for scalars >= 2^252 — a bench run that landed on a 2^15-power sample corrupted the heap there."""
    cv, g = R.BLS12_377, R.BLS12_377.g1
    rng = random.Random(2024)
    for n in (1, 31, 200):
        pts = [g.mul(g.gen, rng.randrange(1, cv.r)) for _ in range(n)]
        ks = [rng.randrange(cv.r) for _ in range(n)]
        blob = b"".join(g.encode(p, False) for p in pts)
        assert O.msm_pippenger(0, 0, blob, False, n, ks) == O.msm(0, 0, blob, False, n, ks), n
    n = 20000
    mult = [rng.randrange(1, cv.r) for _ in range(8)]
    enc = [g.encode(g.mul(g.gen, a), False) for a in mult]
    ks = [rng.randrange(1 << 252, cv.r) if i % 3 else rng.randrange(cv.r) for i in range(n)]
    ks[0], ks[1] = cv.r - 1, 1 << 252
    blob = b"".join(enc[i % 8] for i in range(n))
    want = g.encode(g.mul(g.gen, sum(k * mult[i % 8] for i, k in enumerate(ks)) % cv.r), False)
    assert O.msm_pippenger(0, 0, blob, False, n, ks) == want
