"""In-process multi-GPU sharding behind ONE C-ABI call (ss_init with several devices): results must be
byte-identical to the single-device run, and the summed partial (s, sx) must satisfy the ratio."""
import random

import pytest
import torch

import coracle as O
import pyref as R
import snark_setup_b200 as S
from snark_setup_b200 import ffi

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_devices_one_call():
    cv, cid = R.BLS12_377, S.BLS12_377
    power, batch = 5, 8
    rng = random.Random(77)
    rp = R.Phase1Parameters(cv, power, batch)
    sp = S.Phase1Parameters(cid, power, batch)
    k0 = [rng.randrange(2, cv.r) for _ in range(3)]
    k1 = [rng.randrange(2, cv.r) for _ in range(3)]
    acc0 = bytes(R.phase1_initialization(rp, False))
    try:
        ffi.init([0, 1])
        acc1 = bytearray(sp.get_length(False))
        S.phase1_computation(sp, acc0, acc1, False, False, S.CHECK_NO, *k0)
        resp = bytearray(sp.get_length(True))
        S.phase1_computation(sp, bytes(acc1), resp, False, True, S.CHECK_NO, *k1)
        want1 = O.phase1_computation(0, acc0, sp.get_length(False), False, False, 3, rp.g1_chunk_size, rp.other_chunk_size, 0, *k0)
        assert bytes(acc1[64:]) == want1[64:]
        want = O.phase1_computation(0, bytes(acc1), sp.get_length(True), False, True, 3, rp.g1_chunk_size, rp.other_chunk_size, 0, *k1)
        assert bytes(resp) == want
        newc = bytearray(sp.get_length(False))
        pairs = S.phase1_verification_vectors(sp, bytes(resp), True, newc, False, seed=bytes(range(32)))
        full = O.phase1_computation(0, bytes(acc1), sp.get_length(False), False, False, 3, rp.g1_chunk_size, rp.other_chunk_size, 0, *k1)
        assert bytes(newc[64:]) == full[64:]
        tau = k0[0] * k1[0] % cv.r
        for (s, sx), g in zip(pairs, (cv.g1, cv.g2, cv.g1, cv.g1)):
            Sp = g.decode(s, False)
            assert Sp is not None and g.mul(Sp, tau) == g.decode(sx, False)
        # tamper: the failing vector is still found when the bad pair straddles the two shards
        offs = rp.split_offsets(True)
        o, c, sz = offs[3]
        bad = bytearray(resp)
        bad[o + (c // 2) * sz:o + (c // 2 + 1) * sz] = cv.g1.encode(cv.g1.mul(cv.g1.gen, 99), True)
        pairs = S.phase1_verification_vectors(sp, bytes(bad), True, None, False, seed=bytes(range(32)))
        verdicts = [g.mul(g.decode(s, False), tau) == g.decode(sx, False) for (s, sx), g in zip(pairs, (cv.g1, cv.g2, cv.g1, cv.g1))]
        assert verdicts == [True, True, True, False]
    finally:
        ffi.init([0])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_prepare_phase2_over_two_devices():
    """Groth16Params::new spreads its four transforms over the selected devices; bytes must not depend on that."""
    cv, cid = R.BLS12_377, S.BLS12_377
    power = 6
    rp = R.Phase1Parameters(cv, power, 64)
    sp = S.Phase1Parameters(cid, power, 64)
    keys = [random.Random(5).randrange(2, cv.r) for _ in range(3)]
    acc = bytearray(sp.get_length(False))
    S.phase1_computation(sp, bytes(R.phase1_initialization(rp, False)), acc, False, False, S.CHECK_NO, *keys)
    one = S.groth16_params_new(sp, bytes(acc), False, 1 << power, True)
    try:
        ffi.init([0, 1])
        two = S.groth16_params_new(sp, bytes(acc), False, 1 << power, True)
    finally:
        ffi.init([0])
    assert one == two
    # against the oracle on the first coefficients vector
    s1c, s2c = cv.g1.size(True), cv.g2.size(True)
    offs = rp.split_offsets(False)
    tau_g1 = cv.g1.read_batch(bytes(acc[offs[0][0]:offs[0][0] + (1 << power) * cv.g1.size(False)]), False)
    want = cv.g1.write_batch(R.group_ifft_fast(cv.g1, tau_g1), True)
    assert one[2 * s1c + s2c:2 * s1c + s2c + (1 << power) * s1c] == want
