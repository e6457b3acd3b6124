"""ctypes binding of oracle/_build/liboracle.so (C++ restatement).  TEST INFRASTRUCTURE ONLY — see the
header of oracle/oracle.cpp.  Built by `make -C oracle` / __graft_entry__.build()."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        L = C.CDLL(_PATH)
        L.oracle_apply_powers.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_size_t,
                                          C.c_void_p, C.c_char_p, C.c_uint64, C.c_char_p, C.POINTER(C.c_uint64)]
        L.oracle_transcode.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_size_t,
                                       C.c_int, C.POINTER(C.c_uint64)]
        L.oracle_msm.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
        L.oracle_group_ifft.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_size_t,
                                        C.POINTER(C.c_uint64)]
        L.oracle_powers.argtypes = [C.c_int, C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.oracle_phase1_computation.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                                C.c_uint64, C.c_uint64, C.c_char_p, C.c_char_p, C.c_char_p]
        L.oracle_msm_pippenger.argtypes = L.oracle_msm.argtypes
        L.oracle_verify_vector.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_size_t, C.c_int,
                                           C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.oracle_phase1_verification_vectors.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_uint64,
                                                         C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
        L.oracle_fq_mul_ns.restype = C.c_double
        L.oracle_fq_mul_ns.argtypes = [C.c_int]
        _lib = L
    return _lib


FR_BYTES = {0: 32, 1: 48}
SIZES = {(0, 0): (96, 48), (0, 1): (192, 96), (1, 0): (192, 96), (1, 1): (192, 96)}


class OracleError(Exception):
    def __init__(self, code, index):
        super().__init__(f"oracle error {code} at element {index}")
        self.code, self.index = code, index


def _s(curve, v):
    if v is None:
        return None
    return int(v).to_bytes(FR_BYTES[curve], "little") if isinstance(v, int) else bytes(v)


def threads():
    return lib().oracle_threads()


def set_threads(n):
    lib().oracle_set_threads(n)


def apply_powers(curve, group, inp, in_c, check, out_c, n, powers=None, tau=None, first_power=0, coeff=None):
    osz = SIZES[(curve, group)][1 if out_c else 0]
    out = C.create_string_buffer(max(1, n * osz))
    bad = C.c_uint64(0)
    pw = None
    if powers is not None:
        pw = b"".join(_s(curve, p) for p in powers)
    rc = lib().oracle_apply_powers(curve, group, bytes(inp), int(in_c), check, out, int(out_c), n, pw, _s(curve, tau),
                                   first_power, _s(curve, coeff), C.byref(bad))
    if rc:
        raise OracleError(rc, bad.value)
    return out.raw[:n * osz]


def transcode(curve, group, inp, in_c, check, out_c, n, rmul_subgroup=False, want_output=True):
    osz = SIZES[(curve, group)][1 if out_c else 0]
    out = C.create_string_buffer(max(1, n * osz)) if want_output else None
    bad = C.c_uint64(0)
    rc = lib().oracle_transcode(curve, group, bytes(inp), int(in_c), check, out, int(out_c), n, int(rmul_subgroup),
                                C.byref(bad))
    if rc:
        raise OracleError(rc, bad.value)
    return out.raw[:n * osz] if want_output else None


def msm(curve, group, pts, compressed, n, scalars):
    out = C.create_string_buffer(SIZES[(curve, group)][0])
    rc = lib().oracle_msm(curve, group, bytes(pts), int(compressed), n, b"".join(_s(curve, s) for s in scalars), out)
    if rc:
        raise OracleError(rc, 0)
    return out.raw


def group_ifft(curve, group, inp, in_c, out_c, check=3):
    """to_coeffs (setup-utils/src/groth16_utils.rs:44-53) by the C++ restatement (iterative decimation in frequency)."""
    isz = SIZES[(curve, group)][1 if in_c else 0]
    osz = SIZES[(curve, group)][1 if out_c else 0]
    n = len(inp) // isz
    out = C.create_string_buffer(max(1, n * osz))
    bad = C.c_uint64(0)
    rc = lib().oracle_group_ifft(curve, group, bytes(inp), int(in_c), check, out, int(out_c), n, C.byref(bad))
    if rc:
        raise OracleError(rc, bad.value)
    return out.raw[:n * osz]


def powers(curve, tau, start, end):
    fb = FR_BYTES[curve]
    out = C.create_string_buffer(max(1, (end - start) * fb))
    lib().oracle_powers(curve, _s(curve, tau), start, end, out)
    return [int.from_bytes(out.raw[i * fb:(i + 1) * fb], "little") for i in range(end - start)]


def phase1_computation(curve, inp, out_len, cin, cout, check, n_g1, n_other, first_power, tau, alpha, beta):
    out = C.create_string_buffer(out_len)
    rc = lib().oracle_phase1_computation(curve, bytes(inp), out, int(cin), int(cout), check, n_g1, n_other, first_power,
                                         _s(curve, tau), _s(curve, alpha), _s(curve, beta))
    if rc:
        raise OracleError(rc, 0)
    return out.raw


def msm_pippenger(curve, group, pts, compressed, n, scalars):
    """VariableBaseMSM::msm_bigint restated (signed-digit bucket method, ark-ec 0.4.2) with explicit scalars."""
    out = C.create_string_buffer(SIZES[(curve, group)][0])
    rc = lib().oracle_msm_pippenger(curve, group, bytes(pts), int(compressed), n, b"".join(_s(curve, s) for s in scalars), out)
    if rc:
        raise OracleError(rc, 0)
    return out.raw


def phase1_verification_vectors(curve, resp, cin, out_len, cout, n_g1, n_other, seed=1, decode_passes=2, want_output=True):
    """Per-vector loop of Phase1::verification as the reference runs it (decode twice, r*P subgroup test, power_pairs
    with full-width random scalars, uncompressed re-emit).  Returns (new_challenge bytes | None, [(s, sx)] x 4)."""
    out = C.create_string_buffer(out_len) if want_output else None
    u1, u2 = SIZES[(curve, 0)][0], SIZES[(curve, 1)][0]
    pairs = C.create_string_buffer(2 * (3 * u1 + u2))
    bad = C.c_uint64(0)
    rc = lib().oracle_phase1_verification_vectors(curve, bytes(resp), int(cin), out, int(cout), n_g1, n_other, seed,
                                                  decode_passes, pairs, C.byref(bad))
    if rc:
        raise OracleError(rc, bad.value)
    res, o = [], 0
    for usz in (u1, u2, u1, u1):
        res.append((pairs.raw[o:o + usz], pairs.raw[o + usz:o + 2 * usz]))
        o += 2 * usz
    return (out.raw if want_output else None), res


def fq_mul_ns(iters=2000000):
    """ns per BLS12-377 Fq Montgomery multiplication on one core (dependent chain)."""
    return lib().oracle_fq_mul_ns(iters)


def has_asm_mul():
    return bool(lib().oracle_has_asm_mul())


def force_portable_mul(on=True):
    lib().oracle_force_portable_mul(int(on))
