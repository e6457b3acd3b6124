"""snark-setup_b200 — B200-native batch-exponentiation engine behind the setup-utils / phase1 API
of nimiq/snark-setup.

The product is the CUDA shared library `csrc/libsnarksetup_b200.so` (C ABI in
include/snark_setup_b200.h).  This package is only the thin ctypes binding used by the tests and
bench.py; it mirrors the reference's function names (setup-utils/src/helpers.rs,
phase1/src/helpers/buffers.rs, phase1/src/computation.rs) and raises the reference's error
variants.  There is no CPU fallback: if the library is missing or no CUDA device is usable, every
compute call raises.
"""
from .ffi import (  # noqa: F401
    BLS12_377,
    BW6_761,
    CHECK_FULL,
    CHECK_NO,
    CHECK_ONLY_IN_GROUP,
    CHECK_ONLY_NON_ZERO,
    G1,
    G2,
    DeviceError,
    IncorrectSubgroup,
    InvalidData,
    InvalidLength,
    InvalidRatio,
    Phase1Parameters,
    PointAtInfinity,
    SetupError,
    UnexpectedFlags,
    apply_powers,
    batch_exp,
    batch_mul,
    build,
    check_and_ratio,
    check_same_ratio,
    check_same_ratio_batch,
    same_ratio,
    check_subgroup,
    element_size,
    merge_pairs,
    phase1_verification_ratios,
    phase1_verification_vectors,
    phase1_verification_vectors_dev,
    power_pairs,
    qap_dot_product,
    generate_powers_of_tau,
    group_ifft,
    groth16_params_new,
    groth16_params_size,
    h_query_groth16,
    lib,
    lib_path,
    phase1_aggregate_chunk,
    phase1_computation,
    phase1_decompress,
    phase1_split_chunk,
    phase1_computation_dev,
    scalar_size,
    transcode,
)
