// Kernel instantiations of the reduced Tate pairing ratio check (pairing.cuh) for both curves.
// The pairing is ONE warp per check: pure latency, tiny per-lane state (one Fq^k coefficient).  Here the out-of-line
// Fq2 units with row-interleaved base products (fp2.cuh) pay off — three carry chains in flight from the one warp — while
// the throughput kernels of kern_bls377_g2.cu lose with them (profiles/r02_ab_variants.md).  Device code is per
// translation unit (no -rdc), so the setting is local to this file.
#ifndef SS_FP2_UNITS
#define SS_FP2_UNITS 1
#endif
#include "pairing.cuh"

namespace ss {
const PairingOps& pairing_ops_bls377() {
    static const PairingOps o = PairingLaunch<Bls377Pairing>::ops();
    return o;
}
const PairingOps& pairing_ops_bw6() {
    static const PairingOps o = PairingLaunch<Bw6Pairing>::ops();
    return o;
}
}  // namespace ss
