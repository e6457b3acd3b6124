// Kernel instantiations for Bls377G1 (one translation unit per group keeps nvcc compile times parallel).
#include "kernels.cuh"

namespace ss {
const GroupOps& ops_bls377_g1() {
    static const GroupOps o = GroupLaunch<Bls377G1>::ops();
    return o;
}
}  // namespace ss
