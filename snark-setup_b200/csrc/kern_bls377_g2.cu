// Kernel instantiations for Bls377G2 (one translation unit per group keeps nvcc compile times parallel).
#include "fft.cuh"
#include "msm.cuh"
#include "qap.cuh"

namespace ss {
const GroupOps& ops_bls377_g2() {
    static const GroupOps o = GroupLaunch<Bls377G2>::ops();
    return o;
}
const MsmOps& msm_ops_bls377_g2() {
    static const MsmOps o = MsmLaunch<Bls377G2>::ops();
    return o;
}
const FftOps& fft_ops_bls377_g2() {
    static const FftOps o = FftLaunch<Bls377G2>::ops();
    return o;
}
const QapOps& qap_ops_bls377_g2() {
    static const QapOps o = QapLaunch<Bls377G2>::ops();
    return o;
}
}  // namespace ss
