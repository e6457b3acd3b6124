// Kernel instantiations for Bw6G1 (one translation unit per group keeps nvcc compile times parallel).
#include "fft.cuh"
#include "msm.cuh"
#include "qap.cuh"

namespace ss {
const GroupOps& ops_bw6_g1() {
    static const GroupOps o = GroupLaunch<Bw6G1>::ops();
    return o;
}
const MsmOps& msm_ops_bw6_g1() {
    static const MsmOps o = MsmLaunch<Bw6G1>::ops();
    return o;
}
const FftOps& fft_ops_bw6_g1() {
    static const FftOps o = FftLaunch<Bw6G1>::ops();
    return o;
}
const QapOps& qap_ops_bw6_g1() {
    static const QapOps o = QapLaunch<Bw6G1>::ops();
    return o;
}
}  // namespace ss
