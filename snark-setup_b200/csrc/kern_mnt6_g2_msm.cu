// Bucket-MSM kernel instantiations for Mnt6G2 (see kern_mnt6_g2.cu).
#include "msm.cuh"

namespace ss {
const MsmOps& msm_ops_mnt6_g2() {
    static const MsmOps o = MsmLaunch<Mnt6G2>::ops();
    return o;
}
}  // namespace ss
