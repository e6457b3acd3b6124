// Bucket-MSM kernel instantiations for Mnt4G2 (see kern_mnt4_g2.cu).
#include "msm.cuh"

namespace ss {
const MsmOps& msm_ops_mnt4_g2() {
    static const MsmOps o = MsmLaunch<Mnt4G2>::ops();
    return o;
}
}  // namespace ss
