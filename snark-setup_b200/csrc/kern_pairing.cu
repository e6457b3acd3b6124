// Kernel instantiations of the reduced Tate pairing ratio check (pairing.cuh) for both curves.
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
