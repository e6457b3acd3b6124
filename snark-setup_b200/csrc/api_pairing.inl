// Included by api.cu: same_ratio / check_same_ratio (setup-utils/src/helpers.rs:406-424) on the device, and the
// ratio half of Phase1::verification that uses it (phase1/src/verification.rs:217-411; accumulator.rs:56-91).

namespace {

const PairingOps* pairing_ops(int curve) {
    if (curve == SS_CURVE_BLS12_377) return &pairing_ops_bls377();
    if (curve == SS_CURVE_BW6_761) return &pairing_ops_bw6();
    return nullptr;
}

// verdict[i]: bit 0 = same ratio, bit 1 = a point was the identity; < 0: -ERR_* from decoding
int same_ratio_batch(int curve, const uint8_t* g1_pairs, const uint8_t* g2_pairs, int count, int* verdict) {
    const PairingOps* po = pairing_ops(curve);
    const GroupOps* g1 = group_ops(curve, SS_G1);
    const GroupOps* g2 = group_ops(curve, SS_G2);
    if (!po || !g1 || !g2) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "unknown curve");
    if (count <= 0) return SS_OK;
    if (!g1_pairs || !g2_pairs || !verdict) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "null buffer");
    int rc = ensure_init();
    if (rc) return rc;
    const int device = g_devices[0];
    const size_t b1 = (size_t)2 * g1->usize * count, b2 = (size_t)2 * g2->usize * count;
    LaneGuard lg;
    if ((rc = lane_acquire(device, align_up(b1, 256) + align_up(b2, 256) + align_up((size_t)4 * count, 256), &lg.l))) return rc;
    cudaStream_t s = lg.l->stream;
    Carver cv(lg.l->buf);
    uint8_t* d1 = cv.take<uint8_t>(b1);
    uint8_t* d2 = cv.take<uint8_t>(b2);
    int* dv = cv.take<int>((size_t)4 * count);
    CU(cudaMemcpyAsync(d1, g1_pairs, b1, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d2, g2_pairs, b2, cudaMemcpyHostToDevice, s));
    {
        ProfScope ps("k_same_ratio", curve == SS_CURVE_BLS12_377 ? "bls12_377" : "bw6_761", count, s);
        po->same_ratio(reinterpret_cast<const uint32_t*>(d1), reinterpret_cast<const uint32_t*>(d2), count, dv, s);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(verdict, dv, (size_t)4 * count, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    prof_flush();
    for (int i = 0; i < count; i++)
        if (verdict[i] < 0) return fail(-verdict[i], i, 0, 0, "same_ratio: pair %d: undecodable point (error %d)", i, -verdict[i]);
    return SS_OK;
}

}  // namespace

extern "C" {

int ss_same_ratio(int curve, const uint8_t* g1_pair, const uint8_t* g2_pair, int* same) {
    if (!same) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "null output");
    int v = 0;
    int rc = same_ratio_batch(curve, g1_pair, g2_pair, 1, &v);
    if (rc) return rc;
    *same = v & 1;
    return SS_OK;
}

int ss_check_same_ratio(int curve, const uint8_t* g1_pair, const uint8_t* g2_pair) {
    int v = 0;
    int rc = same_ratio_batch(curve, g1_pair, g2_pair, 1, &v);
    if (rc) return rc;
    if (v & 2) return fail(SS_ERR_INVALID_RATIO, 0, 0, 0, "Invalid Ratio: zero");
    if (!(v & 1)) return fail(SS_ERR_INVALID_RATIO, 0, 0, 0, "Invalid Ratio: wrong pairing");
    return SS_OK;
}

int ss_check_same_ratio_batch(int curve, const uint8_t* g1_pairs, const uint8_t* g2_pairs, int count, int* first_bad) {
    if (count < 0 || count > 65535) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "bad count");
    std::vector<int> v((size_t)std::max(count, 1), 0);
    int rc = same_ratio_batch(curve, g1_pairs, g2_pairs, count, v.data());
    if (rc) return rc;
    for (int i = 0; i < count; i++) {
        if ((v[i] & 2) || !(v[i] & 1)) {
            if (first_bad) *first_bad = i;
            return fail(SS_ERR_INVALID_RATIO, i, 0, 0, (v[i] & 2) ? "Invalid Ratio: zero (check %d)" : "Invalid Ratio: wrong pairing (check %d)", i);
        }
    }
    return SS_OK;
}

}  // extern "C"
