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
        ProfScope ps("k_same_ratio", curve == SS_CURVE_BLS12_377 ? "bls12_377" : "bw6_761", count, s);  // (no MNT pairing: pairing_ops() is null)
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

// check_power_ratios / check_power_ratios_g2 (phase1/src/helpers/accumulator.rs:56-91) for the four vectors of a
// response, given their (s, sx) `pairs` (ss_phase1_verification_vectors, or the ss_phase1_reduce_partial_pairs of the
// shards' partial blobs) and g1_check = (tau_g1[0], tau_g1[1]), g2_check = (tau_g2[0], tau_g2[1]), uncompressed
// (verification.rs:58-71): tau_g1 / alpha_g1 / beta_g1 powers against g2_check, tau_g2 powers against g1_check, the
// four check_same_ratio in one launch.  SS_ERR_INVALID_RATIO: index = 0 tau_g1, 1 tau_g2, 2 alpha_g1, 3 beta_g1.
int ss_phase1_check_ratio_pairs(int curve, const uint8_t* pairs, const uint8_t* g1_check, const uint8_t* g2_check) {
    const GroupOps* g1 = group_ops(curve, SS_G1);
    const GroupOps* g2 = group_ops(curve, SS_G2);
    if (!g1 || !g2) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "unknown curve");
    if (!pairs || !g1_check || !g2_check) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "null argument");
    const size_t u1 = g1->usize, u2 = g2->usize;
    std::vector<uint8_t> a(4 * 2 * u1), b(4 * 2 * u2);
    const uint8_t* pr = pairs;
    memcpy(a.data(), pr, 2 * u1);
    memcpy(b.data(), g2_check, 2 * u2);
    memcpy(a.data() + 2 * u1, g1_check, 2 * u1);
    memcpy(b.data() + 2 * u2, pr + 2 * u1, 2 * u2);
    memcpy(a.data() + 4 * u1, pr + 2 * u1 + 2 * u2, 2 * u1);
    memcpy(b.data() + 4 * u2, g2_check, 2 * u2);
    memcpy(a.data() + 6 * u1, pr + 4 * u1 + 2 * u2, 2 * u1);
    memcpy(b.data() + 6 * u2, g2_check, 2 * u2);
    int bad = -1;
    int rc = ss_check_same_ratio_batch(curve, a.data(), b.data(), 4, &bad);
    if (rc == SS_ERR_INVALID_RATIO) {
        static const char* names[4] = {"tau_g1", "tau_g2", "alpha_g1", "beta_g1"};
        return fail(SS_ERR_INVALID_RATIO, (uint64_t)bad, 0, 0, "Invalid ratio! Context: Power pairs: %s", names[bad & 3]);
    }
    return rc;
}

// The per-vector half of Phase1::verification with its verdict (phase1/src/verification.rs:44-80,217-411):
// ss_phase1_verification_vectors, then check_power_ratios / check_power_ratios_g2 of every vector
// (phase1/src/helpers/accumulator.rs:56-91) against g2_check = (tau_g2[0], tau_g2[1]) resp.
// g1_check = (tau_g1[0], tau_g1[1]) read from the response itself (verification.rs:58-71), the four
// check_same_ratio run as one launch.  index of the error = vector (0 tau_g1, 1 tau_g2, 2 alpha_g1, 3 beta_g1).
int ss_phase1_verification_ratios(const ss_phase1_params* p, const uint8_t* output, size_t output_len, int compressed_output,
                                  int check_output, uint8_t* new_challenge, size_t new_challenge_len,
                                  int compressed_new_challenge, int subgroup_mode, const uint8_t* rho_seed) {
    if (!p) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "null parameters");
    const GroupOps* g1 = group_ops(p->curve, SS_G1);
    const GroupOps* g2 = group_ops(p->curve, SS_G2);
    if (!g1 || !g2) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "unknown curve");
    const size_t u1 = g1->usize, u2 = g2->usize;
    std::vector<uint8_t> pairs(2 * (3 * u1 + u2));
    int rc = ss_phase1_verification_vectors(p, output, output_len, compressed_output, new_challenge, new_challenge_len,
                                            compressed_new_challenge, subgroup_mode, 1, rho_seed, pairs.data());
    if (rc) return rc;
    ss_phase1_sizes z;
    if ((rc = phase1_sizes(p, &z))) return rc;
    if (z.g1_chunk_size < 2 || z.other_chunk_size < 2) return fail(SS_ERR_BATCH_TOO_SMALL, 0, 0, 0, "ratio check needs two elements");
    uint64_t off[5];
    vector_offsets(p->curve, z, compressed_output, off);
    std::vector<uint8_t> g1_check(2 * u1), g2_check(2 * u2);
    // read_initial_elements (verification.rs:58-65)
    if ((rc = ss_transcode(p->curve, SS_G1, output + off[0], compressed_output, check_output, g1_check.data(), 0, 2))) return rc;
    if ((rc = ss_transcode(p->curve, SS_G2, output + off[1], compressed_output, check_output, g2_check.data(), 0, 2))) return rc;
    return ss_phase1_check_ratio_pairs(p->curve, pairs.data(), g1_check.data(), g2_check.data());
}

}  // extern "C"
