// Included by api.cu (shares its anonymous-namespace helpers): the verification side of the path.
//
//   check_elements_are_nonzero_and_in_prime_order_subgroup   phase1/src/helpers/accumulator.rs:95-145
//   check_power_ratios / check_power_ratios_g2               phase1/src/helpers/accumulator.rs:56-91
//   merge_pairs / power_pairs                                setup-utils/src/helpers.rs:371-390
//   per-vector loop of Phase1::verification                  phase1/src/verification.rs:243-411
//   aggregate_verification                                   phase1/src/verification.rs:505-769
//   H/L ratio checks of MPCParameters::verify                phase2/src/parameters.rs:393-407
//
// The reference decodes every window twice, checks the subgroup, runs two MSMs and re-encodes, window
// by window, with 2 pairings per window.  Here a vector is decoded ONCE per tile and the same affine
// scratch feeds the subgroup kernel, the bucket MSM (whose accumulators persist across tiles, so one
// (s, sx) pair comes out per vector) and the re-encode.  `check_same_ratio` (2 pairings per vector)
// stays with the caller (helpers.rs:410-424).

namespace {

const MsmOps* msm_ops(int curve, int group) {
    if (curve == SS_CURVE_BLS12_377) return group == SS_G1 ? &msm_ops_bls377_g1() : group == SS_G2 ? &msm_ops_bls377_g2() : nullptr;
    if (curve == SS_CURVE_BW6_761) return group == SS_G1 ? &msm_ops_bw6_g1() : group == SS_G2 ? &msm_ops_bw6_g2() : nullptr;
    if (curve == SS_CURVE_MNT4_753) return group == SS_G1 ? &msm_ops_mnt4_g1() : group == SS_G2 ? &msm_ops_mnt4_g2() : nullptr;
    if (curve == SS_CURVE_MNT6_753) return group == SS_G1 ? &msm_ops_mnt6_g1() : group == SS_G2 ? &msm_ops_mnt6_g2() : nullptr;
    return nullptr;
}

struct RatioJob {
    int curve, group;
    const uint8_t* v1;       // n elements (host, or device when host == false)
    const uint8_t* v2;       // second vector for merge_pairs, or nullptr for power_pairs on v1
    int compressed;
    int check;               // CheckForCorrectness used when decoding
    uint64_t n;              // elements in v1
    int subgroup;            // 1: p.mul_bigint(r).is_zero() for every element of v1
    int do_ratio;            // 1: produce (s, sx)
    const uint8_t* rho;      // explicit scalars (HOST memory, one per pair) or nullptr
    const uint8_t* seed;     // 32-byte ChaCha20 key (HOST memory) when rho == nullptr
    uint8_t* out;            // re-encoded v1 (same memory kind as v1) or nullptr
    int out_compressed;
    uint8_t* out_s;          // HOST: uncompressed s, sx
    uint8_t* out_sx;
    const char* what;
    uint64_t rho_base = 0;   // global index of element 0 (ChaCha20 counter base) when the vector is a shard
    uint64_t own = 0;        // elements to subgroup-check / re-emit (0 = all n); n - own = 1 overlap element of a shard
    int prio = 0;            // 1: run on a high-priority lane (lane_acquire)
};

bool priority_lanes() {  // $SS_PRIORITY_LANES=0 switches the high-priority G2 lane off (A/B)
    static const bool on = [] {
        const char* e = getenv("SS_PRIORITY_LANES");
        return !(e && atoi(e) == 0);
    }();
    return on;
}

int pick_window_bits(uint64_t pairs_per_tile) {
    // c = log2(pairs) - offset: 2^offset points per bucket on average.  Measured (profiles/r02_ab_variants.md): offset 4
    // with c <= 16 beats round 1's offset 5 / c <= 15 (2^20 pairs: c = 16, W = 8 windows instead of 9 and twice the
    // bucket threads; accumulate 64.7 -> 54.0 ms per 2^20-power response, the reductions grow by 2 ms).
    // $SS_MSM_C_OFFSET / $SS_MSM_C_MAX override for A/B runs.
    static const int offset = [] { const char* e = getenv("SS_MSM_C_OFFSET"); return e ? atoi(e) : 4; }();
    static const int cmax = [] { const char* e = getenv("SS_MSM_C_MAX"); int v = e ? atoi(e) : 16; return v < 2 ? 2 : (v > 16 ? 16 : v); }();
    int lg = 0;
    while ((1ull << (lg + 1)) <= pairs_per_tile) lg++;
    int c = lg - offset;
    if (c < 2) c = 2;
    if (c > cmax) c = cmax;
    return c;
}

size_t ratio_tile_elems() {
    static size_t t = [] {
        const char* e = getenv("SS_RATIO_TILE_LOG2");
        int l = e ? atoi(e) : 21;  // 2^21: -7 ms per 2^22 verify round against 2^20 (profiles/r02_ab_variants.md)
        if (l < 8) l = 8;
        if (l > 24) l = 24;
        return (size_t)1 << l;
    }();
    return t;
}

int run_ratio_vector(int device, const RatioJob& j, bool host, cudaStream_t user_stream) {
    const GroupOps* op = group_ops(j.curve, j.group);
    const MsmOps* mo = msm_ops(j.curve, j.group);
    if (!op || !mo) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "unknown curve/group %d/%d", j.curve, j.group);
    const GroupOps& o = *op;
    if (j.n == 0) return SS_OK;
    const bool two = j.v2 != nullptr;
    const uint64_t pairs = j.do_ratio ? (two ? j.n : j.n - 1) : 0;
    if (j.do_ratio && pairs == 0) return fail(SS_ERR_BATCH_TOO_SMALL, 0, 0, 0, "%s: ratio check needs at least 2 elements", j.what);
    const size_t isz = j.compressed ? o.csize : o.usize, osz = j.out_compressed ? o.csize : o.usize;
    // ratio jobs use larger tiles than batch_exp: W*B bucket threads per tile must be several waves of the
    // machine (2^21 pairs -> c = 16 -> 8 x 65536 buckets, 16 points per bucket on average), see
    // profiles/r02_ab_variants.md
    const uint64_t own_total = j.own ? j.own : j.n;
    const size_t T = std::min<uint64_t>(j.do_ratio ? ratio_tile_elems() : tile_elems(), own_total);
    const size_t ntiles = (own_total + T - 1) / T;
    // Window width c and scalar width.  Every window must be FULL: a top window of b < c bits has only
    // 2^b - 1 buckets holding n / 2^b points each, and one thread per bucket then serialises the whole
    // tile (measured: 3.2 s instead of 9 ms at n = 2^19, c = 14, 128-bit rho).  Generated rho therefore
    // gets W*c >= 128 bits (ChaCha20 supplies 512); for caller-supplied full-width scalars c is lowered
    // until it (almost) divides the field size (253 = 23 * 11, 377 = 29 * 13).
    int c = pick_window_bits(std::min<uint64_t>(T, std::max<uint64_t>(pairs, 1)));
    int nbits;
    if (!j.rho) {
        nbits = ((128 + c - 1) / c) * c;
    } else {
        nbits = o.fr_bits;
        while (c > 2 && !(nbits % c == 0 || nbits % c >= c - 1)) c--;
    }
    const int W = (nbits + c - 1) / c;
    const uint32_t B = 1u << c;
    // bucket segments of the running-sum reduction: one thread per (sum, window, segment), and the kernel is pure
    // latency (2 * seglen + ~20 additions in sequence per thread), so segments shrink while all the threads still fit
    // the machine at once (2 blocks of 128 threads per SM at ~250 registers)
    uint32_t seglen = B >= 1024 ? 32 : (B >= 256 ? B / 256 : 1);
    while (seglen > 8 && (uint64_t)2 * W * (B / (seglen / 2)) <= 148u * 2 * 128) seglen /= 2;
    const uint32_t nseg = B / seglen;
    const size_t fw = o.coord_words;

    // slab layout
    size_t need = 256 + align_up(ntiles * 16, 256);
    const size_t aff_b = align_up((size_t)2 * fw * 4 * (T + 1), 256), inf_b = align_up(T + 1, 256);
    need += (two ? 2 : 1) * (aff_b + inf_b);
    if (host) need += (two ? 2 : 1) * align_up(isz * (T + 1), 256) + (j.out ? align_up(osz * T, 256) : 0);
    size_t sort_b = 0, bucket_b = 0;
    if (j.do_ratio) {
        sort_b = 4 * align_up((size_t)W * B * 4, 256) + 1024 + align_up((size_t)W * T * 4, 256) + (j.rho ? align_up(scalar_stride(o) * T, 256) : 0);
        bucket_b = align_up((size_t)3 * fw * 4 * 2 * W * B, 256) + align_up((size_t)3 * fw * 4 * 2 * W * nseg, 256) +
                   align_up((size_t)3 * fw * 4 * 2 * W, 256) + 2 * align_up(o.usize, 256);
    }
    need += sort_b + bucket_b;
    LaneGuard lg;
    std::vector<std::vector<uint8_t>> keep;  // re-packed scalar staging, alive until the stream is synchronised
    int rc = lane_acquire(device, need, &lg.l, j.prio);
    if (rc) return rc;
    cudaStream_t s = (!host && user_stream) ? user_stream : lg.l->stream;
    Carver cv(lg.l->buf);
    unsigned long long* d_status = cv.take<unsigned long long>(ntiles * 16);
    uint32_t* aff1 = cv.take<uint32_t>((size_t)2 * fw * 4 * (T + 1));
    uint8_t* inf1 = cv.take<uint8_t>(T + 1);
    uint32_t* aff2 = nullptr;
    uint8_t* inf2 = nullptr;
    if (two) {
        aff2 = cv.take<uint32_t>((size_t)2 * fw * 4 * (T + 1));
        inf2 = cv.take<uint8_t>(T + 1);
    }
    uint8_t *bi1 = nullptr, *bi2 = nullptr, *bo = nullptr;
    if (host) {
        bi1 = cv.take<uint8_t>(isz * (T + 1));
        if (two) bi2 = cv.take<uint8_t>(isz * (T + 1));
        if (j.out) bo = cv.take<uint8_t>(osz * T);
    }
    uint32_t *hist = nullptr, *cursor = nullptr, *counts = nullptr, *order = nullptr, *bins = nullptr, *idx = nullptr, *buckets = nullptr, *segres = nullptr,
             *winres = nullptr, *d_s = nullptr, *d_sx = nullptr;
    uint8_t* d_rho = nullptr;
    if (j.do_ratio) {
        hist = cv.take<uint32_t>((size_t)W * B * 4);
        cursor = cv.take<uint32_t>((size_t)W * B * 4);
        counts = cv.take<uint32_t>((size_t)W * B * 4);
        order = cv.take<uint32_t>((size_t)W * B * 4);
        bins = cv.take<uint32_t>(1024);
        idx = cv.take<uint32_t>((size_t)W * T * 4);
        if (j.rho) d_rho = cv.take<uint8_t>(scalar_stride(o) * T);
        buckets = cv.take<uint32_t>((size_t)3 * fw * 4 * 2 * W * B);
        segres = cv.take<uint32_t>((size_t)3 * fw * 4 * 2 * W * nseg);
        winres = cv.take<uint32_t>((size_t)3 * fw * 4 * 2 * W);
        d_s = cv.take<uint32_t>(o.usize);
        d_sx = cv.take<uint32_t>(o.usize);
        ProfScope ps("k_jac_fill_identity", o.name, (uint64_t)2 * W * B, s);
        mo->fill_identity(buckets, (uint64_t)2 * W * B, s);
    }
    CU(cudaMemsetAsync(d_status, 0xff, ntiles * 16, s));

    for (size_t t = 0; t < ntiles; t++) {
        const uint64_t e0 = t * T;
        const uint64_t own = std::min<uint64_t>(T, own_total - e0);            // elements this tile owns
        const uint64_t ne = two ? own : std::min<uint64_t>(T + 1, j.n - e0);   // decoded (one overlap for power_pairs)
        const uint64_t np = j.do_ratio ? (two ? own : (e0 < pairs ? std::min<uint64_t>(T, pairs - e0) : 0)) : 0;
        const uint64_t stride = ne;
        const uint8_t* in1 = j.v1 + e0 * isz;
        if (host) {
            CU(cudaMemcpyAsync(bi1, in1, ne * isz, cudaMemcpyHostToDevice, s));
            in1 = bi1;
        }
        {
            DecodeArgs da = {reinterpret_cast<const uint32_t*>(in1), j.compressed, j.check, ne, aff1, inf1, d_status + 2 * t};
            ProfScope ps("k_decode", o.name, ne, s);
            o.decode(da, s);
        }
        if (two) {
            const uint8_t* in2 = j.v2 + e0 * isz;
            if (host) {
                CU(cudaMemcpyAsync(bi2, in2, ne * isz, cudaMemcpyHostToDevice, s));
                in2 = bi2;
            }
            DecodeArgs da = {reinterpret_cast<const uint32_t*>(in2), j.compressed, j.check, ne, aff2, inf2, d_status + 2 * t};
            ProfScope ps("k_decode", o.name, ne, s);
            o.decode(da, s);
        }
        if (j.subgroup) {
            // only the first `own` elements: the overlap element belongs to the next tile.  The SoA
            // stride is `ne`, so pass n = own through a strided view: k_subgroup indexes [i] with stride n,
            // hence it is launched over `ne` with the count check inside the kernel args.
            SubgroupArgs sa = {aff1, inf1, stride, d_status + 2 * t + 1, own, j.compressed ? 0 : 1};
            ProfScope ps("k_subgroup", o.name, own, s);
            o.subgroup(sa, s);
        }
        if (np) {
            MsmSortArgs sa;
            sa.rho.explicit_rho = nullptr;
            if (j.rho) {
                const uint8_t* src = repack_scalars(o, j.rho + e0 * o.fr_bytes, np, keep);
                CU(cudaMemcpyAsync(d_rho, src, np * scalar_stride(o), cudaMemcpyHostToDevice, s));
                sa.rho.explicit_rho = reinterpret_cast<const uint32_t*>(d_rho);
            } else {
                memcpy(sa.rho.key, j.seed, 32);
            }
            sa.rho.first_index = j.rho_base + e0;
            sa.rho.frw = o.fr_words;
            sa.rho.nbits = nbits;
            sa.n = np;
            sa.c = c;
            sa.W = W;
            sa.hist = hist;
            sa.cursor = cursor;
            sa.idx = idx;
            {
                ProfScope ps("k_msm_sort", o.name, np, s);
                msm_sort(sa, counts, s);
                msm_order(counts, (uint64_t)W << c, bins, order, s);
            }
            MsmAccArgs aa;
            aa.aff1 = aff1;
            aa.inf1 = inf1;
            aa.aff2 = two ? aff2 : aff1 + 2 * fw;  // power_pairs: v2_i = v1_{i+1} (next AoS element)
            aa.inf2 = two ? inf2 : inf1 + 1;
            aa.stride = stride;
            aa.n = np;
            aa.c = c;
            aa.W = W;
            aa.offsets = hist;
            aa.counts = counts;
            aa.idx = idx;
            aa.order = order;
            aa.buckets = buckets;
            ProfScope ps("k_msm_accumulate", o.name, np, s);
            mo->accumulate(aa, s);
        }
        if (j.out) {
            uint8_t* dst = host ? bo : j.out + e0 * osz;
            EncodeArgs ea = {aff1, inf1, stride, reinterpret_cast<uint32_t*>(dst), j.out_compressed, own};
            {
                ProfScope ps("k_encode", o.name, own, s);
                o.encode(ea, s);
            }
            if (host) CU(cudaMemcpyAsync(j.out + e0 * osz, bo, own * osz, cudaMemcpyDeviceToHost, s));
        }
    }
    if (j.do_ratio) {
        MsmReduceArgs ra = {buckets, c, W, seglen, nseg, segres, winres, d_s, d_sx};
        {
            ProfScope ps("k_msm_reduce", o.name, (uint64_t)2 * W * B, s);
            mo->reduce(ra, s);
        }
        CU(cudaMemcpyAsync(j.out_s, d_s, o.usize, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(j.out_sx, d_sx, o.usize, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaGetLastError());
    std::vector<unsigned long long> st(ntiles * 2);
    CU(cudaMemcpyAsync(st.data(), d_status, ntiles * 16, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    prof_flush();
    for (size_t t = 0; t < ntiles; t++) {
        if ((rc = decode_status(st[2 * t], t * T, j.what))) return rc;
        if ((rc = decode_status(st[2 * t + 1], t * T, j.what))) return rc;
    }
    return SS_OK;
}

// pairs <- element-wise sums of `count` blobs laid out like the `pairs` output of ss_phase1_verification_vectors
// (4 x (s || sx), uncompressed: G1, G2, G1, G1).  Identity slots (0x40 flag) are neutral.  Eight tiny kernels on
// one stream, one synchronisation.
int sum_pair_blobs(int curve, const uint8_t* const* blobs, int count, uint8_t* pairs) {
    const GroupOps* g1 = group_ops(curve, SS_G1);
    const GroupOps* g2 = group_ops(curve, SS_G2);
    if (!g1 || !g2) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "unknown curve %d", curve);
    if (count <= 0 || !blobs || !pairs) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "bad argument");
    int rc = ensure_init();
    if (rc) return rc;
    const GroupOps* gs[4] = {g1, g2, g1, g1};
    const size_t blob_b = 2 * (3 * (size_t)g1->usize + g2->usize);
    const int dev0 = g_devices[0];
    CU(cudaSetDevice(dev0));
    LaneGuard lg;
    const size_t in_b = align_up((size_t)count * blob_b, 256);
    if ((rc = lane_acquire(dev0, 256 + in_b + align_up(blob_b, 256), &lg.l))) return rc;
    cudaStream_t s = lg.l->stream;
    unsigned long long* d_status = reinterpret_cast<unsigned long long*>(lg.l->buf);
    uint8_t* d_in = lg.l->buf + 256;
    uint8_t* d_out = d_in + in_b;
    // stage: for every (vector, half) the `count` points contiguously
    std::vector<uint8_t> stage((size_t)count * blob_b);
    size_t off = 0, slot_off[8], src_off = 0;
    for (int v = 0; v < 4; v++)
        for (int half = 0; half < 2; half++) {
            const size_t usz = gs[v]->usize;
            slot_off[2 * v + half] = off;
            for (int k = 0; k < count; k++) memcpy(stage.data() + off + k * usz, blobs[k] + src_off, usz);
            off += (size_t)count * usz;
            src_off += usz;
        }
    CU(cudaMemsetAsync(d_status, 0xff, 8, s));
    CU(cudaMemcpyAsync(d_in, stage.data(), stage.size(), cudaMemcpyHostToDevice, s));
    src_off = 0;
    for (int v = 0; v < 4; v++)
        for (int half = 0; half < 2; half++) {
            gs[v]->sum_points(reinterpret_cast<const uint32_t*>(d_in + slot_off[2 * v + half]), count,
                              reinterpret_cast<uint32_t*>(d_out + src_off), d_status, s);
            src_off += gs[v]->usize;
        }
    CU(cudaGetLastError());
    unsigned long long st;
    CU(cudaMemcpyAsync(pairs, d_out, blob_b, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(&st, d_status, 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return decode_status(st, 0, "sum of partial ratio points");
}

int ratio_common_checks(int curve, int group, const void* v, size_t n, int check) {
    if (!group_ops(curve, group)) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "unknown curve/group %d/%d", curve, group);
    if (n && !v) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "null input");
    if (check < 0 || check > 3) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "bad check mode %d", check);
    return SS_OK;
}

}  // namespace

extern "C" {

int ss_merge_pairs(int curve, int group, const uint8_t* v1, const uint8_t* v2, int compressed, int check, size_t n,
                   const uint8_t* rho, const uint8_t* rho_seed, uint8_t* out_s, uint8_t* out_sx) {
    int rc = ratio_common_checks(curve, group, v1, n, check);
    if (rc) return rc;
    if (!v2 || !out_s || !out_sx || (!rho && !rho_seed)) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "null argument");
    if (n == 0) return fail(SS_ERR_BATCH_TOO_SMALL, 0, 0, 0, "merge_pairs of empty vectors");
    if ((rc = ensure_init())) return rc;
    RatioJob j = {curve, group, v1, v2, compressed, check, n, 0, 1, rho, rho_seed, nullptr, 0, out_s, out_sx, "merge_pairs"};
    return run_ratio_vector(g_devices[0], j, true, nullptr);
}

int ss_power_pairs(int curve, int group, const uint8_t* v, int compressed, int check, size_t n, const uint8_t* rho,
                   const uint8_t* rho_seed, uint8_t* out_s, uint8_t* out_sx) {
    int rc = ratio_common_checks(curve, group, v, n, check);
    if (rc) return rc;
    if (!out_s || !out_sx || (!rho && !rho_seed)) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "null argument");
    if (n < 2) return fail(SS_ERR_BATCH_TOO_SMALL, 0, 0, 0, "power_pairs needs at least 2 elements");
    if ((rc = ensure_init())) return rc;
    RatioJob j = {curve, group, v, nullptr, compressed, check, n, 0, 1, rho, rho_seed, nullptr, 0, out_s, out_sx, "power_pairs"};
    return run_ratio_vector(g_devices[0], j, true, nullptr);
}

int ss_check_and_ratio(int curve, int group, const uint8_t* in, int in_compressed, size_t n, int subgroup_mode,
                       int do_ratio, const uint8_t* rho, const uint8_t* rho_seed, uint8_t* out, int out_compressed,
                       uint8_t* out_s, uint8_t* out_sx) {
    int rc = ratio_common_checks(curve, group, in, n, SS_CHECK_ONLY_NON_ZERO);
    if (rc) return rc;
    if (subgroup_mode < 0 || subgroup_mode > 3) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "bad subgroup mode");
    if (do_ratio && (!out_s || !out_sx || (!rho && !rho_seed))) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "null argument");
    if (do_ratio && n < 2) return fail(SS_ERR_BATCH_TOO_SMALL, 0, 0, 0, "ratio check needs at least 2 elements");
    if (n == 0) return SS_OK;
    if ((rc = ensure_init())) return rc;
    RatioJob j = {curve, group, in, nullptr, in_compressed, SS_CHECK_ONLY_NON_ZERO, n, subgroup_mode != SS_SUBGROUP_NO,
                  do_ratio, rho, rho_seed, out, out_compressed, out_s, out_sx, "check_and_ratio"};
    return run_ratio_vector(g_devices[0], j, true, nullptr);
}

// Per-vector hot loop of Phase1::verification (phase1/src/verification.rs:217-411, Groth16) over the
// whole response: for tau_g1, tau_g2, alpha_g1, beta_g1 -> nonzero + subgroup check, (s, sx) for the
// caller's check_same_ratio, and the re-encoded vector written into new_challenge; beta_g2 is re-encoded
// too (verification.rs:199-201).  `pairs` receives 4 x (s || sx) uncompressed in that vector order
// (G1: 2*g1_usize, G2: 2*g2_usize bytes each).  PoK / generator / before-after checks on the first
// elements (verification.rs:83-213) are O(1) pairing work and stay with the caller.
static int phase1_verification_impl(const ss_phase1_params* p, const uint8_t* output, size_t output_len,
                                    int compressed_output, uint8_t* new_challenge, size_t new_challenge_len,
                                    int compressed_new_challenge, int subgroup_mode, int ratio_check,
                                    const uint8_t* rho_seed, uint8_t* pairs, bool host, cudaStream_t stream,
                                    uint32_t shard_index = 0, uint32_t shard_count = 1) {
    if (!p || !output) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "null argument");
    if (shard_count == 0 || shard_index >= shard_count)
        return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "shard %u of %u", shard_index, shard_count);
    if (p->proving_system != SS_GROTH16 && p->proving_system != SS_MARLIN)
        return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "unknown proving system %d", p->proving_system);
    if (ratio_check && (!rho_seed || !pairs)) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "ratio check needs rho_seed and pairs");
    ss_phase1_sizes z;
    int rc = phase1_sizes(p, &z);
    if (rc) return rc;
    const GroupOps& g1 = *group_ops(p->curve, SS_G1);
    const GroupOps& g2 = *group_ops(p->curve, SS_G2);
    const uint64_t need_out = compressed_output ? z.contribution_size - z.public_key_size : z.accumulator_size;
    const uint64_t need_nc = compressed_new_challenge ? z.contribution_size - z.public_key_size : z.accumulator_size;
    if (output_len < need_out) return fail(SS_ERR_INVALID_LENGTH, 0, need_out, output_len, "response buffer too short");
    if (new_challenge && new_challenge_len < need_nc) return fail(SS_ERR_INVALID_LENGTH, 0, need_nc, new_challenge_len, "new_challenge buffer too short");
    if ((rc = ensure_init())) return rc;
    auto sz = [&](const GroupOps& g, int c) { return (uint64_t)(c ? g.csize : g.usize); };
    const uint64_t n1 = z.g1_chunk_size, n2 = z.other_chunk_size;
    // Marlin (verification.rs:413-483): tau_g1 over the chunk, plus k+2 tau_g2 and 3+3k alpha_g1 elements on chunk 0;
    // only tau_g1 is a power sequence (its ratio is what aggregate_verification checks, :649-672)
    const bool marlin = p->proving_system == SS_MARLIN;
    const bool chunk0 = is_chunk0(p);
    const uint64_t mk = p->total_size_in_log2;
    const uint64_t cnt[5] = {n1, marlin ? (chunk0 ? mk + 2 : 0) : n2, marlin ? (chunk0 ? 3 + 3 * mk : 0) : n2, marlin ? 0 : n2,
                             marlin ? 0u : 1u};
    const bool has_ratio[5] = {true, !marlin, !marlin, !marlin, false};
    const GroupOps* gs[5] = {&g1, &g2, &g1, &g1, &g2};
    const int grp[5] = {SS_G1, SS_G2, SS_G1, SS_G1, SS_G2};
    const char* names[5] = {"tau_g1", "tau_g2", "alpha_g1", "beta_g1", "beta_g2"};
    uint64_t oa[5], ob[5], op[5];
    {
        uint64_t a = 64, b = 64, po = 0;
        for (int v = 0; v < 5; v++) {
            oa[v] = a;
            ob[v] = b;
            op[v] = po;
            po += 2 * (uint64_t)gs[v]->usize;
            a += cnt[v] * sz(*gs[v], compressed_output);
            b += cnt[v] * sz(*gs[v], compressed_new_challenge);
        }
    }
    for (int v = 0; v < 4; v++)
        // a vector of one element cannot be ratio-checked (verification.rs:238-241 -> BatchTooSmall)
        if (ratio_check && has_ratio[v] && cnt[v] == 1) return fail(SS_ERR_BATCH_TOO_SMALL, 0, 0, 0, "%s: batch too small", names[v]);

    // One worker per device: every vector is cut into shard_count * D contiguous parts, this call owns parts
    // [shard_index * D, (shard_index + 1) * D), one per device; a part reads one element of overlap for the ratio
    // pairs that continue into the next part (helpers.rs:388-390) and produces partial (s, sx); rho_i is indexed
    // by the GLOBAL element index so the parts of one vector use disjoint ChaCha20 blocks.  The <= D partial
    // points per vector are added on device 0 (k_sum_points); no other inter-device traffic.  With shard_count > 1
    // `pairs` receives this shard's PARTIAL sums (ss_phase1_reduce_partial_pairs adds the shards' blobs).
    const int D = host ? (int)g_devices.size() : 1;
    const uint64_t parts = (uint64_t)shard_count * D;
    std::vector<std::vector<uint8_t>> partial(D);
    // every partial slot starts as the identity encoding (all-zero bytes would decode as the finite point (0, 0))
    auto identity_pairs = [&](std::vector<uint8_t>& blob) {
        blob.assign(2 * (3 * (size_t)g1.usize + g2.usize), 0);
        for (int v = 0; v < 4; v++) {
            blob[op[v] + gs[v]->usize - 1] = 0x40;
            blob[op[v] + 2 * gs[v]->usize - 1] = 0x40;
        }
    };
    for (int di = 0; di < D; di++) identity_pairs(partial[di]);
    auto worker = [&](int di, ss_error_info* err) -> int {
        const int device = host ? g_devices[di] : g_devices[0];
        auto one_vector = [&](int v) -> int {
            if (!cnt[v]) return SS_OK;
            cudaStream_t st = concurrent_vectors() ? nullptr : stream;
            uint8_t* out = new_challenge ? new_challenge + ob[v] : nullptr;
            int r;
            if (v == 4) {
                if (di != 0 || shard_index != 0) return SS_OK;
                // beta_g2: ALWAYS read with check_output_for_correctness (Full by default, verification.rs:199-201),
                // re-emitted only when there is a new challenge
                RatioJob j = {p->curve, grp[v], output + oa[v], nullptr, compressed_output, SS_CHECK_FULL, 1, 0, 0, nullptr,
                              nullptr, out, compressed_new_challenge, nullptr, nullptr, names[v]};
                return run_ratio_vector(device, j, host, st);
            }
            // own elements [s0, e0); read one more when pairs continue into the next shard
            uint64_t s0, e0;
            part_range(cnt[v], (uint64_t)shard_index * D + di, parts, &s0, &e0);
            if (e0 == s0) return SS_OK;  // empty part: its partial stays the identity
            const bool last = e0 == cnt[v];
            const bool want_ratio = ratio_check && has_ratio[v];
            const uint64_t nread = (e0 - s0) + ((want_ratio && !last) ? 1 : 0);
            const bool do_ratio = want_ratio && nread >= 2;
            uint8_t* ps = partial[di].data() + op[v];
            RatioJob j = {p->curve, grp[v], output + oa[v] + s0 * sz(*gs[v], compressed_output), nullptr, compressed_output,
                          SS_CHECK_ONLY_NON_ZERO, nread, subgroup_mode != SS_SUBGROUP_NO, do_ratio, nullptr, rho_seed,
                          out ? out + s0 * sz(*gs[v], compressed_new_challenge) : nullptr, compressed_new_challenge,
                          ps, ps + gs[v]->usize, names[v], s0};
            j.own = e0 - s0;
            j.prio = (grp[v] == SS_G2 && want_ratio && priority_lanes()) ? 1 : 0;
            if ((r = run_ratio_vector(device, j, host, st))) {
                g_err.index += s0;
                return r;
            }
            return SS_OK;  // a part without a pair (one-element tail) leaves the identity in its slot
        };
        auto run = [&]() -> int {
            if (!concurrent_vectors()) {
                for (int v = 0; v < 5; v++) {
                    int r = one_vector(v);
                    if (r) return r;
                }
                return SS_OK;
            }
            // vectors run concurrently (one thread + lane each): the latency-bound MSM reductions of one vector
            // overlap the subgroup / bucket kernels of the others
            CU(cudaSetDevice(device));
            if (stream) CU(cudaStreamSynchronize(stream));
            int rv[5] = {0, 0, 0, 0, 0};
            ss_error_info ev[5];
            std::vector<std::thread> vt;
            for (int v = 0; v < 5; v++)
                vt.emplace_back([&, v] {
                    cudaSetDevice(device);
                    rv[v] = one_vector(v);
                    if (rv[v]) ev[v] = g_err;
                });
            for (auto& t : vt) t.join();
            for (int v = 0; v < 5; v++)
                if (rv[v]) {
                    g_err = ev[v];
                    return rv[v];
                }
            return SS_OK;
        };
        int r = run();
        if (r && err) *err = g_err;
        return r;
    };
    if (D == 1) {
        if ((rc = worker(0, nullptr))) return rc;
        if (pairs) memcpy(pairs, partial[0].data(), partial[0].size());
        return SS_OK;
    }
    std::vector<std::thread> th;
    std::vector<int> rcs(D, SS_OK);
    std::vector<ss_error_info> errs(D);
    for (int di = 0; di < D; di++) th.emplace_back([&, di] { rcs[di] = worker(di, &errs[di]); });
    for (auto& t : th) t.join();
    for (int di = 0; di < D; di++)
        if (rcs[di]) {
            g_err = errs[di];
            return rcs[di];
        }
    if (!ratio_check || !pairs) return SS_OK;
    std::vector<const uint8_t*> blobs(D);
    for (int di = 0; di < D; di++) blobs[di] = partial[di].data();
    return sum_pair_blobs(p->curve, blobs.data(), D, pairs);
}

int ss_phase1_verification_vectors(const ss_phase1_params* p, const uint8_t* output, size_t output_len,
                                   int compressed_output, uint8_t* new_challenge, size_t new_challenge_len,
                                   int compressed_new_challenge, int subgroup_mode, int ratio_check,
                                   const uint8_t* rho_seed, uint8_t* pairs) {
    return phase1_verification_impl(p, output, output_len, compressed_output, new_challenge, new_challenge_len,
                                    compressed_new_challenge, subgroup_mode, ratio_check, rho_seed, pairs, true, nullptr);
}

int ss_phase1_verification_vectors_dev(const ss_phase1_params* p, const void* d_output, size_t output_len,
                                       int compressed_output, void* d_new_challenge, size_t new_challenge_len,
                                       int compressed_new_challenge, int subgroup_mode, int ratio_check,
                                       const uint8_t* rho_seed, uint8_t* pairs, void* stream) {
    return phase1_verification_impl(p, static_cast<const uint8_t*>(d_output), output_len, compressed_output,
                                    static_cast<uint8_t*>(d_new_challenge), new_challenge_len, compressed_new_challenge,
                                    subgroup_mode, ratio_check, rho_seed, pairs, false, static_cast<cudaStream_t>(stream));
}

// Index-range shard of the verification loop (see ss_phase1_computation_shard): `output` / `new_challenge` are the WHOLE
// buffers, only the shard's ranges (plus one overlap element of the response for the ratio pairs) are touched, and
// `pairs` receives the shard's PARTIAL (s, sx) — independent processes add theirs with ss_phase1_reduce_partial_pairs.
int ss_phase1_verification_vectors_shard(const ss_phase1_params* p, const uint8_t* output, size_t output_len,
                                         int compressed_output, uint8_t* new_challenge, size_t new_challenge_len,
                                         int compressed_new_challenge, int subgroup_mode, int ratio_check,
                                         const uint8_t* rho_seed, uint8_t* pairs, uint32_t shard_index, uint32_t shard_count) {
    return phase1_verification_impl(p, output, output_len, compressed_output, new_challenge, new_challenge_len,
                                    compressed_new_challenge, subgroup_mode, ratio_check, rho_seed, pairs, true, nullptr,
                                    shard_index, shard_count);
}

int ss_phase1_verification_vectors_shard_dev(const ss_phase1_params* p, const void* d_output, size_t output_len,
                                             int compressed_output, void* d_new_challenge, size_t new_challenge_len,
                                             int compressed_new_challenge, int subgroup_mode, int ratio_check,
                                             const uint8_t* rho_seed, uint8_t* pairs, uint32_t shard_index,
                                             uint32_t shard_count, void* stream) {
    return phase1_verification_impl(p, static_cast<const uint8_t*>(d_output), output_len, compressed_output,
                                    static_cast<uint8_t*>(d_new_challenge), new_challenge_len, compressed_new_challenge,
                                    subgroup_mode, ratio_check, rho_seed, pairs, false, static_cast<cudaStream_t>(stream),
                                    shard_index, shard_count);
}

size_t ss_phase1_pairs_size(int curve) {
    const GroupOps* g1 = group_ops(curve, SS_G1);
    const GroupOps* g2 = group_ops(curve, SS_G2);
    return (g1 && g2) ? 2 * (3 * (size_t)g1->usize + g2->usize) : 0;
}

int ss_phase1_reduce_partial_pairs(int curve, const uint8_t* partials, int count, uint8_t* pairs) {
    const size_t blob_b = ss_phase1_pairs_size(curve);
    if (!blob_b || !partials || count <= 0 || !pairs) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "bad argument");
    std::vector<const uint8_t*> blobs(count);
    for (int k = 0; k < count; k++) blobs[k] = partials + (size_t)k * blob_b;
    return sum_pair_blobs(curve, blobs.data(), count, pairs);
}

}  // extern "C"
