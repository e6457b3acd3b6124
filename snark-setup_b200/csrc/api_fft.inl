// Included by api.cu (shares its helpers): Groth16Params::new + ::write, i.e. `prepare_phase2`
// (SURVEY.md §8f rank 3).
//
//   to_coeffs (group IFFT + normalize_batch)      setup-utils/src/groth16_utils.rs:44-53
//   h_query_groth16                               setup-utils/src/groth16_utils.rs:59-63
//   Groth16Params::new / ::write                  setup-utils/src/groth16_utils.rs:81-131,134-168
//   prepare_phase2                                phase2-cli/src/prepare_phase2.rs:16-70
//
// The reference converts the in-memory vectors with ark-poly's generic FFT over `C::Group` on rayon
// threads (n/2 * log n full-width scalar multiplications per vector).  Here one vector stays resident
// in HBM for the whole transform; see fft.cuh for the stage structure.

namespace {

const FftOps* fft_ops(int curve, int group) {
    if (curve == SS_CURVE_BLS12_377) return group == SS_G1 ? &fft_ops_bls377_g1() : group == SS_G2 ? &fft_ops_bls377_g2() : nullptr;
    if (curve == SS_CURVE_BW6_761) return group == SS_G1 ? &fft_ops_bw6_g1() : group == SS_G2 ? &fft_ops_bw6_g2() : nullptr;
    return nullptr;
}

int log2_exact(uint64_t n) {
    if (n == 0 || (n & (n - 1))) return -1;
    int l = 0;
    while ((1ull << l) < n) l++;
    return l;
}

// device scratch one fft_vector call carves from its lane
size_t fft_scratch_bytes(const GroupOps& o, int log_n, uint64_t h_count, int in_c, int out_c) {
    const uint64_t n = 1ull << log_n, nd = n + h_count, half = n >> 1;
    const size_t isz = in_c ? o.csize : o.usize, osz = out_c ? o.csize : o.usize;
    const size_t cw = (size_t)o.coord_words * 4, frw = o.fr_words;
    const size_t io_b = std::max(isz * nd, osz * n);
    const size_t seq_b = (size_t)(log_n + 64) * frw * 4;
    return 256 + align_up(io_b, 256) + align_up(2 * cw * nd, 256) + align_up(nd, 256) + align_up(2 * cw * n, 256) +
           align_up(n, 256) + align_up(3 * cw * std::max<uint64_t>(half, 2), 256) + align_up(3 * cw * n, 256) +
           align_up(cw * n, 256) + align_up(seq_b, 256) + 256;
}

// One vector: decode n + h_count elements (n = 2^log_n), optionally emit the H query
// h_i = P_(i+n) - P_i for i < h_count, then the n Lagrange coefficients.  in/out/h_out are HOST buffers.
int fft_vector(int curve, int group, const uint8_t* in, int in_c, int check, int log_n, uint64_t h_count, uint8_t* out,
               uint8_t* h_out, int out_c, const char* what, int device_slot = 0) {
    const GroupOps* op = group_ops(curve, group);
    const FftOps* fp = fft_ops(curve, group);
    if (!op || !fp) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "unknown curve/group");
    if (log_n < 0 || log_n > fp->max_log_n || log_n > 40)
        return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "%s: no radix-2 domain of size 2^%d", what, log_n);
    if (check < 0 || check > 3) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "bad check mode");
    if (!in || !out || (h_count && !h_out)) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "null buffer");
    const GroupOps& o = *op;
    const FftOps& f = *fp;
    int rc = ensure_init();
    if (rc) return rc;
    const int device = g_devices[(size_t)device_slot % g_devices.size()];
    const uint64_t n = 1ull << log_n, nd = n + h_count, half = n >> 1;
    const size_t isz = in_c ? o.csize : o.usize, osz = out_c ? o.csize : o.usize;
    const size_t cw = (size_t)o.coord_words * 4, frw = o.fr_words;
    const size_t io_b = std::max(isz * nd, osz * n);  // staged input, later reused for the serialized output
    const size_t seq_b = (size_t)(log_n + 64) * frw * 4;
    const size_t need = fft_scratch_bytes(o, log_n, h_count, in_c, out_c);
    size_t free_b = 0, total_b = 0;
    CU(cudaSetDevice(device));
    CU(cudaMemGetInfo(&free_b, &total_b));
    if (need > total_b) return fail(SS_ERR_DEVICE, 0, total_b, need, "%s: 2^%d elements need %zu bytes of HBM", what, log_n, need);
    LaneGuard lg;
    if ((rc = lane_acquire(device, need, &lg.l))) return rc;
    cudaStream_t s = lg.l->stream;
    Carver cv(lg.l->buf);
    unsigned long long* d_status = cv.take<unsigned long long>(8);
    uint8_t* io = cv.take<uint8_t>(io_b);
    uint32_t* aff_s = cv.take<uint32_t>(2 * cw * nd);  // decoded input, natural order (the H query reads it)
    uint8_t* inf_s = cv.take<uint8_t>(nd);
    uint32_t* aff = cv.take<uint32_t>(2 * cw * n);  // working set of the transform
    uint8_t* inf = cv.take<uint8_t>(n);
    uint32_t* jac_t = cv.take<uint32_t>(3 * cw * std::max<uint64_t>(half, 2));
    uint32_t* jac_o = cv.take<uint32_t>(3 * cw * n);
    uint32_t* prefix = cv.take<uint32_t>(cw * n);
    uint32_t* seq = cv.take<uint32_t>(seq_b);
    uint32_t* ninv_m = cv.take<uint32_t>(frw * 4);

    CU(cudaMemsetAsync(d_status, 0xff, 8, s));
    CU(cudaMemcpyAsync(io, in, isz * nd, cudaMemcpyHostToDevice, s));
    DecodeArgs da = {reinterpret_cast<const uint32_t*>(io), in_c, check, nd, aff_s, inf_s, d_status};
    { ProfScope ps("k_decode", o.name, nd, s); o.decode(da, s); }
    unsigned long long st = STATUS_OK;
    CU(cudaMemcpyAsync(&st, d_status, 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if ((rc = decode_status(st, 0, what))) {
        prof_flush();
        return rc;
    }
    auto normalize = [&](const uint32_t* jac, uint64_t cnt, uint8_t* bytes_out) {
        NormalizeArgs na;
        na.jac = jac;
        na.n = cnt;
        na.prefix = prefix;
        na.out = reinterpret_cast<uint32_t*>(bytes_out);
        na.out_compressed = out_c;
        // mid-size transforms: more, shorter batch inversions (n/32 threads would leave most SMs idle)
        na.threads = std::max<uint32_t>(normalize_threads(cnt), (uint32_t)std::min<uint64_t>(std::max<uint64_t>(cnt / 4, 1), 32768));
        if (!bytes_out) {
            na.aff_out = aff;
            na.inf_out = inf;
        }
        ProfScope ps("k_normalize_encode", o.name, cnt, s);
        o.normalize_encode(na, s);
    };
    if (h_count) {  // h_query_groth16: the input bytes in `io` are no longer needed
        { ProfScope ps("k_h_query", o.name, h_count, s); f.h_query(aff_s, inf_s, n, h_count, jac_o, s); }
        normalize(jac_o, h_count, io);
        CU(cudaMemcpyAsync(h_out, io, osz * h_count, cudaMemcpyDeviceToHost, s));
    }
    if (log_n == 0) {  // domain of size 1: the coefficient is the point itself
        EncodeArgs ea = {aff_s, inf_s, 1, reinterpret_cast<uint32_t*>(io), out_c, 1};
        ProfScope ps("k_encode", o.name, 1, s);
        o.encode(ea, s);
    } else {
        f.prepare(log_n, seq, ninv_m, s);
        { ProfScope ps("k_fft_bitrev", o.name, n, s); f.bitrev(aff_s, inf_s, aff, inf, log_n, s); }
        ScalarMulArgs a;
        a.aff = aff;
        a.inf = inf;
        a.exps = nullptr;
        a.first_power = 0;
        a.coeff_m = ninv_m;
        // ifft's scaling by n^-1 rides on the twiddles: block 0 of stage s is lo + t with lo = block 0 of
        // stage s-1 (already scaled, by induction) and t = (n^-1 * w_s^j) * hi, one Fr multiplication more per
        // thread; only elements 0 and 1 (stage 0 has no twiddles) are scaled by a scalar multiplication of their own.
        a.n = 2;
        a.tau_tab = seq;
        a.has_coeff = 1;
        a.power_mask = 0;
        a.src_log_m = -1;
        a.jac = jac_t;
        { ProfScope ps("k_scalar_mul", o.name, 2, s); o.scalar_mul(a, s); }
        normalize(jac_t, 2, nullptr);
        for (int st_i = 0; st_i < log_n; st_i++) {
            const uint32_t* tw = nullptr;
            uint64_t tw_count = half;
            if (st_i > 0) {
                const uint64_t m = 1ull << st_i;
                tw_count = half - half / m + 1;  // every j >= 1, plus (block 0, j 0) whose twiddle is n^-1
                a.fft_compact = 1;
                a.n = tw_count;
                a.tau_tab = seq + (size_t)(log_n - 1 - st_i) * frw;  // table of w_s^(2^j), w_s = w^(-n/2m)
                a.has_coeff = 1;
                a.coeff_limit = 1ull << st_i;  // block 0
                a.power_mask = (1ull << st_i) - 1;
                a.src_log_m = st_i;
                a.jac = jac_t;
                ProfScope ps("k_scalar_mul", o.name, tw_count, s);
                o.scalar_mul(a, s);
                tw = jac_t;
            }
            ButterflyArgs b = {aff, inf, tw, n, st_i, jac_o, tw_count, st_i > 0 ? 1 : 0};
            { ProfScope ps("k_fft_butterfly", o.name, half, s); f.butterfly(b, s); }
            normalize(jac_o, n, st_i == log_n - 1 ? io : nullptr);
        }
    }
    CU(cudaMemcpyAsync(out, io, osz * n, cudaMemcpyDeviceToHost, s));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s));
    prof_flush();
    return SS_OK;
}

}  // namespace

extern "C" {

int ss_group_ifft(int curve, int group, const uint8_t* in, int in_compressed, int check, size_t n, uint8_t* out,
                  int out_compressed) {
    const int l = log2_exact(n);
    if (l < 0) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, n, "group_ifft: %zu is not a power of two", n);
    return fft_vector(curve, group, in, in_compressed, check, l, 0, out, nullptr, out_compressed, "to_coeffs");
}

int ss_h_query_groth16(int curve, const uint8_t* powers, int in_compressed, int check, size_t n_powers, size_t degree,
                       uint8_t* out, int out_compressed) {
    const GroupOps* o = group_ops(curve, SS_G1);
    if (!o) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "unknown curve");
    if (degree == 0) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "h_query: degree 0");
    if (degree == 1) return SS_OK;  // 0..degree-1 is empty
    // powers[i + degree] for i <= degree - 2 (groth16_utils.rs:61): index out of bounds panics in the reference
    if (n_powers < 2 * degree - 1) return fail(SS_ERR_INVALID_LENGTH, 0, 2 * degree - 1, n_powers, "h_query: too few powers");
    if (!powers || !out) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "null buffer");
    int rc = ensure_init();
    if (rc) return rc;
    const int device = g_devices[0];
    const uint64_t nd = 2 * degree - 1, hc = degree - 1;
    const size_t isz = in_compressed ? o->csize : o->usize, osz = out_compressed ? o->csize : o->usize;
    const size_t cw = (size_t)o->coord_words * 4;
    const FftOps* f = fft_ops(curve, SS_G1);
    LaneGuard lg;
    const size_t need = 256 + align_up(isz * nd, 256) + align_up(2 * cw * nd, 256) + align_up(nd, 256) +
                        align_up(3 * cw * hc, 256) + align_up(cw * hc, 256);
    if ((rc = lane_acquire(device, need, &lg.l))) return rc;
    cudaStream_t s = lg.l->stream;
    Carver cv(lg.l->buf);
    unsigned long long* d_status = cv.take<unsigned long long>(8);
    uint8_t* io = cv.take<uint8_t>(isz * nd);
    uint32_t* aff = cv.take<uint32_t>(2 * cw * nd);
    uint8_t* inf = cv.take<uint8_t>(nd);
    uint32_t* jac = cv.take<uint32_t>(3 * cw * hc);
    uint32_t* prefix = cv.take<uint32_t>(cw * hc);
    CU(cudaMemsetAsync(d_status, 0xff, 8, s));
    CU(cudaMemcpyAsync(io, powers, isz * nd, cudaMemcpyHostToDevice, s));
    DecodeArgs da = {reinterpret_cast<const uint32_t*>(io), in_compressed, check, nd, aff, inf, d_status};
    { ProfScope ps("k_decode", o->name, nd, s); o->decode(da, s); }
    { ProfScope ps("k_h_query", o->name, hc, s); f->h_query(aff, inf, degree, hc, jac, s); }
    NormalizeArgs na;
    na.jac = jac;
    na.n = hc;
    na.prefix = prefix;
    na.out = reinterpret_cast<uint32_t*>(io);
    na.out_compressed = out_compressed;
    na.threads = normalize_threads(hc);
    { ProfScope ps("k_normalize_encode", o->name, hc, s); o->normalize_encode(na, s); }
    unsigned long long st = STATUS_OK;
    CU(cudaMemcpyAsync(&st, d_status, 8, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(out, io, osz * hc, cudaMemcpyDeviceToHost, s));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s));
    prof_flush();
    return decode_status(st, 0, "h_query");
}

int ss_groth16_params_size(int curve, uint64_t phase2_size, int compressed, uint64_t* domain_size, size_t* bytes) {
    const GroupOps* g1 = group_ops(curve, SS_G1);
    const GroupOps* g2 = group_ops(curve, SS_G2);
    if (!g1 || !g2) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "unknown curve");
    // GeneralEvaluationDomain::new(phase2_size).size(): the next power of two (groth16_utils.rs:65-69,98-99)
    uint64_t m = 1;
    while (m < phase2_size) {
        m <<= 1;
        if (!m) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, phase2_size, "phase2_size too large");
    }
    const size_t s1 = compressed ? g1->csize : g1->usize, s2 = compressed ? g2->csize : g2->usize;
    if (domain_size) *domain_size = m;
    if (bytes) *bytes = 2 * s1 + s2 + 3 * m * s1 + m * s2 + (m - 1) * s1;
    return SS_OK;
}

int ss_groth16_params_new(const ss_phase1_params* p, const uint8_t* accumulator, size_t accumulator_len,
                          int compressed_input, int check, uint64_t phase2_size, uint8_t* out, size_t out_len,
                          int compressed_output) {
    if (!p || !accumulator || !out) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "null argument");
    if (p->proving_system != SS_GROTH16 || p->contribution_mode != SS_MODE_FULL)
        return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "Groth16Params::new takes a full Groth16 accumulator");
    ss_phase1_sizes z;
    int rc = phase1_sizes(p, &z);
    if (rc) return rc;
    const GroupOps& g1 = *group_ops(p->curve, SS_G1);
    const GroupOps& g2 = *group_ops(p->curve, SS_G2);
    const size_t i1 = compressed_input ? g1.csize : g1.usize, o1 = compressed_output ? g1.csize : g1.usize,
                 o2 = compressed_output ? g2.csize : g2.usize;
    uint64_t off[5];
    vector_offsets(p->curve, z, compressed_input, off);
    const uint64_t acc_need = off[4] + (compressed_input ? g2.csize : g2.usize);
    if (accumulator_len < acc_need) return fail(SS_ERR_INVALID_LENGTH, 0, acc_need, accumulator_len, "accumulator too short");
    uint64_t m = 0;
    size_t need = 0;
    if ((rc = ss_groth16_params_size(p->curve, phase2_size, compressed_output, &m, &need))) return rc;
    // `&tau_powers_g2[0..phase2_size]` panics in the reference when the domain exceeds the accumulator
    if (m > z.powers_length) return fail(SS_ERR_INVALID_LENGTH, 0, z.powers_length, m, "phase2 domain larger than the powers of tau");
    if (out_len < need) return fail(SS_ERR_INVALID_LENGTH, 0, need, out_len, "output buffer too short");
    const int lm = log2_exact(m);
    uint8_t* w = out;
    // alpha_g1 = alpha_tau_powers_g1[0], beta_g1 = beta_tau_powers_g1[0], beta_g2
    if ((rc = ss_transcode(p->curve, SS_G1, accumulator + off[2], compressed_input, check, w, compressed_output, 1))) return rc;
    w += o1;
    if ((rc = ss_transcode(p->curve, SS_G1, accumulator + off[3], compressed_input, check, w, compressed_output, 1))) return rc;
    w += o1;
    if ((rc = ss_transcode(p->curve, SS_G2, accumulator + off[4], compressed_input, check, w, compressed_output, 1))) return rc;
    w += o2;
    uint8_t* coeffs_g1 = w;
    uint8_t* coeffs_g2 = coeffs_g1 + m * o1;
    uint8_t* alpha_g1 = coeffs_g2 + m * o2;
    uint8_t* beta_g1 = alpha_g1 + m * o1;
    uint8_t* h_g1 = beta_g1 + m * o1;
    (void)i1;
    // The four transforms are independent (the reference spawns one crossbeam thread each, groth16_utils.rs:103-111):
    // one host thread + lane each, spread over the selected devices (ss_init / $SNARK_SETUP_GPUS) round-robin;
    // on one device they overlap, so the low-occupancy normalisation tails of one hide behind another's
    // scalar multiplications.  The heaviest (G2) goes first.
    if ((rc = ensure_init())) return rc;
    struct Job {
        int group;
        const uint8_t* in;
        uint64_t h;
        uint8_t* out;
        uint8_t* h_out;
        const char* what;
    };
    const Job jobs[4] = {{SS_G2, accumulator + off[1], 0, coeffs_g2, nullptr, "tau_g2 coefficients"},
                         {SS_G1, accumulator + off[0], m - 1, coeffs_g1, h_g1, "tau_g1 coefficients"},
                         {SS_G1, accumulator + off[2], 0, alpha_g1, nullptr, "alpha_g1 coefficients"},
                         {SS_G1, accumulator + off[3], 0, beta_g1, nullptr, "beta_g1 coefficients"}};
    int rcs[4] = {0, 0, 0, 0};
    ss_error_info errs[4];
    auto run = [&](int v) {
        rcs[v] = fft_vector(p->curve, jobs[v].group, jobs[v].in, compressed_input, check, lm, jobs[v].h, jobs[v].out,
                            jobs[v].h_out, compressed_output, jobs[v].what, v);
        if (rcs[v]) errs[v] = g_err;
    };
    // concurrent lanes need all four scratch slabs at once: fall back to one transform at a time when the
    // device holding the most of them could not fit its share
    bool concurrent = concurrent_vectors();
    if (concurrent) {
        const size_t D = g_devices.size();
        std::vector<size_t> per_dev(D, 0);
        for (int v = 0; v < 4; v++)
            per_dev[(size_t)v % D] += fft_scratch_bytes(*group_ops(p->curve, jobs[v].group), lm, jobs[v].h, compressed_input, compressed_output);
        for (size_t d = 0; d < D && concurrent; d++) {
            size_t free_b = 0, total_b = 0;
            if (cudaSetDevice(g_devices[d]) != cudaSuccess || cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) concurrent = false;
            else {
                // what the transforms can really get: free HBM plus the idle cached slabs of this device (lane_acquire
                // re-grows them), not the card's total — other lanes / processes may hold memory
                size_t idle = 0;
                {
                    std::lock_guard<std::mutex> lk(g_mu);
                    for (Lane* l : g_lanes)
                        if (!l->busy && l->device == g_devices[d]) idle += l->cap;
                }
                if (per_dev[d] > (free_b + idle) / 10 * 9 || per_dev[d] > total_b / 10 * 8) concurrent = false;
            }
        }
    }
    if (concurrent) {
        std::vector<std::thread> th;
        for (int v = 0; v < 4; v++) th.emplace_back(run, v);
        for (auto& t : th) t.join();
    } else {
        for (int v = 0; v < 4; v++) run(v);
    }
    for (int v = 0; v < 4; v++)
        if (rcs[v]) {
            g_err = errs[v];
            return rcs[v];
        }
    return SS_OK;
}

}  // extern "C"
