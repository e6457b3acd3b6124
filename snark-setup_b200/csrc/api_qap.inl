// Included by api.cu: the sparse dot products of the phase-2 QAP evaluation (phase2/src/polynomial.rs:11-94,
// called from MPCParameters::new, phase2/src/parameters.rs:110-131).

namespace {

const QapOps* qap_ops(int curve, int group) {
    if (curve == SS_CURVE_BLS12_377) return group == SS_G1 ? &qap_ops_bls377_g1() : group == SS_G2 ? &qap_ops_bls377_g2() : nullptr;
    if (curve == SS_CURVE_BW6_761) return group == SS_G1 ? &qap_ops_bw6_g1() : group == SS_G2 ? &qap_ops_bw6_g2() : nullptr;
    return nullptr;
}

constexpr uint64_t QAP_SEGMENT = 256;          // entries per partial sum
constexpr uint64_t QAP_BLOCK_ENTRIES = 1u << 21;  // entries per row block (a longer single row is its own block)
constexpr uint64_t QAP_BLOCK_ROWS = 1u << 18;

struct QapBlock {
    uint64_t row0, row1, ent0, ent1, nseg, ngeneral;
};

}  // namespace

extern "C" {

int ss_qap_dot_product(int curve, int group, const uint8_t* bases, int bases_compressed, int check, size_t n_bases,
                       const uint64_t* row_ptr, const uint32_t* index, const uint8_t* coeffs, size_t rows, uint8_t* out,
                       int out_compressed) {
    const GroupOps* op = group_ops(curve, group);
    const QapOps* qp = qap_ops(curve, group);
    if (!op || !qp) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "unknown curve/group");
    if (rows == 0) return SS_OK;
    if (!row_ptr || !out) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "null buffer");
    if (check < 0 || check > 3) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "bad check mode");
    const GroupOps& o = *op;
    const uint64_t nnz = row_ptr[rows];
    if (row_ptr[0] != 0) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, row_ptr[0], "row_ptr[0] must be 0");
    for (size_t v = 0; v < rows; v++)
        if (row_ptr[v + 1] < row_ptr[v]) return fail(SS_ERR_INVALID_ARGUMENT, v, 0, 0, "row_ptr is not monotone at row %zu", v);
    if (nnz && (!index || !coeffs || !bases)) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "null buffer");
    if (n_bases > 0xffffffffull || nnz > 0xffffffffull) return fail(SS_ERR_INVALID_ARGUMENT, 0, 0, 0, "too many bases / entries");
    const size_t fb = o.fr_bytes;
    int rn;
    const uint32_t* rmod = scalar_modulus(curve, &rn);
    std::vector<uint8_t> one(fb, 0), minus_one(fb, 0);
    one[0] = 1;
    memcpy(minus_one.data(), rmod, fb);  // r - 1: r is odd with low limb 1
    minus_one[0] = 0;
    // classify the entries: 0 = +1, 1 = -1, 2 = general, 3 = zero
    std::vector<uint8_t> kind(nnz);
    for (uint64_t e = 0; e < nnz; e++) {
        if (index[e] >= n_bases)  // `coeffs[ind]` panics in the reference
            return fail(SS_ERR_INVALID_LENGTH, e, n_bases, index[e], "entry %llu: base index %u out of range", (unsigned long long)e, index[e]);
        const uint8_t* c = coeffs + e * fb;
        if (!scalar_is_canonical(curve, c)) return fail(SS_ERR_INVALID_DATA, e, 0, 0, "entry %llu: scalar not canonical", (unsigned long long)e);
        bool zero = true;
        for (size_t i = 0; i < fb && zero; i++) zero = c[i] == 0;
        kind[e] = zero ? 3 : !memcmp(c, one.data(), fb) ? 0 : !memcmp(c, minus_one.data(), fb) ? 1 : 2;
    }
    // row blocks
    std::vector<QapBlock> blocks;
    uint64_t max_ent = 1, max_rows = 1, max_seg = 1, max_gen = 1;
    for (uint64_t v = 0; v < rows;) {
        QapBlock b = {v, v, row_ptr[v], row_ptr[v], 0, 0};
        while (b.row1 < rows && b.row1 - b.row0 < QAP_BLOCK_ROWS) {
            const uint64_t len = row_ptr[b.row1 + 1] - row_ptr[b.row1];
            if (b.row1 > b.row0 && b.ent1 - b.ent0 + len > QAP_BLOCK_ENTRIES) break;
            b.ent1 += len;
            b.nseg += (len + QAP_SEGMENT - 1) / QAP_SEGMENT;
            b.row1++;
        }
        for (uint64_t e = b.ent0; e < b.ent1; e++) b.ngeneral += kind[e] == 2;
        max_ent = std::max(max_ent, b.ent1 - b.ent0);
        max_rows = std::max(max_rows, b.row1 - b.row0);
        max_seg = std::max(max_seg, b.nseg);
        max_gen = std::max(max_gen, b.ngeneral);
        blocks.push_back(b);
        v = b.row1;
    }
    int rc = ensure_init();
    if (rc) return rc;
    const int device = g_devices[0];
    const size_t isz = bases_compressed ? o.csize : o.usize, osz = out_compressed ? o.csize : o.usize;
    const size_t cw = (size_t)o.coord_words * 4;
    const size_t nb = std::max<size_t>(n_bases, 1);
    const size_t need = 256 + align_up(isz * nb, 256) + align_up(2 * cw * nb, 256) + align_up(nb, 256) + align_up(4 * max_ent, 256) +
                        align_up(max_ent, 256) + align_up(8 * (max_seg + 1), 256) + align_up(8 * (max_rows + 1), 256) +
                        align_up(4 * max_gen, 256) + align_up(fb * max_gen, 256) + align_up(3 * cw * max_gen, 256) +
                        2 * align_up(3 * cw * max_seg, 256) + align_up(8 * (max_seg + 1), 256) + align_up(3 * cw * max_rows, 256) +
                        align_up(cw * max_rows, 256) + align_up(osz * max_rows, 256);
    LaneGuard lg;
    if ((rc = lane_acquire(device, need, &lg.l))) return rc;
    cudaStream_t s = lg.l->stream;
    Carver cv(lg.l->buf);
    unsigned long long* d_status = cv.take<unsigned long long>(8);
    uint8_t* d_in = cv.take<uint8_t>(isz * nb);
    uint32_t* aff = cv.take<uint32_t>(2 * cw * nb);
    uint8_t* inf = cv.take<uint8_t>(nb);
    uint32_t* d_index = cv.take<uint32_t>(4 * max_ent);
    uint8_t* d_kind = cv.take<uint8_t>(max_ent);
    uint64_t* d_seg = cv.take<uint64_t>(8 * (max_seg + 1));
    uint64_t* d_rowseg = cv.take<uint64_t>(8 * (max_rows + 1));
    uint32_t* d_gather = cv.take<uint32_t>(4 * max_gen);
    uint8_t* d_gcoeff = cv.take<uint8_t>(fb * max_gen);
    uint32_t* jac_g = cv.take<uint32_t>(3 * cw * max_gen);
    uint32_t* partial = cv.take<uint32_t>(3 * cw * max_seg);
    uint32_t* partial2 = cv.take<uint32_t>(3 * cw * max_seg);
    uint64_t* d_seg2 = cv.take<uint64_t>(8 * (max_seg + 1));
    uint32_t* jac_rows = cv.take<uint32_t>(3 * cw * max_rows);
    uint32_t* prefix = cv.take<uint32_t>(cw * max_rows);
    uint8_t* d_out = cv.take<uint8_t>(osz * max_rows);

    CU(cudaMemsetAsync(d_status, 0xff, 8, s));
    if (n_bases) {
        CU(cudaMemcpyAsync(d_in, bases, isz * n_bases, cudaMemcpyHostToDevice, s));
        DecodeArgs da = {reinterpret_cast<const uint32_t*>(d_in), bases_compressed, check, n_bases, aff, inf, d_status};
        ProfScope ps("k_decode", o.name, n_bases, s);
        o.decode(da, s);
    }
    unsigned long long st = STATUS_OK;
    CU(cudaMemcpyAsync(&st, d_status, 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if ((rc = decode_status(st, 0, "dot_product bases"))) {
        prof_flush();
        return rc;
    }
    std::vector<uint32_t> h_index, h_gather;
    std::vector<uint64_t> h_seg, h_seg2, h_rowseg2;
    std::vector<uint8_t> h_gcoeff;
    for (const QapBlock& b : blocks) {
        const uint64_t ne = b.ent1 - b.ent0, nr = b.row1 - b.row0;
        h_index.assign(ne, 0);
        h_gather.clear();
        h_gcoeff.clear();
        h_seg.clear();
        // two levels of segments: entries -> partial sums of <= 256 entries -> partial sums of <= 256 partials
        // -> row sums (a row of 2^24 entries is still only 256 additions deep at every level)
        h_seg2.clear();
        h_rowseg2.clear();
        for (uint64_t v = b.row0; v < b.row1; v++) {
            const uint64_t first_seg = h_seg.size();
            for (uint64_t e = row_ptr[v]; e < row_ptr[v + 1]; e += QAP_SEGMENT) h_seg.push_back(e - b.ent0);
            h_rowseg2.push_back(h_seg2.size());
            for (uint64_t sg = first_seg; sg < h_seg.size(); sg += QAP_SEGMENT) h_seg2.push_back(sg);
        }
        h_rowseg2.push_back(h_seg2.size());
        h_seg2.push_back(h_seg.size());
        h_seg.push_back(ne);
        for (uint64_t e = b.ent0; e < b.ent1; e++) {
            if (kind[e] == 2) {
                h_index[e - b.ent0] = (uint32_t)h_gather.size();
                h_gather.push_back(index[e]);
                h_gcoeff.insert(h_gcoeff.end(), coeffs + e * fb, coeffs + (e + 1) * fb);
            } else {
                h_index[e - b.ent0] = index[e];
            }
        }
        const uint64_t ng = h_gather.size(), nseg = h_seg.size() - 1;
        if (ne) {
            CU(cudaMemcpyAsync(d_index, h_index.data(), 4 * ne, cudaMemcpyHostToDevice, s));
            CU(cudaMemcpyAsync(d_kind, kind.data() + b.ent0, ne, cudaMemcpyHostToDevice, s));
        }
        const uint64_t nseg2 = h_seg2.size() - 1;
        CU(cudaMemcpyAsync(d_seg, h_seg.data(), 8 * (nseg + 1), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(d_seg2, h_seg2.data(), 8 * (nseg2 + 1), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(d_rowseg, h_rowseg2.data(), 8 * (nr + 1), cudaMemcpyHostToDevice, s));
        if (ng) {
            CU(cudaMemcpyAsync(d_gather, h_gather.data(), 4 * ng, cudaMemcpyHostToDevice, s));
            CU(cudaMemcpyAsync(d_gcoeff, h_gcoeff.data(), fb * ng, cudaMemcpyHostToDevice, s));
            ScalarMulArgs a;
            a.aff = aff;
            a.inf = inf;
            a.n = ng;
            a.exps = reinterpret_cast<const uint32_t*>(d_gcoeff);
            a.tau_tab = nullptr;
            a.first_power = 0;
            a.coeff_m = nullptr;
            a.has_coeff = 0;
            a.jac = jac_g;
            a.gather = d_gather;
            ProfScope ps("k_scalar_mul", o.name, ng, s);
            o.scalar_mul(a, s);
        }
        QapSegArgs qa = {aff, inf, d_index, d_kind, d_seg, nseg, jac_g, std::max<uint64_t>(ng, 1), partial};
        { ProfScope ps("k_qap_segment_sum", o.name, ne, s); qp->segment_sum(qa, s); }
        { ProfScope ps("k_qap_row_sum", o.name, nseg2, s); qp->row_sum(partial, nseg, d_seg2, nseg2, partial2, s); }
        { ProfScope ps("k_qap_row_sum", o.name, nr, s); qp->row_sum(partial2, nseg2, d_rowseg, nr, jac_rows, s); }
        NormalizeArgs na;
        na.jac = jac_rows;
        na.n = nr;
        na.prefix = prefix;
        na.out = reinterpret_cast<uint32_t*>(d_out);
        na.out_compressed = out_compressed;
        na.threads = normalize_threads(nr);
        { ProfScope ps("k_normalize_encode", o.name, nr, s); o.normalize_encode(na, s); }
        CU(cudaMemcpyAsync(out + b.row0 * osz, d_out, osz * nr, cudaMemcpyDeviceToHost, s));
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(s));  // the host staging vectors are reused by the next block
    }
    prof_flush();
    return SS_OK;
}

}  // extern "C"
