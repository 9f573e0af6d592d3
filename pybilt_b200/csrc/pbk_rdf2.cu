// pbk_rdf2.cu -- COM lateral RDF, cell-sorted half-shell sweep (the production RDF kernel).
//
// Same arithmetic contract as k_rdf in pbk_pairs.cu (which stays as the dense fallback for
// leaflets that do not fit in shared memory), restating
//   pybilt/bilayer_analyzer/analysis_protocols.py:4900-5006  (COMLateralRDFProtocol)
//   pybilt/common/distance_cutoff_clustering.py:27-71        (distance_euclidean_pbc)
// but organised for the hardware:
//   * one work item = (frame, leaflet); a persistent grid loops over the items;
//   * the leaflet's lipids of type A or B are counting-sorted into a 2-D cell grid in shared
//     memory (cell edge ~ range_outer/2), and every lipid sweeps only the forward half shell
//     of neighbouring cells, so each UNORDERED pair {p,q} is evaluated once;
//   * the reference evaluates ordered pairs dist(a=i in A, b=j in B).  Without a minimum-image
//     wrap d^2(p->q) == d^2(q->p) bit for bit, so one evaluation serves both directions; with a
//     wrap the two orders round differently ((L-|a'|)-|b'|) and both are evaluated;
//   * histogram bins are found from d^2 against thresholds thr[k] = min{t : sqrt_rn(t) >= edge_k}
//     (thr[n_bins] = min{t : sqrt_rn(t) > edge_last}), which is exactly np.histogram's
//     half-open-bins-last-closed rule on r = sqrt_rn(d^2) without a device sqrt or division;
//   * bins accumulate in warp-private shared histograms, flushed to global uint64 once per CTA.
#include "pbk_common.cuh"

#define RDF2_THREADS 256
#define RDF2_MAXCELLS 1024          // per dimension product cap (<= 32 x 32)

struct Rdf2Params {
    const double *com;
    const uint8_t *labels;
    const int32_t *type_code;
    const float *box;
    int64_t F, N;
    int lat0, lat1, ta, tb, leaflet_mask, nb;
    float lo_f, inv_w_f;
    double range_hi;
    const double *thr;
    unsigned long long *count;
    int32_t *frame_counts;
    int cap;                        // capacity of the per-leaflet shared arrays
    int hist_copies;
    int lut_n;                      // entries of the d^2 -> bin look-up table
    float lut_scale;                // lut index = (d2 - thr[0]) * lut_scale
    float thr0_f;
};

// bin of d2: look-up table on uniformly quantised d^2 gives a bin at most one off (every
// table cell holds at most one threshold when lut_n > n_bins^2), corrected against thr[].
__device__ __forceinline__ int rdf2_bin(double d2, const double *s_thr, const unsigned char *s_lut,
                                        const Rdf2Params &P)
{
    if (!(d2 >= s_thr[0]) || !(d2 < s_thr[P.nb])) return -1;
    int c = (int)(((float)d2 - P.thr0_f) * P.lut_scale);
    c = max(0, min(P.lut_n - 1, c));
    int k = s_lut[c];
    while (d2 >= s_thr[k + 1]) k++;
    while (d2 < s_thr[k]) k--;
    return k;
}

struct Rdf2Params;
__device__ __forceinline__ int rdf2_bin(double d2, const double *s_thr, const unsigned char *s_lut,
                                        const Rdf2Params &P);

template <bool SAME>
__global__ void __launch_bounds__(RDF2_THREADS)
k_rdf2(Rdf2Params P)
{
    extern __shared__ unsigned char smem_raw[];
    const int nb = P.nb, cap = P.cap;
    double *s_thr = reinterpret_cast<double *>(smem_raw);                   // nb+1
    double *s_bx = s_thr + (nb + 1);                                        // cap   b' = x - hx
    double *s_by = s_bx + cap;                                              // cap
    unsigned *s_hist = reinterpret_cast<unsigned *>(s_by + cap);            // copies*nb
    int *s_cell_start = reinterpret_cast<int *>(s_hist + (size_t)P.hist_copies * nb);  // MAXCELLS+1
    int *s_cell_fill = s_cell_start + (RDF2_MAXCELLS + 1);                  // MAXCELLS
    unsigned short *s_cellid = reinterpret_cast<unsigned short *>(s_cell_fill + RDF2_MAXCELLS);  // N
    unsigned short *s_scell = s_cellid + P.N;                               // cap (sorted): cell of slot
    unsigned char *s_flag = reinterpret_cast<unsigned char *>(s_scell + cap);    // cap (sorted)
    unsigned char *s_lut = s_flag + cap;                                    // lut_n
    __shared__ int s_cnt[3];
    __shared__ int s_scan[RDF2_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int q = tid; q <= nb; q += blockDim.x) s_thr[q] = P.thr[q];
    for (int q = tid; q < P.hist_copies * nb; q += blockDim.x) s_hist[q] = 0u;
    // look-up table: bin of the lower edge of every d^2 cell
    for (int c = tid; c < P.lut_n; c += blockDim.x) {
        const double t = (double)P.thr0_f + (double)c / (double)P.lut_scale;
        int k = 0;
        while (k < nb - 1 && t >= P.thr[k + 1]) k++;
        s_lut[c] = (unsigned char)k;
    }
    unsigned *myhist = s_hist + (size_t)(warp % P.hist_copies) * nb;
    __syncthreads();

    const int64_t n_items = P.F * 2;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int64_t f = item >> 1;
        const int leaf = (int)(item & 1);            // 0 = upper (label 1), 1 = lower (label 0)
        const unsigned want = leaf == 0 ? 1u : 0u;
        const bool selected = (P.leaflet_mask >> leaf) & 1;
        const double *cf = P.com + f * P.N * 3;
        const uint8_t *lab = P.labels + f * P.N;
        const float bxf = P.box[f * 3 + P.lat0], byf = P.box[f * 3 + P.lat1];
        const float cxf = bxf * 0.5f, cyf = byf * 0.5f;
        const double Lx = (double)bxf, Ly = (double)byf, hx = (double)cxf, hy = (double)cyf;
        // cell grid: edge >= 0.51*range_hi so that the half shell reaches 2 cells
        const double wmin = 0.51 * P.range_hi + 1e-6;
        int ncx = (int)floor(Lx / wmin), ncy = (int)floor(Ly / wmin);
        ncx = max(1, min(32, ncx));
        ncy = max(1, min(32, ncy));
        const double wx = Lx / ncx, wy = Ly / ncy;
        int mx = (int)floor((P.range_hi + 1e-6) / wx) + 1, my = (int)floor((P.range_hi + 1e-6) / wy) + 1;
        if (2 * mx + 1 > ncx) { ncx = 1; mx = 0; }
        if (2 * my + 1 > ncy) { ncy = 1; my = 0; }
        const double iwx = ncx / Lx, iwy = ncy / Ly;
        const int ncell = ncx * ncy;

        if (tid < 3) s_cnt[tid] = 0;
        for (int c = tid; c < ncell; c += blockDim.x) s_cell_fill[c] = 0;
        __syncthreads();
        // pass 1: classify, count, bin into cells
        int cA = 0, cB = 0, cL = 0;
        for (int64_t i = tid; i < P.N; i += blockDim.x) {
            unsigned short cid = 0xFFFF;
            if (lab[i] == want) {
                int t = P.type_code[i];
                bool a = (P.ta == -1 || t == P.ta), b = (P.tb == -1 || t == P.tb);
                cL++; cA += a; cB += b;
                if (a || b) {
                    int cx = (int)(cf[i * 3 + P.lat0] * iwx), cy = (int)(cf[i * 3 + P.lat1] * iwy);
                    cx = max(0, min(ncx - 1, cx));
                    cy = max(0, min(ncy - 1, cy));
                    cid = (unsigned short)(cy * ncx + cx);
                    atomicAdd(&s_cell_fill[cid], 1);
                }
            }
            s_cellid[i] = cid;
        }
        cA = pbk_warp_sum(cA); cB = pbk_warp_sum(cB); cL = pbk_warp_sum(cL);
        if (lane == 0) { atomicAdd(&s_cnt[0], cA); atomicAdd(&s_cnt[1], cB); atomicAdd(&s_cnt[2], cL); }
        __syncthreads();
        const bool on = selected && s_cnt[0] > 0 && s_cnt[1] > 0;
        if (tid < 3) {
            int v = s_cnt[tid];
            if (tid < 2 && !on) v = 0;
            if (tid == 2 && !selected) v = 0;
            P.frame_counts[f * 6 + leaf * 3 + tid] = v;
        }
        if (!on) { __syncthreads(); continue; }
        // exclusive scan of the cell counts (ncell <= 1024, blockDim = 256: 4 per thread)
        {
            int v[4], s = 0;
            for (int q = 0; q < 4; q++) { int c = tid * 4 + q; v[q] = c < ncell ? s_cell_fill[c] : 0; s += v[q]; }
            int incl = s;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            if (lane == 31) s_scan[warp] = incl;
            __syncthreads();
            int base = 0;
            for (int w = 0; w < warp; w++) base += s_scan[w];
            int excl = base + incl - s;
            for (int q = 0; q < 4; q++) { int c = tid * 4 + q; if (c < ncell) { s_cell_start[c] = excl; s_cell_fill[c] = excl; } excl += v[q]; }
            if (tid == blockDim.x - 1) s_cell_start[ncell] = excl;
        }
        __syncthreads();
        const int n_u = s_cell_start[ncell];
        // pass 2: scatter into cell order
        for (int64_t i = tid; i < P.N; i += blockDim.x) {
            unsigned short cid = s_cellid[i];
            if (cid == 0xFFFF) continue;
            int slot = atomicAdd(&s_cell_fill[cid], 1);
            int t = P.type_code[i];
            s_bx[slot] = __dsub_rn(cf[i * 3 + P.lat0], hx);
            s_by[slot] = __dsub_rn(cf[i * 3 + P.lat1], hy);
            s_flag[slot] = (unsigned char)(((P.ta == -1 || t == P.ta) ? 1 : 0) | ((P.tb == -1 || t == P.tb) ? 2 : 0));
            s_scell[slot] = cid;
        }
        __syncthreads();
        // pair sweep: 8 lanes per sorted lipid p; per row of the forward half shell the
        // candidate cells are one or two contiguous slot ranges (cells of a row are adjacent
        // in the sorted order), swept by the 8 lanes in lockstep
        const int sub = tid & 7, grp = tid >> 3, ngrp = blockDim.x >> 3;
        for (int p0 = 0; p0 < n_u; p0 += ngrp) {
            const int p = p0 + grp;
            const bool pact = p < n_u;
            double apx = 0, apy = 0, tpx = 0, tpy = 0;
            unsigned fp = 0;
            int pcx = 0, pcy = 0;
            if (pact) {
                apx = s_bx[p]; apy = s_by[p];
                fp = s_flag[p];
                tpx = __dsub_rn(Lx, fabs(apx)); tpy = __dsub_rn(Ly, fabs(apy));
                const int pc = s_scell[p];
                pcy = pc / ncx; pcx = pc - pcy * ncx;
            }
            for (int oy = 0; oy <= my; oy++) {
                int b1 = 0, n1 = 0, b2 = 0, n2 = 0;
                if (pact) {
                    int cy = pcy + oy; if (cy >= ncy) cy -= ncy;
                    const int row = cy * ncx;
                    int xlo = (oy == 0) ? pcx : pcx - mx, xhi = pcx + mx;
                    if (xlo < 0) {                      // wraps on the left
                        b2 = s_cell_start[row + xlo + ncx]; n2 = s_cell_start[row + ncx] - b2;
                        xlo = 0;
                    }
                    if (xhi >= ncx) {                   // wraps on the right
                        b2 = s_cell_start[row]; n2 = s_cell_start[row + xhi - ncx + 1] - b2;
                        xhi = ncx - 1;
                    }
                    b1 = s_cell_start[row + xlo]; n1 = s_cell_start[row + xhi + 1] - b1;
                    if (oy == 0) { n1 -= (p + 1 - b1); b1 = p + 1; }   // own cell: later entries only
                }
                const int ntot = n1 + n2;
                for (int t = sub; t < ntot; t += 8) {
                    const int q = t < n1 ? b1 + t : b2 + (t - n1);
                    int w = 2;
                    int w1 = 1, w2 = 1;
                    if (!SAME) {
                        const unsigned fq = s_flag[q];
                        w1 = (fp & 1u) && (fq & 2u);            // ordered p -> q
                        w2 = (fq & 1u) && (fp & 2u);            // ordered q -> p
                        w = w1 + w2;
                        if (!w) continue;
                    }
                    const double bqx = s_bx[q], bqy = s_by[q];
                    double dx = __dsub_rn(apx, bqx), dy = __dsub_rn(apy, bqy);
                    const bool wrx = fabs(dx) > hx, wry = fabs(dy) > hy;
                    if (!(wrx | wry)) {
                        const double d2 = fma(dy, dy, __dmul_rn(dx, dx));
                        const int k = rdf2_bin(d2, s_thr, s_lut, P);
                        if (k >= 0) atomicAdd(&myhist[k], (unsigned)w);
                    } else {
                        double dx2 = dx, dy2 = dy;              // the q -> p evaluation
                        if (wrx) { dx = __dsub_rn(tpx, fabs(bqx)); dx2 = __dsub_rn(__dsub_rn(Lx, fabs(bqx)), fabs(apx)); }
                        if (wry) { dy = __dsub_rn(tpy, fabs(bqy)); dy2 = __dsub_rn(__dsub_rn(Ly, fabs(bqy)), fabs(apy)); }
                        if (w1) {
                            const int k = rdf2_bin(fma(dy, dy, __dmul_rn(dx, dx)), s_thr, s_lut, P);
                            if (k >= 0) atomicAdd(&myhist[k], 1u);
                        }
                        if (w2) {
                            const int k = rdf2_bin(fma(dy2, dy2, __dmul_rn(dx2, dx2)), s_thr, s_lut, P);
                            if (k >= 0) atomicAdd(&myhist[k], 1u);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    for (int q = tid; q < nb; q += blockDim.x) {
        unsigned long long t = 0;
        for (int c = 0; c < P.hist_copies; c++) t += s_hist[(size_t)c * nb + q];
        if (t) atomicAdd(&P.count[q], t);
    }
}

// defined in pbk_pairs.cu: dense fallback
int pbk_rdf_dense(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                  const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                  int leaflet_mask, int32_t n_bins, double range_lo, double range_hi,
                  const double *edges, unsigned long long *count, int32_t *frame_counts,
                  cudaStream_t st);

extern "C" int pbk_rdf(pbk_ctx *ctx, const double *com, const uint8_t *labels,
                       const int32_t *type_code, const float *box, int64_t F, int64_t N, int lat0,
                       int lat1, int type_a, int type_b, int leaflet_mask, int32_t n_bins,
                       double range_lo, double range_hi, const double *edges, const double *thr,
                       unsigned long long *count, int32_t *frame_counts, void *stream)
{
    if (!ctx || !com || !labels || !type_code || !box || !edges || !count || !frame_counts ||
        F < 0 || N <= 0 || n_bins <= 0 || !(range_hi > range_lo) || (leaflet_mask & ~3) ||
        lat0 < 0 || lat0 > 2 || lat1 < 0 || lat1 > 2)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_rdf: bad argument");
    if (F == 0) return PBK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int copies = RDF2_THREADS / 32;
    int lut_n = n_bins * n_bins + n_bins + 16;
    if (lut_n > 16384) lut_n = 16384;
    auto smem_for = [&](int c) {
        return (size_t)(n_bins + 1) * 8 + (size_t)N * 16 + (size_t)c * n_bins * 4 +
               (size_t)(2 * RDF2_MAXCELLS + 1) * 4 + (size_t)N * 4 + (size_t)N + (size_t)lut_n + 16;
    };
    while (copies > 1 && smem_for(copies) > 64 * 1024) copies >>= 1;
    size_t smem = smem_for(copies);
    if (!thr || smem > 200 * 1024 || N >= 65535 || n_bins > 255)
        return pbk_rdf_dense(ctx, com, labels, type_code, box, F, N, lat0, lat1, type_a, type_b,
                             leaflet_mask, n_bins, range_lo, range_hi, edges, count, frame_counts, st);
    const bool same = (type_a == type_b);
    auto kern = same ? k_rdf2<true> : k_rdf2<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "k_rdf2 smem attribute", e);
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RDF2_THREADS, smem);
    if (e != cudaSuccess || per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)ctx->sm_count * per_sm;
    if (grid > F * 2) grid = F * 2;
    Rdf2Params P;
    P.com = com; P.labels = labels; P.type_code = type_code; P.box = box; P.F = F; P.N = N;
    P.lat0 = lat0; P.lat1 = lat1; P.ta = type_a; P.tb = type_b; P.leaflet_mask = leaflet_mask;
    P.nb = n_bins; P.lo_f = (float)range_lo; P.inv_w_f = (float)((double)n_bins / (range_hi - range_lo));
    P.range_hi = range_hi; P.thr = thr; P.count = count; P.frame_counts = frame_counts;
    P.cap = (int)N; P.hist_copies = copies;
    P.lut_n = lut_n;
    // thr[] lives on the device; the table is laid over [range_lo^2, range_hi^2] (thr[0] and
    // thr[n] differ from those by rounding only, and the kernel corrects +-1 cell anyway)
    double t0 = range_lo > 0 ? range_lo * range_lo : 0.0, t1 = range_hi * range_hi;
    P.thr0_f = (float)t0;
    P.lut_scale = (float)((double)lut_n / (t1 - t0));
    kern<<<(unsigned)grid, RDF2_THREADS, smem, st>>>(P);
    PBK_CHECK_LAUNCH(ctx, "k_rdf2");
    return PBK_OK;
}
