// pbk_rdf2.cu -- COM lateral RDF, cell-sorted half-shell sweep (the production RDF kernel).
//
// Same arithmetic contract as k_rdf in pbk_pairs.cu (which stays as the dense fallback for
// leaflets that do not fit in shared memory), restating
//   pybilt/bilayer_analyzer/analysis_protocols.py:4900-5006  (COMLateralRDFProtocol)
//   pybilt/common/distance_cutoff_clustering.py:27-71        (distance_euclidean_pbc)
// but organised for the hardware:
//   * one work item = (frame, leaflet); a persistent grid loops over the items;
//   * the leaflet's lipids of type A or B are counting-sorted into a 2-D cell grid in shared
//     memory (cell edge ~ range_outer/2); every lipid p (8 lanes each) sweeps the forward half
//     shell of neighbouring cells -- per cell row one or two contiguous slot ranges -- so each
//     UNORDERED pair {p,q} is visited once;
//   * phase 1 (all candidates, float32): minimum-image d^2 against (range_outer + 0.01)^2; the
//     survivors (p,q) are compacted into a per-warp ring queue;
//   * phase 2 (survivors, 32 lanes in lockstep, float64): the reference's exact arithmetic.
//     It evaluates ordered pairs dist(a=i in A, b=j in B); without a minimum-image wrap
//     d^2(p->q) == d^2(q->p) bit for bit, so one evaluation serves both directions; with a
//     wrap the two orders round differently ((L-|a'|)-|b'|) and both are formed;
//   * bins come from d^2 against thresholds thr[k] = min{t : sqrt_rn(t) >= edge_k}
//     (thr[n_bins] = min{t : sqrt_rn(t) > edge_last}) -- exactly np.histogram's rule on
//     r = sqrt_rn(d^2) without a device sqrt or division -- through a look-up table over
//     d^2 cells of width 2^-m (index and cell edges exact in binary floating point, at most
//     one threshold per cell);
//   * counts accumulate in warp-private shared histograms, flushed to global uint64 once per CTA.
#include "pbk_common.cuh"

#define RDF2_THREADS 256
#define RDF2_WARPS (RDF2_THREADS / 32)
#define RDF2_MAXCELLS 1024          // cells per leaflet (<= 32 x 32)
#define RDF2_MAXROWS 8              // my + 1
#define RDF2_QCAP 128               // ring queue entries per warp

struct Rdf2Params {
    const double *com;
    const uint8_t *labels;
    const int32_t *type_code;
    const float *box;
    int64_t F, N;
    int lat0, lat1, ta, tb, leaflet_mask, nb;
    double range_hi;
    float cut2_f;                   // (range_hi + margin)^2 for the float32 prefilter
    const double *thr;
    unsigned long long *count;
    int32_t *frame_counts;
    int cap;                        // capacity of the per-leaflet shared arrays
    int hist_copies;
    int lut_n;                      // entries of the d^2 -> bin look-up table
    double lut_scale;               // 2^m: lut index = floor(d2 * 2^m) - lut_c0
    int lut_c0;
};

// bin of d2 (or -1).  The table holds, per d^2 cell, the bin of the cell's lower edge; a cell
// contains at most one threshold, so the bin is that or the next one.
__device__ __forceinline__ int rdf2_bin(double d2, const double *s_thr, const unsigned char *s_lut,
                                        const Rdf2Params &P)
{
    if (!(d2 >= s_thr[0]) || !(d2 < s_thr[P.nb])) return -1;
    int c = __double2int_rd(d2 * P.lut_scale) - P.lut_c0;
    c = max(0, min(P.lut_n - 1, c));
    int k = s_lut[c];
    k += (d2 >= s_thr[k + 1]) ? 1 : 0;
    return k;
}

template <bool SAME>
__device__ __forceinline__ void rdf2_exact(unsigned entry, const double *s_bx, const double *s_by,
                                           const unsigned char *s_flag, const double *s_thr,
                                           const unsigned char *s_lut, unsigned *myhist,
                                           const Rdf2Params &P, double Lx, double Ly, double hx,
                                           double hy)
{
    const int p = entry & 0xFFFFu, q = entry >> 16;
    int w1 = 1, w2 = 1;
    if (!SAME) {
        const unsigned fp = s_flag[p], fq = s_flag[q];
        w1 = (fp & 1u) && (fq & 2u);            // ordered p -> q
        w2 = (fq & 1u) && (fp & 2u);            // ordered q -> p
        if (!(w1 | w2)) return;
    }
    const double apx = s_bx[p], apy = s_by[p], bqx = s_bx[q], bqy = s_by[q];
    double dx = __dsub_rn(apx, bqx), dy = __dsub_rn(apy, bqy);
    const bool wrx = fabs(dx) > hx, wry = fabs(dy) > hy;
    if (!(wrx | wry)) {
        const int k = rdf2_bin(fma(dy, dy, __dmul_rn(dx, dx)), s_thr, s_lut, P);
        if (k >= 0) atomicAdd(&myhist[k], (unsigned)(w1 + w2));
        return;
    }
    double dx2 = dx, dy2 = dy;                  // the q -> p evaluation
    if (wrx) {
        dx = __dsub_rn(__dsub_rn(Lx, fabs(apx)), fabs(bqx));
        dx2 = __dsub_rn(__dsub_rn(Lx, fabs(bqx)), fabs(apx));
    }
    if (wry) {
        dy = __dsub_rn(__dsub_rn(Ly, fabs(apy)), fabs(bqy));
        dy2 = __dsub_rn(__dsub_rn(Ly, fabs(bqy)), fabs(apy));
    }
    const double d2a = fma(dy, dy, __dmul_rn(dx, dx));
    const double d2b = fma(dy2, dy2, __dmul_rn(dx2, dx2));
    const int ka = rdf2_bin(d2a, s_thr, s_lut, P);
    int kb = ka;
    if (d2b != d2a) kb = rdf2_bin(d2b, s_thr, s_lut, P);
    if (ka == kb) {
        if (ka >= 0) atomicAdd(&myhist[ka], (unsigned)(w1 + w2));
    } else {
        if (w1 && ka >= 0) atomicAdd(&myhist[ka], 1u);
        if (w2 && kb >= 0) atomicAdd(&myhist[kb], 1u);
    }
}

template <bool SAME>
__global__ void __launch_bounds__(RDF2_THREADS)
k_rdf2(Rdf2Params P)
{
    extern __shared__ unsigned char smem_raw[];
    const int nb = P.nb, cap = P.cap;
    double *s_thr = reinterpret_cast<double *>(smem_raw);                   // nb+1
    double *s_bx = s_thr + (nb + 1);                                        // cap   b' = x - hx
    double *s_by = s_bx + cap;                                              // cap
    float2 *s_fxy = reinterpret_cast<float2 *>(s_by + cap);                 // cap   float32 copy of b'
    unsigned *s_hist = reinterpret_cast<unsigned *>(s_fxy + cap);           // copies*nb
    int *s_cell_start = reinterpret_cast<int *>(s_hist + (size_t)P.hist_copies * nb);  // MAXCELLS+1
    int *s_cell_fill = s_cell_start + (RDF2_MAXCELLS + 1);                  // MAXCELLS
    unsigned *s_queue = reinterpret_cast<unsigned *>(s_cell_fill + RDF2_MAXCELLS);     // WARPS*QCAP
    unsigned short *s_cellid = reinterpret_cast<unsigned short *>(s_queue + RDF2_WARPS * RDF2_QCAP);  // N
    unsigned short *s_scell = s_cellid + P.N;                               // cap (sorted): cell of slot
    unsigned char *s_flag = reinterpret_cast<unsigned char *>(s_scell + cap);    // cap (sorted)
    unsigned char *s_lut = s_flag + cap;                                    // lut_n
    __shared__ int s_cnt[3];
    __shared__ int s_scan[RDF2_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int q = tid; q <= nb; q += blockDim.x) s_thr[q] = P.thr[q];
    for (int q = tid; q < P.hist_copies * nb; q += blockDim.x) s_hist[q] = 0u;
    // look-up table: bin of the lower edge of every d^2 cell (edges c * 2^-m are exact)
    for (int c = tid; c < P.lut_n; c += blockDim.x) {
        const double t = (double)(c + P.lut_c0) / P.lut_scale;
        int k = 0;
        while (k < nb - 1 && t >= P.thr[k + 1]) k++;
        s_lut[c] = (unsigned char)k;
    }
    unsigned *myhist = s_hist + (size_t)(warp % P.hist_copies) * nb;
    unsigned *myq = s_queue + warp * RDF2_QCAP;
    const unsigned lt_mask = (1u << lane) - 1u;
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
        if (2 * my + 1 > ncy || my + 1 > RDF2_MAXROWS) { ncy = 1; my = 0; }
        double iwx = ncx / Lx, iwy = ncy / Ly;
        int ncell = ncx * ncy;
        // Cells and the float32 prefilter assume wrapped coordinates (0 <= x <= L).  If any
        // member lies outside the box the item is redone as a single cell with the prefilter
        // off, i.e. every pair goes through the exact path (reference formula for any input).
        bool all_pairs = false;
        for (;;) {
            if (tid < 3) s_cnt[tid] = 0;
            for (int c = tid; c < ncell; c += blockDim.x) s_cell_fill[c] = 0;
            __syncthreads();
            // pass 1: classify, count, bin into cells
            int cA = 0, cB = 0, cL = 0, oob = 0;
            for (int64_t i = tid; i < P.N; i += blockDim.x) {
                unsigned short cid = 0xFFFF;
                if (lab[i] == want) {
                    int t = P.type_code[i];
                    bool a = (P.ta == -1 || t == P.ta), b = (P.tb == -1 || t == P.tb);
                    cL++; cA += a; cB += b;
                    if (a || b) {
                        const double x = cf[i * 3 + P.lat0], y = cf[i * 3 + P.lat1];
                        oob |= !(x >= 0.0 && x <= Lx && y >= 0.0 && y <= Ly);
                        int cx = (int)(x * iwx), cy = (int)(y * iwy);
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
            const int any_oob = __syncthreads_or(oob);
            if (any_oob && !all_pairs) {
                all_pairs = true;
                ncx = ncy = 1; mx = my = 0; ncell = 1;
                iwx = 1.0 / Lx; iwy = 1.0 / Ly;
                continue;
            }
            break;
        }
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
            const double bx = __dsub_rn(cf[i * 3 + P.lat0], hx), by = __dsub_rn(cf[i * 3 + P.lat1], hy);
            s_bx[slot] = bx;
            s_by[slot] = by;
            s_fxy[slot] = make_float2((float)bx, (float)by);
            s_flag[slot] = (unsigned char)(((P.ta == -1 || t == P.ta) ? 1 : 0) | ((P.tb == -1 || t == P.tb) ? 2 : 0));
            s_scell[slot] = cid;
        }
        __syncthreads();

        // ---- pair sweep -------------------------------------------------------------------
        const int sub = tid & 7, grp = tid >> 3, ngrp = blockDim.x >> 3;
        const unsigned gshift = (unsigned)(lane & ~7);          // first lane of my group
        int qhead = 0, qtail = 0;                               // ring queue (warp uniform)
        for (int p0 = 0; p0 < n_u; p0 += ngrp) {
            const int p = p0 + grp;
            const bool pact = p < n_u;
            float fpx = 0.f, fpy = 0.f;
            // lane `sub` of the group prepares the slot ranges of half-shell row oy = sub
            int b1 = 0, n1 = 0, b2 = 0, n2 = 0;
            if (pact) {
                const float2 fp = s_fxy[p];
                fpx = fp.x; fpy = fp.y;
                if (sub <= my) {
                    const int pc = s_scell[p];
                    const int pcy = pc / ncx, pcx = pc - pcy * ncx;
                    int cy = pcy + sub; if (cy >= ncy) cy -= ncy;
                    const int row = cy * ncx;
                    int xlo = (sub == 0) ? pcx : pcx - mx, xhi = pcx + mx;
                    if (xlo < 0) {                      // wraps on the left
                        b2 = s_cell_start[row + xlo + ncx]; n2 = s_cell_start[row + ncx] - b2;
                        xlo = 0;
                    }
                    if (xhi >= ncx) {                   // wraps on the right
                        b2 = s_cell_start[row]; n2 = s_cell_start[row + xhi - ncx + 1] - b2;
                        xhi = ncx - 1;
                    }
                    b1 = s_cell_start[row + xlo]; n1 = s_cell_start[row + xhi + 1] - b1;
                    if (sub == 0) { n1 -= (p + 1 - b1); b1 = p + 1; }   // own cell: later entries only
                }
            }
            for (int oy = 0; oy <= my; oy++) {
                const int rb1 = __shfl_sync(0xffffffffu, b1, gshift + oy);
                const int rn1 = __shfl_sync(0xffffffffu, n1, gshift + oy);
                const int rb2 = __shfl_sync(0xffffffffu, b2, gshift + oy);
                const int rn2 = __shfl_sync(0xffffffffu, n2, gshift + oy);
                const int ntot = rn1 + rn2;
                // all lanes of the warp iterate together (max trip count over the 4 groups)
                int nmax = ntot;
                nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 8));
                nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 16));
                for (int t0 = 0; t0 < nmax; t0 += 8) {
                    const int t = t0 + sub;
                    bool hit = false;
                    int q = 0;
                    if (t < ntot) {
                        q = t < rn1 ? rb1 + t : rb2 + (t - rn1);
                        const float2 fq = s_fxy[q];
                        float dx = fabsf(fpx - fq.x), dy = fabsf(fpy - fq.y);
                        dx = fminf(dx, bxf - dx);
                        dy = fminf(dy, byf - dy);
                        hit = all_pairs || fmaf(dy, dy, dx * dx) < P.cut2_f;
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (hit) myq[(qtail + __popc(m & lt_mask)) & (RDF2_QCAP - 1)] = (unsigned)p | ((unsigned)q << 16);
                    qtail += __popc(m);
                    if (qtail - qhead >= 32) {
                        __syncwarp();
                        rdf2_exact<SAME>(myq[(qhead + lane) & (RDF2_QCAP - 1)], s_bx, s_by, s_flag, s_thr,
                                         s_lut, myhist, P, Lx, Ly, hx, hy);
                        qhead += 32;
                        __syncwarp();
                    }
                }
            }
        }
        __syncwarp();
        if (lane < qtail - qhead)
            rdf2_exact<SAME>(myq[(qhead + lane) & (RDF2_QCAP - 1)], s_bx, s_by, s_flag, s_thr, s_lut,
                             myhist, P, Lx, Ly, hx, hy);
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

// defined in pbk_rdf3.cu: warp-per-leaflet kernel for small systems
int pbk_rdf3_launch(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                    const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                    int leaflet_mask, int32_t n_bins, double range_lo, double range_hi, const double *thr,
                    unsigned long long *count, int32_t *frame_counts, cudaStream_t st);

// defined in pbk_rdf5.cu: CTA-per-leaflet kernel with a Verlet pair list over frame chunks
int pbk_rdf5_launch(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                    const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                    int leaflet_mask, int32_t n_bins, double range_lo, double range_hi, const double *thr,
                    unsigned long long *count, int32_t *frame_counts, cudaStream_t st);

// defined in pbk_cells.cu: global-memory cell lists for large leaflets
int pbk_rdf_cells(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                  const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                  int leaflet_mask, int32_t n_bins, double range_lo, double range_hi, const double *edges,
                  const double *thr, unsigned long long *count, int32_t *frame_counts, cudaStream_t st);
bool pbk_use_cells(int64_t N, int64_t threshold);

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
    if (thr && pbk_use_cells(N, 4096)) {
        const int rc = pbk_rdf_cells(ctx, com, labels, type_code, box, F, N, lat0, lat1, type_a, type_b,
                                     leaflet_mask, n_bins, range_lo, range_hi, edges, thr, count, frame_counts, st);
        if (rc != PBK_EUNSUPPORTED) return rc;
    }
    {
        const int rc = pbk_rdf5_launch(ctx, com, labels, type_code, box, F, N, lat0, lat1, type_a, type_b,
                                       leaflet_mask, n_bins, range_lo, range_hi, thr, count, frame_counts, st);
        if (rc != PBK_EUNSUPPORTED) return rc;
    }
    {
        const int rc = pbk_rdf3_launch(ctx, com, labels, type_code, box, F, N, lat0, lat1, type_a, type_b,
                                       leaflet_mask, n_bins, range_lo, range_hi, thr, count, frame_counts, st);
        if (rc != PBK_EUNSUPPORTED) return rc;
    }
    // d^2 look-up table: cells of width 2^-m no wider than the narrowest threshold spacing,
    // which for uniform bins of width w starting at lo >= 0 is (lo+w)^2 - lo^2 >= w^2
    const double w = (range_hi - range_lo) / n_bins;
    const double lo = range_lo > 0 ? range_lo : 0.0;
    const double min_gap = 0.999 * ((lo + w) * (lo + w) - lo * lo);
    int m = 0;
    while (ldexp(1.0, -m) > min_gap && m < 40) m++;
    while (m > -40 && ldexp(1.0, -(m - 1)) <= min_gap) m--;
    const double scale = ldexp(1.0, m);
    const double c0d = floor(lo * lo * scale) - 1.0, c1d = floor(range_hi * range_hi * scale) + 2.0;
    const double lut_nd = c1d - c0d + 1.0;
    int copies = RDF2_WARPS;
    const bool lut_ok = lut_nd <= 32768.0 && fabs(c0d) < 1e9;
    const int lut_n = lut_ok ? (int)lut_nd : 0;
    auto smem_for = [&](int c) {
        return (size_t)(n_bins + 1) * 8 + (size_t)N * 24 + (size_t)c * n_bins * 4 +
               (size_t)(2 * RDF2_MAXCELLS + 1) * 4 + (size_t)RDF2_WARPS * RDF2_QCAP * 4 +
               (size_t)N * 4 + (size_t)N + (size_t)lut_n + 16;
    };
    while (copies > 1 && smem_for(copies) > 64 * 1024) copies >>= 1;
    size_t smem = smem_for(copies);
    if (!thr || !lut_ok || smem > 200 * 1024 || N >= 65535 || n_bins > 255 || range_lo < 0)
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
    P.nb = n_bins; P.range_hi = range_hi;
    P.cut2_f = (float)((range_hi + 0.01) * (range_hi + 0.01) * 1.0001);
    P.thr = thr; P.count = count; P.frame_counts = frame_counts;
    P.cap = (int)N; P.hist_copies = copies;
    P.lut_n = lut_n; P.lut_scale = scale; P.lut_c0 = (int)c0d;
    kern<<<(unsigned)grid, RDF2_THREADS, smem, st>>>(P);
    PBK_CHECK_LAUNCH(ctx, "k_rdf2");
    return PBK_OK;
}
