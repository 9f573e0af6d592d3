// pbk_rdf3.cu -- COM lateral RDF for small leaflets (<= 640 lipids per frame): one WARP per
// (frame, leaflet), no block-wide barrier anywhere.
//
// Same arithmetic contract as k_rdf2 / k_rdf, restating
//   pybilt/bilayer_analyzer/analysis_protocols.py:4900-5006  (COMLateralRDFProtocol)
//   pybilt/common/distance_cutoff_clustering.py:27-71        (distance_euclidean_pbc)
//   numpy/lib/_histograms_impl.py                            (np.histogram on uniform bins)
// organised so that almost every pair costs ~25 float32 instructions and no divergence:
//   * the warp counting-sorts the leaflet's lipids of type A or B into a 2-D cell grid (<= 16 x 16
//     cells, edge ~ range_outer / 2) held in its private slice of shared memory;
//   * every lane owns one lipid p at a time and walks the forward half shell of cells, so each
//     UNORDERED pair {p, q} is met once;
//   * a pair's bin comes from its FLOAT32 minimum-image d^2 through a look-up table over d^2
//     cells of width 2^-m.  Each entry holds the bin of the whole cell and the distance (in
//     cells) from the cell to the nearest bin threshold; when that distance exceeds the
//     rounding bound of the float32 d^2 (computed per frame from the box) the float64 reference
//     arithmetic provably lands in the same bin, and the lane adds 1 or 2 to its warp's
//     histogram straight away;
//   * the few pairs closer to a threshold (~1 %) go to a per-warp queue and are evaluated with
//     the reference's exact float64 formulas from the original coordinates -- in both orders
//     when the minimum image wrapped, because (L-|a'|)-|b'| is not symmetric in rounding.
// Thresholds on d^2 (thr[k] = min{t : sqrt_rn(t) >= edge_k}) replace the square root exactly as
// in pbk_rdf2.cu.  Frames with a coordinate outside [0, L] go through the exact path entirely.
#include <stdlib.h>

#include "pbk_common.cuh"

#define R3_THREADS 256
#define R3_WARPS (R3_THREADS / 32)
#define R3_MAXDIM 16
#define R3_MAXCELLS (R3_MAXDIM * R3_MAXDIM)
#define R3_MAXROWS 8                // my + 1
#define R3_QCAP 64                  // exact-path queue entries per warp
#define R3_MAXN 640

struct Rdf3Params {
    const double *com;
    const uint8_t *labels;
    const int32_t *type_code;
    const float *box;
    int64_t F, N;
    int lat0, lat1, ta, tb, leaflet_mask, nb;
    double range_hi;
    const double *thr;
    unsigned long long *count;
    int32_t *frame_counts;
    int lut_n;                      // entries of the d^2 look-up table
    int lut_m;                      // cell width 2^-m
    int *work;                      // item counter (zeroed before the launch)
};

// exact bin of d2 against the thresholds (or -1): thr[k] <= d2 < thr[k+1]
__device__ __forceinline__ int r3_bin_exact(double d2, const double *s_thr, int nb)
{
    if (!(d2 >= s_thr[0]) || !(d2 < s_thr[nb])) return -1;
    int lo = 0, hi = nb;            // invariant: thr[lo] <= d2 < thr[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (d2 >= s_thr[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

// the reference's float64 arithmetic for the unordered pair of sorted slots (p, q)
template <bool SAME>
__device__ __noinline__ void r3_exact(unsigned entry, const double *cf, int lat0, int lat1,
                                      const unsigned short *s_idx, const unsigned char *s_flag,
                                      const double *s_thr, int nb, unsigned *myhist, double Lx, double Ly,
                                      double hx, double hy)
{
    const int p = entry & 0xFFFFu, q = entry >> 16;
    int w1 = 1, w2 = 1;
    if (!SAME) {
        const unsigned fp = s_flag[p], fq = s_flag[q];
        w1 = (fp & 1u) && (fq & 2u);            // ordered p -> q
        w2 = (fq & 1u) && (fp & 2u);            // ordered q -> p
        if (!(w1 | w2)) return;
    }
    const int ip = s_idx[p], iq = s_idx[q];
    const double apx = __dsub_rn(cf[ip * 3 + lat0], hx), apy = __dsub_rn(cf[ip * 3 + lat1], hy);
    const double bqx = __dsub_rn(cf[iq * 3 + lat0], hx), bqy = __dsub_rn(cf[iq * 3 + lat1], hy);
    double dx = __dsub_rn(apx, bqx), dy = __dsub_rn(apy, bqy);
    const bool wrx = fabs(dx) > hx, wry = fabs(dy) > hy;
    if (!(wrx | wry)) {
        const int k = r3_bin_exact(fma(dy, dy, __dmul_rn(dx, dx)), s_thr, nb);
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
    const int ka = r3_bin_exact(d2a, s_thr, nb);
    int kb = ka;
    if (d2b != d2a) kb = r3_bin_exact(d2b, s_thr, nb);
    if (ka == kb) {
        if (ka >= 0) atomicAdd(&myhist[ka], (unsigned)(w1 + w2));
    } else {
        if (w1 && ka >= 0) atomicAdd(&myhist[ka], 1u);
        if (w2 && kb >= 0) atomicAdd(&myhist[kb], 1u);
    }
}

template <bool SAME>
__global__ void __launch_bounds__(R3_THREADS)
k_rdf3(Rdf3Params P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nb = P.nb, N = (int)P.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // CTA-shared: thresholds, look-up table
    double *s_thr = reinterpret_cast<double *>(smem_raw);                                   // nb+1
    unsigned short *s_lut = reinterpret_cast<unsigned short *>(s_thr + (nb + 1));           // lut_n
    // warp-private slices
    const size_t lut_bytes = ((size_t)P.lut_n * 2 + 15) & ~(size_t)15;
    const size_t per_warp = (size_t)N * 8 + (size_t)nb * 4 + (size_t)R3_QCAP * 4 + (size_t)(2 * R3_MAXCELLS + 2) * 4 +
                            (size_t)N * 2 * 2 + (((size_t)N + 15) & ~(size_t)15);
    unsigned char *wbase = reinterpret_cast<unsigned char *>(s_lut) + lut_bytes + (size_t)warp * ((per_warp + 15) & ~(size_t)15);
    float2 *s_fxy = reinterpret_cast<float2 *>(wbase);                                       // N  (sorted) float32 b'
    unsigned *myhist = reinterpret_cast<unsigned *>(s_fxy + N);                              // nb
    unsigned *myq = myhist + nb;                                                             // QCAP
    int *s_cell_start = reinterpret_cast<int *>(myq + R3_QCAP);                              // MAXCELLS+1
    int *s_cell_fill = s_cell_start + (R3_MAXCELLS + 1);                                     // MAXCELLS+1
    unsigned short *s_idx = reinterpret_cast<unsigned short *>(s_cell_fill + (R3_MAXCELLS + 1));  // N (sorted): lipid
    unsigned short *s_cid = s_idx + N;                                                       // N (by lipid): cell or 0xFFFF
    unsigned char *s_flag = reinterpret_cast<unsigned char *>(s_cid + N);                    // N (sorted): bit0 A, bit1 B

    for (int q = tid; q <= nb; q += R3_THREADS) s_thr[q] = P.thr[q];
    for (int q = lane; q < nb; q += 32) myhist[q] = 0u;
    __syncthreads();
    // look-up table over d^2 cells [c, c+1) * 2^-m: low byte = bin of the cell (0xFF outside the
    // range), high byte = distance in cells to the nearest threshold (saturated at 255; 0 when a
    // threshold lies inside the cell)
    {
        const double inv = ldexp(1.0, -P.lut_m), sc = ldexp(1.0, P.lut_m);
        for (int c = tid; c < P.lut_n; c += R3_THREADS) {
            const double lo = (double)c * inv, hi = (double)(c + 1) * inv;
            int k = -1;                         // thr[k] <= lo < thr[k+1]
            if (lo >= s_thr[0]) {
                int a = 0, b = nb + 1;          // thr[a] <= lo, (b == nb+1 or lo < thr[b])
                while (b - a > 1) { const int mid = (a + b) >> 1; if (lo >= s_thr[mid]) a = mid; else b = mid; }
                k = a;
            }
            double dist = 1e300;
            if (k >= 0) dist = fmin(dist, lo - s_thr[k]);
            if (k + 1 <= nb) dist = fmin(dist, s_thr[k + 1] - hi);
            int dc = dist <= 0.0 ? 0 : (int)fmin(255.0, floor(dist * sc));
            if (c == P.lut_n - 1) { k = nb; }   // clamp cell: everything beyond the table is out of range
            const int bin = (k >= 0 && k < nb) ? k : 0xFF;
            s_lut[c] = (unsigned short)(bin | (dc << 8));
        }
    }
    __syncthreads();
    const unsigned lt_mask = (1u << lane) - 1u;
    const float lut_scale = ldexpf(1.0f, P.lut_m);
    const int lut_last = P.lut_n - 1;
    const int64_t n_items = P.F * 2;

    for (;;) {
        long long item = 0;
        if (lane == 0) item = (long long)atomicAdd(P.work, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int64_t f = item >> 1;
        const int leaf = (int)(item & 1);            // 0 = upper (label 1), 1 = lower (label 0)
        const unsigned want = leaf == 0 ? 1u : 0u;
        const bool selected = (P.leaflet_mask >> leaf) & 1;
        const double *cf = P.com + f * P.N * 3;
        const uint8_t *lab = P.labels + f * P.N;
        const float bxf = __ldg(P.box + f * 3 + P.lat0), byf = __ldg(P.box + f * 3 + P.lat1);
        const float cxf = bxf * 0.5f, cyf = byf * 0.5f;
        const double Lx = (double)bxf, Ly = (double)byf, hx = (double)cxf, hy = (double)cyf;
        // cell grid: edge >= 0.51*range_hi so that the half shell reaches 2 cells
        const double wmin = 0.51 * P.range_hi + 1e-6;
        int ncx = (int)floor(Lx / wmin), ncy = (int)floor(Ly / wmin);
        ncx = max(1, min(R3_MAXDIM, ncx));
        ncy = max(1, min(R3_MAXDIM, ncy));
        const double wx = Lx / ncx, wy = Ly / ncy;
        int mx = (int)floor((P.range_hi + 1e-6) / wx) + 1, my = (int)floor((P.range_hi + 1e-6) / wy) + 1;
        if (2 * mx + 1 > ncx) { ncx = 1; mx = 0; }
        if (2 * my + 1 > ncy || my + 1 > R3_MAXROWS) { ncy = 1; my = 0; }
        double iwx = ncx / Lx, iwy = ncy / Ly;
        int ncell = ncx * ncy;
        // rounding bound of the float32 d^2 near the thresholds, in table cells (+1 cell of slack):
        // |d2_f32 - d2_exact| <= 2^-23 (r+1) (8 L + 4 (r+1))  with r = range_hi, L = larger box edge
        const float Lm = fmaxf(bxf, byf), rr = (float)P.range_hi + 1.0f;
        const float tol = 1.1920929e-7f * rr * (8.0f * Lm + 4.0f * rr);
        int tolc = (int)ceilf(tol * lut_scale) + 1;
        bool all_exact = !(tol * lut_scale < 200.0f) || !(bxf > 0.0f) || !(byf > 0.0f);

        // ---- pass 1: classify, count, bin into cells ------------------------------------------
        int cA = 0, cB = 0, cL = 0, oob = 0;
        for (int c = lane; c <= ncell; c += 32) s_cell_fill[c] = 0;
        __syncwarp();
        for (int i = lane; i < N; i += 32) {
            unsigned short cid = 0xFFFF;
            if (lab[i] == want) {
                const int t = __ldg(P.type_code + i);
                const bool a = (P.ta == -1 || t == P.ta), b = (P.tb == -1 || t == P.tb);
                cL++; cA += a; cB += b;
                if (a || b) {
                    const double x = cf[i * 3 + P.lat0], y = cf[i * 3 + P.lat1];
                    oob |= !(x >= 0.0 && x <= Lx && y >= 0.0 && y <= Ly);
                    int cx = (int)(x * iwx), cy = (int)(y * iwy);
                    cx = max(0, min(ncx - 1, cx));
                    cy = max(0, min(ncy - 1, cy));
                    cid = (unsigned short)(cy * ncx + cx);
                }
            }
            s_cid[i] = cid;
        }
        cA = pbk_warp_sum(cA); cB = pbk_warp_sum(cB); cL = pbk_warp_sum(cL);
        if (__any_sync(0xffffffffu, oob)) {
            // a member outside the box: one cell, every pair through the exact formulas
            all_exact = true;
            ncx = ncy = 1; mx = my = 0; ncell = 1;
            for (int i = lane; i < N; i += 32) if (s_cid[i] != 0xFFFF) s_cid[i] = 0;
        }
        const bool on = selected && cA > 0 && cB > 0;
        if (lane < 3) {
            int v = lane == 0 ? cA : (lane == 1 ? cB : cL);
            if (lane < 2 && !on) v = 0;
            if (lane == 2 && !selected) v = 0;
            P.frame_counts[f * 6 + leaf * 3 + lane] = v;
        }
        if (!on) continue;
        __syncwarp();
        for (int i = lane; i < N; i += 32) {
            const unsigned short cid = s_cid[i];
            if (cid != 0xFFFF) atomicAdd(&s_cell_fill[cid], 1);
        }
        __syncwarp();
        // exclusive scan of the cell counts (<= 256 cells: 8 per lane)
        {
            int v[8], s = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) { const int c = lane * 8 + q; v[q] = c < ncell ? s_cell_fill[c] : 0; s += v[q]; }
            int incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            int excl = incl - s;
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int c = lane * 8 + q;
                if (c < ncell) { s_cell_start[c] = excl; s_cell_fill[c] = excl; }
                excl += v[q];
            }
            if (lane == 31) s_cell_start[ncell] = excl;
        }
        __syncwarp();
        const int n_u = s_cell_start[ncell];
        // ---- pass 2: scatter into cell order -----------------------------------------------------
        for (int i = lane; i < N; i += 32) {
            const unsigned short cid = s_cid[i];
            if (cid == 0xFFFF) continue;
            const int slot = atomicAdd(&s_cell_fill[cid], 1);
            const double bx = __dsub_rn(cf[i * 3 + P.lat0], hx), by = __dsub_rn(cf[i * 3 + P.lat1], hy);
            s_fxy[slot] = make_float2((float)bx, (float)by);
            s_idx[slot] = (unsigned short)i;
            if (!SAME) {
                const int t = __ldg(P.type_code + i);
                s_flag[slot] = (unsigned char)(((P.ta == -1 || t == P.ta) ? 1 : 0) | ((P.tb == -1 || t == P.tb) ? 2 : 0));
            }
        }
        __syncwarp();

        // ---- pair sweep: lane = lipid p, forward half shell of cells ---------------------------
        int qn = 0;                                            // queued exact pairs (warp uniform)
        for (int p0 = 0; p0 < n_u; p0 += 32) {
            const int p = p0 + lane;
            const bool pact = p < n_u;
            float fpx = 0.f, fpy = 0.f;
            int pcx = 0, pcy = 0;
            unsigned fpflag = 3u;
            if (pact) {
                const float2 fp = s_fxy[p];
                fpx = fp.x; fpy = fp.y;
                const int pc = s_cid[s_idx[p]];
                pcy = pc / ncx; pcx = pc - pcy * ncx;
                if (!SAME) fpflag = s_flag[p];
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
                for (int t = 0; __any_sync(0xffffffffu, t < ntot); t++) {
                    const bool act = t < ntot;
                    const int q = act ? (t < n1 ? b1 + t : b2 + (t - n1)) : 0;
                    const float2 fq = s_fxy[q];
                    float dx = fabsf(fpx - fq.x), dy = fabsf(fpy - fq.y);
                    dx = fminf(dx, bxf - dx);
                    dy = fminf(dy, byf - dy);
                    const float d2 = fmaf(dy, dy, dx * dx);
                    const int c = min(__float2int_rd(d2 * lut_scale), lut_last);
                    const unsigned e = s_lut[c];
                    unsigned w = 2u;
                    if (!SAME) {
                        const unsigned fqf = s_flag[q];
                        w = ((fpflag & 1u) & ((fqf >> 1) & 1u)) + ((fqf & 1u) & ((fpflag >> 1) & 1u));
                    }
                    const bool sure = (int)(e >> 8) >= tolc && !all_exact;
                    if (act && sure && (e & 0xFFu) != 0xFFu && w) atomicAdd(&myhist[e & 0xFFu], w);
                    const bool amb = act && !sure && w;
                    const unsigned m = __ballot_sync(0xffffffffu, amb);
                    if (m) {                            // rare: queue for the exact float64 evaluation
                        if (amb) myq[qn + __popc(m & lt_mask)] = (unsigned)p | ((unsigned)q << 16);
                        qn += __popc(m);
                        __syncwarp();
                        if (qn >= 32) {
                            r3_exact<SAME>(myq[lane], cf, P.lat0, P.lat1, s_idx, s_flag, s_thr, nb, myhist, Lx, Ly, hx, hy);
                            __syncwarp();
                            const unsigned moved = lane + 32 < qn ? myq[lane + 32] : 0u;
                            __syncwarp();
                            if (lane + 32 < qn) myq[lane] = moved;
                            qn -= 32;
                            __syncwarp();
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane < qn)
            r3_exact<SAME>(myq[lane], cf, P.lat0, P.lat1, s_idx, s_flag, s_thr, nb, myhist, Lx, Ly, hx, hy);
        __syncwarp();
    }
    __syncwarp();
    for (int q = lane; q < nb; q += 32) {
        const unsigned t = myhist[q];
        if (t) atomicAdd(&P.count[q], (unsigned long long)t);
    }
}

// shared-memory bytes of k_rdf3 (0 if the shape is not taken)
static size_t r3_smem(int64_t N, int nb, int lut_n)
{
    const size_t lut_bytes = ((size_t)lut_n * 2 + 15) & ~(size_t)15;
    const size_t per_warp = (size_t)N * 8 + (size_t)nb * 4 + (size_t)R3_QCAP * 4 + (size_t)(2 * R3_MAXCELLS + 2) * 4 +
                            (size_t)N * 2 * 2 + (((size_t)N + 15) & ~(size_t)15);
    return (size_t)(nb + 1) * 8 + lut_bytes + (size_t)R3_WARPS * ((per_warp + 15) & ~(size_t)15) + 16;
}

// Returns PBK_EUNSUPPORTED when the shape belongs to the block-per-leaflet kernel (pbk_rdf2.cu).
int pbk_rdf3_launch(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                    const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                    int leaflet_mask, int32_t n_bins, double range_lo, double range_hi, const double *thr,
                    unsigned long long *count, int32_t *frame_counts, cudaStream_t st)
{
    if (!thr || N > R3_MAXN || n_bins > 254 || range_lo < 0 || !ctx->work || getenv("PBK_RDF_V2")) return PBK_EUNSUPPORTED;
    // d^2 table: cells of width 2^-m, at least 16 per narrowest threshold gap (w^2 for uniform bins of
    // width w starting at lo >= 0), capped at 16384 entries
    const double w = (range_hi - range_lo) / n_bins;
    const double lo = range_lo > 0 ? range_lo : 0.0;
    const double min_gap = (lo + w) * (lo + w) - lo * lo;
    int m = 0;
    while (ldexp(1.0, -m) > min_gap / 16.0 && m < 30) m++;
    while (m > -20 && (range_hi * range_hi + 2.0) * ldexp(1.0, m) + 520.0 > 16384.0) m--;
    const double lut_nd = floor((range_hi * range_hi + 1.0) * ldexp(1.0, m)) + 516.0;
    // (a cell wider than a threshold gap is still handled correctly -- it is marked ambiguous)
    if (lut_nd > 16384.0 || lut_nd < 8.0) return PBK_EUNSUPPORTED;
    const int lut_n = (int)lut_nd;
    const size_t smem = r3_smem(N, n_bins, lut_n);
    if (smem > 110 * 1024) return PBK_EUNSUPPORTED;
    const bool same = (type_a == type_b);
    auto kern = same ? k_rdf3<true> : k_rdf3<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "k_rdf3 smem attribute", e);
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, R3_THREADS, smem);
    if (e != cudaSuccess || per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)ctx->sm_count * per_sm;
    const int64_t need = (F * 2 + R3_WARPS - 1) / R3_WARPS;
    if (grid > need) grid = need;
    int *work = ctx->work + (ctx->work_next++ % PBK_WORK_SLOTS);    // launches in flight do not share a counter
    e = cudaMemsetAsync(work, 0, sizeof(int), st);
    if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "k_rdf3 work counter", e);
    Rdf3Params P;
    P.com = com; P.labels = labels; P.type_code = type_code; P.box = box; P.F = F; P.N = N;
    P.lat0 = lat0; P.lat1 = lat1; P.ta = type_a; P.tb = type_b; P.leaflet_mask = leaflet_mask;
    P.nb = n_bins; P.range_hi = range_hi; P.thr = thr; P.count = count; P.frame_counts = frame_counts;
    P.lut_n = lut_n; P.lut_m = m; P.work = work;
    kern<<<(unsigned)grid, R3_THREADS, smem, st>>>(P);
    PBK_CHECK_LAUNCH(ctx, "k_rdf3");
    return PBK_OK;
}
