// pbk_rdf5.cu -- COM lateral RDF for small leaflets: one pair list per leaflet and frame chunk.
//
// Same arithmetic contract as k_rdf3 / k_rdf2 / k_rdf, restating
//   pybilt/bilayer_analyzer/analysis_protocols.py:4900-5006  (COMLateralRDFProtocol)
//   pybilt/common/distance_cutoff_clustering.py:27-71        (distance_euclidean_pbc)
//   numpy/lib/_histograms_impl.py                            (np.histogram on uniform bins)
// One 256-thread CTA owns one leaflet over a CHUNK of consecutive frames (about 30 KB of shared
// memory, four CTAs per SM).  Lipids move a fraction of an Angstrom per frame, so the cell sort
// and the half-shell sweep -- two thirds of the per-frame kernels' instructions -- are done
// once per chunk instead of once per frame:
//   scan    the CTA reads the chunk's frames once and records, per lipid, the bounding box of
//           its (minimum-image) positions over the run of frames with an identical member set
//           and a bit-identical box;
//   build   lipids are cell-sorted by box centre; a pair enters the list iff the gap between
//           the two bounding boxes is within range_outer (+ margin), i.e. iff the pair can be
//           within range in SOME frame of the run.  The half-shell sweep covers lipids whose
//           half-extent is <= 3 A; the few with larger boxes -- lipids whose WRAPPED centre of
//           mass jumps because a bead crossed the box edge (com_frame.py:71 averages the raw
//           wrapped atoms) -- are paired with every member directly;
//   walk    per frame: refresh the float32 coordinates, walk the flat list (two 8-bit slots
//           per pair), four independent pairs per thread per iteration.
// A pair's bin comes from its FLOAT32 minimum-image d^2 through a look-up table over d^2 cells
// of width 2^-m; each entry holds the bin of the whole cell and the distance (in cells) to the
// nearest bin threshold.  When that distance exceeds the rounding bound of the float32 d^2 the
// float64 reference arithmetic provably lands in the same bin; the ~1 % of pairs that are
// closer to a threshold are queued and evaluated with the reference's exact float64 formulas
// from the original coordinates, in both orders when the minimum image wrapped ((L-|a'|)-|b'|
// is not symmetric in rounding).  Thresholds on d^2 (thr[k] = min{t : sqrt_rn(t) >= edge_k})
// replace the square root exactly as in pbk_rdf2.cu.
// Frames that cannot join a run (member outside the box, more than 256 members, list overflow)
// are binned one by one from all pairs of members with the same fast path / exact path split.
#include <stdlib.h>

#include "pbk_common.cuh"

#define R5_THREADS 256
#define R5_WARPS (R5_THREADS / 32)
#define R5_MAXDIM 16
#define R5_MAXCELLS (R5_MAXDIM * R5_MAXDIM)
#define R5_MAXROWS 8
#define R5_QCAP 64
#define R5_MAXN 512
#define R5_LPT ((R5_MAXN + R5_THREADS - 1) / R5_THREADS)   // lipids per thread in the scan
#define R5_MAXW 64                  // lipids with large boxes per run

struct Rdf5Params {
    const double *com;
    const uint8_t *labels;
    const int32_t *type_code;
    const float *box;
    int64_t F, N;
    int lat0, lat1, ta, tb, leaflet_mask, nb;
    double range_hi;
    float ecap;                     // half-extent up to which a lipid takes part in the cell sweep
    const double *thr;
    unsigned long long *count;
    int32_t *frame_counts;
    int lut_n, lut_m;
    int chunk;                      // frames per work item
    int list_cap;                   // pair-list entries
    int *work;
    const unsigned short *lut;      // [lut_n] global (L1-resident): bin | distance-to-threshold << 8
};

struct R5Sh {                       // the CTA's shared memory
    float2 *fxy;                    // N (sorted): float32 b' = x - box/2 of the current frame / box centres
    float2 *ext;                    // N (sorted): half-extents of the bounding boxes
    unsigned short *list;           // list_cap: p | q << 8
    unsigned *hist;                 // this warp's nb counters
    unsigned *q;                    // this warp's R5_QCAP exact-path queue
    int *cell_start, *cell_fill;    // R5_MAXCELLS + 1 each
    unsigned short *idx;            // N (sorted): lipid of the slot
    unsigned short *cid;            // N (by lipid): cell, 0xFFFF = not a member
    unsigned char *flag;            // N (sorted): bit0 type A, bit1 type B, bit2 large box
    unsigned short *wslot;          // R5_MAXW slots with large boxes
    int *cnt;                       // 8 counters: cA, cB, cL, n_wild, n_list, item, -, -
};

struct R5Frame {                    // per-frame constants
    const double *cf;
    float bxf, byf;
    double Lx, Ly, hx, hy;
    int tolc;
    bool all_exact;
};

__device__ __forceinline__ void r5_frame(R5Frame &Fr, const Rdf5Params &P, int64_t f, float lut_scale)
{
    Fr.cf = P.com + f * P.N * 3;
    Fr.bxf = __ldg(P.box + f * 3 + P.lat0); Fr.byf = __ldg(P.box + f * 3 + P.lat1);
    const float cxf = Fr.bxf * 0.5f, cyf = Fr.byf * 0.5f;
    Fr.Lx = (double)Fr.bxf; Fr.Ly = (double)Fr.byf; Fr.hx = (double)cxf; Fr.hy = (double)cyf;
    // rounding bound of the float32 d^2 near the thresholds, in table cells (+1 cell of slack):
    // |d2_f32 - d2_exact| <= 2^-23 (r+1) (8 L + 4 (r+1))  with r = range_hi, L = larger box edge
    const float Lm = fmaxf(Fr.bxf, Fr.byf), rr = (float)P.range_hi + 1.0f;
    const float tol = 1.1920929e-7f * rr * (8.0f * Lm + 4.0f * rr);
    Fr.tolc = (int)ceilf(tol * lut_scale) + 1;
    Fr.all_exact = !(tol * lut_scale < 200.0f) || !(Fr.bxf > 0.0f) || !(Fr.byf > 0.0f);
}

__device__ __forceinline__ int r5_bin_exact(double d2, const double *s_thr, int nb)
{
    if (!(d2 >= s_thr[0]) || !(d2 < s_thr[nb])) return -1;
    int lo = 0, hi = nb;            // invariant: thr[lo] <= d2 < thr[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (d2 >= s_thr[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

// the reference's float64 arithmetic for the unordered pair of sorted slots (p, q) of frame cf
template <bool SAME>
__device__ __noinline__ void r5_exact(int p, int q, const double *cf, int lat0, int lat1,
                                      const unsigned short *s_idx, const unsigned char *s_flag,
                                      const double *s_thr, int nb, unsigned *myhist, double Lx, double Ly,
                                      double hx, double hy)
{
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
        const int k = r5_bin_exact(fma(dy, dy, __dmul_rn(dx, dx)), s_thr, nb);
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
    const int ka = r5_bin_exact(d2a, s_thr, nb);
    int kb = ka;
    if (d2b != d2a) kb = r5_bin_exact(d2b, s_thr, nb);
    if (ka == kb) {
        if (ka >= 0) atomicAdd(&myhist[ka], (unsigned)(w1 + w2));
    } else {
        if (w1 && ka >= 0) atomicAdd(&myhist[ka], 1u);
        if (w2 && kb >= 0) atomicAdd(&myhist[kb], 1u);
    }
}

// Exact-path queue of a warp.  Entries: p | q << 10 | frame-in-run << 20.  `base` = first frame
// of the run; box constants are those of the run (bit-identical in all its frames).
template <bool SAME>
__device__ __forceinline__ void r5_drain(int n, int lane, const R5Sh &W, const R5Frame &Fr, const Rdf5Params &P,
                                         const double *s_thr, int64_t base)
{
    if (lane < n) {
        const unsigned e = W.q[lane];
        r5_exact<SAME>((int)(e & 0x3FFu), (int)((e >> 10) & 0x3FFu), P.com + (base + (e >> 20)) * P.N * 3, P.lat0, P.lat1,
                       W.idx, W.flag, s_thr, P.nb, W.hist, Fr.Lx, Fr.Ly, Fr.hx, Fr.hy);
    }
}

// queue the pairs flagged `amb` (all 32 lanes call this together)
template <bool SAME>
__device__ __forceinline__ void r5_push(bool amb, int p, int q, unsigned df, int &qn, int lane, const R5Sh &W,
                                        const R5Frame &Fr, const Rdf5Params &P, const double *s_thr, int64_t base)
{
    const unsigned m = __ballot_sync(0xffffffffu, amb);
    if (!m) return;
    if (amb) W.q[qn + __popc(m & ((1u << lane) - 1u))] = (unsigned)p | ((unsigned)q << 10) | (df << 20);
    qn += __popc(m);
    __syncwarp();
    if (qn >= 32) {
        r5_drain<SAME>(32, lane, W, Fr, P, s_thr, base);
        __syncwarp();
        const unsigned moved = lane + 32 < qn ? W.q[lane + 32] : 0u;
        __syncwarp();
        if (lane + 32 < qn) W.q[lane] = moved;
        qn -= 32;
        __syncwarp();
    }
}

// float32 minimum-image d^2 of two slots
__device__ __forceinline__ float r5_d2(float2 fp, float2 fq, float bxf, float byf)
{
    float dx = fabsf(fp.x - fq.x), dy = fabsf(fp.y - fq.y);
    dx = fminf(dx, bxf - dx);
    dy = fminf(dy, byf - dy);
    return fmaf(dy, dy, dx * dx);
}

// smallest possible squared distance between the bounding boxes (centre c, half-extent e)
__device__ __forceinline__ float r5_gap2(float2 cp, float2 ep, float2 cq, float2 eq, float bxf, float byf)
{
    float dx = fabsf(cp.x - cq.x), dy = fabsf(cp.y - cq.y);
    dx = fminf(dx, bxf - dx);
    dy = fminf(dy, byf - dy);
    dx = fmaxf(0.0f, dx - (ep.x + eq.x));
    dy = fmaxf(0.0f, dy - (ep.y + eq.y));
    return fmaf(dy, dy, dx * dx);
}

// look-up table over d^2 cells [c, c+1) * 2^-m: low byte = bin of the cell (0xFF outside the
// range), high byte = distance in cells to the nearest threshold (saturated at 255; 0 when a
// threshold lies inside the cell)
__global__ void k_rdf5_lut(const double *__restrict__ thr, int nb, int lut_n, int lut_m, unsigned short *__restrict__ lut)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= lut_n) return;
    const double inv = ldexp(1.0, -lut_m), sc = ldexp(1.0, lut_m);
    const double lo = (double)c * inv, hi = (double)(c + 1) * inv;
    int k = -1;                         // thr[k] <= lo < thr[k+1]
    if (lo >= thr[0]) {
        int a = 0, b = nb + 1;
        while (b - a > 1) { const int mid = (a + b) >> 1; if (lo >= thr[mid]) a = mid; else b = mid; }
        k = a;
    }
    double dist = 1e300;
    if (k >= 0) dist = fmin(dist, lo - thr[k]);
    if (k + 1 <= nb) dist = fmin(dist, thr[k + 1] - hi);
    const int dc = dist <= 0.0 ? 0 : (int)fmin(255.0, floor(dist * sc));
    if (c == lut_n - 1) k = nb;         // clamp cell: everything beyond the table is out of range
    const int bin = (k >= 0 && k < nb) ? k : 0xFF;
    lut[c] = (unsigned short)(bin | (dc << 8));
}

// one candidate pair: count it, or queue it for the exact path
template <bool SAME>
__device__ __forceinline__ void r5_pair(bool act, int p, int q, float2 fp, float2 fq, unsigned fpflag, unsigned df, int &qn,
                                        int lane, const R5Sh &W, const R5Frame &Fr, const Rdf5Params &P,
                                        const double *s_thr, float lut_scale, int lut_last, int64_t base)
{
    const float d2 = r5_d2(fp, fq, Fr.bxf, Fr.byf);
    const unsigned e = __ldg(P.lut + min(__float2int_rd(d2 * lut_scale), lut_last));
    unsigned w = 2u;
    if (!SAME) {
        const unsigned fqf = W.flag[q];
        w = ((fpflag & 1u) & ((fqf >> 1) & 1u)) + ((fqf & 1u) & ((fpflag >> 1) & 1u));
    }
    const bool sure = (int)(e >> 8) >= Fr.tolc && !Fr.all_exact;
    if (act && sure && (e & 0xFFu) != 0xFFu && w) atomicAdd(&W.hist[e & 0xFFu], w);
    r5_push<SAME>(act && !sure && w, p, q, df, qn, lane, W, Fr, P, s_thr, base);
}

template <bool SAME>
__global__ void __launch_bounds__(R5_THREADS, 4)
k_rdf5(Rdf5Params P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nb = P.nb, N = (int)P.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t npad = ((size_t)N + 15) & ~(size_t)15;
    unsigned char *wb = smem_raw;
    double *s_thr = reinterpret_cast<double *>(wb);                        wb += (((size_t)(nb + 1) * 8) + 15) & ~(size_t)15;
    R5Sh W;
    W.fxy = reinterpret_cast<float2 *>(wb);                                wb += npad * 8;
    W.ext = reinterpret_cast<float2 *>(wb);                                wb += npad * 8;
    W.list = reinterpret_cast<unsigned short *>(wb);                       wb += ((size_t)P.list_cap * 2 + 15) & ~(size_t)15;
    unsigned *hist_all = reinterpret_cast<unsigned *>(wb);                 wb += (((size_t)R5_WARPS * nb * 4) + 15) & ~(size_t)15;
    W.hist = hist_all + (size_t)warp * nb;
    W.q = reinterpret_cast<unsigned *>(wb) + (size_t)warp * R5_QCAP;       wb += (size_t)R5_WARPS * R5_QCAP * 4;
    W.cell_start = reinterpret_cast<int *>(wb);                            wb += (size_t)(R5_MAXCELLS + 1) * 4;
    W.cell_fill = reinterpret_cast<int *>(wb);                             wb += (size_t)(R5_MAXCELLS + 1) * 4;
    W.cnt = reinterpret_cast<int *>(wb);                                   wb += 8 * 4;
    W.idx = reinterpret_cast<unsigned short *>(wb);                        wb += npad * 2;
    W.cid = reinterpret_cast<unsigned short *>(wb);                        wb += npad * 2;
    W.wslot = reinterpret_cast<unsigned short *>(wb);                      wb += (size_t)R5_MAXW * 2;
    W.flag = reinterpret_cast<unsigned char *>(wb);

    for (int q = tid; q <= nb; q += R5_THREADS) s_thr[q] = P.thr[q];
    for (int q = tid; q < R5_WARPS * nb; q += R5_THREADS) hist_all[q] = 0u;
    __syncthreads();
    const unsigned lt_mask = (1u << lane) - 1u;
    const float lut_scale = ldexpf(1.0f, P.lut_m);
    const int lut_last = P.lut_n - 1;
    const int64_t chunks = (P.F + P.chunk - 1) / P.chunk;
    const int64_t n_items = chunks * 2;
    const float Rc = (float)P.range_hi + 0.05f;                   // pair criterion on the box gap
    const float Rc2 = Rc * Rc;
    const float Rb = Rc + 2.0f * P.ecap;                          // reach of the cell sweep (box centres)

    for (;;) {
        __syncthreads();
        if (tid == 0) W.cnt[5] = atomicAdd(P.work, 1);
        __syncthreads();
        const long long item = W.cnt[5];
        if (item >= n_items) break;
        const int leaf = (int)(item & 1);            // 0 = upper (label 1), 1 = lower (label 0)
        const int64_t fa = (item >> 1) * P.chunk, fb = min(P.F, fa + (int64_t)P.chunk);
        const unsigned want = leaf == 0 ? 1u : 0u;
        const bool selected = (P.leaflet_mask >> leaf) & 1;

        int64_t f0 = fa;
        while (f0 < fb) {
            // ================= scan: the run [f0, fe) of frames that share member set and box ========
            R5Frame Fr;
            r5_frame(Fr, P, f0, lut_scale);
            float c0x[R5_LPT], c0y[R5_LPT], lox[R5_LPT], hix[R5_LPT], loy[R5_LPT], hiy[R5_LPT];
            bool in0[R5_LPT], mem0[R5_LPT];                      // in the leaflet / of type A or B as well
            int cA = 0, cB = 0, cL = 0, oob0 = 0;
            if (tid < 5) W.cnt[tid] = 0;
            __syncthreads();
#pragma unroll
            for (int u = 0; u < R5_LPT; u++) {
                const int i = tid + u * R5_THREADS;
                in0[u] = mem0[u] = false; c0x[u] = c0y[u] = 0.f; lox[u] = hix[u] = loy[u] = hiy[u] = 0.f;
                if (i < N && P.labels[f0 * P.N + i] == want) {
                    const int t = __ldg(P.type_code + i);
                    const bool a = (P.ta == -1 || t == P.ta), b = (P.tb == -1 || t == P.tb);
                    in0[u] = true;
                    cL++; cA += a; cB += b;
                    if (a || b) {
                        mem0[u] = true;
                        const double x = Fr.cf[i * 3 + P.lat0], y = Fr.cf[i * 3 + P.lat1];
                        oob0 |= !(x >= 0.0 && x <= Fr.Lx && y >= 0.0 && y <= Fr.Ly);
                        c0x[u] = (float)__dsub_rn(x, Fr.hx); c0y[u] = (float)__dsub_rn(y, Fr.hy);
                    }
                }
            }
            cA = pbk_warp_sum(cA); cB = pbk_warp_sum(cB); cL = pbk_warp_sum(cL);
            if (lane == 0) { atomicAdd(&W.cnt[0], cA); atomicAdd(&W.cnt[1], cB); atomicAdd(&W.cnt[2], cL); }
            const int any_oob0 = __syncthreads_or(oob0);
            cA = W.cnt[0]; cB = W.cnt[1]; cL = W.cnt[2];
            const bool on = selected && cA > 0 && cB > 0;
            const int vA = on ? cA : 0, vB = on ? cB : 0, vL = selected ? cL : 0;
            int64_t fe = f0 + 1;
            if (any_oob0) Fr.all_exact = true;                     // a member outside the box: exact formulas only
            const bool single = !on || Fr.all_exact;                // frame f0 cannot start a run
            if (!single) {
                for (; fe < fb; fe++) {
                    const float bxe = __ldg(P.box + fe * 3 + P.lat0), bye = __ldg(P.box + fe * 3 + P.lat1);
                    int stop = (bxe != Fr.bxf) || (bye != Fr.byf);
                    const double *cfe = P.com + fe * P.N * 3;
                    float nlx[R5_LPT], nhx[R5_LPT], nly[R5_LPT], nhy[R5_LPT];
#pragma unroll
                    for (int u = 0; u < R5_LPT; u++) {
                        const int i = tid + u * R5_THREADS;
                        nlx[u] = lox[u]; nhx[u] = hix[u]; nly[u] = loy[u]; nhy[u] = hiy[u];
                        if (i < N) {
                            // the leaflet (hence N_a, N_b, N_leaflet and the member set) must not change
                            stop |= ((P.labels[fe * P.N + i] == want) != in0[u]);
                            if (mem0[u]) {
                                const double x = cfe[i * 3 + P.lat0], y = cfe[i * 3 + P.lat1];
                                stop |= !(x >= 0.0 && x <= Fr.Lx && y >= 0.0 && y <= Fr.Ly);
                                float dx = (float)__dsub_rn(x, Fr.hx) - c0x[u], dy = (float)__dsub_rn(y, Fr.hy) - c0y[u];
                                if (dx > 0.5f * Fr.bxf) dx -= Fr.bxf; else if (dx < -0.5f * Fr.bxf) dx += Fr.bxf;
                                if (dy > 0.5f * Fr.byf) dy -= Fr.byf; else if (dy < -0.5f * Fr.byf) dy += Fr.byf;
                                nlx[u] = fminf(nlx[u], dx); nhx[u] = fmaxf(nhx[u], dx);
                                nly[u] = fminf(nly[u], dy); nhy[u] = fmaxf(nhy[u], dy);
                            }
                        }
                    }
                    if (__syncthreads_or(stop)) break;              // frame fe starts the next run
#pragma unroll
                    for (int u = 0; u < R5_LPT; u++) { lox[u] = nlx[u]; hix[u] = nhx[u]; loy[u] = nly[u]; hiy[u] = nhy[u]; }
                }
            }
            for (int64_t f = f0 + tid / 3; f < fe; f += R5_THREADS / 3) {      // (N_a, N_b, N_leaflet) of every frame
                const int k = tid % 3;
                if (tid < (R5_THREADS / 3) * 3) P.frame_counts[f * 6 + leaf * 3 + k] = k == 0 ? vA : (k == 1 ? vB : vL);
            }
            if (!on) { f0 = fe; continue; }

            // ================= build: cell sort by box centre, pair list ===========================
            // cell grid: edge >= 0.51 * Rb so that the half shell reaches 2 cells
            int ncx, ncy, mx, my;
            {
                const double wmin = 0.51 * (double)Rb + 1e-6;
                ncx = (int)floor(Fr.Lx / wmin); ncy = (int)floor(Fr.Ly / wmin);
                ncx = max(1, min(R5_MAXDIM, ncx));
                ncy = max(1, min(R5_MAXDIM, ncy));
                const double wx = Fr.Lx / ncx, wy = Fr.Ly / ncy;
                mx = (int)floor(((double)Rb + 1e-6) / wx) + 1; my = (int)floor(((double)Rb + 1e-6) / wy) + 1;
                if (2 * mx + 1 > ncx) { ncx = 1; mx = 0; }
                if (2 * my + 1 > ncy || my + 1 > R5_MAXROWS) { ncy = 1; my = 0; }
                if (single) { ncx = ncy = 1; mx = my = 0; }         // (no sweep: one cell)
            }
            const int ncell = ncx * ncy;
            const float iwx = (float)ncx / Fr.bxf, iwy = (float)ncy / Fr.byf;
            for (int c = tid; c <= R5_MAXCELLS; c += R5_THREADS) W.cell_fill[c] = 0;
            float cmx[R5_LPT], cmy[R5_LPT], ehx[R5_LPT], ehy[R5_LPT];
#pragma unroll
            for (int u = 0; u < R5_LPT; u++) {
                const int i = tid + u * R5_THREADS;
                cmx[u] = c0x[u] + 0.5f * (lox[u] + hix[u]); cmy[u] = c0y[u] + 0.5f * (loy[u] + hiy[u]);
                // half-extent + the float32 rounding of the box bookkeeping
                ehx[u] = 0.5f * (hix[u] - lox[u]) + 1.0e-5f * Fr.bxf; ehy[u] = 0.5f * (hiy[u] - loy[u]) + 1.0e-5f * Fr.byf;
                if (i < N) {
                    unsigned short cid = 0xFFFF;
                    if (mem0[u]) {
                        float gx = cmx[u] + 0.5f * Fr.bxf, gy = cmy[u] + 0.5f * Fr.byf;   // back to [0, L) for the cell index
                        if (gx < 0.0f) gx += Fr.bxf; else if (gx >= Fr.bxf) gx -= Fr.bxf;
                        if (gy < 0.0f) gy += Fr.byf; else if (gy >= Fr.byf) gy -= Fr.byf;
                        int cx = (int)(gx * iwx), cy = (int)(gy * iwy);
                        cx = max(0, min(ncx - 1, cx));
                        cy = max(0, min(ncy - 1, cy));
                        cid = (unsigned short)(cy * ncx + cx);
                    }
                    W.cid[i] = cid;
                }
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < R5_LPT; u++) if (mem0[u]) atomicAdd(&W.cell_fill[W.cid[tid + u * R5_THREADS]], 1);
            __syncthreads();
            if (warp == 0) {   // exclusive scan of the cell counts (<= 256 cells: 8 per lane)
                int v[8], s = 0;
#pragma unroll
                for (int q = 0; q < 8; q++) { const int c = lane * 8 + q; v[q] = c < ncell ? W.cell_fill[c] : 0; s += v[q]; }
                int incl = s;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
                int excl = incl - s;
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int c = lane * 8 + q;
                    if (c < ncell) { W.cell_start[c] = excl; W.cell_fill[c] = excl; }
                    excl += v[q];
                }
                if (lane == 31) W.cell_start[ncell] = excl;
            }
            __syncthreads();
            const int n_u = W.cell_start[ncell];
#pragma unroll
            for (int u = 0; u < R5_LPT; u++) {
                if (!mem0[u]) continue;
                const int i = tid + u * R5_THREADS;
                const int slot = atomicAdd(&W.cell_fill[W.cid[i]], 1);
                const bool wild = !(ehx[u] <= P.ecap && ehy[u] <= P.ecap);
                W.fxy[slot] = make_float2(cmx[u], cmy[u]);
                W.ext[slot] = make_float2(ehx[u], ehy[u]);
                W.idx[slot] = (unsigned short)i;
                const int t = __ldg(P.type_code + i);
                W.flag[slot] = (unsigned char)(((P.ta == -1 || t == P.ta) ? 1 : 0) | ((P.tb == -1 || t == P.tb) ? 2 : 0) | (wild ? 4 : 0));
                if (wild) { const int k = atomicAdd(&W.cnt[3], 1); if (k < R5_MAXW) W.wslot[k] = (unsigned short)slot; }
            }
            __syncthreads();
            const int nw = W.cnt[3];
            // a list needs 8-bit slots, every large box listed, and a frame that may start a run
            bool use_list = !single && n_u <= 256 && nw <= R5_MAXW;
            int qn = 0;                                           // this warp's queued exact pairs
            int n_list = 0;
            if (use_list) {
                // the sweep: thread = lipid p, forward half shell of cells
                for (int p0 = warp * 32; p0 < n_u; p0 += R5_THREADS) {
                    const int p = p0 + lane;
                    const bool pact = p < n_u;
                    float2 fp = make_float2(0.f, 0.f), ep = make_float2(0.f, 0.f);
                    int pcx = 0, pcy = 0;
                    bool pwild = true;
                    if (pact) {
                        fp = W.fxy[p]; ep = W.ext[p];
                        const int pc = W.cid[W.idx[p]];
                        pcy = pc / ncx; pcx = pc - pcy * ncx;
                        pwild = (W.flag[p] & 4u) != 0u;
                    }
                    for (int oy = 0; oy <= my; oy++) {
                        int b1 = 0, n1 = 0, b2 = 0, n2 = 0;
                        if (!pwild) {
                            int cy = pcy + oy; if (cy >= ncy) cy -= ncy;
                            const int row = cy * ncx;
                            int xlo = (oy == 0) ? pcx : pcx - mx, xhi = pcx + mx;
                            if (xlo < 0) {                      // wraps on the left
                                b2 = W.cell_start[row + xlo + ncx]; n2 = W.cell_start[row + ncx] - b2;
                                xlo = 0;
                            }
                            if (xhi >= ncx) {                   // wraps on the right
                                b2 = W.cell_start[row]; n2 = W.cell_start[row + xhi - ncx + 1] - b2;
                                xhi = ncx - 1;
                            }
                            b1 = W.cell_start[row + xlo]; n1 = W.cell_start[row + xhi + 1] - b1;
                            if (oy == 0) { n1 -= (p + 1 - b1); b1 = p + 1; }   // own cell: later entries only
                        }
                        const int ntot = n1 + n2;
                        for (int t = 0; __any_sync(0xffffffffu, t < ntot); t++) {
                            const bool act = t < ntot;
                            const int q = act ? (t < n1 ? b1 + t : b2 + (t - n1)) : 0;
                            const bool hit = act && !(W.flag[q] & 4u) &&
                                             r5_gap2(fp, ep, W.fxy[q], W.ext[q], Fr.bxf, Fr.byf) < Rc2;
                            const unsigned m = __ballot_sync(0xffffffffu, hit);
                            if (m) {
                                int base = 0;
                                if (lane == 0) base = atomicAdd(&W.cnt[4], __popc(m));
                                base = __shfl_sync(0xffffffffu, base, 0);
                                const int pos = base + __popc(m & lt_mask);
                                if (hit && pos < P.list_cap) W.list[pos] = (unsigned short)(p | (q << 8));
                            }
                        }
                    }
                }
                // lipids with large boxes: against every other member (two of them meet once, by slot order)
                for (int jw = 0; jw < nw; jw++) {
                    const int sw = W.wslot[jw];
                    const float2 cw = W.fxy[sw], ew = W.ext[sw];
                    for (int s0 = warp * 32; s0 < n_u; s0 += R5_THREADS) {
                        const int s = s0 + lane;
                        bool hit = false;
                        if (s < n_u && s != sw && !((W.flag[s] & 4u) && s < sw))
                            hit = r5_gap2(cw, ew, W.fxy[s], W.ext[s], Fr.bxf, Fr.byf) < Rc2;
                        const unsigned m = __ballot_sync(0xffffffffu, hit);
                        if (m) {
                            int base = 0;
                            if (lane == 0) base = atomicAdd(&W.cnt[4], __popc(m));
                            base = __shfl_sync(0xffffffffu, base, 0);
                            const int pos = base + __popc(m & lt_mask);
                            if (hit && pos < P.list_cap) W.list[pos] = (unsigned short)(sw | (s << 8));
                        }
                    }
                }
                __syncthreads();
                n_list = W.cnt[4];
                if (n_list > P.list_cap) use_list = false;          // list too small
            }

            if (use_list) {
                // ================= walk: per frame of the run, refresh the coordinates, bin the list ======
                for (int64_t f = f0; f < fe; f++) {
                    const double *cf = P.com + f * P.N * 3;
                    __syncthreads();                               // (the list is complete / the previous walk is done)
                    for (int s = tid; s < n_u; s += R5_THREADS) {
                        const int i = W.idx[s];
                        W.fxy[s] = make_float2((float)__dsub_rn(cf[i * 3 + P.lat0], Fr.hx), (float)__dsub_rn(cf[i * 3 + P.lat1], Fr.hy));
                    }
                    __syncthreads();
                    const unsigned df = (unsigned)(f - f0);
                    for (int e0 = 0; e0 < n_list; e0 += 4 * R5_THREADS) {
                        int pp[4], qq[4];
                        bool amb[4];
                        unsigned ent[4], ee[4];
                        float d2[4];
                        // four independent chains first (list entry -> coordinates -> d^2 -> table), then the counts
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int k = e0 + R5_THREADS * j + tid;
                            ent[j] = k < n_list ? (unsigned)W.list[k] : 0u;        // (slot 0 with itself, never counted)
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            pp[j] = ent[j] & 0xFFu; qq[j] = ent[j] >> 8;
                            d2[j] = r5_d2(W.fxy[pp[j]], W.fxy[qq[j]], Fr.bxf, Fr.byf);
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) ee[j] = __ldg(P.lut + min(__float2int_rd(d2[j] * lut_scale), lut_last));
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const bool act = e0 + R5_THREADS * j + tid < n_list;
                            const unsigned e = ee[j];
                            unsigned w = 2u;
                            if (!SAME) {
                                const unsigned fpf = W.flag[pp[j]], fqf = W.flag[qq[j]];
                                w = ((fpf & 1u) & ((fqf >> 1) & 1u)) + ((fqf & 1u) & ((fpf >> 1) & 1u));
                            }
                            const bool sure = (int)(e >> 8) >= Fr.tolc;
                            if (act && sure && (e & 0xFFu) != 0xFFu && w) atomicAdd(&W.hist[e & 0xFFu], w);
                            amb[j] = act && !sure && w;
                        }
                        if (__any_sync(0xffffffffu, amb[0] | amb[1] | amb[2] | amb[3])) {
#pragma unroll
                            for (int j = 0; j < 4; j++) r5_push<SAME>(amb[j], pp[j], qq[j], df, qn, lane, W, Fr, P, s_thr, f0);
                        }
                    }
                }
            } else {
                // ================= no list: frame f0 alone, every pair of members straight from its
                // coordinates (out-of-box members, > 256 members, overflow); the run is cut to f0 ========
                fe = f0 + 1;
                __syncthreads();
                for (int s = tid; s < n_u; s += R5_THREADS) {
                    const int i = W.idx[s];
                    W.fxy[s] = make_float2((float)__dsub_rn(Fr.cf[i * 3 + P.lat0], Fr.hx), (float)__dsub_rn(Fr.cf[i * 3 + P.lat1], Fr.hy));
                }
                __syncthreads();
                for (int p = 0; p < n_u; p++) {
                    const float2 fp = W.fxy[p];
                    const unsigned fpflag = W.flag[p];
                    for (int s0 = warp * 32; s0 < n_u; s0 += R5_THREADS) {
                        if (s0 + 31 <= p) continue;
                        const int s = s0 + lane;
                        const bool act = s > p && s < n_u;
                        r5_pair<SAME>(act, p, act ? s : 0, fp, W.fxy[act ? s : 0], fpflag, 0u, qn, lane, W, Fr, P, s_thr, lut_scale, lut_last, f0);
                    }
                }
            }
            __syncwarp();
            r5_drain<SAME>(qn, lane, W, Fr, P, s_thr, f0);
            __syncthreads();
            f0 = fe;
        }
    }
    __syncthreads();
    for (int q = tid; q < nb; q += R5_THREADS) {
        unsigned long long t = 0;
        for (int c = 0; c < R5_WARPS; c++) t += hist_all[(size_t)c * nb + q];
        if (t) atomicAdd(&P.count[q], t);
    }
}

static size_t r5_smem(int64_t N, int nb, int list_cap)
{
    const size_t npad = ((size_t)N + 15) & ~(size_t)15;
    return ((((size_t)(nb + 1) * 8) + 15) & ~(size_t)15) + npad * 16 + (((size_t)list_cap * 2 + 15) & ~(size_t)15) +
           ((((size_t)R5_WARPS * nb * 4) + 15) & ~(size_t)15) + (size_t)R5_WARPS * R5_QCAP * 4 + (size_t)(2 * R5_MAXCELLS + 2) * 4 + 32 +
           npad * 4 + (size_t)R5_MAXW * 2 + npad + 16;
}

// d^2 table: cells of width 2^-m, up to 8 per narrowest threshold gap (w^2 for uniform bins of
// width w starting at lo >= 0), capped at PBK_LUT_ENTRIES (a cell wider than a gap is still
// handled correctly -- it is marked ambiguous).  Shared with pbk_cells.cu.
bool pbk_rdf_lut16_shape(int32_t n_bins, double range_lo, double range_hi, int *m_out, int *lut_n_out)
{
    const double w = (range_hi - range_lo) / n_bins;
    const double lo = range_lo > 0 ? range_lo : 0.0;
    const double min_gap = (lo + w) * (lo + w) - lo * lo;
    int m = 0;
    while (ldexp(1.0, -m) > min_gap / 8.0 && m < 30) m++;
    while (m > -20 && (range_hi * range_hi + 2.0) * ldexp(1.0, m) + 520.0 > (double)PBK_LUT_ENTRIES) m--;
    const double lut_nd = floor((range_hi * range_hi + 1.0) * ldexp(1.0, m)) + 516.0;
    if (lut_nd > (double)PBK_LUT_ENTRIES || lut_nd < 8.0) return false;
    *m_out = m;
    *lut_n_out = (int)lut_nd;
    return true;
}

int pbk_rdf_lut16_build(pbk_ctx *ctx, const double *thr, int nb, int lut_n, int lut_m, unsigned short *lut,
                        cudaStream_t st)
{
    k_rdf5_lut<<<(lut_n + 255) / 256, 256, 0, st>>>(thr, nb, lut_n, lut_m, lut);
    PBK_CHECK_LAUNCH(ctx, "k_rdf5_lut");
    return PBK_OK;
}

// Returns PBK_EUNSUPPORTED when the shape belongs to another RDF kernel.
int pbk_rdf5_launch(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                    const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                    int leaflet_mask, int32_t n_bins, double range_lo, double range_hi, const double *thr,
                    unsigned long long *count, int32_t *frame_counts, cudaStream_t st)
{
    if (!thr || N > R5_MAXN || n_bins > 254 || range_lo < 0 || !ctx->work || !ctx->lut || getenv("PBK_RDF_NOLIST") ||
        getenv("PBK_RDF_V2"))
        return PBK_EUNSUPPORTED;
    int m = 0, lut_n = 0;
    if (!pbk_rdf_lut16_shape(n_bins, range_lo, range_hi, &m, &lut_n)) return PBK_EUNSUPPORTED;
    float ecap = 3.0f;
    if (const char *ev = getenv("PBK_RDF_SKIN")) { const float v = (float)atof(ev); if (v > 0.0f && v < 100.0f) ecap = 0.5f * v; }
    // pair list: room for a leaflet of 256 lipids at 64 A^2 each within range_hi + 2 ecap, plus a third
    const double rb = range_hi + 2.0 * ecap + 0.05;
    double want_cap = 128.0 * 3.14159 * rb * rb / 64.0 * 1.33 + 256.0;
    if (want_cap > 32760.0) want_cap = 32760.0;
    int list_cap = ((int)want_cap + 7) & ~7;
    size_t smem = r5_smem(N, n_bins, list_cap);
    if (smem > 100 * 1024) { list_cap = 2048; smem = r5_smem(N, n_bins, list_cap); }
    if (smem > 200 * 1024) return PBK_EUNSUPPORTED;
    const bool same = (type_a == type_b);
    auto kern = same ? k_rdf5<true> : k_rdf5<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "k_rdf5 smem attribute", e);
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, R5_THREADS, smem);
    if (e != cudaSuccess || per_sm < 1) per_sm = 1;
    // frames per work item: one item per resident CTA when the run is short enough (a single wave,
    // the longest list reuse), at most 64 frames
    const int64_t ctas = (int64_t)ctx->sm_count * per_sm;
    int64_t chunk = (2 * F + ctas - 1) / ctas;
    if (chunk < 1) chunk = 1;
    if (chunk > 64) chunk = 64;
    if (const char *ev = getenv("PBK_RDF_CHUNK")) { const int v = atoi(ev); if (v >= 1 && v <= 1024) chunk = v; }   // tests
    const int64_t n_items = 2 * ((F + chunk - 1) / chunk);
    int64_t grid = ctas < n_items ? ctas : n_items;
    const unsigned slot = ctx->work_next++;
    int *work = ctx->work + (slot % PBK_WORK_SLOTS);                // launches in flight do not share a counter
    unsigned short *lut = ctx->lut + (size_t)(slot % PBK_LUT_SLOTS) * PBK_LUT_ENTRIES;
    e = cudaMemsetAsync(work, 0, sizeof(int), st);
    if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "k_rdf5 work counter", e);
    k_rdf5_lut<<<(lut_n + 255) / 256, 256, 0, st>>>(thr, n_bins, lut_n, m, lut);
    PBK_CHECK_LAUNCH(ctx, "k_rdf5_lut");
    Rdf5Params P;
    P.com = com; P.labels = labels; P.type_code = type_code; P.box = box; P.F = F; P.N = N;
    P.lat0 = lat0; P.lat1 = lat1; P.ta = type_a; P.tb = type_b; P.leaflet_mask = leaflet_mask;
    P.nb = n_bins; P.range_hi = range_hi; P.ecap = ecap; P.thr = thr; P.count = count;
    P.frame_counts = frame_counts; P.lut_n = lut_n; P.lut_m = m; P.chunk = (int)chunk; P.list_cap = list_cap;
    P.work = work; P.lut = lut;
    kern<<<(unsigned)grid, R5_THREADS, smem, st>>>(P);
    PBK_CHECK_LAUNCH(ctx, "k_rdf5");
    return PBK_OK;
}
