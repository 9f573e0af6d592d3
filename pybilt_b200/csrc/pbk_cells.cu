// pbk_cells.cu -- minimum-image pair work for LARGE leaflets on global-memory cell lists:
// the COM lateral RDF, the k nearest neighbours of every lipid (nearest-neighbour fraction,
// the 24-neighbour table of the heuristic lipid grid).
//
// Same arithmetic contract as the small-system kernels (pbk_rdf2.cu / pbk_pairs.cu /
// pbk_grid.cu), restating
//   pybilt/common/distance_cutoff_clustering.py:27-71            (distance_euclidean_pbc)
//   pybilt/bilayer_analyzer/analysis_protocols.py:4900-5006      (COMLateralRDFProtocol)
//   pybilt/bilayer_analyzer/analysis_protocols.py:2174-2257      (NNFProtocol)
//   pybilt/lipid_grid/lipid_grid_opt.py:435-468                  (_knn, 24 neighbours)
// but for leaflets of 10^3 .. 10^6 lipids, where neither a leaflet in shared memory nor an
// O(N^2) sweep is an option (BASELINE configs C3, C4, C5).
//
// Data flow per chunk of frames (scratch from the context's arena, ~33 B per lipid-frame):
//   k_cl_count    every member lipid -> its (leaflet, cell); cell populations by atomics;
//                 per-frame class counts (N_a, N_b, N_leaflet) and the out-of-box flag
//   k_cl_scan     exclusive scan of the cell populations per frame
//   k_cl_scatter  lipids into cell order: b' = x - box/2 as float64 (exact phase) and as
//                 float32 (prefilter), original index, class flags, cell id
//   k_rdf_cl      thread per lipid, sweeps the forward half shell of cells (every unordered
//                 pair once); float32 prefilter -> per-warp ring queue -> the reference's
//                 float64 formulas; d^2 -> bin through the exact threshold table
//   k_knn_cl      thread per lipid, rings of cells of growing radius until the k-th best
//                 distance is inside the searched radius; exact float64 distances, ties in
//                 ascending lipid index (Python's stable sort)
// Cells assume wrapped coordinates (0 <= x <= L).  A frame with a member outside the box is
// flagged and handed to the dense kernels, which evaluate the reference formula for any input.
#include <math.h>
#include <stdlib.h>

#include "pbk_common.cuh"

#define CL_THREADS 256
#define CL_QCAP 128
#define CL_MAX_TYPES 8

struct ClGeom {
    int ncx, ncy, ncell, mx, my;
    double iwx, iwy, wx, wy;
};

// cell geometry of a frame: a pure function of the box, so every kernel derives the same one.
//   wmin > 0 : RDF, cell edge >= wmin (= 0.51 range_hi: the half shell reaches two cells)
//   wmin <= 0: kNN, about `occ` lipids per cell
__device__ __forceinline__ ClGeom cl_geom(double Lx, double Ly, double wmin, double range_hi,
                                          double occ, int64_t N, int ncap)
{
    ClGeom g;
    double w = wmin;
    if (!(w > 0.0)) w = sqrt(occ * Lx * Ly / fmax(1.0, 0.5 * (double)N));
    int ncx = 1, ncy = 1;
    if (Lx > 0.0 && Ly > 0.0 && w > 0.0) {
        const double fx = floor(Lx / w), fy = floor(Ly / w);
        ncx = fx < (double)ncap ? (int)fx : ncap;
        ncy = fy < (double)ncap ? (int)fy : ncap;
        ncx = max(1, ncx);
        ncy = max(1, ncy);
    }
    g.mx = g.my = 0;
    if (wmin > 0.0) {
        g.mx = (int)floor((range_hi + 1e-6) / (Lx / ncx)) + 1;
        g.my = (int)floor((range_hi + 1e-6) / (Ly / ncy)) + 1;
        if (2 * g.mx + 1 > ncx) { ncx = 1; g.mx = 0; }
        if (2 * g.my + 1 > ncy || g.my > 7) { ncy = 1; g.my = 0; }
    }
    g.ncx = ncx; g.ncy = ncy; g.ncell = ncx * ncy;
    g.wx = Lx / ncx; g.wy = Ly / ncy;
    g.iwx = ncx / Lx; g.iwy = ncy / Ly;
    return g;
}

struct ClParams {
    const double *com;
    const uint8_t *labels;
    const int32_t *type_code;
    const float *box;
    int64_t F, N;
    int lat0, lat1, ta, tb, leaflet_mask;
    int members_all;                // 1: every leaflet member is listed; 0: only type A or B
    double wmin, range_hi, occ;
    int ncap;
    int64_t cstride;                // >= 2*ncap*ncap + 1 ints of cell data per frame, multiple of 4
    int *cell_fill;                 // [F][cstride] populations, then running fill pointers
    int *cell_start;                // [F][cstride]
    int *fcounts;                   // [F][8]: (N_a, N_b, N_leaf) upper, lower; out-of-box; listed
    double2 *sxy;                   // [F][N] b' = (x, y) - box/2 in cell order (one 16-byte load per lipid)
    float2 *sf;                     // [F][N] float32 copy
    int *sidx;                      // [F][N] original lipid index
    int *scell;                     // [F][N] leaf*ncell + cell
    unsigned char *sflag;           // [F][N] bit0 type A, bit1 type B
    int32_t *frame_counts;          // RDF output [F][2][3] (may be NULL)
    int multi, n_types;             // multi != 0: every type pair at once; sflag holds the type code
    int32_t *type_counts;           // multi: [F][2][n_types] lipids per leaflet and type
    int32_t *dummy_fc;              // [F][6] scratch for the dense kernel's per-frame counts
};

struct ClFrame {
    double Lx, Ly, hx, hy;
    float bxf, byf;
    ClGeom g;
};

__device__ __forceinline__ ClFrame cl_frame(const ClParams &P, int64_t f)
{
    ClFrame fr;
    fr.bxf = P.box[f * 3 + P.lat0];
    fr.byf = P.box[f * 3 + P.lat1];
    const float cxf = fr.bxf * 0.5f, cyf = fr.byf * 0.5f;
    fr.Lx = (double)fr.bxf; fr.Ly = (double)fr.byf; fr.hx = (double)cxf; fr.hy = (double)cyf;
    fr.g = cl_geom(fr.Lx, fr.Ly, P.wmin, P.range_hi, P.occ, P.N, P.ncap);
    return fr;
}

// (leaflet row, flags, listed?) of lipid i; leaflet row 0 = upper (label 1), 1 = lower
__device__ __forceinline__ bool cl_classify(const ClParams &P, const uint8_t *lab, int64_t i, int *leaf,
                                            unsigned *flag)
{
    *leaf = lab[i] ? 0 : 1;
    if (P.multi) {
        const int t = P.type_code[i];
        *flag = (unsigned)t;
        return t >= 0 && t < P.n_types;          // a code outside the table is not listed (never indexes it)
    }
    const int t = (P.ta == -1 && P.tb == -1) ? 0 : P.type_code[i];
    const bool a = (P.ta == -1 || t == P.ta), b = (P.tb == -1 || t == P.tb);
    *flag = (a ? 1u : 0u) | (b ? 2u : 0u);
    return P.members_all || a || b;
}

__device__ __forceinline__ int cl_cell(const ClFrame &fr, double x, double y)
{
    int cx = (int)(x * fr.g.iwx), cy = (int)(y * fr.g.iwy);
    cx = max(0, min(fr.g.ncx - 1, cx));
    cy = max(0, min(fr.g.ncy - 1, cy));
    return cy * fr.g.ncx + cx;
}

__global__ void __launch_bounds__(CL_THREADS)
k_cl_count(ClParams P)
{
    const int64_t f = blockIdx.y;
    const ClFrame fr = cl_frame(P, f);
    const double *cf = P.com + f * P.N * 3;
    const uint8_t *lab = P.labels + f * P.N;
    int *fill = P.cell_fill + f * P.cstride;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int c[7] = {0, 0, 0, 0, 0, 0, 0};
    __shared__ int s_tc[2 * CL_MAX_TYPES];
    if (threadIdx.x < 2 * CL_MAX_TYPES) s_tc[threadIdx.x] = 0;
    __syncthreads();
    if (i < P.N) {
        int leaf;
        unsigned flag;
        const bool listed = cl_classify(P, lab, i, &leaf, &flag);
        if (P.multi) {
            c[leaf * 3 + 0] = c[leaf * 3 + 1] = 1;
            if ((int)flag < P.n_types) atomicAdd(&s_tc[leaf * P.n_types + (int)flag], 1);
        } else {
            c[leaf * 3 + 0] = flag & 1u;
            c[leaf * 3 + 1] = (flag >> 1) & 1u;
        }
        c[leaf * 3 + 2] = 1;
        if (listed) {
            const double x = cf[i * 3 + P.lat0], y = cf[i * 3 + P.lat1];
            c[6] = !(x >= 0.0 && x <= fr.Lx && y >= 0.0 && y <= fr.Ly);
            atomicAdd(&fill[leaf * fr.g.ncell + cl_cell(fr, x, y)], 1);
        }
    }
    // class counts: warp sums -> shared -> one global atomic per counter and CTA
    __shared__ int s_c[7];
    if (threadIdx.x < 7) s_c[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 7; q++) {
        const int v = pbk_warp_sum(c[q]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_c[q], v);
    }
    __syncthreads();
    if (threadIdx.x < 7 && s_c[threadIdx.x]) atomicAdd(&P.fcounts[f * 8 + threadIdx.x], s_c[threadIdx.x]);
    if (P.multi && threadIdx.x < 2 * P.n_types && s_tc[threadIdx.x])
        atomicAdd(&P.type_counts[f * 2 * P.n_types + threadIdx.x], s_tc[threadIdx.x]);
}

// one CTA per frame: exclusive scan of the 2*ncell populations (tiles of 4096 entries, 128-bit
// loads; the per-frame stride is a multiple of 4 ints); fill pointers = starts
__global__ void __launch_bounds__(1024)
k_cl_scan(ClParams P)
{
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int64_t f = blockIdx.x;
    const ClFrame fr = cl_frame(P, f);
    const int n = 2 * fr.g.ncell;
    int *fill = P.cell_fill + f * P.cstride;
    int *start = P.cell_start + f * P.cstride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 4096) {
        const int c0 = base + tid * 4;
        int4 v = make_int4(0, 0, 0, 0);
        if (c0 + 3 < n) v = *reinterpret_cast<const int4 *>(fill + c0);
        else {
            if (c0 < n) v.x = fill[c0];
            if (c0 + 1 < n) v.y = fill[c0 + 1];
            if (c0 + 2 < n) v.z = fill[c0 + 2];
        }
        const int s = v.x + v.y + v.z + v.w;
        int incl = s;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = s_warp[lane];
            int t = w;
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
            s_warp[lane] = t - w;
        }
        __syncthreads();
        const int carry = s_carry;
        int4 e;
        e.x = carry + s_warp[warp] + incl - s;
        e.y = e.x + v.x; e.z = e.y + v.y; e.w = e.z + v.z;
        if (c0 + 3 < n) {
            *reinterpret_cast<int4 *>(start + c0) = e;
            *reinterpret_cast<int4 *>(fill + c0) = e;
        } else {
            if (c0 < n) { start[c0] = e.x; fill[c0] = e.x; }
            if (c0 + 1 < n) { start[c0 + 1] = e.y; fill[c0 + 1] = e.y; }
            if (c0 + 2 < n) { start[c0 + 2] = e.z; fill[c0 + 2] = e.z; }
        }
        __syncthreads();
        if (tid == 1023) s_carry = e.w + v.w;
        __syncthreads();
    }
    if (tid == 0) {
        start[n] = s_carry;
        if (P.frame_counts && !P.fcounts[f * 8 + 6]) {
            for (int l = 0; l < 2; l++) {
                const bool sel = (P.leaflet_mask >> l) & 1;
                const int ca = P.fcounts[f * 8 + l * 3], cb = P.fcounts[f * 8 + l * 3 + 1];
                const bool on = sel && ca > 0 && cb > 0;
                P.frame_counts[f * 6 + l * 3 + 0] = on ? ca : 0;
                P.frame_counts[f * 6 + l * 3 + 1] = on ? cb : 0;
                P.frame_counts[f * 6 + l * 3 + 2] = sel ? P.fcounts[f * 8 + l * 3 + 2] : 0;
            }
        }
    }
}

__global__ void __launch_bounds__(CL_THREADS)
k_cl_scatter(ClParams P)
{
    const int64_t f = blockIdx.y;
    if (P.fcounts[f * 8 + 6]) return;            // out-of-box frame: dense kernels take it
    const ClFrame fr = cl_frame(P, f);
    const double *cf = P.com + f * P.N * 3;
    const uint8_t *lab = P.labels + f * P.N;
    int *fill = P.cell_fill + f * P.cstride;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N) return;
    int leaf;
    unsigned flag;
    if (!cl_classify(P, lab, i, &leaf, &flag)) return;
    const double x = cf[i * 3 + P.lat0], y = cf[i * 3 + P.lat1];
    const int cell = leaf * fr.g.ncell + cl_cell(fr, x, y);
    const int64_t slot = f * P.N + atomicAdd(&fill[cell], 1);
    const double bx = __dsub_rn(x, fr.hx), by = __dsub_rn(y, fr.hy);
    P.sxy[slot] = make_double2(bx, by);
    P.sf[slot] = make_float2((float)bx, (float)by);
    P.sidx[slot] = (int)i;
    P.scell[slot] = cell;
    P.sflag[slot] = (unsigned char)flag;
}

// ---------------------------------------------------------------------------------------
// RDF sweep
// ---------------------------------------------------------------------------------------
struct ClRdf {
    int nb, hist_copies, lut_n, lut_m;
    const double *thr;
    const unsigned short *lut;      // d^2 cell -> bin | distance to the nearest threshold << 8 (k_rdf5_lut)
    unsigned long long *count;
    int *work;                      // dynamic work counter
    int chunk;                      // slots per work item
};

__device__ __forceinline__ int cl_bin(double d2, const double *s_thr, int nb)
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
// MODE 0: type_a == type_b (every listed lipid is both); 1: flags per lipid; 2: multi -- every
// ordered type pair has its own histogram, hist[((leaf*T + t_from)*T + t_to)*nb + bin]
template <int MODE>
__device__ __forceinline__ void cl_exact(uint2 e, const double2 *gxy,
                                         const unsigned char *gflag, const double *s_thr, unsigned *myhist,
                                         int nb, double Lx, double Ly, double hx, double hy,
                                         const int *gcell, int ncell, int T)
{
    const int p = (int)e.x, q = (int)e.y;
    int w1 = 1, w2 = 1;
    unsigned *h1 = myhist, *h2 = myhist;
    if (MODE == 2) {
        const int tp = gflag[p], tq = gflag[q], leaf = gcell[p] >= ncell;
        h1 = myhist + (size_t)((leaf * T + tp) * T + tq) * nb;
        h2 = myhist + (size_t)((leaf * T + tq) * T + tp) * nb;
    }
    if (MODE == 1) {
        const unsigned fp = gflag[p], fq = gflag[q];
        w1 = (fp & 1u) && (fq & 2u);            // ordered p -> q
        w2 = (fq & 1u) && (fp & 2u);            // ordered q -> p
        if (!(w1 | w2)) return;
    }
    const double2 ap = gxy[p], bq = gxy[q];
    const double apx = ap.x, apy = ap.y, bqx = bq.x, bqy = bq.y;
    double dx = __dsub_rn(apx, bqx), dy = __dsub_rn(apy, bqy);
    const bool wrx = fabs(dx) > hx, wry = fabs(dy) > hy;
    if (!(wrx | wry)) {
        const int k = cl_bin(fma(dy, dy, __dmul_rn(dx, dx)), s_thr, nb);
        if (k >= 0) {
            if (MODE == 2 && h1 != h2) { atomicAdd(&h1[k], 1u); atomicAdd(&h2[k], 1u); }
            else atomicAdd(&h1[k], (unsigned)(w1 + w2));
        }
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
    const int ka = cl_bin(d2a, s_thr, nb);
    int kb = ka;
    if (d2b != d2a) kb = cl_bin(d2b, s_thr, nb);
    if (ka == kb && h1 == h2) {
        if (ka >= 0) atomicAdd(&h1[ka], (unsigned)(w1 + w2));
    } else {
        if (w1 && ka >= 0) atomicAdd(&h1[ka], 1u);
        if (w2 && kb >= 0) atomicAdd(&h2[kb], 1u);
    }
}

// A candidate's bin comes from its float32 minimum-image d^2 through the 16-bit table; the
// table entry also says how many table cells separate the d^2 cell from the nearest bin
// threshold, and a pair closer to a threshold than the float32 rounding bound (`tolc` cells,
// see DESIGN.md "Why the RDF fast path is exact") goes to the exact float64 queue instead.
template <int MODE>
__global__ void __launch_bounds__(CL_THREADS, 4)
k_rdf_cl(ClParams P, ClRdf R)
{
    constexpr bool SAME = (MODE == 0);
    extern __shared__ unsigned char smem_raw[];
    const int nb = R.nb;
    const int T = P.n_types;
    const int hsize = (MODE == 2) ? 2 * T * T * nb : nb;     // entries per histogram copy
    double *s_thr = reinterpret_cast<double *>(smem_raw);                              // nb+1
    uint2 *s_queue = reinterpret_cast<uint2 *>(s_thr + (nb + 1));                      // warps*QCAP
    unsigned *s_hist = reinterpret_cast<unsigned *>(s_queue + (CL_THREADS / 32) * CL_QCAP);   // copies*nb
    unsigned short *s_lut = reinterpret_cast<unsigned short *>(s_hist + (size_t)R.hist_copies * hsize);
    __shared__ int s_item;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int q = tid; q <= nb; q += blockDim.x) s_thr[q] = R.thr[q];
    for (int q = tid; q < R.hist_copies * hsize; q += blockDim.x) s_hist[q] = 0u;
    for (int c = tid; c < R.lut_n; c += blockDim.x) s_lut[c] = R.lut[c];
    unsigned *myhist = s_hist + (size_t)(warp % R.hist_copies) * hsize;
    uint2 *myq = s_queue + warp * CL_QCAP;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t per_frame = (P.N + R.chunk - 1) / R.chunk;
    const int64_t n_items = P.F * per_frame;
    const float lut_scale = ldexpf(1.0f, R.lut_m);
    const int lut_last = R.lut_n - 1;
    __syncthreads();

    for (;;) {
        if (tid == 0) s_item = atomicAdd(R.work, 1);
        __syncthreads();
        const int64_t item = s_item;
        __syncthreads();
        if (item >= n_items) break;
        const int64_t f = item / per_frame;
        const int p_lo = (int)(item - f * per_frame) * R.chunk;
        const int *fc = P.fcounts + f * 8;
        if (fc[6]) continue;                                     // out-of-box frame
        const ClFrame fr = cl_frame(P, f);
        const ClGeom &g = fr.g;
        const int *cstart = P.cell_start + f * P.cstride;
        const int n_u = cstart[2 * g.ncell];
        if (p_lo >= n_u) continue;
        const int p_hi = min(n_u, p_lo + R.chunk);
        bool on[2];
        for (int l = 0; l < 2; l++) on[l] = ((P.leaflet_mask >> l) & 1) && fc[l * 3] > 0 && fc[l * 3 + 1] > 0;
        const double2 *gxy = P.sxy + f * P.N;
        const float2 *gf = P.sf + f * P.N;
        const int *gcell = P.scell + f * P.N;
        const unsigned char *gflag = P.sflag + f * P.N;
        const float bxf = fr.bxf, byf = fr.byf;
        // rounding bound of the float32 d^2 near the thresholds, in table cells (+1 of slack):
        // |d2_f32 - d2_exact| <= 2^-23 (r+1) (8 L + 4 (r+1)), r = range_hi, L = larger box edge
        const float Lm = fmaxf(bxf, byf), rr = (float)P.range_hi + 1.0f;
        const float tol = 1.1920929e-7f * rr * (8.0f * Lm + 4.0f * rr);
        const int tolc = (int)ceilf(tol * lut_scale) + 1;
        const bool all_exact = !(tol * lut_scale < 200.0f);
        const float cut = rr + 0.5f * tol;
        const float cut2_f = cut * cut;                          // all_exact frames: plain cut-off
        const int mx = g.mx, my = g.my, ncx = g.ncx, ncy = g.ncy;

        // one lipid per thread; the warp walks the half-shell rows in lockstep, per row first the
        // main slot range of every lane, then the (rare) range beyond the periodic edge
        int qhead = 0, qtail = 0;
        const unsigned sure_min = all_exact ? 0xFFFFFFFFu : ((unsigned)tolc << 8);
        for (int p0 = p_lo; p0 < p_hi; p0 += blockDim.x) {
            const int p = p0 + tid;
            bool pact = p < p_hi;
            float fpx = 0.f, fpy = 0.f;
            unsigned pflag = 3u, ptype = 0u, pleaf = 0u;
            int cbase = 0, pcx = 0, pcy = 0;
            if (pact) {
                int pc = gcell[p];
                const int leaf = pc >= g.ncell;
                pact = on[leaf];
                const float2 fp = gf[p];
                fpx = fp.x; fpy = fp.y;
                if (!SAME) pflag = gflag[p];
                if (MODE == 2) { ptype = pflag; pleaf = (unsigned)leaf; pflag = (pleaf * T + ptype) * (unsigned)T; }   // row of the pair table
                cbase = leaf * g.ncell;
                pc -= cbase;
                pcy = pc / ncx;
                pcx = pc - pcy * ncx;
            }
            for (int oy = 0; oy <= my; oy++) {
                int seg_b[2] = {0, 0}, seg_n[2] = {0, 0};
                if (pact) {
                    int cy = pcy + oy; if (cy >= ncy) cy -= ncy;
                    const int row = cbase + cy * ncx;
                    int xlo = (oy == 0) ? pcx : pcx - mx, xhi = pcx + mx;
                    if (xlo < 0) {
                        seg_b[1] = cstart[row + xlo + ncx]; seg_n[1] = cstart[row + ncx] - seg_b[1];
                        xlo = 0;
                    }
                    if (xhi >= ncx) {
                        seg_b[1] = cstart[row]; seg_n[1] = cstart[row + xhi - ncx + 1] - seg_b[1];
                        xhi = ncx - 1;
                    }
                    seg_b[0] = cstart[row + xlo]; seg_n[0] = cstart[row + xhi + 1] - seg_b[0];
                    if (oy == 0) { seg_n[0] -= (p + 1 - seg_b[0]); seg_b[0] = p + 1; }   // own cell: later entries only
                }
#pragma unroll
                for (int sg = 0; sg < 2; sg++) {
                    const int n = seg_n[sg], b = seg_b[sg];
                    const int nmax = __reduce_max_sync(0xffffffffu, n);
                    const float2 *gp = gf + b;
                    for (int t = 0; t < nmax; t++) {
                        bool hit = false;
                        if (t < n) {
                            const float2 fq = gp[t];
                            float dx = fabsf(fpx - fq.x), dy = fabsf(fpy - fq.y);
                            dx = fminf(dx, bxf - dx);
                            dy = fminf(dy, byf - dy);
                            const float d2 = fmaf(dy, dy, dx * dx);
                            const unsigned e = s_lut[min(__float2int_rd(d2 * lut_scale), lut_last)];
                            const unsigned bin = e & 0xFFu;
                            if (e >= sure_min) {
                                if (bin != 0xFFu) {
                                    if (MODE == 2) {
                                        // ordered pairs p -> q and q -> p land in their own tables
                                        const unsigned tq = gflag[b + t];
                                        const unsigned i1 = (pflag + tq) * (unsigned)nb + bin;
                                        const unsigned i2 = ((pleaf * T + tq) * T + ptype) * (unsigned)nb + bin;
                                        if (i1 == i2) atomicAdd(&myhist[i1], 2u);
                                        else { atomicAdd(&myhist[i1], 1u); atomicAdd(&myhist[i2], 1u); }
                                    } else {
                                        unsigned w = 2u;
                                        if (MODE == 1) {
                                            const unsigned fqf = gflag[b + t];
                                            w = ((pflag & 1u) & ((fqf >> 1) & 1u)) + ((fqf & 1u) & ((pflag >> 1) & 1u));
                                        }
                                        if (w) atomicAdd(&myhist[bin], w);
                                    }
                                }
                            } else {
                                hit = !all_exact || d2 < cut2_f;
                            }
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, hit);
                        if (m) {
                            if (hit) myq[(qtail + __popc(m & lt_mask)) & (CL_QCAP - 1)] = make_uint2((unsigned)p, (unsigned)(b + t));
                            qtail += __popc(m);
                            if (qtail - qhead >= 32) {
                                __syncwarp();
                                cl_exact<MODE>(myq[(qhead + lane) & (CL_QCAP - 1)], gxy, gflag, s_thr, myhist, nb,
                                               fr.Lx, fr.Ly, fr.hx, fr.hy, gcell, g.ncell, T);
                                qhead += 32;
                                __syncwarp();
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane < qtail - qhead)
            cl_exact<MODE>(myq[(qhead + lane) & (CL_QCAP - 1)], gxy, gflag, s_thr, myhist, nb, fr.Lx, fr.Ly,
                           fr.hx, fr.hy, gcell, g.ncell, T);
        __syncwarp();
    }
    __syncthreads();
    for (int q = tid; q < hsize; q += blockDim.x) {
        unsigned long long t = 0;
        for (int c = 0; c < R.hist_copies; c++) t += s_hist[(size_t)c * hsize + q];
        if (t) atomicAdd(&R.count[q], t);
    }
}

// ---------------------------------------------------------------------------------------
// k nearest leaflet members of every lipid
// ---------------------------------------------------------------------------------------
struct ClKnn {
    int k;
    int32_t *n_b;                   // NNF: [F][N] count of type-B lipids among the k nearest (-1: not a reference lipid);
                                    // multi: [F][N][n_types] hits per type
    int32_t *knn;                   // grid: [F][N][24] neighbour indices, -1 padded
};

// GRID: the reference's _knn evaluates each unordered pair once as dist(lower index, higher
// index) (lipid_grid_opt.py:452-459); NNF: dist(reference lipid, other) (:2204).
// MODE 0: NNF counts of type-B neighbours; 1 (GRID): the neighbour table; 2: NNF for every type at
// once -- hits[f][i][t] = lipids of type t among the k nearest (every lipid is a reference lipid)
template <int KMAX, int MODE>
__global__ void __launch_bounds__(128)
k_knn_cl(ClParams P, ClKnn K)
{
    constexpr bool GRID = (MODE == 1);
    const int64_t f = blockIdx.y;
    const int *fc = P.fcounts + f * 8;
    if (fc[6]) return;                                           // out-of-box frame
    const ClFrame fr = cl_frame(P, f);
    const ClGeom &g = fr.g;
    const int *cstart = P.cell_start + f * P.cstride;
    const int n_u = cstart[2 * g.ncell];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_u) return;
    const double2 *gxy = P.sxy + f * P.N;
    const int *gidx = P.sidx + f * P.N;
    const int i = gidx[p];
    int pc = P.scell[f * P.N + p];
    const int leaf = pc >= g.ncell;
    const int cbase = leaf * g.ncell;
    pc -= cbase;
    if (MODE == 0) {
        const bool on = ((P.leaflet_mask >> leaf) & 1) && fc[leaf * 3] > 0 && fc[leaf * 3 + 1] > 0;
        if (!on || !(P.sflag[f * P.N + p] & 1u)) {
            K.n_b[f * P.N + i] = -1;
            return;
        }
    }
    const int k = K.k;
    const double apx = gxy[p].x, apy = gxy[p].y;
    const int ncx = g.ncx, ncy = g.ncy;
    const int pcy = pc / ncx, pcx = pc - pcy * ncx;
    const int lox = (ncx - 1) / 2, hix = ncx / 2, loy = (ncy - 1) / 2, hiy = ncy / 2;
    const int Rmax = max(hix, hiy);
    // the k best so far live in a binary MAX-heap on the key (distance, lipid index): its root is
    // the entry a candidate has to beat and stays in registers (r0, j0), so a rejected candidate
    // costs no local-memory access; an accepted one costs one sift (log2 k steps), not a shift of
    // a sorted array.  All keys are distinct, so the k smallest are the same set however they are
    // kept; the sorted order is produced once at the end.
    // The heap lives in SHARED memory as [slot][thread]: a thread's column sits in one bank whatever the
    // slot, so the data-dependent indexing of the sifts is conflict free (as a per-thread local array the
    // same accesses scatter over 32 cache lines per warp instruction and dominated the kernel).
    extern __shared__ __align__(16) unsigned char knn_raw[];
    double *const sbest = reinterpret_cast<double *>(knn_raw) + threadIdx.x;
    int *const sbj = reinterpret_cast<int *>(knn_raw + (size_t)KMAX * 128 * 8) + threadIdx.x;
#define best(h) sbest[(h) * 128]
#define bj(h) sbj[(h) * 128]
    int nbest = 0;
    double r0 = 0.0;
    int j0 = 0;
    auto key_gt = [](double ra, int ja, double rb, int jb) { return ra > rb || (ra == rb && ja > jb); };
    // place (r, j) into the heap of `n` entries starting at the vacant root, sifting it down
    auto sift_down = [&](double r, int j, int n) {
        int h = 0;
        for (;;) {
            int c = 2 * h + 1;
            if (c >= n) break;
            if (c + 1 < n && key_gt(best(c + 1), bj(c + 1), best(c), bj(c))) c++;
            if (!key_gt(best(c), bj(c), r, j)) break;
            best(h) = best(c);
            bj(h) = bj(c);
            h = c;
        }
        best(h) = r;
        bj(h) = j;
    };

    // float32 screen: once k candidates are held, a candidate whose float32 minimum-image d^2 exceeds the
    // k-th best by more than the float32 rounding bound (the bound of k_rdf_cl: 2^-23 (r+1)(8L+4(r+1)))
    // cannot enter the heap and costs one 8-byte load and a few float operations, no float64 at all
    const float2 *gf = P.sf + f * P.N;
    const float fpx = (float)apx, fpy = (float)apy;
    const float Lmaxf = fmaxf(fr.bxf, fr.byf);
    float screen = 3.0e38f;
    auto set_screen = [&]() {
        const float rf = (float)r0 * 1.0000002f;
        screen = rf * rf * 1.0000003f + 2.4e-7f * (rf + 1.0f) * (8.0f * Lmaxf + 4.0f * (rf + 1.0f));
    };
    auto visit = [&](int s0, int s1) {
        for (int q = s0; q < s1; q++) {
            if (q == p) continue;
            {
                const float2 fq = gf[q];
                float dx = fabsf(fpx - fq.x), dy = fabsf(fpy - fq.y);
                dx = fminf(dx, fr.bxf - dx);
                dy = fminf(dy, fr.byf - dy);
                if (fmaf(dy, dy, dx * dx) > screen) continue;
            }
            const int j = gidx[q];
            const double2 qv = gxy[q];
            const double qx = qv.x, qy = qv.y;
            const double d2 = (!GRID || j > i) ? pbk_pbc_dist2(apx, apy, qx, qy, fr.Lx, fr.Ly, fr.hx, fr.hy)
                                               : pbk_pbc_dist2(qx, qy, apx, apy, fr.Lx, fr.Ly, fr.hx, fr.hy);
            const double r = sqrt(d2);
            if (nbest == k) {
                if (!key_gt(r0, j0, r, j)) continue;          // not better than the current k-th
                sift_down(r, j, k);                           // replaces the root
            } else {
                int h = nbest++;                              // append and sift up
                while (h > 0) {
                    const int par = (h - 1) >> 1;
                    if (!key_gt(r, j, best(par), bj(par))) break;
                    best(h) = best(par);
                    bj(h) = bj(par);
                    h = par;
                }
                best(h) = r;
                bj(h) = j;
            }
            r0 = best(0);
            j0 = bj(0);
            if (nbest == k) set_screen();
        }
    };
    // cells [c0, c1] (offsets from pcx, c1 - c0 + 1 <= ncx) of row `row`
    auto visit_row = [&](int row, int c0, int c1) {
        int a = pcx + c0, b = pcx + c1;
        if (a < 0) {
            if (b < 0) { a += ncx; b += ncx; }
            else { visit(cstart[row + a + ncx], cstart[row + ncx]); a = 0; }
        }
        if (b >= ncx) {
            if (a >= ncx) { a -= ncx; b -= ncx; }
            else { visit(cstart[row], cstart[row + b - ncx + 1]); b = ncx - 1; }
        }
        visit(cstart[row + a], cstart[row + b + 1]);
    };

    for (int R = 0; R <= Rmax; R++) {
        const int dy0 = -min(R, loy), dy1 = min(R, hiy);
        for (int dy = dy0; dy <= dy1; dy++) {
            int cy = pcy + dy;
            if (cy < 0) cy += ncy;
            if (cy >= ncy) cy -= ncy;
            const int row = cbase + cy * ncx;
            if (dy == -R || dy == R) {
                visit_row(row, -min(R, lox), min(R, hix));
            } else {
                if (R <= lox) visit_row(row, -R, -R);
                if (R <= hix) visit_row(row, R, R);
            }
        }
        if (nbest == k) {
            const double bx = (R >= hix) ? 1e300 : R * g.wx, by = (R >= hiy) ? 1e300 : R * g.wy;
            const double bound = fmin(bx, by);
            if (r0 < bound * (1.0 - 1e-9) - 1e-9) break;
        }
    }
    if (GRID) {
        // ascending (distance, index) order: pop the maximum into the last free slot, repeatedly
        int32_t *o = K.knn + (f * P.N + i) * 24;
        for (int t = nbest; t < 24; t++) o[t] = -1;
        for (int n = nbest; n > 0; n--) {
            o[n - 1] = bj(0);
            if (n > 1) sift_down(best(n - 1), bj(n - 1), n - 1);
        }
    } else if (MODE == 2) {
        int h[CL_MAX_TYPES];
#pragma unroll
        for (int t = 0; t < CL_MAX_TYPES; t++) h[t] = 0;
        for (int t = 0; t < nbest; t++) {
            const int tj = P.type_code[bj(t)];
#pragma unroll
            for (int u = 0; u < CL_MAX_TYPES; u++) h[u] += (tj == u) ? 1 : 0;
        }
        int32_t *o = K.n_b + (f * P.N + i) * P.n_types;
#pragma unroll
        for (int u = 0; u < CL_MAX_TYPES; u++)
            if (u < P.n_types) o[u] = h[u];
    } else {
        int c = 0;
        for (int t = 0; t < nbest; t++) {
            const int tj = P.type_code[bj(t)];
            c += (P.tb == -1 || tj == P.tb) ? 1 : 0;
        }
        K.n_b[f * P.N + i] = c;
    }
}
#undef best
#undef bj

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct ClPlan {
    int ncap;
    int64_t cstride, per_frame_bytes, Fc;
};

static ClPlan cl_plan(int64_t F, int64_t N)
{
    ClPlan pl;
    double s = sqrt(0.5 * (double)N) + 2.0;
    if (s > 1024.0) s = 1024.0;
    pl.ncap = (int)s;
    pl.cstride = (2 * (int64_t)pl.ncap * pl.ncap + 1 + 3) & ~(int64_t)3;     // 16-byte aligned frames
    pl.per_frame_bytes = 2 * pl.cstride * 4 + 64 + N * 40;      // 8+8+8+4+4+1 rounded up for alignment
    const int64_t budget = (int64_t)768 << 20;
    pl.Fc = budget / pl.per_frame_bytes;
    if (pl.Fc < 1) pl.Fc = 1;
    if (pl.Fc > F) pl.Fc = F;
    if (pl.Fc > 65535) pl.Fc = 65535;                            // gridDim.y
    return pl;
}

static bool cl_carve(pbk_ctx *ctx, const ClPlan &pl, int64_t N, ClParams &P)
{
    const int64_t Fc = pl.Fc;
    auto up = [](int64_t v) { return (v + 255) & ~(int64_t)255; };
    const int64_t b_cells = up(Fc * pl.cstride * 4), b_fc = up(Fc * 8 * 4), b_d = up(Fc * N * 8),
                  b_i = up(Fc * N * 4), b_c = up(Fc * N);
    const int64_t b_dummy = up(Fc * 6 * 4);
    const int64_t total = 2 * b_cells + b_fc + 3 * b_d + 2 * b_i + b_c + b_dummy;
    unsigned char *a = (unsigned char *)pbk_arena(ctx, (size_t)total);
    if (!a) return false;
    P.cell_fill = (int *)a; a += b_cells;
    P.fcounts = (int *)a; a += b_fc;                             // zeroed together with cell_fill
    P.cell_start = (int *)a; a += b_cells;
    P.sxy = (double2 *)a; a += 2 * b_d;
    P.sf = (float2 *)a; a += b_d;
    P.sidx = (int *)a; a += b_i;
    P.scell = (int *)a; a += b_i;
    P.sflag = a; a += b_c;
    P.dummy_fc = (int32_t *)a;
    return true;
}

// zero the counters, count, scan, scatter for frames [0, P.F) of the chunk
static int cl_build(pbk_ctx *ctx, const ClPlan &pl, ClParams &P, cudaStream_t st)
{
    const size_t zero_bytes = (size_t)((unsigned char *)P.cell_start - (unsigned char *)P.cell_fill);
    cudaError_t e = cudaMemsetAsync(P.cell_fill, 0, zero_bytes, st);
    if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "cell list memset", e);
    dim3 g((unsigned)((P.N + CL_THREADS - 1) / CL_THREADS), (unsigned)P.F);
    k_cl_count<<<g, CL_THREADS, 0, st>>>(P);
    PBK_CHECK_LAUNCH(ctx, "k_cl_count");
    k_cl_scan<<<(unsigned)P.F, 1024, 0, st>>>(P);
    PBK_CHECK_LAUNCH(ctx, "k_cl_scan");
    k_cl_scatter<<<g, CL_THREADS, 0, st>>>(P);
    PBK_CHECK_LAUNCH(ctx, "k_cl_scatter");
    return PBK_OK;
}

// defined in pbk_rdf5.cu: the 16-bit d^2 -> (bin, distance to threshold) table
bool pbk_rdf_lut16_shape(int32_t n_bins, double range_lo, double range_hi, int *m_out, int *lut_n_out);
int pbk_rdf_lut16_build(pbk_ctx *ctx, const double *thr, int nb, int lut_n, int lut_m, unsigned short *lut,
                        cudaStream_t st);

// defined in pbk_pairs.cu / pbk_grid.cu: dense kernels restricted to flagged frames
int pbk_rdf_dense_flagged(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                          const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                          int leaflet_mask, int32_t n_bins, double range_lo, double range_hi,
                          const double *edges, unsigned long long *count, int32_t *frame_counts,
                          const int *frame_flag, int flag_stride, cudaStream_t stream);
int pbk_nnf_dense_flagged(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                          const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                          int leaflet_mask, int k, int32_t *n_b, const int *frame_flag, int flag_stride,
                          cudaStream_t stream);
int pbk_grid_knn_dense_flagged(pbk_ctx *ctx, const double *com, const uint8_t *labels, const float *box,
                               int64_t F, int64_t N, int lat0, int lat1, int32_t *knn, const int *frame_flag,
                               int flag_stride, cudaStream_t stream);
int pbk_nnf_dense_strided(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                          const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                          int leaflet_mask, int k, int32_t *n_b, int out_stride, const int *frame_flag,
                          int flag_stride, cudaStream_t stream);

// n_types > 0: multi mode (count[2][T][T][n_bins], type_counts[F][2][T]; type_a/type_b/leaflet_mask unused)
static int cl_rdf_run(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                      const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                      int leaflet_mask, int32_t n_bins, double range_lo, double range_hi, const double *edges,
                      const double *thr, unsigned long long *count, int32_t *frame_counts, int n_types,
                      int32_t *type_counts, cudaStream_t st)
{
    const bool multi = n_types > 0;
    const int T = multi ? n_types : 1;
    if (multi) { type_a = type_b = -1; leaflet_mask = 3; }
    const size_t hsize = multi ? (size_t)2 * T * T * n_bins : (size_t)n_bins;
    if (!thr || !ctx->work || n_bins > 255 || range_lo < 0 || N > 0x7fffff00LL) return PBK_EUNSUPPORTED;
    int m = 0, lut_n = 0;
    if (!ctx->lut || n_bins > 254 || !pbk_rdf_lut16_shape(n_bins, range_lo, range_hi, &m, &lut_n)) return PBK_EUNSUPPORTED;
    const unsigned lslot = ctx->work_next++;
    unsigned short *lut = ctx->lut + (size_t)(lslot % PBK_LUT_SLOTS) * PBK_LUT_ENTRIES;
    {
        const int rc = pbk_rdf_lut16_build(ctx, thr, n_bins, lut_n, m, lut, st);
        if (rc) return rc;
    }
    ClRdf R;
    R.nb = n_bins; R.lut_n = lut_n; R.lut_m = m; R.thr = thr; R.lut = lut;
    R.count = count; R.chunk = 2048;
    R.hist_copies = CL_THREADS / 32;
    auto smem_for = [&](int c) {
        return (size_t)(n_bins + 1) * 8 + (size_t)(CL_THREADS / 32) * CL_QCAP * 8 + (size_t)c * hsize * 4 +
               (size_t)lut_n * 2 + 16;
    };
    while (R.hist_copies > 1 && smem_for(R.hist_copies) > 48 * 1024) R.hist_copies >>= 1;
    const size_t smem = smem_for(R.hist_copies);
    if (smem > 100 * 1024) return PBK_EUNSUPPORTED;
    const bool same = (type_a == type_b);
    auto kern = multi ? k_rdf_cl<2> : (same ? k_rdf_cl<0> : k_rdf_cl<1>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "k_rdf_cl smem attribute", e);
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CL_THREADS, smem);
    if (e != cudaSuccess || per_sm < 1) per_sm = 1;

    ClPlan pl = cl_plan(F, N);
    ClParams P;
    P.N = N; P.lat0 = lat0; P.lat1 = lat1; P.ta = type_a; P.tb = type_b; P.leaflet_mask = leaflet_mask;
    P.members_all = 0; P.wmin = 0.51 * range_hi + 1e-6; P.range_hi = range_hi; P.occ = 0.0;
    P.ncap = pl.ncap; P.cstride = pl.cstride; P.type_code = type_code;
    P.multi = multi ? 1 : 0; P.n_types = T; P.type_counts = nullptr;
    if (multi) P.members_all = 1;
    if (!cl_carve(ctx, pl, N, P)) return pbk_fail(ctx, PBK_ECUDA, "pbk_rdf: scratch arena");
    for (int64_t f0 = 0; f0 < F; f0 += pl.Fc) {
        const int64_t Fc = (F - f0 < pl.Fc) ? F - f0 : pl.Fc;
        P.F = Fc;
        P.com = com + f0 * N * 3; P.labels = labels + f0 * N; P.box = box + f0 * 3;
        P.frame_counts = multi ? nullptr : frame_counts + f0 * 6;
        P.type_counts = multi ? type_counts + f0 * 2 * T : nullptr;
        int rc = cl_build(ctx, pl, P, st);
        if (rc) return rc;
        R.work = ctx->work + (ctx->work_next++ % PBK_WORK_SLOTS);
        e = cudaMemsetAsync(R.work, 0, sizeof(int), st);
        if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "k_rdf_cl work counter", e);
        const int64_t items = Fc * ((N + R.chunk - 1) / R.chunk);
        int64_t grid = (int64_t)ctx->sm_count * per_sm;
        if (grid > items) grid = items;
        kern<<<(unsigned)grid, CL_THREADS, smem, st>>>(P, R);
        PBK_CHECK_LAUNCH(ctx, "k_rdf_cl");
        if (!multi) {
            rc = pbk_rdf_dense_flagged(ctx, P.com, P.labels, type_code, P.box, Fc, N, lat0, lat1, type_a, type_b,
                                       leaflet_mask, n_bins, range_lo, range_hi, edges, count, P.frame_counts,
                                       P.fcounts + 6, 8, st);
            if (rc) return rc;
        } else {
            // out-of-box frames: the dense kernel per (leaflet, type pair); its per-frame counts go
            // to the scratch rows behind the cell data (unused here)
            for (int l = 0; l < 2; l++)
                for (int ta_ = 0; ta_ < T; ta_++)
                    for (int tb_ = 0; tb_ < T; tb_++) {
                        rc = pbk_rdf_dense_flagged(ctx, P.com, P.labels, type_code, P.box, Fc, N, lat0, lat1, ta_, tb_,
                                                   1 << l, n_bins, range_lo, range_hi, edges,
                                                   count + (size_t)((l * T + ta_) * T + tb_) * n_bins, P.dummy_fc,
                                                   P.fcounts + 6, 8, st);
                        if (rc) return rc;
                    }
        }
    }
    return PBK_OK;
}

int pbk_rdf_cells(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                  const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                  int leaflet_mask, int32_t n_bins, double range_lo, double range_hi, const double *edges,
                  const double *thr, unsigned long long *count, int32_t *frame_counts, cudaStream_t st)
{
    return cl_rdf_run(ctx, com, labels, type_code, box, F, N, lat0, lat1, type_a, type_b, leaflet_mask, n_bins,
                      range_lo, range_hi, edges, thr, count, frame_counts, 0, nullptr, st);
}

extern "C" int pbk_rdf_multi(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                             const float *box, int64_t F, int64_t N, int lat0, int lat1, int n_types,
                             int32_t n_bins, double range_lo, double range_hi, const double *edges,
                             const double *thr, unsigned long long *count, int32_t *type_counts, void *stream)
{
    if (!ctx || !com || !labels || !type_code || !box || !edges || !thr || !count || !type_counts || F < 0 ||
        N <= 0 || n_bins <= 0 || !(range_hi > range_lo) || n_types < 1 || lat0 < 0 || lat0 > 2 || lat1 < 0 ||
        lat1 > 2)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_rdf_multi: bad argument");
    if (n_types > CL_MAX_TYPES) return PBK_EUNSUPPORTED;
    if (F == 0) return PBK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(type_counts, 0, (size_t)F * 2 * n_types * sizeof(int32_t), st);
    if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "pbk_rdf_multi memset", e);
    return cl_rdf_run(ctx, com, labels, type_code, box, F, N, lat0, lat1, -1, -1, 3, n_bins, range_lo, range_hi,
                      edges, thr, count, nullptr, n_types, type_counts, st);
}

template <int MODE>
static int cl_knn_run(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                      const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                      int leaflet_mask, int k, int32_t *n_b, int32_t *knn, int n_types, int32_t *type_counts,
                      cudaStream_t st)
{
    constexpr bool GRID = (MODE == 1);
    if (N > 0x7fffff00LL) return PBK_EUNSUPPORTED;
    ClPlan pl = cl_plan(F, N);
    ClParams P;
    P.N = N; P.lat0 = lat0; P.lat1 = lat1; P.ta = type_a; P.tb = type_b; P.leaflet_mask = leaflet_mask;
    P.members_all = 1; P.wmin = 0.0; P.range_hi = 0.0; P.occ = 4.0;
    P.ncap = pl.ncap; P.cstride = pl.cstride; P.type_code = type_code; P.frame_counts = nullptr;
    P.multi = (MODE == 2) ? 1 : 0; P.n_types = (MODE == 2) ? n_types : 1; P.type_counts = nullptr;
    if (!cl_carve(ctx, pl, N, P)) return pbk_fail(ctx, PBK_ECUDA, "pbk knn: scratch arena");
    ClKnn K;
    K.k = k;
    for (int64_t f0 = 0; f0 < F; f0 += pl.Fc) {
        const int64_t Fc = (F - f0 < pl.Fc) ? F - f0 : pl.Fc;
        P.F = Fc;
        P.com = com + f0 * N * 3; P.labels = labels + f0 * N; P.box = box + f0 * 3;
        K.n_b = n_b ? n_b + f0 * N * (MODE == 2 ? n_types : 1) : nullptr;
        P.type_counts = (MODE == 2) ? type_counts + f0 * 2 * n_types : nullptr;
        K.knn = knn ? knn + f0 * N * 24 : nullptr;
        int rc = cl_build(ctx, pl, P, st);
        if (rc) return rc;
        dim3 g((unsigned)((N + 127) / 128), (unsigned)Fc);
        if (GRID) k_knn_cl<24, 1><<<g, 128, 24 * 128 * 12, st>>>(P, K);
        else if (k <= 8) k_knn_cl<8, MODE><<<g, 128, 8 * 128 * 12, st>>>(P, K);
        else if (k <= 24) k_knn_cl<24, MODE><<<g, 128, 24 * 128 * 12, st>>>(P, K);
        else {
            cudaFuncSetAttribute(k_knn_cl<64, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 128 * 12);
            k_knn_cl<64, MODE><<<g, 128, 64 * 128 * 12, st>>>(P, K);
        }
        PBK_CHECK_LAUNCH(ctx, "k_knn_cl");
        if (GRID)
            rc = pbk_grid_knn_dense_flagged(ctx, P.com, P.labels, P.box, Fc, N, lat0, lat1, K.knn,
                                            P.fcounts + 6, 8, st);
        else if (MODE == 2) {
            // out-of-box frames: the dense kernel once per type (every lipid a reference lipid)
            for (int t = 0; t < n_types && !rc; t++)
                rc = pbk_nnf_dense_strided(ctx, P.com, P.labels, type_code, P.box, Fc, N, lat0, lat1, -1, t, 3, k,
                                           K.n_b + t, n_types, P.fcounts + 6, 8, st);
        } else
            rc = pbk_nnf_dense_flagged(ctx, P.com, P.labels, type_code, P.box, Fc, N, lat0, lat1, type_a,
                                       type_b, leaflet_mask, k, K.n_b, P.fcounts + 6, 8, st);
        if (rc) return rc;
    }
    return PBK_OK;
}

int pbk_nnf_cells(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                  const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                  int leaflet_mask, int k, int32_t *n_b, cudaStream_t st)
{
    return cl_knn_run<0>(ctx, com, labels, type_code, box, F, N, lat0, lat1, type_a, type_b, leaflet_mask,
                         k, n_b, nullptr, 0, nullptr, st);
}

int pbk_grid_knn_cells(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code_or_null,
                       const float *box, int64_t F, int64_t N, int lat0, int lat1, int32_t *knn, cudaStream_t st)
{
    (void)type_code_or_null;
    return cl_knn_run<1>(ctx, com, labels, nullptr, box, F, N, lat0, lat1, -1, -1, 3, 24, nullptr, knn, 0, nullptr, st);
}

// which systems take the cell-list kernels: N above the entry point's threshold, or every
// system when PBK_FORCE_CELLS is set (tests), none when PBK_NO_CELLS is set.
bool pbk_use_cells(int64_t N, int64_t threshold)
{
    if (getenv("PBK_NO_CELLS")) return false;
    if (getenv("PBK_FORCE_CELLS")) return true;
    return N > threshold;
}

// ---- D3, many analyses at once: the k nearest neighbours found once, their types tallied ----
extern "C" int pbk_nnf_multi(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                             const float *box, int64_t F, int64_t N, int lat0, int lat1, int n_types, int k,
                             int32_t *type_hits, int32_t *type_counts, void *stream)
{
    if (!ctx || !com || !labels || !type_code || !box || !type_hits || !type_counts || F < 0 || N <= 0 ||
        n_types < 1 || k < 1 || k > 64 || lat0 < 0 || lat0 > 2 || lat1 < 0 || lat1 > 2)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_nnf_multi: bad argument (1 <= k <= 64)");
    if (n_types > CL_MAX_TYPES) return PBK_EUNSUPPORTED;
    if (F == 0) return PBK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(type_counts, 0, (size_t)F * 2 * n_types * sizeof(int32_t), st);
    if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "pbk_nnf_multi memset", e);
    return cl_knn_run<2>(ctx, com, labels, type_code, box, F, N, lat0, lat1, -1, -1, 3, k, type_hits, nullptr,
                         n_types, type_counts, st);
}
