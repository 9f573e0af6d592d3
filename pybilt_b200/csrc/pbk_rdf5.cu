// pbk_rdf5.cu -- COM lateral RDF for small leaflets with a Verlet pair list per CTA.
//
// Same arithmetic contract as k_rdf3 (pbk_rdf3.cu), restating
//   pybilt/bilayer_analyzer/analysis_protocols.py:4900-5006  (COMLateralRDFProtocol)
//   pybilt/common/distance_cutoff_clustering.py:27-71        (distance_euclidean_pbc)
// One 256-thread CTA owns one leaflet over a CHUNK of consecutive frames (about 30 KB of shared
// memory, so seven CTAs share an SM).  Lipids move a fraction of an Angstrom per frame, so the
// cell sort and the half-shell sweep -- two thirds of k_rdf3's instructions -- are not repeated
// every frame: the CTA sweeps once with the cutoff widened by a skin, keeps the unordered pairs
// it found as a flat list in shared memory (two 8-bit slots per pair), and for the following
// frames only refreshes the float32 coordinates and walks the list, four independent pairs per
// thread per iteration.  The list stays valid for every lipid that has not moved further than
// skin/2 (minimum image) from where it was when the list was built.  The few that have --
// mostly lipids whose WRAPPED centre of mass jumps because one of their beads crossed the box
// edge (com_frame.py:71 averages the raw wrapped atoms) -- are taken out of the list walk by
// poisoning their float32 slot and are paired with every other member directly.  The list is
// rebuilt when more than 32 lipids have left their skin, when the box differs by a bit or when
// the leaflet's member set changes.  Anything the list cannot hold (more than 256 members,
// overflow, coordinates outside the box) is binned by the direct sweep of k_rdf3, so every
// input gives the reference's counts.
// A pair's bin comes from its float32 d^2 through the same look-up table with per-entry
// distance to the nearest threshold; pairs within the float32 rounding bound of a threshold
// are evaluated with the exact float64 formulas in both orders (see pbk_rdf3.cu).
#include <stdlib.h>

#include "pbk_common.cuh"

#define R5_THREADS 256
#define R5_WARPS (R5_THREADS / 32)
#define R5_MAXDIM 16
#define R5_MAXCELLS (R5_MAXDIM * R5_MAXDIM)
#define R5_MAXROWS 8
#define R5_QCAP 64
#define R5_MAXN 640

struct Rdf5Params {
    const double *com;
    const uint8_t *labels;
    const int32_t *type_code;
    const float *box;
    int64_t F, N;
    int lat0, lat1, ta, tb, leaflet_mask, nb;
    double range_hi;
    float skin;                     // list radius = range_hi + skin (+ margin)
    const double *thr;
    unsigned long long *count;
    int32_t *frame_counts;
    int lut_n, lut_m;
    int chunk;                      // frames per work item
    int list_cap;                   // pair-list entries per warp
    int *work;
    const unsigned short *lut;      // [lut_n] global (L1-resident): bin | distance-to-threshold << 8
};

#define R5_MAXJ 32                  // lipids handled outside the list per frame
struct R5Sh {                       // the CTA's shared memory
    float2 *fxy;                    // N (sorted): float32 b' = x - box/2 of the current frame
    float2 *ref;                    // N (sorted): the same when the list was built
    unsigned short *list;           // list_cap: p | q << 8
    unsigned *hist;                 // this warp's nb counters
    unsigned *q;                    // this warp's R5_QCAP exact-path queue
    int *cell_start, *cell_fill;    // R5_MAXCELLS + 1 each
    unsigned short *idx;            // N (sorted): lipid of the slot
    unsigned short *cid;            // N (by lipid): cell at build time, 0xFFFF = not a member
    unsigned char *flag;            // N (sorted): bit0 type A, bit1 type B
    unsigned short *jslot;          // R5_MAXJ slots out of their skin this frame
    float2 *jxy;                    // R5_MAXJ their float32 coordinates
    int *cnt;                       // 8 counters: cA, cB, cL, n_jump, n_list, item, n_u, -
};

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

// the reference's float64 arithmetic for the unordered pair of sorted slots (p, q)
template <bool SAME>
__device__ __noinline__ void r5_exact(unsigned entry, const double *cf, int lat0, int lat1,
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

// per-frame constants of the float32 fast path
struct R5Frame {
    const double *cf;
    float bxf, byf;
    double Lx, Ly, hx, hy;
    int tolc;
    bool all_exact;
};

// queue the pairs flagged `amb` for the exact evaluation (all 32 lanes call this together)
template <bool SAME>
__device__ __forceinline__ void r5_push(bool amb, int p, int q, int &qn, int lane, const R5Sh &W,
                                        const R5Frame &Fr, const Rdf5Params &P, const double *s_thr)
{
    const unsigned m = __ballot_sync(0xffffffffu, amb);
    if (!m) return;
    if (amb) W.q[qn + __popc(m & ((1u << lane) - 1u))] = (unsigned)p | ((unsigned)q << 16);
    qn += __popc(m);
    __syncwarp();
    if (qn >= 32) {
        r5_exact<SAME>(W.q[lane], Fr.cf, P.lat0, P.lat1, W.idx, W.flag, s_thr, P.nb, W.hist, Fr.Lx, Fr.Ly, Fr.hx, Fr.hy);
        __syncwarp();
        const unsigned moved = lane + 32 < qn ? W.q[lane + 32] : 0u;
        __syncwarp();
        if (lane + 32 < qn) W.q[lane] = moved;
        qn -= 32;
        __syncwarp();
    }
}

// float32 minimum-image d^2 of two slots and its table entry
__device__ __forceinline__ unsigned r5_lookup(float2 fp, float2 fq, const R5Frame &Fr, float lut_scale, int lut_last,
                                              const unsigned short *s_lut, float *d2_out)
{
    float dx = fabsf(fp.x - fq.x), dy = fabsf(fp.y - fq.y);
    dx = fminf(dx, Fr.bxf - dx);
    dy = fminf(dy, Fr.byf - dy);
    const float d2 = fmaf(dy, dy, dx * dx);
    *d2_out = d2;
    const int c = min(__float2int_rd(d2 * lut_scale), lut_last);
    return __ldg(s_lut + c);
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

// one candidate pair in direct / jumper mode: count it or queue it for the exact path
template <bool SAME>
__device__ __forceinline__ void r5_pair(bool act, int p, int q, float2 fp, float2 fq, unsigned fpflag, int &qn, int lane,
                                        const R5Sh &W, const R5Frame &Fr, const Rdf5Params &P, const double *s_thr,
                                        float lut_scale, int lut_last)
{
    float d2;
    const unsigned e = r5_lookup(fp, fq, Fr, lut_scale, lut_last, P.lut, &d2);
    unsigned w = 2u;
    if (!SAME) {
        const unsigned fqf = W.flag[q];
        w = ((fpflag & 1u) & ((fqf >> 1) & 1u)) + ((fqf & 1u) & ((fpflag >> 1) & 1u));
    }
    const bool sure = (int)(e >> 8) >= Fr.tolc && !Fr.all_exact;
    if (act && sure && (e & 0xFFu) != 0xFFu && w) atomicAdd(&W.hist[e & 0xFFu], w);
    r5_push<SAME>(act && !sure && w, p, q, qn, lane, W, Fr, P, s_thr);
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
    W.ref = reinterpret_cast<float2 *>(wb);                                wb += npad * 8;
    W.jxy = reinterpret_cast<float2 *>(wb);                                wb += (size_t)R5_MAXJ * 8;
    W.list = reinterpret_cast<unsigned short *>(wb);                       wb += ((size_t)P.list_cap * 2 + 15) & ~(size_t)15;
    unsigned *hist_all = reinterpret_cast<unsigned *>(wb);                 wb += (((size_t)R5_WARPS * nb * 4) + 15) & ~(size_t)15;
    W.hist = hist_all + (size_t)warp * nb;
    W.q = reinterpret_cast<unsigned *>(wb) + (size_t)warp * R5_QCAP;       wb += (size_t)R5_WARPS * R5_QCAP * 4;
    W.cell_start = reinterpret_cast<int *>(wb);                            wb += (size_t)(R5_MAXCELLS + 1) * 4;
    W.cell_fill = reinterpret_cast<int *>(wb);                             wb += (size_t)(R5_MAXCELLS + 1) * 4;
    W.cnt = reinterpret_cast<int *>(wb);                                   wb += 8 * 4;
    W.idx = reinterpret_cast<unsigned short *>(wb);                        wb += npad * 2;
    W.cid = reinterpret_cast<unsigned short *>(wb);                        wb += npad * 2;
    W.jslot = reinterpret_cast<unsigned short *>(wb);                      wb += (size_t)R5_MAXJ * 2;
    W.flag = reinterpret_cast<unsigned char *>(wb);

    for (int q = tid; q <= nb; q += R5_THREADS) s_thr[q] = P.thr[q];
    for (int q = tid; q < R5_WARPS * nb; q += R5_THREADS) hist_all[q] = 0u;
    __syncthreads();
    const unsigned lt_mask = (1u << lane) - 1u;
    const float lut_scale = ldexpf(1.0f, P.lut_m);
    const int lut_last = P.lut_n - 1;
    const int64_t chunks = (P.F + P.chunk - 1) / P.chunk;
    const int64_t n_items = chunks * 2;
    const float Rb = (float)P.range_hi + P.skin + 0.05f;          // list radius
    const float Rb2 = Rb * Rb;
    const float lim2 = 0.25f * P.skin * P.skin * 0.98f;           // (skin/2)^2 with a little slack

    for (;;) {
        if (tid == 0) W.cnt[5] = atomicAdd(P.work, 1);
        __syncthreads();
        const long long item = W.cnt[5];
        if (item >= n_items) break;
        const int leaf = (int)(item & 1);            // 0 = upper (label 1), 1 = lower (label 0)
        const int64_t fa = (item >> 1) * P.chunk, fb = min(P.F, fa + (int64_t)P.chunk);
        const unsigned want = leaf == 0 ? 1u : 0u;
        const bool selected = (P.leaflet_mask >> leaf) & 1;
        bool have_list = false;                      // CTA uniform, like everything that steers the flow below
        int n_list = 0, n_u = 0;
        float bx0 = 0.f, by0 = 0.f;
        int ncx = 1, ncy = 1, mx = 0, my = 0, ncell = 1;

        for (int64_t f = fa; f < fb; f++) {
            R5Frame Fr;
            Fr.cf = P.com + f * P.N * 3;
            const uint8_t *lab = P.labels + f * P.N;
            Fr.bxf = __ldg(P.box + f * 3 + P.lat0); Fr.byf = __ldg(P.box + f * 3 + P.lat1);
            const float cxf = Fr.bxf * 0.5f, cyf = Fr.byf * 0.5f;
            Fr.Lx = (double)Fr.bxf; Fr.Ly = (double)Fr.byf; Fr.hx = (double)cxf; Fr.hy = (double)cyf;
            // rounding bound of the float32 d^2 near the thresholds, in table cells (+1 cell of slack):
            // |d2_f32 - d2_exact| <= 2^-23 (r+1) (8 L + 4 (r+1))  with r = range_hi, L = larger box edge
            const float Lm = fmaxf(Fr.bxf, Fr.byf), rr = (float)P.range_hi + 1.0f;
            const float tol = 1.1920929e-7f * rr * (8.0f * Lm + 4.0f * rr);
            Fr.tolc = (int)ceilf(tol * lut_scale) + 1;
            Fr.all_exact = !(tol * lut_scale < 200.0f) || !(Fr.bxf > 0.0f) || !(Fr.byf > 0.0f);
            int qn = 0;                                           // this warp's queued exact pairs

            // ---- member set and counts of the frame ----------------------------------------------
            if (tid < 5) W.cnt[tid] = 0;
            __syncthreads();
            int cA = 0, cB = 0, cL = 0, changed = 0;
            for (int i = tid; i < N; i += R5_THREADS) {
                bool m2 = false;
                if (lab[i] == want) {
                    const int t = __ldg(P.type_code + i);
                    const bool a = (P.ta == -1 || t == P.ta), b = (P.tb == -1 || t == P.tb);
                    cL++; cA += a; cB += b;
                    m2 = a || b;
                }
                if (have_list) changed |= (m2 != (W.cid[i] != 0xFFFF));
            }
            cA = pbk_warp_sum(cA); cB = pbk_warp_sum(cB); cL = pbk_warp_sum(cL);
            if (lane == 0) { atomicAdd(&W.cnt[0], cA); atomicAdd(&W.cnt[1], cB); atomicAdd(&W.cnt[2], cL); }
            const int any_changed = __syncthreads_or(changed);
            cA = W.cnt[0]; cB = W.cnt[1]; cL = W.cnt[2];
            const bool on = selected && cA > 0 && cB > 0;
            if (tid < 3) {
                int v = tid == 0 ? cA : (tid == 1 ? cB : cL);
                if (tid < 2 && !on) v = 0;
                if (tid == 2 && !selected) v = 0;
                P.frame_counts[f * 6 + leaf * 3 + tid] = v;
            }
            if (!on) { have_list = false; __syncthreads(); continue; }
            if (have_list && (any_changed || Fr.bxf != bx0 || Fr.byf != by0 || Fr.all_exact)) have_list = false;

            // ---- list still valid?  refresh the coordinates; lipids out of their skin leave the list ----
            int nj = 0;
            if (have_list) {
                int oob = 0;
                for (int s = tid; s < n_u; s += R5_THREADS) {
                    const int i = W.idx[s];
                    const double x = Fr.cf[i * 3 + P.lat0], y = Fr.cf[i * 3 + P.lat1];
                    oob |= !(x >= 0.0 && x <= Fr.Lx && y >= 0.0 && y <= Fr.Ly);
                    float2 c = make_float2((float)__dsub_rn(x, Fr.hx), (float)__dsub_rn(y, Fr.hy));
                    const float2 r = W.ref[s];
                    float dx = fabsf(c.x - r.x), dy = fabsf(c.y - r.y);
                    dx = fminf(dx, Fr.bxf - dx);
                    dy = fminf(dy, Fr.byf - dy);
                    if (!(fmaf(dy, dy, dx * dx) < lim2)) {
                        const int k = atomicAdd(&W.cnt[3], 1);
                        if (k < R5_MAXJ) { W.jslot[k] = (unsigned short)s; W.jxy[k] = c; }
                        // poisoned slot: every list pair that touches it lands beyond the table
                        c = make_float2(1.0e30f * (float)(s + 1), 1.0e30f * (float)(s + 1));
                    }
                    W.fxy[s] = c;
                }
                const int any_oob = __syncthreads_or(oob);
                nj = W.cnt[3];
                if (any_oob || nj > R5_MAXJ) { have_list = false; nj = 0; }
            }

            // ---- (re)build: cell sort of the current frame, then one half-shell sweep that either
            // fills the pair list (radius Rb) or, when a list is not possible, bins directly --------
            if (!have_list) {
                // cell grid: edge >= 0.51 * Rb so that the half shell reaches 2 cells
                const double wmin = 0.51 * (double)Rb + 1e-6;
                ncx = (int)floor(Fr.Lx / wmin); ncy = (int)floor(Fr.Ly / wmin);
                ncx = max(1, min(R5_MAXDIM, ncx));
                ncy = max(1, min(R5_MAXDIM, ncy));
                const double wx = Fr.Lx / ncx, wy = Fr.Ly / ncy;
                mx = (int)floor(((double)Rb + 1e-6) / wx) + 1; my = (int)floor(((double)Rb + 1e-6) / wy) + 1;
                if (2 * mx + 1 > ncx) { ncx = 1; mx = 0; }
                if (2 * my + 1 > ncy || my + 1 > R5_MAXROWS) { ncy = 1; my = 0; }
                const double iwx = ncx / Fr.Lx, iwy = ncy / Fr.Ly;
                ncell = ncx * ncy;
                int oob = 0;
                for (int c = tid; c <= R5_MAXCELLS; c += R5_THREADS) W.cell_fill[c] = 0;
                for (int i = tid; i < N; i += R5_THREADS) {
                    unsigned short cid = 0xFFFF;
                    if (lab[i] == want) {
                        const int t = __ldg(P.type_code + i);
                        const bool a = (P.ta == -1 || t == P.ta), b = (P.tb == -1 || t == P.tb);
                        if (a || b) {
                            const double x = Fr.cf[i * 3 + P.lat0], y = Fr.cf[i * 3 + P.lat1];
                            oob |= !(x >= 0.0 && x <= Fr.Lx && y >= 0.0 && y <= Fr.Ly);
                            int cx = (int)(x * iwx), cy = (int)(y * iwy);
                            cx = max(0, min(ncx - 1, cx));
                            cy = max(0, min(ncy - 1, cy));
                            cid = (unsigned short)(cy * ncx + cx);
                        }
                    }
                    W.cid[i] = cid;
                }
                bool direct = Fr.all_exact;
                if (__syncthreads_or(oob)) {
                    // a member outside the box: one cell, every pair through the exact formulas
                    Fr.all_exact = true;
                    direct = true;
                    ncx = ncy = 1; mx = my = 0; ncell = 1;
                    for (int i = tid; i < N; i += R5_THREADS) if (W.cid[i] != 0xFFFF) W.cid[i] = 0;
                    __syncthreads();
                }
                for (int i = tid; i < N; i += R5_THREADS) {
                    const unsigned short cid = W.cid[i];
                    if (cid != 0xFFFF) atomicAdd(&W.cell_fill[cid], 1);
                }
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
                n_u = W.cell_start[ncell];
                for (int i = tid; i < N; i += R5_THREADS) {
                    const unsigned short cid = W.cid[i];
                    if (cid == 0xFFFF) continue;
                    const int slot = atomicAdd(&W.cell_fill[cid], 1);
                    const float2 c = make_float2((float)__dsub_rn(Fr.cf[i * 3 + P.lat0], Fr.hx),
                                                 (float)__dsub_rn(Fr.cf[i * 3 + P.lat1], Fr.hy));
                    W.fxy[slot] = c;
                    W.ref[slot] = c;
                    W.idx[slot] = (unsigned short)i;
                    if (!SAME) {
                        const int t = __ldg(P.type_code + i);
                        W.flag[slot] = (unsigned char)(((P.ta == -1 || t == P.ta) ? 1 : 0) | ((P.tb == -1 || t == P.tb) ? 2 : 0));
                    }
                }
                __syncthreads();
                if (n_u > 256) direct = true;               // slots do not fit the 8-bit list entries

                // the sweep: thread = lipid p, forward half shell of cells.  emit = fill the list.
                for (int attempt = 0; attempt < 2; attempt++) {
                    const bool emit = !direct;
                    for (int p0 = warp * 32; p0 < n_u; p0 += R5_THREADS) {
                        const int p = p0 + lane;
                        const bool pact = p < n_u;
                        float2 fp = make_float2(0.f, 0.f);
                        int pcx = 0, pcy = 0;
                        unsigned fpflag = 3u;
                        if (pact) {
                            fp = W.fxy[p];
                            const int pc = W.cid[W.idx[p]];
                            pcy = pc / ncx; pcx = pc - pcy * ncx;
                            if (!SAME) fpflag = W.flag[p];
                        }
                        for (int oy = 0; oy <= my; oy++) {
                            int b1 = 0, n1 = 0, b2 = 0, n2 = 0;
                            if (pact) {
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
                                if (emit) {
                                    float d2;
                                    r5_lookup(fp, W.fxy[q], Fr, lut_scale, 0, P.lut, &d2);
                                    const bool hit = act && d2 < Rb2;
                                    const unsigned m = __ballot_sync(0xffffffffu, hit);
                                    if (m) {
                                        int base = 0;
                                        if (lane == 0) base = atomicAdd(&W.cnt[4], __popc(m));
                                        base = __shfl_sync(0xffffffffu, base, 0);
                                        const int pos = base + __popc(m & lt_mask);
                                        if (hit && pos < P.list_cap) W.list[pos] = (unsigned short)(p | (q << 8));
                                    }
                                } else {
                                    r5_pair<SAME>(act, p, q, fp, W.fxy[q], fpflag, qn, lane, W, Fr, P, s_thr, lut_scale, lut_last);
                                }
                            }
                        }
                    }
                    __syncthreads();
                    n_list = W.cnt[4];
                    const bool overflow = emit && n_list > P.list_cap;
                    if (emit && !overflow) { have_list = true; bx0 = Fr.bxf; by0 = Fr.byf; }
                    if (!overflow) break;
                    direct = true;                              // list too small: bin this frame directly
                }
            }

            if (have_list) {
                // ---- bin the frame from the pair list: four independent pairs per thread per iteration ----
                for (int e0 = 0; e0 < n_list; e0 += 4 * R5_THREADS) {
                    int pp[4], qq[4];
                    bool amb[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int k = e0 + R5_THREADS * j + tid;
                        const bool act = k < n_list;
                        const unsigned ent = act ? W.list[k] : 0u;
                        pp[j] = ent & 0xFFu; qq[j] = ent >> 8;
                        float d2;
                        const unsigned e = r5_lookup(W.fxy[pp[j]], W.fxy[qq[j]], Fr, lut_scale, lut_last, P.lut, &d2);
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
                        for (int j = 0; j < 4; j++) r5_push<SAME>(amb[j], pp[j], qq[j], qn, lane, W, Fr, P, s_thr);
                    }
                }
                // ---- lipids out of their skin: paired with every other member directly (their slots
                // are poisoned, so two of them never meet here) and with each other below --------------
                for (int jj = 0; jj < nj; jj++) {
                    const int sj = W.jslot[jj];
                    const float2 cj = W.jxy[jj];
                    unsigned fjflag = 3u;
                    if (!SAME) fjflag = W.flag[sj];
                    for (int s0 = warp * 32; s0 < n_u; s0 += R5_THREADS) {
                        const int s = s0 + lane;
                        const bool act = s < n_u && s != sj;
                        r5_pair<SAME>(act, sj, act ? s : 0, cj, W.fxy[act ? s : 0], fjflag, qn, lane, W, Fr, P, s_thr, lut_scale, lut_last);
                    }
                }
                for (int t0 = warp * 32; t0 < nj * nj; t0 += R5_THREADS) {
                    const int t = t0 + lane;
                    const int a = t / max(nj, 1), b = t - a * max(nj, 1);
                    const bool act = t < nj * nj && a < b;
                    const int sa = W.jslot[act ? a : 0], sb = W.jslot[act ? b : 0];
                    unsigned faflag = 3u;
                    if (!SAME) faflag = W.flag[sa];
                    r5_pair<SAME>(act, sa, sb, W.jxy[act ? a : 0], W.jxy[act ? b : 0], faflag, qn, lane, W, Fr, P, s_thr, lut_scale, lut_last);
                }
            }
            __syncwarp();
            if (lane < qn)
                r5_exact<SAME>(W.q[lane], Fr.cf, P.lat0, P.lat1, W.idx, W.flag, s_thr, nb, W.hist, Fr.Lx, Fr.Ly, Fr.hx, Fr.hy);
            __syncthreads();
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
    return ((((size_t)(nb + 1) * 8) + 15) & ~(size_t)15) + npad * 16 + (size_t)R5_MAXJ * 8 + (((size_t)list_cap * 2 + 15) & ~(size_t)15) +
           ((((size_t)R5_WARPS * nb * 4) + 15) & ~(size_t)15) + (size_t)R5_WARPS * R5_QCAP * 4 + (size_t)(2 * R5_MAXCELLS + 2) * 4 + 32 +
           npad * 4 + (size_t)R5_MAXJ * 2 + npad + 16;
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
    // d^2 table: cells of width 2^-m, up to 8 per narrowest threshold gap (w^2 for uniform bins of
    // width w starting at lo >= 0), capped at PBK_LUT_ENTRIES (a cell wider than a gap is still
    // handled correctly -- it is marked ambiguous)
    const double w = (range_hi - range_lo) / n_bins;
    const double lo = range_lo > 0 ? range_lo : 0.0;
    const double min_gap = (lo + w) * (lo + w) - lo * lo;
    int m = 0;
    while (ldexp(1.0, -m) > min_gap / 8.0 && m < 30) m++;
    while (m > -20 && (range_hi * range_hi + 2.0) * ldexp(1.0, m) + 520.0 > (double)PBK_LUT_ENTRIES) m--;
    const double lut_nd = floor((range_hi * range_hi + 1.0) * ldexp(1.0, m)) + 516.0;
    if (lut_nd > (double)PBK_LUT_ENTRIES || lut_nd < 8.0) return PBK_EUNSUPPORTED;
    const int lut_n = (int)lut_nd;
    float skin = 4.0f;
    if (const char *ev = getenv("PBK_RDF_SKIN")) { const float v = (float)atof(ev); if (v > 0.0f && v < 100.0f) skin = v; }
    // pair list: room for a leaflet of 256 lipids at 64 A^2 each within range_hi + skin, plus a third
    const double rb = range_hi + skin + 0.05;
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
    if (const char *ev = getenv("PBK_RDF_CHUNK")) { const int v = atoi(ev); if (v >= 1 && v <= 4096) chunk = v; }   // tests
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
    P.nb = n_bins; P.range_hi = range_hi; P.skin = skin; P.thr = thr; P.count = count;
    P.frame_counts = frame_counts; P.lut_n = lut_n; P.lut_m = m; P.chunk = (int)chunk; P.list_cap = list_cap;
    P.work = work; P.lut = lut;
    kern<<<(unsigned)grid, R5_THREADS, smem, st>>>(P);
    PBK_CHECK_LAUNCH(ctx, "k_rdf5");
    return PBK_OK;
}
