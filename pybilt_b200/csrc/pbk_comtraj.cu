// pbk_comtraj.cu -- pair searches of the stand-alone COM trajectory.
//
// Reference arithmetic restated (pybilt/com_trajectory/COMTraj.py):
//   calc_membrane_thickness                      :1244-1408  nearest cross-leaflet partner of every lipid
//   calc_area_per_lipid_closest_neighbor_circle  :1655-1766  nearest member of the same leaflet
//   check_clustering                             :1411-1576  pairs within a cut-off (clusters are grown on the host)
// All three use COMTraj's own minimum image on the WRAPPED COMs (not distance_euclidean_pbc):
//   d = a - b;  if |d| > h:  d = L - |a - h| - |b - h|      with h = f32(L / 2), L the float32 box edge,
//   rxy = sqrt(dx*dx + dy*dy)  (two rounded products and a rounded sum -- no FMA),
// scan the candidates in list order and keep the first strict minimum below the start value 10000.0.
#include "pbk_common.cuh"

// mode 1: distance_euclidean_pbc(a, b, box, center='box_half') of the BilayerAnalyzer protocols
// (pybilt/common/distance_cutoff_clustering.py:27-71; msd_dist :5776-5805, dc_cluster :2936-2944 of
// analysis_protocols.py): a' = a - c, b' = b - c with c = f32 box / 2, d^2 as np.dot forms it
static __device__ __forceinline__ double ct_d1(double ax, double ay, double bx, double by, double Lx, double Ly,
                                               double hx, double hy)
{
    return __dsqrt_rn(pbk_pbc_dist2(__dsub_rn(ax, hx), __dsub_rn(ay, hy), __dsub_rn(bx, hx), __dsub_rn(by, hy), Lx, Ly, hx, hy));
}

static __device__ __forceinline__ double ct_rxy(double ax, double ay, double bx, double by, double Lx, double Ly,
                                                double hx, double hy)
{
    double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by);
    if (fabs(dx) > hx) dx = __dsub_rn(__dsub_rn(Lx, fabs(__dsub_rn(ax, hx))), fabs(__dsub_rn(bx, hx)));
    if (fabs(dy) > hy) dy = __dsub_rn(__dsub_rn(Ly, fabs(__dsub_rn(ay, hy))), fabs(__dsub_rn(by, hy)));
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

#define CT_THREADS 128
#define CT_TILE 256

// COMFrame._rewrap (pybilt/bilayer_analyzer/com_frame.py:233-270, rep setting com_frame:rewrap): the
// wrapped COM of a lipid is rebuilt from its unwrapped one when the two disagree by more than 1 A:
//   pos_c = com_unwrap + mem_com;  |pos_c - com| > 1  (np.dot of three float64: fma(c,c, fma(b,b, a*a)))
//   every component is then moved by whole (float32) box edges into [0, box].
__global__ void k_rewrap(double *__restrict__ com, const double *__restrict__ com_unwrap, const double *__restrict__ mem_com,
                         const float *__restrict__ box, int64_t F, int64_t N, int32_t *__restrict__ status)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= F * N) return;
    const int64_t f = t / N;
    double pc[3], d[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        pc[k] = __dadd_rn(com_unwrap[3 * t + k], mem_com[3 * f + k]);
        d[k] = __dsub_rn(pc[k], com[3 * t + k]);
    }
    const double d2 = fma(d[2], d[2], fma(d[1], d[1], __dmul_rn(d[0], d[0])));
    if (!(__dsqrt_rn(d2) > 1.0)) return;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double b = (double)box[3 * f + k];
        double p = pc[k];
        if (!(b > 0.0)) { atomicOr(status, PBK_ST_IMAGE_OVERFLOW); com[3 * t + k] = p; continue; }
        int it = 0;
        while (p < 0.0 && ++it < (1 << 20)) p = __dadd_rn(p, b);
        while (p > b && ++it < (1 << 20)) p = __dsub_rn(p, b);
        com[3 * t + k] = p;
    }
}

extern "C" int pbk_rewrap(pbk_ctx *ctx, double *com, const double *com_unwrap, const double *mem_com, const float *box,
                          int64_t F, int64_t N, int32_t *status, void *stream)
{
    if (!ctx || !com || !com_unwrap || !mem_com || !box || !status || F < 0 || N <= 0)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_rewrap: bad argument");
    if (F == 0) return PBK_OK;
    k_rewrap<<<(unsigned)((F * N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(com, com_unwrap, mem_com, box, F, N, status);
    PBK_CHECK_LAUNCH(ctx, "k_rewrap");
    return PBK_OK;
}

// grid (ceil(na / CT_THREADS), F): thread = one lipid of list A, candidates of list B staged in
// shared memory tile by tile, in list order
__global__ void __launch_bounds__(CT_THREADS)
k_ct_nearest(const double *__restrict__ com, const float *__restrict__ box, const int32_t *__restrict__ idx_a,
             int na, const int32_t *__restrict__ idx_b, int nb, int xi, int yi, int exclude_self, int mode, int64_t N,
             int32_t *__restrict__ out_pos, double *__restrict__ out_dist, double *__restrict__ out_dist2)
{
    __shared__ double s_x[CT_TILE], s_y[CT_TILE];
    __shared__ int s_i[CT_TILE];
    const int64_t f = blockIdx.y;
    const double *c = com + f * N * 3;
    const float Lxf = box[f * 3 + xi], Lyf = box[f * 3 + yi];
    const double Lx = (double)Lxf, Ly = (double)Lyf, hx = (double)(Lxf * 0.5f), hy = (double)(Lyf * 0.5f);
    const int a = blockIdx.x * CT_THREADS + threadIdx.x;
    const int ia = a < na ? idx_a[a] : -1;
    const double ax = ia >= 0 ? c[3 * (int64_t)ia + xi] : 0.0, ay = ia >= 0 ? c[3 * (int64_t)ia + yi] : 0.0;
    double best = 10000.0, second = 10000.0;
    int pos = -1;
    for (int b0 = 0; b0 < nb; b0 += CT_TILE) {
        __syncthreads();
        for (int j = threadIdx.x; j < CT_TILE && b0 + j < nb; j += CT_THREADS) {
            const int ib = idx_b[b0 + j];
            s_i[j] = ib; s_x[j] = c[3 * (int64_t)ib + xi]; s_y[j] = c[3 * (int64_t)ib + yi];
        }
        __syncthreads();
        if (ia >= 0) {
            const int m = min(CT_TILE, nb - b0);
            for (int j = 0; j < m; j++) {
                if (exclude_self && s_i[j] == ia) continue;
                const double r = mode ? ct_d1(ax, ay, s_x[j], s_y[j], Lx, Ly, hx, hy)
                                      : ct_rxy(ax, ay, s_x[j], s_y[j], Lx, Ly, hx, hy);
                if (r < best) { second = best; best = r; pos = b0 + j; }
                else if (r < second) second = r;
            }
        }
    }
    if (ia >= 0) {
        out_pos[f * na + a] = pos;
        out_dist[f * na + a] = best;
        if (out_dist2) out_dist2[f * na + a] = second;
    }
}

// grid (ceil(n / CT_THREADS), F): row a of the frame's adjacency matrix, bit b of word b / 32 set when
// rxy(idx[b], idx[a]) <= dist (evaluated as the reference does: candidate first, cluster seed second)
__global__ void __launch_bounds__(CT_THREADS)
k_ct_adjacency(const double *__restrict__ com, const float *__restrict__ box, const int32_t *__restrict__ idx, int n,
               int xi, int yi, double dist, int mode, int64_t N, uint32_t *__restrict__ bits)
{
    __shared__ double s_x[CT_TILE], s_y[CT_TILE];
    const int64_t f = blockIdx.y;
    const double *c = com + f * N * 3;
    const float Lxf = box[f * 3 + xi], Lyf = box[f * 3 + yi];
    const double Lx = (double)Lxf, Ly = (double)Lyf, hx = (double)(Lxf * 0.5f), hy = (double)(Lyf * 0.5f);
    const int a = blockIdx.x * CT_THREADS + threadIdx.x;
    const int ia = a < n ? idx[a] : -1;
    const double ax = ia >= 0 ? c[3 * (int64_t)ia + xi] : 0.0, ay = ia >= 0 ? c[3 * (int64_t)ia + yi] : 0.0;
    const int nw = (n + 31) / 32;
    uint32_t *row = bits + (f * n + a) * nw;
    for (int b0 = 0; b0 < n; b0 += CT_TILE) {
        __syncthreads();
        for (int j = threadIdx.x; j < CT_TILE && b0 + j < n; j += CT_THREADS) {
            const int ib = idx[b0 + j];
            s_x[j] = c[3 * (int64_t)ib + xi]; s_y[j] = c[3 * (int64_t)ib + yi];
        }
        __syncthreads();
        if (ia >= 0) {
            const int m = min(CT_TILE, n - b0);
            for (int w = 0; w < m; w += 32) {
                uint32_t word = 0u;
                for (int j = w; j < min(m, w + 32); j++) {
                    // comc = candidate (b), coms = seed (a):  dx = comc - coms
                    // mode 1: the pair is evaluated once, dist(vectors[i], vectors[j]) with i < j
                    const bool lower_first = a < b0 + j;
                    const double r = !mode ? ct_rxy(s_x[j], s_y[j], ax, ay, Lx, Ly, hx, hy)
                                   : (lower_first ? ct_d1(ax, ay, s_x[j], s_y[j], Lx, Ly, hx, hy)
                                                  : ct_d1(s_x[j], s_y[j], ax, ay, Lx, Ly, hx, hy));
                    if (r <= dist && (!mode || a != b0 + j)) word |= 1u << (j - w);
                }
                row[(b0 + w) / 32] = word;
            }
        }
    }
}

extern "C" int pbk_ct_nearest(pbk_ctx *ctx, const double *com, const float *box, const int32_t *idx_a, int32_t na,
                              const int32_t *idx_b, int32_t nb, int xi, int yi, int exclude_self, int mode, int64_t F,
                              int64_t N, int32_t *out_pos, double *out_dist, double *out_dist2, void *stream)
{
    if (!ctx || !com || !box || !out_pos || !out_dist || na < 0 || nb < 0 || F < 0 || N <= 0 || xi < 0 || xi > 2 ||
        yi < 0 || yi > 2 || (na > 0 && !idx_a) || (nb > 0 && !idx_b))
        return pbk_fail(ctx, PBK_EINVAL, "pbk_ct_nearest: bad argument");
    if (F == 0 || na == 0) return PBK_OK;
    if (F > 65535) return PBK_EUNSUPPORTED;
    k_ct_nearest<<<dim3((unsigned)((na + CT_THREADS - 1) / CT_THREADS), (unsigned)F), CT_THREADS, 0, (cudaStream_t)stream>>>(
        com, box, idx_a, na, idx_b, nb, xi, yi, exclude_self, mode, N, out_pos, out_dist, out_dist2);
    PBK_CHECK_LAUNCH(ctx, "k_ct_nearest");
    return PBK_OK;
}

extern "C" int pbk_ct_adjacency(pbk_ctx *ctx, const double *com, const float *box, const int32_t *idx, int32_t n,
                                int xi, int yi, double dist, int mode, int64_t F, int64_t N, uint32_t *bits, void *stream)
{
    if (!ctx || !com || !box || !bits || n < 0 || F < 0 || N <= 0 || xi < 0 || xi > 2 || yi < 0 || yi > 2 || (n > 0 && !idx))
        return pbk_fail(ctx, PBK_EINVAL, "pbk_ct_adjacency: bad argument");
    if (F == 0 || n == 0) return PBK_OK;
    if (F > 65535) return PBK_EUNSUPPORTED;
    k_ct_adjacency<<<dim3((unsigned)((n + CT_THREADS - 1) / CT_THREADS), (unsigned)F), CT_THREADS, 0, (cudaStream_t)stream>>>(
        com, box, idx, n, xi, yi, dist, mode, N, bits);
    PBK_CHECK_LAUNCH(ctx, "k_ct_adjacency");
    return PBK_OK;
}

// ------------------------------------------------------------------------------------------
// leaflets:assign_method = 'orientation'   (BilayerAnalyzer._build_leaflets_orientation,
// pybilt/bilayer_analyzer/bilayer_analyzer.py:849-876)
// The reference takes the head and tail atoms of every lipid out of the CURRENT timestep, which COMFrame
// has overwritten with v = f32(f64(u) - mem_com) (com_frame.py:172-181), forms their centres of mass
// (rows added in atom order, divided by the mass sum) and labels the lipid by the sign of
// (head - tail)[norm] (the cosine's denominator is positive; a zero or NaN vector component leaves the
// label empty -> KeyError('')).  Only the normal component of the selected atoms is needed, so one
// thread per lipid walks the frames in order with the exact unwrap rule (mda_unwrap_vectorized.py:12-93)
// for its few atoms; labels are written on the frames where the leaflets are rebuilt.
// ------------------------------------------------------------------------------------------
static __device__ int ct_shift(float d, float L, int *bad)
{
    const float h = L * 0.5f;
    int s = 0;
    if (!(L > 0.0f)) { *bad = 1; return 0; }
    const double dd = (double)d, Ld = (double)L, hd = (double)h;
    if (d > h) {
        do { s -= 1; } while (__dadd_rn(dd, __dmul_rn((double)s, Ld)) > hd && s > -32768);
    } else if (d < -h) {
        do { s += 1; } while (__dadd_rn(dd, __dmul_rn((double)s, Ld)) < -hd && s < 32768);
    }
    if (s <= -32768 || s >= 32768) { *bad = 1; s = 0; }
    return s;
}

__global__ void k_orient_labels(const float *__restrict__ x, const float *__restrict__ box, const double *__restrict__ mem_com,
                                const int32_t *__restrict__ sel_atom, const double *__restrict__ sel_mass,
                                const int32_t *__restrict__ seg, float *__restrict__ u_state, int first_frame,
                                int64_t F, int64_t M, int64_t N, int norm, int64_t frame0, int interval,
                                uint8_t *__restrict__ labels, int32_t *__restrict__ status)
{
    // seg[2 i] .. seg[2 i + 1]: tail atoms of lipid i in sel_*, seg[2 i + 1] .. seg[2 i + 2]: head atoms
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int a0 = seg[2 * i], a1 = seg[2 * i + 1], a2 = seg[2 * i + 2];
    int bad = 0, tie = 0;
    for (int64_t t = 0; t < F; t++) {
        const float L = box[t * 3 + norm];
        const double mem = mem_com[t * 3 + norm];
        const bool rebuild = ((frame0 + t) % interval) == 0;
        double st = 0.0, mt = 0.0, sh = 0.0, mh = 0.0;
        for (int a = a0; a < a2; a++) {
            const float xv = x[(t * M + sel_atom[a]) * 3 + norm];
            float u = xv;
            if (!(first_frame && t == 0)) {
                const int s = ct_shift(__fsub_rn(xv, u_state[a]), L, &bad);
                u = pbk_unwrapped(xv, s, L);
            }
            u_state[a] = u;
            if (rebuild) {
                const double v = (double)__double2float_rn(__dsub_rn((double)u, mem));
                const double m = sel_mass[a];
                if (a < a1) { st = (a == a0) ? __dmul_rn(v, m) : __dadd_rn(st, __dmul_rn(v, m)); mt = (a == a0) ? m : __dadd_rn(mt, m); }
                else { sh = (a == a1) ? __dmul_rn(v, m) : __dadd_rn(sh, __dmul_rn(v, m)); mh = (a == a1) ? m : __dadd_rn(mh, m); }
            }
        }
        if (rebuild) {
            const double vec = __dsub_rn(__ddiv_rn(sh, mh), __ddiv_rn(st, mt));
            if (vec > 0.0) labels[t * N + i] = 1;
            else if (vec < 0.0) labels[t * N + i] = 0;
            else { labels[t * N + i] = 0; tie = 1; }
        }
    }
    if (bad) atomicOr(status, PBK_ST_IMAGE_OVERFLOW);
    if (tie) atomicOr(status, PBK_ST_LEAFLET_TIE);
}

extern "C" int pbk_orient_labels(pbk_ctx *ctx, const float *x, const float *box, const double *mem_com,
                                 const int32_t *sel_atom, const double *sel_mass, const int32_t *seg, float *u_state,
                                 int first_frame, int64_t F, int64_t M, int64_t N, int norm, int64_t frame0, int interval,
                                 uint8_t *labels, int32_t *status, void *stream)
{
    if (!ctx || !x || !box || !mem_com || !sel_atom || !sel_mass || !seg || !u_state || !labels || !status || F < 0 ||
        M <= 0 || N <= 0 || norm < 0 || norm > 2 || interval < 1 || frame0 < 0)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_orient_labels: bad argument");
    if (F == 0) return PBK_OK;
    k_orient_labels<<<(unsigned)((N + 63) / 64), 64, 0, (cudaStream_t)stream>>>(x, box, mem_com, sel_atom, sel_mass, seg, u_state,
                                                                               first_frame, F, M, N, norm, frame0, interval,
                                                                               labels, status);
    PBK_CHECK_LAUNCH(ctx, "k_orient_labels");
    return PBK_OK;
}

// ------------------------------------------------------------------------------------------
// self-check of pbk_div_r (pbk_common.cuh) against the hardware's correctly rounded division on
// pseudo-random operands: mismatches[0] = number of (a, b) with pbk_div_r(a, b, 1/b) != a / b.
// b cycles through typical bead masses and random values in [2^-4, 2^16), a through +-[2^-10, 2^30).
// ------------------------------------------------------------------------------------------
__global__ void k_div_check(unsigned long long seed, int64_t n, unsigned long long *mismatches)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long s = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    next(); next();
    const double masses[8] = {72.0, 864.0, 576.0, 760.076, 1.008, 30.974, 786.113, 12.011};
    unsigned long long r = next();
    double b = masses[r & 7];
    if (r & 8) b = ldexp((double)(next() >> 11) * 0x1p-53 + 0.5, (int)(next() % 20) - 4);
    r = next();
    double a = ldexp((double)(r >> 11) * 0x1p-53 + 0.5, (int)(next() % 40) - 10);
    if (r & 1) a = -a;
    const double rb = __ddiv_rn(1.0, b);
    if (pbk_div_r(a, b, rb) != __ddiv_rn(a, b)) atomicAdd(mismatches, 1ull);
}

extern "C" int pbk_debug_div_check(pbk_ctx *ctx, uint64_t seed, int64_t n, uint64_t *mismatches, void *stream)
{
    if (!ctx || !mismatches || n < 0) return pbk_fail(ctx, PBK_EINVAL, "pbk_debug_div_check: bad argument");
    if (n == 0) return PBK_OK;
    k_div_check<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, n, reinterpret_cast<unsigned long long *>(mismatches));
    PBK_CHECK_LAUNCH(ctx, "k_div_check");
    return PBK_OK;
}
