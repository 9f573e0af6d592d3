// pbk_sweep.cu -- frame reduction for LARGE systems (a frame does not fit in shared memory):
// unwrap + membrane COM + per-lipid COMs + leaflet labels with TWO reads of the positions and no
// per-frame image-count array.
//
// Reference arithmetic restated (paths relative to the PyBILT tree), identical to pbk_com.cu:
//   unwrap        pybilt/mda_tools/mda_unwrap_vectorized.py:12-93
//   membrane COM  pybilt/bilayer_analyzer/com_frame.py:172-181 (NumPy pairwise order)
//   lipid COM     pybilt/bilayer_analyzer/com_frame.py:43-96,163-185
//   leaflets      pybilt/bilayer_analyzer/bilayer_analyzer.py:826-847
//
//   k_sweep_a   one CTA per <=128-atom leaf of NumPy's pairwise summation walks the frames of the
//               batch in time order; the leaf's slice of every frame arrives by TMA bulk copy
//               (cp.async.bulk + mbarrier) in a shared-memory ring several frames ahead of the
//               arithmetic, so the unwrap recurrence (frame t needs the image counts of frame
//               t-1) never waits for HBM.  Warps 0-2 own one float4 column of the leaf each
//               (previous raw values and image counts in registers), decide the image counts of
//               frame t and write the products m*f64(u) to shared memory; warp 3 adds them in
//               NumPy's order one frame behind (8 accumulators x 3 axes = 24 lanes, 16 sequential
//               adds, butterfly, tail) and stores the leaf sums.  Every SW_CHUNK frames the image
//               counts are written out (int16 per chain), which is what lets the second sweep run
//               time-parallel.  Reads x once; writes 24 B per leaf and frame.
//   k_mem_tree  (pbk_com.cu) combines the leaf sums along the recursion tree -> mem_com[t].
//   k_sweep_b   one thread per (lipid, axis) and chunk of SW_CHUNK frames: starts from the image
//               counts A left at the chunk start, repeats the recurrence for its <= NA atoms
//               (previous raw values and counts in registers, next frame's loads in flight while
//               the current one is reduced), accumulates m*f64(x) and m*f64(f32(f64(u) - mem_com))
//               in atom order, divides, stores both COMs (8-byte stores, consecutive threads =
//               consecutive addresses) and the bilayer-normal component in a compact array.
//               Reads x a second time; writes 48 + 8 B per lipid and frame.
//   k_leaflets_c  leaflet labels from the compact normal components, one CTA per frame.
//
// The image-count shortcut is the one of the fused path (pbk_same_image, pbk_common.cuh): a raw
// jump that stays clear of +-L/2 by the rounding bound keeps the count; everything else runs
// the reference's float64 loops.  Counts beyond +-127 switch the owning thread to the exact
// rule for every frame (the bound is derived for |s| <= 127).
#include "pbk_common.cuh"

// pbk_com.cu: combines leaf sums [F][L+K][3] along the recursion tree -> mem_com[F][3]
int pbk_launch_mem_tree(pbk_ctx *ctx, const int32_t *node_left, const int32_t *node_right, const int32_t *level_off,
                        int32_t H, int32_t L, int32_t K, double total_mass, int64_t F, double *node_vals,
                        double *mem_com, cudaStream_t st);

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
static __device__ __forceinline__ uint32_t sw_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void sw_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sw_smem_u32(bar)), "r"(count) : "memory");
}
static __device__ __forceinline__ void sw_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sw_smem_u32(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void sw_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sw_smem_u32(dst)), "l"(src), "r"(bytes), "r"(sw_smem_u32(bar)) : "memory");
}
static __device__ __forceinline__ void sw_cp4(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sw_smem_u32(dst)), "l"(src) : "memory");
}
static __device__ __forceinline__ void sw_cp_arrive(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(sw_smem_u32(bar)) : "memory");
}
static __device__ __forceinline__ bool sw_mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = sw_smem_u32(bar);
    for (int it = 0; it < (1 << 20); it++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
static __device__ __forceinline__ void sw_fence_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// image shift of the reference rule for the float32 difference d and box length L
// (mda_unwrap_vectorized.py:31-45: strict comparisons, dz + s*C in float64)
static __device__ __noinline__ int sw_shift(float d, float L, int *bad)
{
    const float h = L * 0.5f;
    int s = 0;
    if (!(L > 0.0f)) { *bad |= 1; return 0; }
    const double dd = (double)d, Ld = (double)L, hd = (double)h;
    if (d > h) {
        do { s -= 1; } while (__dadd_rn(dd, __dmul_rn((double)s, Ld)) > hd && s > -32768);
    } else if (d < -h) {
        do { s += 1; } while (__dadd_rn(dd, __dmul_rn((double)s, Ld)) < -hd && s < 32768);
    }
    if (s <= -32768 || s >= 32768) { *bad |= 1; s = 0; }
    return s;
}

static __device__ __noinline__ float sw_unwrap_slow(float x, int s, float L)
{
    if (s == 0) return x;
    return __double2float_rn(__dadd_rn((double)x, __dmul_rn((double)s, (double)L)));
}

// ------------------------------------------------------------------------------------------
// box table: per frame {L0, L1, L2, he0, he1, he2, Lbig, ok}; he = L/2 - rounding bound against
// the frame before (-1 when the box is bad: every same-image test fails and the exact rule
// reports it)
// ------------------------------------------------------------------------------------------
__global__ void k_box_table(const float *__restrict__ box, int64_t F, float *__restrict__ tab)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= F) return;
    const int64_t tp = t > 0 ? t - 1 : 0;
    float L[3], he[3];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; k++) { L[k] = box[t * 3 + k]; ok = ok && (L[k] > 0.0f); }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float Lp = box[tp * 3 + k];
        he[k] = ok ? L[k] * 0.5f - pbk_image_eps(L[k], Lp) : -1.0f;
    }
    float4 *o = reinterpret_cast<float4 *>(tab + t * 8);
    o[0] = make_float4(L[0], L[1], L[2], he[0]);
    o[1] = make_float4(he[1], he[2], (L[0] < 8192.0f && L[1] < 8192.0f && L[2] < 8192.0f) ? 0.0f : 1.0f, ok ? 1.0f : 0.0f);
}

// ------------------------------------------------------------------------------------------
// k_sweep_a
// ------------------------------------------------------------------------------------------
#define SW_CHUNK 32                  // frames per time chunk of the second sweep
#define SA_THREADS 128
#define SA_DEPTH 6
#define SA_LEAF_FLOATS 384           // 128 atoms x 3

struct SweepAParams {
    const float *x; const float *boxtab; const double *mass;
    const int32_t *leaf_start, *leaf_len;
    int L, nodes_per_frame;
    const float *prev0;             // u_state (later batches) or x (first frame: frame 0 is its own reference)
    const int32_t *init_images;
    int64_t F; int64_t M; int64_t net_frame;
    int32_t *net_images; float *u_out; double *node_vals; int32_t *status;
    int16_t *snap;                  // [n_chunks][3M] image counts at frame c*SW_CHUNK - 1 (row 0 unused)
    int store_sums;
};

template <bool BULK>
__global__ void __launch_bounds__(SA_THREADS)
k_sweep_a(SweepAParams P)
{
    __shared__ __align__(128) float s_x[SA_DEPTH][SA_LEAF_FLOATS];
    __shared__ __align__(16) double s_p[2][SA_LEAF_FLOATS];
    __shared__ uint64_t s_full[SA_DEPTH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int leaf = blockIdx.x;
    const int ls = P.leaf_start[leaf], n = P.leaf_len[leaf];
    const int nfl = 3 * n;                       // floats of the leaf per frame
    const int ncol = (nfl + 3) >> 2;
    const int64_t C3 = 3 * P.M;
    const float *xl = P.x + 3 * (int64_t)ls;     // the leaf inside frame 0
    int bad = 0;

    if (tid == 0)
        for (int d = 0; d < SA_DEPTH; d++) sw_mbar_init(&s_full[d], BULK ? 1u : 32u);
    sw_fence_async();
    __syncthreads();

    auto issue = [&](int64_t t) {                // warp 3 only
        const int slot = (int)(t % SA_DEPTH);
        const float *src = xl + t * C3;
        if (BULK) {
            if (lane == 0) {
                sw_mbar_expect_tx(&s_full[slot], (uint32_t)nfl * 4u);
                sw_bulk_g2s(&s_x[slot][0], src, (uint32_t)nfl * 4u, &s_full[slot]);
            }
        } else {
            for (int c = lane; c < nfl; c += 32) sw_cp4(&s_x[slot][c], src + c);
            sw_cp_arrive(&s_full[slot]);
        }
    };
    if (warp == 3)
        for (int64_t t = 0; t < P.F && t < SA_DEPTH; t++) issue(t);

    // ---- column role (warps 0-2) -------------------------------------------------------------
    const int col = tid;
    const bool col_on = warp < 3 && col < ncol;
    const int nvalid = col_on ? min(4, nfl - 4 * col) : 0;
    const int k0 = col % 3;                       // axis of comp e is (k0 + e) % 3
    float pv[4] = {0.f, 0.f, 0.f, 0.f};
    int sc[4] = {0, 0, 0, 0};
    float sf[4] = {0.f, 0.f, 0.f, 0.f};
    double ma = 0.0, mb = 0.0;
    int wide = 0;
    if (col_on) {
        const int a0 = ls + (4 * col) / 3;
        ma = P.mass[a0];
        mb = (a0 + 1 < ls + n) ? P.mass[a0 + 1] : 0.0;
        const float *p0 = P.prev0 + 3 * (int64_t)ls + 4 * col;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (e < nvalid) {
                pv[e] = p0[e];
                int s0 = P.init_images ? P.init_images[3 * (int64_t)ls + 4 * col + e] : 0;
                if (s0 > 32767 || s0 < -32767) { bad |= 1; s0 = 0; }
                sc[e] = s0;
                sf[e] = (float)s0;
                if (s0 > 127 || s0 < -127) wide = 1;
            }
        }
    }
    const int n_first = 3 - k0;                   // comps e < n_first belong to the column's first atom

    // ---- reducing role (warp 3, lanes 0-23) ---------------------------------------------------
    const int rj = lane / 3, rk = lane % 3;
    const bool red_on = warp == 3 && lane < 24;
    const int nq = n >= 8 ? n / 8 : 0;

    float Lp[3] = {0.f, 0.f, 0.f};
    float4 nb0 = __ldg(reinterpret_cast<const float4 *>(P.boxtab)), nb1 = __ldg(reinterpret_cast<const float4 *>(P.boxtab) + 1);
    Lp[0] = nb0.x; Lp[1] = nb0.y; Lp[2] = nb0.z;
    for (int64_t t = 0; t <= P.F; t++) {
        if (warp < 3) {
            if (t < P.F) {
                const int slot = (int)(t % SA_DEPTH);
                const float4 b0 = nb0, b1 = nb1;
                if (t + 1 < P.F) {             // box row of the next frame: its latency hides behind this frame
                    nb0 = __ldg(reinterpret_cast<const float4 *>(P.boxtab + (t + 1) * 8));
                    nb1 = __ldg(reinterpret_cast<const float4 *>(P.boxtab + (t + 1) * 8) + 1);
                }
                if (!sw_mbar_wait(&s_full[slot], (uint32_t)(t / SA_DEPTH) & 1u)) bad |= 2;
                if (col_on) {
                    const float Lc[3] = {b0.x, b0.y, b0.z}, hc[3] = {b0.w, b1.x, b1.y};
                    const bool Lbig = b1.z != 0.0f;
                    const float4 v = reinterpret_cast<const float4 *>(&s_x[slot][0])[col];
                    const float xv[4] = {v.x, v.y, v.z, v.w};
                    float Le[4], he[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int k = (k0 + e) % 3;
                        Le[e] = k == 0 ? Lc[0] : (k == 1 ? Lc[1] : Lc[2]);
                        he[e] = k == 0 ? hc[0] : (k == 1 ? hc[1] : hc[2]);
                    }
                    float sl[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) sl[e] = (e < nvalid) ? he[e] - fabsf(__fsub_rn(xv[e], pv[e])) : 1.0f;
                    const float mn = fminf(fminf(sl[0], sl[1]), fminf(sl[2], sl[3]));
                    if (!(mn > 0.0f) || !(sl[0] > 0.0f) || !(sl[1] > 0.0f) || !(sl[2] > 0.0f) || !(sl[3] > 0.0f) || wide) {
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            if (e < nvalid && (wide || !(sl[e] > 0.0f))) {
                                const int k = (k0 + e) % 3;
                                const float Lpk = k == 0 ? Lp[0] : (k == 1 ? Lp[1] : Lp[2]);
                                const int sn = sw_shift(__fsub_rn(xv[e], pbk_unwrapped(pv[e], sc[e], Lpk)), Le[e], &bad);
                                sc[e] = sn;
                                sf[e] = (float)sn;
                                if (sn > 127 || sn < -127) wide = 1;
                            }
                        }
                    }
                    float uv[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) uv[e] = __fmaf_rn(sf[e], Le[e], xv[e]);
                    {
                        const float tiny = fminf(fminf(fabsf(xv[0]), fabsf(xv[1])), fminf(fabsf(xv[2]), fabsf(xv[3])));
                        if (!(tiny >= 0.00390625f) || Lbig || wide) {
#pragma unroll
                            for (int e = 0; e < 4; e++) uv[e] = (e < nvalid) ? sw_unwrap_slow(xv[e], sc[e], Le[e]) : 0.0f;
                        }
                    }
                    double pr[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) pr[e] = __dmul_rn(e < n_first ? ma : mb, (double)uv[e]);
                    double2 *dst = reinterpret_cast<double2 *>(&s_p[t & 1][4 * col]);
                    dst[0] = make_double2(pr[0], pr[1]);
                    dst[1] = make_double2(pr[2], pr[3]);
#pragma unroll
                    for (int e = 0; e < 4; e++) pv[e] = xv[e];
                    if (t == P.F - 1 && P.u_out) {
                        float *uo = P.u_out + 3 * (int64_t)ls + 4 * col;
#pragma unroll
                        for (int e = 0; e < 4; e++) if (e < nvalid) uo[e] = pbk_unwrapped(xv[e], sc[e], Le[e]);
                    }
                    if (P.snap && ((t + 1) % SW_CHUNK) == 0 && t + 1 < P.F) {
                        int16_t *so = P.snap + ((t + 1) / SW_CHUNK) * C3 + 3 * (int64_t)ls + 4 * col;
#pragma unroll
                        for (int e = 0; e < 4; e++) if (e < nvalid) so[e] = (int16_t)sc[e];
                    }
                    if (t == P.net_frame && P.net_images) {
                        int32_t *no = P.net_images + 3 * (int64_t)ls + 4 * col;
#pragma unroll
                        for (int e = 0; e < 4; e++) if (e < nvalid) no[e] = sc[e];
                    }
                }
                Lp[0] = b0.x; Lp[1] = b0.y; Lp[2] = b0.z;
                sw_fence_async();               // generic reads of the slot before its refill by the async proxy
            }
        } else if (t > 0 && P.store_sums) {
            // NumPy's pairwise leaf sum of frame t-1 (numpy/_core/src/umath/loops_utils.h.src):
            //   r[j] = a[j]; r[j] += a[8q + j];  res = ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); tail sequentially
            const double *pp = &s_p[(t - 1) & 1][0];
            double r = 0.0;
            if (red_on && nq > 0) {
                r = pp[lane];
#pragma unroll 4
                for (int q = 1; q < nq; q++) r = __dadd_rn(r, pp[q * 24 + lane]);
            }
            {
                const bool act = lane < 24;
                const int s1 = ((rj ^ 1) * 3 + rk) & 31, s2 = ((rj ^ 2) * 3 + rk) & 31, s4 = ((rj ^ 4) * 3 + rk) & 31;
                double o = __shfl_sync(0xffffffffu, r, act ? s1 : lane);
                r = __dadd_rn(r, o);
                o = __shfl_sync(0xffffffffu, r, act ? s2 : lane);
                r = __dadd_rn(r, o);
                o = __shfl_sync(0xffffffffu, r, act ? s4 : lane);
                r = __dadd_rn(r, o);
            }
            if (lane < 3) {
                int i0 = 8 * nq;
                if (nq == 0) { r = 0.0; i0 = 0; }
                for (int i = i0; i < n; i++) r = __dadd_rn(r, pp[3 * i + lane]);
                P.node_vals[((t - 1) * (int64_t)P.nodes_per_frame + leaf) * 3 + lane] = r;
            }
        }
        __syncthreads();
        // the slot of frame t is free again: refill it with frame t + DEPTH
        if (warp == 3 && t + SA_DEPTH < P.F) issue(t + SA_DEPTH);
    }
    if (bad) atomicOr(P.status, (bad & 1 ? PBK_ST_IMAGE_OVERFLOW : 0) | (bad & 2 ? PBK_ST_INTERNAL : 0));
}

// ------------------------------------------------------------------------------------------
// k_sweep_b<NA>: NA > 0 -- every bead has at most NA atoms, state in registers;
//                NA == 0 -- any bead size, image counts in shared memory [atom][thread], previous
//                           frame re-read (cache hit)
// ------------------------------------------------------------------------------------------
#define SB_THREADS 128

struct SweepBParams {
    const float *x; const float *boxtab; const double *mass; const int32_t *seg; const int32_t *gather;
    const double *lipid_mass; const double *mem_com;
    const float *prev0; const int32_t *init_images; const int16_t *snap;
    int64_t F; int64_t M; int64_t N; int norm; int max_atoms;
    double *com, *com_unwrap, *zc; int32_t *status;
};

// one atom component: image count of frame t from (raw previous value, its count); returns the count
static __device__ __forceinline__ int sb_count(float xv, float pvv, int s, float hc, float Lc, float Lp, bool wide, int *bad)
{
    if (!(hc - fabsf(__fsub_rn(xv, pvv)) > 0.0f) || wide)
        s = sw_shift(__fsub_rn(xv, pbk_unwrapped(pvv, s, Lp)), Lc, bad);
    return s;
}

// adds m*f64(x) to the wrapped sum and m*f64(f32(f64(u) - mem)) to the unwrapped one (com_frame.py:71-96,180)
static __device__ __forceinline__ void sb_accumulate(float xv, int s, float Lc, bool Lbig, double mem, double m,
                                                     double &cw, double &cu)
{
    const double xd = (double)xv;
    cw = __dadd_rn(cw, __dmul_rn(xd, m));
    double ud = xd;
    if (s != 0) {
        float u = __fmaf_rn((float)s, Lc, xv);
        if (!(fabsf(xv) >= 0.00390625f) || Lbig || s > 255 || s < -255) u = sw_unwrap_slow(xv, s, Lc);
        ud = (double)u;
    }
    const float v = __double2float_rn(__dsub_rn(ud, mem));
    cu = __dadd_rn(cu, __dmul_rn((double)v, m));
}

template <int NA>
__global__ void __launch_bounds__(SB_THREADS)
k_sweep_b(SweepBParams P)
{
    extern __shared__ __align__(16) unsigned char sb_raw[];
    __shared__ float s_box[SW_CHUNK][8];
    __shared__ double s_mem[SW_CHUNK][3];
    int16_t *s_img = reinterpret_cast<int16_t *>(sb_raw);       // NA == 0: [max_atoms][SB_THREADS]
    const int tid = threadIdx.x;
    const int64_t C3 = 3 * P.M;
    const int64_t f0 = (int64_t)blockIdx.y * SW_CHUNK, f1 = min(P.F, f0 + SW_CHUNK);
    const int nf = (int)(f1 - f0);
    for (int r = tid; r < nf * 8; r += SB_THREADS) s_box[r >> 3][r & 7] = P.boxtab[f0 * 8 + r];
    for (int r = tid; r < nf * 3; r += SB_THREADS) s_mem[r / 3][r % 3] = P.mem_com[f0 * 3 + r];
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * SB_THREADS + tid;
    if (g >= 3 * P.N) return;
    const int64_t i = g / 3;
    const int k = (int)(g - 3 * i);
    const int p0 = P.seg[i], n = P.seg[i + 1] - p0;
    const double lm = P.lipid_mass[i];
    const int32_t *__restrict__ gat = P.gather;
    const double *__restrict__ gm = P.mass;
    const bool first_chunk = blockIdx.y == 0;
    const float *prow = first_chunk ? P.prev0 : P.x + (f0 - 1) * C3;      // "previous" frame of the chunk's first frame
    int bad = 0, wide = 0;
    float Lp = first_chunk ? s_box[0][k] : P.boxtab[(f0 - 1) * 8 + k];
    double *oc = P.com + f0 * 3 * P.N + g, *ou = P.com_unwrap + f0 * 3 * P.N + g;
    double *oz = (k == P.norm) ? P.zc + f0 * P.N + i : nullptr;

    if (NA > 0) {
        float pv[NA > 0 ? NA : 1], xn[NA > 0 ? NA : 1];
        int sc[NA > 0 ? NA : 1], ix[NA > 0 ? NA : 1];
#pragma unroll
        for (int j = 0; j < NA; j++) {
            pv[j] = 0.f; xn[j] = 0.f; sc[j] = 0; ix[j] = 0;
            if (j < n) {
                const int a = gat ? gat[p0 + j] : p0 + j;
                ix[j] = 3 * a + k;
                pv[j] = prow[ix[j]];
                int s0 = first_chunk ? (P.init_images ? P.init_images[ix[j]] : 0) : (int)P.snap[(int64_t)blockIdx.y * C3 + ix[j]];
                if (s0 > 32767 || s0 < -32767) { bad |= 1; s0 = 0; }
                if (s0 > 127 || s0 < -127) wide = 1;
                sc[j] = s0;
                xn[j] = P.x[f0 * C3 + ix[j]];
            }
        }
        for (int f = 0; f < nf; f++) {
            float xv[NA > 0 ? NA : 1];
            const float *nxt = P.x + (f0 + f + 1) * C3;
            const bool more = f + 1 < nf;
#pragma unroll
            for (int j = 0; j < NA; j++) {
                xv[j] = xn[j];
                if (more && j < n) xn[j] = __ldg(nxt + ix[j]);
            }
            const float Lc = s_box[f][k], hc = s_box[f][3 + k];
            const bool Lbig = s_box[f][6] != 0.0f;
            const double mem = s_mem[f][k];
            double cw = 0.0, cu = 0.0;
#pragma unroll
            for (int j = 0; j < NA; j++) {
                if (j < n) {
                    const int sn = sb_count(xv[j], pv[j], sc[j], hc, Lc, Lp, wide != 0, &bad);
                    if (sn > 127 || sn < -127) wide = 1;
                    sc[j] = sn;
                    pv[j] = xv[j];
                    sb_accumulate(xv[j], sn, Lc, Lbig, mem, __ldg(gm + ix[j] / 3), cw, cu);
                }
            }
            oc[(int64_t)f * 3 * P.N] = __ddiv_rn(cw, lm);
            const double uu = __ddiv_rn(cu, lm);
            ou[(int64_t)f * 3 * P.N] = uu;
            if (oz) oz[(int64_t)f * P.N] = uu;
            Lp = Lc;
        }
    } else {
        for (int j = 0; j < n; j++) {
            const int a = gat ? gat[p0 + j] : p0 + j;
            int s0 = first_chunk ? (P.init_images ? P.init_images[3 * a + k] : 0) : (int)P.snap[(int64_t)blockIdx.y * C3 + 3 * a + k];
            if (s0 > 32767 || s0 < -32767) { bad |= 1; s0 = 0; }
            if (s0 > 127 || s0 < -127) wide = 1;
            s_img[j * SB_THREADS + tid] = (int16_t)s0;
        }
        for (int f = 0; f < nf; f++) {
            const float *cur = P.x + (f0 + f) * C3;
            const float Lc = s_box[f][k], hc = s_box[f][3 + k];
            const bool Lbig = s_box[f][6] != 0.0f;
            const double mem = s_mem[f][k];
            double cw = 0.0, cu = 0.0;
            for (int j = 0; j < n; j++) {
                const int a = gat ? __ldg(gat + p0 + j) : p0 + j;
                const float xv = __ldg(cur + 3 * a + k), pvv = prow[3 * a + k];
                const int s = (int)s_img[j * SB_THREADS + tid];
                const int sn = sb_count(xv, pvv, s, hc, Lc, Lp, wide != 0, &bad);
                if (sn != s) {
                    s_img[j * SB_THREADS + tid] = (int16_t)sn;
                    if (sn > 127 || sn < -127) wide = 1;
                }
                sb_accumulate(xv, sn, Lc, Lbig, mem, __ldg(gm + a), cw, cu);
            }
            oc[(int64_t)f * 3 * P.N] = __ddiv_rn(cw, lm);
            const double uu = __ddiv_rn(cu, lm);
            ou[(int64_t)f * 3 * P.N] = uu;
            if (oz) oz[(int64_t)f * P.N] = uu;
            Lp = Lc;
            prow = cur;
        }
    }
    if (bad) atomicOr(P.status, PBK_ST_IMAGE_OVERFLOW);
}

// ------------------------------------------------------------------------------------------
// leaflet labels from the compact normal components zc[F][N], one CTA per frame
// (BilayerAnalyzer._build_leaflets, bilayer_analyzer.py:826-847; running_stats.py:34-52).
// The label only depends on the sign of z - zavg, so a tree mean decides every lipid that is
// further from it than the worst-case difference between the tree mean and the reference's
// sequential Welford mean; if any lipid is closer, thread 0 replays the recurrence.
// Frame g = frame0 + f takes the labels of src = g - g % interval (bilayer_analyzer.py:969-981);
// src before the batch: the row before the batch (or a copy of it).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
k_leaflets_c(const double *__restrict__ zc, int64_t N, int64_t frame0, int interval, const uint8_t *__restrict__ prev,
             uint8_t *__restrict__ labels, int32_t *status)
{
    __shared__ double red[32];
    __shared__ double zexact;
    __shared__ int need_exact;
    const int64_t f = blockIdx.x;
    const int64_t g = frame0 + f;
    const int64_t src = g - (g % interval);
    uint8_t *out = labels + f * N;
    if (src < frame0) {
        for (int64_t i = threadIdx.x; i < N; i += blockDim.x) out[i] = prev ? prev[i] : 0;
        return;
    }
    const double *z = zc + (src - frame0) * N;
    double s = 0.0, mx = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        const double v = z[i];
        s += v;
        mx = fmax(mx, fabs(v));
    }
    s = pbk_block_sum(s, red);
    mx = pbk_block_max(mx, red);
    double mean = s / (double)N;
    const double tol = 64.0 * (double)N * 2.220446049250313e-16 * mx;
    if (threadIdx.x == 0) need_exact = 0;
    __syncthreads();
    int close = 0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x)
        if (fabs(z[i] - mean) <= tol) close = 1;
    if (close) need_exact = 1;
    __syncthreads();
    if (need_exact) {
        if (threadIdx.x == 0) {
            double m = 0.0;                       // running_stats.py:34-52
            for (int64_t i = 0; i < N; i++) {
                const double v = z[i];
                m = (i == 0) ? v : __dadd_rn(m, __ddiv_rn(__dsub_rn(v, m), (double)(i + 1)));
            }
            zexact = m;
        }
        __syncthreads();
        mean = zexact;
    }
    int tie = 0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        const double v = z[i];
        if (v > mean) out[i] = 1;
        else if (v < mean) out[i] = 0;
        else { out[i] = 0; tie = 1; }
    }
    if (tie) atomicOr(status, PBK_ST_LEAFLET_TIE);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline size_t sw_align(size_t v) { return (v + 255) & ~(size_t)255; }

struct SweepScratch {
    float *boxtab; double *node_vals; float *u_out; double *zc; int16_t *snap;
    size_t total;
};

static SweepScratch sw_layout(void *base, int64_t F, int64_t M, int64_t N, int32_t L, int32_t K)
{
    SweepScratch s;
    size_t o = 0;
    unsigned char *b = reinterpret_cast<unsigned char *>(base);
    const int64_t nch = (F + SW_CHUNK - 1) / SW_CHUNK;
    s.boxtab = reinterpret_cast<float *>(b + o); o += sw_align((size_t)F * 32);
    s.node_vals = reinterpret_cast<double *>(b + o); o += sw_align((size_t)F * (L + K) * 24);
    s.u_out = reinterpret_cast<float *>(b + o); o += sw_align((size_t)M * 12);
    s.zc = reinterpret_cast<double *>(b + o); o += sw_align((size_t)F * N * 8);
    s.snap = reinterpret_cast<int16_t *>(b + o); o += sw_align((size_t)nch * M * 6);
    s.total = o;
    return s;
}

extern "C" int pbk_reduce_stream_scratch_bytes(int64_t F, int64_t M, int64_t N, int32_t L, int32_t K, int64_t *bytes)
{
    if (F <= 0 || M <= 0 || N <= 0 || L <= 0 || K < 0 || !bytes) return PBK_EINVAL;
    *bytes = (int64_t)sw_layout(nullptr, F, M, N, L, K).total;
    return PBK_OK;
}

template <int NA>
static cudaError_t sw_launch_b(const SweepBParams &B, int64_t nch, cudaStream_t st)
{
    dim3 g((unsigned)((3 * B.N + SB_THREADS - 1) / SB_THREADS), (unsigned)nch);
    const size_t smem = NA == 0 ? (size_t)B.max_atoms * SB_THREADS * 2 : 0;
    if (NA == 0) {
        cudaError_t e = cudaFuncSetAttribute(k_sweep_b<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_sweep_b<NA><<<g, SB_THREADS, smem, st>>>(B);
    return cudaGetLastError();
}

extern "C" int pbk_reduce_stream(pbk_ctx *ctx, const float *x, const float *box, const double *mass,
                                 const int32_t *seg, const int32_t *gather, const double *lipid_mass,
                                 int32_t max_bead_atoms, const int32_t *leaf_start, const int32_t *leaf_len,
                                 int32_t L, const int32_t *node_left, const int32_t *node_right, int32_t K,
                                 const int32_t *level_off, int32_t H, double total_mass, float *u_state,
                                 int first_frame, int64_t F, int64_t M, int64_t N, int norm, int phase,
                                 const int32_t *init_images, int32_t *net_images, int64_t net_frames,
                                 int64_t frame0, int update_interval, const uint8_t *prev_labels, void *scratch,
                                 int64_t scratch_bytes, double *com, double *com_unwrap, double *mem_com,
                                 uint8_t *labels, int32_t *status, void *stream)
{
    if ((phase & ~3) || phase == 0 || (init_images && !first_frame))
        return pbk_fail(ctx, PBK_EINVAL, "pbk_reduce_stream: bad phase / init_images needs first_frame");
    if (!ctx || !x || !box || !mass || !leaf_start || !leaf_len || !u_state || !scratch || !status || F < 0 ||
        M <= 0 || N <= 0 || L <= 0 || K < 0 || norm < 0 || norm > 2 || update_interval < 1 || frame0 < 0 ||
        (K > 0 && (!node_left || !node_right || !level_off)))
        return pbk_fail(ctx, PBK_EINVAL, "pbk_reduce_stream: bad argument");
    if ((phase & 2) && (!seg || !lipid_mass || !com || !com_unwrap || !mem_com || !labels || max_bead_atoms <= 0))
        return pbk_fail(ctx, PBK_EINVAL, "pbk_reduce_stream: bad argument (phase 2)");
    if (F == 0) return PBK_OK;
    if (3 * M >= ((int64_t)1 << 31) || max_bead_atoms > 800) return PBK_EUNSUPPORTED;
    const SweepScratch S = sw_layout(scratch, F, M, N, L, K);
    if ((int64_t)S.total > scratch_bytes) return pbk_fail(ctx, PBK_EINVAL, "pbk_reduce_stream: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const bool bulk = (M % 4) == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0;
    const int64_t nch = (F + SW_CHUNK - 1) / SW_CHUNK;
    k_box_table<<<(unsigned)((F + 127) / 128), 128, 0, st>>>(box, F, S.boxtab);
    PBK_CHECK_LAUNCH(ctx, "k_box_table");
    {
        SweepAParams A;
        A.x = x; A.boxtab = S.boxtab; A.mass = mass; A.leaf_start = leaf_start; A.leaf_len = leaf_len;
        A.L = L; A.nodes_per_frame = L + K; A.prev0 = first_frame ? x : u_state; A.init_images = init_images;
        A.F = F; A.M = M; A.net_frame = ((net_frames > 0 && net_frames < F) ? net_frames : F) - 1;
        A.net_images = net_images; A.u_out = S.u_out; A.node_vals = S.node_vals; A.status = status;
        A.snap = (phase & 2) ? S.snap : nullptr;
        A.store_sums = (phase & 2) ? 1 : 0;
        if (bulk) k_sweep_a<true><<<(unsigned)L, SA_THREADS, 0, st>>>(A);
        else k_sweep_a<false><<<(unsigned)L, SA_THREADS, 0, st>>>(A);
        PBK_CHECK_LAUNCH(ctx, "k_sweep_a");
    }
    if (phase & 2) {
        const int rc = pbk_launch_mem_tree(ctx, node_left, node_right, level_off, H, L, K, total_mass, F, S.node_vals,
                                           mem_com, st);
        if (rc != PBK_OK) return rc;
        SweepBParams B;
        B.x = x; B.boxtab = S.boxtab; B.mass = mass; B.seg = seg; B.gather = gather; B.lipid_mass = lipid_mass;
        B.mem_com = mem_com; B.prev0 = first_frame ? x : u_state; B.init_images = init_images; B.snap = S.snap;
        B.F = F; B.M = M; B.N = N; B.norm = norm; B.max_atoms = max_bead_atoms;
        B.com = com; B.com_unwrap = com_unwrap; B.zc = S.zc; B.status = status;
        cudaError_t e;
        if (max_bead_atoms == 1) e = sw_launch_b<1>(B, nch, st);
        else if (max_bead_atoms <= 4) e = sw_launch_b<4>(B, nch, st);
        else if (max_bead_atoms <= 8) e = sw_launch_b<8>(B, nch, st);
        else if (max_bead_atoms <= 12) e = sw_launch_b<12>(B, nch, st);
        else if (max_bead_atoms <= 16) e = sw_launch_b<16>(B, nch, st);
        else e = sw_launch_b<0>(B, nch, st);
        if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "k_sweep_b", e);
        k_leaflets_c<<<(unsigned)F, 512, 0, st>>>(S.zc, N, frame0, update_interval, prev_labels, labels, status);
        PBK_CHECK_LAUNCH(ctx, "k_leaflets_c");
    }
    cudaError_t e = cudaMemcpyAsync(u_state, S.u_out, (size_t)M * 12, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "u_state copy", e);
    return PBK_OK;
}
