// pbk_com.cu -- unwrap, membrane COM, per-lipid COM, leaflets, MSD (generic multi-kernel path).
//
// Reference arithmetic restated here (paths relative to the PyBILT tree):
//   unwrap        pybilt/mda_tools/mda_unwrap_vectorized.py:12-93
//   membrane COM  pybilt/bilayer_analyzer/com_frame.py:172-181 (NumPy pairwise sums)
//   lipid COM     pybilt/bilayer_analyzer/com_frame.py:43-96,163-185 (MDAnalysis center_of_mass)
//   leaflets      pybilt/bilayer_analyzer/bilayer_analyzer.py:826-847
//   msd           pybilt/bilayer_analyzer/analysis_protocols.py:582-667, 781-895
#include <stdlib.h>

#include "pbk_common.cuh"

// ---------------------------------------------------------------------------------------
// K1  image counts: one thread per atom component, marching over the frames of the batch.
// Loads are independent of the recurrence, so they are issued UNR frames ahead.
// ---------------------------------------------------------------------------------------
#define UNR 8

__device__ __forceinline__ int pbk_image_shift(float x, float uprev, float L, int *bad)
{
    float d = __fsub_rn(x, uprev);
    float h = L * 0.5f;
    int s = 0;
    if (!(L > 0.0f)) { *bad = 1; return 0; }
    double dd = (double)d, Ld = (double)L, hd = (double)h;
    if (d > h) {
        do { s -= 1; } while (__dadd_rn(dd, __dmul_rn((double)s, Ld)) > hd && s > -32768);
    } else if (d < -h) {
        do { s += 1; } while (__dadd_rn(dd, __dmul_rn((double)s, Ld)) < -hd && s < 32768);
    }
    if (s <= -32768 || s >= 32768) { *bad = 1; s = 0; }
    return s;
}

// mda_unwrap.py:12-85 (the scalar unwrap COMTraj uses, pybilt/com_trajectory/COMTraj.py:1067-1081):
// same rule, but `dz + shift*C` is evaluated on NumPy float32 scalars, i.e. in float32.
__device__ __forceinline__ int pbk_image_shift_f32(float x, float uprev, float L, int *bad)
{
    float d = __fsub_rn(x, uprev);
    float h = L * 0.5f;
    int s = 0;
    if (!(L > 0.0f)) { *bad = 1; return 0; }
    if (d > h) {
        do { s -= 1; } while (__fadd_rn(d, __fmul_rn((float)s, L)) > h && s > -32768);
    } else if (d < -h) {
        do { s += 1; } while (__fadd_rn(d, __fmul_rn((float)s, L)) < -h && s < 32768);
    }
    if (s <= -32768 || s >= 32768) { *bad = 1; s = 0; }
    return s;
}

// Chain state = (raw previous coordinate, its image count).  The common case -- the raw
// jump is nowhere near half a box -- keeps the image count without touching float64
// (pbk_same_image, pbk_common.cuh); otherwise the reference's loops run.
__device__ __forceinline__ int pbk_next_image(float xv, float xprev, int s, float L, float Lprev, int flags,
                                              int *bad)
{
    const int as = s < 0 ? -s : s;
    const float eps = (float)(as + 2) * fmaxf(L, Lprev) * 2.3841858e-7f + (float)as * fabsf(L - Lprev);
    if (pbk_same_image(xv, xprev, L, eps) && L > 0.0f) return s;
    const float up = pbk_unwrapped(xprev, s, Lprev);
    return (flags & PBK_UNWRAP_SCALAR_F32) ? pbk_image_shift_f32(xv, up, L, bad) : pbk_image_shift(xv, up, L, bad);
}

__global__ void __launch_bounds__(128)
k_unwrap_images(const float *__restrict__ x, const float *__restrict__ box, float *u_state,
                int first_frame, const int32_t *__restrict__ init_images, int flags, int64_t F, int64_t C,
                int16_t *__restrict__ img, int32_t *status)
{
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    int k = (int)(c % 3);
    int bad = 0;
    float xprev, Lprev;
    int s = 0;
    int64_t t = 0;
    if (first_frame) {
        // frame 0 is its own reference; in a frame shard it is a raw halo frame whose image
        // counts (init_images) are already known from the ranks before
        xprev = x[c];
        Lprev = box[k];
        s = init_images ? init_images[c] : 0;
        if (s > 32767 || s < -32767) { bad = 1; s = 0; }
        img[c] = (int16_t)s;
        t = 1;
    } else {
        xprev = u_state[c];            // an unwrapped coordinate is a raw one with image count 0
        Lprev = box[k];
    }
    for (; t + UNR <= F; t += UNR) {
        float xv[UNR], Lv[UNR];
#pragma unroll
        for (int j = 0; j < UNR; j++) {
            xv[j] = __ldg(x + (t + j) * C + c);
            Lv[j] = __ldg(box + (t + j) * 3 + k);
        }
#pragma unroll
        for (int j = 0; j < UNR; j++) {
            s = pbk_next_image(xv[j], xprev, s, Lv[j], Lprev, flags, &bad);
            xprev = xv[j];
            Lprev = Lv[j];
            img[(t + j) * C + c] = (int16_t)s;
        }
    }
    for (; t < F; t++) {
        float xv = __ldg(x + t * C + c), L = __ldg(box + t * 3 + k);
        s = pbk_next_image(xv, xprev, s, L, Lprev, flags, &bad);
        xprev = xv;
        Lprev = L;
        img[t * C + c] = (int16_t)s;
    }
    u_state[c] = pbk_unwrapped(xprev, s, Lprev);
    if (bad) atomicOr(status, PBK_ST_IMAGE_OVERFLOW);
}

extern "C" int pbk_unwrap_images(pbk_ctx *ctx, const float *x, const float *box, float *u_state,
                                 int first_frame, const int32_t *init_images, int flags, int64_t F,
                                 int64_t M, int16_t *img, int32_t *status, void *stream)
{
    if (init_images && !first_frame)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_unwrap_images: init_images needs first_frame");
    if (!ctx || !x || !box || !u_state || !img || !status || F < 0 || M <= 0)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_unwrap_images: bad argument");
    if (F == 0) return PBK_OK;
    int64_t C = 3 * M;
    int threads = 128;
    int64_t blocks = (C + threads - 1) / threads;
    k_unwrap_images<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        x, box, u_state, first_frame, init_images, flags, F, C, img, status);
    PBK_CHECK_LAUNCH(ctx, "k_unwrap_images");
    return PBK_OK;
}

// ---------------------------------------------------------------------------------------
// K2  membrane COM, step 1: NumPy pairwise leaf sums.  One warp per (frame, leaf); lanes
// 0..23 = 8 accumulator lanes x 3 axes reading 96 contiguous bytes per step.
//   r[j] = a[j]; for i = 8; i < n - n%8; i += 8: r[j] += a[i+j];
//   res = ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); then the n%8 tail sequentially.
// with a[i] = m_i * f64(u_i).   n < 8: res = 0 + a0 + a1 + ... sequentially.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_mem_leaf_sums(const float *__restrict__ x, const float *__restrict__ box,
                const int16_t *__restrict__ img, const double *__restrict__ mass, int64_t F,
                int64_t M, const int32_t *__restrict__ leaf_start,
                const int32_t *__restrict__ leaf_len, int32_t L, int32_t nodes_per_frame,
                double *__restrict__ node_vals)
{
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= F * (int64_t)L) return;
    int64_t f = w / L;
    int leaf = (int)(w % L);
    int start = leaf_start[leaf], n = leaf_len[leaf];
    int j = lane / 3, k = lane % 3;             // accumulator lane, axis (lanes >= 24 idle)
    bool act = lane < 24;
    const float *xf = x + (f * M + start) * 3;
    const int16_t *sf = img + (f * M + start) * 3;
    const double *mf = mass + start;
    float Lk = box[f * 3 + k];
    double r = 0.0;
    int nfull = n - (n % 8);
    if (n >= 8 && act) {
        r = __dmul_rn(mf[j], (double)pbk_unwrapped(xf[j * 3 + k], sf[j * 3 + k], Lk));
        for (int i = 8; i < nfull; i += 8) {
            int a = i + j;
            double p = __dmul_rn(mf[a], (double)pbk_unwrapped(xf[a * 3 + k], sf[a * 3 + k], Lk));
            r = __dadd_rn(r, p);
        }
    }
    // butterfly over the accumulator index j (lane = 3*j + k): j^1, j^2, j^4
    {
        int src1 = ((j ^ 1) * 3 + k) & 31, src2 = ((j ^ 2) * 3 + k) & 31, src4 = ((j ^ 4) * 3 + k) & 31;
        double o = __shfl_sync(0xffffffffu, r, act ? src1 : lane);
        r = __dadd_rn(r, o);
        o = __shfl_sync(0xffffffffu, r, act ? src2 : lane);
        r = __dadd_rn(r, o);
        o = __shfl_sync(0xffffffffu, r, act ? src4 : lane);
        r = __dadd_rn(r, o);
    }
    if (lane < 3) {
        int i0 = nfull;
        if (n < 8) { r = 0.0; i0 = 0; }
        for (int i = i0; i < n; i++) {
            double p = __dmul_rn(mf[i], (double)pbk_unwrapped(xf[i * 3 + k], sf[i * 3 + k], Lk));
            r = __dadd_rn(r, p);
        }
        node_vals[(f * nodes_per_frame + leaf) * 3 + k] = r;
    }
}

// K2b  step 2: combine the leaf sums along the recursion tree, level by level.
__global__ void __launch_bounds__(256)
k_mem_tree(const int32_t *__restrict__ node_left, const int32_t *__restrict__ node_right,
           const int32_t *__restrict__ level_off, int32_t H, int32_t L, int32_t K,
           double total_mass, double *__restrict__ node_vals, double *__restrict__ mem_com)
{
    int64_t f = blockIdx.x;
    double *v = node_vals + f * (int64_t)(L + K) * 3;
    for (int h = 0; h < H; h++) {
        int a = level_off[h], b = level_off[h + 1];
        for (int w = threadIdx.x; w < (b - a) * 3; w += blockDim.x) {
            int node = a + w / 3, k = w % 3;
            v[(L + node) * 3 + k] = __dadd_rn(v[node_left[node] * 3 + k], v[node_right[node] * 3 + k]);
        }
        __syncthreads();
    }
    if (threadIdx.x < 3) {
        int root = L + K - 1;
        mem_com[f * 3 + threadIdx.x] = __ddiv_rn(v[root * 3 + threadIdx.x], total_mass);
    }
}

int pbk_launch_mem_tree(pbk_ctx *ctx, const int32_t *node_left, const int32_t *node_right, const int32_t *level_off,
                        int32_t H, int32_t L, int32_t K, double total_mass, int64_t F, double *node_vals,
                        double *mem_com, cudaStream_t st)
{
    k_mem_tree<<<(unsigned)F, 256, 0, st>>>(node_left, node_right, level_off, H, L, K, total_mass, node_vals, mem_com);
    PBK_CHECK_LAUNCH(ctx, "k_mem_tree");
    return PBK_OK;
}

extern "C" int pbk_membrane_com(pbk_ctx *ctx, const float *x, const float *box, const int16_t *img,
                                const double *mass, int64_t F, int64_t M,
                                const int32_t *leaf_start, const int32_t *leaf_len, int32_t L,
                                const int32_t *node_left, const int32_t *node_right, int32_t K,
                                const int32_t *level_off, int32_t H, double total_mass,
                                double *node_vals, double *mem_com, void *stream)
{
    if (!ctx || !x || !box || !img || !mass || !leaf_start || !leaf_len || !node_vals || !mem_com ||
        F < 0 || M <= 0 || L <= 0 || K < 0 || (K > 0 && (!node_left || !node_right || !level_off)))
        return pbk_fail(ctx, PBK_EINVAL, "pbk_membrane_com: bad argument");
    if (F == 0) return PBK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t warps = F * (int64_t)L;
    int64_t blocks = (warps * 32 + 255) / 256;
    k_mem_leaf_sums<<<(unsigned)blocks, 256, 0, st>>>(x, box, img, mass, F, M, leaf_start, leaf_len,
                                                      L, L + K, node_vals);
    PBK_CHECK_LAUNCH(ctx, "k_mem_leaf_sums");
    k_mem_tree<<<(unsigned)F, 256, 0, st>>>(node_left, node_right, level_off, H, L, K, total_mass,
                                            node_vals, mem_com);
    PBK_CHECK_LAUNCH(ctx, "k_mem_tree");
    return PBK_OK;
}

// K2c  membrane COM the way COMTraj forms it: mem_sel.center_of_mass()
// (pybilt/com_trajectory/COMTraj.py:1084) = (pos.astype(f64) * m[:,None]).sum(axis=0) / m.sum(),
// whose axis-0 sum accumulates the rows one after the other.  One thread per (frame, axis).
__global__ void __launch_bounds__(128)
k_mem_seq(const float *__restrict__ x, const float *__restrict__ box, const int16_t *__restrict__ img,
          const double *__restrict__ mass, int64_t F, int64_t M, double total_mass,
          double *__restrict__ mem_com)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= F * 3) return;
    int64_t f = g / 3;
    int k = (int)(g % 3);
    const float *xf = x + f * M * 3 + k;
    const int16_t *sf = img + f * M * 3 + k;
    float L = box[f * 3 + k];
    double acc = 0.0;
    int64_t a = 0;
    for (; a + 4 <= M; a += 4) {
        double p[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
            p[j] = __dmul_rn((double)pbk_unwrapped(xf[(a + j) * 3], sf[(a + j) * 3], L), mass[a + j]);
#pragma unroll
        for (int j = 0; j < 4; j++) acc = __dadd_rn(acc, p[j]);
    }
    for (; a < M; a++)
        acc = __dadd_rn(acc, __dmul_rn((double)pbk_unwrapped(xf[a * 3], sf[a * 3], L), mass[a]));
    mem_com[g] = __ddiv_rn(acc, total_mass);
}

extern "C" int pbk_membrane_com_seq(pbk_ctx *ctx, const float *x, const float *box, const int16_t *img,
                                    const double *mass, int64_t F, int64_t M, double total_mass,
                                    double *mem_com, void *stream)
{
    if (!ctx || !x || !box || !img || !mass || !mem_com || F < 0 || M <= 0)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_membrane_com_seq: bad argument");
    if (F == 0) return PBK_OK;
    k_mem_seq<<<(unsigned)((F * 3 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, box, img, mass, F, M,
                                                                                total_mass, mem_com);
    PBK_CHECK_LAUNCH(ctx, "k_mem_seq");
    return PBK_OK;
}

// ---------------------------------------------------------------------------------------
// K3  per-lipid COM, wrapped (raw x) and unwrapped (v = f32(f64(u) - mem_com)), one thread
// per (frame, lipid), atoms accumulated in list order.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_lipid_com(const float *__restrict__ x, const float *__restrict__ box,
            const int16_t *__restrict__ img, const double *__restrict__ mass,
            const double *__restrict__ mem_com, const int32_t *__restrict__ seg,
            const int32_t *__restrict__ gather, const double *__restrict__ lipid_mass, int64_t F,
            int64_t M, int64_t N, double *__restrict__ com, double *__restrict__ com_unwrap)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= F * N) return;
    int64_t f = g / N;
    int i = (int)(g % N);
    const float *xf = x + f * M * 3;
    const int16_t *sf = img + f * M * 3;
    float L0 = box[f * 3], L1 = box[f * 3 + 1], L2 = box[f * 3 + 2];
    double m0 = mem_com[f * 3], m1 = mem_com[f * 3 + 1], m2 = mem_com[f * 3 + 2];
    double c0 = 0, c1 = 0, c2 = 0, w0 = 0, w1 = 0, w2 = 0;
    int a0 = seg[i], a1 = seg[i + 1];
    for (int p = a0; p < a1; p++) {
        int a = gather ? gather[p] : p;
        double m = mass[a];
        float x0 = xf[a * 3], x1 = xf[a * 3 + 1], x2 = xf[a * 3 + 2];
        c0 = __dadd_rn(c0, __dmul_rn((double)x0, m));
        c1 = __dadd_rn(c1, __dmul_rn((double)x1, m));
        c2 = __dadd_rn(c2, __dmul_rn((double)x2, m));
        float u0 = pbk_unwrapped(x0, sf[a * 3], L0);
        float u1 = pbk_unwrapped(x1, sf[a * 3 + 1], L1);
        float u2 = pbk_unwrapped(x2, sf[a * 3 + 2], L2);
        float v0 = __double2float_rn(__dsub_rn((double)u0, m0));
        float v1 = __double2float_rn(__dsub_rn((double)u1, m1));
        float v2 = __double2float_rn(__dsub_rn((double)u2, m2));
        w0 = __dadd_rn(w0, __dmul_rn((double)v0, m));
        w1 = __dadd_rn(w1, __dmul_rn((double)v1, m));
        w2 = __dadd_rn(w2, __dmul_rn((double)v2, m));
    }
    double lm = lipid_mass[i];
    double *co = com + g * 3, *cu = com_unwrap + g * 3;
    co[0] = __ddiv_rn(c0, lm); co[1] = __ddiv_rn(c1, lm); co[2] = __ddiv_rn(c2, lm);
    cu[0] = __ddiv_rn(w0, lm); cu[1] = __ddiv_rn(w1, lm); cu[2] = __ddiv_rn(w2, lm);
}

extern "C" int pbk_lipid_com(pbk_ctx *ctx, const float *x, const float *box, const int16_t *img,
                             const double *mass, const double *mem_com, const int32_t *seg,
                             const int32_t *gather, const double *lipid_mass, int64_t F, int64_t M,
                             int64_t N, double *com, double *com_unwrap, void *stream)
{
    if (!ctx || !x || !box || !img || !mass || !mem_com || !seg || !lipid_mass || !com ||
        !com_unwrap || F < 0 || M <= 0 || N <= 0)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_lipid_com: bad argument");
    if (F == 0) return PBK_OK;
    int64_t total = F * N;
    int64_t blocks = (total + 127) / 128;
    k_lipid_com<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(
        x, box, img, mass, mem_com, seg, gather, lipid_mass, F, M, N, com, com_unwrap);
    PBK_CHECK_LAUNCH(ctx, "k_lipid_com");
    return PBK_OK;
}

// ---------------------------------------------------------------------------------------
// K4  leaflets.  One CTA per frame.  The label only depends on the sign of z - zavg, so a
// tree mean decides every lipid that is further from it than the worst-case difference
// between the tree mean and the reference's sequential Welford mean; if any lipid is
// closer, thread 0 replays the sequential Welford recurrence and all lipids are compared
// with that exact value.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_leaflets(const double *__restrict__ com_unwrap, int64_t F, int64_t N, int norm, int64_t frame0,
           int interval, const uint8_t *__restrict__ prev, uint8_t *__restrict__ labels,
           int32_t *status)
{
    __shared__ double red[32];
    __shared__ double zexact;
    __shared__ int need_exact;
    int64_t f = blockIdx.x;
    int64_t g = frame0 + f;
    int64_t src = g - (g % interval);
    uint8_t *out = labels + f * N;
    if (src < frame0) {
        // copied from the row before this batch (or from an earlier copy of it)
        for (int64_t i = threadIdx.x; i < N; i += blockDim.x) out[i] = prev ? prev[i] : 0;
        return;
    }
    const double *z = com_unwrap + (src - frame0) * N * 3 + norm;
    double s = 0.0, mx = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        double v = z[i * 3];
        s += v;
        mx = fmax(mx, fabs(v));
    }
    s = pbk_block_sum(s, red);
    mx = pbk_block_max(mx, red);
    double mean = s / (double)N;
    double tol = 64.0 * (double)N * 2.220446049250313e-16 * mx;
    if (threadIdx.x == 0) need_exact = 0;
    __syncthreads();
    int close = 0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x)
        if (fabs(z[i * 3] - mean) <= tol) close = 1;
    if (close) need_exact = 1;
    __syncthreads();
    if (need_exact) {
        if (threadIdx.x == 0) {
            double m = 0.0;                       // running_stats.py:34-52
            for (int64_t i = 0; i < N; i++) {
                double v = z[i * 3];
                m = (i == 0) ? v : __dadd_rn(m, __ddiv_rn(__dsub_rn(v, m), (double)(i + 1)));
            }
            zexact = m;
        }
        __syncthreads();
        mean = zexact;
    }
    int tie = 0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        double v = z[i * 3];
        if (v > mean) out[i] = 1;
        else if (v < mean) out[i] = 0;
        else { out[i] = 0; tie = 1; }
    }
    if (tie) atomicOr(status, PBK_ST_LEAFLET_TIE);
}

extern "C" int pbk_leaflets(pbk_ctx *ctx, const double *com_unwrap, int64_t F, int64_t N, int norm,
                            int64_t frame0, int update_interval, const uint8_t *prev_labels,
                            uint8_t *labels, int32_t *status, void *stream)
{
    if (!ctx || !com_unwrap || !labels || !status || F < 0 || N <= 0 || norm < 0 || norm > 2 ||
        update_interval < 1 || frame0 < 0)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_leaflets: bad argument");
    if (F == 0) return PBK_OK;
    k_leaflets<<<(unsigned)F, 256, 0, (cudaStream_t)stream>>>(com_unwrap, F, N, norm, frame0,
                                                             update_interval, prev_labels, labels,
                                                             status);
    PBK_CHECK_LAUNCH(ctx, "k_leaflets");
    return PBK_OK;
}

// ---------------------------------------------------------------------------------------
// K5  MSD over (origin, end) frame pairs; one CTA per pair, fixed-order tree mean.
// ---------------------------------------------------------------------------------------
// mode 0: mean squared displacement; mode 1: mean displacement length (ald: sqrt(np.dot(dr, dr)))
__global__ void __launch_bounds__(256)
k_msd_pairs(const double *__restrict__ cu, int64_t N, const int32_t *__restrict__ idx, int32_t n,
            int lat0, int lat1, const int32_t *__restrict__ po, const int32_t *__restrict__ pe,
            double *__restrict__ out, int mode)
{
    __shared__ double red[32];
    int p = blockIdx.x;
    const double *a = cu + (int64_t)pe[p] * N * 3;
    const double *b = cu + (int64_t)po[p] * N * 3;
    double s = 0.0;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        int64_t i = idx[t];
        double dx = __dsub_rn(a[i * 3 + lat0], b[i * 3 + lat0]);
        double dy = __dsub_rn(a[i * 3 + lat1], b[i * 3 + lat1]);
        if (mode == 0) s += __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        else s += sqrt(fma(dy, dy, __dmul_rn(dx, dx)));
    }
    s = pbk_block_sum(s, red);
    if (threadIdx.x == 0) out[p] = n > 0 ? s / (double)n : 0.0;
}

// Large systems: the index list of an MSD analysis ("upper members, then lower members") walks the COM rows
// with a stride, so one CTA per pair reads every row's sector without using half of it and has little in
// flight.  Instead the list becomes a per-lipid weight (how often the lipid is listed), and a grid of
// (slice of lipids, pair) CTAs streams the rows of both frames in order -- the sum does not depend on the
// order of the list -- with four rows in flight per thread; a second small kernel adds the slice partials in
// slice order and divides.
#define MSD_SLICE 4096
__global__ void k_msd_weights(const int32_t *__restrict__ idx, int32_t n, int32_t *__restrict__ w)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) atomicAdd(&w[idx[t]], 1);
}

__global__ void __launch_bounds__(256)
k_msd_slices(const double *__restrict__ cu, int64_t N, const int32_t *__restrict__ w, int lat0, int lat1,
             const int32_t *__restrict__ po, const int32_t *__restrict__ pe, double *__restrict__ part, int mode)
{
    __shared__ double red[32];
    const int p = blockIdx.y;
    const double *a = cu + (int64_t)pe[p] * N * 3;
    const double *b = cu + (int64_t)po[p] * N * 3;
    const int64_t i0 = (int64_t)blockIdx.x * MSD_SLICE;
    double s = 0.0;
#pragma unroll 4
    for (int r = threadIdx.x; r < MSD_SLICE; r += 256) {
        const int64_t i = i0 + r;
        if (i < N) {
            const int wi = __ldg(w + i);
            const double dx = __dsub_rn(a[i * 3 + lat0], b[i * 3 + lat0]);
            const double dy = __dsub_rn(a[i * 3 + lat1], b[i * 3 + lat1]);
            const double v = mode == 0 ? __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) : sqrt(fma(dy, dy, __dmul_rn(dx, dx)));
            s += (double)wi * v;
        }
    }
    s = pbk_block_sum(s, red);
    if (threadIdx.x == 0) part[(int64_t)p * gridDim.x + blockIdx.x] = s;
}

__global__ void k_msd_final(const double *__restrict__ part, int n_slices, int32_t n, int32_t n_pairs, double *__restrict__ out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    double s = 0.0;
    for (int q = 0; q < n_slices; q++) s += part[(int64_t)p * n_slices + q];
    out[p] = n > 0 ? s / (double)n : 0.0;
}

static int pbk_msd_run(pbk_ctx *ctx, const double *com_unwrap, int64_t N, const int32_t *idx, int32_t n_idx, int lat0,
                       int lat1, const int32_t *pair_origin, const int32_t *pair_end, int32_t n_pairs, double *out,
                       int mode, cudaStream_t st)
{
    const int64_t n_slices = (N + MSD_SLICE - 1) / MSD_SLICE;
    if (N >= 16384 && !getenv("PBK_MSD_SIMPLE")) {
        const size_t need = (size_t)N * 4 + (size_t)n_pairs * n_slices * 8 + 256;
        unsigned char *ar = reinterpret_cast<unsigned char *>(pbk_arena(ctx, need));
        if (ar) {
            int32_t *w = reinterpret_cast<int32_t *>(ar);
            double *part = reinterpret_cast<double *>(ar + (((size_t)N * 4 + 255) & ~(size_t)255));
            cudaError_t e = cudaMemsetAsync(w, 0, (size_t)N * 4, st);
            if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "msd weights memset", e);
            if (n_idx > 0) k_msd_weights<<<(n_idx + 255) / 256, 256, 0, st>>>(idx, n_idx, w);
            for (int32_t p0 = 0; p0 < n_pairs; p0 += 65535) {      // (pairs are the y dimension of the grid)
                const int32_t np = n_pairs - p0 < 65535 ? n_pairs - p0 : 65535;
                k_msd_slices<<<dim3((unsigned)n_slices, (unsigned)np), 256, 0, st>>>(
                    com_unwrap, N, w, lat0, lat1, pair_origin + p0, pair_end + p0, part + (size_t)p0 * n_slices, mode);
            }
            k_msd_final<<<(n_pairs + 127) / 128, 128, 0, st>>>(part, (int)n_slices, n_idx, n_pairs, out);
            PBK_CHECK_LAUNCH(ctx, "k_msd_slices");
            return PBK_OK;
        }
    }
    k_msd_pairs<<<(unsigned)n_pairs, 256, 0, st>>>(com_unwrap, N, idx, n_idx, lat0, lat1, pair_origin, pair_end, out, mode);
    PBK_CHECK_LAUNCH(ctx, "k_msd_pairs");
    return PBK_OK;
}

extern "C" int pbk_msd_pairs(pbk_ctx *ctx, const double *com_unwrap, int64_t F, int64_t N,
                             const int32_t *idx, int32_t n_idx, int lat0, int lat1,
                             const int32_t *pair_origin, const int32_t *pair_end, int32_t n_pairs,
                             double *out, void *stream)
{
    if (!ctx || !com_unwrap || !idx || !pair_origin || !pair_end || !out || F <= 0 || N <= 0 ||
        n_idx < 0 || n_pairs < 0 || lat0 < 0 || lat0 > 2 || lat1 < 0 || lat1 > 2)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_msd_pairs: bad argument");
    if (n_pairs == 0) return PBK_OK;
    return pbk_msd_run(ctx, com_unwrap, N, idx, n_idx, lat0, lat1, pair_origin, pair_end, n_pairs, out, 0,
                       (cudaStream_t)stream);
}

extern "C" int pbk_ald_pairs(pbk_ctx *ctx, const double *com_unwrap, int64_t F, int64_t N,
                             const int32_t *idx, int32_t n_idx, int lat0, int lat1,
                             const int32_t *pair_origin, const int32_t *pair_end, int32_t n_pairs,
                             double *out, void *stream)
{
    if (!ctx || !com_unwrap || !idx || !pair_origin || !pair_end || !out || F <= 0 || N <= 0 ||
        n_idx < 0 || n_pairs < 0 || lat0 < 0 || lat0 > 2 || lat1 < 0 || lat1 > 2)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_ald_pairs: bad argument");
    if (n_pairs == 0) return PBK_OK;
    return pbk_msd_run(ctx, com_unwrap, N, idx, n_idx, lat0, lat1, pair_origin, pair_end, n_pairs, out, 1,
                       (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------
// K0  trajectory ingestion: a DCD frame holds the coordinates as three records X[], Y[], Z[] of
// ALL atoms (between Fortran record markers, behind an optional unit-cell record); the reduction
// wants MDAnalysis' layout [selected atom][xyz].  The frames arrive on the device as the raw
// bytes of the file; one thread per (frame, selected atom) makes three coalesced record reads
// (gathered through `sel` when the selection is not everything, byte-swapped for files written
// on a big-endian machine) and one 12-byte write; a warp writes 384 contiguous bytes.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_planes_to_aos(const float *__restrict__ frames, const int32_t *__restrict__ sel, int64_t frame_stride,
                int64_t off_x, int64_t off_y, int64_t off_z, int byteswap, int64_t M, float *__restrict__ out)
{
    const int64_t f = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int64_t a = sel ? (int64_t)sel[i] : i;
    const unsigned *p = reinterpret_cast<const unsigned *>(frames + f * frame_stride);
    unsigned vx = p[off_x + a], vy = p[off_y + a], vz = p[off_z + a];
    if (byteswap) {
        vx = __byte_perm(vx, 0, 0x0123);
        vy = __byte_perm(vy, 0, 0x0123);
        vz = __byte_perm(vz, 0, 0x0123);
    }
    unsigned *o = reinterpret_cast<unsigned *>(out + (f * M + i) * 3);
    o[0] = vx;
    o[1] = vy;
    o[2] = vz;
}

extern "C" int pbk_planes_to_aos(pbk_ctx *ctx, const float *frames, const int32_t *sel, int64_t F,
                                 int64_t frame_stride, int64_t off_x, int64_t off_y, int64_t off_z,
                                 int byteswap, int64_t M, float *out, void *stream)
{
    if (!ctx || !frames || !out || F < 0 || frame_stride <= 0 || off_x < 0 || off_y < 0 || off_z < 0 ||
        M <= 0 || F > 65535)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_planes_to_aos: bad argument (F <= 65535 per call)");
    if (F == 0) return PBK_OK;
    dim3 g((unsigned)((M + 255) / 256), (unsigned)F);
    k_planes_to_aos<<<g, 256, 0, (cudaStream_t)stream>>>(frames, sel, frame_stride, off_x, off_y, off_z,
                                                         byteswap, M, out);
    PBK_CHECK_LAUNCH(ctx, "k_planes_to_aos");
    return PBK_OK;
}

// ---------------------------------------------------------------------------------------
// K10  thickness from the leaflets' centres of mass (thickness_leaflet_distance):
// |mean_upper(com[norm]) - mean_lower(com[norm])| per frame.  NumPy's mean(axis=0) of the [n, 3]
// coordinate array adds the rows in lipid order, so the sums are sequential: one warp per frame
// loads labels and coordinates coalesced and replays the order through ballots (two independent
// chains, one per leaflet).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_leaflet_distance(const double *__restrict__ com, const uint8_t *__restrict__ labels, int64_t F, int64_t N,
                   int norm, double *__restrict__ out)
{
    const int64_t f = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (f >= F) return;
    const double *cf = com + f * N * 3;
    const uint8_t *lab = labels + f * N;
    double su = 0.0, sl = 0.0;
    long long nu = 0, nl = 0;
    for (int64_t base = 0; base < N; base += 32) {
        const int64_t i = base + lane;
        double z = 0.0;
        int l = -1;
        if (i < N) { z = cf[i * 3 + norm]; l = lab[i] ? 1 : 0; }
        unsigned mu = __ballot_sync(0xffffffffu, l == 1), ml = __ballot_sync(0xffffffffu, l == 0);
        nu += __popc(mu);
        nl += __popc(ml);
        while (mu) { const int s = __ffs(mu) - 1; mu &= mu - 1; su = __dadd_rn(su, __shfl_sync(0xffffffffu, z, s)); }
        while (ml) { const int s = __ffs(ml) - 1; ml &= ml - 1; sl = __dadd_rn(sl, __shfl_sync(0xffffffffu, z, s)); }
    }
    if (lane == 0) {
        // an empty leaflet: numpy's mean of nothing is nan
        const double mu_ = nu ? __ddiv_rn(su, (double)nu) : nan("");
        const double ml_ = nl ? __ddiv_rn(sl, (double)nl) : nan("");
        out[f] = fabs(__dsub_rn(mu_, ml_));
    }
}

extern "C" int pbk_leaflet_distance(pbk_ctx *ctx, const double *com, const uint8_t *labels, int64_t F, int64_t N,
                                    int norm, double *out, void *stream)
{
    if (!ctx || !com || !labels || !out || F < 0 || N <= 0 || norm < 0 || norm > 2)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_leaflet_distance: bad argument");
    if (F == 0) return PBK_OK;
    k_leaflet_distance<<<(unsigned)((F + 3) / 4), 128, 0, (cudaStream_t)stream>>>(com, labels, F, N, norm, out);
    PBK_CHECK_LAUNCH(ctx, "k_leaflet_distance");
    return PBK_OK;
}

// ---------------------------------------------------------------------------------------
// K12  flip-flop detection (flip_flop): how many lipids changed leaflet between consecutive frames.
// One warp per frame; frame 0 of the call is compared with prev_row (NULL: nothing to compare, 0).
// The events themselves (which lipid, in the reference's set-iteration order) are listed on the host
// for the flagged frames only.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_label_changes(const uint8_t *__restrict__ labels, const uint8_t *__restrict__ prev_row, int64_t F, int64_t N,
                int32_t *__restrict__ changed)
{
    const int64_t f = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (f >= F) return;
    const uint8_t *cur = labels + f * N;
    const uint8_t *prv = f > 0 ? labels + (f - 1) * N : prev_row;
    int c = 0;
    if (prv)
        for (int64_t i = lane; i < N; i += 32) c += (cur[i] != prv[i]) ? 1 : 0;
    c = pbk_warp_sum(c);
    if (lane == 0) changed[f] = c;
}

extern "C" int pbk_label_changes(pbk_ctx *ctx, const uint8_t *labels, const uint8_t *prev_row, int64_t F, int64_t N,
                                 int32_t *changed, void *stream)
{
    if (!ctx || !labels || !changed || F < 0 || N <= 0)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_label_changes: bad argument");
    if (F == 0) return PBK_OK;
    k_label_changes<<<(unsigned)((F + 3) / 4), 128, 0, (cudaStream_t)stream>>>(labels, prev_row, F, N, changed);
    PBK_CHECK_LAUNCH(ctx, "k_label_changes");
    return PBK_OK;
}
