// pbk_fused.cu -- fused frame reduction for small systems (one frame fits in shared memory):
// unwrap + membrane COM + per-lipid COMs + leaflet labels with ONE pass over the positions
// (plus a streaming pre-pass that only counts periodic images).
//
// Reference arithmetic restated (paths relative to the PyBILT tree), identical to pbk_com.cu:
//   unwrap        pybilt/mda_tools/mda_unwrap_vectorized.py:12-93
//   membrane COM  pybilt/bilayer_analyzer/com_frame.py:172-181 (NumPy pairwise order)
//   lipid COM     pybilt/bilayer_analyzer/com_frame.py:43-96,163-185
//   leaflets      pybilt/bilayer_analyzer/bilayer_analyzer.py:826-847
//
// Why three kernels.  The unwrap is a recurrence over frames, but with only 3*M chains a small
// system cannot fill the GPU chain-by-chain.  So the frames are cut into one contiguous chunk
// per SM and
//   k_fused_agg    streams the positions once and, per (chunk, chain), sums the image-count
//                  increments delta_t, which depend on x_t and x_{t-1} only as long as the jump
//                  is not within rounding distance of half a box (the "slack" it records);
//   k_fused_scan   prefix-sums the chunk totals per chain -> image count at every chunk start,
//                  and checks every recorded slack against the rounding bound that the now
//                  known image counts imply; a violation raises PBK_ST_UNWRAP_AMBIGUOUS and
//                  the host redoes the batch with the sequential kernels of pbk_com.cu;
//   k_fused_frames one CTA per chunk walks its frames with the EXACT reference recurrence:
//                  frame t+1 arrives by TMA bulk copy (cp.async.bulk + mbarrier) into one of two
//                  shared-memory frame buffers while frame t is reduced out of the other:
//                  image counts (int8, shared), NumPy-pairwise membrane COM (warp per leaf,
//                  tree in shared memory), per-lipid wrapped / unwrapped COMs (one thread per
//                  lipid and role, atoms in order), leaflet labels.
// HBM traffic: positions twice (once per streaming pass), outputs once.
#include <stdlib.h>

#include "pbk_common.cuh"

#define FUSED_THREADS 1024
#define FUSED_TAB 1024             // ints of summation-plan tables staged in shared memory
#define FUSED_UNR 8

// ------------------------------------------------------------------------------------------
// image shift of the reference rule for difference d (float32) and box length L; also returns
// the distance of the shifted difference from the nearer window boundary +-L/2
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int fused_shift(float d, float L, float *slack, int *bad)
{
    const float h = L * 0.5f;
    int s = 0;
    if (!(L > 0.0f)) { *bad = 1; *slack = 0.f; return 0; }
    if (d > h) {
        const double dd = (double)d, Ld = (double)L, hd = (double)h;
        do { s -= 1; } while (__dadd_rn(dd, __dmul_rn((double)s, Ld)) > hd && s > -32768);
        *slack = (float)(hd - fabs(__dadd_rn(dd, __dmul_rn((double)s, Ld))));
    } else if (d < -h) {
        const double dd = (double)d, Ld = (double)L, hd = (double)h;
        do { s += 1; } while (__dadd_rn(dd, __dmul_rn((double)s, Ld)) < -hd && s < 32768);
        *slack = (float)(hd - fabs(__dadd_rn(dd, __dmul_rn((double)s, Ld))));
    } else {
        *slack = h - fabsf(d);
    }
    if (s <= -32768 || s >= 32768) { *bad = 1; s = 0; }
    return s;
}

// Out-of-line slow path: |d| > L/2.  The image count is found with float32 arithmetic and
// confirmed with the reference's float64 comparisons (the loops then run 0 extra times unless
// the estimate sits within rounding of a window boundary).
__device__ __noinline__ int fused_shift_far(float d, float L, float *slack, int *bad)
{
    return fused_shift(d, L, slack, bad);
}

__device__ __forceinline__ int fused_step(float xv, float prevx, float L, float *smin, int *bad)
{
    const float d = __fsub_rn(xv, prevx);
    const float h = L * 0.5f;
    const float sl = h - fabsf(d);
    if (sl < 0.0f || !(L > 0.0f)) {               // |d| > h  (rare): out of line
        float s2;
        const int s = fused_shift_far(d, L, &s2, bad);
        *smin = fminf(*smin, s2);
        return s;
    }
    *smin = fminf(*smin, sl);
    return 0;
}

// k_fused_agg: per (chunk, chain) the sum of image-count increments, its running |max| and
// the smallest slack.  VEC = 4: each thread owns 4 consecutive chains and streams them with
// 128-bit loads (needs 3*M % 4 == 0); VEC = 1 is the generic variant.
#define AGG_THREADS 512
#define AGG_SPLIT 2                 // CTAs per chunk (each takes a slice of the chains)
template <int VEC>
__global__ void __launch_bounds__(AGG_THREADS, 2)
k_fused_agg(const float *__restrict__ x, const float *__restrict__ box,
            const float *__restrict__ u_state, int first_frame, int64_t F, int C3, int chunk,
            int32_t *__restrict__ agg, int32_t *__restrict__ amax, float *__restrict__ slack_out,
            float *__restrict__ chunk_box, int32_t *status)
{
    const int kc = blockIdx.x;
    const int64_t f0 = (int64_t)kc * chunk, f1 = min(F, f0 + chunk);
    int bad = 0;
    if (threadIdx.x == 0 && blockIdx.y == 0) {   // largest box edge and largest frame-to-frame change
        float Lmax = 0.f, dL = 0.f;
        for (int64_t t = f0; t < f1; t++)
            for (int k = 0; k < 3; k++) {
                float L = box[t * 3 + k];
                Lmax = fmaxf(Lmax, L);
                if (t > 0) dL = fmaxf(dL, fabsf(L - box[(t - 1) * 3 + k]));
            }
        chunk_box[kc * 2] = Lmax;
        chunk_box[kc * 2 + 1] = dL;
    }
    const int nq = C3 / VEC;
    for (int q = blockIdx.y * blockDim.x + threadIdx.x; q < nq; q += blockDim.x * gridDim.y) {
        const int c0 = q * VEC;
        int kk[VEC];
        float prevx[VEC], smin[VEC];
        int acc[VEC], am[VEC];
#pragma unroll
        for (int e = 0; e < VEC; e++) { kk[e] = (c0 + e) % 3; smin[e] = 3.0e38f; acc[e] = 0; am[e] = 0; }
        int64_t t = f0;
        if (kc == 0) {
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                prevx[e] = x[c0 + e];
                if (!first_frame) {         // exact rule against the carried unwrapped state
                    float sl;
                    acc[e] = fused_shift(__fsub_rn(prevx[e], u_state[c0 + e]), box[kk[e]], &sl, &bad);
                    am[e] = abs(acc[e]);
                }
            }
            t = 1;
        } else {
#pragma unroll
            for (int e = 0; e < VEC; e++) prevx[e] = x[(f0 - 1) * C3 + c0 + e];
        }
        for (; t < f1; t += 4) {
            float xv[4][VEC];
            float Lb[4][3];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int64_t tt = min(t + j, f1 - 1);
                if (VEC == 4) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(x + tt * C3 + c0));
                    xv[j][0] = v.x; xv[j][1 % VEC] = v.y; xv[j][2 % VEC] = v.z; xv[j][3 % VEC] = v.w;
                } else {
                    xv[j][0] = __ldg(x + tt * C3 + c0);
                }
                Lb[j][0] = __ldg(box + tt * 3); Lb[j][1] = __ldg(box + tt * 3 + 1); Lb[j][2] = __ldg(box + tt * 3 + 2);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (t + j < f1) {
#pragma unroll
                    for (int e = 0; e < VEC; e++) {
                        const float L = kk[e] == 0 ? Lb[j][0] : (kk[e] == 1 ? Lb[j][1] : Lb[j][2]);
                        const int s = fused_step(xv[j][e], prevx[e], L, &smin[e], &bad);
                        if (s) { acc[e] += s; am[e] = max(am[e], abs(acc[e])); }
                        prevx[e] = xv[j][e];
                    }
                }
            }
        }
#pragma unroll
        for (int e = 0; e < VEC; e++) {
            agg[(int64_t)kc * C3 + c0 + e] = acc[e];
            amax[(int64_t)kc * C3 + c0 + e] = am[e];
            slack_out[(int64_t)kc * C3 + c0 + e] = smin[e];
        }
    }
    if (bad) atomicOr(status, PBK_ST_IMAGE_OVERFLOW);
}

// k_fused_agg12: the streaming variant for 3*M % 12 == 0.  A thread owns 4 consecutive atoms
// (12 chains, three 128-bit loads per frame), so the axis of every register is fixed (x,y,z,
// x,y,z,...) and the common case costs one subtract, one |.|-subtract and one min per chain.
// The slack and |image| maxima are kept per thread (the minimum / maximum over its 12 chains
// is a conservative stand-in for each of them) and written out per chain.
__global__ void __launch_bounds__(AGG_THREADS, 2)
k_fused_agg12(const float *__restrict__ x, const float *__restrict__ box,
              const float *__restrict__ u_state, int first_frame, int64_t F, int C3, int chunk,
              int32_t *__restrict__ agg, int32_t *__restrict__ amax, float *__restrict__ slack_out,
              float *__restrict__ chunk_box, int32_t *status)
{
    const int kc = blockIdx.x;
    const int64_t f0 = (int64_t)kc * chunk, f1 = min(F, f0 + chunk);
    int bad = 0;
    if (threadIdx.x == 0 && blockIdx.y == 0) {
        float Lmax = 0.f, dL = 0.f;
        for (int64_t t = f0; t < f1; t++)
            for (int k = 0; k < 3; k++) {
                float L = box[t * 3 + k];
                Lmax = fmaxf(Lmax, L);
                if (t > 0) dL = fmaxf(dL, fabsf(L - box[(t - 1) * 3 + k]));
            }
        chunk_box[kc * 2] = Lmax;
        chunk_box[kc * 2 + 1] = dL;
    }
    const int nq = C3 / 12;
    for (int q = blockIdx.y * blockDim.x + threadIdx.x; q < nq; q += blockDim.x * gridDim.y) {
        const int c0 = q * 12;
        float prevx[12];
        int acc[12];
        float smin = 3.0e38f;
        int am = 0;
#pragma unroll
        for (int e = 0; e < 12; e++) acc[e] = 0;
        int64_t t = f0;
        if (kc == 0) {
#pragma unroll
            for (int e = 0; e < 12; e++) {
                prevx[e] = x[c0 + e];
                if (!first_frame) {         // exact rule against the carried unwrapped state
                    float sl;
                    acc[e] = fused_shift(__fsub_rn(prevx[e], u_state[c0 + e]), box[e % 3], &sl, &bad);
                    am = max(am, abs(acc[e]));
                }
            }
            t = 1;
        } else {
#pragma unroll
            for (int e = 0; e < 12; e++) prevx[e] = x[(f0 - 1) * C3 + c0 + e];
        }
        for (; t < f1; t += 2) {
            float xv[2][12], hb[2][3], Lb[2][3];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int64_t tt = min(t + j, f1 - 1);
                const float4 *src = reinterpret_cast<const float4 *>(x + tt * C3 + c0);
#pragma unroll
                for (int v = 0; v < 3; v++) {
                    const float4 w = __ldg(src + v);
                    xv[j][v * 4] = w.x; xv[j][v * 4 + 1] = w.y; xv[j][v * 4 + 2] = w.z; xv[j][v * 4 + 3] = w.w;
                }
#pragma unroll
                for (int k = 0; k < 3; k++) { Lb[j][k] = __ldg(box + tt * 3 + k); hb[j][k] = Lb[j][k] * 0.5f; }
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                if (t + j < f1) {
#pragma unroll
                    for (int e = 0; e < 12; e++) {
                        const float sl = hb[j][e % 3] - fabsf(__fsub_rn(xv[j][e], prevx[e]));
                        if (!(sl > 0.0f)) {                 // rare: |jump| >= L/2 (or bad box)
                            float s2;
                            acc[e] += fused_shift_far(__fsub_rn(xv[j][e], prevx[e]), Lb[j][e % 3], &s2, &bad);
                            am = max(am, abs(acc[e]));
                            smin = fminf(smin, s2);
                        } else {
                            smin = fminf(smin, sl);
                        }
                        prevx[e] = xv[j][e];
                    }
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 12; e++) {
            agg[(int64_t)kc * C3 + c0 + e] = acc[e];
            amax[(int64_t)kc * C3 + c0 + e] = am;
            slack_out[(int64_t)kc * C3 + c0 + e] = smin;
        }
    }
    if (bad) atomicOr(status, PBK_ST_IMAGE_OVERFLOW);
}

// k_fused_scan: per chain, exclusive prefix of the chunk totals (= image count at every chunk
// start) and the slack check.  256 threads = 32 chains x 8 chunk groups, so the loads of one
// chain's chunks are spread over 8 threads instead of serialised in one.
#define SCAN_GROUPS 8
__global__ void __launch_bounds__(256)
k_fused_scan(const int32_t *__restrict__ agg, const int32_t *__restrict__ amax,
             const float *__restrict__ slack, const float *__restrict__ chunk_box, int C3, int n_chunks,
             const int32_t *__restrict__ init_images, int32_t *__restrict__ sstart, int32_t *status)
{
    __shared__ int tot[SCAN_GROUPS][32];
    const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const int cpg = (n_chunks + SCAN_GROUPS - 1) / SCAN_GROUPS;
    const int k0 = g * cpg, k1 = min(n_chunks, k0 + cpg);
    int sum = 0;
    if (c < C3)
        for (int kc = k0; kc < k1; kc++) sum += __ldg(agg + (int64_t)kc * C3 + c);
    tot[g][cl] = sum;
    __syncthreads();
    if (c >= C3) return;
    int S = init_images ? init_images[c] : 0, flags = 0;
    for (int gg = 0; gg < g; gg++) S += tot[gg][cl];
    for (int kb = k0; kb < k1; kb += 4) {
        int a[4], m[4];
        float sl[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {              // independent loads first
            const int64_t o = (int64_t)min(kb + j, k1 - 1) * C3 + c;
            a[j] = __ldg(agg + o); m[j] = __ldg(amax + o); sl[j] = __ldg(slack + o);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int kc = kb + j;
            if (kc < k1) {
                sstart[(int64_t)kc * C3 + c] = S;
                const int smax = abs(S) + m[j] + 1;
                const float Lmax = chunk_box[kc * 2], dL = chunk_box[kc * 2 + 1];
                // |true shifted difference - delta-form one| <= ulp32(u)/2 + ulp32(d)/2 + ulp32(d0)/2
                //                                              + |s| * |L_t - L_{t-1}|
                const float eps = (float)(smax + 2) * Lmax * 2.3841858e-7f + (float)smax * dL;
                if (!(sl[j] > eps)) flags |= PBK_ST_UNWRAP_AMBIGUOUS;
                if (smax > 126) flags |= PBK_ST_FUSED_RANGE;
                S += a[j];
            }
        }
    }
    if (flags) atomicOr(status, flags);
}

// net image-count change over the whole batch (sum of the sub-chunk totals): what a frame
// shard hands to the ranks after it
__global__ void k_fused_net(const int32_t *__restrict__ agg, int C3, int n_chunks, int32_t *__restrict__ net)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C3) return;
    int s = 0;
    for (int kc = 0; kc < n_chunks; kc++) s += __ldg(agg + (int64_t)kc * C3 + c);
    net[c] = s;
}

// ------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers (PTX ISA: mbarrier, cp.async.bulk)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// bounded wait: returns false if the phase never completed (reported, never spins forever)
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    for (int it = 0; it < (1 << 22); it++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

struct FusedParams {
    const float *x; const float *box; const double *mass; const int32_t *seg; const int32_t *gather;
    const double *lipid_mass;
    const int32_t *leaf_start, *leaf_len, *node_left, *node_right, *level_off;
    int L, K, H;
    double total_mass;
    const float *u_state; float *u_out; int first_frame;
    int64_t F; int M, N, norm, chunk, n_chunks, do_labels, sub;
    const int32_t *sstart;
    const int32_t *init_images;
    double *com, *com_unwrap, *mem_com; uint8_t *labels; int32_t *status;
    int use_tma, mass_in_smem;
    int skip;                   // debug ablation mask (PBK_FUSED_SKIP): 1 U, 2 L, 4 C role0, 8 C role1, 16 Z, 32 stores
    long long *prof;            // optional [n_chunks][8] phase cycle counters (debug), may be NULL
};

template <bool MASS_SMEM>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
k_fused_frames(FusedParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C3 = 3 * P.M;
    const int C3pad = (C3 + 3) & ~3;
    float *XA = reinterpret_cast<float *>(smem_raw);
    float *XB = XA + C3pad;
    double *s_leaf = reinterpret_cast<double *>(XB + C3pad);           // (L+K)*3
    double *s_z = s_leaf + (size_t)(P.L + P.K) * 3;                     // N   com_unwrap[norm]
    double *s_mass = s_z + P.N;                                         // M (optional)
    signed char *S = reinterpret_cast<signed char *>(s_mass + (MASS_SMEM ? P.M : 0));  // C3
    __shared__ int s_tab[FUSED_TAB];            // leaf_start | leaf_len | node_left | node_right | level_off
    __shared__ uint64_t bar[2];
    __shared__ double s_mem[3];
    __shared__ double red[32];
    __shared__ double s_pmax[32];
    __shared__ float s_box[2][3];
    __shared__ double s_zexact;
    __shared__ int s_flag;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int kc = blockIdx.x;
    const int64_t f0 = (int64_t)kc * P.chunk, f1 = min(P.F, f0 + P.chunk);
    if (f0 >= f1) return;
    // address space known at compile time: LDS for the shared copy, LDG for the global one
    const double *__restrict__ gmass = P.mass;
#define MASS(a) (MASS_SMEM ? s_mass[a] : __ldg(gmass + (a)))
    const uint32_t frame_bytes = (uint32_t)C3 * 4u;
    int bad = 0;

    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); s_flag = 0; }
    int *t_ls = s_tab, *t_ll = t_ls + P.L, *t_nl = t_ll + P.L, *t_nr = t_nl + P.K, *t_lo = t_nr + P.K;
    for (int i = tid; i < P.L; i += blockDim.x) { t_ls[i] = P.leaf_start[i]; t_ll[i] = P.leaf_len[i]; }
    for (int i = tid; i < P.K; i += blockDim.x) { t_nl[i] = P.node_left[i]; t_nr[i] = P.node_right[i]; }
    for (int i = tid; i <= P.H; i += blockDim.x) t_lo[i] = P.level_off[i];
    if (MASS_SMEM) for (int a = tid; a < P.M; a += blockDim.x) s_mass[a] = P.mass[a];
    // image counts at the chunk start; "previous" frame buffer
    //   chunk 0, first batch : prev = x_0 itself, counts 0  (u_0 = x_0)
    //   chunk 0, later batch : prev = carried u_state, counts 0 (u_prev = prev + 0*L)
    //   chunk k > 0          : prev = raw x_{f0-1}, counts from the scan
    float *cur = XA, *prev = XB;
    const float *prev_src = (kc == 0) ? (P.first_frame ? P.x : P.u_state) : P.x + (f0 - 1) * (int64_t)C3;
    for (int c = tid; c < C3; c += blockDim.x) {
        prev[c] = prev_src[c];
        int s0 = (kc == 0) ? (P.init_images ? P.init_images[c] : 0) : P.sstart[(int64_t)kc * P.sub * C3 + c];
        if (s0 > 127 || s0 < -127) { bad |= 4; s0 = 0; }
        S[c] = (signed char)s0;
    }
    if (tid < 3) s_box[1][tid] = (kc == 0) ? P.box[tid] : P.box[(f0 - 1) * 3 + tid];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // prefetch the first frame of the chunk
    uint32_t parbits = 0u;                      // bit b = parity to wait for on bar[b]
    int cb = 0;                                 // barrier / buffer index of `cur`
    if (P.use_tma) {
        if (tid == 0) { mbar_expect_tx(&bar[0], frame_bytes); bulk_g2s(cur, P.x + f0 * (int64_t)C3, frame_bytes, &bar[0]); }
    } else {
        for (int c = tid; c < C3; c += blockDim.x) cur[c] = P.x[f0 * (int64_t)C3 + c];
    }

    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tk = clock64();
#define PROF_MARK(i) do { if (P.prof && tid == 0) { long long n_ = clock64(); pc[i] += n_ - tk; tk = n_; } } while (0)
    for (int64_t t = f0; t < f1; t++) {
        if (tid < 3) s_box[0][tid] = P.box[t * 3 + tid];
        if (P.use_tma) {
            if (!mbar_wait(&bar[cb], (parbits >> cb) & 1u)) bad = 2;
            parbits ^= (1u << cb);
        }
        __syncthreads();
        PROF_MARK(0);                           // wait for the frame
        const float Lc0 = s_box[0][0], Lc1 = s_box[0][1], Lc2 = s_box[0][2];
        const float Lp0 = s_box[1][0], Lp1 = s_box[1][1], Lp2 = s_box[1][2];
        // ---- phase U: image counts of frame t.  Same-image test on the raw frames (two adds and
        // a compare per chain); only chains that fail it run the exact reference rule -----------
        if (!(P.skip & 1)) {
            const float eps0 = pbk_image_eps(Lc0, Lp0), eps1 = pbk_image_eps(Lc1, Lp1), eps2 = pbk_image_eps(Lc2, Lp2);
            const float h0 = Lc0 * 0.5f - eps0, h1 = Lc1 * 0.5f - eps1, h2 = Lc2 * 0.5f - eps2;
            const bool boxok = Lc0 > 0.0f && Lc1 > 0.0f && Lc2 > 0.0f;
            for (int a = tid; a < P.M; a += blockDim.x) {
                const int c = a * 3;
                const float x0 = cur[c], x1 = cur[c + 1], x2 = cur[c + 2];
                const float p0 = prev[c], p1 = prev[c + 1], p2 = prev[c + 2];
                const bool same = (h0 - fabsf(__fsub_rn(x0, p0)) > 0.0f) & (h1 - fabsf(__fsub_rn(x1, p1)) > 0.0f) &
                                  (h2 - fabsf(__fsub_rn(x2, p2)) > 0.0f) & boxok;
                if (!same) {                                    // rare: exact rule for the atom
                    const float xs[3] = {x0, x1, x2}, ps[3] = {p0, p1, p2};
                    const float Lcs[3] = {Lc0, Lc1, Lc2}, Lps[3] = {Lp0, Lp1, Lp2};
                    for (int k = 0; k < 3; k++) {
                        float sl;
                        int sn = fused_shift_far(__fsub_rn(xs[k], pbk_unwrapped(ps[k], (int)S[c + k], Lps[k])), Lcs[k], &sl, &bad);
                        if (sn > 127 || sn < -127) { bad |= 4; sn = 0; }
                        S[c + k] = (signed char)sn;
                    }
                }
            }
        }
        __syncthreads();
        PROF_MARK(1);                           // phase U
        // ---- phase L: NumPy-pairwise leaf sums of m*f64(u); one warp per leaf, lanes 0..23 =
        // 8 accumulator lanes x 3 axes ---------------------------------------------------------------
        if (!(P.skip & 2)) {
            const int j = lane / 3, k = lane % 3;
            const bool act = lane < 24;
            const float Lc = k == 0 ? Lc0 : (k == 1 ? Lc1 : Lc2);
            const int src1 = ((j ^ 1) * 3 + k) & 31, src2 = ((j ^ 2) * 3 + k) & 31, src4 = ((j ^ 4) * 3 + k) & 31;
            // two leaves per pass: two independent accumulation chains per lane
            for (int leafA = warp; leafA < P.L; leafA += 2 * nwarp) {
                const int leafB = leafA + nwarp;
                const bool hasB = leafB < P.L;
                const int startA = t_ls[leafA], nA = t_ll[leafA];
                const int startB = hasB ? t_ls[leafB] : 0, nB = hasB ? t_ll[leafB] : 0;
                const float *xA = cur + (size_t)startA * 3 + k, *xB = cur + (size_t)startB * 3 + k;
                const signed char *sA = S + (size_t)startA * 3 + k, *sB = S + (size_t)startB * 3 + k;
                double rA = 0.0, rB = 0.0;
                const int fullA = nA - (nA % 8), fullB = nB - (nB % 8);
                if (act) {
                    if (nA >= 8) rA = __dmul_rn(MASS(startA + j), (double)pbk_unwrapped_i8(xA[j * 3], (int)sA[j * 3], Lc));
                    if (nB >= 8) rB = __dmul_rn(MASS(startB + j), (double)pbk_unwrapped_i8(xB[j * 3], (int)sB[j * 3], Lc));
                    const int fa = nA >= 8 ? fullA : 0, fb = nB >= 8 ? fullB : 0;
                    const int common = min(fa, fb);
                    int i = 8 + j;
#pragma unroll 2
                    for (; i < common; i += 8) {
                        const double pA = __dmul_rn(MASS(startA + i), (double)pbk_unwrapped_i8(xA[i * 3], (int)sA[i * 3], Lc));
                        const double pB = __dmul_rn(MASS(startB + i), (double)pbk_unwrapped_i8(xB[i * 3], (int)sB[i * 3], Lc));
                        rA = __dadd_rn(rA, pA);
                        rB = __dadd_rn(rB, pB);
                    }
                    for (int ia = i; ia < fa; ia += 8)
                        rA = __dadd_rn(rA, __dmul_rn(MASS(startA + ia), (double)pbk_unwrapped_i8(xA[ia * 3], (int)sA[ia * 3], Lc)));
                    for (int ib = i; ib < fb; ib += 8)
                        rB = __dadd_rn(rB, __dmul_rn(MASS(startB + ib), (double)pbk_unwrapped_i8(xB[ib * 3], (int)sB[ib * 3], Lc)));
                }
                rA = __dadd_rn(rA, __shfl_sync(0xffffffffu, rA, act ? src1 : lane));
                rB = __dadd_rn(rB, __shfl_sync(0xffffffffu, rB, act ? src1 : lane));
                rA = __dadd_rn(rA, __shfl_sync(0xffffffffu, rA, act ? src2 : lane));
                rB = __dadd_rn(rB, __shfl_sync(0xffffffffu, rB, act ? src2 : lane));
                rA = __dadd_rn(rA, __shfl_sync(0xffffffffu, rA, act ? src4 : lane));
                rB = __dadd_rn(rB, __shfl_sync(0xffffffffu, rB, act ? src4 : lane));
                if (lane < 3) {
                    int i0 = fullA;
                    if (nA < 8) { rA = 0.0; i0 = 0; }
                    for (int i = i0; i < nA; i++)
                        rA = __dadd_rn(rA, __dmul_rn(MASS(startA + i), (double)pbk_unwrapped_i8(xA[i * 3], (int)sA[i * 3], Lc)));
                    s_leaf[leafA * 3 + k] = rA;
                    if (hasB) {
                        i0 = fullB;
                        if (nB < 8) { rB = 0.0; i0 = 0; }
                        for (int i = i0; i < nB; i++)
                            rB = __dadd_rn(rB, __dmul_rn(MASS(startB + i), (double)pbk_unwrapped_i8(xB[i * 3], (int)sB[i * 3], Lc)));
                        s_leaf[leafB * 3 + k] = rB;
                    }
                }
            }
        }
        PROF_MARK(7);                           // (own leaves done)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of `prev` before the async write
        __syncthreads();
        // the other buffer is free now: fetch frame t+1 into it while frame t is reduced
        if (t + 1 < f1 && P.use_tma && tid == 0) {
            mbar_expect_tx(&bar[cb ^ 1], frame_bytes);
            bulk_g2s(prev, P.x + (t + 1) * (int64_t)C3, frame_bytes, &bar[cb ^ 1]);
        }
        PROF_MARK(2);                           // phase L
        if (warp == 0) {                        // summation tree + division in one warp
            for (int h = 0; h < P.H; h++) {
                const int a = t_lo[h], b = t_lo[h + 1];
                for (int w = lane; w < (b - a) * 3; w += 32) {
                    const int node = a + w / 3, k = w % 3;
                    s_leaf[(P.L + node) * 3 + k] = __dadd_rn(s_leaf[t_nl[node] * 3 + k], s_leaf[t_nr[node] * 3 + k]);
                }
                __syncwarp();
            }
            if (lane < 3) {
                const double mc = __ddiv_rn(s_leaf[(P.L + P.K - 1) * 3 + lane], P.total_mass);
                s_mem[lane] = mc;
                P.mem_com[t * 3 + lane] = mc;
            }
        }
        __syncthreads();
        PROF_MARK(3);                           // tree
        // ---- phase C: per-lipid COMs; thread (i, role): role 0 wrapped, role 1 unwrapped -----
        {
            const double m0 = s_mem[0], m1 = s_mem[1], m2 = s_mem[2];
            const int32_t *__restrict__ gat = P.gather;
            for (int w0 = 0; w0 < 2 * P.N; w0 += blockDim.x) {
                const int w = w0 + tid;
                double zv = 0.0;
                bool zact = false;
                if (w < 2 * P.N) {
                    const int role = w >= P.N, i = role ? w - P.N : w;
                    const int a0 = P.seg[i], a1 = P.seg[i + 1];
                    double c0 = 0, c1 = 0, c2 = 0;
                    if (!role && !(P.skip & 4)) {
#pragma unroll 4
                        for (int p = a0; p < a1; p++) {
                            const int a = gat ? gat[p] : p;
                            const double m = MASS(a);
                            c0 = __dadd_rn(c0, __dmul_rn((double)cur[a * 3], m));
                            c1 = __dadd_rn(c1, __dmul_rn((double)cur[a * 3 + 1], m));
                            c2 = __dadd_rn(c2, __dmul_rn((double)cur[a * 3 + 2], m));
                        }
                    } else if (role && !(P.skip & 8)) {
#pragma unroll 4
                        for (int p = a0; p < a1; p++) {
                            const int a = gat ? gat[p] : p;
                            const double m = MASS(a);
                            const float v0 = __double2float_rn(__dsub_rn((double)pbk_unwrapped_i8(cur[a * 3], S[a * 3], Lc0), m0));
                            const float v1 = __double2float_rn(__dsub_rn((double)pbk_unwrapped_i8(cur[a * 3 + 1], S[a * 3 + 1], Lc1), m1));
                            const float v2 = __double2float_rn(__dsub_rn((double)pbk_unwrapped_i8(cur[a * 3 + 2], S[a * 3 + 2], Lc2), m2));
                            c0 = __dadd_rn(c0, __dmul_rn((double)v0, m));
                            c1 = __dadd_rn(c1, __dmul_rn((double)v1, m));
                            c2 = __dadd_rn(c2, __dmul_rn((double)v2, m));
                        }
                    }
                    const double lm = P.lipid_mass[i];
                    c0 = __ddiv_rn(c0, lm); c1 = __ddiv_rn(c1, lm); c2 = __ddiv_rn(c2, lm);
                    double *o = (role ? P.com_unwrap : P.com) + (t * (int64_t)P.N + i) * 3;
                    if (!(P.skip & 32)) { o[0] = c0; o[1] = c1; o[2] = c2; }
                    if (role) { zv = P.norm == 0 ? c0 : (P.norm == 1 ? c1 : c2); s_z[i] = zv; zact = true; }
                }
            }
        }
        __syncthreads();
        if (P.do_labels) {                      // warp partials of z over lipids 32w .. 32w+31
            const int np = (P.N + 31) / 32;
            double sacc = 0.0, macc = 0.0;
            for (int w = warp; w < np; w += nwarp) {
                const int i = w * 32 + lane;
                const double v = i < P.N ? s_z[i] : 0.0;
                double sm = pbk_warp_sum(v), mx = fabs(v);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                sacc += sm;
                macc = fmax(macc, mx);
            }
            if (lane == 0 && warp < 32) { red[warp] = sacc; s_pmax[warp] = macc; }
            __syncthreads();
        }
        PROF_MARK(4);                           // phase C
        // ---- phase Z: leaflet labels (tree mean decides unless a lipid is within the rounding
        // bound of it; then thread 0 replays the sequential Welford mean) -------------------------
        if (P.do_labels && !(P.skip & 16)) {
            // warp partials (sum, max |z|) were written by the role-1 warps in phase C; every warp
            // combines them in the same fixed order, so no second barrier is needed
            double s2 = lane < nwarp ? red[lane] : 0.0, m2 = lane < nwarp ? s_pmax[lane] : 0.0;
            s2 = pbk_warp_sum(s2);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, o));
            double mean = s2 / (double)P.N;
            const double tol = 64.0 * (double)P.N * 2.220446049250313e-16 * m2;
            int close = 0;
            for (int i = tid; i < P.N; i += blockDim.x) if (fabs(s_z[i] - mean) <= tol) close = 1;
            if (__syncthreads_or(close)) {
                if (tid == 0) {
                    double m = 0.0;
                    for (int i = 0; i < P.N; i++) {
                        const double v = s_z[i];
                        m = (i == 0) ? v : __dadd_rn(m, __ddiv_rn(__dsub_rn(v, m), (double)(i + 1)));
                    }
                    s_zexact = m;
                }
                __syncthreads();
                mean = s_zexact;
            }
            uint8_t *out = P.labels + t * (int64_t)P.N;
            for (int i = tid; i < P.N; i += blockDim.x) {
                const double v = s_z[i];
                if (v > mean) out[i] = 1;
                else if (v < mean) out[i] = 0;
                else { out[i] = 0; bad |= 8; }
            }
        }
        PROF_MARK(5);                           // phase Z
        // carried state for the next batch
        if (t == P.F - 1) {
            for (int c = tid; c < C3; c += blockDim.x) {
                const int k = c % 3;
                P.u_out[c] = pbk_unwrapped(cur[c], (int)S[c], k == 0 ? Lc0 : (k == 1 ? Lc1 : Lc2));
            }
        }
        if (tid < 3) s_box[1][tid] = s_box[0][tid];
        if (!P.use_tma && t + 1 < f1) {
            __syncthreads();
            for (int c = tid; c < C3; c += blockDim.x) prev[c] = P.x[(t + 1) * (int64_t)C3 + c];
        }
        // swap: the buffer being filled becomes `cur`, the frame just reduced becomes `prev`
        float *tmp = cur; cur = prev; prev = tmp;
        cb ^= 1;
        __syncthreads();
        PROF_MARK(6);                           // tail
    }
    if (P.prof && tid == 0) for (int i = 0; i < 8; i++) P.prof[kc * 8 + i] = pc[i];
    if (bad) {
        int st = 0;
        if (bad & 1) st |= PBK_ST_IMAGE_OVERFLOW;
        if (bad & 2) st |= PBK_ST_INTERNAL;
        if (bad & 4) st |= PBK_ST_FUSED_RANGE;
        if (bad & 8) st |= PBK_ST_LEAFLET_TIE;
        atomicOr(P.status, st);
    }
}

static int fused_v1_scratch_bytes(int64_t F, int64_t M, int sm_count, int64_t *bytes,
                                             int32_t *n_chunks_out, int32_t *chunk_out)
{
    if (F <= 0 || M <= 0 || sm_count <= 0 || !bytes) return PBK_EINVAL;
    int64_t chunk = (F + sm_count - 1) / sm_count;
    if (chunk < 4) chunk = 4;
    int64_t n_chunks = (F + chunk - 1) / chunk;
    *bytes = 4 * (n_chunks * 3 * M * 16 + n_chunks * 8) + 3 * M * 4 + 256;   // up to 4 sub-chunks per chunk
    if (n_chunks_out) *n_chunks_out = (int32_t)n_chunks;
    if (chunk_out) *chunk_out = (int32_t)chunk;
    return PBK_OK;
}

int pbk_reduce_fused_v1(pbk_ctx *ctx, const float *x, const float *box, const double *mass,
                                const int32_t *seg, const int32_t *gather, const double *lipid_mass,
                                const int32_t *leaf_start, const int32_t *leaf_len, int32_t L,
                                const int32_t *node_left, const int32_t *node_right, int32_t K,
                                const int32_t *level_off, int32_t H, double total_mass, float *u_state,
                                int first_frame, int64_t F, int64_t M, int64_t N, int norm,
                                int do_labels, int phase, const int32_t *init_images,
                                int32_t *net_images, void *scratch, int64_t scratch_bytes, double *com,
                                double *com_unwrap, double *mem_com, uint8_t *labels, int32_t *status,
                                void *stream)
{
    if ((phase & ~3) || phase == 0 || (init_images && !first_frame))
        return pbk_fail(ctx, PBK_EINVAL, "pbk_reduce_fused: bad phase / init_images needs first_frame");
    if (!ctx || !x || !box || !mass || !seg || !lipid_mass || !leaf_start || !leaf_len || !u_state ||
        !scratch || !com || !com_unwrap || !mem_com || !status || (do_labels && !labels) || F < 0 ||
        M <= 0 || N <= 0 || L <= 0 || K < 0 || norm < 0 || norm > 2)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_reduce_fused: bad argument");
    if (F == 0) return PBK_OK;
    if (M > (1 << 20) || 2 * (int64_t)L + 2 * (int64_t)K + H + 1 > FUSED_TAB) return PBK_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t need; int32_t n_chunks, chunk;
    fused_v1_scratch_bytes(F, M, ctx->sm_count, &need, &n_chunks, &chunk);
    if (scratch_bytes < need) return pbk_fail(ctx, PBK_EINVAL, "pbk_reduce_fused: scratch too small");
    const int C3 = (int)(3 * M), C3pad = (C3 + 3) & ~3;
    size_t base = (size_t)C3pad * 8 + (size_t)(L + K) * 24 + (size_t)N * 8 + (size_t)C3 + 64;
    int mass_in_smem = 1;
    size_t smem = base + (size_t)M * 8;
    const size_t limit = 227 * 1024 - 5 * 1024;  // static shared of the kernel is < 5 KB
    if (smem > limit) { mass_in_smem = 0; smem = base; }
    if (smem > limit) return PBK_EUNSUPPORTED;   // frame does not fit: caller uses pbk_com.cu path
    // the image-count pre-pass works on sub-chunks (more, shorter work items: the streaming
    // kernel is latency bound otherwise); the frames kernel picks every `sub`-th scan entry
    int sub = 1;
    for (int c = 4; c >= 2; c--) if (chunk % c == 0) { sub = c; break; }
    const int chunk_sub = chunk / sub;
    const int n_sub = (int)((F + chunk_sub - 1) / chunk_sub);
    // scratch layout
    int32_t *agg = reinterpret_cast<int32_t *>(scratch);
    int32_t *amax = agg + (int64_t)n_sub * C3;
    float *slack = reinterpret_cast<float *>(amax + (int64_t)n_sub * C3);
    int32_t *sstart = reinterpret_cast<int32_t *>(slack + (int64_t)n_sub * C3);
    float *chunk_box = reinterpret_cast<float *>(sstart + (int64_t)n_sub * C3);
    float *u_out = chunk_box + 2 * (int64_t)n_sub;
    if (phase & 1) {
    if (C3 % 12 == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0)
        k_fused_agg12<<<dim3(n_sub, AGG_SPLIT), AGG_THREADS, 0, st>>>(x, box, u_state, first_frame, F, C3, chunk_sub,
                                                                     agg, amax, slack, chunk_box, status);
    else if (C3 % 4 == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0)
        k_fused_agg<4><<<dim3(n_sub, AGG_SPLIT), AGG_THREADS, 0, st>>>(x, box, u_state, first_frame, F, C3, chunk_sub, agg,
                                                           amax, slack, chunk_box, status);
    else
        k_fused_agg<1><<<dim3(n_sub, AGG_SPLIT), AGG_THREADS, 0, st>>>(x, box, u_state, first_frame, F, C3, chunk_sub, agg,
                                                           amax, slack, chunk_box, status);
    PBK_CHECK_LAUNCH(ctx, "k_fused_agg");
    if (net_images) {
        k_fused_net<<<(C3 + 255) / 256, 256, 0, st>>>(agg, C3, n_sub, net_images);
        PBK_CHECK_LAUNCH(ctx, "k_fused_net");
    }
    }
    if (!(phase & 2)) return PBK_OK;
    k_fused_scan<<<(C3 + 31) / 32, 256, 0, st>>>(agg, amax, slack, chunk_box, C3, n_sub, init_images, sstart, status);
    PBK_CHECK_LAUNCH(ctx, "k_fused_scan");
    FusedParams P;
    P.x = x; P.box = box; P.mass = mass; P.seg = seg; P.gather = gather; P.lipid_mass = lipid_mass;
    P.leaf_start = leaf_start; P.leaf_len = leaf_len; P.node_left = node_left; P.node_right = node_right;
    P.level_off = level_off; P.L = L; P.K = K; P.H = H; P.total_mass = total_mass;
    P.u_state = u_state; P.u_out = u_out; P.first_frame = first_frame; P.F = F; P.M = (int)M; P.N = (int)N; P.norm = norm;
    P.sub = sub; P.chunk = chunk; P.n_chunks = n_chunks; P.init_images = init_images; P.do_labels = do_labels; P.sstart = sstart;
    P.com = com; P.com_unwrap = com_unwrap; P.mem_com = mem_com; P.labels = labels; P.status = status;
    P.use_tma = ((C3 * 4) % 16 == 0) && ((reinterpret_cast<uintptr_t>(x) % 16) == 0);
    P.mass_in_smem = mass_in_smem;
    P.skip = getenv("PBK_FUSED_SKIP") ? atoi(getenv("PBK_FUSED_SKIP")) : 0;
    P.prof = (scratch_bytes >= need + 8 * 8 * (int64_t)n_chunks + 64 && getenv("PBK_FUSED_PROF")) ? reinterpret_cast<long long *>(reinterpret_cast<char *>(scratch) + ((need + 63) / 64) * 64) : nullptr;
    auto kern = mass_in_smem ? k_fused_frames<true> : k_fused_frames<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "k_fused_frames smem attribute", e);
    kern<<<n_chunks, FUSED_THREADS, smem, st>>>(P);
    PBK_CHECK_LAUNCH(ctx, "k_fused_frames");
    // u_state is read by chunk 0 and replaced by the last chunk: the new state is produced in
    // scratch and copied over once the kernel is done
    e = cudaMemcpyAsync(u_state, u_out, (size_t)C3 * 4, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "u_state copy", e);
    return PBK_OK;
}
