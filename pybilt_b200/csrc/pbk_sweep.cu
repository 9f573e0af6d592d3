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
#include <stdlib.h>

#include "pbk_common.cuh"

// pbk_com.cu: combines leaf sums [F][L+K][3] along the recursion tree -> mem_com[F][3]
int pbk_launch_mem_tree(pbk_ctx *ctx, const int32_t *node_left, const int32_t *node_right, const int32_t *level_off,
                        int32_t H, int32_t L, int32_t K, double total_mass, int64_t F, double *node_vals,
                        double *mem_com, cudaStream_t st);

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
static __device__ __forceinline__ uint32_t sw_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// (all helpers take 32-bit shared-window addresses formed ONCE per kernel: converting a generic
//  pointer inside a loop costs an S2R SR_CgaCtaId round trip every time)
static __device__ __forceinline__ void sw_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
static __device__ __forceinline__ void sw_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void sw_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
static __device__ __forceinline__ void sw_cp4(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
static __device__ __forceinline__ void sw_cp_arrive(uint32_t bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
static __device__ __forceinline__ bool sw_mbar_wait(uint32_t a, uint32_t parity)
{
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
// k_sweep_a: persistent CTAs over tasks (leaf, block of SW_CHUNK frames) in task order
// task = block * L + leaf.  Task (leaf, b) starts from the image counts task (leaf, b-1) left in
// `snap` (a flag per leaf says how many blocks are done); since every CTA takes its tasks in
// increasing order and all CTAs are resident, the lowest unfinished task can always run.
// ------------------------------------------------------------------------------------------
#define SW_CHUNK 32                  // frames per block of sweep A = frames per time chunk of sweep B
#define SA_THREADS 160                  // warps 0-2 columns, warp 3 reducer, warp 4 producer (TMA issue, box rows)
#define SA_SUB 4                     // frames per barrier interval = frames per group of the shared ring
#define SA_GROUPS 3                  // groups in the ring: SA_GROUPS - 1 sub-blocks of positions in flight per CTA
#define SA_LEAF_FLOATS 384           // 128 atoms x 3

struct SweepAParams {
    const float *x; const float *boxtab; const double *mass;
    const int32_t *leaf_start, *leaf_len;
    int L, nodes_per_frame, nblk;
    const float *prev0;             // u_state (later batches) or x (first frame: frame 0 is its own reference)
    const int32_t *init_images;
    int F; int64_t M; int net_frame;
    int32_t *net_images; float *u_out; double *node_vals; int32_t *status;
    int16_t *snap;                  // [nblk][3M] image counts at frame b*SW_CHUNK - 1 (row 0 unused)
    int *flags;                     // [L] blocks finished per leaf
    int store_sums;
};

// exact reference rule for one component (rare): new image count as a float
static __device__ __noinline__ float sa_exact(float xv, float pv, float sf, float Lc, float Lp, int *bad)
{
    const int s = (int)sf;
    return (float)sw_shift(__fsub_rn(xv, pbk_unwrapped(pv, s, Lp)), Lc, bad);
}

// phase cycle counters (tools/prof_sweep.py), compiled in with -DPBK_SWEEP_PROFILE only
#ifdef PBK_SWEEP_PROFILE
__device__ unsigned long long g_sa_prof[16];
// (a barrier only blocks at the next use of protected state: the clock read is made to depend on a
//  shared-memory load issued after it)
#define SA_MARK(i) do { long long n_; const int dep_ = *(volatile int *)&s_full[0]; \
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(n_) : "r"(dep_)); pc[i] += n_ - tk; tk = n_; } while (0)
extern "C" int pbk_debug_sweep_prof(unsigned long long *out)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_sa_prof, sizeof(unsigned long long) * 16);
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_sa_prof, z, sizeof(z));
    return 0;
}
#else
#define SA_MARK(i) do { } while (0)
#endif

// one frame of the reducing phase (generic: any number of pending frames)
static __device__ __forceinline__ void sa_reduce_one(const double *pp, int lane, int nq, int n, int s1, int s2, int s4,
                                                     double *out)
{
    double r = 0.0;
    if (lane < 24 && nq > 0) {
        r = pp[lane];
        for (int q = 1; q < nq; q++) r = __dadd_rn(r, pp[q * 24 + lane]);
    }
    const bool act = lane < 24;
    double o = __shfl_sync(0xffffffffu, r, act ? s1 : lane);
    r = __dadd_rn(r, o);
    o = __shfl_sync(0xffffffffu, r, act ? s2 : lane);
    r = __dadd_rn(r, o);
    o = __shfl_sync(0xffffffffu, r, act ? s4 : lane);
    r = __dadd_rn(r, o);
    if (lane < 3) {
        int i0 = 8 * nq;
        if (nq == 0) { r = 0.0; i0 = 0; }
        for (int i = i0; i < n; i++) r = __dadd_rn(r, pp[3 * i + lane]);
        out[lane] = r;
    }
}

template <bool BULK>
__global__ void __launch_bounds__(SA_THREADS, 4)
k_sweep_a(SweepAParams P)
{
#ifdef PBK_SWEEP_PROFILE
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tk = clock64();
#endif
    __shared__ __align__(128) float s_x[SA_GROUPS][SA_SUB][SA_LEAF_FLOATS];
    __shared__ __align__(16) double s_p[2][SA_SUB][SA_LEAF_FLOATS];
    __shared__ uint64_t s_full[SA_GROUPS];       // one mbarrier per group
    __shared__ float s_boxes[2][SW_CHUNK][8];    // box rows of the task's frames (no global load inside the frame loop);
                                                 // the producer warp stages the next task's rows during the current one
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const int n_tasks = P.L * P.nblk;
    const int64_t C3 = 3 * P.M;
    int bad = 0;

    const uint32_t a_full = sw_smem_u32(&s_full[0]), a_x = sw_smem_u32(&s_x[0][0][0]);
    if (tid == 0)
        for (int g = 0; g < SA_GROUPS; g++) sw_mbar_init(a_full + 8u * g, BULK ? 1u : 32u);
    int prev_leaf = -1, prev_blk = -1;           // the task before (column state stays in registers when the leaf continues)
    sw_fence_async();
    __syncthreads();

    // ---- prefetch cursor (warp 4): one group (= one sub-block of frames) per call, two groups ahead
    // of the arithmetic, across tasks.  (Leaf geometry is reloaded only when the cursor moves to
    // another task: no global load per sub-block.)
    int pf_task = blockIdx.x, pf_left = 0, pf_grp = 0, pf_nfl = 0;
    const float *pf_src = nullptr;
    auto prefetch = [&]() {
        if (pf_left == 0) {
            if (pf_task >= n_tasks) return;
            const int leaf = pf_task % P.L, blk = pf_task / P.L;
            const int f0 = blk * SW_CHUNK;
            pf_left = min(P.F, f0 + SW_CHUNK) - f0;
            pf_nfl = 3 * P.leaf_len[leaf];
            pf_src = P.x + (int64_t)f0 * C3 + 3 * (int64_t)P.leaf_start[leaf];
            pf_task += G;
        }
        const int nf = min(SA_SUB, pf_left);
        const uint32_t bar = a_full + 8u * (pf_grp % SA_GROUPS), dst = a_x + (uint32_t)(pf_grp % SA_GROUPS) * (SA_SUB * SA_LEAF_FLOATS * 4u);
        if (BULK) {
            if (lane == 0) {
                sw_mbar_expect_tx(bar, (uint32_t)(nf * pf_nfl) * 4u);
                for (int u = 0; u < nf; u++)
                    sw_bulk_g2s(dst + (uint32_t)u * (SA_LEAF_FLOATS * 4u), pf_src + (int64_t)u * C3, (uint32_t)pf_nfl * 4u, bar);
            }
        } else {
            for (int u = 0; u < nf; u++)
                for (int c = lane; c < pf_nfl; c += 32) sw_cp4(dst + (uint32_t)u * (SA_LEAF_FLOATS * 4u) + 4u * c, pf_src + (int64_t)u * C3 + c);
            sw_cp_arrive(bar);
        }
        pf_grp++;
        pf_left -= nf;
        pf_src += (int64_t)nf * C3;
    };
    if (warp == 4)
        for (int g = 0; g < SA_GROUPS; g++) prefetch();

    // roles.  Warps 0-2 (columns): warp = axis k, lane = atoms lane + 32 j.  Warp 3 (reducer), lanes 0-23:
    // (accumulator rj, axis rk) of the frames of the sub-block before, their chains interleaved.
    const int k = warp < 3 ? warp : 0;             // a column warp = one axis; its lane owns atoms lane, lane+32, lane+64, lane+96
    const int rj = lane / 3, rk = lane % 3;        // (stride-3 shared accesses: conflict-free for both floats and doubles)
    const int s1 = ((rj ^ 1) * 3 + rk) & 31, s2 = ((rj ^ 2) * 3 + rk) & 31, s4 = ((rj ^ 4) * 3 + rk) & 31;
    int sub = 0;                                  // sub-blocks processed by this CTA (group = sub % SA_GROUPS)
    float pv[4] = {0.f, 0.f, 0.f, 0.f}, sf[4] = {0.f, 0.f, 0.f, 0.f}, Lp = 0.f;     // column state
    double mj[4] = {0.0, 0.0, 0.0, 0.0};
    int cj[4] = {0, 0, 0, 0};
    int wide = 0;

    int tpar = 0;
    for (int task = blockIdx.x; task < n_tasks; task += G, tpar ^= 1) {
        const int leaf = task % P.L, blk = task / P.L;
        const int f0 = blk * SW_CHUNK, f1 = min(P.F, f0 + SW_CHUNK);
        const int ls = P.leaf_start[leaf], n = P.leaf_len[leaf];
        const int nq = n >= 8 ? n / 8 : 0;
        float (*s_box)[8] = s_boxes[tpar];
        if (task == (int)blockIdx.x)
            for (int r = tid; r < (f1 - f0) * 8; r += SA_THREADS) s_box[r >> 3][r & 7] = P.boxtab[f0 * 8 + r];
        if (blk > 0 && tid == 0 && !(leaf == prev_leaf && blk == prev_blk + 1)) {
            int it = 0;
            while (*(volatile int *)(P.flags + leaf) < blk && ++it < (1 << 26)) __nanosleep(32);
            if (it >= (1 << 26)) bad |= 2;
            __threadfence();
        }
        __syncthreads();
        if (warp == 4 && task + G < n_tasks) {       // box rows of the next task
            const int nb = (task + G) / P.L;
            const int g0 = nb * SW_CHUNK, g1 = min(P.F, g0 + SW_CHUNK);
            for (int r = lane; r < (g1 - g0) * 8; r += 32) s_boxes[tpar ^ 1][r >> 3][r & 7] = P.boxtab[g0 * 8 + r];
        }
        SA_MARK(0);      // task start: box rows, flag wait
        // ---- column state ---------------------------------------------------------------------
        const int nval = warp < 3 ? max(0, min(4, (n - lane + 31) >> 5)) : 0;
        const bool cont = leaf == prev_leaf && blk == prev_blk + 1;
        prev_leaf = leaf; prev_blk = blk;
        if (!cont) { wide = 0; Lp = P.boxtab[(blk == 0 ? 0 : f0 - 1) * 8 + k]; }
        if (nval > 0 && !cont) {
            const float *prow = blk == 0 ? P.prev0 : P.x + (int64_t)(f0 - 1) * C3;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int a = min(lane + 32 * j, n - 1);        // atoms past the leaf's end repeat its last one with mass 0
                cj[j] = 3 * a + k;
                const int64_t gc = 3 * (int64_t)ls + cj[j];
                mj[j] = (j < nval) ? P.mass[ls + a] : 0.0;
                pv[j] = prow[gc];
                int s0 = blk == 0 ? (P.init_images ? P.init_images[gc] : 0) : (int)P.snap[(int64_t)blk * C3 + gc];
                if (s0 > 32767 || s0 < -32767) { bad |= 1; s0 = 0; }
                if (s0 > 127 || s0 < -127) wide = 1;
                sf[j] = (float)s0;
            }
        }
        const bool has_net = P.net_images && P.net_frame >= f0 && P.net_frame < f1;
        bool task_slow = false;                   // a box edge >= 2^13 in the task: the float64 unwrap for every frame
        for (int r = 0; r < f1 - f0; r++) task_slow = task_slow || s_box[r][6] != 0.0f;

        int pend = 0, pend_t0 = 0;
        SA_MARK(1);      // state init
        for (int t0 = f0; t0 < f1 || pend > 0; t0 += SA_SUB) {
            const int nfu = max(0, min(SA_SUB, f1 - t0));
            if (warp == 3) {
                // ---- reducer: NumPy's pairwise leaf sums of the frames of the sub-block before
                //   r[j] = a[j]; r[j] += a[8q + j];  res = ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); tail sequentially
                //   (numpy/_core/src/umath/loops_utils.h.src); the chains of the frames are interleaved
                if (pend > 0 && P.store_sums) {
                    const int pb = (sub + 1) & 1;
                    double *out = P.node_vals + ((int64_t)pend_t0 * P.nodes_per_frame + leaf) * 3;
                    const int64_t ostr = (int64_t)P.nodes_per_frame * 3;
                    if (pend == SA_SUB && nq > 0 && (n & 7) == 0) {
                        double r[SA_SUB];
                        const int l24 = lane < 24 ? lane : 0;
#pragma unroll
                        for (int u = 0; u < SA_SUB; u++) r[u] = s_p[pb][u][l24];
#pragma unroll 4
                        for (int q = 1; q < nq; q++) {
#pragma unroll
                            for (int u = 0; u < SA_SUB; u++) r[u] = __dadd_rn(r[u], s_p[pb][u][q * 24 + l24]);
                        }
                        const bool act = lane < 24;
#pragma unroll
                        for (int u = 0; u < SA_SUB; u++) r[u] = __dadd_rn(r[u], __shfl_sync(0xffffffffu, r[u], act ? s1 : lane));
#pragma unroll
                        for (int u = 0; u < SA_SUB; u++) r[u] = __dadd_rn(r[u], __shfl_sync(0xffffffffu, r[u], act ? s2 : lane));
#pragma unroll
                        for (int u = 0; u < SA_SUB; u++) r[u] = __dadd_rn(r[u], __shfl_sync(0xffffffffu, r[u], act ? s4 : lane));
                        if (lane < 3) {
#pragma unroll
                            for (int u = 0; u < SA_SUB; u++) out[u * ostr + lane] = r[u];
                        }
                    } else {
                        for (int u = 0; u < pend; u++) sa_reduce_one(&s_p[pb][u][0], lane, nq, n, s1, s2, s4, out + u * ostr);
                    }
                }
                SA_MARK(2);
            } else if (warp < 3 && nfu > 0) {
                // ---- columns: image counts and products m*f64(u) of frames t0 .. t0+nfu-1 -------------
                const int grp2 = sub % SA_GROUPS, pb2 = sub & 1;
                if (!sw_mbar_wait(a_full + 8u * grp2, (uint32_t)(sub / SA_GROUPS) & 1u)) bad |= 2;
                SA_MARK(3);  // waiting for the frames
                if (nval > 0) {
                    // Fast path (almost every sub-block): the same-image tests only compare consecutive RAW
                    // frames, so all SA_SUB frames are tested at once; when every test passes the image
                    // counts do not change and the 4 x SA_SUB products are independent of each other.
                    bool fast = nfu == SA_SUB && !wide && !task_slow && !(has_net && P.net_frame >= t0 && P.net_frame < t0 + SA_SUB);
                    float xa[SA_SUB][4], Lcs[SA_SUB];
                    if (fast) {
                        float mn = 3.0e38f, tiny = 3.0e38f;
#pragma unroll
                        for (int u = 0; u < SA_SUB; u++) {
                            Lcs[u] = s_box[t0 - f0 + u][k];
                            const float hc = s_box[t0 - f0 + u][3 + k];
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                xa[u][j] = s_x[grp2][u][cj[j]];
                                mn = fminf(mn, hc - fabsf(__fsub_rn(xa[u][j], u == 0 ? pv[j] : xa[u - 1][j])));
                                tiny = fminf(tiny, fabsf(xa[u][j]));
                            }
                        }
                        fast = mn > 0.0f && tiny >= 0.00390625f;
                    }
                    if (fast) {
#pragma unroll
                        for (int u = 0; u < SA_SUB; u++) {
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const double pr = __dmul_rn(mj[j], (double)__fmaf_rn(sf[j], Lcs[u], xa[u][j]));
                                if (nval == 4 || j < nval) s_p[pb2][u][cj[j]] = pr;
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) pv[j] = xa[SA_SUB - 1][j];
                        Lp = Lcs[SA_SUB - 1];
                    } else {
#pragma unroll
                        for (int u = 0; u < SA_SUB; u++) {
                            if (u < nfu) {
                                const int t = t0 + u;
                                const float Lc = s_box[t - f0][k], hc = wide ? -1.0f : s_box[t - f0][3 + k];
                                const bool slow = s_box[t - f0][6] != 0.0f || wide;
                                float xv[4], d[4];
#pragma unroll
                                for (int j = 0; j < 4; j++) xv[j] = s_x[grp2][u][cj[j]];
#pragma unroll
                                for (int j = 0; j < 4; j++) d[j] = hc - fabsf(__fsub_rn(xv[j], pv[j]));
                                if (!(fminf(fminf(d[0], d[1]), fminf(d[2], d[3])) > 0.0f)) {
#pragma unroll
                                    for (int j = 0; j < 4; j++) {
                                        if (!(d[j] > 0.0f)) {
                                            sf[j] = sa_exact(xv[j], pv[j], sf[j], Lc, Lp, &bad);
                                            if (fabsf(sf[j]) > 127.0f) wide = 1;
                                        }
                                    }
                                }
                                float uu[4];
#pragma unroll
                                for (int j = 0; j < 4; j++) uu[j] = __fmaf_rn(sf[j], Lc, xv[j]);
                                if (!(fminf(fminf(fabsf(xv[0]), fabsf(xv[1])), fminf(fabsf(xv[2]), fabsf(xv[3]))) >= 0.00390625f) || slow) {
#pragma unroll
                                    for (int j = 0; j < 4; j++) uu[j] = sw_unwrap_slow(xv[j], (int)sf[j], Lc);
                                }
                                if (nval == 4) {
#pragma unroll
                                    for (int j = 0; j < 4; j++) s_p[pb2][u][cj[j]] = __dmul_rn(mj[j], (double)uu[j]);
                                } else {
#pragma unroll
                                    for (int j = 0; j < 4; j++) if (j < nval) s_p[pb2][u][cj[j]] = __dmul_rn(mj[j], (double)uu[j]);
                                }
#pragma unroll
                                for (int j = 0; j < 4; j++) pv[j] = xv[j];
                                if (has_net && t == P.net_frame) {
#pragma unroll
                                    for (int j = 0; j < 4; j++)
                                        if (j < nval) P.net_images[3 * (int64_t)ls + cj[j]] = (int)sf[j];
                                }
                                Lp = Lc;
                            }
                        }
                    }
                    if (t0 + nfu == f1) {           // state after the task's last frame
                        if (f1 < P.F) {
#pragma unroll
                            for (int j = 0; j < 4; j++)
                                if (j < nval) P.snap[(int64_t)(blk + 1) * C3 + 3 * (int64_t)ls + cj[j]] = (int16_t)(int)sf[j];
                        } else if (P.u_out) {
#pragma unroll
                            for (int j = 0; j < 4; j++)
                                if (j < nval) P.u_out[3 * (int64_t)ls + cj[j]] = pbk_unwrapped(pv[j], (int)sf[j], Lp);
                        }
                    }
                }
                SA_MARK(4);  // column arithmetic
                // (no proxy fence: the group is only READ here, and its refill by the async proxy is ordered
                //  behind those reads by the barrier below -- the consumer-release pattern of TMA pipelines)
            }
            __syncthreads();
            SA_MARK(6);  // barrier
            if (nfu > 0) {
                // the group of this sub-block is free again: the prefetch cursor moves one group on
                if (warp == 4) prefetch();
                sub++;
            }
            SA_MARK(7);
            pend = nfu;
            pend_t0 = t0;
        }
        // this leaf's block is done: its image counts (snap) are visible before the flag
        if (f1 < P.F && tid == 0) {
            __threadfence();
            *(volatile int *)(P.flags + leaf) = blk + 1;
        }
    }
#ifdef PBK_SWEEP_PROFILE
    if (tid == 0 || tid == 96)
        for (int i = 0; i < 8; i++) atomicAdd(&g_sa_prof[(tid == 0 ? 0 : 8) + i], (unsigned long long)pc[i]);
#endif
    if (bad) atomicOr(P.status, (bad & 1 ? PBK_ST_IMAGE_OVERFLOW : 0) | (bad & 2 ? PBK_ST_INTERNAL : 0));
}

// ------------------------------------------------------------------------------------------
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
#define LB_THREADS 256
#define LB_VPT 8
#define LB_SLICE (LB_THREADS * LB_VPT)

// per (slice of 2,048 lipids, frame): (sum, max |z|) of the compact normal components
__global__ void __launch_bounds__(LB_THREADS)
k_zpart(const double *__restrict__ zc, int64_t N, double *__restrict__ zpart)
{
    __shared__ double red[32];
    const int64_t f = blockIdx.y;
    const double *z = zc + f * N;
    const int64_t i0 = (int64_t)blockIdx.x * LB_SLICE + threadIdx.x;
    double s = 0.0, mx = 0.0;
#pragma unroll
    for (int e = 0; e < LB_VPT; e++) {
        const int64_t i = i0 + (int64_t)e * LB_THREADS;
        const double v = i < N ? z[i] : 0.0;
        s += v;
        mx = fmax(mx, fabs(v));
    }
    s = pbk_block_sum(s, red);
    mx = pbk_block_max(mx, red);
    if (threadIdx.x == 0) {
        zpart[(f * gridDim.x + blockIdx.x) * 2] = s;
        zpart[(f * gridDim.x + blockIdx.x) * 2 + 1] = mx;
    }
}

// zstat[f] = (tree mean of the normal components, rounding bound) from the slice partials
__global__ void __launch_bounds__(256)
k_zmean(const double *__restrict__ zpart, int n_part, int64_t N, double *__restrict__ zstat, int *__restrict__ flags)
{
    __shared__ double red[32];
    const int64_t f = blockIdx.x;
    double s = 0.0, mx = 0.0;
    for (int q = threadIdx.x; q < n_part; q += blockDim.x) {
        s += zpart[(f * n_part + q) * 2];
        mx = fmax(mx, zpart[(f * n_part + q) * 2 + 1]);
    }
    s = pbk_block_sum(s, red);
    mx = pbk_block_max(mx, red);
    if (threadIdx.x == 0) {
        zstat[f * 2] = s / (double)N;
        zstat[f * 2 + 1] = 64.0 * (double)N * 2.220446049250313e-16 * mx;
        flags[f] = 0;
    }
}

// grid (slices, frames): labels of 2,048 lipids against the tree mean of frame src; a lipid within
// the rounding bound of the mean flags the frame for k_leaflets_c (sequential Welford mean)
__global__ void __launch_bounds__(LB_THREADS)
k_labels(const double *__restrict__ zc, const double *__restrict__ zstat, int64_t N, int64_t frame0, int interval,
         const uint8_t *__restrict__ prev, uint8_t *__restrict__ labels, int *flags, int32_t *status)
{
    const int64_t f = blockIdx.y;
    const int64_t g = frame0 + f;
    const int64_t src = g - (g % interval);
    uint8_t *out = labels + f * N;
    const int64_t i0 = (int64_t)blockIdx.x * LB_SLICE + threadIdx.x;
    if (src < frame0) {
        for (int e = 0; e < LB_VPT; e++) {
            const int64_t i = i0 + (int64_t)e * LB_THREADS;
            if (i < N) out[i] = prev ? prev[i] : 0;
        }
        return;
    }
    const int64_t fs = src - frame0;
    const double *z = zc + fs * N;
    const double mean = zstat[fs * 2], tol = zstat[fs * 2 + 1];
    double v[LB_VPT];
#pragma unroll
    for (int e = 0; e < LB_VPT; e++) {
        const int64_t i = i0 + (int64_t)e * LB_THREADS;
        v[e] = i < N ? z[i] : mean + 2.0 * tol + 1.0;
    }
    int close = 0, tie = 0;
#pragma unroll
    for (int e = 0; e < LB_VPT; e++) {
        const int64_t i = i0 + (int64_t)e * LB_THREADS;
        if (i < N) {
            if (fabs(v[e] - mean) <= tol) close = 1;
            if (v[e] > mean) out[i] = 1;
            else { out[i] = 0; if (!(v[e] < mean)) tie = 1; }
        }
    }
    if (close) atomicOr(flags + f, 1);
    else if (tie) atomicOr(status, PBK_ST_LEAFLET_TIE);
}

__global__ void __launch_bounds__(512)
k_leaflets_c(const double *__restrict__ zc, int64_t N, int64_t frame0, int interval, const uint8_t *__restrict__ prev,
             uint8_t *__restrict__ labels, const int *__restrict__ only_flagged, int32_t *status)
{
    __shared__ double red[32];
    __shared__ double zexact;
    __shared__ int need_exact;
    const int64_t f = blockIdx.x;
    if (only_flagged && !only_flagged[f]) return;
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
    float *boxtab; double *node_vals; float *u_out; double *zc; int16_t *snap; int *flags;
    double *zpart, *zstat; int *lflags;   // labels: sweep-B partials [F][n_part][2], (mean, bound) [F][2], flags [F]
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
    s.flags = reinterpret_cast<int *>(b + o); o += sw_align((size_t)L * 4);
    s.zpart = reinterpret_cast<double *>(b + o); o += sw_align((size_t)F * ((N + LB_SLICE - 1) / LB_SLICE) * 16);
    s.zstat = reinterpret_cast<double *>(b + o); o += sw_align((size_t)F * 16);
    s.lflags = reinterpret_cast<int *>(b + o); o += sw_align((size_t)F * 4);
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
        if (F >= ((int64_t)1 << 26)) return PBK_EUNSUPPORTED;
        SweepAParams A;
        A.x = x; A.boxtab = S.boxtab; A.mass = mass; A.leaf_start = leaf_start; A.leaf_len = leaf_len;
        A.L = L; A.nodes_per_frame = L + K; A.nblk = (int)nch; A.prev0 = first_frame ? x : u_state;
        A.init_images = init_images; A.F = (int)F; A.M = M;
        A.net_frame = (int)(((net_frames > 0 && net_frames < F) ? net_frames : F) - 1);
        A.net_images = net_images; A.u_out = S.u_out; A.node_vals = S.node_vals; A.status = status;
        A.snap = S.snap; A.flags = S.flags;
        A.store_sums = (phase & 2) ? 1 : 0;
        cudaError_t e = cudaMemsetAsync(S.flags, 0, (size_t)L * 4, st);
        if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "flags memset", e);
        // persistent grid: every CTA resident (tasks wait for each other through flags)
        void (*ka)(SweepAParams) = bulk ? k_sweep_a<true> : k_sweep_a<false>;
        cudaFuncSetAttribute(ka, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ka, SA_THREADS, 0);
        if (e != cudaSuccess || per_sm < 1) return pbk_fail(ctx, PBK_ECUDA, "k_sweep_a occupancy", e);
        // grid = L when every leaf can have its own CTA (the CTA then continues its leaf block after
        // block, no waiting on another CTA); otherwise every resident slot, tasks round-robin
        int64_t grid = (int64_t)per_sm * ctx->sm_count;
        if (grid > L) grid = L;
        ka<<<(unsigned)grid, SA_THREADS, 0, st>>>(A);
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
        const int64_t n_slices = (N + LB_SLICE - 1) / LB_SLICE;
        if (F <= 65535) {                                 // (frames are the y dimension of the grid)
            k_zpart<<<dim3((unsigned)n_slices, (unsigned)F), LB_THREADS, 0, st>>>(S.zc, N, S.zpart);
            PBK_CHECK_LAUNCH(ctx, "k_zpart");
            k_zmean<<<(unsigned)F, 256, 0, st>>>(S.zpart, (int)n_slices, N, S.zstat, S.lflags);
            PBK_CHECK_LAUNCH(ctx, "k_zmean");
            k_labels<<<dim3((unsigned)n_slices, (unsigned)F), LB_THREADS, 0, st>>>(
                S.zc, S.zstat, N, frame0, update_interval, prev_labels, labels, S.lflags, status);
            PBK_CHECK_LAUNCH(ctx, "k_labels");
        }
        k_leaflets_c<<<(unsigned)F, 512, 0, st>>>(S.zc, N, frame0, update_interval, prev_labels, labels,
                                                  F <= 65535 ? S.lflags : nullptr, status);
        PBK_CHECK_LAUNCH(ctx, "k_leaflets_c");
    }
    cudaError_t e = cudaMemcpyAsync(u_state, S.u_out, (size_t)M * 12, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "u_state copy", e);
    return PBK_OK;
}
