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
    uint8_t *chg_meta;              // sweep A2: [nblk][3M] how each chain's count moved inside the chunk
    uint32_t *chg_mask;             //           [nblk][3M] frames of the chunk that use the second count
    int log2_depth;                 //           ring slots per warp = 1 << log2_depth
    const int *fb_flag;             //           non-NULL: run only when the time-parallel front end flagged the batch
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
// k_sweep_a2: one WARP per leaf (one 32-thread CTA), all frames of the batch in time order, no
// CTA-level barrier and no producer / reducer hand-over.
//   lane c < 24 = (accumulator rj = c / 3, axis rk = c % 3) of NumPy's 8-accumulator loop
//   (numpy/_core/src/umath/loops_utils.h.src): the lane owns float 24 q + c of every 8-atom row q of
//   the leaf -- consecutive lanes read consecutive shared words -- keeps the raw previous value and
//   the image count of its <= 16 rows in registers, and adds the products m*f64(u) of a frame in row
//   order, which IS accumulator rj of axis rk.  A 3-step butterfly gives the leaf sum, lanes 0-2
//   append the n % 8 tail atoms and store.  Two frames per step (their chains interleave).
//   The leaf's slice of every frame arrives by TMA bulk copy into a per-warp ring of `depth` frames
//   (one mbarrier per slot; lane 0 refills a slot as soon as the warp has read it).
//   A same-image test that fails sends only the failing lane through the exact rule (sa_exact);
//   anything unusual (count beyond +-127, box edge >= 2^13, |x| < 2^-8, a tail row, the frame whose
//   counts a shard reports) takes the frame-by-frame generic path for the whole warp.
// Besides the leaf sums the sweep leaves, per chunk of SW_CHUNK frames and chain: the count before
// the chunk (snap), and how the count moved inside it -- meta 0: not at all; 1 / 2: between the start
// count s0 and s0 + 1 / s0 - 1, bit f of cmask set = frame f of the chunk uses the second value;
// 255: anything else -- which lets sweep B form every frame of a chunk independently.
// ------------------------------------------------------------------------------------------
#define S2_ROWS 16
#define S2_SLOT 1536u                // bytes per ring slot: 128 atoms x 12 B

static __device__ __forceinline__ float s2_lds(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}

// a change of the image count of row q at local frame f of the chunk: old -> new
static __device__ __noinline__ void s2_note(uint32_t *cm, int16_t *ss, uint8_t *mt, int idx, bool first, int f, float so, float sn)
{
    if (first) { cm[idx] = 0u; ss[idx] = (int16_t)(int)so; mt[idx] = 0; }
    const int s0 = ss[idx];
    uint8_t m = mt[idx];
    const int isn = (int)sn;
    if (m == 0) m = (isn == s0 + 1) ? 1 : ((isn == s0 - 1) ? 2 : 255);
    const int alt = s0 + (m == 1 ? 1 : -1);
    if (m != 255 && isn != s0 && isn != alt) m = 255;
    const uint32_t msk = cm[idx] ^ (0xffffffffu << f);
    if (m != 255 && (((msk >> f) & 1u) != 0u) != (isn == alt)) m = 255;
    cm[idx] = msk;
    mt[idx] = m;
}

template <bool BULK>
__global__ void __launch_bounds__(32, 14)
k_sweep_a2(SweepAParams P)
{
    extern __shared__ __align__(128) unsigned char s2_raw[];
    if (P.fb_flag && *P.fb_flag == 0) return;
    const int lane = threadIdx.x, leaf = blockIdx.x;
    const int D = 1 << P.log2_depth;
    const uint32_t a_x = sw_smem_u32(s2_raw), a_bar = a_x + (uint32_t)D * S2_SLOT;
    uint32_t *s_cm = reinterpret_cast<uint32_t *>(s2_raw + (size_t)D * S2_SLOT + 8 * (size_t)D);   // [17][32]
    int16_t *s_ss = reinterpret_cast<int16_t *>(s_cm + 17 * 32);                                   // [17][32]
    uint8_t *s_mt = reinterpret_cast<uint8_t *>(s_ss + 17 * 32);                                   // [17][32]
    const int ls = P.leaf_start[leaf], n = P.leaf_len[leaf];
    const int nq = n >= 8 ? n / 8 : 0, rem = n - 8 * nq;
    const int c = lane < 24 ? lane : lane - 24;           // lanes 24-31 shadow lanes 0-7 and store nothing
    const bool act = lane < 24;
    const int rj = c / 3, rk = c % 3;
    const int s1 = (rj ^ 1) * 3 + rk, s2 = (rj ^ 2) * 3 + rk, s4 = (rj ^ 4) * 3 + rk;
    const int64_t C3 = 3 * P.M;
    const int F = P.F;
    const uint32_t nbytes = 12u * (uint32_t)n;
    const float *src0 = P.x + 3 * (int64_t)ls;
    const int64_t gc0 = 3 * (int64_t)ls + c;               // chain of row q: gc0 + 24 q
    const bool tailv = rem > 0 && c < 3 * rem;
    int bad = 0;
    if (lane == 0)
        for (int s = 0; s < D; s++) sw_mbar_init(a_bar + 8u * s, BULK ? 1u : 32u);
    sw_fence_async();
    __syncwarp();
    auto issue = [&](int t) {
        const uint32_t slot = (uint32_t)t & (uint32_t)(D - 1);
        const uint32_t bar = a_bar + 8u * slot, dst = a_x + slot * S2_SLOT;
        const float *src = src0 + (int64_t)t * C3;
        if (BULK) {
            if (lane == 0) { sw_mbar_expect_tx(bar, nbytes); sw_bulk_g2s(dst, src, nbytes, bar); }
        } else {
            for (int w = lane; w < 3 * n; w += 32) sw_cp4(dst + 4u * w, src + w);
            sw_cp_arrive(bar);
        }
    };
    for (int t = 0; t < D && t < F; t++) issue(t);

    // uniform masses (coarse-grained models): the mass is a register, not a load
    const double m_first = P.mass[ls];
    int uni = 1;
    for (int a = lane; a < n; a += 32) if (P.mass[ls + a] != m_first) uni = 0;
    const bool um = __all_sync(0xffffffffu, uni) != 0;
    const double *__restrict__ gm = P.mass + ls + rj;
#define S2_MASS(q) (um ? m_first : __ldg(gm + 8 * (q)))

    float pv[S2_ROWS], sf[S2_ROWS], pvT = 0.f, sfT = 0.f;
    int wide = 0;
#pragma unroll
    for (int q = 0; q < S2_ROWS; q++) {
        pv[q] = 0.f; sf[q] = 0.f;
        if (q < nq) {
            const int64_t gc = gc0 + 24 * q;
            pv[q] = P.prev0[gc];
            int s0 = P.init_images ? P.init_images[gc] : 0;
            if (s0 > 32767 || s0 < -32767) { bad |= 1; s0 = 0; }
            if (s0 > 127 || s0 < -127) wide = 1;
            sf[q] = (float)s0;
        }
    }
    if (tailv) {
        const int64_t gc = gc0 + 24 * nq;
        pvT = P.prev0[gc];
        int s0 = P.init_images ? P.init_images[gc] : 0;
        if (s0 > 32767 || s0 < -32767) { bad |= 1; s0 = 0; }
        if (s0 > 127 || s0 < -127) wide = 1;
        sfT = (float)s0;
    }
    const double mT = tailv ? P.mass[ls + 8 * nq + rj] : 0.0;
    float Lp = P.boxtab[rk];
    unsigned chm = 0u;                                    // rows (bit 16: tail) whose count moved in this chunk
    const bool has_net = P.net_images != nullptr;
    const bool plain = rem == 0 && nq > 0;

    // box rows of the next step are loaded a step ahead
    float bL[2], bh[2], bb[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
        const int t = min(u, F - 1);
        bL[u] = P.boxtab[t * 8 + rk]; bh[u] = P.boxtab[t * 8 + 3 + rk]; bb[u] = P.boxtab[t * 8 + 6];
    }

    for (int t0 = 0; t0 < F; t0 += 2) {
        const int nfu = min(2, F - t0);
        const float Lc0 = bL[0], Lc1 = bL[1], hc0 = bh[0], hc1 = bh[1];
        const bool big = bb[0] != 0.0f || bb[1] != 0.0f;
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int t = min(t0 + 2 + u, F - 1);
            bL[u] = __ldg(P.boxtab + t * 8 + rk); bh[u] = __ldg(P.boxtab + t * 8 + 3 + rk); bb[u] = __ldg(P.boxtab + t * 8 + 6);
        }
        const uint32_t slot0 = (uint32_t)t0 & (uint32_t)(D - 1), slot1 = (uint32_t)(t0 + 1) & (uint32_t)(D - 1);
        if (!sw_mbar_wait(a_bar + 8u * slot0, (uint32_t)(t0 >> P.log2_depth) & 1u)) bad |= 2;
        if (nfu == 2 && !sw_mbar_wait(a_bar + 8u * slot1, (uint32_t)((t0 + 1) >> P.log2_depth) & 1u)) bad |= 2;
        const uint32_t b0 = a_x + slot0 * S2_SLOT + 4u * (uint32_t)c, b1 = a_x + slot1 * S2_SLOT + 4u * (uint32_t)c;
        const int fl0 = t0 & (SW_CHUNK - 1);              // local frame of t0 in its chunk (t0 is even: t0 + 1 is in the same chunk)
        bool generic = wide || big || nfu < 2 || !plain || (has_net && P.net_frame >= t0 && P.net_frame < t0 + 2);
        generic = __any_sync(0xffffffffu, generic) != 0;
        double r0 = 0.0, r1 = 0.0;
        if (!generic) {
            float x0[S2_ROWS], x1[S2_ROWS];
            float mx0 = 0.f, mx1 = 0.f, tiny = 3.0e38f;
#pragma unroll
            for (int q = 0; q < S2_ROWS; q++) {
                x0[q] = 1.0f; x1[q] = 1.0f;
                if (q < nq) { x0[q] = s2_lds(b0 + 96u * q); x1[q] = s2_lds(b1 + 96u * q); }
            }
#pragma unroll
            for (int q = 0; q < S2_ROWS; q++) {
                if (q < nq) {
                    mx0 = fmaxf(mx0, fabsf(__fsub_rn(x0[q], pv[q])));
                    mx1 = fmaxf(mx1, fabsf(__fsub_rn(x1[q], x0[q])));
                    tiny = fminf(tiny, fminf(fabsf(x0[q]), fabsf(x1[q])));
                }
            }
            generic = __any_sync(0xffffffffu, !(tiny >= 0.00390625f)) != 0;
            if (!generic) {
                float sm[S2_ROWS];
#pragma unroll
                for (int q = 0; q < S2_ROWS; q++) sm[q] = sf[q];
                if (!(mx0 < hc0) || !(mx1 < hc1)) {        // this lane only: the exact rule for the rows that failed
#pragma unroll
                    for (int q = 0; q < S2_ROWS; q++) {
                        if (q < nq) {
                            if (!(fabsf(__fsub_rn(x0[q], pv[q])) < hc0)) {
                                const float sn = sa_exact(x0[q], pv[q], sf[q], Lc0, Lp, &bad);
                                if (sn != sf[q]) {
                                    s2_note(s_cm, s_ss, s_mt, q * 32 + lane, !((chm >> q) & 1u), fl0, sf[q], sn);
                                    chm |= 1u << q;
                                    sf[q] = sn;
                                    if (fabsf(sn) > 127.0f) wide = 1;
                                }
                            }
                            sm[q] = sf[q];
                            if (wide || !(fabsf(__fsub_rn(x1[q], x0[q])) < hc1)) {
                                const float sn = sa_exact(x1[q], x0[q], sf[q], Lc1, Lc0, &bad);
                                if (sn != sf[q]) {
                                    s2_note(s_cm, s_ss, s_mt, q * 32 + lane, !((chm >> q) & 1u), fl0 + 1, sf[q], sn);
                                    chm |= 1u << q;
                                    sf[q] = sn;
                                    if (fabsf(sn) > 127.0f) wide = 1;
                                }
                            }
                        }
                    }
                }
                // (counts up to +-255 still take the single-FMA unwrap exactly, pbk_common.cuh; a lane that
                //  went wide here sends the warp through the generic path from the next step on)
#pragma unroll
                for (int q = 0; q < S2_ROWS; q++) {
                    if (q < nq) {
                        const double m = S2_MASS(q);
                        const double p0 = __dmul_rn(m, (double)__fmaf_rn(sm[q], Lc0, x0[q]));
                        const double p1 = __dmul_rn(m, (double)__fmaf_rn(sf[q], Lc1, x1[q]));
                        r0 = (q == 0) ? p0 : __dadd_rn(r0, p0);
                        r1 = (q == 0) ? p1 : __dadd_rn(r1, p1);
                        pv[q] = x1[q];
                    }
                }
                Lp = Lc1;
                __syncwarp();
                r0 = __dadd_rn(r0, __shfl_sync(0xffffffffu, r0, s1)); r1 = __dadd_rn(r1, __shfl_sync(0xffffffffu, r1, s1));
                r0 = __dadd_rn(r0, __shfl_sync(0xffffffffu, r0, s2)); r1 = __dadd_rn(r1, __shfl_sync(0xffffffffu, r1, s2));
                r0 = __dadd_rn(r0, __shfl_sync(0xffffffffu, r0, s4)); r1 = __dadd_rn(r1, __shfl_sync(0xffffffffu, r1, s4));
                if (lane < 3 && P.store_sums) {
                    double *out = P.node_vals + ((int64_t)t0 * P.nodes_per_frame + leaf) * 3 + lane;
                    out[0] = r0;
                    out[(int64_t)P.nodes_per_frame * 3] = r1;
                }
            }
        }
        if (generic) {
            // ---- frame by frame, every case (the arithmetic of k_sweep_a's slow path) -------------------
            for (int u = 0; u < nfu; u++) {
                const int t = t0 + u;
                const uint32_t bx = u ? b1 : b0;
                const float Lc = u ? Lc1 : Lc0, hc = wide ? -1.0f : (u ? hc1 : hc0);
                const bool Lbig = __ldg(P.boxtab + t * 8 + 6) != 0.0f;
                double r = 0.0;
#pragma unroll
                for (int q = 0; q < S2_ROWS; q++) {
                    if (q < nq) {
                        const float xv = s2_lds(bx + 96u * q);
                        if (!(fabsf(__fsub_rn(xv, pv[q])) < hc)) {
                            const float sn = sa_exact(xv, pv[q], sf[q], Lc, Lp, &bad);
                            if (sn != sf[q]) {
                                s2_note(s_cm, s_ss, s_mt, q * 32 + lane, !((chm >> q) & 1u), fl0 + u, sf[q], sn);
                                chm |= 1u << q;
                                sf[q] = sn;
                                if (fabsf(sn) > 127.0f) wide = 1;
                            }
                        }
                        float uu = __fmaf_rn(sf[q], Lc, xv);
                        if (!(fabsf(xv) >= 0.00390625f) || Lbig || wide) uu = sw_unwrap_slow(xv, (int)sf[q], Lc);
                        const double p = __dmul_rn(S2_MASS(q), (double)uu);
                        r = (q == 0) ? p : __dadd_rn(r, p);
                        pv[q] = xv;
                    }
                }
                double pT = 0.0;
                if (tailv) {
                    const float xv = s2_lds(bx + 96u * nq);
                    if (!(fabsf(__fsub_rn(xv, pvT)) < hc)) {
                        const float sn = sa_exact(xv, pvT, sfT, Lc, Lp, &bad);
                        if (sn != sfT) {
                            s2_note(s_cm, s_ss, s_mt, 16 * 32 + lane, !((chm >> 16) & 1u), fl0 + u, sfT, sn);
                            chm |= 1u << 16;
                            sfT = sn;
                            if (fabsf(sn) > 127.0f) wide = 1;
                        }
                    }
                    pT = __dmul_rn(mT, (double)sw_unwrap_slow(xv, (int)sfT, Lc));
                    pvT = xv;
                }
                if (has_net && t == P.net_frame && act) {
#pragma unroll
                    for (int q = 0; q < S2_ROWS; q++) if (q < nq) P.net_images[gc0 + 24 * q] = (int)sf[q];
                    if (tailv) P.net_images[gc0 + 24 * nq] = (int)sfT;
                }
                Lp = Lc;
                __syncwarp();
                r = __dadd_rn(r, __shfl_sync(0xffffffffu, r, s1));
                r = __dadd_rn(r, __shfl_sync(0xffffffffu, r, s2));
                r = __dadd_rn(r, __shfl_sync(0xffffffffu, r, s4));
                if (nq == 0) r = 0.0;
                for (int a = 0; a < rem; a++) {
                    const double o = __shfl_sync(0xffffffffu, pT, (3 * a + lane) & 31);
                    if (lane < 3) r = __dadd_rn(r, o);
                }
                if (lane < 3 && P.store_sums) P.node_vals[((int64_t)t * P.nodes_per_frame + leaf) * 3 + lane] = r;
            }
        }
        // both slots are read: refill them (consumer release: the reads above are ordered before the
        // async-proxy writes by the warp barrier and the issue that follows it)
        __syncwarp();
        if (t0 + D < F) issue(t0 + D);
        if (nfu == 2 && t0 + 1 + D < F) issue(t0 + 1 + D);
        // ---- end of a chunk: how the counts moved, and the counts the next chunk starts from -----------
        const int te = t0 + nfu;                          // frames done
        if ((te & (SW_CHUNK - 1)) == 0 || te == F) {
            const int blk = (te - 1) / SW_CHUNK;
            if (act) {
                uint8_t *mo = P.chg_meta + (int64_t)blk * C3 + gc0;
                uint32_t *co = P.chg_mask + (int64_t)blk * C3 + gc0;
#pragma unroll
                for (int q = 0; q < S2_ROWS; q++) {
                    if (q < nq) {
                        const bool ch = (chm >> q) & 1u;
                        mo[24 * q] = ch ? s_mt[q * 32 + lane] : (uint8_t)0;
                        co[24 * q] = ch ? s_cm[q * 32 + lane] : 0u;
                    }
                }
                if (tailv) {
                    const bool ch = (chm >> 16) & 1u;
                    mo[24 * nq] = ch ? s_mt[16 * 32 + lane] : (uint8_t)0;
                    co[24 * nq] = ch ? s_cm[16 * 32 + lane] : 0u;
                }
                if (te < F) {
                    int16_t *so = P.snap + (int64_t)(blk + 1) * C3 + gc0;
#pragma unroll
                    for (int q = 0; q < S2_ROWS; q++) if (q < nq) so[24 * q] = (int16_t)(int)sf[q];
                    if (tailv) so[24 * nq] = (int16_t)(int)sfT;
                } else if (P.u_out) {
#pragma unroll
                    for (int q = 0; q < S2_ROWS; q++) if (q < nq) P.u_out[gc0 + 24 * q] = pbk_unwrapped(pv[q], (int)sf[q], Lp);
                    if (tailv) P.u_out[gc0 + 24 * nq] = pbk_unwrapped(pvT, (int)sfT, Lp);
                }
            }
            chm = 0u;
        }
    }
#undef S2_MASS
    if (bad) atomicOr(P.status, (bad & 1 ? PBK_ST_IMAGE_OVERFLOW : 0) | (bad & 2 ? PBK_ST_INTERNAL : 0));
}

// ------------------------------------------------------------------------------------------
// Time-parallel front end (the default): the image counts never have to be walked frame by frame,
// because the same-image test compares consecutive RAW frames.
//   k_sweep_p   thread = one float4 column (4 chains) x one chunk of SW_CHUNK frames: per chain the net
//               change of the count over the chunk (dsum) and how it moved inside it (meta / cmask as
//               described above), all relative to the unknown count at the chunk start.  A jump that
//               fails the test changes the count by the reference's shift of the raw jump; if the shifted
//               jump is again within the rounding bound of +-L/2, or the motion cannot be written as
//               (s0, s0 +- 1), the batch is flagged and the warp-per-leaf sweep k_sweep_a2 redoes it
//               frame by frame (it then overwrites everything this front end wrote).
//   k_sweep_s   thread = chain: prefix over the chunks -> snap (counts before every chunk), the unwrapped
//               last frame (state carried to the next batch), a shard's net counts; counts beyond +-127
//               (the range the rounding bound is derived for) flag the batch as well.
//   k_sweep_l   warp = (leaf, SL_FRAMES frames of one chunk): lanes as in k_sweep_a2, the counts of the
//               lane's rows are constants of the chunk (two values and a mask), every frame independent.
// ------------------------------------------------------------------------------------------
#define SP_THREADS 128
#define SP_UNR 8

struct SweepPParams {
    const float *x; const float *boxtab; const float *prev0;
    int64_t F, C3; int nblk;
    int8_t *dsum; uint8_t *meta; uint32_t *mask; int *fb_flag; int32_t *status;
    const double *mass; int *nonuniform_mass;
};

// one failed same-image test: the reference's shift of the raw jump d0 (0x7fff: the shifted jump is
// again within the rounding bound of +-L/2, or the box is bad -> the batch is flagged)
static __device__ __noinline__ int sp_shift(float d0, float Lc, float hc)
{
    int bad = 0;
    const int ds = sw_shift(d0, Lc, &bad);
    const double sh = __dadd_rn((double)d0, __dmul_rn((double)ds, (double)Lc));
    if (bad || ds == 0 || !((float)fabs(sh) < hc)) return 0x7fff;
    return ds;
}
// the move of the relative count rel by ds at local frame f
#define SP_MOVE(d0, Lc, hc, e) do { \
        const int ds_ = sp_shift(d0, Lc, hc); \
        if (ds_ == 0x7fff) flag = 1; \
        else { \
            const int nr_ = rel[e] + ds_; \
            if (mt[e] == 0) mt[e] = (nr_ == 1) ? 1 : ((nr_ == -1) ? 2 : 255); \
            const int alt_ = mt[e] == 1 ? 1 : -1; \
            if (nr_ != 0 && nr_ != alt_) mt[e] = 255; \
            msk[e] ^= 0xffffffffu << f; \
            if ((((msk[e] >> f) & 1u) != 0u) != (nr_ == alt_)) mt[e] = 255; \
            rel[e] = nr_; \
        } } while (0)

__global__ void __launch_bounds__(SP_THREADS)
k_sweep_p(SweepPParams P)
{
    __shared__ float s_box[SW_CHUNK][8];
    const int blk = blockIdx.y;
    const int64_t f0 = (int64_t)blk * SW_CHUNK, f1 = min(P.F, f0 + SW_CHUNK);
    const int nf = (int)(f1 - f0);
    for (int r = threadIdx.x; r < nf * 8; r += SP_THREADS) s_box[r >> 3][r & 7] = P.boxtab[f0 * 8 + r];
    __syncthreads();
    const int64_t col = (int64_t)blockIdx.x * SP_THREADS + threadIdx.x;
    const int64_t ncol = P.C3 >> 2;
    if (col >= ncol) return;
    if (blk == 0) {                                   // any atom whose mass differs from the first one's?
        const int64_t a0 = (4 * col) / 3, a1 = (4 * col + 3) / 3;
        const double m0 = P.mass[0];
        if (P.mass[a0] != m0 || P.mass[a1] != m0) *P.nonuniform_mass = 1;
    }
    const int ke0 = (int)((4 * col) % 3);
    const float4 *xc = reinterpret_cast<const float4 *>(P.x) + col;
    float4 pvv = blk == 0 ? reinterpret_cast<const float4 *>(P.prev0)[col] : xc[(f0 - 1) * ncol];
    float pv[4] = {pvv.x, pvv.y, pvv.z, pvv.w};
    int rel[4] = {0, 0, 0, 0}, mt[4] = {0, 0, 0, 0};
    uint32_t msk[4] = {0u, 0u, 0u, 0u};
    int flag = 0;
    const int k0 = ke0, k1 = (ke0 + 1) % 3, k2 = (ke0 + 2) % 3;       // axes of chains 0..3: k0, k1, k2, k0
    for (int fb = 0; fb < nf; fb += SP_UNR) {
        float4 v[SP_UNR];
#pragma unroll
        for (int u = 0; u < SP_UNR; u++)
            if (fb + u < nf) v[u] = __ldcs(xc + (f0 + fb + u) * ncol);
#pragma unroll
        for (int u = 0; u < SP_UNR; u++) {
            if (fb + u < nf) {
                const int f = fb + u;
                const float xv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                const float h0 = s_box[f][3 + k0], h1 = s_box[f][3 + k1], h2 = s_box[f][3 + k2];
                const float d0 = __fsub_rn(xv[0], pv[0]), d1 = __fsub_rn(xv[1], pv[1]), d2 = __fsub_rn(xv[2], pv[2]),
                            d3 = __fsub_rn(xv[3], pv[3]);
                const float m = fminf(fminf(h0 - fabsf(d0), h1 - fabsf(d1)), fminf(h2 - fabsf(d2), h0 - fabsf(d3)));
                if (!(m > 0.0f)) {
                    if (!(h0 - fabsf(d0) > 0.0f)) SP_MOVE(d0, s_box[f][k0], h0, 0);
                    if (!(h1 - fabsf(d1) > 0.0f)) SP_MOVE(d1, s_box[f][k1], h1, 1);
                    if (!(h2 - fabsf(d2) > 0.0f)) SP_MOVE(d2, s_box[f][k2], h2, 2);
                    if (!(h0 - fabsf(d3) > 0.0f)) SP_MOVE(d3, s_box[f][k0], h0, 3);
                }
#pragma unroll
                for (int e = 0; e < 4; e++) pv[e] = xv[e];
            }
        }
    }
    int cplx = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) if (mt[e] == 255 || rel[e] > 127 || rel[e] < -127) cplx = 1;
    if (flag || cplx) atomicOr(P.fb_flag, 1);
    const int64_t o = (int64_t)blk * ncol + col;
    reinterpret_cast<uint32_t *>(P.dsum)[o] = (uint32_t)(rel[0] & 0xff) | ((uint32_t)(rel[1] & 0xff) << 8) |
                                              ((uint32_t)(rel[2] & 0xff) << 16) | ((uint32_t)(rel[3] & 0xff) << 24);
    reinterpret_cast<uint32_t *>(P.meta)[o] = (uint32_t)mt[0] | ((uint32_t)mt[1] << 8) | ((uint32_t)mt[2] << 16) | ((uint32_t)mt[3] << 24);
    reinterpret_cast<uint4 *>(P.mask)[o] = make_uint4(msk[0], msk[1], msk[2], msk[3]);
}

struct SweepSParams {
    const float *x; const float *boxtab; const int32_t *init_images;
    int64_t F, C3; int nblk; int net_frame;
    const int8_t *dsum; const uint8_t *meta; const uint32_t *mask;
    int16_t *snap; float *u_out; int32_t *net_images; int *fb_flag;
};

__global__ void __launch_bounds__(256)
k_sweep_s(SweepSParams P)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.C3) return;
    int s = P.init_images ? P.init_images[c] : 0;
    int over = 0;
    const int nb = P.net_images ? P.net_frame / SW_CHUNK : -1;
    for (int b = 0; b < P.nblk; b++) {
        if (s > 126 || s < -126) over = 1;
        P.snap[(int64_t)b * P.C3 + c] = (int16_t)max(-32767, min(32767, s));
        if (b == nb) {
            const int mt = P.meta[(int64_t)b * P.C3 + c];
            const int on = mt != 0 && ((P.mask[(int64_t)b * P.C3 + c] >> (P.net_frame & (SW_CHUNK - 1))) & 1u);
            P.net_images[c] = s + (on ? (mt == 1 ? 1 : -1) : 0);
        }
        s += (int)P.dsum[(int64_t)b * P.C3 + c];
    }
    if (s > 126 || s < -126) over = 1;
    if (over) atomicOr(P.fb_flag, 1);
    if (P.u_out) {
        const int k = (int)(c % 3);
        P.u_out[c] = pbk_unwrapped(P.x[(P.F - 1) * P.C3 + c], s, P.boxtab[(P.F - 1) * 8 + k]);
    }
}

#define SL_FRAMES SW_CHUNK          // a warp = one leaf x one chunk of frames
#define SL_DEPTH 8                  // ring slots per warp
#define SL_WARPS 4

struct SweepLParams {
    const float *x; const float *boxtab; const double *mass; const int32_t *leaf_start, *leaf_len;
    int L, nodes_per_frame; int64_t F, C3;
    const int16_t *snap; const uint8_t *meta; const uint32_t *mask;
    double *node_vals; const int *fb_flag; const int *nonuniform_mass; int32_t *status;
};

// NQ: rows held per lane.  TIGHT: every leaf of the batch has NQ - 1 or NQ full rows (what NumPy's
// halving produces for most atom counts), so only the last row needs its guard; otherwise every row does.
template <int NQ, bool TIGHT>
__global__ void __launch_bounds__(SL_WARPS * 32)
k_sweep_l(SweepLParams P)
{
    // per warp: a ring of SL_DEPTH leaf slices (TMA bulk copies, one mbarrier each), refilled as they are read
    extern __shared__ __align__(128) unsigned char sl_raw[];
    if (*P.fb_flag) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int leaf = blockIdx.x * SL_WARPS + warp;
    if (leaf >= P.L) return;
    const int64_t t0 = (int64_t)blockIdx.y * SL_FRAMES, t1 = min(P.F, t0 + SL_FRAMES);
    const int nfr = (int)(t1 - t0);
    const int blk = (int)(t0 / SW_CHUNK);
    const int fl0 = (int)(t0 & (SW_CHUNK - 1));
    const int ls = P.leaf_start[leaf], n = P.leaf_len[leaf];
    const float *s_x = reinterpret_cast<const float *>(sl_raw) + (size_t)warp * (SL_DEPTH * (S2_SLOT / 4));
    const uint32_t a_x = sw_smem_u32(sl_raw) + (uint32_t)warp * (SL_DEPTH * S2_SLOT);
    const uint32_t a_bar = sw_smem_u32(sl_raw) + SL_WARPS * SL_DEPTH * S2_SLOT + (uint32_t)warp * (SL_DEPTH * 8u);
    const float *src = P.x + t0 * P.C3 + 3 * (int64_t)ls;
    if (lane == 0) {
        for (int u = 0; u < SL_DEPTH; u++) sw_mbar_init(a_bar + 8u * u, 1u);
        sw_fence_async();
        for (int u = 0; u < nfr && u < SL_DEPTH; u++) {
            sw_mbar_expect_tx(a_bar + 8u * u, 12u * (uint32_t)n);
            sw_bulk_g2s(a_x + (uint32_t)u * S2_SLOT, src + (int64_t)u * P.C3, 12u * (uint32_t)n, a_bar + 8u * u);
        }
    }
    __syncwarp();
    const int nq = n >= 8 ? n / 8 : 0, rem = n - 8 * nq;
    const int c = lane < 24 ? lane : lane - 24;
    const int rj = c / 3, rk = c % 3;
    const int s1 = (rj ^ 1) * 3 + rk, s2 = (rj ^ 2) * 3 + rk, s4 = (rj ^ 4) * 3 + rk;
    const int64_t gc0 = 3 * (int64_t)ls + c;
    const bool tailv = rem > 0 && c < 3 * rem;
    const bool um = *P.nonuniform_mass == 0;          // (k_sweep_p looked at every atom)
    const double m_first = P.mass[0];
    const double *__restrict__ gm = P.mass + ls + rj;
#define SL_ROW(q) (TIGHT ? ((q) < NQ - 1 || nq == NQ) : ((q) < nq))
    // counts of the lane's rows in this chunk: sa (start), sb (the other value), mk (frames that use sb)
    float sa[NQ], sb[NQ], saT = 0.f, sbT = 0.f;
    uint32_t mk[NQ], mkT = 0u;
    const int64_t so = (int64_t)blk * P.C3 + gc0;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        sa[q] = 0.f; sb[q] = 0.f; mk[q] = 0u;
        if (SL_ROW(q)) {
            const int s0 = P.snap[so + 24 * q], mt = P.meta[so + 24 * q];
            sa[q] = (float)s0;
            sb[q] = (float)(s0 + (mt == 1 ? 1 : 0) - (mt == 2 ? 1 : 0));
            mk[q] = P.mask[so + 24 * q];
        }
    }
    if (tailv) {
        const int s0 = P.snap[so + 24 * nq], mt = P.meta[so + 24 * nq];
        saT = (float)s0;
        sbT = (float)(s0 + (mt == 1 ? 1 : 0) - (mt == 2 ? 1 : 0));
        if (mt) mkT = P.mask[so + 24 * nq];
    }
    const double mT = tailv ? P.mass[ls + 8 * nq + rj] : 0.0;
    int bad = 0;
    for (int u = 0; u < nfr; u++) {
        const int64_t t = t0 + u;
        const int f = fl0 + u;
        const float Lc = __ldg(P.boxtab + t * 8 + rk);
        const bool big = __ldg(P.boxtab + t * 8 + 6) != 0.0f;
        const int slot = u & (SL_DEPTH - 1);
        if (!sw_mbar_wait(a_bar + 8u * slot, (uint32_t)(u / SL_DEPTH) & 1u)) bad = 1;
        const float *xs = s_x + (size_t)slot * (S2_SLOT / 4) + c;
        float xv[NQ];
        float tiny = 3.0e38f;
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            xv[q] = 1.0f;
            if (SL_ROW(q)) xv[q] = xs[24 * q];
            tiny = fminf(tiny, fabsf(xv[q]));
        }
        const bool slow = __any_sync(0xffffffffu, big || !(tiny >= 0.00390625f)) != 0;
        double r = 0.0;
        if (!slow && um) {
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                if (SL_ROW(q)) {
                    const float cq = ((mk[q] >> f) & 1u) ? sb[q] : sa[q];
                    const double p = __dmul_rn(m_first, (double)__fmaf_rn(cq, Lc, xv[q]));
                    r = (q == 0) ? p : __dadd_rn(r, p);
                }
            }
        } else if (!slow) {
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                if (SL_ROW(q)) {
                    const float cq = ((mk[q] >> f) & 1u) ? sb[q] : sa[q];
                    const double p = __dmul_rn(__ldg(gm + 8 * q), (double)__fmaf_rn(cq, Lc, xv[q]));
                    r = (q == 0) ? p : __dadd_rn(r, p);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                if (SL_ROW(q)) {
                    const float cq = ((mk[q] >> f) & 1u) ? sb[q] : sa[q];
                    const double p = __dmul_rn(__ldg(gm + 8 * q), (double)sw_unwrap_slow(xv[q], (int)cq, Lc));
                    r = (q == 0) ? p : __dadd_rn(r, p);
                }
            }
        }
        double pT = 0.0;
        if (tailv) {
            const float xt = xs[24 * nq];
            const float cq = ((mkT >> f) & 1u) ? sbT : saT;
            pT = __dmul_rn(mT, (double)sw_unwrap_slow(xt, (int)cq, Lc));
        }
        __syncwarp();
        r = __dadd_rn(r, __shfl_sync(0xffffffffu, r, s1));
        r = __dadd_rn(r, __shfl_sync(0xffffffffu, r, s2));
        r = __dadd_rn(r, __shfl_sync(0xffffffffu, r, s4));
        if (rem != 0 || nq == 0) {
            if (nq == 0) r = 0.0;
            for (int a = 0; a < rem; a++) {
                const double o = __shfl_sync(0xffffffffu, pT, (3 * a + lane) & 31);
                if (lane < 3) r = __dadd_rn(r, o);
            }
        }
        if (lane < 3) P.node_vals[(t * P.nodes_per_frame + leaf) * 3 + lane] = r;
        if (lane == 0 && u + SL_DEPTH < nfr) {      // (the shuffles above ordered every lane's reads of the slot before this)
            sw_mbar_expect_tx(a_bar + 8u * slot, 12u * (uint32_t)n);
            sw_bulk_g2s(a_x + (uint32_t)slot * S2_SLOT, src + (int64_t)(u + SL_DEPTH) * P.C3, 12u * (uint32_t)n, a_bar + 8u * slot);
        }
    }
    if (bad) atomicOr(P.status, PBK_ST_INTERNAL);
#undef SL_ROW
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
    double *com, *com_unwrap; float *zc; int32_t *status;    // zc: float32 copies of com_unwrap[norm] (label screen)
    const uint8_t *chg_meta; const uint32_t *chg_mask;   // from sweep A2 (NULL: sequential chunks only)
    const int *fb_flag;          // k_sweep_b: run only if set (NULL: always); k_sweep_bt: run only if clear
    const int *nonuniform_mass;
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

// (CTA size by bead size: this kernel is launched after every time-parallel reduction only to find
//  the flag clear, so few large CTAs for the common single-atom shapes)
#define SBS_THREADS(NA) ((NA) == 0 ? 128 : ((NA) <= 4 ? 512 : 256))
template <int NA>
__global__ void __launch_bounds__(SBS_THREADS(NA))
k_sweep_b(SweepBParams P)
{
    extern __shared__ __align__(16) unsigned char sb_raw[];
    __shared__ float s_box[SW_CHUNK][8];
    __shared__ double s_mem[SW_CHUNK][3];
    int16_t *s_img = reinterpret_cast<int16_t *>(sb_raw);       // NA == 0: [max_atoms][SBS_THREADS(NA)]
    if (P.fb_flag && *P.fb_flag == 0) return;
    const int tid = threadIdx.x;
    const int64_t C3 = 3 * P.M;
    const int64_t f0 = (int64_t)blockIdx.y * SW_CHUNK, f1 = min(P.F, f0 + SW_CHUNK);
    const int nf = (int)(f1 - f0);
    for (int r = tid; r < nf * 8; r += SBS_THREADS(NA)) s_box[r >> 3][r & 7] = P.boxtab[f0 * 8 + r];
    for (int r = tid; r < nf * 3; r += SBS_THREADS(NA)) s_mem[r / 3][r % 3] = P.mem_com[f0 * 3 + r];
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * SBS_THREADS(NA) + tid;
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
    float *oz = (k == P.norm) ? P.zc + f0 * P.N + i : nullptr;

    if (NA > 0) {
        float pv[NA > 0 ? NA : 1], xn[NA > 0 ? NA : 1];
        int sc[NA > 0 ? NA : 1], ix[NA > 0 ? NA : 1];
        // ---- counts at the chunk start and how sweep A saw them move inside the chunk ------------
        bool fastb = P.chg_meta != nullptr;
        int alt[NA > 0 ? NA : 1];
        uint32_t mk[NA > 0 ? NA : 1];
        bool umb = true;                              // every atom of the bead has the same mass
        const double mb0 = n > 0 ? gm[gat ? gat[p0] : p0] : 0.0;
#pragma unroll
        for (int j = 0; j < NA; j++) {
            sc[j] = 0; ix[j] = 0; alt[j] = 0; mk[j] = 0u;
            if (j < n) {
                const int a = gat ? gat[p0 + j] : p0 + j;
                ix[j] = 3 * a + k;
                int s0 = first_chunk ? (P.init_images ? P.init_images[ix[j]] : 0) : (int)P.snap[(int64_t)blockIdx.y * C3 + ix[j]];
                if (s0 > 32767 || s0 < -32767) { bad |= 1; s0 = 0; }
                if (s0 > 127 || s0 < -127) wide = 1;
                sc[j] = s0;
                if (gm[a] != mb0) umb = false;
                if (fastb) {
                    const int mt = P.chg_meta[(int64_t)blockIdx.y * C3 + ix[j]];
                    if (mt == 255) fastb = false;
                    alt[j] = s0 + (mt == 1 ? 1 : 0) - (mt == 2 ? 1 : 0);
                    if (mt != 0) mk[j] = P.chg_mask[(int64_t)blockIdx.y * C3 + ix[j]];
                }
            }
        }
        if (fastb && !wide) {
            // every frame of the chunk on its own: the count of atom j in frame f is sc[j] or alt[j]
            const double rlm = 1.0 / lm;
            const float *xf = P.x + f0 * C3;
            constexpr int SBU = NA == 1 ? 16 : (NA <= 4 ? 4 : 2);      // frames whose loads are in flight together
            for (int fb = 0; fb < nf; fb += SBU) {
                float xr[SBU][NA > 0 ? NA : 1];
#pragma unroll
                for (int u = 0; u < SBU; u++) {
#pragma unroll
                    for (int j = 0; j < NA; j++) {
                        xr[u][j] = 0.f;
                        if (fb + u < nf && j < n) xr[u][j] = __ldg(xf + (int64_t)(fb + u) * C3 + ix[j]);
                    }
                }
#pragma unroll
                for (int u = 0; u < SBU; u++) {
                    const int f = fb + u;
                    if (f < nf) {
                        const float Lc = s_box[f][k];
                        const bool Lbig = s_box[f][6] != 0.0f;
                        const double mem = s_mem[f][k];
                        double cw = 0.0, cu = 0.0;
#pragma unroll
                        for (int j = 0; j < NA; j++) {
                            if (j < n) {
                                const int sn = ((mk[j] >> f) & 1u) ? alt[j] : sc[j];
                                sb_accumulate(xr[u][j], sn, Lc, Lbig, mem, umb ? mb0 : __ldg(gm + ix[j] / 3), cw, cu);
                            }
                        }
                        oc[(int64_t)f * 3 * P.N] = pbk_div_r(cw, lm, rlm);
                        const double uu = pbk_div_r(cu, lm, rlm);
                        ou[(int64_t)f * 3 * P.N] = uu;
                        if (oz) oz[(int64_t)f * P.N] = (float)uu;
                    }
                }
            }
            if (bad) atomicOr(P.status, PBK_ST_IMAGE_OVERFLOW);
            return;
        }
#pragma unroll
        for (int j = 0; j < NA; j++) {
            pv[j] = 0.f; xn[j] = 0.f;
            if (j < n) {
                pv[j] = prow[ix[j]];
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
            if (oz) oz[(int64_t)f * P.N] = (float)uu;
            Lp = Lc;
        }
    } else {
        for (int j = 0; j < n; j++) {
            const int a = gat ? gat[p0 + j] : p0 + j;
            int s0 = first_chunk ? (P.init_images ? P.init_images[3 * a + k] : 0) : (int)P.snap[(int64_t)blockIdx.y * C3 + 3 * a + k];
            if (s0 > 32767 || s0 < -32767) { bad |= 1; s0 = 0; }
            if (s0 > 127 || s0 < -127) wide = 1;
            s_img[j * SBS_THREADS(NA) + tid] = (int16_t)s0;
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
                const int s = (int)s_img[j * SBS_THREADS(NA) + tid];
                const int sn = sb_count(xv, pvv, s, hc, Lc, Lp, wide != 0, &bad);
                if (sn != s) {
                    s_img[j * SBS_THREADS(NA) + tid] = (int16_t)sn;
                    if (sn > 127 || sn < -127) wide = 1;
                }
                sb_accumulate(xv, sn, Lc, Lbig, mem, __ldg(gm + a), cw, cu);
            }
            oc[(int64_t)f * 3 * P.N] = __ddiv_rn(cw, lm);
            const double uu = __ddiv_rn(cu, lm);
            ou[(int64_t)f * 3 * P.N] = uu;
            if (oz) oz[(int64_t)f * P.N] = (float)uu;
            Lp = Lc;
            prow = cur;
        }
    }
    if (bad) atomicOr(P.status, PBK_ST_IMAGE_OVERFLOW);
}

// ------------------------------------------------------------------------------------------
// k_sweep_bt<NA, GATHER>: sweep B behind the time-parallel front end.  Thread = (lipid, axis) x
// SBT_FRAMES frames of one chunk; the counts of the bead's atoms are two values and a mask per atom
// (snap / meta / cmask), so every frame is formed on its own, SBU frames of loads in flight together,
// no branch per atom: u = fma(count, L, x) is exact for count 0 too, and the rare frames that need the
// float64 unwrap (|x| < 2^-8 with a count, box edge >= 2^13) are redone by sb_accumulate.
// Runs only when the front end described the whole batch (fb_flag == 0); k_sweep_b does the rest.
// ------------------------------------------------------------------------------------------
#define SBT_FRAMES 16

template <int NA, bool GATHER>
__global__ void __launch_bounds__(SB_THREADS)
k_sweep_bt(SweepBParams P)
{
    constexpr int SBTF = NA == 1 ? 2 * SBT_FRAMES : SBT_FRAMES;      // frames per thread (single-atom beads: the whole chunk)
    __shared__ float s_box[SBTF][8];
    __shared__ double s_mem[SBTF][3];
    if (*P.fb_flag) return;
    const int tid = threadIdx.x;
    const int64_t C3 = 3 * P.M;
    const int64_t f0 = (int64_t)blockIdx.y * SBTF, f1 = min(P.F, f0 + SBTF);
    const int nf = (int)(f1 - f0);
    const int64_t blk = f0 / SW_CHUNK;
    const int fl0 = (int)(f0 & (SW_CHUNK - 1));
    for (int r = tid; r < nf * 8; r += SB_THREADS) s_box[r >> 3][r & 7] = P.boxtab[f0 * 8 + r];
    for (int r = tid; r < nf * 3; r += SB_THREADS) s_mem[r / 3][r % 3] = P.mem_com[f0 * 3 + r];
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * SB_THREADS + tid;
    if (g >= 3 * P.N) return;
    const int64_t i = g / 3;
    const int k = (int)(g - 3 * i);
    const int p0 = P.seg[i], n = P.seg[i + 1] - p0;
    const double lm = P.lipid_mass[i];
    const double rlm = 1.0 / lm;
    const int32_t *__restrict__ gat = P.gather;
    const double *__restrict__ gm = P.mass;
    const bool um = *P.nonuniform_mass == 0;
    const double m_first = gm[0];
    float sa[NA], sb[NA];
    uint32_t mk[NA];
    constexpr bool MREG = NA <= 4;                    // small beads: the atoms' masses stay in registers
    double mj[MREG ? NA : 1];
    int ix[GATHER ? NA : 1];
    const int ix0 = 3 * p0 + k;
    bool hasc = false;                                // some atom of the bead carries an image count in this chunk
#pragma unroll
    for (int j = 0; j < NA; j++) {
        sa[j] = 0.f; sb[j] = 0.f; mk[j] = 0u;
        if (GATHER) ix[j] = 0;
        if (MREG) mj[j] = 0.0;
        if (j < n) {
            const int c = GATHER ? 3 * gat[p0 + j] + k : ix0 + 3 * j;
            if (GATHER) ix[j] = c;
            const int s0 = P.snap[blk * C3 + c], mt = P.chg_meta[blk * C3 + c];
            sa[j] = (float)s0;
            sb[j] = (float)(s0 + (mt == 1 ? 1 : 0) - (mt == 2 ? 1 : 0));
            mk[j] = P.chg_mask[blk * C3 + c] >> fl0;
            if (s0 != 0 || mt != 0) hasc = true;
            if (MREG) mj[j] = gm[GATHER ? gat[p0 + j] : p0 + j];
        }
    }
    double *oc = P.com + f0 * 3 * P.N + g, *ou = P.com_unwrap + f0 * 3 * P.N + g;
    float *oz = (k == P.norm) ? P.zc + f0 * P.N + i : nullptr;
    const float *xf = P.x + f0 * C3 + (GATHER ? 0 : ix0);
    const uint32_t sC = (uint32_t)C3, sN3 = (uint32_t)(3 * P.N), sN = (uint32_t)P.N;   // (3 M < 2^31, SBTF frames: 32-bit offsets)
    constexpr int SBU = NA == 1 ? 16 : (NA <= 4 ? 4 : 2);
    for (int fb = 0; fb < nf; fb += SBU) {
        float xr[SBU][NA];
#pragma unroll
        for (int u = 0; u < SBU; u++) {
            const float *cur = xf + (uint32_t)min(fb + u, nf - 1) * sC;
#pragma unroll
            for (int j = 0; j < NA; j++) {
                xr[u][j] = 1.0f;
                if (j < n) xr[u][j] = __ldg(cur + (GATHER ? ix[j] : 3 * j));
            }
        }
#pragma unroll
        for (int u = 0; u < SBU; u++) {
            const int f = fb + u;
            if (f < nf) {
                const float Lc = s_box[f][k];
                const double mem = s_mem[f][k];
                double cw = 0.0, cu = 0.0;
                float tiny = 3.0e38f;
#pragma unroll
                for (int j = 0; j < NA; j++) {
                    if (j < n) {
                        const float xv = xr[u][j];
                        const float cq = ((mk[j] >> f) & 1u) ? sb[j] : sa[j];
                        const double m = MREG ? mj[j] : (um ? m_first : __ldg(gm + (GATHER ? ix[j] / 3 : p0 + j)));
                        const double xd = (double)xv;
                        cw = __dadd_rn(cw, __dmul_rn(xd, m));
                        const float v = __double2float_rn(__dsub_rn((double)__fmaf_rn(cq, Lc, xv), mem));
                        cu = __dadd_rn(cu, __dmul_rn((double)v, m));
                        tiny = fminf(tiny, fabsf(xv));
                    }
                }
                if (s_box[f][6] != 0.0f || (hasc && !(tiny >= 0.00390625f))) {      // rare: the float64 unwrap
                    const bool Lbig = s_box[f][6] != 0.0f;
                    cw = 0.0; cu = 0.0;
#pragma unroll
                    for (int j = 0; j < NA; j++) {
                        if (j < n) {
                            const float cq = ((mk[j] >> f) & 1u) ? sb[j] : sa[j];
                            const double m = MREG ? mj[j] : (um ? m_first : __ldg(gm + (GATHER ? ix[j] / 3 : p0 + j)));
                            sb_accumulate(xr[u][j], (int)cq, Lc, Lbig, mem, m, cw, cu);
                        }
                    }
                }
                oc[(uint32_t)f * sN3] = pbk_div_r(cw, lm, rlm);
                const double uu = pbk_div_r(cu, lm, rlm);
                ou[(uint32_t)f * sN3] = uu;
                if (oz) oz[(uint32_t)f * sN] = (float)uu;
            }
        }
    }
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

// eight consecutive copies of a frame's row (two 128-bit loads when the rows are 32-byte multiples)
static __device__ __forceinline__ void lb_load8(const float *z, int64_t i, int64_t N, float pad, float v[LB_VPT])
{
    if ((N & 7) == 0 && i + 8 <= N) {
        const float4 a = *reinterpret_cast<const float4 *>(z + i), b = *reinterpret_cast<const float4 *>(z + i + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int e = 0; e < LB_VPT; e++) v[e] = i + e < N ? z[i + e] : pad;
    }
}

// per (slice of 2,048 lipids, frame): (sum, max |z|) of the compact normal components
__global__ void __launch_bounds__(LB_THREADS)
k_zpart(const float *__restrict__ zc, int64_t N, double *__restrict__ zpart)
{
    __shared__ double red[32];
    const int64_t f = blockIdx.y;
    const int64_t i0 = (int64_t)blockIdx.x * LB_SLICE + (int64_t)threadIdx.x * LB_VPT;
    float v[LB_VPT];
    lb_load8(zc + f * N, i0, N, 0.0f, v);
    double s = 0.0;
    float mxf = 0.0f;
#pragma unroll
    for (int e = 0; e < LB_VPT; e++) {
        s += (double)v[e];
        mxf = fmaxf(mxf, fabsf(v[e]));
    }
    double mx = (double)mxf;
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
        // bound on |mean of the copies - Welford mean of the float64 values| + the rounding of a copy
        zstat[f * 2 + 1] = 64.0 * (double)N * 2.220446049250313e-16 * mx + 1.2e-7 * mx;
        flags[f] = 0;
    }
}

// grid (slices, frames): labels of 2,048 lipids against the tree mean of frame src; a lipid within
// the rounding bound of the mean flags the frame for k_leaflets_c (sequential Welford mean)
__global__ void __launch_bounds__(LB_THREADS)
k_labels(const float *__restrict__ zc, const double *__restrict__ zstat, int64_t N, int64_t frame0, int interval,
         const uint8_t *__restrict__ prev, uint8_t *__restrict__ labels, int *flags, int32_t *status)
{
    const int64_t f = blockIdx.y;
    const int64_t g = frame0 + f;
    const int64_t src = g - (g % interval);
    uint8_t *out = labels + f * N;
    const int64_t i0 = (int64_t)blockIdx.x * LB_SLICE + (int64_t)threadIdx.x * LB_VPT;
    if (i0 >= N) return;
    if (src < frame0) {
        for (int e = 0; e < LB_VPT; e++)
            if (i0 + e < N) out[i0 + e] = prev ? prev[i0 + e] : 0;
        return;
    }
    const int64_t fs = src - frame0;
    const double mean = zstat[fs * 2], tol = zstat[fs * 2 + 1];
    float v[LB_VPT];
    lb_load8(zc + fs * N, i0, N, 0.0f, v);
    int close = 0, tie = 0;
    unsigned long long bits = 0ull;
#pragma unroll
    for (int e = 0; e < LB_VPT; e++) {
        if (i0 + e < N) {
            const double d = (double)v[e] - mean;
            if (fabs(d) <= tol) close = 1;
            if (d > 0.0) bits |= 1ull << (8 * e);
            else if (!(d < 0.0)) tie = 1;
        }
    }
    if ((N & 7) == 0) {
        *reinterpret_cast<unsigned long long *>(out + i0) = bits;
    } else {
#pragma unroll
        for (int e = 0; e < LB_VPT; e++)
            if (i0 + e < N) out[i0 + e] = (uint8_t)((bits >> (8 * e)) & 1ull);
    }
    if (close) atomicOr(flags + f, 1);
    else if (tie) atomicOr(status, PBK_ST_LEAFLET_TIE);
}

__global__ void __launch_bounds__(512)
k_leaflets_c(const double *__restrict__ com_unwrap, int norm, int64_t N, int64_t frame0, int interval, const uint8_t *__restrict__ prev,
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
    const double *z = com_unwrap + (src - frame0) * N * 3 + norm;      // z[3 i]: the float64 values themselves
    double s = 0.0, mx = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        const double v = z[3 * i];
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
        if (fabs(z[3 * i] - mean) <= tol) close = 1;
    if (close) need_exact = 1;
    __syncthreads();
    if (need_exact) {
        if (threadIdx.x == 0) {
            double m = 0.0;                       // running_stats.py:34-52
            for (int64_t i = 0; i < N; i++) {
                const double v = z[3 * i];
                m = (i == 0) ? v : __dadd_rn(m, __ddiv_rn(__dsub_rn(v, m), (double)(i + 1)));
            }
            zexact = m;
        }
        __syncthreads();
        mean = zexact;
    }
    int tie = 0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        const double v = z[3 * i];
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

// smallest / largest number of full 8-atom rows among the leaves of NumPy's pairwise split of m
// elements (pybilt_b200/pairwise.py: build_plan)
static void sw_leaf_rows(int64_t m, int *mn, int *mx)
{
    if (m <= 128) {
        const int nq = m >= 8 ? (int)(m / 8) : 0;
        if (nq < *mn) *mn = nq;
        if (nq > *mx) *mx = nq;
        return;
    }
    int64_t n2 = m / 2;
    n2 -= n2 % 8;
    sw_leaf_rows(n2, mn, mx);
    sw_leaf_rows(m - n2, mn, mx);
}

struct SweepScratch {
    float *boxtab; double *node_vals; float *u_out; float *zc; int16_t *snap; int *flags;
    uint8_t *chg_meta; uint32_t *chg_mask; int8_t *dsum; int *fb_flag;
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
    s.zc = reinterpret_cast<float *>(b + o); o += sw_align((size_t)F * N * 4);
    s.snap = reinterpret_cast<int16_t *>(b + o); o += sw_align((size_t)nch * M * 6);
    s.flags = reinterpret_cast<int *>(b + o); o += sw_align((size_t)L * 4);
    s.chg_meta = reinterpret_cast<uint8_t *>(b + o); o += sw_align((size_t)nch * M * 3);
    s.chg_mask = reinterpret_cast<uint32_t *>(b + o); o += sw_align((size_t)nch * M * 12);
    s.dsum = reinterpret_cast<int8_t *>(b + o); o += sw_align((size_t)nch * M * 3);
    s.fb_flag = reinterpret_cast<int *>(b + o); o += sw_align(8);        // [0] fall back to k_sweep_a2, [1] masses differ
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
    dim3 g((unsigned)((3 * B.N + SBS_THREADS(NA) - 1) / SBS_THREADS(NA)), (unsigned)nch);
    const size_t smem = NA == 0 ? (size_t)B.max_atoms * SBS_THREADS(NA) * 2 : 0;
    if (NA == 0) {
        cudaError_t e = cudaFuncSetAttribute(k_sweep_b<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_sweep_b<NA><<<g, SBS_THREADS(NA), smem, st>>>(B);
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
        A.chg_meta = nullptr; A.chg_mask = nullptr; A.log2_depth = 0; A.fb_flag = nullptr;
        const bool v2 = !getenv("PBK_SWEEP_V1");
        if (v2) {
            A.chg_meta = S.chg_meta; A.chg_mask = S.chg_mask;
            // time-parallel front end (needs 16-byte columns); a batch it cannot describe is flagged
            // on the device and redone by the warp-per-leaf sweep, which otherwise returns at once
            const bool tp = bulk && !getenv("PBK_SWEEP_SEQ_A");
            cudaError_t e = cudaMemsetAsync(S.fb_flag, tp ? 0 : 1, 1, st);
            if (e == cudaSuccess) e = cudaMemsetAsync(reinterpret_cast<char *>(S.fb_flag) + 1, 0, 7, st);
            if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "flag memset", e);
            if (tp) {
                SweepPParams Pp;
                Pp.x = x; Pp.boxtab = S.boxtab; Pp.prev0 = A.prev0; Pp.F = F; Pp.C3 = 3 * M; Pp.nblk = (int)nch;
                Pp.dsum = S.dsum; Pp.meta = S.chg_meta; Pp.mask = S.chg_mask; Pp.fb_flag = S.fb_flag; Pp.status = status; Pp.mass = mass; Pp.nonuniform_mass = S.fb_flag + 1;
                const int64_t ncol = (3 * M) / 4;
                k_sweep_p<<<dim3((unsigned)((ncol + SP_THREADS - 1) / SP_THREADS), (unsigned)nch), SP_THREADS, 0, st>>>(Pp);
                PBK_CHECK_LAUNCH(ctx, "k_sweep_p");
                SweepSParams Sp;
                Sp.x = x; Sp.boxtab = S.boxtab; Sp.init_images = init_images; Sp.F = F; Sp.C3 = 3 * M; Sp.nblk = (int)nch;
                Sp.net_frame = A.net_frame; Sp.dsum = S.dsum; Sp.meta = S.chg_meta; Sp.mask = S.chg_mask; Sp.snap = S.snap;
                Sp.u_out = S.u_out; Sp.net_images = net_images; Sp.fb_flag = S.fb_flag;
                k_sweep_s<<<(unsigned)((3 * M + 255) / 256), 256, 0, st>>>(Sp);
                PBK_CHECK_LAUNCH(ctx, "k_sweep_s");
                if (phase & 2) {
                    SweepLParams Lp;
                    Lp.x = x; Lp.boxtab = S.boxtab; Lp.mass = mass; Lp.leaf_start = leaf_start; Lp.leaf_len = leaf_len;
                    Lp.L = L; Lp.nodes_per_frame = L + K; Lp.F = F; Lp.C3 = 3 * M; Lp.snap = S.snap; Lp.meta = S.chg_meta;
                    Lp.mask = S.chg_mask; Lp.node_vals = S.node_vals; Lp.fb_flag = S.fb_flag; Lp.nonuniform_mass = S.fb_flag + 1; Lp.status = status;
                    const size_t lsm = (size_t)SL_WARPS * SL_DEPTH * (S2_SLOT + 8);
                    static int64_t rows_m = -1;
                    static int rows_mn = 0, rows_mx = 0;
                    if (rows_m != M) { int mn = 1 << 30, mx = 0; sw_leaf_rows(M, &mn, &mx); rows_mn = mn; rows_mx = mx; rows_m = M; }
                    const int nq_tight = (rows_mn >= rows_mx - 1 && rows_mx >= 2) ? rows_mx : 0;
                    void (*kl)(SweepLParams) = k_sweep_l<16, false>;
                    if (nq_tight == 16) kl = k_sweep_l<16, true>;
                    else if (nq_tight == 12) kl = k_sweep_l<12, true>;
                    else if (nq_tight == 8) kl = k_sweep_l<8, true>;
                    cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm);
                    cudaFuncSetAttribute(kl, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
                    kl<<<dim3((unsigned)((L + SL_WARPS - 1) / SL_WARPS), (unsigned)((F + SL_FRAMES - 1) / SL_FRAMES)),
                         SL_WARPS * 32, lsm, st>>>(Lp);
                    PBK_CHECK_LAUNCH(ctx, "k_sweep_l");
                }
            }
            A.fb_flag = S.fb_flag;
            // one warp per leaf; ring depth from the shared memory an SM can give the leaves it hosts
            const bool bulk2 = bulk;
            const int64_t per_sm_leaves = (L + ctx->sm_count - 1) / ctx->sm_count;
            int lg = 4;
            while (lg > 1 && per_sm_leaves * ((int64_t)(1 << lg) * 1544 + 3808 + 1024) > 200 * 1024) lg--;
            A.log2_depth = lg;
            const size_t smem = (size_t)(1 << lg) * 1544 + 3808;
            void (*k2)(SweepAParams) = bulk2 ? k_sweep_a2<true> : k_sweep_a2<false>;
            cudaFuncSetAttribute(k2, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            k2<<<(unsigned)L, 32, smem, st>>>(A);
            PBK_CHECK_LAUNCH(ctx, "k_sweep_a2");
        } else {
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
        const bool v2b = !getenv("PBK_SWEEP_V1") && !getenv("PBK_SWEEP_SEQ_B");
        B.chg_meta = v2b ? S.chg_meta : nullptr; B.chg_mask = v2b ? S.chg_mask : nullptr;
        cudaError_t e = cudaSuccess;
        const bool bt = v2b && !getenv("PBK_SWEEP_SEQ_A") && bulk && max_bead_atoms <= 16 && 3 * M < ((int64_t)1 << 27);
        B.fb_flag = bt ? S.fb_flag : nullptr; B.nonuniform_mass = S.fb_flag + 1;
        if (bt) {
            const int64_t sbtf = max_bead_atoms == 1 ? 2 * SBT_FRAMES : SBT_FRAMES;
            dim3 g((unsigned)((3 * N + SB_THREADS - 1) / SB_THREADS), (unsigned)((F + sbtf - 1) / sbtf));
#define SBT_LAUNCH(NA) do { if (gather) k_sweep_bt<NA, true><<<g, SB_THREADS, 0, st>>>(B); \
                            else k_sweep_bt<NA, false><<<g, SB_THREADS, 0, st>>>(B); } while (0)
            if (max_bead_atoms == 1) SBT_LAUNCH(1);
            else if (max_bead_atoms <= 4) SBT_LAUNCH(4);
            else if (max_bead_atoms <= 8) SBT_LAUNCH(8);
            else if (max_bead_atoms <= 12) SBT_LAUNCH(12);
            else SBT_LAUNCH(16);
#undef SBT_LAUNCH
            PBK_CHECK_LAUNCH(ctx, "k_sweep_bt");
        }
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
        k_leaflets_c<<<(unsigned)F, 512, 0, st>>>(com_unwrap, norm, N, frame0, update_interval, prev_labels, labels,
                                                  F <= 65535 ? S.lflags : nullptr, status);
        PBK_CHECK_LAUNCH(ctx, "k_leaflets_c");
    }
    cudaError_t e = cudaMemcpyAsync(u_state, S.u_out, (size_t)M * 12, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "u_state copy", e);
    return PBK_OK;
}
