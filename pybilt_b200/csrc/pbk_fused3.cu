// pbk_fused3.cu -- fused frame reduction for small systems, second generation:
// unwrap + membrane COM + per-lipid COMs + leaflet labels out of shared memory, one frame
// at a time per SM, frames arriving by TMA bulk copy one full frame ahead.
//
// Reference arithmetic restated (paths relative to the PyBILT tree), identical to pbk_com.cu:
//   unwrap        pybilt/mda_tools/mda_unwrap_vectorized.py:12-93
//   membrane COM  pybilt/bilayer_analyzer/com_frame.py:172-181 (NumPy pairwise order)
//   lipid COM     pybilt/bilayer_analyzer/com_frame.py:43-96,163-185
//   leaflets      pybilt/bilayer_analyzer/bilayer_analyzer.py:826-847
//
// Three kernels, as in pbk_fused.cu (which stays as the fallback for shapes this file does
// not take: atom counts that are not a multiple of 4, more than 80 summation leaves):
//   k_img3     streams the positions once (one thread per float4 column of a frame chunk, 4
//              frames of loads in flight), sums the image-count increments per chunk and lists
//              the moves of the counts for k_frames4;
//   k_scan3    exclusive prefix over the chunks -> int8 image counts at every chunk start,
//              plus the slack check that makes the chunk-local increments exact;
//   k_frames3  one 512-thread CTA per chunk.  Per frame:
//                leaf phase   6 lanes per <=128-atom NumPy summation leaf; a lane owns 4 of
//                             the leaf's 24 (accumulator, axis) chains and walks them with one
//                             128-bit shared load per 8-atom row.  The previous frame lives in
//                             registers, so the same-image test costs no shared traffic and the
//                             other frame buffer is free for the TMA prefetch at once.
//                tree         warp 0 combines the leaves (table driven) while the other warps
//                             write the leaflet labels of the PREVIOUS frame;
//                lipid phase  one thread per lipid, atoms in order, 128-bit shared loads when
//                             every bead is a 4-aligned run of atoms, both COMs at once.
#include <stdlib.h>

#include "pbk_common.cuh"

#define F3_THREADS 512
#define F3_WARPS (F3_THREADS / 32)
#define F3_LPW 4                    // leaves per warp in the leaf phase (8 lanes each, 6 of them active)
#define F3_NQ 16                    // 8-atom rows per leaf (128 / 8)
#define F3_TAB 1024                 // ints of summation-plan tables staged in shared memory
#define F4_TAB 384                  // the same for k_frames4 (two CTAs per SM: every KB counts)
#define F4_EVCAP 4096               // count moves per chunk that k_frames4 takes (more: the batch is handed back)
#define F4_MAXCHUNK 1024            // frames per chunk that k_evsort buckets
#define F3_PIECES 4                 // bulk copies per frame (all complete on one mbarrier)

int pbk_reduce_fused_v1(pbk_ctx *ctx, const float *x, const float *box, const double *mass,
                        const int32_t *seg, const int32_t *gather, const double *lipid_mass,
                        const int32_t *leaf_start, const int32_t *leaf_len, int32_t L,
                        const int32_t *node_left, const int32_t *node_right, int32_t K,
                        const int32_t *level_off, int32_t H, double total_mass, float *u_state,
                        int first_frame, int64_t F, int64_t M, int64_t N, int norm, int do_labels,
                        int phase, const int32_t *init_images, int32_t *net_images, void *scratch,
                        int64_t scratch_bytes, double *com, double *com_unwrap, double *mem_com,
                        uint8_t *labels, int32_t *status, void *stream);

// ------------------------------------------------------------------------------------------
// image shift of the reference rule for difference d (float32) and box length L; also returns
// the distance of the shifted difference from the nearer window boundary +-L/2
// ------------------------------------------------------------------------------------------
static __device__ __forceinline__ int f3_shift(float d, float L, float *slack, int *bad)
{
    const float h = L * 0.5f;
    int s = 0;
    if (!(L > 0.0f)) { *bad |= 1; *slack = 0.f; return 0; }
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
    if (s <= -32768 || s >= 32768) { *bad |= 1; s = 0; }
    return s;
}

static __device__ __noinline__ int f3_shift_far(float d, float L, float *slack, int *bad)
{
    return f3_shift(d, L, slack, bad);
}

// ------------------------------------------------------------------------------------------
// k_img3: image-count increments per (chunk, chain).  One thread per float4 column.
// ------------------------------------------------------------------------------------------
struct __align__(16) ImgCol {
    short acc[4];       // sum of the increments of the column's 4 chains over the chunk
    float smin;         // smallest distance of a shifted jump from +-L/2 (over the 4 chains)
    int amax;           // largest |running sum| (over the 4 chains)
};

#define IMG_THREADS 128
#define IMG_UNR 4                   // frames of loads in flight per thread (8 spilled at 64 registers and was slower)

// one frame of one column: same-image test of the 4 chains at once; the exact rule only when
// one of them fails (rare)
// (a count move of chain c at local frame tl is appended to the chunk's event list inside img3_step: k_frames4
//  reads it)
static __device__ __forceinline__ void img3_step(const float4 v, float pv[4], const float he[4], const float Le[4],
                                                 int acc[4], int &am, float &smin, int &bad, uint2 *ev, int *ev_n,
                                                 int c0, int tl)
{
    const float xv[4] = {v.x, v.y, v.z, v.w};
    float sl[4];
#pragma unroll
    for (int e = 0; e < 4; e++) sl[e] = he[e] - fabsf(__fsub_rn(xv[e], pv[e]));
    const float m = fminf(fminf(sl[0], sl[1]), fminf(sl[2], sl[3]));
    if (!(m > 0.0f) || !(sl[0] > 0.0f) || !(sl[1] > 0.0f) || !(sl[2] > 0.0f) || !(sl[3] > 0.0f)) {   // (the second form catches NaN)
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (!(sl[e] > 0.0f)) {                          // |jump| >= L/2 (or bad box)
                float s2;
                int ds;
                const float d = __fsub_rn(xv[e], pv[e]), h = Le[e] * 0.5f;
                if (fabsf(d) > h && fabsf(d) < 3.0f * h) {
                    // the usual crossing: one image (the loop of the reference rule stops after its first step,
                    // d -+ L lies inside the window); same float64 expressions as f3_shift, no call
                    ds = d > 0.0f ? -1 : 1;
                    s2 = (float)((double)h - fabs(__dadd_rn((double)d, __dmul_rn((double)ds, (double)Le[e]))));
                } else {
                    ds = f3_shift_far(d, Le[e], &s2, &bad);
                    if (ds > 127 || ds < -127) bad |= 2;
                }
                if (ev && ds != 0) {
                    const int slot = atomicAdd(ev_n, 1);
                    if (slot < F4_EVCAP) ev[slot] = make_uint2((uint32_t)(c0 + e), ((uint32_t)tl << 8) | ((uint32_t)ds & 0xffu));
                }
                acc[e] += ds;
                am = max(am, abs(acc[e]));
                smin = fminf(smin, s2);
            } else {
                smin = fminf(smin, sl[e]);
            }
        }
    } else {
        smin = fminf(smin, m);
    }
#pragma unroll
    for (int e = 0; e < 4; e++) pv[e] = xv[e];
}

// Frames of chunk kc.  Chunks of `chunk` frames tile [0, split) and then, starting again at
// `split`, [split, F): a frame shard asks for the image counts after its owned frames (`split`),
// which must be a chunk boundary; split == F is the plain uniform tiling.
static __device__ __forceinline__ void f3_chunk_range(int kc, int chunk, int64_t split, int64_t F, int64_t *f0,
                                                      int64_t *f1)
{
    const int n1 = (int)((split + chunk - 1) / chunk);
    if (kc < n1) {
        *f0 = (int64_t)kc * chunk;
        *f1 = min(split, *f0 + (int64_t)chunk);
    } else {
        *f0 = split + (int64_t)(kc - n1) * chunk;
        *f1 = min(F, *f0 + (int64_t)chunk);
    }
}

// per chunk: largest box edge, largest frame-to-frame change of an edge (for the slack check of
// k_scan3) and whether the box is the same in every frame of the chunk (NVT runs)
__global__ void __launch_bounds__(128)
k_chunk_box(const float *__restrict__ box, int64_t F, int64_t split, int chunk, float *__restrict__ chunk_box,
            int *__restrict__ chunk_const, int *__restrict__ ev_count)
{
    __shared__ int s_red[3];
    const int kc = blockIdx.x;
    int64_t f0, f1;
    f3_chunk_range(kc, chunk, split, F, &f0, &f1);
    const int nf = (int)(f1 - f0);
    if (threadIdx.x < 3) s_red[threadIdx.x] = 0;
    __syncthreads();
    float Lmax = 0.f, dL = 0.f;
    int diff = 0;
    for (int r = threadIdx.x; r < nf * 3; r += 128) {
        const int64_t t = f0 + r / 3;
        const int k = r % 3;
        const float L = box[t * 3 + k];
        Lmax = fmaxf(Lmax, L);
        if (t > 0) dL = fmaxf(dL, fabsf(L - box[(t - 1) * 3 + k]));
        diff |= (L != box[f0 * 3 + k]);
    }
    atomicMax(&s_red[0], __float_as_int(Lmax));      // (non-negative floats order like their bits; NaN/negative
    atomicMax(&s_red[1], __float_as_int(dL));        //  boxes are caught by the L > 0 tests of k_img3)
    if (diff) atomicOr(&s_red[2], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        chunk_box[kc * 2] = __int_as_float(s_red[0]);
        chunk_box[kc * 2 + 1] = __int_as_float(s_red[1]);
        chunk_const[kc] = s_red[2] == 0;
        if (ev_count) ev_count[kc] = 0;              // the chunk's event list starts empty (k_img3 appends)
    }
}

__global__ void __launch_bounds__(IMG_THREADS, 8)
k_img3(const float4 *__restrict__ x4, const float *__restrict__ box, const float4 *__restrict__ u4,
       int first_frame, int64_t F, int64_t split, int C3q, int chunk, ImgCol *__restrict__ cols,
       const int *__restrict__ chunk_const, int32_t *status, uint2 *__restrict__ events, int *__restrict__ ev_count,
       int8_t *__restrict__ first_shift, int n_chunks, int span)
{
    // a thread walks `span` consecutive chunks (they are contiguous in time, also across `split`): the last
    // frame of one chunk is the previous frame of the next, and the per-thread set-up is paid once
    const int q4 = blockIdx.x * IMG_THREADS + threadIdx.x;      // column blocks vary fastest: neighbouring CTAs read one frame
    if (q4 >= C3q) return;
    const int k0 = (4 * q4) % 3;                 // axis of the column's first chain
    int bad = 0;
    float pv[4] = {0.f, 0.f, 0.f, 0.f};
    for (int sub = 0; sub < span; sub++) {
    const int kc = blockIdx.y * span + sub;
    if (kc >= n_chunks) break;
    int64_t f0, f1;
    f3_chunk_range(kc, chunk, split, F, &f0, &f1);
    const bool cbox = chunk_const[kc] != 0;
    uint2 *ev = events ? events + (int64_t)kc * F4_EVCAP : nullptr;
    int *ev_n = ev_count ? ev_count + kc : nullptr;
    int acc[4] = {0, 0, 0, 0};
    int am = 0;
    float smin = 3.0e38f;
    int64_t t = f0;
    if (kc == 0) {
        const float4 p = x4[q4];
        pv[0] = p.x; pv[1] = p.y; pv[2] = p.z; pv[3] = p.w;
        if (!first_frame) {                      // exact rule against the carried unwrapped state
            const float4 u = u4[q4];
            const float uv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float sl;
                acc[e] = f3_shift(__fsub_rn(pv[e], uv[e]), box[(k0 + e) % 3], &sl, &bad);
                // (k_frames4 starts chunk 0 from these counts: they are the whole state, not a few moves)
                if (first_shift) {
                    if (acc[e] > 127 || acc[e] < -127) bad |= 2;
                    first_shift[4 * q4 + e] = (int8_t)max(-127, min(127, acc[e]));
                }
                am = max(am, abs(acc[e]));
            }
        }
        t = 1;
    } else if (sub == 0) {
        const float4 p = x4[(f0 - 1) * C3q + q4];
        pv[0] = p.x; pv[1] = p.y; pv[2] = p.z; pv[3] = p.w;
    }
    const float4 *ptr = x4 + t * C3q + q4;
    if (cbox) {
        // constant box over the chunk (NVT runs): half edges in registers
        float he[4], Le[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            Le[e] = box[f0 * 3 + (k0 + e) % 3];
            he[e] = Le[e] > 0.0f ? Le[e] * 0.5f : -1.0f;    // bad box: every test fails
        }
        for (; t + IMG_UNR <= f1; t += IMG_UNR) {
            float4 v[IMG_UNR];
#pragma unroll
            for (int j = 0; j < IMG_UNR; j++) v[j] = __ldg(ptr + (int64_t)j * C3q);
            ptr += (int64_t)IMG_UNR * C3q;
#pragma unroll
            for (int j = 0; j < IMG_UNR; j++) img3_step(v[j], pv, he, Le, acc, am, smin, bad, ev, ev_n, 4 * q4, (int)(t + j - f0));
        }
        for (; t < f1; t++) {
            const float4 v = __ldg(ptr);
            ptr += C3q;
            img3_step(v, pv, he, Le, acc, am, smin, bad, ev, ev_n, 4 * q4, (int)(t - f0));
        }
    } else {
        for (; t < f1; t += IMG_UNR) {
            float4 v[IMG_UNR];
#pragma unroll
            for (int j = 0; j < IMG_UNR; j++) v[j] = __ldg(ptr + (int64_t)min((int64_t)j, f1 - 1 - t) * C3q);
            ptr += (int64_t)IMG_UNR * C3q;
#pragma unroll
            for (int j = 0; j < IMG_UNR; j++) {
                if (t + j < f1) {
                    float he[4], Le[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int k = (k0 + e) % 3;
                        Le[e] = __ldg(box + (t + j) * 3 + k);
                        he[e] = Le[e] > 0.0f ? Le[e] * 0.5f : -1.0f;
                    }
                    img3_step(v[j], pv, he, Le, acc, am, smin, bad, ev, ev_n, 4 * q4, (int)(t + j - f0));
                }
            }
        }
    }
    ImgCol c;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        if (acc[e] > 32767 || acc[e] < -32767) { bad |= 1; acc[e] = 0; }
        c.acc[e] = (short)acc[e];
    }
    c.smin = smin;
    c.amax = am;
    cols[(int64_t)kc * C3q + q4] = c;
    if ((bad & 2) && !ev) bad &= ~2;
    }
    if (bad & 1) atomicOr(status, PBK_ST_IMAGE_OVERFLOW);
    if (bad & 2) atomicOr(status, PBK_ST_FUSED_RANGE);     // a move of more than 127 images in one frame (k_frames4 only)
}

// k_scan3: per chain, exclusive prefix of the chunk totals (= image count at every chunk start,
// int8) and the slack check.  One warp per float4 column, one lane per chunk (rounds of 32
// chunks, all loads of a column issued before the first shuffle): a 5-step warp scan per
// chain instead of a serial walk over the chunks.
#define SCAN_ROUNDS 8
__global__ void __launch_bounds__(128)
k_scan3(const ImgCol *__restrict__ cols, const float *__restrict__ chunk_box, int C3, int n_chunks,
        const int32_t *__restrict__ init_images, int8_t *__restrict__ sstart, int32_t *status)
{
    const int lane = threadIdx.x & 31;
    const int q4 = blockIdx.x * 4 + (threadIdx.x >> 5), C3q = C3 >> 2;
    if (q4 >= C3q) return;
    int carry[4], flags = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) carry[e] = init_images ? init_images[4 * q4 + e] : 0;
    for (int k0 = 0; k0 < n_chunks; k0 += 32 * SCAN_ROUNDS) {
        int4 raw[SCAN_ROUNDS];
#pragma unroll
        for (int r = 0; r < SCAN_ROUNDS; r++) {
            const int kc = k0 + r * 32 + lane;
            raw[r] = make_int4(0, 0, 0x7f7fffff, 0);
            if (kc < n_chunks) raw[r] = __ldg(reinterpret_cast<const int4 *>(cols + (int64_t)kc * C3q + q4));
        }
#pragma unroll
        for (int r = 0; r < SCAN_ROUNDS; r++) {
            if (k0 + r * 32 >= n_chunks) break;
            const int kc = k0 + r * 32 + lane;
            int a[4] = {(int)(short)(raw[r].x & 0xffff), (int)(short)(raw[r].x >> 16),
                        (int)(short)(raw[r].y & 0xffff), (int)(short)(raw[r].y >> 16)};
            int st[4], worst = 0;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                int inc = a[e];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                st[e] = carry[e] + inc - a[e];                    // image count at the start of chunk kc
                carry[e] += __shfl_sync(0xffffffffu, inc, 31);
                worst = max(worst, abs(st[e]));
            }
            if (kc < n_chunks) {
                uint32_t w = 0;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    int sb = st[e];
                    if (sb > 127 || sb < -127) { flags |= PBK_ST_FUSED_RANGE; sb = 0; }
                    w |= ((uint32_t)sb & 0xffu) << (8 * e);
                }
                *reinterpret_cast<uint32_t *>(sstart + (int64_t)kc * C3 + 4 * q4) = w;
                const float sl = __int_as_float(raw[r].z);
                const int smax = worst + raw[r].w + 1;
                const float Lmax = chunk_box[kc * 2], dL = chunk_box[kc * 2 + 1];
                // |true shifted difference - delta-form one| <= ulp32(u)/2 + ulp32(d)/2 + ulp32(d0)/2
                //                                              + |s| * |L_t - L_{t-1}|
                const float eps = (float)(smax + 2) * Lmax * 2.3841858e-7f + (float)smax * dL;
                if (!(sl > eps)) flags |= PBK_ST_UNWRAP_AMBIGUOUS;
                if (smax > 126) flags |= PBK_ST_FUSED_RANGE;
            }
        }
    }
    if (flags) atomicOr(status, flags);
}

// k_evsort: one CTA per chunk -- the chunk's count moves bucketed by frame (counting sort):
// sorted[off[tl] .. off[tl + 1]) are the moves of local frame tl.  Too many moves, or a chunk longer
// than the buckets: the batch is handed back (PBK_ST_FUSED_RANGE) and the generic kernels take it.
__global__ void __launch_bounds__(256)
k_evsort(const uint2 *__restrict__ events, const int *__restrict__ ev_count, int chunk, uint2 *__restrict__ sorted,
         int *__restrict__ off, int32_t *status)
{
    __shared__ int s_cnt[F4_MAXCHUNK + 1];
    const int kc = blockIdx.x, tid = threadIdx.x;
    int n = ev_count[kc];
    int *o = off + (int64_t)kc * (chunk + 1);
    if (n > F4_EVCAP || chunk > F4_MAXCHUNK) {
        if (tid == 0) atomicOr(status, PBK_ST_FUSED_RANGE);
        for (int t = tid; t <= chunk; t += 256) o[t] = 0;
        return;
    }
    for (int t = tid; t <= chunk; t += 256) s_cnt[t] = 0;
    __syncthreads();
    const uint2 *ev = events + (int64_t)kc * F4_EVCAP;
    for (int j = tid; j < n; j += 256) atomicAdd(&s_cnt[min((int)(ev[j].y >> 8), chunk - 1) + 1], 1);
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int t = 0; t <= chunk; t++) { run += s_cnt[t]; s_cnt[t] = run; o[t] = run; }   // s_cnt[t] = start of bucket t
    }
    __syncthreads();
    // (the order inside a bucket does not matter: a chain moves at most once per frame)
    uint2 *so = sorted + (int64_t)kc * F4_EVCAP;
    for (int j = tid; j < n; j += 256) {
        const uint2 e = ev[j];
        const int tl = min((int)(e.y >> 8), chunk - 1);
        so[atomicAdd(&s_cnt[tl], 1)] = e;
    }
}

// net image-count change over the whole batch: what a frame shard hands to the ranks after it
// (one warp per float4 column, lanes over the chunks: the serial walk over ~300 chunks per chain took 25 us)
__global__ void __launch_bounds__(128)
k_net3(const ImgCol *__restrict__ cols, int C3, int n_chunks, int32_t *__restrict__ net)
{
    const int lane = threadIdx.x & 31, C3q = C3 >> 2;
    const int q4 = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (q4 >= C3q) return;
    int s[4] = {0, 0, 0, 0};
    for (int kc = lane; kc < n_chunks; kc += 32) {
        const int4 raw = __ldg(reinterpret_cast<const int4 *>(cols + (int64_t)kc * C3q + q4));
        s[0] += (int)(short)(raw.x & 0xffff); s[1] += (int)(short)(raw.x >> 16);
        s[2] += (int)(short)(raw.y & 0xffff); s[3] += (int)(short)(raw.y >> 16);
    }
#pragma unroll
    for (int e = 0; e < 4; e++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[e] += __shfl_xor_sync(0xffffffffu, s[e], o);
    }
    if (lane == 0) *reinterpret_cast<int4 *>(net + 4 * q4) = make_int4(s[0], s[1], s[2], s[3]);
}

// ------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers (PTX ISA: mbarrier, cp.async.bulk)
// ------------------------------------------------------------------------------------------
static __device__ __forceinline__ uint32_t f3_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

static __device__ __forceinline__ void f3_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(f3_smem_u32(bar)), "r"(count) : "memory");
}
static __device__ __forceinline__ void f3_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(f3_smem_u32(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void f3_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(f3_smem_u32(dst)), "l"(src), "r"(bytes), "r"(f3_smem_u32(bar)) : "memory");
}
// bounded wait: returns false if the phase never completed (reported, never spins forever)
static __device__ __forceinline__ bool f3_mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = f3_smem_u32(bar);
    for (int it = 0; it < (1 << 22); it++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

struct F3Params {
    const float *x; const float *box; const double *mass; const int32_t *seg; const int32_t *gather;
    const double *lipid_mass;
    const int32_t *leaf_start, *leaf_len, *node_left, *node_right, *level_off;
    int L, K, H;
    int perfect;                // log2(L) when the summation tree is a perfect binary tree, else -1
    double total_mass;
    const float *u_state; float *u_out; int first_frame;
    int64_t F, split; int M, N, norm, chunk, n_chunks, do_labels;
    const int8_t *sstart;       // [n_chunks][3M] image counts at the chunk starts
    const int32_t *init_images;
    double *com, *com_unwrap, *mem_com; uint8_t *labels; int32_t *status;
    long long *prof;            // optional [n_chunks][8] phase cycle counters (debug), may be NULL
    const uint2 *ev_sorted; const int *ev_off; int ev_cap;   // k_frames4: count moves per chunk, sorted by frame
    const int8_t *first_shift;   // k_frames4, later batches: counts of frame 0 against the carried state
};

// Image counts live in shared memory as BIASED bytes (s + 128): byte e of a word drops straight
// into the mantissa of 2^23 * 1.5 with one PRMT, and (that - (12582912 + 128)) is (float)s.
#define F3_ZERO4 0x80808080u
static __device__ __forceinline__ float f3_sfloat(uint32_t w, int e)
{
    return __int_as_float((int)__byte_perm(w, 0x4B400000u, 0x7650u + (uint32_t)e)) - 12583040.0f;
}
static __device__ __forceinline__ int f3_sint(uint32_t w, int e) { return (int)((w >> (8 * e)) & 0xffu) - 128; }

// u = f32(f64(x) + s*L) outside the range where one float32 FMA is exact (|x| < 2^-8 or L >= 2^13)
static __device__ __noinline__ float f3_unwrap_slow(float x, int s, float L)
{
    if (s == 0) return x;
    return __double2float_rn(__dadd_rn((double)x, __dmul_rn((double)s, (double)L)));
}

// Exact reference rule for one float4 column whose same-image test failed: new biased image
// counts in the low word, PBK_ST_* style `bad` bits (1 overflow / bad box, 4 outside int8) above.
static __device__ __noinline__ unsigned long long
f3_column_exact(float x0, float x1, float x2, float x3, float p0, float p1, float p2, float p3, uint32_t sw,
                float Lc0, float Lc1, float Lc2, float Lp0, float Lp1, float Lp2, int ke0, int nvalid)
{
    const float xs[4] = {x0, x1, x2, x3}, ps[4] = {p0, p1, p2, p3};
    int bad = 0;
    uint32_t nw = sw;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        if (e < nvalid) {
            const int k = (ke0 + e) % 3;
            const float Lc = k == 0 ? Lc0 : (k == 1 ? Lc1 : Lc2), Lp = k == 0 ? Lp0 : (k == 1 ? Lp1 : Lp2);
            float sl;
            int sn = f3_shift(__fsub_rn(xs[e], pbk_unwrapped(ps[e], f3_sint(sw, e), Lp)), Lc, &sl, &bad);
            if (sn > 127 || sn < -127) { bad |= 4; sn = 0; }
            nw = (nw & ~(0xffu << (8 * e))) | ((uint32_t)(sn + 128) << (8 * e));
        }
    }
    return (unsigned long long)nw | ((unsigned long long)(uint32_t)bad << 32);
}

// per-axis count of non-zero image counts (shared): lets the lipid phase skip f64(u) on axes
// where u == x for every atom
static __device__ __forceinline__ void f3_nnz_update(uint32_t oldw, uint32_t neww, int ke0, int *s_nnz)
{
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const bool z0 = ((oldw >> (8 * e)) & 0xffu) == 0x80u, z1 = ((neww >> (8 * e)) & 0xffu) == 0x80u;
        if (z0 != z1) atomicAdd(&s_nnz[(ke0 + e) % 3], z1 ? -1 : 1);
    }
}

// unwrapped float32 coordinates of one column (biased counts sw != ZERO4)
static __device__ __forceinline__ void f3_unwrap4(const float xv[4], uint32_t sw, const float Le[4], bool Lbig, float uv[4])
{
#pragma unroll
    for (int e = 0; e < 4; e++) uv[e] = __fmaf_rn(f3_sfloat(sw, e), Le[e], xv[e]);
    const float tiny = fminf(fminf(fabsf(xv[0]), fabsf(xv[1])), fminf(fabsf(xv[2]), fabsf(xv[3])));
    if (tiny < 0.00390625f || Lbig) {
#pragma unroll
        for (int e = 0; e < 4; e++) uv[e] = f3_unwrap_slow(xv[e], f3_sint(sw, e), Le[e]);
    }
}

// One atom component of the lipid phase: adds m*f64(x) to the wrapped sum and
// m*f64(f32(f64(u) - mem)) to the unwrapped one (com_frame.py:71-96,180).
static __device__ __forceinline__ void f3_comp0(float xv, double mem, double m, double &cw, double &cu)
{
    const double xd = (double)xv;
    cw = __dadd_rn(cw, __dmul_rn(xd, m));
    const float v = __double2float_rn(__dsub_rn(xd, mem));
    cu = __dadd_rn(cu, __dmul_rn((double)v, m));
}
static __device__ __forceinline__ void f3_comp1(float xv, float uv, double mem, double m, double &cw, double &cu)
{
    cw = __dadd_rn(cw, __dmul_rn((double)xv, m));
    const float v = __double2float_rn(__dsub_rn((double)uv, mem));
    cu = __dadd_rn(cu, __dmul_rn((double)v, m));
}

// Four consecutive atoms of a bead (three float4 of positions, three words of image counts).
// ZM: bit k set = no atom carries an image count on axis k (u == x there).
template <int ZM>
static __device__ __forceinline__ void
f3_lipid_group(const float4 *x4, const uint32_t *s4, double ma, double mb, double mc, double md, float Lc0,
               float Lc1, float Lc2, bool Lbig, double m0, double m1, double m2, double &cw0, double &cw1,
               double &cw2, double &cu0, double &cu1, double &cu2)
{
    const float4 A = x4[0], B = x4[1], C = x4[2];
    const float xs[12] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, C.x, C.y, C.z, C.w};
    float us[12];
#pragma unroll
    for (int c = 0; c < 12; c++) us[c] = xs[c];
    if (ZM != 7) {
        const uint32_t sw[3] = {s4[0], s4[1], s4[2]};
        const float Ls[3] = {Lc0, Lc1, Lc2};
        float tiny = 1.0f;
#pragma unroll
        for (int c = 0; c < 12; c++) {
            if (!((ZM >> (c % 3)) & 1)) {
                us[c] = __fmaf_rn(f3_sfloat(sw[c / 4], c % 4), Ls[c % 3], xs[c]);
                tiny = fminf(tiny, fabsf(xs[c]));
            }
        }
        if (tiny < 0.00390625f || Lbig) {
#pragma unroll
            for (int c = 0; c < 12; c++)
                if (!((ZM >> (c % 3)) & 1)) us[c] = f3_unwrap_slow(xs[c], f3_sint(sw[c / 4], c % 4), Ls[c % 3]);
        }
    }
    const double ms[4] = {ma, mb, mc, md};
    const double mems[3] = {m0, m1, m2};
    double cw[3] = {cw0, cw1, cw2}, cu[3] = {cu0, cu1, cu2};
#pragma unroll
    for (int c = 0; c < 12; c++) {
        const int k = c % 3;
        if ((ZM >> k) & 1) f3_comp0(xs[c], mems[k], ms[c / 3], cw[k], cu[k]);
        else f3_comp1(xs[c], us[c], mems[k], ms[c / 3], cw[k], cu[k]);
    }
    cw0 = cw[0]; cw1 = cw[1]; cw2 = cw[2]; cu0 = cu[0]; cu1 = cu[1]; cu2 = cu[2];
}

// FULL: every leaf is exactly NQ rows of 8 atoms (perfect tree, no tail row)
template <int NQ, bool FULL>
__global__ void __launch_bounds__(F3_THREADS, 1)
k_frames3(F3Params P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C3 = 3 * P.M;                       // multiple of 12
    // doubles first, then the two frame buffers (128-byte aligned), then the image counts
    double *s_acc = reinterpret_cast<double *>(smem_raw);               // L*24 accumulators
    double *s_leaf = s_acc + (size_t)P.L * 24;                          // (L+K)*3
    double *s_z = s_leaf + (size_t)(P.L + P.K) * 3;                     // N   com_unwrap[norm]
    double *s_tail = s_z + P.N;                                         // 24
    size_t off = ((size_t)(reinterpret_cast<unsigned char *>(s_tail + 24) - smem_raw) + 127) & ~(size_t)127;
    float *XB0 = reinterpret_cast<float *>(smem_raw + off);
    float *XB1 = XB0 + C3;
    uint32_t *S32 = reinterpret_cast<uint32_t *>(XB1 + C3);             // C3 biased int8 image counts
    unsigned char *S8 = reinterpret_cast<unsigned char *>(S32);
    __shared__ int s_tab[F3_TAB];               // leaf_start | leaf_len | node_left | node_right | level_off
    __shared__ uint64_t bar[2];
    __shared__ double s_mem[3];
    __shared__ double red[F3_WARPS];
    __shared__ double s_pmax[F3_WARPS];
    __shared__ int s_nnz[3];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kc = blockIdx.x;
    int64_t f0, f1;
    f3_chunk_range(kc, P.chunk, P.split, P.F, &f0, &f1);
    if (f0 >= f1) return;
    const uint32_t frame_bytes = (uint32_t)C3 * 4u;
    int bad = 0;

    if (tid == 0) { f3_mbar_init(&bar[0], 1); f3_mbar_init(&bar[1], 1); }
    if (tid < 3) s_nnz[tid] = 0;
    __syncthreads();
    int *t_ls = s_tab, *t_ll = t_ls + P.L, *t_nl = t_ll + P.L, *t_nr = t_nl + P.K, *t_lo = t_nr + P.K;
    for (int i = tid; i < P.L; i += F3_THREADS) { t_ls[i] = P.leaf_start[i]; t_ll[i] = P.leaf_len[i]; }
    for (int i = tid; i < P.K; i += F3_THREADS) { t_nl[i] = P.node_left[i]; t_nr[i] = P.node_right[i]; }
    for (int i = tid; i <= P.H; i += F3_THREADS) t_lo[i] = P.level_off[i];
    // uniform masses (coarse-grained models): the mass is a register, not a load
    int uni = 1;
    const double m_first = P.mass[0];
    for (int a = tid; a < P.M; a += F3_THREADS) if (P.mass[a] != m_first) uni = 0;
    // vector lipid phase: contiguous beads, every start and length a multiple of 4 atoms
    int vec = P.gather ? 0 : 1;
    for (int i = tid; i < P.N; i += F3_THREADS) {
        const int a0 = P.seg[i], a1 = P.seg[i + 1];
        if ((a0 & 3) || ((a1 - a0) & 3)) vec = 0;
    }
    // image counts at the chunk start; "previous" frame
    //   chunk 0, first batch : prev = x_0 itself, counts 0 / init_images  (u_0 = x_0 + s*L)
    //   chunk 0, later batch : prev = carried u_state, counts 0 (u_prev = prev + 0*L)
    //   chunk k > 0          : prev = raw x_{f0-1}, counts from the scan
    const float *prev_src = (kc == 0) ? (P.first_frame ? P.x : P.u_state) : P.x + (f0 - 1) * (int64_t)C3;
    for (int c = tid; c < C3; c += F3_THREADS) {
        int s0 = (kc == 0) ? (P.init_images ? P.init_images[c] : 0) : (int)P.sstart[(int64_t)kc * C3 + c];
        if (s0 > 127 || s0 < -127) { bad |= 4; s0 = 0; }
        S8[c] = (unsigned char)(s0 + 128);
        if (s0 != 0) atomicAdd(&s_nnz[c % 3], 1);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const bool um = __syncthreads_and(uni) != 0;        // also publishes tables, S, the barriers
    const bool vec_beads = __syncthreads_and(vec) != 0;
    const double um_val = m_first;
    const double *__restrict__ gmass = P.mass;
#define F3_MASS(a) (um ? um_val : __ldg(gmass + (a)))

    // ---- leaf-phase role: lane -> (leaf, li); column li of the leaf's 6-float4 rows ----------
    const int lsub = lane >> 3, li = lane & 7;
    const int leaf = warp * F3_LPW + lsub;
    const bool leaf_on = li < 6 && leaf < P.L;
    int ls = 0, nq = 0, rem = 0;
    if (leaf_on) {
        ls = t_ls[leaf];
        const int n = t_ll[leaf];
        nq = n >= 8 ? n / 8 : 0;
        rem = n - 8 * nq;
    }
    const int tail_valid = leaf_on ? max(0, min(4, 3 * rem - 4 * li)) : 0;   // comps of the tail row in this column
    const int base4 = 6 * (ls >> 3) + li;          // float4 index of the lane's column in row 0
    const int ke0 = li % 3;                        // axis of comp e is (ke0 + e) % 3
    const int n_first = 3 - ke0;                   // comps e < n_first belong to the column's first atom
    const int a_first = (4 * li) / 3;              // ... which is atom a_first of the 8-atom row
    float4 prev[NQ], prev_tail = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        prev[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < nq) prev[q] = reinterpret_cast<const float4 *>(prev_src)[base4 + 6 * q];
    }
    if (tail_valid) prev_tail = reinterpret_cast<const float4 *>(prev_src)[base4 + 6 * nq];
    // bead of this thread in the lipid phase (first round), kept in registers
    const int my_a0 = tid < P.N ? P.seg[tid] : 0, my_a1 = tid < P.N ? P.seg[tid + 1] : 0;
    const double my_lm = tid < P.N ? P.lipid_mass[tid] : 1.0;
    const double my_rlm = 1.0 / my_lm;
    float Lp0, Lp1, Lp2, Ln0, Ln1, Ln2;
    {
        const int64_t tp = (kc == 0) ? 0 : f0 - 1;
        Lp0 = __ldg(P.box + tp * 3); Lp1 = __ldg(P.box + tp * 3 + 1); Lp2 = __ldg(P.box + tp * 3 + 2);
        Ln0 = __ldg(P.box + f0 * 3); Ln1 = __ldg(P.box + f0 * 3 + 1); Ln2 = __ldg(P.box + f0 * 3 + 2);
    }
    auto issue_frame = [&](float *dst, int64_t t, uint64_t *b) {
        f3_mbar_expect_tx(b, frame_bytes);
        const uint32_t piece = ((frame_bytes / F3_PIECES) + 15u) & ~15u;
        const unsigned char *src = reinterpret_cast<const unsigned char *>(P.x + t * (int64_t)C3);
        for (uint32_t o = 0; o < frame_bytes; o += piece)
            f3_bulk_g2s(reinterpret_cast<unsigned char *>(dst) + o, src + o, min(piece, frame_bytes - o), b);
    };
    if (tid == 0) issue_frame(XB0, f0, &bar[0]);

    unsigned pc[5] = {0u, 0u, 0u, 0u, 0u};
    unsigned tk = (unsigned)clock();
// (a barrier only blocks at the next use of protected state: the clock read is made to depend
// on a shared-memory load issued after it)
#define F3_MARK(i) do { if (P.prof && tid == 0) { unsigned n_; const int dep_ = *(volatile int *)&s_tab[0]; \
        asm volatile("mov.u32 %0, %%clock;" : "=r"(n_) : "r"(dep_)); pc[i] += n_ - tk; tk = n_; } } while (0)

    // leaflet labels of frame tf from s_z / red / s_pmax (written by the lipid phase of tf);
    // thread zt of nzt label threads.  The tree mean decides unless a lipid is within the
    // rounding bound of it; then the warp replays the sequential Welford mean (running_stats.py:34-52).
    auto write_labels = [&](int64_t tf, int zt, int nzt) {
        double s2 = lane < F3_WARPS ? red[lane] : 0.0, m2 = lane < F3_WARPS ? s_pmax[lane] : 0.0;
        s2 = pbk_warp_sum(s2);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, o));
        double mean = s2 / (double)P.N;
        const double tol = 64.0 * (double)P.N * 2.220446049250313e-16 * m2;
        int close = 0;
        for (int i = zt; i < P.N; i += nzt) if (fabs(s_z[i] - mean) <= tol) close = 1;
        if (__any_sync(0xffffffffu, close)) {
            double m = 0.0;
            if (lane == 0)
                for (int i = 0; i < P.N; i++) {
                    const double v = s_z[i];
                    m = (i == 0) ? v : __dadd_rn(m, __ddiv_rn(__dsub_rn(v, m), (double)(i + 1)));
                }
            mean = __shfl_sync(0xffffffffu, m, 0);
        }
        uint8_t *out = P.labels + tf * (int64_t)P.N;
        for (int i = zt; i < P.N; i += nzt) {
            const double v = s_z[i];
            if (v > mean) out[i] = 1;
            else if (v < mean) out[i] = 0;
            else { out[i] = 0; bad |= 8; }
        }
    };

    for (int64_t t = f0; t < f1; t++) {
        const int b = (int)(t - f0) & 1;
        float *cur = b ? XB1 : XB0;
        const float Lc0 = Ln0, Lc1 = Ln1, Lc2 = Ln2;
        const bool Lbig = !(Lc0 < 8192.0f && Lc1 < 8192.0f && Lc2 < 8192.0f);
        if (!f3_mbar_wait(&bar[b], (uint32_t)((t - f0) >> 1) & 1u)) bad |= 2;
        F3_MARK(0);                             // wait for the frame
        // ---- leaf phase: image counts of frame t and the NumPy-pairwise leaf sums of m*f64(u).
        // Pass A tests every row against the previous raw frame (registers) and only notes the
        // rows that fail; the failing lanes of a warp then run the exact reference rule TOGETHER
        // (previous values re-read from global memory / L2); pass B forms the products. --------
        if (leaf_on) {
            const bool boxok = Lc0 > 0.0f && Lc1 > 0.0f && Lc2 > 0.0f;
            float Le[4], he[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int k = (ke0 + e) % 3;
                const float Lc = k == 0 ? Lc0 : (k == 1 ? Lc1 : Lc2), Lp = k == 0 ? Lp0 : (k == 1 ? Lp1 : Lp2);
                Le[e] = Lc;
                he[e] = boxok ? Lc * 0.5f - pbk_image_eps(Lc, Lp) : -1.0f;   // bad box: every test fails
            }
            const float4 *cur4 = reinterpret_cast<const float4 *>(cur);
            uint32_t failrows = 0u;
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                if (FULL || q < nq) {
                    const float4 cv = cur4[base4 + 6 * q];
                    const float4 pv = prev[q];
                    const float m01 = fminf(he[0] - fabsf(__fsub_rn(cv.x, pv.x)), he[1] - fabsf(__fsub_rn(cv.y, pv.y)));
                    const float m23 = fminf(he[2] - fabsf(__fsub_rn(cv.z, pv.z)), he[3] - fabsf(__fsub_rn(cv.w, pv.w)));
                    if (!(fminf(m01, m23) > 0.0f)) failrows |= 1u << q;
                }
            }
            if (failrows) {                     // rare per lane; the warp's failing lanes run this together
                uint32_t todo = failrows;
                while (todo) {
                    const int q = __ffs(todo) - 1;
                    todo &= todo - 1u;
                    const int f4 = base4 + 6 * q;
                    const float4 cv = cur4[f4];
                    float4 pv = prev[0];
#pragma unroll
                    for (int qq = 1; qq < NQ; qq++) if (qq == q) pv = prev[qq];
                    const uint32_t sw = S32[f4];
                    const float xs[4] = {cv.x, cv.y, cv.z, cv.w}, ps[4] = {pv.x, pv.y, pv.z, pv.w};
                    uint32_t nw = sw;
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        if (!(he[e] - fabsf(__fsub_rn(xs[e], ps[e])) > 0.0f)) {
                            const int k = (ke0 + e) % 3;
                            const float Lp = k == 0 ? Lp0 : (k == 1 ? Lp1 : Lp2);
                            float sl;
                            int sn = f3_shift(__fsub_rn(xs[e], pbk_unwrapped(ps[e], f3_sint(sw, e), Lp)), Le[e], &sl, &bad);
                            if (sn > 127 || sn < -127) { bad |= 4; sn = 0; }
                            nw = (nw & ~(0xffu << (8 * e))) | ((uint32_t)(sn + 128) << (8 * e));
                        }
                    }
                    if (nw != sw) { S32[f4] = nw; f3_nnz_update(sw, nw, ke0, s_nnz); }
                }
            }
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            if (um) {
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    if (FULL || q < nq) {
                        const int f4 = base4 + 6 * q;
                        const float4 cv = cur4[f4];
                        const uint32_t sw = S32[f4];
                        prev[q] = cv;
                        const float xv[4] = {cv.x, cv.y, cv.z, cv.w};
                        float uv[4];
                        f3_unwrap4(xv, sw, Le, Lbig, uv);
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const double pr = __dmul_rn(um_val, (double)uv[e]);
                            acc[e] = (q == 0) ? pr : __dadd_rn(acc[e], pr);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    if (FULL || q < nq) {
                        const int f4 = base4 + 6 * q;
                        const float4 cv = cur4[f4];
                        const uint32_t sw = S32[f4];
                        prev[q] = cv;
                        const float xv[4] = {cv.x, cv.y, cv.z, cv.w};
                        float uv[4];
                        f3_unwrap4(xv, sw, Le, Lbig, uv);
                        const int arow = ls + 8 * q + a_first;
                        const double ma = __ldg(gmass + arow), mb = __ldg(gmass + arow + 1);
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const double pr = __dmul_rn(e < n_first ? ma : mb, (double)uv[e]);
                            acc[e] = (q == 0) ? pr : __dadd_rn(acc[e], pr);
                        }
                    }
                }
            }
            if (FULL || nq > 0) {
                double *my = s_acc + (size_t)leaf * 24 + 4 * li;
                my[0] = acc[0]; my[1] = acc[1]; my[2] = acc[2]; my[3] = acc[3];
            }
            if (!FULL && tail_valid) {          // the n % 8 atoms after the last full row (last leaf only)
                const int f4 = base4 + 6 * nq;
                const float4 cv = cur4[f4];
                uint32_t sw = S32[f4];
                const float4 pv = prev_tail;
                prev_tail = cv;
                const float xv[4] = {cv.x, cv.y, cv.z, cv.w}, pvv[4] = {pv.x, pv.y, pv.z, pv.w};
                bool same = true;
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (e < tail_valid) same = same && (he[e] - fabsf(__fsub_rn(xv[e], pvv[e])) > 0.0f);
                if (!same) {
                    const unsigned long long r = f3_column_exact(cv.x, cv.y, cv.z, cv.w, pv.x, pv.y, pv.z, pv.w, sw,
                                                                 Lc0, Lc1, Lc2, Lp0, Lp1, Lp2, ke0, tail_valid);
                    f3_nnz_update(sw, (uint32_t)r, ke0, s_nnz);
                    sw = (uint32_t)r;
                    bad |= (int)(r >> 32);
                    S32[f4] = sw;
                }
                const int arow = ls + 8 * nq + a_first;
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (e < tail_valid) {
                        const float u = f3_unwrap_slow(xv[e], f3_sint(sw, e), Le[e]);
                        s_tail[4 * li + e] = __dmul_rn(F3_MASS(e < n_first ? arow : arow + 1), (double)u);
                    }
            }
        }
        __syncwarp();
        if (leaf_on && li < 3) {                // lane li combines axis li of its leaf
            const double *r = s_acc + (size_t)leaf * 24 + li;
            double v = 0.0;
            if (FULL || nq > 0)
                v = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[3]), __dadd_rn(r[6], r[9])),
                              __dadd_rn(__dadd_rn(r[12], r[15]), __dadd_rn(r[18], r[21])));
            if (!FULL) for (int a = 0; a < rem; a++) v = __dadd_rn(v, s_tail[3 * a + li]);
            s_leaf[leaf * 3 + li] = v;
        }
        Lp0 = Lc0; Lp1 = Lc1; Lp2 = Lc2;
        __syncthreads();
        F3_MARK(1);                             // leaf phase
        if (warp == 0) {                        // summation tree + division in one warp
            if (P.perfect >= 0) {
                // perfect binary tree over L = 2^perfect leaves in order: butterfly (IEEE addition is
                // commutative, so every lane ends with the root of NumPy's recursion)
                double v0 = 0.0, v1 = 0.0, v2 = 0.0;
                if (P.L == 64) {
                    const double *a = s_leaf + 6 * lane;
                    v0 = __dadd_rn(a[0], a[3]); v1 = __dadd_rn(a[1], a[4]); v2 = __dadd_rn(a[2], a[5]);
                } else if (lane < P.L) {
                    const double *a = s_leaf + 3 * lane;
                    v0 = a[0]; v1 = a[1]; v2 = a[2];
                }
                const int top = P.L == 64 ? 16 : (P.L >> 1);
                for (int o = 1; o <= top; o <<= 1) {
                    v0 = __dadd_rn(v0, __shfl_xor_sync(0xffffffffu, v0, o));
                    v1 = __dadd_rn(v1, __shfl_xor_sync(0xffffffffu, v1, o));
                    v2 = __dadd_rn(v2, __shfl_xor_sync(0xffffffffu, v2, o));
                }
                v1 = __shfl_sync(0xffffffffu, v1, 0);       // (lanes >= L hold nothing when L < 32)
                v2 = __shfl_sync(0xffffffffu, v2, 0);
                if (lane < 3) {
                    const double mc = __ddiv_rn(lane == 0 ? v0 : (lane == 1 ? v1 : v2), P.total_mass);
                    s_mem[lane] = mc;
                    P.mem_com[t * 3 + lane] = mc;
                }
            } else {
                for (int h = 0; h < P.H; h++) {
                    const int a = t_lo[h], bnd = t_lo[h + 1];
                    for (int w = lane; w < (bnd - a) * 3; w += 32) {
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
        } else {
            // the other buffer was last read by the lipid phase of frame t-1 (two barriers ago):
            // one lane of an otherwise idle warp fetches frame t+1 into it now, a lipid phase ahead
            if (tid == 32 && t + 1 < f1) issue_frame(b ? XB0 : XB1, t + 1, &bar[b ^ 1]);
            // meanwhile: leaflet labels of the previous frame
            if (P.do_labels && t > f0) write_labels(t - 1, tid - 32, F3_THREADS - 32);
        }
        if (t + 1 < f1) {                       // box of the next frame (latency hidden by the lipid phase)
            Ln0 = __ldg(P.box + (t + 1) * 3); Ln1 = __ldg(P.box + (t + 1) * 3 + 1); Ln2 = __ldg(P.box + (t + 1) * 3 + 2);
        }
        __syncthreads();
        F3_MARK(2);                             // tree (+ labels of t-1)
        // ---- lipid phase: wrapped and unwrapped COM of every lipid, atoms in order -------------
        {
            const double m0 = s_mem[0], m1 = s_mem[1], m2 = s_mem[2];
            // axes on which no atom carries an image count: u == x there, f64(u) is f64(x)
            const int zmask = (s_nnz[0] == 0 ? 1 : 0) | (s_nnz[1] == 0 ? 2 : 0) | (s_nnz[2] == 0 ? 4 : 0);
            const int32_t *__restrict__ gat = P.gather;
            double sacc = 0.0, macc = 0.0;
            for (int i0 = 0; i0 < P.N; i0 += F3_THREADS) {
                const int i = i0 + tid;
                double zv = 0.0;
                if (i < P.N) {
                    const int a0 = i0 == 0 ? my_a0 : __ldg(P.seg + i), a1 = i0 == 0 ? my_a1 : __ldg(P.seg + i + 1);
                    double cw0 = 0, cw1 = 0, cw2 = 0, cu0 = 0, cu1 = 0, cu2 = 0;
                    if (vec_beads) {
                        const float4 *x4 = reinterpret_cast<const float4 *>(cur) + (a0 * 3) / 4;
                        const uint32_t *s4 = S32 + (a0 * 3) / 4;
                        const int ng = (a1 - a0) >> 2;
                        if (zmask == 7) {
                            for (int g = 0; g < ng; g++)
                                f3_lipid_group<7>(x4 + 3 * g, s4 + 3 * g, F3_MASS(a0 + 4 * g), F3_MASS(a0 + 4 * g + 1), F3_MASS(a0 + 4 * g + 2),
                                                  F3_MASS(a0 + 4 * g + 3), Lc0, Lc1, Lc2, Lbig, m0, m1, m2, cw0, cw1, cw2, cu0, cu1, cu2);
                        } else if (zmask == 4) {
                            for (int g = 0; g < ng; g++)
                                f3_lipid_group<4>(x4 + 3 * g, s4 + 3 * g, F3_MASS(a0 + 4 * g), F3_MASS(a0 + 4 * g + 1), F3_MASS(a0 + 4 * g + 2),
                                                  F3_MASS(a0 + 4 * g + 3), Lc0, Lc1, Lc2, Lbig, m0, m1, m2, cw0, cw1, cw2, cu0, cu1, cu2);
                        } else {
                            for (int g = 0; g < ng; g++)
                                f3_lipid_group<0>(x4 + 3 * g, s4 + 3 * g, F3_MASS(a0 + 4 * g), F3_MASS(a0 + 4 * g + 1), F3_MASS(a0 + 4 * g + 2),
                                                  F3_MASS(a0 + 4 * g + 3), Lc0, Lc1, Lc2, Lbig, m0, m1, m2, cw0, cw1, cw2, cu0, cu1, cu2);
                        }
                    } else {
                        for (int p = a0; p < a1; p++) {
                            const int a = gat ? __ldg(gat + p) : p;
                            const double m = F3_MASS(a);
                            const float x0 = cur[a * 3], x1 = cur[a * 3 + 1], x2 = cur[a * 3 + 2];
                            const int s0 = (int)S8[a * 3] - 128, s1 = (int)S8[a * 3 + 1] - 128, s2 = (int)S8[a * 3 + 2] - 128;
                            if ((s0 | s1 | s2) == 0) {
                                f3_comp0(x0, m0, m, cw0, cu0); f3_comp0(x1, m1, m, cw1, cu1); f3_comp0(x2, m2, m, cw2, cu2);
                            } else {
                                f3_comp1(x0, pbk_unwrapped(x0, s0, Lc0), m0, m, cw0, cu0);
                                f3_comp1(x1, pbk_unwrapped(x1, s1, Lc1), m1, m, cw1, cu1);
                                f3_comp1(x2, pbk_unwrapped(x2, s2, Lc2), m2, m, cw2, cu2);
                            }
                        }
                    }
                    const double lm = i0 == 0 ? my_lm : __ldg(P.lipid_mass + i);
                    const double rlm = i0 == 0 ? my_rlm : 1.0 / lm;        // pbk_div_r: RN(c / lm) without the division routine
                    double *ow = P.com + (t * (int64_t)P.N + i) * 3;
                    double *ou = P.com_unwrap + (t * (int64_t)P.N + i) * 3;
                    ow[0] = pbk_div_r(cw0, lm, rlm); ow[1] = pbk_div_r(cw1, lm, rlm); ow[2] = pbk_div_r(cw2, lm, rlm);
                    const double u0 = pbk_div_r(cu0, lm, rlm), u1 = pbk_div_r(cu1, lm, rlm), u2 = pbk_div_r(cu2, lm, rlm);
                    ou[0] = u0; ou[1] = u1; ou[2] = u2;
                    zv = P.norm == 0 ? u0 : (P.norm == 1 ? u1 : u2);
                    s_z[i] = zv;
                }
                if (P.do_labels) {              // warp partials of z for the tree mean of the labels
                    double mx = fabs(zv);
                    sacc += pbk_warp_sum(zv);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    macc = fmax(macc, mx);
                }
            }
            if (P.do_labels && lane == 0) { red[warp] = sacc; s_pmax[warp] = macc; }
        }
        // carried state for the next batch: u of the last frame of the batch
        if (t == P.F - 1) {
            for (int c = tid; c < C3; c += F3_THREADS) {
                const int k = c % 3;
                P.u_out[c] = pbk_unwrapped(cur[c], (int)S8[c] - 128, k == 0 ? Lc0 : (k == 1 ? Lc1 : Lc2));
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of `cur` before the async write
        __syncthreads();
        F3_MARK(3);                             // lipid phase
    }
    if (P.do_labels) write_labels(f1 - 1, tid, F3_THREADS);
    F3_MARK(4);
    if (P.prof && tid == 0) for (int i = 0; i < 8; i++) P.prof[kc * 8 + i] = i < 5 ? (long long)pc[i] : 0;
    if (bad) {
        int st = 0;
        if (bad & 1) st |= PBK_ST_IMAGE_OVERFLOW;
        if (bad & 2) st |= PBK_ST_INTERNAL;
        if (bad & 4) st |= PBK_ST_FUSED_RANGE;
        if (bad & 8) st |= PBK_ST_LEAFLET_TIE;
        atomicOr(P.status, st);
    }
}

// ------------------------------------------------------------------------------------------
// k_frames4: the same three phases per frame, TWO CTAs per SM.  The image counts do not come from a
// same-image test against the previous frame any more: k_img3 already looks at every jump, so it now
// also notes WHEN a count moves (an event list per chunk, sorted by frame in k_evsort), and a frame
// only has to apply its few events to the shared int8 counts.  That frees the 48 registers and the
// second frame buffer the previous frame took: one frame buffer (the next frame's bulk copy is issued
// when the lipid phase is done and hides under the other CTA of the SM), 64 registers per thread,
// 32 warps per SM whose phases interleave instead of waiting on each other's barriers.
// ------------------------------------------------------------------------------------------
template <int NQ, bool FULL>
__global__ void __launch_bounds__(F3_THREADS, 2)
k_frames4(F3Params P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C3 = 3 * P.M;                       // multiple of 12
    // doubles first, then the two frame buffers (128-byte aligned), then the image counts
    double *s_acc = reinterpret_cast<double *>(smem_raw);               // L*24 accumulators
    double *s_leaf = s_acc + (size_t)P.L * 24;                          // (L+K)*3
    double *s_z = s_leaf + (size_t)(P.L + P.K) * 3;                     // N   com_unwrap[norm]
    double *s_tail = s_z + P.N;                                         // 24
    size_t off = ((size_t)(reinterpret_cast<unsigned char *>(s_tail + 24) - smem_raw) + 127) & ~(size_t)127;
    float *XB0 = reinterpret_cast<float *>(smem_raw + off);
    uint32_t *S32 = reinterpret_cast<uint32_t *>(XB0 + C3);             // C3 biased int8 image counts
    unsigned char *S8 = reinterpret_cast<unsigned char *>(S32);
    __shared__ int s_tab[F4_TAB];               // leaf_start | leaf_len | node_left | node_right | level_off
    __shared__ uint64_t bar[1];
    __shared__ double s_mem[3];
    __shared__ double red[F3_WARPS];
    __shared__ double s_pmax[F3_WARPS];
    __shared__ int s_nnz[3];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kc = blockIdx.x;
    int64_t f0, f1;
    f3_chunk_range(kc, P.chunk, P.split, P.F, &f0, &f1);
    if (f0 >= f1) return;
    const uint32_t frame_bytes = (uint32_t)C3 * 4u;
    int bad = 0;

    if (tid == 0) f3_mbar_init(&bar[0], 1);
    if (tid < 3) s_nnz[tid] = 0;
    __syncthreads();
    int *t_ls = s_tab, *t_ll = t_ls + P.L, *t_nl = t_ll + P.L, *t_nr = t_nl + P.K, *t_lo = t_nr + P.K;
    for (int i = tid; i < P.L; i += F3_THREADS) { t_ls[i] = P.leaf_start[i]; t_ll[i] = P.leaf_len[i]; }
    for (int i = tid; i < P.K; i += F3_THREADS) { t_nl[i] = P.node_left[i]; t_nr[i] = P.node_right[i]; }
    for (int i = tid; i <= P.H; i += F3_THREADS) t_lo[i] = P.level_off[i];
    // uniform masses (coarse-grained models): the mass is a register, not a load
    int uni = 1;
    const double m_first = P.mass[0];
    for (int a = tid; a < P.M; a += F3_THREADS) if (P.mass[a] != m_first) uni = 0;
    // vector lipid phase: contiguous beads, every start and length a multiple of 4 atoms
    int vec = P.gather ? 0 : 1;
    for (int i = tid; i < P.N; i += F3_THREADS) {
        const int a0 = P.seg[i], a1 = P.seg[i + 1];
        if ((a0 & 3) || ((a1 - a0) & 3)) vec = 0;
    }
    // image counts before the chunk's first frame: 0 / init_images for chunk 0 (a later batch's jump from
    // the carried state is frame 0's event), the scan's counts otherwise
    for (int c = tid; c < C3; c += F3_THREADS) {
        int s0 = (kc == 0) ? (P.init_images ? P.init_images[c] : (P.first_frame ? 0 : (int)P.first_shift[c]))
                           : (int)P.sstart[(int64_t)kc * C3 + c];
        if (s0 > 127 || s0 < -127) { bad |= 4; s0 = 0; }
        S8[c] = (unsigned char)(s0 + 128);
        if (s0 != 0) atomicAdd(&s_nnz[c % 3], 1);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const bool um = __syncthreads_and(uni) != 0;        // also publishes tables, S, the barriers
    const bool vec_beads = __syncthreads_and(vec) != 0;
    const double um_val = m_first;
    const double *__restrict__ gmass = P.mass;
#define F3_MASS(a) (um ? um_val : __ldg(gmass + (a)))

    // ---- leaf-phase role: lane -> (leaf, li); column li of the leaf's 6-float4 rows ----------
    const int lsub = lane >> 3, li = lane & 7;
    const int leaf = warp * F3_LPW + lsub;
    const bool leaf_on = li < 6 && leaf < P.L;
    int ls = 0, nq = 0, rem = 0;
    if (leaf_on) {
        ls = t_ls[leaf];
        const int n = t_ll[leaf];
        nq = n >= 8 ? n / 8 : 0;
        rem = n - 8 * nq;
    }
    const int tail_valid = leaf_on ? max(0, min(4, 3 * rem - 4 * li)) : 0;   // comps of the tail row in this column
    const int base4 = 6 * (ls >> 3) + li;          // float4 index of the lane's column in row 0
    const int ke0 = li % 3;                        // axis of comp e is (ke0 + e) % 3
    const int n_first = 3 - ke0;                   // comps e < n_first belong to the column's first atom
    const int a_first = (4 * li) / 3;              // ... which is atom a_first of the 8-atom row
    // this chunk's events, sorted by frame: ev_off[t - f0] .. ev_off[t - f0 + 1]
    const uint2 *__restrict__ evs = P.ev_sorted + (int64_t)kc * P.ev_cap;
    const int *__restrict__ ev_off = P.ev_off + (int64_t)kc * (P.chunk + 1);
    // bead of this thread in the lipid phase (first round), kept in registers
    const int my_a0 = tid < P.N ? P.seg[tid] : 0, my_a1 = tid < P.N ? P.seg[tid + 1] : 0;
    const double my_lm = tid < P.N ? P.lipid_mass[tid] : 1.0;
    const double my_rlm = 1.0 / my_lm;
    float Ln0, Ln1, Ln2;
    Ln0 = __ldg(P.box + f0 * 3); Ln1 = __ldg(P.box + f0 * 3 + 1); Ln2 = __ldg(P.box + f0 * 3 + 2);
    auto issue_frame = [&](float *dst, int64_t t, uint64_t *b) {
        f3_mbar_expect_tx(b, frame_bytes);
        const uint32_t piece = ((frame_bytes / F3_PIECES) + 15u) & ~15u;
        const unsigned char *src = reinterpret_cast<const unsigned char *>(P.x + t * (int64_t)C3);
        for (uint32_t o = 0; o < frame_bytes; o += piece)
            f3_bulk_g2s(reinterpret_cast<unsigned char *>(dst) + o, src + o, min(piece, frame_bytes - o), b);
    };
    if (tid == 0) issue_frame(XB0, f0, &bar[0]);

    unsigned pc[5] = {0u, 0u, 0u, 0u, 0u};
    unsigned tk = (unsigned)clock();
// (a barrier only blocks at the next use of protected state: the clock read is made to depend
// on a shared-memory load issued after it)
#define F3_MARK(i) do { if (P.prof && tid == 0) { unsigned n_; const int dep_ = *(volatile int *)&s_tab[0]; \
        asm volatile("mov.u32 %0, %%clock;" : "=r"(n_) : "r"(dep_)); pc[i] += n_ - tk; tk = n_; } } while (0)

    // leaflet labels of frame tf from s_z / red / s_pmax (written by the lipid phase of tf);
    // thread zt of nzt label threads.  The tree mean decides unless a lipid is within the
    // rounding bound of it; then the warp replays the sequential Welford mean (running_stats.py:34-52).
    auto write_labels = [&](int64_t tf, int zt, int nzt) {
        double s2 = lane < F3_WARPS ? red[lane] : 0.0, m2 = lane < F3_WARPS ? s_pmax[lane] : 0.0;
        s2 = pbk_warp_sum(s2);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, o));
        double mean = s2 / (double)P.N;
        const double tol = 64.0 * (double)P.N * 2.220446049250313e-16 * m2;
        int close = 0;
        for (int i = zt; i < P.N; i += nzt) if (fabs(s_z[i] - mean) <= tol) close = 1;
        if (__any_sync(0xffffffffu, close)) {
            double m = 0.0;
            if (lane == 0)
                for (int i = 0; i < P.N; i++) {
                    const double v = s_z[i];
                    m = (i == 0) ? v : __dadd_rn(m, __ddiv_rn(__dsub_rn(v, m), (double)(i + 1)));
                }
            mean = __shfl_sync(0xffffffffu, m, 0);
        }
        uint8_t *out = P.labels + tf * (int64_t)P.N;
        for (int i = zt; i < P.N; i += nzt) {
            const double v = s_z[i];
            if (v > mean) out[i] = 1;
            else if (v < mean) out[i] = 0;
            else { out[i] = 0; bad |= 8; }
        }
    };

    for (int64_t t = f0; t < f1; t++) {
        float *cur = XB0;
        const float Lc0 = Ln0, Lc1 = Ln1, Lc2 = Ln2;
        const bool Lbig = !(Lc0 < 8192.0f && Lc1 < 8192.0f && Lc2 < 8192.0f);
        {   // the frame's events: count moves of a few chains (k_img3 found them, k_evsort ordered them)
            const int e0 = __ldg(ev_off + (int)(t - f0)), e1 = __ldg(ev_off + (int)(t - f0) + 1);
            for (int j = e0 + tid; j < e1; j += F3_THREADS) {
                const uint2 ev = __ldg(evs + j);
                const int c = (int)ev.x, sn = (int)S8[c] - 128 + (int)(int8_t)(ev.y & 0xffu);
                if (sn > 127 || sn < -127) { bad |= 4; continue; }
                if ((S8[c] == 128) != (sn == 0)) atomicAdd(&s_nnz[c % 3], sn == 0 ? -1 : 1);
                S8[c] = (unsigned char)(sn + 128);
            }
        }
        if (!f3_mbar_wait(&bar[0], (uint32_t)(t - f0) & 1u)) bad |= 2;
        __syncthreads();                        // the counts of frame t are in place
        F3_MARK(0);                             // wait for the frame
        // ---- leaf phase: the NumPy-pairwise leaf sums of m*f64(u), u from the frame's counts -------
        if (leaf_on) {
            float Le[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int k = (ke0 + e) % 3;
                Le[e] = k == 0 ? Lc0 : (k == 1 ? Lc1 : Lc2);
            }
            const float4 *cur4 = reinterpret_cast<const float4 *>(cur);
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            if (um) {
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    if (FULL || q < nq) {
                        const int f4 = base4 + 6 * q;
                        const float4 cv = cur4[f4];
                        const uint32_t sw = S32[f4];
                        const float xv[4] = {cv.x, cv.y, cv.z, cv.w};
                        float uv[4];
                        f3_unwrap4(xv, sw, Le, Lbig, uv);
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const double pr = __dmul_rn(um_val, (double)uv[e]);
                            acc[e] = (q == 0) ? pr : __dadd_rn(acc[e], pr);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    if (FULL || q < nq) {
                        const int f4 = base4 + 6 * q;
                        const float4 cv = cur4[f4];
                        const uint32_t sw = S32[f4];
                        const float xv[4] = {cv.x, cv.y, cv.z, cv.w};
                        float uv[4];
                        f3_unwrap4(xv, sw, Le, Lbig, uv);
                        const int arow = ls + 8 * q + a_first;
                        const double ma = __ldg(gmass + arow), mb = __ldg(gmass + arow + 1);
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const double pr = __dmul_rn(e < n_first ? ma : mb, (double)uv[e]);
                            acc[e] = (q == 0) ? pr : __dadd_rn(acc[e], pr);
                        }
                    }
                }
            }
            if (FULL || nq > 0) {
                double *my = s_acc + (size_t)leaf * 24 + 4 * li;
                my[0] = acc[0]; my[1] = acc[1]; my[2] = acc[2]; my[3] = acc[3];
            }
            if (!FULL && tail_valid) {          // the n % 8 atoms after the last full row (last leaf only)
                const int f4 = base4 + 6 * nq;
                const float4 cv = cur4[f4];
                const uint32_t sw = S32[f4];
                const float xv[4] = {cv.x, cv.y, cv.z, cv.w};
                const int arow = ls + 8 * nq + a_first;
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (e < tail_valid) {
                        const float u = f3_unwrap_slow(xv[e], f3_sint(sw, e), Le[e]);
                        s_tail[4 * li + e] = __dmul_rn(F3_MASS(e < n_first ? arow : arow + 1), (double)u);
                    }
            }
        }
        __syncwarp();
        if (leaf_on && li < 3) {                // lane li combines axis li of its leaf
            const double *r = s_acc + (size_t)leaf * 24 + li;
            double v = 0.0;
            if (FULL || nq > 0)
                v = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[3]), __dadd_rn(r[6], r[9])),
                              __dadd_rn(__dadd_rn(r[12], r[15]), __dadd_rn(r[18], r[21])));
            if (!FULL) for (int a = 0; a < rem; a++) v = __dadd_rn(v, s_tail[3 * a + li]);
            s_leaf[leaf * 3 + li] = v;
        }
        __syncthreads();
        F3_MARK(1);                             // leaf phase
        if (warp == 0) {                        // summation tree + division in one warp
            if (P.perfect >= 0) {
                // perfect binary tree over L = 2^perfect leaves in order: butterfly (IEEE addition is
                // commutative, so every lane ends with the root of NumPy's recursion)
                double v0 = 0.0, v1 = 0.0, v2 = 0.0;
                if (P.L == 64) {
                    const double *a = s_leaf + 6 * lane;
                    v0 = __dadd_rn(a[0], a[3]); v1 = __dadd_rn(a[1], a[4]); v2 = __dadd_rn(a[2], a[5]);
                } else if (lane < P.L) {
                    const double *a = s_leaf + 3 * lane;
                    v0 = a[0]; v1 = a[1]; v2 = a[2];
                }
                const int top = P.L == 64 ? 16 : (P.L >> 1);
                for (int o = 1; o <= top; o <<= 1) {
                    v0 = __dadd_rn(v0, __shfl_xor_sync(0xffffffffu, v0, o));
                    v1 = __dadd_rn(v1, __shfl_xor_sync(0xffffffffu, v1, o));
                    v2 = __dadd_rn(v2, __shfl_xor_sync(0xffffffffu, v2, o));
                }
                v1 = __shfl_sync(0xffffffffu, v1, 0);       // (lanes >= L hold nothing when L < 32)
                v2 = __shfl_sync(0xffffffffu, v2, 0);
                if (lane < 3) {
                    const double mc = __ddiv_rn(lane == 0 ? v0 : (lane == 1 ? v1 : v2), P.total_mass);
                    s_mem[lane] = mc;
                    P.mem_com[t * 3 + lane] = mc;
                }
            } else {
                for (int h = 0; h < P.H; h++) {
                    const int a = t_lo[h], bnd = t_lo[h + 1];
                    for (int w = lane; w < (bnd - a) * 3; w += 32) {
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
        } else {
            // meanwhile: leaflet labels of the previous frame
            if (P.do_labels && t > f0) write_labels(t - 1, tid - 32, F3_THREADS - 32);
        }
        if (t + 1 < f1) {                       // box of the next frame (latency hidden by the lipid phase)
            Ln0 = __ldg(P.box + (t + 1) * 3); Ln1 = __ldg(P.box + (t + 1) * 3 + 1); Ln2 = __ldg(P.box + (t + 1) * 3 + 2);
        }
        __syncthreads();
        F3_MARK(2);                             // tree (+ labels of t-1)
        // ---- lipid phase: wrapped and unwrapped COM of every lipid, atoms in order -------------
        {
            const double m0 = s_mem[0], m1 = s_mem[1], m2 = s_mem[2];
            // axes on which no atom carries an image count: u == x there, f64(u) is f64(x)
            const int zmask = (s_nnz[0] == 0 ? 1 : 0) | (s_nnz[1] == 0 ? 2 : 0) | (s_nnz[2] == 0 ? 4 : 0);
            const int32_t *__restrict__ gat = P.gather;
            double sacc = 0.0, macc = 0.0;
            for (int i0 = 0; i0 < P.N; i0 += F3_THREADS) {
                const int i = i0 + tid;
                double zv = 0.0;
                if (i < P.N) {
                    const int a0 = i0 == 0 ? my_a0 : __ldg(P.seg + i), a1 = i0 == 0 ? my_a1 : __ldg(P.seg + i + 1);
                    double cw0 = 0, cw1 = 0, cw2 = 0, cu0 = 0, cu1 = 0, cu2 = 0;
                    if (vec_beads) {
                        const float4 *x4 = reinterpret_cast<const float4 *>(cur) + (a0 * 3) / 4;
                        const uint32_t *s4 = S32 + (a0 * 3) / 4;
                        const int ng = (a1 - a0) >> 2;
                        if (zmask == 7) {
                            for (int g = 0; g < ng; g++)
                                f3_lipid_group<7>(x4 + 3 * g, s4 + 3 * g, F3_MASS(a0 + 4 * g), F3_MASS(a0 + 4 * g + 1), F3_MASS(a0 + 4 * g + 2),
                                                  F3_MASS(a0 + 4 * g + 3), Lc0, Lc1, Lc2, Lbig, m0, m1, m2, cw0, cw1, cw2, cu0, cu1, cu2);
                        } else if (zmask == 4) {
                            for (int g = 0; g < ng; g++)
                                f3_lipid_group<4>(x4 + 3 * g, s4 + 3 * g, F3_MASS(a0 + 4 * g), F3_MASS(a0 + 4 * g + 1), F3_MASS(a0 + 4 * g + 2),
                                                  F3_MASS(a0 + 4 * g + 3), Lc0, Lc1, Lc2, Lbig, m0, m1, m2, cw0, cw1, cw2, cu0, cu1, cu2);
                        } else {
                            for (int g = 0; g < ng; g++)
                                f3_lipid_group<0>(x4 + 3 * g, s4 + 3 * g, F3_MASS(a0 + 4 * g), F3_MASS(a0 + 4 * g + 1), F3_MASS(a0 + 4 * g + 2),
                                                  F3_MASS(a0 + 4 * g + 3), Lc0, Lc1, Lc2, Lbig, m0, m1, m2, cw0, cw1, cw2, cu0, cu1, cu2);
                        }
                    } else {
                        for (int p = a0; p < a1; p++) {
                            const int a = gat ? __ldg(gat + p) : p;
                            const double m = F3_MASS(a);
                            const float x0 = cur[a * 3], x1 = cur[a * 3 + 1], x2 = cur[a * 3 + 2];
                            const int s0 = (int)S8[a * 3] - 128, s1 = (int)S8[a * 3 + 1] - 128, s2 = (int)S8[a * 3 + 2] - 128;
                            if ((s0 | s1 | s2) == 0) {
                                f3_comp0(x0, m0, m, cw0, cu0); f3_comp0(x1, m1, m, cw1, cu1); f3_comp0(x2, m2, m, cw2, cu2);
                            } else {
                                f3_comp1(x0, pbk_unwrapped(x0, s0, Lc0), m0, m, cw0, cu0);
                                f3_comp1(x1, pbk_unwrapped(x1, s1, Lc1), m1, m, cw1, cu1);
                                f3_comp1(x2, pbk_unwrapped(x2, s2, Lc2), m2, m, cw2, cu2);
                            }
                        }
                    }
                    const double lm = i0 == 0 ? my_lm : __ldg(P.lipid_mass + i);
                    const double rlm = i0 == 0 ? my_rlm : 1.0 / lm;        // pbk_div_r: RN(c / lm) without the division routine
                    double *ow = P.com + (t * (int64_t)P.N + i) * 3;
                    double *ou = P.com_unwrap + (t * (int64_t)P.N + i) * 3;
                    ow[0] = pbk_div_r(cw0, lm, rlm); ow[1] = pbk_div_r(cw1, lm, rlm); ow[2] = pbk_div_r(cw2, lm, rlm);
                    const double u0 = pbk_div_r(cu0, lm, rlm), u1 = pbk_div_r(cu1, lm, rlm), u2 = pbk_div_r(cu2, lm, rlm);
                    ou[0] = u0; ou[1] = u1; ou[2] = u2;
                    zv = P.norm == 0 ? u0 : (P.norm == 1 ? u1 : u2);
                    s_z[i] = zv;
                }
                if (P.do_labels) {              // warp partials of z for the tree mean of the labels
                    double mx = fabs(zv);
                    sacc += pbk_warp_sum(zv);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    macc = fmax(macc, mx);
                }
            }
            if (P.do_labels && lane == 0) { red[warp] = sacc; s_pmax[warp] = macc; }
        }
        // carried state for the next batch: u of the last frame of the batch
        if (t == P.F - 1) {
            for (int c = tid; c < C3; c += F3_THREADS) {
                const int k = c % 3;
                P.u_out[c] = pbk_unwrapped(cur[c], (int)S8[c] - 128, k == 0 ? Lc0 : (k == 1 ? Lc1 : Lc2));
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of `cur` before the async write
        __syncthreads();
        // the frame buffer is free: the next frame's bulk copy runs while the SM's other CTA computes
        if (tid == 0 && t + 1 < f1) issue_frame(XB0, t + 1, &bar[0]);
        F3_MARK(3);                             // lipid phase
    }
    if (P.do_labels) write_labels(f1 - 1, tid, F3_THREADS);
    F3_MARK(4);
    if (P.prof && tid == 0) for (int i = 0; i < 8; i++) P.prof[kc * 8 + i] = i < 5 ? (long long)pc[i] : 0;
    if (bad) {
        int st = 0;
        if (bad & 1) st |= PBK_ST_IMAGE_OVERFLOW;
        if (bad & 2) st |= PBK_ST_INTERNAL;
        if (bad & 4) st |= PBK_ST_FUSED_RANGE;
        if (bad & 8) st |= PBK_ST_LEAFLET_TIE;
        atomicOr(P.status, st);
    }
}


// largest leaf of NumPy's pairwise split of m elements (pybilt_b200/pairwise.py: build_plan)
static int64_t f3_max_leaf(int64_t m)
{
    if (m <= 128) return m;
    int64_t n2 = m / 2;
    n2 -= n2 % 8;
    const int64_t a = f3_max_leaf(n2), b = f3_max_leaf(m - n2);
    return a > b ? a : b;
}

// log2(number of leaves) when NumPy's recursion over m elements is a perfect binary tree of
// equal leaves (every split exactly in half), else -1
static int f3_perfect(int64_t m)
{
    int lv = 0;
    while (m > 128) {
        if ((m & 1) || ((m / 2) % 8)) return -1;
        m /= 2;
        lv++;
    }
    return lv;
}

// chunks per SM: two when the two-CTA kernel (k_frames4) may run -- sized for that case always
static void f3_tiling(int64_t F, int sm_count, int per_sm, int64_t *chunk, int64_t *n_chunks)
{
    const int64_t slots = (int64_t)sm_count * per_sm;
    int64_t c = (F + slots - 1) / slots;
    if (c < 4) c = 4;
    *chunk = c;
    *n_chunks = (F + c - 1) / c;
}

extern "C" int pbk_reduce_fused_scratch_bytes(int64_t F, int64_t M, int sm_count, int64_t *bytes,
                                             int32_t *n_chunks_out, int32_t *chunk_out)
{
    if (F <= 0 || M <= 0 || sm_count <= 0 || !bytes) return PBK_EINVAL;
    int64_t chunk, n_chunks, chunk2, n_chunks2;
    f3_tiling(F, sm_count, 1, &chunk, &n_chunks);
    f3_tiling(F, sm_count, 2, &chunk2, &n_chunks2);
    // sized for the v1 layout (the larger) and for one extra chunk (a split tiling, see f3_chunk_range)
    const int64_t v1 = 4 * ((n_chunks + 1) * 3 * M * 16 + (n_chunks + 1) * 8) + 3 * M * 4 + 256;
    // k_frames4: two chunks per SM -- columns, counts at the chunk starts, the event lists (raw + sorted) and offsets
    const int64_t v4 = (n_chunks2 + 2) * (3 * M / 4 + 1) * 16 + 3 * M * 4 + (n_chunks2 + 2) * 8 + 64 +
                       (n_chunks2 + 2) * 3 * M + (n_chunks2 + 2) * ((int64_t)F4_EVCAP * 16 + 8 + (chunk2 + 1) * 4) + 3 * M + 1024;
    *bytes = v1 > v4 ? v1 : v4;
    if (n_chunks_out) *n_chunks_out = (int32_t)n_chunks;
    if (chunk_out) *chunk_out = (int32_t)chunk;
    return PBK_OK;
}

extern "C" int pbk_reduce_fused(pbk_ctx *ctx, const float *x, const float *box, const double *mass,
                                const int32_t *seg, const int32_t *gather, const double *lipid_mass,
                                const int32_t *leaf_start, const int32_t *leaf_len, int32_t L,
                                const int32_t *node_left, const int32_t *node_right, int32_t K,
                                const int32_t *level_off, int32_t H, double total_mass, float *u_state,
                                int first_frame, int64_t F, int64_t M, int64_t N, int norm,
                                int do_labels, int phase, const int32_t *init_images,
                                int32_t *net_images, int64_t net_frames, void *scratch, int64_t scratch_bytes,
                                double *com, double *com_unwrap, double *mem_com, uint8_t *labels,
                                int32_t *status, void *stream)
{
    if ((phase & ~3) || phase == 0 || (init_images && !first_frame))
        return pbk_fail(ctx, PBK_EINVAL, "pbk_reduce_fused: bad phase / init_images needs first_frame");
    if (!ctx || !x || !box || !mass || !seg || !lipid_mass || !leaf_start || !leaf_len || !u_state ||
        !scratch || !com || !com_unwrap || !mem_com || !status || (do_labels && !labels) || F < 0 ||
        M <= 0 || N <= 0 || L <= 0 || K < 0 || norm < 0 || norm > 2)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_reduce_fused: bad argument");
    if (F == 0) return PBK_OK;
    const int C3 = (int)(3 * M);
    size_t smem = (size_t)L * 24 * 8 + (size_t)(L + K) * 24 + (size_t)N * 8 + 24 * 8 + 128 + (size_t)C3 * 8 + (size_t)C3;
    const size_t limit = 227 * 1024 - 6 * 1024;  // static shared of the kernel is < 6 KB
    const bool v3 = !getenv("PBK_FUSED_V1") && (M % 4) == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0 &&
                    (reinterpret_cast<uintptr_t>(u_state) % 16) == 0 && L <= F3_LPW * F3_WARPS &&
                    2 * (int64_t)L + 2 * (int64_t)K + H + 1 <= F3_TAB && smem <= limit && M <= (1 << 20);
    const int64_t split = (net_frames > 0 && net_frames < F) ? net_frames : F;
    if (!v3) {
        if (split != F) return PBK_EUNSUPPORTED;      // the first-generation path has no split tiling
        return pbk_reduce_fused_v1(ctx, x, box, mass, seg, gather, lipid_mass, leaf_start, leaf_len, L,
                                   node_left, node_right, K, level_off, H, total_mass, u_state, first_frame,
                                   F, M, N, norm, do_labels, phase, init_images, net_images, scratch,
                                   scratch_bytes, com, com_unwrap, mem_com, labels, status, stream);
    }
    cudaStream_t st = (cudaStream_t)stream;
    int64_t need; int32_t n_chunks, chunk;
    pbk_reduce_fused_scratch_bytes(F, M, ctx->sm_count, &need, &n_chunks, &chunk);
    if (scratch_bytes < need) return pbk_fail(ctx, PBK_EINVAL, "pbk_reduce_fused: scratch too small");
    // two CTAs per SM (k_frames4) when a CTA with ONE frame buffer and the small table fits twice
    const size_t smem4 = smem - (size_t)C3 * 4;
    bool v4 = !getenv("PBK_FUSED_V3") && 2 * (int64_t)L + 2 * (int64_t)K + H + 1 <= F4_TAB &&
              2 * (smem4 + 3 * 1024) <= 228 * 1024;
    if (v4) {
        int64_t c2, n2;
        f3_tiling(F, ctx->sm_count, 2, &c2, &n2);
        if (c2 <= F4_MAXCHUNK) { chunk = (int32_t)c2; n_chunks = (int32_t)n2; }
        else v4 = false;
    }
    const int32_t n_net = (int32_t)((split + chunk - 1) / chunk);       // chunks tiling [0, split)
    n_chunks = n_net + (int32_t)((F - split + chunk - 1) / chunk);
    const int C3q = C3 / 4;
    // scratch layout
    unsigned char *sp = reinterpret_cast<unsigned char *>(scratch);
    ImgCol *cols = reinterpret_cast<ImgCol *>(sp);
    sp += (size_t)n_chunks * C3q * sizeof(ImgCol);
    float *u_out = reinterpret_cast<float *>(sp);
    sp += (size_t)C3 * 4;
    float *chunk_box = reinterpret_cast<float *>(sp);
    sp += (((size_t)n_chunks * 8) + 15) & ~(size_t)15;
    int *chunk_const = reinterpret_cast<int *>(sp);
    sp += (((size_t)n_chunks * 4) + 15) & ~(size_t)15;
    int8_t *sstart = reinterpret_cast<int8_t *>(sp);
    sp += (((size_t)n_chunks * C3) + 15) & ~(size_t)15;
    uint2 *ev_raw = nullptr, *ev_sorted = nullptr;
    int *ev_count = nullptr, *ev_off = nullptr;
    int8_t *first_shift = nullptr;
    if (v4) {
        ev_raw = reinterpret_cast<uint2 *>(sp); sp += (size_t)n_chunks * F4_EVCAP * 8;
        ev_sorted = reinterpret_cast<uint2 *>(sp); sp += (size_t)n_chunks * F4_EVCAP * 8;
        ev_count = reinterpret_cast<int *>(sp); sp += (((size_t)n_chunks * 4) + 15) & ~(size_t)15;
        ev_off = reinterpret_cast<int *>(sp); sp += (((size_t)n_chunks * (chunk + 1) * 4) + 15) & ~(size_t)15;
        first_shift = reinterpret_cast<int8_t *>(sp); sp += (size_t)C3;
        if ((int64_t)(sp - reinterpret_cast<unsigned char *>(scratch)) > scratch_bytes)
            return pbk_fail(ctx, PBK_EINVAL, "pbk_reduce_fused: scratch too small (events)");
    }
    if (phase & 1) {
        k_chunk_box<<<n_chunks, 128, 0, st>>>(box, F, split, chunk, chunk_box, chunk_const, ev_count);
        PBK_CHECK_LAUNCH(ctx, "k_chunk_box");
        // chunks walked per thread: 1 measured best (2 saves the re-read of a chunk's previous frame but costs registers)
        const int span = getenv("PBK_IMG_SPAN") ? atoi(getenv("PBK_IMG_SPAN")) : 1;
        k_img3<<<dim3((C3q + IMG_THREADS - 1) / IMG_THREADS, (n_chunks + span - 1) / span), IMG_THREADS, 0, st>>>(
            reinterpret_cast<const float4 *>(x), box, reinterpret_cast<const float4 *>(u_state), first_frame, F,
            split, C3q, chunk, cols, chunk_const, status, ev_raw, ev_count, first_shift, n_chunks, span);
        PBK_CHECK_LAUNCH(ctx, "k_img3");
        if (net_images) {
            k_net3<<<(C3q + 3) / 4, 128, 0, st>>>(cols, C3, n_net, net_images);
            PBK_CHECK_LAUNCH(ctx, "k_net3");
        }
    }
    if (!(phase & 2)) return PBK_OK;
    k_scan3<<<(C3q + 3) / 4, 128, 0, st>>>(cols, chunk_box, C3, n_chunks, init_images, sstart, status);
    PBK_CHECK_LAUNCH(ctx, "k_scan3");
    F3Params P;
    P.x = x; P.box = box; P.mass = mass; P.seg = seg; P.gather = gather; P.lipid_mass = lipid_mass;
    P.leaf_start = leaf_start; P.leaf_len = leaf_len; P.node_left = node_left; P.node_right = node_right;
    P.level_off = level_off; P.L = L; P.K = K; P.H = H; P.total_mass = total_mass;
    P.perfect = f3_perfect(M);
    if (P.perfect >= 0 && ((1 << P.perfect) != L || L > 64)) P.perfect = -1;
    P.u_state = u_state; P.u_out = u_out; P.first_frame = first_frame; P.F = F; P.M = (int)M; P.N = (int)N;
    P.split = split; P.norm = norm; P.chunk = chunk; P.n_chunks = n_chunks; P.do_labels = do_labels; P.sstart = sstart;
    P.init_images = init_images; P.com = com; P.com_unwrap = com_unwrap; P.mem_com = mem_com;
    P.labels = labels; P.status = status;
    P.prof = (scratch_bytes >= need + 8 * 8 * (int64_t)n_chunks + 64 && getenv("PBK_FUSED_PROF")) ? reinterpret_cast<long long *>(reinterpret_cast<char *>(scratch) + ((need + 63) / 64) * 64) : nullptr;
    // rows of 8 atoms per leaf; FULL when every leaf is exactly that many rows
    const int64_t mleaf = f3_max_leaf(M);
    const bool full = P.perfect >= 0 && M == mleaf * L && (mleaf % 8) == 0;
    void (*kern)(F3Params) = k_frames3<F3_NQ, false>;
    if (full && mleaf == 96) kern = k_frames3<12, true>;
    else if (full && mleaf == 128) kern = k_frames3<16, true>;
    else if (mleaf <= 96) kern = k_frames3<12, false>;
    cudaError_t e;
    if (v4) {
        k_evsort<<<n_chunks, 256, 0, st>>>(ev_raw, ev_count, chunk, ev_sorted, ev_off, status);
        PBK_CHECK_LAUNCH(ctx, "k_evsort");
        P.ev_sorted = ev_sorted; P.ev_off = ev_off; P.ev_cap = F4_EVCAP; P.first_shift = first_shift;
        P.u_out = u_state;      // k_frames4 never reads the carried state (k_img3 did): the new one goes straight there
        kern = k_frames4<F3_NQ, false>;
        if (full && mleaf == 96) kern = k_frames4<12, true>;
        else if (full && mleaf == 128) kern = k_frames4<16, true>;
        else if (mleaf <= 96) kern = k_frames4<12, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "k_frames4 smem attribute", e);
        kern<<<n_chunks, F3_THREADS, smem4, st>>>(P);
        PBK_CHECK_LAUNCH(ctx, "k_frames4");
    } else {
        P.ev_sorted = nullptr; P.ev_off = nullptr; P.ev_cap = 0; P.first_shift = nullptr;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "k_frames3 smem attribute", e);
        kern<<<n_chunks, F3_THREADS, smem, st>>>(P);
        PBK_CHECK_LAUNCH(ctx, "k_frames3");
    }
    // k_frames3: u_state is read by chunk 0 and replaced by the last chunk, so the new state is produced in
    // scratch and copied over once the kernel is done
    if (!v4) {
        e = cudaMemcpyAsync(u_state, u_out, (size_t)C3 * 4, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "u_state copy", e);
    }
    return PBK_OK;
}
