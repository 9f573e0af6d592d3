// pbk_common.cuh -- shared device helpers of libpbk (sm_100a).
//
// Arithmetic contract (see DESIGN.md "Arithmetic"): the library is compiled with
// -fmad=false so that no multiply-add is fused behind our back; the only fused operation
// is the explicit fma() in pbc_dist2(), which restates the FMA accumulation OpenBLAS' ddot
// performs for the reference's np.dot(d_v, d_v).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pbk.h"

#define PBK_WORK_SLOTS 64
#define PBK_LUT_SLOTS 4
#define PBK_LUT_ENTRIES 8192
struct pbk_ctx {
    int device;
    int sm_count;
    char err[512];
    int *work;                  // device: PBK_WORK_SLOTS work counters of persistent kernels (may be NULL)
    unsigned work_next;
    unsigned short *lut;        // device: PBK_LUT_SLOTS d^2 -> bin tables of PBK_LUT_ENTRIES entries (may be NULL)
    void *arena;                // device scratch of the cell-list kernels (pbk_cells.cu), grown on demand
    size_t arena_bytes;
};

// scratch owned by the context; grows with cudaMalloc (which synchronises the device, so no kernel
// still reads the block that is freed).  NULL when the allocation fails.
void *pbk_arena(pbk_ctx *ctx, size_t bytes);

#define PBK_WARP 32

__device__ __forceinline__ float pbk_unwrapped(float x, int s, float L)
{
    // u = f32(f64(x) + s*L)      mda_unwrap_vectorized.py:49,70,92
    if (s == 0) return x;
    // s*L is exact in float64 and, for |s| <= 255, L < 2^13 and |x| >= 2^-8, so is f64(x) + s*L
    // (all its bits lie between 2^21 and 2^-31): the only rounding is the final one to float32,
    // which is what a single float32 FMA performs.  Outside that range: the float64 path.
    if (s >= -255 && s <= 255 && L < 8192.0f && fabsf(x) >= 0.00390625f)
        return __fmaf_rn((float)s, L, x);
    return __double2float_rn(__dadd_rn((double)x, __dmul_rn((double)s, (double)L)));
}

// Branch-free variant for |s| <= 127 (int8 image counts) and L < 2^13: (float)s through the
// 2^23 magic constant (no I2F), one FMA; the float64 route only for |x| < 2^-8 with s != 0.
__device__ __forceinline__ float pbk_unwrapped_i8(float x, int s, float L)
{
    const float sf = __int_as_float(0x4B400000 + s) - 12582912.0f;
    float u = __fmaf_rn(sf, L, x);
    if (s != 0 && (fabsf(x) < 0.00390625f || !(L < 8192.0f)))
        u = __double2float_rn(__dadd_rn((double)x, __dmul_rn((double)s, (double)L)));
    return u;
}

// Image count of frame t from the RAW previous frame.  The reference compares
// d = f32(x_t - u_{t-1}) with +-L/2 where u_{t-1} = f32(x_{t-1} + s_{t-1} L_{t-1}); writing
// d0 = f32(x_t - x_{t-1}), the shifted difference d + s_t L_t equals d0 + (s_t - s_{t-1}) L_t up to
// ulp32(u)/2 + ulp32(d)/2 + ulp32(d0)/2 + |s| |L_t - L_{t-1}|.  So when d0 is further than that
// bound (`eps`, supplied by the caller for |s| <= 127) from +-L/2 the image count cannot change.
// Returns true and leaves *s untouched in that case; false = caller must run the exact rule.
__device__ __forceinline__ bool pbk_same_image(float x, float xprev_raw, float L, float eps)
{
    const float d0 = __fsub_rn(x, xprev_raw);
    return (L * 0.5f - fabsf(d0)) > eps;
}

__device__ __forceinline__ float pbk_image_eps(float Lc, float Lp)
{
    // (|s|+2) * L * 2^-22 + |s| * |dL| with |s| <= 127, L = max(Lc, Lp)
    return 130.0f * fmaxf(Lc, Lp) * 2.3841858e-7f + 127.0f * fabsf(Lc - Lp);
}

// distance_euclidean_pbc(a, b, l_box, center='box_half') with a' = a - c, b' = b - c already
// formed (c = f32 box / 2).  Returns d^2 = fma(dy, dy, dx*dx); distance_cutoff_clustering.py:50-70.
__device__ __forceinline__ double pbk_pbc_dist2(double apx, double apy, double bpx, double bpy,
                                                double Lx, double Ly, double hx, double hy)
{
    double dx = __dsub_rn(apx, bpx);
    double dy = __dsub_rn(apy, bpy);
    if (fabs(dx) > hx) dx = __dsub_rn(__dsub_rn(Lx, fabs(apx)), fabs(bpx));
    if (fabs(dy) > hy) dy = __dsub_rn(__dsub_rn(Ly, fabs(apy)), fabs(bpy));
    return fma(dy, dy, __dmul_rn(dx, dx));
}

// np.histogram uniform-bin rule (numpy/lib/_histograms_impl.py): -1 when out of range.
__device__ __forceinline__ int pbk_hist_bin(double r, double lo, double hi, int nb,
                                            const double *edges)
{
    if (!(r >= lo) || !(r <= hi)) return -1;
    double f = __dmul_rn(__ddiv_rn(__dsub_rn(r, lo), __dsub_rn(hi, lo)), (double)nb);
    int k = (int)f;
    if (k == nb) k -= 1;
    if (r < edges[k]) k -= 1;
    else if (r >= edges[k + 1] && k != nb - 1) k += 1;
    return k;
}

// a / b for a divisor b that is reused (a bead's mass): rb = RN(1 / b) formed once.  q1 is a faithful
// quotient (one Newton correction of a * rb with an exact residual), and a second correction step
// from a faithful quotient with the correctly rounded reciprocal returns RN(a / b) (Markstein's
// theorem; Muller et al., Handbook of Floating-Point Arithmetic, "division with an FMA") -- the value
// __ddiv_rn gives, in five multiply-adds instead of the division subroutine.  Normal range only
// (masses and coordinate sums are far from overflow / underflow).
__device__ __forceinline__ double pbk_div_r(double a, double b, double rb)
{
    const double q0 = __dmul_rn(a, rb);
    const double q1 = __fma_rn(__fma_rn(-q0, b, a), rb, q0);
    return __fma_rn(__fma_rn(-q1, b, a), rb, q1);
}

template <typename T>
__device__ __forceinline__ T pbk_warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block sum (fixed tree), result valid in every thread; red must hold 32 T's
template <typename T>
__device__ __forceinline__ T pbk_block_sum(T v, T *red)
{
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = pbk_warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    T t = (lane < nw) ? red[lane] : T(0);
    t = pbk_warp_sum(t);
    return t;
}

__device__ __forceinline__ double pbk_block_max(double v, double *red)
{
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = (lane < nw) ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
    return t;
}

int pbk_fail(pbk_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess);
#define PBK_CHECK_LAUNCH(ctx, name)                                      \
    do {                                                                 \
        cudaError_t _e = cudaGetLastError();                             \
        if (_e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, name, _e); \
    } while (0)
