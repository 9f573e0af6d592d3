// ubench_conv.cu -- cost of the per-component arithmetic of the lipid phase of k_frames3 under
// its real launch shape (512 threads, one CTA per SM), for several ways of forming
//     cw += m * f64(x);   v = f32(f64(u) - mem);   cu += m * f64(v)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench_conv ubench_conv.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define NCOMP 12
#define ITER 2048

__device__ __forceinline__ double rn24_int(double d, unsigned &flag)
{
    // round-to-nearest-even of a double to 24 significant bits, on the bit pattern
    unsigned lo = (unsigned)__double2loint(d), hi = (unsigned)__double2hiint(d);
    const unsigned e = (hi & 0x7ff00000u) - 0x38100000u;          // exponent - 897 (f32 normal range)
    flag |= (e > 0x0fc00000u) ? 1u : 0u;                           // outside [2^-126, 2^127): slow path
    const unsigned odd = (lo >> 29) & 1u;
    const unsigned t = lo + 0x0fffffffu + odd;
    const unsigned carry = t < lo ? 1u : 0u;
    hi += carry;
    lo = t & 0xe0000000u;
    return __hiloint2double((int)hi, (int)lo);
}

__device__ __forceinline__ double rn24_veltkamp(double d, unsigned &flag)
{
    const unsigned lo = (unsigned)__double2loint(d), hi = (unsigned)__double2hiint(d);
    const unsigned e = (hi & 0x7ff00000u) - 0x38100000u;
    flag |= (e > 0x0fc00000u) ? 1u : 0u;
    flag |= ((lo & 0x1fffffffu) == 0x10000000u) ? 1u : 0u;         // exact tie: slow path
    const double g = __dmul_rn(d, 536870913.0);                    // 2^29 + 1
    const double dl = __dsub_rn(d, g);
    return __dadd_rn(g, dl);
}

__device__ __forceinline__ double f2d_int(float x, unsigned &flag)
{
    const unsigned b = (unsigned)__float_as_int(x);
    flag |= ((b & 0x7f800000u) == 0u) ? 1u : 0u;                   // zero / subnormal: slow path
    const unsigned a = (unsigned)((int)b >> 3);
    const unsigned hi = (a & 0x8fffffffu) + 0x38000000u;
    return __hiloint2double((int)hi, (int)(b << 29));
}

template <int V>
__global__ void __launch_bounds__(512, 1) k(const float *__restrict__ in, double *out, double mem, double m, float L)
{
    extern __shared__ float sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 512 * NCOMP; i += 512) sm[i] = in[i];
    __syncthreads();
    double cw[3] = {0, 0, 0}, cu[3] = {0, 0, 0};
    unsigned flag = 0;
    for (int it = 0; it < ITER; it++) {
        float xs[NCOMP];
        const float4 *p = reinterpret_cast<const float4 *>(sm + tid * NCOMP);
        const float4 A = p[0], B = p[1], C = p[2];
        xs[0] = A.x; xs[1] = A.y; xs[2] = A.z; xs[3] = A.w; xs[4] = B.x; xs[5] = B.y; xs[6] = B.z; xs[7] = B.w;
        xs[8] = C.x; xs[9] = C.y; xs[10] = C.z; xs[11] = C.w;
#pragma unroll
        for (int c = 0; c < NCOMP; c++) {
            const int kx = c % 3;
            const float x = xs[c] + (float)it * 1e-7f;
            const float u = __fmaf_rn((float)(it & 1), L, x);     // stands for the unwrap FMA
            double xd, ud, vd;
            if (V == 0) {                          // 4 conversions
                xd = (double)x; ud = (double)u;
                vd = (double)__double2float_rn(__dsub_rn(ud, mem));
            } else if (V == 1) {                   // 3 conversions (u == x on this axis)
                xd = (double)x; ud = xd;
                vd = (double)__double2float_rn(__dsub_rn(ud, mem));
            } else if (V == 2) {                   // f32 rounding on the bit pattern
                xd = (double)x; ud = (double)u;
                vd = rn24_int(__dsub_rn(ud, mem), flag);
            } else if (V == 3) {                   // f32 rounding by Veltkamp splitting
                xd = (double)x; ud = (double)u;
                vd = rn24_veltkamp(__dsub_rn(ud, mem), flag);
            } else if (V == 4) {                   // no conversion instruction at all
                xd = f2d_int(x, flag); ud = f2d_int(u, flag);
                vd = rn24_int(__dsub_rn(ud, mem), flag);
            } else if (V == 5) {                   // one conversion on XU, one on ALU, Veltkamp rounding
                xd = (double)x; ud = f2d_int(u, flag);
                vd = rn24_veltkamp(__dsub_rn(ud, mem), flag);
            } else {                               // V == 6: shared xd + Veltkamp (1 conversion)
                xd = (double)x; ud = xd;
                vd = rn24_veltkamp(__dsub_rn(ud, mem), flag);
            }
            cw[kx] = __dadd_rn(cw[kx], __dmul_rn(xd, m));
            cu[kx] = __dadd_rn(cu[kx], __dmul_rn(vd, m));
        }
    }
    out[blockIdx.x * 512 + tid] = cw[0] + cw[1] + cw[2] + cu[0] + cu[1] + cu[2] + (double)flag;
}

template <int V> void run(const char *name, const float *in, double *out)
{
    cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V><<<148, 512, 200 * 1024>>>(in, out, 63.99, 72.0, 128.0f); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<V><<<148, 512, 200 * 1024>>>(in, out, 63.99, 72.0, 128.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // cycles per warp-component per SMSP (4 warps on each of the 4 sub-partitions)
    const double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-44s %8.3f ms  %6.1f cycles per (warp, component) per SMSP; %6.1f cycles per lipid-frame of 36 comps on the SM\n",
           name, ms, cyc / ((double)ITER * NCOMP * 4), cyc / ((double)ITER * NCOMP * 16) * 36 * 16);
}

int main()
{
    float *in; double *out;
    cudaMalloc(&in, 512 * NCOMP * 4); cudaMalloc(&out, 148 * 512 * 8);
    float h[512 * NCOMP];
    for (int i = 0; i < 512 * NCOMP; i++) h[i] = 1.0f + (float)(i % 997) * 0.125f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>("V0 4 F2F", in, out);
    run<1>("V1 3 F2F (shared f64(x))", in, out);
    run<2>("V2 2 F2F + integer RN-even", in, out);
    run<3>("V3 2 F2F + Veltkamp", in, out);
    run<4>("V4 0 F2F (integer everything)", in, out);
    run<5>("V5 1 F2F + 1 int cvt + Veltkamp", in, out);
    run<6>("V6 1 F2F (shared) + Veltkamp", in, out);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
