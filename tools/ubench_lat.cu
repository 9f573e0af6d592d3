// ubench_lat.cu -- dependent-chain latencies (cycles) of the instructions the COM kernels lean on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/build/ubench_lat tools/ubench_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double a0, float f0, int warps_note)
{
    double a = a0 + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; i++) a = __dadd_rn(a, 1.0);
    long long t1 = clock64();
    double m = a;
#pragma unroll 16
    for (int i = 0; i < 1024; i++) m = __dmul_rn(m, 1.0000001);
    long long t2 = clock64();
    float f = f0 + (float)m;
#pragma unroll 16
    for (int i = 0; i < 1024; i++) f = (float)((double)f + 1.0) ;   // F2F.F64.F32 -> DADD -> F2F.F32.F64
    long long t3 = clock64();
    float g = f;
#pragma unroll 16
    for (int i = 0; i < 1024; i++) g = __fmaf_rn(g, 1.0001f, 0.5f);
    long long t4 = clock64();
    double s = m;
#pragma unroll 16
    for (int i = 0; i < 1024; i++) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
    long long t5 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + m + f + g + s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
}
int main()
{
    double *out; long long *cyc, h[5];
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 64);
    const int cfgs[][2] = {{1, 32}, {1, 128}, {148, 128}, {148 * 5, 128}, {148 * 8, 256}};
    for (auto &c : cfgs) {
        k<<<c[0], c[1]>>>(out, cyc, 1.0, 2.0f, 0);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 40, cudaMemcpyDeviceToHost);
        printf("grid %4d x %3d: DADD %.1f  DMUL %.1f  F2F+DADD+F2F %.1f  FFMA %.1f  SHFL64+DADD %.1f cycles per dependent op\n", c[0], c[1],
               h[0] / 1024.0, h[1] / 1024.0, h[2] / 1024.0, h[3] / 1024.0, h[4] / 1024.0);
    }
    return 0;
}
