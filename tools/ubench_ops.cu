// ubench_ops.cu -- per-SM throughput of the conversion / fp64 instructions the kernels lean on.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_ops ubench_ops.cu && ./ubench_ops
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
#define UNR 8
template <int OP>
__global__ void k(float *out, float seed, double dseed)
{
    float a[UNR]; double d[UNR]; int n[UNR];
    for (int j = 0; j < UNR; j++) { a[j] = seed + j + threadIdx.x; d[j] = dseed + j + threadIdx.x; n[j] = j + threadIdx.x; }
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int j = 0; j < UNR; j++) {
            if (OP == 0) { d[j] = (double)a[j]; a[j] = __int_as_float(__double2hiint(d[j]) ^ i); }       // F2F f32->f64 (+ lop)
            if (OP == 1) { a[j] = (float)d[j]; d[j] = __hiloint2double(__float_as_int(a[j]), i); }       // F2F f64->f32
            if (OP == 2) { d[j] = __dadd_rn(d[j], dseed); }                                               // DADD
            if (OP == 3) { d[j] = fma(d[j], dseed, dseed); }                                              // DFMA
            if (OP == 4) { a[j] = fmaf(a[j], seed, seed); }                                               // FFMA
            if (OP == 5) { n[j] = __double2int_rd(d[j]) + i; d[j] = __hiloint2double(n[j], n[j]); }       // F2I f64
            if (OP == 6) { n[j] += (d[j] > dseed + n[j]) ? 1 : 0; }                                        // DSETP + I2F.F64?
            if (OP == 7) { d[j] = __dmul_rn(d[j], dseed); }                                               // DMUL
            if (OP == 8) { n[j] = (n[j] ^ i) + (n[j] >> 3); }                                             // int alu (lop+shf+iadd)
        }
    }
    float s = 0; for (int j = 0; j < UNR; j++) s += a[j] + (float)d[j] + n[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char *name, float opsPerIter)
{
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148 * 2, 1024>>>(out, 1.5f, 1.25); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<148 * 2, 1024>>>(out, 1.5f, 1.25);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 2 * 1024 * ITER * UNR * opsPerIter;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %7.1f ops/clk/SM (at %d MHz nominal)\n", name, ms, ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
    cudaFree(out);
}
int main()
{
    run<0>("F2F.F64.F32 (+LOP)", 1); run<1>("F2F.F32.F64 (+mov)", 1); run<2>("DADD", 1); run<3>("DFMA", 1);
    run<7>("DMUL", 1); run<4>("FFMA", 1); run<5>("F2I.F64 floor (+I2D-ish)", 1); run<6>("DSETP+...", 1); run<8>("int alu x3", 3);
    return 0;
}
