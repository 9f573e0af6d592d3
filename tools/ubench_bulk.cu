// ubench_bulk.cu -- throughput of small cp.async.bulk (TMA 1-D) copies: bytes/clk/SM against copy size
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/ubench_bulk tools/ubench_bulk.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
static __device__ __forceinline__ uint32_t su32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const unsigned char *src, size_t span, int bytes, int ncopy, int depth, long long *cyc, long long *issue)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t bar[16];
    const uint32_t a_bar = su32(bar), a_sm = su32(sm);
    if (threadIdx.x == 0)
        for (int d = 0; d < depth; d++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a_bar + 8 * d));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x != 0) return;
    const unsigned char *p = src + ((size_t)blockIdx.x * 7919 * 4096) % span;
    long long t0 = clock64(), ti = 0;
    for (int i = 0; i < ncopy + depth; i++) {
        const int slot = i % depth;
        if (i >= depth) {      // wait for the copy issued `depth` ago
            uint32_t ok = 0, par = ((i / depth) - 1) & 1;
            while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a_bar + 8 * slot), "r"(par) : "memory");
        }
        if (i < ncopy) {
            long long a = clock64();
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a_bar + 8 * slot), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(a_sm + slot * bytes), "l"(p), "r"(bytes), "r"(a_bar + 8 * slot) : "memory");
            ti += clock64() - a;
            p += 786432;       // next "frame"
            if (p + bytes > src + span) p = src;
        }
    }
    if (blockIdx.x == 0) { cyc[0] = clock64() - t0; issue[0] = ti; }
}
int main()
{
    const size_t span = (size_t)1 << 31;
    unsigned char *src; long long *cyc, h, hi;
    cudaMalloc(&src, span); cudaMemset(src, 1, span); cudaMalloc(&cyc, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int per_sm : {1, 4, 8}) for (int bytes : {768, 1536, 3072, 6144, 12288}) for (int depth : {4, 8}) {
        if ((size_t)bytes * depth * per_sm > 200 * 1024) continue;
        const int ncopy = 2000;
        k<<<148 * per_sm, 32, bytes * depth>>>(src, span, bytes, ncopy, depth, cyc, cyc + 1);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&hi, cyc + 1, 8, cudaMemcpyDeviceToHost);
        printf("CTAs/SM %d  copy %5d B  depth %d : %.0f cycles per copy per CTA (issue %.0f), %.1f B/clk/SM, ~%.2f TB/s %s\n", per_sm, bytes, depth,
               (double)h / ncopy, (double)hi / ncopy, (double)bytes * per_sm * ncopy / h, (double)bytes * per_sm * ncopy / h * 148 * 1.9e9 / 1e12,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
