// pbk_comm.cu -- the two exchange steps of a frame-sharded run over NVLink peer memory
// (pbk_comm_init, pbk_allreduce_u64), one process per GPU on one node.
//
// The reference has no distributed path; the exchanges are the ones pybilt_b200/distributed.py
// defines: (1) the exclusive prefix over ranks of the net periodic-image counts [M][3] int32 that
// sits between the two phases of a shard's reduction (the only exchange on the critical path), and
// (2) the sum over ranks of partial RDF histograms / per-frame slot vectors.
//
// Every rank owns one exchange window in device memory (cudaMalloc) that the other ranks map through
// CUDA IPC: `world` slots of `slot_bytes` plus one 64-bit flag per slot.  A collective is one kernel
// that PUSHES the caller's contribution into slot[rank] of every peer's window with plain stores over
// NVLink, publishes it with a system-scope release of the call's epoch number into the peers' flags,
// and then acquires the flags of the ranks it needs and combines their slots out of its OWN window.
// No NCCL launch, no host round trip, no separate prefix kernel.  Epochs only grow, so flags are never
// reset; two alternating halves of the window make call k+1 safe against a peer still reading call k.
#include <stdio.h>
#include <string.h>

#include "pbk_common.cuh"

#define PBK_COMM_MAX_WORLD 16

struct pbk_comm {
    pbk_ctx *ctx;
    int rank, world;
    size_t slot_bytes;              // payload bytes per slot (multiple of 16)
    size_t half_bytes;              // world * slot_bytes + world * 8 (flags), rounded to 256
    unsigned char *win[PBK_COMM_MAX_WORLD];     // win[rank] = own allocation, others = IPC mappings
    bool opened[PBK_COMM_MAX_WORLD];
    unsigned long long epoch;
};

struct CommView {
    unsigned char *win[PBK_COMM_MAX_WORLD];
    int rank, world;
    size_t slot_bytes, half_off;
    unsigned long long epoch;
};

static __device__ __forceinline__ unsigned long long *comm_flag(const CommView &V, int owner, int slot)
{
    return reinterpret_cast<unsigned long long *>(V.win[owner] + V.half_off + (size_t)V.world * V.slot_bytes) + slot;
}
static __device__ __forceinline__ unsigned char *comm_slot(const CommView &V, int owner, int slot)
{
    return V.win[owner] + V.half_off + (size_t)slot * V.slot_bytes;
}
static __device__ __forceinline__ void comm_release(unsigned long long *flag, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(v) : "memory");
}
static __device__ __forceinline__ unsigned long long comm_acquire(const unsigned long long *flag)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    return v;
}

// grid: `world` CTAs push (CTA q writes the contribution into peer q's slot[rank] and releases the flag),
// then every CTA takes part in the combine over int32 elements once the needed flags are visible.
// mode 0: out = sum over ranks q < rank (exclusive prefix);  mode 1: out = sum over all ranks.
template <typename T>
__global__ void __launch_bounds__(512)
k_comm_combine(CommView V, const T *__restrict__ src, T *__restrict__ out, int64_t n, int mode, int32_t *status)
{
    const int q = blockIdx.x;                        // peer this CTA pushes to
    {
        T *dst = reinterpret_cast<T *>(comm_slot(V, q, V.rank));
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) comm_release(comm_flag(V, q, V.rank), V.epoch);
    }
    // every rank's flag is awaited, also by a prefix that only reads the ranks before it: a call is thereby a
    // barrier, which is what lets two alternating window halves suffice (nobody can be two calls ahead)
    const int need = mode == 0 ? V.rank : V.world;   // ranks whose slots are read from the own window
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        int ok = 1;
        for (int r = 0; r < V.world; r++) {
            int it = 0;
            while (comm_acquire(comm_flag(V, V.rank, r)) < V.epoch && ++it < (1 << 24)) __nanosleep(40);
            if (it >= (1 << 24)) ok = 0;
        }
        s_ok = ok;
    }
    __syncthreads();
    if (!s_ok) { if (threadIdx.x == 0) atomicOr(status, PBK_ST_INTERNAL); return; }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        T acc = 0;
        for (int r = 0; r < need; r++) acc += reinterpret_cast<const volatile T *>(comm_slot(V, V.rank, r))[i];
        out[i] = acc;
    }
}

extern "C" int pbk_comm_create(pbk_ctx *ctx, int rank, int world, int64_t slot_bytes, pbk_comm **out)
{
    if (!ctx || !out || rank < 0 || world < 1 || rank >= world || world > PBK_COMM_MAX_WORLD || slot_bytes <= 0)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_comm_create: bad argument (world <= 16)");
    pbk_comm *c = new pbk_comm();
    memset(c, 0, sizeof(*c));
    c->ctx = ctx; c->rank = rank; c->world = world;
    c->slot_bytes = ((size_t)slot_bytes + 15) & ~(size_t)15;
    c->half_bytes = ((size_t)world * c->slot_bytes + (size_t)world * 8 + 255) & ~(size_t)255;
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != ctx->device) cudaSetDevice(ctx->device);
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, 2 * c->half_bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, 2 * c->half_bytes);
    if (cur != ctx->device && cur >= 0) cudaSetDevice(cur);
    if (e != cudaSuccess) { delete c; return pbk_fail(ctx, PBK_ECUDA, "pbk_comm_create: window allocation", e); }
    c->win[rank] = reinterpret_cast<unsigned char *>(p);
    *out = c;
    return PBK_OK;
}

extern "C" int pbk_comm_handle(pbk_comm *c, unsigned char *handle64)
{
    if (!c || !handle64) return PBK_EINVAL;
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, c->win[c->rank]);
    if (e != cudaSuccess) return pbk_fail(c->ctx, PBK_ECUDA, "pbk_comm_handle: cudaIpcGetMemHandle", e);
    static_assert(sizeof(h) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    return PBK_OK;
}

extern "C" int pbk_comm_connect(pbk_comm *c, const unsigned char *handles)
{
    if (!c || !handles) return PBK_EINVAL;
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != c->ctx->device) cudaSetDevice(c->ctx->device);
    int rc = PBK_OK;
    for (int r = 0; r < c->world && rc == PBK_OK; r++) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { rc = pbk_fail(c->ctx, PBK_ECUDA, "pbk_comm_connect: cudaIpcOpenMemHandle", e); break; }
        c->win[r] = reinterpret_cast<unsigned char *>(p);
        c->opened[r] = true;
    }
    if (cur != c->ctx->device && cur >= 0) cudaSetDevice(cur);
    return rc;
}

extern "C" void pbk_comm_destroy(pbk_comm *c)
{
    if (!c) return;
    for (int r = 0; r < c->world; r++)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->win[r]);
    if (c->win[c->rank]) cudaFree(c->win[c->rank]);
    delete c;
}

template <typename T>
static int comm_run(pbk_comm *c, const T *src, T *out, int64_t n, int mode, int32_t *status, void *stream, const char *what)
{
    if (!c || !src || !out || !status || n < 0) return pbk_fail(c ? c->ctx : nullptr, PBK_EINVAL, what);
    if ((size_t)n * sizeof(T) > c->slot_bytes) return pbk_fail(c->ctx, PBK_EINVAL, "pbk_comm: payload larger than the window slot");
    for (int r = 0; r < c->world; r++)
        if (!c->win[r]) return pbk_fail(c->ctx, PBK_EINVAL, "pbk_comm: not connected");
    CommView V;
    for (int r = 0; r < PBK_COMM_MAX_WORLD; r++) V.win[r] = r < c->world ? c->win[r] : nullptr;
    c->epoch += 1;
    V.rank = c->rank; V.world = c->world; V.slot_bytes = c->slot_bytes;
    V.half_off = (c->epoch & 1) ? c->half_bytes : 0; V.epoch = c->epoch;
    k_comm_combine<T><<<c->world, 512, 0, (cudaStream_t)stream>>>(V, src, out, n, mode, status);
    PBK_CHECK_LAUNCH(c->ctx, "k_comm_combine");
    return PBK_OK;
}

extern "C" int pbk_comm_prefix_i32(pbk_comm *c, const int32_t *src, int32_t *out, int64_t n, int32_t *status, void *stream)
{
    return comm_run<int32_t>(c, src, out, n, 0, status, stream, "pbk_comm_prefix_i32: bad argument");
}

extern "C" int pbk_allreduce_u64(pbk_comm *c, const unsigned long long *src, unsigned long long *out, int64_t n,
                                 int32_t *status, void *stream)
{
    return comm_run<unsigned long long>(c, src, out, n, 1, status, stream, "pbk_allreduce_u64: bad argument");
}

extern "C" int pbk_allreduce_f64(pbk_comm *c, const double *src, double *out, int64_t n, int32_t *status, void *stream)
{
    return comm_run<double>(c, src, out, n, 1, status, stream, "pbk_allreduce_f64: bad argument");
}
