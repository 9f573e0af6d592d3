// pbk_ctx.cu -- context handling and error reporting of libpbk.
#include <stdio.h>
#include <string.h>

#include "pbk_common.cuh"

int pbk_fail(pbk_ctx *ctx, int code, const char *what, cudaError_t e)
{
    if (ctx) {
        if (e != cudaSuccess)
            snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, cudaGetErrorString(e));
        else
            snprintf(ctx->err, sizeof(ctx->err), "%s", what);
    }
    return code;
}

extern "C" int pbk_abi_version(void) { return PBK_ABI_VERSION; }

extern "C" int pbk_ctx_create(int device, pbk_ctx **out)
{
    if (!out) return PBK_EINVAL;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) return PBK_ECUDA;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return PBK_ECUDA;
    pbk_ctx *c = new pbk_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->err[0] = 0;
    c->work = nullptr;
    c->work_next = 0;
    c->lut = nullptr;
    c->arena = nullptr;
    c->arena_bytes = 0;
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess) {
        if (cur != device) cudaSetDevice(device);
        if (cudaMalloc(&c->work, PBK_WORK_SLOTS * sizeof(int)) != cudaSuccess) { c->work = nullptr; cudaGetLastError(); }
        if (cudaMalloc(&c->lut, (size_t)PBK_LUT_SLOTS * PBK_LUT_ENTRIES * sizeof(unsigned short)) != cudaSuccess) { c->lut = nullptr; cudaGetLastError(); }
        if (cur != device) cudaSetDevice(cur);
    }
    *out = c;
    return PBK_OK;
}

extern "C" void pbk_ctx_destroy(pbk_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->work) cudaFree(ctx->work);
    if (ctx->lut) cudaFree(ctx->lut);
    if (ctx->arena) cudaFree(ctx->arena);
    delete ctx;
}

extern "C" const char *pbk_last_error(const pbk_ctx *ctx) { return ctx ? ctx->err : "null context"; }

extern "C" int pbk_sm_count(const pbk_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

void *pbk_arena(pbk_ctx *ctx, size_t bytes)
{
    if (!ctx) return nullptr;
    if (bytes <= ctx->arena_bytes) return ctx->arena;
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != ctx->device) cudaSetDevice(ctx->device);
    if (ctx->arena) { cudaDeviceSynchronize(); cudaFree(ctx->arena); }
    ctx->arena = nullptr;
    ctx->arena_bytes = 0;
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) == cudaSuccess) { ctx->arena = p; ctx->arena_bytes = bytes; }
    else cudaGetLastError();
    if (cur != ctx->device && cur >= 0) cudaSetDevice(cur);
    return ctx->arena;
}
