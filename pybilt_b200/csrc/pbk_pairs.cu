// pbk_pairs.cu -- minimum-image lateral pair work per leaflet: COM lateral RDF histograms and
// nearest-neighbour fractions.
//
// Reference arithmetic restated here (paths relative to the PyBILT tree):
//   distance       pybilt/common/distance_cutoff_clustering.py:27-71
//   com_lateral_rdf pybilt/bilayer_analyzer/analysis_protocols.py:4900-5006
//   nnf            pybilt/bilayer_analyzer/analysis_protocols.py:2174-2257
#include <stdlib.h>

#include "pbk_common.cuh"

#define PAIR_TJ 1024         // lipids staged in shared memory per tile
#define PAIR_THREADS 256

// class byte of a lipid in a frame: bit0 = leaflet label, bit1 = is type A, bit2 = is type B
__device__ __forceinline__ unsigned pbk_cls(uint8_t label, int type, int ta, int tb)
{
    return (unsigned)(label & 1) | ((ta == -1 || type == ta) ? 2u : 0u) | ((tb == -1 || type == tb) ? 4u : 0u);
}

// counts[leaf][0..2] = (N_a, N_b, N_leaflet) of the frame; every thread gets the result
__device__ void pbk_frame_class_counts(const uint8_t *lab, const int32_t *type_code, int64_t N,
                                       int ta, int tb, int *sh_counts /*[6]*/)
{
    if (threadIdx.x < 6) sh_counts[threadIdx.x] = 0;
    __syncthreads();
    int c[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        unsigned cl = pbk_cls(lab[i], type_code[i], ta, tb);
        int leaf = (cl & 1) ? 0 : 1;                 // row 0 = upper (label 1), row 1 = lower
        c[leaf * 3 + 0] += (cl >> 1) & 1;
        c[leaf * 3 + 1] += (cl >> 2) & 1;
        c[leaf * 3 + 2] += 1;
    }
#pragma unroll
    for (int q = 0; q < 6; q++) {
        int v = pbk_warp_sum(c[q]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh_counts[q], v);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// K6  RDF.  grid = (frames, i-tiles); each thread owns one reference lipid i and sweeps all
// j staged through shared memory; bins go to warp-private shared histograms.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PAIR_THREADS)
k_rdf(const double *__restrict__ com, const uint8_t *__restrict__ labels,
      const int32_t *__restrict__ type_code, const float *__restrict__ box, int64_t N, int lat0,
      int lat1, int ta, int tb, int leaflet_mask, int nb, double lo, double hi,
      const double *__restrict__ edges, unsigned long long *__restrict__ count,
      int32_t *__restrict__ frame_counts, int hist_copies, const int *__restrict__ frame_flag,
      int flag_stride)
{
    if (frame_flag && !frame_flag[(int64_t)blockIdx.x * flag_stride]) return;   // only flagged frames
    extern __shared__ unsigned char smem_raw[];
    double *s_edges = reinterpret_cast<double *>(smem_raw);                 // nb+1
    double *s_px = s_edges + (nb + 1);                                      // PAIR_TJ
    double *s_py = s_px + PAIR_TJ;                                          // PAIR_TJ
    unsigned *s_hist = reinterpret_cast<unsigned *>(s_py + PAIR_TJ);        // hist_copies*nb
    unsigned char *s_cls = reinterpret_cast<unsigned char *>(s_hist + (size_t)hist_copies * nb);
    __shared__ int sh_counts[6];

    int64_t f = blockIdx.x;
    const double *cf = com + f * N * 3;
    const uint8_t *lab = labels + f * N;
    float bxf = box[f * 3 + lat0], byf = box[f * 3 + lat1];
    float cxf = bxf * 0.5f, cyf = byf * 0.5f;
    double Lx = (double)bxf, Ly = (double)byf, hx = (double)cxf, hy = (double)cyf;

    for (int q = threadIdx.x; q <= nb; q += blockDim.x) s_edges[q] = edges[q];
    for (int q = threadIdx.x; q < hist_copies * nb; q += blockDim.x) s_hist[q] = 0u;
    pbk_frame_class_counts(lab, type_code, N, ta, tb, sh_counts);
    // a leaflet takes part iff selected and it holds both groups (has_group, :4965)
    bool leaf_on[2];
    for (int l = 0; l < 2; l++)
        leaf_on[l] = ((leaflet_mask >> l) & 1) && sh_counts[l * 3 + 0] > 0 && sh_counts[l * 3 + 1] > 0;
    if (blockIdx.y == 0 && threadIdx.x < 6) {
        int l = threadIdx.x / 3, q = threadIdx.x % 3;
        int v = sh_counts[threadIdx.x];
        if (q < 2 && !leaf_on[l]) v = 0;
        if (q == 2 && !((leaflet_mask >> l) & 1)) v = 0;
        frame_counts[f * 6 + threadIdx.x] = v;
    }

    int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    bool iact = false;
    double apx = 0, apy = 0;
    unsigned icl = 0;
    if (i < N) {
        icl = pbk_cls(lab[i], type_code[i], ta, tb);
        int leaf = (icl & 1) ? 0 : 1;
        iact = (icl & 2u) && leaf_on[leaf];
        apx = __dsub_rn(cf[i * 3 + lat0], hx);
        apy = __dsub_rn(cf[i * 3 + lat1], hy);
    }
    unsigned *myhist = s_hist + (size_t)((threadIdx.x >> 5) % hist_copies) * nb;

    for (int64_t j0 = 0; j0 < N; j0 += PAIR_TJ) {
        int tj = (int)min((int64_t)PAIR_TJ, N - j0);
        __syncthreads();
        for (int q = threadIdx.x; q < tj; q += blockDim.x) {
            int64_t j = j0 + q;
            s_px[q] = __dsub_rn(cf[j * 3 + lat0], hx);
            s_py[q] = __dsub_rn(cf[j * 3 + lat1], hy);
            s_cls[q] = (unsigned char)pbk_cls(lab[j], type_code[j], ta, tb);
        }
        __syncthreads();
        if (iact) {
            for (int q = 0; q < tj; q++) {
                unsigned jc = s_cls[q];
                if (!(jc & 4u) || ((jc ^ icl) & 1u) || (j0 + q) == i) continue;
                double d2 = pbk_pbc_dist2(apx, apy, s_px[q], s_py[q], Lx, Ly, hx, hy);
                double r = sqrt(d2);
                int k = pbk_hist_bin(r, lo, hi, nb, s_edges);
                if (k >= 0) atomicAdd(&myhist[k], 1u);
            }
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nb; q += blockDim.x) {
        unsigned long long t = 0;
        for (int c = 0; c < hist_copies; c++) t += s_hist[(size_t)c * nb + q];
        if (t) atomicAdd(&count[q], t);
    }
}

// dense O(N^2) fallback of pbk_rdf (pbk_rdf2.cu) for leaflets that do not fit in shared memory
int pbk_rdf_dense_flagged(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                          const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                          int leaflet_mask, int32_t n_bins, double range_lo, double range_hi,
                          const double *edges, unsigned long long *count, int32_t *frame_counts,
                          const int *frame_flag, int flag_stride, cudaStream_t stream)
{
    int warps = PAIR_THREADS / 32;
    int copies = warps;
    auto smem_for = [&](int c) {
        return (size_t)(n_bins + 1) * 8 + (size_t)PAIR_TJ * 16 + (size_t)c * n_bins * 4 + PAIR_TJ;
    };
    while (copies > 1 && smem_for(copies) > 96 * 1024) copies >>= 1;
    size_t smem = smem_for(copies);
    if (smem > 200 * 1024) return pbk_fail(ctx, PBK_EINVAL, "pbk_rdf: n_bins too large");
    cudaError_t e = cudaFuncSetAttribute(k_rdf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return pbk_fail(ctx, PBK_ECUDA, "k_rdf smem attribute", e);
    dim3 grid((unsigned)F, (unsigned)((N + PAIR_THREADS - 1) / PAIR_THREADS));
    k_rdf<<<grid, PAIR_THREADS, smem, (cudaStream_t)stream>>>(
        com, labels, type_code, box, N, lat0, lat1, type_a, type_b, leaflet_mask, n_bins, range_lo,
        range_hi, edges, count, frame_counts, copies, frame_flag, flag_stride);
    PBK_CHECK_LAUNCH(ctx, "k_rdf");
    return PBK_OK;
}

int pbk_rdf_dense(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                  const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                  int leaflet_mask, int32_t n_bins, double range_lo, double range_hi,
                  const double *edges, unsigned long long *count, int32_t *frame_counts,
                  cudaStream_t stream)
{
    return pbk_rdf_dense_flagged(ctx, com, labels, type_code, box, F, N, lat0, lat1, type_a, type_b,
                                 leaflet_mask, n_bins, range_lo, range_hi, edges, count, frame_counts,
                                 nullptr, 0, stream);
}

// ---------------------------------------------------------------------------------------
// K7  NNF.  One thread per reference lipid keeps the k best (distance, is-type-b) entries
// in ascending order; a candidate is inserted only if strictly closer than the current
// k-th, and behind entries of equal distance (Python's stable sort over candidates
// visited in ascending lipid index).
// ---------------------------------------------------------------------------------------
#define NNF_KMAX 64

__global__ void __launch_bounds__(PAIR_THREADS)
k_nnf(const double *__restrict__ com, const uint8_t *__restrict__ labels,
      const int32_t *__restrict__ type_code, const float *__restrict__ box, int64_t N, int lat0,
      int lat1, int ta, int tb, int leaflet_mask, int k, int32_t *__restrict__ n_b, int out_stride,
      const int *__restrict__ frame_flag, int flag_stride)
{
    if (frame_flag && !frame_flag[(int64_t)blockIdx.x * flag_stride]) return;   // only flagged frames
    __shared__ double s_px[PAIR_TJ], s_py[PAIR_TJ];
    __shared__ unsigned char s_cls[PAIR_TJ];
    __shared__ int sh_counts[6];
    int64_t f = blockIdx.x;
    const double *cf = com + f * N * 3;
    const uint8_t *lab = labels + f * N;
    float bxf = box[f * 3 + lat0], byf = box[f * 3 + lat1];
    float cxf = bxf * 0.5f, cyf = byf * 0.5f;
    double Lx = (double)bxf, Ly = (double)byf, hx = (double)cxf, hy = (double)cyf;
    pbk_frame_class_counts(lab, type_code, N, ta, tb, sh_counts);
    bool leaf_on[2];
    for (int l = 0; l < 2; l++)
        leaf_on[l] = ((leaflet_mask >> l) & 1) && sh_counts[l * 3 + 0] > 0 && sh_counts[l * 3 + 1] > 0;

    int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    bool iact = false;
    double apx = 0, apy = 0;
    unsigned icl = 0;
    if (i < N) {
        icl = pbk_cls(lab[i], type_code[i], ta, tb);
        iact = (icl & 2u) && leaf_on[(icl & 1) ? 0 : 1];
        apx = __dsub_rn(cf[i * 3 + lat0], hx);
        apy = __dsub_rn(cf[i * 3 + lat1], hy);
    }
    double best[NNF_KMAX];
    unsigned long long bmask = 0ull;      // bit t = entry t is of type b
    int nbest = 0;
    for (int64_t j0 = 0; j0 < N; j0 += PAIR_TJ) {
        int tj = (int)min((int64_t)PAIR_TJ, N - j0);
        __syncthreads();
        for (int q = threadIdx.x; q < tj; q += blockDim.x) {
            int64_t j = j0 + q;
            s_px[q] = __dsub_rn(cf[j * 3 + lat0], hx);
            s_py[q] = __dsub_rn(cf[j * 3 + lat1], hy);
            s_cls[q] = (unsigned char)pbk_cls(lab[j], type_code[j], ta, tb);
        }
        __syncthreads();
        if (iact) {
            for (int q = 0; q < tj; q++) {
                unsigned jc = s_cls[q];
                if (((jc ^ icl) & 1u) || (j0 + q) == i) continue;
                double r = sqrt(pbk_pbc_dist2(apx, apy, s_px[q], s_py[q], Lx, Ly, hx, hy));
                if (nbest == k && !(r < best[k - 1])) continue;
                int pos = nbest < k ? nbest : k - 1;          // slot that is freed / appended
                while (pos > 0 && best[pos - 1] > r) pos--;
                // shift entries [pos, last) one slot up
                int last = nbest < k ? nbest : k - 1;
                for (int t = last; t > pos; t--) best[t] = best[t - 1];
                unsigned long long lowm = (pos == 0) ? 0ull : ((~0ull) >> (64 - pos));
                unsigned long long keep = bmask & lowm;
                unsigned long long up = (bmask & ~lowm) << 1;
                bmask = keep | up | (((jc >> 2) & 1ull) << pos);
                if (k < 64) bmask &= ((1ull << k) - 1ull);
                best[pos] = r;
                if (nbest < k) nbest++;
            }
        }
    }
    if (i < N) n_b[(f * N + i) * out_stride] = iact ? __popcll(bmask) : -1;
}

// Welford mean of n_b/k over the reference lipids of a frame, upper members first
// (analysis_protocols.py:2219-2250).  One warp per frame: coalesced loads, a ballot picks the
// reference lipids, and every lane replays the sequential recurrence on the broadcast values
// (the recurrence itself cannot be reordered without changing the rounding).
__global__ void __launch_bounds__(128)
k_nnf_frame_mean(const int32_t *__restrict__ n_b, const uint8_t *__restrict__ labels,
                 int64_t F, int64_t N, int k, double *__restrict__ frac)
{
    const int64_t f = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (f >= F) return;
    const int32_t *c = n_b + f * N;
    const uint8_t *lab = labels + f * N;
    double m = 0.0;
    int64_t n = 0;
    for (int pass = 1; pass >= 0; pass--) {
        for (int64_t base = 0; base < N; base += 32) {
            const int64_t i = base + lane;
            int ci = -1;
            if (i < N && lab[i] == pass) ci = c[i];
            unsigned mask = __ballot_sync(0xffffffffu, ci >= 0);
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                const double v = __ddiv_rn((double)__shfl_sync(0xffffffffu, ci, src), (double)k);
                n++;
                m = (n == 1) ? v : __dadd_rn(m, __ddiv_rn(__dsub_rn(v, m), (double)n));
            }
        }
    }
    if (lane == 0) frac[f] = m;
}

// Parallel form of the frame mean (default; PBK_SEQ_WELFORD=1 selects the sequential replays): the
// running mean of RunningStats.push equals sum / n in exact arithmetic, and the values are the
// k + 1 quotients c / k, so a lane-strided float64 sum with a fixed combination tree differs from
// the sequential recurrence by rounding noise only (~1e-15 relative).  One CTA per (analysis, frame).
__global__ void __launch_bounds__(256)
k_nnf_frame_mean_par(const int32_t *__restrict__ n_b, const uint8_t *__restrict__ labels, int64_t AF, int64_t F,
                     int64_t N, int k, double *__restrict__ frac)
{
    // one CTA per (analysis, frame); the k + 1 possible quotients c / k come from a table (one division each,
    // not one per lipid); fixed combination tree (pbk_block_sum)
    __shared__ double s_q[65];
    __shared__ double red[32];
    const int64_t af = blockIdx.x;
    const int tid = threadIdx.x;
    const double kd = (double)k;
    if (tid <= 64) s_q[tid] = __ddiv_rn((double)tid, kd);
    __syncthreads();
    const int32_t *c = n_b + af * N;
    double s = 0.0, n = 0.0;
    for (int64_t i = tid; i < N; i += 256) {
        const int ci = c[i];
        if (ci >= 0) { s += ci <= 64 ? s_q[ci] : __ddiv_rn((double)ci, kd); n += 1.0; }
    }
    s = pbk_block_sum(s, red);
    n = pbk_block_sum(n, red);
    if (tid == 0) frac[af] = n > 0.0 ? s / n : 0.0;
}

static bool pbk_seq_welford() { return getenv("PBK_SEQ_WELFORD") != nullptr; }

int pbk_nnf_dense_strided(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                          const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                          int leaflet_mask, int k, int32_t *n_b, int out_stride, const int *frame_flag,
                          int flag_stride, cudaStream_t stream)
{
    dim3 grid((unsigned)F, (unsigned)((N + PAIR_THREADS - 1) / PAIR_THREADS));
    k_nnf<<<grid, PAIR_THREADS, 0, stream>>>(com, labels, type_code, box, N, lat0, lat1, type_a, type_b,
                                             leaflet_mask, k, n_b, out_stride, frame_flag, flag_stride);
    PBK_CHECK_LAUNCH(ctx, "k_nnf");
    return PBK_OK;
}

int pbk_nnf_dense_flagged(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                          const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                          int leaflet_mask, int k, int32_t *n_b, const int *frame_flag, int flag_stride,
                          cudaStream_t stream)
{
    return pbk_nnf_dense_strided(ctx, com, labels, type_code, box, F, N, lat0, lat1, type_a, type_b,
                                 leaflet_mask, k, n_b, 1, frame_flag, flag_stride, stream);
}

// n_b of one (type_a, type_b, leaflets) analysis from the per-type neighbour tallies of
// pbk_nnf_multi: a lipid of type_a in a selected leaflet that holds both types gets the number of
// type_b lipids among its k nearest, everyone else -1.
__global__ void __launch_bounds__(256)
k_nnf_select(const int32_t *__restrict__ hits, const int32_t *__restrict__ type_counts,
             const uint8_t *__restrict__ labels, const int32_t *__restrict__ type_code, int64_t N, int T,
             int ta, int tb, int leaflet_mask, int32_t *__restrict__ n_b)
{
    const int64_t f = blockIdx.x;
    const int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int leaf = labels[f * N + i] ? 0 : 1;
    const int32_t *tc = type_counts + (f * 2 + leaf) * T;
    int ca = 0, cb = 0;
    for (int t = 0; t < T; t++) {
        ca += (ta == -1 || t == ta) ? tc[t] : 0;
        cb += (tb == -1 || t == tb) ? tc[t] : 0;
    }
    const int ti = type_code[i];
    const bool ref = ((leaflet_mask >> leaf) & 1) && ca > 0 && cb > 0 && (ta == -1 || ti == ta);
    int c = -1;
    if (ref) {
        c = 0;
        const int32_t *h = hits + (f * N + i) * T;
        for (int t = 0; t < T; t++)
            if (tb == -1 || t == tb) c += max(0, h[t]);
    }
    n_b[f * N + i] = c;
}

// the same for A analyses in one launch each: specs[a] = (type_a, type_b, leaflet_mask); outputs
// n_b[A][F][N], frac[A][F].  One grid over (analysis, frame) keeps the sequential frame means of all
// analyses running side by side instead of one launch latency after the other.
__global__ void __launch_bounds__(256)
k_nnf_select_many(const int32_t *__restrict__ hits, const int32_t *__restrict__ type_counts,
                  const uint8_t *__restrict__ labels, const int32_t *__restrict__ type_code, int64_t F, int64_t N,
                  int T, const int32_t *__restrict__ specs, int32_t *__restrict__ n_b)
{
    const int64_t af = blockIdx.x;                    // a * F + f
    const int64_t a = af / F, f = af - a * F;
    const int ta = specs[a * 3], tb = specs[a * 3 + 1], leaflet_mask = specs[a * 3 + 2];
    const int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int leaf = labels[f * N + i] ? 0 : 1;
    const int32_t *tc = type_counts + (f * 2 + leaf) * T;
    int ca = 0, cb = 0;
    for (int t = 0; t < T; t++) {
        ca += (ta == -1 || t == ta) ? tc[t] : 0;
        cb += (tb == -1 || t == tb) ? tc[t] : 0;
    }
    const int ti = type_code[i];
    const bool ref = ((leaflet_mask >> leaf) & 1) && ca > 0 && cb > 0 && (ta == -1 || ti == ta);
    int c = -1;
    if (ref) {
        c = 0;
        const int32_t *h = hits + (f * N + i) * T;
        for (int t = 0; t < T; t++)
            if (tb == -1 || t == tb) c += max(0, h[t]);
    }
    n_b[af * N + i] = c;
}

__global__ void __launch_bounds__(128)
k_nnf_frame_mean_many(const int32_t *__restrict__ n_b, const uint8_t *__restrict__ labels, int64_t A, int64_t F,
                      int64_t N, int k, double *__restrict__ frac)
{
    const int64_t af = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (af >= A * F) return;
    const int64_t f = af % F;
    const int32_t *c = n_b + af * N;
    const uint8_t *lab = labels + f * N;
    double m = 0.0;
    int64_t n = 0;
    for (int pass = 1; pass >= 0; pass--) {
        for (int64_t base = 0; base < N; base += 32) {
            const int64_t i = base + lane;
            int ci = -1;
            if (i < N && lab[i] == pass) ci = c[i];
            unsigned mask = __ballot_sync(0xffffffffu, ci >= 0);
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                const double v = __ddiv_rn((double)__shfl_sync(0xffffffffu, ci, src), (double)k);
                n++;
                m = (n == 1) ? v : __dadd_rn(m, __ddiv_rn(__dsub_rn(v, m), (double)n));
            }
        }
    }
    if (lane == 0) frac[af] = m;
}

extern "C" int pbk_nnf_select_many(pbk_ctx *ctx, const int32_t *type_hits, const int32_t *type_counts,
                                   const uint8_t *labels, const int32_t *type_code, int64_t F, int64_t N,
                                   int n_types, int n_specs, const int32_t *specs, int k, int32_t *n_b,
                                   double *frac, void *stream)
{
    if (!ctx || !type_hits || !type_counts || !labels || !type_code || !specs || !n_b || !frac || F < 0 ||
        N <= 0 || n_types < 1 || n_specs < 1 || k < 1 || (int64_t)n_specs * F > 0x7fffffffLL)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_nnf_select_many: bad argument");
    if (F == 0) return PBK_OK;
    dim3 g((unsigned)((int64_t)n_specs * F), (unsigned)((N + 255) / 256));
    k_nnf_select_many<<<g, 256, 0, (cudaStream_t)stream>>>(type_hits, type_counts, labels, type_code, F, N,
                                                           n_types, specs, n_b);
    PBK_CHECK_LAUNCH(ctx, "k_nnf_select_many");
    const int64_t n_af = (int64_t)n_specs * F;
    if (!pbk_seq_welford()) {
        k_nnf_frame_mean_par<<<(unsigned)n_af, 256, 0, (cudaStream_t)stream>>>(n_b, labels, n_af, F, N, k, frac);
        PBK_CHECK_LAUNCH(ctx, "k_nnf_frame_mean_par");
        return PBK_OK;
    }
    k_nnf_frame_mean_many<<<(unsigned)((n_af + 3) / 4), 128, 0, (cudaStream_t)stream>>>(n_b, labels, n_specs, F, N,
                                                                                         k, frac);
    PBK_CHECK_LAUNCH(ctx, "k_nnf_frame_mean_many");
    return PBK_OK;
}

extern "C" int pbk_nnf_select(pbk_ctx *ctx, const int32_t *type_hits, const int32_t *type_counts,
                              const uint8_t *labels, const int32_t *type_code, int64_t F, int64_t N,
                              int n_types, int type_a, int type_b, int leaflet_mask, int k, int32_t *n_b,
                              double *frac, void *stream)
{
    if (!ctx || !type_hits || !type_counts || !labels || !type_code || !n_b || !frac || F < 0 || N <= 0 ||
        n_types < 1 || k < 1 || (leaflet_mask & ~3))
        return pbk_fail(ctx, PBK_EINVAL, "pbk_nnf_select: bad argument");
    if (F == 0) return PBK_OK;
    dim3 g((unsigned)F, (unsigned)((N + 255) / 256));
    k_nnf_select<<<g, 256, 0, (cudaStream_t)stream>>>(type_hits, type_counts, labels, type_code, N, n_types,
                                                      type_a, type_b, leaflet_mask, n_b);
    PBK_CHECK_LAUNCH(ctx, "k_nnf_select");
    if (pbk_seq_welford()) k_nnf_frame_mean<<<(unsigned)((F + 3) / 4), 128, 0, (cudaStream_t)stream>>>(n_b, labels, F, N, k, frac);
    else k_nnf_frame_mean_par<<<(unsigned)F, 256, 0, (cudaStream_t)stream>>>(n_b, labels, F, F, N, k, frac);
    PBK_CHECK_LAUNCH(ctx, "k_nnf_frame_mean");
    return PBK_OK;
}

// defined in pbk_cells.cu: cell-list search for large systems
int pbk_nnf_cells(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                  const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
                  int leaflet_mask, int k, int32_t *n_b, cudaStream_t st);
bool pbk_use_cells(int64_t N, int64_t threshold);

extern "C" int pbk_nnf(pbk_ctx *ctx, const double *com, const uint8_t *labels,
                       const int32_t *type_code, const float *box, int64_t F, int64_t N, int lat0,
                       int lat1, int type_a, int type_b, int leaflet_mask, int k, int32_t *n_b,
                       double *frac, void *stream)
{
    if (!ctx || !com || !labels || !type_code || !box || !n_b || !frac || F < 0 || N <= 0 ||
        k < 1 || k > NNF_KMAX || (leaflet_mask & ~3) || lat0 < 0 || lat0 > 2 || lat1 < 0 || lat1 > 2)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_nnf: bad argument (1 <= k <= 64)");
    if (F == 0) return PBK_OK;
    int rc = PBK_EUNSUPPORTED;
    if (pbk_use_cells(N, 1024))
        rc = pbk_nnf_cells(ctx, com, labels, type_code, box, F, N, lat0, lat1, type_a, type_b, leaflet_mask,
                           k, n_b, (cudaStream_t)stream);
    if (rc == PBK_EUNSUPPORTED)
        rc = pbk_nnf_dense_flagged(ctx, com, labels, type_code, box, F, N, lat0, lat1, type_a, type_b,
                                   leaflet_mask, k, n_b, nullptr, 0, (cudaStream_t)stream);
    if (rc) return rc;
    if (pbk_seq_welford()) k_nnf_frame_mean<<<(unsigned)((F + 3) / 4), 128, 0, (cudaStream_t)stream>>>(n_b, labels, F, N, k, frac);
    else k_nnf_frame_mean_par<<<(unsigned)F, 256, 0, (cudaStream_t)stream>>>(n_b, labels, F, F, N, k, frac);
    PBK_CHECK_LAUNCH(ctx, "k_nnf_frame_mean");
    return PBK_OK;
}
