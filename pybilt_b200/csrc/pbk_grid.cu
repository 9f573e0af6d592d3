// pbk_grid.cu -- GridMAT-style lipid grids: nearest-lipid assignment per grid cell and the
// area-per-lipid / thickness reductions.
//
// Reference arithmetic restated here (paths relative to the PyBILT tree):
//   exact grid     pybilt/lipid_grid/lipid_grid.py:150-248
//   heuristic grid pybilt/lipid_grid/lipid_grid_opt.py:155-334 (+ _knn :435-468)
//   reductions     pybilt/lipid_grid/lipid_grid_opt.py:496-526, 553-597
#include <stdlib.h>

#include "pbk_common.cuh"

#define GRID_NNK 24
#define GRID_TJ 1024

// cell centre j of an n-bin grid over [0, L]: float32 linspace arithmetic of NumPy >= 2
// (lipid_grid_opt.py:204-215): edges = f32(j)*step, step = L/n, centre = edges[j] + step/2.
__device__ __forceinline__ double pbk_cell_centre(int j, float step)
{
    return (double)__fadd_rn(__fmul_rn((float)j, step), step * 0.5f);
}

// cell-to-lipid distance with the reference's quirk: the dy test uses bxh
// (lipid_grid_opt.py:259,303,323).  Squares are exact products (DESIGN.md "grid squares").
__device__ __forceinline__ double pbk_cell_dist(double cx, double cy, double xi, double yi,
                                                double Lx, double Ly, double hx, double hy)
{
    double dx = __dsub_rn(cx, xi), dy = __dsub_rn(cy, yi);
    if (fabs(dx) > hx) dx = __dsub_rn(__dsub_rn(Lx, fabs(__dsub_rn(cx, hx))), fabs(__dsub_rn(xi, hx)));
    if (fabs(dy) > hx) dy = __dsub_rn(__dsub_rn(Ly, fabs(__dsub_rn(cy, hy))), fabs(__dsub_rn(yi, hy)));
    return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// ---------------------------------------------------------------------------------------
// K8a  24 nearest same-leaflet neighbours of every lipid (stable order).  The reference
// evaluates each unordered pair once as dist(lower member, higher member).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_grid_knn(const double *__restrict__ com, const uint8_t *__restrict__ labels,
           const float *__restrict__ box, int64_t N, int lat0, int lat1, int32_t *__restrict__ knn,
           const int *__restrict__ frame_flag, int flag_stride)
{
    if (frame_flag && !frame_flag[(int64_t)blockIdx.x * flag_stride]) return;   // only flagged frames
    __shared__ double s_px[GRID_TJ], s_py[GRID_TJ];
    __shared__ unsigned char s_lab[GRID_TJ];
    int64_t f = blockIdx.x;
    const double *cf = com + f * N * 3;
    const uint8_t *lab = labels + f * N;
    float bxf = box[f * 3 + lat0], byf = box[f * 3 + lat1];
    float cxf = bxf * 0.5f, cyf = byf * 0.5f;
    double Lx = (double)bxf, Ly = (double)byf, hx = (double)cxf, hy = (double)cyf;
    int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    double apx = 0, apy = 0;
    unsigned il = 0;
    if (i < N) {
        il = lab[i];
        apx = __dsub_rn(cf[i * 3 + lat0], hx);
        apy = __dsub_rn(cf[i * 3 + lat1], hy);
    }
    double best[GRID_NNK];
    int bj[GRID_NNK];
    int nbest = 0;
    for (int64_t j0 = 0; j0 < N; j0 += GRID_TJ) {
        int tj = (int)min((int64_t)GRID_TJ, N - j0);
        __syncthreads();
        for (int q = threadIdx.x; q < tj; q += blockDim.x) {
            int64_t j = j0 + q;
            s_px[q] = __dsub_rn(cf[j * 3 + lat0], hx);
            s_py[q] = __dsub_rn(cf[j * 3 + lat1], hy);
            s_lab[q] = lab[j];
        }
        __syncthreads();
        if (i < N) {
            for (int q = 0; q < tj; q++) {
                int64_t j = j0 + q;
                if (s_lab[q] != il || j == i) continue;
                double d2 = (j > i) ? pbk_pbc_dist2(apx, apy, s_px[q], s_py[q], Lx, Ly, hx, hy)
                                    : pbk_pbc_dist2(s_px[q], s_py[q], apx, apy, Lx, Ly, hx, hy);
                double r = sqrt(d2);
                if (nbest == GRID_NNK && !(r < best[GRID_NNK - 1])) continue;
                int pos = nbest < GRID_NNK ? nbest : GRID_NNK - 1;
                int last = pos;
                while (pos > 0 && best[pos - 1] > r) pos--;
                for (int t = last; t > pos; t--) { best[t] = best[t - 1]; bj[t] = bj[t - 1]; }
                best[pos] = r;
                bj[pos] = (int)j;
                if (nbest < GRID_NNK) nbest++;
            }
        }
    }
    if (i < N) {
        int32_t *o = knn + (f * N + i) * GRID_NNK;
        for (int t = 0; t < GRID_NNK; t++) o[t] = t < nbest ? bj[t] : -1;
    }
}

// ---------------------------------------------------------------------------------------
// K8b  cell assignment, one CTA per (frame, leaflet).
//   exact: every cell scans the leaflet members in ascending order, strict '<'.
//   opt  : cell (0,0) exact; row 0 is a chain (0,cy) <- (0,cy-1); every later row only
//          depends on the row above: (cx,cy) <- (cx-1, max(0,cy-1)), candidates p then
//          knn24(p), first minimum wins.
// ---------------------------------------------------------------------------------------
__device__ void pbk_exact_cell(const double *cf, const uint8_t *lab, int64_t N, int lat0, int lat1,
                               unsigned want, double x, double y, double Lx, double Ly, double hx,
                               double hy, double *rmin_out, int *imin_out)
{
    // warp-cooperative scan; the result is valid in every lane
    int lane = threadIdx.x & 31;
    double rmin = 1.0e10;
    int imin = 0x7fffffff;
    for (int64_t i = lane; i < N; i += 32) {
        if (lab[i] != want) continue;
        double r = pbk_cell_dist(x, y, cf[i * 3 + lat0], cf[i * 3 + lat1], Lx, Ly, hx, hy);
        if (r < rmin) { rmin = r; imin = (int)i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double r2 = __shfl_xor_sync(0xffffffffu, rmin, o);
        int i2 = __shfl_xor_sync(0xffffffffu, imin, o);
        if (r2 < rmin || (r2 == rmin && i2 < imin)) { rmin = r2; imin = i2; }
    }
    *rmin_out = rmin;
    *imin_out = (imin == 0x7fffffff) ? 0 : imin;   // empty leaflet: i_min stays 0 (:244)
}

__device__ __forceinline__ int pbk_opt_cell(const double *cf, const int32_t *knn_f, int lat0, int lat1,
                                            int p, double x, double y, double Lx, double Ly,
                                            double hx, double hy)
{
    double rmin = 1.0e10;
    int imin = p;
    double r = pbk_cell_dist(x, y, cf[(int64_t)p * 3 + lat0], cf[(int64_t)p * 3 + lat1], Lx, Ly, hx, hy);
    if (r < rmin) { rmin = r; imin = p; }
    const int32_t *row = knn_f + (int64_t)p * GRID_NNK;
    for (int t = 0; t < GRID_NNK; t++) {
        int q = row[t];
        if (q < 0) break;
        r = pbk_cell_dist(x, y, cf[(int64_t)q * 3 + lat0], cf[(int64_t)q * 3 + lat1], Lx, Ly, hx, hy);
        if (r < rmin) { rmin = r; imin = q; }
    }
    return imin;
}

__global__ void __launch_bounds__(128)
k_grid_assign(const double *__restrict__ com, const double *__restrict__ com_unwrap,
              const uint8_t *__restrict__ labels, const float *__restrict__ box, int64_t N,
              int lat0, int lat1, int norm, int nx, int ny, int mode,
              const int32_t *__restrict__ knn, int32_t *__restrict__ grid_idx,
              double *__restrict__ grid_z)
{
    int64_t f = blockIdx.x;
    int leaf = blockIdx.y;                       // 0 = upper (label 1), 1 = lower (label 0)
    unsigned want = leaf == 0 ? 1u : 0u;
    const double *cf = com + f * N * 3;
    const double *cu = com_unwrap + f * N * 3;
    const uint8_t *lab = labels + f * N;
    const int32_t *knn_f = knn ? knn + f * N * GRID_NNK : nullptr;
    float bxf = box[f * 3 + lat0], byf = box[f * 3 + lat1];
    float cxf = bxf * 0.5f, cyf = byf * 0.5f;
    double Lx = (double)bxf, Ly = (double)byf, hx = (double)cxf, hy = (double)cyf;
    float stepx = __fdiv_rn(bxf, (float)nx), stepy = __fdiv_rn(byf, (float)ny);
    int32_t *gi = grid_idx + (f * 2 + leaf) * (int64_t)nx * ny;
    double *gz = grid_z + (f * 2 + leaf) * (int64_t)nx * ny;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    // y_centers are only filled for the first min(nx, ny) entries (lipid_grid_opt.py:225)
    auto ycentre = [&](int cy) { return cy < nx ? pbk_cell_centre(cy, stepy) : 0.0; };

    if (mode == 0) {
        for (int c = warp; c < nx * ny; c += nwarp) {
            int cx = c / ny, cy = c % ny;
            double r;
            int im;
            pbk_exact_cell(cf, lab, N, lat0, lat1, want, pbk_cell_centre(cx, stepx), ycentre(cy), Lx,
                           Ly, hx, hy, &r, &im);
            if (lane == 0) { gi[c] = im; gz[c] = cu[(int64_t)im * 3 + norm]; }
        }
        return;
    }
    // ---- opt ----
    if (warp == 0) {
        double r;
        int im;
        pbk_exact_cell(cf, lab, N, lat0, lat1, want, pbk_cell_centre(0, stepx), ycentre(0), Lx, Ly,
                       hx, hy, &r, &im);
        if (lane == 0) {
            gi[0] = im;
            int p = im;
            for (int cy = 1; cy < ny; cy++) {          // row 0: (0,cy) <- (0,cy-1)
                p = pbk_opt_cell(cf, knn_f, lat0, lat1, p, pbk_cell_centre(0, stepx), ycentre(cy), Lx,
                                 Ly, hx, hy);
                gi[cy] = p;
            }
        }
    }
    __syncthreads();
    for (int cx = 1; cx < nx; cx++) {
        double x = pbk_cell_centre(cx, stepx);
        for (int cy = threadIdx.x; cy < ny; cy += blockDim.x) {
            int p = gi[(cx - 1) * ny + (cy > 0 ? cy - 1 : 0)];
            gi[cx * ny + cy] = pbk_opt_cell(cf, knn_f, lat0, lat1, p, x, ycentre(cy), Lx, Ly, hx, hy);
        }
        __syncthreads();
    }
    for (int c = threadIdx.x; c < nx * ny; c += blockDim.x) gz[c] = cu[(int64_t)gi[c] * 3 + norm];
}

int pbk_grid_knn_dense_flagged(pbk_ctx *ctx, const double *com, const uint8_t *labels, const float *box,
                               int64_t F, int64_t N, int lat0, int lat1, int32_t *knn, const int *frame_flag,
                               int flag_stride, cudaStream_t stream)
{
    dim3 g((unsigned)F, (unsigned)((N + 255) / 256));
    k_grid_knn<<<g, 256, 0, stream>>>(com, labels, box, N, lat0, lat1, knn, frame_flag, flag_stride);
    PBK_CHECK_LAUNCH(ctx, "k_grid_knn");
    return PBK_OK;
}

// defined in pbk_cells.cu: cell-list search for large systems
int pbk_grid_knn_cells(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code_or_null,
                       const float *box, int64_t F, int64_t N, int lat0, int lat1, int32_t *knn, cudaStream_t st);
bool pbk_use_cells(int64_t N, int64_t threshold);

// the 24 nearest same-class neighbours of every lipid (class = its label byte): cell lists for large
// systems, the dense kernel otherwise and for frames with a lipid outside the box
extern "C" int pbk_knn24(pbk_ctx *ctx, const double *com, const uint8_t *labels, const float *box, int64_t F,
                         int64_t N, int lat0, int lat1, int32_t *knn, void *stream)
{
    if (!ctx || !com || !labels || !box || !knn || F < 0 || N <= 0 || lat0 < 0 || lat0 > 2 || lat1 < 0 ||
        lat1 > 2)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_knn24: bad argument");
    if (F == 0) return PBK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = PBK_EUNSUPPORTED;
    if (pbk_use_cells(N, 1024))
        rc = pbk_grid_knn_cells(ctx, com, labels, nullptr, box, F, N, lat0, lat1, knn, st);
    if (rc == PBK_EUNSUPPORTED)
        rc = pbk_grid_knn_dense_flagged(ctx, com, labels, box, F, N, lat0, lat1, knn, nullptr, 0, st);
    return rc;
}

// ---------------------------------------------------------------------------------------
// K11  Halperin-Nelson hexatic invariant (halperin_nelson).  Per member lipid l: its six nearest
// members (first six entries of the 24-neighbour table, already in (distance, index) order),
// a = x component of the unit separation vector in the reference's rounding (the pair is
// evaluated as dist(lower index, higher index), the vector handed to l points from l to j),
// phi_l = |(1/6) sum exp(6 i a)|^2; out[f] = mean over the members.  One CTA per frame.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_halperin_nelson(const double *__restrict__ com, const uint8_t *__restrict__ member,
                  const float *__restrict__ box, const int32_t *__restrict__ knn, int64_t N, int lat0, int lat1,
                  int64_t n_members, double *__restrict__ out)
{
    __shared__ double red[32];
    const int64_t f = blockIdx.x;
    const double *cf = com + f * N * 3;
    const float bxf = box[f * 3 + lat0], byf = box[f * 3 + lat1];
    const float cxf = bxf * 0.5f, cyf = byf * 0.5f;
    const double Lx = (double)bxf, Ly = (double)byf, hx = (double)cxf, hy = (double)cyf;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        if (!member[i]) continue;
        const double ix = __dsub_rn(cf[i * 3 + lat0], hx), iy = __dsub_rn(cf[i * 3 + lat1], hy);
        const int32_t *row = knn + (f * N + i) * GRID_NNK;
        double re = 0.0, im = 0.0;
        for (int t = 0; t < 6; t++) {
            const int j = row[t];
            if (j < 0) break;
            const double jx = __dsub_rn(cf[(int64_t)j * 3 + lat0], hx), jy = __dsub_rn(cf[(int64_t)j * 3 + lat1], hy);
            // pair evaluated as (a = lower index, b = higher index): d = a' - b', minimum image per axis
            const bool i_low = i < j;
            const double ax = i_low ? ix : jx, ay = i_low ? iy : jy, bx = i_low ? jx : ix, by = i_low ? jy : iy;
            double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by);
            if (fabs(dx) > hx) dx = __dsub_rn(__dsub_rn(Lx, fabs(ax)), fabs(bx));
            if (fabs(dy) > hy) dy = __dsub_rn(__dsub_rn(Ly, fabs(ay)), fabs(by));
            const double r = sqrt(fma(dy, dy, __dmul_rn(dx, dx)));
            const double vx = i_low ? -dx : dx;            // vector stored for lipid i
            const double a6 = __dmul_rn(6.0, __ddiv_rn(vx, r));
            double sn, cs;
            sincos(a6, &sn, &cs);
            re = __dadd_rn(re, cs);
            im = __dadd_rn(im, sn);
        }
        const double h = hypot(__ddiv_rn(re, 6.0), __ddiv_rn(im, 6.0));
        acc += __dmul_rn(h, h);
    }
    acc = pbk_block_sum(acc, red);
    if (threadIdx.x == 0) out[f] = n_members > 0 ? acc / (double)n_members : nan("");
}

extern "C" int pbk_halperin_nelson(pbk_ctx *ctx, const double *com, const uint8_t *member, const float *box,
                                   const int32_t *knn, int64_t F, int64_t N, int lat0, int lat1, int64_t n_members,
                                   double *out, void *stream)
{
    if (!ctx || !com || !member || !box || !knn || !out || F < 0 || N <= 0 || lat0 < 0 || lat0 > 2 ||
        lat1 < 0 || lat1 > 2)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_halperin_nelson: bad argument");
    if (F == 0) return PBK_OK;
    k_halperin_nelson<<<(unsigned)F, 256, 0, (cudaStream_t)stream>>>(com, member, box, knn, N, lat0, lat1, n_members, out);
    PBK_CHECK_LAUNCH(ctx, "k_halperin_nelson");
    return PBK_OK;
}

extern "C" int pbk_lipid_grid(pbk_ctx *ctx, const double *com, const double *com_unwrap,
                              const uint8_t *labels, const float *box, int64_t F, int64_t N,
                              int lat0, int lat1, int norm, int nx, int ny, int mode,
                              int32_t *knn_scratch, int32_t *grid_idx, double *grid_z, void *stream)
{
    if (!ctx || !com || !com_unwrap || !labels || !box || !grid_idx || !grid_z || F < 0 || N <= 0 ||
        nx <= 0 || ny <= 0 || (mode != 0 && mode != 1) || (mode == 1 && !knn_scratch) ||
        lat0 < 0 || lat0 > 2 || lat1 < 0 || lat1 > 2 || norm < 0 || norm > 2)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_lipid_grid: bad argument");
    if (F == 0) return PBK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 1) {
        const int rc = pbk_knn24(ctx, com, labels, box, F, N, lat0, lat1, knn_scratch, stream);
        if (rc) return rc;
    }
    dim3 g2((unsigned)F, 2);
    k_grid_assign<<<g2, 128, 0, st>>>(com, com_unwrap, labels, box, N, lat0, lat1, norm, nx, ny,
                                      mode, mode == 1 ? knn_scratch : nullptr, grid_idx, grid_z);
    PBK_CHECK_LAUNCH(ctx, "k_grid_assign");
    return PBK_OK;
}

// ---------------------------------------------------------------------------------------
// K9  reductions.  k_grid_reduce, one CTA per frame: cells per lipid, thickness mean /
// population std.  k_grid_apl, one thread per (frame, Welford state): replays the reference's
// pushes (float32 areas, first push keeps float32; pybilt/common/running_stats.py:34-52) over
// upper then lower members -- state 0 is the composite mean, state 1+t the lipids of type t.
// The recurrence is sequential by nature; one thread per state keeps it in registers and
// lets all frames' replays run side by side.
// ---------------------------------------------------------------------------------------
#define GRID_MAX_TYPES 16

__global__ void __launch_bounds__(256)
k_grid_reduce(const int32_t *__restrict__ grid_idx, const double *__restrict__ grid_z,
              int64_t N, int nx, int ny, int32_t *__restrict__ cells, double *__restrict__ thickness)
{
    __shared__ double red[32];
    int64_t f = blockIdx.x;
    int64_t nc = (int64_t)nx * ny;
    const int32_t *gu = grid_idx + f * 2 * nc, *gl = gu + nc;
    const double *zu = grid_z + f * 2 * nc, *zl = zu + nc;
    int32_t *cnt = cells + f * N;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    double s = 0.0;
    for (int64_t c = threadIdx.x; c < nc; c += blockDim.x) {
        atomicAdd(&cnt[gu[c]], 1);
        atomicAdd(&cnt[gl[c]], 1);
        s += __dsub_rn(zu[c], zl[c]);
    }
    s = pbk_block_sum(s, red);
    double mean = s / (double)nc;
    double q = 0.0;
    for (int64_t c = threadIdx.x; c < nc; c += blockDim.x) {
        double d = __dsub_rn(__dsub_rn(zu[c], zl[c]), mean);
        q += d * d;
    }
    q = pbk_block_sum(q, red);
    if (threadIdx.x != 0) return;
    thickness[f * 2] = mean;
    thickness[f * 2 + 1] = sqrt(q / (double)nc);
}

__global__ void __launch_bounds__(64)
k_grid_apl(const int32_t *__restrict__ cells, const uint8_t *__restrict__ labels,
           const int32_t *__restrict__ type_code, const float *__restrict__ box, int64_t F, int64_t N,
           int lat0, int lat1, int nx, int ny, int n_types, double *__restrict__ apl)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n_states = n_types + 1;
    if (gid >= F * n_states) return;
    // consecutive threads = consecutive frames of one state (no divergence on the state test)
    const int t = (int)(gid / F);
    const int64_t f = gid - (int64_t)t * F;
    const int32_t *cnt = cells + f * N;
    const uint8_t *lab = labels + f * N;
    const float stepx = __fdiv_rn(box[f * 3 + lat0], (float)nx), stepy = __fdiv_rn(box[f * 3 + lat1], (float)ny);
    const float apb = __fmul_rn(stepx, stepy);                  // x_incr*y_incr, float32
    double M = 0.0, S = 0.0;
    float M32 = 0.0f;
    long long n = 0;
    for (int pass = 1; pass >= 0; pass--) {
        // the recurrence is sequential, its inputs are not: eight lipids' label / type / cell count
        // are loaded together (independent loads), then pushed one after the other
        for (int64_t i0 = 0; i0 < N; i0 += 8) {
            int use[8], cn[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int64_t i = i0 + u;
                use[u] = 0;
                cn[u] = 0;
                if (i < N) {
                    use[u] = (lab[i] == pass) && (t == 0 || type_code[i] == t - 1);
                    cn[u] = cnt[i];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (!use[u]) continue;
                const float a = __fmul_rn(apb, (float)cn[u]);   // area_per_bin*nlocs, float32
                n++;
                if (n == 1) {
                    M32 = a; M = (double)a; S = 0.0;
                } else {
                    const double diff = (n == 2) ? (double)__fsub_rn(a, M32) : __dsub_rn((double)a, M);
                    const double mn = __dadd_rn(M, __ddiv_rn(diff, (double)n));
                    S = __dadd_rn(S, __dmul_rn(diff, __dsub_rn((double)a, mn)));
                    M = mn;
                }
            }
        }
    }
    double *o = apl + f * (1 + 2 * n_types);
    if (t == 0) {
        o[0] = n ? M : 0.0;
    } else {
        o[2 * t - 1] = n ? M : nan("");
        o[2 * t] = n > 1 ? sqrt(__ddiv_rn(S, (double)(n - 1))) : (n ? 0.0 : nan(""));
    }
}

// Parallel form of the same statistics (default; PBK_SEQ_WELFORD=1 selects the sequential replay
// above, which the tests keep as the bit-exact restatement).  One CTA per (frame, state).
// The reference's first two pushes are replayed literally (the second one subtracts in float32,
// running_stats.py:34-52 with NumPy float32 inputs); the other n - 2 samples are reduced in
// parallel -- their sum in float64 is exact (float32 values of similar magnitude), the squared
// deviations are summed about their own mean -- and merged with Chan's pairwise update, which is
// what the sequential recurrence computes in exact arithmetic.  Differs from the sequential
// replay by its accumulated rounding only (~1e-14 relative at 65,536 lipids).
__global__ void __launch_bounds__(256)
k_grid_apl_par(const int32_t *__restrict__ cells, const uint8_t *__restrict__ labels,
               const int32_t *__restrict__ type_code, const float *__restrict__ box, int64_t N,
               int lat0, int lat1, int nx, int ny, int n_types, double *__restrict__ apl)
{
    __shared__ double red[32];
    __shared__ long long s_k[2];
    __shared__ double s_v[4];
    const int64_t f = blockIdx.x;
    const int t = blockIdx.y;
    const int32_t *cnt = cells + f * N;
    const uint8_t *lab = labels + f * N;
    const float stepx = __fdiv_rn(box[f * 3 + lat0], (float)nx), stepy = __fdiv_rn(box[f * 3 + lat1], (float)ny);
    const float apb = __fmul_rn(stepx, stepy);                  // x_incr*y_incr, float32
    const long long INF = 4 * (long long)N;
    // push order: upper members by index, then lower members by index -> key
    long long k1 = INF, k2 = INF, n_loc = 0;
    double sum = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        if (t != 0 && type_code[i] != t - 1) continue;
        const long long key = (lab[i] == 1 ? 0 : (long long)N) + i;
        const double a = (double)__fmul_rn(apb, (float)cnt[i]);
        n_loc++;
        sum += a;
        if (key < k1) { k2 = k1; k1 = key; }
        else if (key < k2) k2 = key;
    }
    sum = pbk_block_sum(sum, red);
    const long long n = (long long)(pbk_block_sum((double)n_loc, red) + 0.5);
    // the two smallest keys of the block
    if (threadIdx.x == 0) { s_k[0] = INF; s_k[1] = INF; }
    __syncthreads();
    atomicMin((unsigned long long *)&s_k[0], (unsigned long long)k1);
    __syncthreads();
    const long long m1 = s_k[0];
    atomicMin((unsigned long long *)&s_k[1], (unsigned long long)(k1 == m1 ? k2 : k1));
    __syncthreads();
    const long long m2 = s_k[1];
    double *o = apl + f * (1 + 2 * n_types);
    if (n == 0) {
        if (threadIdx.x == 0) {
            if (t == 0) o[0] = 0.0;
            else { o[2 * t - 1] = nan(""); o[2 * t] = nan(""); }
        }
        return;
    }
    if (threadIdx.x == 0) {
        const float a1 = __fmul_rn(apb, (float)cnt[m1 % N]);
        double M = (double)a1, S = 0.0, a2d = 0.0;
        if (n >= 2) {
            const float a2 = __fmul_rn(apb, (float)cnt[m2 % N]);
            a2d = (double)a2;
            const double diff = (double)__fsub_rn(a2, a1);
            const double mn = __dadd_rn(M, __ddiv_rn(diff, 2.0));
            S = __dmul_rn(diff, __dsub_rn(a2d, mn));
            M = mn;
        }
        s_v[0] = M; s_v[1] = S; s_v[2] = (double)a1; s_v[3] = a2d;
    }
    __syncthreads();
    double M = s_v[0], S = s_v[1];
    if (n > 2) {
        const double nr = (double)(n - 2);
        const double sum_r = sum - s_v[2] - s_v[3];
        const double m_r = sum_r / nr;
        double q = 0.0;
        for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
            if (t != 0 && type_code[i] != t - 1) continue;
            const long long key = (lab[i] == 1 ? 0 : (long long)N) + i;
            if (key == m1 || key == m2) continue;
            const double d = (double)__fmul_rn(apb, (float)cnt[i]) - m_r;
            q += d * d;
        }
        q = pbk_block_sum(q, red);
        const double dm = m_r - M;
        S = S + q + dm * dm * (2.0 * nr / (double)n);
        M = (2.0 * M + sum_r) / (double)n;
    }
    if (threadIdx.x == 0) {
        if (t == 0) o[0] = M;
        else {
            o[2 * t - 1] = M;
            o[2 * t] = n > 1 ? sqrt(S / (double)(n - 1)) : 0.0;
        }
    }
}

extern "C" int pbk_grid_reduce(pbk_ctx *ctx, const int32_t *grid_idx, const double *grid_z,
                               const uint8_t *labels, const int32_t *type_code, const float *box,
                               int64_t F, int64_t N, int lat0, int lat1, int nx, int ny, int n_types,
                               int32_t *cells_per_lipid, double *thickness, double *apl, void *stream)
{
    if (!ctx || !grid_idx || !grid_z || !labels || !type_code || !box || !cells_per_lipid ||
        !thickness || !apl || F < 0 || N <= 0 || nx <= 0 || ny <= 0 || n_types < 1 ||
        n_types > GRID_MAX_TYPES)
        return pbk_fail(ctx, PBK_EINVAL, "pbk_grid_reduce: bad argument (n_types <= 16)");
    if (F == 0) return PBK_OK;
    k_grid_reduce<<<(unsigned)F, 256, 0, (cudaStream_t)stream>>>(grid_idx, grid_z, N, nx, ny,
                                                                cells_per_lipid, thickness);
    PBK_CHECK_LAUNCH(ctx, "k_grid_reduce");
    if (!getenv("PBK_SEQ_WELFORD") && F <= 0x7fffffff) {
        k_grid_apl_par<<<dim3((unsigned)F, (unsigned)(n_types + 1)), 256, 0, (cudaStream_t)stream>>>(
            cells_per_lipid, labels, type_code, box, N, lat0, lat1, nx, ny, n_types, apl);
        PBK_CHECK_LAUNCH(ctx, "k_grid_apl_par");
        return PBK_OK;
    }
    const int64_t n_thr = F * (n_types + 1);
    k_grid_apl<<<(unsigned)((n_thr + 63) / 64), 64, 0, (cudaStream_t)stream>>>(
        cells_per_lipid, labels, type_code, box, F, N, lat0, lat1, nx, ny, n_types, apl);
    PBK_CHECK_LAUNCH(ctx, "k_grid_apl");
    return PBK_OK;
}
