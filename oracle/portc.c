/* oracle/portc.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the scalar inner loops of PyBILT's lipid-reduction hot
 * path, used by oracle/port.py (the CPU checker and the reported CPU baseline).
 * Each function cites the reference file:line (relative to /root/reference) whose
 * arithmetic it follows.  Built by oracle/build.py with
 *     gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared
 * (-ffp-contract=off: the only fused multiply-add is the explicit fma() below).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- NumPy float64 pairwise summation (numpy/_core/src/umath/loops_utils.h.src,
 * DOUBLE_pairwise_sum) -- the order `(positions[index][:,k]*masses).sum()` uses at
 * pybilt/bilayer_analyzer/com_frame.py:175-178.  SURVEY.md A.4. */
static double pw_sum_rec(const double *a, int64_t n)
{
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return pw_sum_rec(a, n2) + pw_sum_rec(a + n2, n - n2);
    }
}

double pc_pairwise_sum(const double *a, int64_t n) { return pw_sum_rec(a, n); }

/* ---- Welford running mean (pybilt/common/running_stats.py:34-62) over each row. */
void pc_welford_mean_rows(const double *vals, int64_t rows, int64_t n, double *out)
{
    for (int64_t r = 0; r < rows; r++) {
        const double *v = vals + r * n;
        double m = 0.0;
        for (int64_t i = 0; i < n; i++) {
            if (i == 0) m = v[0];
            else m = m + (v[i] - m) / (double)(i + 1);
        }
        out[r] = m;            /* n == 0 -> 0.0 (running_stats.py:61-62) */
    }
}

/* ---- leaflet assignment (pybilt/bilayer_analyzer/bilayer_analyzer.py:826-847):
 * zavg = Welford mean in lipid order; z > zavg -> upper (1), z < zavg -> lower (0);
 * z == zavg raises KeyError('') in the reference -> return -1 - index here. */
int64_t pc_leaflet_labels(const double *z, int64_t n, uint8_t *label)
{
    double m = 0.0;
    for (int64_t i = 0; i < n; i++) {
        if (i == 0) m = z[0];
        else m = m + (z[i] - m) / (double)(i + 1);
    }
    for (int64_t i = 0; i < n; i++) {
        if (z[i] > m) label[i] = 1;
        else if (z[i] < m) label[i] = 0;
        else return -1 - i;
    }
    return 0;
}

/* ---- distance_euclidean_pbc(v_a, v_b, l_box, center='box_half')
 * (pybilt/common/distance_cutoff_clustering.py:27-71); l_box is float32
 * (analysis_protocols.py:2212,4950; lipid_grid_opt.py:445).  The final
 * np.dot(d_v, d_v) of two float64 goes through OpenBLAS ddot, whose x86 kernels
 * accumulate with FMA: dot = fma(d1, d1, d0*d0) (verified against numpy 2.3.5 /
 * OpenBLAS 0.3.30 in the build container, see DESIGN.md "np.dot"). */
static inline double pbc_dist(double ax, double ay, double bx, double by, float boxx, float boxy)
{
    float cx = boxx / 2.0f, cy = boxy / 2.0f;
    double apx = ax - (double)cx, apy = ay - (double)cy;
    double bpx = bx - (double)cx, bpy = by - (double)cy;
    double dx = apx - bpx, dy = apy - bpy;
    if (fabs(dx) > (double)cx) dx = ((double)boxx - fabs(apx)) - fabs(bpx);
    if (fabs(dy) > (double)cy) dy = ((double)boxy - fabs(apy)) - fabs(bpy);
    return sqrt(fma(dy, dy, dx * dx));
}

void pc_pbc_dist_pairs(const double *pos, const int64_t *ia, int64_t na, const int64_t *ib,
                       int64_t nb, float boxx, float boxy, double *out)
{
    for (int64_t i = 0; i < na; i++)
        for (int64_t j = 0; j < nb; j++)
            out[i * nb + j] = pbc_dist(pos[2 * ia[i]], pos[2 * ia[i] + 1],
                                       pos[2 * ib[j]], pos[2 * ib[j] + 1], boxx, boxy);
}

/* ---- np.histogram(dists, bins=n, range=[lo,hi]) for uniform bins
 * (numpy/lib/_histograms_impl.py, "fast path"): keep lo<=r<=hi, estimate the index,
 * then correct against the linspace edges; last bin closed.  SURVEY.md A.6. */
static inline int hist_bin(double r, double lo, double hi, int nb, const double *edges)
{
    if (!(r >= lo) || !(r <= hi)) return -1;
    double f = ((r - lo) / (hi - lo)) * (double)nb;
    int k = (int)f;
    if (k == nb) k -= 1;
    if (r < edges[k]) k -= 1;
    else if (r >= edges[k + 1] && k != nb - 1) k += 1;
    return k;
}

/* ---- COMLateralRDFProtocol.run_analysis inner loops
 * (pybilt/bilayer_analyzer/analysis_protocols.py:4973-4998): ordered pairs i in A,
 * j in B, i != j, wrapped lateral COMs. */
void pc_rdf_accumulate(const double *pos, const int64_t *ia, int64_t na, const int64_t *ib,
                       int64_t nb_idx, float boxx, float boxy, int nbins, double lo, double hi,
                       const double *edges, int64_t *count)
{
    for (int64_t i = 0; i < na; i++) {
        int64_t a = ia[i];
        for (int64_t j = 0; j < nb_idx; j++) {
            int64_t b = ib[j];
            if (a == b) continue;
            double r = pbc_dist(pos[2 * a], pos[2 * a + 1], pos[2 * b], pos[2 * b + 1], boxx, boxy);
            int k = hist_bin(r, lo, hi, nbins, edges);
            if (k >= 0) count[k] += 1;
        }
    }
}

typedef struct { double d; int64_t j; } nb_t;

/* stable ascending sort by distance (Python list.sort is stable: ties keep insertion order) */
static void merge_sort(nb_t *a, nb_t *tmp, int64_t n)
{
    if (n < 2) return;
    if (n <= 16) {
        for (int64_t i = 1; i < n; i++) {
            nb_t x = a[i];
            int64_t k = i - 1;
            while (k >= 0 && a[k].d > x.d) { a[k + 1] = a[k]; k--; }
            a[k + 1] = x;
        }
        return;
    }
    int64_t h = n / 2;
    merge_sort(a, tmp, h);
    merge_sort(a + h, tmp, n - h);
    int64_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].d < a[i].d) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, (size_t)n * sizeof(nb_t));
}

/* ---- NNFProtocol.run_analysis inner loops (analysis_protocols.py:2219-2249):
 * for every reference lipid i: distances to all other leaflet members j in member
 * order, stable sort, first k; count neighbours whose type flag is set. */
void pc_nnf_counts(const double *pos, const int64_t *ia, int64_t na, const int64_t *iall,
                   int64_t nall, const uint8_t *is_b, int k, float boxx, float boxy,
                   int32_t *n_b)
{
    nb_t *buf = (nb_t *)malloc((size_t)(nall + 1) * sizeof(nb_t));
    nb_t *tmp = (nb_t *)malloc((size_t)(nall + 1) * sizeof(nb_t));
    for (int64_t i = 0; i < na; i++) {
        int64_t a = ia[i], m = 0;
        for (int64_t j = 0; j < nall; j++) {
            int64_t b = iall[j];
            if (a == b) continue;
            buf[m].d = pbc_dist(pos[2 * a], pos[2 * a + 1], pos[2 * b], pos[2 * b + 1], boxx, boxy);
            buf[m].j = b;
            m++;
        }
        merge_sort(buf, tmp, m);
        int32_t c = 0;
        for (int64_t t = 0; t < k && t < m; t++) c += is_b[buf[t].j] ? 1 : 0;
        n_b[i] = c;
    }
    free(buf);
    free(tmp);
}

/* ---- LipidGrid2d._knn (pybilt/lipid_grid/lipid_grid_opt.py:435-468): distance of
 * each unordered pair evaluated once as dist(pos[idx[i]], pos[idx[j]]) with i<j,
 * appended to both lists in ascending position order, stable sort, first nn_k.
 * out_nn[p*nn_k + t] = lipid index of the t-th neighbour of member p (-1 if fewer). */
void pc_knn(const double *pos, const int64_t *idx, int64_t n, int nn_k, float boxx, float boxy,
            int64_t *out_nn)
{
    double *D = (double *)malloc((size_t)n * (size_t)n * sizeof(double));
    for (int64_t i = 0; i < n; i++) {
        D[i * n + i] = 0.0;
        for (int64_t j = i + 1; j < n; j++) {
            double r = pbc_dist(pos[2 * idx[i]], pos[2 * idx[i] + 1],
                                pos[2 * idx[j]], pos[2 * idx[j] + 1], boxx, boxy);
            D[i * n + j] = r;
            D[j * n + i] = r;
        }
    }
    nb_t *buf = (nb_t *)malloc((size_t)(n + 1) * sizeof(nb_t));
    nb_t *tmp = (nb_t *)malloc((size_t)(n + 1) * sizeof(nb_t));
    for (int64_t p = 0; p < n; p++) {
        int64_t m = 0;
        for (int64_t q = 0; q < n; q++) {
            if (q == p) continue;
            buf[m].d = D[p * n + q];
            buf[m].j = idx[q];
            m++;
        }
        merge_sort(buf, tmp, m);
        for (int t = 0; t < nn_k; t++) out_nn[p * nn_k + t] = (t < m) ? buf[t].j : -1;
    }
    free(buf);
    free(tmp);
    free(D);
}

/* cell-centre to lipid distance of the lipid grids (lipid_grid_opt.py:252-263,
 * lipid_grid.py:228-239): minimum image with the reference's quirk that the dy test
 * uses bxh.  The reference squares with `dx**2` on NumPy float64 scalars, i.e. libm
 * pow(dx, 2.0), which is within 1 ulp of (and 99.9 % of the time equal to) dx*dx;
 * this restatement uses the exact square (DESIGN.md "grid squares"). */
static inline double cell_dist(double x, double y, double xi, double yi, float boxx, float boxy)
{
    float bxh = boxx / 2.0f, byh = boxy / 2.0f;
    double dx = x - xi, dy = y - yi;
    if (fabs(dx) > (double)bxh) dx = ((double)boxx - fabs(x - (double)bxh)) - fabs(xi - (double)bxh);
    if (fabs(dy) > (double)bxh) dy = ((double)boxy - fabs(y - (double)byh)) - fabs(yi - (double)byh);
    return sqrt(dx * dx + dy * dy);
}

/* ---- lipid_grid.LipidGrid2d.__init__ (pybilt/lipid_grid/lipid_grid.py:217-248): exact
 * nearest lipid per cell, members scanned in order, strict '<'. */
void pc_grid_exact(const double *pos, const double *z, const int64_t *idx, int64_t n,
                   const double *xc, int nx, const double *yc, int ny, float boxx, float boxy,
                   int64_t *gidx, double *gz)
{
    for (int cx = 0; cx < nx; cx++)
        for (int cy = 0; cy < ny; cy++) {
            double rmin = 1.0e10, zmin = 0.0;
            int64_t imin = 0;
            for (int64_t m = 0; m < n; m++) {
                int64_t i = idx[m];
                double r = cell_dist(xc[cx], yc[cy], pos[2 * i], pos[2 * i + 1], boxx, boxy);
                if (r < rmin) { rmin = r; imin = i; zmin = z[i]; }
            }
            gidx[cx * ny + cy] = imin;
            gz[cx * ny + cy] = zmin;
        }
}

/* ---- lipid_grid_opt.LipidGrid2d.__init__ (pybilt/lipid_grid/lipid_grid_opt.py:237-334):
 * cell (0,0) exact; every other cell searches p = grid[max(0,cx-1)][max(0,cy-1)] and
 * then the 24 nearest neighbours of p, first minimum wins.  `knn` rows are indexed by
 * member position; `pos_of[lipid]` maps a lipid index to its member position. */
void pc_grid_opt(const double *pos, const double *z, const int64_t *idx, int64_t n,
                 const int64_t *knn, int nn_k, const int64_t *pos_of,
                 const double *xc, int nx, const double *yc, int ny, float boxx, float boxy,
                 int64_t *gidx, double *gz)
{
    for (int cx = 0; cx < nx; cx++)
        for (int cy = 0; cy < ny; cy++) {
            double rmin = 1.0e10, zmin = 0.0;
            int64_t imin = 0;
            if (cx == 0 && cy == 0) {
                for (int64_t m = 0; m < n; m++) {
                    int64_t i = idx[m];
                    double r = cell_dist(xc[cx], yc[cy], pos[2 * i], pos[2 * i + 1], boxx, boxy);
                    if (r < rmin) { rmin = r; imin = i; zmin = z[i]; }
                }
            } else {
                int px = cx - 1 < 0 ? 0 : cx - 1, py = cy - 1 < 0 ? 0 : cy - 1;
                int64_t p = gidx[px * ny + py];
                imin = p;
                zmin = (double)p;       /* lipid_grid_opt.py:276 seeds z_min with the index */
                double r = cell_dist(xc[cx], yc[cy], pos[2 * p], pos[2 * p + 1], boxx, boxy);
                if (r < rmin) { rmin = r; imin = p; zmin = z[p]; }
                const int64_t *row = knn + pos_of[p] * nn_k;
                for (int t = 0; t < nn_k; t++) {
                    int64_t q = row[t];
                    if (q < 0) break;
                    r = cell_dist(xc[cx], yc[cy], pos[2 * q], pos[2 * q + 1], boxx, boxy);
                    if (r < rmin) { rmin = r; imin = q; zmin = z[q]; }
                }
            }
            gidx[cx * ny + cy] = imin;
            gz[cx * ny + cy] = zmin;
        }
}

/* ---- wrap_coordinates (pybilt/mda_tools/mda_unwrap_vectorized.py:12-93), one axis of
 * one frame, scalar form of SURVEY.md B.1: d = f32(x - u_prev); shift count s from the
 * strict '>'/'<' loops evaluated in float64; u = f32(f64(x) + s*L). */
void pc_unwrap_axis(const float *x, const float *uprev, int64_t n, int64_t stride, float L,
                    float *u)
{
    float h = L / 2.0f;
    for (int64_t i = 0; i < n; i++) {
        float xv = x[i * stride];
        float d = xv - uprev[i * stride];
        double s = 0.0;
        if (d > h) {
            do { s -= 1.0; } while ((double)d + s * (double)L > (double)h);
        } else if (d < -h) {
            do { s += 1.0; } while ((double)d + s * (double)L < -(double)h);
        }
        u[i * stride] = (float)((double)xv + s * (double)L);
    }
}

/* ======================================================================================
 * Large-shape variants (BASELINE configs C3-C5).  Same arithmetic as the functions above --
 * the identical pbc_dist / hist_bin calls in the identical argument order -- reorganised so
 * that 32,768- and 131,072-lipid leaflets finish in seconds: rows spread over OpenMP threads,
 * no n x n matrix, and for the RDF a cell list that only skips pairs which cannot land in a
 * bin.  tests/test_oracle_golden.py pins each of them to its dense counterpart above.
 * ====================================================================================== */

/* ---- LipidGrid2d._knn (lipid_grid_opt.py:435-468) without the n x n matrix: row p keeps its
 * nn_k best (distance, position) in ascending order; a candidate enters behind every entry of
 * equal distance (candidates arrive in ascending position order, so this is the stable sort
 * of pc_knn).  Each pair is evaluated as dist(lower position, higher position) like there. */
void pc_knn_lean(const double *pos, const int64_t *idx, int64_t n, int nn_k, float boxx, float boxy,
                 int64_t *out_nn)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t p = 0; p < n; p++) {
        double bd[64];
        int64_t bj[64];
        int m = 0;
        for (int64_t q = 0; q < n; q++) {
            if (q == p) continue;
            int64_t lo = p < q ? p : q, hi = p < q ? q : p;
            double r = pbc_dist(pos[2 * idx[lo]], pos[2 * idx[lo] + 1], pos[2 * idx[hi]], pos[2 * idx[hi] + 1],
                                boxx, boxy);
            if (m == nn_k && !(r < bd[m - 1])) continue;
            int k = m < nn_k ? m : nn_k - 1;
            while (k > 0 && bd[k - 1] > r) { bd[k] = bd[k - 1]; bj[k] = bj[k - 1]; k--; }
            bd[k] = r;
            bj[k] = idx[q];
            if (m < nn_k) m++;
        }
        for (int t = 0; t < nn_k; t++) out_nn[p * nn_k + t] = (t < m) ? bj[t] : -1;
    }
}

/* ---- COMLateralRDFProtocol.run_analysis inner loops (analysis_protocols.py:4973-4998) over a
 * cell list.  Only valid when every position lies inside [0, box) on both lateral axes (then
 * the reference's formula is the minimum image and |d| <= r_out implies neighbouring cells);
 * returns -1 without counting otherwise.  Cells have edge >= hi (the outer range), at least 3
 * per axis, so the 3 x 3 block around a cell holds every partner within range exactly once.
 * The pairs are visited as ordered pairs (a in A, b in B, a != b) with pbc_dist(a, b) -- the
 * argument order matters for the rounding of the wrapped branch. */
int64_t pc_rdf_accumulate_cells(const double *pos, const int64_t *ia, int64_t na, const int64_t *ib,
                                int64_t nb_idx, float boxx, float boxy, int nbins, double lo, double hi,
                                const double *edges, int64_t *count)
{
    if (!(hi > 0.0)) return -1;
    int ncx = (int)floor((double)boxx / hi), ncy = (int)floor((double)boxy / hi);
    if (ncx < 3 || ncy < 3) return -1;
    if (ncx > 1024) ncx = 1024;
    if (ncy > 1024) ncy = 1024;
    const double wx = (double)boxx / ncx, wy = (double)boxy / ncy;
    for (int64_t j = 0; j < nb_idx; j++) {
        double x = pos[2 * ib[j]], y = pos[2 * ib[j] + 1];
        if (!(x >= 0.0 && x < (double)boxx && y >= 0.0 && y < (double)boxy)) return -1;
    }
    for (int64_t i = 0; i < na; i++) {
        double x = pos[2 * ia[i]], y = pos[2 * ia[i] + 1];
        if (!(x >= 0.0 && x < (double)boxx && y >= 0.0 && y < (double)boxy)) return -1;
    }
    int64_t ncell = (int64_t)ncx * ncy;
    int64_t *start = (int64_t *)calloc((size_t)ncell + 1, sizeof(int64_t));
    int64_t *cell_of = (int64_t *)malloc((size_t)nb_idx * sizeof(int64_t));
    int64_t *order = (int64_t *)malloc((size_t)nb_idx * sizeof(int64_t));
    for (int64_t j = 0; j < nb_idx; j++) {
        int cx = (int)(pos[2 * ib[j]] / wx), cy = (int)(pos[2 * ib[j] + 1] / wy);
        if (cx >= ncx) cx = ncx - 1;
        if (cy >= ncy) cy = ncy - 1;
        cell_of[j] = (int64_t)cx * ncy + cy;
        start[cell_of[j] + 1]++;
    }
    for (int64_t c = 0; c < ncell; c++) start[c + 1] += start[c];
    int64_t *fill = (int64_t *)malloc((size_t)ncell * sizeof(int64_t));
    memcpy(fill, start, (size_t)ncell * sizeof(int64_t));
    for (int64_t j = 0; j < nb_idx; j++) order[fill[cell_of[j]]++] = ib[j];
#pragma omp parallel
    {
        int64_t *mine = (int64_t *)calloc((size_t)nbins, sizeof(int64_t));
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < na; i++) {
            int64_t a = ia[i];
            int cx = (int)(pos[2 * a] / wx), cy = (int)(pos[2 * a + 1] / wy);
            if (cx >= ncx) cx = ncx - 1;
            if (cy >= ncy) cy = ncy - 1;
            for (int ox = -1; ox <= 1; ox++)
                for (int oy = -1; oy <= 1; oy++) {
                    int64_t c = (int64_t)((cx + ox + ncx) % ncx) * ncy + (cy + oy + ncy) % ncy;
                    for (int64_t s = start[c]; s < start[c + 1]; s++) {
                        int64_t b = order[s];
                        if (a == b) continue;
                        double r = pbc_dist(pos[2 * a], pos[2 * a + 1], pos[2 * b], pos[2 * b + 1], boxx, boxy);
                        int k = hist_bin(r, lo, hi, nbins, edges);
                        if (k >= 0) mine[k] += 1;
                    }
                }
        }
#pragma omp critical
        for (int k = 0; k < nbins; k++) count[k] += mine[k];
        free(mine);
    }
    free(start); free(cell_of); free(order); free(fill);
    return 0;
}
