/* pbk.h -- C ABI of libpbk, the B200 (sm_100a) kernels behind PyBILT's BilayerAnalyzer
 * analysis-plugin API (frame-batched lipid reduction hot path).
 *
 * Every entry point replaces a piece of the reference's pure-Python per-frame loop; the
 * file:line each one restates is given beside it (paths relative to the PyBILT tree).
 * The reference has no FFI of its own (it is 100 % Python), so the binding a maintainer
 * adds is the ctypes stub shown in INTEGRATION.md; pybilt_b200/_native.py is that stub.
 *
 * Conventions
 *   - all array arguments are DEVICE pointers owned by the caller (allocated e.g. with
 *     torch.empty(..., device='cuda')); the library allocates nothing that outlives a call;
 *   - F = analysed frames in this batch, M = selected atoms, N = lipids (COM beads);
 *   - positions are float32 [F][M][3] exactly as MDAnalysis hands them over, boxes float32
 *     [F][3], masses float64 [M]; COM outputs float64 [F][N][3]; labels uint8 [F][N]
 *     (1 = upper, 0 = lower);
 *   - every call enqueues on the given cudaStream_t (passed as void*) and returns without
 *     synchronising; data-dependent failures (a lipid exactly on the leaflet mid-plane, an
 *     image count that does not fit int16) are written to the int32 `status` word the
 *     caller supplies and reads after synchronising;
 *   - return value: 0 on success, <0 = PBK_E*; pbk_last_error() gives the text.
 *     No C++ exception crosses this boundary.
 */
#ifndef PBK_H
#define PBK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBK_ABI_VERSION 8

#define PBK_OK 0
#define PBK_EINVAL (-1) /* bad argument */
#define PBK_ECUDA (-2)  /* CUDA runtime error (launch, attribute ...) */
#define PBK_EUNSUPPORTED (-3) /* shape outside this entry point's fast path: use the generic one */

/* bits of the device-side status word */
#define PBK_ST_IMAGE_OVERFLOW 1 /* |image count| > 32767 or non-positive box length */
#define PBK_ST_LEAFLET_TIE 2    /* com_unwrap[norm] == zavg: reference raises KeyError('') */
#define PBK_ST_UNWRAP_AMBIGUOUS 4 /* fused path: a jump within rounding distance of L/2; redo with
                                     pbk_unwrap_images/pbk_membrane_com/pbk_lipid_com */
#define PBK_ST_FUSED_RANGE 8    /* fused path: |image count| > 126; redo with the generic path */
#define PBK_ST_INTERNAL 16      /* a bulk copy never completed (should not happen) */

typedef struct pbk_ctx pbk_ctx;

int pbk_abi_version(void);
int pbk_ctx_create(int device, pbk_ctx **out);
void pbk_ctx_destroy(pbk_ctx *ctx);
const char *pbk_last_error(const pbk_ctx *ctx);

/* ---- N2  trajectory ingestion (the wire format before the path) ---------------------------
 * MDAData / mda.Universe(psf, dcd)             pybilt/bilayer_analyzer/mda_data.py:14-38
 * frame.positions[index]                       pybilt/bilayer_analyzer/bilayer_analyzer.py:936
 * A DCD frame stores the records X[], Y[], Z[] of all atoms.  `frames` are F frames exactly as
 * they lie in the file (raw bytes viewed as float32, frame_stride floats apart, the X / Y / Z
 * records starting off_x / off_y / off_z floats into a frame; byteswap != 0 for a file of the
 * other endianness); out = positions float32 [F][M][3] of the selected atoms sel[M] (int32 atom
 * indices, ascending; NULL = the first M atoms in file order).  F <= 65535 per call. */
int pbk_planes_to_aos(pbk_ctx *ctx, const float *frames, const int32_t *sel, int64_t F,
                      int64_t frame_stride, int64_t off_x, int64_t off_y, int64_t off_z, int byteswap,
                      int64_t M, float *out, void *stream);

/* ---- A1  PBC unwrap ---------------------------------------------------------------------
 * wrap_coordinates(abc, coord, refcoord)      pybilt/mda_tools/mda_unwrap_vectorized.py:12-93
 * + the _oldcoords recurrence                 pybilt/bilayer_analyzer/bilayer_analyzer.py:936-954
 * Per atom and axis: d = f32(x_t - u_{t-1}); image shift s_t from the strict >/< loops
 * evaluated in float64; u_t = f32(f64(x_t) + s_t*L_t).  Writes the integer image counts
 * img[F][M][3]; u_state[M][3] (float32) carries u of the last frame in and out so batches
 * chain.  first_frame != 0: frame 0 of the batch is its own reference (u_0 = x_0).
 * init_images (may be NULL; needs first_frame): frame 0 is a raw frame whose image counts
 * [M][3] int32 are already known -- how a frame shard continues the ranks before it
 * (pybilt_b200/distributed.py). */
#define PBK_UNWRAP_SCALAR_F32 1 /* flags: the scalar unwrap of pybilt/mda_tools/mda_unwrap.py:12-85
                                  (what COMTraj calls): `dz + shift*C` evaluated in float32 */
int pbk_unwrap_images(pbk_ctx *ctx, const float *x, const float *box, float *u_state,
                      int first_frame, const int32_t *init_images, int flags, int64_t F, int64_t M,
                      int16_t *img, int32_t *status, void *stream);

/* ---- A3  membrane centre of mass --------------------------------------------------------
 * mem_com_k = (f64(u[:,k]) * masses).sum() / masses.sum()   pybilt/bilayer_analyzer/com_frame.py:175-179
 * evaluated in NumPy's pairwise-summation order so that it is bit-identical (the value is
 * subtracted from every coordinate before a float32 rounding, com_frame.py:180).  The
 * summation tree is passed as a plan built by the host (pybilt_b200/pairwise.py):
 * leaf_start/leaf_len[L] (<=128-element blocks), node_left/node_right[K] (internal nodes,
 * children numbered leaves 0..L-1 then nodes L..L+K-1, ordered by height),
 * level_off[H+1] (node ranges per height).  node_vals is scratch [F][L+K][3] float64. */
int pbk_membrane_com(pbk_ctx *ctx, const float *x, const float *box, const int16_t *img,
                     const double *mass, int64_t F, int64_t M, const int32_t *leaf_start,
                     const int32_t *leaf_len, int32_t L, const int32_t *node_left,
                     const int32_t *node_right, int32_t K, const int32_t *level_off, int32_t H,
                     double total_mass, double *node_vals, double *mem_com, void *stream);

/* ---- T1  membrane centre of mass as COMTraj forms it -------------------------------------
 * mem_sel.center_of_mass()                       pybilt/com_trajectory/COMTraj.py:1084
 * = (pos.astype(f64) * m[:,None]).sum(axis=0) / m.sum(): rows accumulated in atom order. */
int pbk_membrane_com_seq(pbk_ctx *ctx, const float *x, const float *box, const int16_t *img,
                         const double *mass, int64_t F, int64_t M, double total_mass,
                         double *mem_com, void *stream);

/* ---- A2 + A4  per-lipid centres of mass -------------------------------------------------
 * LipidCOM.extract wrapped / unwrapped passes  pybilt/bilayer_analyzer/com_frame.py:43-96,163-185
 * com_i        = sum_a m_a*f64(x_a)                    / lipid_mass_i   (raw wrapped positions)
 * com_unwrap_i = sum_a m_a*f64(f32(f64(u_a) - mem_com)) / lipid_mass_i
 * atoms of lipid i are gather[seg[i] .. seg[i+1]) in that order (gather == NULL: identity,
 * i.e. lipids are contiguous atom ranges); accumulated in atom order like MDAnalysis'
 * (pos*w[:,None]).sum(axis=0). */
int pbk_lipid_com(pbk_ctx *ctx, const float *x, const float *box, const int16_t *img,
                  const double *mass, const double *mem_com, const int32_t *seg,
                  const int32_t *gather, const double *lipid_mass, int64_t F, int64_t M, int64_t N,
                  double *com, double *com_unwrap, void *stream);

/* ---- A1-A4 + B1 fused  (small systems: one frame of positions fits in shared memory) -------
 * Same results as pbk_unwrap_images + pbk_membrane_com + pbk_lipid_com (+ pbk_leaflets with
 * update_interval 1 when do_labels != 0) with one reducing pass over the positions:
 * a streaming pre-pass counts periodic images per frame chunk, a scan turns them into the
 * image count at every chunk start, and one CTA per chunk then applies the exact reference
 * recurrence frame by frame out of shared memory (frames arrive by TMA bulk copy).
 * Returns PBK_EUNSUPPORTED when a frame does not fit (caller falls back to the calls above).
 * The pre-pass decides image increments from consecutive raw frames; if any jump lies
 * within rounding distance of half a box, PBK_ST_UNWRAP_AMBIGUOUS is raised in `status` and
 * the outputs must be recomputed with the generic entry points.
 * scratch: pbk_reduce_fused_scratch_bytes(F, M, sm_count) bytes of device memory.
 * phase: 3 = everything; 1 = image-count pre-pass only (writes net_images[M][3], the net
 * image-count change over the first net_frames frames of the batch -- 0 = all of it -- when not
 * NULL); 2 = finish from the pre-pass left in scratch, with init_images[M][3] (image counts of
 * frame 0, may be NULL) added -- the two halves of a frame shard, around the exchange of net
 * counts between ranks.  A shard pushes [back halo | owned | forward halo] in one batch with
 * net_frames = back + owned: the next rank's stream starts at its last owned frame.  Phases 1 and
 * 2 of one batch must pass the same net_frames. */
int pbk_reduce_fused_scratch_bytes(int64_t F, int64_t M, int sm_count, int64_t *bytes,
                                   int32_t *n_chunks, int32_t *chunk);
int pbk_reduce_fused(pbk_ctx *ctx, const float *x, const float *box, const double *mass,
                     const int32_t *seg, const int32_t *gather, const double *lipid_mass,
                     const int32_t *leaf_start, const int32_t *leaf_len, int32_t L,
                     const int32_t *node_left, const int32_t *node_right, int32_t K,
                     const int32_t *level_off, int32_t H, double total_mass, float *u_state,
                     int first_frame, int64_t F, int64_t M, int64_t N, int norm, int do_labels,
                     int phase, const int32_t *init_images, int32_t *net_images, int64_t net_frames,
                     void *scratch, int64_t scratch_bytes, double *com, double *com_unwrap,
                     double *mem_com, uint8_t *labels, int32_t *status, void *stream);
int pbk_sm_count(const pbk_ctx *ctx);

/* ---- A1-A4 + B1 streamed  (large systems: a frame does not fit in shared memory) -----------
 * Same results as pbk_unwrap_images + pbk_membrane_com + pbk_lipid_com + pbk_leaflets with TWO
 * reads of the positions and no per-frame image-count array.  Sweep A (one CTA per summation
 * leaf, frames in time order, the leaf's slice of every frame arriving by TMA bulk copy in a
 * shared-memory ring) decides the image counts, forms the NumPy-pairwise leaf sums of m*f64(u)
 * -> mem_com, and leaves the image counts of every 32nd frame behind; sweep B (a thread per
 * lipid, axis and 32-frame chunk) repeats the recurrence from there and forms both COMs; the
 * labels come from the compact bilayer-normal components.
 * max_bead_atoms: the largest seg[i+1]-seg[i] (<= 800, else PBK_EUNSUPPORTED).  No atom may
 * belong to two beads.  phase / init_images / net_images / net_frames as for pbk_reduce_fused
 * (phase 1 = image counts only, which is also how a shard advances u_state without reducing;
 * net_images may also be asked for in phase 2 / 3: the image counts of frame net_frames-1);
 * frame0 / update_interval / prev_labels as for pbk_leaflets.
 * scratch: pbk_reduce_stream_scratch_bytes(F, M, N, L, K). */
int pbk_reduce_stream_scratch_bytes(int64_t F, int64_t M, int64_t N, int32_t L, int32_t K, int64_t *bytes);
int pbk_reduce_stream(pbk_ctx *ctx, const float *x, const float *box, const double *mass,
                      const int32_t *seg, const int32_t *gather, const double *lipid_mass,
                      int32_t max_bead_atoms, const int32_t *leaf_start, const int32_t *leaf_len,
                      int32_t L, const int32_t *node_left, const int32_t *node_right, int32_t K,
                      const int32_t *level_off, int32_t H, double total_mass, float *u_state,
                      int first_frame, int64_t F, int64_t M, int64_t N, int norm, int phase,
                      const int32_t *init_images, int32_t *net_images, int64_t net_frames,
                      int64_t frame0, int update_interval, const uint8_t *prev_labels, void *scratch,
                      int64_t scratch_bytes, double *com, double *com_unwrap, double *mem_com,
                      uint8_t *labels, int32_t *status, void *stream);

/* ---- B1  leaflet assignment -------------------------------------------------------------
 * BilayerAnalyzer._build_leaflets             pybilt/bilayer_analyzer/bilayer_analyzer.py:826-847
 * label = 1 (upper) if com_unwrap[norm] > zavg, 0 (lower) if < zavg, zavg = Welford running
 * mean in lipid order (pybilt/common/running_stats.py:34-52).  Frames whose global index
 * (frame0 + f) is not a multiple of update_interval copy the previous frame's labels
 * (bilayer_analyzer.py:969-981); prev_labels[N] is the row before this batch (may be NULL
 * when frame0 % update_interval == 0). */
int pbk_leaflets(pbk_ctx *ctx, const double *com_unwrap, int64_t F, int64_t N, int norm,
                 int64_t frame0, int update_interval, const uint8_t *prev_labels,
                 uint8_t *labels, int32_t *status, void *stream);

/* ---- N3  thickness from the leaflets' centres of mass --------------------------------------
 * LeaftoLeafDistProtocol.run_analysis   pybilt/bilayer_analyzer/analysis_protocols.py:1154-1170
 * COMFrame.coordinates(leaflet=...)      pybilt/bilayer_analyzer/com_frame.py:369-385
 * out[f] = | mean over upper lipids of com[f][:, norm] - mean over lower lipids |, the means as
 * ndarray.mean(axis=0) forms them (rows added in lipid order, divided by the count); WRAPPED COMs;
 * nan when a leaflet is empty. */
int pbk_leaflet_distance(pbk_ctx *ctx, const double *com, const uint8_t *labels, int64_t F, int64_t N,
                         int norm, double *out, void *stream);

/* ---- N3  flip-flop detection ----------------------------------------------------------------
 * FlipFlopProtocol.run_analysis          pybilt/bilayer_analyzer/analysis_protocols.py:3784-3833
 * changed[f] = number of lipids whose leaflet label differs between frame f and frame f-1 (frame 0
 * of the call against prev_row[N]; prev_row NULL: 0).  The reference's reference leaflets are always
 * the previous frame's, so changed[f] != 0 marks the frames that produce events. */
int pbk_label_changes(pbk_ctx *ctx, const uint8_t *labels, const uint8_t *prev_row, int64_t F, int64_t N,
                      int32_t *changed, void *stream);

/* ---- N3  com_frame:rewrap -----------------------------------------------------------------------
 * COMFrame._rewrap                       pybilt/bilayer_analyzer/com_frame.py:233-270
 * com[f, i] <- (com_unwrap[f, i] + mem_com[f]) moved by whole box edges into [0, box] wherever it lies
 * more than 1 A from com[f, i]; in place. */
int pbk_rewrap(pbk_ctx *ctx, double *com, const double *com_unwrap, const double *mem_com, const float *box,
               int64_t F, int64_t N, int32_t *status, void *stream);

/* ---- self-check: division by a reused mass ---------------------------------------------------------
 * The per-lipid COM division (com_frame.py:91-96, MDAnalysis center_of_mass) is done as pbk_div_r: a
 * reciprocal formed once per lipid and two FMA corrections.  mismatches[0] (device, caller zeroes it) +=
 * number of n pseudo-random operand pairs for which that differs from the correctly rounded quotient. */
int pbk_debug_div_check(pbk_ctx *ctx, uint64_t seed, int64_t n, uint64_t *mismatches, void *stream);

/* ---- N3  leaflets from the lipids' orientation -----------------------------------------------------
 * BilayerAnalyzer._build_leaflets_orientation   pybilt/bilayer_analyzer/bilayer_analyzer.py:849-876
 * labels[f, i] = 1 / 0 by the sign of (COM of the head atoms - COM of the tail atoms)[norm], both taken
 * from the unwrapped, membrane-COM-free float32 coordinates of the frame, on the frames where the
 * leaflets are rebuilt ((frame0 + f) % interval == 0; other rows are left alone).  seg[2 i .. 2 i + 2]
 * delimits lipid i's tail and head atoms in sel_atom / sel_mass; u_state[n_sel] carries the unwrapped
 * normal components from batch to batch (first_frame: frame 0 is its own reference).  A zero component:
 * PBK_ST_LEAFLET_TIE (the reference raises KeyError('')). */
int pbk_orient_labels(pbk_ctx *ctx, const float *x, const float *box, const double *mem_com,
                      const int32_t *sel_atom, const double *sel_mass, const int32_t *seg, float *u_state,
                      int first_frame, int64_t F, int64_t M, int64_t N, int norm, int64_t frame0, int interval,
                      uint8_t *labels, int32_t *status, void *stream);

/* ---- N4  pair searches of the stand-alone COM trajectory ---------------------------------------
 * COMTraj.calc_membrane_thickness                       pybilt/com_trajectory/COMTraj.py:1244-1408
 * COMTraj.calc_area_per_lipid_closest_neighbor_circle   pybilt/com_trajectory/COMTraj.py:1655-1766
 * COMTraj.check_clustering                              pybilt/com_trajectory/COMTraj.py:1411-1576
 * COMTraj's own minimum image on the wrapped COMs (d = a - b; |d| > h: d = L - |a - h| - |b - h| with
 * h = f32(L/2)), rxy = sqrt(dx*dx + dy*dy).
 * mode 1 (N3: MSDDistProtocol._ordered_com_dists analysis_protocols.py:5776-5805, DCClusterProtocol
 * :2936-2944 with distance_cutoff_clustering.py:27-71,127-143): the distance is
 * distance_euclidean_pbc(a, b, box, center='box_half') instead.
 * pbk_ct_nearest: for every lipid idx_a[a] the candidates idx_b[0..nb) are scanned in list order
 * (idx_b[j] == idx_a[a] skipped when exclude_self); out_pos[f, a] = position j of the first strict
 * minimum below 10000.0 (-1: none), out_dist[f, a] = that distance (10000.0: none), out_dist2[f, a]
 * (may be NULL) = the second smallest distance.
 * pbk_ct_adjacency: bits[f, a, w] bit b: rxy(candidate idx[32 w + b], seed idx[a]) <= dist; mode 1:
 * dist(idx[min(a, b')], idx[max(a, b')]) <= dist for b' = 32 w + b != a (each pair evaluated once). */
int pbk_ct_nearest(pbk_ctx *ctx, const double *com, const float *box, const int32_t *idx_a, int32_t na,
                   const int32_t *idx_b, int32_t nb, int xi, int yi, int exclude_self, int mode, int64_t F,
                   int64_t N, int32_t *out_pos, double *out_dist, double *out_dist2, void *stream);
int pbk_ct_adjacency(pbk_ctx *ctx, const double *com, const float *box, const int32_t *idx, int32_t n,
                     int xi, int yi, double dist, int mode, int64_t F, int64_t N, uint32_t *bits, void *stream);

/* ---- C1 / C2  mean squared displacement -------------------------------------------------
 * MSDProtocol.run_analysis / MSDMultiProtocol.run_analysis
 *                                  pybilt/bilayer_analyzer/analysis_protocols.py:582-667, 781-895
 * out[p] = mean_{i in idx} |r_{end[p], i} - r_{origin[p], i}|^2 over the two lateral
 * components lat0, lat1 of com_unwrap.  Single-origin MSD is pairs (0, t); multi-origin
 * MSD passes one pair per finished time block. */
int pbk_msd_pairs(pbk_ctx *ctx, const double *com_unwrap, int64_t F, int64_t N,
                  const int32_t *idx, int32_t n_idx, int lat0, int lat1,
                  const int32_t *pair_origin, const int32_t *pair_end, int32_t n_pairs,
                  double *out, void *stream);

/* ---- N3  average lateral displacement ------------------------------------------------------
 * ALDProtocol.run_analysis             pybilt/bilayer_analyzer/analysis_protocols.py:3262-3339
 * Same arguments as pbk_msd_pairs; out[p] = mean_{i in idx} |r_{end[p], i} - r_{origin[p], i}|
 * (the length, not its square). */
int pbk_ald_pairs(pbk_ctx *ctx, const double *com_unwrap, int64_t F, int64_t N,
                  const int32_t *idx, int32_t n_idx, int lat0, int lat1,
                  const int32_t *pair_origin, const int32_t *pair_end, int32_t n_pairs,
                  double *out, void *stream);

/* ---- D1 + D2  COM lateral RDF -----------------------------------------------------------
 * COMLateralRDFProtocol.run_analysis   pybilt/bilayer_analyzer/analysis_protocols.py:4900-5006
 * distance_euclidean_pbc               pybilt/common/distance_cutoff_clustering.py:27-71
 * np.histogram fast path               numpy/lib/_histograms_impl.py
 * Per frame and per leaflet in leaflet_mask (bit0 upper, bit1 lower) that holds at least one
 * lipid of type_a and one of type_b (type -1 = 'all', any other negative = no lipid): every ordered pair i in A, j in B,
 * i != j, minimum-image lateral distance of the wrapped COMs, binned on edges[n_bins+1]
 * (the float64 linspace np.histogram builds).  thr[n_bins+1] are the same edges expressed on
 * d^2: thr[k] = min{t : sqrt_rn(t) >= edges[k]}, thr[n_bins] = min{t : sqrt_rn(t) > edges[n_bins]}
 * (pybilt_b200/engine.py: rdf_thresholds), which lets the kernel bin without sqrt; thr == NULL
 * selects the dense sqrt-based fallback kernel.  count[n_bins] (uint64) is accumulated;
 * frame_counts[F][2][3] int32 receives (N_a, N_b, N_leaflet) per frame and leaflet
 * (upper row 0, lower row 1), N_a = N_b = 0 where the leaflet was skipped. */
int pbk_rdf(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
            const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
            int leaflet_mask, int32_t n_bins, double range_lo, double range_hi,
            const double *edges, const double *thr, unsigned long long *count,
            int32_t *frame_counts, void *stream);

/* ---- D2, many analyses at once ----------------------------------------------------------
 * The prefab drivers request one com_lateral_rdf per ordered resname pair and leaflet with the
 * same bins (pybilt/bilayer_analyzer/prefab_analysis_protocols.py:1508-1521: 18 analyses for
 * three lipid types).  pbk_rdf_multi makes ONE pair sweep per frame and leaflet and bins every
 * ordered pair (i, j) into the table of (leaflet, type_i, type_j):
 * count[2][n_types][n_types][n_bins] uint64 (accumulated; leaflet row 0 = upper);
 * type_counts[F][2][n_types] int32 = lipids per frame, leaflet and type (overwritten).
 * Any (resname_1, resname_2, leaflet) analysis -- 'all' included -- is a sum of table entries,
 * its (N_a, N_b, N_leaflet) sums follow from type_counts (pybilt_b200/analysis_protocols.py).
 * Returns PBK_EUNSUPPORTED for more than 8 types or tables that do not fit in shared memory. */
int pbk_rdf_multi(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                  const float *box, int64_t F, int64_t N, int lat0, int lat1, int n_types,
                  int32_t n_bins, double range_lo, double range_hi, const double *edges,
                  const double *thr, unsigned long long *count, int32_t *type_counts, void *stream);

/* ---- D3  nearest-neighbour fraction -----------------------------------------------------
 * NNFProtocol.run_analysis             pybilt/bilayer_analyzer/analysis_protocols.py:2174-2257
 * For every lipid i of type_a in a leaflet that holds both types: the k nearest other
 * members of the leaflet (stable order: equal distances keep ascending lipid index);
 * n_b[F][N] = number of them of type_b (-1 for lipids that are not reference lipids);
 * frac[F] = Welford mean over reference lipids (upper members first) of n_b/k, 0 if none. */
int pbk_nnf(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
            const float *box, int64_t F, int64_t N, int lat0, int lat1, int type_a, int type_b,
            int leaflet_mask, int k, int32_t *n_b, double *frac, void *stream);

/* ---- D3, many analyses at once ------------------------------------------------------------
 * The prefab drivers request one nnf analysis per ordered resname pair
 * (pybilt/bilayer_analyzer/prefab_analysis_protocols.py:880-890).  pbk_nnf_multi searches the k
 * nearest leaflet members of EVERY lipid once and tallies their types:
 * type_hits[F][N][n_types] int32, type_counts[F][2][n_types] int32 (lipids per leaflet and type).
 * pbk_nnf_select then forms the n_b[F][N] / frac[F] outputs of pbk_nnf for one
 * (type_a, type_b, leaflet_mask) from those tallies.  PBK_EUNSUPPORTED for more than 8 types. */
int pbk_nnf_multi(pbk_ctx *ctx, const double *com, const uint8_t *labels, const int32_t *type_code,
                  const float *box, int64_t F, int64_t N, int lat0, int lat1, int n_types, int k,
                  int32_t *type_hits, int32_t *type_counts, void *stream);
int pbk_nnf_select(pbk_ctx *ctx, const int32_t *type_hits, const int32_t *type_counts,
                   const uint8_t *labels, const int32_t *type_code, int64_t F, int64_t N, int n_types,
                   int type_a, int type_b, int leaflet_mask, int k, int32_t *n_b, double *frac,
                   void *stream);

/* pbk_nnf_select for n_specs analyses at once: specs[n_specs][3] int32 = (type_a, type_b,
 * leaflet_mask) on the device; n_b[n_specs][F][N], frac[n_specs][F]. */
int pbk_nnf_select_many(pbk_ctx *ctx, const int32_t *type_hits, const int32_t *type_counts,
                        const uint8_t *labels, const int32_t *type_code, int64_t F, int64_t N, int n_types,
                        int n_specs, const int32_t *specs, int k, int32_t *n_b, double *frac, void *stream);

/* ---- G1 / G2  lipid grids ---------------------------------------------------------------
 * lipid_grid_opt.LipidGrid2d.__init__ + _knn  pybilt/lipid_grid/lipid_grid_opt.py:155-334,435-468
 * lipid_grid.LipidGrid2d.__init__ (exact)     pybilt/lipid_grid/lipid_grid.py:150-248
 * mode 1 ('opt', what BilayerAnalyzer builds): 24 nearest neighbours per lipid, cell (0,0)
 * exact, every other cell searches the lipid of cell (max(0,cx-1), max(0,cy-1)) and its 24
 * neighbours; mode 0 ('exact'): nearest lipid of the leaflet per cell.  Grid geometry is
 * the float32 linspace of NumPy >= 2 (lipid_grid_opt.py:204-227).  grid_idx[F][2][nx][ny]
 * int32 (leaflet 0 = upper), grid_z[F][2][nx][ny] float64 = com_unwrap[norm] of that lipid.
 * knn_scratch: int32 [F][N][24] (mode 1 only). */
int pbk_lipid_grid(pbk_ctx *ctx, const double *com, const double *com_unwrap,
                   const uint8_t *labels, const float *box, int64_t F, int64_t N, int lat0,
                   int lat1, int norm, int nx, int ny, int mode, int32_t *knn_scratch,
                   int32_t *grid_idx, double *grid_z, void *stream);

/* ---- N3  Halperin-Nelson hexatic invariant ------------------------------------------------
 * HalperinNelsonProtocol._6nn / run_analysis   pybilt/bilayer_analyzer/analysis_protocols.py:4362-4425
 * pbk_knn24: the 24 nearest same-class neighbours of every lipid, knn[F][N][24] int32 (-1 padded),
 * in (distance, index) order, each pair evaluated as dist(lower index, higher index) -- the table of
 * lipid_grid_opt._knn, exported on its own; `labels` is any per-frame class byte (leaflet labels, or
 * a membership mask repeated for every frame).
 * pbk_halperin_nelson: out[f] = mean over the lipids with member[i] != 0 of
 * |(1/6) sum_{j in first six neighbours} exp(6 i a_ij)|^2, a_ij = x component of the unit vector
 * from i to j in the reference's arithmetic (it feeds that cosine to exp as an angle). */
int pbk_knn24(pbk_ctx *ctx, const double *com, const uint8_t *labels, const float *box, int64_t F, int64_t N,
              int lat0, int lat1, int32_t *knn, void *stream);
int pbk_halperin_nelson(pbk_ctx *ctx, const double *com, const uint8_t *member, const float *box,
                        const int32_t *knn, int64_t F, int64_t N, int lat0, int lat1, int64_t n_members,
                        double *out, void *stream);

/* ---- G3  grid reductions ----------------------------------------------------------------
 * LipidGrids.average_thickness / area_per_lipid   pybilt/lipid_grid/lipid_grid_opt.py:496-526,553-597
 * thickness[F][2] = (mean, population std) over cells of grid_z[upper]-grid_z[lower];
 * cells_per_lipid[F][N] int32; apl[F][1+2*T]: composite Welford mean, then per type code
 * (mean, sample deviation) in the float32/float64 promotion order of the reference. */
int pbk_grid_reduce(pbk_ctx *ctx, const int32_t *grid_idx, const double *grid_z,
                    const uint8_t *labels, const int32_t *type_code, const float *box, int64_t F,
                    int64_t N, int lat0, int lat1, int nx, int ny, int n_types,
                    int32_t *cells_per_lipid, double *thickness, double *apl, void *stream);

/* ---- (e)  exchanges of a frame-sharded run over NVLink peer memory -------------------------
 * The reference has no distributed path (its only parallelism is multiprocessing.Pool over frame
 * ranges, pybilt/com_trajectory/COMTraj.py:2139-2232); these are the exchange steps of SURVEY.md 8e
 * as pybilt_b200/distributed.py defines them, one process per GPU on one node.
 * pbk_comm_create allocates this rank's exchange window (world slots of slot_bytes + flags);
 * pbk_comm_handle returns its 64-byte CUDA IPC handle, which the host exchanges by any means
 * (torch.distributed.all_gather in pybilt_b200/distributed.py); pbk_comm_connect maps the peers'
 * windows from handles[world][64] (the own entry is ignored).
 * Each collective is ONE kernel: it pushes the contribution into the peers' windows with plain
 * stores over NVLink, releases a per-call epoch flag, acquires the other ranks' flags and combines
 * their slots out of the own window -- no NCCL launch and no host round trip on the critical path.
 *   pbk_comm_prefix_i32  out = sum over ranks q < rank of src_q   (net periodic-image counts
 *                        [M][3] between the two phases of a shard's reduction)
 *   pbk_allreduce_u64 / pbk_allreduce_f64  out = sum over all ranks, in rank order on every rank
 *                        (partial RDF histograms; per-frame slot vectors)
 * n * sizeof(element) <= slot_bytes.  A rank that never sees a peer's flag raises PBK_ST_INTERNAL in
 * `status` instead of spinning forever.  All ranks must issue the same sequence of collectives. */
typedef struct pbk_comm pbk_comm;
int pbk_comm_create(pbk_ctx *ctx, int rank, int world, int64_t slot_bytes, pbk_comm **out);
int pbk_comm_handle(pbk_comm *comm, unsigned char *handle64);
int pbk_comm_connect(pbk_comm *comm, const unsigned char *handles);
void pbk_comm_destroy(pbk_comm *comm);
int pbk_comm_prefix_i32(pbk_comm *comm, const int32_t *src, int32_t *out, int64_t n, int32_t *status, void *stream);
int pbk_allreduce_u64(pbk_comm *comm, const unsigned long long *src, unsigned long long *out, int64_t n,
                      int32_t *status, void *stream);
int pbk_allreduce_f64(pbk_comm *comm, const double *src, double *out, int64_t n, int32_t *status, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PBK_H */
