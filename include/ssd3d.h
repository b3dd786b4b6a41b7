/*
 * ssd3d.h -- C ABI of libssd3d.so, the B200 (sm_100a) implementation of the set-abstraction
 * operators of dvlab-research/3DSSD.
 *
 * Drop-in boundary (seam iii): the reference's TF custom ops call free C++
 * "Launcher" functions that take plain ints and raw device pointers; each entry point below
 * replaces one of them with the same argument list, plus a CUDA stream and an int status
 * (0 = cudaSuccess, otherwise a cudaError_t or SSD3D_ERR_*).  Citations are file:line under
 * /root/reference/lib/utils/tf_ops/ (declaration in the .cpp that the TF op binds, definition at
 * the tail of the matching *_g.cu).
 *
 * Conventions, identical to the reference: dense row-major tensors, batch-major; fp32 values,
 * int32 indices; every pointer is a DEVICE pointer owned by the caller; the library allocates
 * nothing, keeps no state between calls, never synchronises the host, and is CUDA-graph
 * capturable.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 */
#ifndef SSD3D_H_
#define SSD3D_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSD3D_ERR_INVALID_ARGUMENT (-1) /* mirrors the reference's OP_REQUIRES -> InvalidArgument */
#define SSD3D_ERR_UNSUPPORTED      (-2)

typedef void *ssd3d_stream_t;

/* ABI version of this header (bumped on any signature change). */
int ssd3d_version(void);   /* 2 since round 2 (explicit-placement *_ex entry points, last_scale_nonneg) */
/* Human-readable description of the last non-zero status returned on this thread. */
const char *ssd3d_last_error(void);

/* ---- sampling ------------------------------------------------------------------------------ */

/* replaces farthestpointsamplingLauncher(b,n,c,m,inp,temp,out)
 *   sampling/tf_sampling.cpp:131 (decl), sampling/tf_sampling_g.cu:392-394 (def), kernel :124-178.
 * D-FPS over inp[b,n,c] (any c; c==3 takes the on-chip cluster kernel).  out[b,m] int32.
 * temp[b,n] fp32 scratch is only touched by the large-n fallback; it may be NULL when
 * ssd3d_fps_needs_temp(n,c)==0.  Bit-exact, including the reference's tie-break (value desc,
 * k mod 1024 asc, k asc). */
int ssd3d_farthest_point_sample(int b, int n, int c, int m, const float *inp, float *temp, int *out,
                                ssd3d_stream_t stream);
int ssd3d_fps_needs_temp(int n, int c);

/* The same operator with explicit placement -- what lets an SA layer run its sampling without any copy / add /
 * concat kernels around it (lib/utils/layers_util.py:84-111):
 *   in_stride   floats between consecutive scenes of inp (>= n*c): a [:, a:b] slice of a dense [b,N,c] tensor;
 *   ldo         ints between consecutive rows of out (>= m): a column block of the concatenated fps_idx tensor;
 *   idx_offset  added to every stored index (the `+ last_fps_end_index` of :109);
 *   j0, j1      run only rounds [j0, j1) of 0..m; the running state travels through temp[b, E] between the launches
 *               (required when the range is partial), E = ssd3d_fps_temp_elems(n, c, m, flags) floats per scene.  A
 *               sample is final as soon as its round is done, so work on the first samples can overlap the remaining
 *               rounds.  Partial ranges need ssd3d_fps_supports_rounds(n, c);
 *   cluster     CTAs per scene: 0 = heuristic, 1/2/4/8/16 = exactly that, negative = heuristic capped at -cluster
 *               (FPS is latency-bound: fewer CTAs cost little time and leave SMs to concurrent work);
 *   flags       bit 0: use the general cluster kernel (coordinates travel in the packets) even where the
 *               resident-scene kernel applies; bit 1: never / bit 2: always (64 <= n <= 16384) take the single-CTA
 *               kernel with spatial pruning (csrc/fps_bucket.cu) that c == 3 scenes of 8192 < n <= 16384 points and
 *               m >= 256 get by default -- points grouped into spatially compact buckets of 32, a round updates only
 *               the buckets whose bounding box lies closer to the new sample than their largest running distance
 *               (the others provably keep every distance), so one SM per scene suffices; `cluster` is ignored there.
 * No state is kept in the library: every choice is an argument. */
int ssd3d_farthest_point_sample_ex(int b, int n, int c, int m, const float *inp, long long in_stride, float *temp,
                                   int *out, int ldo, int idx_offset, int j0, int j1, int cluster, int flags,
                                   ssd3d_stream_t stream);
int ssd3d_fps_supports_rounds(int n, int c);
long ssd3d_fps_temp_elems(int n, int c, int m, int flags);

/* replaces farthestpointsamplingwithdistLauncher(b,n,m,inp,temp,out)
 *   sampling/tf_sampling.cpp:164, sampling/tf_sampling_g.cu:396-398, kernel :181-230.
 * F-FPS over a precomputed distance matrix dist[b,n,n]. */
int ssd3d_farthest_point_sample_with_distance(int b, int n, int m, const float *dist, float *temp, int *out,
                                              ssd3d_stream_t stream);
/* ... with explicit output placement and cluster request (see ssd3d_farthest_point_sample_ex). */
int ssd3d_farthest_point_sample_with_distance_ex(int b, int n, int m, const float *dist, float *temp, int *out, int ldo,
                                                 int idx_offset, int cluster, ssd3d_stream_t stream);

/* The F-FPS branch of the SA layer in one call (lib/utils/layers_util.py:94-96 and :102-104):
 *   farthest_point_sample_with_distance(m, calc_square_dist(concat[fa, fb]))
 * without materialising the [b,n,n] matrix -- each round evaluates the picked point's matrix row on the fly from
 * feature rows kept on chip.  fa[b,n,ca] (xyz, ca = 3 in the model), fb[b,n,cb] (features; may be NULL when cb = 0).
 * Same indices as the two-call route (same pinned fp32 arithmetic as ssd3d_calc_square_dist).  Covers
 * ca+cb <= 68 with n <= 4096 and ca+cb <= 132 with n <= 2048 (ssd3d_ffps_supported); other shapes return
 * SSD3D_ERR_UNSUPPORTED and the caller takes the two-call route. */
int ssd3d_farthest_point_sample_features(int b, int n, int ca, int cb, int m, const float *fa, const float *fb,
                                         int *out, ssd3d_stream_t stream);
/* ... with scene strides of fa / fb in floats, explicit output placement and a range of rounds [j0, j1) with the
 * running distances in temp[b,n] (see ssd3d_farthest_point_sample_ex; temp may be NULL for the full range). */
int ssd3d_farthest_point_sample_features_ex(int b, int n, int ca, int cb, int m, const float *fa, long long fa_stride,
                                            const float *fb, long long fb_stride, float *temp, int *out, int ldo,
                                            int idx_offset, int j0, int j1, ssd3d_stream_t stream);
int ssd3d_ffps_supported(int n, int c);

/* replaces gatherpointLauncher(b,n,m,c,inp,idx,out)
 *   sampling/tf_sampling.cpp:235, sampling/tf_sampling_g.cu:403-407, kernel :320-331.
 * out[b,m,c] = inp[b, idx[b,m], c] */
int ssd3d_gather_point(int b, int n, int m, int c, const float *inp, const int *idx, float *out,
                       ssd3d_stream_t stream);
/* ... reading idx rows ld_idx ints apart (a column block of a wider [b, L] index tensor). */
int ssd3d_gather_point_ex(int b, int n, int m, int c, const float *inp, const int *idx, int ld_idx, float *out,
                          ssd3d_stream_t stream);

/* ---- grouping ------------------------------------------------------------------------------ */

/* replaces queryBallPointLauncher(b,n,m,radius,nsample,xyz1,xyz2,idx,pts_cnt)
 *   grouping/tf_grouping.cpp:270, grouping/tf_grouping_g.cu:461-464, kernel :215-255.
 * idx[b,m,nsample], pts_cnt[b,m].  Rows with pts_cnt==0 are written as zeros (the reference leaves
 * them uninitialised and its caller masks them, lib/utils/layers_util.py:157-159). */
int ssd3d_query_ball_point(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2,
                           int *idx, int *pts_cnt, ssd3d_stream_t stream);

/* replaces queryBallPointDilatedLauncher(b,n,m,min_radius,max_radius,nsample,xyz1,xyz2,idx,pts_cnt)
 *   grouping/tf_grouping.cpp:363, grouping/tf_grouping_g.cu:465-468, kernel :308-357. */
int ssd3d_query_ball_point_dilated(int b, int n, int m, float min_radius, float max_radius, int nsample,
                                   const float *xyz1, const float *xyz2, int *idx, int *pts_cnt,
                                   ssd3d_stream_t stream);

/* B200 fast path used by pointnet_sa_module_msg: up to 4 (min_r,max_r,nsample) queries over the same
 * (xyz1,xyz2) answered in ONE pass over the candidates (the reference scans once per radius,
 * lib/utils/layers_util.py:134-147).  dilated!=0 selects the :308-357 predicate, else the :215-255 one
 * (min_radius ignored).  idx[s] / pts_cnt[s] are per-query outputs as above. */
int ssd3d_query_ball_point_multi(int b, int n, int m, int nqueries, int dilated, const float *min_radius,
                                 const float *max_radius, const int *nsample, const float *xyz1, const float *xyz2,
                                 int *const *idx, int *const *pts_cnt, ssd3d_stream_t stream);
/* The same search with spatial culling (csrc/ball_query_grid.cu) for large candidate sets: `workspace` is caller-owned
 * scratch of ssd3d_query_ball_point_workspace(b, n) bytes (16-byte aligned; 0 = this size has no culled kernel).  One
 * small kernel bins the candidates of every scene into a uniform 2-D grid with cells >= r_max, the search then visits
 * only the 3x3 cell neighbourhood of a query and restores "first nsample hits in ascending index" through a per-shell
 * bitmap over candidate indices.  Same outputs, bit for bit, non-finite coordinates included.  workspace == NULL
 * falls through to ssd3d_query_ball_point_multi.
 * units (optional, host array of nqueries device pointers, entries may be NULL): per-shell UNIT LISTS for the grouped MLP.
 * units[s] holds 1 + b*m*ceil(nsample[s]/8) ints: [0] = number of units, [1 + u] = (group << 4) | j naming rows 8j..8j+7
 * of neighbour list `group` (= scene*m + query).  A group with cnt hits gets ceil(cnt/8) units -- slots beyond cnt repeat
 * the first hit (tf_grouping_g.cu:245-248) and cannot change the max-pool that follows (layers_util.py:178), so the
 * grouped MLP needs only these rows (ssd3d_sa_mlp_fused*, ssd3d_linear_tc*_units).  The list order is unspecified.
 * Without a workspace the list is built from the finished counts by one more small kernel. */
size_t ssd3d_query_ball_point_workspace(int b, int n);
int ssd3d_query_ball_point_multi_ws(int b, int n, int m, int nqueries, int dilated, const float *min_radius,
                                    const float *max_radius, const int *nsample, const float *xyz1, const float *xyz2,
                                    int *const *idx, int *const *pts_cnt, int *const *units, void *workspace,
                                    size_t workspace_bytes, ssd3d_stream_t stream);

/* replaces groupPointLauncher(b,n,c,m,nsample,points,idx,out)
 *   grouping/tf_grouping.cpp:446, grouping/tf_grouping_g.cu:476-479, kernel :362-379.
 * out[b,m,nsample,c] = points[b, idx[b,m,nsample], c]; idx==-1 -> 0 */
int ssd3d_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out,
                      ssd3d_stream_t stream);

/* ---- interpolation ------------------------------------------------------------------------- */

/* replaces ThreeNNLauncher(b,n,m,xyz1,xyz2,dist,idx)
 *   interpolation/tf_interpolate.cpp:215, interpolation/tf_interpolate_g.cu:191-195, kernel :24-84. */
int ssd3d_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx,
                   ssd3d_stream_t stream);

/* replaces ThreeInterpolateLauncher(b,m,c,n,points,idx,weight,out)
 *   interpolation/tf_interpolate.cpp:285, interpolation/tf_interpolate_g.cu:198-200, kernel :87-113. */
int ssd3d_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx, const float *weight,
                            float *out, ssd3d_stream_t stream);

/* ---- dense parts of the SA layer (TF stock ops in the reference) ---------------------------- */

/* model_util.calc_square_dist(a, a, norm=False)  (lib/utils/model_util.py:144-160) for a[b,n,c]:
 * out[b,i,j] = (|a_i|^2 + |a_j|^2) - 2 a_i.a_j, fp32, channel sums as sequential fma chains
 * (order pinned so the oracle can be bit-exact; the reference's cuBLAS order is unspecified). */
int ssd3d_calc_square_dist(int b, int n, int c, const float *a, float *out, ssd3d_stream_t stream);

/* Fused gather + concat of lib/utils/layers_util.py:160-165:
 *   x[b,m,k,:] = concat( points[b, idx[b,m,k], 0:c],  xyz[b, idx[b,m,k], :] - new_xyz[b,m,:] )
 * written with row stride ldx (>= c+3; columns c+3..ldx-1 are zero-filled). */
int ssd3d_group_concat(int b, int n, int c, int m, int nsample, const float *xyz, const float *points,
                       const float *new_xyz, const int *idx, float *x, int ldx, ssd3d_stream_t stream);

/* One tf_util.conv2d/conv1d 1x1 layer at inference (lib/utils/tf_util.py:51-201, BN :424-444) with
 * bias and BatchNorm pre-folded by the caller into per-channel (scale, shift):
 *   y[r, o] = act( (sum_k x[r,k] * w[k,o]) * scale[o] + shift[o] ),  act = relu or identity.
 * x[rows, ldx] (first cin columns used), w[cin, cout] (the TF kernel [1,1,cin,cout]), y[rows, ldy].
 * pool > 1: instead of y, write ymax[rows/pool, cout] = max over each run of `pool` consecutive rows
 * (tf.reduce_max(axis=2), layers_util.py:178) times rowmask[rows/pool] when rowmask != NULL (:180). */
int ssd3d_linear_bn_relu(long rows, int cin, int cout, const float *x, int ldx, const float *w,
                         const float *scale, const float *shift, int relu, int pool, const int *rowmask,
                         float *y, int ldy, ssd3d_stream_t stream);

/* ---- tensor-core (tcgen05) path of the same layer -------------------------------------------------
 * fp32 operands travel as TWO bf16 matrices (x = hi + lo, 16 mantissa bits); the product is evaluated as
 * hi.hi + lo.hi + hi.lo with fp32 accumulation in TMEM (error ~2^-16, inside the 1e-3 parity budget).
 *   a_hi/a_lo [rows, kp] bf16 row-major (kp % 16 == 0, zero padded); b_hi/b_lo [n, kp] bf16 = W^T split;
 *   outputs (any subset): out_f32 [rows or rows/pool, ld_f32] and/or out_hi/out_lo [.., ld_split] bf16 for the
 *   next layer (columns n..ld_split-1 are written as zeros when pool == 1).  pool / rowmask as above. */
int ssd3d_linear_tc(long rows, int kp, int n, const void *a_hi, const void *a_lo, const void *b_hi, const void *b_lo,
                    const float *scale, const float *shift, int relu, int pool, const int *rowmask, float *out_f32,
                    int ld_f32, void *out_hi, void *out_lo, int ld_split, ssd3d_stream_t stream);
/* ssd3d_linear_tc for the FIRST layer of an SA scale with the gather fused into the operand load: the A operand
 * x[b,m,k,:] = concat(points[b, idx[b,m,k], :], xyz[b, idx[b,m,k], :] - new_xyz[b,m,:]) (layers_util.py:160-165) is
 * built in shared memory by producer warps instead of being materialised in HBM.  b_hi/b_lo [nout, round16(c+3)]. */
int ssd3d_linear_tc_gather(int b, int n, int c, int m, int nsample, const float *xyz, const float *points,
                           const float *new_xyz, const int *idx, int nout, const void *b_hi, const void *b_lo,
                           const float *scale, const float *shift, int relu, int pool, const int *rowmask,
                           float *out_f32, int ld_f32, void *out_hi, void *out_lo, int ld_split, ssd3d_stream_t stream);

/* "Hoisted first layer": with the first conv of an SA scale split by input rows, W1 = [Wf ; Wx],
 *   relu((concat(f_j, x_j - c_i) . W1) * s1 + t1) = relu(z[j] + (x_j - c_i) . (Wx * s1)),   z = (f . Wf) * s1 + t1,
 * the feature part is one small per-POINT GEMM (ssd3d_linear_tc on the [b*n, c] features, no ReLU) instead of a
 * per-grouped-row one, and this call runs the SECOND conv of the scale with its operand rebuilt on the fly from z
 * (z[b,n,ldz], this scale's n1 columns start at the pointer; wx = Wx*s1 as [3][n1]).  Replaces conv #1 and #2 of
 * layers_util.py:167-176 plus the grouping of :157-165; outputs as ssd3d_linear_tc. */
int ssd3d_linear_tc_hoisted(int b, int n, int n1, int m, int nsample, const float *xyz, const float *z, int ldz,
                            const float *wx, const float *new_xyz, const int *idx, int nout, const void *b_hi,
                            const void *b_lo, const float *scale, const float *shift, int relu, int pool,
                            const int *rowmask, float *out_f32, int ld_f32, void *out_hi, void *out_lo, int ld_split,
                            ssd3d_stream_t stream);
/* The operand of ssd3d_linear_tc_hoisted materialised instead of built inside the GEMM: hi/lo[b*m*nsample, kp] bf16 =
 * split(relu(z[idx] + (xyz[idx] - new_xyz) . wx)), zero padded to kp.  For wide layers (K >= 256) an elementwise pass at
 * memory speed + the plain TMA-fed ssd3d_linear_tc beats in-kernel production (lib/utils/layers_util.py:160-176). */
int ssd3d_hoist_expand_split(int b, int n, int n1, int m, int nsample, const float *xyz, const float *z, int ldz,
                             const float *wx, const float *new_xyz, const int *idx, void *hi, void *lo, int kp,
                             ssd3d_stream_t stream);
/* UNIT-LIST forms of the three calls above (units: ssd3d_query_ball_point_multi_ws).  The grouped matrix then has
 * units[0] * 8 rows: compact row 8u + e is neighbour slot 8j + e of the group unit u names, and the buffers keep their
 * full capacity (rows / b*m*nsample rows).  ssd3d_hoist_expand_split_units and ssd3d_linear_tc_hoisted_units look their
 * source rows up through the list; ssd3d_linear_tc_units consumes compact rows.  unit_pool != 0 (last conv of a scale,
 * layers_util.py:176-180): every 8-row unit is max-pooled and combined into out_f32[group, :] with atomicMax -- requires
 * relu != 0 (values >= 0), no split output, and out_f32 ZERO-FILLED by the caller (which also is the cnt == 0 mask). */
int ssd3d_hoist_expand_split_units(int b, int n, int n1, int m, int nsample, const float *xyz, const float *z, int ldz,
                                   const float *wx, const float *new_xyz, const int *idx, const int *units, void *hi,
                                   void *lo, int kp, ssd3d_stream_t stream);
int ssd3d_linear_tc_units(long rows, int kp, int n, const void *a_hi, const void *a_lo, const void *b_hi, const void *b_lo,
                          const float *scale, const float *shift, int relu, const int *units, int unit_pool,
                          float *out_f32, int ld_f32, void *out_hi, void *out_lo, int ld_split, ssd3d_stream_t stream);
int ssd3d_linear_tc_hoisted_units(int b, int n, int n1, int m, int nsample, const float *xyz, const float *z, int ldz,
                                  const float *wx, const float *new_xyz, const int *idx, const int *units, int nout,
                                  const void *b_hi, const void *b_lo, const float *scale, const float *shift, int relu,
                                  int unit_pool, float *out_f32, int ld_f32, void *out_hi, void *out_lo, int ld_split,
                                  ssd3d_stream_t stream);
/* Same idea for a scale that fits the fused kernel (ssd3d_sa_mlp_fused): the stack passed here starts at the scale's
 * SECOND conv, the first operand row is relu(z[idx] + (xyz[idx] - new_xyz) . wx), built during the gather. */
int ssd3d_sa_mlp_fused_hoisted(int b, int n, int n1, int m, int nsample, const float *xyz, const float *z, int ldz,
                               const float *wx, const float *new_xyz, const int *idx, const int *pts_cnt, const int *units,
                               int nl, const int *nout, const void *w_blob, const float *ss_blob, int last_scale_nonneg,
                               float *out_f32, int ld_f32, void *out_hi, void *out_lo, int ld_split, ssd3d_stream_t stream);
/* hi/lo[row, 0:kp] = split(x[row, 0:c]), zero padded (kp % 8 == 0). */
int ssd3d_split_rows(long rows, int c, const float *x, int ldx, void *hi, void *lo, int kp, ssd3d_stream_t stream);
/* ssd3d_group_concat fused with the split: hi/lo [b*m*nsample, kp] bf16. */
int ssd3d_group_concat_split(int b, int n, int c, int m, int nsample, const float *xyz, const float *points,
                             const float *new_xyz, const int *idx, void *hi, void *lo, int kp, ssd3d_stream_t stream);

/* A whole SA scale in one kernel (gather + concat + up to 3 conv/BN/ReLU layers + max-pool + mask,
 * lib/utils/layers_util.py:157-180) for layer stacks whose weights fit in shared memory.
 * ssd3d_sa_fused_smem returns the shared-memory bytes needed, or 0 if the stack does not fit (use the per-layer
 * path then).  w_blob / ss_blob are the pre-swizzled split weights and folded scale/shift built by the host
 * (3dssd_b200/params.py: FusedStack).  Outputs as ssd3d_linear_tc with pool = nsample.
 * last_scale_nonneg != 0: the caller guarantees scale >= 0 in the LAST layer (fold the sign of a negative BatchNorm
 * gamma into that layer's weight column -- FusedStack does); the kernel then max-pools the raw accumulators and applies
 * scale / shift / ReLU to the pooled values only (exactly the same result: the affine map is monotone).
 * units != NULL: the unit list of this scale from ssd3d_query_ball_point_multi_ws.  The kernel then convolves only the
 * listed 8-row units instead of all nsample rows of every group -- the rows it skips repeat a group's first neighbour and
 * cannot change the max-pool -- and combines the units of a group with atomicMax on out_f32, which the caller must have
 * ZERO-FILLED (zero is also the masked result of a group without neighbours); out_hi / out_lo must be NULL.  Identical
 * results; the work shrinks from b*m*nsample rows to 8 * sum(ceil(cnt / 8)). */
size_t ssd3d_sa_fused_smem(int c, int nl, const int *nout);
int ssd3d_sa_mlp_fused(int b, int n, int c, int m, int nsample, const float *xyz, const float *points,
                       const float *new_xyz, const int *idx, const int *pts_cnt, const int *units, int nl,
                       const int *nout, const void *w_blob, const float *ss_blob, int last_scale_nonneg, float *out_f32,
                       int ld_f32, void *out_hi, void *out_lo, int ld_split, ssd3d_stream_t stream);

/* ---- backward operators (training graphs) ---------------------------------------------------
 * Each zero-fills its output first, as the reference's TF ops do with cudaMemset before the launcher. */

/* replaces scatteraddpointLauncher(b,n,m,c,out_g,idx,inp_g)
 *   sampling/tf_sampling.cpp:261 (decl, memset :286), sampling/tf_sampling_g.cu:408-410, kernel :335-346.
 * inp_g[b, idx[b,j], :] += out_g[b,j,:] */
int ssd3d_gather_point_grad(int b, int n, int m, int c, const float *out_g, const int *idx, float *inp_g,
                            ssd3d_stream_t stream);

/* replaces groupPointGradLauncher(b,n,c,m,nsample,grad_out,idx,grad_points)
 *   grouping/tf_grouping.cpp:479 (memset :510), grouping/tf_grouping_g.cu:480-484, kernel :383-398 (idx == -1 skipped). */
int ssd3d_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                           float *grad_points, ssd3d_stream_t stream);

/* replaces ThreeInterpolateGradLauncher(b,n,c,m,grad_out,idx,weight,grad_points)
 *   interpolation/tf_interpolate.cpp:363 (memset :398), interpolation/tf_interpolate_g.cu:202-206, kernel :115-140. */
int ssd3d_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx,
                                 const float *weight, float *grad_points, ssd3d_stream_t stream);

/* Per-scene greedy BEV NMS on the GPU: replaces box_3d_to_anchor + project_to_bev + tf.image.non_max_suppression
 * of lib/builder/postprocessor.py:76-88 (one class per call).  boxes [b,n,7] = (x,y,z,l,h,w,ry), scores [b,n];
 * candidates in descending score order (ties: lower index), dropped when axis-aligned BEV IoU with a kept box is
 * > iou_threshold, at most max_output kept.  out_block [b,max_output,9] = (box7, score, class) zero padded,
 * out_cnt [b] = number kept.  n <= 512. */
int ssd3d_bev_nms(int b, int n, const float *boxes, const float *scores, float iou_threshold, int max_output,
                  int cls_id, float *out_block, int *out_cnt, ssd3d_stream_t stream);

/* ymax[g, 0:c] = max over rows g*pool .. g*pool+pool-1 of y[., 0:c] (times rowmask[g] != 0): the
 * tf.reduce_max(axis=2) * mask of layers_util.py:178-180 for nsample values the fused epilogue of
 * ssd3d_linear_bn_relu does not cover (pool must divide 128 there). */
int ssd3d_rowgroup_max(long groups, int pool, int c, const float *y, int ldy, const int *rowmask, float *out,
                       ssd3d_stream_t stream);

/* Training-mode BatchNorm + activation of one conv output (lib/utils/tf_util.py:424-444 with is_training=True:
 * tf.contrib.layers.batch_norm(decay, updates_collections=None, fused=False), epsilon 0.001 passed as `eps`):
 * batch mean / population variance of x[rows, c] over the rows, out = act(x * inv + (beta - mean * inv)) with
 * inv = gamma * rsqrt(var + eps), and -- when moving_mean / moving_var are given -- their in-place update
 * v -= (v - batch) * (1 - decay).  scale / shift [c] receive (inv, beta - mean*inv): the folded form the inference
 * kernels take.  batch_mean / batch_var [c] are optional outputs.  workspace: ssd3d_bn_train_workspace(c) bytes.
 * x and out may alias.  Statistics couple all rows: a batch sharded over GPUs needs its own reduction (not built). */
size_t ssd3d_bn_train_workspace(int c);
int ssd3d_bn_train(long rows, int c, const float *x, int ldx, const float *gamma, const float *beta, float *moving_mean,
                   float *moving_var, float decay, float eps, void *workspace, float *scale, float *shift,
                   float *batch_mean, float *batch_var, int relu, float *out, int ldo, ssd3d_stream_t stream);

/* ---- the small elementwise stages of the path (csrc/misc.cu) --------------------------------- */

/* points[rows, c] -> xyz[rows, 3], feat[rows, c-3]: the tf.slice pair of
 * lib/modeling/single_stage_detector.py:116-117.  feat may be NULL when c == 3. */
int ssd3d_split_points(long rows, int c, const float *points, float *xyz, float *feat, ssd3d_stream_t stream);
/* x[0:count] = 0 (fp32): the zero fill the unit-list mode of ssd3d_sa_mlp_fused needs, as a kernel of this library. */
int ssd3d_fill_zero(float *x, long count, ssd3d_stream_t stream);
/* out[s*ldo + j] = start + j, s < b, j < m: tf.tile(tf.range(npoint)) of lib/utils/layers_util.py:91-92, :100-101. */
int ssd3d_iota_idx(int b, int m, int start, int *out, int ldo, ssd3d_stream_t stream);
/* out[b,n,ca+cb] = concat(a[b,n,ca], bsrc[b,n,cb]) with scene strides in floats: tf.concat([xyz, points], -1) in
 * front of calc_square_dist (lib/utils/layers_util.py:94, :102). */
int ssd3d_concat_cols(int b, int n, int ca, int cb, const float *a, long long a_stride, const float *bsrc,
                      long long b_stride, float *out, ssd3d_stream_t stream);
/* out[b, sum(m), c] = concat along axis 1 of parts src[i][b, m[i], c] (i < parts <= 8; src / m are HOST arrays of
 * device pointers / row counts): joins the per-part results of an SA layer whose sampling was consumed in parts. */
int ssd3d_concat_rows(int b, int parts, const float *const *src, const int *m, int c, float *out, ssd3d_stream_t stream);
/* out[r, 0:3] = xyz[r, 0:3] + min(max(offsets[r, 0:3], min_xyz), -min_xyz): the clamp + add of vote_layer,
 * lib/utils/layers_util.py:20-23 (min_xyz = MODEL.MAX_TRANSLATE_RANGE, negative). */
int ssd3d_vote_translate(long rows, const float *xyz, const float *offsets, int ld_offsets, float min_x, float min_y,
                         float min_z, float *out, ssd3d_stream_t stream);
/* decode_dist_anchor_free + decode_class2angle + the score sigmoid (lib/utils/anchor_decoder.py:86-112, :6-14,
 * lib/modeling/single_stage_detector.py:210-211) in one pass.  pred_reg[rows, ld_reg] = [6 face distances |
 * angle_bins logits | angle_bins residuals], pred_cls[rows, ld_cls] (column 0 = the class logit), center_xyz[rows,3]
 * -> boxes[rows,7] = (x, y, z, l, h, w, ry), scores[rows]. */
int ssd3d_decode_dist_anchor_free(long rows, int angle_bins, const float *center_xyz, const float *pred_reg, int ld_reg,
                                  const float *pred_cls, int ld_cls, float *boxes, float *scores, ssd3d_stream_t stream);

/* ---- sharded step: exchange over NVLink peer memory ---------------------------------------- */

/* All-gather of one slice per rank as ONE kernel (csrc/peer_gather.cu): no reference counterpart (the reference's
 * inference is single-GPU, lib/core/evaluator.py:145-147); replaces the ncclAllGather of the per-scene detection blocks at
 * the end of a sharded step.  peer_base: DEVICE array [world] with the base address of every rank's symmetric buffer as
 * mapped in this process (e.g. torch.distributed._symmetric_memory buffer_ptrs); inside it, for parity q in {0,1}:
 * recv_off<q> = world slices of slice_bytes, flag_off<q> = world int32 flags, all zero before the first call.  The call
 * stores src into slot `rank` of every peer's receive area, publishes replay number s (kept in state[0], so a captured
 * launch needs no changing argument; state = 3 zero-initialised ints owned by this exchange) with release semantics at
 * system scope, waits for the peers' flags and copies the received slices to out[world][slice_bytes].  Every rank must
 * make the same sequence of calls; two parities make reuse safe without acknowledgements.  A wait of ~4 s gives up and
 * increments state[2] instead of hanging the GPU. */
int ssd3d_peer_allgather(const void *src, size_t slice_bytes, void *const *peer_base, int world, int rank,
                         size_t recv_off0, size_t recv_off1, size_t flag_off0, size_t flag_off1, int *state, void *out,
                         ssd3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SSD3D_H_ */
