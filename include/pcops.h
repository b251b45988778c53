/*
 * pcops.h -- C ABI of libpcops.so: the PointNet++ set-abstraction / feature-propagation geometry ops
 * (and the per-neighbourhood attention contraction) as hand-written sm_100a CUDA kernels.
 *
 * This is the drop-in seam between a framework's op kernels and the device code.  In the reference
 * (tpfeifle/pointcloud-segmentation-attention, paths below relative to pointnet2_tensorflow/tf_ops/) that seam is
 * the set of C++ "launcher" prototypes each TensorFlow OpKernel::Compute calls; every entry point here names the
 * launcher (file:line) it replaces and keeps that launcher's (b, n, m, ...) argument order, so a TF OpKernel, a
 * ctypes stub or a torch wrapper forwards its arguments 1:1.
 *
 * Conventions
 *   - all pointers are DEVICE pointers to dense row-major float32 / int32 arrays, 4-byte aligned; rows need not
 *     be 16-byte aligned (vector paths are chosen at run time from the actual addresses);
 *   - the caller owns every buffer, including workspaces (size from the matching *_workspace_bytes query);
 *     the library allocates nothing and keeps no state besides per-device kernel attributes set lazily;
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and the call returns immediately;
 *   - outputs are fully overwritten (gradient outputs need no pre-zeroing by the caller);
 *   - indices are trusted to lie in range exactly as in the reference kernels;
 *   - return value: PC_OK (0); a negative PC_ERR_* for a rejected argument (nothing enqueued); or a positive
 *     cudaError_t from the launch.  Calls with an empty problem (b*n*m == 0 ...) return PC_OK without launching.
 *   - re-entrant and thread-safe; no host synchronisation inside any call (CUDA-graph capturable after one
 *     warm-up call per device).
 *
 * Arithmetic contract: fp32, never fused (no FMA contraction), evaluated in the reference's written order, so that
 * FPS / ball-query / kNN / three_nn / gather indices and grouped or gathered values are bit-exact against the
 * reference ALGORITHM in un-fused arithmetic -- i.e. against the reference's CPU code, or its .cu files compiled with
 * --fmad=false.  The reference's own compile scripts use plain `nvcc -O2` (fmad on), which contracts the distance
 * into FMAs: against such a build FPS tie / arg-max choices and ball-boundary hits can differ.
 * three_interpolate, all gradients and the attention contraction are deterministic (fixed summation order, no float
 * atomics).  Non-finite coordinates follow the registered ops: a NaN distance is a ball-query hit (max = fmaxf,
 * tf_grouping_g.cu:24), three_nn never selects an infinite or NaN distance (tf_interpolate.cpp:66,74).
 * The tensor-core entry points (pc_dense_*, pc_attention_layer_*) compute 3xTF32 split products with fp32
 * accumulation: ~2e-6 of the output scale against a float64 product (documented per call).
 */
#ifndef PCOPS_H_
#define PCOPS_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PC_OK 0
#define PC_ERR_INVALID_ARGUMENT (-1) /* shape / attribute the reference op would reject with InvalidArgument */
#define PC_ERR_UNSUPPORTED (-2)      /* valid, but outside what this build implements (documented per call)   */
#define PC_ERR_WORKSPACE (-3)        /* workspace pointer is NULL although *_workspace_bytes() > 0             */

typedef void *pc_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define PC_API __attribute__((visibility("default")))
#else
#define PC_API
#endif

PC_API int pc_version(void);                   /* 100*major + minor */
PC_API const char *pc_error_string(int code);  /* PC_ERR_* names, or cudaGetErrorString for positive codes */
PC_API int pc_num_sms(void);                   /* SM count of the current device (grid sizing / tests); <0 on error */
/* Concurrency hint (process-wide, default 1): how many independent launches of this library the caller keeps in
 * flight on the device at once (e.g. the number of batches a multi-stream / multi-graph pipeline overlaps).  Results
 * never depend on it.  The streaming kernels (group_point, three_interpolate, the feature pro- / epilogues, the attention
 * contraction) size their grids and shared-memory rings for a lone launch when it is 1, and as fewer, longer-lived CTAs
 * with small reservations when it is larger: inside a mix of co-resident kernels that raised the whole ScanNet step
 * by 4 %, while a lone launch sized that way is up to 2 x slower.  n >= 1, else PC_ERR_INVALID_ARGUMENT. */
PC_API int pc_set_concurrency_hint(int n);
PC_API int pc_get_concurrency_hint(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Sampling library (reference: sampling/tf_sampling_g.cu, sampling/tf_sampling.cpp)
 * ------------------------------------------------------------------------------------------------------------- */

/* Farthest point sampling.  Replaces farthestpointsamplingLauncher(b,n,m,inp,temp,out)
 * (tf_sampling_g.cu:203-205, prototype tf_sampling.cpp:94; op FarthestPointSample tf_sampling.cpp:28-40,95-123).
 *   xyz (b,n,3) f32 -> out_idx (b,m) i32.  out_idx[:,0] = 0; ties resolve to the smallest (k mod 512, k) as the
 *   reference's 512-thread partition + left-biased tree does (tf_sampling_g.cu:130,153-163).
 * workspace: pc_fps_workspace_bytes(b,n,m) bytes (0 for n <= 262144: all state stays on chip -- one CTA per scene up
 * to 8192 points, one thread-block cluster of 2..16 CTAs exchanging winners through DSMEM beyond); the reference needed
 * a (32,n) f32 temp tensor (tf_sampling.cpp:115).  m <= 0 returns PC_OK at once (tf_sampling_g.cu:106-107). */
PC_API size_t pc_fps_workspace_bytes(int b, int n, int m);
PC_API int pc_fps(int b, int n, int m, const float *xyz, void *workspace, int *out_idx, pc_stream_t stream);

/* farthest_point_sample + gather_point in one launch (the pair pointnet_util.py:34 always issues together):
 * additionally writes out_xyz (b,m,3) = xyz[b, out_idx[b,j], :] when out_xyz is not NULL.  Same indices as pc_fps. */
PC_API int pc_fps_gather(int b, int n, int m, const float *xyz, void *workspace, int *out_idx, float *out_xyz,
                  pc_stream_t stream);

/* ProbSample: categorical sampling by inverse CDF.  Replaces probsampleLauncher(b,n,m,inp_p,inp_r,temp,out)
 * (tf_sampling_g.cu:197-200 = cumsumKernel :7-88 + binarysearchKernel :90-104; prototype tf_sampling.cpp:65; op
 * ProbSample tf_sampling.cpp:14-27,66-92).
 *   inp_p (b,n) f32 weights, inp_r (b,m) f32 uniforms -> out (b,m) i32: smallest index whose cumulative weight is
 *   >= r * total.  temp: (b,n) f32, receives the cumulative sums (the reference's allocate_temp, tf_sampling.cpp:85-88);
 *   they follow the reference's blocked summation tree bit for bit.  pc_cumsum is cumsumLauncher (:193-195) alone. */
PC_API int pc_cumsum(int b, int n, const float *inp, float *out, pc_stream_t stream);
PC_API int pc_prob_sample(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp, int *out,
                   pc_stream_t stream);

/* gather_point.  Replaces gatherpointLauncher(b,n,m,inp,idx,out) (tf_sampling_g.cu:206-208, prototype
 * tf_sampling.cpp:125; op GatherPoint tf_sampling.cpp:41-54,126-148).
 *   inp (b,n,3), idx (b,m) -> out (b,m,3);  out[b,j,:] = inp[b,idx[b,j],:] */
PC_API int pc_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out, pc_stream_t stream);

/* gather_point gradient.  Replaces cudaMemset + scatteraddpointLauncher(b,n,m,out_g,idx,inp_g)
 * (tf_sampling.cpp:174-175, tf_sampling_g.cu:209-211, prototype tf_sampling.cpp:150; op GatherPointGrad).
 *   out_g (b,m,3), idx (b,m) -> inp_g (b,n,3), fully overwritten; contributions to one point are summed in
 *   ascending j (deterministic; the reference's float atomicAdd order is undefined). */
PC_API size_t pc_gather_point_grad_workspace_bytes(int b, int n, int m);
PC_API int pc_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g, void *workspace,
                         pc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Grouping library (reference: grouping/tf_grouping_g.cu, grouping/tf_grouping.cpp, grouping/tf_grouping.py)
 * ------------------------------------------------------------------------------------------------------------- */

/* Ball query.  Replaces queryBallPointLauncher(b,n,m,radius,nsample,xyz1,xyz2,idx,pts_cnt)
 * (tf_grouping_g.cu:125-128, prototype tf_grouping.cpp:66; op QueryBallPoint tf_grouping.cpp:13-30,67-106).
 *   xyz1 (b,n,3) dataset, xyz2 (b,m,3) queries -> idx (b,m,nsample) i32, pts_cnt (b,m) i32.
 *   Per query: the first nsample dataset indices k (ascending) with max(sqrtf(d2),1e-20f) < radius; unfilled
 *   slots repeat the first hit; pts_cnt = hits (saturating at nsample); rows of empty balls are zero (the
 *   reference leaves them unwritten).  radius <= 0 or nsample <= 0 -> PC_ERR_INVALID_ARGUMENT
 *   (tf_grouping.cpp:70-74). */
PC_API int pc_query_ball(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx,
                  int *pts_cnt, pc_stream_t stream);

/* Ball query through a per-scene cell grid: SAME outputs as pc_query_ball, bit for bit, but a query only meets the
 * candidates of its 3x3x3 cell neighbourhood (cell edge >= radius) instead of all n.  The TF shim / wrappers call this
 * one; it forwards to pc_query_ball for shapes the grid path does not cover (n > ~15000 points per scene).
 * workspace: pc_query_ball_grid_workspace_bytes(b,n,m) bytes, 16-byte aligned, contents scratch. */
PC_API size_t pc_query_ball_grid_workspace_bytes(int b, int n, int m);
PC_API int pc_query_ball_grid(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2,
                       int *idx, int *pts_cnt, void *workspace, pc_stream_t stream);

/* group_point.  Replaces groupPointLauncher(b,n,c,m,nsample,points,idx,out) (tf_grouping_g.cu:133-136, prototype
 * tf_grouping.cpp:142; op GroupPoint tf_grouping.cpp:41-54,143-171).
 *   points (b,n,c), idx (b,m,nsample) -> out (b,m,nsample,c);  out[b,j,k,:] = points[b,idx[b,j,k],:] */
PC_API int pc_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out,
                   pc_stream_t stream);

/* group_point gradient.  Replaces cudaMemset + groupPointGradLauncher(b,n,c,m,nsample,grad_out,idx,grad_points)
 * (tf_grouping.cpp:204-205, tf_grouping_g.cu:137-141, prototype tf_grouping.cpp:173; op GroupPointGrad).
 *   grad_out (b,m,nsample,c), idx (b,m,nsample) -> grad_points (b,n,c), fully overwritten, each row summed in
 *   ascending (j,k) order = the serial order of grouping/test/query_ball_point.cpp:70-84. */
PC_API size_t pc_group_point_grad_workspace_bytes(int b, int n, int c, int m, int nsample);
PC_API int pc_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                        float *grad_points, void *workspace, pc_stream_t stream);

/* SelectionSort.  Replaces selectionSortLauncher(b,n,m,k,dist,outi,out) (tf_grouping_g.cu:129-132, prototype
 * tf_grouping.cpp:108; op SelectionSort tf_grouping.cpp:31-40,109-139).
 *   dist (b,m,n) -> outi (b,m,n) i32, out (b,m,n) f32: the arrays the reference's k swap steps leave behind
 *   (first k columns = k smallest, swap-induced tie order).  k <= 0 -> PC_ERR_INVALID_ARGUMENT. */
PC_API int pc_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out, pc_stream_t stream);

/* knn_point, fused.  Replaces the TF graph of grouping/tf_grouping.py:48-73 (tile, subtract, square, reduce_sum,
 * SelectionSort, slice) without materialising the (b,m,n) matrix.
 *   xyz1 (b,n,c) dataset, xyz2 (b,m,c) queries -> val (b,m,k) f32 squared distances, idx (b,m,k) i32, identical
 *   to the first k columns of SelectionSort on that matrix, swap-induced tie order included.
 *   Supported: 1 <= k <= 128, k <= n, 1 <= c <= 16; otherwise PC_ERR_UNSUPPORTED. */
PC_API int pc_knn(int b, int n, int m, int k, int c, const float *xyz1, const float *xyz2, float *val, int *idx,
           pc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Interpolation library (reference: interpolation_3d/tf_interpolate.cpp -- CPU-only ops in the reference)
 * ------------------------------------------------------------------------------------------------------------- */

/* three_nn.  Replaces threenn_cpu(b,n,m,xyz1,xyz2,dist,idx) (tf_interpolate.cpp:60-103; op ThreeNN :12-21,157-187).
 *   xyz1 (b,n,3) unknown/dense, xyz2 (b,m,3) known/sparse -> dist (b,n,3) f32 SQUARED distances ascending,
 *   idx (b,n,3) i32; equal distances keep ascending index order; missing neighbours (m < 3) are (+inf, 0). */
PC_API int pc_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx,
                pc_stream_t stream);

/* three_nn through a per-scene cell grid over the known cloud xyz2: SAME outputs as pc_three_nn, bit for bit (ties by
 * ascending index), scanning cell rings around the query until the third-best distance is certified.
 * workspace: pc_three_nn_grid_workspace_bytes(b,n,m) bytes, 16-byte aligned, contents scratch. */
PC_API size_t pc_three_nn_grid_workspace_bytes(int b, int n, int m);
PC_API int pc_three_nn_grid(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx,
                     void *workspace, pc_stream_t stream);

/* Inverse-distance weights of pointnet_fp_module (utils/pointnet_util.py:219-222, stock TF ops in the reference):
 *   d = max(dist,1e-10); w_t = (1/d_t) / ((1/d_0 + 1/d_1) + 1/d_2).   dist, weight: (rows,3). */
PC_API int pc_three_weights(size_t rows, const float *dist, float *weight, pc_stream_t stream);

/* three_interpolate.  Replaces threeinterpolate_cpu(b,m,c,n,points,idx,weight,out) (tf_interpolate.cpp:107-127;
 * op ThreeInterpolate :22-36,191-222).
 *   points (b,m,c), idx (b,n,3), weight (b,n,3) -> out (b,n,c);  out = (p1*w1 + p2*w2) + p3*w3, un-fused. */
PC_API int pc_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx, const float *weight,
                         float *out, pc_stream_t stream);

/* three_interpolate gradient.  Replaces memset + threeinterpolate_grad_cpu(b,n,c,m,grad_out,idx,weight,grad_points)
 * (tf_interpolate.cpp:258-259,131-153; op ThreeInterpolateGrad :37-46,225-262).
 *   grad_out (b,n,c), idx (b,n,3), weight (b,n,3) -> grad_points (b,m,c), fully overwritten, each row summed in
 *   the reference's serial (j,t) order. */
PC_API size_t pc_three_interpolate_grad_workspace_bytes(int b, int n, int c, int m);
PC_API int pc_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx,
                              const float *weight, float *grad_points, void *workspace, pc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused layer front ends (reference: the stock-TF glue around the ops in utils/pointnet_util.py)
 * ------------------------------------------------------------------------------------------------------------- */

/* The grouping half of sample_and_group (pointnet_util.py:39-52, use_xyz=True) in one pass:
 *   new_points (b,m,nsample,3+c) = concat(xyz[idx] - new_xyz[:, :, None, :], points[idx])
 *   grouped_xyz (b,m,nsample,3)  = xyz[idx] - new_xyz[:, :, None, :]            (optional, may be NULL)
 * xyz (b,n,3), points (b,n,c) or NULL with c = 0, idx (b,m,nsample), new_xyz (b,m,3).  Bit-identical to
 * pc_group_point x2 + an fp32 subtract + a concat. */
PC_API int pc_sa_group(int b, int n, int c, int m, int nsample, const float *xyz, const float *points, const int *idx,
                const float *new_xyz, float *new_points, float *grouped_xyz, pc_stream_t stream);

/* The interpolation half of pointnet_fp_module (pointnet_util.py:219-226) in one pass:
 *   weight = inverse-distance weights of dist (as pc_three_weights);  interpolated = three_interpolate(points2, idx, weight)
 *   out (b,n,c2+c1) = concat(interpolated, points1);  weight (b,n,3) is also written when not NULL (the backward needs it)
 * dist, idx (b,n,3) from three_nn; points2 (b,m,c2); points1 (b,n,c1) or NULL with c1 = 0.  Bit-identical to the
 * op-by-op composition. */
PC_API int pc_fp_interpolate(int b, int n, int m, int c2, int c1, const float *dist, const int *idx, const float *points2,
                      const float *points1, float *out, float *weight, pc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Per-neighbourhood attention contraction (reference: attention_points/attention_scannet/attention_layer.py:35-42,
 * i.e. AttentionLayer.call after its three Dense projections)
 * ------------------------------------------------------------------------------------------------------------- */

/* G neighbourhoods of S samples, H heads of key_dim D (= output_dim), width HD = H*D:
 *   Q (G,HD), K (G,S,HD), V (G,S,HD) -> out (G,HD)
 * with the reference's RAW reshape of each neighbourhood's (S,HD) buffer to (H,S,D) (attention_layer.py:35):
 *   logit[h,s] = sum_d Q[h*D+d] * Kflat[h*S*D + s*D + d] / sqrt(D);  a = softmax_s(logit)
 *   out[h*D+d] = sum_s a[h,s] * Vflat[h*S*D + s*D + d]
 * Supported: D in {1,2,4,8,16,32,64}, 1 <= S <= 128 (else PC_ERR_UNSUPPORTED).  fp32 throughout.  S = 32, D = 4 (the
 * ScanNet models) runs a lane-per-head kernel with cp.async.bulk staging. */
PC_API int pc_attention_fwd(int G, int S, int H, int D, const float *Q, const float *K, const float *V, float *out,
                     pc_stream_t stream);

/* The whole AttentionLayer.call (attention_layer.py:29-45: Dense Q from the query row, Dense K and V from the S grouped
 * rows, raw reshape to heads of key_dim 4, softmax(QK^T/2) V) in ONE kernel on the tcgen05 tensor cores: the K | V
 * projection is a 3xTF32 split-accumulation UMMA into TMEM, the per-head softmax runs in the TMEM epilogue, K and V are
 * never written to memory.
 *   xq (G,C) query rows, x (G,S,C) grouped rows, wq/wk/wv (C,C) Dense kernels laid out [in][out], bq/bk/bv (C) or NULL
 *   -> out (G,C), within 1e-5 relative of the fp32 composition.
 * Supported in this build: S = 32, C in {64, 128, 256, 512} (the four ScanNet attention levels; for C >= 128 both
 * operands stream through a shared-memory ring, W from a pre-split image); otherwise PC_ERR_UNSUPPORTED (use a Dense GEMM +
 * pc_attention_fwd).  workspace: pc_attention_layer_workspace_bytes(G,S,C) bytes, 16-byte aligned. */
PC_API size_t pc_attention_layer_workspace_bytes(int G, int S, int C);
PC_API int pc_attention_layer_fwd(int G, int S, int C, const float *xq, const float *x, const float *wq, const float *bq,
                           const float *wk, const float *bk, const float *wv, const float *bv, float *out,
                           void *workspace, pc_stream_t stream);
/* The same layer in two steps for callers that keep the workspace across calls: _prepare builds the tensor-core operand
 * image of the weights at the start of the workspace (once per set of weights), _fwd_prepared runs the layer on it
 * without rebuilding (the weight / bias pointers are still read by the Q projection and the epilogue); xq_stride is the
 * row stride of xq in floats (0 = C; S * C with xq = x reads sample 0 of every group, attention_layer.py:259).  The
 * workspace must have been sized for the largest G used: pc_attention_layer_workspace_bytes(G_max, S, C). */
PC_API int pc_attention_layer_prepare(int S, int C, const float *wq, const float *bq, const float *wk, const float *bk,
                               const float *wv, const float *bv, void *workspace, pc_stream_t stream);
PC_API int pc_attention_layer_fwd_prepared(int G, int S, int C, const float *xq, size_t xq_stride, const float *x,
                                    const float *wq, const float *bq, const float *wk, const float *bk, const float *wv,
                                    const float *bv, float *out, void *workspace, pc_stream_t stream);

/* InnerAttentionLayer's contraction (attention_layer.py:61-75; experimental layers): attention ACROSS the 5
 * heads of one point.  Q, K, V (rows, 5 * key_dim) -> out (rows, 5 * key_dim): per row weights = softmax_j(Q_i . K_j /
 * sqrt(key_dim)) (5 x 5), out_i = sum_j w_ij V_j.  key_dim in {4, 8, 16, 32, 64}. */
PC_API int pc_inner_attention_fwd(size_t rows, int key_dim, const float *Q, const float *K, const float *V, float *out,
                                  pc_stream_t stream);

/* Gradient of pc_attention_fwd w.r.t. Q, K, V given dout (G,HD); dQ (G,HD), dK and dV (G,S,HD) fully overwritten. */
PC_API int pc_attention_bwd(int G, int S, int H, int D, const float *Q, const float *K, const float *V, const float *dout,
                     float *dQ, float *dK, float *dV, pc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Whole-scene chunker and map-back (reference: attention_points/scannet_dataset/complete_scene_loader.py:4-117,
 * attention_points/benchmark/generate_predictions.py:19-37) -- SURVEY.md 8f rank 4, BASELINE config 4
 * ------------------------------------------------------------------------------------------------------------- */

/* Membership of every point of a scan in every padded 1.5 m cell, compacted per cell in ascending point index (the
 * order of numpy's boolean mask, complete_scene_loader.py:35-36), plus the un-padded membership flag (:41).
 *   points (n,3) f32; boxes (ncells,12) f32 = padded lo[3], padded hi[3], inner lo[3], inner hi[3] (inclusive compares;
 *   the caller rounds the reference's float64 thresholds to fp32 so the compare is identical)
 *   -> cell_base (ncells+1) i32: list offsets of each cell, [ncells] = total; list (capacity 4n) i32 point indices;
 *      inner (capacity 4n) u8.   workspace: pc_scene_cells_workspace_bytes(n, ncells). */
PC_API size_t pc_scene_cells_workspace_bytes(int n, int ncells);
/* coordmin[3], coordmax[3] of a scan (complete_scene_loader.py:21-22) -> out6 (device). */
PC_API int pc_scene_bbox(int n, const float *points, float *out6, pc_stream_t stream);
PC_API int pc_scene_cells(int n, int ncells, const float *points, const float *boxes, int *cell_base, int *list,
                   unsigned char *inner, void *workspace, pc_stream_t stream);

/* Candidate chunks are described by 5 ints {list_base, order_off, start, rest, fill_off}: row t < rest is cell position
 * order[order_off+start+t] (the host-drawn np.random.shuffle order, :45-48), row t >= rest is the fill-up position
 * order[order_off + fill[fill_off+t-rest]] (np.random.choice, :87-90); the source point is list[list_base+position].
 * masksum[chunk] = number of rows t < rest inside the un-padded cell (chunks with 0 are dropped, :63,:99). */
PC_API int pc_scene_chunk_masksum(int nchunks, int npoints, const int *desc, const int *order, const unsigned char *inner,
                           int *masksum, pc_stream_t stream);
/* The kept chunks: src_index (nchunks,npoints) i32 source point of every row, point_sets (nchunks,npoints,3) f32,
 * masks (nchunks,npoints) u8 (0 on fill-up rows, :92), orig_idx (nchunks,npoints) i64 (0 on fill-up rows, :93-94). */
PC_API int pc_scene_chunk_assemble(int nchunks, int npoints, const int *desc, const int *order, const int *fill,
                            const int *list, const unsigned char *inner, const float *points, int *src_index,
                            float *point_sets, unsigned char *masks, long long *orig_idx, pc_stream_t stream);
/* out[r,:] = src[src_index[r],:] for rows of row_bytes bytes of any dtype (labels, colours, normals); rows whose index
 * is negative are zero-filled. */
PC_API int pc_gather_rows_bytes(size_t rows, int row_bytes, const void *src, const int *src_index, void *out,
                         pc_stream_t stream);
/* sample_weight (nchunks,npoints) f64 = label_weights[label] (ones(21) with [0]=0, :12-13; 1 when labels is NULL),
 * times the mask for full chunks (:70) but not for the fill-up chunk (:100-103).  labels: gathered (nchunks,npoints) i32. */
PC_API int pc_scene_sample_weights(int nchunks, int npoints, const int *desc, const int *labels, const unsigned char *masks,
                            double *out, pc_stream_t stream);
/* map_back (generate_predictions.py:19-37): winner[i] = last row r with mask[r] and orig_idx[r] == i, or -1; the
 * remapped array is then pc_gather_rows_bytes(nres, ..., values, winner, res) (zeros where nobody wrote). */
PC_API int pc_map_back_winner(size_t rows, int nres, const long long *orig_idx, const unsigned char *mask, int *winner,
                       pc_stream_t stream);

/* HOST function (no device work): the permutation np.random.shuffle(arange(n)) produces on a legacy MT19937 RandomState,
 * complete_scene_loader.py:17-18 -- the per-cell shuffle of the chunker, numpy's stream restated in C (same draws, ~5 x
 * faster than numpy's generic shuffle).  mt_key[624] / mt_pos: the state as np.random.get_state() returns it, advanced
 * in place (hand it back with np.random.set_state()).  perm: n ints. */
PC_API int pc_host_legacy_shuffle(unsigned int *mt_key, int *mt_pos, int n, int *perm);
/* HOST function: np.random.choice(high, count, replace=True) (= legacy randint(0, high, size=count)) on the same state,
 * complete_scene_loader.py:87 -- the fill-up indices of a cell's last chunk.  out: count ints in [0, high). */
PC_API int pc_host_legacy_randint(unsigned int *mt_key, int *mt_pos, int high, int count, int *out);

/* ---------------------------------------------------------------------------------------------------------------
 * Dense layers of the path on the tcgen05 tensor cores (csrc/gemm_tf32.cu), 3xTF32 split precision: fp32 in / out,
 * ~2e-6 of the output scale against a float64 product.  Reference: the 1x1 conv2d + batch norm + ReLU stack and the
 * max over nsample of pointnet_sa_module (utils/pointnet_util.py:119-135, utils/tf_util.py:120-186,512-530: stock cuDNN
 * layers there), the Dense projections of AttentionLayer (attention_layer.py:24-34) and their gradients.
 * ------------------------------------------------------------------------------------------------------------- */

/* Bytes of the tensor-core weight image of a (K, N) layer (TF32 hi / lo parts, 128-byte-swizzled K blocks). */
PC_API size_t pc_dense_image_bytes(int K, int N);
/* Builds the image once per set of weights (cache it across calls).  transpose = 0: w is (K, N) row-major, the layer
 * computes X w (Keras / tf.layers kernel layout [in][out]); transpose = 1: w is (N, K) row-major and the layer computes
 * X w^T -- the input-gradient product dX = dY W^T of a layer whose kernel W is (N_in = N here, ...) see INTEGRATION.md.
 * N <= 1024 (any K; widths that are not multiples of 16 / 32 are zero-padded inside the image). */
PC_API int pc_dense_prepare(int K, int N, const float *w, int transpose, void *image, pc_stream_t stream);
/* y (rows, N; row stride ldy) = act(x (rows, K; row stride ldx) . W + bias), act = ReLU when relu != 0, bias (N) or NULL.
 * A batch norm in inference mode folds into W and bias on the caller's side. */
PC_API int pc_dense_fwd(size_t rows, int K, int N, const float *x, size_t ldx, const void *image, const float *bias,
                        int relu, float *y, size_t ldy, pc_stream_t stream);
/* The same layer followed by the maximum over each group of group_size (= 32 = nsample) consecutive rows
 * (pointnet_util.py:134-135): y_pooled (groups, N).  y_full = NULL: the (groups * 32, N) activation is never written
 * (pointnet_sa_module); y_full != NULL: it is written as well (pointnet_sa_module_attention_and_pooling needs both,
 * attention_layer.py:306-323). */
PC_API int pc_dense_pool_fwd(size_t groups, int group_size, int K, int N, const float *x, size_t ldx, const void *image,
                             const float *bias, int relu, float *y_full, size_t ldy, float *y_pooled, size_t ldp,
                             pc_stream_t stream);

/* Weight and bias gradients of the layer: dw (K, N) = x^T . dy, db (N) = column sums of dy (db may be NULL), x (rows, K; row
 * stride ldx), dy (rows, N; row stride ldy).  Same 3xTF32 tensor-core scheme; the reduction over the rows is split
 * over the chip and the partial products are added in a fixed order (deterministic, no float atomics).  The input
 * gradient dx = dy . W^T is pc_dense_fwd with an image prepared with transpose = 1. */
PC_API size_t pc_dense_bwd_weight_workspace_bytes(size_t rows, int K, int N);
PC_API int pc_dense_bwd_weight(size_t rows, int K, int N, const float *x, size_t ldx, const float *dy, size_t ldy,
                               float *dw, float *db, void *workspace, pc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Host-boundary packing (csrc/io_pack.cu)
 * ------------------------------------------------------------------------------------------------------------- */

/* The feature prologue of the reference's input pipeline (attention_points/train.py:95-98):
 * feat (rows,6) f32 = concat(float(colors (rows,3) u8) / 255, normals (rows,3) f32).  IEEE fp32 division: the same
 * bits as tf.div / numpy.  Lets a host ship colours as the bytes they are stored as. */
PC_API int pc_unpack_features(size_t rows, const unsigned char *colors, const float *normals, float *feat,
                       pc_stream_t stream);
/* dst[i] = (uint16) src[i] for index tensors whose values are < 65536 (indices into clouds of at most 65536 points,
 * pts_cnt <= nsample); values outside [0, 65535] saturate.  Halves the device-to-host bytes of integer results. */
PC_API int pc_narrow_indices_u16(size_t count, const int *src, unsigned short *dst, pc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PCOPS_H_ */
