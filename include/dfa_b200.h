/*
 * dfa_b200.h — C ABI of libdfa_b200.so: SimPB's deformable feature aggregation hot path,
 * hand-written CUDA for sm_100a (B200).
 *
 * This is the drop-in boundary.  Every entry point takes plain pointers and sizes (device
 * pointers unless the name ends in `_host`), an explicit CUDA stream, and returns an int:
 *   0                      success
 *   > 0                    a cudaError_t raised by the launch / runtime
 *   DFA_ERR_* (< 0)        argument validation failed — nothing was launched
 * Nothing throws across this boundary and every entry point is re-entrant from any host thread
 * (PyTorch's autograd thread included).  The only process-wide state is read-only after first use:
 * the SM count of each device and the DFA_* tuning knobs, which are environment variables read once
 * per process (tests and tools switch kernels with them; dfa_debug_reload_knobs() makes the library
 * read them again).  No launch path calls getenv() after that.
 *
 * Reference interface each entry point replaces (paths under
 * /root/reference/projects/mmdet3d_plugin/):
 *   dfa_forward            ops/src/deformable_aggregation_cuda.cu:265-288  (deformable_aggregation)
 *   dfa_backward           ops/src/deformable_aggregation_cuda.cu:291-318  (deformable_aggregation_grad)
 *   dfa_flatten_maps       ops/__init__.py:63-92                           (feature_maps_format)
 *   dfa_keypoints_project  models/detection3d/blocks.py:181-207 + models/blocks.py:198-213
 *   dfa_keypoints_project_backward   autograd of the two above (torch ops in the reference)
 *   dfa_softmax_weights    models/blocks.py:175-195 (softmax over cams x levels x points, attn-drop
 *                          mask) + the permute of models/blocks.py:133-144
 *   dfa_softmax_weights_backward     autograd of the above
 *   dfa_softmax_weights_split(_backward)  the same for models/blocks.py:166-174's camera-embedding
 *                          branch, with the broadcast add folded into the kernel
 *   dfa_msda_forward / _backward     mmcv-full 1.7.1 MultiScaleDeformableAttnFunction as called at
 *                          models/group_attn.py:229-233 (third-party kernel, not vendored)
 *   dfa_forward_fused      models/blocks.py:118-147 in one launch (inference)
 *   dfa_forward_host       the same forward, called with HOST buffers (copies inside)
 *
 * Tensor layouts (row-major, innermost last) — ops/src/deformable_aggregation.cpp:22-28:
 *   mc_ms_feat         [bs, num_feat, C]         float32 or bfloat16 (feat_dtype)
 *   spatial_shape      [K, L, 2]  = (H, W)       int32
 *   scale_start_index  [K, L]                    int32, first row of (camera k, level l)
 *   sampling_location  [bs, A, P, K, 2] = (x, y) float32, normalised to the image
 *   weights            [bs, A, P, K, L, G]       float32
 *   output             [bs, A, C]                float32
 */
#ifndef DFA_B200_H_
#define DFA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFA_B200_VERSION 100

/* argument-validation errors (negative so they never collide with cudaError_t) */
#define DFA_ERR_NULL_POINTER (-1)
#define DFA_ERR_BAD_DIMS (-2)        /* non-positive dim, C % G != 0, int32 index overflow */
#define DFA_ERR_BAD_DTYPE (-3)
#define DFA_ERR_MISALIGNED (-4)      /* a pointer is not aligned for its element type */
#define DFA_ERR_UNSUPPORTED (-5)

/* feature element types */
#define DFA_F32 0
#define DFA_BF16 1

/* dfa_backward flags */
#define DFA_BWD_ACCUMULATE 0      /* reference contract: add into caller-zeroed gradient buffers  */
#define DFA_BWD_OVERWRITE_SMALL 1 /* grad_sampling_location / grad_weights are fully written
                                     (zeros for masked samples): no memset needed for them     */
#define DFA_BWD_ZERO_GRAD_FEAT 2  /* the library zero-fills grad_mc_ms_feat on `stream` first     */

/* Same eight sizes the reference launchers take (…_cuda.cu:272-279), in the same order. */
typedef struct dfa_dims {
  int32_t batch_size;  /* bs */
  int32_t num_cams;    /* K  */
  int32_t num_feat;    /* rows of mc_ms_feat per batch item */
  int32_t num_embeds;  /* C  */
  int32_t num_scale;   /* L  */
  int32_t num_anchors; /* A  */
  int32_t num_pts;     /* P  */
  int32_t num_groups;  /* G  */
} dfa_dims;

int dfa_version(void);
/* Tests / tools: re-read the DFA_* tuning environment variables on their next use. */
void dfa_debug_reload_knobs(void);
const char *dfa_error_string(int code);

/* out[b,a,c] = sum_{p,k,l} valid(b,a,p,k) * w[b,a,p,k,l,c/(C/G)] * bilinear(feat, loc)  —
 * `output` is written, not accumulated: it needs no zero-fill (the reference needs at::zeros,
 * deformable_aggregation.cpp:55). */
int dfa_forward(const void *mc_ms_feat, int feat_dtype, const int32_t *spatial_shape,
                const int32_t *scale_start_index, const float *sampling_location,
                const float *weights, float *output, const dfa_dims *dims, void *stream);

/* Gradients wrt features, sampling locations and weights.  grad_mc_ms_feat is float32
 * [bs,num_feat,C] whatever feat_dtype is, and is always accumulated into (scatter); it may be NULL
 * (frozen backbone): the scatter and its zero-fill are then skipped altogether. */
int dfa_backward(const void *mc_ms_feat, int feat_dtype, const int32_t *spatial_shape,
                 const int32_t *scale_start_index, const float *sampling_location,
                 const float *weights, const float *grad_output, float *grad_mc_ms_feat,
                 float *grad_sampling_location, float *grad_weights, const dfa_dims *dims,
                 int flags, void *stream);

/* The module's forward in ONE kernel (inference): key points + camera projection (as
 * dfa_keypoints_project), softmax over cams x levels x points of the attention logits (as
 * dfa_softmax_weights[_split], no keep mask) and the aggregation.  logits_anchor is [bs,A,L*P*G]
 * with logits_cam [bs,K,L*P*G] (split form), or the full [bs,A,K,L*P*G] tensor with
 * logits_cam = NULL.  sampling_location_out (may be NULL) receives the [bs,A,P,K,2] locations.
 * Replaces models/blocks.py:118-147 (everything between the Linear layers).  Returns
 * DFA_ERR_UNSUPPORTED for shapes outside the row-sliced fast path or G not a power of two <= 32:
 * the caller then runs the three separate entry points. */
int dfa_forward_fused(const void *mc_ms_feat, int feat_dtype, const int32_t *spatial_shape,
                      const int32_t *scale_start_index, const float *anchor, const float *fix_scale,
                      int num_fix, const float *learnable_logits, const float *projection_mat,
                      const float *image_wh, const float *logits_anchor, const float *logits_cam,
                      float *output, float *sampling_location_out, const dfa_dims *dims, void *stream);

/* Test side channel: the geometry the kernels use, for the bit-exact checks.
 * valid [bs,A,P,K] uint8; corner_rows [bs,A,P,K,L,4] int32 (row in [0,num_feat) or -1). */
int dfa_debug_indices(const int32_t *spatial_shape, const int32_t *scale_start_index,
                      const float *sampling_location, uint8_t *valid, int32_t *corner_rows,
                      const dfa_dims *dims, void *stream);

/* Flatten L feature maps [bs, K, C, H_l, W_l] (NCHW, float32) into the multi-camera /
 * multi-scale channel-last layout [bs, K*sum(H_l*W_l), C] in one pass; out_dtype selects
 * float32 or bfloat16 output.  level_ptrs / level_hw are HOST arrays (L pointers, L (H,W)). */
int dfa_flatten_maps(const float *const *level_ptrs, const int32_t *level_hw, int num_levels,
                     int bs, int num_cams, int channels, void *col_feats, int out_dtype,
                     void *stream);

/* Key-point generation + camera projection:
 *   anchor [bs,A,11], fix_scale [F,3], learnable_logits [bs,A,(P-F)*3] or NULL (then P == F),
 *   projection_mat [bs,K,4,4], image_wh [bs,K,2] or NULL
 *   → key_points [bs,A,P,3] (may be NULL) and sampling_location [bs,A,P,K,2]. */
int dfa_keypoints_project(const float *anchor, const float *fix_scale, int num_fix,
                          const float *learnable_logits, const float *projection_mat,
                          const float *image_wh, float *key_points, float *sampling_location,
                          int bs, int num_anchors, int num_pts, int num_cams, void *stream);

/* Gradients of dfa_keypoints_project wrt the anchor ([bs,A,11]; velocity entries get 0) and the
 * learnable-offset logits ([bs,A,(P-F)*3], may be NULL), given grad_sampling_location
 * [bs,A,P,K,2].  Both outputs are fully written (no zero-fill needed). */
int dfa_keypoints_project_backward(const float *anchor, const float *fix_scale, int num_fix,
                                   const float *learnable_logits, const float *projection_mat,
                                   const float *image_wh, const float *grad_sampling_location,
                                   float *grad_anchor, float *grad_learnable_logits, int bs,
                                   int num_anchors, int num_pts, int num_cams, void *stream);

/* Attention weights of the module: logits [bs,A,K,L,P,G] (the weights_fc output, any view of that
 * memory order) -> weights [bs,A,P,K,L,G] = softmax over the K*L*P entries of every (b,a,g),
 * times keep_mask[b,a,k,p] * scale when keep_mask (uint8 [bs,A,K,P]) is given (training-time
 * attn-drop, scale = 1/(1-p)).  Needs 256 % G == 0. */
int dfa_softmax_weights(const float *logits, const uint8_t *keep_mask, float scale, float *weights,
                        int bs, int num_anchors, int num_cams, int num_scale, int num_pts,
                        int num_groups, void *stream);
int dfa_softmax_weights_backward(const float *logits, const uint8_t *keep_mask, float scale,
                                 const float *grad_weights, float *grad_logits, int bs,
                                 int num_anchors, int num_cams, int num_scale, int num_pts,
                                 int num_groups, void *stream);

/* The same with SPLIT logits.  weights_fc is linear, so with camera embedding
 * weights_fc(feature[b,a] + cam[b,k]) = logits_anchor[b,a,:] + logits_cam[b,k,:]  with
 * logits_anchor [bs,A,L*P*G] = weights_fc(feature) (bias included) and logits_cam [bs,K,L*P*G] =
 * cam @ W^T: the module runs the GEMM on bs*(A+K) rows instead of bs*A*K and the two parts are
 * added inside the kernel.  The backward writes grad_logits_full [bs,A,K,L*P*G] (the caller sums it
 * over anchors for the camera part) and grad_logits_anchor [bs,A,L*P*G] (summed over cameras). */
int dfa_softmax_weights_split(const float *logits_anchor, const float *logits_cam,
                              const uint8_t *keep_mask, float scale, float *weights, int bs,
                              int num_anchors, int num_cams, int num_scale, int num_pts,
                              int num_groups, void *stream);
int dfa_softmax_weights_split_backward(const float *logits_anchor, const float *logits_cam,
                                       const uint8_t *keep_mask, float scale,
                                       const float *grad_weights, float *grad_logits_full,
                                       float *grad_logits_anchor, int bs, int num_anchors,
                                       int num_cams, int num_scale, int num_pts, int num_groups,
                                       void *stream);

/* Multi-scale deformable attention of the 2-D query branch — replaces mmcv-full 1.7.1's
 * MultiScaleDeformableAttnFunction (ms_deform_attn_forward / _backward) at the reference call site
 * models/group_attn.py:229-233.  value [bs, S, M, D] (float32 or bfloat16; for SimPB a view of one
 * camera's slice of the mc_ms_feat table), spatial_shapes [L,2] (H,W), level_start_index [L],
 * sampling_loc [bs,Q,M,L,P,2] (x,y), attn_weight [bs,Q,M,L,P] → output [bs,Q,M*D] float32, written
 * (no zero-fill needed).  Backward: grad_sampling_loc and grad_attn_weight are fully written;
 * grad_value [bs,S,M,D] float32 is accumulated into (zero_grad_value != 0: zero-filled on the stream
 * first; NULL: skipped).
 * Camera groups in ONE launch: with num_tables = K > 1, value is [bs, K, S, M, D] and query q samples
 * table query_table[q] (int32 [Q], device) — the reference loops over the K groups with one mmcv
 * call each.  num_tables = 1 / query_table = NULL is the plain function. */
int dfa_msda_forward(const void *value, int value_dtype, const int32_t *spatial_shapes,
                     const int32_t *level_start_index, const float *sampling_loc,
                     const float *attn_weight, float *output, int bs, int num_value, int num_heads,
                     int head_dim, int num_query, int num_levels, int num_points, int num_tables,
                     const int32_t *query_table, void *stream);
/* Inference variant on the UNPROJECTED table [bs, num_tables, num_value, channels] (no head dimension;
 * a row is 512 or 1024 bytes): gathers whole rows per (query, head) — out_gathered [bs, num_query,
 * num_heads, channels] — and the sum of attention x in-map bilinear weights out_weight_sum [bs, num_query,
 * num_heads], so that the caller applies value_proj AFTER the gather:
 *   out[b,q,m,:] = W_m . out_gathered[b,q,m,:] + bias_m * out_weight_sum[b,q,m]
 * (sampling is linear; this replaces the [bs*tables*num_value, C] x [C, C] GEMM of models/group_attn.py:
 * 172 by num_heads products of [num_query, C] x [C, head_dim]). */
int dfa_msda_forward_raw(const void *table, int table_dtype, const int32_t *spatial_shapes,
                         const int32_t *level_start_index, const float *sampling_loc,
                         const float *attn_weight, float *out_gathered, float *out_weight_sum, int bs,
                         int num_value, int channels, int num_heads, int num_query, int num_levels,
                         int num_points, int num_tables, const int32_t *query_table, void *stream);
int dfa_msda_backward(const void *value, int value_dtype, const int32_t *spatial_shapes,
                      const int32_t *level_start_index, const float *sampling_loc,
                      const float *attn_weight, const float *grad_output, float *grad_value,
                      float *grad_sampling_loc, float *grad_attn_weight, int bs, int num_value,
                      int num_heads, int head_dim, int num_query, int num_levels, int num_points,
                      int num_tables, const int32_t *query_table, int zero_grad_value, void *stream);

/* Forward with HOST buffers: the inputs go host→device, the kernel runs, the output comes back
 * device→host, all on `stream`, then a stream synchronise.  `workspace` is a device buffer of at least
 * dfa_forward_host_workspace_bytes() bytes (the library never allocates).
 *
 * How the inputs travel depends on the host buffers:
 *   - pinned AND mapped (cudaHostAlloc / cudaHostRegister(…Mapped): device-accessible): PULL mode.  The
 *     small operands are copied whole; the device then marks the feature rows the forward will read
 *     (the op's own validity test and corner geometry, so the set is exact) and reads exactly those
 *     rows — and the weight lines of the valid samples — straight from host memory.  With camera-rig
 *     inputs that is about a quarter of the table.  The output is bit-identical to the whole-copy path.
 *   - anything else (pageable, or DFA_HOST_PULL=0): whole copies of all five inputs; pinned buffers
 *     make them asynchronous.
 * dfa_forward_host_stats() reports what the LAST call on this workspace moved host→device: total
 * bytes, feature rows and weight bytes (whole-copy mode: everything). */
int64_t dfa_forward_host_workspace_bytes(int feat_dtype, const dfa_dims *dims);
int dfa_forward_host(const void *h_mc_ms_feat, int feat_dtype, const int32_t *h_spatial_shape,
                     const int32_t *h_scale_start_index, const float *h_sampling_location,
                     const float *h_weights, float *h_output, const dfa_dims *dims,
                     void *workspace, int64_t workspace_bytes, void *stream);
int dfa_forward_host_stats(const void *workspace, int feat_dtype, const dfa_dims *dims, void *stream,
                           int64_t *h2d_bytes, int64_t *rows_moved, int64_t *weight_bytes_moved);

#ifdef __cplusplus
}
#endif
#endif /* DFA_B200_H_ */
