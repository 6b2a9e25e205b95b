/*
 * acn_b200.h -- C ABI of libacn_b200.so: the per-ray volumetric rendering hot path of
 * psklavos1/adaptive-city-nerf as hand-written sm_100a kernels.
 *
 * Conventions
 *   - every entry point returns ACN_OK (0) or a negative ACN_E* code; the message for the
 *     calling thread is in acn_last_error().  No exception crosses the ABI, nothing exits.
 *   - all tensor arguments are DEVICE pointers to contiguous row-major arrays unless the
 *     parameter is documented as "stride"; the caller owns every buffer (Python: torch
 *     allocator).  The library keeps no state besides the opaque acn_ctx.
 *   - calls are asynchronous on `stream` (a cudaStream_t); there are no hidden syncs.
 *   - "reference" citations are file:line in psklavos1/adaptive-city-nerf.
 */
#ifndef ACN_B200_H
#define ACN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define ACN_VERSION 200 /* major*100 + minor */

typedef struct acn_ctx acn_ctx;
typedef void* acn_stream; /* cudaStream_t */

enum {
    ACN_OK = 0,
    ACN_EINVAL = -1,       /* bad argument (null pointer, negative size, ...) */
    ACN_EUNSUPPORTED = -2, /* shape / dtype outside what the kernels are built for */
    ACN_ECUDA = -3,        /* CUDA runtime / launch error */
    ACN_ENODEVICE = -4     /* no sm_100 device */
};

enum { ACN_INTERP_NEAREST = 0, ACN_INTERP_LINEAR = 1, ACN_INTERP_SMOOTHSTEP = 2 };
enum { ACN_F32 = 0, ACN_F16 = 1 };

/* ---- context ----------------------------------------------------------------------------- */
int acn_version(void);
const char* acn_last_error(void);
/* One context per device; records SM count / L2 size and opts kernels into large shared memory. */
int acn_create(int device, acn_ctx** out);
int acn_destroy(acn_ctx* ctx);
int acn_device_info(acn_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes);

/* ---- stage 1: camera rays, scene-box clipping, stratified sampling ------------------------- */

/* nerfs/ray_sampling.py:111-136 get_ray_directions -> dirs (H,W,3) */
int acn_ray_directions(acn_ctx*, int H, int W, float fx, float fy, float cx, float cy,
                       int center_pixels, float* dirs, acn_stream);

/* nerfs/scene_box.py:45-107 SceneBox.ray_aabb_intersect.  o,d are (N,*) with the given row
 * strides (in floats); aabb6 = [min xyz, max xyz] on the device. */
int acn_aabb_intersect(acn_ctx*, const float* o, const float* d, int64_t N, int stride_o,
                       int stride_d, const float* aabb6, float eps, float max_bound,
                       float invalid, float* tmin, float* tmax, acn_stream);

/* nerfs/ray_sampling.py:10-24,50-108 _rays_cam_to_world + get_rays + pack_rays -> rays (N,8)
 * [o, d, near, far].  c2w: device pointer to a row-major (3|4, 4) matrix.  aabb6_or_null == NULL
 * selects the constant near/far branch. */
int acn_get_rays(acn_ctx*, const float* dirs_cam, int64_t N, const float* c2w,
                 const float* aabb6_or_null, float near_c, float far_c, float max_bound,
                 float invalid, float* rays8, acn_stream);

/* nerfs/ray_sampling.py:139-176 clamp_rays_near_far (in place).  has_override = 0 is the
 * `near_far_override is None` branch; NaN means "no override" for n / f. */
int acn_clamp_near_far(acn_ctx*, float* rays8, int64_t N, int has_override, float n_or_nan,
                       float f_or_nan, float eps, float invalid, uint8_t* valid, acn_stream);

/* nerfs/ray_rendering.py:262-287 stratified_t_vals.  u_lin = torch.linspace(0,1,S) (device);
 * jitter_or_null = the rand tensor (N,S) in training mode.  Bins are bit-exact w.r.t. the
 * reference CPU path. */
int acn_sample_stratified(acn_ctx*, const float* rays8, int64_t N, int S, const float* u_lin,
                          const float* jitter_or_null, float* t_vals, acn_stream);

/* nerfs/ray_rendering.py:317-319 -> id6 (N*S,6) = [o + d*t, d] */
int acn_points(acn_ctx*, const float* rays8, int64_t N, int S, const float* t_vals, float* id6,
               acn_stream);

/* Performance hint, not part of the reference: *flag (device int32) = 1 when consecutive rays look like adjacent pixels
 * of a frame (median gap between rays r, r+1 at mid depth < threshold x sample spacing, over the first 65 rays). */
int acn_rays_coherent(acn_ctx*, const float* rays8, int64_t N, int S, float threshold, int32_t* flag, acn_stream);

/* ---- stage 2: multiresolution hash grid (models/encodings.py:160-381, torch branch) --------- */

/* ROW RANGES (the routed container path: an expert's bucket, whose size only the device knows -- nothing is read back
 * to the host, SURVEY 8b "dispatch counts stay on device").  Entry points that take `range_or_null` process, when it is
 * given (device, 2 int32 [first, end)), only those rows of their row-indexed arrays -- all pointers stay the bases of
 * the FULL arrays -- and P is then an UPPER BOUND on end - first, used to size the launch.  Per-point directions only
 * (dirs_group == 1).
 *
 * x: (P,>=3) with row stride x_stride floats.  box6_or_null = [min xyz, extent xyz] (device):
 * when given, x is in world coordinates and models/inr/meta_ngp.py:155-158 _world_to_unit is
 * applied first.  table: (L*2^log2T, F) fp32.  res: (L) int32 on the device.
 * out: (P, L*F) fp32 or fp16.  idx_out_or_null: (P,L,8) int32 table rows (parity checks). */
int acn_hashgrid_fwd(acn_ctx*, const float* x, int64_t P, int x_stride, const int32_t* range_or_null,
                     const float* box6_or_null, const float* table, int L, int F, int log2T, const int32_t* res,
                     int interp, void* out, int out_dtype, int32_t* idx_out_or_null, acn_stream);

/* Autograd of the above w.r.t. the table: dtable (L*2^log2T, F) fp32 is ACCUMULATED into. */
int acn_hashgrid_bwd(acn_ctx*, const float* x, int64_t P, int x_stride, const int32_t* range_or_null,
                     const float* box6_or_null, int L, int F, int log2T, const int32_t* res, int interp,
                     const void* dout, int dout_dtype, float* dtable, acn_stream);

/* Same, with the points formed on the fly from rays (N,8) and t_vals (N,S): p = o + d*t
 * (nerfs/ray_rendering.py:317); the reference's (N*S,6) id6 tensor is never materialised.
 * out / dout are (N*S, L*F).  ray_major != 0 (forward): a warp encodes one sample of 32 consecutive rays instead of 32
 * consecutive samples of one ray -- same output, fewer distinct cache lines per gather when consecutive rays are
 * adjacent pixels of a frame.  ray_major_dev_or_null (device, 1 int32, e.g. from acn_rays_coherent) overrides
 * ray_major when given, so the choice needs no host read. */
int acn_hashgrid_fwd_rays(acn_ctx*, const float* rays8, const float* t_vals, int64_t N, int S,
                          const float* box6_or_null, const float* table, int L, int F, int log2T,
                          const int32_t* res, int interp, void* out, int out_dtype, int ray_major,
                          const int32_t* ray_major_dev_or_null, acn_stream);
int acn_hashgrid_bwd_rays(acn_ctx*, const float* rays8, const float* t_vals, int64_t N, int S,
                          const float* box6_or_null, int L, int F, int log2T, const int32_t* res,
                          int interp, const void* dout, int dout_dtype, float* dtable, acn_stream);

/* ---- stage 3: field MLPs (models/inr/meta_ngp.py:171-241, metamodule.py:129-192) ------------ */

/* Device pointers, fp32, nn.Linear layout (out,in), in this order:
 *  0 sigma_trunk.0.linear.weight (H,E)   1 .bias (H)     2 sigma_trunk.1.linear.weight (H,H)  3 .bias
 *  4 sigma_head.weight (1,H)  5 .bias    6 geo_head.weight (G,H)  7 .bias
 *  8 color_mlp.0.linear.weight (C,G+16)  9 .bias         10 color_mlp.1.linear.weight (C,C)  11 .bias
 * 12 color_mlp.2.weight (3,C) 13 .bias
 * These may be the module's own parameters or `params=` fast weights. */
typedef struct { const float* p[14]; } acn_field_weights;
typedef struct { float* p[14]; } acn_field_grads; /* accumulated into; entries may be NULL */

/* models/encodings.py:27-81,133-151 SH degree 3 (16 comps) after the expert's double normalise */
int acn_sh16(acn_ctx*, const float* dirs, int64_t P, int stride, float* out, acn_stream);

/* enc (P,E) fp32|fp16.  Point p reads its direction at dirs[(p / dirs_group) * dirs_stride]:
 * (P,6) id6 rows -> dirs = id6 + 3, stride 6, group 1; rays (N,8) -> dirs = rays + 3, stride 8,
 * group S.  precision ACN_F32: SIMT fp32 math.  ACN_F16: tcgen05 tensor cores, fp16
 * operands / fp32 accumulate (the reference's autocast path).  -> rgb_sigma (P,4) fp32. */
int acn_field_fwd(acn_ctx*, const void* enc, int enc_dtype, const float* dirs, int dirs_stride,
                  int dirs_group, int64_t P, const int32_t* range_or_null, int E, int H, int G, int C,
                  const acn_field_weights* w, int precision, float* rgb_sigma, acn_stream);

/* d_rgb_sigma (P,4) -> weight grads (accumulated) and d_enc (P,E) (NULL to skip). */
int acn_field_bwd(acn_ctx*, const void* enc, int enc_dtype, const float* dirs, int dirs_stride,
                  int dirs_group, int64_t P, const int32_t* range_or_null, int E, int H, int G, int C,
                  const acn_field_weights* w, int precision, const float* d_rgb_sigma,
                  const acn_field_grads* g, void* d_enc_or_null, int d_enc_dtype, acn_stream);

/* Fused per-expert forward for the `active_module` / routed-bucket render path (SURVEY 8b acn_render_expert_fwd; replaces
 * nerfs/ray_rendering.py:317-325 -> models/inr/meta_ngp.py:226-241 -> models/encodings.py:293-381 under autocast):
 * world->unit, hash-grid encode and the tcgen05 field MLPs in ONE persistent, warp-specialised kernel -- producer warps
 * write each point's fp16 encoding row into shared memory as the A operand of the first trunk layer; the encoding is
 * never read back from HBM.  Positions / directions / row range as in acn_render_expert_bwd; ray_major as in
 * acn_hashgrid_fwd_rays.  enc_f16_out_or_null (P, L*F): when given, the encoding is also written for the backward
 * (bit-identical to acn_hashgrid_fwd's fp16 rows).  res_host_or_null: the same (L) resolutions in HOST memory; when
 * given (and box6 is), the coarsest levels that fit are staged as dense lattices in shared memory.  F = 2, L in {8,16},
 * Linear / Smoothstep, H = C = 64.  -> rgb_sigma (P,4) fp32, equal to acn_hashgrid_fwd(f16) + acn_field_fwd(ACN_F16). */
int acn_render_expert_fwd(acn_ctx*, const float* x_or_null, int x_stride, const float* rays8_or_null,
                          const float* t_vals_or_null, int64_t P, int S, int ray_major,
                          const int32_t* ray_major_dev_or_null, const int32_t* range_or_null,
                          const float* box6_or_null, const float* table, int L, int F, int log2T,
                          const int32_t* res, const int32_t* res_host_or_null, int interp, const float* dirs,
                          int dirs_stride, int dirs_group, int H, int G, int C, const acn_field_weights* w,
                          void* enc_f16_out_or_null, float* rgb_sigma, acn_stream);

/* Fused per-expert backward for the `active_module` / routed-bucket render path (SURVEY 8b acn_render_expert_bwd;
 * replaces autograd through nerfs/ray_rendering.py:317-325 -> models/inr/meta_ngp.py:226-241 ->
 * models/encodings.py:331-381): the tcgen05 MLP backward of acn_field_bwd(ACN_F16) and the hash-table gradient scatter
 * of acn_hashgrid_bwd in ONE warp-specialised kernel (csrc/expert_bwd.cu: chain warps run the MLP backward, scatter warps take
 * each tile's encoding gradient out of a TMEM window and issue the REDs) -- the (P, L*F) gradient of the encoding never exists in HBM.
 * Positions: x_or_null (P,>=3) rows of stride x_stride, or rays8 (N,8) + t_vals (N,S) with P = N*S (p = o + d*t).
 * enc_f16 (P, L*F): the fp16 encoding the forward saved.  F = 2, L in {8,16}, Linear / Smoothstep.  Weight gradients
 * are accumulated into g, the table gradient into dtable (L*2^log2T, 2) fp32. */
int acn_render_expert_bwd(acn_ctx*, const float* x_or_null, int x_stride, const float* rays8_or_null,
                          const float* t_vals_or_null, int64_t P, int S, const int32_t* range_or_null,
                          const float* box6_or_null, int L, int F,
                          int log2T, const int32_t* res, int interp, const void* enc_f16, const float* dirs,
                          int dirs_stride, int dirs_group, int H, int G, int C, const acn_field_weights* w,
                          const float* d_rgb_sigma, const acn_field_grads* g, float* dtable, acn_stream);

/* Background head of the container (models/inr/meta_container.py:79-93 bg_dir_enc + bg_mlp, :347-382
 * background_color): per ray direction F.normalize -> SH16 -> Linear(16,hidden) -> ReLU -> Linear(hidden,3) -> Sigmoid,
 * one kernel forward, one backward.  dirs (N,>=3) rows of `stride` floats; w1 (hidden,16), b1 (hidden), w2 (3,hidden),
 * b2 (3) fp32 nn.Linear layout; hidden <= 64.  rgb (N,3) fp32 or fp16.  The backward takes d_rgb (N,3) fp32 and
 * ACCUMULATES the four weight gradients (each may be NULL); directions carry no gradient (rays are built under no_grad). */
int acn_background_fwd(acn_ctx*, const float* dirs, int64_t N, int stride, const float* w1, const float* b1,
                       const float* w2, const float* b2, int hidden, void* rgb, int rgb_dtype, acn_stream);
int acn_background_bwd(acn_ctx*, const float* dirs, int64_t N, int stride, const float* w1, const float* b1,
                       const float* w2, const float* b2, int hidden, const float* d_rgb, float* g_w1, float* g_b1,
                       float* g_w2, float* g_b2, acn_stream);

/* ---- stage 4: alpha compositing (nerfs/ray_rendering.py:114-165 volume_render) -------------- */
int acn_composite_fwd(acn_ctx*, const float* rgb_sigma, const float* t_vals,
                      const float* bg_or_null, int64_t N, int S, float sigma_scale, float* rgb,
                      float* depth, float* weights, float* acc, acn_stream);

/* Incoming grads may each be NULL (= zero).  d_bg_or_null (N,3) is written when bg is given. */
int acn_composite_bwd(acn_ctx*, const float* rgb_sigma, const float* t_vals,
                      const float* bg_or_null, int64_t N, int S, float sigma_scale,
                      const float* g_rgb, const float* g_depth, const float* g_weights,
                      const float* g_acc, float* d_rgb_sigma, float* d_bg_or_null, acn_stream);

/* ---- stage 5: Voronoi routing --------------------------------------------------------------- */

/* models/inr/meta_container.py:97-134 _routing.  dims = 2 (cluster_2d: columns y,z) or 3.
 * margin > 1: weights (P,K) fp32; else hard (P) int32.  counts_or_null (K) int32 is ADDED to:
 * number of points with w_k > 0 (or assigned to k). */
int acn_route_points(acn_ctx*, const float* pts, int64_t P, int stride, const float* centroids,
                     int K, int dims, float margin, float* weights_or_null,
                     int32_t* hard_or_null, int32_t* counts_or_null, acn_stream);

/* scripts/create_clusters.py:559-634 compute_voronoi_orig: mask (N,K) uint8.
 * mins_or_null / maxs_or_null (K,3) fp32 and counts_or_null (K) int64: the per-expert sample boxes streamed by
 * compute_voronoi_opt (:386-556 update_aabbs, mins_out / maxs_out / counts_out) -- UPDATED, not overwritten (start them
 * at +inf / -inf / 0 and call once per image): every sample o + d*t with D_c / (min_c' D + 1e-8) <= margin extends
 * expert c's box and counts once; non-finite samples contribute nothing.  K <= 16 with the boxes. */
int acn_route_rays_voronoi(acn_ctx*, const float* rays8, int64_t N, int S, const float* u_lin,
                           const float* centroids, int K, int dims, float margin, uint8_t* mask,
                           float* mins_or_null, float* maxs_or_null, int64_t* counts_or_null, acn_stream);

/* Device-side dispatch for meta_container.py:306-337 (replaces nonzero/index_select and its
 * K host syncs).  For expert k, the points with weight > 0 (or hard == k) are written, in
 * point order, to sel[offsets[k] .. offsets[k]+counts[k]) where offsets = exclusive scan of
 * counts.  `cursor` (K) int32 must be zero on entry.  xd_out (sum counts, 6) receives the
 * gathered [xyz,dir] rows and w_out (sum counts) the blend weight (1 for hard routing). */
int acn_bucket_points(acn_ctx*, const float* id6, int64_t P, const float* weights_or_null,
                      const int32_t* hard_or_null, int K, const int32_t* offsets, int32_t* cursor,
                      int32_t* sel, float* xd_out, float* w_out, acn_stream);

/* Expert sharding, fused dispatch: acn_bucket_points whose [xyz,dir] rows go straight into the owning GPU's receive
 * buffer.  row_base (K) holds, per expert, that buffer's DEVICE ADDRESS as mapped into this process (peer memory over
 * NVLink, e.g. torch symmetric memory; local memory for experts this rank owns) and row_off (K) the first row reserved
 * there for (this source rank, expert k); rows are 6 floats, 8-byte aligned.  sel / w_out are local, as above.  The
 * caller orders the stores against the owner's reads with a cross-GPU barrier on the same stream.  There is no
 * reference counterpart (the reference is single-GPU, meta_container.py:306-337). */
int acn_dispatch_points(acn_ctx*, const float* id6, int64_t P, const float* weights_or_null,
                        const int32_t* hard_or_null, int K, const int32_t* offsets, int32_t* cursor,
                        int32_t* sel, float* w_out, const uint64_t* row_base, const int32_t* row_off, acn_stream);

/* out[sel[i]] += y[i] * w[i]  (index_add_, meta_container.py:321) over M routed rows of 4.  y may be peer memory.
 * range_or_null: rows [range[0], range[1]) of (w, sel) only (see "ROW RANGES").  y_row0_or_null (device, 1 int32): y is
 * then read at row *y_row0 + (i - range[0]) -- the rows an expert's OWNER holds for this rank start elsewhere in its
 * buffer than in the local bucket. */
int acn_blend_add(acn_ctx*, const float* y, const float* w, const int32_t* sel, int64_t M,
                  const int32_t* range_or_null, const int32_t* y_row0_or_null, float* out, acn_stream);
/* d_y[i] = d_out[sel[i]] * w[i]  (same addressing of d_y as of y above) */
int acn_blend_bwd(acn_ctx*, const float* d_out, const float* w, const int32_t* sel, int64_t M,
                  const int32_t* range_or_null, const int32_t* y_row0_or_null, float* d_y, acn_stream);

/* The container's render path without the point and weight matrices: routing (meta_container.py:97-134) of the samples
 * o + d*t of packed rays (nerfs/ray_rendering.py:317-319) and their bucketing (:306-337), straight from (rays8, t_vals).
 *   acn_route_count_rays : counts (K) int32 += rows per expert (the one host read that sizes the buckets)
 *   acn_route_bucket_rays: sel (total) = sample index r*S+s, w_out (total) = blend weight (1 when margin == 1),
 *                          xd_out (total,6) = [xyz, dir] rows, expert k's in [offsets[k], offsets[k]+counts[k]);
 *                          cursor (K) int32 must be zero on entry.
 * support (N*S) uint16: bit k set = expert k is in the sample's support set; written by the count pass and, when handed to
 * the bucket pass, saves it the routing (distances are then evaluated only for the experts in the set).
 * ray_major != 0 orders the rows of a bucket as (sample, 32 adjacent rays) instead of (ray, 32 consecutive samples): for
 * frames, where consecutive rays are adjacent pixels, a warp of the experts' gather kernels then works on neighbouring
 * cells.  The set of rows per expert does not depend on it.  ray_major_dev_or_null (device, 1 int32) overrides it.
 * row_base / row_off (K each, as in acn_dispatch_points): when given, expert k's [xyz, dir] rows are stored into that
 * buffer (peer memory of the GPU that owns k) instead of xd_out.  row_limit (K): at most that many rows per expert are
 * written (see acn_bucket_plan).
 * Same arithmetic as acn_points + acn_route_points + acn_bucket_points (rows and weights are bit-identical); K <= 16. */
int acn_route_count_rays(acn_ctx*, const float* rays8, const float* t_vals, int64_t N, int S,
                         const float* centroids, int K, int dims, float margin, int ray_major,
                         const int32_t* ray_major_dev_or_null, uint16_t* support_or_null, int32_t* counts, acn_stream);
int acn_route_bucket_rays(acn_ctx*, const float* rays8, const float* t_vals, int64_t N, int S,
                          const float* centroids, int K, int dims, float margin, int ray_major,
                          const int32_t* ray_major_dev_or_null, const uint16_t* support_or_null,
                          const int32_t* offsets, int32_t* cursor, int32_t* sel, float* xd_out, float* w_out,
                          const uint64_t* row_base_or_null, const int32_t* row_off_or_null,
                          const int32_t* row_limit_or_null, acn_stream);

/* The bucket layout from DEVICE-side counts, nothing read back (SURVEY 8b "dispatch counts stay on device"; the reference
 * syncs K times per container forward, meta_container.py:309-313): seg (K+1) int32 = exclusive scan of counts (K),
 * clamped so that the total stays within cap rows (the size the caller allocated the bucket arrays with); limit (K) =
 * rows of each expert that fit (hand it to acn_route_bucket_rays as row_limit: rows beyond are dropped); cursor (K) is
 * zeroed for the bucket pass; *overflow_or_null is SET to 1 when rows were cut (never cleared: the caller polls it). */
int acn_bucket_plan(acn_ctx*, const int32_t* counts, int K, int64_t cap, int32_t* seg, int32_t* limit,
                    int32_t* cursor, int32_t* overflow_or_null, acn_stream);

/* The same for EXPERT SHARDING (expert k lives on rank k % world; no reference counterpart, the reference is single-GPU):
 * from the all-gathered counts of every rank, all_counts (world, K) int32 on the device, every rank derives identically
 * where each (source rank, expert) segment lies in the receive buffer of the expert's owner -- buffer order: local
 * expert, then source rank; every buffer holds at most cap_peer rows -- and takes its part: seg_local (K+1) its local
 * bucket layout (at most cap_local rows), limit (K) the rows of expert k it may write, row_off (K) where they start in
 * the owner's buffer, seg_recv (K/world + 1) the row ranges of its OWN experts in its buffer; cursor (K) is zeroed;
 * *overflow_or_null is set to 1 when anything was cut.  With these, acn_route_bucket_rays stores rows straight into peer
 * memory and the owners' kernels take device-side row ranges: a routed step needs no host read. */
int acn_shard_plan(acn_ctx*, const int32_t* all_counts, int world, int K, int rank, int64_t cap_local, int64_t cap_peer,
                   int32_t* seg_local, int32_t* limit, int32_t* row_off, int32_t* seg_recv, int32_t* cursor,
                   int32_t* overflow_or_null, acn_stream);

/* ---- around the render: loss epilogue and optimizer tail --------------- */
enum { ACN_COLOR_LINEAR = 0, ACN_COLOR_SRGB = 1, ACN_COLOR_IDENTITY = 2 };
#define ACN_LOSS_PARTIALS 1024   /* doubles of workspace acn_color_mse may use */
#define ACN_ADAM_MAX_TENSORS 192 /* tensors per acn_grad_sqnorm / acn_adam_prepare / acn_adam_apply call */

/* nerfs/color_space.py:22-66 color_space_transformer + nerfs/losses.py:32 F.mse_loss in one pass over n = 3N
 * elements: pred is the rendered LINEAR rgb, gt the sRGB ground truth.  loss_or_null (1): mean (mean != 0) or sum of
 * the squared errors; elem_or_null (n): the squared errors (reduction="none"); dpred_or_null (n): d loss / d pred
 * (already divided by n when mean != 0).  partial: ACN_LOSS_PARTIALS doubles of workspace, summed in a fixed order.
 * In sRGB mode the gradient at pred == 0 is the linear branch's (the reference's is NaN there). */
int acn_color_mse(acn_ctx*, const float* pred, const float* gt, int64_t n, int color_space, int mean,
                  float* loss_or_null, float* elem_or_null, float* dpred_or_null, double* partial, acn_stream);

/* One parameter tensor of an optimizer step (host struct, copied into the kernel arguments). */
typedef struct {
    float* p;            /* parameter */
    float* g;            /* gradient, possibly still multiplied by the GradScaler scale */
    float* m;            /* exp_avg */
    float* v;            /* exp_avg_sq */
    int64_t n;
    double lr;           /* the param group's lr */
    double weight_decay; /* the param group's weight decay */
    float* step;         /* device, 1 float: THIS tensor's step count, as torch.optim.Adam keeps it per parameter (a tensor
                          * that had no gradient on some steps lags behind the others); NULL = use the global count */
    double* bias;        /* device, 4 doubles of workspace for this tensor (zero before the first step):
                          * [bias_correction1, sqrt(bias_correction2), gradient-has-a-non-zero mark, takes-part-in-this-step],
                          * written by acn_grad_sqnorm / acn_adam_prepare / acn_adam_advance, read by acn_adam_apply;
                          * NULL with step == NULL */
} acn_adam_tensor;

/* Optimizer tail of pipelines/offline_stage/meta_core.py:123-141 maml_meta_update (scaler.unscale_ ->
 * clip_all_grads :181-190 -> scaler.step) and pipelines/online_stage/runtime_adapt.py:262-268, for
 * common/utils.py:16-76 get_optimizer's Adam / AdamW, without reading anything back to the host:
 *   acn_grad_sqnorm : acc2[0] += sum (g / scale)^2, acc2[1] += [any element non-finite]; call once per <= 48 tensors.
 *   acn_adam_prepare: one thread decides the step.  state8 (doubles, persistent; zero before the first step) =
 *                     [step, coef = clip / scale, skip, bias_correction1, sqrt(bias_correction2), total_norm, -, -];
 *                     skip when any gradient (or *found_inf_or_null) is non-finite, as GradScaler.step does;
 *                     clip = min(1, max_norm / (total_norm + 1e-6)) when max_norm > 0 (torch clip_grad_norm_);
 *                     acc2 is cleared; found_inf_out_or_null (1 float) is for GradScaler.update().
 *                     tensors_or_null / count: the tensors taking part in THIS step (those with a gradient): unless
 *                     the step is skipped, each one's *step is incremented and its bias corrections are written to
 *                     *bias from ITS step (torch.optim.Adam: a parameter whose grad is None keeps its step).
 *                     skip_zero_grads != 0: a tensor whose gradient is identically zero is left alone like one whose
 *                     grad is None -- the sync-free routed path returns zeros, not None, for an expert without rows.
 *   acn_adam_advance: the per-tensor part of acn_adam_prepare for a further batch of tensors (more than
 *                     ACN_ADAM_MAX_TENSORS in one step); call after acn_adam_prepare.
 *   acn_adam_apply  : torch.optim.Adam (adamw = 0) / AdamW (adamw = 1) update of <= 48 tensors with the gradient
 *                     multiplied by coef on the fly; write_grads != 0 also stores the unscaled, clipped gradient
 *                     back (what the reference leaves in .grad). */
int acn_grad_sqnorm(acn_ctx*, const acn_adam_tensor* tensors, int count, const float* grad_scale_or_null,
                    double* acc2, acn_stream);
int acn_adam_prepare(acn_ctx*, double* acc2, const float* grad_scale_or_null, const float* found_inf_or_null,
                     float max_norm, double beta1, double beta2, double* state8, float* found_inf_out_or_null,
                     const acn_adam_tensor* tensors_or_null, int count, int skip_zero_grads, acn_stream);
int acn_adam_advance(acn_ctx*, const acn_adam_tensor* tensors, int count, const double* state8, double beta1,
                     double beta2, int skip_zero_grads, acn_stream);
int acn_adam_apply(acn_ctx*, const acn_adam_tensor* tensors, int count, const double* state8, double beta1,
                   double beta2, double eps, int adamw, int write_grads, acn_stream);

/* ---- ray-batch producer ------------------------------------------------------ */
/* data/task_dataset.py:544-627 TaskDataset._route_and_bin with routing_policy="dda" (nerf_runner.py:207), per ray:
 * clip to the region box and (near, far) (:155-172), walk the nx*ny*nz task grid for at most max_steps voxels keeping the
 * cell with the longest in-cell parametric length (:239-351), drop the ray when its overlap with that cell is below
 * tol[cell] (:590-603).  aabb6 = region [lo, hi]; cell3 = clamp((hi-lo)/cells, 1e-12) (3 floats), cell_bounds (C,2,3) and
 * tol (C) are the small tensors the reference builds on the host (:174-197, :241-245, :595-597), all on the device.
 * cid_out (N) int32 = cell or -1; best_len_or_null (N); counts_or_null (C) int32 is ADDED to.  Bins: feed cid_out to
 * acn_bucket_points as `hard`. */
int acn_dda_route_rays(acn_ctx*, const float* rays8, int64_t N, const float* aabb6, int nx, int ny, int nz,
                       const float* cell3, const float* cell_bounds, const float* tol, int max_steps,
                       int32_t* cid_out, float* best_len_or_null, int32_t* counts_or_null, acn_stream);

/* Plain per-(point, level) table scatter (one thread per point and level, 8 corner REDs): what acn_hashgrid_bwd runs for
 * F != 2 and "Nearest"; exported so tests can cross-check the run-length kernels against it on the same inputs. */
int acn_hashgrid_bwd_plain(acn_ctx*, const float* x, int64_t P, int x_stride, const float* box6_or_null,
                           int L, int F, int log2T, const int32_t* res, int interp, const void* dout,
                           int dout_dtype, float* dtable, acn_stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* ACN_B200_H */
