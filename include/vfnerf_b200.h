/* vfnerf_b200 -- C ABI of the B200-native VF-NeRF render() hot path.
 *
 * The reference (albertgassol1/vf-nerf) has no FFI: its "operator interface" for this path is the
 * Python method VectorFieldNerf.render() (models/nerf/vector_field_nerf.py:216-338) and the bare
 * nn.Module call VectorFieldNetwork.__call__ (models/vector_field/vector_field_network.py:140-208,
 * used by evaluation/utils/mc_utils.py:88-104 and train/vector_field_nerf_train.py:191,203,215).
 * The entry points below are what a ctypes binding of those two call sites needs (INTEGRATION.md
 * shows the stub).  Conventions:
 *   - every pointer is a DEVICE pointer to contiguous fp32 unless stated otherwise;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default);
 *   - no hidden allocation: scratch memory is a caller-provided workspace whose size is queried first;
 *   - return value 0 = ok, non-zero = error; vfnerf_last_error() returns a thread-local message;
 *   - no exceptions cross the boundary, no torch types appear in any signature.
 */
#ifndef VFNERF_B200_H
#define VFNERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFNERF_MAX_LAYERS 16
/* render_cfg.flags.  The reference evaluates the VF MLP on the coarse points twice (once in the coarse sweep,
 * vector_field_nerf.py:252-272, once among the merged points, :289-312).  On the bf16 forward-only path both MLPs run on
 * the coarse points once and the merged pass evaluates only the new fine candidates; results are bit-identical to
 * recomputing because the merged coarse points carry the same bits and the fused chain is a pure per-point function.
 * This flag restores the literal two-evaluation schedule (parity tests compare the two). */
#define VFNERF_FLAG_RECOMPUTE_COARSE 1
/* Two paths the reference's render() has but cannot execute, offered behind explicit
 * opt-in flags with their evident intent:
 *   WHITE_BG: white background, rgb += 1 - sum_j w_j after the final composite (vector_field_nerf.py:325-329; upstream
 *             the same statement in the coarse block, :273-277, reads rgb before it is assigned and raises);
 *   NERF_WEIGHTS: rendering="nerf", w_j = a_j * prod_{i<=j}(1 - a_i + 1e-10) (utils/rendering.py:98-119, inclusive
 *             cumprod) with the arguments in the function's own order (upstream passes (z_vals, density) swapped). */
#define VFNERF_FLAG_WHITE_BG 2
#define VFNERF_FLAG_NERF_WEIGHTS 4
/* forward-only calls on the tensor-core paths: the workspace already holds the packed weight images of these arenas (an
 * earlier vfnerf_render_fwd with the same workspace, cfg and network descriptions, parameters unchanged since) -- skip
 * the re-tiling of the weights (2 launches, ~30 us).  Chunked evaluation loops render hundreds of 1024-ray chunks with
 * one set of weights (evaluation/methods.py:510-530). */
#define VFNERF_FLAG_WEIGHTS_PACKED 8
#define VFNERF_MAX_SAMPLES 256 /* samples per ray (coarse + fine) handled by one warp */

/* precision of the two MLPs */
#define VFNERF_PREC_FP32 0    /* CUDA-core fp32 (generic layer widths) -- parity path, 1e-3 contract      */
#define VFNERF_PREC_BF16 1    /* tcgen05 bf16 x bf16 -> fp32 in TMEM (256-wide nets) -- 5e-3 contract       */
#define VFNERF_PREC_BF16X3 2  /* tcgen05, hi/lo split operands, 3 MMAs per product -- fp32-class accuracy   */
#define VFNERF_PREC_FP16F8 3  /* tcgen05, fp16 product + two 8-bit remainder products (kind::f8f6f4) per VF product:
                               * 2 tensor-core units instead of 3, ~1e-3 on the reference goldens (forward only).
                               * Hidden VF activations and folded weights travel as fp16: they must stay below 65504
                               * (any BatchNorm-ed net does); VFNERF_PREC_BF16X3 keeps fp32's range              */

/* One MLP (Linear [+ BatchNorm1d in eval mode] per layer).  All parameters AND running statistics of
 * the network live in one contiguous fp32 "arena"; the offsets below are in floats.  -1 = absent.
 * Mirrors the state_dict of VectorFieldNetwork / RenderingNetwork: layers.{i}.0.{weight,bias},
 * layers.{i}.1.{weight,bias,running_mean,running_var}, last layer layers.{n-1}.{weight,bias}.
 * The gradient arena produced by the backward calls has the same layout (running-stat slots = 0). */
typedef struct vfnerf_mlp_desc {
  int32_t n_layers;
  int32_t in_dim[VFNERF_MAX_LAYERS];
  int32_t out_dim[VFNERF_MAX_LAYERS];
  int64_t w_off[VFNERF_MAX_LAYERS];     /* Linear.weight [out,in] row-major                      */
  int64_t b_off[VFNERF_MAX_LAYERS];     /* Linear.bias [out]                                     */
  int64_t gamma_off[VFNERF_MAX_LAYERS]; /* BatchNorm1d.weight, -1 on layers without BN           */
  int64_t beta_off[VFNERF_MAX_LAYERS];  /* BatchNorm1d.bias                                      */
  int64_t mean_off[VFNERF_MAX_LAYERS];  /* running_mean                                          */
  int64_t var_off[VFNERF_MAX_LAYERS];   /* running_var                                           */
  int64_t arena_floats;
} vfnerf_mlp_desc;

/* Everything render() reads from the reference's config objects at call time
 * (config_parser/vf_nerf_config.py:62-124, models/samplers/ray_sampler.py near/far/N_samples). */
typedef struct vfnerf_render_cfg {
  int32_t n_rays;
  int32_t n_coarse;         /* ray_sampler.N_samples                                         */
  int32_t n_fine;           /* min(fine_sampler.N_samples, fine_sampler.max_samples)         */
  int32_t perturb;          /* 0: deterministic (eval)  1: stratified with U1/U2             */
  int32_t pose_is_quat;     /* 0: pose [R,4,4]   1: pose [R,7] = (qr,qi,qj,qk,tx,ty,tz)      */
  int32_t window;           /* len(cos_sim_weights), 11 in the shipped config                */
  int32_t normalize;        /* normalize_rendering                                           */
  int32_t multires;         /* VF embedder_multires (6)                                      */
  int32_t multires_view;    /* rendering-net embedder_multires (4)                           */
  int32_t skip_layer;       /* VF skip_connection_in[0], -1 = none                           */
  int32_t precision;        /* VFNERF_PREC_*                                                 */
  int32_t flags;            /* VFNERF_FLAG_*                                                 */
  double near_, far_, fine_range;   /* python floats of the samplers (kept in double: the reference
                                       forms far-near and 2*range/(Nf-1) in double before rounding) */
  float dir_to_normal_th;
  float beta_lo, beta_hi, mean_lo, mean_hi, scale_min;   /* LaplaceDensity bounds                */
  float bn_eps;
  double fine_near_, fine_far_;     /* fine_sampler.near / .far: the range of the fallback samples z_add
                                       (ray_sampler.py:296-299); the reference's callers set them equal to near/far */
} vfnerf_render_cfg;

/* Outputs of render() == fields of NerfOutput (models/nerf/output.py:7-22) as filled at
 * vector_field_nerf.py:331-338, plus optional intermediates (may be NULL). */
typedef struct vfnerf_render_out {
  float* points;        /* [R,N,3]  NerfOutput.points_coarse (merged coarse+fine points)          */
  float* normals;       /* [R,N,3]  NerfOutput.coarse_normals (raw tanh VF vectors)               */
  float* rgb;           /* [R,3]    NerfOutput.coarse_rgb_values                                  */
  float* depth;         /* [R,1]    NerfOutput.coarse_depth_map                                   */
  float* z_vals;        /* [R,N]    NerfOutput.z_vals                                             */
  float* ray_dirs_rep;  /* [R*N,3]  NerfOutput.ray_dirs (unit dirs repeated per sample), nullable */
  float* colors;        /* [R*N,3]  NerfOutput.coarse_colors                                      */
  float* weights;       /* [R,N]    compositing weights (additive extra), nullable                */
  float* z_coarse;      /* [R,Nc]   nullable                                                      */
  float* weights_coarse;/* [R,Nc]   nullable                                                      */
} vfnerf_render_out;

/* The binding side (vfnerf_b200/_lib.py) mirrors these layouts field by field; sizes are part of the ABI. */
#ifdef __cplusplus
#define VFNERF_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define VFNERF_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif
VFNERF_STATIC_ASSERT(sizeof(vfnerf_mlp_desc) == 912, "vfnerf_mlp_desc: 4 + pad 4 ... = 8 + 2*16*4 + 6*16*8 + 8 bytes");
VFNERF_STATIC_ASSERT(sizeof(vfnerf_render_cfg) == 120, "vfnerf_render_cfg: 12*4 + 3*8 + 7*4 + 4 (pad) + 2*8 bytes");
VFNERF_STATIC_ASSERT(sizeof(vfnerf_render_out) == 80, "vfnerf_render_out: 10 pointers");


int vfnerf_abi_version(void);
const char* vfnerf_last_error(void);
/* number of kernels this library has launched in this process (monotonic; for gpu_launches accounting) */
long long vfnerf_launch_count(void);

/* ---- whole path ------------------------------------------------------------------------------- */

/* Bytes of workspace render_fwd needs for cfg->n_rays rays.  keep_for_backward != 0 reserves the
 * activation stash that vfnerf_render_bwd consumes. */
int64_t vfnerf_render_workspace_bytes(const vfnerf_render_cfg* cfg, const vfnerf_mlp_desc* vf,
                                      const vfnerf_mlp_desc* rn, int keep_for_backward);

/* VectorFieldNerf.render(pose, pixels, intrinsics, epoch) in eval mode, rendering="volsdf"
 * (vector_field_nerf.py:216-338).  uv [R,2]; pose [R,4,4] or [R,7]; intrinsics [R,4,4];
 * t_vals [Nc] = torch.linspace(0,1,Nc) made by the host (ray_sampler.py:129);
 * U1 [R,Nc], U2 [R,Nf] (read only when cfg->perturb), U3 [R,Nf] (always read) are the three
 * uniform draws of ray_sampler.py:138,292,297.  z_override (nullable, [R,N]) replaces the merged z
 * values of the second pass (parity protocol: the argmax that places fine samples is discontinuous).
 * density_params = {beta, scale, mean} raw (unclamped) parameters on the device. */
int vfnerf_render_fwd(const vfnerf_render_cfg* cfg,
                      const vfnerf_mlp_desc* vf, const float* vf_arena,
                      const vfnerf_mlp_desc* rn, const float* rn_arena,
                      const float* density_params,
                      const float* uv, const float* pose, const float* intrinsics,
                      const float* t_vals, const float* U1, const float* U2, const float* U3,
                      const float* z_override,
                      const vfnerf_render_out* out,
                      void* workspace, int64_t workspace_bytes, int keep_for_backward,
                      void* stream);

/* Backward of the second (merged) pass -- the autograd graph the reference builds at
 * vector_field_nerf.py:281-323 (the coarse pass runs under no_grad, :252; sample positions carry
 * no grad; normals are detached on the way into the colour net, rendering_network.py:76-77).
 * d_rgb [R,3], d_depth [R,1], d_normals [R,N,3] (nullable), d_colors [R*N,3] (nullable).
 * Gradients are WRITTEN (not accumulated) into grad arenas laid out like the parameter arenas and
 * into d_density[3] = d(beta, scale, mean).  `workspace` must be the one render_fwd filled with
 * keep_for_backward != 0 for the same cfg; `out` the same output block. */
int vfnerf_render_bwd(const vfnerf_render_cfg* cfg,
                      const vfnerf_mlp_desc* vf, const float* vf_arena,
                      const vfnerf_mlp_desc* rn, const float* rn_arena,
                      const float* density_params,
                      const vfnerf_render_out* out,
                      const float* d_rgb, const float* d_depth, const float* d_normals,
                      const float* d_colors,
                      float* vf_grad_arena, float* rn_grad_arena, float* d_density,
                      void* workspace, int64_t workspace_bytes, void* stream);

/* ---- VF-only query: VectorFieldNetwork.__call__ in eval mode -------------------------------- */
/* (vector_field_network.py:177-208; the marching-cubes grid query of mc_utils.py:88-104 and the
 * supervision-point queries of train/vector_field_nerf_train.py:191,203,215).
 * points [P,3] -> out [P,out_ld] with the first n_out_cols columns of [v(3), feat(F)] written
 * (n_out_cols = 3 for the grid query, 3+F for the module call). */
int64_t vfnerf_vf_workspace_bytes(const vfnerf_mlp_desc* vf, int64_t n_points, int multires,
                                  int keep_for_backward, int precision);
int vfnerf_vf_fwd(const vfnerf_mlp_desc* vf, const float* vf_arena, int multires, int skip_layer,
                  float bn_eps, int precision, const float* points, int64_t n_points,
                  float* out, int64_t out_ld, int n_out_cols,
                  void* workspace, int64_t workspace_bytes, int keep_for_backward, void* stream);
/* d_out [P,d_ld] holds dL/d(out[:, :n_out_cols]); grads are ACCUMULATED into vf_grad_arena when
 * accumulate != 0 (the trainer sums the render() and supervision-point contributions). */
int vfnerf_vf_bwd(const vfnerf_mlp_desc* vf, const float* vf_arena, int multires, int skip_layer,
                  float bn_eps, int precision, int64_t n_points,
                  const float* out, int64_t out_ld, const float* d_out, int64_t d_ld, int n_out_cols,
                  float* vf_grad_arena, int accumulate,
                  void* workspace, int64_t workspace_bytes, void* stream);

/* ---- fused VFLoss (models/losses/vf_loss.py:34-87) ------------------------------------------
 * Forward: all six terms in one reduction launch, no host synchronisation.  rgb/rgb_gt [n_rays,3]; depth/depth_gt
 * [n_rays,1] (depth_gt NULL: the batch has no depth, term = 0); normals [n_points,3] (raw VF vectors); sup/sup_gt
 * [n_sup,3] (n_sup may be 0); dd [n_dd] directional derivatives or NULL.  weights[6] in the order rgb, depth, unit_norm,
 * supervision, norm_smaller_than_one, directional_derivatives; norm_lt1_active = (epoch >= norm_smaller_than_one_start).
 * terms: 16 device floats -- [0..5] the unweighted terms in that order, [6] the weighted total, the rest scratch.
 * Backward: gradients of the total times the device scalar *grad_loss; any output pointer may be NULL. */
int vfnerf_vf_loss_fwd(int64_t n_rays, int64_t n_points, int64_t n_sup, int64_t n_dd, const float* rgb,
                       const float* rgb_gt, const float* depth, const float* depth_gt, const float* normals,
                       const float* sup, const float* sup_gt, const float* dd, const float* weights,
                       float depth_clamp, int norm_lt1_active, float* terms, void* stream);
int vfnerf_vf_loss_bwd(int64_t n_rays, int64_t n_points, int64_t n_sup, const float* rgb, const float* rgb_gt,
                       const float* depth, const float* depth_gt, const float* normals, const float* sup,
                       const float* sup_gt, const float* weights, float depth_clamp, int norm_lt1_active,
                       const float* grad_loss, float* d_rgb, float* d_depth, float* d_normals, float* d_sup,
                       void* stream);

/* ---- clip_grad_norm_ + Adam on the flat arenas (train/vector_field_nerf_train.py:252-258) --------------------------
 * vfnerf_sqnorm_accumulate: *out_sq += sum (pre_scale * g[i])^2 (call once per arena, after zeroing *out_sq).  The result
 * is reproducible run to run (per-block partials summed in a fixed order, no float atomics): `scratch` holds
 * VFNERF_SQNORM_SCRATCH_FLOATS floats, zeroed once by the caller before the first call and left zero by every call.
 * vfnerf_adam_step: torch.optim.Adam's update (amsgrad off) of p[0..n) from g, first scaling g IN PLACE by pre_scale
 * (1/world_size after a summing all-reduce of the gradient arena, else 1) and by
 * min(1, max_norm / (sqrt(*total_sqnorm) + 1e-6)) when max_norm > 0.  mask (optional, one byte per element) marks the
 * trainable elements; lr and step are device scalars (step = number of updates including this one), so the call is
 * CUDA-graph capturable and a scheduler can change the rate without touching the graph. */
#define VFNERF_SQNORM_SCRATCH_FLOATS 1024
int vfnerf_sqnorm_accumulate(const float* g, int64_t n, float pre_scale, float* out_sq, float* scratch, void* stream);
int vfnerf_adam_step(float* p, float* g, float* m, float* v, const uint8_t* mask, int64_t n, const float* lr,
                     const float* step, float beta1, float beta2, float eps, float weight_decay, float max_norm,
                     const float* total_sqnorm, float pre_scale, void* stream);

/* ---- both MLPs on given points: the VF + colour evaluation of VectorFieldNerf.get_colors ---------------- */
/* (vector_field_nerf.py:341-375 / the merged pass of render(), :292-321).  points [P,3]; ray_dirs [P/samples_per_ray,3]
 * unit view directions, one per ray; normals [P,3] = tanh VF vectors; colors [P,3] = sigmoid colour-net output.
 * One fused tcgen05 launch (tensor-core precisions only: bf16, bf16x3, fp16f8); repack != 0 re-tiles the weights first (needed after any
 * parameter update; 0 reuses the images already in `workspace`). */
int64_t vfnerf_mlp_points_workspace_bytes(const vfnerf_mlp_desc* vf, const vfnerf_mlp_desc* rn, int multires,
                                          int multires_view, int skip_layer);
int vfnerf_mlp_points_fwd(const vfnerf_mlp_desc* vf, const float* vf_arena, const vfnerf_mlp_desc* rn,
                          const float* rn_arena, int multires, int multires_view, int skip_layer, float bn_eps,
                          int precision, const float* points, const float* ray_dirs, int samples_per_ray,
                          int64_t n_points, float* normals, float* colors, void* workspace,
                          int64_t workspace_bytes, int repack, void* stream);

/* Same query with grid coordinates generated in-kernel (evaluation/methods.py:194-208): point i has
 * integer coordinates (ix,iy,iz) = unravel(i0 + i, [res,res,res]) (z fastest) and position
 * ((index * voxel + origin) + translation) + centroid per axis -- the reference's fp32 op order.
 * origin/translation/centroid are 3 floats each in HOST memory.  out[i, 0:3] = v. */
int vfnerf_vf_grid_query(const vfnerf_mlp_desc* vf, const float* vf_arena, int multires, int skip_layer,
                         float bn_eps, int precision, int res, int64_t i0, int64_t n_points,
                         const float* origin3_host, const float* translation3_host,
                         const float* centroid3_host, float voxel, float* out,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* ---- host helper ------------------------------------------------------------------------------ */
/* n uniform floats in [0,1) from torch's CPU generator engine (MT19937, 24-bit float conversion), identical to what
 * torch.rand would draw from the same state; state624 / left / next are the engine fields of torch.get_rng_state()
 * and are advanced in place.  HOST pointers; no device work.  Replaces the draws of ray_sampler.py:138,292,297. */
int vfnerf_mt19937_uniform(uint32_t* state624, int32_t* left, uint32_t* next, int64_t n, float* out);

/* ---- marching-cubes preprocessing of the dense grid --------------------- */
/* evaluation/methods.py:209-278 (default flags) = mc_utils.extract_divergence (:34-85) + unify_direction (:107-166) +
 * make_comb_format (:169-223) + the block-ordered compaction, fused per cell.  pred [res^3,3] is the grid query's
 * output (x index slowest).  Cells are visited in the reference's order: 2x2x2 blocks in C order, `inc` order inside.
 *   vfnerf_mc_count: keep [8*(res/2)^3] = 1 for cells whose corner pairs differ; cta_counts [ceil(n/256)] kept cells per
 *     256-cell CTA.  Optional dense outputs in grid order for tests: div_raw [res^3] (raw divergence of interior cells,
 *     caller zero-fills), choice [res^3] (bit s = side of corner s).  surface (nullable) [res^3] u8 in grid order: surface
 *     cells decided elsewhere (the reference's smooth_after: divergence of the raw field, sides of the smoothed one).
 *   vfnerf_mc_emit: cta_offsets = inclusive scan of cta_counts (int64); writes cells [M,3], comb [M,28], udf [M,28,2]
 *     -- exactly the arrays contrastive_marching_cubes receives (methods.py:272-283) before their flattening reshape. */
#define VFNERF_MC_CTA 256
int vfnerf_mc_count(const float* pred, int resolution, uint8_t* keep, int32_t* cta_counts, float* div_raw,
                    uint8_t* choice, const uint8_t* surface, void* stream);
int vfnerf_mc_emit(const float* pred, int resolution, const uint8_t* keep, const int64_t* cta_offsets,
                   int32_t* cells, float* comb, float* udf, void* stream);

/* smooth_vf (evaluation/utils/guassian_smoothing.py:81-97; the optional smooth_after / smooth_all steps of
 * methods.py:211-218): replicate-padded depthwise 3-D gaussian of the [res^3,3] grid as three 1-D passes.  taps_host: the
 * kernel_size (odd, <= 31) normalised 1-D taps, HOST floats; tmp: a second [res^3,3] device buffer; out may be `in`'s size
 * only (no aliasing with tmp). */
int vfnerf_smooth_vf(const float* in, float* tmp, float* out, int resolution, int kernel_size,
                     const float* taps_host, void* stream);

/* ---- stage entry points (one per SURVEY.md §8(a) row; used by the parity tests) ------------- */
/* a1: get_ray_directions_and_cam_location, utils/rendering.py:12-60 */
int vfnerf_ray_geometry(int n_rays, int pose_is_quat, const float* uv, const float* pose,
                        const float* intrinsics, float* directions, float* ray_dirs, float* cam_loc,
                        void* stream);
/* a2: UniformSampler.get_z_vals + RaySampler.sample, ray_sampler.py:113-142, 49-80 */
int vfnerf_coarse_sample(int n_rays, int n_coarse, double near_, double far_, int perturb,
                         const float* t_vals, const float* U1, const float* directions,
                         const float* cam_loc, float* z, float* points, void* stream);
/* a7: RangeFineSampler.get_z_vals + sample, ray_sampler.py:264-302 */
int vfnerf_fine_sample(int n_rays, int n_coarse, int n_fine, double near_, double far_,
                       double fine_range, int perturb, const float* z_coarse, const float* w_coarse,
                       const float* U2, const float* U3, const float* directions, const float* cam_loc,
                       float* z, float* points, void* stream);
/* SURVEY.md 8f rank 4 -- the inverse-CDF importance sampler, FineSampler (ray_sampler.py:145-237); not called by the
 * reference's render().  sample_pdf (:163-215): bins [R,n_bins], weights [R,n_bins-1] -> samples [R,n_samples];
 * u is the reference's uniform tensor made explicit: [n_samples] (its linspace, deterministic, u_per_ray = 0) or
 * [R,n_samples] (its torch.rand draw, u_per_ray = 1).  pdf_fine_sample = get_z_vals (:217-237): midpoint bins of
 * z_coarse [R,Nc], weights w_coarse[:, 1:-1], result merged and sorted with z_coarse -> z [R,Nc+Nf] (+ points,
 * nullable).  Warp prefix scan + binary search; parity is to 1e-5 (aten's CPU summation order is not reproduced). */
int vfnerf_sample_pdf(int n_rays, int n_bins, int n_samples, const float* bins, const float* weights,
                      const float* u, int u_per_ray, float* samples, void* stream);
int vfnerf_pdf_fine_sample(int n_rays, int n_coarse, int n_fine, const float* z_coarse, const float* w_coarse,
                           const float* u, int u_per_ray, const float* directions, const float* cam_loc,
                           float* z, float* points, void* stream);
/* a4+a5+a6: window_cosine_similarity + get_density + volsdf_volume_rendering
 * (functions.py:41-72, vector_field_nerf.py:442-474, rendering.py:122-148).
 * normals [R,N,*] with row stride normals_ld floats per sample. */
int vfnerf_density_weights(const vfnerf_render_cfg* cfg, int n_samples, const float* density_params,
                           const float* normals, int64_t normals_ld, const float* ray_dirs,
                           const float* z, float* cosw, float* sigma, float* weights, void* stream);
/* utils/rendering.py as stand-alone ops on a given density sigma [R,N] and z [R,N] -> weights [R,N]:
 * mode 0 = volsdf_volume_rendering (:122-148), mode 1 = nerf_volume_rendering (:98-119, inclusive cumprod, +1e-10).
 * The reference's render() calls the latter with swapped arguments, so rendering="nerf" stays rejected in
 * render(); this is the corrected-order op of SURVEY.md 8f rank 4. */
int vfnerf_volume_weights(int n_rays, int n_samples, int mode, int normalize, const float* sigma,
                          const float* z, float* weights, void* stream);
/* a9: rgb = sum_j w_j c_j, depth = sum_j w_j z_j, vector_field_nerf.py:322-323 */
int vfnerf_composite(int n_rays, int n_samples, const float* weights, const float* colors,
                     const float* z, float* rgb, float* depth, void* stream);
/* same with the white-background term: rgb += 1 - sum_j w_j (vector_field_nerf.py:325-329) */
int vfnerf_composite_white(int n_rays, int n_samples, const float* weights, const float* colors,
                           const float* z, float* rgb, float* depth, void* stream);

/* ---- supervision points of the trainer (train/vector_field_nerf_train.py:180-216) --------------------------------- */
/* SphereSampler.sample (models/samplers/sampler.py:160-193) + sample_border_points / sample_center_points
 * (models/helpers/functions.py:99-130) from GIVEN float64 draws phi in [0,2pi), cos_theta in [-1,1), u in [0,1) (device
 * pointers; the reference draws them with np.random.uniform in this order): points[n,3] = fp32(sphere point) + centroid,
 * gt[n,3] = normalize(centroid - p) (inward = 1: border points) or normalize(p - centroid) (inward = 0: centre points). */
int vfnerf_sphere_points(int64_t n, const double* phi, const double* cos_theta, const double* u, double r_max, double r_min,
                         const float* centroid3_host, int inward, float* points, float* gt, void* stream);
/* get_border_indices_and_gt (mode 0: |p - centroid| > threshold, gt = normalize(centroid - p), functions.py:75-97) and
 * get_center_indices_and_gt (mode 1: |p - centroid| < threshold, gt = normalize(p - centroid), :132-154) over all n ray
 * samples in one pass: flag[n] (uint8) marks the selected samples, gt[n,3] holds every sample's target.  The caller
 * compacts (the reference indexes with the boolean mask, which synchronises the same way). */
int vfnerf_select_supervised(int64_t n, const float* points, const float* centroid3_host, float threshold, int mode,
                             uint8_t* flag, float* gt, void* stream);

/* ---- train mode: BatchNorm batch statistics, the autograd "Jacobian", directional derivatives ---- */
/* What render() and the VF-only call compute after model.train() (models/nerf/vector_field_nerf.py:84-101 puts both
 * networks into training mode): every hidden layer normalises with the mean and biased variance of ITS BATCH
 * (vector_field_network.py:177-208, rendering_network.py:62-108 with nn.BatchNorm1d in training mode) and folds them into
 * the running statistics IN THE ARENA (momentum bn_momentum = 0.1, unbiased variance) -- hence the non-const arenas; the
 * module's num_batches_tracked counters are the caller's to advance (VF net: +2 per render, colour net: +1).
 * fp32 layer-wise path only (cfg.precision must be VFNERF_PREC_FP32).
 *
 * vfnerf_render_train_fwd: render() as in vfnerf_render_fwd plus dir_derivs [4 * n_rays * n_coarse] (optional) =
 * NerfOutput.directional_derivtives: per-row norms of compute_directional_derivatives (vector_field_nerf.py:476-498) on the
 * COARSE pass, the block concatenated with itself as upstream does (:305).  The Jacobian behind it is the reference's
 * (vector_field_network.py:146-171): three reverse sweeps of output-column SUMS over the batch through the batch
 * statistics.  The workspace always keeps what vfnerf_render_train_bwd needs.
 * vfnerf_render_train_bwd: same contract as vfnerf_render_bwd; gradients flow through the batch statistics (exact
 * BatchNorm-training backward).  Linear biases in front of a BatchNorm get exactly 0 (autograd: rounding noise ~1e-8). */
int64_t vfnerf_render_train_workspace_bytes(const vfnerf_render_cfg* cfg, const vfnerf_mlp_desc* vf,
                                            const vfnerf_mlp_desc* rn);
int vfnerf_render_train_fwd(const vfnerf_render_cfg* cfg, const vfnerf_mlp_desc* vf, float* vf_arena,
                            const vfnerf_mlp_desc* rn, float* rn_arena, const float* density_params, const float* uv,
                            const float* pose, const float* intrinsics, const float* t_vals, const float* U1,
                            const float* U2, const float* U3, const float* z_override, const vfnerf_render_out* out,
                            float* dir_derivs, float bn_momentum, void* workspace, int64_t workspace_bytes, void* stream);
int vfnerf_render_train_bwd(const vfnerf_render_cfg* cfg, const vfnerf_mlp_desc* vf, const float* vf_arena,
                            const vfnerf_mlp_desc* rn, const float* rn_arena, const float* density_params,
                            const vfnerf_render_out* out, const float* d_rgb, const float* d_depth, const float* d_normals,
                            const float* d_colors, float* vf_grad_arena, float* rn_grad_arena, float* d_density,
                            void* workspace, int64_t workspace_bytes, void* stream);
/* VectorFieldNetwork.forward in training mode (vector_field_network.py:140-175), the call the trainer makes on its
 * supervision points (train/vector_field_nerf_train.py:191,204,217): out[:, :n_out_cols] = tanh outputs,
 * jacobian [n,9] (optional, row stride jacobian_ld) = cat(d sum(y0)/dx, d sum(y1)/dx, d sum(y2)/dx).  The workspace keeps
 * what vfnerf_vf_train_bwd needs (same arguments as vfnerf_vf_bwd, fp32). */
int64_t vfnerf_vf_train_workspace_bytes(const vfnerf_mlp_desc* vf, int64_t n_points, int multires);
int vfnerf_vf_train_fwd(const vfnerf_mlp_desc* vf, float* vf_arena, int multires, int skip_layer, float bn_eps,
                        float bn_momentum, const float* points, int64_t n_points, float* out, int64_t out_ld,
                        int n_out_cols, float* jacobian, int64_t jacobian_ld, void* workspace, int64_t workspace_bytes,
                        void* stream);
int vfnerf_vf_train_bwd(const vfnerf_mlp_desc* vf, const float* vf_arena, int multires, int skip_layer, int64_t n_points,
                        const float* out, int64_t out_ld, const float* d_out, int64_t d_ld, int n_out_cols,
                        float* vf_grad_arena, int accumulate, void* workspace, int64_t workspace_bytes, void* stream);

/* Test-only entry points (UMMA descriptor probes, micro-benchmarks, activation-stash read-back) are NOT part of this
 * library: they are declared in vfnerf_b200_debug.h and built into a separate libvfnerf_b200_debug.so by the tests. */

#ifdef __cplusplus
}
#endif
#endif /* VFNERF_B200_H */
