// C-ABI entry points (include/vfnerf_b200.h) and the host-side orchestration of the render() path.
//
// Order of execution follows VectorFieldNerf.render() (vector_field_nerf.py:216-338):
//   rays -> coarse z/points -> VF MLP (no grad) -> density -> weights -> fine z/points (merged, sorted)
//   -> VF MLP on ALL merged points -> density -> weights -> colour MLP -> composite.
// Everything is enqueued on the caller's stream; the only memory used is the caller's workspace.
#include <cstdarg>
#include <cstring>
#include <cmath>

#include "common.cuh"
#include "host_plan.cuh"
#include "mlp_tc.cuh"
#ifdef VFNERF_DEBUG_EXPORTS
#include "../../include/vfnerf_b200_debug.h"
#endif

namespace vfn {

static thread_local char g_err[1024] = "";
long long g_launches = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// split-precision tile variant of a precision (mlp_tc.cuh): 0 plain bf16, 1 bf16x3, 2 fp16 + fp8 remainders
static int split_mode(int precision) {
  return precision == VFNERF_PREC_BF16X3 ? 1 : (precision == VFNERF_PREC_FP16F8 ? 2 : 0);
}

// ---------------------------------------------------------------------------------------------
// fp32 MLP forward / backward (generic widths)
// ---------------------------------------------------------------------------------------------

// VF MLP on n points.  x0 [n, in_dim[0]] = embedded input.  The last layer writes the first
// n_out_cols columns of tanh(...) to out (row stride out_ld).
static int vf_forward_fp32(const vfnerf_mlp_desc& d, const float* arena, const MlpBufs& b, int skip_layer,
                           const float* x0, int64_t n, float* out, int64_t out_ld, int n_out_cols,
                           cudaStream_t s) {
  const float* x = x0;
  int64_t x_ld = d.in_dim[0];
  for (int l = 0; l < d.n_layers; ++l) {
    const bool last = (l == d.n_layers - 1);
    GemmArgs g{};
    g.A = x; g.a_rs = x_ld; g.a_cs = 1; g.a_kscale = nullptr;
    g.B = arena + d.w_off[l]; g.b_rs = 1; g.b_cs = d.in_dim[l];
    g.M = n; g.K = d.in_dim[l];
    g.scale = b.scale + b.soff[l]; g.shift = b.shift + b.soff[l];
    g.split_k = 1; g.mask = nullptr; g.mask_rs = 0;
    if (last) {
      g.C = out; g.c_rs = out_ld; g.N = n_out_cols; g.act = ACT_TANH; g.post_div = 0.f;
    } else {
      g.C = b.act[l]; g.c_rs = d.in_dim[l + 1]; g.N = d.out_dim[l]; g.act = ACT_RELU;
      g.post_div = (l + 1 == skip_layer) ? kSqrt2 : 0.f;
    }
    if (int e = launch_gemm(g, s)) return e;
    x = b.act[l];
    x_ld = last ? 0 : d.in_dim[l + 1];
  }
  return 0;
}

// writes the embedding of `points` to emb [n,E] and, if the net has a skip layer, emb/sqrt(2) into
// the tail columns of the skip layer's input buffer.
static int vf_embed_fp32(const vfnerf_mlp_desc& d, const MlpBufs& b, int multires, int skip_layer,
                         const float* points, int64_t n, float* emb, cudaStream_t s) {
  const int E = 3 + 6 * multires;
  if (int e = launch_embed(points, 3, n, multires, 0.f, emb, E, s)) return e;
  if (skip_layer > 0)
    if (int e = launch_embed(points, 3, n, multires, kSqrt2, b.act[skip_layer - 1] + d.out_dim[skip_layer - 1],
                             d.in_dim[skip_layer], s)) return e;
  return 0;
}

// Generic MLP backward.  dY [n, out_dim[L-1]] (row stride dy_ld) is the gradient wrt the last layer's
// pre-activation output.  x0 is the input of layer 0.  Hidden activations are ReLU; `skip_layer`'s
// producer was divided by sqrt(2).  If d_in0 != NULL, the gradient wrt columns
// [in0_col0, in0_col0 + in0_cols) of x0 is written there (row stride d_in0_ld).
static int mlp_backward_fp32(const vfnerf_mlp_desc& d, const float* arena, const MlpBufs& b, int skip_layer,
                             float bn_eps, const float* x0, int64_t x0_ld, int64_t n, const float* dY,
                             int64_t dy_ld, const BwdBufs& w, float* grad_arena, int accumulate,
                             float* d_in0, int64_t d_in0_ld, int in0_col0, int in0_cols, cudaStream_t s) {
  const float* dy = dY;
  int64_t ld = dy_ld;
  float* pp[2] = {w.dA, w.dB};
  int flip = 0;
  const int split = (int)std::min<int64_t>(128, std::max<int64_t>(1, n / 1024));
  for (int l = d.n_layers - 1; l >= 0; --l) {
    const float* x = (l == 0) ? x0 : b.act[l - 1];
    const int64_t x_ld = (l == 0) ? x0_ld : d.in_dim[l];
    const int No = d.out_dim[l], Ki = d.in_dim[l];
    if (int e = launch_colsum(dy, ld, n, No, nullptr, w.colsum, s)) return e;
    VFN_CHECK_CUDA(cudaMemsetAsync(w.G, 0, sizeof(float) * (int64_t)No * Ki, s));
    GemmArgs g{};
    g.A = dy; g.a_rs = 1; g.a_cs = ld;            // A(m = out channel, k = point)
    g.B = x; g.b_rs = x_ld; g.b_cs = 1;           // B(k = point, n = in channel)
    g.C = w.G; g.c_rs = Ki; g.M = No; g.N = Ki; g.K = n; g.split_k = split;
    if (int e = launch_gemm(g, s)) return e;
    if (int e = launch_grad_finalize(d, l, arena, bn_eps, w.colsum, grad_arena, accumulate, w.G, s)) return e;
    if (l > 0) {
      // dX[:, :out_dim[l-1]] = ((dY * scale) W)[:, :out_dim[l-1]] masked by ReLU of the producer
      GemmArgs h{};
      h.A = dy; h.a_rs = ld; h.a_cs = 1; h.a_kscale = b.scale + b.soff[l];
      h.B = arena + d.w_off[l]; h.b_rs = Ki; h.b_cs = 1;
      h.C = pp[flip]; h.c_rs = d.out_dim[l - 1]; h.M = n; h.N = d.out_dim[l - 1]; h.K = No; h.split_k = 1;
      h.post_div = (l == skip_layer) ? kSqrt2 : 0.f;
      h.mask = b.act[l - 1]; h.mask_rs = d.in_dim[l];
      if (int e = launch_gemm(h, s)) return e;
      dy = pp[flip]; ld = d.out_dim[l - 1];
      flip ^= 1;
    } else if (d_in0) {
      GemmArgs h{};
      h.A = dy; h.a_rs = ld; h.a_cs = 1; h.a_kscale = b.scale + b.soff[0];
      h.B = arena + d.w_off[0] + in0_col0; h.b_rs = Ki; h.b_cs = 1;
      h.C = d_in0; h.c_rs = d_in0_ld; h.M = n; h.N = in0_cols; h.K = No; h.split_k = 1;
      if (int e = launch_gemm(h, s)) return e;
    }
  }
  return 0;
}

static int rn_forward_fp32(const vfnerf_mlp_desc& d, const float* arena, const MlpBufs& b, const float* cin,
                           int64_t cin_ld, int64_t n, float* colors, cudaStream_t s) {
  const float* x = cin;
  int64_t x_ld = cin_ld;
  for (int l = 0; l < d.n_layers; ++l) {
    const bool last = (l == d.n_layers - 1);
    GemmArgs g{};
    g.A = x; g.a_rs = x_ld; g.a_cs = 1;
    g.B = arena + d.w_off[l]; g.b_rs = 1; g.b_cs = d.in_dim[l];
    g.M = n; g.K = d.in_dim[l]; g.N = d.out_dim[l];
    g.scale = b.scale + b.soff[l]; g.shift = b.shift + b.soff[l]; g.split_k = 1;
    if (last) { g.C = colors; g.c_rs = d.out_dim[l]; g.act = ACT_SIGMOID; }
    else { g.C = b.act[l]; g.c_rs = d.in_dim[l + 1]; g.act = ACT_RELU; }
    if (int e = launch_gemm(g, s)) return e;
    if (!last) { x = b.act[l]; x_ld = d.in_dim[l + 1]; }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// render plan
// ---------------------------------------------------------------------------------------------
struct RenderPlan {
  int R, Nc, Nf, N, E, Ev, F, cin_ld;
  int64_t P, Pc;
  float *directions, *ray_dirs, *cam_loc, *z_c, *pts_c, *normals_c, *w_c, *emb, *cin, *weights;
  // bf16 inference: coarse results are reused in the merged pass, only the fine candidates are evaluated again
  int reuse_coarse;
  float *normals_cf, *colors_cf;   // [P,3] in evaluation order: all coarse candidates of all rays, then all fine candidates
  float *colors_c, *pts_f, *normals_f, *colors_f;
  uint8_t* src;      // [R,N] candidate index (coarse 0..Nc-1, fine Nc..N-1) at every merged position
  MlpBufs vf, rn;
  BwdBufs bw;
  float* d_out;      // [P, 3+F]   gradient wrt the VF net's tanh outputs / pre-activations
  float* d_colors;   // [P, 3]
  TcPlan tc;         // tensor-core path buffers (mlp_tc.cu)
  int64_t bytes;
};

static int make_plan(const vfnerf_render_cfg& cfg, const vfnerf_mlp_desc& vf, const vfnerf_mlp_desc& rn,
                     int keep, void* ws, RenderPlan& p, int have_z_override = 0) {
  p.R = cfg.n_rays; p.Nc = cfg.n_coarse; p.Nf = cfg.n_fine; p.N = p.Nc + p.Nf;
  p.P = (int64_t)p.R * p.N; p.Pc = (int64_t)p.R * p.Nc;
  p.E = 3 + 6 * cfg.multires; p.Ev = 3 + 6 * cfg.multires_view;
  p.F = vf.out_dim[vf.n_layers - 1] - 3;
  p.cin_ld = 3 + p.Ev + 3 + p.F;
  VFN_REQUIRE(p.R >= 0 && p.Nc >= 2, "render: n_rays=%d n_coarse=%d invalid", p.R, p.Nc);
  VFN_REQUIRE(p.Nf >= 2, "render: n_fine=%d; the reference's render() needs fine sampling (n_importance > 0)", p.Nf);
  VFN_REQUIRE(p.N <= VFNERF_MAX_SAMPLES, "render: %d samples per ray exceed %d", p.N, VFNERF_MAX_SAMPLES);
  if (int e = validate_vf(vf, cfg.multires, cfg.skip_layer)) return e;
  VFN_REQUIRE(rn.in_dim[0] == p.cin_ld, "colour net: in_dim[0]=%d, expected %d (mode 'idr')", rn.in_dim[0], p.cin_ld);
  VFN_REQUIRE(rn.out_dim[rn.n_layers - 1] == 3, "colour net: output dim must be 3");
  Carver c(ws);
  p.directions = c.f(3 * p.R); p.ray_dirs = c.f(3 * p.R); p.cam_loc = c.f(3 * p.R);
  p.z_c = c.f(p.Pc); p.pts_c = c.f(3 * p.Pc); p.normals_c = c.f(3 * p.Pc); p.w_c = c.f(p.Pc);
  p.weights = c.f(p.P);
  p.cin = nullptr;
  p.d_out = nullptr; p.d_colors = nullptr;
  if (cfg.precision == VFNERF_PREC_FP32) {
    p.cin = c.f(p.P * p.cin_ld);
    p.emb = c.f(p.P * p.E);
    carve_mlp(c, vf, p.P, p.vf);
    carve_mlp(c, rn, p.P, p.rn);
    if (keep) {
      carve_bwd(c, vf, rn, p.P, p.bw);
      p.d_out = c.f(p.P * (3 + p.F));
      p.d_colors = c.f(3 * p.P);
    }
  } else {
    if (int e = tc_carve(c.base, c.off, cfg.multires, cfg.multires_view, cfg.skip_layer, vf, &rn, p.tc, p.P, keep,
                         split_mode(cfg.precision))) return e;
  }
  // Sized for every call (the workspace query does not know about z_override); used when it applies.  With a stash
  // (training) the two launches fill consecutive tile ranges of the same stash, so the coarse block must end on a tile
  // boundary; the backward then runs over the points in evaluation order (render_tail_bwd scatters through `src`).
  p.reuse_coarse = cfg.precision != VFNERF_PREC_FP32 && !(cfg.flags & VFNERF_FLAG_RECOMPUTE_COARSE) &&
                   (!keep || p.Pc % 128 == 0);
  p.normals_cf = p.colors_cf = p.colors_c = p.pts_f = p.normals_f = p.colors_f = nullptr; p.src = nullptr;
  if (p.reuse_coarse) {
    const int64_t Pf = (int64_t)p.R * p.Nf;
    p.normals_cf = c.f(3 * p.P); p.colors_cf = c.f(3 * p.P); p.pts_f = c.f(3 * Pf);
    p.src = reinterpret_cast<uint8_t*>(c.f((p.P + 3) / 4));
    p.normals_c = p.normals_cf; p.normals_f = p.normals_cf + 3 * p.Pc;
    p.colors_c = p.colors_cf; p.colors_f = p.colors_cf + 3 * p.Pc;
    if (have_z_override) p.reuse_coarse = 0;
  }
  p.bytes = c.off;
  return 0;
}

static int check_precision(const vfnerf_render_cfg& cfg) {
  VFN_REQUIRE(cfg.precision == VFNERF_PREC_FP32 || cfg.precision == VFNERF_PREC_BF16 ||
              cfg.precision == VFNERF_PREC_BF16X3 || cfg.precision == VFNERF_PREC_FP16F8, "unknown precision %d", cfg.precision);
  return 0;
}

}  // namespace vfn

using namespace vfn;

extern "C" {

int vfnerf_abi_version(void) { return 1; }
long long vfnerf_launch_count(void) { return g_launches; }
const char* vfnerf_last_error(void) { return g_err; }

int64_t vfnerf_render_workspace_bytes(const vfnerf_render_cfg* cfg, const vfnerf_mlp_desc* vf,
                                      const vfnerf_mlp_desc* rn, int keep_for_backward) {
  if (!cfg || !vf || !rn) { set_error("null argument"); return -1; }
  if (check_precision(*cfg)) return -1;
  RenderPlan p;
  if (make_plan(*cfg, *vf, *rn, keep_for_backward, nullptr, p)) return -1;
  return p.bytes + 256;
}

int vfnerf_render_fwd(const vfnerf_render_cfg* cfg, const vfnerf_mlp_desc* vf, const float* vf_arena,
                      const vfnerf_mlp_desc* rn, const float* rn_arena, const float* density_params,
                      const float* uv, const float* pose, const float* intrinsics, const float* t_vals,
                      const float* U1, const float* U2, const float* U3, const float* z_override,
                      const vfnerf_render_out* out, void* workspace, int64_t workspace_bytes,
                      int keep_for_backward, void* stream) {
  VFN_REQUIRE(cfg && vf && rn && out, "render_fwd: null argument");
  if (int e = check_precision(*cfg)) return e;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(s);
  RenderPlan p;
  if (int e = make_plan(*cfg, *vf, *rn, keep_for_backward, workspace, p, z_override != nullptr)) return e;
  if (p.R == 0) return 0;          // an empty batch is a no-op (its output tensors have no storage to point at)
  VFN_REQUIRE(workspace && workspace_bytes >= p.bytes, "render_fwd: workspace %lld B < required %lld B",
              (long long)workspace_bytes, (long long)p.bytes);
  VFN_REQUIRE(out->points && out->normals && out->rgb && out->depth && out->z_vals && out->colors,
              "render_fwd: required output pointer is null");
  VFN_REQUIRE(!(z_override && keep_for_backward && cfg->precision != VFNERF_PREC_FP32 &&
                !(cfg->flags & VFNERF_FLAG_RECOMPUTE_COARSE)),
              "render_fwd: z_override with keep_for_backward needs VFNERF_FLAG_RECOMPUTE_COARSE in cfg.flags (render_bwd "
              "derives the order of the stash from cfg alone)");
  if (p.R == 0) return 0;
  float* weights = out->weights ? out->weights : p.weights;
  float* z_c = out->z_coarse ? out->z_coarse : p.z_c;
  float* w_c = out->weights_coarse ? out->weights_coarse : p.w_c;

  // a1 + a2 in one launch (render_fused.cu)
  if (int e = launch_ray_head(p.R, cfg->pose_is_quat, uv, pose, intrinsics, p.Nc, cfg->near_, cfg->far_, cfg->perturb, t_vals,
                              U1, p.directions, p.ray_dirs, p.cam_loc, z_c, p.pts_c, s)) return e;

  if (cfg->precision == VFNERF_PREC_FP32) {
    if (int e = launch_fold_bn(*vf, vf_arena, cfg->bn_eps, p.vf.scale, p.vf.shift, s)) return e;
    if (int e = launch_fold_bn(*rn, rn_arena, cfg->bn_eps, p.rn.scale, p.rn.shift, s)) return e;
    // ---- coarse pass (vector_field_nerf.py:252-272), only the 3 vector outputs are needed
    if (!z_override) {
      if (int e = vf_embed_fp32(*vf, p.vf, cfg->multires, cfg->skip_layer, p.pts_c, p.Pc, p.emb, s)) return e;
      if (int e = vf_forward_fp32(*vf, vf_arena, p.vf, cfg->skip_layer, p.emb, p.Pc, p.normals_c, 3, 3, s)) return e;
    }
  } else {
    if (!(cfg->flags & VFNERF_FLAG_WEIGHTS_PACKED) || keep_for_backward)
      if (int e = tc_prepare(*vf, vf_arena, rn, rn_arena, cfg->bn_eps, p.tc, s)) return e;
    if (p.reuse_coarse) {
      // both MLPs on the coarse points now: the merged pass moves these results instead of recomputing them
      if (int e = tc_forward(p.tc, keep_for_backward ? TC_MODE_RENDER_STASH : TC_MODE_RENDER, p.pts_c, nullptr, 0, 0, p.Pc,
                             p.ray_dirs, p.Nc, p.normals_c, 3, nullptr, 0, p.colors_c, s, 0)) return e;
    } else if (!z_override) {
      if (int e = tc_forward(p.tc, TC_MODE_V_ONLY, p.pts_c, nullptr, 0, 0, p.Pc, nullptr, 0, p.normals_c, 3,
                             nullptr, 0, nullptr, s)) return e;
    }
  }
  // ---- coarse weights -> argmax -> fine candidates -> merged, sorted z values and points (vector_field_nerf.py:266-287),
  // one launch; a prescribed z (parity protocol) only needs its points
  if (!z_override) {
    if (int e = launch_coarse_to_fine(*cfg, p.R, p.Nc, p.Nf, density_params, p.normals_c, 3, p.ray_dirs, z_c, U2, U3,
                                      p.directions, p.cam_loc, w_c, out->z_vals, out->points,
                                      p.reuse_coarse ? p.src : nullptr, p.reuse_coarse ? p.pts_f : nullptr, s)) return e;
  } else {
    if (int e = launch_fine_sample(p.R, p.Nc, p.Nf, cfg->fine_near_, cfg->fine_far_, cfg->fine_range, cfg->perturb, z_c, w_c,
                                   U2, U3, z_override, p.directions, p.cam_loc, out->z_vals, out->points, nullptr, nullptr,
                                   s)) return e;
  }
  // ---- merged pass
  const int white = (cfg->flags & VFNERF_FLAG_WHITE_BG) ? 1 : 0;
  if (cfg->precision == VFNERF_PREC_FP32) {
    if (int e = launch_color_input_head(out->points, p.ray_dirs, p.R, p.N, cfg->multires_view, p.cin, p.cin_ld,
                                        out->ray_dirs_rep, s)) return e;
    float* vf_out = p.cin + 3 + p.Ev;      // [v(3), feat(F)] written in place into the colour-net input
    if (int e = vf_embed_fp32(*vf, p.vf, cfg->multires, cfg->skip_layer, out->points, p.P, p.emb, s)) return e;
    if (int e = vf_forward_fp32(*vf, vf_arena, p.vf, cfg->skip_layer, p.emb, p.P, vf_out, p.cin_ld, 3 + p.F, s)) return e;
    if (int e = launch_copy_cols(vf_out, p.cin_ld, out->normals, 3, p.P, 3, s)) return e;
    if (int e = rn_forward_fp32(*rn, rn_arena, p.rn, p.cin, p.cin_ld, p.P, out->colors, s)) return e;
    // density -> weights -> composite in one launch
    return launch_render_tail(*cfg, p.R, p.N, p.Nc, density_params, nullptr, vf_out, p.cin_ld, out->colors, p.ray_dirs,
                              out->z_vals, nullptr, nullptr, weights, out->rgb, out->depth, nullptr, white, s);
  }
  // one fused tcgen05 launch: VF MLP -> normals, colour MLP -> colours (features stay on chip)
  if (p.reuse_coarse) {
    // the fine candidates are the only points not evaluated yet (the merged coarse points carry the same bits as
    // the coarse sweep's, and the fused chain is a pure per-point function): evaluate them, then one launch gathers the
    // per-candidate results into sample order by `src`, forms the weights and composites
    if (int e = tc_forward(p.tc, keep_for_backward ? TC_MODE_RENDER_STASH : TC_MODE_RENDER, p.pts_f, nullptr, 0, 0,
                           (int64_t)p.R * p.Nf, p.ray_dirs, p.Nf, p.normals_f, 3, nullptr, 0, p.colors_f, s,
                           keep_for_backward ? p.Pc / 128 : 0)) return e;
    return launch_render_tail(*cfg, p.R, p.N, p.Nc, density_params, p.src, p.normals_cf, 3, p.colors_cf, p.ray_dirs,
                              out->z_vals, out->normals, out->colors, weights, out->rgb, out->depth, out->ray_dirs_rep,
                              white, s);
  }
  if (int e = tc_forward(p.tc, keep_for_backward ? TC_MODE_RENDER_STASH : TC_MODE_RENDER, out->points, nullptr, 0, 0, p.P,
                         p.ray_dirs, p.N, out->normals, 3, nullptr, 0, out->colors, s)) return e;
  return launch_render_tail(*cfg, p.R, p.N, p.Nc, density_params, nullptr, out->normals, 3, out->colors, p.ray_dirs,
                            out->z_vals, nullptr, nullptr, weights, out->rgb, out->depth, out->ray_dirs_rep, white, s);
}

int vfnerf_render_bwd(const vfnerf_render_cfg* cfg, const vfnerf_mlp_desc* vf, const float* vf_arena,
                      const vfnerf_mlp_desc* rn, const float* rn_arena, const float* density_params,
                      const vfnerf_render_out* out, const float* d_rgb, const float* d_depth,
                      const float* d_normals, const float* d_colors, float* vf_grad_arena,
                      float* rn_grad_arena, float* d_density, void* workspace, int64_t workspace_bytes,
                      void* stream) {
  VFN_REQUIRE(cfg && vf && rn && out && d_rgb && d_depth && vf_grad_arena && rn_grad_arena && d_density,
              "render_bwd: null argument");
  if (int e = check_precision(*cfg)) return e;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(s);
  RenderPlan p;
  if (int e = make_plan(*cfg, *vf, *rn, 1, workspace, p)) return e;
  VFN_REQUIRE(workspace && workspace_bytes >= p.bytes, "render_bwd: workspace %lld B < required %lld B",
              (long long)workspace_bytes, (long long)p.bytes);
  VFN_CHECK_CUDA(cudaMemsetAsync(d_density, 0, 3 * sizeof(float), s));
  VFN_CHECK_CUDA(cudaMemsetAsync(vf_grad_arena, 0, sizeof(float) * vf->arena_floats, s));
  VFN_CHECK_CUDA(cudaMemsetAsync(rn_grad_arena, 0, sizeof(float) * rn->arena_floats, s));
  if (p.R == 0) return 0;
  if (cfg->precision != VFNERF_PREC_FP32) {
    // tensor-core training path (mlp_tc_bwd.cu): the forward left the activation stash and the packed transposed
    // weights in the workspace; ray_dirs are recomputed by the caller-visible forward only, so they must still be there
    float* dcol = p.tc.d3;
    float* dv = p.tc.d3 + 3 * p.P;
    if (p.reuse_coarse) {
      // the forward evaluated (and stashed) the coarse candidates, then the fine ones: the per-sample gradients are
      // written in that order, and the per-point forward outputs are the evaluation-order buffers
      if (int e = launch_render_tail_bwd(*cfg, p.R, p.N, density_params, out->normals, 3, p.ray_dirs, out->z_vals,
                                         out->colors, d_rgb, d_depth, d_normals, d_colors, dcol, dv, 3, d_density, s,
                                         p.src, p.Nc)) return e;
      return tc_backward(p.tc, *vf, vf_arena, *rn, rn_arena, cfg->bn_eps, p.P, p.colors_cf, p.normals_cf, dcol, dv,
                         vf_grad_arena, rn_grad_arena, s);
    }
    if (int e = launch_render_tail_bwd(*cfg, p.R, p.N, density_params, out->normals, 3, p.ray_dirs, out->z_vals,
                                       out->colors, d_rgb, d_depth, d_normals, d_colors, dcol, dv, 3, d_density, s)) return e;
    return tc_backward(p.tc, *vf, vf_arena, *rn, rn_arena, cfg->bn_eps, p.P, out->colors, out->normals, dcol, dv,
                       vf_grad_arena, rn_grad_arena, s);
  }
  const float* vf_out = p.cin + 3 + p.Ev;
  const int Dv = 3 + p.F;
  // a9..a4 fused: d_colors, dL/dv (into columns 0..2 of d_out), density parameter grads
  if (int e = launch_render_tail_bwd(*cfg, p.R, p.N, density_params, vf_out, p.cin_ld, p.ray_dirs, out->z_vals,
                                     out->colors, d_rgb, d_depth, d_normals, d_colors, p.d_colors, p.d_out, Dv,
                                     d_density, s)) return e;
  // colour net: sigmoid', then layers; only the feature columns of its input carry grad
  // (points / view dirs have none, normals are detached: rendering_network.py:76-77)
  if (int e = launch_act_bwd(out->colors, 3, p.d_colors, 3, p.P, 3, ACT_SIGMOID, p.d_colors, 3, s)) return e;
  if (int e = mlp_backward_fp32(*rn, rn_arena, p.rn, -1, cfg->bn_eps, p.cin, p.cin_ld, p.P, p.d_colors, 3, p.bw,
                                rn_grad_arena, 0, p.d_out + 3, Dv, 3 + p.Ev + 3, p.F, s)) return e;
  // VF net: tanh' on [v, feat], then layers
  if (int e = launch_act_bwd(vf_out, p.cin_ld, p.d_out, Dv, p.P, Dv, ACT_TANH, p.d_out, Dv, s)) return e;
  if (int e = mlp_backward_fp32(*vf, vf_arena, p.vf, cfg->skip_layer, cfg->bn_eps, p.emb, p.E, p.P, p.d_out, Dv,
                                p.bw, vf_grad_arena, 0, nullptr, 0, 0, 0, s)) return e;
  return 0;
}

#ifdef VFNERF_DEBUG_EXPORTS
// test-only (include/vfnerf_b200_debug.h): compiled into libvfnerf_b200_debug.so, not into the product library
const char* vfnerf_debug_last_error(void) { return g_err; }

int vfnerf_debug_stash_read(const vfnerf_render_cfg* cfg, const vfnerf_mlp_desc* vf, const vfnerf_mlp_desc* rn,
                            void* workspace, int tensor, float* out, int* n_cols, void* stream) {
  VFN_REQUIRE(cfg && vf && rn && workspace, "stash_read: null argument");
  VFN_REQUIRE(cfg->precision == VFNERF_PREC_BF16, "stash_read: only the bf16 path keeps a stash");
  RenderPlan p;
  if (int e = make_plan(*cfg, *vf, *rn, 1, workspace, p)) return e;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(s);
  if (!p.reuse_coarse || !out) return tc_debug_stash_read(p.tc, tensor, p.P, out, n_cols, s);
  // the stash is in evaluation order (coarse candidates, then fine ones): hand the rows back in merged sample order
  int cols = 0;
  if (int e = tc_debug_stash_read(p.tc, tensor, p.P, nullptr, &cols, s)) return e;
  float* tmp = nullptr;
  VFN_CHECK_CUDA(cudaMalloc(&tmp, sizeof(float) * (size_t)p.P * cols));
  int e = tc_debug_stash_read(p.tc, tensor, p.P, tmp, n_cols, s);
  if (!e) e = launch_rows_to_merged_order(p.R, p.Nc, p.Nf, p.src, tmp, out, cols, s);
  cudaStreamSynchronize(s);
  cudaFree(tmp);
  return e;
}

// host only: the chunk records behind one tensor-core program of the render() plan (no GPU needed)
int vfnerf_debug_chunk_table(const vfnerf_render_cfg* cfg, const vfnerf_mlp_desc* vf, const vfnerf_mlp_desc* rn, int keep_for_backward,
                             int program, uint32_t* records, int* n_chunks, int* step_facts, int* n_steps) {
  VFN_REQUIRE(cfg && vf && rn && records && n_chunks && step_facts && n_steps, "chunk_table: null argument");
  VFN_REQUIRE(cfg->precision != VFNERF_PREC_FP32, "chunk_table: tensor-core precisions only");
  RenderPlan p;
  if (int e = make_plan(*cfg, *vf, *rn, keep_for_backward, nullptr, p)) return e;
  return tc_debug_chunk_table(p.tc, program, records, n_chunks, step_facts, n_steps);
}
#endif  // VFNERF_DEBUG_EXPORTS

// ---- VF-only query -----------------------------------------------------------------------------
struct VfPlan {
  float* emb;
  float* pts;     // grid query only
  MlpBufs b;
  BwdBufs bw;
  float* d_pre;
  int64_t bytes;
};
static void make_vf_plan(const vfnerf_mlp_desc& vf, int64_t n, int multires, int keep, int grid, void* ws, VfPlan& p) {
  Carver c(ws);
  p.emb = c.f(n * (3 + 6 * multires));
  p.pts = grid ? c.f(3 * n) : nullptr;
  carve_mlp(c, vf, n, p.b);
  if (keep) {
    vfnerf_mlp_desc none{};
    carve_bwd(c, vf, none, n, p.bw);
    p.d_pre = c.f(n * vf.out_dim[vf.n_layers - 1]);
  }
  p.bytes = c.off;
}

static int tc_precision(int precision, const char* who) {
  VFN_REQUIRE(precision == VFNERF_PREC_BF16 || precision == VFNERF_PREC_BF16X3 || precision == VFNERF_PREC_FP16F8,
              "%s: unknown precision %d", who, precision);
  return 0;
}

static int vf_tc_plan(const vfnerf_mlp_desc& vf, int multires, int skip_layer, void* ws, TcPlan& plan, int64_t& bytes,
                      int64_t n_points = 0, int keep = 0, int x3 = 0) {
  int64_t off = 0;
  if (int e = tc_carve(reinterpret_cast<char*>(ws), off, multires, 0, skip_layer, vf, nullptr, plan, n_points, keep, x3)) return e;
  bytes = off;
  return 0;
}

int64_t vfnerf_vf_workspace_bytes(const vfnerf_mlp_desc* vf, int64_t n_points, int multires,
                                  int keep_for_backward, int precision) {
  if (!vf) { set_error("null argument"); return -1; }
  if (precision != VFNERF_PREC_FP32) {
    if (tc_precision(precision, "vf_workspace_bytes")) return -1;
    TcPlan plan;
    int64_t bytes = 0;
    int skip = -1;
    for (int l = 1; l < vf->n_layers; ++l) if (vf->in_dim[l] != vf->out_dim[l - 1]) skip = l;
    if (vf_tc_plan(*vf, multires, skip, nullptr, plan, bytes, n_points, keep_for_backward, split_mode(precision))) return -1;
    return bytes + 1024;
  }
  VfPlan p;
  make_vf_plan(*vf, n_points, multires, keep_for_backward, 1, nullptr, p);
  return p.bytes + 256;
}

int vfnerf_vf_fwd(const vfnerf_mlp_desc* vf, const float* vf_arena, int multires, int skip_layer, float bn_eps,
                  int precision, const float* points, int64_t n_points, float* out, int64_t out_ld,
                  int n_out_cols, void* workspace, int64_t workspace_bytes, int keep_for_backward,
                  void* stream) {
  VFN_REQUIRE(vf && vf_arena && out, "vf_fwd: null argument");
  if (int e = validate_vf(*vf, multires, skip_layer)) return e;
  VFN_REQUIRE(n_out_cols >= 1 && n_out_cols <= vf->out_dim[vf->n_layers - 1], "vf_fwd: n_out_cols=%d invalid", n_out_cols);
  if (n_points == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(s);
  if (precision != VFNERF_PREC_FP32) {
    if (int e = tc_precision(precision, "vf_fwd")) return e;
    VFN_REQUIRE(n_out_cols == 3 || n_out_cols == vf->out_dim[vf->n_layers - 1], "vf_fwd(bf16): n_out_cols must be 3 or all");
    TcPlan plan;
    int64_t bytes = 0;
    if (int e = vf_tc_plan(*vf, multires, skip_layer, workspace, plan, bytes, n_points, keep_for_backward,
                           split_mode(precision))) return e;
    VFN_REQUIRE(workspace && workspace_bytes >= bytes, "vf_fwd: workspace too small");
    if (int e = tc_prepare(*vf, vf_arena, nullptr, nullptr, bn_eps, plan, s)) return e;
    const bool full = n_out_cols > 3;
    const int mode = keep_for_backward ? (full ? TC_MODE_VF_FULL_STASH : TC_MODE_V_ONLY_STASH)
                                       : (full ? TC_MODE_VF_FULL : TC_MODE_V_ONLY);
    return tc_forward(plan, mode, points, nullptr, 0, 0, n_points, nullptr, 0, out, out_ld, full ? out + 3 : nullptr, out_ld,
                      nullptr, s);
  }
  VfPlan p;
  make_vf_plan(*vf, n_points, multires, keep_for_backward, 1, workspace, p);
  VFN_REQUIRE(workspace && workspace_bytes >= p.bytes, "vf_fwd: workspace %lld B < required %lld B",
              (long long)workspace_bytes, (long long)p.bytes);
  if (int e = launch_fold_bn(*vf, vf_arena, bn_eps, p.b.scale, p.b.shift, s)) return e;
  if (int e = vf_embed_fp32(*vf, p.b, multires, skip_layer, points, n_points, p.emb, s)) return e;
  return vf_forward_fp32(*vf, vf_arena, p.b, skip_layer, p.emb, n_points, out, out_ld, n_out_cols, s);
}

int64_t vfnerf_mlp_points_workspace_bytes(const vfnerf_mlp_desc* vf, const vfnerf_mlp_desc* rn, int multires,
                                          int multires_view, int skip_layer) {
  if (!vf || !rn) { set_error("null argument"); return -1; }
  TcPlan plan;
  int64_t off = 0;
  if (tc_carve(nullptr, off, multires, multires_view, skip_layer, *vf, rn, plan, 0, 0, 1)) return -1;   // the larger (bf16x3) images
  return off + 1024;
}

int vfnerf_mlp_points_fwd(const vfnerf_mlp_desc* vf, const float* vf_arena, const vfnerf_mlp_desc* rn,
                          const float* rn_arena, int multires, int multires_view, int skip_layer, float bn_eps,
                          int precision, const float* points, const float* ray_dirs, int samples_per_ray,
                          int64_t n_points, float* normals, float* colors, void* workspace,
                          int64_t workspace_bytes, int repack, void* stream) {
  VFN_REQUIRE(vf && vf_arena && rn && rn_arena && points && ray_dirs && normals && colors, "mlp_points_fwd: null argument");
  VFN_REQUIRE(precision == VFNERF_PREC_BF16 || precision == VFNERF_PREC_BF16X3 || precision == VFNERF_PREC_FP16F8,
              "mlp_points_fwd: only the tensor-core paths (bf16, bf16x3, fp16f8) implement this entry");
  if (int e = validate_vf(*vf, multires, skip_layer)) return e;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(s);
  TcPlan plan;
  int64_t off = 0;
  if (int e = tc_carve(reinterpret_cast<char*>(workspace), off, multires, multires_view, skip_layer, *vf, rn, plan, 0, 0,
                       split_mode(precision))) return e;
  VFN_REQUIRE(workspace && workspace_bytes >= off, "mlp_points_fwd: workspace too small");
  if (repack)
    if (int e = tc_prepare(*vf, vf_arena, rn, rn_arena, bn_eps, plan, s)) return e;
  return tc_forward(plan, TC_MODE_RENDER, points, nullptr, 0, 0, n_points, ray_dirs, samples_per_ray, normals, 3,
                    nullptr, 0, colors, s);
}

int vfnerf_vf_bwd(const vfnerf_mlp_desc* vf, const float* vf_arena, int multires, int skip_layer, float bn_eps,
                  int precision, int64_t n_points, const float* out, int64_t out_ld, const float* d_out,
                  int64_t d_ld, int n_out_cols, float* vf_grad_arena, int accumulate, void* workspace,
                  int64_t workspace_bytes, void* stream) {
  VFN_REQUIRE(vf && vf_arena && out && d_out && vf_grad_arena, "vf_bwd: null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(s);
  if (precision != VFNERF_PREC_FP32) {
    // tensor-core path: the forward (keep_for_backward) left the activation stash and the transposed weight images
    VFN_REQUIRE(precision == VFNERF_PREC_BF16, "vf_bwd: precision bf16x3 is forward-only; train with bf16 or fp32");
    VFN_REQUIRE(n_out_cols == 3 || n_out_cols == vf->out_dim[vf->n_layers - 1], "vf_bwd(bf16): n_out_cols must be 3 or all");
    if (!accumulate) VFN_CHECK_CUDA(cudaMemsetAsync(vf_grad_arena, 0, sizeof(float) * vf->arena_floats, s));
    if (n_points == 0) return 0;
    TcPlan plan;
    int64_t bytes = 0;
    if (int e = vf_tc_plan(*vf, multires, skip_layer, workspace, plan, bytes, n_points, 1)) return e;
    VFN_REQUIRE(workspace && workspace_bytes >= bytes, "vf_bwd: workspace too small");
    return tc_backward_vf(plan, *vf, vf_arena, bn_eps, n_points, out, out_ld, d_out, d_ld, n_out_cols, vf_grad_arena,
                          accumulate, s);
  }
  const int Do = vf->out_dim[vf->n_layers - 1];
  VFN_REQUIRE(n_out_cols >= 1 && n_out_cols <= Do, "vf_bwd: n_out_cols invalid");
  VfPlan p;
  make_vf_plan(*vf, n_points, multires, 1, 1, workspace, p);
  VFN_REQUIRE(workspace && workspace_bytes >= p.bytes, "vf_bwd: workspace too small");
  if (!accumulate) VFN_CHECK_CUDA(cudaMemsetAsync(vf_grad_arena, 0, sizeof(float) * vf->arena_floats, s));
  if (n_points == 0) return 0;
  // d_pre = d_out * (1 - out^2) on the first n_out_cols columns, zero elsewhere
  if (n_out_cols < Do) VFN_CHECK_CUDA(cudaMemsetAsync(p.d_pre, 0, sizeof(float) * n_points * Do, s));
  if (int e = launch_act_bwd(out, out_ld, d_out, d_ld, n_points, n_out_cols, ACT_TANH, p.d_pre, Do, s)) return e;
  return mlp_backward_fp32(*vf, vf_arena, p.b, skip_layer, bn_eps, p.emb, 3 + 6 * multires, n_points, p.d_pre, Do,
                           p.bw, vf_grad_arena, 1, nullptr, 0, 0, 0, s);
}

int vfnerf_vf_grid_query(const vfnerf_mlp_desc* vf, const float* vf_arena, int multires, int skip_layer,
                         float bn_eps, int precision, int res, int64_t i0, int64_t n_points,
                         const float* origin3_host, const float* translation3_host,
                         const float* centroid3_host, float voxel, float* out, void* workspace,
                         int64_t workspace_bytes, void* stream) {
  VFN_REQUIRE(vf && vf_arena && out && origin3_host && translation3_host && centroid3_host, "grid_query: null argument");
  if (int e = validate_vf(*vf, multires, skip_layer)) return e;
  if (n_points == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(s);
  GridSpec gs;
  for (int c = 0; c < 3; ++c) { gs.origin[c] = origin3_host[c]; gs.translation[c] = translation3_host[c]; gs.centroid[c] = centroid3_host[c]; }
  gs.voxel = voxel;
  if (precision != VFNERF_PREC_FP32) {
    if (int e = tc_precision(precision, "grid_query")) return e;
    TcPlan plan;
    int64_t bytes = 0;
    if (int e = vf_tc_plan(*vf, multires, skip_layer, workspace, plan, bytes, 0, 0, split_mode(precision))) return e;
    VFN_REQUIRE(workspace && workspace_bytes >= bytes, "grid_query: workspace too small");
    if (int e = tc_prepare(*vf, vf_arena, nullptr, nullptr, bn_eps, plan, s)) return e;
    return tc_forward(plan, TC_MODE_V_ONLY, nullptr, &gs, res, i0, n_points, nullptr, 0, out, 3, nullptr, 0, nullptr, s);
  }
  VfPlan p;
  make_vf_plan(*vf, n_points, multires, 0, 1, workspace, p);
  VFN_REQUIRE(workspace && workspace_bytes >= p.bytes, "grid_query: workspace %lld B < required %lld B",
              (long long)workspace_bytes, (long long)p.bytes);
  if (int e = launch_grid_points(res, i0, n_points, gs, p.pts, s)) return e;
  if (int e = launch_fold_bn(*vf, vf_arena, bn_eps, p.b.scale, p.b.shift, s)) return e;
  if (int e = vf_embed_fp32(*vf, p.b, multires, skip_layer, p.pts, n_points, p.emb, s)) return e;
  return vf_forward_fp32(*vf, vf_arena, p.b, skip_layer, p.emb, n_points, out, 3, 3, s);
}

// ---- stage entry points ----------------------------------------------------------------------
int vfnerf_ray_geometry(int n_rays, int pose_is_quat, const float* uv, const float* pose,
                        const float* intrinsics, float* directions, float* ray_dirs, float* cam_loc,
                        void* stream) {
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  return launch_ray_geometry(n_rays, pose_is_quat, uv, pose, intrinsics, directions, ray_dirs, cam_loc,
                             reinterpret_cast<cudaStream_t>(stream));
}

int vfnerf_coarse_sample(int n_rays, int n_coarse, double near_, double far_, int perturb,
                         const float* t_vals, const float* U1, const float* directions,
                         const float* cam_loc, float* z, float* points, void* stream) {
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  return launch_coarse_sample(n_rays, n_coarse, near_, far_, perturb, t_vals, U1, directions, cam_loc, z, points,
                              reinterpret_cast<cudaStream_t>(stream));
}

int vfnerf_fine_sample(int n_rays, int n_coarse, int n_fine, double near_, double far_, double fine_range,
                       int perturb, const float* z_coarse, const float* w_coarse, const float* U2,
                       const float* U3, const float* directions, const float* cam_loc, float* z,
                       float* points, void* stream) {
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  return launch_fine_sample(n_rays, n_coarse, n_fine, near_, far_, fine_range, perturb, z_coarse, w_coarse, U2, U3,
                            nullptr, directions, cam_loc, z, points, nullptr, nullptr,
                            reinterpret_cast<cudaStream_t>(stream));
}

int vfnerf_mc_count(const float* pred, int resolution, uint8_t* keep, int32_t* cta_counts, float* div_raw,
                    uint8_t* choice, const uint8_t* surface, void* stream) {
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  return launch_mc_count(pred, resolution, keep, cta_counts, div_raw, choice, surface, reinterpret_cast<cudaStream_t>(stream));
}

int vfnerf_mc_emit(const float* pred, int resolution, const uint8_t* keep, const int64_t* cta_offsets, int32_t* cells,
                   float* comb, float* udf, void* stream) {
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  return launch_mc_emit(pred, resolution, keep, cta_offsets, cells, comb, udf, reinterpret_cast<cudaStream_t>(stream));
}

int vfnerf_smooth_vf(const float* in, float* tmp, float* out, int resolution, int kernel_size, const float* taps_host,
                     void* stream) {
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  return launch_smooth_vf(in, tmp, out, resolution, kernel_size, taps_host, reinterpret_cast<cudaStream_t>(stream));
}

int vfnerf_sample_pdf(int n_rays, int n_bins, int n_samples, const float* bins, const float* weights, const float* u,
                      int u_per_ray, float* samples, void* stream) {
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  return launch_sample_pdf(n_rays, n_bins, n_samples, bins, weights, u, u_per_ray, samples,
                           reinterpret_cast<cudaStream_t>(stream));
}

int vfnerf_pdf_fine_sample(int n_rays, int n_coarse, int n_fine, const float* z_coarse, const float* w_coarse,
                           const float* u, int u_per_ray, const float* directions, const float* cam_loc, float* z,
                           float* points, void* stream) {
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  return launch_pdf_fine_sample(n_rays, n_coarse, n_fine, z_coarse, w_coarse, u, u_per_ray, directions, cam_loc, z,
                                points, reinterpret_cast<cudaStream_t>(stream));
}

int vfnerf_density_weights(const vfnerf_render_cfg* cfg, int n_samples, const float* density_params,
                           const float* normals, int64_t normals_ld, const float* ray_dirs, const float* z,
                           float* cosw, float* sigma, float* weights, void* stream) {
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  VFN_REQUIRE(cfg, "density_weights: null cfg");
  return launch_density_weights(*cfg, cfg->n_rays, n_samples, density_params, normals, normals_ld, ray_dirs, z,
                                cosw, sigma, weights, reinterpret_cast<cudaStream_t>(stream));
}

int vfnerf_volume_weights(int n_rays, int n_samples, int mode, int normalize, const float* sigma, const float* z,
                          float* weights, void* stream) {
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  return launch_volume_weights(n_rays, n_samples, mode, normalize, sigma, z, weights, reinterpret_cast<cudaStream_t>(stream));
}

int vfnerf_composite(int n_rays, int n_samples, const float* weights, const float* colors, const float* z,
                     float* rgb, float* depth, void* stream) {
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  return launch_composite(n_rays, n_samples, weights, colors, z, rgb, depth, reinterpret_cast<cudaStream_t>(stream));
}

int vfnerf_composite_white(int n_rays, int n_samples, const float* weights, const float* colors, const float* z,
                           float* rgb, float* depth, void* stream) {
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  return launch_composite(n_rays, n_samples, weights, colors, z, rgb, depth, reinterpret_cast<cudaStream_t>(stream), 1);
}

}  // extern "C"
