// Host-side planning helpers shared by the orchestration files (api.cu: eval-mode render / VF query; mlp_train.cu:
// train-mode path): workspace carving and the fp32 layer-wise buffers of one MLP.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace vfn {

// ---------------------------------------------------------------------------------------------
// workspace carving
// ---------------------------------------------------------------------------------------------
struct Carver {
  char* base;
  int64_t off = 0;
  explicit Carver(void* p) : base(reinterpret_cast<char*>(p)) {}
  float* f(int64_t n_floats) {
    float* r = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += align_up(n_floats * (int64_t)sizeof(float), 256);
    return r;
  }
};

inline int sum_out(const vfnerf_mlp_desc& d) {
  int s = 0;
  for (int l = 0; l < d.n_layers; ++l) s += d.out_dim[l];
  return s;
}
inline int max_dim(const vfnerf_mlp_desc& d) {
  int m = 0;
  for (int l = 0; l < d.n_layers; ++l) m = std::max(m, std::max(d.in_dim[l], d.out_dim[l]));
  return m;
}
inline int64_t max_wsize(const vfnerf_mlp_desc& d) {
  int64_t m = 0;
  for (int l = 0; l < d.n_layers; ++l) m = std::max<int64_t>(m, (int64_t)d.in_dim[l] * d.out_dim[l]);
  return m;
}

// fp32-path buffers of one MLP evaluated on n points: the (post-activation) output of every hidden
// layer l lives in act[l] with row stride in_dim[l+1] (so a skip layer finds its concatenated input
// in place); `scale`/`shift` hold the folded BatchNorm affine of all layers.
struct MlpBufs {
  float* act[VFNERF_MAX_LAYERS];
  float* scale;
  float* shift;
  int soff[VFNERF_MAX_LAYERS];
};
inline void carve_mlp(Carver& c, const vfnerf_mlp_desc& d, int64_t n, MlpBufs& b) {
  int so = 0;
  for (int l = 0; l < d.n_layers; ++l) { b.soff[l] = so; so += d.out_dim[l]; }
  b.scale = c.f(so);
  b.shift = c.f(so);
  for (int l = 0; l + 1 < d.n_layers; ++l) b.act[l] = c.f(n * d.in_dim[l + 1]);
}

struct BwdBufs {
  float *dA, *dB, *G, *colsum;
};
inline void carve_bwd(Carver& c, const vfnerf_mlp_desc& vf, const vfnerf_mlp_desc& rn, int64_t n, BwdBufs& b) {
  int w = std::max(max_dim(vf), max_dim(rn));
  b.dA = c.f(n * w);
  b.dB = c.f(n * w);
  b.G = c.f(std::max(max_wsize(vf), max_wsize(rn)));
  b.colsum = c.f(w);
}

inline int validate_vf(const vfnerf_mlp_desc& vf, int multires, int skip_layer) {
  const int E = 3 + 6 * multires;
  VFN_REQUIRE(vf.n_layers >= 2 && vf.n_layers <= VFNERF_MAX_LAYERS, "VF net: n_layers=%d unsupported", vf.n_layers);
  VFN_REQUIRE(vf.in_dim[0] == E, "VF net: in_dim[0]=%d but embedding width is %d", vf.in_dim[0], E);
  for (int l = 1; l < vf.n_layers; ++l) {
    int want = vf.out_dim[l - 1] + (l == skip_layer ? E : 0);
    VFN_REQUIRE(vf.in_dim[l] == want, "VF net: in_dim[%d]=%d, expected %d", l, vf.in_dim[l], want);
  }
  VFN_REQUIRE(skip_layer != 0 && skip_layer < vf.n_layers, "VF net: skip_layer=%d unsupported", skip_layer);
  VFN_REQUIRE(vf.out_dim[vf.n_layers - 1] >= 3, "VF net: output narrower than 3");
  return 0;
}


static const float kSqrt2 = 1.41421354f;   // torch.sqrt(torch.tensor([2.]).float()), vector_field_network.py:193

}  // namespace vfn
