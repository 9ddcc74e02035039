// Host-side helper (no device code): the reference draws its sampler uniforms with torch.rand on the global CPU generator
// (ray_sampler.py:138,292,297) -- 52 M numbers per 1200x680 image, drawn even when only a handful are used.  Reproducing
// that stream bit for bit is part of the contract (sample positions), and once the kernels render an image in 170 ms the
// scalar generator behind torch.rand (60 ms per image) bounds the end-to-end rate.  This is the same generator -- MT19937
// with torch's tempering and its 24-bit float conversion (ATen/core/MT19937RNGEngine.h, ATen/core/TransformationHelper.h:
// uniform_real<float>) -- written so that the compiler vectorises the block regeneration, the tempering and the
// conversion.  The caller (vfnerf_b200/nerf.py) hands in the engine fields of torch.get_rng_state() and writes them back,
// so torch's generator continues exactly where it would have; a self-check against torch.rand guards the state layout.
#include <cstdint>

#include "common.cuh"

namespace {
constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;

inline uint32_t twist(uint32_t u, uint32_t v) { return (((u & kUpper) | (v & kLower)) >> 1) ^ ((v & 1u) ? kMatrixA : 0u); }

// next_state() of at::mt19937: regenerates the 624 words in place
void next_state(uint32_t* s) {
  uint32_t t[kN];
  for (int i = 0; i < kN - kM; ++i) t[i] = s[i + kM] ^ twist(s[i], s[i + 1]);                 // 0 .. 226: old words only
  for (int i = kN - kM; i < 2 * (kN - kM); ++i) t[i] = t[i - (kN - kM)] ^ twist(s[i], s[i + 1]);   // 227 .. 453
  for (int i = 2 * (kN - kM); i < kN - 1; ++i) t[i] = t[i - (kN - kM)] ^ twist(s[i], s[i + 1]);    // 454 .. 622
  t[kN - 1] = t[kM - 1] ^ twist(s[kN - 1], t[0]);
  for (int i = 0; i < kN; ++i) s[i] = t[i];
}

inline float temper_to_float(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return (float)(y & 0xFFFFFFu) * 5.9604644775390625e-08f;     // (y & (2^24 - 1)) * 2^-24, exact in fp32
}
}  // namespace

extern "C" int vfnerf_mt19937_uniform(uint32_t* state624, int32_t* left, uint32_t* next, int64_t n, float* out) {
  if (!state624 || !left || !next || (n > 0 && !out)) { vfn::set_error("mt19937_uniform: null argument"); return 1; }
  int64_t i = 0;
  int l = *left;
  uint32_t nx = *next;
  while (i < n) {
    // at::mt19937::operator(): if (--left == 0) next_state();  y = state[next++]
    if (l == 1) { next_state(state624); l = kN + 1; nx = 0; }      // the decrement below brings it to 624 - consumed
    const int64_t avail = l - 1;                                   // numbers obtainable before the next regeneration
    const int64_t take = (n - i < avail) ? (n - i) : avail;
    const uint32_t* src = state624 + nx;
    float* dst = out + i;
    for (int64_t k = 0; k < take; ++k) dst[k] = temper_to_float(src[k]);
    i += take; nx += (uint32_t)take; l -= (int)take;
  }
  *left = l; *next = nx;
  return 0;
}
