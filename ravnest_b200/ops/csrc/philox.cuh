// Counter-based RNG for replayable dropout (masks must be identical in forward, recompute and
// backward).  Philox4x32-10 keyed by a 64-bit seed; the counter is (element group, call step) so a mask depends only on
// (seed, step, position) — both live in a 2 x int64 device tensor, which keeps CUDA-graph replays advancing correctly.
#pragma once
#include <stdint.h>

namespace rvn {

__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 7; ++r) {             // Philox4x32-7 (still passes BigCrush; 30% fewer multiplies than -10)
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

struct DropoutRng {
  uint32_t k0, k1, step_lo, step_hi;
  uint32_t threshold;        // drop when r < threshold  (threshold = p * 2^32)
  float inv_keep;
  __device__ __forceinline__ void init(const long long* state, float p) {
    const unsigned long long seed = static_cast<unsigned long long>(state[0]);
    const unsigned long long step = static_cast<unsigned long long>(state[1]);
    k0 = static_cast<uint32_t>(seed); k1 = static_cast<uint32_t>(seed >> 32);
    step_lo = static_cast<uint32_t>(step); step_hi = static_cast<uint32_t>(step >> 32);
    threshold = static_cast<uint32_t>(fminf(p, 0.99999f) * 4294967296.0f);
    inv_keep = 1.0f / (1.0f - p);
  }
  // eight keep-scales from ONE Philox call (16 random bits per element) for the 8 consecutive elements of group ``g8``;
  // uses a different counter word than scales4 so the two group spaces never collide
  __device__ __forceinline__ void scales8(unsigned long long g8, float (&s)[8]) const {
    const uint4 r = philox4x32(static_cast<uint32_t>(g8), static_cast<uint32_t>(g8 >> 32), step_lo, step_hi ^ 0x80000000u, k0, k1);
    const uint32_t t16 = threshold >> 16;
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s[2 * i] = (w[i] & 0xFFFFu) >= t16 ? inv_keep : 0.f;
      s[2 * i + 1] = (w[i] >> 16) >= t16 ? inv_keep : 0.f;
    }
  }
  // four keep-scales for the 4 consecutive elements of group ``g`` (64-bit group index)
  __device__ __forceinline__ void scales4(unsigned long long g, float (&s)[4]) const {
    const uint4 r = philox4x32(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), step_lo, step_hi, k0, k1);
    s[0] = r.x >= threshold ? inv_keep : 0.f;
    s[1] = r.y >= threshold ? inv_keep : 0.f;
    s[2] = r.z >= threshold ? inv_keep : 0.f;
    s[3] = r.w >= threshold ? inv_keep : 0.f;
  }
};

}  // namespace rvn
