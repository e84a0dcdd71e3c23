// lgk_rng.cuh -- Philox4x32-10 counter-based RNG (Salmon et al. SC'11; Random123 constants).
// The reference draws from torch's global generator in a data-dependent order (LR:353-366, 405, 425,
// 431, 442, 465, 230) that no kernel can mirror; the product defines its own stream instead:
//   key = (seed_lo, seed_hi), counter = (global_env_id, word/4, stream, step)
// and oracle/philox.py restates it bit-exactly so that oracle and kernels consume identical numbers.
#pragma once
#include "lgk_common.cuh"

namespace lgk {

struct U4 { uint32_t x, y, z, w; };

LGK_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

LGK_HD U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return U4{c0, c1, c2, c3};
}

LGK_HD float u32_to_uniform(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-08f; }  // 2^-24, exact

struct RngKey {
  uint32_t k0, k1, step;
};

LGK_HD RngKey make_key(uint64_t seed, int32_t step) {
  return RngKey{(uint32_t)(seed & 0xFFFFFFFFull), (uint32_t)(seed >> 32), (uint32_t)step};
}

// block `blk` (4 words) of `stream` for global env id `env`
LGK_HD U4 rng_block(const RngKey& k, uint32_t env, uint32_t stream, uint32_t blk) {
  return philox4x32_10(env, blk, stream, k.step, k.k0, k.k1);
}

LGK_HD uint32_t pick(const U4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

}  // namespace lgk
