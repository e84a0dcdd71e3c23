// lgk_policy_common.cuh -- pieces shared by the two ActorCritic.act paths (tcgen05 in lgk_policy_tc.cu, FP32 in
// lgk_policy.cu): the distribution epilogue and the host-side plan of the tensor-core kernel.
#pragma once
#include "lgk_math.cuh"
#include <cstring>

namespace lgk {

// PPO.act's distribution step for one environment (rsl_rl ActorCritic.act + get_actions_log_prob):
//   actions = mu + std * eps,  eps ~ N(0,1) by Box-Muller on the Philox ACT stream (pair pr uses words 2(pr&1), 2(pr&1)+1
//   of block pr>>1);  log_prob = sum_a log N(a; mu, std).  `mu` lives in registers (compile-time indexed).
// FAST selects the MUFU-based log / sin / cos / sqrt (absolute error ~1e-6 on a N(0,1) draw, inside the 1e-3 bar).
template <int MAXA, bool FAST>
__device__ __forceinline__ void policy_finish_row(const LgkPolicyParams& p, int n, const float (&mu)[MAXA],
                                                  const float* __restrict__ sd_arr) {
  const int A = p.num_actions;
  const RngKey key = make_key(p.seed, p.step);
  const uint32_t genv = (uint32_t)(p.env_id_offset + n);
  float logp = 0.f;
#pragma unroll
  for (int pr = 0; pr < MAXA / 2; ++pr) {
    if (2 * pr < A) {
      float z0 = 0.f, z1 = 0.f;
      if (p.sample) {
        const U4 r = rng_block(key, genv, LGK_STREAM_ACT, (uint32_t)(pr >> 1));
        const uint32_t wa = (pr & 1) ? r.z : r.x, wb = (pr & 1) ? r.w : r.y;
        const float u1 = 1.0f - u32_to_uniform(wa), u2 = u32_to_uniform(wb);
        const float th = 6.283185307179586f * u2;
        if (FAST) {
          const float rad = __fsqrt_rn(-2.0f * __logf(u1));
          float sn, cs;
          __sincosf(th, &sn, &cs);
          z0 = rad * cs; z1 = rad * sn;
        } else {
          const float rad = sqrtf(-2.0f * logf(u1));
          z0 = rad * cosf(th); z1 = rad * sinf(th);
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int a = 2 * pr + h;
        if (a < A) {
          const float m = mu[a], sd = sd_arr[a];
          const float act = m + sd * (h ? z1 : z0);
          p.action_mean[(size_t)n * A + a] = m;
          p.actions[(size_t)n * A + a] = act;
          p.action_sigma[(size_t)n * A + a] = sd;
          const float d = act - m;
          logp += -(d * d) / (2.0f * sd * sd) - (FAST ? __logf(sd) : logf(sd)) - 0.9189385332046727f;   // log(sqrt(2*pi))
        }
      }
    }
  }
  p.actions_log_prob[n] = logp;
}

// The same distribution step for the four actions 4*blk .. 4*blk+3 of one environment (= the two Box-Muller pairs of Philox
// block `blk` of the ACT stream): lets four warps share a row's epilogue.  Returns the partial log-prob of these actions.
// sd3 = [std | 1/std | log(std)] x 16 (precomputed once per CTA): log N(a; mu, sd) = -((a - mu)/sd)^2 / 2 - log sd - log sqrt(2 pi)
// without a division or a logarithm per action.
// the four N(0,1) draws of actions 4*blk .. 4*blk+3 (independent of the network's output: the tensor-core kernel draws them
// while the last MMAs are still running)
__device__ __forceinline__ void policy_draw_quad(const LgkPolicyParams& p, int n, int blk, float (&z)[4]) {
  z[0] = z[1] = z[2] = z[3] = 0.f;
  if (p.sample) {
    const RngKey key = make_key(p.seed, p.step);
    const U4 r = rng_block(key, (uint32_t)(p.env_id_offset + n), LGK_STREAM_ACT, (uint32_t)blk);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u1 = 1.0f - u32_to_uniform(w[2 * h]), u2 = u32_to_uniform(w[2 * h + 1]);
      const float th = 6.283185307179586f * u2;
      const float rad = __fsqrt_rn(-2.0f * __logf(u1));
      float sn, cs;
      __sincosf(th, &sn, &cs);
      z[2 * h] = rad * cs; z[2 * h + 1] = rad * sn;
    }
  }
}

__device__ __forceinline__ float policy_finish_quad(const LgkPolicyParams& p, int n, int blk, const float (&mu)[4],
                                                    const float* __restrict__ sd3, const float (&z)[4]) {
  const int A = p.num_actions;
  float logp = 0.f;
  float act[4], sd[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int a = 4 * blk + j;
    sd[j] = sd3[a];
    act[j] = mu[j] + sd[j] * z[j];
    if (a < A) {
      const float t = (act[j] - mu[j]) * sd3[16 + a];
      logp += -0.5f * (t * t) - sd3[32 + a] - 0.9189385332046727f;     // log(sqrt(2*pi))
    }
  }
  const size_t o = (size_t)n * A + 4 * blk;
  if ((A & 3) == 0) {      // rows are 16-byte aligned: one vector store per output
    *reinterpret_cast<float4*>(p.action_mean + o) = make_float4(mu[0], mu[1], mu[2], mu[3]);
    *reinterpret_cast<float4*>(p.actions + o) = make_float4(act[0], act[1], act[2], act[3]);
    *reinterpret_cast<float4*>(p.action_sigma + o) = make_float4(sd[0], sd[1], sd[2], sd[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * blk + j < A) { p.action_mean[o + j] = mu[j]; p.actions[o + j] = act[j]; p.action_sigma[o + j] = sd[j]; }
  }
  return logp;
}

struct TcPlan;
bool policy_tc_plan(const LgkPolicyParams* p, TcPlan* pl);
long long policy_tc_workspace_bytes(const TcPlan& pl);
int policy_tc_launch(const LgkPolicyParams* p, const TcPlan& pl, cudaStream_t st);
void policy_tc_set_timeline(long long* dev, int flags);

}  // namespace lgk
