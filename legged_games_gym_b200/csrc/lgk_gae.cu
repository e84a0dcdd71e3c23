// lgk_gae.cu -- rsl_rl RolloutStorage.compute_returns (not vendored in the reference; call site
// PPO.compute_returns): reverse GAE scan over the [T,N] rollout + global advantage
// normalisation (mean, unbiased std + 1e-8).
//
// Pass 1: one thread per env walks t = T-1..0 (loads coalesced across envs), writes returns and the raw
//         advantages, and reduces sum / sum-of-squares in double (warp shuffle -> one atomicAdd pair per CTA).
// Pass 2: grid-stride normalisation of the [T,N] advantages (they are still L2-resident: T*N*4 B << 126 MB).
#include "lgk_common.cuh"

namespace lgk {

__global__ void __launch_bounds__(256) gae_scan_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                      const uint8_t* __restrict__ dones, const float* __restrict__ last_values,
                                                      int T, int N, float gamma, float lam, float* __restrict__ returns,
                                                      float* __restrict__ adv, double* __restrict__ acc) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0, s2 = 0.0;
  if (n < N) {
    float next_v = last_values[n];
    float a = 0.f;
    for (int t = T - 1; t >= 0; --t) {
      const size_t i = (size_t)t * N + n;
      const float v = values[i];
      const float nt = 1.0f - (float)dones[i];
      const float delta = rewards[i] + nt * gamma * next_v - v;
      a = delta + nt * gamma * lam * a;
      const float ret = a + v;
      returns[i] = ret;
      const float ad = ret - v;               // advantages = returns - values
      adv[i] = ad;
      s += ad; s2 += (double)ad * ad;
      next_v = v;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  __shared__ double sh[2][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[0][w] = s; sh[1][w] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < 8; ++i) { a += sh[0][i]; b += sh[1][i]; }
    atomicAdd(acc, a); atomicAdd(acc + 1, b);
  }
}

__global__ void __launch_bounds__(256) gae_norm_kernel(float* __restrict__ adv, long long total, const double* __restrict__ acc) {
  const double mean = acc[0] / (double)total;
  double var = (acc[1] - (double)total * mean * mean) / (double)(total - 1);   // unbiased (torch .std())
  var = var > 0 ? var : 0;
  const float m = (float)mean, inv = 1.0f / ((float)sqrt(var) + 1e-8f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    adv[i] = (adv[i] - m) * inv;
}

}  // namespace lgk

using namespace lgk;

extern "C" int lgk_gae(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                       int32_t T, int32_t N, float gamma, float lam, float* returns, float* advantages,
                       double* scratch, void* stream) {
  LGK_REQUIRE(rewards && values && dones && last_values && returns && advantages && scratch, "gae: null pointer");
  LGK_REQUIRE(T > 0 && N > 0 && (long long)T * N > 1, "gae: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = check_cuda(cudaMemsetAsync(scratch, 0, 4 * sizeof(double), st), "gae memset")) return rc;
  gae_scan_kernel<<<(N + 255) / 256, 256, 0, st>>>(rewards, values, dones, last_values, T, N, gamma, lam, returns, advantages, scratch);
  const long long total = (long long)T * N;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  gae_norm_kernel<<<blocks, 256, 0, st>>>(advantages, total, scratch);
  count_launch(2);
  return check_cuda(cudaGetLastError(), "gae kernels launch");
}

// ------------------------------------------------------------------ rsl_rl OnPolicyRunner: episode bookkeeping of a rollout step
// cur_return += reward; cur_length += 1; for every env whose episode ended: stats += (cur_return, cur_length, 1) and both
// running values restart at 0 (rsl_rl on_policy_runner.py: cur_reward_sum / cur_episode_length / rewbuffer / lenbuffer,
// with the finished episodes reduced to three sums that stay on the device).  One launch instead of six torch ops.
namespace lgk {
__global__ void __launch_bounds__(256) episode_stats_kernel(const float* __restrict__ rewards, const uint8_t* __restrict__ dones,
                                                            float* __restrict__ cur_return, float* __restrict__ cur_length,
                                                            double* __restrict__ stats, int n) {
  __shared__ double s_part[3][8];
  double ret = 0.0, len = 0.0, cnt = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float r = cur_return[i] + rewards[i], l = cur_length[i] + 1.0f;
    const bool d = dones[i] != 0;
    if (d) { ret += (double)r; len += (double)l; cnt += 1.0; }
    cur_return[i] = d ? 0.f : r;
    cur_length[i] = d ? 0.f : l;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ret += __shfl_xor_sync(0xffffffffu, ret, o); len += __shfl_xor_sync(0xffffffffu, len, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_part[0][warp] = ret; s_part[1][warp] = len; s_part[2][warp] = cnt; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_part[threadIdx.x][w];
    if (t != 0.0) atomicAdd(stats + threadIdx.x, t);
  }
}
}  // namespace lgk

extern "C" int lgk_episode_stats(const float* rewards, const uint8_t* dones, float* cur_return, float* cur_length,
                                 double* stats3, int32_t n, void* stream) {
  LGK_REQUIRE(rewards && dones && cur_return && cur_length && stats3 && n > 0, "episode_stats: bad arguments");
  const int blocks = (n + 255) / 256 < 296 ? (n + 255) / 256 : 296;
  lgk::episode_stats_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rewards, dones, cur_return, cur_length, stats3, n);
  lgk::count_launch();
  return lgk::check_cuda(cudaGetLastError(), "episode_stats_kernel launch");
}

