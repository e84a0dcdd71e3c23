// lgk_policy.cu -- rsl_rl ActorCritic.act / evaluate / get_actions_log_prob (rsl_rl is not vendored in the
// reference; call sites utils/task_registry.py:37-38,154; restated API in SURVEY App. C.2):
//   actor  = Linear(O,h0) ELU Linear(h0,h1) ELU Linear(h1,h2) ELU Linear(h2,A)
//   critic = Linear(Oc,h0) ELU ... Linear(h2,1);  a ~ Normal(mu, std);  logp = sum log N(a; mu, std)
//
// lgk_policy_act dispatches to the tcgen05 kernel (lgk_policy_tc.cu) whenever the shape fits it.  This file holds the
// FP32 path for every other shape: a 64x64x16 shared-memory FFMA GEMM with fused bias+ELU epilogue per layer and a
// sampling/log-prob epilogue kernel; hidden activations go through the caller-provided workspace.
#include "lgk_policy_common.cuh"
#include "lgk_policy_tc_plan.h"

namespace lgk {

constexpr int BM = 64, BN = 64, BK = 16;

// Y[M, Nout] = act(X[M, K] @ W[Nout, K]^T + b)   (nn.Linear layout)
template <bool ELU>
__global__ void __launch_bounds__(256) gemm_bias_act(const float* __restrict__ X, const float* __restrict__ W,
                                                     const float* __restrict__ b, float* __restrict__ Y, int M, int K,
                                                     int Nout) {
  __shared__ float sx[BK][BM + 1], sw[BK][BN + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int i = threadIdx.x; i < BM * BK; i += 256) {
      const int r = i / BK, c = i % BK;
      const int m = m0 + r, k = k0 + c;
      sx[c][r] = (m < M && k < K) ? X[(size_t)m * K + k] : 0.f;
      const int n = n0 + r;
      sw[c][r] = (n < Nout && k < K) ? W[(size_t)n * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sx[k][ty * 4 + i]; w[i] = sw[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= Nout) continue;
      float v = acc[i][j] + b[n];
      if (ELU) v = v > 0.f ? v : expm1f(v);
      Y[(size_t)m * Nout + n] = v;
    }
  }
}

// actions = mu + std*eps (Box-Muller on the ACT stream), log_prob = sum_a log N(a; mu, std)
__global__ void sample_kernel(const __grid_constant__ LgkPolicyParams p) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.num_envs) return;
  const int A = p.num_actions;
  const RngKey key = make_key(p.seed, p.step);
  const uint32_t genv = (uint32_t)(p.env_id_offset + n);
  float logp = 0.f;
  for (int pr = 0; pr < (A + 1) / 2; ++pr) {
    float z0 = 0.f, z1 = 0.f;
    if (p.sample) {
      const U4 r = rng_block(key, genv, LGK_STREAM_ACT, (uint32_t)(pr >> 1));
      const uint32_t wa = (pr & 1) ? r.z : r.x, wb = (pr & 1) ? r.w : r.y;
      const float u1 = 1.0f - u32_to_uniform(wa), u2 = u32_to_uniform(wb);
      const float rad = sqrtf(-2.0f * logf(u1)), th = 6.283185307179586f * u2;
      z0 = rad * cosf(th); z1 = rad * sinf(th);
    }
    for (int h = 0; h < 2; ++h) {
      const int a = 2 * pr + h;
      if (a >= A) break;
      const float mu = p.action_mean[(size_t)n * A + a], sd = p.std[a];
      const float act = mu + sd * (h ? z1 : z0);
      p.actions[(size_t)n * A + a] = act;
      p.action_sigma[(size_t)n * A + a] = sd;
      const float d = act - mu;
      logp += -(d * d) / (2.0f * sd * sd) - logf(sd) - 0.9189385332046727f;   // log(sqrt(2*pi))
    }
  }
  p.actions_log_prob[n] = logp;
}

}  // namespace lgk

using namespace lgk;

static int g_variant = 0;

extern "C" int lgk_policy_set_variant(int variant) {
  const int prev = g_variant;
  if (variant >= 0 && variant <= 2) g_variant = variant;
  return prev;
}

extern "C" int lgk_policy_debug_timeline(int64_t* device_buf16, int flags) {
  policy_tc_set_timeline(reinterpret_cast<long long*>(device_buf16), flags);
  return LGK_OK;
}

static int64_t fp32_workspace_bytes(const LgkPolicyParams* p) {
  const int64_t hmax = p->hidden[0] > p->hidden[1] ? (p->hidden[0] > p->hidden[2] ? p->hidden[0] : p->hidden[2])
                                                    : (p->hidden[1] > p->hidden[2] ? p->hidden[1] : p->hidden[2]);
  return (int64_t)2 * p->num_envs * hmax * (int64_t)sizeof(float);
}

// large enough for either path, so that switching variants never needs a new workspace
extern "C" int64_t lgk_policy_workspace_bytes(const LgkPolicyParams* p) {
  if (!p) return -1;
  int64_t need = fp32_workspace_bytes(p);
  TcPlan pl;
  if (policy_tc_plan(p, &pl)) { const int64_t t = policy_tc_workspace_bytes(pl); if (t > need) need = t; }
  return need;
}

template <bool ELU>
static void launch_gemm(const float* X, const float* W, const float* b, float* Y, int M, int K, int Nout, cudaStream_t st) {
  dim3 grid((Nout + BN - 1) / BN, (M + BM - 1) / BM);
  gemm_bias_act<ELU><<<grid, 256, 0, st>>>(X, W, b, Y, M, K, Nout);
  count_launch();
}

extern "C" int lgk_policy_act(const LgkPolicyParams* p, void* stream) {
  LGK_REQUIRE(p != nullptr && p->num_envs > 0, "policy: bad params");
  LGK_REQUIRE(p->nets >= 0 && p->nets <= 3, "policy: nets must be 0..3");
  const bool run_actor = p->nets != 2, run_critic = p->nets != 1;
  LGK_REQUIRE(p->workspace != nullptr, "policy: null buffer");
  if (run_actor) LGK_REQUIRE(p->obs && p->std && p->actions && p->action_mean && p->action_sigma && p->actions_log_prob, "policy: null buffer");
  if (run_critic) LGK_REQUIRE(p->critic_obs && p->values, "policy: null buffer");
  for (int i = 0; i < 4; ++i) LGK_REQUIRE(p->actor_w[i] && p->actor_b[i] && p->critic_w[i] && p->critic_b[i], "policy: null weights");
  LGK_REQUIRE(p->workspace_bytes >= lgk_policy_workspace_bytes(p), "policy: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = p->num_envs;
  {
    TcPlan pl;
    const bool fits = policy_tc_plan(p, &pl);
    LGK_REQUIRE(g_variant != 2 || fits, "policy: shape does not fit the tcgen05 kernel");
    if (fits && g_variant != 1) return policy_tc_launch(p, pl, st);
  }
  const int64_t hmax = fp32_workspace_bytes(p) / (2 * (int64_t)N * (int64_t)sizeof(float));
  float* w0 = reinterpret_cast<float*>(p->workspace);
  float* w1 = w0 + (size_t)N * hmax;
  for (int net = 0; net < 2; ++net) {
    if (net == 0 ? !run_actor : !run_critic) continue;
    const float* const* W = net ? p->critic_w : p->actor_w;
    const float* const* B = net ? p->critic_b : p->actor_b;
    const float* x = net ? p->critic_obs : p->obs;
    const int in = net ? p->num_critic_obs : p->num_obs;
    launch_gemm<true>(x, W[0], B[0], w0, N, in, p->hidden[0], st);
    launch_gemm<true>(w0, W[1], B[1], w1, N, p->hidden[0], p->hidden[1], st);
    launch_gemm<true>(w1, W[2], B[2], w0, N, p->hidden[1], p->hidden[2], st);
    launch_gemm<false>(w0, W[3], B[3], net ? p->values : p->action_mean, N, p->hidden[2], net ? 1 : p->num_actions, st);
  }
  if (run_actor) {
    sample_kernel<<<(N + 127) / 128, 128, 0, st>>>(*p);
    count_launch();
  }
  return check_cuda(cudaGetLastError(), "policy kernels launch");
}
