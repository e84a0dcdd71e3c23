// lgk_torque.cu -- _compute_torques: PD law (reference LR:371-395) and the ANYdrive SEA actuator network
// (ANY:71-81 + resources/actuator_nets/anydrive_v3_lstm.pt: in_scale -> LSTM(2,8,2 layers) -> Linear(8,1) -> out_scale).
//
// One thread per (env, joint) sequence.  The 969 LSTM weights live in __constant__ memory and every weight
// reference is a compile-time offset, so each MAC is one FFMA with a constant-bank operand (no weight loads);
// the 32 floats of h/c state per sequence stream through registers as 128-bit loads/stores.
// Gate nonlinearities use MUFU ex2/rcp with the products sigma(i)*tanh(g) and sigma(o)*tanh(c') sharing one
// reciprocal each (8 MUFU per hidden unit instead of 10) -- the MUFU pipe, not HBM, is the next limiter.
#include "lgk_math.cuh"

namespace lgk {

// Device-side weight image.  lgk_set_lstm_weights folds every constant factor into the weights once:
//   * gate rows are pre-multiplied by -log2(e) (i, f, o) or -2*log2(e) (g) and the two bias vectors summed, so a
//     gate pre-activation IS the ex2 argument of its sigmoid / tanh;
//   * in_scale is folded into w_ih0, out_scale into the output layer;
//   * matrices are stored INPUT-major ([k][gate row]) so that two adjacent gate rows form one 64-bit operand
//     of the Blackwell packed-fp32 FMA (fma.rn.f32x2 -> SASS FFMA2 with a uniform-register weight pair).
struct LstmDev {
  float w_ih0[2 * 32], w_hh0[8 * 32], b0[32];
  float w_ih1[8 * 32], w_hh1[8 * 32], b1[32];
  float lin_w[8], lin_b;
};
__constant__ LstmDev c_lstm;

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kTanhClamp = 43.0f;   // = 2*log2(e)*14.9: tanh saturates in fp32 long before, keeps 1 - 2^x finite

// one LSTM layer, hidden 8; IN = input width.  torch gate order i, f, g, o (rows 0-7, 8-15, 16-23, 24-31).
//   c' = sigmoid(f)*c + sigmoid(i)*tanh(g) ; h' = sigmoid(o)*tanh(c')
// with sigmoid(a)*tanh(b) = (1 - eb) / ((1 + ea)(1 + eb)), ea = e^-a, eb = e^-2b: 5 ex2 + 3 rcp per unit.
template <int IN>
__device__ __forceinline__ void lstm_layer(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                           const float* __restrict__ bias, const float (&x)[IN], float (&h)[8],
                                           float (&c)[8]) {
  unsigned long long acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = pack2(bias[2 * j], bias[2 * j + 1]);
#pragma unroll
  for (int k = 0; k < IN; ++k) {
    const unsigned long long xx = pack2(x[k], x[k]);
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = fma2(pack2(w_ih[k * 32 + 2 * j], w_ih[k * 32 + 2 * j + 1]), xx, acc[j]);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const unsigned long long hh = pack2(h[k], h[k]);
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = fma2(pack2(w_hh[k * 32 + 2 * j], w_hh[k * 32 + 2 * j + 1]), hh, acc[j]);
  }
  float g[32];
#pragma unroll
  for (int j = 0; j < 16; ++j) unpack2(acc[j], g[2 * j], g[2 * j + 1]);
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const float ei = ex2_approx(g[u]), ef = ex2_approx(g[8 + u]), eo = ex2_approx(g[24 + u]);
    const float eg = ex2_approx(fminf(g[16 + u], kTanhClamp));
    const float cn = fmaf(c[u], rcp_approx(1.0f + ef), (1.0f - eg) * rcp_approx((1.0f + ei) * (1.0f + eg)));
    const float ec = ex2_approx(fminf(cn * (-2.0f * kLog2e), kTanhClamp));
    c[u] = cn;
    h[u] = (1.0f - ec) * rcp_approx((1.0f + eo) * (1.0f + ec));
  }
}

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

template <bool LSTM>
__global__ void __launch_bounds__(128, LSTM ? 8 : 4) torque_kernel(const __grid_constant__ LgkTorqueParams p) {
  const int idx = blockIdx.x * 128 + threadIdx.x;
  const int total = p.num_envs * kDof;
  if (idx >= total) return;
  const int d = idx % kDof;
  float a = p.actions_in[idx];
  a = fminf(fmaxf(a, -p.clip_actions), p.clip_actions);                       // LR:86-87
  if (p.actions_clipped) p.actions_clipped[idx] = a;
  const float2 qs = *reinterpret_cast<const float2*>(p.dof_state + 2 * (size_t)idx);   // (pos, vel)
  const float a_s = f_mul(a, p.action_scale);
  if (LSTM) {
    // ANY:75-76 sea_input, then x * in_scale inside the TorchScript module
    float x[2];
    x[0] = f_sub(f_add(a_s, p.default_dof_pos[d]), qs.x);     // in_scale is folded into w_ih0
    x[1] = qs.y;
    const size_t layer = (size_t)total * 8;
    float* H = p.sea_hidden_state + (size_t)idx * 8;
    float* C = p.sea_cell_state + (size_t)idx * 8;
    float h0[8], c0[8], h1[8], c1[8];
    load8(H, h0); load8(C, c0); load8(H + layer, h1); load8(C + layer, c1);
    lstm_layer<2>(c_lstm.w_ih0, c_lstm.w_hh0, c_lstm.b0, x, h0, c0);
    lstm_layer<8>(c_lstm.w_ih1, c_lstm.w_hh1, c_lstm.b1, h0, h1, c1);
    float y = c_lstm.lin_b;
#pragma unroll
    for (int k = 0; k < 8; ++k) y = fmaf(c_lstm.lin_w[k], h1[k], y);
    p.torques[idx] = y;                                                        // out_scale folded; no clip (ANY:77-78)
    store8(H, h0); store8(C, c0); store8(H + layer, h1); store8(C + layer, c1);
  } else {
    // PD law, op-for-op as torch evaluates it on fp32 tensors (each op rounded; LR:383-395)
    float tq;
    if (p.control_type == LGK_CTRL_P) {
      tq = f_sub(f_mul(p.p_gains[d], f_sub(f_add(a_s, p.default_dof_pos[d]), qs.x)), f_mul(p.d_gains[d], qs.y));
    } else if (p.control_type == LGK_CTRL_V) {
      const float lqd = p.last_dof_vel[idx];
      tq = f_sub(f_mul(p.p_gains[d], f_sub(a_s, qs.y)), f_div(f_mul(p.d_gains[d], f_sub(qs.y, lqd)), p.sim_dt));
    } else {
      tq = a_s;
    }
    p.torques[idx] = fminf(fmaxf(tq, -p.torque_limits[d]), p.torque_limits[d]);
  }
}

}  // namespace lgk

using namespace lgk;

extern "C" int lgk_set_lstm_weights(const LgkLstmWeights* w, void* stream) {
  LGK_REQUIRE(w != nullptr, "lstm weights are null");
  static LstmDev img;   // must outlive the async copy
  const double l2e = 1.4426950408889634;
  auto gate_scale = [&](int row) { return (row >= 16 && row < 24) ? -2.0 * l2e : -l2e; };
  for (int j = 0; j < 32; ++j) {
    const double s = gate_scale(j);
    for (int k = 0; k < 2; ++k) img.w_ih0[k * 32 + j] = (float)(s * (double)w->w_ih0[j * 2 + k] * (double)w->in_scale[k]);
    for (int k = 0; k < 8; ++k) {
      img.w_hh0[k * 32 + j] = (float)(s * (double)w->w_hh0[j * 8 + k]);
      img.w_ih1[k * 32 + j] = (float)(s * (double)w->w_ih1[j * 8 + k]);
      img.w_hh1[k * 32 + j] = (float)(s * (double)w->w_hh1[j * 8 + k]);
    }
    img.b0[j] = (float)(s * ((double)w->b_ih0[j] + (double)w->b_hh0[j]));
    img.b1[j] = (float)(s * ((double)w->b_ih1[j] + (double)w->b_hh1[j]));
  }
  for (int k = 0; k < 8; ++k) img.lin_w[k] = (float)((double)w->out_scale[0] * (double)w->lin_w[k]);
  img.lin_b = (float)((double)w->out_scale[0] * (double)w->lin_b[0]);
  return check_cuda(cudaMemcpyToSymbolAsync(c_lstm, &img, sizeof(LstmDev), 0, cudaMemcpyHostToDevice,
                                            (cudaStream_t)stream), "cudaMemcpyToSymbolAsync(c_lstm)");
}

extern "C" int lgk_compute_torques(const LgkTorqueParams* p, void* stream) {
  LGK_REQUIRE(p != nullptr && p->num_envs > 0, "torques: bad params");
  LGK_REQUIRE(p->actions_in && p->dof_state && p->torques, "torques: null buffer");
  LGK_REQUIRE(p->control_type >= LGK_CTRL_P && p->control_type <= LGK_CTRL_T, "Unknown controller type");
  if (p->use_lstm) {
    LGK_REQUIRE(p->sea_hidden_state && p->sea_cell_state, "torques: LSTM state is null");
    LGK_ALIGNED16(p->sea_hidden_state, "sea_hidden_state"); LGK_ALIGNED16(p->sea_cell_state, "sea_cell_state");
  } else if (p->control_type == LGK_CTRL_V) {
    LGK_REQUIRE(p->last_dof_vel != nullptr, "torques: last_dof_vel is null");
  }
  if ((reinterpret_cast<uintptr_t>(p->dof_state) & 7u) != 0) return set_error(LGK_ERR_ALIGN, "dof_state must be 8-byte aligned");
  const int blocks = (p->num_envs * kDof + 127) / 128;
  if (p->use_lstm) torque_kernel<true><<<blocks, 128, 0, (cudaStream_t)stream>>>(*p);
  else torque_kernel<false><<<blocks, 128, 0, (cudaStream_t)stream>>>(*p);
  count_launch();
  return check_cuda(cudaGetLastError(), "torque_kernel launch");
}
