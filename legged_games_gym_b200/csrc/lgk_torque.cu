// lgk_torque.cu -- _compute_torques: PD law (reference LR:371-395) and the ANYdrive SEA actuator network
// (ANY:71-81 + resources/actuator_nets/anydrive_v3_lstm.pt: in_scale -> LSTM(2,8,2 layers) -> Linear(8,1) -> out_scale).
//
// One thread per (env, joint) sequence.  The 969 LSTM weights live in __constant__ memory and every weight
// reference is a compile-time offset, so each MAC is one FFMA with a constant-bank operand (no weight loads);
// the 32 floats of h/c state per sequence stream through registers as 128-bit loads/stores.
// Gate nonlinearities use MUFU ex2/rcp with the products sigma(i)*tanh(g) and sigma(o)*tanh(c') sharing one
// reciprocal each (8 MUFU per hidden unit instead of 10) -- the MUFU pipe, not HBM, is the next limiter.
#include "lgk_math.cuh"
#include <cstdlib>

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
  // per-role copies for torque_lstm_split_kernel: role r owns gate rows 8g + 2r + {0,1}; for each input k the 8 floats
  // [g][2] are contiguous, so a role's slice of a matrix row is two LDCU.128 with compile-time offsets
  float role_w[4][(2 + 8 + 8 + 8) * 8];     // w_ih0 (k<2) | w_hh0 | w_ih1 | w_hh1
  float role_b[4][2][8];                     // layer 0 / layer 1 biases [g][2]
};
__constant__ LstmDev c_lstm;

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kTanhClamp = 43.0f;   // = 2*log2(e)*14.9: tanh saturates in fp32 long before, keeps 1 - 2^x finite

// one LSTM layer, hidden 8; IN = input width.  torch gate order i, f, g, o (rows 0-7, 8-15, 16-23, 24-31).
//   c' = sigmoid(f)*c + sigmoid(i)*tanh(g) ; h' = sigmoid(o)*tanh(c')
// with sigmoid(a)*tanh(b) = (1 - eb) / ((1 + ea)(1 + eb)), ea = e^-a, eb = e^-2b: 5 ex2 + 3 rcp per unit.
template <int IN, bool PACKED = false>
__device__ __forceinline__ void lstm_layer(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                           const float* __restrict__ bias, const float (&x)[IN], float (&h)[8],
                                           float (&c)[8]) {
  unsigned long long acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = pack2(bias[2 * j], bias[2 * j + 1]);
  // recurrent part first: in this order ptxas keeps the hoisted LDCU.128 weight loads inside the 63 uniform registers
  // (8 spill instructions instead of 112 with the input part first; 1176 instead of 1280 instructions per sequence)
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const unsigned long long hh = pack2(h[k], h[k]);
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = fma2(pack2(w_hh[k * 32 + 2 * j], w_hh[k * 32 + 2 * j + 1]), hh, acc[j]);
  }
#pragma unroll
  for (int k = 0; k < IN; ++k) {
    const unsigned long long xx = pack2(x[k], x[k]);
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = fma2(pack2(w_ih[k * 32 + 2 * j], w_ih[k * 32 + 2 * j + 1]), xx, acc[j]);
  }
  if (PACKED) {
  // gates two hidden units at a time on the packed fp32 pipe (acc[j] already pairs the units 2j', 2j'+1 of one gate type:
  // rows 0-7 i, 8-15 f, 16-23 g, 24-31 o); only the MUFU ops and the clamps are scalar
  const f2_t one2 = pack2(1.0f, 1.0f), mone2 = pack2(-1.0f, -1.0f), scale2 = pack2(-2.0f * kLog2e, -2.0f * kLog2e);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float gi0, gi1, gf0, gf1, gg0, gg1, go0, go1;
    unpack2(acc[q], gi0, gi1); unpack2(acc[4 + q], gf0, gf1); unpack2(acc[8 + q], gg0, gg1); unpack2(acc[12 + q], go0, go1);
    const f2_t ei = pack2(ex2_approx(gi0), ex2_approx(gi1)), ef = pack2(ex2_approx(gf0), ex2_approx(gf1));
    const f2_t eo = pack2(ex2_approx(go0), ex2_approx(go1));
    const f2_t eg = pack2(ex2_approx(fminf(gg0, kTanhClamp)), ex2_approx(fminf(gg1, kTanhClamp)));
    float d0, d1, f0, f1;
    unpack2(mul2(add2(ei, one2), add2(eg, one2)), d0, d1);                  // (1 + ei)(1 + eg)
    unpack2(add2(ef, one2), f0, f1);                                        // 1 + ef
    const f2_t num = fma2(eg, mone2, one2);                                 // 1 - eg
    const f2_t cn = fma2(pack2(c[2 * q], c[2 * q + 1]), pack2(rcp_approx(f0), rcp_approx(f1)),
                         mul2(num, pack2(rcp_approx(d0), rcp_approx(d1))));
    float t0, t1;
    unpack2(mul2(cn, scale2), t0, t1);
    const f2_t ec = pack2(ex2_approx(fminf(t0, kTanhClamp)), ex2_approx(fminf(t1, kTanhClamp)));
    float e0, e1;
    unpack2(mul2(add2(eo, one2), add2(ec, one2)), e0, e1);                  // (1 + eo)(1 + ec)
    const f2_t hn = mul2(fma2(ec, mone2, one2), pack2(rcp_approx(e0), rcp_approx(e1)));
    unpack2(cn, c[2 * q], c[2 * q + 1]);
    unpack2(hn, h[2 * q], h[2 * q + 1]);
  }
  } else {
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
}

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// HOSTIO: dof_state may be pinned host memory (uncached loads) and the torques are mirrored into torques_mirror; the
// device-resident instantiation carries neither (the uncached load alone cost 3 % at 65 536 envs).
// one wave of the device-resident LSTM kernel: 8 CTAs of 128 sequences on each of the 148 SMs
constexpr size_t kLstmWaveSeqs = (size_t)148 * 8 * 128;

template <bool LSTM, bool HOSTIO = false, bool PACKED = false>
__global__ void __launch_bounds__(128, LSTM ? 8 : 4) torque_kernel(const __grid_constant__ LgkTorqueParams p) {
  pdl_launch_dependents();
  const int idx = blockIdx.x * 128 + threadIdx.x;
  const int total = p.num_envs * kDof;
  if (idx >= total) return;
  const int d = idx % kDof;
  pdl_wait();                       // the predecessor on the stream produced actions / dof state / LSTM state
  float a = p.actions_in[idx];
  a = fminf(fmaxf(a, -p.clip_actions), p.clip_actions);                       // LR:86-87
  if (p.actions_clipped) p.actions_clipped[idx] = a;
  const float2* qsp = reinterpret_cast<const float2*>(p.dof_state + 2 * (size_t)idx);   // (pos, vel)
  const float2 qs = HOSTIO ? __ldcv(qsp) : *qsp;
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
    if (!HOSTIO) {  // the state of the block that takes this CTA's slot one wave later: into L2 now, a DRAM round trip saved then
      const size_t pf = (size_t)idx + kLstmWaveSeqs;
      if (pf < (size_t)total) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.sea_hidden_state + pf * 8));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.sea_cell_state + pf * 8));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.sea_hidden_state + layer + pf * 8));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.sea_cell_state + layer + pf * 8));
      }
    }
    lstm_layer<2, PACKED>(c_lstm.w_ih0, c_lstm.w_hh0, c_lstm.b0, x, h0, c0);
    lstm_layer<8, PACKED>(c_lstm.w_ih1, c_lstm.w_hh1, c_lstm.b1, h0, h1, c1);
    float y = c_lstm.lin_b;
#pragma unroll
    for (int k = 0; k < 8; ++k) y = fmaf(c_lstm.lin_w[k], h1[k], y);
    p.torques[idx] = y;                                                        // out_scale folded; no clip (ANY:77-78)
    if (HOSTIO && p.torques_mirror) p.torques_mirror[idx] = y;
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
    const float tc = fminf(fmaxf(tq, -p.torque_limits[d]), p.torque_limits[d]);
    p.torques[idx] = tc;
    if (HOSTIO && p.torques_mirror) p.torques_mirror[idx] = tc;
  }
}


// ------------------------------------------------------------------ role-split variant (small and medium batches)
// CTA = 4 warps x 32 sequences.  Warp w owns hidden units {2w, 2w+1} of BOTH layers, i.e. the 8 gate rows
// {2w, 2w+1} + {0, 8, 16, 24}: four packed (FFMA2) accumulators per layer, a quarter of the matvec and of the
// MUFU work per thread.  The 32 floats of h/c state per sequence enter and leave through shared memory with fully
// coalesced 128-bit global accesses; the new h vectors are exchanged through shared memory between the layers.
// Per-thread instruction stream is ~4x shorter than in torque_kernel<true> (and the code ~4x smaller), which is what
// matters when the batch cannot fill the machine (4096 envs = 49k sequences).
constexpr int kSeqPerCta = 32;

__device__ __forceinline__ void gates_to_state(const f2_t (&acc)[4], float (&c)[2], float (&h)[2]) {
  float gi[2], gf[2], gg[2], go[2];
  unpack2(acc[0], gi[0], gi[1]); unpack2(acc[1], gf[0], gf[1]); unpack2(acc[2], gg[0], gg[1]); unpack2(acc[3], go[0], go[1]);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const float ei = ex2_approx(gi[u]), ef = ex2_approx(gf[u]), eo = ex2_approx(go[u]);
    const float eg = ex2_approx(fminf(gg[u], kTanhClamp));
    const float cn = fmaf(c[u], rcp_approx(1.0f + ef), (1.0f - eg) * rcp_approx((1.0f + ei) * (1.0f + eg)));
    const float ec = ex2_approx(fminf(cn * (-2.0f * kLog2e), kTanhClamp));
    c[u] = cn;
    h[u] = (1.0f - ec) * rcp_approx((1.0f + eo) * (1.0f + ec));
  }
}

// acc[g] += W_role[k][g][0..1] * x[k] for the four gates g; W points at the role's [k][8] slice (compile-time offsets)
template <int IN>
__device__ __forceinline__ void role_matvec(const float* __restrict__ W, const float (&x)[IN], f2_t (&acc)[4]) {
#pragma unroll
  for (int k = 0; k < IN; ++k) {
    const f2_t xx = pack2(x[k], x[k]);
#pragma unroll
    for (int g = 0; g < 4; ++g) acc[g] = fma2(pack2(W[k * 8 + 2 * g], W[k * 8 + 2 * g + 1]), xx, acc[g]);
  }
}

// both layers for role R; shared-memory exchange of the new layer-0 hidden vector in the middle
template <int R>
__device__ __forceinline__ void role_body(const float (&x)[2], int lane, float (*s_state)[kSeqPerCta][8],
                                          float (*s_hnew)[kSeqPerCta][8]) {
  constexpr int w2 = 2 * R;
  const float* W = c_lstm.role_w[R];
  float hin[8], c[2], h[2];
  f2_t acc[4];
  // ---- layer 0
  *reinterpret_cast<float4*>(hin) = *reinterpret_cast<const float4*>(&s_state[0][lane][0]);
  *reinterpret_cast<float4*>(hin + 4) = *reinterpret_cast<const float4*>(&s_state[0][lane][4]);
  c[0] = s_state[2][lane][w2]; c[1] = s_state[2][lane][w2 + 1];
#pragma unroll
  for (int g = 0; g < 4; ++g) acc[g] = pack2(c_lstm.role_b[R][0][2 * g], c_lstm.role_b[R][0][2 * g + 1]);
  role_matvec<2>(W, x, acc);
  role_matvec<8>(W + 2 * 8, hin, acc);
  gates_to_state(acc, c, h);
  s_hnew[0][lane][w2] = h[0]; s_hnew[0][lane][w2 + 1] = h[1];
  const float c0n0 = c[0], c0n1 = c[1];
  // ---- layer 1: the recurrent part does not depend on layer 0 and goes first
  float h1in[8];
  *reinterpret_cast<float4*>(h1in) = *reinterpret_cast<const float4*>(&s_state[1][lane][0]);
  *reinterpret_cast<float4*>(h1in + 4) = *reinterpret_cast<const float4*>(&s_state[1][lane][4]);
  c[0] = s_state[3][lane][w2]; c[1] = s_state[3][lane][w2 + 1];
#pragma unroll
  for (int g = 0; g < 4; ++g) acc[g] = pack2(c_lstm.role_b[R][1][2 * g], c_lstm.role_b[R][1][2 * g + 1]);
  role_matvec<8>(W + (2 + 8 + 8) * 8, h1in, acc);
  __syncthreads();                                        // s_hnew[0] complete; every role has read its old state
  *reinterpret_cast<float4*>(hin) = *reinterpret_cast<const float4*>(&s_hnew[0][lane][0]);
  *reinterpret_cast<float4*>(hin + 4) = *reinterpret_cast<const float4*>(&s_hnew[0][lane][4]);
  role_matvec<8>(W + (2 + 8) * 8, hin, acc);
  gates_to_state(acc, c, h);
  s_hnew[1][lane][w2] = h[0]; s_hnew[1][lane][w2 + 1] = h[1];
  // new cell states overwrite the staged old ones (each role only ever touches its own two units)
  s_state[2][lane][w2] = c0n0; s_state[2][lane][w2 + 1] = c0n1;
  s_state[3][lane][w2] = c[0]; s_state[3][lane][w2 + 1] = c[1];
}

__global__ void __launch_bounds__(128) torque_lstm_split_kernel(const __grid_constant__ LgkTorqueParams p) {
  // [array: h0, h1, c0, c1][seq 32][8]
  __shared__ __align__(16) float s_state[4][kSeqPerCta][8];
  __shared__ __align__(16) float s_hnew[2][kSeqPerCta][8];
  pdl_launch_dependents();
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int total = p.num_envs * kDof;
  const int seq0 = blockIdx.x * kSeqPerCta;
  pdl_wait();
  const int nseq = min(kSeqPerCta, total - seq0);
  const size_t layer = (size_t)total * 8;
  // ---- coalesced state load: 4 arrays x (32 seq x 8 floats) = 4 x 64 float4
  for (int i = tid; i < 4 * 64; i += 128) {
    const int arr = i >> 6, q = i & 63;                   // q = float4 index inside the 32x8 block
    if ((q >> 1) < nseq) {
      const float* src = ((arr & 2) ? p.sea_cell_state : p.sea_hidden_state) + (size_t)(arr & 1) * layer + (size_t)seq0 * 8;
      reinterpret_cast<float4*>(&s_state[arr][0][0])[q] = reinterpret_cast<const float4*>(src)[q];
    }
  }
  const int idx = seq0 + lane;
  const bool live = lane < nseq;
  float x[2] = {0.f, 0.f};
  if (live) {
    const int d = idx % kDof;
    float a = p.actions_in[idx];
    a = fminf(fmaxf(a, -p.clip_actions), p.clip_actions);                     // LR:86-87
    if (p.actions_clipped && w == 0) p.actions_clipped[idx] = a;
    const float2 qs = *reinterpret_cast<const float2*>(p.dof_state + 2 * (size_t)idx);
    x[0] = f_sub(f_add(f_mul(a, p.action_scale), p.default_dof_pos[d]), qs.x);   // ANY:75 (in_scale folded into w_ih0)
    x[1] = qs.y;                                                                 // ANY:76
  }
  __syncthreads();
  switch (w) {            // warp-uniform: each role runs its own copy with compile-time weight offsets
    case 0: role_body<0>(x, lane, s_state, s_hnew); break;
    case 1: role_body<1>(x, lane, s_state, s_hnew); break;
    case 2: role_body<2>(x, lane, s_state, s_hnew); break;
    default: role_body<3>(x, lane, s_state, s_hnew); break;
  }
  __syncthreads();
  // ---- output layer (warp 0) and coalesced state store
  if (w == 0 && live) {
    float y = c_lstm.lin_b;
#pragma unroll
    for (int k = 0; k < 8; ++k) y = fmaf(c_lstm.lin_w[k], s_hnew[1][lane][k], y);
    p.torques[idx] = y;                                   // out_scale folded; no clip on this path (ANY:77-78)
  }
  for (int i = tid; i < 4 * 64; i += 128) {
    const int arr = i >> 6, q = i & 63;
    if ((q >> 1) < nseq) {
      float* dst = ((arr & 2) ? p.sea_cell_state : p.sea_hidden_state) + (size_t)(arr & 1) * layer + (size_t)seq0 * 8;
      const float* src = (arr & 2) ? &s_state[arr][0][0] : &s_hnew[arr & 1][0][0];
      reinterpret_cast<float4*>(dst)[q] = reinterpret_cast<const float4*>(src)[q];
    }
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
  for (int r = 0; r < 4; ++r) {
    const float* mats[4] = {img.w_ih0, img.w_hh0, img.w_ih1, img.w_hh1};
    const int ks[4] = {2, 8, 8, 8};
    int o = 0;
    for (int m = 0; m < 4; ++m)
      for (int k = 0; k < ks[m]; ++k)
        for (int g = 0; g < 4; ++g)
          for (int h = 0; h < 2; ++h) img.role_w[r][o++] = mats[m][k * 32 + 8 * g + 2 * r + h];
    for (int g = 0; g < 4; ++g)
      for (int h = 0; h < 2; ++h) {
        img.role_b[r][0][2 * g + h] = img.b0[8 * g + 2 * r + h];
        img.role_b[r][1][2 * g + h] = img.b1[8 * g + 2 * r + h];
      }
  }
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
  LGK_REQUIRE(p->lstm_variant >= 0 && p->lstm_variant <= 3, "lstm_variant must be 0..3");
  LGK_REQUIRE(p->host_io || p->torques_mirror == nullptr, "torques_mirror needs host_io = 1");
  LGK_REQUIRE(!(p->host_io && p->use_lstm && p->lstm_variant == 2), "the role-split LSTM kernel has no host_io path");
  // auto = one thread per sequence (measured faster than the role-split CTA at 4096, 16384 and 65536 envs), with the gate
  // arithmetic on the packed fp32 pipe while the batch fits one wave (latency-bound: 6.2 against 6.6 us at 4096 envs) and
  // scalar beyond (throughput-bound: 37.1 against 41.1 us at 65 536 envs); profiles/torque_probe.py compares the variants
  const bool split = p->lstm_variant == 2;
  const bool packed = p->lstm_variant == 3 || (p->lstm_variant == 0 && (size_t)p->num_envs * kDof <= kLstmWaveSeqs);
  cudaError_t e;
  if (p->use_lstm && split)
    e = launch_chained(torque_lstm_split_kernel, dim3((p->num_envs * kDof + kSeqPerCta - 1) / kSeqPerCta), dim3(128), 0, (cudaStream_t)stream, *p);
  else if (p->use_lstm && packed && !p->host_io)
    e = launch_chained(torque_kernel<true, false, true>, dim3(blocks), dim3(128), 0, (cudaStream_t)stream, *p);
  else if (p->use_lstm) e = p->host_io ? launch_chained(torque_kernel<true, true>, dim3(blocks), dim3(128), 0, (cudaStream_t)stream, *p)
                                       : launch_chained(torque_kernel<true, false>, dim3(blocks), dim3(128), 0, (cudaStream_t)stream, *p);
  else e = p->host_io ? launch_chained(torque_kernel<false, true>, dim3(blocks), dim3(128), 0, (cudaStream_t)stream, *p)
                      : launch_chained(torque_kernel<false, false>, dim3(blocks), dim3(128), 0, (cudaStream_t)stream, *p);
  count_launch();
  return check_cuda(e, "torque_kernel launch");
}
