// lgk_math.cuh -- per-environment maths of the post-physics step, host+device so that
// tests/hostcheck can exercise the exact same source on the CPU (it is NOT a product path).
// Citations: LR = legged_gym/envs/base/legged_robot.py, MATH = legged_gym/utils/math.py,
// CAS = legged_gym/envs/cassie/cassie.py of the reference; TU = isaacgym.torch_utils.
#pragma once
#include "lgk_common.cuh"
#include "lgk_rng.cuh"

namespace lgk {

struct V3 { float x, y, z; };

// TU quat_rotate_inverse: a = v*(2w^2-1); b = cross(qv, v)*w*2; c = qv*dot(qv, v)*2; a - b + c
LGK_HD V3 quat_rotate_inverse(float qx, float qy, float qz, float qw, V3 v) {
  const float s = 2.0f * qw * qw - 1.0f;
  const float cx = qy * v.z - qz * v.y, cy = qz * v.x - qx * v.z, cz = qx * v.y - qy * v.x;
  const float d = qx * v.x + qy * v.y + qz * v.z;
  V3 r;
  r.x = v.x * s - cx * qw * 2.0f + qx * d * 2.0f;
  r.y = v.y * s - cy * qw * 2.0f + qy * d * 2.0f;
  r.z = v.z * s - cz * qw * 2.0f + qz * d * 2.0f;
  return r;
}

// TU quat_apply(q, [1,0,0]) -> heading = atan2(fwd_y, fwd_x) (LR:338-339)
LGK_HD float heading_of(float qx, float qy, float qz, float qw) {
  // t = cross(qv, b)*2 with b=(1,0,0): (0, qz, -qy)*2 ; fwd = b + w*t + cross(qv, t)
  const float ty = 2.0f * qz, tz = -2.0f * qy;
  const float fx = 1.0f + (qy * tz - qz * ty);
  const float fy = qw * ty + (qz * 0.0f - qx * tz);
  return atan2f(fy, fx);
}

// MATH:45-48 wrap_to_pi: torch remainder (fmod + sign fix) by 2*pi, then -2*pi where > pi
LGK_HD float wrap_to_pi(float a) {
  const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
  float m = fmodf(a, two_pi);
  if (m != 0.0f && m < 0.0f) m += two_pi;
  if (m > pi) m -= two_pi;
  return m;
}

// ---- height scan (LR:831-869, MATH:38-42).  Index path is op-exact fp32: every
// product / sum / quotient individually rounded, true IEEE division by the horizontal scale.
struct YawFrame { float zn, wn, rx, ry; };

LGK_HD YawFrame yaw_frame(float qz, float qw, float root_x, float root_y) {
  // normalize((0,0,z,w)): x / max(sqrt(z*z + w*w), 1e-9)
  float n = f_sqrt(f_add(f_mul(qz, qz), f_mul(qw, qw)));
  n = n < 1e-9f ? 1e-9f : n;
  return YawFrame{f_div(qz, n), f_div(qw, n), root_x, root_y};
}

LGK_HD void height_index(const YawFrame& f, float bx, float by, float border, float hscale, int rows,
                         int cols, int& ix, int& iy) {
  // quat_apply((0,0,zn,wn), (bx,by,0)):  t = cross(qv,b)*2 ; out = b + w*t + cross(qv,t)
  const float t0 = f_mul(-f_mul(f.zn, by), 2.0f);
  const float t1 = f_mul(f_mul(f.zn, bx), 2.0f);
  float px = f_add(f_add(bx, f_mul(f.wn, t0)), -f_mul(f.zn, t1));
  float py = f_add(f_add(by, f_mul(f.wn, t1)), f_mul(f.zn, t0));
  px = f_add(f_add(px, f.rx), border);            // + root pos (LR:853-854), + border_size (LR:856)
  py = f_add(f_add(py, f.ry), border);
  const float qx = f_div(px, hscale), qy = f_div(py, hscale);   // LR:857, then .long() = trunc
  // x86 float->int64 conversion (what .long() compiles to) returns INT64_MIN for NaN, inf and |q| >= 2^63; the
  // clip then maps those to 0.  In-range values saturate harmlessly to int32 before the same clip.
#if defined(__CUDA_ARCH__)
  ix = qx < 9.2233720368547758e18f ? __float2int_rz(qx) : 0;    // saturating; NaN -> 0
  iy = qy < 9.2233720368547758e18f ? __float2int_rz(qy) : 0;
#else
  ix = (qx != qx || qx >= 9.2233720368547758e18f) ? 0 : (qx >= 2147483520.f ? 2147483647 : (qx <= -2147483648.f ? (-2147483647 - 1) : (int)qx));
  iy = (qy != qy || qy >= 9.2233720368547758e18f) ? 0 : (qy >= 2147483520.f ? 2147483647 : (qy <= -2147483648.f ? (-2147483647 - 1) : (int)qy));
#endif
  ix = ix < 0 ? 0 : (ix > rows - 2 ? rows - 2 : ix);            // LR:860-861
  iy = iy < 0 ? 0 : (iy > cols - 2 ? cols - 2 : iy);
}

#if defined(__CUDACC__)
// ---- packed variant of the index path (device only).  Same individually-rounded fp32 operations as
// height_index(), issued two-at-a-time as f32x2 instructions (x and y coordinate in one register pair), and the
// IEEE division by the horizontal scale c replaced by the FMA sequence  q0 = x*r; q = fma(fma(-q0, c, x), r, q0)
// with r = RN(1/c), which oracle/divcheck.c proves bit-identical to x / c for EVERY fp32 x in [1e-20, 1e20) for
// c in {0.1f, 0.05f, 0.25f} (outside that range the clipped index is 0 or rows-2 either way).
struct YawFrame2 {
  f2_t T;    // (-2zn, +2zn): times (by, bx) gives (t0, t1)
  f2_t Ts;   // (+2zn, -2zn): times (bx, by) gives (t1, t0)
  f2_t W;    // (wn, wn)
  f2_t Zn;   // (-zn, +zn)
  f2_t R;    // (root_x, root_y)
};

__device__ __forceinline__ YawFrame2 yaw_frame2(const YawFrame& f) {
  const float z2 = f_mul(f.zn, 2.0f);      // exact doubling: fl(2zn*b) == 2*fl(zn*b)
  return YawFrame2{pack2(-z2, z2), pack2(z2, -z2), pack2(f.wn, f.wn), pack2(-f.zn, f.zn), pack2(f.rx, f.ry)};
}

// B = (bx, by), Bs = (by, bx)
// kInRange: the caller has checked that the env's root position is below 1e17 m, so no quotient can reach 2^63 and the
// saturating conversion alone reproduces .long() + clip (NaN -> 0 either way, negative overflow clips to 0 either way).
template <bool kRecipDiv, bool kInRange = false>
__device__ __forceinline__ void height_index2(const YawFrame2& f, f2_t B, f2_t Bs, float border, float hscale,
                                              float hrecip, float one, int rows, int cols, int& ix, int& iy) {
  const f2_t T = mul2(f.T, Bs);                 // (t0, t1) = (-(2zn*by), 2zn*bx)
  const f2_t Tsw = mul2(f.Ts, B);               // (t1, t0)
  // (bx + wn*t0) + (-(zn*t1)),  (by + wn*t1) + zn*t0.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (it
  // honours .rn only on scalar f32, and -fmad=false does not reach it), which flips truncated indices; writing
  // each "product + addend" as fma(product, 1, addend) keeps the product individually rounded.
  // `one` must be a RUN-TIME 1.0f (the compiler folds fma(x, 1.0f, y) back into x + y and contracts again).
  const f2_t One = pack2(one, one);
  f2_t Pp = fma2(mul2(f.Zn, Tsw), One, fma2(mul2(f.W, T), One, B));
  Pp = add2(add2(Pp, f.R), pack2(border, border));
  float qx, qy;
  if (kRecipDiv) {
    const f2_t Rr = pack2(hrecip, hrecip);
    const f2_t Q0 = mul2(Pp, Rr);
    const f2_t Rem = fma2(Q0, pack2(-hscale, -hscale), Pp);
    unpack2(fma2(Rem, Rr, Q0), qx, qy);
  } else {
    float px, py;
    unpack2(Pp, px, py);
    qx = f_div(px, hscale); qy = f_div(py, hscale);
  }
  // .long(): truncation; x86 turns NaN / inf / |q| >= 2^63 into INT64_MIN, which the clip maps to 0
  if (kInRange) {
    ix = __float2int_rz(qx);
    iy = __float2int_rz(qy);
  } else {
    ix = qx < 9.2233720368547758e18f ? __float2int_rz(qx) : 0;
    iy = qy < 9.2233720368547758e18f ? __float2int_rz(qy) : 0;
  }
  ix = min(max(ix, 0), rows - 2);
  iy = min(max(iy, 0), cols - 2);
}
#endif

// torch.norm(dim=-1) over 3 / 2 elements on CPU: sqrt(fma(z,z,fma(y,y,x*x)))
LGK_HD float norm3(float x, float y, float z) { return f_sqrt(f_fma(z, z, f_fma(y, y, f_mul(x, x)))); }
LGK_HD float norm2(float x, float y) { return f_sqrt(f_fma(y, y, f_mul(x, x))); }

// LR:353-366 torch_rand_float: (upper - lower) * u + lower, product and sum rounded separately
LGK_HD float scale_uniform(float range, float lo, float u) { return f_add(f_mul(range, u), lo); }

// LR:347-369 _resample_commands for one env.  cmd = pointer to the env's 4 floats.
LGK_HD void resample_commands(const LgkStepParams& p, float* cmd, const U4& r) {
  cmd[0] = scale_uniform(p.cmd_range[0], p.cmd_lo[0], u32_to_uniform(r.x));
  cmd[1] = scale_uniform(p.cmd_range[1], p.cmd_lo[1], u32_to_uniform(r.y));
  if (p.heading_command) cmd[3] = scale_uniform(p.cmd_range[3], p.cmd_lo[3], u32_to_uniform(r.z));
  else                   cmd[2] = scale_uniform(p.cmd_range[2], p.cmd_lo[2], u32_to_uniform(r.z));
  const float keep = norm2(cmd[0], cmd[1]) > 0.2f ? 1.0f : 0.0f;   // LR:369
  cmd[0] *= keep; cmd[1] *= keep;
}

LGK_HD float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

}  // namespace lgk
