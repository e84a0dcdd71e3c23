// lgk_post_physics.cu -- post-physics step (reference LR:106-230, 329-508, 831-969) for sm_100a: K2, the explicit
// reset_idx kernel, the finalize kernel and the host-side ordering of the launches.  (K1, the per-env / per-joint scalar
// work, lives in lgk_post_k1.cu.)
//
//  K2  scan_obs_fast_kernel (specialised, branch-free) / scan_obs_kernel (generic)   one WARP per env, lanes over
//      columns: the 187-point height scan (packed f32x2 op-exact index path, all int16 gathers of the env issued back
//      to back from the precomputed min3 field), measured_heights, and the finished observation row: the 48 head columns
//      K1 left un-noised in the compact [N,48] hand-over buffer, the height columns, in-kernel Philox noise, clip --
//      lane l owns columns l+32m, so one Philox block serves four coalesced 128-byte row segments.  Persistent CTAs;
//      point grid, noise scales and column masks live in registers.
//
// lgk_post_physics orders them (K1 then K2; when the base_height reward is active the scan runs first) and runs the
// phases of K1 separately when Python code has to run in between.
#include <stdlib.h>
#include <type_traits>
#include "lgk_tile.cuh"

namespace lgk {

// ------------------------------------------------------------------ K2
enum { kScan = 1, kObs = 2 };
constexpr int kK2Threads = 128;
constexpr int kMaxGroupsAny = 12;  // 32-column groups per row: O <= 384

template <int kMaxGroups>
__global__ void __launch_bounds__(kK2Threads, 8) scan_obs_kernel(const __grid_constant__ LgkStepParams p, int mode) {
  extern __shared__ __align__(16) float s_pts[];       // (bx, by, by, bx) per height point
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.num_envs, P = p.num_height_points, O = p.num_obs;
  const bool scan = (mode & kScan) != 0 && p.measure_heights && !p.terrain_is_plane && P > 0;
  const bool obs = (mode & kObs) != 0;
  const bool hcols = p.measure_heights != 0;
  pdl_launch_dependents();
  if (scan) {
    for (int i = tid; i < P; i += kK2Threads) {
      const float bx = p.height_points_xy[2 * i], by = p.height_points_xy[2 * i + 1];
      *reinterpret_cast<float4*>(s_pts + 4 * i) = make_float4(bx, by, by, bx);
    }
  }
  __syncthreads();
  const bool recip_div = p.horizontal_scale_recip != 0.f;
  const float rt_one = __int_as_float(0x3f800000u | ((uint32_t)p.num_envs >> 31));   // 1.0f the compiler cannot see
  const bool noisy = p.add_noise != 0;
  const float clip = p.clip_obs, vs = p.vertical_scale, hsc = p.obs_scale_height;
  // lane l owns columns l + 32g; its noise scales are loop-invariant
  float nz[kMaxGroups];
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g) {
    const int j = 32 * g + lane;
    nz[g] = (obs && noisy && j < O) ? __ldg(p.noise_scale_vec + j) : 0.f;
  }
  const int ngroups = obs ? (O + 31) >> 5 : ((48 + P + 31) >> 5);
  pdl_wait();              // constants above (point grid, noise scales) never change; everything below is per-step state
  const int step_eff = p.step_counter_dev ? (*p.step_counter_dev + 1) : p.step;
  const RngKey key = make_key(p.seed, step_eff);

  // per-env inputs (frame from K1 or the root pose, the two head segments) are fetched one env AHEAD of their use
  struct EnvIn { float4 f; float rz, head0, head1; };
  auto fetch = [&](int env) {
    EnvIn in;
    in.f = make_float4(0.f, 0.f, 0.f, 0.f); in.rz = 0.f; in.head0 = 0.f; in.head1 = 0.f;
    if (env >= N) return in;
    const float* r = p.root_states + ((size_t)env * p.actors_per_env + p.root_actor_offset) * 13;
    if (scan) {
      if (p.scan_frames && obs) in.f = __ldg(reinterpret_cast<const float4*>(p.scan_frames + (size_t)env * kFrameFloats));
      else in.f = make_float4(r[5], r[6], r[0], r[1]);       // scan-only pass before K1: current root pose
    }
    if (obs && hcols) in.rz = p.scan_frames ? p.scan_frames[(size_t)env * kFrameFloats + 4] : r[2] - 0.5f;
    if (obs) {
      const float* orow = p.obs_head ? p.obs_head + (size_t)env * kHeadCols : p.obs_buf + (size_t)env * O;
      in.head0 = orow[lane];                         // columns 0..31, written un-noised by K1
      if (lane < 16) in.head1 = orow[32 + lane];     // columns 32..47
    }
    return in;
  };
  const int stride = gridDim.x * (kK2Threads / 32);
  int env = blockIdx.x * (kK2Threads / 32) + warp;
  EnvIn nxt = fetch(env);
  for (; env < N; env += stride) {
    const EnvIn cur = nxt;
    nxt = fetch(env + stride);
    YawFrame2 yf;
    if (scan) {
      if (p.scan_frames && obs) yf = yaw_frame2(YawFrame{cur.f.x, cur.f.y, cur.f.z, cur.f.w});   // K1's pre-reset frame
      else yf = yaw_frame2(yaw_frame(cur.f.x, cur.f.y, cur.f.z, cur.f.w));
    }
    const float rz = cur.rz;
    float* orow = p.obs_buf + (size_t)env * O;
    float* hrow = p.measured_heights + (size_t)env * P;
    const uint32_t genv = (uint32_t)(p.env_id_offset + env);

    // ---- pass 1: sample offsets of every height column this lane owns (column j <-> point j - 48)
    int off[kMaxGroups];
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      off[g] = -1;
      const int pt = 32 * g + lane - 48;
      if (scan && g >= 1 && g < ngroups && pt >= 0 && pt < P) {
        const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(s_pts + 4 * pt);
        int ix, iy;
        if (recip_div) height_index2<true>(yf, q.x, q.y, p.border_size, p.horizontal_scale, p.horizontal_scale_recip, rt_one, p.hf_rows, p.hf_cols, ix, iy);
        else height_index2<false>(yf, q.x, q.y, p.border_size, p.horizontal_scale, 0.f, rt_one, p.hf_rows, p.hf_cols, ix, iy);
        off[g] = ix * p.hf_cols + iy;
      }
    }
    // ---- pass 2: all gathers back to back (memory-level parallelism)
    float h[kMaxGroups];
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      h[g] = 0.f;
      const int pt = 32 * g + lane - 48;
      if (g >= 1 && g < ngroups && pt >= 0 && pt < P) {
        if (scan) h[g] = f_mul((float)__ldg(p.height_min3 + off[g]), vs);               // LR:869
        else if (obs && hcols && !p.terrain_is_plane) h[g] = hrow[pt];                    // scan ran in an earlier launch
      }
    }
    // ---- pass 3: measured_heights + finished observation columns
    const float head0 = cur.head0, head1 = cur.head1;
#pragma unroll
    for (int sc = 0; sc < (kMaxGroups + 3) / 4; ++sc) {
      if (sc * 4 < ngroups) {
        U4 r = U4{0, 0, 0, 0};
        if (obs && noisy) r = rng_block(key, genv, LGK_STREAM_OBS, (uint32_t)(32 * sc + lane));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int g = sc * 4 + k;
          const int j = 32 * g + lane, pt = j - 48;
          if (g < kMaxGroups && g < ngroups) {
            if ((scan || (p.terrain_is_plane && (mode & kScan) && hcols)) && pt >= 0 && pt < P) hrow[pt] = h[g];
            if (obs && j < O) {
              float v;
              if (g == 0) v = head0;
              else if (g == 1 && lane < 16) v = head1;
              else v = hcols ? f_mul(clampf(rz - h[g], -1.f, 1.f), hsc) : 0.f;
              v = noisy_obs(v, pick(r, k), nz[g]);
              orow[j] = clampf(v, -clip, clip);
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ K2, specialised
// Same work split as scan_obs_kernel for the configurations that matter for throughput (height field present, or no height
// columns at all), written branch-free: the column-group count G, the pass (scan / observations / both) and the division
// variant are template parameters; every lane computes an in-range sample index for each of its G columns whether or not
// the column is a height point (the clip of LR:860-861 makes any input a valid index), so the only predication left is on
// the stores.  Per-lane constants (grid points, noise scales, column masks) live in registers across the env loop.
#ifndef LGK_K2_MINBLOCKS
#define LGK_K2_MINBLOCKS 5
#endif
template <int G, int MODE, bool RECIP>
__global__ void __launch_bounds__(kK2Threads, LGK_K2_MINBLOCKS) scan_obs_fast_kernel(const __grid_constant__ LgkStepParams p) {
  constexpr bool SCAN = (MODE & kScan) != 0, OBS = (MODE & kObs) != 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = p.num_envs, P = p.num_height_points, O = p.num_obs;
  pdl_launch_dependents();
  // ---- lane constants
  f2_t Bv[G], Bsv[G];
  float nz[G];
  uint32_t pmask = 0, omask = 0;      // bit g: column 32g+lane is a height point / an observation column
  const bool noisy = OBS && p.add_noise != 0;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int j = 32 * g + lane, pt = j - 48;
    const bool isp = G > 2 && pt >= 0 && pt < P;
    float bx = 0.f, by = 0.f;
    if (isp && SCAN) { const float2 b = __ldg(reinterpret_cast<const float2*>(p.height_points_xy) + pt); bx = b.x; by = b.y; }
    Bv[g] = pack2(bx, by); Bsv[g] = pack2(by, bx);
    pmask |= (isp ? 1u : 0u) << g;
    omask |= ((OBS && j < O) ? 1u : 0u) << g;
    nz[g] = (noisy && j < O) ? __ldg(p.noise_scale_vec + j) : 0.f;
  }
  const float rt_one = __int_as_float(0x3f800000u | ((uint32_t)p.num_envs >> 31));   // 1.0f the compiler cannot see
  const float clip = p.clip_obs, vs = p.vertical_scale, hsc = p.obs_scale_height;
  const float border = p.border_size, hscale = p.horizontal_scale, hrecip = p.horizontal_scale_recip;
  const int rows = p.hf_rows, cols = p.hf_cols;
  const bool k1_frames = OBS && p.scan_frames != nullptr;     // K1 ran before this pass: use its pre-reset yaw frame
  const float* __restrict__ head_buf = p.obs_head;           // compact [N,48] hand-over buffer (NULL: through obs_buf)
  // a height column is clamp(.,-1,1) * scale + noise: when scale + noise_scale <= clip_obs the final clip (LR:100-101) cannot
  // act on it (the inner clamp also removes NaN), so groups >= 2 (height columns only) skip it
  bool hclip = false;
#pragma unroll
  for (int g = 2; g < G; ++g) hclip = hclip || !(fabsf(hsc) + fabsf(nz[g]) <= clip);
  hclip = __any_sync(0xffffffffu, hclip);
  pdl_wait();
  const int step_eff = p.step_counter_dev ? (*p.step_counter_dev + 1) : p.step;
  const RngKey key = make_key(p.seed, step_eff);

  struct EnvIn { float4 f; float rz, head0, head1; };
  auto fetch = [&](int env) {
    EnvIn in;
    in.f = make_float4(0.f, 0.f, 0.f, 0.f); in.rz = 0.f; in.head0 = 0.f; in.head1 = 0.f;
    if (env >= N) return in;
    const float* r = p.root_states + ((size_t)env * p.actors_per_env + p.root_actor_offset) * 13;
    if (G > 2) {
      if (k1_frames) {
        in.f = __ldg(reinterpret_cast<const float4*>(p.scan_frames + (size_t)env * kFrameFloats));
        in.rz = p.scan_frames[(size_t)env * kFrameFloats + 4];
      } else {
        in.f = make_float4(r[5], r[6], r[0], r[1]);
        in.rz = r[2] - 0.5f;
      }
    }
    if (OBS) {
      const float* orow = head_buf ? head_buf + (size_t)env * kHeadCols : p.obs_buf + (size_t)env * O;
      in.head0 = orow[lane];                         // columns 0..31, written un-noised by K1
      if (lane < 16) in.head1 = orow[32 + lane];     // columns 32..47
    }
    return in;
  };
  const int stride = gridDim.x * (kK2Threads / 32);
  int env = blockIdx.x * (kK2Threads / 32) + warp;
  EnvIn nxt = fetch(env);
  for (; env < N; env += stride) {
    const EnvIn cur = nxt;
    nxt = fetch(env + stride);
    float* orow = p.obs_buf + (size_t)env * O;
    float* hrow = p.measured_heights + (size_t)env * P;
    float h[G];
#pragma unroll
    for (int g = 0; g < G; ++g) h[g] = 0.f;
    if (G > 2) {
      if (SCAN) {
        const YawFrame2 yf = k1_frames ? yaw_frame2(YawFrame{cur.f.x, cur.f.y, cur.f.z, cur.f.w})
                                       : yaw_frame2(yaw_frame(cur.f.x, cur.f.y, cur.f.z, cur.f.w));
        int off[G];
        if (cur.f.z < 1e17f && cur.f.w < 1e17f) {      // (warp-uniform) no quotient can reach 2^63: skip the per-point guard
#pragma unroll
          for (int g = 1; g < G; ++g) {
            int ix, iy;
            height_index2<RECIP, true>(yf, Bv[g], Bsv[g], border, hscale, hrecip, rt_one, rows, cols, ix, iy);
            off[g] = ix * cols + iy;
          }
        } else {
#pragma unroll
          for (int g = 1; g < G; ++g) {
            int ix, iy;
            height_index2<RECIP>(yf, Bv[g], Bsv[g], border, hscale, hrecip, rt_one, rows, cols, ix, iy);
            off[g] = ix * cols + iy;
          }
        }
#pragma unroll
        for (int g = 1; g < G; ++g) h[g] = f_mul((float)__ldg(p.height_min3 + off[g]), vs);      // LR:869
#pragma unroll
        for (int g = 1; g < G; ++g) if ((pmask >> g) & 1u) hrow[32 * g + lane - 48] = h[g];
      } else {
#pragma unroll
        for (int g = 1; g < G; ++g) if ((pmask >> g) & 1u) h[g] = hrow[32 * g + lane - 48];     // scan ran in an earlier launch
      }
    }
    if (OBS) {
      const uint32_t genv = (uint32_t)(p.env_id_offset + env);
      auto finish_row = [&](auto clip_heights) {
        constexpr bool kClipHeights = decltype(clip_heights)::value;
#pragma unroll
        for (int sc = 0; sc < (G + 3) / 4; ++sc) {
          U4 r = U4{0, 0, 0, 0};
          if (noisy) r = rng_block(key, genv, LGK_STREAM_OBS, (uint32_t)(32 * sc + lane));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int g = sc * 4 + k;
            if (g < G) {
              float v;
              if (g == 0) v = cur.head0;
              else {
                v = f_mul(clampf(cur.rz - h[g], -1.f, 1.f), hsc);
                if (g == 1) v = lane < 16 ? cur.head1 : v;
              }
              v = noisy_obs(v, pick(r, k), nz[g]);
              if (g < 2 || kClipHeights) v = clampf(v, -clip, clip);
              if ((omask >> g) & 1u) orow[32 * g + lane] = v;
            }
          }
        }
      };
      if (hclip) finish_row(std::true_type{}); else finish_row(std::false_type{});
    }
  }
}

// ------------------------------------------------------------------ K2, hot path
// lgk_post_physics_finalize: the arguments of the finalize pass when it rides in K2's grid as CTA 0 (see finalize_body)
struct FinArgs {
  int32_t* reset_ids; int32_t* reset_count; float* episode_means; uint8_t* time_outs_extras;
  int enabled;
};
__device__ __noinline__ void finalize_body(const LgkStepParams& p, int32_t* reset_ids, int32_t* reset_count,
                                           float* episode_means, uint8_t* time_outs_extras, int advance, bool in_k2);

// The configuration every rough-terrain step takes (scan + observations in one pass, K1's scan frames and compact head
// buffer present, FMA division): same arithmetic and the same column -> lane / Philox word mapping as
// scan_obs_fast_kernel<G, kScan | kObs, true>, with (a) no generic fetch path, (b) the int16 gathers of an env ISSUED before
// its two Philox blocks are computed and CONSUMED after them, so the ~130 instructions of the noise draw cover the
// gather latency instead of a stalled warp, (c) the per-point 2^63 guard decided once per env.
template <int G>
__global__ void __launch_bounds__(kK2Threads, LGK_K2_MINBLOCKS) scan_obs_hot_kernel(const __grid_constant__ LgkStepParams p,
                                                                                     const __grid_constant__ FinArgs fin) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = p.num_envs, P = p.num_height_points, O = p.num_obs;
  pdl_launch_dependents();
  f2_t Bv[G], Bsv[G];
  float nz[G];
  uint32_t pmask = 0, omask = 0;      // bit g: column 32g+lane is a height point / an observation column
  const bool noisy = p.add_noise != 0;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int j = 32 * g + lane, pt = j - 48;
    const bool isp = pt >= 0 && pt < P;
    float bx = 0.f, by = 0.f;
    if (isp) { const float2 b = __ldg(reinterpret_cast<const float2*>(p.height_points_xy) + pt); bx = b.x; by = b.y; }
    Bv[g] = pack2(bx, by); Bsv[g] = pack2(by, bx);
    pmask |= (isp ? 1u : 0u) << g;
    omask |= ((j < O) ? 1u : 0u) << g;
    nz[g] = (noisy && j < O) ? __ldg(p.noise_scale_vec + j) : 0.f;
  }
  const float rt_one = __int_as_float(0x3f800000u | ((uint32_t)p.num_envs >> 31));   // 1.0f the compiler cannot see
  const float clip = p.clip_obs, vs = p.vertical_scale, hsc = p.obs_scale_height;
  const float border = p.border_size, hscale = p.horizontal_scale, hrecip = p.horizontal_scale_recip;
  const int rows = p.hf_rows, cols = p.hf_cols;
  const float4* __restrict__ frames = reinterpret_cast<const float4*>(p.scan_frames);
  const float* __restrict__ head_buf = p.obs_head;
  const int16_t* __restrict__ field = p.height_min3;
  bool hclip = false;                 // see scan_obs_fast_kernel: the final clip cannot act on a height column unless ...
#pragma unroll
  for (int g = 2; g < G; ++g) hclip = hclip || !(fabsf(hsc) + fabsf(nz[g]) <= clip);
  hclip = __any_sync(0xffffffffu, hclip);
  pdl_wait();
  // with the finalize pass on board (CTA 0) the step comes from word [1] of the counter, which K1 wrote: that CTA
  // advances word [0] while the others are still starting
  if (fin.enabled && blockIdx.x == 0) {
    finalize_body(p, fin.reset_ids, fin.reset_count, fin.episode_means, fin.time_outs_extras, 1, true);
    return;
  }
  const int step_eff = p.step_counter_dev ? (fin.enabled ? p.step_counter_dev[1] : *p.step_counter_dev + 1) : p.step;
  const RngKey key = make_key(p.seed, step_eff);
  const int cta = blockIdx.x - (fin.enabled ? 1 : 0), nctas = gridDim.x - (fin.enabled ? 1 : 0);

  struct EnvIn { float4 f; float rz, head0, head1; };
  auto fetch = [&](int env) {
    EnvIn in;
    in.f = make_float4(0.f, 0.f, 0.f, 0.f); in.rz = 0.f; in.head0 = 0.f; in.head1 = 0.f;
    if (env < N) {
      in.f = frames[2 * (size_t)env];                                  // (zn, wn, root_x, root_y), pre-reset
      in.rz = reinterpret_cast<const float*>(frames + 2 * (size_t)env + 1)[0];      // post-reset root z - 0.5
      const float* hr = head_buf + (size_t)env * kHeadCols;
      in.head0 = hr[lane];
      if (lane < 16) in.head1 = hr[32 + lane];
    }
    return in;
  };
  const int stride = nctas * (kK2Threads / 32);
  int env = cta * (kK2Threads / 32) + warp;
  EnvIn nxt = fetch(env);
  for (; env < N; env += stride) {
    const EnvIn cur = nxt;
    nxt = fetch(env + stride);
    float* orow = p.obs_buf + (size_t)env * O;
    float* hrow = p.measured_heights + (size_t)env * P;
    const YawFrame2 yf = yaw_frame2(YawFrame{cur.f.x, cur.f.y, cur.f.z, cur.f.w});
    int off[G];
    if (cur.f.z < 1e17f && cur.f.w < 1e17f) {      // (warp-uniform) no quotient can reach 2^63: skip the per-point guard
#pragma unroll
      for (int g = 1; g < G; ++g) {
        int ix, iy;
        height_index2<true, true>(yf, Bv[g], Bsv[g], border, hscale, hrecip, rt_one, rows, cols, ix, iy);
        off[g] = ix * cols + iy;
      }
    } else {
#pragma unroll
      for (int g = 1; g < G; ++g) {
        int ix, iy;
        height_index2<true>(yf, Bv[g], Bsv[g], border, hscale, hrecip, rt_one, rows, cols, ix, iy);
        off[g] = ix * cols + iy;
      }
    }
    // ---- gathers in flight ...
    int raw[G];
#pragma unroll
    for (int g = 1; g < G; ++g) raw[g] = (int)__ldg(field + off[g]);
    // ---- ... while the row's noise is drawn
    const uint32_t genv = (uint32_t)(p.env_id_offset + env);
    U4 r[(G + 3) / 4];
#pragma unroll
    for (int sc = 0; sc < (G + 3) / 4; ++sc) {
      r[sc] = U4{0, 0, 0, 0};
      if (noisy) r[sc] = rng_block(key, genv, LGK_STREAM_OBS, (uint32_t)(32 * sc + lane));
    }
    float u[G];                        // 2u - 1 per column group
#pragma unroll
    for (int g = 0; g < G; ++g) u[g] = noise_unit(pick(r[g >> 2], g & 3));
    // ---- heights (LR:869), measured_heights, finished row
    float h[G];
    h[0] = 0.f;
#pragma unroll
    for (int g = 1; g < G; ++g) h[g] = f_mul((float)raw[g], vs);
#pragma unroll
    for (int g = 1; g < G; ++g) if ((pmask >> g) & 1u) hrow[32 * g + lane - 48] = h[g];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float v;
      if (g == 0) v = cur.head0;
      else {
        v = f_mul(clampf(cur.rz - h[g], -1.f, 1.f), hsc);
        if (g == 1) v = lane < 16 ? cur.head1 : v;
      }
      v = f_fma(u[g], nz[g], v);
      if (g < 2 || hclip) v = clampf(v, -clip, clip);
      if ((omask >> g) & 1u) orow[32 * g + lane] = v;
    }
  }
}

// ------------------------------------------------------------------ reset_idx on an explicit id list
__global__ void __launch_bounds__(128) reset_idx_kernel(const __grid_constant__ LgkStepParams p,
                                                       const int64_t* __restrict__ ids, int n) {
  // one warp per listed env: lane 0 performs the per-env reset on a small shared scratch row set, then the
  // warp writes rows back cooperatively (same write path as the step kernel).
  __shared__ float s_root[4][13], s_dof[4][24], s_cmd[4][4], s_fat[4][LGK_MAX_FEET];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * 4 + w;
  const bool live = i < n;
  const int env = live ? (int)ids[i] : 0;
  const int N = p.num_envs, F = p.num_feet;
  const int step_eff = p.step_counter_dev ? *p.step_counter_dev : p.step;
  const RngKey key = make_key(p.seed, step_eff);
  const size_t rrow = ((size_t)env * p.actors_per_env + p.root_actor_offset) * 13;
  if (live) {
    if (lane < 13) s_root[w][lane] = p.root_states[rrow + lane];
    if (lane < 24) s_dof[w][lane] = p.dof_state[(size_t)env * 24 + lane];
    if (lane < 4) s_cmd[w][lane] = p.commands[(size_t)env * 4 + lane];
    if (lane < F) s_fat[w][lane] = p.feet_air_time[(size_t)env * F + lane];
  }
  __syncwarp();
  if (live && lane == 0) {
    long long ep = 0;
    env_reset(p, key, (uint32_t)(p.env_id_offset + env), env, s_root[w], s_dof[w], s_cmd[w], s_fat[w], ep);
    p.episode_length_buf[env] = 0;
    p.reset_buf[env] = 1;                                   // LR:177
  }
  __syncwarp();
  float* stats = p.reset_stats + (size_t)(step_eff & 1) * (p.num_reward_slots + 2);
  if (live) {
    if (lane < 13) p.root_states[rrow + lane] = s_root[w][lane];
    if (lane < 24) p.dof_state[(size_t)env * 24 + lane] = s_dof[w][lane];
    if (lane < 4) p.commands[(size_t)env * 4 + lane] = s_cmd[w][lane];
    if (lane < F) p.feet_air_time[(size_t)env * F + lane] = s_fat[w][lane];
    if (lane < 12) { p.last_actions[(size_t)env * 12 + lane] = 0.f; p.last_dof_vel[(size_t)env * 12 + lane] = 0.f; }  // LR:173-174
    for (int k = lane; k < p.num_reward_slots; k += 32) {
      float* s = p.episode_sums + (size_t)k * N + env;
      atomicAdd(stats + k, *s);
      *s = 0.f;
    }
    if (lane == 0) atomicAdd(stats + p.num_reward_slots, 1.0f);
    if (p.zero_lstm_on_reset && p.sea_hidden_state) {
      const size_t layer = (size_t)N * 96;
      for (int j = lane; j < 96; j += 32) {
        p.sea_hidden_state[(size_t)env * 96 + j] = 0.f; p.sea_hidden_state[layer + (size_t)env * 96 + j] = 0.f;
        p.sea_cell_state[(size_t)env * 96 + j] = 0.f;   p.sea_cell_state[layer + (size_t)env * 96 + j] = 0.f;
      }
    }
  }
}

// sum(terrain_levels) for the explicit reset path (LR:186 takes the mean over ALL envs)
__global__ void terrain_level_sum_kernel(const __grid_constant__ LgkStepParams p) {
  float v = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.num_envs; i += gridDim.x * blockDim.x)
    v += (float)p.terrain_levels[i];
  v = warp_sum(v);
  const int step_eff = p.step_counter_dev ? *p.step_counter_dev : p.step;
  if ((threadIdx.x & 31) == 0) atomicAdd(p.reset_stats + (size_t)(step_eff & 1) * (p.num_reward_slots + 2) + p.num_reward_slots + 1, v);
}

// ------------------------------------------------------------------ finalize: id compaction + extras
// Single CTA, single sweep: thread t owns the contiguous flag range [t*chunk, (t+1)*chunk) (chunk a multiple of 16 so
// every load is one aligned 16-byte vector), counts it, one block-wide exclusive scan, then emits the ids of its range
// -- ascending like reset_buf.nonzero() (LR:128).  The kernel sits at the end of the step's dependent chain, so what
// matters is its latency: every global load it needs (flags, time-out flags, both parities of the statistics, the
// step counter) is issued right after the dependency wait, before the first instruction that consumes one -- one L2
// round trip instead of four -- and the CTA is sized to the batch (lgk_finalize_step) so that it becomes resident while
// K2's CTAs still hold most of every SM's registers.
__device__ __forceinline__ int count_flags16(uint4 v) {
  // flags are 0/1 bytes (bool tensors): the byte sum is the popcount
  return __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
}

constexpr int kFinRegs = 2;      // 16-byte vectors of flags (and of time-out flags) a thread keeps in registers
constexpr int kFinBatch = 8;     // vectors per round trip on the longer sweeps

// in_k2: the body runs as CTA 0 of K2's grid (lgk_post_physics_finalize) instead of as a kernel of its own at the end of
// the chain: the step is then word [1] of the counter (K1 put it there; K2's other CTAs read the same word) and the
// advance goes to word [0], which nothing running concurrently reads.
__device__ __noinline__ void finalize_body(const LgkStepParams& p, int32_t* reset_ids, int32_t* reset_count,
                                           float* episode_means, uint8_t* time_outs_extras, int advance, bool in_k2) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nwarps = nthr >> 5;
  const int N = p.num_envs, ns = p.num_reward_slots;
  // ---- loads
  const int ctr = p.step_counter_dev ? (in_k2 ? p.step_counter_dev[1] - 1 : *p.step_counter_dev) : 0;
  const int chunk = (((N + nthr - 1) / nthr) + 15) & ~15;
  const int first = tid * chunk, last = min(N, first + chunk);
  const bool vec = (reinterpret_cast<uintptr_t>(p.reset_buf) & 15u) == 0;
  const bool in_regs = vec && (N & 15) == 0 && chunk <= 16 * kFinRegs;     // whole range of every thread in registers
  uint4 fl[kFinRegs];
#pragma unroll
  for (int q = 0; q < kFinRegs; ++q) {
    fl[q] = make_uint4(0u, 0u, 0u, 0u);
    if (in_regs && first + 16 * q < last) fl[q] = *reinterpret_cast<const uint4*>(p.reset_buf + first + 16 * q);
  }
  const bool want_to = p.send_timeouts && time_outs_extras;
  const bool v16 = ((reinterpret_cast<uintptr_t>(p.time_out_buf) | reinterpret_cast<uintptr_t>(time_outs_extras)) & 15u) == 0;
  const int n16 = v16 ? N / 16 : 0;
  const bool to_regs = want_to && n16 <= nthr * kFinRegs;
  uint4 tv[kFinRegs];
#pragma unroll
  for (int q = 0; q < kFinRegs; ++q) {
    tv[q] = make_uint4(0u, 0u, 0u, 0u);
    if (to_regs && tid + q * nthr < n16) tv[q] = reinterpret_cast<const uint4*>(p.time_out_buf)[tid + q * nthr];
  }
  const float st0 = tid < ns + 2 ? p.reset_stats[tid] : 0.f;
  const float st1 = tid < ns + 2 ? p.reset_stats[ns + 2 + tid] : 0.f;
  // ---- count + block-wide exclusive scan
  int c = 0;
  if (in_regs) {
#pragma unroll
    for (int q = 0; q < kFinRegs; ++q) c += count_flags16(fl[q]);
  } else {
    // longer ranges: kFinBatch vectors in flight per round trip (one load per iteration would make the sweep a chain of
    // L2 latencies: the riding CTA then outlasts K2 at 65 536 envs)
    int i = first;
    if (vec) {
      for (; i + 16 * kFinBatch <= last; i += 16 * kFinBatch) {
        uint4 v[kFinBatch];
#pragma unroll
        for (int q = 0; q < kFinBatch; ++q) v[q] = *reinterpret_cast<const uint4*>(p.reset_buf + i + 16 * q);
#pragma unroll
        for (int q = 0; q < kFinBatch; ++q) c += count_flags16(v[q]);
      }
      for (; i + 16 <= last; i += 16) c += count_flags16(*reinterpret_cast<const uint4*>(p.reset_buf + i));
    }
    for (; i < last; ++i) c += p.reset_buf[i] != 0;
  }
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int v = lane < nwarps ? s_warp[lane] : 0;
    int wi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
    s_warp[lane] = wi - v;
    if (lane == 31) s_total = wi;
  }
  __syncthreads();
  const int count = s_total;
  // ---- ids of this thread's range
  if (reset_ids && c) {
    int off = s_warp[warp] + incl - c;
    if (in_regs) {
#pragma unroll
      for (int q = 0; q < kFinRegs; ++q) {
        const uint32_t w[4] = {fl[q].x, fl[q].y, fl[q].z, fl[q].w};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          uint32_t m = w[r];
          while (m) { const int b = __ffs(m) - 1; m &= m - 1; reset_ids[off++] = first + 16 * q + 4 * r + (b >> 3); }
        }
      }
    } else {
      auto emit16 = [&](uint4 v, int base) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          uint32_t m = w[r];
          while (m) { const int b = __ffs(m) - 1; m &= m - 1; reset_ids[off++] = base + 4 * r + (b >> 3); }
        }
      };
      int i = first;
      if (vec) {
        for (; i + 16 * kFinBatch <= last; i += 16 * kFinBatch) {
          uint4 v[kFinBatch];
#pragma unroll
          for (int q = 0; q < kFinBatch; ++q) v[q] = *reinterpret_cast<const uint4*>(p.reset_buf + i + 16 * q);
#pragma unroll
          for (int q = 0; q < kFinBatch; ++q) emit16(v[q], i + 16 * q);
        }
        for (; i + 16 <= last; i += 16) emit16(*reinterpret_cast<const uint4*>(p.reset_buf + i), i);
      }
      for (; i < last; ++i) if (p.reset_buf[i]) reset_ids[off++] = i;
    }
  }
  const int step_eff = p.step_counter_dev ? (ctr + (advance ? 1 : 0)) : p.step;
  const float cur = (step_eff & 1) ? st1 : st0;                  // this thread's entry of the current parity
  float* other = p.reset_stats + (size_t)((step_eff + 1) & 1) * (ns + 2);
  if (tid == 0 && reset_count) *reset_count = count;
  if (count > 0) {     // the reference refreshes extras only inside reset_idx with a non-empty id list
    if (episode_means) {
      if (tid < ns) episode_means[tid] = cur / (float)count / p.max_episode_length_s;
      if (tid == ns + 1) episode_means[ns] = p.terrain_curriculum ? cur / (float)N : 0.f;
    }
    if (want_to) {
      if (to_regs) {
#pragma unroll
        for (int q = 0; q < kFinRegs; ++q)
          if (tid + q * nthr < n16) reinterpret_cast<uint4*>(time_outs_extras)[tid + q * nthr] = tv[q];
      } else {
        const uint4* src = reinterpret_cast<const uint4*>(p.time_out_buf);
        uint4* dst = reinterpret_cast<uint4*>(time_outs_extras);
        int i = tid;
        for (; i + (kFinBatch - 1) * nthr < n16; i += kFinBatch * nthr) {
          uint4 v[kFinBatch];
#pragma unroll
          for (int q = 0; q < kFinBatch; ++q) v[q] = src[i + q * nthr];
#pragma unroll
          for (int q = 0; q < kFinBatch; ++q) dst[i + q * nthr] = v[q];
        }
        for (; i < n16; i += nthr) dst[i] = src[i];
      }
      for (int i = n16 * 16 + tid; i < N; i += nthr) time_outs_extras[i] = p.time_out_buf[i];
    }
  }
  if (tid < ns + 2) other[tid] = 0.f;
  if (advance && p.step_counter_dev && tid == 0) *p.step_counter_dev = step_eff;
}

__global__ void __launch_bounds__(1024) finalize_kernel(const __grid_constant__ LgkStepParams p, int32_t* reset_ids,
                                                        int32_t* reset_count, float* episode_means,
                                                        uint8_t* time_outs_extras, int advance) {
  pdl_launch_dependents();
  pdl_wait();
  finalize_body(p, reset_ids, reset_count, episode_means, time_outs_extras, advance, false);
}

}  // namespace lgk

using namespace lgk;

static int validate_step(const LgkStepParams* p) {
  LGK_REQUIRE(p != nullptr, "params is null");
  LGK_REQUIRE(p->num_envs > 0, "num_envs must be positive");
  LGK_REQUIRE(p->num_bodies > 0 && p->num_bodies <= LGK_MAX_BODIES, "num_bodies out of range");
  LGK_REQUIRE(p->num_feet >= 0 && p->num_feet <= LGK_MAX_FEET, "num_feet out of range");
  LGK_REQUIRE(p->num_pen >= 0 && p->num_pen <= LGK_MAX_PEN, "num_pen out of range");
  LGK_REQUIRE(p->num_term >= 0 && p->num_term <= LGK_MAX_TERM, "num_term out of range");
  LGK_REQUIRE(p->actors_per_env >= 1 && p->root_actor_offset >= 0 && p->root_actor_offset < p->actors_per_env, "bad actor layout");
  LGK_REQUIRE(!p->predator_spawn || (p->predator_actor_offset >= 0 && p->predator_actor_offset < p->actors_per_env &&
                                     p->predator_actor_offset != p->root_actor_offset), "bad predator actor offset");
  LGK_REQUIRE(p->num_obs == 48 + (p->measure_heights ? p->num_height_points : 0), "num_obs must be 48 + P");
  LGK_REQUIRE(p->resample_period > 0, "resample_period must be positive");
  LGK_REQUIRE(p->root_states && p->dof_state && p->contact_forces && p->actions && p->torques && p->commands &&
              p->episode_length_buf && p->last_actions && p->last_dof_vel && p->last_root_vel && p->episode_sums &&
              p->base_lin_vel && p->base_ang_vel && p->projected_gravity && p->obs_buf && p->rew_buf && p->reset_buf &&
              p->time_out_buf && p->reset_stats && p->noise_scale_vec, "a required buffer is null");
  LGK_REQUIRE(p->num_feet == 0 || (p->feet_air_time && p->last_contacts), "feet buffers are null");
  if (p->measure_heights) {
    LGK_REQUIRE(p->measured_heights && p->num_height_points > 0, "measured_heights missing");
    if (!p->terrain_is_plane) LGK_REQUIRE(p->height_min3 && p->height_points_xy && p->hf_rows >= 2 && p->hf_cols >= 2, "height field missing");
  }
  if (p->terrain_curriculum) LGK_REQUIRE(p->terrain_levels && p->terrain_types && p->terrain_origins && p->env_origins, "terrain curriculum buffers are null");
  for (int k = 0; k < LGK_R_COUNT; ++k)
    if (p->reward_active[k]) LGK_REQUIRE(p->reward_slot[k] >= 0 && p->reward_slot[k] < p->num_reward_slots, "bad reward slot");
  LGK_ALIGNED16(p->root_states, "root_states"); LGK_ALIGNED16(p->dof_state, "dof_state");
  LGK_ALIGNED16(p->contact_forces, "contact_forces"); LGK_ALIGNED16(p->actions, "actions");
  LGK_ALIGNED16(p->torques, "torques"); LGK_ALIGNED16(p->commands, "commands");
  LGK_ALIGNED16(p->last_actions, "last_actions"); LGK_ALIGNED16(p->last_dof_vel, "last_dof_vel");
  LGK_ALIGNED16(p->last_root_vel, "last_root_vel"); LGK_ALIGNED16(p->base_lin_vel, "base_lin_vel");
  LGK_ALIGNED16(p->base_ang_vel, "base_ang_vel"); LGK_ALIGNED16(p->projected_gravity, "projected_gravity");
  if (p->num_feet) { LGK_ALIGNED16(p->feet_air_time, "feet_air_time"); LGK_ALIGNED16(p->last_contacts, "last_contacts"); }
  if (p->sea_hidden_state) { LGK_ALIGNED16(p->sea_hidden_state, "sea_hidden_state"); LGK_ALIGNED16(p->sea_cell_state, "sea_cell_state"); }
  return LGK_OK;
}

// the specialised kernel of the rough-terrain step (scan_obs_hot_kernel) applies
static bool k2_hot(const LgkStepParams* p, int mode) {
  const bool field = p->measure_heights && !p->terrain_is_plane && p->num_height_points > 0;
  return field && mode == (kScan | kObs) && p->horizontal_scale_recip != 0.f && p->scan_frames != nullptr &&
         p->obs_head != nullptr && p->actors_per_env == 1 && getenv("LGK_K2_NO_HOT") == nullptr;
}

// fin (optional, only when k2_hot): the finalize pass rides in the grid as CTA 0
static int launch_k2(const LgkStepParams* p, int mode, cudaStream_t st, const FinArgs* fin = nullptr) {
  const int wpb = kK2Threads / 32;
  int blocks = (p->num_envs + wpb - 1) / wpb;
  const int cap = 148 * 16;                          // persistent beyond one full wave of resident CTAs
  if (blocks > cap) blocks = cap;
  const size_t smem = (size_t)(p->num_height_points > 0 ? p->num_height_points : 1) * 16;
  const int groups = (48 + p->num_height_points + 31) / 32;
  cudaError_t e;
  // specialised branch-free kernels for the throughput configurations: a height field behind every height column, or
  // no height columns at all; everything else (plane terrain with measure_heights, ...) takes the generic kernel
  const bool field = p->measure_heights && !p->terrain_is_plane && p->num_height_points > 0;
  const bool flat = !p->measure_heights;
  if (field || (flat && mode == kObs)) {
    const int fblocks = blocks < 148 * 2 * LGK_K2_MINBLOCKS ? blocks : 148 * 2 * LGK_K2_MINBLOCKS;
    const bool rc = p->horizontal_scale_recip != 0.f;
    const dim3 gd(fblocks), bd(kK2Threads);
#define LGK_K2(GG, MM) (rc ? launch_chained(scan_obs_fast_kernel<GG, MM, true>, gd, bd, 0, st, *p) \
                           : launch_chained(scan_obs_fast_kernel<GG, MM, false>, gd, bd, 0, st, *p))
    const bool hot = k2_hot(p, mode);
    if (fin && !hot) return set_error(LGK_ERR_ARG, "fused finalize needs the hot K2 path");
    FinArgs fa = {};
    if (fin) { fa = *fin; fa.enabled = 1; }
    const dim3 gh(fblocks + (fin ? 1 : 0));
    if (hot && groups <= 8) e = launch_chained(scan_obs_hot_kernel<8>, gh, bd, 0, st, *p, fa);
    else if (hot) e = launch_chained(scan_obs_hot_kernel<12>, gh, bd, 0, st, *p, fa);
    else if (flat) e = launch_chained(scan_obs_fast_kernel<2, kObs, false>, gd, bd, 0, st, *p);
    else if (groups <= 8) e = mode == kScan ? LGK_K2(8, kScan) : (mode == kObs ? LGK_K2(8, kObs) : LGK_K2(8, kScan | kObs));
    else e = mode == kScan ? LGK_K2(12, kScan) : (mode == kObs ? LGK_K2(12, kObs) : LGK_K2(12, kScan | kObs));
#undef LGK_K2
    count_launch();
    return check_cuda(e, "scan_obs_fast_kernel launch");
  }
  if (fin) return set_error(LGK_ERR_ARG, "fused finalize needs the hot K2 path");
  if (groups <= 2) e = launch_chained(scan_obs_kernel<2>, dim3(blocks), dim3(kK2Threads), smem, st, *p, mode);
  else if (groups <= 8) e = launch_chained(scan_obs_kernel<8>, dim3(blocks), dim3(kK2Threads), smem, st, *p, mode);
  else e = launch_chained(scan_obs_kernel<12>, dim3(blocks), dim3(kK2Threads), smem, st, *p, mode);
  count_launch();
  return check_cuda(e, "scan_obs_kernel launch");
}

extern "C" int lgk_post_physics(const LgkStepParams* p, void* stream) {
  if (int rc = validate_step(p)) return rc;
  const int all = LGK_PHASE_PRE | LGK_PHASE_POST | LGK_PHASE_POST_REWARD | LGK_PHASE_POST_OBS;
  LGK_REQUIRE((p->phase_mask & all) != 0 && (p->phase_mask & ~all) == 0, "phase_mask selects nothing");
  LGK_REQUIRE(p->num_height_points <= 256 && p->num_obs <= 32 * kMaxGroupsAny, "at most 256 height points");
  cudaStream_t st = (cudaStream_t)stream;
  const bool pre = (p->phase_mask & LGK_PHASE_PRE) != 0;
  const bool obsph = (p->phase_mask & (LGK_PHASE_POST | LGK_PHASE_POST_OBS)) != 0;
  const bool heights = p->measure_heights != 0;
  // the base_height reward needs this step's heights inside K1 (LR:884-887): run the scan first in that case
  const bool scan_first = heights && p->reward_active[LGK_R_BASE_HEIGHT] != 0;
  if (heights && !p->terrain_is_plane) LGK_REQUIRE(p->scan_frames != nullptr, "scan_frames buffer is null");
  if (pre && obsph) {             // whole step in one pass
    static const int only = getenv("LGK_PP_ONLY") ? atoi(getenv("LGK_PP_ONLY")) : 0;     // timing aid (profiles/pp_probe.py)
    if (only == 1) return launch_k1(p, st);
    if (only == 2) return launch_k2(p, kScan | kObs, st);
    if (scan_first)
      if (int rc = launch_k2(p, kScan, st)) return rc;
    if (int rc = launch_k1(p, st)) return rc;
    if (!heights) return LGK_OK;                    // flat tasks: K1 finished the 48-column rows itself
    return launch_k2(p, scan_first ? kObs : (kScan | kObs), st);
  }
  if (pre) {                      // Python code follows (user reward terms may read measured_heights): scan now
    if (heights)
      if (int rc = launch_k2(p, kScan, st)) return rc;
    return launch_k1(p, st);
  }
  if (int rc = launch_k1(p, st)) return rc;      // POST / POST_REWARD / POST_OBS
  return (heights && obsph) ? launch_k2(p, kObs, st) : LGK_OK;
}

extern "C" int lgk_reset_idx(const LgkStepParams* p, const int64_t* env_ids, int32_t num_ids, void* stream) {
  if (int rc = validate_step(p)) return rc;
  LGK_REQUIRE(num_ids >= 0, "num_ids negative");
  if (num_ids == 0) return LGK_OK;
  LGK_REQUIRE(env_ids != nullptr, "env_ids is null");
  reset_idx_kernel<<<(num_ids + 3) / 4, 128, 0, (cudaStream_t)stream>>>(*p, env_ids, num_ids);
  count_launch();
  if (p->terrain_curriculum) {
    terrain_level_sum_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(*p);
    count_launch();
  }
  return check_cuda(cudaGetLastError(), "reset_idx_kernel launch");
}

extern "C" int lgk_finalize_step(const LgkStepParams* p, int32_t* reset_ids, int32_t* reset_count,
                                 float* episode_means, uint8_t* time_outs_extras, int32_t advance, void* stream) {
  if (int rc = validate_step(p)) return rc;
  LGK_REQUIRE(p->num_reward_slots + 2 <= 128, "too many reward slots");
  // one 16-byte flag vector per thread up to 16 384 envs (128 ... 1024 threads)
  int threads = ((p->num_envs + 15) / 16 + 31) & ~31;
  threads = threads < 128 ? 128 : (threads > 1024 ? 1024 : threads);
  static const int force = getenv("LGK_FINALIZE_THREADS") ? atoi(getenv("LGK_FINALIZE_THREADS")) : 0;   // A/B aid
  if (force) threads = force;
  const cudaError_t e = launch_chained(finalize_kernel, dim3(1), dim3(threads), 0, (cudaStream_t)stream, *p, reset_ids, reset_count,
                                       episode_means, time_outs_extras, (int)advance);
  count_launch();
  return check_cuda(e, "finalize_kernel launch");
}

extern "C" int lgk_post_physics_finalize(const LgkStepParams* p, int32_t* reset_ids, int32_t* reset_count,
                                         float* episode_means, uint8_t* time_outs_extras, void* stream) {
  if (int rc = validate_step(p)) return rc;
  LGK_REQUIRE(p->phase_mask == (LGK_PHASE_PRE | LGK_PHASE_POST), "lgk_post_physics_finalize runs the whole step (PRE | POST)");
  const bool scan_first = p->measure_heights && p->reward_active[LGK_R_BASE_HEIGHT] != 0;
  static const int no_fuse = getenv("LGK_NO_FUSED_FINALIZE") ? 1 : 0;      // A/B aid
  // the riding CTA has K2's 128 threads: up to 1024 flags per thread (three sweeps of eight 16-byte vectors per round
  // trip); beyond that the stand-alone 1024-thread kernel.  At 4096 / 16 384 envs the fused form saves 3.4 / 3.8 us of 45.5 / 81.2
  static const int limit = getenv("LGK_FUSED_FINALIZE_MAX") ? atoi(getenv("LGK_FUSED_FINALIZE_MAX")) : 1024 * kK2Threads;   // A/B aid
  const bool small = p->num_envs <= limit;
  if (no_fuse || !small || scan_first || !k2_hot(p, kScan | kObs) || p->num_height_points > 256 || p->num_reward_slots + 2 > kK2Threads) {
    if (int rc = lgk_post_physics(p, stream)) return rc;                   // no K2 after K1, or not the specialised one
    return lgk_finalize_step(p, reset_ids, reset_count, episode_means, time_outs_extras, 1, stream);
  }
  LgkStepParams q = *p;
  q.phase_mask |= kPhaseFusedFin;
  if (int rc = launch_k1(&q, (cudaStream_t)stream)) return rc;
  const FinArgs fin = {reset_ids, reset_count, episode_means, time_outs_extras, 1};
  return launch_k2(&q, kScan | kObs, (cudaStream_t)stream, &fin);
}
