// lgk_post_physics.cu -- fused post-physics step (reference LR:106-230, 329-508, 831-969) for sm_100a.
//
// One CTA (4 warps) owns a TILE of 32 consecutive environments.  Because every reference tensor is
// env-major row-major, the tile's slice of root_states / dof_state / contact_forces / actions / torques /
// last_actions / last_dof_vel / commands / feet_air_time is ONE contiguous chunk per tensor: each is
// fetched with a single TMA bulk copy (cp.async.bulk.shared::cluster.global, completion on an mbarrier)
// and the whole-tile results (commands, feet_air_time, last_*, base_*) leave through bulk stores
// (cp.async.bulk.global.shared::cta).  Compute threads only touch shared memory:
//   warp 0, lane = env : all per-env scalar work (rotations, commands, termination, the reward terms in
//                        the reference's alphabetical order, reset_idx, the 48 proprioceptive columns)
//   warps 1-3 (+ warp 0 when done): the 187-point height scan, 32 points per warp-step, one int16 gather
//                        per point from the precomputed min3 field (lgk_height_min3)
//   all warps          : observation rows (height columns + in-kernel Philox noise + clip), coalesced
//                        128-byte row segments; cooperative write-back of reset rows and LSTM-state zeroing
// Cross-env sums for extras["episode"] go through warp shuffles + one atomicAdd per tile into a
// ping-pong accumulator consumed by lgk_finalize_step.
#include "lgk_step_device.cuh"
#include <cooperative_groups.h>

namespace lgk {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;

// ------------------------------------------------------------------ PTX helpers (TMA bulk copy + mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  while (!mbar_try_wait(bar, phase)) {}
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ shared-memory carve-up (bytes, 16-aligned)
struct TileLayout {
  int root, dof, contact, act, tq, lact, ldv, cmd, fat, lc, head, blv, bav, pg, lrv, pts, yaw, h16, hstride, misc, total;
};

__host__ __device__ inline int al16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline TileLayout make_layout(int kTile, int nb, int nfeet, int npts) {
  TileLayout L;
  int o = 0;
  L.root = o;    o += al16(kTile * 13 * 4);
  L.dof = o;     o += al16(kTile * 24 * 4);
  L.contact = o; o += al16(kTile * nb * 3 * 4);
  L.act = o;     o += al16(kTile * 12 * 4);
  L.tq = o;      o += al16(kTile * 12 * 4);
  L.lact = o;    o += al16(kTile * 12 * 4);
  L.ldv = o;     o += al16(kTile * 12 * 4);
  L.cmd = o;     o += al16(kTile * 4 * 4);
  L.fat = o;     o += al16(kTile * (nfeet > 0 ? nfeet : 1) * 4);
  L.lc = o;      o += al16(kTile * (nfeet > 0 ? nfeet : 1));
  L.head = o;    o += al16(kTile * 48 * 4);
  L.blv = o;     o += al16(kTile * 3 * 4);
  L.bav = o;     o += al16(kTile * 3 * 4);
  L.pg = o;      o += al16(kTile * 3 * 4);
  L.lrv = o;     o += al16(kTile * 6 * 4);
  L.pts = o;     o += al16((npts > 0 ? npts : 1) * 4 * 4);   // (bx, by, by, bx) per point
  L.yaw = o;     o += al16(kTile * 40);                     // YawFrame2 per env
  L.hstride = (npts + 7) & ~7;                       // int16 samples per env row
  L.h16 = o;     o += al16(kTile * (L.hstride > 0 ? L.hstride : 8) * 2);
  L.misc = o;    o += 64;   // mbarrier (8 B) + chunk counter + reset mask
  L.total = o;
  return L;
}

struct Misc {
  uint64_t bar;
  int chunk_counter;
  uint32_t reset_mask;
  uint32_t valid_mask;
};

// cooperative copy helpers for the non-bulk (partial tile / strided root) path
__device__ __forceinline__ void copy_f32(float* dst, const float* src, int n, int tid) {
  for (int i = tid; i < n; i += kThreads) dst[i] = src[i];
}
__device__ __forceinline__ void copy_u8(uint8_t* dst, const uint8_t* src, int n, int tid) {
  for (int i = tid; i < n; i += kThreads) dst[i] = src[i];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ the kernel
template <int kTile>
__global__ void __launch_bounds__(kThreads, 8) post_physics_kernel(const __grid_constant__ LgkStepParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const TileLayout L = make_layout(kTile, p.num_bodies, p.num_feet, p.num_height_points);
  float* s_root = reinterpret_cast<float*>(smem + L.root);
  float* s_dof = reinterpret_cast<float*>(smem + L.dof);
  float* s_contact = reinterpret_cast<float*>(smem + L.contact);
  float* s_act = reinterpret_cast<float*>(smem + L.act);
  float* s_tq = reinterpret_cast<float*>(smem + L.tq);
  float* s_lact = reinterpret_cast<float*>(smem + L.lact);
  float* s_ldv = reinterpret_cast<float*>(smem + L.ldv);
  float* s_cmd = reinterpret_cast<float*>(smem + L.cmd);
  float* s_fat = reinterpret_cast<float*>(smem + L.fat);
  uint8_t* s_lc = smem + L.lc;
  float* s_head = reinterpret_cast<float*>(smem + L.head);
  float* s_blv = reinterpret_cast<float*>(smem + L.blv);
  float* s_bav = reinterpret_cast<float*>(smem + L.bav);
  float* s_pg = reinterpret_cast<float*>(smem + L.pg);
  float* s_lrv = reinterpret_cast<float*>(smem + L.lrv);
  float* s_pts = reinterpret_cast<float*>(smem + L.pts);
  YawFrame2* s_yaw = reinterpret_cast<YawFrame2*>(smem + L.yaw);
  int16_t* s_h16 = reinterpret_cast<int16_t*>(smem + L.h16);
  const int HS = L.hstride;
  Misc* misc = reinterpret_cast<Misc*>(smem + L.misc);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int env0 = blockIdx.x * kTile;
  const int nval = min(kTile, p.num_envs - env0);     // valid envs in this tile
  const int NB = p.num_bodies, F = p.num_feet, P = p.num_height_points, O = p.num_obs, N = p.num_envs;
  const bool pre = (p.phase_mask & LGK_PHASE_PRE) != 0, post = (p.phase_mask & LGK_PHASE_POST) != 0;
  const bool fat_active = p.reward_active[LGK_R_FEET_AIR_TIME] != 0 && F > 0;
  // bulk (TMA) path needs a full tile (sizes are then multiples of 16 B) and unit root stride
  const bool bulk = (nval == kTile) && (p.actors_per_env == 1) && ((kTile * F) % 16 == 0);
  const int step_eff = p.step_counter_dev ? (*p.step_counter_dev + 1) : p.step;
  const bool do_push = p.step_counter_dev ? (p.push_interval > 0 && step_eff % p.push_interval == 0) : (p.do_push != 0);
  const RngKey key = make_key(p.seed, step_eff);

  // ---------------- stage the tile
  if (tid == 0) {
    misc->chunk_counter = 0;
    misc->reset_mask = 0;
    mbar_init(&misc->bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (bulk) {
    if (tid == 0) {
      uint32_t bytes = kTile * (13 + 24 + 3 * NB + 12 + 12 + 12 + 12 + 4) * 4;
      if (F > 0) bytes += kTile * F * 4 + kTile * F;
      mbar_expect_tx(&misc->bar, bytes);
      bulk_g2s(s_root, p.root_states + (size_t)env0 * 13, kTile * 13 * 4, &misc->bar);
      bulk_g2s(s_dof, p.dof_state + (size_t)env0 * 24, kTile * 24 * 4, &misc->bar);
      bulk_g2s(s_contact, p.contact_forces + (size_t)env0 * NB * 3, kTile * NB * 3 * 4, &misc->bar);
      bulk_g2s(s_act, p.actions + (size_t)env0 * 12, kTile * 12 * 4, &misc->bar);
      bulk_g2s(s_tq, p.torques + (size_t)env0 * 12, kTile * 12 * 4, &misc->bar);
      bulk_g2s(s_lact, p.last_actions + (size_t)env0 * 12, kTile * 12 * 4, &misc->bar);
      bulk_g2s(s_ldv, p.last_dof_vel + (size_t)env0 * 12, kTile * 12 * 4, &misc->bar);
      bulk_g2s(s_cmd, p.commands + (size_t)env0 * 4, kTile * 4 * 4, &misc->bar);
      if (F > 0) {
        bulk_g2s(s_fat, p.feet_air_time + (size_t)env0 * F, kTile * F * 4, &misc->bar);
        bulk_g2s(s_lc, p.last_contacts + (size_t)env0 * F, kTile * F, &misc->bar);
      }
    }
  } else {
    for (int i = tid; i < nval * 13; i += kThreads) {
      const int e = i / 13, c = i - e * 13;
      s_root[i] = p.root_states[((size_t)(env0 + e) * p.actors_per_env + p.root_actor_offset) * 13 + c];
    }
    copy_f32(s_dof, p.dof_state + (size_t)env0 * 24, nval * 24, tid);
    copy_f32(s_contact, p.contact_forces + (size_t)env0 * NB * 3, nval * NB * 3, tid);
    copy_f32(s_act, p.actions + (size_t)env0 * 12, nval * 12, tid);
    copy_f32(s_tq, p.torques + (size_t)env0 * 12, nval * 12, tid);
    copy_f32(s_lact, p.last_actions + (size_t)env0 * 12, nval * 12, tid);
    copy_f32(s_ldv, p.last_dof_vel + (size_t)env0 * 12, nval * 12, tid);
    copy_f32(s_cmd, p.commands + (size_t)env0 * 4, nval * 4, tid);
    if (F > 0) {
      copy_f32(s_fat, p.feet_air_time + (size_t)env0 * F, nval * F, tid);
      copy_u8(s_lc, p.last_contacts + (size_t)env0 * F, nval * F, tid);
    }
  }
  // constants that are not part of the tile: height grid -> smem (P*2 floats, L2-resident)
  if (p.measure_heights && !p.terrain_is_plane)
    for (int i = tid; i < P; i += kThreads) {
      const float bx = p.height_points_xy[2 * i], by = p.height_points_xy[2 * i + 1];
      *reinterpret_cast<float4*>(s_pts + 4 * i) = make_float4(bx, by, by, bx);
    }
  if (bulk) mbar_wait(&misc->bar, 0);
  __syncthreads();

  // ---------------- yaw frames for the height scan (pre-reset root pose, LR:853-854)
  const bool scan = pre && p.measure_heights && !p.terrain_is_plane && P > 0;
  if (warp == 0 && lane < nval && scan) {
    const float* r = s_root + lane * 13;
    s_yaw[lane] = yaw_frame2(yaw_frame(r[5], r[6], r[0], r[1]));
  }
  const bool heights_first = scan && p.reward_active[LGK_R_BASE_HEIGHT];
  __syncthreads();

  // ---------------- height scan: dynamic 32-point chunks (env e, points 32c..32c+31)
  // one work item = one env: all ceil(P/32) point chunks of the env are indexed first, then their gathers are
  // issued back to back (memory-level parallelism), then stored.  Items are claimed dynamically so warp 0 can
  // join after its scalar phase.
  const bool recip_div = p.horizontal_scale_recip != 0.f;
  const float rt_one = __int_as_float(0x3f800000u | ((uint32_t)p.num_envs >> 31));   // 1.0f the compiler cannot see
  auto height_scan = [&]() {
    constexpr int kMaxChunks = 8;                       // P <= 256
    const int cpe = (P + 31) >> 5;
    while (true) {
      int e = 0;
      if (lane == 0) e = atomicAdd(&misc->chunk_counter, 1);
      e = __shfl_sync(0xffffffffu, e, 0);
      if (e >= nval) break;
      const YawFrame2 yf = s_yaw[e];
      int off[kMaxChunks];
#pragma unroll
      for (int c = 0; c < kMaxChunks; ++c) {
        const int j = (c << 5) + lane;
        off[c] = -1;
        if (c < cpe && j < P) {
          int ix, iy;
          const ulonglong2 pt = *reinterpret_cast<const ulonglong2*>(s_pts + 4 * j);
          if (recip_div) height_index2<true>(yf, pt.x, pt.y, p.border_size, p.horizontal_scale, p.horizontal_scale_recip, rt_one, p.hf_rows, p.hf_cols, ix, iy);
          else height_index2<false>(yf, pt.x, pt.y, p.border_size, p.horizontal_scale, 0.f, rt_one, p.hf_rows, p.hf_cols, ix, iy);
          off[c] = ix * p.hf_cols + iy;
        }
      }
      int16_t h[kMaxChunks];
#pragma unroll
      for (int c = 0; c < kMaxChunks; ++c) h[c] = off[c] >= 0 ? __ldg(p.height_min3 + off[c]) : (int16_t)0;
#pragma unroll
      for (int c = 0; c < kMaxChunks; ++c) {
        if (off[c] >= 0) {
          const int j = (c << 5) + lane;
          s_h16[e * HS + j] = h[c];
          p.measured_heights[(size_t)(env0 + e) * P + j] = f_mul((float)h[c], p.vertical_scale);   // LR:869
        }
      }
    }
  };
  if (pre && p.measure_heights && p.terrain_is_plane)      // LR:844-845: zeros
    for (int i = tid; i < nval * P; i += kThreads) { p.measured_heights[(size_t)env0 * P + i] = 0.f; s_h16[(i / P) * HS + (i % P)] = 0; }
  if (heights_first) {
    height_scan();
    __syncthreads();
  }

  // ---------------- warp 0: per-env scalar work, lane = env
  if (warp == 0) {
    const int e = lane, env = env0 + e;
    const bool valid = e < nval;
    const uint32_t genv = (uint32_t)(p.env_id_offset + env);
    EnvScalars s;
    s.reset = false; s.time_out = false; s.rew = 0.f; s.ep_len = 0;
    float* root = s_root + e * 13;
    float* dof = s_dof + e * 24;
    float* cmd = s_cmd + e * 4;
    float* fat = s_fat + e * F;
    uint8_t* lc = s_lc + e * F;
    float* sums = p.episode_sums + env;
    if (valid) {
      if (pre) {
        float mh = 0.f;
        if (p.reward_active[LGK_R_BASE_HEIGHT]) {        // mean_p(z - h_p), LR:886
          if (p.measure_heights) {
            for (int j = 0; j < P; ++j) mh += root[2] - f_mul((float)s_h16[e * HS + j], p.vertical_scale);
            mh /= (float)P;
          } else {
            mh = root[2];                                 // measured_heights is the int 0 (LR:562)
          }
        }
        env_pre(p, do_push, key, genv, root, dof, s_contact + e * NB * 3, s_act + e * 12, s_tq + e * 12, s_lact + e * 12,
                s_ldv + e * 12, cmd, fat, lc, sums, N, p.episode_length_buf[env], mh, s);
        s_blv[3 * e] = s.blv.x; s_blv[3 * e + 1] = s.blv.y; s_blv[3 * e + 2] = s.blv.z;
        s_bav[3 * e] = s.bav.x; s_bav[3 * e + 1] = s.bav.y; s_bav[3 * e + 2] = s.bav.z;
        s_pg[3 * e] = s.pg.x; s_pg[3 * e + 1] = s.pg.y; s_pg[3 * e + 2] = s.pg.z;
      } else {   // split mode: PRE ran in an earlier launch, pick its results up from global memory
        s.blv = V3{p.base_lin_vel[3 * env], p.base_lin_vel[3 * env + 1], p.base_lin_vel[3 * env + 2]};
        s.bav = V3{p.base_ang_vel[3 * env], p.base_ang_vel[3 * env + 1], p.base_ang_vel[3 * env + 2]};
        s.pg = V3{p.projected_gravity[3 * env], p.projected_gravity[3 * env + 1], p.projected_gravity[3 * env + 2]};
        s.ep_len = p.episode_length_buf[env];
        s.reset = p.reset_buf[env] != 0;
        s.time_out = p.time_out_buf[env] != 0;
        s.rew = p.rew_buf[env];
      }
      if (post) {
        s.rew = env_finish_reward(p, s.rew, s.reset, s.time_out, sums, N);
        if (s.reset) env_reset(p, key, genv, env, root, dof, cmd, fat, s.ep_len);
        env_obs_head(p, s, dof, cmd, s_act + e * 12, s_head + e * 48);
        for (int d = 0; d < 12; ++d) s_ldv[e * 12 + d] = dof[2 * d + 1];    // LR:133 (post-reset dof_vel)
        for (int i = 0; i < 6; ++i) s_lrv[e * 6 + i] = root[7 + i];         // LR:134 (post push/reset)
      }
      p.rew_buf[env] = s.rew;
      p.episode_length_buf[env] = s.ep_len;
      if (pre) {
        p.reset_buf[env] = s.reset ? 1 : 0;
        p.time_out_buf[env] = s.time_out ? 1 : 0;
      }
    }
    const uint32_t rmask = __ballot_sync(0xffffffffu, valid && s.reset && post);
    // extras["episode"] sums over the reset set + zeroing (LR:179-183), terrain-level mean (LR:186)
    if (post) {
      float* stats = p.reset_stats + (size_t)(step_eff & 1) * (p.num_reward_slots + 2);
      if (rmask != 0) {
        for (int k = 0; k < p.num_reward_slots; ++k) {
          float v = 0.f;
          if (valid && s.reset) { v = sums[(size_t)k * N]; sums[(size_t)k * N] = 0.f; }
          v = warp_sum(v);
          if (lane == 0) atomicAdd(stats + k, v);
        }
        if (lane == 0) atomicAdd(stats + p.num_reward_slots, (float)__popc(rmask));
      }
      if (p.terrain_curriculum) {
        float lv = valid ? (float)p.terrain_levels[env] : 0.f;
        lv = warp_sum(lv);
        if (lane == 0) atomicAdd(stats + p.num_reward_slots + 1, lv);
      }
    }
    if (lane == 0) misc->reset_mask = rmask;
  }
  if (scan && !heights_first) height_scan();     // warps 1-3 start immediately; warp 0 joins when done
  fence_async_smem();
  __syncthreads();

  // ---------------- whole-tile write-backs
  if (bulk) {
    if (tid == 0) {
      if (pre) {
        bulk_s2g(p.base_lin_vel + (size_t)env0 * 3, s_blv, kTile * 3 * 4);
        bulk_s2g(p.base_ang_vel + (size_t)env0 * 3, s_bav, kTile * 3 * 4);
        bulk_s2g(p.projected_gravity + (size_t)env0 * 3, s_pg, kTile * 3 * 4);
        if (do_push && !post) bulk_s2g(p.root_states + (size_t)env0 * 13, s_root, kTile * 13 * 4);
      }
      if (pre || post) bulk_s2g(p.commands + (size_t)env0 * 4, s_cmd, kTile * 4 * 4);
      if (fat_active || (post && F > 0)) {
        bulk_s2g(p.feet_air_time + (size_t)env0 * F, s_fat, kTile * F * 4);
        if (pre) bulk_s2g(p.last_contacts + (size_t)env0 * F, s_lc, kTile * F);
      }
      if (post) {
        bulk_s2g(p.last_actions + (size_t)env0 * 12, s_act, kTile * 12 * 4);
        bulk_s2g(p.last_dof_vel + (size_t)env0 * 12, s_ldv, kTile * 12 * 4);
        bulk_s2g(p.last_root_vel + (size_t)env0 * 6, s_lrv, kTile * 6 * 4);
        if (do_push && pre) bulk_s2g(p.root_states + (size_t)env0 * 13, s_root, kTile * 13 * 4);
      }
      bulk_commit();
    }
  } else {
    if (pre) {
      copy_f32(p.base_lin_vel + (size_t)env0 * 3, s_blv, nval * 3, tid);
      copy_f32(p.base_ang_vel + (size_t)env0 * 3, s_bav, nval * 3, tid);
      copy_f32(p.projected_gravity + (size_t)env0 * 3, s_pg, nval * 3, tid);
    }
    copy_f32(p.commands + (size_t)env0 * 4, s_cmd, nval * 4, tid);
    if (fat_active || (post && F > 0)) {
      copy_f32(p.feet_air_time + (size_t)env0 * F, s_fat, nval * F, tid);
      if (pre) copy_u8(p.last_contacts + (size_t)env0 * F, s_lc, nval * F, tid);
    }
    if (post) {
      copy_f32(p.last_actions + (size_t)env0 * 12, s_act, nval * 12, tid);
      copy_f32(p.last_dof_vel + (size_t)env0 * 12, s_ldv, nval * 12, tid);
      copy_f32(p.last_root_vel + (size_t)env0 * 6, s_lrv, nval * 6, tid);
    }
    if (do_push) {
      for (int i = tid; i < nval * 13; i += kThreads) {
        const int e = i / 13, c = i - e * 13;
        if (c == 7 || c == 8)
          p.root_states[((size_t)(env0 + e) * p.actors_per_env + p.root_actor_offset) * 13 + c] = s_root[i];
      }
    }
  }

  if (post) {
    // ---------------- reset rows: dof_state / root_states write-back + LSTM state zeroing (ANY:56-60)
    uint32_t rm = misc->reset_mask;
    int idx = 0;
    while (rm) {
      const int e = __ffs(rm) - 1;
      rm &= rm - 1;
      if ((idx++ & (kWarps - 1)) != warp) continue;
      const int env = env0 + e;
      if (lane < 24) p.dof_state[(size_t)env * 24 + lane] = s_dof[e * 24 + lane];
      if (lane < 13)
        p.root_states[((size_t)env * p.actors_per_env + p.root_actor_offset) * 13 + lane] = s_root[e * 13 + lane];
      if (p.zero_lstm_on_reset && p.sea_hidden_state) {
        // [2, N*12, 8]: per layer the env's 12 joints x 8 = 96 contiguous floats
        const size_t layer = (size_t)N * 96;
        float4* h0 = reinterpret_cast<float4*>(p.sea_hidden_state + (size_t)env * 96);
        float4* h1 = reinterpret_cast<float4*>(p.sea_hidden_state + layer + (size_t)env * 96);
        float4* c0 = reinterpret_cast<float4*>(p.sea_cell_state + (size_t)env * 96);
        float4* c1 = reinterpret_cast<float4*>(p.sea_cell_state + layer + (size_t)env * 96);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < 24) { h0[lane] = z; h1[lane] = z; c0[lane] = z; c1[lane] = z; }
      }
    }

    // ---------------- observation rows (LR:212-230 + clip LR:100-101): warp w takes envs w, w+4, ...
    // Lane l owns columns l + 32m; its noise scales are loop-invariant and live in registers.
    constexpr int kMaxSuper = 3;                       // O <= 48 + 256
    const bool hcols = p.measure_heights != 0;
    const bool noisy = p.add_noise != 0;
    const float clip = p.clip_obs, vs = p.vertical_scale, hsc = p.obs_scale_height;
    float nz[kMaxSuper][4];
#pragma unroll
    for (int sc = 0; sc < kMaxSuper; ++sc)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = sc * 128 + 32 * k + lane;
        nz[sc][k] = (noisy && j < O) ? __ldg(p.noise_scale_vec + j) : 0.f;
      }
    for (int e = warp; e < nval; e += kWarps) {
      const int env = env0 + e;
      const uint32_t genv = (uint32_t)(p.env_id_offset + env);
      const float rz = s_root[e * 13 + 2] - 0.5f;                 // post-reset root z, LR:225
      float* orow = p.obs_buf + (size_t)env * O;
      const float* hrow = p.measured_heights + (size_t)env * P;   // split mode only (PRE ran in another launch)
#pragma unroll
      for (int sc = 0; sc < kMaxSuper; ++sc) {
        if (sc * 128 < O) {
          U4 r = U4{0, 0, 0, 0};
          if (noisy) r = rng_block(key, genv, LGK_STREAM_OBS, (uint32_t)(32 * sc + lane));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int j = sc * 128 + 32 * k + lane;
            if (j < O) {
              float v;
              if (sc == 0 && k < 2 && j < 48) {
                v = s_head[e * 48 + j];
              } else {
                float h = 0.f;
                if (hcols) h = pre ? f_mul((float)s_h16[e * HS + (j - 48)], vs) : hrow[j - 48];
                v = hcols ? clampf(rz - h, -1.f, 1.f) * hsc : 0.f;
              }
              v = v + (2.0f * u32_to_uniform(pick(r, k)) - 1.0f) * nz[sc][k];
              orow[j] = clampf(v, -clip, clip);
            }
          }
        }
      }
    }
  }
  if (bulk && tid == 0) bulk_wait_read0();
}

// ------------------------------------------------------------------ reset_idx on an explicit id list
__global__ void __launch_bounds__(128) reset_idx_kernel(const __grid_constant__ LgkStepParams p,
                                                       const int64_t* __restrict__ ids, int n) {
  // one warp per listed env: lane 0 performs the per-env reset on a small shared scratch row set, then the
  // warp writes rows back cooperatively (same write path as the fused kernel).
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
// Single CTA.  Each thread owns 16 consecutive envs per sweep (one 16-byte load of reset_buf), so a sweep
// covers 16384 envs; ids come out in ascending order like reset_buf.nonzero() (LR:128).
__global__ void __launch_bounds__(1024) finalize_kernel(const __grid_constant__ LgkStepParams p, int32_t* reset_ids,
                                                        int32_t* reset_count, float* episode_means,
                                                        uint8_t* time_outs_extras, int advance) {
  __shared__ int s_warp[32];
  __shared__ int s_base, s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.num_envs;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int start = 0; start < N; start += 16384) {
    const int first = start + tid * 16;
    uint8_t flags[16];
    int c = 0;
    if (first + 16 <= N) {
      *reinterpret_cast<uint4*>(flags) = *reinterpret_cast<const uint4*>(p.reset_buf + first);
#pragma unroll
      for (int i = 0; i < 16; ++i) c += flags[i] != 0;
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) { flags[i] = (first + i < N) ? p.reset_buf[first + i] : 0; c += flags[i] != 0; }
    }
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int v = s_warp[lane];
      int wi = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
      s_warp[lane] = wi - v;
      if (lane == 31) s_total = wi;
    }
    __syncthreads();
    int off = s_base + s_warp[warp] + incl - c;
    if (reset_ids && c) {
#pragma unroll
      for (int i = 0; i < 16; ++i) if (flags[i]) reset_ids[off++] = first + i;
    }
    __syncthreads();
    if (tid == 0) s_base += s_total;
    __syncthreads();
  }
  const int count = s_base;
  const int ns = p.num_reward_slots;
  const int step_eff = p.step_counter_dev ? (*p.step_counter_dev + (advance ? 1 : 0)) : p.step;
  float* cur = p.reset_stats + (size_t)(step_eff & 1) * (ns + 2);
  float* other = p.reset_stats + (size_t)((step_eff + 1) & 1) * (ns + 2);
  if (tid == 0 && reset_count) *reset_count = count;
  if (count > 0) {     // the reference refreshes extras only inside reset_idx with a non-empty id list
    if (episode_means) {
      for (int k = tid; k < ns; k += 1024) episode_means[k] = cur[k] / (float)count / p.max_episode_length_s;
      if (tid == 0) episode_means[ns] = p.terrain_curriculum ? cur[ns + 1] / (float)N : 0.f;
    }
    if (p.send_timeouts && time_outs_extras)
      for (int i = tid; i < N; i += 1024) time_outs_extras[i] = p.time_out_buf[i];
  }
  for (int k = tid; k < ns + 2; k += 1024) other[k] = 0.f;
  if (advance && p.step_counter_dev && tid == 0) *p.step_counter_dev = step_eff;
}

}  // namespace lgk

// ====================================================================== C ABI
using namespace lgk;

static int validate_step(const LgkStepParams* p) {
  LGK_REQUIRE(p != nullptr, "params is null");
  LGK_REQUIRE(p->num_envs > 0, "num_envs must be positive");
  LGK_REQUIRE(p->num_bodies > 0 && p->num_bodies <= LGK_MAX_BODIES, "num_bodies out of range");
  LGK_REQUIRE(p->num_feet >= 0 && p->num_feet <= LGK_MAX_FEET, "num_feet out of range");
  LGK_REQUIRE(p->num_pen >= 0 && p->num_pen <= LGK_MAX_PEN, "num_pen out of range");
  LGK_REQUIRE(p->num_term >= 0 && p->num_term <= LGK_MAX_TERM, "num_term out of range");
  LGK_REQUIRE(p->actors_per_env >= 1 && p->root_actor_offset >= 0 && p->root_actor_offset < p->actors_per_env, "bad actor layout");
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

template <int kTile>
static int launch_post_physics(const LgkStepParams* p, cudaStream_t st) {
  const TileLayout L = make_layout(kTile, p->num_bodies, p->num_feet, p->num_height_points);
  static int smem_set = 0;
  if (L.total > smem_set) {
    if (int rc = check_cuda(cudaFuncSetAttribute(post_physics_kernel<kTile>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total),
                            "cudaFuncSetAttribute(post_physics_kernel)")) return rc;
    smem_set = L.total;
  }
  const int tiles = (p->num_envs + kTile - 1) / kTile;
  post_physics_kernel<kTile><<<tiles, kThreads, L.total, st>>>(*p);
  count_launch();
  return check_cuda(cudaGetLastError(), "post_physics_kernel launch");
}

extern "C" int lgk_post_physics(const LgkStepParams* p, void* stream) {
  if (int rc = validate_step(p)) return rc;
  LGK_REQUIRE((p->phase_mask & (LGK_PHASE_PRE | LGK_PHASE_POST)) != 0, "phase_mask selects nothing");
  LGK_REQUIRE(p->num_height_points <= 256 && p->num_obs <= 384, "at most 256 height points");
  int tile = p->tile_envs;
  if (tile == 0) tile = p->num_envs <= 16384 ? 8 : 16;     // small batches: more, shorter CTAs (latency-bound regime)
  switch (tile) {
    case 8: return launch_post_physics<8>(p, (cudaStream_t)stream);
    case 16: return launch_post_physics<16>(p, (cudaStream_t)stream);
    case 32: return launch_post_physics<32>(p, (cudaStream_t)stream);
    default: return set_error(LGK_ERR_ARG, "tile_envs must be 0 (auto), 8, 16 or 32");
  }
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
  finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(*p, reset_ids, reset_count, episode_means, time_outs_extras, advance);
  count_launch();
  return check_cuda(cudaGetLastError(), "finalize_kernel launch");
}
