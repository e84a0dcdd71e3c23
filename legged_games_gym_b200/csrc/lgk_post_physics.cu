// lgk_post_physics.cu -- post-physics step (reference LR:106-230, 329-508, 831-969) for sm_100a, one kernel per phase:
//
//  K1  post_kernel<0>       one CTA (4 warps) per tile of 32 consecutive envs, lane = env, warp = role.  Every reference
//      tensor is env-major row-major, so the tile's slice of root_states / dof_state / contact_forces / actions / torques /
//      last_actions / last_dof_vel / commands / feet_air_time is ONE contiguous chunk per tensor: each arrives by a single
//      TMA bulk copy (cp.async.bulk.shared::cluster.global, mbarrier completion) and whole-tile results (commands,
//      feet_air_time, last_*, base_*) leave by bulk stores.  Phase A: every role reduces its three joints, its foot and
//      its share of the penalised / termination bodies to partial sums.  Phase B: role 0 does the once-per-env work:
//      rotations, command resampling / heading, push, termination, the reward terms in the reference's alphabetical
//      order, reset of root / commands / terrain level.  Phase C: every role finishes its joints (reset draw, the 48
//      proprioceptive observation columns un-noised, histories).  Reset rows + LSTM-state zeroing are written
//      cooperatively; cross-env sums for extras["episode"] use warp shuffles + one atomicAdd per tile.
//  K2  scan_obs_fast_kernel (specialised, branch-free) / scan_obs_kernel (generic)   one WARP per env, lanes over
//      columns: the 187-point height scan (packed f32x2 op-exact index path, all int16 gathers of the env issued back
//      to back from the precomputed min3 field), measured_heights, and the finished observation row: height columns,
//      in-kernel Philox noise, clip -- lane l owns columns l+32m, so one Philox block serves four coalesced 128-byte
//      row segments.  Persistent CTAs; point grid, noise scales and column masks live in registers.
//
// lgk_post_physics orders them (K1 then K2; when the base_height reward is active the scan runs first) and runs the
// PRE / POST phases of K1 separately when Python code has to run in between.
//
//  Fused variant (opt-in, lgk_set_fused): post_kernel<G, RECIP> with G > 0 adds 1..8 scan warps to the K1 CTA; they run
//  K2's arithmetic for the tile's 32 envs concurrently with the role warps (named barriers; speculative height columns with
//  the pre-reset root z, redone by a role warp for the rare reset env; the 48-column head never leaves shared memory).
//  Bit-identical to the chain, one launch less, but measured slower on B200.
#include <stdlib.h>
#include <type_traits>
#include "lgk_step_device.cuh"

namespace lgk {

constexpr int kTile = 32;          // envs per K1 warp

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

// 16-byte asynchronous copies global -> shared (LDGSTS): every thread of the CTA moves pieces of the tile's contiguous
// chunks, no register staging, completion by cp.async.wait_group
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void stage_in16(void* dst_smem, const void* src, int bytes, int tid, int nthreads) {
  const int n16 = bytes >> 4;
  for (int i = tid; i < n16; i += nthreads)
    cp_async16(reinterpret_cast<uint8_t*>(dst_smem) + 16 * i, reinterpret_cast<const uint8_t*>(src) + 16 * i);
}
__device__ __forceinline__ void stage_out16(void* dst, const void* src_smem, int bytes, int tid, int nthreads) {
  const int n16 = bytes >> 4;
  for (int i = tid; i < n16; i += nthreads)
    reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src_smem)[i];
}
#ifndef LGK_K1_STAGE
#define LGK_K1_STAGE 0     // 0: TMA bulk copies in and out; 1: cp.async in, bulk out; 2: cp.async in, vector stores out
#endif

// ------------------------------------------------------------------ K1 shared-memory carve-up (bytes, 16-aligned)
struct TileLayout {
  int root, dof, contact, act, tq, lact, ldv, cmd, fat, lc, head, blv, bav, pg, lrv, frame, sums, ep, rew, flags, part, noise, misc, total;
};

__host__ __device__ inline int al16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline TileLayout make_layout(int nb, int nfeet, int nslots, bool fused = false) {
  TileLayout L;
  int o = 0;
  L.root = o;    o += al16(kTile * 13 * 4);
  L.dof = o;     o += al16(kTile * 24 * 4);
  // the 48-column observation head (row stride 49: conflict-free lane = env writes) reuses the contact tile, which
  // is dead once env_pre has run (a __syncwarp separates the two uses)
  { const int c = al16(kTile * nb * 3 * 4), h = al16(kTile * 49 * 4); L.contact = o; L.head = o; o += c > h ? c : h; }
  L.act = o;     o += al16(kTile * 12 * 4);
  L.tq = o;      o += al16(kTile * 12 * 4);
  L.lact = o;    o += al16(kTile * 12 * 4);
  L.ldv = o;     o += al16(kTile * 12 * 4);
  L.cmd = o;     o += al16(kTile * 4 * 4);
  L.fat = o;     o += al16(kTile * (nfeet > 0 ? nfeet : 1) * 4);
  L.lc = o;      o += al16(kTile * (nfeet > 0 ? nfeet : 1));
  L.blv = o;     o += al16(kTile * 3 * 4);
  L.bav = o;     o += al16(kTile * 3 * 4);
  L.pg = o;      o += al16(kTile * 3 * 4);
  L.lrv = o;     o += al16(kTile * 6 * 4);
  L.frame = o;   o += al16(kTile * 8 * 4);
  L.sums = o;    o += al16((nslots > 0 ? nslots : 1) * kTile * 4);   // episode_sums rows of the tile: [K][32]
  L.ep = o;      o += al16(kTile * 8);                                // episode_length_buf (int64)
  L.rew = o;     o += al16(kTile * 4);
  L.flags = o;   o += al16(kTile * 2);                                // reset flags [32] then time_out flags [32]
  L.part = o;    o += al16(PS_COUNT * 3 * kTile * 4);                 // partial sums of roles 1..3: [slot][role-1][env]
  L.noise = o;   o += fused ? al16(kTile * 48 * 4) : 0;             // fused kernel: 2u-1 of the 48 head columns, from the scan warps
  L.misc = o;    o += 16;   // mbarrier
  L.total = o;
  return L;
}

// used only by the partial-tile fallback path: kept out of line and rolled so the hot path stays compact
__device__ __noinline__ void copy_f32(float* dst, const float* src, int n, int tid) {
#pragma unroll 1
  for (int i = tid; i < n; i += 4 * kTile) dst[i] = src[i];
}
__device__ __noinline__ void copy_u8(uint8_t* dst, const uint8_t* src, int n, int tid) {
#pragma unroll 1
  for (int i = tid; i < n; i += 4 * kTile) dst[i] = src[i];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// observation noise (LR:229-230): obs + (2u - 1) * noise_scale, as ONE fused multiply-add on the individually rounded
// observation -- written out so that every kernel variant rounds identically
__device__ __forceinline__ float noise_unit(uint32_t word) { return f_fma(2.0f, u32_to_uniform(word), -1.0f); }
__device__ __forceinline__ float noisy_obs(float v, uint32_t word, float scale) { return f_fma(noise_unit(word), scale, v); }

// scan frame of one env: [zn, wn, root_x, root_y] (pre-reset yaw frame, LR:853-854) + [root_z_post_reset - 0.5] (LR:225)
constexpr int kFrameFloats = 8;

// ------------------------------------------------------------------ K1
// CTA = 4 warps = one tile of 32 envs, lane = env, warp = role (see lgk_step_device.cuh): phase A every role reduces its
// joints / foot / bodies to partial sums, phase B role 0 does the once-per-env work, phase C every role finishes its
// joints (reset, observation columns, histories).  Total work per tile is about a third of running all four lanes of an
// env through everything, and the chain on the critical path about half of that of one thread doing a whole env.
constexpr int kK1Threads = 4 * kTile;
__device__ long long* g_k1_timeline = nullptr;     // profiling hook (lgk_step_debug_timeline): stamps of CTA 0
__device__ __forceinline__ void k1_stamp(int slot) {
  if (g_k1_timeline != nullptr && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_k1_timeline[blockIdx.x == 0 ? slot : 16 + slot] = (long long)t;      // [16..24]: the same stamps of the LAST CTA
  }
}

// experiments (-DLGK_EXP_CTA_STAMPS): every CTA records stamps 0..6 and its SM (slot 7) in buf[32 + 8 * blockIdx.x + slot]
__device__ __forceinline__ void cta_stamp(int which) {
#ifdef LGK_EXP_CTA_STAMPS
  if (g_k1_timeline != nullptr && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_k1_timeline[32 + 8 * blockIdx.x + which] = (long long)t;
    if (which == 0) { uint32_t sm; asm volatile("mov.u32 %0, %smid;" : "=r"(sm)); g_k1_timeline[32 + 8 * blockIdx.x + 7] = sm; }
  }
#endif
}

__device__ __forceinline__ void scan_stamp(int slot) {      // first scan warp of CTA 0
  if (g_k1_timeline != nullptr && blockIdx.x == 0 && threadIdx.x == kK1Threads) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_k1_timeline[slot] = (long long)t;
  }
}

// named barriers: 1 = the four role warps (what __syncthreads() is to the unfused kernel), 2 = scan warps -> role warps
// hand-over of the fused kernel (scan warps arrive, role warps wait)
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void role_sync() { named_sync(1, kK1Threads); }

template <int G, bool RECIP>
__device__ __forceinline__ void scan_tile(const LgkStepParams& p, const RngKey& key, int env0, int nval, int e0, int e1,
                                          int de, int lane, float* s_noise);
LGK_COLD void refresh_reset_height_obs(const LgkStepParams& p, const RngKey& key, int env, float rz, int lane);

// G == 0: K1 alone (128 threads).  G > 0: the fused post-physics kernel -- warps 0..3 are the role warps of K1, warps
// 4.. are scan warps that run K2's work for the same 32 envs concurrently (the role warps' dependent chain leaves the
// issue slots the scan warps need); the un-noised observation head never leaves shared memory.
template <int G, bool RECIP>
#ifndef LGK_K1_MINBLOCKS
#define LGK_K1_MINBLOCKS 7
#endif
__global__ void __launch_bounds__(G > 0 ? kK1Threads + 256 : kK1Threads, G > 0 ? 2 : LGK_K1_MINBLOCKS)
post_kernel(const __grid_constant__ LgkStepParams p) {
  constexpr bool FUSED = G > 0;
  extern __shared__ __align__(128) uint8_t smem[];
  const TileLayout L = make_layout(p.num_bodies, p.num_feet, p.num_reward_slots, FUSED);
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
  float* s_frame = reinterpret_cast<float*>(smem + L.frame);
  float* s_sums = reinterpret_cast<float*>(smem + L.sums);
  long long* s_ep = reinterpret_cast<long long*>(smem + L.ep);
  float* s_rew = reinterpret_cast<float*>(smem + L.rew);
  uint8_t* s_flags = smem + L.flags;
  float* s_part = reinterpret_cast<float*>(smem + L.part);
  const int K = p.num_reward_slots;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.misc);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int env0 = blockIdx.x * kTile;
  const int nval = min(kTile, p.num_envs - env0);
  const int NB = p.num_bodies, F = p.num_feet, P = p.num_height_points, O = p.num_obs, N = p.num_envs;
  const bool pre = (p.phase_mask & LGK_PHASE_PRE) != 0, post = (p.phase_mask & LGK_PHASE_POST) != 0;
  const bool fat_active = p.reward_active[LGK_R_FEET_AIR_TIME] != 0 && F > 0;
  // bulk (TMA) path needs a full tile (sizes are then multiples of 16 B) and unit root stride
  const bool bulk = (nval == kTile) && (p.actors_per_env == 1) && (p.num_envs % 4 == 0);   // 16-B aligned rows
  pdl_launch_dependents();
  k1_stamp(0);
  cta_stamp(0);
  if (bulk && tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();              // everything below reads state written by the previous kernels of the step
  cta_stamp(2);
  if constexpr (FUSED) if (warp >= 4) {        // scan warps: heights + height observation columns + the noise of the whole row
    scan_stamp(9);
    const int step_s = p.step_counter_dev ? (*p.step_counter_dev + 1) : p.step;
    scan_tile<G, RECIP>(p, make_key(p.seed, step_s), env0, nval, warp - 4, nval, (int)(blockDim.x >> 5) - 4, lane,
                        reinterpret_cast<float*>(smem + L.noise));
    scan_stamp(11);
    __threadfence_block();
    named_arrive(2, blockDim.x);
    return;
  }

  // ---------------- stage the tile
#if LGK_K1_STAGE >= 1
#ifdef LGK_EXP_NO_LOADS
  if (bulk && p.num_envs < 0) {
#else
  if (bulk) {
#endif
    stage_in16(s_root, p.root_states + (size_t)env0 * 13, kTile * 13 * 4, tid, kK1Threads);
    stage_in16(s_dof, p.dof_state + (size_t)env0 * 24, kTile * 24 * 4, tid, kK1Threads);
    stage_in16(s_contact, p.contact_forces + (size_t)env0 * NB * 3, kTile * NB * 3 * 4, tid, kK1Threads);
    stage_in16(s_act, p.actions + (size_t)env0 * 12, kTile * 12 * 4, tid, kK1Threads);
    stage_in16(s_tq, p.torques + (size_t)env0 * 12, kTile * 12 * 4, tid, kK1Threads);
    stage_in16(s_lact, p.last_actions + (size_t)env0 * 12, kTile * 12 * 4, tid, kK1Threads);
    stage_in16(s_ldv, p.last_dof_vel + (size_t)env0 * 12, kTile * 12 * 4, tid, kK1Threads);
    stage_in16(s_cmd, p.commands + (size_t)env0 * 4, kTile * 4 * 4, tid, kK1Threads);
    if (F > 0) {
      stage_in16(s_fat, p.feet_air_time + (size_t)env0 * F, kTile * F * 4, tid, kK1Threads);
      stage_in16(s_lc, p.last_contacts + (size_t)env0 * F, kTile * F, tid, kK1Threads);
    }
    cp_async_commit();
  }
#else
  if (bulk) {
    if (tid == 0) {
      uint32_t bytes = kTile * (13 + 24 + 3 * NB + 12 + 12 + 12 + 12 + 4) * 4;
      if (F > 0) bytes += kTile * F * 4 + kTile * F;
      mbar_expect_tx(bar, bytes);
      bulk_g2s(s_root, p.root_states + (size_t)env0 * 13, kTile * 13 * 4, bar);
      bulk_g2s(s_dof, p.dof_state + (size_t)env0 * 24, kTile * 24 * 4, bar);
      bulk_g2s(s_contact, p.contact_forces + (size_t)env0 * NB * 3, kTile * NB * 3 * 4, bar);
      bulk_g2s(s_act, p.actions + (size_t)env0 * 12, kTile * 12 * 4, bar);
      bulk_g2s(s_tq, p.torques + (size_t)env0 * 12, kTile * 12 * 4, bar);
      bulk_g2s(s_lact, p.last_actions + (size_t)env0 * 12, kTile * 12 * 4, bar);
      bulk_g2s(s_ldv, p.last_dof_vel + (size_t)env0 * 12, kTile * 12 * 4, bar);
      bulk_g2s(s_cmd, p.commands + (size_t)env0 * 4, kTile * 4 * 4, bar);
      if (F > 0) {
        bulk_g2s(s_fat, p.feet_air_time + (size_t)env0 * F, kTile * F * 4, bar);
        bulk_g2s(s_lc, p.last_contacts + (size_t)env0 * F, kTile * F, bar);
      }
    }
  }
#endif
  // per-env scalars (one 128-byte row segment per tensor and tile) come by plain coalesced loads, lane = env, issued while
  // the bulk copies are in flight: 20 fewer bulk operations per tile (measured neutral for the kernel time)
  if (bulk) {
    for (int k = warp; k < K; k += 4) s_sums[k * kTile + lane] = p.episode_sums[(size_t)k * N + env0 + lane];
    if (warp == (K & 3)) s_ep[lane] = p.episode_length_buf[env0 + lane];
  }
  // the device step counter is read while the tile loads are in flight
  const int step_eff = p.step_counter_dev ? (*p.step_counter_dev + 1) : p.step;
  const bool do_push = p.step_counter_dev ? (p.push_interval > 0 && step_eff % p.push_interval == 0) : (p.do_push != 0);
  const RngKey key = make_key(p.seed, step_eff);
  k1_stamp(1);
  if (bulk) {
#if LGK_K1_STAGE >= 1
    cp_async_wait0();
    role_sync();
#else
    mbar_wait(bar, 0);
#endif
  } else {
    for (int i = tid; i < nval * 13; i += kK1Threads) {
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
    for (int i = tid; i < K * nval; i += kK1Threads) {
      const int k = i / nval, e = i - k * nval;
      s_sums[k * kTile + e] = p.episode_sums[(size_t)k * N + env0 + e];
    }
    if (tid < nval) s_ep[tid] = p.episode_length_buf[env0 + tid];
    role_sync();
  }
  k1_stamp(2);
  cta_stamp(3);

  // ---------------- phase A: every role, its share of the per-joint / per-foot / per-body terms
  const int e = lane, role = warp, env = env0 + e;
  const bool valid = e < nval;
  const uint32_t genv = (uint32_t)(p.env_id_offset + env);
  float* root = s_root + e * 13;
  float* dof = s_dof + e * 24;
  float* cmd = s_cmd + e * 4;
  float* fat = s_fat + e * F;
  uint8_t* lc = s_lc + e * F;
  float* sums = s_sums + e;                 // tile-local episode sums, row stride kTile
  RolePartials mine;
#ifdef LGK_EXP_SKIP_A
#pragma unroll
  for (int k = 0; k < PS_COUNT; ++k) mine.v[k] = 0.f;
  if (false) {
#else
  if (pre) {
#endif
    const float* hrow = (valid && p.measure_heights && p.reward_active[LGK_R_BASE_HEIGHT])
                            ? p.measured_heights + (size_t)env * P : nullptr;       // the scan ran before this kernel
    role_partials(p, role, dof, s_contact + e * NB * 3, s_act + e * 12, s_tq + e * 12, s_lact + e * 12, s_ldv + e * 12,
                  fat, lc, root[2], hrow, mine);
    if (role > 0) {
#pragma unroll
      for (int k = 0; k < PS_COUNT; ++k) s_part[(k * 3 + role - 1) * kTile + e] = mine.v[k];
    }
  }
  role_sync();
  k1_stamp(3);

  // ---------------- phase B: role 0, everything that exists once per env
  EnvScalars s;
  s.reset = false; s.time_out = false; s.rew = 0.f; s.ep_len = 0;
  const bool want_frames = !FUSED && p.measure_heights && !p.terrain_is_plane && p.scan_frames != nullptr;
#ifdef LGK_EXP_SKIP_B
  if (false) {
#else
  if (role == 0) {
#endif
    if (pre) {
      uint32_t bits = __float_as_uint(mine.v[PS_BITS]);
      uint32_t any = bits & 5u, feet_down = (bits >> 1) & 1u;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int k = 0; k < PS_COUNT; ++k) {
          const float v = s_part[(k * 3 + r) * kTile + e];
          if (k == PS_BITS) { const uint32_t b = __float_as_uint(v); any |= b & 5u; feet_down += (b >> 1) & 1u; }
          else mine.v[k] += v;
        }
      }
      mine.v[PS_BITS] = __uint_as_float(any | (feet_down << 8));
      if (want_frames) {     // pre-reset yaw frame for K2 (LR:853-854 uses the pose of THIS step before reset_idx)
        const YawFrame yf = yaw_frame(root[5], root[6], root[0], root[1]);
        *reinterpret_cast<float4*>(s_frame + e * kFrameFloats) = make_float4(yf.zn, yf.wn, yf.rx, yf.ry);
      }
      if (p.base_quat != nullptr && valid)      // LowLevelGame.base_quat: the pose the rotations below use, kept past reset_idx
        *reinterpret_cast<float4*>(p.base_quat + (size_t)env * 4) = make_float4(root[3], root[4], root[5], root[6]);
      env_finish(p, do_push, key, genv, root, cmd, sums, kTile, s_ep[e], mine, s);
      s_blv[3 * e] = s.blv.x; s_blv[3 * e + 1] = s.blv.y; s_blv[3 * e + 2] = s.blv.z;
      s_bav[3 * e] = s.bav.x; s_bav[3 * e + 1] = s.bav.y; s_bav[3 * e + 2] = s.bav.z;
      s_pg[3 * e] = s.pg.x; s_pg[3 * e + 1] = s.pg.y; s_pg[3 * e + 2] = s.pg.z;
    } else {   // split mode: PRE ran in an earlier launch, pick its results up from global memory
      const int envc = valid ? env : env0;
      s.blv = V3{p.base_lin_vel[3 * envc], p.base_lin_vel[3 * envc + 1], p.base_lin_vel[3 * envc + 2]};
      s.bav = V3{p.base_ang_vel[3 * envc], p.base_ang_vel[3 * envc + 1], p.base_ang_vel[3 * envc + 2]};
      s.pg = V3{p.projected_gravity[3 * envc], p.projected_gravity[3 * envc + 1], p.projected_gravity[3 * envc + 2]};
      s.ep_len = s_ep[e];
      s.reset = p.reset_buf[envc] != 0;
      s.time_out = p.time_out_buf[envc] != 0;
      s.rew = p.rew_buf[envc];
    }
    const bool resetting = post && valid && s.reset;
    if (post) {
      s.rew = env_finish_reward(p, s.rew, s.reset, s.time_out, sums, kTile);
      // cross-env sums for extras["episode"] over the envs that reset, using their PRE-reset episode sums (LR:179-183)
      float* stats = p.reset_stats + (size_t)(step_eff & 1) * (p.num_reward_slots + 2);
      const uint32_t rmask = __ballot_sync(0xffffffffu, resetting);
      if (rmask != 0) {
        for (int k = 0; k < p.num_reward_slots; ++k) {
          float v = 0.f;
          if (resetting) { v = sums[k * kTile]; sums[k * kTile] = 0.f; }
          v = warp_sum(v);
          if (lane == 0) atomicAdd(stats + k, v);
        }
        if (lane == 0) atomicAdd(stats + p.num_reward_slots, (float)__popc(rmask));
      }
      if (resetting) { env_reset_base(p, key, genv, env, root, cmd); s.ep_len = 0; }      // LR:176
      if (p.terrain_curriculum) {                       // mean terrain level over ALL envs (LR:186), after the level updates
        float lv = valid ? (float)p.terrain_levels[env] : 0.f;
        lv = warp_sum(lv);
        if (lane == 0) atomicAdd(stats + p.num_reward_slots + 1, lv);
      }
    }
    s_rew[e] = s.rew;
    s_ep[e] = s.ep_len;
    s_flags[e] = s.reset ? 1 : 0;
    s_flags[kTile + e] = s.time_out ? 1 : 0;
  }
  role_sync();           // contact rows are dead from here on: their region becomes the observation head
  k1_stamp(4);

  // ---------------- phase C: every role finishes its joints
  const bool reset_e = post && valid && s_flags[e] != 0;
#ifdef LGK_EXP_SKIP_C
  if (false) {
#else
  if (post) {
#endif
    if (reset_e) env_reset_joints(p, key, genv, role, dof, fat);
    env_obs_head_role(p, role, dof, s_act + e * 12, s_head + e * 49);
    if (role == 0) env_obs_head_base(p, s, cmd, s_head + e * 49);
#pragma unroll
    for (int d = 3 * role; d < 3 * role + 3; ++d) s_ldv[e * 12 + d] = dof[2 * d + 1];   // LR:133 (post-reset dof_vel)
    if (role == 1) { for (int i = 0; i < 6; ++i) s_lrv[e * 6 + i] = root[7 + i]; }      // LR:134 (post push/reset)
    if (!FUSED && role == 2 && valid && p.scan_frames) p.scan_frames[(size_t)env * kFrameFloats + 4] = root[2] - 0.5f;   // post-reset z
  }
  fence_async_smem();
  role_sync();
  k1_stamp(5);
  cta_stamp(4);
#ifdef LGK_EXP_NO_STORES
  if (p.num_envs > 0) { cta_stamp(1); return; }
#endif

  // ---------------- whole-tile write-backs
#ifndef LGK_EXP_STORE_MASK
#define LGK_EXP_STORE_MASK 15      // experiments: 1 tile vectors, 2 per-env scalar rows, 4 scan frames, 8 obs head
#endif
#if LGK_K1_STAGE >= 2
  if (bulk) {
    if (pre && (LGK_EXP_STORE_MASK & 1)) {
      stage_out16(p.base_lin_vel + (size_t)env0 * 3, s_blv, kTile * 3 * 4, tid, kK1Threads);
      stage_out16(p.base_ang_vel + (size_t)env0 * 3, s_bav, kTile * 3 * 4, tid, kK1Threads);
      stage_out16(p.projected_gravity + (size_t)env0 * 3, s_pg, kTile * 3 * 4, tid, kK1Threads);
      if (do_push && (!post || !FUSED)) stage_out16(p.root_states + (size_t)env0 * 13, s_root, kTile * 13 * 4, tid, kK1Threads);
    }
    if (LGK_EXP_STORE_MASK & 1) stage_out16(p.commands + (size_t)env0 * 4, s_cmd, kTile * 4 * 4, tid, kK1Threads);
    if ((LGK_EXP_STORE_MASK & 1) && (fat_active || (post && F > 0))) {
      stage_out16(p.feet_air_time + (size_t)env0 * F, s_fat, kTile * F * 4, tid, kK1Threads);
      if (pre) stage_out16(p.last_contacts + (size_t)env0 * F, s_lc, kTile * F, tid, kK1Threads);
    }
    if (post && (LGK_EXP_STORE_MASK & 1)) {
      stage_out16(p.last_actions + (size_t)env0 * 12, s_act, kTile * 12 * 4, tid, kK1Threads);
      stage_out16(p.last_dof_vel + (size_t)env0 * 12, s_ldv, kTile * 12 * 4, tid, kK1Threads);
      stage_out16(p.last_root_vel + (size_t)env0 * 6, s_lrv, kTile * 6 * 4, tid, kK1Threads);
    }
    if (LGK_EXP_STORE_MASK & 2) {
    for (int k = warp; k < K; k += 4) p.episode_sums[(size_t)k * N + env0 + lane] = s_sums[k * kTile + lane];
    if (warp == (K & 3)) p.episode_length_buf[env0 + lane] = s_ep[lane];
    if (warp == ((K + 1) & 3)) p.rew_buf[env0 + lane] = s_rew[lane];
    if (pre && warp == ((K + 2) & 3)) { p.reset_buf[env0 + lane] = s_flags[lane]; p.time_out_buf[env0 + lane] = s_flags[kTile + lane]; }
    }
  } else if (false) {
#else
  if (bulk) {
#endif
    if (tid == 0) {
      if (pre) {
        bulk_s2g(p.base_lin_vel + (size_t)env0 * 3, s_blv, kTile * 3 * 4);
        bulk_s2g(p.base_ang_vel + (size_t)env0 * 3, s_bav, kTile * 3 * 4);
        bulk_s2g(p.projected_gravity + (size_t)env0 * 3, s_pg, kTile * 3 * 4);
        if (do_push && !post) bulk_s2g(p.root_states + (size_t)env0 * 13, s_root, kTile * 13 * 4);
      }
      bulk_s2g(p.commands + (size_t)env0 * 4, s_cmd, kTile * 4 * 4);
      if (fat_active || (post && F > 0)) {
        bulk_s2g(p.feet_air_time + (size_t)env0 * F, s_fat, kTile * F * 4);
        if (pre) bulk_s2g(p.last_contacts + (size_t)env0 * F, s_lc, kTile * F);
      }
      if (post) {
        bulk_s2g(p.last_actions + (size_t)env0 * 12, s_act, kTile * 12 * 4);
        bulk_s2g(p.last_dof_vel + (size_t)env0 * 12, s_ldv, kTile * 12 * 4);
        bulk_s2g(p.last_root_vel + (size_t)env0 * 6, s_lrv, kTile * 6 * 4);
        // (fused: the scan warps read root poses from global memory, so the pushed tile goes out after their barrier)
        if (do_push && pre && !FUSED) bulk_s2g(p.root_states + (size_t)env0 * 13, s_root, kTile * 13 * 4);
      }
      bulk_commit();
    }
    for (int k = warp; k < K; k += 4) p.episode_sums[(size_t)k * N + env0 + lane] = s_sums[k * kTile + lane];
    if (warp == (K & 3)) p.episode_length_buf[env0 + lane] = s_ep[lane];
    if (warp == ((K + 1) & 3)) p.rew_buf[env0 + lane] = s_rew[lane];
    if (pre && warp == ((K + 2) & 3)) { p.reset_buf[env0 + lane] = s_flags[lane]; p.time_out_buf[env0 + lane] = s_flags[kTile + lane]; }
  } else {
    if (pre) {
      copy_f32(p.base_lin_vel + (size_t)env0 * 3, s_blv, nval * 3, tid);
      copy_f32(p.base_ang_vel + (size_t)env0 * 3, s_bav, nval * 3, tid);
      copy_f32(p.projected_gravity + (size_t)env0 * 3, s_pg, nval * 3, tid);
    }
    copy_f32(p.commands + (size_t)env0 * 4, s_cmd, nval * 4, tid);
    for (int i = tid; i < K * nval; i += kK1Threads) {
      const int k = i / nval, ee = i - k * nval;
      p.episode_sums[(size_t)k * N + env0 + ee] = s_sums[k * kTile + ee];
    }
    if (tid < nval) {
      p.episode_length_buf[env0 + tid] = s_ep[tid];
      p.rew_buf[env0 + tid] = s_rew[tid];
      if (pre) { p.reset_buf[env0 + tid] = s_flags[tid]; p.time_out_buf[env0 + tid] = s_flags[kTile + tid]; }
    }
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
      for (int i = tid; i < nval * 13; i += kK1Threads) {
        const int ee = i / 13, c = i - ee * 13;
        if (c == 7 || c == 8)
          p.root_states[((size_t)(env0 + ee) * p.actors_per_env + p.root_actor_offset) * 13 + c] = s_root[i];
      }
    }
  }
  // scan frames (pose part): 4 floats per env, strided rows of 8
  if (pre && want_frames && valid && role == 0 && (LGK_EXP_STORE_MASK & 4)) {
    *reinterpret_cast<float4*>(p.scan_frames + (size_t)env * kFrameFloats) =
        *reinterpret_cast<const float4*>(s_frame + e * kFrameFloats);
  }
  k1_stamp(6);

  if (post) {
    // fused: from here on the scan warps' speculative height columns (pre-reset z), measured_heights and the head noise
    // are complete and visible
    if (FUSED) {
      named_sync(2, blockDim.x);
      if (bulk && tid == 0 && do_push) { bulk_s2g(p.root_states + (size_t)env0 * 13, s_root, kTile * 13 * 4); bulk_commit(); }
    }
    // ---------------- reset rows: dof_state / root_states write-back + LSTM state zeroing (ANY:56-60); every warp sees the
    // same reset mask (lane = env) and takes every fourth reset env
    uint32_t rm = __ballot_sync(0xffffffffu, reset_e);
    int turn = 0;
    while (rm) {
      const int ee = __ffs(rm) - 1;
      rm &= rm - 1;
      if ((turn++ & 3) != warp) continue;
      const int en = env0 + ee;
      // height columns of a reset env use the post-reset root z (LR:225 after LR:160)
      if (FUSED) refresh_reset_height_obs(p, key, en, s_root[ee * 13 + 2] - 0.5f, lane);
      if (lane < 24) p.dof_state[(size_t)en * 24 + lane] = s_dof[ee * 24 + lane];
      if (lane < 13)
        p.root_states[((size_t)en * p.actors_per_env + p.root_actor_offset) * 13 + lane] = s_root[ee * 13 + lane];
      if (p.zero_lstm_on_reset && p.sea_hidden_state) {
        // [2, N*12, 8]: per layer the env's 12 joints x 8 = 96 contiguous floats
        const size_t layer = (size_t)N * 96;
        float4* h0 = reinterpret_cast<float4*>(p.sea_hidden_state + (size_t)en * 96);
        float4* h1 = reinterpret_cast<float4*>(p.sea_hidden_state + layer + (size_t)en * 96);
        float4* c0 = reinterpret_cast<float4*>(p.sea_cell_state + (size_t)en * 96);
        float4* c1 = reinterpret_cast<float4*>(p.sea_cell_state + layer + (size_t)en * 96);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < 24) { h0[lane] = z; h1[lane] = z; c0[lane] = z; c1[lane] = z; }
      }
    }
    if (FUSED) {
      // ---------------- the 48 proprioceptive columns, finished: noise (the scan warps' uniforms) + clip (LR:100-101, 229-230)
      const float* s_noise = reinterpret_cast<const float*>(smem + L.noise);
      const bool noisy = p.add_noise != 0;
      const float nz0 = noisy ? __ldg(p.noise_scale_vec + lane) : 0.f;
      const float nz1 = (noisy && lane < 16) ? __ldg(p.noise_scale_vec + 32 + lane) : 0.f;
      const float clip = p.clip_obs;
      for (int ee = warp * 8; ee < min(warp * 8 + 8, nval); ++ee) {
        float* orow = p.obs_buf + (size_t)(env0 + ee) * O;
        const float v0 = f_fma(s_noise[ee * 48 + lane], nz0, s_head[ee * 49 + lane]);
        orow[lane] = clampf(v0, -clip, clip);
        if (lane < 16) {
          const float v1 = f_fma(s_noise[ee * 48 + 32 + lane], nz1, s_head[ee * 49 + 32 + lane]);
          orow[32 + lane] = clampf(v1, -clip, clip);
        }
      }
    } else if (!p.measure_heights) {
      // ---------------- no height columns (flat tasks): the row is finished here -- noise (one Philox block per env and
      // lane, the words K2 would use) + clip -- and lgk_post_physics launches no K2
      const bool noisy = p.add_noise != 0;
      const float nz0 = noisy ? __ldg(p.noise_scale_vec + lane) : 0.f;
      const float nz1 = (noisy && lane < 16) ? __ldg(p.noise_scale_vec + 32 + lane) : 0.f;
      const float clip = p.clip_obs;
      for (int ee = warp * 8; ee < min(warp * 8 + 8, nval); ++ee) {
        float* orow = p.obs_buf + (size_t)(env0 + ee) * O;
        U4 r = U4{0, 0, 0, 0};
        if (noisy) r = rng_block(key, (uint32_t)(p.env_id_offset + env0 + ee), LGK_STREAM_OBS, (uint32_t)lane);
        orow[lane] = clampf(noisy_obs(s_head[ee * 49 + lane], r.x, nz0), -clip, clip);
        if (lane < 16) orow[32 + lane] = clampf(noisy_obs(s_head[ee * 49 + 32 + lane], r.y, nz1), -clip, clip);
      }
    } else if (LGK_EXP_STORE_MASK & 8) {
      // ---------------- the 48 proprioceptive columns, un-noised (K2 adds noise + clip): coalesced row segments
      for (int ee = warp * 8; ee < min(warp * 8 + 8, nval); ++ee) {
        float* orow = p.obs_buf + (size_t)(env0 + ee) * O;
        orow[lane] = s_head[ee * 49 + lane];
        if (lane < 16) orow[32 + lane] = s_head[ee * 49 + 32 + lane];
      }
    }
  }
  k1_stamp(7);
#if LGK_K1_STAGE < 2
  if (bulk && tid == 0) bulk_wait_read0();
#endif
  k1_stamp(8);
  cta_stamp(1);
}

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
      const float* orow = p.obs_buf + (size_t)env * O;
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
      const float* orow = p.obs_buf + (size_t)env * O;
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

// ------------------------------------------------------------------ scan warps of the fused kernel
// K2's work for the 32 envs of one K1 tile, spread over the CTA's nsw scan warps (warp per env, lane = column, exactly
// the arithmetic of scan_obs_fast_kernel<G, kScan | kObs, RECIP>).  Nothing here waits for the role warps: the yaw frames
// come straight from root_states (pre-reset pose, LR:853-854), the height columns are finished with the pre-reset root z
// (the role warps redo the rare reset env, refresh_reset_height_obs), and of the 48 head columns only the uniforms are
// produced (2u-1, parked in shared memory for the role warps' final write).
template <int G, bool RECIP>
__device__ __forceinline__ void scan_tile(const LgkStepParams& p, const RngKey& key, int env0, int nval, int e0, int e1,
                                          int de, int lane, float* s_noise) {
  const int P = p.num_height_points, O = p.num_obs;
  f2_t Bv[G], Bsv[G];
  float nz[G];
  uint32_t pmask = 0, omask = 0;
  const bool noisy = p.add_noise != 0;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int j = 32 * g + lane, pt = j - 48;
    const bool isp = pt >= 0 && pt < P;
    float bx = 0.f, by = 0.f;
    if (isp) { const float2 b = __ldg(reinterpret_cast<const float2*>(p.height_points_xy) + pt); bx = b.x; by = b.y; }
    Bv[g] = pack2(bx, by); Bsv[g] = pack2(by, bx);
    pmask |= (isp ? 1u : 0u) << g;
    omask |= ((j < O && j >= 48) ? 1u : 0u) << g;
    nz[g] = (noisy && j < O) ? __ldg(p.noise_scale_vec + j) : 0.f;
  }
  const float rt_one = __int_as_float(0x3f800000u | ((uint32_t)p.num_envs >> 31));   // 1.0f the compiler cannot see
  const float clip = p.clip_obs, vs = p.vertical_scale, hsc = p.obs_scale_height;
  const float border = p.border_size, hscale = p.horizontal_scale, hrecip = p.horizontal_scale_recip;
  const int rows = p.hf_rows, cols = p.hf_cols;
  // lane l holds the frame of env l of the tile; the env loop broadcasts it by shuffle
  const float* r = p.root_states + ((size_t)(env0 + min(lane, nval - 1)) * p.actors_per_env + p.root_actor_offset) * 13;
  const YawFrame fl = yaw_frame(r[5], r[6], r[0], r[1]);
  const float rzl = r[2] - 0.5f;
  scan_stamp(10);
#pragma unroll 1
  for (int e = e0; e < e1; e += de) {
    const int env = env0 + e;
    const YawFrame2 yf = yaw_frame2(YawFrame{__shfl_sync(0xffffffffu, fl.zn, e), __shfl_sync(0xffffffffu, fl.wn, e),
                                             __shfl_sync(0xffffffffu, fl.rx, e), __shfl_sync(0xffffffffu, fl.ry, e)});
    const float rz = __shfl_sync(0xffffffffu, rzl, e);
    float* orow = p.obs_buf + (size_t)env * O;
    float* hrow = p.measured_heights + (size_t)env * P;
    float h[G];
    int off[G];
    h[0] = 0.f;
#pragma unroll
    for (int g = 1; g < G; ++g) {
      int ix, iy;
      height_index2<RECIP>(yf, Bv[g], Bsv[g], border, hscale, hrecip, rt_one, rows, cols, ix, iy);
      off[g] = ix * cols + iy;
    }
#pragma unroll
    for (int g = 1; g < G; ++g) h[g] = f_mul((float)__ldg(p.height_min3 + off[g]), vs);      // LR:869
#pragma unroll
    for (int g = 1; g < G; ++g) if ((pmask >> g) & 1u) hrow[32 * g + lane - 48] = h[g];
    const uint32_t genv = (uint32_t)(p.env_id_offset + env);
#pragma unroll
    for (int sc = 0; sc < (G + 3) / 4; ++sc) {
      U4 rr = U4{0, 0, 0, 0};
      if (noisy) rr = rng_block(key, genv, LGK_STREAM_OBS, (uint32_t)(32 * sc + lane));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int g = sc * 4 + k;
        if (g < G) {
          const float t = noise_unit(pick(rr, k));
          if (g == 0) s_noise[e * 48 + lane] = t;
          if (g == 1 && lane < 16) s_noise[e * 48 + 32 + lane] = t;
          if (g >= 1) {
            const float v = f_fma(t, nz[g], f_mul(clampf(rz - h[g], -1.f, 1.f), hsc));
            if ((omask >> g) & 1u) orow[32 * g + lane] = clampf(v, -clip, clip);
          }
        }
      }
    }
  }
}

// height columns of an env that reset this step, redone with its post-reset root z (cold: a warp of the role group per env)
LGK_COLD void refresh_reset_height_obs(const LgkStepParams& p, const RngKey& key, int env, float rz, int lane) {
  const int P = p.num_height_points, O = p.num_obs;
  const bool noisy = p.add_noise != 0;
  const uint32_t genv = (uint32_t)(p.env_id_offset + env);
  const float* hrow = p.measured_heights + (size_t)env * P;
  float* orow = p.obs_buf + (size_t)env * O;
#pragma unroll 1
  for (int sc = 0; sc * 128 < O; ++sc) {
    U4 rr = U4{0, 0, 0, 0};
    if (noisy) rr = rng_block(key, genv, LGK_STREAM_OBS, (uint32_t)(32 * sc + lane));
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const int j = 128 * sc + 32 * k + lane;
      if (j >= 48 && j < O) {
        const float v = f_fma(noise_unit(pick(rr, k)), noisy ? p.noise_scale_vec[j] : 0.f,
                              f_mul(clampf(rz - hrow[j - 48], -1.f, 1.f), p.obs_scale_height));
        orow[j] = clampf(v, -p.clip_obs, p.clip_obs);
      }
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
// every load is one aligned 16-byte vector), counts it, one block-wide exclusive scan, then re-reads its (L1-resident)
// range and emits the ids -- ascending like reset_buf.nonzero() (LR:128).
__device__ __forceinline__ int count_flags16(uint4 v) {
  // flags are 0/1 bytes (bool tensors): the byte sum is the popcount
  return __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
}

__global__ void __launch_bounds__(1024) finalize_kernel(const __grid_constant__ LgkStepParams p, int32_t* reset_ids,
                                                        int32_t* reset_count, float* episode_means,
                                                        uint8_t* time_outs_extras, int advance) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.num_envs;
  pdl_launch_dependents();
  pdl_wait();
  const int chunk = (((N + 1023) / 1024) + 15) & ~15;
  const int first = tid * chunk, last = min(N, first + chunk);
  const bool vec = (reinterpret_cast<uintptr_t>(p.reset_buf) & 15u) == 0;
  int c = 0;
  for (int i = first; i < last; i += 16) {
    if (vec && i + 16 <= last) c += count_flags16(*reinterpret_cast<const uint4*>(p.reset_buf + i));
    else for (int k = i; k < min(i + 16, last); ++k) c += p.reset_buf[k] != 0;
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
  const int count = s_total;
  if (reset_ids && c) {
    int off = s_warp[warp] + incl - c;
    for (int i = first; i < last; i += 16) {
      if (vec && i + 16 <= last) {
        const uint4 v = *reinterpret_cast<const uint4*>(p.reset_buf + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t m = w[q];
          while (m) { const int b = __ffs(m) - 1; m &= m - 1; reset_ids[off++] = i + 4 * q + (b >> 3); }
        }
      } else {
        for (int k = i; k < min(i + 16, last); ++k) if (p.reset_buf[k]) reset_ids[off++] = k;
      }
    }
  }
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
    if (p.send_timeouts && time_outs_extras) {
      const bool v16 = ((reinterpret_cast<uintptr_t>(p.time_out_buf) | reinterpret_cast<uintptr_t>(time_outs_extras)) & 15u) == 0;
      const int n16 = v16 ? N / 16 : 0;
      for (int i = tid; i < n16; i += 1024)
        reinterpret_cast<uint4*>(time_outs_extras)[i] = reinterpret_cast<const uint4*>(p.time_out_buf)[i];
      for (int i = n16 * 16 + tid; i < N; i += 1024) time_outs_extras[i] = p.time_out_buf[i];
    }
  }
  for (int k = tid; k < ns + 2; k += 1024) other[k] = 0.f;
  if (advance && p.step_counter_dev && tid == 0) *p.step_counter_dev = step_eff;
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

static int launch_k1(const LgkStepParams* p, cudaStream_t st) {
  const TileLayout L = make_layout(p->num_bodies, p->num_feet, p->num_reward_slots);
  // ~25 KB of staged tiles per CTA: ask for the largest shared-memory carve-out so that 7 CTAs are resident per SM
  if (int rc = ensure_func_attr(reinterpret_cast<const void*>(post_kernel<0, false>), L.total, "cudaFuncSetAttribute(post_kernel)", true)) return rc;
  static const int extra_smem = getenv("LGK_K1_EXTRA_SMEM") ? atoi(getenv("LGK_K1_EXTRA_SMEM")) : 0;    // occupancy experiments
  if (extra_smem)
    if (int rc = ensure_func_attr(reinterpret_cast<const void*>(post_kernel<0, false>), L.total + extra_smem, "cudaFuncSetAttribute(post_kernel)", true)) return rc;
  const cudaError_t e = launch_chained(post_kernel<0, false>, dim3((p->num_envs + kTile - 1) / kTile), dim3(kK1Threads), (size_t)(L.total + extra_smem), st, *p);
  count_launch();
  return check_cuda(e, "post_kernel launch");
}

// ---- fused K1 + K2 (one launch): a height field behind the height columns, heights not needed by a reward term.
// LGK_FUSED=0 / lgk_set_fused(0) keeps the two-kernel path; LGK_FUSED_SCAN_WARPS overrides the scan-warp count.
static int g_fused = -1, g_fused_sw = 0;
extern "C" int lgk_set_fused(int enable) { const int prev = g_fused; g_fused = enable ? 1 : 0; return prev < 0 ? 0 : prev; }
static bool g_fused_sw_set = false;
extern "C" int lgk_set_fused_scan_warps(int n) { const int prev = g_fused_sw; g_fused_sw = n; g_fused_sw_set = true; return prev; }
static bool fused_eligible(const LgkStepParams* p) {
  static bool env_read = false;
  if (!env_read) {
    env_read = true;
    const char* e = getenv("LGK_FUSED");
    if (g_fused < 0) g_fused = (e && e[0] == '1') ? 1 : 0;
    const char* w = getenv("LGK_FUSED_SCAN_WARPS");
    if (!g_fused_sw_set) g_fused_sw = w ? atoi(w) : 0;
  }
  const int groups = (48 + p->num_height_points + 31) / 32;
  return g_fused && p->measure_heights && !p->terrain_is_plane && p->num_height_points > 0 && groups > 2 && groups <= 12 &&
         !p->reward_active[LGK_R_BASE_HEIGHT];
}

template <int G, bool RECIP>
static int launch_fused_t(const LgkStepParams* p, int scan_warps, cudaStream_t st) {
  const TileLayout L = make_layout(p->num_bodies, p->num_feet, p->num_reward_slots, true);
  if (int rc = ensure_func_attr(reinterpret_cast<const void*>(post_kernel<G, RECIP>), L.total, "cudaFuncSetAttribute(post_kernel fused)", true)) return rc;
  const cudaError_t e = launch_chained(post_kernel<G, RECIP>, dim3((p->num_envs + kTile - 1) / kTile), dim3(kK1Threads + 32 * scan_warps),
                                       (size_t)L.total, st, *p);
  count_launch();
  return check_cuda(e, "post_kernel (fused) launch");
}

static int launch_fused(const LgkStepParams* p, cudaStream_t st) {
  const int groups = (48 + p->num_height_points + 31) / 32;
  // few tiles per SM: 8 scan warps keep the scan shorter than the role warps' chain; many tiles: 4, so that three tiles
  // are resident per SM
  int sw = g_fused_sw > 0 ? g_fused_sw : ((p->num_envs + kTile - 1) / kTile <= 2 * 148 ? 8 : 4);
  sw = sw < 1 ? 1 : (sw > 8 ? 8 : sw);
  const bool rc = p->horizontal_scale_recip != 0.f;
  if (groups <= 8) return rc ? launch_fused_t<8, true>(p, sw, st) : launch_fused_t<8, false>(p, sw, st);
  return rc ? launch_fused_t<12, true>(p, sw, st) : launch_fused_t<12, false>(p, sw, st);
}

static int launch_k2(const LgkStepParams* p, int mode, cudaStream_t st) {
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
    if (flat) e = launch_chained(scan_obs_fast_kernel<2, kObs, false>, gd, bd, 0, st, *p);
    else if (groups <= 8) e = mode == kScan ? LGK_K2(8, kScan) : (mode == kObs ? LGK_K2(8, kObs) : LGK_K2(8, kScan | kObs));
    else e = mode == kScan ? LGK_K2(12, kScan) : (mode == kObs ? LGK_K2(12, kObs) : LGK_K2(12, kScan | kObs));
#undef LGK_K2
    count_launch();
    return check_cuda(e, "scan_obs_fast_kernel launch");
  }
  if (groups <= 2) e = launch_chained(scan_obs_kernel<2>, dim3(blocks), dim3(kK2Threads), smem, st, *p, mode);
  else if (groups <= 8) e = launch_chained(scan_obs_kernel<8>, dim3(blocks), dim3(kK2Threads), smem, st, *p, mode);
  else e = launch_chained(scan_obs_kernel<12>, dim3(blocks), dim3(kK2Threads), smem, st, *p, mode);
  count_launch();
  return check_cuda(e, "scan_obs_kernel launch");
}

extern "C" int lgk_post_physics(const LgkStepParams* p, void* stream) {
  if (int rc = validate_step(p)) return rc;
  LGK_REQUIRE((p->phase_mask & (LGK_PHASE_PRE | LGK_PHASE_POST)) != 0, "phase_mask selects nothing");
  LGK_REQUIRE(p->num_height_points <= 256 && p->num_obs <= 32 * kMaxGroupsAny, "at most 256 height points");
  cudaStream_t st = (cudaStream_t)stream;
  const bool pre = (p->phase_mask & LGK_PHASE_PRE) != 0, post = (p->phase_mask & LGK_PHASE_POST) != 0;
  const bool heights = p->measure_heights != 0;
  // the base_height reward needs this step's heights inside K1 (LR:884-887): run the scan first in that case
  const bool scan_first = heights && p->reward_active[LGK_R_BASE_HEIGHT] != 0;
  if (heights && !p->terrain_is_plane) LGK_REQUIRE(p->scan_frames != nullptr, "scan_frames buffer is null");
  if (pre && post) {              // fused step
    if (fused_eligible(p)) return launch_fused(p, st);
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
  if (int rc = launch_k1(p, st)) return rc;      // POST only
  return heights ? launch_k2(p, kObs, st) : LGK_OK;
}

extern "C" int lgk_step_debug_timeline(int64_t* device_buf16) {
  long long* ptr = reinterpret_cast<long long*>(device_buf16);
  return check_cuda(cudaMemcpyToSymbol(g_k1_timeline, &ptr, sizeof(ptr)), "cudaMemcpyToSymbol(g_k1_timeline)");
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
  const cudaError_t e = launch_chained(finalize_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, *p, reset_ids, reset_count,
                                       episode_means, time_outs_extras, (int)advance);
  count_launch();
  return check_cuda(e, "finalize_kernel launch");
}
