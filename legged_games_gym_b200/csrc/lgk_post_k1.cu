// lgk_post_k1.cu -- K1 of the post-physics step (reference LR:106-230, 329-508, 872-969): everything that exists once per
// environment or once per joint / foot / body.
//
// post_kernel: PERSISTENT CTAs (4 warps, at most one resident set of CTAs per SM) loop over tiles of 32 consecutive envs,
// lane = env, warp = role.  Every reference tensor is env-major row-major, so the tile's slice of root_states /
// dof_state / contact_forces / actions / torques / last_actions / last_dof_vel / commands / feet_air_time is ONE
// contiguous chunk per tensor: each arrives by a single TMA bulk copy (cp.async.bulk + mbarrier), while the chunks of the
// CTA's NEXT tile are pulled into L2 (cp.async.bulk.prefetch.L2) so that its loads find them there; whole-tile results
// (commands, feet_air_time, last_*, base_*, the 48-column observation head, the scan frames for K2) leave by bulk stores.
//   phase A  every role reduces its three joints, its foot and its share of the penalised / termination bodies to 13
//            partial sums (shared memory)
//   phase B  split over the four roles (lgk_step_device.cuh): role 0 base_lin_vel + push + linear-velocity terms, role 1
//            base_ang_vel + heading command + angular terms, role 2 projected gravity + yaw frame + half of the summed
//            terms, role 3 episode length, time-out / termination flags + the other half
//   phase F  role 0: reward sum, positive clip, termination term, cross-env extras sums, reset of root / commands /
//            terrain level (cold, out of line); meanwhile every role finishes its joints (reset draw, nine observation
//            columns, histories)
//   stores   bulk stores + coalesced rows; reset rows and LSTM-state zeroing cooperatively
// Phases are selected by LgkStepParams.phase_mask so that Python code can run between them (user reward terms, a
// subclass's reset_idx).
#include <stdlib.h>
#include "lgk_tile.cuh"

namespace lgk {

// ------------------------------------------------------------------ shared-memory carve-up (bytes, 16-aligned)
struct TileLayout {
  int root, dof, contact, act, tq, lact, ldv, cmd, fat, lc, head, blv, bav, pg, lrv, frame, sums, ep, epo, lvl, typ, org, rew,
      rewp, flags, part, misc, total;
};

__host__ __device__ inline int al16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline TileLayout make_layout(int nb, int nfeet, int nslots) {
  TileLayout L;
  int o = 0;
  L.root = o;    o += al16(kTile * 13 * 4);
  L.dof = o;     o += al16(kTile * 24 * 4);
  // the 48-column observation head (row stride 49: conflict-free lane = env writes) reuses the contact tile, which
  // is dead once phase A has run (a CTA barrier separates the two uses)
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
  L.frame = o;   o += al16(kTile * kFrameFloats * 4);
  L.sums = o;    o += al16((nslots > 0 ? nslots : 1) * kTile * 4);   // episode_sums rows of the tile: [K][32]
  L.ep = o;      o += al16(kTile * 8);                                // episode_length_buf as loaded (int64)
  L.epo = o;     o += al16(kTile * 8);                                // ... as it leaves
  L.lvl = o;     o += al16(kTile * 8);                                // terrain_levels (int64)
  L.typ = o;     o += al16(kTile * 8);                                // terrain_types (int64)
  L.org = o;     o += al16(kTile * 3 * 4);                            // env_origins
  L.rew = o;     o += al16(kTile * 4);
  L.rewp = o;    o += al16(4 * kTile * 4);                            // reward partials of the four roles
  L.flags = o;   o += al16(kTile * 2);                                // reset flags [32] then time_out flags [32]
  L.part = o;    o += al16(PS_COUNT * 4 * kTile * 4);                 // partial sums: [slot][role][env]
  L.misc = o;    o += 16;   // mbarrier
  L.total = o;
  return L;
}

// used only by the partial-tile fallback path: kept out of line and rolled so the hot path stays compact
__device__ __noinline__ void copy_f32(float* dst, const float* src, int n, int tid) {
#pragma unroll 1
  for (int i = tid; i < n; i += kK1Threads) dst[i] = src[i];
}
__device__ __noinline__ void copy_u8(uint8_t* dst, const uint8_t* src, int n, int tid) {
#pragma unroll 1
  for (int i = tid; i < n; i += kK1Threads) dst[i] = src[i];
}

__device__ long long* g_k1_timeline = nullptr;     // profiling hook (lgk_step_debug_timeline)
// stamps of the first tile of CTA 0 ([0..8]) and of the last CTA ([16..24])
__device__ __forceinline__ void k1_stamp(int slot, bool first_tile) {
#ifdef LGK_K1_TIMELINE          // (the stamps cost a global load each: compiled in only for timeline builds)
  if (g_k1_timeline != nullptr && first_tile && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_k1_timeline[blockIdx.x == 0 ? slot : 16 + slot] = (long long)t;
  }
#endif
}

__device__ __forceinline__ void role_sync() { named_sync(1, kK1Threads); }

// ep_len % period == 0 (LR:334) without the 64-bit division routine on the hot path: episode lengths fit 31 bits
__device__ __forceinline__ bool divisible(long long ep_len, int period) {
  if ((unsigned long long)ep_len < 0x80000000ull) return (uint32_t)ep_len % (uint32_t)period == 0;
  return ep_len % (long long)period == 0;
}

#ifndef LGK_K1_MINBLOCKS
#define LGK_K1_MINBLOCKS 7
#endif
#ifndef LGK_K1_PREFETCH
#define LGK_K1_PREFETCH 1
#endif

constexpr int kMaxRowsPerWarp = 5;      // episode_sums rows per role warp: ceil(LGK_R_COUNT / 4)

// FAST: the whole step in one pass (PRE | POST) over full, 16-byte aligned tiles -- the phase flags and the partial-tile
// fallback are compiled out (half the code size: the kernel is instruction-fetch sensitive).  !FAST: every other case.
template <bool FAST>
__global__ void __launch_bounds__(kK1Threads, LGK_K1_MINBLOCKS)
post_kernel(const __grid_constant__ LgkStepParams p, const __grid_constant__ TileLayout L, int ntiles) {
  extern __shared__ __align__(128) uint8_t smem[];
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
  long long* s_epo = reinterpret_cast<long long*>(smem + L.epo);
  long long* s_lvl = reinterpret_cast<long long*>(smem + L.lvl);
  long long* s_typ = reinterpret_cast<long long*>(smem + L.typ);
  float* s_org = reinterpret_cast<float*>(smem + L.org);
  float* s_rew = reinterpret_cast<float*>(smem + L.rew);
  float* s_rewp = reinterpret_cast<float*>(smem + L.rewp);
  uint8_t* s_flags = smem + L.flags;
  float* s_part = reinterpret_cast<float*>(smem + L.part);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.misc);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NB = p.num_bodies, F = p.num_feet, P = p.num_height_points, O = p.num_obs, N = p.num_envs;
  const int K = p.num_reward_slots;
  const int mask = p.phase_mask;
  const bool pre = FAST || (mask & LGK_PHASE_PRE) != 0;
  const bool fin = FAST || (mask & (LGK_PHASE_POST | LGK_PHASE_POST_REWARD)) != 0;      // positive clip + termination term
  const bool rst = FAST || (mask & LGK_PHASE_POST) != 0;                                // in-kernel reset_idx
  const bool obsph = FAST || (mask & (LGK_PHASE_POST | LGK_PHASE_POST_OBS)) != 0;       // observations + histories
  const bool fat_active = p.reward_active[LGK_R_FEET_AIR_TIME] != 0 && F > 0;
  const bool curriculum = p.terrain_curriculum != 0;
  const bool want_frames = p.measure_heights && p.scan_frames != nullptr;      // K2 takes the yaw frame and the post-reset z from here
  const bool head_to_k2 = p.measure_heights != 0;      // K2 finishes the rows (noise + clip); flat tasks finish them here
  // whole-tile paths need unit root stride and 16-byte aligned tile chunks
  const bool bulk_ok = FAST || ((p.actors_per_env == 1) && (N % 4 == 0));
  const bool has_org = p.env_origins != nullptr;
  pdl_launch_dependents();
  k1_stamp(0, true);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
#if LGK_K1_PREFETCH
  // the CTA's first tile: start pulling its chunks into L2 while the previous kernel of the step is still draining (a
  // prefetch carries no data into the SM: whatever that kernel still writes lands in the same L2)
  if (FAST && tid == 0 && (int)blockIdx.x < ntiles) {
    const size_t n0 = (size_t)blockIdx.x * kTile;
    if (!p.host_state) {
      bulk_prefetch_l2(p.root_states + n0 * 13, kTile * 13 * 4);
      bulk_prefetch_l2(p.dof_state + n0 * 24, kTile * 24 * 4);
      bulk_prefetch_l2(p.contact_forces + n0 * NB * 3, kTile * NB * 3 * 4);
    }
    bulk_prefetch_l2(p.actions + n0 * 12, kTile * 12 * 4);
    bulk_prefetch_l2(p.last_actions + n0 * 12, kTile * 12 * 4);
    bulk_prefetch_l2(p.last_dof_vel + n0 * 12, kTile * 12 * 4);
    bulk_prefetch_l2(p.commands + n0 * 4, kTile * 4 * 4);
  }
#endif
  pdl_wait();              // everything below reads state written by the previous kernels of the step
  const int step_eff = p.step_counter_dev ? (*p.step_counter_dev + 1) : p.step;
  const bool do_push = pre && (p.step_counter_dev ? (p.push_interval > 0 && step_eff % p.push_interval == 0) : (p.do_push != 0));
  const RngKey key = make_key(p.seed, step_eff);
  float* stats = p.reset_stats + (size_t)(step_eff & 1) * (K + 2);
  // lgk_post_physics_finalize: K2 and the finalize CTA riding in its grid take the step from word [1] of the counter (no
  // kernel that runs concurrently with this one reads it), so that the finalize CTA may advance word [0] while K2 runs
  if ((p.phase_mask & kPhaseFusedFin) && p.step_counter_dev && blockIdx.x == 0 && tid == 0) p.step_counter_dev[1] = step_eff;

  // per-env scalar rows (one 128-byte row segment per tensor and tile; lane = env) travel through registers, loaded one
  // tile AHEAD: the loads of tile i+1 are in flight while tile i is processed
  float pf_sum[kMaxRowsPerWarp];
  long long pf_ep = 0, pf_lvl = 0, pf_typ = 0;
  auto fetch_rows = [&](int t) {
    const int e0 = t * kTile;
    if (t < ntiles && e0 + lane < N) {
#pragma unroll
      for (int j = 0; j < kMaxRowsPerWarp; ++j) {
        const int k = warp + 4 * j;
        if (k < K) pf_sum[j] = p.episode_sums[(size_t)k * N + e0 + lane];
      }
      if (warp == (K & 3)) pf_ep = p.episode_length_buf[e0 + lane];
      if (curriculum && warp == ((K + 1) & 3)) pf_lvl = p.terrain_levels[e0 + lane];
      if (curriculum && warp == ((K + 2) & 3)) pf_typ = p.terrain_types[e0 + lane];
    }
  };
  fetch_rows(blockIdx.x);

  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const bool first = it == 0;
    const int env0 = tile * kTile;
    const int nval = FAST ? kTile : min(kTile, N - env0);
    const bool bulk = FAST || (bulk_ok && nval == kTile);

    // ---------------- stage the tile
    if (bulk && tid == 0) {
      uint32_t bytes = kTile * (13 + 24 + 3 * NB + 12 + 12 + 12 + 12 + 4) * 4;
      if (F > 0) bytes += kTile * F * 4 + kTile * F;
      if (has_org) bytes += kTile * 3 * 4;
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
      if (has_org) bulk_g2s(s_org, p.env_origins + (size_t)env0 * 3, kTile * 3 * 4, bar);
#if LGK_K1_PREFETCH
      // the CTA's next tile: pull its chunks into L2 now, a whole tile's processing time ahead of their use
      const int nt = tile + gridDim.x;
      if (nt < ntiles && (nt + 1) * kTile <= N) {
        const size_t n0 = (size_t)nt * kTile;
        if (!p.host_state) {        // (sim state in pinned host memory: a prefetch would pull the chunk over PCIe twice)
          bulk_prefetch_l2(p.root_states + n0 * 13, kTile * 13 * 4);
          bulk_prefetch_l2(p.dof_state + n0 * 24, kTile * 24 * 4);
          bulk_prefetch_l2(p.contact_forces + n0 * NB * 3, kTile * NB * 3 * 4);
        }
        bulk_prefetch_l2(p.actions + n0 * 12, kTile * 12 * 4);
        bulk_prefetch_l2(p.torques + n0 * 12, kTile * 12 * 4);
        bulk_prefetch_l2(p.last_actions + n0 * 12, kTile * 12 * 4);
        bulk_prefetch_l2(p.last_dof_vel + n0 * 12, kTile * 12 * 4);
        bulk_prefetch_l2(p.commands + n0 * 4, kTile * 4 * 4);
        if (F > 0) bulk_prefetch_l2(p.feet_air_time + n0 * F, kTile * F * 4);
      }
#endif
    }
    // the scalar rows of THIS tile are in registers (fetched one tile ago); park them and start the next tile's loads
#pragma unroll
    for (int j = 0; j < kMaxRowsPerWarp; ++j) {
      const int k = warp + 4 * j;
      if (k < K) s_sums[k * kTile + lane] = pf_sum[j];
    }
    if (warp == (K & 3)) s_ep[lane] = pf_ep;
    if (curriculum && warp == ((K + 1) & 3)) s_lvl[lane] = pf_lvl;
    if (curriculum && warp == ((K + 2) & 3)) s_typ[lane] = pf_typ;
    fetch_rows(tile + gridDim.x);
    k1_stamp(1, first);
    if (bulk) {
      mbar_wait(bar, it & 1);
    } else if (!FAST) {
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
      if (has_org) copy_f32(s_org, p.env_origins + (size_t)env0 * 3, nval * 3, tid);
      role_sync();
    }
    k1_stamp(2, first);

    // ---------------- phase A: every role, its share of the per-joint / per-foot / per-body terms
    const int e = lane, role = warp, env = env0 + e;
    const bool valid = e < nval;
    const int envc = valid ? env : env0;      // clamped index for loads of padded lanes
    const uint32_t genv = (uint32_t)(p.env_id_offset + env);
    float* root = s_root + e * 13;
    float* dof = s_dof + e * 24;
    float* cmd = s_cmd + e * 4;
    float* fat = s_fat + e * F;
    uint8_t* lc = s_lc + e * F;
    float* sums = s_sums + e;                 // tile-local episode sums, row stride kTile
    float* head = s_head + e * 49;
    CmdRegs c{cmd[0], cmd[1], cmd[2], cmd[3]};   // read before the barrier: role 1 rewrites the row in phase B
    if (pre) {
      const float* hrow = (valid && p.measure_heights && p.reward_active[LGK_R_BASE_HEIGHT])
                              ? p.measured_heights + (size_t)env * P : nullptr;       // the scan ran before this kernel
      RolePartials mine;
      role_partials(p, role, dof, s_contact + e * NB * 3, s_act + e * 12, s_tq + e * 12, s_lact + e * 12, s_ldv + e * 12,
                    fat, lc, root[2], hrow, mine);
#pragma unroll
      for (int k = 0; k < PS_COUNT; ++k) s_part[(k * 4 + role) * kTile + e] = mine.v[k];
    }
    role_sync();           // contact rows are dead from here on: their region becomes the observation head
    k1_stamp(3, first);

    // ---------------- phase B: the once-per-env work, split over the four roles
    if (pre) {
      const long long ep_len = s_ep[e] + 1;                                // LR:114
      if (divisible(ep_len, p.resample_period)) resample_cmd_regs(p, key, genv, c);   // LR:333-336 (every role, same result)
      const float* part = s_part + e;
      float rp;
      if (role == 0) {
        V3 v;
        rp = phase_b_lin(p, do_push, key, genv, root, c, part, kTile, sums, kTile, v);
        s_blv[3 * e] = v.x; s_blv[3 * e + 1] = v.y; s_blv[3 * e + 2] = v.z;
        head[0] = v.x * p.obs_scale_lin_vel; head[1] = v.y * p.obs_scale_lin_vel; head[2] = v.z * p.obs_scale_lin_vel;
      } else if (role == 1) {
        V3 v;
        rp = phase_b_ang(p, root, c, cmd, sums, kTile, v);
        s_bav[3 * e] = v.x; s_bav[3 * e + 1] = v.y; s_bav[3 * e + 2] = v.z;
        head[3] = v.x * p.obs_scale_ang_vel; head[4] = v.y * p.obs_scale_ang_vel; head[5] = v.z * p.obs_scale_ang_vel;
      } else if (role == 2) {
        V3 v;
        rp = phase_b_grav(p, root, part, kTile, sums, kTile, v);
        s_pg[3 * e] = v.x; s_pg[3 * e + 1] = v.y; s_pg[3 * e + 2] = v.z;
        head[6] = v.x; head[7] = v.y; head[8] = v.z;
        if (want_frames) {     // pre-reset yaw frame for K2 (LR:853-854 uses the pose of THIS step before reset_idx)
          const YawFrame yf = yaw_frame(root[5], root[6], root[0], root[1]);
          *reinterpret_cast<float4*>(s_frame + e * kFrameFloats) = make_float4(yf.zn, yf.wn, yf.rx, yf.ry);
        }
        if (p.base_quat != nullptr && valid)      // LowLevelGame.base_quat: the pose the rotations use, kept past reset_idx
          *reinterpret_cast<float4*>(p.base_quat + (size_t)env * 4) = make_float4(root[3], root[4], root[5], root[6]);
      } else {
        bool reset, time_out;
        rp = phase_b_flag(p, root, part, kTile, sums, kTile, ep_len, reset, time_out);
        s_flags[e] = reset ? 1 : 0;
        s_flags[kTile + e] = time_out ? 1 : 0;
        s_epo[e] = ep_len;
      }
      s_rewp[role * kTile + e] = rp;
    }
    role_sync();
    k1_stamp(4, first);

    // ---------------- phase F: role 0 finishes the reward and resets; every role finishes its joints
    bool reset_e, time_out_e;
    if (pre) { reset_e = s_flags[e] != 0; time_out_e = s_flags[kTile + e] != 0; }
    else { reset_e = p.reset_buf[envc] != 0; time_out_e = p.time_out_buf[envc] != 0; }
    const bool resetting = rst && valid && reset_e;
    if (role == 0) {
      float rew = pre ? (s_rewp[e] + s_rewp[kTile + e]) + (s_rewp[2 * kTile + e] + s_rewp[3 * kTile + e]) : p.rew_buf[envc];
      long long ep_len = pre ? s_epo[e] : s_ep[e];
      if (fin) rew = env_finish_reward(p, rew, reset_e, time_out_e, sums, kTile);
      if (rst) {
        // cross-env sums for extras["episode"] over the envs that reset, using their PRE-reset episode sums (LR:179-183)
        const uint32_t rmask = __ballot_sync(0xffffffffu, resetting);
        if (rmask != 0) {
          for (int k = 0; k < K; ++k) {
            float v = 0.f;
            if (resetting) { v = sums[k * kTile]; sums[k * kTile] = 0.f; }
            v = warp_sum(v);
            if (lane == 0) atomicAdd(stats + k, v);
          }
          if (lane == 0) atomicAdd(stats + K, (float)__popc(rmask));
        }
      }
      s_rew[e] = rew;
      s_epo[e] = ep_len;
      if (obsph) {
        if (!pre) {      // split mode: PRE ran in an earlier launch, pick its rotations up from global memory
          head[0] = p.base_lin_vel[3 * envc] * p.obs_scale_lin_vel; head[1] = p.base_lin_vel[3 * envc + 1] * p.obs_scale_lin_vel;
          head[2] = p.base_lin_vel[3 * envc + 2] * p.obs_scale_lin_vel;
          head[3] = p.base_ang_vel[3 * envc] * p.obs_scale_ang_vel; head[4] = p.base_ang_vel[3 * envc + 1] * p.obs_scale_ang_vel;
          head[5] = p.base_ang_vel[3 * envc + 2] * p.obs_scale_ang_vel;
          head[6] = p.projected_gravity[3 * envc]; head[7] = p.projected_gravity[3 * envc + 1]; head[8] = p.projected_gravity[3 * envc + 2];
        }
        obs_head_cmd(p, cmd, head);                                        // post-reset commands
#pragma unroll
        for (int i = 0; i < 6; ++i) s_lrv[e * 6 + i] = root[7 + i];        // LR:134 (post push / reset)
        *reinterpret_cast<float4*>(s_frame + e * kFrameFloats + 4) = make_float4(root[2] - 0.5f, 0.f, 0.f, 0.f);   // post-reset z for the height columns (LR:225)
      }
    }
    if (obsph) {
      env_obs_head_role(p, role, dof, s_act + e * 12, head);
#pragma unroll
      for (int d = 3 * role; d < 3 * role + 3; ++d) s_ldv[e * 12 + d] = dof[2 * d + 1];   // LR:133 (post-reset dof_vel)
    }
    fence_async_smem();
    role_sync();
    // ---------------- reset tail (LR:147-191): every warp sees the same reset mask (lane = env) and takes every fourth reset
    // env; the env's rows inside the tile are redrawn warp-cooperatively, so the phases above carry no reset branch
    bool tile_reset = false;
    if (rst) {
      uint32_t rm = __ballot_sync(0xffffffffu, resetting);
      tile_reset = rm != 0;                // (CTA-uniform)
      if (tile_reset) {
        int turn = 0;
        while (rm) {
          const int ee = __ffs(rm) - 1;
          rm &= rm - 1;
          if ((turn++ & 3) != warp) continue;
          TileRows t{s_root + ee * 13, s_dof + ee * 24, s_cmd + ee * 4, s_fat + ee * F, s_head + ee * 49, s_ldv + ee * 12,
                     s_lrv + ee * 6, s_frame + ee * kFrameFloats, has_org ? s_org + ee * 3 : nullptr, s_lvl + ee, s_typ + ee,
                     s_epo + ee};
          tile_reset_env(p, key, (uint32_t)(p.env_id_offset + env0 + ee), env0 + ee, lane, t, kFrameFloats);
        }
        fence_async_smem();
        role_sync();
      }
      if (curriculum && role == 1) {             // mean terrain level over ALL envs (LR:186), after the level updates
        float lv = valid ? (float)s_lvl[e] : 0.f;
        lv = warp_sum(lv);
        if (lane == 0) atomicAdd(stats + K + 1, lv);
      }
    }
    k1_stamp(5, first);

    // ---------------- whole-tile write-backs: lane 0 of every warp issues its share of the bulk stores
    const bool frames_tile = bulk && want_frames && pre && obsph;      // both halves of the frame rows are fresh
    if (bulk) {
      if (lane == 0) {
        if (warp == 0) {
          if (pre) {
            bulk_s2g(p.base_lin_vel + (size_t)env0 * 3, s_blv, kTile * 3 * 4);
            bulk_s2g(p.base_ang_vel + (size_t)env0 * 3, s_bav, kTile * 3 * 4);
            bulk_s2g(p.projected_gravity + (size_t)env0 * 3, s_pg, kTile * 3 * 4);
          }
        } else if (warp == 1) {
          if (pre || rst) bulk_s2g(p.commands + (size_t)env0 * 4, s_cmd, kTile * 4 * 4);
          if (F > 0 && ((pre && fat_active) || (rst && obsph))) {
            bulk_s2g(p.feet_air_time + (size_t)env0 * F, s_fat, kTile * F * 4);
            if (pre && fat_active) bulk_s2g(p.last_contacts + (size_t)env0 * F, s_lc, kTile * F);
          }
          if (do_push) bulk_s2g(p.root_states + (size_t)env0 * 13, s_root, kTile * 13 * 4);
        } else if (warp == 2) {
          if (obsph) {
            bulk_s2g(p.last_actions + (size_t)env0 * 12, s_act, kTile * 12 * 4);
            bulk_s2g(p.last_dof_vel + (size_t)env0 * 12, s_ldv, kTile * 12 * 4);
          }
        } else {
          if (obsph) bulk_s2g(p.last_root_vel + (size_t)env0 * 6, s_lrv, kTile * 6 * 4);
          if (frames_tile) bulk_s2g(p.scan_frames + (size_t)env0 * kFrameFloats, s_frame, kTile * kFrameFloats * 4);
          if (tile_reset && curriculum && has_org) bulk_s2g(p.env_origins + (size_t)env0 * 3, s_org, kTile * 3 * 4);
        }
        bulk_commit();
      }
    } else if (!FAST) {
      if (pre) {
        copy_f32(p.base_lin_vel + (size_t)env0 * 3, s_blv, nval * 3, tid);
        copy_f32(p.base_ang_vel + (size_t)env0 * 3, s_bav, nval * 3, tid);
        copy_f32(p.projected_gravity + (size_t)env0 * 3, s_pg, nval * 3, tid);
      }
      if (pre || rst) copy_f32(p.commands + (size_t)env0 * 4, s_cmd, nval * 4, tid);
      if (F > 0 && ((pre && fat_active) || (rst && obsph))) {
        copy_f32(p.feet_air_time + (size_t)env0 * F, s_fat, nval * F, tid);
        if (pre && fat_active) copy_u8(p.last_contacts + (size_t)env0 * F, s_lc, nval * F, tid);
      }
      if (obsph) {
        copy_f32(p.last_actions + (size_t)env0 * 12, s_act, nval * 12, tid);
        copy_f32(p.last_dof_vel + (size_t)env0 * 12, s_ldv, nval * 12, tid);
        copy_f32(p.last_root_vel + (size_t)env0 * 6, s_lrv, nval * 6, tid);
      }
      if (tile_reset && curriculum && has_org) copy_f32(p.env_origins + (size_t)env0 * 3, s_org, nval * 3, tid);
      if (do_push) {
        for (int i = tid; i < nval * 13; i += kK1Threads) {
          const int ee = i / 13, cc = i - ee * 13;
          if (cc == 7 || cc == 8)
            p.root_states[((size_t)(env0 + ee) * p.actors_per_env + p.root_actor_offset) * 13 + cc] = s_root[i];
        }
      }
    }
    // per-env scalar rows: coalesced, lane = env
    if (lane < nval) {
      for (int k = warp; k < K; k += 4) p.episode_sums[(size_t)k * N + env0 + lane] = s_sums[k * kTile + lane];
      if (warp == (K & 3)) p.episode_length_buf[env0 + lane] = s_epo[lane];
      if (warp == ((K + 1) & 3)) p.rew_buf[env0 + lane] = s_rew[lane];
      if (pre && warp == ((K + 2) & 3)) { p.reset_buf[env0 + lane] = s_flags[lane]; p.time_out_buf[env0 + lane] = s_flags[kTile + lane]; }
      if (rst && curriculum && warp == ((K + 3) & 3)) p.terrain_levels[env0 + lane] = s_lvl[lane];
    }
    // scan frames when they do not leave as one tile: pose part (pre) and post-reset z (observation phase) of the [N,8] rows
    if (want_frames && !frames_tile && valid) {
      if (pre && role == 2)
        *reinterpret_cast<float4*>(p.scan_frames + (size_t)env * kFrameFloats) = *reinterpret_cast<const float4*>(s_frame + e * kFrameFloats);
      if (obsph && role == 0) p.scan_frames[(size_t)env * kFrameFloats + 4] = s_frame[e * kFrameFloats + 4];
    }
    k1_stamp(6, first);

    if (tile_reset) {
      // ---------------- reset rows: dof_state / root_states write-back + LSTM state zeroing (ANY:56-60); every warp sees
      // the same reset mask (lane = env) and takes every fourth reset env
      uint32_t rm = __ballot_sync(0xffffffffu, resetting);
      int turn = 0;
      while (rm) {
        const int ee = __ffs(rm) - 1;
        rm &= rm - 1;
        if ((turn++ & 3) != warp) continue;
        const int en = env0 + ee;
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
    }
    if (obsph) {
      if (!head_to_k2) {
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
      } else if (p.obs_head != nullptr) {
        // ---------------- the 48 proprioceptive columns, un-noised, into the compact [N,48] hand-over buffer of K2: the
        // tile is ONE contiguous 6 KB block (K2 adds noise + clip and writes the final row)
        for (int ee = warp * 8; ee < min(warp * 8 + 8, nval); ++ee) {
          float* dst = p.obs_head + (size_t)(env0 + ee) * kHeadCols;
          dst[lane] = s_head[ee * 49 + lane];
          if (lane < 16) dst[32 + lane] = s_head[ee * 49 + 32 + lane];
        }
      } else {
        // (callers without a hand-over buffer: un-noised head columns go through obs_buf)
        for (int ee = warp * 8; ee < min(warp * 8 + 8, nval); ++ee) {
          float* orow = p.obs_buf + (size_t)(env0 + ee) * O;
          orow[lane] = s_head[ee * 49 + lane];
          if (lane < 16) orow[32 + lane] = s_head[ee * 49 + 32 + lane];
        }
      }
    }
    k1_stamp(7, first);
    if (bulk && lane == 0) bulk_wait_read0();      // this lane's bulk stores have read their shared-memory sources
    role_sync();           // every warp is done with the tile's shared memory: the next tile may be staged
    k1_stamp(8, first);
  }
}

int launch_k1(const LgkStepParams* p, cudaStream_t st) {
  const TileLayout L = make_layout(p->num_bodies, p->num_feet, p->num_reward_slots);
  const bool fast = (p->phase_mask & ~kPhaseFusedFin) == (LGK_PHASE_PRE | LGK_PHASE_POST) && p->actors_per_env == 1 && p->num_envs % kTile == 0;
  const void* fn = fast ? reinterpret_cast<const void*>(post_kernel<true>) : reinterpret_cast<const void*>(post_kernel<false>);
  // ~31 KB of staged tiles per CTA: ask for the largest shared-memory carve-out so that LGK_K1_MINBLOCKS CTAs are resident per SM
  if (int rc = ensure_func_attr(fn, L.total, "cudaFuncSetAttribute(post_kernel)", true)) return rc;
  const int ntiles = (p->num_envs + kTile - 1) / kTile;
  static const int per_sm = getenv("LGK_K1_CTAS_PER_SM") ? atoi(getenv("LGK_K1_CTAS_PER_SM")) : LGK_K1_MINBLOCKS;   // tuning aid
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = sms * (per_sm > 0 ? per_sm : 1);
  if (grid > ntiles) grid = ntiles;
  const cudaError_t e = fast ? launch_chained(post_kernel<true>, dim3(grid), dim3(kK1Threads), (size_t)L.total, st, *p, L, ntiles)
                             : launch_chained(post_kernel<false>, dim3(grid), dim3(kK1Threads), (size_t)L.total, st, *p, L, ntiles);
  count_launch();
  return check_cuda(e, "post_kernel launch");
}

}  // namespace lgk

extern "C" int lgk_step_debug_timeline(int64_t* device_buf32) {
  long long* ptr = reinterpret_cast<long long*>(device_buf32);
#ifndef LGK_K1_TIMELINE
  if (ptr != nullptr) return lgk::set_error(LGK_ERR_ARG, "K1 timeline stamps are compiled out (build with -DLGK_K1_TIMELINE)");
#endif
  return lgk::check_cuda(cudaMemcpyToSymbol(lgk::g_k1_timeline, &ptr, sizeof(ptr)), "cudaMemcpyToSymbol(g_k1_timeline)");
}
