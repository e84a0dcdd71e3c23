// lgk_step_device.cuh -- per-environment pieces of post_physics_step shared by the scalar tile kernel and
// the explicit reset_idx kernel (device code).
// "tile-local" pointers address one env's row inside a staged tile (shared memory on the device).
#pragma once
#include "lgk_math.cuh"

namespace lgk {

// OBS-stream word mapping: column j of the observation row uses word (j/32)%4 of Philox block
// 32*(j/128) + j%32, so that lane l of a warp computes ONE block per 128 columns and stores four
// coalesced 128-byte segments (oracle/philox.py obs_uniforms restates the same mapping).
LGK_HD uint32_t obs_block_of(int j) { return (uint32_t)(32 * (j >> 7) + (j & 31)); }
LGK_HD int obs_word_of(int j) { return (j >> 5) & 3; }

// rng_block, NOT inlined: for the rarely-taken branches (resampling, pushes, resets) of kernels whose hot path must stay
// small (each inlined Philox4x32-10 is ~100 instructions)
LGK_COLD U4 rng_block_cold(const RngKey& k, uint32_t env, uint32_t stream, uint32_t blk) { return rng_block(k, env, stream, blk); }

// LLG:419-432: re-spawn the predator sphere relative to the freshly reset prey position (`root` = prey row)
LGK_D void spawn_predator(const LgkStepParams& p, const RngKey& key, uint32_t genv, int env, const float* root) {
  const U4 r = rng_block_cold(key, genv, LGK_STREAM_PREDATOR, 0);
  const float sgn = u32_to_uniform(r.w) < 0.5f ? -1.f : 1.f;
  float* pred = p.root_states + ((size_t)env * p.actors_per_env + p.predator_actor_offset) * 13;
  pred[0] = f_sub(root[0], sgn * scale_uniform(9.0f, 1.0f, u32_to_uniform(r.x)));
  pred[1] = f_sub(root[1], sgn * scale_uniform(9.0f, 1.0f, u32_to_uniform(r.y)));
  pred[2] = 0.3f;          // the z offset (word r.z) is drawn and then overwritten, LLG:432
}

struct EnvScalars {
  V3 blv, bav, pg;       // base_lin_vel, base_ang_vel, projected_gravity (LR:119-121)
  long long ep_len;      // episode_length_buf after += 1 (LR:114)
  bool reset, time_out;  // LR:139-145
  float rew;             // rew_buf (LR:193-210)
};

// ------------------------------------------------------------------ role-split variants (scalar kernel)
// The scalar kernel runs a tile of 32 envs on four warps, lane = env.  Warp ("role") r owns joints {3r,3r+1,3r+2}, foot r,
// the penalised bodies {r, r+4, ...} and the termination bodies {r, r+4}: in a first phase every role reduces its share of
// the per-joint / per-foot / per-body reward terms to a handful of partial sums (role_partials); after a CTA barrier
// role 0 alone does everything that exists once per env -- rotations, commands, termination, the reward terms in the
// reference's alphabetical order (env_finish) -- reading the other roles' partials from shared memory.  Same reference
// lines as env_pre above; only the summation order inside a term differs (fp32, within the 1e-5 bar).
#if defined(__CUDACC__)
enum {   // partial-sum slots, one float per (role, env)
  PS_ACTION_RATE = 0, PS_DOF_ACC, PS_DOF_POS_LIMITS, PS_DOF_VEL, PS_DOF_VEL_LIMITS, PS_STAND_STILL, PS_TORQUE_LIMITS,
  PS_TORQUES, PS_COLLISION, PS_FEET_AIR_TIME, PS_FEET_CONTACT_FORCES, PS_HEIGHT_ERR, PS_BITS, PS_COUNT
};
// PS_BITS packs the role's boolean facts: bit 0 termination contact, bit 1 foot in the air test of no_fly (contact z > 0.1),
// bit 2 stumble
struct RolePartials { float v[PS_COUNT]; };

// role r's share of LR:139-145 (termination contacts) and of the per-joint / per-foot / per-body terms of LR:872-969.
// dof/act/tq/lact/ldv/contact/fat/lc are the env's rows inside the staged tile.
LGK_D void role_partials(const LgkStepParams& p, int role, const float* dof, const float* contact, const float* act,
                         const float* tq, const float* lact, const float* ldv, float* fat, uint8_t* lc, float root_z,
                         const float* heights_row, RolePartials& o) {
  const int d0 = 3 * role;
  const float inv_dt = __frcp_rn(p.dt);       // (x / dt differs from x * (1 / dt) by at most one ulp: inside the 1e-5 bar)
#pragma unroll
  for (int k = 0; k < PS_COUNT; ++k) o.v[k] = 0.f;
  uint32_t bits = 0;
  for (int t = role; t < p.num_term; t += 4) {                        // LR:142
    const float* f = contact + 3 * p.term_idx[t];
    if (norm3(f[0], f[1], f[2]) > 1.0f) bits |= 1u;
  }
  if (p.reward_active[LGK_R_ACTION_RATE]) {                           // LR:901-903
#pragma unroll
    for (int d = d0; d < d0 + 3; ++d) { const float e = lact[d] - act[d]; o.v[PS_ACTION_RATE] += e * e; }
  }
  if (p.reward_active[LGK_R_COLLISION]) {                             // LR:905-908
    for (int b = role; b < p.num_pen; b += 4) {
      const float* f = contact + 3 * p.pen_idx[b];
      o.v[PS_COLLISION] += norm3(f[0], f[1], f[2]) > 0.1f ? 1.f : 0.f;
    }
  }
  if (p.reward_active[LGK_R_DOF_ACC]) {                               // LR:897-899
#pragma unroll
    for (int d = d0; d < d0 + 3; ++d) { const float a = (ldv[d] - dof[2 * d + 1]) * inv_dt; o.v[PS_DOF_ACC] += a * a; }
  }
  if (p.reward_active[LGK_R_DOF_POS_LIMITS]) {                        // LR:914-918
#pragma unroll
    for (int d = d0; d < d0 + 3; ++d) {
      const float q = dof[2 * d];
      o.v[PS_DOF_POS_LIMITS] += -fminf(q - p.dof_pos_lo[d], 0.f) + fmaxf(q - p.dof_pos_hi[d], 0.f);
    }
  }
  if (p.reward_active[LGK_R_DOF_VEL]) {                               // LR:893-895
#pragma unroll
    for (int d = d0; d < d0 + 3; ++d) { const float v = dof[2 * d + 1]; o.v[PS_DOF_VEL] += v * v; }
  }
  if (p.reward_active[LGK_R_DOF_VEL_LIMITS]) {                        // LR:920-925
#pragma unroll
    for (int d = d0; d < d0 + 3; ++d)
      o.v[PS_DOF_VEL_LIMITS] += clampf(fabsf(dof[2 * d + 1]) - p.dof_vel_limits[d] * p.soft_dof_vel_limit, 0.f, 1.f);
  }
  if (role < p.num_feet) {
    const float* cf = contact + 3 * p.feet_idx[role];                 // this role's foot
    if (p.reward_active[LGK_R_FEET_AIR_TIME]) {                       // LR:942-954 (stateful)
      const int f = role;
      const bool c = cf[2] > 1.0f;
      const bool filt = c || (lc[f] != 0);
      lc[f] = c ? 1 : 0;
      const bool first = (fat[f] > 0.f) && filt;
      const float air = fat[f] + p.dt;
      o.v[PS_FEET_AIR_TIME] = (air - 0.5f) * (first ? 1.f : 0.f);
      fat[f] = filt ? 0.f * air : air;     // air *= ~filt
    }
    if (p.reward_active[LGK_R_FEET_CONTACT_FORCES])                   // LR:966-969
      o.v[PS_FEET_CONTACT_FORCES] = fmaxf(norm3(cf[0], cf[1], cf[2]) - p.max_contact_force, 0.f);
    if (p.reward_active[LGK_R_NO_FLY] && cf[2] > 0.1f) bits |= 2u;    // CAS:43-46
    if (p.reward_active[LGK_R_STUMBLE] && norm2(cf[0], cf[1]) > 5.f * fabsf(cf[2])) bits |= 4u;   // LR:956-959
  }
  if (p.reward_active[LGK_R_STAND_STILL]) {                           // LR:961-964
#pragma unroll
    for (int d = d0; d < d0 + 3; ++d) o.v[PS_STAND_STILL] += fabsf(dof[2 * d] - p.default_dof_pos[d]);
  }
  if (p.reward_active[LGK_R_TORQUE_LIMITS]) {                         // LR:927-930
#pragma unroll
    for (int d = d0; d < d0 + 3; ++d)
      o.v[PS_TORQUE_LIMITS] += fmaxf(fabsf(tq[d]) - p.torque_limits[d] * p.soft_torque_limit, 0.f);
  }
  if (p.reward_active[LGK_R_TORQUES]) {                               // LR:889-891
#pragma unroll
    for (int d = d0; d < d0 + 3; ++d) o.v[PS_TORQUES] += tq[d] * tq[d];
  }
  if (p.reward_active[LGK_R_BASE_HEIGHT] && heights_row != nullptr) { // LR:886: sum_p (z - h_p), points p = role, role+4, ...
    for (int j = role; j < p.num_height_points; j += 4) o.v[PS_HEIGHT_ERR] += root_z - heights_row[j];
  }
  o.v[PS_BITS] = __uint_as_float(bits);
}

// ---- phase B, split over the four role warps (lane = env).  Every role reads the partial sums of ALL four roles from
// shared memory (`part`: slot k of role r at part[(k * 4 + r) * tile]) for the terms it owns, evaluates its share of
// LR:114-127 / 193-203 and returns its partial of the reward sum; the per-term episode sums are disjoint rows, so no two
// roles touch the same one.  Summation order: terms are added role by role instead of alphabetically (fp32, inside the
// 1e-5 bar).
//   role 0 "lin":  base_lin_vel (LR:119), push (LR:438-444), lin_vel_z, tracking_lin_vel, feet_air_time, stand_still
//   role 1 "ang":  base_ang_vel (LR:120), heading -> yaw command (LR:337-340), ang_vel_xy, tracking_ang_vel; owns `cmd`
//   role 2 "grav": projected_gravity (LR:121), orientation, action_rate, collision, dof_acc, dof_pos_limits, dof_vel,
//                  dof_vel_limits
//   role 3 "flag": episode length, time-out / termination / reset flags (LR:139-145), base_height, feet_contact_forces,
//                  no_fly, stumble, torque_limits, torques
// Command resampling (LR:347-369, rare) is evaluated by every role on its own register copy of the command so that no
// role waits for another; role 1 writes the row back.
struct CmdRegs { float c0, c1, c2, c3; };

LGK_D float part_sum(const float* part, int slot, int tile) {
  const float* q = part + slot * 4 * tile;
  return (q[0] + q[tile]) + (q[2 * tile] + q[3 * tile]);
}
LGK_D uint32_t part_bits(const float* part, int r, int tile) { return __float_as_uint(part[(PS_BITS * 4 + r) * tile]); }

LGK_COLD void resample_cmd_regs(const LgkStepParams& p, const RngKey& key, uint32_t genv, CmdRegs& c) {
  float tmp[4] = {c.c0, c.c1, c.c2, c.c3};
  resample_commands(p, tmp, rng_block(key, genv, LGK_STREAM_CMD, 0));
  c.c0 = tmp[0]; c.c1 = tmp[1]; c.c2 = tmp[2]; c.c3 = tmp[3];
}

#define LGK_TERM(ID, EXPR)                                            \
  if (p.reward_active[ID]) {                                          \
    const float r_ = (EXPR) * p.reward_scale[ID];                     \
    rew += r_;                                                        \
    sums[(size_t)p.reward_slot[ID] * sums_stride] += r_;              \
  }

// role 0.  root: the env's staged row (read 3..9, push writes 7..8).  Returns the reward partial.
LGK_D float phase_b_lin(const LgkStepParams& p, bool do_push, const RngKey& key, uint32_t genv, float* root,
                        const CmdRegs& c, const float* part, int tile, float* sums, int sums_stride, V3& blv) {
  const float qx = root[3], qy = root[4], qz = root[5], qw = root[6];
  blv = quat_rotate_inverse(qx, qy, qz, qw, V3{root[7], root[8], root[9]});
  if (do_push) {   // LR:438-444; rewards / obs of this step keep the pre-push base_lin_vel
    const U4 r = rng_block_cold(key, genv, LGK_STREAM_PUSH, 0);
    const float range = 2.0f * p.max_push_vel, lo = -p.max_push_vel;
    root[7] = scale_uniform(range, lo, u32_to_uniform(r.x));
    root[8] = scale_uniform(range, lo, u32_to_uniform(r.y));
  }
  float rew = 0.f;
  const float cmd_norm = norm2(c.c0, c.c1);
  LGK_TERM(LGK_R_FEET_AIR_TIME, part_sum(part, PS_FEET_AIR_TIME, tile) * (cmd_norm > 0.1f ? 1.f : 0.f))   // LR:942-954
  LGK_TERM(LGK_R_LIN_VEL_Z, blv.z * blv.z)                            // LR:872-874
  LGK_TERM(LGK_R_STAND_STILL, part_sum(part, PS_STAND_STILL, tile) * (cmd_norm < 0.1f ? 1.f : 0.f))       // LR:961-964
  {                                                                   // LR:932-935
    const float ex = c.c0 - blv.x, ey = c.c1 - blv.y;
    LGK_TERM(LGK_R_TRACKING_LIN_VEL, expf(-(ex * ex + ey * ey) / p.tracking_sigma))
  }
  return rew;
}

// role 1.  Updates c.c2 (heading controller) and writes the command row.
LGK_D float phase_b_ang(const LgkStepParams& p, const float* root, CmdRegs& c, float* cmd_row, float* sums, int sums_stride,
                        V3& bav) {
  const float qx = root[3], qy = root[4], qz = root[5], qw = root[6];
  bav = quat_rotate_inverse(qx, qy, qz, qw, V3{root[10], root[11], root[12]});
  if (p.heading_command) {                                            // LR:337-340
    const float h = heading_of(qx, qy, qz, qw);
    c.c2 = clampf(0.5f * wrap_to_pi(c.c3 - h), -1.f, 1.f);
  }
  cmd_row[0] = c.c0; cmd_row[1] = c.c1; cmd_row[2] = c.c2; cmd_row[3] = c.c3;
  float rew = 0.f;
  LGK_TERM(LGK_R_ANG_VEL_XY, bav.x * bav.x + bav.y * bav.y)           // LR:876-878
  {                                                                   // LR:937-940
    const float e = c.c2 - bav.z;
    LGK_TERM(LGK_R_TRACKING_ANG_VEL, expf(-(e * e) / p.tracking_sigma))
  }
  return rew;
}

// role 2
LGK_D float phase_b_grav(const LgkStepParams& p, const float* root, const float* part, int tile, float* sums,
                         int sums_stride, V3& pg) {
  pg = quat_rotate_inverse(root[3], root[4], root[5], root[6], V3{0.f, 0.f, -1.f});
  float rew = 0.f;
  LGK_TERM(LGK_R_ACTION_RATE, part_sum(part, PS_ACTION_RATE, tile))
  LGK_TERM(LGK_R_COLLISION, part_sum(part, PS_COLLISION, tile))
  LGK_TERM(LGK_R_DOF_ACC, part_sum(part, PS_DOF_ACC, tile))
  LGK_TERM(LGK_R_DOF_POS_LIMITS, part_sum(part, PS_DOF_POS_LIMITS, tile))
  LGK_TERM(LGK_R_DOF_VEL, part_sum(part, PS_DOF_VEL, tile))
  LGK_TERM(LGK_R_DOF_VEL_LIMITS, part_sum(part, PS_DOF_VEL_LIMITS, tile))
  LGK_TERM(LGK_R_ORIENTATION, pg.x * pg.x + pg.y * pg.y)              // LR:880-882
  return rew;
}

// role 3.  ep_len: episode_length_buf after += 1 (LR:114).
LGK_D float phase_b_flag(const LgkStepParams& p, const float* root, const float* part, int tile, float* sums,
                         int sums_stride, long long ep_len, bool& reset, bool& time_out) {
  const uint32_t b0 = part_bits(part, 0, tile), b1 = part_bits(part, 1, tile), b2 = part_bits(part, 2, tile),
                 b3 = part_bits(part, 3, tile);
  const uint32_t any = b0 | b1 | b2 | b3;
  const uint32_t feet_down = ((b0 >> 1) & 1u) + ((b1 >> 1) & 1u) + ((b2 >> 1) & 1u) + ((b3 >> 1) & 1u);
  time_out = (float)ep_len > p.max_episode_length;                    // LR:139-145
  reset = ((any & 1u) != 0) || time_out;
  float rew = 0.f;
  if (p.reward_active[LGK_R_BASE_HEIGHT]) {                           // LR:884-887
    const float mh = p.measure_heights ? part_sum(part, PS_HEIGHT_ERR, tile) / (float)p.num_height_points : root[2];   // LR:562: heights = 0
    const float e = mh - p.base_height_target;
    LGK_TERM(LGK_R_BASE_HEIGHT, e * e)
  }
  LGK_TERM(LGK_R_FEET_CONTACT_FORCES, part_sum(part, PS_FEET_CONTACT_FORCES, tile))
  LGK_TERM(LGK_R_NO_FLY, feet_down == 1u ? 1.f : 0.f)                 // CAS:43-46: exactly one foot on the ground
  LGK_TERM(LGK_R_STUMBLE, (any & 4u) ? 1.f : 0.f)                     // LR:956-959
  LGK_TERM(LGK_R_TORQUE_LIMITS, part_sum(part, PS_TORQUE_LIMITS, tile))
  LGK_TERM(LGK_R_TORQUES, part_sum(part, PS_TORQUES, tile))
  return rew;
}
#undef LGK_TERM

// the twelve base / command columns: role 0 lin vel (0..2), role 1 ang vel (3..5), role 2 gravity (6..8) right after their
// rotations; the command columns (9..11) once the command row is final (after a possible reset)
LGK_D void obs_head_cmd(const LgkStepParams& p, const float* cmd, float* out48) {
  out48[9] = cmd[0] * p.obs_scale_lin_vel; out48[10] = cmd[1] * p.obs_scale_lin_vel; out48[11] = cmd[2] * p.obs_scale_ang_vel;
}
// ---- reset_idx for one env of a staged tile, WARP-cooperative (LR:147-191 minus the cross-env means and the LSTM-state
// zeroing, which the caller does): the eight Philox blocks of the reset are drawn in parallel (lane j < 8 owns one), the
// terrain curriculum runs on lane 0 from staged rows (level, type, origin), and lanes write the env's rows inside the
// tile -- root (13), dof (12 x 2), command, feet_air_time -- plus everything the step has already derived from them
// (observation head columns, last_dof_vel, last_root_vel, scan-frame z), so no other phase has a reset branch.
// Same streams / blocks / words and the same individually rounded fp32 ops as env_reset() below.
struct TileRows {
  float* root; float* dof; float* cmd; float* fat; float* head; float* ldv; float* lrv; float* frame; float* origin;
  long long* level; long long* type; long long* ep_out;
};

LGK_D void tile_reset_env(const LgkStepParams& p, const RngKey& key, uint32_t genv, int env, int lane, const TileRows& t,
                          int frame_floats) {
  const uint32_t FULL = 0xffffffffu;
  // lane -> (stream, block): 0..2 dof, 3..4 root, 5 command, 6 terrain level wrap, 7 predator spawn
  const uint32_t strm = lane < 3 ? LGK_STREAM_RESET_DOF : (lane < 5 ? LGK_STREAM_RESET_ROOT : (lane == 5 ? LGK_STREAM_RESET_CMD
                        : (lane == 6 ? LGK_STREAM_TERRAIN : LGK_STREAM_PREDATOR)));
  const uint32_t blk = lane < 3 ? (uint32_t)lane : (lane < 5 ? (uint32_t)(lane - 3) : 0u);
  const U4 r = rng_block(key, genv, strm, blk);
  // ---- terrain curriculum + origin (lane 0; LR:446-469).  The curriculum reads the PRE-reset root position and command.
  float ox = 0.f, oy = 0.f, oz = 0.f;
  const uint32_t terr_word = __shfl_sync(FULL, r.x, 6);
  if (lane == 0) {
    if (t.origin) { ox = t.origin[0]; oy = t.origin[1]; oz = t.origin[2]; }
    if (p.terrain_curriculum) {
      const float dist = norm2(t.root[0] - ox, t.root[1] - oy);
      const bool up = dist > p.half_env_length;
      const bool down = (dist < norm2(t.cmd[0], t.cmd[1]) * p.max_episode_length_s * 0.5f) && !up;
      long long lvl = *t.level + (up ? 1 : 0) - (down ? 1 : 0);
      if (lvl >= p.max_terrain_level) lvl = (long long)(terr_word % (uint32_t)p.max_terrain_level);
      else if (lvl < 0) lvl = 0;
      *t.level = lvl;
      const float* o = p.terrain_origins + 3 * ((size_t)lvl * p.terrain_num_cols + (size_t)*t.type);
      ox = o[0]; oy = o[1]; oz = o[2];
      t.origin[0] = ox; t.origin[1] = oy; t.origin[2] = oz;
    }
  }
  ox = __shfl_sync(FULL, ox, 0); oy = __shfl_sync(FULL, oy, 0); oz = __shfl_sync(FULL, oz, 0);
  // ---- root row (LR:414-432): lane i < 13 owns column i; uniform k = i (i < 2) or i - 5 (i >= 7) of the eight root words
  {
    const int k = lane < 2 ? lane : lane - 5;
    const int src = 3 + ((k >> 2) & 1);
    const uint32_t w0 = __shfl_sync(FULL, r.x, src), w1 = __shfl_sync(FULL, r.y, src), w2 = __shfl_sync(FULL, r.z, src),
                   w3 = __shfl_sync(FULL, r.w, src);
    const float u = u32_to_uniform((k & 3) == 0 ? w0 : ((k & 3) == 1 ? w1 : ((k & 3) == 2 ? w2 : w3)));
    if (lane < 13) {
      float v = p.base_init_state[lane];
      if (lane < 3) v = f_add(v, lane == 0 ? ox : (lane == 1 ? oy : oz));
      if (lane < 2 && p.custom_origins) v = f_add(v, scale_uniform(2.0f, -1.0f, u));
      if (lane >= 7) v = scale_uniform(1.0f, -0.5f, u);
      t.root[lane] = v;
      if (lane >= 7) t.lrv[lane - 7] = v;                                 // LR:134 after the reset
      if (lane == 2) t.frame[4] = v - 0.5f;                               // height columns use the post-reset z (LR:225)
    }
    (void)frame_floats;
  }
  // ---- dofs (LR:397-407): joint d uses word d % 4 of block d / 4
  {
    const int d = lane < 12 ? lane : 0, src = d >> 2;
    const uint32_t w0 = __shfl_sync(FULL, r.x, src), w1 = __shfl_sync(FULL, r.y, src), w2 = __shfl_sync(FULL, r.z, src),
                   w3 = __shfl_sync(FULL, r.w, src);
    if (lane < 12) {
      const uint32_t w = (d & 3) == 0 ? w0 : ((d & 3) == 1 ? w1 : ((d & 3) == 2 ? w2 : w3));
      const float q = f_mul(p.default_dof_pos[d], scale_uniform(1.0f, 0.5f, u32_to_uniform(w)));
      t.dof[2 * d] = q;
      t.dof[2 * d + 1] = 0.f;
      t.head[12 + d] = (q - p.default_dof_pos[d]) * p.obs_scale_dof_pos;
      t.head[24 + d] = 0.f * p.obs_scale_dof_vel;
      t.ldv[d] = 0.f;                                                     // LR:133 after the reset
    }
  }
  if (lane < p.num_feet) t.fat[lane] = 0.f;                               // LR:175
  // ---- predator re-spawn (LLG:419-432) relative to the fresh prey position, then the command (LR:170), lane 0
  {
    const uint32_t px = __shfl_sync(FULL, r.x, 7), py = __shfl_sync(FULL, r.y, 7), pw = __shfl_sync(FULL, r.w, 7);
    const uint32_t cx = __shfl_sync(FULL, r.x, 5), cy = __shfl_sync(FULL, r.y, 5), cz = __shfl_sync(FULL, r.z, 5);
    __syncwarp();
    if (lane == 0) {
      if (p.predator_spawn) {
        const float sgn = u32_to_uniform(pw) < 0.5f ? -1.f : 1.f;
        float* pred = p.root_states + ((size_t)env * p.actors_per_env + p.predator_actor_offset) * 13;
        pred[0] = f_sub(t.root[0], sgn * scale_uniform(9.0f, 1.0f, u32_to_uniform(px)));
        pred[1] = f_sub(t.root[1], sgn * scale_uniform(9.0f, 1.0f, u32_to_uniform(py)));
        pred[2] = 0.3f;
      }
      resample_commands(p, t.cmd, U4{cx, cy, cz, 0u});
      obs_head_cmd(p, t.cmd, t.head);
      *t.ep_out = 0;                                                      // LR:176
    }
  }
}

// LR:212-222: the 48 proprioceptive columns, before noise -- role r writes the nine columns of its joints, role 0 also
// the twelve base / command columns
LGK_D void env_obs_head_role(const LgkStepParams& p, int role, const float* dof, const float* act, float* out48) {
#pragma unroll
  for (int d = 3 * role; d < 3 * role + 3; ++d) {
    out48[12 + d] = (dof[2 * d] - p.default_dof_pos[d]) * p.obs_scale_dof_pos;
    out48[24 + d] = dof[2 * d + 1] * p.obs_scale_dof_vel;
    out48[36 + d] = act[d];
  }
}
#endif

// LR:204-210: positive clip, then the termination term
LGK_D float env_finish_reward(const LgkStepParams& p, float rew, bool reset, bool time_out, float* sums,
                               int sums_stride) {
  if (p.only_positive_rewards) rew = fmaxf(rew, 0.f);
  if (p.reward_active[LGK_R_TERMINATION]) {
    const float r_ = ((reset && !time_out) ? 1.f : 0.f) * p.reward_scale[LGK_R_TERMINATION];
    rew += r_;
    sums[(size_t)p.reward_slot[LGK_R_TERMINATION] * sums_stride] += r_;
  }
  return rew;
}
#undef LGK_TERM

// ---- reset_idx for one env (LR:147-191) minus the cross-env means and the LSTM-state zeroing, which
// the callers do cooperatively.  Mutates the env's staged rows; terrain tables are global.
// Returns the env's (possibly updated) terrain level.
// rare (a few % of the envs per step): kept out of line so that the hot path of the scalar kernel stays compact
LGK_COLD void env_reset(const LgkStepParams& p, const RngKey& key, uint32_t genv, int env, float* root, float* dof,
                      float* cmd, float* fat, long long& ep_len) {
  float ox = 0.f, oy = 0.f, oz = 0.f;
  if (p.env_origins) { ox = p.env_origins[3 * env]; oy = p.env_origins[3 * env + 1]; oz = p.env_origins[3 * env + 2]; }
  if (p.terrain_curriculum) {                                         // LR:446-469
    const float dist = norm2(root[0] - ox, root[1] - oy);
    const bool up = dist > p.half_env_length;
    const bool down = (dist < norm2(cmd[0], cmd[1]) * p.max_episode_length_s * 0.5f) && !up;
    long long lvl = p.terrain_levels[env] + (up ? 1 : 0) - (down ? 1 : 0);
    if (lvl >= p.max_terrain_level) {
      const U4 r = rng_block_cold(key, genv, LGK_STREAM_TERRAIN, 0);
      lvl = (long long)(r.x % (uint32_t)p.max_terrain_level);
    } else if (lvl < 0) {
      lvl = 0;
    }
    p.terrain_levels[env] = lvl;
    const float* o = p.terrain_origins + 3 * ((size_t)lvl * p.terrain_num_cols + (size_t)p.terrain_types[env]);
    ox = o[0]; oy = o[1]; oz = o[2];
    p.env_origins[3 * env] = ox; p.env_origins[3 * env + 1] = oy; p.env_origins[3 * env + 2] = oz;
  }
  // _reset_dofs LR:397-407
  for (int b = 0; b < 3; ++b) {
    const U4 r = rng_block_cold(key, genv, LGK_STREAM_RESET_DOF, b);
    for (int i = 0; i < 4; ++i) {
      const int d = 4 * b + i;
      dof[2 * d] = f_mul(p.default_dof_pos[d], scale_uniform(1.0f, 0.5f, u32_to_uniform(pick(r, i))));
      dof[2 * d + 1] = 0.f;
    }
  }
  // _reset_root_states LR:414-432
  const U4 r0 = rng_block_cold(key, genv, LGK_STREAM_RESET_ROOT, 0);
  const U4 r1 = rng_block_cold(key, genv, LGK_STREAM_RESET_ROOT, 1);
  for (int i = 0; i < 13; ++i) root[i] = p.base_init_state[i];
  root[0] = f_add(root[0], ox); root[1] = f_add(root[1], oy); root[2] = f_add(root[2], oz);
  if (p.custom_origins) {
    root[0] = f_add(root[0], scale_uniform(2.0f, -1.0f, u32_to_uniform(r0.x)));
    root[1] = f_add(root[1], scale_uniform(2.0f, -1.0f, u32_to_uniform(r0.y)));
  }
  root[7] = scale_uniform(1.0f, -0.5f, u32_to_uniform(r0.z));
  root[8] = scale_uniform(1.0f, -0.5f, u32_to_uniform(r0.w));
  root[9] = scale_uniform(1.0f, -0.5f, u32_to_uniform(r1.x));
  root[10] = scale_uniform(1.0f, -0.5f, u32_to_uniform(r1.y));
  root[11] = scale_uniform(1.0f, -0.5f, u32_to_uniform(r1.z));
  root[12] = scale_uniform(1.0f, -0.5f, u32_to_uniform(r1.w));
  if (p.predator_spawn) spawn_predator(p, key, genv, env, root);
  resample_commands(p, cmd, rng_block_cold(key, genv, LGK_STREAM_RESET_CMD, 0));   // LR:170
  for (int f = 0; f < p.num_feet; ++f) fat[f] = 0.f;                            // LR:175
  ep_len = 0;                                                                   // LR:176
}

// LR:212-222: the 48 proprioceptive columns, before noise
LGK_D void env_obs_head(const LgkStepParams& p, const EnvScalars& s, const float* dof, const float* cmd,
                         const float* act, float* out48) {
  out48[0] = s.blv.x * p.obs_scale_lin_vel; out48[1] = s.blv.y * p.obs_scale_lin_vel; out48[2] = s.blv.z * p.obs_scale_lin_vel;
  out48[3] = s.bav.x * p.obs_scale_ang_vel; out48[4] = s.bav.y * p.obs_scale_ang_vel; out48[5] = s.bav.z * p.obs_scale_ang_vel;
  out48[6] = s.pg.x; out48[7] = s.pg.y; out48[8] = s.pg.z;
  out48[9] = cmd[0] * p.obs_scale_lin_vel; out48[10] = cmd[1] * p.obs_scale_lin_vel; out48[11] = cmd[2] * p.obs_scale_ang_vel;
  for (int d = 0; d < kDof; ++d) {
    out48[12 + d] = (dof[2 * d] - p.default_dof_pos[d]) * p.obs_scale_dof_pos;
    out48[24 + d] = dof[2 * d + 1] * p.obs_scale_dof_vel;
    out48[36 + d] = act[d];
  }
}

// LR:225-230 + LR:100-101 for one column: value (+ noise) clipped
LGK_HD float obs_finish(const LgkStepParams& p, float v, float noise_scale, uint32_t word) {
  if (p.add_noise) v = v + (2.0f * u32_to_uniform(word) - 1.0f) * noise_scale;
  return clampf(v, -p.clip_obs, p.clip_obs);
}

LGK_HD float obs_height_col(const LgkStepParams& p, float root_z, float h) {
  return clampf(root_z - 0.5f - h, -1.f, 1.f) * p.obs_scale_height;
}

}  // namespace lgk
