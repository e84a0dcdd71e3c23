// lgk_game.cu -- the hierarchical predator / prey games on top of LowLevelGame: everything HighLevelGame.step does after
// ll_env.step (reference legged_gym/envs/a1_game/high_level_game.py:178-239 = HLG) and DecHighLevelGame.step /
// post_physics_step (dec_high_level_game.py:204-261 = DHLG).  One thread per env; ~400 bytes of state per env, all of
// it streamed once (root rows of prey + predator, 16 + 3 observation columns, a handful of scalars), so the kernel is
// a single coalescing-friendly pass with no cross-env communication except the DHLG extras sums (warp shuffle + atomics).
#include "lgk_math.cuh"

namespace lgk {

__device__ __forceinline__ float warp_sum_g(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// HLG:357-378 / DHLG:321-361 for one agent.  `base` = ll_rew_weight * ll_rew (prey / single agent) or 0 (predator).
__device__ __forceinline__ float game_reward(const LgkGameAgent& a, float base, float dist, bool prev_reset, bool time_out,
                                             int env, int N) {
  float rew = base;
  if (a.active[LGK_G_EVASION]) {
    const float r = f_mul(dist, a.scale[LGK_G_EVASION]);
    rew = f_add(rew, r);
    a.sums[(size_t)a.slot[LGK_G_EVASION] * N + env] += r;
  }
  if (a.active[LGK_G_PURSUIT]) {
    const float r = f_mul(-dist, a.scale[LGK_G_PURSUIT]);
    rew = f_add(rew, r);
    a.sums[(size_t)a.slot[LGK_G_PURSUIT] * N + env] += r;
  }
  if (a.only_positive) rew = fmaxf(rew, 0.f);
  if (a.active[LGK_G_TERMINATION]) {                                  // reset_buf * ~time_out_buf (HLG:584-586)
    const float r = f_mul((prev_reset && !time_out) ? 1.f : 0.f, a.scale[LGK_G_TERMINATION]);
    rew = f_add(rew, r);
    a.sums[(size_t)a.slot[LGK_G_TERMINATION] * N + env] += r;
  }
  return rew;
}

__device__ __forceinline__ void game_observe_and_store(const LgkGameParams& p, int env, int e, bool dec, bool done, bool time_out,
                                                       long long ep_len, long long ces, float rew_prey, float rew_pred,
                                                       const float (&pr)[13], float px, float py, float pz, const float (&o)[16],
                                                       float* op, float* predrow, bool& occluded);

__global__ void __launch_bounds__(128) game_step_kernel(const __grid_constant__ LgkGameParams p) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = p.num_envs;
  const bool valid = env < N;
  const int e = valid ? env : N - 1;
  const bool dec = p.variant == 1;
  float* prey = p.root_states + (size_t)(2 * e) * 13;
  float* predrow = prey + 13;
  float pr[13];
#pragma unroll
  for (int i = 0; i < 13; ++i) pr[i] = prey[i];

  // ---- step_predator_single_integrator (HLG:265-287): `decimation` sequential fp32 adds of fl(sim_dt * u) to the game's
  // own copy of the predator position, then written over the simulator row (also over a re-spawn the low-level reset
  // may just have put there -- the reference does the same)
  const float* cmd = p.command_pred + (size_t)e * p.command_pred_stride;
  float px = p.predator_pos[3 * e], py = p.predator_pos[3 * e + 1], pz = p.predator_pos[3 * e + 2];
  const float dx = f_mul(p.sim_dt, cmd[0]), dy = f_mul(p.sim_dt, cmd[1]);
  for (int k = 0; k < p.decimation; ++k) { px = f_add(px, dx); py = f_add(py, dy); }

  const bool ro = p.reset_only != 0;
  long long ep_len = p.episode_length_buf[e], ces = p.curr_episode_step[e] + (ro ? 0 : 1);      // HLG:183, DHLG:243-244
  if (dec && !ro) ep_len += 1;
  const bool prev_reset = p.reset_buf[e] != 0;
  bool time_out = p.time_out_buf[e] != 0;

  // ---- rewards and terminations
  const float ex = f_sub(px, pr[0]), ey = f_sub(py, pr[1]), ez = f_sub(pz, pr[2]);
  const float dist3 = norm3(ex, ey, ez);                                              // HLG:574-582
  const float dist2 = norm2(f_sub(pr[0], px), f_sub(pr[1], py));                      // HLG:190
  bool done = dist2 < p.capture_dist;
  float rew_prey = 0.f, rew_pred = 0.f;
  const float base = f_mul(p.ll_rew_weight, p.ll_rews[e]);
  if (ro) {
    done = false; rew_prey = 0.f;
  } else if (dec) {                                                                   // check_termination precedes the rewards
    time_out = (float)ep_len > p.max_episode_length;                                  // DHLG:267
    done = done || time_out;
    if (valid) {
      rew_prey = game_reward(p.prey, base, dist3, done, time_out, env, N);
      rew_pred = game_reward(p.pred, 0.f, dist3, done, time_out, env, N);
    }
  } else if (valid) {
    rew_prey = game_reward(p.prey, base, dist3, prev_reset, time_out, env, N);       // previous step's reset_buf (HLG:202)
  }
  if (!dec && p.has_env_radius && !ro) {                                                     // HLG:197-210
    const float ox = p.env_origins[3 * e], oy = p.env_origins[3 * e + 1];
    done = done || norm2(f_sub(pr[0], ox), f_sub(pr[1], oy)) > p.env_radius || norm2(f_sub(px, ox), f_sub(py, oy)) > p.env_radius;
  }
  done = done || p.ll_dones[e] != 0;

  // ---- observation history before the reset touches it
  float* op = p.obs_prey + (size_t)e * p.obs_prey_stride;
  float o[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = op[i];

  // ---- reset_idx (HLG:326-349, DHLG:270-312)
  if (done) {
    const RngKey key = make_key(p.seed, p.step);
    const uint32_t genv = (uint32_t)(p.env_id_offset + env);
    if (p.reset_dofs && valid) {                                                      // LLG:384-399
      float* dof = p.dof_state + (size_t)e * 2 * kDof;
      for (int b = 0; b < kDof / 4; ++b) {
        const U4 r = rng_block(key, genv, LGK_STREAM_GAME_DOF, (uint32_t)b);
        for (int k = 0; k < 4; ++k) {
          const int d = 4 * b + k;
          dof[2 * d] = f_mul(p.default_dof_pos[d], scale_uniform(1.0f, 0.5f, u32_to_uniform(pick(r, k))));
          dof[2 * d + 1] = 0.f;
        }
      }
    }
    const U4 r0 = rng_block(key, genv, LGK_STREAM_GAME_ROOT, 0), r1 = rng_block(key, genv, LGK_STREAM_GAME_ROOT, 1);
#pragma unroll
    for (int i = 0; i < 13; ++i) pr[i] = p.base_init_state[i];                        // LLG:409-418
    pr[0] = f_add(pr[0], p.env_origins[3 * e]); pr[1] = f_add(pr[1], p.env_origins[3 * e + 1]); pr[2] = f_add(pr[2], p.env_origins[3 * e + 2]);
    if (p.custom_origins) {
      pr[0] = f_add(pr[0], scale_uniform(2.0f, -1.0f, u32_to_uniform(r0.x)));
      pr[1] = f_add(pr[1], scale_uniform(2.0f, -1.0f, u32_to_uniform(r0.y)));
    }
    pr[7] = scale_uniform(1.0f, -0.5f, u32_to_uniform(r0.z)); pr[8] = scale_uniform(1.0f, -0.5f, u32_to_uniform(r0.w));
    pr[9] = scale_uniform(1.0f, -0.5f, u32_to_uniform(r1.x)); pr[10] = scale_uniform(1.0f, -0.5f, u32_to_uniform(r1.y));
    pr[11] = scale_uniform(1.0f, -0.5f, u32_to_uniform(r1.z)); pr[12] = scale_uniform(1.0f, -0.5f, u32_to_uniform(r1.w));
    const U4 rp = rng_block(key, genv, LGK_STREAM_GAME_PREDATOR, 0);                  // LLG:419-432
    const float sgn = u32_to_uniform(rp.w) < 0.5f ? -1.f : 1.f;
    px = f_sub(pr[0], sgn * scale_uniform(9.0f, 1.0f, u32_to_uniform(rp.x)));
    py = f_sub(pr[1], sgn * scale_uniform(9.0f, 1.0f, u32_to_uniform(rp.y)));
    pz = 0.3f;
#pragma unroll
    for (int i = 0; i < 12; ++i) o[i] = p.max_rel_pos;                                // HLG:341-343
    o[12] = o[13] = o[14] = o[15] = 0.f;
    if (ro && valid) {
      float* od = p.obs_pred + (size_t)e * p.obs_pred_stride;
      od[0] = od[1] = od[2] = -p.max_rel_pos;
#pragma unroll
      for (int i = 0; i < 16; ++i) op[i] = o[i];
    }
    ep_len = 0; ces = 0;
    if (valid) {
#pragma unroll
      for (int i = 0; i < 13; ++i) prey[i] = pr[i];
    }
  }
  // DHLG:298-306: extras["episode"] means over the reset envs, from the pre-reset episode sums
  if (dec && p.reset_stats) {
    const bool r = done && valid;
    const uint32_t m = __ballot_sync(0xffffffffu, r);
    if (m) {
      const int lane = threadIdx.x & 31;
      for (int a = 0; a < 2; ++a) {
        const LgkGameAgent& ag = a == 0 ? p.prey : p.pred;
        const int ns = a == 0 ? p.num_prey_slots : p.num_pred_slots, off = a == 0 ? 0 : p.num_prey_slots;
        for (int k = 0; k < ns; ++k) {
          float v = 0.f;
          if (r) { v = ag.sums[(size_t)k * N + env]; ag.sums[(size_t)k * N + env] = 0.f; }
          v = warp_sum_g(v);
          if (lane == 0) atomicAdd(p.reset_stats + off + k, v);
        }
      }
      if (lane == 0) atomicAdd(p.reset_stats + p.num_prey_slots + p.num_pred_slots, (float)__popc(m));
    }
  }
  bool occluded = false;
  if (valid && !ro) game_observe_and_store(p, env, e, dec, done, time_out, ep_len, ces, rew_prey, rew_pred, pr, px, py, pz, o, op,
                                           predrow, occluded);
  if (valid && ro) {             // reset_idx alone: states and counters of the flagged envs, no observation update
    if (done) {
      predrow[0] = px; predrow[1] = py; predrow[2] = pz;
      p.predator_pos[3 * e] = px; p.predator_pos[3 * e + 1] = py; p.predator_pos[3 * e + 2] = pz;
      if (p.prey_states) { for (int i = 0; i < 13; ++i) p.prey_states[(size_t)e * 13 + i] = pr[i]; }
      if (dec) p.reset_buf[e] = 1;
      p.episode_length_buf[e] = 0;
      p.curr_episode_step[e] = 0;
    }
  }
  if (ro) return;
  // ---- env 0 and the flattened [K,2] id list of HLG:456-458: `any env occluded` puts index 0 into the occluded list.
  // CTA-wide OR -> global flag; the last CTA to arrive patches env 0's sensed position and re-arms the scratch words.
  const int any_occ = __syncthreads_or(occluded ? 1 : 0);
  if (threadIdx.x == 0) {
    if (any_occ) atomicOr(p.scratch, 1);
    __threadfence();
    const int ticket = atomicAdd(p.scratch + 1, 1);
    if (ticket == (int)gridDim.x - 1) {
      __threadfence();
      if (atomicOr(p.scratch, 0)) {
        const volatile int32_t* sv = p.scratch;
        p.obs_prey[9] = __int_as_float(sv[2]); p.obs_prey[10] = __int_as_float(sv[3]); p.obs_prey[11] = __int_as_float(sv[4]);
      }
      p.scratch[0] = 0; p.scratch[1] = 0;
    }
  }
}

__device__ __forceinline__ void game_observe_and_store(const LgkGameParams& p, int env, int e, bool dec, bool done, bool time_out,
                                                       long long ep_len, long long ces, float rew_prey, float rew_pred,
                                                       const float (&pr)[13], float px, float py, float pz, const float (&o)[16],
                                                       float* op, float* predrow, bool& occluded) {
  // ---- sense_predator (HLG:418-482): is the predator inside the prey's horizontal field of view?
  const float rx = f_sub(px, pr[0]), ry = f_sub(py, pr[1]), rz = f_sub(pz, pr[2]);
  const float4 q = *reinterpret_cast<const float4*>(p.base_quat + (size_t)e * 4);      // LowLevelGame.base_quat: pre-reset copy
  const YawFrame yf = yaw_frame(q.z, q.w, 0.f, 0.f);
  // quat_apply((0,0,zn,wn), (1,0,0)): t = (0, 2zn, 0); forward = (1 - zn*t1, wn*t1, 0)
  const float t1 = f_mul(yf.zn, 2.0f);
  const float fx = f_add(1.0f, -f_mul(yf.zn, t1)), fy = f_mul(yf.wn, t1);
  const float dot = f_add(f_add(f_mul(fx, rx), f_mul(fy, ry)), f_mul(0.f, rz));
  const float denom = f_mul(norm3(fx, fy, 0.f), norm3(rx, ry, rz));
  float ang = acosf(f_div(dot, denom));
  ang = wrap_to_pi(ang);
  const bool visible = fabsf(ang) <= p.half_fov;     // torch compares the fp32 tensor with the python scalar rounded to fp32
  float sx = rx, sy = ry, sz = rz;
  if (!visible) { sx = o[9]; sy = o[10]; sz = o[11]; }           // keep the last sensed position (HLG:455-458)
  occluded = !visible;
  if (env == 0) {                                                // see the tail of the kernel
    p.scratch[2] = __float_as_int(o[9]); p.scratch[3] = __float_as_int(o[10]); p.scratch[4] = __float_as_int(o[11]);
  }
  // history shift (HLG:394-409)
#pragma unroll
  for (int i = 0; i < 9; ++i) op[i] = o[i + 3];
  op[9] = sx; op[10] = sy; op[11] = sz;
  op[12] = o[13]; op[13] = o[14]; op[14] = o[15];
  op[15] = visible ? 1.f : 0.f;
  float* od = p.obs_pred + (size_t)e * p.obs_pred_stride;      // prey - predator (HLG:391, DHLG:389-391)
  od[0] = f_sub(pr[0], px); od[1] = f_sub(pr[1], py); od[2] = f_sub(pr[2], pz);

  // ---- state write-back
  predrow[0] = px; predrow[1] = py; predrow[2] = pz;
  p.predator_pos[3 * e] = px; p.predator_pos[3 * e + 1] = py; p.predator_pos[3 * e + 2] = pz;
  if (p.prey_states) {
#pragma unroll
    for (int i = 0; i < 13; ++i) p.prey_states[(size_t)e * 13 + i] = pr[i];
  }
  p.prey.rew[e] = rew_prey;
  if (dec && p.pred.rew) p.pred.rew[e] = rew_pred;
  p.reset_buf[e] = done ? 1 : 0;
  p.time_out_buf[e] = time_out ? 1 : 0;
  p.episode_length_buf[e] = ep_len;
  p.curr_episode_step[e] = ces;
}

// HLG:161-174 / DHLG:181-195: clip the prey's and the predator's high-level commands in place (torch.clip semantics: NaN
// stays NaN), wrap the prey's heading command, and hand the prey command to the low-level env's command buffer.
__device__ __forceinline__ float clip_keep_nan(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

__global__ void game_prepare_kernel(float* prey, long long prey_stride, float* pred, long long pred_stride, float* ll_commands,
                                    int n, float x0, float x1, float y0, float y1, float px0, float px1, float py0, float py1,
                                    int heading_command) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float* a = prey + (size_t)e * prey_stride;
  float* b = pred + (size_t)e * pred_stride;
  const float c0 = clip_keep_nan(a[0], x0, x1), c1 = clip_keep_nan(a[1], y0, y1);
  const float c2 = heading_command ? wrap_to_pi(a[2]) : a[2];
  a[0] = c0; a[1] = c1; a[2] = c2;
  b[0] = clip_keep_nan(b[0], px0, px1);
  b[1] = clip_keep_nan(b[1], py0, py1);
  *reinterpret_cast<float4*>(ll_commands + 4 * (size_t)e) = make_float4(c0, c1, c2, a[3]);
}

}  // namespace lgk

using namespace lgk;

extern "C" int lgk_game_prepare(float* command_prey, int64_t prey_stride, float* command_pred, int64_t pred_stride,
                                float* ll_commands, int32_t num_envs, const float* ranges8, int32_t heading_command, void* stream) {
  LGK_REQUIRE(command_prey && command_pred && ll_commands && ranges8 && num_envs > 0, "game_prepare: bad arguments");
  LGK_REQUIRE(prey_stride >= 4 && pred_stride >= 2, "game_prepare: bad row stride");
  LGK_ALIGNED16(ll_commands, "ll_commands");
  game_prepare_kernel<<<(num_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      command_prey, prey_stride, command_pred, pred_stride, ll_commands, num_envs, ranges8[0], ranges8[1], ranges8[2], ranges8[3],
      ranges8[4], ranges8[5], ranges8[6], ranges8[7], heading_command);
  count_launch();
  return check_cuda(cudaGetLastError(), "game_prepare_kernel launch");
}

extern "C" int lgk_game_step(const LgkGameParams* p, void* stream) {
  LGK_REQUIRE(p != nullptr, "game params is null");
  LGK_REQUIRE(p->num_envs > 0, "num_envs must be positive");
  LGK_REQUIRE(p->variant == 0 || p->variant == 1, "variant must be 0 (HighLevelGame) or 1 (DecHighLevelGame)");
  LGK_REQUIRE(p->decimation >= 0, "decimation negative");
  LGK_REQUIRE(p->root_states && p->env_origins && p->base_quat && p->command_pred && p->ll_rews && p->ll_dones &&
              p->predator_pos && p->obs_prey && p->obs_pred && p->reset_buf && p->time_out_buf && p->episode_length_buf &&
              p->curr_episode_step && p->prey.rew && p->scratch, "a required game buffer is null");
  LGK_REQUIRE(!p->reset_dofs || p->dof_state, "dof_state is null");
  LGK_REQUIRE(p->variant == 0 || p->pred.rew, "predator reward buffer is null");
  LGK_REQUIRE(p->obs_prey_stride >= 16 && p->obs_pred_stride >= 3 && p->command_pred_stride >= 2, "bad row stride");
  const LgkGameAgent* ag[2] = {&p->prey, &p->pred};
  for (int a = 0; a < 2; ++a)
    for (int k = 0; k < LGK_G_COUNT; ++k)
      if (ag[a]->active[k]) LGK_REQUIRE(ag[a]->sums && ag[a]->slot[k] >= 0, "episode sums of an active game reward term missing");
  LGK_ALIGNED16(p->base_quat, "base_quat");
  game_step_kernel<<<(p->num_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*p);
  count_launch();
  return check_cuda(cudaGetLastError(), "game_step_kernel launch");
}
