/* lgk.h -- C ABI of liblgk.so: the B200 (sm_100a) kernels behind legged_gym's per-environment step.
 *
 * Every entry point takes plain device pointers, sizes and a cudaStream_t (as void*), returns
 * 0 on success or a non-zero code (LGK_ERR_* or 1000+cudaError_t), never throws, never allocates
 * and never synchronises: all work is stream-ordered and CUDA-graph capturable.
 * Citations are to /root/reference/legged_gym/... : LR = envs/base/legged_robot.py,
 * ANY = envs/anymal_c/anymal.py, MATH = utils/math.py, CAS = envs/cassie/cassie.py.
 * rsl_rl (not vendored in the reference; call sites utils/task_registry.py:37-38,154) is cited by
 * function name.
 *
 * Tensor layouts are the reference's own: env-major row-major fp32 unless noted.
 */
#ifndef LGK_H
#define LGK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGK_ABI_VERSION 7
#define LGK_NUM_DOF 12          /* every registered task has 12 DOF = 12 actions */
#define LGK_MAX_FEET 4
#define LGK_MAX_PEN 16
#define LGK_MAX_TERM 8
#define LGK_MAX_BODIES 32

enum {
  LGK_OK = 0,
  LGK_ERR_ARG = 1,          /* bad size / null pointer / unsupported configuration */
  LGK_ERR_ALIGN = 2,        /* a pointer violates the 16-byte alignment the kernels assume */
  LGK_ERR_CUDA_BASE = 1000  /* 1000 + cudaError_t */
};

/* reward terms, in the ALPHABETICAL order the reference sums them (helpers.py:41-56 walks dir();
 * LR:583-607).  "termination" is applied after the positive clip (LR:204-210). */
enum {
  LGK_R_ACTION_RATE = 0, LGK_R_ANG_VEL_XY, LGK_R_BASE_HEIGHT, LGK_R_COLLISION, LGK_R_DOF_ACC,
  LGK_R_DOF_POS_LIMITS, LGK_R_DOF_VEL, LGK_R_DOF_VEL_LIMITS, LGK_R_FEET_AIR_TIME,
  LGK_R_FEET_CONTACT_FORCES, LGK_R_LIN_VEL_Z, LGK_R_NO_FLY, LGK_R_ORIENTATION, LGK_R_STAND_STILL,
  LGK_R_STUMBLE, LGK_R_TERMINATION, LGK_R_TORQUE_LIMITS, LGK_R_TORQUES, LGK_R_TRACKING_ANG_VEL,
  LGK_R_TRACKING_LIN_VEL, LGK_R_COUNT
};

/* random streams of the counter-based generator (Philox4x32-10, key = seed,
 * counter = (global env id, word_index/4, stream, step)); uniform = (word >> 8) * 2^-24. */
enum {
  LGK_STREAM_CMD = 0, LGK_STREAM_PUSH = 1, LGK_STREAM_RESET_DOF = 2, LGK_STREAM_RESET_ROOT = 3,
  LGK_STREAM_RESET_CMD = 4, LGK_STREAM_TERRAIN = 5, LGK_STREAM_OBS = 6, LGK_STREAM_ACT = 7,
  LGK_STREAM_PREDATOR = 8,
  /* the high-level games reset the low-level root / dof state a second time within one step (HLG:326-347, DHLG:270-296) */
  LGK_STREAM_GAME_ROOT = 9, LGK_STREAM_GAME_PREDATOR = 10, LGK_STREAM_GAME_DOF = 11
};

enum { LGK_CTRL_P = 0, LGK_CTRL_V = 1, LGK_CTRL_T = 2 };

/* phases of lgk_post_physics (bit mask).  PRE = LR:114-127 up to the per-term reward sum;
 * POST = positive clip + termination term (LR:204-210), reset_idx (LR:147-191), observations
 * (LR:212-230) and the history copies (LR:132-134).  PRE|POST runs in one pass; the host splits them only when
 * Python code must run in between (user reward terms, command curriculum).  When a subclass overrides reset_idx
 * (LR:128-129 calls it every step; ANY:56-60 is such an override) POST is split once more so that the Python
 * method runs exactly where the reference calls it: POST_REWARD = the reward-finishing part of POST alone,
 * POST_OBS = observations + histories alone (no in-kernel reset: the caller has run lgk_reset_idx in between). */
enum { LGK_PHASE_PRE = 1, LGK_PHASE_POST = 2, LGK_PHASE_POST_REWARD = 4, LGK_PHASE_POST_OBS = 8 };

/* ------------------------------------------------------------------ torques (LR:371-395, ANY:71-81) */
typedef struct LgkTorqueParams {
  int32_t num_envs;
  int32_t control_type;               /* LGK_CTRL_*; ignored when lstm != 0 */
  int32_t use_lstm;                   /* ANY:73 cfg.control.use_actuator_network */
  float action_scale;                 /* cfg.control.action_scale */
  float clip_actions;                 /* LR:86-87: actions are clipped to +-clip on load */
  int32_t lstm_variant;               /* 0 = auto (by batch size), 1 = one thread per sequence, 2 = role-split CTA,
                                       * 3 = one thread per sequence with packed-fp32 gate arithmetic (auto's choice up to 12 628 envs) */
  float sim_dt;                       /* LR:390 divides by sim_params.dt ("V" control) */
  float p_gains[LGK_NUM_DOF];         /* LR:566-580 */
  float d_gains[LGK_NUM_DOF];
  float torque_limits[LGK_NUM_DOF];   /* LR:395 clip (PD path only; ANY:77-78 does not clip) */
  float default_dof_pos[LGK_NUM_DOF]; /* LR:565-581 */
  const float* actions_in;            /* [N,12] raw policy output */
  float* actions_clipped;             /* [N,12] or NULL: self.actions (LR:87) */
  const float* dof_state;             /* [N*12,2] (pos,vel) interleaved, LR:524-526 */
  const float* last_dof_vel;          /* [N,12] ("V" control only) */
  float* torques;                     /* [N,12] out */
  float* sea_hidden_state;            /* [2,N*12,8] r/w (ANY:65-69) */
  float* sea_cell_state;              /* [2,N*12,8] r/w */
  /* Host-sim pipelines: `dof_state` may point at PINNED host memory (the kernel pulls it over PCIe with uncached loads),
   * and `torques_mirror` (optional, e.g. pinned host memory) receives a second copy of the torques, so that a sub-step
   * of LR:88-96 is one launch instead of copy-in, kernel, copy-out. */
  float* torques_mirror;
  int32_t host_io;                    /* 1: dof_state is (or may be) host memory / torques_mirror is used; 0: device-resident fast path */
  int32_t pad_;
} LgkTorqueParams;

/* LSTM actuator weights (resources/actuator_nets/anydrive_v3_lstm.pt: LSTM(2,8,layers=2) + Linear(8,1),
 * buffers in_scale[2], out_scale[1]); torch gate order i,f,g,o.  Uploaded once into __constant__. */
typedef struct LgkLstmWeights {
  float w_ih0[32 * 2], w_hh0[32 * 8], b_ih0[32], b_hh0[32];
  float w_ih1[32 * 8], w_hh1[32 * 8], b_ih1[32], b_hh1[32];
  float lin_w[8], lin_b[1];
  float in_scale[2], out_scale[1];
} LgkLstmWeights;

int lgk_set_lstm_weights(const LgkLstmWeights* host_weights, void* stream);
int lgk_compute_torques(const LgkTorqueParams* p, void* stream);

/* ------------------------------------------------------------------ post-physics step (LR:106-230, 329-508, 831-969) */
typedef struct LgkStepParams {
  /* sizes */
  int32_t num_envs;
  int32_t num_bodies;                 /* NB: contact_forces is [N, NB, 3] (LR:529) */
  int32_t num_obs;                    /* O = 48 + P when measure_heights else 48 */
  int32_t num_height_points;          /* P (0 when !measure_heights) */
  int32_t actors_per_env;             /* root_states rows per env (1; 2 in low_level_game, LLG:532) */
  int32_t root_actor_offset;          /* row of the robot inside the env's actor group (prey index) */
  int32_t phase_mask;                 /* LGK_PHASE_* */
  int32_t host_state;                 /* 1: root_states / dof_state / contact_forces live in pinned HOST memory (the kernels reach
                                       * them over the unified address space): no L2 prefetch of those chunks, every byte crosses
                                       * PCIe exactly once */
  int32_t push_interval;              /* used with step_counter_dev: push when step % interval == 0 (0 = never) */
  int32_t step;                       /* common_step_counter AFTER the += 1 of LR:115 (RNG counter) */
  uint64_t seed;
  int64_t env_id_offset;              /* global id of env 0 (multi-GPU sharding keeps streams disjoint) */

  /* flags */
  int32_t heading_command;            /* cfg.commands.heading_command (LR:337, 359) */
  int32_t measure_heights;            /* cfg.terrain.measure_heights (LR:342) */
  int32_t terrain_is_plane;           /* mesh_type == 'plane' -> zeros (LR:844-845) */
  int32_t do_push;                    /* push_robots && counter % push_interval == 0 (LR:344) */
  int32_t add_noise;                  /* LR:229 */
  int32_t only_positive_rewards;      /* LR:204 */
  int32_t terrain_curriculum;         /* LR:160 (cfg.terrain.curriculum after _parse_cfg LR:786-787) */
  int32_t custom_origins;             /* LR:422 */
  int32_t send_timeouts;              /* LR:190 */
  int32_t zero_lstm_on_reset;         /* ANY:56-60 */
  /* low_level_game (envs/a1_game/low_level_game.py:419-432): on reset the env's second actor (a sphere "predator") is
   * re-spawned at prey_pos - sign * U(1,10)^3 with one random sign per env, z forced to 0.3; only its position columns
   * are written.  predator_actor_offset = its row inside the env's actor group. */
  int32_t predator_spawn;
  int32_t predator_actor_offset;

  /* scalars */
  float dt;                           /* decimation * sim.dt (LR:782) */
  int32_t resample_period;            /* int(resampling_time / dt) (LR:334) */
  float max_episode_length;           /* np.ceil(episode_length_s / dt) (LR:789); compared with > (LR:144) */
  float max_episode_length_s;
  float max_push_vel;                 /* LR:441 */
  float cmd_lo[4], cmd_range[4];      /* x, y, yaw, heading: lo and (float)(hi - lo) (LR:353-366) */
  float obs_scale_lin_vel, obs_scale_ang_vel, obs_scale_dof_pos, obs_scale_dof_vel, obs_scale_height;
  float clip_obs;                     /* LR:100-101 */
  float tracking_sigma, base_height_target, max_contact_force;
  float soft_dof_vel_limit, soft_torque_limit;
  float border_size, horizontal_scale, vertical_scale;   /* LR:856-857, 869 */
  float horizontal_scale_recip;       /* RN(1/horizontal_scale) when the FMA division is proven exact for that scale
                                         (oracle/divcheck.c: 0.1f, 0.05f, 0.25f); 0 = use IEEE division */
  int32_t hf_rows, hf_cols;           /* height_samples.shape */
  float half_env_length;              /* terrain.env_length / 2 (LR:458) */
  int32_t max_terrain_level;          /* LR:766 */
  int32_t terrain_num_cols;           /* terrain_origins is [rows, cols, 3] */

  /* per-robot constants */
  float default_dof_pos[LGK_NUM_DOF];
  float dof_pos_lo[LGK_NUM_DOF], dof_pos_hi[LGK_NUM_DOF];  /* soft limits LR:310-313 */
  float dof_vel_limits[LGK_NUM_DOF], torque_limits[LGK_NUM_DOF];
  float base_init_state[13];          /* LR:704-705 */
  int32_t num_feet, num_pen, num_term;
  int32_t feet_idx[LGK_MAX_FEET], pen_idx[LGK_MAX_PEN], term_idx[LGK_MAX_TERM];
  float reward_scale[LGK_R_COUNT];    /* already multiplied by dt (LR:593); 0 = term inactive */
  int32_t reward_active[LGK_R_COUNT]; /* non-zero scale in cfg (LR:588-593) */
  int32_t reward_slot[LGK_R_COUNT];   /* row of episode_sums for the term, -1 if inactive */
  int32_t num_reward_slots;

  /* sim-owned state (LR:523-530) */
  float* root_states;                 /* [N*actors_per_env, 13] pos3 quat4(xyzw) linvel3 angvel3 */
  float* dof_state;                   /* [N*12, 2] */
  const float* contact_forces;        /* [N*NB, 3] */
  /* env-owned state */
  const float* actions;               /* [N,12] clipped (LR:87) */
  const float* torques;               /* [N,12] last sub-step's torques (LR:91) */
  float* commands;                    /* [N,4] */
  int64_t* episode_length_buf;        /* [N] */
  float* last_actions;                /* [N,12] */
  float* last_dof_vel;                /* [N,12] */
  float* last_root_vel;               /* [N,6] */
  float* feet_air_time;               /* [N,F] */
  uint8_t* last_contacts;             /* [N,F] bool */
  float* episode_sums;                /* [num_reward_slots, N] one row per active term */
  int64_t* terrain_levels;            /* [N] (curriculum only) */
  const int64_t* terrain_types;       /* [N] */
  const float* terrain_origins;       /* [rows, cols, 3] */
  float* env_origins;                 /* [N,3] */
  float* sea_hidden_state;            /* [2,N*12,8] or NULL */
  float* sea_cell_state;
  /* outputs */
  float* base_lin_vel;                /* [N,3] */
  float* base_ang_vel;                /* [N,3] */
  float* projected_gravity;           /* [N,3] */
  float* measured_heights;            /* [N,P] */
  float* obs_buf;                     /* [N,O], already clipped to +-clip_obs */
  float* rew_buf;                     /* [N] */
  uint8_t* reset_buf;                 /* [N] bool */
  uint8_t* time_out_buf;              /* [N] bool */
  /* constants in device memory */
  const int16_t* height_min3;         /* [hf_rows, hf_cols] from lgk_height_min3 (see below) */
  const float* height_points_xy;      /* [P,2] base-frame grid (LR:815-829), x outer / y inner */
  const float* noise_scale_vec;       /* [O] (LR:485-508) */
  /* cross-env accumulators, two slots ping-ponged on (step & 1): [2][num_reward_slots + 2] floats:
   * per-term sum of episode_sums over the envs reset this step, then reset count, then
   * sum(terrain_levels) over all envs (LR:181-186) */
  float* reset_stats;
  /* [N,8] fp32 scratch handed from the scalar kernel to the scan/observation kernel: pre-reset yaw frame
   * (zn, wn, root_x, root_y) and post-reset (root_z - 0.5); required when measure_heights on a height field */
  float* scan_frames;
  /* optional device-resident step counter (number of completed steps).  When non-NULL it overrides `step`
   * (post_physics / finalize(advance=1) use counter+1, reset_idx uses counter) and `do_push` (derived from
   * push_interval), and finalize(advance=1) stores counter+1: a whole env step then has no per-step host-side
   * parameter and can be replayed as one CUDA graph.  lgk_post_physics_finalize needs room for a second int32. */
  int32_t* step_counter_dev;
  /* optional output [N,4]: the root quaternion (xyzw) this step's rotations used, i.e. the pose BEFORE reset_idx.
   * LowLevelGame keeps exactly this copy as `base_quat` until its next step (LLG:123), and the high-level games read
   * it after the reset (HLG:432). */
  float* base_quat;
  /* [N,48] fp32 hand-over buffer between the two step kernels: K1 leaves the 48 proprioceptive observation columns
   * (LR:212-222, before noise) here as one contiguous block per 32-env tile and K2 finishes them into obs_buf.
   * NULL: the columns go through obs_buf itself (192-byte fragments at the row stride). */
  float* obs_head;
} LgkStepParams;

int lgk_post_physics(const LgkStepParams* p, void* stream);

/* reset_idx(env_ids) on an explicit id list (LR:147-191), e.g. BaseTask.reset() (base_task.py:114-118).
 * Accumulates into the same reset_stats slot as a step with p->step. */
int lgk_reset_idx(const LgkStepParams* p, const int64_t* env_ids, int32_t num_ids, void* stream);

/* Profiling hook: when non-NULL, CTA 0 of the scalar post-physics kernel writes %globaltimer stamps (ns) into
 * device_buf16[0..8]: entry, step counter read, tile staged, rewards done, reset done, outputs staged, bulk stores issued,
 * rows written, tile done (first tile of the persistent CTA); [16..24]: the same nine stamps of the LAST CTA.  The
 * buffer holds 32 int64. */
int lgk_step_debug_timeline(int64_t* device_buf32);

/* After lgk_post_physics / lgk_reset_idx: single-CTA pass that (a) compacts reset_buf into an ascending
 * int32 id list + count (what set_*_tensor_indexed needs, LR:409-412, 433-436), (b) if count > 0 writes the
 * extras the reference refreshes only when something was reset (LR:157-158, 179-191):
 * episode_means[slot] = sum/count/max_episode_length_s, episode_means[nslots] = mean terrain level, and a copy
 * of time_out_buf into time_outs_extras; (c) clears the other ping-pong slot of reset_stats. */
int lgk_finalize_step(const LgkStepParams* p, int32_t* reset_ids, int32_t* reset_count,
                      float* episode_means, uint8_t* time_outs_extras, int32_t advance, void* stream);

/* lgk_post_physics(PRE | POST) followed by lgk_finalize_step(advance = 1), as one call: the whole of LR:106-137 + the
 * id list / extras of LR:147-191.  Same results as the two calls.  On the rough-terrain path (height field, scan +
 * observations in one K2 pass, hand-over buffers present) the finalize pass does not end the step's dependent chain as
 * a kernel of its own but rides in K2's grid as one more CTA; K1 then hands the step number to K2 through word [1]
 * of the counter, so step_counter_dev, when non-NULL, must point at TWO int32 here ([0] completed steps, [1] scratch).
 * Every other configuration (flat terrain, two actors per env, a plane under a height scan, more than 131 072 envs, the
 * base_height reward, which needs the scan before K1) runs the two calls back to back. */
int lgk_post_physics_finalize(const LgkStepParams* p, int32_t* reset_ids, int32_t* reset_count,
                              float* episode_means, uint8_t* time_outs_extras, void* stream);

/* ------------------------------------------------------------------ hierarchical predator / prey games
 * HighLevelGame.step after ll_env.step (legged_gym/envs/a1_game/high_level_game.py:178-239 = HLG) and
 * DecHighLevelGame.step / post_physics_step (dec_high_level_game.py:204-261 = DHLG), one thread per env: predator
 * single integrator (HLG:265-287), agent states (HLG:510-517), rewards evasion / pursuit / termination on top of the
 * low-level reward (HLG:357-378, DHLG:321-361), capture / radius / time-out / low-level dones (HLG:190-236,
 * DHLG:263-268), reset of the low-level root (and, DHLG, dof) state with predator re-spawn (LLG:384-451) and of the
 * observation history (HLG:341-347), field-of-view sensing and the 4-frame history (HLG:380-482, DHLG:364-472).
 * variant 0 = HighLevelGame (obs_prey and obs_pred point into one [N,19] buffer, columns 0 and 16; one reward: rew_prey
 * with BOTH agents' terms in `prey`), 1 = DecHighLevelGame.  Reward terms are indexed LGK_G_EVASION.. in the
 * reference's alphabetical order. */
enum { LGK_G_EVASION = 0, LGK_G_PURSUIT = 1, LGK_G_TERMINATION = 2, LGK_G_COUNT = 3 };

typedef struct {
  int32_t active[LGK_G_COUNT];        /* scale != 0 */
  float scale[LGK_G_COUNT];           /* cfg scale * low-level dt (HLG:546) */
  int32_t slot[LGK_G_COUNT];          /* row of `sums` */
  int32_t only_positive;
  float* sums;                        /* [num slots, N] episode sums */
  float* rew;                         /* [N] */
} LgkGameAgent;

typedef struct {
  int32_t num_envs, variant, decimation, custom_origins;
  int32_t step;                       /* RNG counter word (the game's own step count) */
  int32_t has_env_radius, reset_dofs;
  int32_t reset_only;                 /* 1: reset_idx on the envs flagged in ll_dones and nothing else (HLG:351-355 reset()) */
  uint64_t seed;
  int64_t env_id_offset;
  float sim_dt, capture_dist, env_radius, max_episode_length, half_fov, max_rel_pos, ll_rew_weight, pad0;
  float base_init_state[13];
  float default_dof_pos[LGK_NUM_DOF];
  float pad1[3];
  LgkGameAgent prey, pred;            /* variant 0 uses `prey` only (rew = ll_rew_weight * ll_rews + terms) */
  float* root_states;                 /* [2N,13]: prey row 2i, predator row 2i+1 (LLG:799-812) */
  float* dof_state;                   /* [N*12,2] (reset_dofs only) */
  const float* env_origins;           /* [N,3] of the low-level env */
  const float* base_quat;             /* [N,4] LowLevelGame.base_quat (pre-reset copy, see LgkStepParams.base_quat) */
  const float* command_pred;          /* predator command rows, 2 used floats */
  int64_t command_pred_stride;        /* floats between rows (6 for HLG's [N,6] command, 2 for DHLG) */
  const float* ll_rews;               /* [N] */
  const uint8_t* ll_dones;            /* [N] bool: the low-level reset_buf */
  float* predator_pos;                /* [N,3] the game's own copy, carried between steps (HLG:280-286, 517) */
  float* prey_states;                 /* [N,13] out (HLG:512) */
  float* obs_prey; int64_t obs_prey_stride;   /* 16 columns used */
  float* obs_pred; int64_t obs_pred_stride;   /* 3 columns used */
  uint8_t* reset_buf;                 /* [N] bool in/out (the termination reward reads the PREVIOUS value) */
  uint8_t* time_out_buf;              /* [N] bool */
  int64_t* episode_length_buf;        /* [N] */
  int64_t* curr_episode_step;         /* [N] */
  float* reset_stats;                 /* [prey slots + pred slots + 1] zeroed by the caller: sums of the episode sums
                                       * over the envs reset this step, then the reset count (DHLG:298-306) */
  int32_t num_prey_slots, num_pred_slots;
  /* 8 int32 of device scratch, zero at allocation and left zero by every launch.  sense_predator flattens a [K,2]
   * nonzero() result (HLG:456-458), so env 0 is in its "occluded" id list whenever ANY env is occluded and keeps its
   * previous sensed position then; the last CTA to finish applies that to env 0's row. */
  int32_t* scratch;
} LgkGameParams;

int lgk_game_step(const LgkGameParams* p, void* stream);
/* HLG:161-174 / DHLG:181-195 before ll_env.step: clip the prey command (x, y) and the predator command (x, y) in place to
 * ranges8 = {lin_vel_x lo, hi, lin_vel_y lo, hi, predator_lin_vel_x lo, hi, predator_lin_vel_y lo, hi} (host array),
 * wrap the prey's heading command to (-pi, pi] when heading_command, and copy the prey's four commands into the
 * low-level env's command buffer [N,4].  Strides in floats. */
int lgk_game_prepare(float* command_prey, int64_t prey_stride, float* command_pred, int64_t pred_stride,
                     float* ll_commands, int32_t num_envs, const float* ranges8, int32_t heading_command, void* stream);

/* ------------------------------------------------------------------ height field (LR:831-869) */
/* min3[r,c] = min(hs[r,c], hs[r+1,c], hs[r,c+1]) for r<=rows-2, c<=cols-2 (the three samples LR:863-867
 * takes after clipping px<=rows-2, py<=cols-2); border row/col copy hs.  Built once per terrain. */
int lgk_height_min3(const int16_t* height_samples, int16_t* min3, int32_t rows, int32_t cols, void* stream);

/* Stand-alone _get_heights (LR:831-869 + MATH:38-42): measured_heights[N,P] and optionally the clipped
 * integer sample indices px,py [N,P] int32 (parity tests assert those bit-exactly). */
int lgk_height_scan(const float* root_states, int32_t actors_per_env, int32_t root_actor_offset,
                    int32_t num_envs, const float* height_points_xy, int32_t num_points,
                    const int16_t* height_min3, int32_t rows, int32_t cols, float border_size,
                    float horizontal_scale, float vertical_scale, float* measured_heights,
                    int32_t* px_out, int32_t* py_out, void* stream);

/* ------------------------------------------------------------------ RNG tap */
/* Dump the uniforms / raw words a step consumes: out[n,count] for global env ids
 * env_id_offset..+n.  kind 0 = float32 uniforms, 1 = raw uint32 words, 2 = Box-Muller normals (ACT). */
int lgk_rng_dump(uint64_t seed, int32_t step, int64_t env_id_offset, int32_t num_envs, int32_t stream_id,
                 int32_t count, int32_t kind, void* out, void* stream);

/* ------------------------------------------------------------------ rsl_rl: ActorCritic.act / evaluate */
/* Fused actor + critic MLP forward (ELU), Normal(mean, std).sample(), log_prob.sum(-1)
 * (rsl_rl ActorCritic.act / evaluate / get_actions_log_prob; PPO.act).  Weights are nn.Linear layout
 * [out,in] fp32.  The three hidden layers run on tcgen05 tensor cores (TF32 operands rounded to nearest, FP32
 * accumulation in TMEM, one CTA per 128-env tile and network, activations never leave tensor memory); the last layer,
 * biases, ELU and the distribution epilogue are FP32.  The tensor-core kernel covers hidden[0] in {128, 256, 512},
 * hidden[1] in {64, 128, 256} with hidden[0]/2 + hidden[1] <= 512, hidden[2] a multiple of 32 up to min(128, hidden[0]/2),
 * observations up to 256 wide, up to 16 actions, 16-byte aligned biases; every other shape runs an FP32 tiled-GEMM path
 * with the same results to 1e-3. */
typedef struct LgkPolicyParams {
  int32_t num_envs, num_obs, num_critic_obs, num_actions;
  int32_t hidden[3];                  /* actor and critic hidden sizes (must match pairwise) */
  const float* obs;                   /* [N, num_obs] */
  const float* critic_obs;            /* [N, num_critic_obs] (== obs when no privileged obs) */
  const float* actor_w[4];            /* [h0,O] [h1,h0] [h2,h1] [A,h2] */
  const float* actor_b[4];
  const float* critic_w[4];           /* [h0,Oc] [h1,h0] [h2,h1] [1,h2] */
  const float* critic_b[4];
  const float* std;                   /* [A] */
  uint64_t seed; int32_t step; int64_t env_id_offset;
  int32_t sample;                     /* 1: actions = mu + std*eps (act); 0: actions = mu (act_inference) */
  float* actions;                     /* [N,A] */
  float* action_mean;                 /* [N,A] */
  float* action_sigma;                /* [N,A] */
  float* values;                      /* [N,1] */
  float* actions_log_prob;            /* [N] */
  void* workspace; int64_t workspace_bytes;   /* see lgk_policy_workspace_bytes */
  /* 0: weights are re-packed (TF32, swizzled tiles) into the workspace on every call.  Non-zero: a caller-maintained
   * version of the weight tensors; the packed image in `workspace` is reused while version, workspace and shapes
   * are unchanged (the Python ActorCritic passes the sum of the parameters' torch version counters + 1). */
  int64_t weights_version;
  /* which networks run: 0 or 3 = actor and critic (PPO.act), 1 = actor only (act / act_inference: critic_obs, values may be
   * NULL and num_critic_obs columns are never read), 2 = critic only (evaluate: obs and the action outputs may be NULL) */
  int32_t nets; int32_t pad_;
} LgkPolicyParams;

int64_t lgk_policy_workspace_bytes(const LgkPolicyParams* p);
int lgk_policy_act(const LgkPolicyParams* p, void* stream);
/* 0 = auto (tcgen05 when the shape fits), 1 = force the FP32 path, 2 = require tcgen05 (error if the shape does not
 * fit).  Process-wide; returns the previous value.  For tests and benchmarks. */
int lgk_policy_set_variant(int variant);
/* Profiling hook: when non-NULL, CTA (0,0) of the tcgen05 kernel writes %globaltimer stamps (ns) of its phases into
 * device_buf128[0..13] (0 setup, 1 obs staged, 2 first half of layer 1 accumulated, 3-4 its two column groups activated in
 * place, 5-7 the same for the second half, 8 layer 2 done, 9-10 its groups activated, 11 layer 3 done, 12 last layer summed,
 * 13 outputs written); slots 16..55 take the issuer's clock64 at the start of each weight tile, 100..103 its cycles spent
 * waiting for activations / for weight tiles / issuing MMAs / committing.  The buffer must hold 128 entries.
 * flags (profiling experiments only, results are then meaningless): bit 0 skips the weight-tile copies, bit 1 the MMAs. */
int lgk_policy_debug_timeline(int64_t* device_buf128, int flags);

/* ------------------------------------------------------------------ rsl_rl: RolloutStorage.compute_returns */
/* Reverse GAE scan over [T,N] + global advantage normalisation (unbiased std, +1e-8).
 * rewards/values/returns/advantages: [T,N] fp32 (the [T,N,1] storage tensors); dones: [T,N] uint8;
 * last_values: [N].  scratch: >= 4 doubles, zeroed by the call. */
int lgk_gae(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
            int32_t T, int32_t N, float gamma, float lam, float* returns, float* advantages,
            double* scratch, void* stream);

/* rsl_rl OnPolicyRunner's episode bookkeeping of one rollout step (cur_reward_sum / cur_episode_length / rewbuffer /
 * lenbuffer of on_policy_runner.py, the finished episodes reduced to three running sums): cur_return += rewards,
 * cur_length += 1; envs with dones != 0 add (cur_return, cur_length, 1) to stats3 (double[3], device) and restart at 0. */
int lgk_episode_stats(const float* rewards, const uint8_t* dones, float* cur_return, float* cur_length,
                      double* stats3, int32_t n, void* stream);

/* ------------------------------------------------------------------ misc */
const char* lgk_last_error_string(void);
int lgk_abi_version(void);
/* Programmatic dependent launch for the kernels of the env step (default on): each kernel of the chain may begin its
 * prologue while its predecessor on the stream drains, and waits (griddepcontrol.wait) before touching memory.
 * Returns the previous setting. */
int lgk_set_pdl(int enable);
/* Host-sim pipeline (the reference's sim_device=cpu: PhysX results are host tensors, LR:515-530 hand back host memory):
 * stream-ordered, graph-capturable copies between PINNED (cudaHostAlloc / torch pin_memory) host memory and device
 * memory done by a kernel over the unified address space instead of a memcpy node.  bytes % 16 == 0, 16-byte aligned. */
int lgk_copy_from_pinned(void* dst_device, const void* src_pinned_host, int64_t bytes, void* stream);
int lgk_copy_to_pinned(void* dst_pinned_host, const void* src_device, int64_t bytes, void* stream);
/* The indexed variant (gym.set_dof_state_tensor_indexed / set_actor_root_state_tensor_indexed, LR:409-412, 433-436) for a
 * host-resident sim: rows `ids[i] * row_stride + row_offset`, i < min(*count, max_ids), of a [rows, row_floats] fp32
 * tensor go from device to the same rows of its pinned host twin.  ids / count: device int32 (lgk_finalize_step). */
int lgk_copy_rows_to_pinned(void* dst_pinned_host, const void* src_device, int32_t row_floats, const int32_t* ids,
                            const int32_t* count, int32_t row_stride, int32_t row_offset, int32_t max_ids, void* stream);
/* Write `bytes` of a scratch buffer (L2 flush between timed iterations; bench only). */
int lgk_l2_flush(void* scratch, int64_t bytes, void* stream);
/* number of kernel launches issued through this library since load (bench's gpu_launches). */
int64_t lgk_launch_count(void);
/* sizeof() of the parameter structs as compiled (0 Torque, 1 LstmWeights, 2 Step, 3 Policy, 4 Game): lets a foreign-language
 * binding (ctypes / cgo / JNI) verify its mirror of the layout before the first call. */
int lgk_struct_size(int which);

#ifdef __cplusplus
}
#endif
#endif /* LGK_H */
