"""ctypes binding of liblgk.so (the C ABI in include/lgk.h).  No torch types cross this boundary: callers pass
``tensor.data_ptr()`` integers and the raw ``cudaStream_t``.  There is NO fallback: if the shared library is
missing or its struct layout differs from this mirror, importing this module raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
ABI_VERSION = 7          # include/lgk.h LGK_ABI_VERSION
LIB_PATH = os.environ.get("LGK_LIB_PATH") or os.path.join(_HERE, "liblgk.so")      # override: A/B of kernel builds

NUM_DOF, MAX_FEET, MAX_PEN, MAX_TERM, MAX_BODIES = 12, 4, 16, 8, 32
PHASE_PRE, PHASE_POST, PHASE_POST_REWARD, PHASE_POST_OBS = 1, 2, 4, 8
CTRL = {"P": 0, "V": 1, "T": 2}
STREAM_CMD, STREAM_PUSH, STREAM_RESET_DOF, STREAM_RESET_ROOT, STREAM_RESET_CMD, STREAM_TERRAIN, STREAM_OBS, STREAM_ACT = range(8)

# alphabetical = summation order of the reference (helpers.py:41-56, legged_robot.py:583-607)
REWARD_TERMS = ["action_rate", "ang_vel_xy", "base_height", "collision", "dof_acc", "dof_pos_limits", "dof_vel",
                "dof_vel_limits", "feet_air_time", "feet_contact_forces", "lin_vel_z", "no_fly", "orientation",
                "stand_still", "stumble", "termination", "torque_limits", "torques", "tracking_ang_vel",
                "tracking_lin_vel"]
R_COUNT = len(REWARD_TERMS)

f32, i32, i64, u64, vp = C.c_float, C.c_int32, C.c_int64, C.c_uint64, C.c_void_p


class TorqueParams(C.Structure):
    _fields_ = [("num_envs", i32), ("control_type", i32), ("use_lstm", i32), ("action_scale", f32),
                ("clip_actions", f32), ("lstm_variant", i32), ("sim_dt", f32),
                ("p_gains", f32 * NUM_DOF), ("d_gains", f32 * NUM_DOF), ("torque_limits", f32 * NUM_DOF),
                ("default_dof_pos", f32 * NUM_DOF),
                ("actions_in", vp), ("actions_clipped", vp), ("dof_state", vp), ("last_dof_vel", vp),
                ("torques", vp), ("sea_hidden_state", vp), ("sea_cell_state", vp), ("torques_mirror", vp),
                ("host_io", i32), ("pad_", i32)]


class LstmWeights(C.Structure):
    _fields_ = [("w_ih0", f32 * 64), ("w_hh0", f32 * 256), ("b_ih0", f32 * 32), ("b_hh0", f32 * 32),
                ("w_ih1", f32 * 256), ("w_hh1", f32 * 256), ("b_ih1", f32 * 32), ("b_hh1", f32 * 32),
                ("lin_w", f32 * 8), ("lin_b", f32 * 1), ("in_scale", f32 * 2), ("out_scale", f32 * 1)]


class StepParams(C.Structure):
    _fields_ = [
        ("num_envs", i32), ("num_bodies", i32), ("num_obs", i32), ("num_height_points", i32),
        ("actors_per_env", i32), ("root_actor_offset", i32), ("phase_mask", i32), ("host_state", i32),
        ("push_interval", i32), ("step", i32),
        ("seed", u64), ("env_id_offset", i64),
        ("heading_command", i32), ("measure_heights", i32), ("terrain_is_plane", i32), ("do_push", i32),
        ("add_noise", i32), ("only_positive_rewards", i32), ("terrain_curriculum", i32), ("custom_origins", i32),
        ("send_timeouts", i32), ("zero_lstm_on_reset", i32), ("predator_spawn", i32), ("predator_actor_offset", i32),
        ("dt", f32), ("resample_period", i32), ("max_episode_length", f32), ("max_episode_length_s", f32),
        ("max_push_vel", f32), ("cmd_lo", f32 * 4), ("cmd_range", f32 * 4),
        ("obs_scale_lin_vel", f32), ("obs_scale_ang_vel", f32), ("obs_scale_dof_pos", f32),
        ("obs_scale_dof_vel", f32), ("obs_scale_height", f32), ("clip_obs", f32),
        ("tracking_sigma", f32), ("base_height_target", f32), ("max_contact_force", f32),
        ("soft_dof_vel_limit", f32), ("soft_torque_limit", f32),
        ("border_size", f32), ("horizontal_scale", f32), ("vertical_scale", f32), ("horizontal_scale_recip", f32),
        ("hf_rows", i32), ("hf_cols", i32), ("half_env_length", f32), ("max_terrain_level", i32),
        ("terrain_num_cols", i32),
        ("default_dof_pos", f32 * NUM_DOF), ("dof_pos_lo", f32 * NUM_DOF), ("dof_pos_hi", f32 * NUM_DOF),
        ("dof_vel_limits", f32 * NUM_DOF), ("torque_limits", f32 * NUM_DOF), ("base_init_state", f32 * 13),
        ("num_feet", i32), ("num_pen", i32), ("num_term", i32),
        ("feet_idx", i32 * MAX_FEET), ("pen_idx", i32 * MAX_PEN), ("term_idx", i32 * MAX_TERM),
        ("reward_scale", f32 * R_COUNT), ("reward_active", i32 * R_COUNT), ("reward_slot", i32 * R_COUNT),
        ("num_reward_slots", i32),
        ("root_states", vp), ("dof_state", vp), ("contact_forces", vp),
        ("actions", vp), ("torques", vp), ("commands", vp), ("episode_length_buf", vp), ("last_actions", vp),
        ("last_dof_vel", vp), ("last_root_vel", vp), ("feet_air_time", vp), ("last_contacts", vp),
        ("episode_sums", vp), ("terrain_levels", vp), ("terrain_types", vp), ("terrain_origins", vp),
        ("env_origins", vp), ("sea_hidden_state", vp), ("sea_cell_state", vp),
        ("base_lin_vel", vp), ("base_ang_vel", vp), ("projected_gravity", vp), ("measured_heights", vp),
        ("obs_buf", vp), ("rew_buf", vp), ("reset_buf", vp), ("time_out_buf", vp),
        ("height_min3", vp), ("height_points_xy", vp), ("noise_scale_vec", vp), ("reset_stats", vp),
        ("scan_frames", vp), ("step_counter_dev", vp), ("base_quat", vp), ("obs_head", vp)]


GAME_TERMS = ["evasion", "pursuit", "termination"]        # LGK_G_* (alphabetical, like class_to_dict)


class GameAgent(C.Structure):
    _fields_ = [("active", i32 * 3), ("scale", f32 * 3), ("slot", i32 * 3), ("only_positive", i32), ("sums", vp), ("rew", vp)]


class GameParams(C.Structure):
    _fields_ = [
        ("num_envs", i32), ("variant", i32), ("decimation", i32), ("custom_origins", i32),
        ("step", i32), ("has_env_radius", i32), ("reset_dofs", i32), ("reset_only", i32),
        ("seed", u64), ("env_id_offset", i64),
        ("sim_dt", f32), ("capture_dist", f32), ("env_radius", f32), ("max_episode_length", f32), ("half_fov", f32),
        ("max_rel_pos", f32), ("ll_rew_weight", f32), ("pad0", f32),
        ("base_init_state", f32 * 13), ("default_dof_pos", f32 * 12), ("pad1", f32 * 3),
        ("prey", GameAgent), ("pred", GameAgent),
        ("root_states", vp), ("dof_state", vp), ("env_origins", vp), ("base_quat", vp),
        ("command_pred", vp), ("command_pred_stride", i64), ("ll_rews", vp), ("ll_dones", vp),
        ("predator_pos", vp), ("prey_states", vp), ("obs_prey", vp), ("obs_prey_stride", i64),
        ("obs_pred", vp), ("obs_pred_stride", i64), ("reset_buf", vp), ("time_out_buf", vp),
        ("episode_length_buf", vp), ("curr_episode_step", vp), ("reset_stats", vp),
        ("num_prey_slots", i32), ("num_pred_slots", i32), ("scratch", vp)]


class PolicyParams(C.Structure):
    _fields_ = [("num_envs", i32), ("num_obs", i32), ("num_critic_obs", i32), ("num_actions", i32),
                ("hidden", i32 * 3), ("obs", vp), ("critic_obs", vp),
                ("actor_w", vp * 4), ("actor_b", vp * 4), ("critic_w", vp * 4), ("critic_b", vp * 4),
                ("std", vp), ("seed", u64), ("step", i32), ("env_id_offset", i64), ("sample", i32),
                ("actions", vp), ("action_mean", vp), ("action_sigma", vp), ("values", vp),
                ("actions_log_prob", vp), ("workspace", vp), ("workspace_bytes", i64), ("weights_version", i64),
                ("nets", i32), ("pad_", i32)]


class LgkError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -m legged_games_gym_b200.csrc.build` "
            "(or __graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.lgk_last_error_string.restype = C.c_char_p
    lib.lgk_launch_count.restype = i64
    lib.lgk_policy_workspace_bytes.restype = i64
    lib.lgk_set_lstm_weights.argtypes = [C.POINTER(LstmWeights), vp]
    lib.lgk_compute_torques.argtypes = [C.POINTER(TorqueParams), vp]
    lib.lgk_post_physics.argtypes = [C.POINTER(StepParams), vp]
    lib.lgk_reset_idx.argtypes = [C.POINTER(StepParams), vp, i32, vp]
    lib.lgk_finalize_step.argtypes = [C.POINTER(StepParams), vp, vp, vp, vp, i32, vp]
    lib.lgk_post_physics_finalize.argtypes = [C.POINTER(StepParams), vp, vp, vp, vp, vp]
    lib.lgk_height_min3.argtypes = [vp, vp, i32, i32, vp]
    lib.lgk_height_scan.argtypes = [vp, i32, i32, i32, vp, i32, vp, i32, i32, f32, f32, f32, vp, vp, vp, vp]
    lib.lgk_rng_dump.argtypes = [u64, i32, i64, i32, i32, i32, i32, vp, vp]
    lib.lgk_policy_workspace_bytes.argtypes = [C.POINTER(PolicyParams)]
    lib.lgk_policy_act.argtypes = [C.POINTER(PolicyParams), vp]
    lib.lgk_policy_set_variant.argtypes = [C.c_int]
    lib.lgk_policy_debug_timeline.argtypes = [vp, C.c_int]
    lib.lgk_gae.argtypes = [vp, vp, vp, vp, i32, i32, f32, f32, vp, vp, vp, vp]
    lib.lgk_episode_stats.argtypes = [vp, vp, vp, vp, vp, i32, vp]
    lib.lgk_l2_flush.argtypes = [vp, i64, vp]
    lib.lgk_copy_from_pinned.argtypes = [vp, vp, i64, vp]
    lib.lgk_copy_to_pinned.argtypes = [vp, vp, i64, vp]
    lib.lgk_copy_rows_to_pinned.argtypes = [vp, vp, i32, vp, vp, i32, i32, i32, vp]
    lib.lgk_struct_size.argtypes = [C.c_int]
    lib.lgk_set_pdl.argtypes = [C.c_int]
    lib.lgk_game_step.argtypes = [C.POINTER(GameParams), vp]
    lib.lgk_game_prepare.argtypes = [vp, i64, vp, i64, vp, i32, C.POINTER(f32 * 8), i32, vp]
    lib.lgk_step_debug_timeline.argtypes = [vp]
    for which, cls in enumerate((TorqueParams, LstmWeights, StepParams, PolicyParams, GameParams)):
        n = lib.lgk_struct_size(which)
        if n != C.sizeof(cls):
            raise ImportError(f"liblgk.so struct {cls.__name__} is {n} bytes, ctypes mirror is {C.sizeof(cls)}: rebuild")
    if lib.lgk_abi_version() != ABI_VERSION:
        raise ImportError("liblgk.so ABI version mismatch")
    return lib


lib = _load()

EXPORTS = ["lgk_set_lstm_weights", "lgk_compute_torques", "lgk_post_physics", "lgk_reset_idx", "lgk_finalize_step", "lgk_post_physics_finalize",
           "lgk_height_min3", "lgk_height_scan", "lgk_rng_dump", "lgk_policy_workspace_bytes", "lgk_policy_act", "lgk_policy_set_variant", "lgk_policy_debug_timeline", "lgk_set_pdl", "lgk_game_step", "lgk_game_prepare", "lgk_step_debug_timeline",
           "lgk_gae", "lgk_episode_stats", "lgk_last_error_string", "lgk_abi_version", "lgk_l2_flush", "lgk_copy_from_pinned", "lgk_copy_to_pinned", "lgk_copy_rows_to_pinned", "lgk_launch_count",
           "lgk_struct_size"]


def check(rc, what=""):
    if rc != 0:
        msg = lib.lgk_last_error_string()
        raise LgkError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def launch_count():
    return int(lib.lgk_launch_count())
