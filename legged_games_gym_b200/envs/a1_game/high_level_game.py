"""HighLevelGame -- the reference's hierarchical predator / prey task (legged_gym/envs/a1_game/high_level_game.py:26-586).

A high-level policy commands the prey robot (4 velocity commands, executed by a frozen low-level locomotion policy on a
LowLevelGame env) and the predator sphere (2 velocity commands, single integrator).  Everything the reference does in
torch after ``ll_env.step`` -- predator integration, rewards, capture / radius / low-level dones, the reset of root states
and of the observation history, field-of-view sensing -- is ONE kernel here (``lgk_game_step``, csrc/lgk_game.cu).

Differences from the reference constructor, all forced by what is absent from its tree:
the low-level policy can be handed in (``ll_policy=`` any callable obs -> actions); without it the checkpoint
``logs/<experiment>/sideways_walking_policy`` is loaded like HLG:95-103 and a missing file raises.  The forked rsl_rl
``LLPolicyRunner`` is replaced by building the ActorCritic of the low-level train cfg directly."""
import copy
import ctypes as C
import os

import numpy as np
import torch

from ... import LEGGED_GYM_ROOT_DIR
from ... import _native as nat
from ...utils.helpers import class_to_dict, get_load_path
from .low_level_game import LowLevelGame


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


class _GameBase:
    VARIANT = 0
    MAX_REL_POS = 100.
    LL_REW_WEIGHT = 2.0                    # HLG:364
    HALF_FOV = 1.20428 / 2.                # HLG:427

    def __init__(self, cfg, sim_params, physics_engine, sim_device, headless, ll_policy=None, ll_env_cfg=None,
                 **ll_env_kwargs):
        print(f"[{type(self).__name__}] initializing ...")
        self.cfg = cfg
        self.sim_params = sim_params
        self.height_samples = None
        self.debug_viz = False
        self.init_done = False
        self.physics_engine = physics_engine
        self.sim_device = sim_device
        self.headless = headless
        self.capture_dist = self.cfg.env.capture_dist
        # ---- low-level env: the A1 cfg with the game's representation (HLG:69-90)
        from ...utils.task_registry import task_registry
        a1_env_cfg, ll_train_cfg = task_registry.get_cfgs(name="a1")
        ll_env_cfg, ll_train_cfg = copy.deepcopy(ll_env_cfg if ll_env_cfg is not None else a1_env_cfg), copy.deepcopy(ll_train_cfg)
        ll_env_cfg.env.num_envs = self.cfg.env.num_envs
        ll_env_cfg.terrain.num_rows = self.cfg.terrain.num_rows
        ll_env_cfg.terrain.num_cols = self.cfg.terrain.num_cols
        ll_env_cfg.terrain.curriculum = self.cfg.terrain.curriculum
        ll_env_cfg.noise.add_noise = self.cfg.noise.add_noise
        ll_env_cfg.domain_rand.randomize_friction = self.cfg.domain_rand.randomize_friction
        ll_env_cfg.domain_rand.push_robots = self.cfg.domain_rand.push_robots
        ll_env_cfg.terrain.mesh_type = self.cfg.terrain.mesh_type
        ll_env_cfg.rewards.scales.torques = -5.                      # HLG:85
        self.ll_env = LowLevelGame(cfg=ll_env_cfg, sim_params=sim_params, physics_engine=physics_engine,
                                   sim_device=sim_device, headless=headless, **ll_env_kwargs)
        self.device = self.ll_env.device
        self.gym = self.ll_env.gym
        self.ll_policy = ll_policy if ll_policy is not None else self._load_ll_policy(ll_env_cfg, ll_train_cfg)
        self._parse_cfg(self.cfg)
        self.num_envs = cfg.env.num_envs
        N, dev = self.num_envs, self.device
        self.reset_buf = torch.ones(N, device=dev, dtype=torch.bool)      # reference: long ones, bool from the first step on
        self.episode_length_buf = torch.zeros(N, device=dev, dtype=torch.long)
        self.time_out_buf = torch.zeros(N, device=dev, dtype=torch.bool)
        self.curr_episode_step = torch.zeros(N, device=dev, dtype=torch.long)
        self.extras = {}
        self.enable_viewer_sync = True
        self.viewer = None
        self._alloc_agent_buffers()
        self._init_buffers()
        self._prepare_rewards()
        self._build_params()
        self.common_step_counter = 0
        self.init_done = True

    # ------------------------------------------------------------------ HLG:95-103
    def _load_ll_policy(self, ll_env_cfg, ll_train_cfg):
        from ...rsl_rl.modules import ActorCritic
        log_root = os.path.join(LEGGED_GYM_ROOT_DIR, "logs", ll_train_cfg.runner.experiment_name)
        path = get_load_path(log_root, load_run="sideways_walking_policy", checkpoint=ll_train_cfg.runner.checkpoint)
        pol = class_to_dict(ll_train_cfg.policy)
        n_obs = ll_env_cfg.env.num_observations
        n_critic = ll_env_cfg.env.num_privileged_obs or n_obs
        ac = ActorCritic(n_obs, n_critic, ll_env_cfg.env.num_actions, **pol).to(self.device)
        ac.load_state_dict(torch.load(path, map_location=self.device)["model_state_dict"])
        ac.eval()
        return ac.act_inference

    # ------------------------------------------------------------------ HLG:562-571
    def _parse_cfg(self, cfg):
        self.command_ranges = class_to_dict(self.cfg.commands.ranges)
        if self.cfg.terrain.mesh_type not in ["heightfield", "trimesh"]:
            self.cfg.terrain.curriculum = False
        self.max_episode_length_s = self.cfg.env.episode_length_s
        self.max_episode_length = np.ceil(self.max_episode_length_s / self.ll_env.dt)

    # ------------------------------------------------------------------ HLG:519-535
    def _init_buffers(self):
        N, dev = self.num_envs, self.device
        ll = self.ll_env
        self.prey_states = ll.root_states[ll.prey_indices, :].contiguous()
        self.init_predator_pos = ll.init_predator_pos.clone()
        self.predator_pos = self.init_predator_pos.contiguous()
        self._stats = torch.zeros(16, device=dev)
        self._scratch = torch.zeros(8, device=dev, dtype=torch.int32)

    @property
    def base_quat(self):
        return self.prey_states[:, 3:7]

    @property
    def base_lin_vel(self):
        from ...utils.math import quat_rotate_inverse
        return quat_rotate_inverse(self.base_quat, self.prey_states[:, 7:10])

    @property
    def base_ang_vel(self):
        from ...utils.math import quat_rotate_inverse
        return quat_rotate_inverse(self.base_quat, self.prey_states[:, 10:13])

    def _agent(self, ag, scales_cfg, only_positive, rew):
        """HLG:537-560: drop zero scales, multiply by the low-level dt; one [K,N] buffer backs the dict of sums."""
        scales = class_to_dict(scales_cfg)
        for k in list(scales.keys()):
            if scales[k] == 0:
                scales.pop(k)
            else:
                scales[k] *= self.ll_env.dt
        for k in scales:
            if k not in nat.GAME_TERMS:
                raise AttributeError(f"'{type(self).__name__}' object has no attribute '_reward_{k}'")
        buf = torch.zeros(max(1, len(scales)), self.num_envs, device=self.device)
        sums = {}
        for slot, k in enumerate(scales):
            i = nat.GAME_TERMS.index(k)
            ag.active[i], ag.scale[i], ag.slot[i] = 1, float(scales[k]), slot
            sums[k] = buf[slot]
        ag.only_positive = int(bool(only_positive))
        ag.sums = buf.data_ptr()
        ag.rew = rew.data_ptr()
        return scales, [k for k in scales if k != "termination"], sums, buf

    def _build_params(self):
        ll, cfg = self.ll_env, self.cfg
        p = self._params
        p.num_envs, p.variant = self.num_envs, self.VARIANT
        p.decimation = int(ll.cfg.control.decimation)
        p.custom_origins = int(ll.custom_origins)
        radius = getattr(cfg.env, "env_radius", None)
        p.has_env_radius, p.env_radius = int(radius is not None), float(radius or 0.)
        p.reset_dofs = int(self.VARIANT == 1)
        p.seed = int(getattr(ll.cfg, "seed", 0) or 0) & 0xFFFFFFFFFFFFFFFF
        p.env_id_offset = int(getattr(ll, "env_id_offset", 0))
        p.sim_dt, p.capture_dist = float(ll.cfg.sim.dt), float(self.capture_dist)
        p.max_episode_length = float(self.max_episode_length)
        p.half_fov, p.max_rel_pos, p.ll_rew_weight = self.HALF_FOV, self.MAX_REL_POS, self.LL_REW_WEIGHT
        p.base_init_state[:] = [float(v) for v in ll.base_init_state.cpu()]
        p.default_dof_pos[:] = [float(v) for v in ll.default_dof_pos.flatten().cpu()]
        p.root_states, p.dof_state = ll.root_states.data_ptr(), ll.dof_state.data_ptr()
        p.env_origins, p.base_quat = ll.env_origins.data_ptr(), ll.base_quat.data_ptr()
        p.ll_rews, p.ll_dones = ll.rew_buf.data_ptr(), ll._reset_bool.data_ptr()
        p.predator_pos, p.prey_states = self.predator_pos.data_ptr(), self.prey_states.data_ptr()
        p.reset_buf, p.time_out_buf = self.reset_buf.data_ptr(), self.time_out_buf.data_ptr()
        p.episode_length_buf, p.curr_episode_step = self.episode_length_buf.data_ptr(), self.curr_episode_step.data_ptr()
        p.reset_stats = self._stats.data_ptr()
        p.scratch = self._scratch.data_ptr()
        r = self.command_ranges
        self._ranges8 = (C.c_float * 8)(*[float(v) for k in ("lin_vel_x", "lin_vel_y", "predator_lin_vel_x", "predator_lin_vel_y")
                                          for v in r[k]])

    def _launch(self, command_pred, reset_only=False, dones=None):
        p = self._params
        assert command_pred.is_cuda and command_pred.dtype == torch.float32 and command_pred.stride(1) == 1
        p.command_pred, p.command_pred_stride = command_pred.data_ptr(), command_pred.stride(0)
        p.step = self.common_step_counter
        p.reset_only = int(reset_only)
        p.decimation = 0 if reset_only else int(self.ll_env.cfg.control.decimation)
        p.ll_dones = (dones if dones is not None else self.ll_env._reset_bool).data_ptr()
        p.ll_rews = self.ll_env.rew_buf.data_ptr()
        p.env_id_offset = int(getattr(self.ll_env, "env_id_offset", 0))
        nat.check(nat.lib.lgk_game_step(C.byref(p), _stream_ptr()), "lgk_game_step")

    def _ll_step(self, command_prey, command_pred):
        """HLG:161-181: clip / wrap the commands in place and hand the prey's to the low-level env (one kernel), then the
        frozen low-level policy and the low-level step.  The reference aliases ``ll_env.commands`` to the command tensor
        (HLG:174); the kernels own a persistent buffer, so the values are copied."""
        ll = self.ll_env
        for t in (command_prey, command_pred):
            if not (t.is_cuda and t.dtype == torch.float32 and t.stride(1) == 1):
                raise TypeError("high-level commands must be fp32 CUDA tensors with unit column stride")
        nat.check(nat.lib.lgk_game_prepare(command_prey.data_ptr(), command_prey.stride(0), command_pred.data_ptr(),
                                           command_pred.stride(0), ll.commands.data_ptr(), self.num_envs,
                                           C.byref(self._ranges8), int(bool(self.cfg.commands.heading_command)),
                                           _stream_ptr()), "lgk_game_prepare")
        ll_obs = ll.get_observations()
        actions = self.ll_policy(ll_obs.detach())
        return ll.step(actions.detach())

    def _push_game_state_to_sim(self):
        """HLG:285-286 and the indexed pushes of LLG:441-451 / 395-399, as full-tensor updates (no id list on the host)"""
        ll = self.ll_env
        ll.gym.set_actor_root_state_tensor(ll.root_states)
        if self.VARIANT == 1 and hasattr(ll.gym, "set_dof_state_tensor"):
            ll.gym.set_dof_state_tensor(ll.dof_state)

    def reset_idx(self, env_ids):
        """HLG:326-349 on an explicit id list"""
        if len(env_ids) == 0:
            return
        mask = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        mask[env_ids] = True
        self._stats.zero_()
        self._launch(torch.zeros(self.num_envs, 2, device=self.device), reset_only=True, dones=mask)
        self._after_reset_stats()
        self._push_game_state_to_sim()

    def _after_reset_stats(self):
        pass

    def get_privileged_observations(self):
        return None

    def render(self, sync_frame_time=True):
        return None


class HighLevelGame(_GameBase):
    VARIANT = 0

    def _alloc_agent_buffers(self):
        cfg, N, dev = self.cfg, self.num_envs, self.device
        self.num_obs = cfg.env.num_observations
        self.num_privileged_obs = cfg.env.num_privileged_obs
        self.num_actions = cfg.env.num_actions
        self.obs_buf = self.MAX_REL_POS * torch.ones(N, self.num_obs, device=dev, dtype=torch.float)
        self.rew_buf = torch.zeros(N, device=dev, dtype=torch.float)
        self.privileged_obs_buf = None
        self._params = nat.GameParams()

    def _prepare_rewards(self):
        p = self._params
        self.reward_scales, self.reward_names, self.episode_sums, self._sums_buf = self._agent(
            p.prey, self.cfg.rewards.scales, self.cfg.rewards.only_positive_rewards, self.rew_buf)
        p.num_prey_slots, p.num_pred_slots = len(self.reward_scales), 0
        p.obs_prey, p.obs_prey_stride = self.obs_buf.data_ptr(), self.obs_buf.stride(0)
        p.obs_pred, p.obs_pred_stride = self.obs_buf[:, 16:].data_ptr(), self.obs_buf.stride(0)

    # ------------------------------------------------------------------ HLG:146-241
    def step(self, command):
        self._ll_step(command[:, 0:4], command[:, 4:6])
        self.common_step_counter += 1
        self._launch(command[:, 4:6])
        self._push_game_state_to_sim()
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def reset(self):
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        obs, privileged_obs, _, _, _ = self.step(torch.zeros(self.num_envs, self.num_actions, device=self.device, requires_grad=False))
        return obs, privileged_obs

    def get_observations(self):
        # the reference recomputes the observation (and shifts its history once more) on every call (HLG:411-413); the
        # runner calls it once before the first step, when the buffer still holds its initial fill
        return self.obs_buf
