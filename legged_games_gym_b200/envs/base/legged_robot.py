"""LeggedRobot -- host-side mirror of the reference env class (legged_gym/envs/base/legged_robot.py, "LR").

Same constructor, attributes, tensor layouts and method names; every per-env computation of LR:80-230,
329-508 and 831-969 is executed by liblgk.so (include/lgk.h) through ``_native``:

    step()                -> 4 x lgk_compute_torques  +  lgk_post_physics_finalize (= lgk_post_physics(PRE|POST) + lgk_finalize_step)
    reset_idx(env_ids)    -> lgk_reset_idx + lgk_finalize_step
    _get_heights()        -> lgk_height_scan

Python keeps only what is host logic in the reference too: cfg parsing, buffer allocation, the push /
curriculum step counters and the extension points.  Supported extension points for subclasses (README.md:29-32,
56-66 of the reference): extra ``_reward_<name>`` terms written in torch (the step then runs PRE, the Python
terms, POST), ``_compute_torques``, ``compute_observations`` (called after the native step when overridden),
``_get_noise_scale_vec``, ``_init_buffers`` and ``reset_idx`` (the step is then split around the Python method, which runs
where LR:128-129 calls it).  Overriding ``check_termination`` / ``_post_physics_step_callback`` / ``compute_reward`` is
rejected at construction: those run inside the step kernels.
"""
import ctypes as C

import numpy as np
import torch

from ... import _native as nat
from ...sim.asset_model import model_for_asset
from ...sim.state_feeder import StateFeeder, synth_height_field, synth_terrain_origins
from ...utils.helpers import class_to_dict
from .base_task import BaseTask
from ...utils.terrain import Terrain
from .legged_robot_config import LeggedRobotCfg


class SyntheticTerrain:
    """Stand-in with the attributes the hot path reads from a Terrain object (utils/terrain.py: cfg, env_length,
    env_origins [rows, cols, 3], heightsamples int16 [tot_rows, tot_cols]) filled with the seeded synthetic field of
    SURVEY.md section 8(d) (uniform int16 heights): the benchmark / parity input.  Pass it as ``terrain=``; the default
    is the generated terrain of utils/terrain.py, like the reference (LR:235)."""

    def __init__(self, cfg, seed):
        self.cfg = cfg
        self.env_length = cfg.terrain_length
        self.env_width = cfg.terrain_width
        self.border = int(cfg.border_size / cfg.horizontal_scale)
        self.tot_rows = int(cfg.num_rows * cfg.terrain_length / cfg.horizontal_scale) + 2 * self.border
        self.tot_cols = int(cfg.num_cols * cfg.terrain_width / cfg.horizontal_scale) + 2 * self.border
        self.heightsamples = synth_height_field(self.tot_rows, self.tot_cols, seed)
        self.env_origins = synth_terrain_origins(cfg)


_Terrain = SyntheticTerrain


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


class _Extras(dict):
    """extras dict whose "episode" entry is produced on read (see LeggedRobot._finalize)."""

    def arm_episode(self, producer):
        self._producer = producer
        dict.__setitem__(self, "episode", None)

    def __getitem__(self, key):
        if key == "episode" and getattr(self, "_producer", None) is not None and dict.__getitem__(self, key) is None:
            return self._producer()
        return dict.__getitem__(self, key)

    def get(self, key, default=None):
        return self[key] if key in self else default


class LeggedRobot(BaseTask):
    _FUSED_HOOKS = ("check_termination", "_post_physics_step_callback", "compute_reward", "_resample_commands",
                    "_push_robots", "_reset_dofs", "_reset_root_states", "_update_terrain_curriculum")

    def __init__(self, cfg: LeggedRobotCfg, sim_params, physics_engine, sim_device, headless, sim_backend=None,
                 terrain=None, init_terrain_levels=None):
        self.cfg = cfg
        self.sim_params = sim_params
        self.height_samples = None
        self.debug_viz = False
        self.init_done = False
        self._terrain_arg = terrain
        self._init_levels_arg = init_terrain_levels
        for name in self._FUSED_HOOKS:
            if getattr(type(self), name) is not getattr(LeggedRobot, name):
                raise NotImplementedError(
                    f"{type(self).__name__}.{name} overrides a stage that runs inside the fused CUDA step; "
                    "extend through _reward_<name>, compute_observations, _compute_torques or _init_buffers instead")
        self._parse_cfg(self.cfg)
        super().__init__(self.cfg, sim_params, physics_engine, sim_device, headless, sim_backend)
        self._init_buffers()
        self._prepare_reward_function()
        self._build_native_params()
        self.init_done = True

    # ------------------------------------------------------------------ LR:80-104
    def step(self, actions):
        if self._graph_ok and not self._needs_host_logic_next_step():
            return self._step_graphed(actions)
        return self._step_eager(actions)

    def set_env_id_offset(self, offset):
        """Global id of this process's env 0 (multi-GPU sharding: rank r passes r * num_envs): keeps the counter-based
        RNG streams of the shards disjoint and identical to a single-process job over all envs."""
        self.env_id_offset = int(offset)
        self._params.env_id_offset = int(offset)
        self._graph = None                      # parameters are baked into the captured graph

    @property
    def action_buffer(self):
        """[num_envs, num_actions] fp32 device buffer the captured step graph reads its actions from.  A policy that
        writes its actions here (ActorCritic's fused kernel takes an output pointer) and calls ``step(env.action_buffer)``
        saves the per-step device-to-device copy."""
        return self._actions_in

    def _needs_host_logic_next_step(self):
        return bool(self.cfg.commands.curriculum) and ((self.common_step_counter + 1) % self.max_episode_length == 0)

    def _step_graphed(self, actions):
        """The whole step (4 torque launches, fused post-physics, finalize) replayed as ONE CUDA graph: every
        per-step parameter (step counter, push flag) lives in device memory, so nothing is patched between replays."""
        if actions.data_ptr() != self._actions_in.data_ptr():     # callers that write into `action_buffer` skip this copy
            self._actions_in.copy_(actions, non_blocking=True)
        if self._graph is None:
            if self._eager_steps < 1:              # first step eager: sets kernel attributes, warms the allocator
                self._eager_steps += 1
                return self._step_eager(self._actions_in)
            g = torch.cuda.CUDAGraph()
            counter = self.common_step_counter
            l0 = nat.launch_count()
            with torch.cuda.graph(g):
                self._step_eager(self._actions_in)
            self._graph_launches = nat.launch_count() - l0     # kernel nodes of the graph (the library counts its launches)
            self.common_step_counter = counter     # capture ran the Python side once without executing kernels
            self._graph = g
        self._graph.replay()
        self.common_step_counter += 1
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def _step_eager(self, actions):
        clip_actions = self.cfg.normalization.clip_actions
        gym = self.gym
        native_tq = self._native_torques
        if native_tq:
            tp = self._tq_params
            tp.actions_in = actions.data_ptr()
            if not (actions.is_cuda and actions.dtype == torch.float32 and actions.is_contiguous()):
                actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
                tp.actions_in = actions.data_ptr()
            tp.actions_clipped = self.actions.data_ptr()
        else:
            torch.clamp(actions.to(self.device), -clip_actions, clip_actions, out=self.actions)
        self.render()
        decimation = self.cfg.control.decimation
        for k in range(decimation):
            zero_copy = False
            if native_tq:
                # a host-resident sim may hand the kernel its pinned buffers for this sub-step (SimBackend.dof_state_source /
                # actuation_force_sink): the launch then pulls the dof state and pushes the torques over PCIe itself
                src, sink = gym.dof_state_source(), gym.actuation_force_sink()
                zero_copy = src is not None
                tp.dof_state = (src if zero_copy else self.dof_state).data_ptr()
                tp.torques_mirror = sink.data_ptr() if sink is not None else None
                tp.host_io = int(zero_copy or sink is not None or bool(getattr(gym, "state_in_host_memory", False)))
                nat.check(nat.lib.lgk_compute_torques(C.byref(tp), _stream_ptr()), "lgk_compute_torques")
                if k == 0:      # later sub-steps read the clipped copy (clip is idempotent)
                    tp.actions_in = self.actions.data_ptr()
                    tp.actions_clipped = None
            else:
                self.torques[:] = self._compute_torques(self.actions).view(self.torques.shape)
            gym.set_dof_actuation_force_tensor(self.torques)
            gym.simulate()
            if not zero_copy or k == decimation - 1:
                gym.refresh_dof_state_tensor()       # zero-copy source: only post-physics needs the device copy
        self.post_physics_step()
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    # ------------------------------------------------------------------ LR:106-137
    def post_physics_step(self):
        gym = self.gym
        gym.refresh_actor_root_state_tensor()
        gym.refresh_net_contact_force_tensor()
        self.common_step_counter += 1          # episode_length_buf += 1 happens in the kernel (LR:114)
        p = self._params
        dr = self.cfg.domain_rand
        do_push = bool(dr.push_robots) and (self.common_step_counter % dr.push_interval == 0)   # device derives the same
        if self.reset_buf.dtype != torch.bool:     # first step: long ones -> persistent bool buffer
            self.reset_buf = self._reset_bool
        st = _stream_ptr()
        cmd_curr = self.cfg.commands.curriculum and (self.common_step_counter % self.max_episode_length == 0)
        if self._python_reward_names or cmd_curr or self._reset_overridden:
            p.phase_mask = nat.PHASE_PRE
            nat.check(nat.lib.lgk_post_physics(C.byref(p), st), "lgk_post_physics(PRE)")
            for name in self._python_reward_names:             # user terms, LR:199-203
                rew = getattr(self, "_reward_" + name)() * self.reward_scales[name]
                self.rew_buf += rew
                self.episode_sums[name] += rew
            if self._reset_overridden:
                # a subclass's reset_idx runs exactly where the reference calls it (LR:126-130): after compute_reward
                # (positive clip + termination term), before compute_observations.  The in-kernel reset is off for this
                # step; the override reaches the kernels through LeggedRobot.reset_idx -> lgk_reset_idx.
                p.phase_mask = nat.PHASE_POST_REWARD
                nat.check(nat.lib.lgk_post_physics(C.byref(p), st), "lgk_post_physics(POST_REWARD)")
                env_ids = self.reset_buf.nonzero(as_tuple=False).flatten()          # LR:128
                self._in_step_reset = True
                try:
                    self.reset_idx(env_ids)
                finally:
                    self._in_step_reset = False
                p.phase_mask = nat.PHASE_POST_OBS
                nat.check(nat.lib.lgk_post_physics(C.byref(p), st), "lgk_post_physics(POST_OBS)")
            else:
                if cmd_curr:
                    ids = self.reset_buf.nonzero(as_tuple=False).flatten()
                    if len(ids) > 0:
                        self.update_command_curriculum(ids)
                p.phase_mask = nat.PHASE_POST
                nat.check(nat.lib.lgk_post_physics(C.byref(p), st), "lgk_post_physics(POST)")
            self._finalize(st, advance=1)
        elif not getattr(self, "fuse_finalize", True):       # the two calls lgk_post_physics_finalize stands for (tests)
            p.phase_mask = nat.PHASE_PRE | nat.PHASE_POST
            nat.check(nat.lib.lgk_post_physics(C.byref(p), st), "lgk_post_physics")
            self._finalize(st, advance=1)
        else:
            # the whole step: K1, K2 and the finalize pass in one call (on rough terrain finalize rides in K2's grid)
            p.phase_mask = nat.PHASE_PRE | nat.PHASE_POST
            nat.check(nat.lib.lgk_post_physics_finalize(C.byref(p), self.reset_env_ids.data_ptr(), self.reset_count.data_ptr(),
                                                        self._episode_means.data_ptr(), self._time_outs_extras.data_ptr(), st),
                      "lgk_post_physics_finalize")
            self._arm_extras()
        if do_push:
            gym.set_actor_root_state_tensor(self.root_states)
        self._push_resets_to_sim()
        if self._obs_overridden:
            self.compute_observations()
            clip_obs = self.cfg.normalization.clip_observations
            self.obs_buf = torch.clip(self.obs_buf, -clip_obs, clip_obs)

    def _finalize(self, st, advance=0):
        nat.check(nat.lib.lgk_finalize_step(C.byref(self._params), self.reset_env_ids.data_ptr(),
                                            self.reset_count.data_ptr(), self._episode_means.data_ptr(),
                                            self._time_outs_extras.data_ptr(), advance, st), "lgk_finalize_step")
        self._arm_extras()

    def _arm_extras(self):
        # extras are refreshed only when something was reset (LR:157-158, 179-191): the kernel keeps the previous
        # values otherwise.  extras["episode"] is materialised (one clone of the means vector) when it is READ, so a
        # runner that keeps the dicts of several steps gets distinct tensors and a step that nobody logs costs nothing.
        self.extras.arm_episode(self._episode_snapshot)
        if self.cfg.env.send_timeouts:
            dict.__setitem__(self.extras, "time_outs", self._time_outs_extras)

    def _episode_snapshot(self):
        means = self._episode_means.clone()
        ep = {"rew_" + k: means[i] for i, k in enumerate(self._sum_names)}
        if self.cfg.terrain.curriculum:
            ep["terrain_level"] = means[len(self._sum_names)]
        if self.cfg.commands.curriculum:
            ep["max_command_x"] = self.command_ranges["lin_vel_x"][1]
        return ep

    # ------------------------------------------------------------------ stages that live in the kernel
    def check_termination(self):
        raise RuntimeError("check_termination runs inside lgk_post_physics")

    def compute_reward(self):
        raise RuntimeError("compute_reward runs inside lgk_post_physics")

    def _post_physics_step_callback(self):
        raise RuntimeError("_post_physics_step_callback runs inside lgk_post_physics")

    def _resample_commands(self, env_ids):
        raise RuntimeError("_resample_commands runs inside lgk_post_physics / lgk_reset_idx")

    def _push_robots(self):
        raise RuntimeError("_push_robots runs inside lgk_post_physics")

    def _reset_dofs(self, env_ids):
        raise RuntimeError("_reset_dofs runs inside lgk_post_physics / lgk_reset_idx")

    def _reset_root_states(self, env_ids):
        raise RuntimeError("_reset_root_states runs inside lgk_post_physics / lgk_reset_idx")

    def _update_terrain_curriculum(self, env_ids):
        raise RuntimeError("_update_terrain_curriculum runs inside lgk_post_physics / lgk_reset_idx")

    # ------------------------------------------------------------------ LR:147-191 on an explicit id list
    def reset_idx(self, env_ids):
        if len(env_ids) == 0:
            return
        env_ids = env_ids.to(device=self.device, dtype=torch.long).contiguous()
        if self.cfg.commands.curriculum and (self.common_step_counter % self.max_episode_length == 0):
            self.update_command_curriculum(env_ids)
        p = self._params
        if self.reset_buf.dtype != torch.bool:
            self.reset_buf = self._reset_bool
            self.reset_buf.fill_(True)
        st = _stream_ptr()
        saved = p.terrain_curriculum, p.step_counter_dev, p.step
        if not self.init_done:
            p.terrain_curriculum = 0            # "don't change on initial reset" (LR:453-455)
        in_step = getattr(self, "_in_step_reset", False)
        if in_step:
            # called from post_physics_step on behalf of a subclass override: the reset belongs to THIS step (same RNG
            # counter and extras slot as the in-kernel reset), and the step's own finalize / sim push follow
            p.step_counter_dev, p.step = None, int(self.common_step_counter)
        nat.check(nat.lib.lgk_reset_idx(C.byref(p), env_ids.data_ptr(), int(env_ids.numel()), st), "lgk_reset_idx")
        p.terrain_curriculum, p.step_counter_dev, p.step = saved
        if not in_step:
            self._finalize(st, advance=0)
            self._push_resets_to_sim()
    reset_idx._lgk_native = True

    # ------------------------------------------------------------------ LR:212-230 (torch version for overriders)
    def compute_observations(self):
        o = self.obs_scales
        self.obs_buf = torch.cat((self.base_lin_vel * o.lin_vel, self.base_ang_vel * o.ang_vel, self.projected_gravity,
                                  self.commands[:, :3] * self.commands_scale,
                                  (self.dof_pos - self.default_dof_pos) * o.dof_pos, self.dof_vel * o.dof_vel,
                                  self.actions), dim=-1)
        if self.cfg.terrain.measure_heights:
            heights = torch.clip(self.root_states[self._root_rows(), 2].unsqueeze(1) - 0.5 - self.measured_heights, -1, 1.) * o.height_measurements
            self.obs_buf = torch.cat((self.obs_buf, heights), dim=-1)
        if self.add_noise:
            self.obs_buf += (2 * torch.rand_like(self.obs_buf) - 1) * self.noise_scale_vec

    # ------------------------------------------------------------------ LR:232-251
    def create_sim(self):
        self.up_axis_idx = 2
        mesh_type = self.cfg.terrain.mesh_type
        if mesh_type in ("heightfield", "trimesh"):
            self.terrain = self._terrain_arg if self._terrain_arg is not None else Terrain(self.cfg.terrain, self.num_envs)
            # env instances built on the SAME terrain object share its device copies (height field + min3 table): one
            # terrain serves every env of a GPU, however many env objects exist
            cache = getattr(self.terrain, "_lgk_device_cache", None)
            if cache is None:
                cache = {}
                try:
                    self.terrain._lgk_device_cache = cache
                except AttributeError:
                    pass
            key = str(self.device)
            if key not in cache:
                cache[key] = {"height_samples": torch.from_numpy(np.ascontiguousarray(self.terrain.heightsamples)).view(
                    self.terrain.tot_rows, self.terrain.tot_cols).to(self.device)}
            self._terrain_cache = cache[key]
            self.height_samples = self._terrain_cache["height_samples"]
        elif mesh_type not in (None, "plane"):
            raise ValueError("Terrain mesh type not recognised. Allowed types are [None, plane, heightfield, trimesh]")
        self._create_envs()

    # ------------------------------------------------------------------ LR:657-750 (asset tables instead of PhysX)
    def _create_envs(self):
        model = model_for_asset(self.cfg.asset)
        self.robot_model = model
        self.num_dof = self.num_dofs = model.num_dof
        self.num_bodies = model.num_bodies                     # robot links (LR:687); the contact view may hold more
        self.dof_names = list(model.dof_names)
        if self.num_dof != nat.NUM_DOF:
            raise ValueError(f"kernels are built for {nat.NUM_DOF} DOF, asset has {self.num_dof}")
        dev = self.device
        a = self.cfg.asset
        self.feet_indices = torch.tensor(model.indices(a.foot_name), dtype=torch.long, device=dev)
        self.penalised_contact_indices = torch.tensor(model.indices(list(a.penalize_contacts_on)), dtype=torch.long, device=dev)
        self.termination_contact_indices = torch.tensor(model.indices(list(a.terminate_after_contacts_on)), dtype=torch.long, device=dev)
        # _process_dof_props LR:299-313 (same fp32 op sequence)
        self.dof_pos_limits = torch.zeros(self.num_dof, 2, dtype=torch.float, device="cpu")
        self.dof_vel_limits = torch.zeros(self.num_dof, dtype=torch.float, device="cpu")
        self.torque_limits = torch.zeros(self.num_dof, dtype=torch.float, device="cpu")
        soft = self.cfg.rewards.soft_dof_pos_limit
        for i in range(self.num_dof):
            self.dof_pos_limits[i, 0] = float(np.float32(model.dof_lower[i]))
            self.dof_pos_limits[i, 1] = float(np.float32(model.dof_upper[i]))
            self.dof_vel_limits[i] = float(np.float32(model.dof_vel_limits[i]))
            self.torque_limits[i] = float(np.float32(model.torque_limits[i]))
            m = (self.dof_pos_limits[i, 0] + self.dof_pos_limits[i, 1]) / 2
            r = self.dof_pos_limits[i, 1] - self.dof_pos_limits[i, 0]
            self.dof_pos_limits[i, 0] = m - 0.5 * r * soft
            self.dof_pos_limits[i, 1] = m + 0.5 * r * soft
        self.dof_pos_limits = self.dof_pos_limits.to(dev)
        self.dof_vel_limits = self.dof_vel_limits.to(dev)
        self.torque_limits = self.torque_limits.to(dev)
        ini = self.cfg.init_state
        self.base_init_state = torch.tensor(ini.pos + ini.rot + ini.lin_vel + ini.ang_vel, dtype=torch.float, device=dev)
        self._get_env_origins()
        self._randomize_physical_props()
        if self.gym is None:
            self.gym = StateFeeder(self.num_envs, self.num_bodies + self._extra_bodies_per_env(), self.num_dof, device=dev,
                                   seed=getattr(self.cfg, "seed", 0) or 0, actors_per_env=self._actors_per_env())

    def _randomize_physical_props(self):
        """Init-time domain randomisation of the reference's creation callbacks: 64 friction buckets drawn once and
        assigned per env (LR:261-283 _process_rigid_shape_props) and a uniform added base mass per env (LR:316-327
        _process_rigid_body_props).  PhysX is what consumes them: they are kept as tensors and offered to the sim backend
        (``set_rigid_shape_friction`` / ``set_base_mass_offsets``); no per-step cost."""
        dr = self.cfg.domain_rand
        self.friction_coeffs = None
        self.added_base_mass = None
        if dr.randomize_friction:
            num_buckets = 64
            bucket_ids = torch.randint(0, num_buckets, (self.num_envs, 1))
            lo, hi = dr.friction_range
            friction_buckets = (hi - lo) * torch.rand(num_buckets, 1) + lo          # torch_rand_float on the CPU (LR:279)
            self.friction_coeffs = friction_buckets[bucket_ids]                      # [N, 1, 1] like the reference
        if dr.randomize_base_mass:
            lo, hi = dr.added_mass_range
            self.added_base_mass = torch.from_numpy(np.random.uniform(lo, hi, self.num_envs).astype(np.float32))
        gym = self.gym
        if gym is not None:
            if self.friction_coeffs is not None and hasattr(gym, "set_rigid_shape_friction"):
                gym.set_rigid_shape_friction(self.friction_coeffs)
            if self.added_base_mass is not None and hasattr(gym, "set_base_mass_offsets"):
                gym.set_base_mass_offsets(self.added_base_mass)

    def _actors_per_env(self):
        return 1

    def _extra_bodies_per_env(self):
        return 0

    def _configure_native(self, p):
        """hook for subclasses that hand extra facts to the kernels (LowLevelGame: predator spawn)"""

    def _push_resets_to_sim(self):
        # LR:409-412, 433-436: indexed state writes for the envs that were reset
        self.gym.set_dof_state_tensor_indexed(self.dof_state, self.reset_env_ids, self.reset_count)
        self.gym.set_actor_root_state_tensor_indexed(self.root_states, self.reset_env_ids, self.reset_count)

    # ------------------------------------------------------------------ LR:752-779
    def _get_env_origins(self):
        dev, N, t = self.device, self.num_envs, self.cfg.terrain
        self.env_origins = torch.zeros(N, 3, device=dev, requires_grad=False)
        if t.mesh_type in ("heightfield", "trimesh"):
            self.custom_origins = True
            max_init_level = t.max_init_terrain_level
            if not t.curriculum:
                max_init_level = t.num_rows - 1
            if self._init_levels_arg is not None:
                self.terrain_levels = torch.as_tensor(self._init_levels_arg, dtype=torch.long).to(dev).clone()
            else:
                self.terrain_levels = torch.randint(0, max_init_level + 1, (N,), device=dev)
            self.terrain_types = torch.div(torch.arange(N, device=dev), (N / t.num_cols), rounding_mode="floor").to(torch.long)
            self.max_terrain_level = t.num_rows
            self.terrain_origins = torch.from_numpy(np.asarray(self.terrain.env_origins)).to(dev).to(torch.float).contiguous()
            self.env_origins[:] = self.terrain_origins[self.terrain_levels, self.terrain_types]
        else:
            self.custom_origins = False
            num_cols = np.floor(np.sqrt(N))
            num_rows = np.ceil(N / num_cols)
            xx, yy = torch.meshgrid(torch.arange(num_rows), torch.arange(num_cols), indexing="ij")
            spacing = self.cfg.env.env_spacing
            self.env_origins[:, 0] = (spacing * xx.flatten()[:N]).to(dev)
            self.env_origins[:, 1] = (spacing * yy.flatten()[:N]).to(dev)
            self.env_origins[:, 2] = 0.

    # ------------------------------------------------------------------ LR:781-791
    def _parse_cfg(self, cfg):
        self.dt = self.cfg.control.decimation * self.sim_params.dt
        self.obs_scales = self.cfg.normalization.obs_scales
        self.reward_scales = class_to_dict(self.cfg.rewards.scales)
        self.command_ranges = class_to_dict(self.cfg.commands.ranges)
        if self.cfg.terrain.mesh_type not in ["heightfield", "trimesh"]:
            self.cfg.terrain.curriculum = False
        self.max_episode_length_s = self.cfg.env.episode_length_s
        self.max_episode_length = np.ceil(self.max_episode_length_s / self.dt)
        self.cfg.domain_rand.push_interval = np.ceil(self.cfg.domain_rand.push_interval_s / self.dt)

    # ------------------------------------------------------------------ LR:485-508
    def _get_noise_scale_vec(self, cfg):
        noise_vec = torch.zeros_like(self.obs_buf[0])
        self.add_noise = self.cfg.noise.add_noise
        ns, lvl, o = self.cfg.noise.noise_scales, self.cfg.noise.noise_level, self.obs_scales
        noise_vec[:3] = ns.lin_vel * lvl * o.lin_vel
        noise_vec[3:6] = ns.ang_vel * lvl * o.ang_vel
        noise_vec[6:9] = ns.gravity * lvl
        noise_vec[9:12] = 0.
        noise_vec[12:24] = ns.dof_pos * lvl * o.dof_pos
        noise_vec[24:36] = ns.dof_vel * lvl * o.dof_vel
        noise_vec[36:48] = 0.
        if self.cfg.terrain.measure_heights:
            noise_vec[48:235] = ns.height_measurements * lvl * o.height_measurements
        return noise_vec

    # ------------------------------------------------------------------ LR:511-581
    def _init_buffers(self):
        dev, N, D = self.device, self.num_envs, self.num_dof
        gym = self.gym
        self.root_states = gym.acquire_actor_root_state_tensor()
        self.dof_state = gym.acquire_dof_state_tensor()
        net_contact_forces = gym.acquire_net_contact_force_tensor()
        gym.refresh_dof_state_tensor()
        gym.refresh_actor_root_state_tensor()
        gym.refresh_net_contact_force_tensor()
        self.dof_pos = self.dof_state.view(N, D, 2)[..., 0]
        self.dof_vel = self.dof_state.view(N, D, 2)[..., 1]
        self.base_quat = self.root_states[self._root_rows(), 3:7]
        self.contact_forces = net_contact_forces.view(N, -1, 3)
        self.common_step_counter = 0
        self.extras = _Extras()
        self.noise_scale_vec = self._get_noise_scale_vec(self.cfg).contiguous()
        self.gravity_vec = torch.tensor([0., 0., -1.], device=dev).repeat((N, 1))
        self.forward_vec = torch.tensor([1., 0., 0.], device=dev).repeat((N, 1))
        z = lambda *s, dt=torch.float: torch.zeros(*s, dtype=dt, device=dev, requires_grad=False)
        self.torques = z(N, self.num_actions)
        self.p_gains = z(self.num_actions)
        self.d_gains = z(self.num_actions)
        self.actions = z(N, self.num_actions)
        self.last_actions = z(N, self.num_actions)
        self.last_dof_vel = z(N, D)
        self.last_root_vel = z(N, 6)
        self.commands = z(N, self.cfg.commands.num_commands)
        self.commands_scale = torch.tensor([self.obs_scales.lin_vel, self.obs_scales.lin_vel, self.obs_scales.ang_vel], device=dev)
        F = self.feet_indices.shape[0]
        self.feet_air_time = z(N, F)
        self.last_contacts = z(N, F, dt=torch.bool)
        self.base_lin_vel = z(N, 3)
        self.base_ang_vel = z(N, 3)
        self.projected_gravity = z(N, 3)
        if self.cfg.terrain.measure_heights:
            self.height_points = self._init_height_points()
            self.measured_heights = z(N, self.num_height_points)
        else:
            self.num_height_points = 0
            self.measured_heights = 0
        self.default_dof_pos = torch.zeros(D, dtype=torch.float, device="cpu")
        pg, dg = torch.zeros(D), torch.zeros(D)
        for i, name in enumerate(self.dof_names):
            self.default_dof_pos[i] = self.cfg.init_state.default_joint_angles[name]
            found = False
            for key in self.cfg.control.stiffness.keys():
                if key in name:
                    pg[i] = self.cfg.control.stiffness[key]
                    dg[i] = self.cfg.control.damping[key]
                    found = True
            if not found and self.cfg.control.control_type in ["P", "V"]:
                print(f"PD gain of joint {name} were not defined, setting them to zero")
        self.p_gains.copy_(pg)
        self.d_gains.copy_(dg)
        self.default_dof_pos = self.default_dof_pos.to(dev).unsqueeze(0)
        # kernel-side persistent buffers
        self._reset_bool = z(N, dt=torch.bool)
        self._time_outs_extras = z(N, dt=torch.bool)
        self.reset_env_ids = z(N, dt=torch.int32)
        self.reset_count = z(1, dt=torch.int32)
        # [0] completed steps, advanced by the finalize pass; [1] K1 -> K2 hand-over of lgk_post_physics_finalize
        self._step_counter_dev = z(2, dt=torch.int32)
        self._actions_in = z(N, self.num_actions)          # staging for the graph-replayed step
        self._graph, self._eager_steps = None, 0

    def _root_rows(self):
        return slice(None)

    # ------------------------------------------------------------------ LR:583-607
    def _prepare_reward_function(self):
        for key in list(self.reward_scales.keys()):
            if self.reward_scales[key] == 0:
                self.reward_scales.pop(key)
            else:
                self.reward_scales[key] *= self.dt
        self.reward_functions, self.reward_names = [], []
        for name in self.reward_scales.keys():
            if name == "termination":
                continue
            self.reward_names.append(name)
            self.reward_functions.append(getattr(self, "_reward_" + name))     # AttributeError like LR:602
        self._sum_names = list(self.reward_scales.keys())
        K = max(len(self._sum_names), 1)
        self._episode_sums_buf = torch.zeros(K, self.num_envs, dtype=torch.float, device=self.device)
        self.episode_sums = {n: self._episode_sums_buf[i] for i, n in enumerate(self._sum_names)}
        self._episode_means = torch.zeros(K + 1, dtype=torch.float, device=self.device)
        self._reset_stats = torch.zeros(2, K + 2, dtype=torch.float, device=self.device)
        # terms whose implementation is the class's own (native in the kernel) vs user-written torch terms
        self._python_reward_names = []
        for name in self.reward_names:
            fn = getattr(type(self), "_reward_" + name)
            native_owner = getattr(fn, "_lgk_native", False)
            if not (native_owner and name in nat.REWARD_TERMS):
                self._python_reward_names.append(name)

    # ------------------------------------------------------------------ native parameter blocks
    def _build_native_params(self):
        cfg, N = self.cfg, self.num_envs
        cls = type(self)
        self._native_torques = getattr(cls._compute_torques, "_lgk_native", False)
        self._obs_overridden = cls.compute_observations is not LeggedRobot.compute_observations
        # reset_idx is an extension point of the reference (README.md:56-66; its own Anymal overrides it, ANY:56-60): an
        # override is honoured by splitting the step around it (see post_physics_step)
        self._reset_overridden = not getattr(cls.reset_idx, "_lgk_native", False)
        self._in_step_reset = False
        # whole-step CUDA graph: needs the native torque path, no torch-written reward terms, the built-in
        # observations and a sim backend whose hooks enqueue nothing between the kernels
        self._graph_ok = (self._native_torques and not self._python_reward_names and not self._obs_overridden
                          and not self._reset_overridden
                          and getattr(self.gym, "graph_safe", False) and getattr(self, "use_cuda_graph", True))
        f = lambda t: [float(x) for x in t.detach().flatten().cpu().tolist()]
        # ---- torques (LR:371-395)
        tp = nat.TorqueParams()
        tp.num_envs = N
        ct = cfg.control.control_type
        if ct not in nat.CTRL:
            raise NameError(f"Unknown controller type: {ct}")
        tp.control_type = nat.CTRL[ct]
        tp.use_lstm = 0
        tp.action_scale = cfg.control.action_scale
        tp.clip_actions = cfg.normalization.clip_actions
        tp.sim_dt = self.sim_params.dt
        tp.p_gains[:] = f(self.p_gains)
        tp.d_gains[:] = f(self.d_gains)
        tp.torque_limits[:] = f(self.torque_limits)
        tp.default_dof_pos[:] = f(self.default_dof_pos)
        tp.dof_state = self.dof_state.data_ptr()
        tp.last_dof_vel = self.last_dof_vel.data_ptr()
        tp.torques = self.torques.data_ptr()
        self._tq_params = tp
        # ---- step
        p = nat.StepParams()
        p.num_envs, p.num_bodies, p.num_obs = N, int(self.contact_forces.shape[1]), self.num_obs      # bodies in the contact view (LR:529)
        mh = bool(cfg.terrain.measure_heights)
        p.num_height_points = self.num_height_points if mh else 0
        p.actors_per_env, p.root_actor_offset = self._actors_per_env(), 0
        self._configure_native(p)
        p.seed = int(getattr(cfg, "seed", 0) or 0) & 0xFFFFFFFFFFFFFFFF
        p.env_id_offset = int(getattr(self, "env_id_offset", 0))
        p.heading_command = int(bool(cfg.commands.heading_command))
        p.measure_heights = int(mh)
        p.terrain_is_plane = int(cfg.terrain.mesh_type == "plane")
        if mh and cfg.terrain.mesh_type == "none":
            raise NameError("Can't measure height with terrain mesh type 'none'")
        p.add_noise = int(bool(self.add_noise))
        p.only_positive_rewards = int(bool(cfg.rewards.only_positive_rewards))
        p.terrain_curriculum = int(bool(cfg.terrain.curriculum))
        p.custom_origins = int(self.custom_origins)
        p.send_timeouts = int(bool(cfg.env.send_timeouts))
        p.zero_lstm_on_reset = 0
        p.dt = self.dt
        p.resample_period = int(cfg.commands.resampling_time / self.dt)
        p.max_episode_length = float(self.max_episode_length)
        p.max_episode_length_s = float(self.max_episode_length_s)
        p.max_push_vel = cfg.domain_rand.max_push_vel_xy
        self._params = p
        self._refresh_command_ranges()
        o = self.obs_scales
        p.obs_scale_lin_vel, p.obs_scale_ang_vel, p.obs_scale_dof_pos = o.lin_vel, o.ang_vel, o.dof_pos
        p.obs_scale_dof_vel, p.obs_scale_height = o.dof_vel, o.height_measurements
        p.clip_obs = cfg.normalization.clip_observations
        r = cfg.rewards
        p.tracking_sigma, p.base_height_target, p.max_contact_force = r.tracking_sigma, r.base_height_target, r.max_contact_force
        p.soft_dof_vel_limit, p.soft_torque_limit = r.soft_dof_vel_limit, r.soft_torque_limit
        t = cfg.terrain
        p.border_size, p.horizontal_scale, p.vertical_scale = t.border_size, t.horizontal_scale, t.vertical_scale
        # FMA-based exact division: only for scales proven bit-identical to IEEE division (oracle/divcheck.c)
        hs32 = np.float32(t.horizontal_scale)
        if any(hs32 == np.float32(c) for c in (0.1, 0.05, 0.25)):
            p.horizontal_scale_recip = float(np.float32(1.0 / float(hs32)))
        p.default_dof_pos[:] = f(self.default_dof_pos)
        p.dof_pos_lo[:] = f(self.dof_pos_limits[:, 0])
        p.dof_pos_hi[:] = f(self.dof_pos_limits[:, 1])
        p.dof_vel_limits[:] = f(self.dof_vel_limits)
        p.torque_limits[:] = f(self.torque_limits)
        p.base_init_state[:] = f(self.base_init_state)
        for name, arr, cap in (("feet", self.feet_indices, nat.MAX_FEET), ("pen", self.penalised_contact_indices, nat.MAX_PEN),
                               ("term", self.termination_contact_indices, nat.MAX_TERM)):
            idx = [int(i) for i in arr.cpu().tolist()]
            if len(idx) > cap:
                raise ValueError(f"too many {name} bodies ({len(idx)} > {cap})")
            setattr(p, "num_" + name, len(idx))
            getattr(p, name + "_idx")[:len(idx)] = idx
        for k, name in enumerate(nat.REWARD_TERMS):
            active = name in self.reward_scales and name not in self._python_reward_names
            p.reward_active[k] = int(active)
            p.reward_scale[k] = float(self.reward_scales[name]) if active else 0.0
            p.reward_slot[k] = self._sum_names.index(name) if active else -1
        p.num_reward_slots = len(self._sum_names)
        ptr = lambda x: x.data_ptr()
        p.root_states, p.dof_state, p.contact_forces = ptr(self.root_states), ptr(self.dof_state), ptr(self.contact_forces)
        p.actions, p.torques, p.commands = ptr(self.actions), ptr(self.torques), ptr(self.commands)
        p.episode_length_buf, p.last_actions, p.last_dof_vel = ptr(self.episode_length_buf), ptr(self.last_actions), ptr(self.last_dof_vel)
        p.last_root_vel, p.feet_air_time, p.last_contacts = ptr(self.last_root_vel), ptr(self.feet_air_time), ptr(self.last_contacts)
        p.episode_sums = ptr(self._episode_sums_buf)
        p.env_origins = ptr(self.env_origins)
        if self.custom_origins:
            p.terrain_levels, p.terrain_types, p.terrain_origins = ptr(self.terrain_levels), ptr(self.terrain_types), ptr(self.terrain_origins)
            p.half_env_length = self.terrain.env_length / 2
            p.max_terrain_level = int(self.max_terrain_level)
            p.terrain_num_cols = int(self.terrain_origins.shape[1])
        p.base_lin_vel, p.base_ang_vel, p.projected_gravity = ptr(self.base_lin_vel), ptr(self.base_ang_vel), ptr(self.projected_gravity)
        p.obs_buf, p.rew_buf = ptr(self.obs_buf), ptr(self.rew_buf)
        p.reset_buf, p.time_out_buf = ptr(self._reset_bool), ptr(self.time_out_buf)
        p.noise_scale_vec = ptr(self.noise_scale_vec)
        p.reset_stats = ptr(self._reset_stats)
        p.step_counter_dev = ptr(self._step_counter_dev)
        self._scan_frames = torch.zeros(N, 8, dtype=torch.float, device=self.device)
        p.scan_frames = ptr(self._scan_frames)
        if mh:         # hand-over buffer of the 48 proprioceptive columns between the two step kernels
            self._obs_head = torch.zeros(N, 48, dtype=torch.float, device=self.device)
            p.obs_head = ptr(self._obs_head)
        dr = cfg.domain_rand
        p.push_interval = int(dr.push_interval) if dr.push_robots else 0
        p.host_state = int(bool(getattr(self.gym, "state_in_host_memory", False)))
        if mh:
            p.measured_heights = ptr(self.measured_heights)
            if not p.terrain_is_plane:
                hs = self.height_samples
                p.hf_rows, p.hf_cols = int(hs.shape[0]), int(hs.shape[1])
                tc = getattr(self, "_terrain_cache", None)
                if tc is not None and "min3" in tc:
                    self._height_min3 = tc["min3"]
                else:
                    self._height_min3 = torch.empty_like(hs)
                    nat.check(nat.lib.lgk_height_min3(hs.data_ptr(), self._height_min3.data_ptr(), p.hf_rows, p.hf_cols,
                                                      _stream_ptr()), "lgk_height_min3")
                    if tc is not None:
                        tc["min3"] = self._height_min3
                self._height_points_xy = self.height_points[0, :, :2].contiguous()
                p.height_min3 = ptr(self._height_min3)
                p.height_points_xy = ptr(self._height_points_xy)
        self._params = p

    def _refresh_command_ranges(self):
        """(lo, float(hi - lo)) per command, as torch_rand_float evaluates them (LR:353-366)."""
        p = getattr(self, "_params", None)
        if p is None:
            return
        r = self.command_ranges
        for i, k in enumerate(("lin_vel_x", "lin_vel_y", "ang_vel_yaw", "heading")):
            p.cmd_lo[i] = float(r[k][0])
            p.cmd_range[i] = float(r[k][1] - r[k][0])

    # ------------------------------------------------------------------ LR:471-483
    def update_command_curriculum(self, env_ids):
        if torch.mean(self.episode_sums["tracking_lin_vel"][env_ids]) / self.max_episode_length > \
                0.8 * self.reward_scales["tracking_lin_vel"]:
            r, mc = self.command_ranges["lin_vel_x"], self.cfg.commands.max_curriculum
            r[0] = np.clip(r[0] - 0.5, -mc, 0.)
            r[1] = np.clip(r[1] + 0.5, 0., mc)
            self._refresh_command_ranges()

    # ------------------------------------------------------------------ LR:371-395
    def _compute_torques(self, actions):
        tp = self._tq_params
        a = actions if (actions.is_cuda and actions.is_contiguous() and actions.dtype == torch.float32) else \
            actions.to(device=self.device, dtype=torch.float32).contiguous()
        tp.actions_in = a.data_ptr()
        tp.actions_clipped = None
        nat.check(nat.lib.lgk_compute_torques(C.byref(tp), _stream_ptr()), "lgk_compute_torques")
        return self.torques
    _compute_torques._lgk_native = True

    # ------------------------------------------------------------------ LR:815-829
    def _init_height_points(self):
        y = torch.tensor(self.cfg.terrain.measured_points_y, device=self.device, requires_grad=False)
        x = torch.tensor(self.cfg.terrain.measured_points_x, device=self.device, requires_grad=False)
        grid_x, grid_y = torch.meshgrid(x, y, indexing="ij")
        self.num_height_points = grid_x.numel()
        points = torch.zeros(self.num_envs, self.num_height_points, 3, device=self.device, requires_grad=False)
        points[:, :, 0] = grid_x.flatten()
        points[:, :, 1] = grid_y.flatten()
        return points

    # ------------------------------------------------------------------ LR:831-869
    def _get_heights(self, env_ids=None):
        if self.cfg.terrain.mesh_type == "plane":
            return torch.zeros(self.num_envs, self.num_height_points, device=self.device, requires_grad=False)
        elif self.cfg.terrain.mesh_type == "none":
            raise NameError("Can't measure height with terrain mesh type 'none'")
        p = self._params
        out = torch.empty(self.num_envs, self.num_height_points, device=self.device)
        nat.check(nat.lib.lgk_height_scan(self.root_states.data_ptr(), p.actors_per_env, p.root_actor_offset,
                                          self.num_envs, p.height_points_xy, self.num_height_points, p.height_min3,
                                          p.hf_rows, p.hf_cols, p.border_size, p.horizontal_scale, p.vertical_scale,
                                          out.data_ptr(), None, None, _stream_ptr()), "lgk_height_scan")
        return out if not env_ids else out[env_ids]


def _native_reward(name):
    def fn(self):
        raise RuntimeError(f"_reward_{name} is evaluated inside lgk_post_physics; it has no Python body")
    fn.__name__ = "_reward_" + name
    fn._lgk_native = True
    return fn


# the 19 terms of LR:872-969 exist as attributes (so _prepare_reward_function finds them, LR:602) but are computed
# by the kernel.  ``stumble`` keeps the reference's naming quirk: the cfg scale is ``feet_stumble`` (no method).
for _n in nat.REWARD_TERMS:
    if _n not in ("termination", "no_fly"):
        setattr(LeggedRobot, "_reward_" + _n, _native_reward(_n))
setattr(LeggedRobot, "_reward_termination", _native_reward("termination"))
