"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch, fp32) of the reference hot path.

Restates, in one explicit-state class, what /root/reference/legged_gym/envs/base/legged_robot.py
(LR), envs/anymal_c/anymal.py (ANY), envs/cassie/cassie.py and utils/math.py (MATH) compute
per environment step around the PhysX call.  Each method cites the reference lines it follows.
The torch op *sequence* inside every expression is kept as in the reference so that CPU rounding
is identical; ``tests/test_oracle_vs_reference.py`` pins this file bit-for-bit against the
unmodified reference executed through ``oracle/ref_loader.py`` (container only), and the
committed fixtures under ``tests/golden/`` (made by ``oracle/make_golden.py`` from the
REFERENCE's outputs) pin it on the GPU box where /root/reference does not exist.

Randomness is an explicit input: ``tables[stream]`` are per-env uniform tables from
``oracle/philox.py`` (see ref_loader.RngTap for how the same tables are fed to the reference).

Third-party maths (isaacgym.torch_utils) comes from oracle/isaac_torch_utils.py: PARITY UNPINNED
for those six helpers (no reference test covers them).
"""

import numpy as np
import torch

from . import philox
from .isaac_torch_utils import quat_rotate_inverse, quat_apply, normalize


def sorted_public_dict(obj):
    """helpers.py:41-56 ``class_to_dict`` restated: walks dir(obj) -> ALPHABETICAL key order."""
    if not hasattr(obj, "__dict__"):
        return obj
    out = {}
    for k in dir(obj):
        if k.startswith("_"):
            continue
        v = getattr(obj, k)
        out[k] = [sorted_public_dict(i) for i in v] if isinstance(v, list) else sorted_public_dict(v)
    return out


def yaw_only_apply(quat, vec):
    """MATH:38-42 quat_apply_yaw."""
    qy = quat.clone().view(-1, 4)
    qy[:, :2] = 0.
    qy = normalize(qy)
    return quat_apply(qy, vec)


def wrap_pi(a):
    """MATH:45-48 wrap_to_pi (in place on its argument, like the reference)."""
    a %= 2 * np.pi
    a -= 2 * np.pi * (a > np.pi)
    return a


class ActuatorLSTM:
    """ANY:53-54, 71-81: the TorchScript SEA model = in_scale * x -> LSTM(2,8,2 layers) -> Linear(8,1)
    * out_scale.  Weights come from an .npz dump of resources/actuator_nets/anydrive_v3_lstm.pt.
    ``forward`` runs aten::lstm (what the TorchScript module dispatches to) so CPU numerics equal
    the reference's; ``forward_plain`` is the gate-by-gate restatement used to document the math."""

    def __init__(self, weights, device="cpu"):
        w = {k: torch.as_tensor(np.asarray(v), dtype=torch.float32).to(device) for k, v in weights.items()}
        self.w = w
        self.lstm = torch.nn.LSTM(input_size=2, hidden_size=8, num_layers=2, batch_first=True).to(device)
        with torch.no_grad():
            for l in (0, 1):
                getattr(self.lstm, f"weight_ih_l{l}").copy_(w[f"weight_ih_l{l}"])
                getattr(self.lstm, f"weight_hh_l{l}").copy_(w[f"weight_hh_l{l}"])
                getattr(self.lstm, f"bias_ih_l{l}").copy_(w[f"bias_ih_l{l}"])
                getattr(self.lstm, f"bias_hh_l{l}").copy_(w[f"bias_hh_l{l}"])
        self.lin_w = w["linear_weight"]
        self.lin_b = w["linear_bias"]
        self.in_scale = w["in_scale"].view(1, 1, 2)
        self.out_scale = w["out_scale"].view(1)

    def forward(self, x, h, c):
        with torch.inference_mode():
            y, (hn, cn) = self.lstm(x * self.in_scale, (h, c))
            out = self.out_scale * torch.squeeze(torch.nn.functional.linear(y, self.lin_w, self.lin_b))
        return out, hn, cn

    def forward_plain(self, x, h, c):
        w = self.w
        xin = (x * self.in_scale)[:, 0, :]
        hs, cs = [], []
        for l in (0, 1):
            g = xin @ w[f"weight_ih_l{l}"].T + w[f"bias_ih_l{l}"] + h[l] @ w[f"weight_hh_l{l}"].T + w[f"bias_hh_l{l}"]
            i, f, gg, o = g[:, 0:8], g[:, 8:16], g[:, 16:24], g[:, 24:32]
            cn = torch.sigmoid(f) * c[l] + torch.sigmoid(i) * torch.tanh(gg)
            hn = torch.sigmoid(o) * torch.tanh(cn)
            hs.append(hn)
            cs.append(cn)
            xin = hn
        out = self.out_scale * (xin @ self.lin_w.T + self.lin_b).squeeze(-1)
        return out, torch.stack(hs), torch.stack(cs)


class OracleEnv:
    """Explicit-state restatement of LeggedRobot / Anymal / Cassie (hot path only).

    cfg     : a LeggedRobotCfg-shaped object (reference class or the product's mirror)
    consts  : see ref_loader.make_ref_env
    state   : dict(root_states[N,13], dof_state[N*D,2], contact_forces[N*NB,3]) CPU tensors, mutated in place
    kind    : "legged" | "anymal" | "cassie"   (which subclass behaviour to add)
    """

    def __init__(self, cfg, consts, state, kind="legged", height_samples=None, terrain_origins=None,
                 init_levels=None, lstm_weights=None):
        self.cfg = cfg
        self.kind = kind
        # the tensors live where the caller's state lives (CPU for every parity use; bench.py's "eager torch on the GPU" leg
        # builds the same object under ``with torch.device("cuda")``)
        self.device = state["root_states"].device
        N = cfg.env.num_envs
        self.N = N
        # ---- _parse_cfg (LR:781-791)
        self.sim_dt = cfg.sim.dt
        self.dt = cfg.control.decimation * cfg.sim.dt
        self.obs_scales = cfg.normalization.obs_scales
        self.reward_scales = sorted_public_dict(cfg.rewards.scales)
        self.command_ranges = sorted_public_dict(cfg.commands.ranges)
        self.curriculum = cfg.terrain.curriculum and cfg.terrain.mesh_type in ("heightfield", "trimesh")
        self.max_episode_length_s = cfg.env.episode_length_s
        self.max_episode_length = np.ceil(self.max_episode_length_s / self.dt)
        self.push_interval = np.ceil(cfg.domain_rand.push_interval_s / self.dt)
        # ---- buffers (BT:71-79)
        self.num_obs = cfg.env.num_observations
        self.num_actions = cfg.env.num_actions
        self.obs_buf = torch.zeros(N, self.num_obs)
        self.rew_buf = torch.zeros(N)
        self.reset_buf = torch.ones(N, dtype=torch.long)
        self.episode_length_buf = torch.zeros(N, dtype=torch.long)
        self.time_out_buf = torch.zeros(N, dtype=torch.bool)
        self.extras = {}
        # ---- asset constants (LR:299-313, 733-750)
        D = len(consts["dof_names"])
        self.D = D
        self.feet = torch.tensor(consts["feet_indices"], dtype=torch.long)
        self.penalised = torch.tensor(consts["penalised_contact_indices"], dtype=torch.long)
        self.term_idx = torch.tensor(consts["termination_contact_indices"], dtype=torch.long)
        self.dof_pos_limits = torch.zeros(D, 2)
        self.dof_vel_limits = torch.zeros(D)
        self.torque_limits = torch.zeros(D)
        for i in range(D):
            self.dof_pos_limits[i, 0] = float(np.float32(consts["dof_lower"][i]))
            self.dof_pos_limits[i, 1] = float(np.float32(consts["dof_upper"][i]))
            self.dof_vel_limits[i] = float(np.float32(consts["dof_vel_limits"][i]))
            self.torque_limits[i] = float(np.float32(consts["torque_limits"][i]))
            m = (self.dof_pos_limits[i, 0] + self.dof_pos_limits[i, 1]) / 2
            r = self.dof_pos_limits[i, 1] - self.dof_pos_limits[i, 0]
            self.dof_pos_limits[i, 0] = m - 0.5 * r * cfg.rewards.soft_dof_pos_limit
            self.dof_pos_limits[i, 1] = m + 0.5 * r * cfg.rewards.soft_dof_pos_limit
        ini = cfg.init_state
        self.base_init_state = torch.tensor(ini.pos + ini.rot + ini.lin_vel + ini.ang_vel, dtype=torch.float)
        # ---- origins (LR:752-779)
        self.custom_origins = cfg.terrain.mesh_type in ("heightfield", "trimesh")
        self.env_origins = torch.zeros(N, 3)
        if self.custom_origins:
            self.terrain_levels = torch.as_tensor(init_levels, dtype=torch.long).clone().to(self.device)
            self.terrain_types = torch.div(torch.arange(N), (N / cfg.terrain.num_cols), rounding_mode="floor").to(torch.long)
            self.max_terrain_level = cfg.terrain.num_rows
            self.terrain_origins = torch.from_numpy(np.asarray(terrain_origins)).to(torch.float).to(self.device)
            self.env_origins[:] = self.terrain_origins[self.terrain_levels, self.terrain_types]
        else:
            ncol = np.floor(np.sqrt(N))
            nrow = np.ceil(N / ncol)
            xx, yy = torch.meshgrid(torch.arange(nrow), torch.arange(ncol), indexing="ij")
            self.env_origins[:, 0] = cfg.env.env_spacing * xx.flatten()[:N]
            self.env_origins[:, 1] = cfg.env.env_spacing * yy.flatten()[:N]
        self.height_samples = height_samples
        self.env_length = cfg.terrain.terrain_length
        # ---- _init_buffers (LR:511-581)
        self.root_states = state["root_states"]
        # low_level_game (LLG:123-125, 760-812): two actors per env, prey (robot) first, then the predator sphere
        self.llg = kind == "llg"
        if self.llg:
            self.prey_indices = torch.arange(N) * 2
            self.predator_indices = torch.arange(N) * 2 + 1
            self.rows = self.prey_indices
        else:
            self.rows = slice(None)
        self.dof_state = state["dof_state"]
        self.dof_pos = self.dof_state.view(N, D, 2)[..., 0]
        self.dof_vel = self.dof_state.view(N, D, 2)[..., 1]
        self.base_quat = self.root_states[self.rows, 3:7]      # a copy for llg (advanced indexing), refreshed every step
        self.contact_forces = state["contact_forces"].view(N, -1, 3)
        self.common_step_counter = 0
        self.gravity_vec = torch.tensor([0., 0., -1.]).repeat((N, 1))
        self.forward_vec = torch.tensor([1., 0., 0.]).repeat((N, 1))
        self.torques = torch.zeros(N, self.num_actions)
        self.p_gains = torch.zeros(self.num_actions)
        self.d_gains = torch.zeros(self.num_actions)
        self.actions = torch.zeros(N, self.num_actions)
        self.last_actions = torch.zeros(N, self.num_actions)
        self.last_dof_vel = torch.zeros_like(self.dof_vel)
        self.last_root_vel = torch.zeros_like(self.root_states[self.rows, 7:13])
        self.commands = torch.zeros(N, cfg.commands.num_commands)
        self.commands_scale = torch.tensor([self.obs_scales.lin_vel, self.obs_scales.lin_vel, self.obs_scales.ang_vel])
        self.feet_air_time = torch.zeros(N, self.feet.shape[0])
        self.last_contacts = torch.zeros(N, len(self.feet), dtype=torch.bool)
        self.base_lin_vel = quat_rotate_inverse(self.base_quat, self.root_states[self.rows, 7:10])
        self.base_ang_vel = quat_rotate_inverse(self.base_quat, self.root_states[self.rows, 10:13])
        self.projected_gravity = quat_rotate_inverse(self.base_quat, self.gravity_vec)
        self.measure_heights = cfg.terrain.measure_heights
        if self.measure_heights:
            ys = torch.tensor(cfg.terrain.measured_points_y)
            xs = torch.tensor(cfg.terrain.measured_points_x)
            gx, gy = torch.meshgrid(xs, ys, indexing="ij")          # LR:821-829 (x outer, y inner)
            self.num_height_points = gx.numel()
            self.height_points = torch.zeros(N, self.num_height_points, 3)
            self.height_points[:, :, 0] = gx.flatten()
            self.height_points[:, :, 1] = gy.flatten()
        self.measured_heights = 0
        self.default_dof_pos = torch.zeros(D)
        for i, name in enumerate(consts["dof_names"]):
            self.default_dof_pos[i] = cfg.init_state.default_joint_angles[name]
            for key in cfg.control.stiffness.keys():
                if key in name:
                    self.p_gains[i] = cfg.control.stiffness[key]
                    self.d_gains[i] = cfg.control.damping[key]
        self.default_dof_pos = self.default_dof_pos.unsqueeze(0)
        self.noise_scale_vec = self._noise_scale_vec()
        # ---- _prepare_reward_function (LR:583-607)
        for k in list(self.reward_scales.keys()):
            if self.reward_scales[k] == 0:
                self.reward_scales.pop(k)
            else:
                self.reward_scales[k] *= self.dt
        self.reward_names = [k for k in self.reward_scales if k != "termination"]
        self.episode_sums = {k: torch.zeros(N) for k in self.reward_scales}
        # ---- Anymal (ANY:62-69)
        self.use_actuator_net = kind == "anymal" and getattr(cfg.control, "use_actuator_network", False)
        if self.use_actuator_net:
            self.actuator = ActuatorLSTM(lstm_weights, self.device)
            self.sea_input = torch.zeros(N * self.num_actions, 1, 2)
            self.sea_hidden_state = torch.zeros(2, N * self.num_actions, 8)
            self.sea_cell_state = torch.zeros(2, N * self.num_actions, 8)
        self.init_done = True

    # ------------------------------------------------------------------ LR:485-508
    def _noise_scale_vec(self):
        c = self.cfg.noise
        v = torch.zeros(self.num_obs)
        self.add_noise = c.add_noise
        s, lvl, o = c.noise_scales, c.noise_level, self.obs_scales
        v[:3] = s.lin_vel * lvl * o.lin_vel
        v[3:6] = s.ang_vel * lvl * o.ang_vel
        v[6:9] = s.gravity * lvl
        v[9:12] = 0.
        v[12:24] = s.dof_pos * lvl * o.dof_pos
        v[24:36] = s.dof_vel * lvl * o.dof_vel
        v[36:48] = 0.
        if self.measure_heights:
            v[48:235] = s.height_measurements * lvl * o.height_measurements
        return v

    # ------------------------------------------------------------------ LR:371-395, ANY:71-81
    def compute_torques(self, actions):
        if self.use_actuator_net:
            self.sea_input[:, 0, 0] = (actions * self.cfg.control.action_scale + self.default_dof_pos - self.dof_pos).flatten()
            self.sea_input[:, 0, 1] = self.dof_vel.flatten()
            tq, h, c = self.actuator.forward(self.sea_input, self.sea_hidden_state, self.sea_cell_state)
            self.sea_hidden_state[:] = h
            self.sea_cell_state[:] = c
            return tq
        a = actions * self.cfg.control.action_scale
        ct = self.cfg.control.control_type
        if ct == "P":
            tq = self.p_gains * (a + self.default_dof_pos - self.dof_pos) - self.d_gains * self.dof_vel
        elif ct == "V":
            tq = self.p_gains * (a - self.dof_vel) - self.d_gains * (self.dof_vel - self.last_dof_vel) / self.sim_dt
        elif ct == "T":
            tq = a
        else:
            raise NameError(f"Unknown controller type: {ct}")
        return torch.clip(tq, -self.torque_limits, self.torque_limits)

    # ------------------------------------------------------------------ LR:80-104
    def step(self, actions, tables, sim=None):
        """sim(substep) is called where PhysX would run (LR:92-96); None = state frozen."""
        ca = self.cfg.normalization.clip_actions
        self.actions = torch.clip(actions, -ca, ca)
        for k in range(self.cfg.control.decimation):
            self.torques = self.compute_torques(self.actions).view(self.torques.shape)
            if sim is not None:
                sim(k)
        self.post_physics_step(tables)
        co = self.cfg.normalization.clip_observations
        self.obs_buf = torch.clip(self.obs_buf, -co, co)
        return self.obs_buf, None, self.rew_buf, self.reset_buf, self.extras

    # ------------------------------------------------------------------ LR:106-137
    def post_physics_step(self, tables):
        self.tables = {k: (v if torch.is_tensor(v) else torch.from_numpy(np.ascontiguousarray(v))).to(self.device)
                       for k, v in tables.items()}
        self.episode_length_buf += 1
        self.common_step_counter += 1
        if self.llg:
            self.base_quat[:] = self.root_states[self.rows, 3:7]                  # LLG:123
        self.base_lin_vel[:] = quat_rotate_inverse(self.base_quat, self.root_states[self.rows, 7:10])
        self.base_ang_vel[:] = quat_rotate_inverse(self.base_quat, self.root_states[self.rows, 10:13])
        self.projected_gravity[:] = quat_rotate_inverse(self.base_quat, self.gravity_vec)
        self._callback()
        self.check_termination()
        self.compute_reward()
        ids = self.reset_buf.nonzero(as_tuple=False).flatten()
        self.reset_idx(ids)
        self.compute_observations()
        self.last_actions[:] = self.actions[:]
        self.last_dof_vel[:] = self.dof_vel[:]
        self.last_root_vel[:] = self.root_states[self.rows, 7:13]

    # ------------------------------------------------------------------ LR:329-345
    def _callback(self):
        period = int(self.cfg.commands.resampling_time / self.dt)
        ids = (self.episode_length_buf % period == 0).nonzero(as_tuple=False).flatten()
        self.resample_commands(ids, philox.STREAM_CMD)
        if self.cfg.commands.heading_command:
            fwd = quat_apply(self.base_quat, self.forward_vec)
            heading = torch.atan2(fwd[:, 1], fwd[:, 0])
            self.commands[:, 2] = torch.clip(0.5 * wrap_pi(self.commands[:, 3] - heading), -1., 1.)
        if self.measure_heights:
            self.measured_heights = self.get_heights()
        if self.cfg.domain_rand.push_robots and (self.common_step_counter % self.push_interval == 0):
            mv = self.cfg.domain_rand.max_push_vel_xy                       # LR:438-444
            u = self.tables[philox.STREAM_PUSH][:, 0:2]
            self.root_states[self.rows, 7:9] = (mv - -mv) * u + -mv

    def _draw(self, lo, hi, stream, ids, c0, c1):
        return (hi - lo) * self.tables[stream][ids, c0:c1] + lo

    # ------------------------------------------------------------------ LR:347-369
    def resample_commands(self, ids, stream):
        r = self.command_ranges
        self.commands[ids, 0] = self._draw(r["lin_vel_x"][0], r["lin_vel_x"][1], stream, ids, 0, 1).squeeze(1)
        self.commands[ids, 1] = self._draw(r["lin_vel_y"][0], r["lin_vel_y"][1], stream, ids, 1, 2).squeeze(1)
        if self.cfg.commands.heading_command:
            self.commands[ids, 3] = self._draw(r["heading"][0], r["heading"][1], stream, ids, 2, 3).squeeze(1)
        else:
            self.commands[ids, 2] = self._draw(r["ang_vel_yaw"][0], r["ang_vel_yaw"][1], stream, ids, 2, 3).squeeze(1)
        self.commands[ids, :2] *= (torch.norm(self.commands[ids, :2], dim=1) > 0.2).unsqueeze(1)

    # ------------------------------------------------------------------ LR:831-869
    def get_heights(self):
        if self.cfg.terrain.mesh_type == "plane":
            return torch.zeros(self.N, self.num_height_points)
        if self.cfg.terrain.mesh_type == "none":
            raise NameError("Can't measure height with terrain mesh type 'none'")
        pts = yaw_only_apply(self.base_quat.repeat(1, self.num_height_points), self.height_points) + \
            (self.root_states[self.rows, :3]).unsqueeze(1)
        pts += self.cfg.terrain.border_size
        pts = (pts / self.cfg.terrain.horizontal_scale).long()
        px = torch.clip(pts[:, :, 0].view(-1), 0, self.height_samples.shape[0] - 2)
        py = torch.clip(pts[:, :, 1].view(-1), 0, self.height_samples.shape[1] - 2)
        self.last_px, self.last_py = px, py      # exposed so tests can assert index bit-exactness
        h = torch.min(self.height_samples[px, py], self.height_samples[px + 1, py])
        h = torch.min(h, self.height_samples[px, py + 1])
        return h.view(self.N, -1) * self.cfg.terrain.vertical_scale

    # ------------------------------------------------------------------ LR:139-145
    def check_termination(self):
        self.reset_buf = torch.any(torch.norm(self.contact_forces[:, self.term_idx, :], dim=-1) > 1., dim=1)
        self.time_out_buf = self.episode_length_buf > self.max_episode_length
        self.reset_buf |= self.time_out_buf

    # ------------------------------------------------------------------ LR:193-210
    def compute_reward(self):
        self.rew_buf[:] = 0.
        self.rew_terms = {}
        for name in self.reward_names:
            rew = getattr(self, "_r_" + name)() * self.reward_scales[name]
            self.rew_buf += rew
            self.episode_sums[name] += rew
            self.rew_terms[name] = rew
        if self.cfg.rewards.only_positive_rewards:
            self.rew_buf[:] = torch.clip(self.rew_buf[:], min=0.)
        if "termination" in self.reward_scales:
            rew = self._r_termination() * self.reward_scales["termination"]
            self.rew_buf += rew
            self.episode_sums["termination"] += rew

    # ------------------------------------------------------------------ LR:147-191, ANY:56-60
    def reset_idx(self, ids):
        if len(ids) == 0:
            return
        if self.curriculum:
            self._terrain_curriculum(ids)
        if self.cfg.commands.curriculum and (self.common_step_counter % self.max_episode_length == 0):
            self._command_curriculum(ids)
        # _reset_dofs LR:397-412
        self.dof_pos[ids] = self.default_dof_pos * self._draw(0.5, 1.5, philox.STREAM_RESET_DOF, ids, 0, self.D)
        self.dof_vel[ids] = 0.
        # _reset_root_states LR:414-436
        rid = self.rows[ids] if self.llg else ids
        self.root_states[rid] = self.base_init_state
        self.root_states[rid, :3] += self.env_origins[ids]
        if self.custom_origins:
            self.root_states[rid, :2] += self._draw(-1., 1., philox.STREAM_RESET_ROOT, ids, 0, 2)
        self.root_states[rid, 7:13] = self._draw(-0.5, 0.5, philox.STREAM_RESET_ROOT, ids, 2, 8)
        if self.llg:                                                               # LLG:419-432
            init_prey_pos = self.root_states[rid, :3].detach().clone()
            rand_offset = self._draw(1.0, 10.0, philox.STREAM_PREDATOR, ids, 0, 3)
            rand_sign = self.tables[philox.STREAM_PREDATOR][ids, 3].clone()
            lo = rand_sign < 0.5
            rand_sign[lo] = -1
            rand_sign[~lo] = 1
            offset = rand_sign.unsqueeze(1) * rand_offset
            self.root_states[self.predator_indices[ids], :3] = init_prey_pos - offset
            self.root_states[self.predator_indices[ids], 2] = 0.3
        self.resample_commands(ids, philox.STREAM_RESET_CMD)
        self.last_actions[ids] = 0.
        self.last_dof_vel[ids] = 0.
        self.feet_air_time[ids] = 0.
        self.episode_length_buf[ids] = 0
        self.reset_buf[ids] = 1
        self.extras["episode"] = {}
        for k in self.episode_sums.keys():
            self.extras["episode"]["rew_" + k] = torch.mean(self.episode_sums[k][ids]) / self.max_episode_length_s
            self.episode_sums[k][ids] = 0.
        if self.curriculum:
            self.extras["episode"]["terrain_level"] = torch.mean(self.terrain_levels.float())
        if self.cfg.commands.curriculum:
            self.extras["episode"]["max_command_x"] = self.command_ranges["lin_vel_x"][1]
        if self.cfg.env.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf
        if self.use_actuator_net:
            self.sea_hidden_state.view(2, self.N, self.num_actions, 8)[:, ids] = 0.
            self.sea_cell_state.view(2, self.N, self.num_actions, 8)[:, ids] = 0.

    # ------------------------------------------------------------------ LR:446-469
    def _terrain_curriculum(self, ids):
        if not self.init_done:
            return
        dist = torch.norm(self.root_states[self.rows[ids] if self.llg else ids, :2] - self.env_origins[ids, :2], dim=1)
        up = dist > self.env_length / 2
        down = (dist < torch.norm(self.commands[ids, :2], dim=1) * self.max_episode_length_s * 0.5) * ~up
        self.terrain_levels[ids] += 1 * up - 1 * down
        rnd = (self.tables[philox.STREAM_TERRAIN][ids, 0].to(torch.int64) % int(self.max_terrain_level))
        self.terrain_levels[ids] = torch.where(self.terrain_levels[ids] >= self.max_terrain_level, rnd,
                                               torch.clip(self.terrain_levels[ids], 0))
        self.env_origins[ids] = self.terrain_origins[self.terrain_levels[ids], self.terrain_types[ids]]

    # ------------------------------------------------------------------ LR:471-483
    def _command_curriculum(self, ids):
        if torch.mean(self.episode_sums["tracking_lin_vel"][ids]) / self.max_episode_length > \
                0.8 * self.reward_scales["tracking_lin_vel"]:
            r, mc = self.command_ranges["lin_vel_x"], self.cfg.commands.max_curriculum
            r[0] = np.clip(r[0] - 0.5, -mc, 0.)
            r[1] = np.clip(r[1] + 0.5, 0., mc)

    # ------------------------------------------------------------------ LR:212-230
    def compute_observations(self):
        o = self.obs_scales
        self.obs_buf = torch.cat((self.base_lin_vel * o.lin_vel, self.base_ang_vel * o.ang_vel,
                                  self.projected_gravity, self.commands[:, :3] * self.commands_scale,
                                  (self.dof_pos - self.default_dof_pos) * o.dof_pos,
                                  self.dof_vel * o.dof_vel, self.actions), dim=-1)
        if self.measure_heights:
            h = torch.clip(self.root_states[self.rows, 2].unsqueeze(1) - 0.5 - self.measured_heights, -1, 1.) * o.height_measurements
            self.obs_buf = torch.cat((self.obs_buf, h), dim=-1)
        if self.add_noise:
            self.obs_buf += (2 * self.tables[philox.STREAM_OBS].clone() - 1) * self.noise_scale_vec

    # ------------------------------------------------------------------ reward terms LR:872-969
    def _r_lin_vel_z(self):
        return torch.square(self.base_lin_vel[:, 2])

    def _r_ang_vel_xy(self):
        return torch.sum(torch.square(self.base_ang_vel[:, :2]), dim=1)

    def _r_orientation(self):
        return torch.sum(torch.square(self.projected_gravity[:, :2]), dim=1)

    def _r_base_height(self):
        bh = torch.mean(self.root_states[self.rows, 2].unsqueeze(1) - self.measured_heights, dim=1)
        return torch.square(bh - self.cfg.rewards.base_height_target)

    def _r_torques(self):
        return torch.sum(torch.square(self.torques), dim=1)

    def _r_dof_vel(self):
        return torch.sum(torch.square(self.dof_vel), dim=1)

    def _r_dof_acc(self):
        return torch.sum(torch.square((self.last_dof_vel - self.dof_vel) / self.dt), dim=1)

    def _r_action_rate(self):
        return torch.sum(torch.square(self.last_actions - self.actions), dim=1)

    def _r_collision(self):
        return torch.sum(1. * (torch.norm(self.contact_forces[:, self.penalised, :], dim=-1) > 0.1), dim=1)

    def _r_termination(self):
        return self.reset_buf * ~self.time_out_buf

    def _r_dof_pos_limits(self):
        out = -(self.dof_pos - self.dof_pos_limits[:, 0]).clip(max=0.)
        out += (self.dof_pos - self.dof_pos_limits[:, 1]).clip(min=0.)
        return torch.sum(out, dim=1)

    def _r_dof_vel_limits(self):
        return torch.sum((torch.abs(self.dof_vel) - self.dof_vel_limits * self.cfg.rewards.soft_dof_vel_limit).clip(min=0., max=1.), dim=1)

    def _r_torque_limits(self):
        return torch.sum((torch.abs(self.torques) - self.torque_limits * self.cfg.rewards.soft_torque_limit).clip(min=0.), dim=1)

    def _r_tracking_lin_vel(self):
        e = torch.sum(torch.square(self.commands[:, :2] - self.base_lin_vel[:, :2]), dim=1)
        return torch.exp(-e / self.cfg.rewards.tracking_sigma)

    def _r_tracking_ang_vel(self):
        e = torch.square(self.commands[:, 2] - self.base_ang_vel[:, 2])
        return torch.exp(-e / self.cfg.rewards.tracking_sigma)

    def _r_feet_air_time(self):
        contact = self.contact_forces[:, self.feet, 2] > 1.
        filt = torch.logical_or(contact, self.last_contacts)
        self.last_contacts = contact
        first = (self.feet_air_time > 0.) * filt
        self.feet_air_time += self.dt
        r = torch.sum((self.feet_air_time - 0.5) * first, dim=1)
        r *= torch.norm(self.commands[:, :2], dim=1) > 0.1
        self.feet_air_time *= ~filt
        return r

    def _r_stumble(self):
        return torch.any(torch.norm(self.contact_forces[:, self.feet, :2], dim=2) >
                         5 * torch.abs(self.contact_forces[:, self.feet, 2]), dim=1)

    def _r_stand_still(self):
        return torch.sum(torch.abs(self.dof_pos - self.default_dof_pos), dim=1) * (torch.norm(self.commands[:, :2], dim=1) < 0.1)

    def _r_feet_contact_forces(self):
        return torch.sum((torch.norm(self.contact_forces[:, self.feet, :], dim=-1) - self.cfg.rewards.max_contact_force).clip(min=0.), dim=1)

    def _r_no_fly(self):                                            # cassie.py:43-46
        contacts = self.contact_forces[:, self.feet, 2] > 0.1
        return 1. * (torch.sum(1. * contacts, dim=1) == 1)
