"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32, same op sequence) of the two hierarchical predator/prey games
of the reference, around an ``OracleEnv(kind="llg")`` low-level env:

    HighLevelGame     legged_gym/envs/a1_game/high_level_game.py      (HLG)  one 19-float observation, one reward
    DecHighLevelGame  legged_gym/envs/a1_game/dec_high_level_game.py  (DHLG) prey (16) / predator (3) observations, two rewards

Pinned bit-for-bit against the unmodified reference classes in tests/test_oracle_vs_reference.py (built there with
oracle/ref_loader.make_ref_game).  The low-level policy is outside the oracle: ``step`` takes the low-level actions.
Random draws of the second root / dof reset inside one step come from the GAME_* Philox streams (oracle/philox.py).
"""
import numpy as np
import torch

from . import philox
from .isaac_torch_utils import quat_rotate_inverse
from .legged_oracle import sorted_public_dict, yaw_only_apply, wrap_pi


class GameOracle:
    MAX_REL_POS = 100.

    def __init__(self, cfg, ll, variant="hl"):
        assert variant in ("hl", "dec")
        self.cfg, self.ll, self.variant = cfg, ll, variant
        N = ll.N
        self.N = N
        self.capture_dist = cfg.env.capture_dist
        # _parse_cfg (HLG:562-571, DHLG:578-588)
        self.command_ranges = sorted_public_dict(cfg.commands.ranges)
        self.max_episode_length_s = cfg.env.episode_length_s
        self.max_episode_length = np.ceil(self.max_episode_length_s / ll.dt)
        self.reset_buf = torch.ones(N, dtype=torch.long)
        self.episode_length_buf = torch.zeros(N, dtype=torch.long)
        self.time_out_buf = torch.zeros(N, dtype=torch.bool)
        self.curr_episode_step = torch.zeros(N, dtype=torch.long)
        self.extras = {}
        if variant == "hl":
            self.obs_buf = self.MAX_REL_POS * torch.ones(N, cfg.env.num_observations)
            self.rew_buf = torch.zeros(N)
            self.agents = {"": (sorted_public_dict(cfg.rewards.scales), cfg.rewards.only_positive_rewards)}
        else:
            self.obs_buf_prey = self.MAX_REL_POS * torch.ones(N, cfg.env.num_observations_prey)
            self.obs_buf_prey[:, 12:16] = 0
            self.obs_buf_pred = self.MAX_REL_POS * torch.ones(N, cfg.env.num_observations_predator)
            self.rew_buf_prey = torch.zeros(N)
            self.rew_buf_pred = torch.zeros(N)
            self.agents = {"pred": (sorted_public_dict(cfg.rewards_predator.scales), cfg.rewards_predator.only_positive_rewards),
                           "prey": (sorted_public_dict(cfg.rewards_prey.scales), cfg.rewards_prey.only_positive_rewards)}
        # _prepare_reward_function* (HLG:537-560, DHLG:527-576): drop zero scales, multiply by the low-level dt
        self.reward_scales, self.reward_names, self.episode_sums = {}, {}, {}
        for a, (scales, _) in self.agents.items():
            for k in list(scales.keys()):
                if scales[k] == 0:
                    scales.pop(k)
                else:
                    scales[k] *= ll.dt
            self.reward_scales[a] = scales
            self.reward_names[a] = [k for k in scales if k != "termination"]
            self.episode_sums[a] = {k: torch.zeros(N) for k in scales}
        # _init_buffers (HLG:519-535)
        self._update_agent_states()
        self.predator_pos = ll.root_states[ll.predator_indices, :3].clone()

    # HLG:510-517
    def _update_agent_states(self):
        ll = self.ll
        self.prey_states = ll.root_states[ll.prey_indices, :]
        self.base_quat = self.prey_states[:, 3:7]
        self.base_lin_vel = quat_rotate_inverse(self.base_quat, self.prey_states[:, 7:10])
        self.base_ang_vel = quat_rotate_inverse(self.base_quat, self.prey_states[:, 10:13])
        self.predator_pos = ll.root_states[ll.predator_indices, :3]

    # HLG:265-287
    def _step_predator(self, command):
        ll = self.ll
        for _ in range(ll.cfg.control.decimation):
            self.predator_pos[:, 0] += ll.cfg.sim.dt * command[:, 0]
            self.predator_pos[:, 1] += ll.cfg.sim.dt * command[:, 1]
        ll.root_states[ll.predator_indices, :3] = self.predator_pos

    def _clip_commands(self, prey, pred):
        r = self.command_ranges
        prey[:, 0] = torch.clip(prey[:, 0], min=r["lin_vel_x"][0], max=r["lin_vel_x"][1])
        prey[:, 1] = torch.clip(prey[:, 1], min=r["lin_vel_y"][0], max=r["lin_vel_y"][1])
        if self.cfg.commands.heading_command:
            prey[:, 2] = wrap_pi(prey[:, 2])
        pred[:, 0] = torch.clip(pred[:, 0], min=r["predator_lin_vel_x"][0], max=r["predator_lin_vel_x"][1])
        pred[:, 1] = torch.clip(pred[:, 1], min=r["predator_lin_vel_y"][0], max=r["predator_lin_vel_y"][1])

    # ------------------------------------------------------------------ HLG:146-241
    def step(self, command, ll_actions, tables):
        assert self.variant == "hl"
        ll = self.ll
        self.tables = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in tables.items()}
        self._clip_commands(command[:, 0:4], command[:, 4:6])
        ll.commands = command[:, :4]
        _, _, ll_rews, ll_dones, _ = ll.step(ll_actions, tables)
        self.curr_episode_step += 1
        self._step_predator(command[:, 4:])
        self._update_agent_states()
        self._compute_reward("", ll_rews)
        dist = torch.norm(self.prey_states[:, :2] - self.predator_pos[:, :2], dim=1)
        ids = (dist < self.capture_dist).nonzero(as_tuple=False).flatten()
        if self.cfg.env.env_radius is not None:
            a = torch.norm(self.prey_states[:, :2] - ll.env_origins[:, :2], dim=1) > self.cfg.env.env_radius
            b = torch.norm(self.predator_pos[:, :2] - ll.env_origins[:, :2], dim=1) > self.cfg.env.env_radius
            ids = torch.unique(torch.cat((ids, torch.logical_or(a, b).nonzero(as_tuple=False).flatten()), dim=-1))
        ids = torch.unique(torch.cat((ids, ll_dones.nonzero(as_tuple=False).flatten()), dim=-1))
        self.reset_idx(ids)
        done = torch.zeros_like(ll_dones)
        done[ids] = True
        self.reset_buf = done
        self.compute_observations()
        return self.obs_buf, None, self.rew_buf, self.reset_buf, self.extras

    # ------------------------------------------------------------------ DHLG:169-261
    def step_dec(self, command_pred, command_prey, ll_actions, tables):
        assert self.variant == "dec"
        ll = self.ll
        self.tables = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in tables.items()}
        self._clip_commands(command_prey, command_pred)
        ll.commands = command_prey
        _, _, ll_rews, ll_dones, _ = ll.step(ll_actions, tables)
        self._step_predator(command_pred)
        self._update_agent_states()
        self.episode_length_buf += 1
        self.curr_episode_step += 1
        # check_termination DHLG:263-268
        self.reset_buf = torch.norm(self.prey_states[:, :2] - self.predator_pos[:, :2], dim=-1) < self.capture_dist
        self.time_out_buf = self.episode_length_buf > self.max_episode_length
        self.reset_buf |= self.time_out_buf
        self._compute_reward("prey", ll_rews)
        self._compute_reward("pred", None)
        self.reset_buf |= ll.reset_buf
        self.reset_idx(self.reset_buf.nonzero(as_tuple=False).flatten())
        self.obs_buf_pred = self.prey_states[:, :3] - self.predator_pos                # DHLG:384-391
        self.compute_observations()
        return (self.obs_buf_pred, self.obs_buf_prey, None, None, self.rew_buf_pred, self.rew_buf_prey, self.reset_buf,
                self.extras)

    # ------------------------------------------------------------------ HLG:357-378, DHLG:321-361
    def _compute_reward(self, agent, ll_rews):
        if ll_rews is not None:
            rew = 2.0 * ll_rews
        else:
            rew = self.rew_buf_pred
            rew[:] = 0.
        scales = self.reward_scales[agent]
        for name in self.reward_names[agent]:
            d = torch.norm(self.predator_pos - self.prey_states[:, :3], dim=1)
            r = (d if name == "evasion" else -d) * scales[name]
            rew += r
            self.episode_sums[agent][name] += r
        if self.agents[agent][1]:
            rew[:] = torch.clip(rew[:], min=0.)
        if "termination" in scales:
            r = (self.reset_buf * ~self.time_out_buf) * scales["termination"]
            rew += r
            self.episode_sums[agent]["termination"] += r
        if agent == "":
            self.rew_buf = rew
        elif agent == "prey":
            self.rew_buf_prey = rew

    # ------------------------------------------------------------------ HLG:326-349, DHLG:270-312
    def reset_idx(self, ids):
        if len(ids) == 0:
            return
        ll = self.ll
        if self.variant == "dec":                                                         # ll_env._reset_dofs LLG:384-399
            u = self.tables[philox.STREAM_GAME_DOF][ids, 0:ll.D]
            ll.dof_pos[ids] = ll.default_dof_pos * ((1.5 - 0.5) * u + 0.5)
            ll.dof_vel[ids] = 0.
        # ll_env._reset_root_states LLG:401-451
        rid = ll.prey_indices[ids]
        ll.root_states[rid] = ll.base_init_state
        ll.root_states[rid, :3] += ll.env_origins[ids]
        root_u = self.tables[philox.STREAM_GAME_ROOT]
        if ll.custom_origins:
            ll.root_states[rid, :2] += (1. - -1.) * root_u[ids, 0:2] + -1.
        ll.root_states[rid, 7:13] = (0.5 - -0.5) * root_u[ids, 2:8] + -0.5
        init_prey_pos = ll.root_states[rid, :3].detach().clone()
        pu = self.tables[philox.STREAM_GAME_PREDATOR]
        rand_offset = (10.0 - 1.0) * pu[ids, 0:3] + 1.0
        rand_sign = pu[ids, 3].clone()
        lo = rand_sign < 0.5
        rand_sign[lo] = -1
        rand_sign[~lo] = 1
        ll.root_states[ll.predator_indices[ids], :3] = init_prey_pos - rand_sign.unsqueeze(1) * rand_offset
        ll.root_states[ll.predator_indices[ids], 2] = 0.3
        self._update_agent_states()
        if self.variant == "hl":
            self.obs_buf[ids, 0:12] = self.MAX_REL_POS
            self.obs_buf[ids, 12:16] = 0
            self.obs_buf[ids, 16:] = -self.MAX_REL_POS
            self.episode_length_buf[ids] = 0
            self.curr_episode_step[ids] = 0
            return
        self.obs_buf_prey[ids, 0:12] = self.MAX_REL_POS
        self.obs_buf_prey[ids, 12:16] = 0
        self.obs_buf_pred[ids, :] = -self.MAX_REL_POS
        self.episode_length_buf[ids] = 0
        self.reset_buf[ids] = 1
        self.curr_episode_step[ids] = 0
        self.extras["episode"] = {}
        for a in ("pred", "prey"):
            for k in self.episode_sums[a].keys():
                self.extras["episode"]["rew_" + a + "_" + k] = torch.mean(self.episode_sums[a][k][ids]) / self.max_episode_length_s
                self.episode_sums[a][k][ids] = 0.
        if self.cfg.env.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf

    # ------------------------------------------------------------------ HLG:380-482, DHLG:364-472
    def compute_observations(self):
        rel, sense = self.sense_predator()
        buf = self.obs_buf if self.variant == "hl" else self.obs_buf_prey
        parts = (buf[:, 3:12].clone(), rel, buf[:, 13:16].clone(), sense.long())
        if self.variant == "hl":
            self.obs_buf = torch.cat(parts + (self.prey_states[:, :3] - self.predator_pos,), dim=-1)
        else:
            self.obs_buf_prey = torch.cat(parts, dim=-1)

    def sense_predator(self):
        half_fov = 1.20428 / 2.
        ll = self.ll
        buf = self.obs_buf if self.variant == "hl" else self.obs_buf_prey
        rel = self.predator_pos - self.prey_states[:, :3]
        forward = yaw_only_apply(ll.base_quat, ll.forward_vec)
        dot = torch.sum(forward * rel, dim=-1).unsqueeze(-1)
        denom = torch.norm(forward, p=2, dim=-1, keepdim=True) * torch.norm(rel, p=2, dim=-1, keepdim=True)
        ang = wrap_pi(torch.acos(dot / denom))
        sense = rel.clone()
        vis = torch.any(torch.abs(ang) <= half_fov, dim=1)
        visible = torch.zeros(self.N, 1)
        visible[vis.nonzero(as_tuple=False).flatten(), :] = 1
        occ = (visible == 0).nonzero(as_tuple=False).flatten()
        sense[occ, :] = buf[occ, 9:12].clone()
        return sense, visible


def snapshot(g):
    """Comparable view of a game (reference instance, GameOracle, or the product class): name -> CPU tensor."""
    d = {}
    names = ["reset_buf", "episode_length_buf", "time_out_buf", "curr_episode_step", "predator_pos"]
    names += ["obs_buf", "rew_buf"] if hasattr(g, "obs_buf") else ["obs_buf_prey", "obs_buf_pred", "rew_buf_prey", "rew_buf_pred"]
    for n in names:
        d[n] = getattr(g, n).detach().cpu().clone()
    ll = g.ll_env if hasattr(g, "ll_env") else g.ll
    d["root_states"] = ll.root_states.detach().cpu().clone()
    d["dof_state"] = ll.dof_state.detach().cpu().clone()
    for k, v in g.extras.get("episode", {}).items():
        d["ex_" + k] = torch.as_tensor(v).detach().cpu().clone().float()
    return d


def game_sums(g):
    """episode sums by name, for the three implementations (reference attribute names differ per variant)."""
    if hasattr(g, "episode_sums_prey"):
        out = {"prey_" + k: v for k, v in g.episode_sums_prey.items()}
        out.update({"pred_" + k: v for k, v in g.episode_sums_pred.items()})
    elif isinstance(g.episode_sums, dict) and set(g.episode_sums.keys()) <= {"", "prey", "pred"}:
        out = {}
        for a, sums in g.episode_sums.items():
            out.update({(a + "_" if a else "") + k: v for k, v in sums.items()})
    else:
        out = dict(g.episode_sums)
    return {k: v.detach().cpu().clone() for k, v in out.items()}
