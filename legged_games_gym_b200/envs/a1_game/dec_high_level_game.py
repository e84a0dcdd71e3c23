"""DecHighLevelGame -- the decentralised variant (reference legged_gym/envs/a1_game/dec_high_level_game.py:26-605): the
prey (robot) and the predator are separate agents with their own observations (16 / 3 floats), commands (4 / 2) and
rewards; episodes also end on a time-out, and a reset re-draws the low-level dof state as well (DHLG:283-284).  Same
kernel as HighLevelGame (``lgk_game_step`` with variant = 1)."""
import torch

from ... import _native as nat
from .high_level_game import _GameBase


class DecHighLevelGame(_GameBase):
    VARIANT = 1

    def _alloc_agent_buffers(self):
        cfg, N, dev = self.cfg, self.num_envs, self.device
        self.num_obs_prey, self.num_obs_pred = cfg.env.num_observations_prey, cfg.env.num_observations_predator
        self.num_privileged_obs_prey, self.num_privileged_obs_pred = cfg.env.num_privileged_obs_prey, cfg.env.num_privileged_obs_predator
        self.num_actions_prey, self.num_actions_pred = cfg.env.num_actions_prey, cfg.env.num_actions_predator
        self.obs_buf_prey = self.MAX_REL_POS * torch.ones(N, self.num_obs_prey, device=dev, dtype=torch.float)
        self.obs_buf_prey[:, 12:16] = 0
        self.rew_buf_prey = torch.zeros(N, device=dev, dtype=torch.float)
        self.obs_buf_pred = self.MAX_REL_POS * torch.ones(N, self.num_obs_pred, device=dev, dtype=torch.float)
        self.rew_buf_pred = torch.zeros(N, device=dev, dtype=torch.float)
        self.privileged_obs_buf_pred = self.privileged_obs_buf_prey = None
        self._params = nat.GameParams()

    def _prepare_rewards(self):
        p, cfg = self._params, self.cfg
        if "termination" in {k for k, v in vars(cfg.rewards_predator.scales).items() if not k.startswith("_") and v != 0}:
            # the reference reads self.reward_scales["termination"] here, an attribute this class never sets (DHLG:358-359)
            raise AttributeError("'DecHighLevelGame' object has no attribute 'reward_scales'")
        self.reward_scales_pred, self.reward_names_pred, self.episode_sums_pred, self._sums_pred = self._agent(
            p.pred, cfg.rewards_predator.scales, cfg.rewards_predator.only_positive_rewards, self.rew_buf_pred)
        self.reward_scales_prey, self.reward_names_prey, self.episode_sums_prey, self._sums_prey = self._agent(
            p.prey, cfg.rewards_prey.scales, cfg.rewards_prey.only_positive_rewards, self.rew_buf_prey)
        p.num_prey_slots, p.num_pred_slots = len(self.reward_scales_prey), len(self.reward_scales_pred)
        p.obs_prey, p.obs_prey_stride = self.obs_buf_prey.data_ptr(), self.obs_buf_prey.stride(0)
        p.obs_pred, p.obs_pred_stride = self.obs_buf_pred.data_ptr(), self.obs_buf_pred.stride(0)
        nk = p.num_prey_slots + p.num_pred_slots
        self._ep_means = torch.zeros(nk, device=self.device)
        names = ["rew_pred_" + k for k in self.reward_scales_pred] + ["rew_prey_" + k for k in self.reward_scales_prey]
        order = list(range(p.num_prey_slots, nk)) + list(range(p.num_prey_slots))      # stats hold the prey slots first
        self._ep_index = torch.tensor(order, device=self.device, dtype=torch.long)
        self._ep_names = names

    # ------------------------------------------------------------------ DHLG:169-261
    def step(self, command_pred, command_prey):
        self._ll_step(command_prey, command_pred)
        self.common_step_counter += 1
        self._stats.zero_()
        self._launch(command_pred)
        self._after_reset_stats()
        self._push_game_state_to_sim()
        return (self.obs_buf_pred, self.obs_buf_prey, self.privileged_obs_buf_pred, self.privileged_obs_buf_prey,
                self.rew_buf_pred, self.rew_buf_prey, self.reset_buf, self.extras)

    def _after_reset_stats(self):
        """DHLG:298-312 without a host synchronisation: the means are refreshed on the device only when an env was reset
        this step (the reference leaves extras untouched otherwise)."""
        p = self._params
        nk = p.num_prey_slots + p.num_pred_slots
        cnt = self._stats[nk]
        fresh = self._stats[:nk][self._ep_index] / cnt.clamp(min=1.0) / self.max_episode_length_s
        self._ep_means.copy_(torch.where(cnt > 0, fresh, self._ep_means))
        self.extras["episode"] = {n: self._ep_means[i] for i, n in enumerate(self._ep_names)}
        if self.cfg.env.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf

    def reset(self):
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        a_pred = torch.zeros(self.num_envs, self.num_actions_pred, device=self.device, requires_grad=False)
        a_prey = torch.zeros(self.num_envs, self.num_actions_prey, device=self.device, requires_grad=False)
        obs_pred, obs_prey, priv_pred, priv_prey, _, _, _, _ = self.step(a_pred, a_prey)
        return obs_pred, obs_prey, priv_pred, priv_prey

    def get_observations_pred(self):
        return self.obs_buf_pred

    def get_observations_prey(self):
        return self.obs_buf_prey

    def get_privileged_observations_pred(self):
        return self.privileged_obs_buf_pred

    def get_privileged_observations_prey(self):
        return self.privileged_obs_buf_prey
