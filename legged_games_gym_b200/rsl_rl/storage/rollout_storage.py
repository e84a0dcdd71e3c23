"""RolloutStorage with the rsl_rl layout ([T, N, *] tensors); compute_returns runs the lgk_gae kernels."""
import torch

from ... import _native as nat


class RolloutStorage:
    class Transition:
        def __init__(self):
            self.observations = self.critic_observations = self.actions = self.rewards = self.dones = None
            self.values = self.actions_log_prob = self.action_mean = self.action_sigma = None
            self.hidden_states = None
            self.in_slot = False          # PPO.act wrote observations / policy outputs straight into the storage slot

        def clear(self):
            self.__init__()

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape, device="cpu"):
        self.device = device
        self.obs_shape, self.privileged_obs_shape, self.actions_shape = obs_shape, privileged_obs_shape, actions_shape
        T, N = num_transitions_per_env, num_envs
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=device)
        self.observations = z(T, N, *obs_shape)
        self.privileged_observations = z(T, N, *privileged_obs_shape) if privileged_obs_shape[0] is not None else None
        self.rewards, self.actions = z(T, N, 1), z(T, N, *actions_shape)
        self.dones = z(T, N, 1, dt=torch.uint8)
        self.actions_log_prob, self.values, self.returns, self.advantages = z(T, N, 1), z(T, N, 1), z(T, N, 1), z(T, N, 1)
        self.mu, self.sigma = z(T, N, *actions_shape), z(T, N, *actions_shape)
        self.num_transitions_per_env, self.num_envs = T, N
        self.step = 0
        self._gae_scratch = torch.zeros(4, dtype=torch.float64, device=device)

    def add_transitions(self, t):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s = self.step
        if t.observations.data_ptr() != self.observations[s].data_ptr():        # PPO.act already snapshots into the slot
            self.observations[s].copy_(t.observations)
        if self.privileged_observations is not None and t.critic_observations.data_ptr() != self.privileged_observations[s].data_ptr():
            self.privileged_observations[s].copy_(t.critic_observations)
        self.rewards[s].copy_(t.rewards.view(-1, 1))
        self.dones[s].copy_(t.dones.view(-1, 1))
        for dst, src in ((self.actions[s], t.actions), (self.values[s], t.values),
                         (self.actions_log_prob[s], t.actions_log_prob.view(-1, 1)), (self.mu[s], t.action_mean),
                         (self.sigma[s], t.action_sigma)):
            if src.data_ptr() != dst.data_ptr():        # PPO.act lets the policy kernel write the slot directly
                dst.copy_(src)
        self.step += 1

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam):
        T, N = self.num_transitions_per_env, self.num_envs
        lv = last_values.contiguous()
        nat.check(nat.lib.lgk_gae(self.rewards.data_ptr(), self.values.data_ptr(), self.dones.data_ptr(), lv.data_ptr(),
                                  T, N, gamma, lam, self.returns.data_ptr(), self.advantages.data_ptr(),
                                  self._gae_scratch.data_ptr(), torch.cuda.current_stream().cuda_stream), "lgk_gae")

    def get_statistics(self):
        done = self.dones.clone()
        done[-1] = 1
        flat = done.permute(1, 0, 2).reshape(-1, 1)
        idx = torch.cat((flat.new_tensor([-1], dtype=torch.int64), flat.nonzero(as_tuple=False)[:, 0]))
        lens = idx[1:] - idx[:-1]
        return lens.float().mean(), self.rewards.mean()

    def mini_batch_generator(self, num_mini_batches, num_epochs=8):
        B = self.num_envs * self.num_transitions_per_env
        mb = B // num_mini_batches
        idx = torch.randperm(num_mini_batches * mb, requires_grad=False, device=self.device)
        obs = self.observations.flatten(0, 1)
        cobs = self.privileged_observations.flatten(0, 1) if self.privileged_observations is not None else obs
        acts, vals, rets = self.actions.flatten(0, 1), self.values.flatten(0, 1), self.returns.flatten(0, 1)
        olp, adv = self.actions_log_prob.flatten(0, 1), self.advantages.flatten(0, 1)
        mu, sg = self.mu.flatten(0, 1), self.sigma.flatten(0, 1)
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                b = idx[i * mb:(i + 1) * mb]
                yield obs[b], cobs[b], acts[b], vals[b], adv[b], rets[b], olp[b], mu[b], sg[b], (None, None), None
