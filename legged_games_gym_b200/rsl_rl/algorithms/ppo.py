"""PPO with the rsl_rl v1.0.2 API.  act / process_env_step / compute_returns feed the CUDA hot path; update() is plain
PyTorch autograd.  Multi-GPU (one process per GPU, envs sharded): gradients are flattened into one bucket and
all-reduced over NCCL once per mini-batch; the KL estimate that drives the adaptive learning rate is all-reduced so
every rank takes the same schedule."""
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.optim as optim

from ..storage import RolloutStorage


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class PPO:
    def __init__(self, actor_critic, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998, lam=0.95,
                 value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, max_grad_norm=1.0,
                 use_clipped_value_loss=True, schedule="fixed", desired_kl=0.01, device="cpu"):
        self.device = device
        self.desired_kl, self.schedule, self.learning_rate = desired_kl, schedule, learning_rate
        self.actor_critic = actor_critic.to(device)
        self.storage = None
        self.optimizer = optim.Adam(self.actor_critic.parameters(), lr=learning_rate)
        self.transition = RolloutStorage.Transition()
        self.clip_param, self.num_learning_epochs, self.num_mini_batches = clip_param, num_learning_epochs, num_mini_batches
        self.value_loss_coef, self.entropy_coef = value_loss_coef, entropy_coef
        self.gamma, self.lam, self.max_grad_norm = gamma, lam, max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self._flat_grad = None
        self.allreduce_calls = 0

    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        self.storage = RolloutStorage(num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape, self.device)

    def test_mode(self):
        self.actor_critic.eval()

    def train_mode(self):
        self.actor_critic.train()

    def act(self, obs, critic_obs):
        out = self.actor_critic.act_and_evaluate(obs, critic_obs)
        t = self.transition
        t.actions, t.values, t.actions_log_prob = out["actions"], out["values"], out["logp"]
        t.action_mean, t.action_sigma = out["mean"], out["sigma"]
        t.observations, t.critic_observations = obs, critic_obs
        return t.actions

    def process_env_step(self, rewards, dones, infos):
        t = self.transition
        t.rewards = rewards.clone()
        t.dones = dones
        if "time_outs" in infos:          # bootstrap on time-outs
            t.rewards += self.gamma * torch.squeeze(t.values * infos["time_outs"].unsqueeze(1).to(self.device), 1)
        self.storage.add_transitions(t)
        t.clear()
        self.actor_critic.reset(dones)

    def compute_returns(self, last_critic_obs):
        with torch.no_grad():
            last_values = self.actor_critic.critic(last_critic_obs)
        self.storage.compute_returns(last_values, self.gamma, self.lam)

    def _allreduce_grads(self):
        ws = _world()
        if ws == 1:
            return
        params = [p for p in self.actor_critic.parameters() if p.grad is not None]
        n = sum(p.grad.numel() for p in params)
        if self._flat_grad is None or self._flat_grad.numel() != n:
            self._flat_grad = torch.empty(n, device=self.device)
        torch.cat([p.grad.reshape(-1) for p in params], out=self._flat_grad)
        dist.all_reduce(self._flat_grad, op=dist.ReduceOp.SUM)
        self._flat_grad.div_(ws)
        off = 0
        for p in params:
            k = p.grad.numel()
            p.grad.copy_(self._flat_grad[off:off + k].view_as(p.grad))
            off += k
        self.allreduce_calls += 1

    def update(self):
        mean_value_loss = mean_surrogate_loss = 0.0
        ac = self.actor_critic
        gen = self.storage.mini_batch_generator(self.num_mini_batches, self.num_learning_epochs)
        for obs, cobs, acts, tvals, adv, rets, old_lp, old_mu, old_sigma, _, _ in gen:
            ac.update_distribution(obs)
            lp = ac.distribution.log_prob(acts).sum(dim=-1)
            value = ac.critic(cobs)
            mu, sigma, entropy = ac.distribution.mean, ac.distribution.stddev, ac.entropy
            if self.desired_kl is not None and self.schedule == "adaptive":
                with torch.inference_mode():
                    kl = torch.sum(torch.log(sigma / old_sigma + 1.e-5) +
                                   (torch.square(old_sigma) + torch.square(old_mu - mu)) / (2.0 * torch.square(sigma)) - 0.5, axis=-1)
                    kl_mean = torch.mean(kl)
                    if _world() > 1:
                        dist.all_reduce(kl_mean, op=dist.ReduceOp.SUM)
                        kl_mean /= _world()
                    if kl_mean > self.desired_kl * 2.0:
                        self.learning_rate = max(1e-5, self.learning_rate / 1.5)
                    elif kl_mean < self.desired_kl / 2.0 and kl_mean > 0.0:
                        self.learning_rate = min(1e-2, self.learning_rate * 1.5)
                    for g in self.optimizer.param_groups:
                        g["lr"] = self.learning_rate
            ratio = torch.exp(lp - torch.squeeze(old_lp))
            a = torch.squeeze(adv)
            surrogate_loss = torch.max(-a * ratio, -a * torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param)).mean()
            if self.use_clipped_value_loss:
                vc = tvals + (value - tvals).clamp(-self.clip_param, self.clip_param)
                value_loss = torch.max((value - rets).pow(2), (vc - rets).pow(2)).mean()
            else:
                value_loss = (rets - value).pow(2).mean()
            loss = surrogate_loss + self.value_loss_coef * value_loss - self.entropy_coef * entropy.mean()
            self.optimizer.zero_grad()
            loss.backward()
            self._allreduce_grads()
            nn.utils.clip_grad_norm_(ac.parameters(), self.max_grad_norm)
            self.optimizer.step()
            mean_value_loss += value_loss.item()
            mean_surrogate_loss += surrogate_loss.item()
        n = self.num_learning_epochs * self.num_mini_batches
        self.storage.clear()
        return mean_value_loss / n, mean_surrogate_loss / n
