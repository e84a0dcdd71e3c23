"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the two rsl_rl functions on the hot path.

rsl_rl is a third-party dependency that is NOT under /root/reference (setup.py:12 names 'rsl-rl' unpinned; README.md:21-23
points at leggedrobotics/rsl_rl, v1.0.2 era; call sites: legged_gym/utils/task_registry.py:37-38, 154, 161).  The
published algorithm of that version is restated here: ActorCritic (modules/actor_critic.py: MLP actor/critic with ELU,
Normal(mean, std) sampling, log_prob summed over actions) and RolloutStorage.compute_returns (storage/rollout_storage.py:
reverse GAE + global advantage normalisation).  PARITY UNPINNED: the reference holds no test or golden vector for these.
"""
import torch
import torch.nn as nn


def build_mlp(n_in, hidden, n_out):
    layers = [nn.Linear(n_in, hidden[0]), nn.ELU()]
    for i in range(len(hidden)):
        if i == len(hidden) - 1:
            layers.append(nn.Linear(hidden[i], n_out))
        else:
            layers += [nn.Linear(hidden[i], hidden[i + 1]), nn.ELU()]
    return nn.Sequential(*layers)


class ActorCriticOracle(nn.Module):
    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=(512, 256, 128),
                 critic_hidden_dims=(512, 256, 128), init_noise_std=1.0):
        super().__init__()
        self.actor = build_mlp(num_actor_obs, list(actor_hidden_dims), num_actions)
        self.critic = build_mlp(num_critic_obs, list(critic_hidden_dims), 1)
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))

    @torch.no_grad()
    def act(self, obs, critic_obs, eps):
        """PPO.act with the N(0,1) draws `eps` explicit: actions = mean + std*eps (what Normal.sample() computes)."""
        mean = self.actor(obs)
        sigma = mean * 0. + self.std
        actions = mean + sigma * eps
        dist = torch.distributions.Normal(mean, sigma)
        logp = dist.log_prob(actions).sum(dim=-1)
        values = self.critic(critic_obs)
        return actions, values, logp, mean, sigma


def compute_returns(rewards, values, dones, last_values, gamma, lam):
    """rewards/values [T,N,1] fp32, dones [T,N,1] uint8, last_values [N,1] -> returns, advantages [T,N,1]."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    advantage = 0
    for step in reversed(range(T)):
        next_values = last_values if step == T - 1 else values[step + 1]
        next_is_not_terminal = 1.0 - dones[step].float()
        delta = rewards[step] + next_is_not_terminal * gamma * next_values - values[step]
        advantage = delta + next_is_not_terminal * gamma * lam * advantage
        returns[step] = advantage + values[step]
    advantages = returns - values
    advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    return returns, advantages
