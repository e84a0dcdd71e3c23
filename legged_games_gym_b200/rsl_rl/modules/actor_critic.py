"""ActorCritic with the rsl_rl API and state_dict layout (actor.{0,2,4,6}.*, critic.{0,2,4,6}.*, std).
Rollout-time calls (no autograd) run the fused lgk_policy_act kernel; calls that need gradients (PPO.update) go through
the nn.Sequential modules so autograd sees ordinary Linear/ELU ops."""
import ctypes as C

import torch
import torch.nn as nn
from torch.distributions import Normal

from ... import _native as nat


def _mlp(n_in, hidden, n_out, act):
    layers = [nn.Linear(n_in, hidden[0]), act()]
    for i in range(len(hidden)):
        if i == len(hidden) - 1:
            layers.append(nn.Linear(hidden[i], n_out))
        else:
            layers += [nn.Linear(hidden[i], hidden[i + 1]), act()]
    return nn.Sequential(*layers)


class ActorCritic(nn.Module):
    is_recurrent = False

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=[256, 256, 256],
                 critic_hidden_dims=[256, 256, 256], activation="elu", init_noise_std=1.0, **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str(list(kwargs)))
        super().__init__()
        if activation != "elu":
            raise NotImplementedError("the fused policy kernel implements ELU (the reference's cfg, LRC:208)")
        self.num_actor_obs, self.num_critic_obs, self.num_actions = num_actor_obs, num_critic_obs, num_actions
        self.hidden = list(actor_hidden_dims)
        if list(critic_hidden_dims) != self.hidden or len(self.hidden) != 3:
            raise NotImplementedError("fused kernel needs three hidden layers, identical for actor and critic")
        self.actor = _mlp(num_actor_obs, self.hidden, num_actions, nn.ELU)
        self.critic = _mlp(num_critic_obs, self.hidden, 1, nn.ELU)
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))
        self.distribution = None
        Normal.set_default_validate_args = False
        self._seed, self._step, self._env_offset = 0, 0, 0
        self._fused = None            # outputs of the last fused call
        self._ws = None

    # -- RNG of the sampling epilogue (Philox ACT stream); the runner bumps `step` once per env step
    def set_rng(self, seed, step, env_id_offset=0):
        self._seed, self._step, self._env_offset = int(seed), int(step), int(env_id_offset)

    def reset(self, dones=None):
        pass

    def forward(self):
        raise NotImplementedError

    @property
    def action_mean(self):
        return self._fused["mean"] if self._fused is not None else self.distribution.mean

    @property
    def action_std(self):
        return self._fused["sigma"] if self._fused is not None else self.distribution.stddev

    @property
    def entropy(self):
        return self.distribution.entropy().sum(dim=-1)

    def update_distribution(self, observations):
        mean = self.actor(observations)
        # validate_args=False: no host-synchronising range checks (rsl_rl disables validation too), CUDA-graph capturable
        self.distribution = Normal(mean, mean * 0. + self.std, validate_args=False)
        self._fused = None

    def _run_fused(self, obs, critic_obs, sample):
        n = obs.shape[0]
        dev = obs.device
        p = nat.PolicyParams()
        p.num_envs, p.num_obs, p.num_critic_obs, p.num_actions = n, self.num_actor_obs, self.num_critic_obs, self.num_actions
        p.hidden[:] = self.hidden
        obs = obs.contiguous()
        critic_obs = critic_obs.contiguous()
        p.obs, p.critic_obs = obs.data_ptr(), critic_obs.data_ptr()
        for i, li in enumerate((0, 2, 4, 6)):
            p.actor_w[i], p.actor_b[i] = self.actor[li].weight.data_ptr(), self.actor[li].bias.data_ptr()
            p.critic_w[i], p.critic_b[i] = self.critic[li].weight.data_ptr(), self.critic[li].bias.data_ptr()
        p.std = self.std.data_ptr()
        p.seed, p.step, p.env_id_offset, p.sample = self._seed, self._step, self._env_offset, int(sample)
        out = dict(actions=torch.empty(n, self.num_actions, device=dev), mean=torch.empty(n, self.num_actions, device=dev),
                   sigma=torch.empty(n, self.num_actions, device=dev), values=torch.empty(n, 1, device=dev),
                   logp=torch.empty(n, device=dev), obs_ptr=critic_obs.data_ptr())
        p.actions, p.action_mean, p.action_sigma = out["actions"].data_ptr(), out["mean"].data_ptr(), out["sigma"].data_ptr()
        p.values, p.actions_log_prob = out["values"].data_ptr(), out["logp"].data_ptr()
        need = int(nat.lib.lgk_policy_workspace_bytes(C.byref(p)))
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        p.workspace, p.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        # torch bumps a tensor's version counter on every in-place update (optimizer.step, load_state_dict): the packed
        # TF32 weight image inside the workspace is rebuilt only when this sum moves
        p.weights_version = 1 + sum(int(q._version) for q in self.parameters())
        nat.check(nat.lib.lgk_policy_act(C.byref(p), torch.cuda.current_stream().cuda_stream), "lgk_policy_act")
        self._last_params, self._last_inputs = p, (obs, critic_obs)      # keeps the launch's buffers alive
        self._fused = out
        return out

    def act_and_evaluate(self, observations, critic_observations):
        """PPO.act in one launch set: actions, values, log-prob, mean, sigma (rollout time, no autograd)."""
        with torch.no_grad():
            return self._run_fused(observations, critic_observations, sample=True)

    def act(self, observations, **kwargs):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and not torch.is_inference_mode_enabled():
            self.update_distribution(observations)
            return self.distribution.sample()
        return self._run_fused(observations, observations, sample=True)["actions"]

    def get_actions_log_prob(self, actions):
        if self._fused is not None:
            return self._fused["logp"]
        return self.distribution.log_prob(actions).sum(dim=-1)

    def act_inference(self, observations):
        with torch.no_grad():
            return self._run_fused(observations, observations, sample=False)["mean"]

    def evaluate(self, critic_observations, **kwargs):
        if self._fused is not None and not torch.is_grad_enabled() and self._fused["obs_ptr"] == critic_observations.data_ptr():
            return self._fused["values"]
        return self.critic(critic_observations)
