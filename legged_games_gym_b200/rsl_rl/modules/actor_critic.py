"""ActorCritic with the rsl_rl API and state_dict layout (actor.{0,2,4,6}.*, critic.{0,2,4,6}.*, std).
Rollout-time calls (no autograd) run the fused lgk_policy_act kernel; calls that need gradients (PPO.update) go through
the nn.Sequential modules so autograd sees ordinary Linear/ELU ops.

The kernel covers what the reference trains (ELU, three hidden layers, actor and critic of equal widths: LRC:203-208 and
every task cfg).  rsl_rl's other options -- selu / relu / crelu / lrelu / tanh / sigmoid, any depth, different actor and
critic widths -- build the same modules and run every call through them (torch on the module's device, torch's sampler):
same API and state_dict, no fused rollout kernel."""
import ctypes as C
import itertools

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions import Normal

from ... import _native as nat


_INSTANCE_IDS = itertools.count(1)      # process-wide: no two modules ever present the same weights_version


def get_activation(act_name):
    """rsl_rl's activation table (modules/actor_critic.py of rsl_rl v1.0.2)."""
    table = {"elu": nn.ELU, "selu": nn.SELU, "relu": nn.ReLU, "lrelu": nn.LeakyReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid}
    if act_name == "crelu":
        raise NotImplementedError("crelu doubles the layer width (rsl_rl maps it to nn.ReLU as well): use relu")
    if act_name not in table:
        raise ValueError(f"invalid activation function {act_name!r}")
    return table[act_name]


def _mlp(n_in, hidden, n_out, act):
    layers = [nn.Linear(n_in, hidden[0]), act()]
    for i in range(len(hidden)):
        if i == len(hidden) - 1:
            layers.append(nn.Linear(hidden[i], n_out))
        else:
            layers += [nn.Linear(hidden[i], hidden[i + 1]), act()]
    return nn.Sequential(*layers)


_ONES = {}


def _ones(n, like):
    key = (n, like.device, like.dtype)
    t = _ONES.get(key)
    if t is None:
        t = _ONES[key] = torch.ones(n, device=like.device, dtype=like.dtype)
    return t


class _Linear(torch.autograd.Function):
    """F.linear with a backward shaped for PPO's mini-batches of 24 576 rows (profiles/update_profile.py): the bias gradient
    is a matrix-vector product with a vector of ones (cuBLAS gemv, one pass over dY at memory speed; autograd's generic
    column reduction took 28 us per layer = 4.5 ms of a 30 ms update) and the weight gradient is split over the batch."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return F.linear(x, w, b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy2, x2 = dy.reshape(-1, dy.shape[-1]), x.reshape(-1, x.shape[-1])
        dx = (dy2 @ w).view_as(x) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            # dW = dY^T X has a tiny output ([512 x 240] at most) and the whole batch as its inner dimension: one GEMM fills
            # 16 CTAs (cuBLAS does not split K here: 56 us per layer).  Eight batch slices as one bmm fill the machine.
            m = dy2.shape[0]
            if m >= 4096 and m % 8 == 0 and dy2.is_contiguous() and x2.is_contiguous():
                dw = torch.bmm(dy2.view(8, m // 8, -1).transpose(1, 2), x2.view(8, m // 8, -1)).sum(0)
            else:
                dw = dy2.t() @ x2
        db = torch.mv(dy2.t(), _ones(dy2.shape[0], dy2)) if ctx.needs_input_grad[2] else None
        return dx, dw, db


class ActorCritic(nn.Module):
    is_recurrent = False

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=[256, 256, 256],
                 critic_hidden_dims=[256, 256, 256], activation="elu", init_noise_std=1.0, **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str(list(kwargs)))
        super().__init__()
        act = get_activation(activation)
        self.num_actor_obs, self.num_critic_obs, self.num_actions = num_actor_obs, num_critic_obs, num_actions
        self.hidden = list(actor_hidden_dims)
        # what lgk_policy_act implements; anything else runs through the torch modules (see the module docstring)
        self.fusable = activation == "elu" and len(self.hidden) == 3 and list(critic_hidden_dims) == self.hidden
        self.actor = _mlp(num_actor_obs, self.hidden, num_actions, act)
        self.critic = _mlp(num_critic_obs, list(critic_hidden_dims), 1, act)
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))
        self.distribution = None
        Normal.set_default_validate_args = False
        self._seed, self._step, self._env_offset = 0, 0, 0
        self._fused = None            # outputs of the last fused call
        self._ws = None
        self._params, self._params_dev = None, None      # cached LgkPolicyParams (pointers to the parameters, workspace)
        # generation of the weight VALUES: the packed TF32 image the kernel keeps in the workspace is rebuilt when it moves.
        # torch's per-tensor version counters see eager in-place updates (optimizer.step, load_state_dict, .to()) but NOT the
        # replay of a captured update graph, so whoever changes the parameters out of torch's sight calls weights_changed().
        self._weights_gen = 0
        self._instance_id = next(_INSTANCE_IDS)

    def weights_changed(self):
        """Tell the rollout kernel that the parameter values changed (PPO calls this after every update())."""
        self._weights_gen += 1

    def _apply(self, fn, *a, **kw):
        self._weights_gen += 1
        self._ws = None
        self._params = None
        self.__dict__.pop("_plist", None)
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        self._weights_gen += 1
        return super().load_state_dict(*a, **kw)

    def _weights_version(self):
        # instance id (bits 40..62) | generation (bits 20..39) | sum of the tensors' version counters (bits 0..19)
        plist = self.__dict__.get("_plist")
        if plist is None:          # nn.Module.parameters() walks the module tree: ~25 us per call at rollout rate
            plist = self.__dict__["_plist"] = list(self.parameters())
        tv = 0
        for q in plist:
            tv += q._version
        return (self._instance_id << 40) | ((self._weights_gen & 0xFFFFF) << 20) | (tv & 0xFFFFF)

    # -- RNG of the sampling epilogue (Philox ACT stream); the runner bumps `step` once per env step
    def set_rng(self, seed, step, env_id_offset=0):
        d = self.__dict__
        d["_seed"], d["_step"], d["_env_offset"] = int(seed), int(step), int(env_id_offset)

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

    @staticmethod
    def _forward_padded(seq, x):
        """seq(x) for an input that may carry zero columns beyond the first layer's width.  PPO's captured update feeds
        mini-batches whose rows are padded to a multiple of 8 floats (235 -> 240): with a 16-byte aligned leading dimension
        cuBLAS runs the first layer's forward and weight-gradient GEMMs on its sm_100 tensor-op kernels instead of the
        unaligned legacy path (4x slower at [24 576 x 235]); the weight is zero-padded to match, so the result is the same sum."""
        h = x
        for i, m in enumerate(seq):
            if isinstance(m, nn.Linear):
                w = m.weight
                if i == 0 and h.shape[-1] != w.shape[1]:
                    w = F.pad(w, (0, h.shape[-1] - w.shape[1]))
                h = _Linear.apply(h, w, m.bias) if torch.is_grad_enabled() else F.linear(h, w, m.bias)
            else:
                h = m(h)
        return h

    def update_distribution(self, observations):
        mean = self._forward_padded(self.actor, observations)
        # validate_args=False: no host-synchronising range checks (rsl_rl disables validation too), CUDA-graph capturable
        self.distribution = Normal(mean, mean * 0. + self.std, validate_args=False)
        self._fused = None

    def _run_fused(self, obs, critic_obs, sample, out=None):
        """critic_obs=None runs the actor alone (act / act_inference): nothing is read through the critic pointer.
        `out` (optional): dict of preallocated contiguous fp32 outputs -- actions / mean / sigma [n, A], values [n, 1],
        logp [n] -- e.g. views of a rollout-storage slot, so that the kernel writes a transition where it is kept."""
        if not obs.is_cuda:
            raise RuntimeError("lgk_policy_act runs on CUDA tensors (there is no CPU path for the rollout kernel)")
        n = obs.shape[0]
        dev = obs.device
        obs = obs.contiguous()
        if obs.shape[1] != self.num_actor_obs:
            raise ValueError(f"observations have {obs.shape[1]} columns, the actor takes {self.num_actor_obs}")
        p = self._params
        if p is None or p.num_envs != n or self._params_dev != dev:
            # everything that does not change from call to call is filled once (the module's parameters keep their storage:
            # optimizer steps and load_state_dict write in place; .to() / _apply drop this cache)
            p = nat.PolicyParams()
            p.num_envs, p.num_obs, p.num_critic_obs, p.num_actions = n, self.num_actor_obs, self.num_critic_obs, self.num_actions
            p.hidden[:] = self.hidden
            for i, li in enumerate((0, 2, 4, 6)):
                p.actor_w[i], p.actor_b[i] = self.actor[li].weight.data_ptr(), self.actor[li].bias.data_ptr()
                p.critic_w[i], p.critic_b[i] = self.critic[li].weight.data_ptr(), self.critic[li].bias.data_ptr()
            p.std = self.std.data_ptr()
            need = int(nat.lib.lgk_policy_workspace_bytes(C.byref(p)))
            if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            p.workspace, p.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
            self.__dict__["_params"], self.__dict__["_params_dev"] = p, dev
        p.obs = obs.data_ptr()
        if critic_obs is not None:
            critic_obs = critic_obs.contiguous()
            if critic_obs.shape[1] != self.num_critic_obs or critic_obs.shape[0] != n:
                raise ValueError(f"critic observations are {tuple(critic_obs.shape)}, the critic takes [{n}, {self.num_critic_obs}]")
            p.critic_obs, p.nets = critic_obs.data_ptr(), 3
        else:
            p.critic_obs, p.nets = None, 1
        p.seed, p.step, p.env_id_offset, p.sample = self._seed, self._step, self._env_offset, int(sample)
        if out is None:
            A = self.num_actions
            buf = torch.empty(n, 3 * A + 2, device=dev)          # one allocation, five outputs
            flat = buf.view(-1)
            out = dict(actions=flat[:n * A].view(n, A), mean=flat[n * A:2 * n * A].view(n, A),
                       sigma=flat[2 * n * A:3 * n * A].view(n, A), values=flat[3 * n * A:3 * n * A + n].view(n, 1),
                       logp=flat[3 * n * A + n:])
        p.actions, p.action_mean, p.action_sigma = out["actions"].data_ptr(), out["mean"].data_ptr(), out["sigma"].data_ptr()
        p.values, p.actions_log_prob = out["values"].data_ptr(), out["logp"].data_ptr()
        p.weights_version = self._weights_version()
        nat.check(nat.lib.lgk_policy_act(C.byref(p), torch.cuda.current_stream().cuda_stream), "lgk_policy_act")
        d = self.__dict__           # plain attributes: skip nn.Module.__setattr__'s parameter / buffer / module checks
        d["_last_params"], d["_last_inputs"] = p, (obs, critic_obs)      # keeps the launch's buffers alive
        d["_fused"] = out
        return out

    def _run_modules(self, obs, critic_obs, out=None):
        """PPO.act for the configurations the kernel does not cover: torch modules, torch's Normal sampler."""
        self.update_distribution(obs)
        actions = self.distribution.sample()
        res = dict(actions=actions, mean=self.distribution.mean, sigma=self.distribution.stddev,
                   values=self.critic(critic_obs), logp=self.distribution.log_prob(actions).sum(dim=-1))
        if out is not None:
            for k, v in res.items():
                out[k].copy_(v.view_as(out[k]))
            res = out
        self._fused = res
        return res

    def act_and_evaluate(self, observations, critic_observations, out=None):
        """PPO.act in one launch set: actions, values, log-prob, mean, sigma (rollout time, no autograd)."""
        with torch.no_grad():
            if not self.fusable:
                return self._run_modules(observations, critic_observations, out)
            return self._run_fused(observations, critic_observations, sample=True, out=out)

    def act(self, observations, **kwargs):
        if (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and not torch.is_inference_mode_enabled()) \
                or not self.fusable:
            self.update_distribution(observations)
            return self.distribution.sample()
        return self._run_fused(observations, None, sample=True)["actions"]

    def get_actions_log_prob(self, actions):
        if self._fused is not None:
            return self._fused["logp"]
        return self.distribution.log_prob(actions).sum(dim=-1)

    def act_inference(self, observations):
        with torch.no_grad():
            if not self.fusable:
                return self.actor(observations)
            return self._run_fused(observations, None, sample=False)["mean"]

    def evaluate(self, critic_observations, **kwargs):
        # always evaluated: env observation buffers are persistent and rewritten in place, so a cache keyed by the tensor's
        # address would hand back the values of the PREVIOUS observation
        return self._forward_padded(self.critic, critic_observations)
