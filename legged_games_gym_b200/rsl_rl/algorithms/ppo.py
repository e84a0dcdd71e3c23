"""PPO with the rsl_rl v1.0.2 API.  act / process_env_step / compute_returns feed the CUDA hot path; update() is plain
PyTorch autograd -- on a single GPU captured once into a CUDA graph (one replay per mini-batch: gathers, forward, losses,
backward, gradient clipping, Adam, the adaptive-KL learning-rate rule and the loss accumulators all on the device, no
host synchronisation inside the loop), eager otherwise.  Multi-GPU (one process per GPU, envs sharded): the parameters'
.grad tensors are views of ONE flat bucket, which is all-reduced over NCCL once per mini-batch -- inside the captured graph
on CUDA (NCCL collectives are capturable), so N > 1 keeps the one-replay-per-mini-batch structure; the KL estimate that
drives the adaptive learning rate is all-reduced on the device too, so every rank takes the same schedule."""
import copy

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
                 use_clipped_value_loss=True, schedule="fixed", desired_kl=0.01, device="cpu", tf32_matmul=None):
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
        # The update's matmuls run on TF32 tensor cores by default: that is what the reference's stack did (its README pins
        # PyTorch 1.10, where torch.backends.cuda.matmul.allow_tf32 defaulted to True; 1.12 flipped the default), and the
        # rollout-time policy kernel is TF32 as well.  tf32_matmul=False or LGK_PPO_TF32=0 keeps cuBLAS in strict fp32.
        import os
        if tf32_matmul is None:
            tf32_matmul = os.environ.get("LGK_PPO_TF32", "1") != "0"
        self.tf32_matmul = bool(tf32_matmul)
        torch.backends.cuda.matmul.allow_tf32 = self.tf32_matmul
        # single-GPU fast path of update(): whole mini-batch step as one CUDA graph (set to False for the eager loop)
        self.use_cuda_graph = True
        self._graph = None

    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        self.storage = RolloutStorage(num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape, self.device)

    def test_mode(self):
        self.actor_critic.eval()

    def train_mode(self):
        self.actor_critic.train()

    def act(self, obs, critic_obs):
        st = self.storage
        slot = None
        if st is not None and st.step < st.num_transitions_per_env and obs.is_cuda and st.actions.is_cuda:
            # the kernel writes the transition straight into its rollout-storage slot (add_transitions then has nothing to copy)
            s = st.step
            slot = dict(actions=st.actions[s], mean=st.mu[s], sigma=st.sigma[s], values=st.values[s],
                        logp=st.actions_log_prob[s].view(-1))
        out = self.actor_critic.act_and_evaluate(obs, critic_obs, out=slot)
        t = self.transition
        t.in_slot = slot is not None
        t.actions, t.values, t.actions_log_prob = out["actions"], out["values"], out["logp"]
        t.action_mean, t.action_sigma = out["mean"], out["sigma"]
        # Snapshot NOW: `obs` is usually the env's persistent obs_buf, which env.step() rewrites in place before
        # process_env_step() stores the transition (the reference's env rebinds obs_buf to a fresh tensor every step, so
        # rsl_rl can keep the reference there).  The copy lands in its final place, the rollout storage slot.
        if st is not None and st.step < st.num_transitions_per_env:
            t.observations = st.observations[st.step]
            t.observations.copy_(obs)
            if st.privileged_observations is not None:
                t.critic_observations = st.privileged_observations[st.step]
                t.critic_observations.copy_(critic_obs)
            else:
                t.critic_observations = t.observations
        else:
            t.observations, t.critic_observations = obs.clone(), critic_obs.clone()
        return t.actions

    def process_env_step(self, rewards, dones, infos):
        t = self.transition
        st = self.storage
        s = st.step
        if t.in_slot and s < st.num_transitions_per_env:
            # rollout fast path: PPO.act already placed observations and the policy outputs in slot s; rewards (with the
            # time-out bootstrap, rsl_rl PPO.process_env_step) and dones go there directly
            r = st.rewards[s].view(-1)
            if "time_outs" in infos:
                torch.addcmul(rewards, t.values.view(-1), infos["time_outs"].to(device=r.device, dtype=r.dtype), value=self.gamma, out=r)
            else:
                r.copy_(rewards)
            st.dones[s].view(-1).copy_(dones)
            st.step = s + 1
        else:
            t.rewards = rewards.clone()
            t.dones = dones
            if "time_outs" in infos:          # bootstrap on time-outs
                t.rewards += self.gamma * torch.squeeze(t.values * infos["time_outs"].unsqueeze(1).to(self.device), 1)
            st.add_transitions(t)
        t.clear()
        self.actor_critic.reset(dones)

    def compute_returns(self, last_critic_obs):
        with torch.no_grad():
            last_values = self.actor_critic.critic(last_critic_obs)
        self.storage.compute_returns(last_values, self.gamma, self.lam)

    def _bucket_grads(self):
        """Make every parameter's .grad a view of one flat fp32 buffer (autograd accumulates into existing .grad tensors
        in place): the bucket IS the gradients, no gather / scatter around the all-reduce."""
        params = list(self.actor_critic.parameters())
        n = sum(p.numel() for p in params)
        if self._flat_grad is None or self._flat_grad.numel() != n + 1 or self._flat_grad.device != params[0].device:
            self._flat_grad = torch.zeros(n + 1, device=params[0].device, dtype=torch.float32)   # + 1: the KL estimate
        off = 0
        for p in params:
            k = p.numel()
            view = self._flat_grad[off:off + k].view_as(p)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                p.grad = view
            off += k

    def _allreduce_grads(self):
        """mean of the shards' gradients: one NCCL all-reduce of the flat bucket (2.29 MB for the 512-256-128 nets)"""
        ws = _world()
        if ws == 1:
            return
        dist.all_reduce(self._flat_grad, op=dist.ReduceOp.SUM)
        self._flat_grad.div_(ws)
        self.allreduce_calls += 1

    # ------------------------------------------------------------------ graphed update (single process, CUDA)
    def _minibatch_loss(self, obs, cobs, acts, tvals, adv, rets, old_lp, old_mu, old_sigma, lr_t):
        """One mini-batch of rsl_rl PPO.update with the learning-rate rule on the device (lr_t: 0-dim tensor)."""
        ac = self.actor_critic
        ac.update_distribution(obs)
        lp = ac.distribution.log_prob(acts).sum(dim=-1)
        value = ac.evaluate(cobs)
        mu, sigma, entropy = ac.distribution.mean, ac.distribution.stddev, ac.entropy
        kl_mean = None
        if self.desired_kl is not None and self.schedule == "adaptive":
            with torch.no_grad():
                kl = torch.sum(torch.log(sigma / old_sigma + 1.e-5) +
                               (torch.square(old_sigma) + torch.square(old_mu - mu)) / (2.0 * torch.square(sigma)) - 0.5, axis=-1)
                kl_mean = torch.mean(kl)
        ratio = torch.exp(lp - torch.squeeze(old_lp))
        a = torch.squeeze(adv)
        surrogate_loss = torch.max(-a * ratio, -a * torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param)).mean()
        if self.use_clipped_value_loss:
            vc = tvals + (value - tvals).clamp(-self.clip_param, self.clip_param)
            value_loss = torch.max((value - rets).pow(2), (vc - rets).pow(2)).mean()
        else:
            value_loss = (rets - value).pow(2).mean()
        loss = surrogate_loss + self.value_loss_coef * value_loss - self.entropy_coef * entropy.mean()
        return loss, value_loss, surrogate_loss, kl_mean

    def _adapt_lr(self, kl_mean, lr_t):
        """rsl_rl's adaptive-KL rule on the device (lr_t: 0-dim tensor the capturable optimizer reads)"""
        down = torch.clamp(lr_t / 1.5, min=1e-5)
        up = torch.clamp(lr_t * 1.5, max=1e-2)
        new_lr = torch.where(kl_mean > self.desired_kl * 2.0, down,
                             torch.where((kl_mean < self.desired_kl / 2.0) & (kl_mean > 0.0), up, lr_t))
        lr_t.copy_(new_lr)

    def _graph_step(self, g):
        """gather the mini-batch named by g['idx'] from the flat rollout tensors, then one optimisation step"""
        st = g["flat"]
        idx = g["idx"]
        take = lambda t: t.index_select(0, idx)
        obs_mb = take(st["obs"])
        cobs_mb = obs_mb if st["cobs"] is st["obs"] else take(st["cobs"])      # no privileged observations: one gather
        loss, vl, sl, kl = self._minibatch_loss(obs_mb, cobs_mb, take(st["acts"]), take(st["vals"]), take(st["adv"]),
                                                take(st["rets"]), take(st["olp"]), take(st["mu"]), take(st["sg"]), g["lr"])
        self.optimizer.zero_grad(set_to_none=False)
        loss.backward()
        # ONE collective per mini-batch: the KL estimate rides in the last element of the gradient bucket, so every rank
        # takes the same learning-rate decision from the same all-reduce that averages the gradients
        with torch.no_grad():
            if kl is not None:
                self._flat_grad[-1] = kl
            self._allreduce_grads()
            if kl is not None:
                self._adapt_lr(self._flat_grad[-1], g["lr"])
        nn.utils.clip_grad_norm_(self.actor_critic.parameters(), self.max_grad_norm, foreach=True)
        self.optimizer.step()
        g["acc"][0] += vl.detach()
        g["acc"][1] += sl.detach()

    def _build_graph(self):
        st, dev = self.storage, self.device
        B = st.num_envs * st.num_transitions_per_env
        mb = B // self.num_mini_batches
        # observation rows padded with zeros to a multiple of 8 floats (refreshed from the storage at every update): see
        # ActorCritic._forward_padded
        pads = []

        def padded(t):
            k, src = t.shape[-1], t.flatten(0, 1)
            if k % 8 == 0:
                return src
            dst = torch.zeros(B, -(-k // 8) * 8, device=dev, dtype=t.dtype)
            pads.append((dst, src))
            return dst
        obs = padded(st.observations)
        cobs = padded(st.privileged_observations) if st.privileged_observations is not None else obs
        flat = dict(obs=obs, cobs=cobs,
                    acts=st.actions.flatten(0, 1), vals=st.values.flatten(0, 1), rets=st.returns.flatten(0, 1),
                    olp=st.actions_log_prob.flatten(0, 1), adv=st.advantages.flatten(0, 1), mu=st.mu.flatten(0, 1),
                    sg=st.sigma.flatten(0, 1))
        # Adam whose step counter and learning rate live on the device (capturable); state carried over from the eager optimizer
        # deep copy: state_dict() hands out references, and load_state_dict() of same-device tensors aliases them, so the
        # warm-up steps and the zero_() below would otherwise wipe the moments that are restored afterwards (--resume)
        state = copy.deepcopy(self.optimizer.state_dict())
        lr_t = torch.tensor(float(self.learning_rate), device=dev)
        self.optimizer = optim.Adam(self.actor_critic.parameters(), lr=lr_t, capturable=True, fused=True)     # one multi-tensor kernel
        carried = bool(state["state"])
        if carried:
            for s_ in state["state"].values():
                s_["step"] = torch.as_tensor(float(s_["step"]), dtype=torch.float32, device=dev)
            for grp in state["param_groups"]:
                grp["lr"], grp["capturable"], grp["fused"], grp["foreach"] = lr_t, True, True, False
            self.optimizer.load_state_dict(copy.deepcopy(state))       # `state` itself stays pristine for the restore below
        for grp in self.optimizer.param_groups:
            grp["lr"] = lr_t
        g = dict(flat=flat, idx=torch.zeros(mb, dtype=torch.long, device=dev), lr=lr_t, mb=mb,
                 acc=torch.zeros(2, device=dev), pads=pads)
        for dst, src in pads:
            dst[:, :src.shape[1]].copy_(src)
        # an autograd graph left over from an eager update (ActorCritic.distribution holds the actor's output) would keep
        # the parameters' AccumulateGrad nodes alive on the stream they were created on: drop it before warm-up / capture
        self.actor_critic.distribution = None
        for p_ in self.actor_critic.parameters():
            p_.grad = None
        self._bucket_grads()
        # the live weights / optimizer state must not be touched by warm-up and capture: snapshot, run, restore
        snap_p = [p.detach().clone() for p in self.actor_critic.parameters()]
        calls = self.allreduce_calls                     # warm-up / capture launches do not count as update steps
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._graph_step(g)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):      # same stream as the warm-up: the AccumulateGrad nodes stay on it
            self._graph_step(g)
        with torch.no_grad():
            for p, q in zip(self.actor_critic.parameters(), snap_p):
                p.copy_(q)
        # the warm-up steps must not count: restore moments and step IN PLACE (the graph holds these tensors' addresses)
        params = [p_ for grp in self.optimizer.param_groups for p_ in grp["params"]]
        for i, p_ in enumerate(params):
            live, src = self.optimizer.state[p_], state["state"].get(i) if carried else None
            for k in ("step", "exp_avg", "exp_avg_sq"):
                if src is not None:
                    live[k].copy_(src[k])
                else:
                    live[k].zero_()
        for grp in self.optimizer.param_groups:
            grp["lr"] = lr_t
        lr_t.fill_(float(self.learning_rate))
        self.allreduce_calls = calls
        g["graph"] = graph
        self._graph = g

    def release_graph(self):
        """Drop the captured update graph (multi-GPU: it holds NCCL work; release it before destroy_process_group)."""
        self._graph = None

    def _update_graphed(self):
        if self._graph is None:
            self._build_graph()
        g = self._graph
        for dst, src in g["pads"]:
            dst[:, :src.shape[1]].copy_(src)
        g["acc"].zero_()
        g["lr"].fill_(float(self.learning_rate))
        B = g["mb"] * self.num_mini_batches
        perm = torch.randperm(B, requires_grad=False, device=self.device)     # one permutation per update(), like the generator
        for _ in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                g["idx"].copy_(perm[i * g["mb"]:(i + 1) * g["mb"]])
                g["graph"].replay()
        n = self.num_learning_epochs * self.num_mini_batches
        if _world() > 1:
            self.allreduce_calls += n                                       # one bucket all-reduce inside every replay
        acc = (g["acc"] / n).tolist()                                       # the only host synchronisation of the update
        self.learning_rate = float(g["lr"].item())
        self.storage.clear()
        return acc[0], acc[1]

    def update(self):
        out = self._update()
        # the parameters changed (a graph replay moves no torch version counter): the rollout kernel re-packs its weights
        self.actor_critic.weights_changed()
        return out

    # ------------------------------------------------------------------ optimizer state in the reference's format
    def optimizer_state_dict(self):
        """Adam state as stock rsl_rl / torch.optim.Adam writes it (plain-float lr, CPU `step`, no capturable / fused
        flags), whichever optimizer flavour is live: checkpoints stay interchangeable with the reference's."""
        sd = copy.deepcopy(self.optimizer.state_dict())
        for s_ in sd["state"].values():
            if torch.is_tensor(s_.get("step")):
                s_["step"] = s_["step"].detach().float().cpu()
        for grp in sd["param_groups"]:
            grp["lr"] = float(grp["lr"])
            grp["capturable"], grp["fused"], grp["foreach"] = False, None, None
        return sd

    def load_optimizer_state_dict(self, sd):
        """Accepts the reference's format (and this class's own): values are converted to the live optimizer's flavour.
        Once the update graph is captured the state tensors' addresses are baked into it, so values are copied IN PLACE."""
        lr = float(sd["param_groups"][0]["lr"])
        if self._graph is not None and len(self.optimizer.state) > 0:
            params = [p for g in self.optimizer.param_groups for p in g["params"]]
            for i, p in enumerate(params):
                src = sd["state"].get(i)
                if src is None:
                    continue
                dst = self.optimizer.state[p]
                dst["exp_avg"].copy_(src["exp_avg"])
                dst["exp_avg_sq"].copy_(src["exp_avg_sq"])
                dst["step"].fill_(float(src["step"]))
            self._graph["lr"].fill_(lr)
        else:
            sd = copy.deepcopy(sd)
            live = self.optimizer.param_groups[0]
            capt = bool(live.get("capturable", False))
            for s_ in sd["state"].values():
                if "step" in s_:
                    s_["step"] = torch.as_tensor(float(s_["step"]), dtype=torch.float32, device=self.device if capt else "cpu")
            for grp, lg in zip(sd["param_groups"], self.optimizer.param_groups):
                for k in ("capturable", "fused", "foreach"):
                    if k in lg:
                        grp[k] = lg[k]
                grp["lr"] = lg["lr"]
                if torch.is_tensor(lg["lr"]):
                    lg["lr"].fill_(lr)
                else:
                    grp["lr"] = lr
            self.optimizer.load_state_dict(sd)
        self.learning_rate = lr

    def _update(self):
        if (self.use_cuda_graph and torch.device(self.device).type == "cuda" and self.storage.observations.is_cuda
                and (_world() == 1 or dist.get_backend() == "nccl")):
            return self._update_graphed()
        self._bucket_grads()
        mean_value_loss = mean_surrogate_loss = 0.0
        ac = self.actor_critic
        gen = self.storage.mini_batch_generator(self.num_mini_batches, self.num_learning_epochs)
        for obs, cobs, acts, tvals, adv, rets, old_lp, old_mu, old_sigma, _, _ in gen:
            ac.update_distribution(obs)
            lp = ac.distribution.log_prob(acts).sum(dim=-1)
            value = ac.evaluate(cobs)
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
            self.optimizer.zero_grad(set_to_none=False)       # (the .grad tensors are views of the bucket: keep them)
            loss.backward()
            self._allreduce_grads()
            nn.utils.clip_grad_norm_(ac.parameters(), self.max_grad_norm)
            self.optimizer.step()
            mean_value_loss += value_loss.item()
            mean_surrogate_loss += surrogate_loss.item()
        n = self.num_learning_epochs * self.num_mini_batches
        self.storage.clear()
        return mean_value_loss / n, mean_surrogate_loss / n
