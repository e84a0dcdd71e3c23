"""OnPolicyRunner with the rsl_rl v1.0.2 API (constructor, learn, save, load, get_inference_policy)."""
import os
import time

import torch
import torch.distributed as dist

from ... import _native as nat
from ..algorithms import PPO
from ..modules import ActorCritic


class OnPolicyRunner:
    def __init__(self, env, train_cfg, log_dir=None, device="cpu"):
        self.cfg, self.alg_cfg, self.policy_cfg = train_cfg["runner"], train_cfg["algorithm"], train_cfg["policy"]
        self.device, self.env = device, env
        num_critic_obs = env.num_privileged_obs if env.num_privileged_obs is not None else env.num_obs
        ac = ActorCritic(env.num_obs, num_critic_obs, env.num_actions, **self.policy_cfg).to(device)
        self.alg = PPO(ac, device=device, **self.alg_cfg)
        self.num_steps_per_env, self.save_interval = self.cfg["num_steps_per_env"], self.cfg["save_interval"]
        self.alg.init_storage(env.num_envs, self.num_steps_per_env, [env.num_obs], [env.num_privileged_obs], [env.num_actions])
        self.log_dir, self.writer = log_dir, None
        self.tot_timesteps, self.tot_time, self.current_learning_iteration = 0, 0, 0
        self.seed = int(train_cfg.get("seed", 1))
        self.collection_time = self.learn_time = 0.0
        _, _ = self.env.reset()

    def learn(self, num_learning_iterations, init_at_random_ep_len=False):
        env, alg = self.env, self.alg
        if init_at_random_ep_len:
            # in place: the env's kernels (and a step graph that may already be captured) hold this buffer's address
            env.episode_length_buf.copy_(torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length)))
        obs = env.get_observations()
        pobs = env.get_privileged_observations()
        critic_obs = pobs if pobs is not None else obs
        obs, critic_obs = obs.to(self.device), critic_obs.to(self.device)
        alg.actor_critic.train()
        ep_infos = []
        # running return / length of the episode in progress, per env: rows of one [2, N] tensor so that a step updates and
        # harvests both with a handful of launches (this bookkeeping is host-bound: every op is a launch per rollout step)
        cur = torch.zeros(2, env.num_envs, dtype=torch.float, device=self.device)
        step_inc = torch.zeros(2, env.num_envs, dtype=torch.float, device=self.device)
        step_inc[1] = 1.0
        # finished-episode statistics live on the device (sum of returns, sum of lengths, count): no host synchronisation
        # inside the rollout; read (and, multi-GPU, all-reduced) once per iteration
        ep_stats = torch.zeros(3, dtype=torch.float64, device=self.device)
        tot_iter = self.current_learning_iteration + num_learning_iterations
        act_step = 0
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        track = self.log_dir is not None or world > 1
        native_stats = None       # decided at the first tracked step (CUDA fp32 rewards + 1-byte dones: one kernel)
        for it in range(self.current_learning_iteration, tot_iter):
            start = time.time()
            with torch.inference_mode():
                for _ in range(self.num_steps_per_env):
                    act_step += 1
                    alg.actor_critic.set_rng(self.seed, act_step, getattr(env, "env_id_offset", 0))
                    actions = alg.act(obs, critic_obs)
                    obs, pobs, rewards, dones, infos = env.step(actions)
                    critic_obs = pobs if pobs is not None else obs
                    alg.process_env_step(rewards, dones, infos)
                    if track:
                        if native_stats is None:
                            native_stats = bool(rewards.is_cuda and rewards.dtype == torch.float32 and rewards.is_contiguous()
                                                and dones.element_size() == 1 and dones.is_contiguous() and cur.is_cuda)
                        if self.log_dir is not None and "episode" in infos:
                            ep_infos.append(infos["episode"])
                        if native_stats:
                            nat.check(nat.lib.lgk_episode_stats(rewards.data_ptr(), dones.data_ptr(), cur[0].data_ptr(),
                                                                cur[1].data_ptr(), ep_stats.data_ptr(), env.num_envs,
                                                                torch.cuda.current_stream().cuda_stream), "lgk_episode_stats")
                        else:
                            step_inc[0].copy_(rewards)
                            cur += step_inc                                  # return += reward, length += 1
                            done = (dones > 0).to(cur.dtype)
                            ep_stats[:2] += torch.mv(cur, done)              # sums over the episodes that just ended
                            ep_stats[2] += done.sum()
                            cur *= (1.0 - done)
                stop = time.time()
                self.collection_time = stop - start
                start = stop
                alg.compute_returns(critic_obs)
            mean_value_loss, mean_surrogate_loss = alg.update()
            self.learn_time = time.time() - start
            self.tot_timesteps += world * self.num_steps_per_env * env.num_envs
            self.tot_time += self.collection_time + self.learn_time
            if track:
                stats = ep_stats.clone()
                if world > 1:      # episode statistics over ALL shards
                    dist.all_reduce(stats, op=dist.ReduceOp.SUM)
                s_rew, s_len, s_cnt = stats.tolist()
                self.global_episode_stats = dict(mean_reward=s_rew / s_cnt if s_cnt > 0 else float("nan"),
                                                 mean_length=s_len / s_cnt if s_cnt > 0 else float("nan"), episodes=int(s_cnt))
            if self.log_dir is not None:
                fps = int(world * self.num_steps_per_env * env.num_envs / (self.collection_time + self.learn_time))
                mr = self.global_episode_stats["mean_reward"]
                print(f"it {it}/{tot_iter} steps/s {fps} collection {self.collection_time:.3f}s learning {self.learn_time:.3f}s "
                      f"value_loss {mean_value_loss:.4f} surrogate {mean_surrogate_loss:.4f} mean_reward {mr:.3f}")
                if it % self.save_interval == 0:
                    self.save(os.path.join(self.log_dir, "model_{}.pt".format(it)))
            ep_infos.clear()
        self.current_learning_iteration += num_learning_iterations
        if self.log_dir is not None:
            self.save(os.path.join(self.log_dir, "model_{}.pt".format(self.current_learning_iteration)))
        return mean_value_loss, mean_surrogate_loss

    def save(self, path, infos=None):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save({"model_state_dict": self.alg.actor_critic.state_dict(),
                    "optimizer_state_dict": self.alg.optimizer_state_dict(),
                    "iter": self.current_learning_iteration, "infos": infos}, path)

    def load(self, path, load_optimizer=True):
        d = torch.load(path, map_location=self.device)
        self.alg.actor_critic.load_state_dict(d["model_state_dict"])
        if load_optimizer:
            self.alg.load_optimizer_state_dict(d["optimizer_state_dict"])
        self.current_learning_iteration = d["iter"]
        return d["infos"]

    def get_inference_policy(self, device=None):
        self.alg.actor_critic.eval()
        if device is not None:
            self.alg.actor_critic.to(device)
        return self.alg.actor_critic.act_inference
