"""OnPolicyRunner with the rsl_rl v1.0.2 API (constructor, learn, save, load, get_inference_policy)."""
import os
import statistics
import time
from collections import deque

import torch
import torch.distributed as dist

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
        ep_infos, rewbuffer, lenbuffer = [], deque(maxlen=100), deque(maxlen=100)
        cur_rew = torch.zeros(env.num_envs, dtype=torch.float, device=self.device)
        cur_len = torch.zeros(env.num_envs, dtype=torch.float, device=self.device)
        tot_iter = self.current_learning_iteration + num_learning_iterations
        act_step = 0
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
                    if self.log_dir is not None or (dist.is_available() and dist.is_initialized()):
                        if "episode" in infos:
                            ep_infos.append(infos["episode"])
                        cur_rew += rewards
                        cur_len += 1
                        new_ids = (dones > 0).nonzero(as_tuple=False)
                        rewbuffer.extend(cur_rew[new_ids][:, 0].cpu().numpy().tolist())
                        lenbuffer.extend(cur_len[new_ids][:, 0].cpu().numpy().tolist())
                        cur_rew[new_ids] = 0
                        cur_len[new_ids] = 0
                stop = time.time()
                self.collection_time = stop - start
                start = stop
                alg.compute_returns(critic_obs)
            mean_value_loss, mean_surrogate_loss = alg.update()
            self.learn_time = time.time() - start
            self.tot_timesteps += self.num_steps_per_env * env.num_envs
            self.tot_time += self.collection_time + self.learn_time
            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
            self.tot_timesteps += (world - 1) * self.num_steps_per_env * env.num_envs
            if world > 1:
                # episode statistics over ALL shards: (sum of finished-episode returns, lengths, count) all-reduced
                stats = torch.tensor([sum(rewbuffer), sum(lenbuffer), float(len(rewbuffer))], device=self.device,
                                     dtype=torch.float64)
                dist.all_reduce(stats, op=dist.ReduceOp.SUM)
                self.global_episode_stats = dict(mean_reward=(stats[0] / stats[2]).item() if stats[2] > 0 else float("nan"),
                                                 mean_length=(stats[1] / stats[2]).item() if stats[2] > 0 else float("nan"),
                                                 episodes=int(stats[2].item()))
            if self.log_dir is not None:
                fps = int(world * self.num_steps_per_env * env.num_envs / (self.collection_time + self.learn_time))
                mr = statistics.mean(rewbuffer) if len(rewbuffer) else float("nan")
                if world > 1:
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
