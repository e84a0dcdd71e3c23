"""DecGamePolicyRunner -- two PPO agents (0 = predator, 1 = prey) on a DecHighLevelGame env, trained in alternating
"evolutions".  The reference imports this class from a fork of rsl_rl that is not in its tree
(legged_gym/utils/task_registry.py:38); what is written here follows its call sites: the constructor
(task_registry.py:207), ``learn(max_num_evolutions, num_learning_iterations, init_at_random_ep_len)``
(scripts/train_dec_game.py:46), ``load(agent_id, path)`` (task_registry.py:218-220),
``get_inference_policy(agent_id, device)`` (scripts/play_dec_game.py:68-69) and the checkpoint names
``pred_model_<it>.pt`` / ``prey_model_<it>.pt`` (utils/helpers.py:141-155).  PARITY UNPINNED for the schedule itself:
in evolution e the agent e % 2 learns with PPO for ``num_learning_iterations`` iterations while the other acts with its
current mean policy."""
import os
import time

import torch

from ..algorithms import PPO
from ..modules import ActorCritic

AGENTS = ("pred", "prey")


class DecGamePolicyRunner:
    def __init__(self, env, train_cfg, log_dir=None, device="cpu"):
        self.cfg, self.alg_cfg, self.policy_cfg = train_cfg["runner"], train_cfg["algorithm"], train_cfg["policy"]
        self.device, self.env = device, env
        self.num_steps_per_env, self.save_interval = self.cfg["num_steps_per_env"], self.cfg["save_interval"]
        dims = ((env.num_obs_pred, env.num_privileged_obs_pred, env.num_actions_pred),
                (env.num_obs_prey, env.num_privileged_obs_prey, env.num_actions_prey))
        self.algs = []
        for n_obs, n_priv, n_act in dims:
            ac = ActorCritic(n_obs, n_priv if n_priv is not None else n_obs, n_act, **self.policy_cfg).to(device)
            alg = PPO(ac, device=device, **self.alg_cfg)
            alg.init_storage(env.num_envs, self.num_steps_per_env, [n_obs], [n_priv], [n_act])
            self.algs.append(alg)
        self.log_dir = log_dir
        self.seed = int(train_cfg.get("seed", 1))
        self.tot_timesteps, self.tot_time = 0, 0.0
        self.current_learning_iteration = [0, 0]
        self.current_evolution = 0
        self.env.reset()

    def _obs(self):
        env = self.env
        return [env.get_observations_pred().to(self.device), env.get_observations_prey().to(self.device)]

    def learn(self, max_num_evolutions, num_learning_iterations, init_at_random_ep_len=False):
        env = self.env
        if init_at_random_ep_len:
            env.episode_length_buf.copy_(torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length)))
        obs = self._obs()
        losses = None
        act_step = 0
        for evo in range(self.current_evolution, self.current_evolution + max_num_evolutions):
            learner = evo % 2
            alg = self.algs[learner]
            other = self.algs[1 - learner].actor_critic
            alg.actor_critic.train()
            other.eval()
            first = self.current_learning_iteration[learner]
            for it in range(first, first + num_learning_iterations):
                start = time.time()
                with torch.inference_mode():
                    for _ in range(self.num_steps_per_env):
                        act_step += 1
                        alg.actor_critic.set_rng(self.seed + learner, act_step, getattr(env.ll_env, "env_id_offset", 0))
                        acts = [None, None]
                        acts[learner] = alg.act(obs[learner], obs[learner])
                        acts[1 - learner] = other.act_inference(obs[1 - learner])
                        o_pred, o_prey, _, _, r_pred, r_prey, dones, infos = env.step(acts[0], acts[1])
                        obs = [o_pred.to(self.device), o_prey.to(self.device)]
                        alg.process_env_step((r_pred, r_prey)[learner], dones, infos)
                    alg.compute_returns(obs[learner])
                losses = alg.update()
                self.tot_timesteps += self.num_steps_per_env * env.num_envs
                self.tot_time += time.time() - start
                if self.log_dir is not None:
                    print(f"evolution {evo} [{AGENTS[learner]}] it {it} value_loss {losses[0]:.4f} surrogate {losses[1]:.4f}")
                    if it % self.save_interval == 0:
                        self.save(learner, os.path.join(self.log_dir, f"{AGENTS[learner]}_model_{it}.pt"))
            self.current_learning_iteration[learner] += num_learning_iterations
            if self.log_dir is not None:
                self.save(learner, os.path.join(self.log_dir, f"{AGENTS[learner]}_model_{self.current_learning_iteration[learner]}.pt"))
        self.current_evolution += max_num_evolutions
        return losses

    def save(self, agent_id, path, infos=None):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        alg = self.algs[agent_id]
        torch.save({"model_state_dict": alg.actor_critic.state_dict(), "optimizer_state_dict": alg.optimizer_state_dict(),
                    "iter": self.current_learning_iteration[agent_id], "infos": infos}, path)

    def load(self, agent_id, path, load_optimizer=True):
        d = torch.load(path, map_location=self.device)
        alg = self.algs[agent_id]
        alg.actor_critic.load_state_dict(d["model_state_dict"])
        if load_optimizer:
            alg.load_optimizer_state_dict(d["optimizer_state_dict"])
        self.current_learning_iteration[agent_id] = d["iter"]
        return d["infos"]

    def get_inference_policy(self, agent_id, device=None):
        ac = self.algs[agent_id].actor_critic
        ac.eval()
        if device is not None:
            ac.to(device)
        return ac.act_inference
