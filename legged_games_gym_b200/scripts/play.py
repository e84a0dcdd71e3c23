"""python -m legged_games_gym_b200.scripts.play --task anymal_c_flat --load_run <run> --checkpoint <n>
Mirror of the reference's legged_gym/scripts/play.py:42-121 without the viewer: loads the last checkpoint, rolls the
inference policy, exports the actor as TorchScript, logs states / episode rewards (Logger prints, it does not plot)."""
import os

import torch

from legged_games_gym_b200 import LEGGED_GYM_ROOT_DIR
from legged_games_gym_b200.envs import *  # noqa: F401,F403
from legged_games_gym_b200.utils import get_args, export_policy_as_jit, task_registry, Logger

EXPORT_POLICY = True


def play(args, num_steps=None, log_root="default"):
    env_cfg, train_cfg = task_registry.get_cfgs(name=args.task)
    # override some parameters for testing (play.py:45-52)
    env_cfg.env.num_envs = min(env_cfg.env.num_envs, 50)
    env_cfg.terrain.num_rows = 5
    env_cfg.terrain.num_cols = 5
    env_cfg.terrain.curriculum = False
    env_cfg.noise.add_noise = False
    env_cfg.domain_rand.randomize_friction = False
    env_cfg.domain_rand.push_robots = False
    env, _ = task_registry.make_env(name=args.task, args=args, env_cfg=env_cfg)
    obs = env.get_observations()
    train_cfg.runner.resume = True
    ppo_runner, train_cfg = task_registry.make_alg_runner(env=env, name=args.task, args=args, train_cfg=train_cfg,
                                                          log_root=log_root)
    policy = ppo_runner.get_inference_policy(device=env.device)
    exported = None
    if EXPORT_POLICY:
        root = log_root if log_root not in ("default", None) else os.path.join(LEGGED_GYM_ROOT_DIR, "logs", train_cfg.runner.experiment_name)
        exported = os.path.join(root, "exported", "policies")
        export_policy_as_jit(ppo_runner.alg.actor_critic, exported)
        print("Exported policy as jit script to: ", exported)
    logger = Logger(env.dt)
    robot_index, joint_index = 0, 1
    stop_state_log = 100
    stop_rew_log = env.max_episode_length + 1
    n = 10 * int(env.max_episode_length) if num_steps is None else num_steps
    for i in range(n):
        actions = policy(obs.detach())
        obs, _, rews, dones, infos = env.step(actions.detach())
        if i < stop_state_log:
            logger.log_states({
                "dof_pos_target": actions[robot_index, joint_index].item() * env.cfg.control.action_scale,
                "dof_pos": env.dof_pos[robot_index, joint_index].item(),
                "dof_vel": env.dof_vel[robot_index, joint_index].item(),
                "dof_torque": env.torques[robot_index, joint_index].item(),
                "command_x": env.commands[robot_index, 0].item(),
                "command_y": env.commands[robot_index, 1].item(),
                "command_yaw": env.commands[robot_index, 2].item(),
                "base_vel_x": env.base_lin_vel[robot_index, 0].item(),
                "base_vel_y": env.base_lin_vel[robot_index, 1].item(),
                "base_vel_z": env.base_lin_vel[robot_index, 2].item(),
                "base_vel_yaw": env.base_ang_vel[robot_index, 2].item(),
                "contact_forces_z": env.contact_forces[robot_index, env.feet_indices, 2].cpu().numpy()})
        elif i == stop_state_log:
            logger.plot_states()
        if 0 < i < stop_rew_log:
            if infos.get("episode"):
                num_episodes = torch.sum(env.reset_buf).item()
                if num_episodes > 0:
                    logger.log_rewards(infos["episode"], num_episodes)
        elif i == stop_rew_log:
            logger.print_rewards()
    return env, logger, exported


if __name__ == "__main__":
    import sys
    play(get_args(sys.argv[1:]))
