"""python -m legged_games_gym_b200.scripts.play_game --task high_level_game --load_run <run>
Mirror of the RL-policy branch of the reference's legged_gym/scripts/play_game.py:53-104 (no viewer / frame recording):
loads the last high-level checkpoint, exports the actor as TorchScript and rolls the inference policy."""
import os

from legged_games_gym_b200 import LEGGED_GYM_ROOT_DIR
from legged_games_gym_b200.envs import *  # noqa: F401,F403
from legged_games_gym_b200.utils import get_args, export_policy_as_jit, task_registry

EXPORT_POLICY = True


def play_game(args, num_steps=None, log_root="default", **env_kwargs):
    env_cfg, train_cfg = task_registry.get_cfgs(name=args.task)
    # override some parameters for testing (play_game.py:56-65)
    env_cfg.env.num_envs = min(env_cfg.env.num_envs, 5)
    env_cfg.terrain.mesh_type = "plane"
    env_cfg.terrain.num_rows = 4
    env_cfg.terrain.num_cols = 4
    env_cfg.terrain.curriculum = False
    env_cfg.noise.add_noise = False
    env_cfg.domain_rand.randomize_friction = False
    env_cfg.domain_rand.push_robots = False
    env, _ = task_registry.make_env(name=args.task, args=args, env_cfg=env_cfg, **env_kwargs)
    obs = env.get_observations()
    train_cfg.runner.resume = True
    ppo_runner, train_cfg = task_registry.make_alg_runner(env=env, name=args.task, args=args, train_cfg=train_cfg, log_root=log_root)
    policy = ppo_runner.get_inference_policy(device=env.device)
    exported = None
    if EXPORT_POLICY:
        root = log_root if log_root not in ("default", None) else os.path.join(LEGGED_GYM_ROOT_DIR, "logs", train_cfg.runner.experiment_name)
        exported = os.path.join(root, "exported", "policies")
        export_policy_as_jit(ppo_runner.alg.actor_critic, exported)
        print("Exported policy as jit script to: ", exported)
    n = 10 * int(env.max_episode_length) if num_steps is None else num_steps
    for _ in range(n):
        actions = policy(obs.detach())
        obs, _, rews, dones, infos = env.step(actions.detach())
    return env, exported


if __name__ == "__main__":
    import sys
    play_game(get_args(sys.argv[1:]))
