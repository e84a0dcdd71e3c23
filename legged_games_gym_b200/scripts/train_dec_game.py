"""python -m legged_games_gym_b200.scripts.train_dec_game --task dec_high_level_game --headless
(same entry point as the reference's legged_gym/scripts/train_dec_game.py:39-50)."""
from legged_games_gym_b200.envs import *  # noqa: F401,F403
from legged_games_gym_b200.utils import get_args, task_registry


def train(args, log_root="default", **env_kwargs):
    print("[train_game] making the high-level environment")
    env, env_cfg = task_registry.make_env(name=args.task, args=args, **env_kwargs)
    print("[train_game] making the algorithm runner...")
    ppo_runner, train_cfg = task_registry.make_dec_alg_runner(env=env, name=args.task, args=args, log_root=log_root)
    print("[train_game] starting the PPO runner...")
    ppo_runner.learn(max_num_evolutions=train_cfg.runner.max_evolutions, num_learning_iterations=train_cfg.runner.max_iterations,
                     init_at_random_ep_len=True)
    return ppo_runner


if __name__ == "__main__":
    import sys
    train(get_args(sys.argv[1:]))
