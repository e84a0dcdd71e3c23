"""python -m legged_games_gym_b200.scripts.train --task anymal_c_rough --num_envs 4096 --headless --max_iterations 100
(same entry point and flags as the reference's legged_gym/scripts/train.py:40-47)."""
from legged_games_gym_b200.envs import *  # noqa: F401,F403  (registers the tasks)
from legged_games_gym_b200.utils import get_args, task_registry


def train(args):
    env, env_cfg = task_registry.make_env(name=args.task, args=args)
    ppo_runner, train_cfg = task_registry.make_alg_runner(env=env, name=args.task, args=args)
    ppo_runner.learn(num_learning_iterations=train_cfg.runner.max_iterations, init_at_random_ep_len=True)
    return ppo_runner


if __name__ == "__main__":
    import sys
    train(get_args(sys.argv[1:]))
