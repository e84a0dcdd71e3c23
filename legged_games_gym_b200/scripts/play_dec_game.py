"""python -m legged_games_gym_b200.scripts.play_dec_game --task dec_high_level_game --load_run <run> --checkpoint <n>
Mirror of the reference's legged_gym/scripts/play_dec_game.py:44-97 without the viewer / frame recording."""
from legged_games_gym_b200.envs import *  # noqa: F401,F403
from legged_games_gym_b200.utils import get_args, task_registry


def play_dec_game(args, num_steps=None, log_root="default", **env_kwargs):
    env_cfg, train_cfg = task_registry.get_cfgs(name=args.task)
    # override some parameters for testing (play_dec_game.py:47-56)
    env_cfg.env.num_envs = min(env_cfg.env.num_envs, 5)
    env_cfg.terrain.mesh_type = "plane"
    env_cfg.terrain.num_rows = 4
    env_cfg.terrain.num_cols = 4
    env_cfg.terrain.curriculum = False
    env_cfg.noise.add_noise = False
    env_cfg.domain_rand.randomize_friction = False
    env_cfg.domain_rand.push_robots = False
    print("[play_dec_game] making environment...")
    env, _ = task_registry.make_env(name=args.task, args=args, env_cfg=env_cfg, **env_kwargs)
    obs_pred, obs_prey = env.get_observations_pred(), env.get_observations_prey()
    train_cfg.runner.resume = True
    runner, train_cfg = task_registry.make_dec_alg_runner(env=env, name=args.task, args=args, train_cfg=train_cfg, log_root=log_root)
    policy_pred = runner.get_inference_policy(agent_id=0, device=env.device)
    policy_prey = runner.get_inference_policy(agent_id=1, device=env.device)
    n = 10 * int(env.max_episode_length) if num_steps is None else num_steps
    for _ in range(n):
        actions_pred = policy_pred(obs_pred.detach())
        actions_prey = policy_prey(obs_prey.detach())
        obs_pred, obs_prey, _, _, rews_pred, rews_prey, dones, infos = env.step(actions_pred.detach(), actions_prey.detach())
    return env


if __name__ == "__main__":
    import sys
    play_dec_game(get_args(sys.argv[1:]))
