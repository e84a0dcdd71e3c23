class VecEnv:
    """Protocol marker (rsl_rl.env.VecEnv): num_envs, num_obs, num_privileged_obs, num_actions, max_episode_length,
    obs_buf, rew_buf, reset_buf, episode_length_buf, extras, device; step / reset / get_observations."""
