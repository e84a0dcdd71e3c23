"""DecHighLevelGame cfg (values: reference legged_gym/envs/a1_game/dec_high_level_game_config.py:4-155)."""
from ..base.base_config import BaseConfig, cfg_from_spec
from .high_level_game_flat_config import GAME_COMMON, GAME_PPO

_spec = dict(GAME_COMMON)
_spec.update(
    env=dict(num_envs=2000, num_observations_prey=16, num_observations_predator=3, num_privileged_obs_prey=None,
             num_privileged_obs_predator=None, num_actions_prey=4, num_actions_predator=2, env_spacing=3.,
             send_timeouts=True, episode_length_s=20, capture_dist=0.5),
    terrain=dict(mesh_type="plane", curriculum=False, num_rows=10, num_cols=20),
    rewards_prey=dict(only_positive_rewards=True, scales=dict(evasion=0.9)),
    rewards_predator=dict(only_positive_rewards=False, scales=dict(pursuit=0.9)),
)

DecHighLevelGameCfg = cfg_from_spec("DecHighLevelGameCfg", (BaseConfig,), _spec, module=__name__)

DecHighLevelGameCfgPPO = cfg_from_spec("DecHighLevelGameCfgPPO", (BaseConfig,), dict(
    GAME_PPO,
    runner=dict(policy_class_name="ActorCritic", algorithm_class_name="PPO", num_steps_per_env=24, max_iterations=200,
                max_evolutions=20, save_interval=50, experiment_name="dec_high_level_game", run_name="", resume=False,
                load_run=-1, checkpoint=-1, resume_path=None),
), module=__name__)
