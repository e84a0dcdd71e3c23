"""HighLevelGame cfg (values: reference legged_gym/envs/a1_game/high_level_game_flat_config.py:4-150).  Like the
reference these classes derive from BaseConfig only, not from LeggedRobotCfg."""
from ..base.base_config import BaseConfig, cfg_from_spec, value
from ..a1.a1_config import A1_DEFAULT_ANGLES

_PHYSX = dict(num_threads=10, solver_type=1, num_position_iterations=4, num_velocity_iterations=0, contact_offset=0.01,
              rest_offset=0.0, bounce_threshold_velocity=0.5, max_depenetration_velocity=1.0,
              max_gpu_contact_pairs=2 ** 23, default_buffer_size_multiplier=5, contact_collection=2)

GAME_COMMON = dict(
    commands=dict(heading_command=True,
                  ranges=dict(lin_vel_x=[-1.0, 1.0], lin_vel_y=[-1.0, 1.0], ang_vel_yaw=[-1, 1], heading=[-3.14, 3.14],
                              predator_lin_vel_x=[-2.0, 2.0], predator_lin_vel_y=[-2.0, 2.0])),
    init_state=dict(predator_pos=[0.0, 0.0, 0.3], pos=[0.0, 0.0, 0.42], rot=[0.0, 0.0, 0.0, 1.0],
                    lin_vel=[0.0, 0.0, 0.0], ang_vel=[0.0, 0.0, 0.0], default_joint_angles=value(A1_DEFAULT_ANGLES)),
    domain_rand=dict(randomize_friction=True, friction_range=[0.5, 1.25], randomize_base_mass=False,
                     added_mass_range=[-1., 1.], push_robots=True, push_interval_s=15, max_push_vel_xy=1.),
    noise=dict(add_noise=True, noise_level=1.0),
    viewer=dict(ref_env=0, pos=[10, 0, 6], lookat=[11., 5, 3.]),
    sim=dict(dt=0.005, substeps=1, gravity=[0., 0., -9.81], up_axis=1, physx=dict(_PHYSX)),
)

_spec = dict(GAME_COMMON)
_spec.update(
    env=dict(num_envs=2000, num_observations=19, num_privileged_obs=None, num_actions=6, env_spacing=3.,
             send_timeouts=True, episode_length_s=20, env_radius=None, capture_dist=0.5),
    terrain=dict(mesh_type="trimesh", curriculum=True, num_rows=10, num_cols=20),
    rewards=dict(only_positive_rewards=True, scales=dict(evasion=0.9, pursuit=0.9)),
)
_spec["commands"] = dict(GAME_COMMON["commands"], num_commands=4)

HighLevelGameFlatCfg = cfg_from_spec("HighLevelGameFlatCfg", (BaseConfig,), _spec, module=__name__)

GAME_PPO = dict(
    seed=1,
    runner_class_name="OnPolicyRunner",
    policy=dict(init_noise_std=1.0, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128], activation="elu"),
    algorithm=dict(value_loss_coef=1.0, use_clipped_value_loss=True, clip_param=0.2, entropy_coef=0.01,
                   num_learning_epochs=5, num_mini_batches=4, learning_rate=1.e-3, schedule="adaptive", gamma=0.99,
                   lam=0.95, desired_kl=0.01, max_grad_norm=1.),
)

HighLevelGameFlatCfgPPO = cfg_from_spec("HighLevelGameFlatCfgPPO", (BaseConfig,), dict(
    GAME_PPO,
    runner=dict(policy_class_name="ActorCritic", algorithm_class_name="PPO", num_steps_per_env=24, max_iterations=1500,
                save_interval=50, experiment_name="high_level_game_flat", run_name="", resume=False, load_run=-1,
                checkpoint=-1, resume_path=None),
), module=__name__)
