"""ANYmal-C flat-ground cfg (values: reference
legged_gym/envs/anymal_c/flat/anymal_c_flat_config.py:33-74)."""
from ...base.base_config import cfg_from_spec
from ..mixed_terrains.anymal_c_rough_config import AnymalCRoughCfg, AnymalCRoughCfgPPO

AnymalCFlatCfg = cfg_from_spec("AnymalCFlatCfg", (AnymalCRoughCfg,), dict(
    env=dict(num_observations=48),
    terrain=dict(mesh_type="plane", measure_heights=False),
    asset=dict(self_collisions=0),
    rewards=dict(max_contact_force=350.,
                 scales=dict(orientation=-5.0, torques=-0.000025, feet_air_time=2.)),
    commands=dict(heading_command=False, resampling_time=4., ranges=dict(ang_vel_yaw=[-1.5, 1.5])),
    domain_rand=dict(friction_range=[0., 1.5]),
), module=__name__)

AnymalCFlatCfgPPO = cfg_from_spec("AnymalCFlatCfgPPO", (AnymalCRoughCfgPPO,), dict(
    policy=dict(actor_hidden_dims=[128, 64, 32], critic_hidden_dims=[128, 64, 32], activation="elu"),
    algorithm=dict(entropy_coef=0.01),
    runner=dict(run_name="", experiment_name="flat_anymal_c", load_run=-1, max_iterations=300),
), module=__name__)
