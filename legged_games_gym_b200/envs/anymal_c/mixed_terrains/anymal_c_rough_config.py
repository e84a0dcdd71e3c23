"""ANYmal-C rough-terrain cfg (values: reference
legged_gym/envs/anymal_c/mixed_terrains/anymal_c_rough_config.py:33-94)."""
from ...base.base_config import cfg_from_spec, value
from ...base.legged_robot_config import LeggedRobotCfg, LeggedRobotCfgPPO

_LEGS = ("LF", "LH", "RF", "RH")
_Q0 = {}
for _leg in _LEGS:
    _front = _leg[1] == "F"
    _Q0[_leg + "_HAA"] = 0.0 if _leg[0] == "L" else -0.0
    _Q0[_leg + "_HFE"] = 0.4 if _front else -0.4
    _Q0[_leg + "_KFE"] = -0.8 if _front else 0.8

AnymalCRoughCfg = cfg_from_spec("AnymalCRoughCfg", (LeggedRobotCfg,), dict(
    env=dict(num_envs=4096, num_actions=12),
    terrain=dict(mesh_type="trimesh"),
    init_state=dict(pos=[0.0, 0.0, 0.6], default_joint_angles=value(_Q0)),
    control=dict(stiffness=value({"HAA": 80., "HFE": 80., "KFE": 80.}),
                 damping=value({"HAA": 2., "HFE": 2., "KFE": 2.}),
                 action_scale=0.5, decimation=4, use_actuator_network=True,
                 actuator_net_file="{LEGGED_GYM_ROOT_DIR}/resources/actuator_nets/anydrive_v3_lstm.pt"),
    asset=dict(file="{LEGGED_GYM_ROOT_DIR}/resources/robots/anymal_c/urdf/anymal_c.urdf", name="anymal_c",
               foot_name="FOOT", penalize_contacts_on=["SHANK", "THIGH"],
               terminate_after_contacts_on=["base"], self_collisions=1),
    domain_rand=dict(randomize_base_mass=True, added_mass_range=[-5., 5.]),
    rewards=dict(base_height_target=0.5, max_contact_force=500., only_positive_rewards=True, scales=dict()),
), module=__name__)

AnymalCRoughCfgPPO = cfg_from_spec("AnymalCRoughCfgPPO", (LeggedRobotCfgPPO,), dict(
    runner=dict(run_name="", experiment_name="rough_anymal_c", load_run=-1),
), module=__name__)
