"""Cassie cfg (values: reference legged_gym/envs/cassie/cassie_config.py:33-109)."""
from ..base.base_config import cfg_from_spec, value
from ..base.legged_robot_config import LeggedRobotCfg, LeggedRobotCfgPPO

_SQUARE = [-0.5, -0.4, -0.3, -0.2, -0.1, 0., 0.1, 0.2, 0.3, 0.4, 0.5]
_Q0 = {}
for _side, _sgn in (("left", 1.0), ("right", -1.0)):
    _Q0["hip_abduction_" + _side] = 0.1 * _sgn
    _Q0["hip_rotation_" + _side] = 0.
    _Q0["hip_flexion_" + _side] = 1.
    _Q0["thigh_joint_" + _side] = -1.8
    _Q0["ankle_joint_" + _side] = 1.57
    _Q0["toe_joint_" + _side] = -1.57

CassieRoughCfg = cfg_from_spec("CassieRoughCfg", (LeggedRobotCfg,), dict(
    env=dict(num_envs=4096, num_observations=169, num_actions=12),
    terrain=dict(measured_points_x=list(_SQUARE), measured_points_y=list(_SQUARE)),
    init_state=dict(pos=[0.0, 0.0, 1.], default_joint_angles=value(_Q0)),
    control=dict(
        stiffness=value({"hip_abduction": 100.0, "hip_rotation": 100.0, "hip_flexion": 200.,
                         "thigh_joint": 200., "ankle_joint": 200., "toe_joint": 40.}),
        damping=value({"hip_abduction": 3.0, "hip_rotation": 3.0, "hip_flexion": 6., "thigh_joint": 6.,
                       "ankle_joint": 6., "toe_joint": 1.}),
        action_scale=0.5, decimation=4),
    asset=dict(file="{LEGGED_GYM_ROOT_DIR}/resources/robots/cassie/urdf/cassie.urdf", name="cassie",
               foot_name="toe", terminate_after_contacts_on=["pelvis"], flip_visual_attachments=False,
               self_collisions=1),
    rewards=dict(soft_dof_pos_limit=0.95, soft_dof_vel_limit=0.9, soft_torque_limit=0.9,
                 max_contact_force=300., only_positive_rewards=False,
                 scales=dict(termination=-200., tracking_ang_vel=1.0, torques=-5.e-6, dof_acc=-2.e-7,
                             lin_vel_z=-0.5, feet_air_time=5., dof_pos_limits=-1., no_fly=0.25,
                             dof_vel=-0.0, ang_vel_xy=-0.0, feet_contact_forces=-0.)),
), module=__name__)

CassieRoughCfgPPO = cfg_from_spec("CassieRoughCfgPPO", (LeggedRobotCfgPPO,), dict(
    runner=dict(run_name="", experiment_name="rough_cassie"),
    algorithm=dict(entropy_coef=0.01),
), module=__name__)
