"""Unitree A1 cfg (values: reference legged_gym/envs/a1/a1_config.py:33-85)."""
from ..base.base_config import cfg_from_spec, value
from ..base.legged_robot_config import LeggedRobotCfg, LeggedRobotCfgPPO

A1_DEFAULT_ANGLES = {}
for _leg in ("FL", "RL", "FR", "RR"):
    A1_DEFAULT_ANGLES[_leg + "_hip_joint"] = 0.1 if _leg[1] == "L" else -0.1
    A1_DEFAULT_ANGLES[_leg + "_thigh_joint"] = 0.8 if _leg[0] == "F" else 1.
    A1_DEFAULT_ANGLES[_leg + "_calf_joint"] = -1.5

A1_SPEC = dict(
    init_state=dict(pos=[0.0, 0.0, 0.42], default_joint_angles=value(A1_DEFAULT_ANGLES)),
    control=dict(control_type="P", stiffness=value({"joint": 20.}), damping=value({"joint": 0.5}),
                 action_scale=0.25, decimation=4),
    asset=dict(file="{LEGGED_GYM_ROOT_DIR}/resources/robots/a1/urdf/a1.urdf", name="a1", foot_name="foot",
               penalize_contacts_on=["thigh", "calf"], terminate_after_contacts_on=["base"],
               self_collisions=1),
    rewards=dict(soft_dof_pos_limit=0.9, base_height_target=0.25,
                 scales=dict(torques=-0.0002, dof_pos_limits=-10.0)),
)

A1RoughCfg = cfg_from_spec("A1RoughCfg", (LeggedRobotCfg,), A1_SPEC, module=__name__)

A1RoughCfgPPO = cfg_from_spec("A1RoughCfgPPO", (LeggedRobotCfgPPO,), dict(
    algorithm=dict(entropy_coef=0.01),
    runner=dict(run_name="", experiment_name="rough_a1"),
), module=__name__)
