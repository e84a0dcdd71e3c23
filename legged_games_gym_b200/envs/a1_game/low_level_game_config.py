"""Low-level predator/prey game cfg: A1 with direct yaw-rate commands (values: reference
legged_gym/envs/a1_game/low_level_game_config.py:34-98)."""
from ..base.base_config import cfg_from_spec
from ..base.legged_robot_config import LeggedRobotCfg, LeggedRobotCfgPPO
from ..a1.a1_config import A1_SPEC

_spec = dict(A1_SPEC)
_spec["env"] = dict(num_envs=2000)
# NB: ``ranges`` does not inherit from the base ranges class in the reference (line 40)
_spec["commands"] = dict(heading_command=False,
                         ranges=dict(__replace__=True, lin_vel_x=[-1.0, 1.0], lin_vel_y=[-1.0, 1.0],
                                     ang_vel_yaw=[-1, 1], heading=[-3.14, 3.14]))

LowLevelGameCfg = cfg_from_spec("LowLevelGameCfg", (LeggedRobotCfg,), _spec, module=__name__)

LowLevelGamePPO = cfg_from_spec("LowLevelGamePPO", (LeggedRobotCfgPPO,), dict(
    algorithm=dict(entropy_coef=0.01),
    runner=dict(run_name="", experiment_name="low_level_game"),
), module=__name__)
