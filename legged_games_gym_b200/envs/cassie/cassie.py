"""Cassie: LeggedRobot + the ``no_fly`` reward (mirror of reference legged_gym/envs/cassie/cassie.py:42-46);
the term itself is LGK_R_NO_FLY inside lgk_post_physics."""
from ..base.legged_robot import LeggedRobot, _native_reward


class Cassie(LeggedRobot):
    _reward_no_fly = _native_reward("no_fly")
