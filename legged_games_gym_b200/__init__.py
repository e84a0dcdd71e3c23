"""legged_games_gym_b200 -- B200-native drop-in for legged_gym's per-environment step.

Host side mirrors the reference's operator interface (``LeggedRobot`` / ``task_registry`` /
cfg classes; /root/reference/legged_gym/__init__.py:31-34 for the two path constants); the
hot path lives in ``csrc/`` behind the C ABI declared in ``include/lgk.h`` and is loaded by
``_native``.  There is no CPU fallback: importing ``_native`` without the built library raises.
"""
import os

LEGGED_GYM_ROOT_DIR = os.path.dirname(os.path.realpath(__file__))
LEGGED_GYM_ENVS_DIR = os.path.join(LEGGED_GYM_ROOT_DIR, "envs")
PACKAGE_ROOT = LEGGED_GYM_ROOT_DIR


def install_as_legged_gym():
    """Alias this package as ``legged_gym`` in sys.modules so reference user code
    (``from legged_gym.envs import *``; ``from legged_gym.utils import task_registry``) runs unchanged."""
    import importlib
    import sys
    sys.modules.setdefault("legged_gym", sys.modules[__name__])
    for sub in ("envs", "utils", "envs.base", "envs.base.legged_robot", "envs.base.legged_robot_config",
                "envs.base.base_config", "envs.base.base_task", "utils.task_registry", "utils.helpers",
                "utils.math"):
        m = importlib.import_module(__name__ + "." + sub)
        sys.modules.setdefault("legged_gym." + sub, m)
