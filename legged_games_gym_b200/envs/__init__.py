"""Registered tasks (mirror of reference legged_gym/envs/__init__.py:31-56): all eight.  The two hierarchical game tasks
drive a frozen low-level policy; the reference loads it from a checkpoint that is not part of its tree (SURVEY.md section 2
row 14), here it can also be handed in (``ll_policy=``).

The task classes (and with them liblgk.so) are loaded on first use of this package's names (``from ...envs import *``,
``envs.task_registry``, ``envs.LeggedRobot`` ...), not on import of a sub-module: the cfg modules below ``envs/`` are plain
value tables and stay importable without the CUDA extension (the CPU reference arm of bench.py reads them)."""
import importlib

_CLASSES = {
    "LeggedRobot": ".base.legged_robot", "Anymal": ".anymal_c.anymal", "Cassie": ".cassie.cassie",
    "LowLevelGame": ".a1_game.low_level_game", "HighLevelGame": ".a1_game.high_level_game",
    "DecHighLevelGame": ".a1_game.dec_high_level_game",
}
_CFGS = {
    "AnymalCRoughCfg": ".anymal_c.mixed_terrains.anymal_c_rough_config", "AnymalCRoughCfgPPO": ".anymal_c.mixed_terrains.anymal_c_rough_config",
    "AnymalCFlatCfg": ".anymal_c.flat.anymal_c_flat_config", "AnymalCFlatCfgPPO": ".anymal_c.flat.anymal_c_flat_config",
    "AnymalBRoughCfg": ".anymal_b.anymal_b_config", "AnymalBRoughCfgPPO": ".anymal_b.anymal_b_config",
    "CassieRoughCfg": ".cassie.cassie_config", "CassieRoughCfgPPO": ".cassie.cassie_config",
    "A1RoughCfg": ".a1.a1_config", "A1RoughCfgPPO": ".a1.a1_config",
    "LowLevelGameCfg": ".a1_game.low_level_game_config", "LowLevelGamePPO": ".a1_game.low_level_game_config",
    "HighLevelGameFlatCfg": ".a1_game.high_level_game_flat_config", "HighLevelGameFlatCfgPPO": ".a1_game.high_level_game_flat_config",
    "DecHighLevelGameCfg": ".a1_game.dec_high_level_game_config", "DecHighLevelGameCfgPPO": ".a1_game.dec_high_level_game_config",
}
# (task name, task class, env cfg, train cfg): reference legged_gym/envs/__init__.py:49-56
_TASKS = (("anymal_c_rough", "Anymal", "AnymalCRoughCfg", "AnymalCRoughCfgPPO"),
          ("anymal_c_flat", "Anymal", "AnymalCFlatCfg", "AnymalCFlatCfgPPO"),
          ("anymal_b", "Anymal", "AnymalBRoughCfg", "AnymalBRoughCfgPPO"),
          ("a1", "LeggedRobot", "A1RoughCfg", "A1RoughCfgPPO"),
          ("cassie", "Cassie", "CassieRoughCfg", "CassieRoughCfgPPO"),
          ("low_level_game", "LowLevelGame", "LowLevelGameCfg", "LowLevelGamePPO"),
          ("high_level_game", "HighLevelGame", "HighLevelGameFlatCfg", "HighLevelGameFlatCfgPPO"),
          ("dec_high_level_game", "DecHighLevelGame", "DecHighLevelGameCfg", "DecHighLevelGameCfgPPO"))
__all__ = list(_CLASSES) + list(_CFGS) + ["task_registry"]
_registered = False


def _load(name):
    table = _CLASSES if name in _CLASSES else _CFGS
    return getattr(importlib.import_module(table[name], __name__), name)


def _register_all():
    global _registered
    if _registered:
        return
    _registered = True
    from ..utils.task_registry import task_registry
    g = globals()
    for name in list(_CLASSES) + list(_CFGS):
        g[name] = _load(name)
    g["task_registry"] = task_registry
    for task, cls, cfg, ppo in _TASKS:
        task_registry.register(task, g[cls], g[cfg](), g[ppo]())


def __getattr__(name):
    if name in _CLASSES or name in _CFGS or name == "task_registry":
        _register_all()
        return globals()[name]
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
