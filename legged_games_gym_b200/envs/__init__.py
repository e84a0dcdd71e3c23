"""Registered tasks (mirror of reference legged_gym/envs/__init__.py:31-56): all eight.  The two hierarchical game tasks
drive a frozen low-level policy; the reference loads it from a checkpoint that is not part of its tree (SURVEY.md section 2
row 14), here it can also be handed in (``ll_policy=``)."""
from .base.legged_robot import LeggedRobot
from .anymal_c.anymal import Anymal
from .anymal_c.mixed_terrains.anymal_c_rough_config import AnymalCRoughCfg, AnymalCRoughCfgPPO
from .anymal_c.flat.anymal_c_flat_config import AnymalCFlatCfg, AnymalCFlatCfgPPO
from .anymal_b.anymal_b_config import AnymalBRoughCfg, AnymalBRoughCfgPPO
from .cassie.cassie import Cassie
from .cassie.cassie_config import CassieRoughCfg, CassieRoughCfgPPO
from .a1.a1_config import A1RoughCfg, A1RoughCfgPPO
from .a1_game.low_level_game import LowLevelGame
from .a1_game.low_level_game_config import LowLevelGameCfg, LowLevelGamePPO
from .a1_game.high_level_game import HighLevelGame
from .a1_game.dec_high_level_game import DecHighLevelGame
from .a1_game.high_level_game_flat_config import HighLevelGameFlatCfg, HighLevelGameFlatCfgPPO
from .a1_game.dec_high_level_game_config import DecHighLevelGameCfg, DecHighLevelGameCfgPPO

from ..utils.task_registry import task_registry

task_registry.register("anymal_c_rough", Anymal, AnymalCRoughCfg(), AnymalCRoughCfgPPO())
task_registry.register("anymal_c_flat", Anymal, AnymalCFlatCfg(), AnymalCFlatCfgPPO())
task_registry.register("anymal_b", Anymal, AnymalBRoughCfg(), AnymalBRoughCfgPPO())
task_registry.register("a1", LeggedRobot, A1RoughCfg(), A1RoughCfgPPO())
task_registry.register("cassie", Cassie, CassieRoughCfg(), CassieRoughCfgPPO())
task_registry.register("low_level_game", LowLevelGame, LowLevelGameCfg(), LowLevelGamePPO())
task_registry.register("high_level_game", HighLevelGame, HighLevelGameFlatCfg(), HighLevelGameFlatCfgPPO())
task_registry.register("dec_high_level_game", DecHighLevelGame, DecHighLevelGameCfg(), DecHighLevelGameCfgPPO())
