"""ANYmal-B cfg (values: reference legged_gym/envs/anymal_b/anymal_b_config.py:33-46)."""
from ..base.base_config import cfg_from_spec
from ..anymal_c.mixed_terrains.anymal_c_rough_config import AnymalCRoughCfg, AnymalCRoughCfgPPO

AnymalBRoughCfg = cfg_from_spec("AnymalBRoughCfg", (AnymalCRoughCfg,), dict(
    asset=dict(file="{LEGGED_GYM_ROOT_DIR}/resources/robots/anymal_b/urdf/anymal_b.urdf", name="anymal_b",
               foot_name="FOOT"),
    rewards=dict(scales=dict()),
), module=__name__)

AnymalBRoughCfgPPO = cfg_from_spec("AnymalBRoughCfgPPO", (AnymalCRoughCfgPPO,), dict(
    runner=dict(run_name="", experiment_name="rough_anymal_b", load_run=-1),
), module=__name__)
