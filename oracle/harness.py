"""TEST INFRASTRUCTURE ONLY -- shared scaffolding for parity tests, golden fixtures and the CPU baseline.

``TASKS`` maps each registered task name (reference legged_gym/envs/__init__.py:49-56) to the oracle
``kind`` and the product-side cfg class; ``build_case`` creates seeded synthetic inputs
(SURVEY.md section 8(d)); ``step_tables`` makes the per-step uniform tables from oracle/philox.py.
"""
import copy
import os

import numpy as np
import torch

from . import philox
from .legged_oracle import OracleEnv

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "legged_games_gym_b200")
LSTM_NPZ = os.path.join(_PKG, "resources", "actuator_nets", "anydrive_v3_lstm.npz")

TASKS = {
    "anymal_c_rough": ("anymal", "envs.anymal_c.mixed_terrains.anymal_c_rough_config", "AnymalCRoughCfg"),
    "anymal_c_flat": ("anymal", "envs.anymal_c.flat.anymal_c_flat_config", "AnymalCFlatCfg"),
    "anymal_b": ("anymal", "envs.anymal_b.anymal_b_config", "AnymalBRoughCfg"),
    "a1": ("legged", "envs.a1.a1_config", "A1RoughCfg"),
    "cassie": ("cassie", "envs.cassie.cassie_config", "CassieRoughCfg"),
    "low_level_game": ("llg", "envs.a1_game.low_level_game_config", "LowLevelGameCfg"),
}


def lstm_weights():
    return dict(np.load(LSTM_NPZ))


def product_cfg(task, num_envs, overrides=None):
    import importlib
    kind, mod, cls = TASKS[task]
    cfg = getattr(importlib.import_module("legged_games_gym_b200." + mod), cls)()
    cfg.env.num_envs = num_envs
    apply_overrides(cfg, overrides)
    return cfg


def apply_overrides(cfg, overrides):
    for path, val in (overrides or {}).items():
        obj = cfg
        parts = path.split(".")
        for p in parts[:-1]:
            obj = getattr(obj, p)
        setattr(obj, parts[-1], val)


def build_case(task, num_envs, seed=0, overrides=None, hf_shape=None, xy_max=(83., 163.)):
    """Returns dict(cfg, consts, state(np), height_samples(np int16|None), terrain_origins, init_levels)."""
    from legged_games_gym_b200.sim.asset_model import model_for_asset
    from legged_games_gym_b200.sim.state_feeder import synth_state, synth_height_field, synth_terrain_origins
    cfg = product_cfg(task, num_envs, overrides)
    model = model_for_asset(cfg.asset)
    consts = model.consts(cfg.asset)
    llg = TASKS[task][0] == "llg"        # two actors per env (robot + predator sphere), 18 bodies in the contact view
    st = synth_state(num_envs, model.num_bodies + (1 if llg else 0), model.num_dof, seed, xy_max=xy_max,
                     actors_per_env=2 if llg else 1)
    rough = cfg.terrain.mesh_type in ("heightfield", "trimesh")
    hs = None
    origins = None
    levels = None
    if rough:
        t = cfg.terrain
        rows = int(t.num_rows * t.terrain_length / t.horizontal_scale) + 2 * int(t.border_size / t.horizontal_scale)
        cols = int(t.num_cols * t.terrain_width / t.horizontal_scale) + 2 * int(t.border_size / t.horizontal_scale)
        if hf_shape is not None:
            rows, cols = hf_shape
        hs = synth_height_field(rows, cols, seed)
        origins = synth_terrain_origins(t)
        g = np.random.default_rng(seed + 104729)
        max_init = t.max_init_terrain_level if t.curriculum else t.num_rows - 1
        levels = g.integers(0, max_init + 1, num_envs).astype(np.int64)
    return dict(task=task, kind=TASKS[task][0], cfg=cfg, consts=consts, state=st, height_samples=hs,
                terrain_origins=origins, init_levels=levels, seed=seed)


def torch_state(case):
    return {k: torch.from_numpy(case["state"][k].copy()) for k in ("root_states", "dof_state", "contact_forces")}


def make_oracle(case, state=None, device="cpu"):
    """device != "cpu": the same torch code with every tensor on that device (the caller must also step it under
    ``with torch.device(device)``): bench.py's "reference algorithm as eager torch on the GPU" figure."""
    state = state if state is not None else torch_state(case)
    state = {k: v.to(device) for k, v in state.items()} if str(device) != "cpu" else state
    hs = None if case["height_samples"] is None else torch.from_numpy(case["height_samples"].copy()).to(device)
    w = lstm_weights() if case["kind"] == "anymal" else None
    with torch.device(device):
        env = OracleEnv(copy.deepcopy(case["cfg"]), case["consts"], state, kind=case["kind"], height_samples=hs,
                        terrain_origins=case["terrain_origins"], init_levels=case["init_levels"], lstm_weights=w)
    env.episode_length_buf[:] = torch.from_numpy(case["state"]["episode_length_buf"]).to(device)
    return env


def step_tables(seed, step, num_envs, num_obs, num_dof=12, env_offset=0):
    ids = np.arange(env_offset, env_offset + num_envs)
    return {
        philox.STREAM_CMD: philox.uniforms(seed, step, ids, philox.STREAM_CMD, 3),
        philox.STREAM_PUSH: philox.uniforms(seed, step, ids, philox.STREAM_PUSH, 2),
        philox.STREAM_RESET_DOF: philox.uniforms(seed, step, ids, philox.STREAM_RESET_DOF, num_dof),
        philox.STREAM_RESET_ROOT: philox.uniforms(seed, step, ids, philox.STREAM_RESET_ROOT, 8),
        philox.STREAM_RESET_CMD: philox.uniforms(seed, step, ids, philox.STREAM_RESET_CMD, 3),
        philox.STREAM_TERRAIN: philox.raw_u32(seed, step, ids, philox.STREAM_TERRAIN, 1),
        philox.STREAM_OBS: philox.obs_uniforms(seed, step, ids, num_obs),
        philox.STREAM_PREDATOR: philox.uniforms(seed, step, ids, philox.STREAM_PREDATOR, 4),
        philox.STREAM_GAME_ROOT: philox.uniforms(seed, step, ids, philox.STREAM_GAME_ROOT, 8),
        philox.STREAM_GAME_PREDATOR: philox.uniforms(seed, step, ids, philox.STREAM_GAME_PREDATOR, 4),
        philox.STREAM_GAME_DOF: philox.uniforms(seed, step, ids, philox.STREAM_GAME_DOF, num_dof),
    }


def game_ll_overrides(game_cfg):
    """What the games' constructors change in the low-level cfg before building the low-level env (HLG:69-85)."""
    return {"terrain.num_rows": game_cfg.terrain.num_rows, "terrain.num_cols": game_cfg.terrain.num_cols,
            "terrain.curriculum": game_cfg.terrain.curriculum, "terrain.mesh_type": game_cfg.terrain.mesh_type,
            "noise.add_noise": game_cfg.noise.add_noise,
            "domain_rand.randomize_friction": game_cfg.domain_rand.randomize_friction,
            "domain_rand.push_robots": game_cfg.domain_rand.push_robots, "rewards.scales.torques": -5.}


def product_game_cfg(variant, num_envs, overrides=None):
    import importlib
    mod, cls = (("envs.a1_game.high_level_game_flat_config", "HighLevelGameFlatCfg") if variant == "hl"
                else ("envs.a1_game.dec_high_level_game_config", "DecHighLevelGameCfg"))
    cfg = getattr(importlib.import_module("legged_games_gym_b200." + mod), cls)()
    cfg.env.num_envs = num_envs
    apply_overrides(cfg, overrides)
    return cfg


def game_inputs(case, step, variant):
    """Seeded high-level commands of one game step: prey command [N,4] and predator command [N,2] (scaled so that
    the clips of HLG:162-169 trigger), plus low-level actions [N,12]."""
    n = case["cfg"].env.num_envs
    g = np.random.default_rng(7919 * step + 13)
    prey = g.normal(0, 1.5, (n, 4)).astype(np.float32)
    prey[:, 2] *= 3.0                                   # heading beyond +-pi: wrap_to_pi path
    pred = g.normal(0, 2.5, (n, 2)).astype(np.float32)
    acts = g.normal(0, 1, (n, 12)).astype(np.float32)
    return torch.from_numpy(prey), torch.from_numpy(pred), torch.from_numpy(acts)


def place_predators(state, seed):
    """Put the predator sphere of every env near its prey (synthetic states scatter them over the whole terrain): a third
    within capture distance, the rest 0.5..6 m away at a random bearing, so that captures, visible and occluded
    predators all occur.  `state` holds torch tensors; mutated in place."""
    r = state["root_states"]
    n = r.shape[0] // 2
    g = np.random.default_rng(seed + 4242)
    dist = np.where(g.random(n) < 0.33, g.uniform(0.05, 0.45, n), g.uniform(0.5, 6.0, n)).astype(np.float32)
    ang = g.uniform(-np.pi, np.pi, n).astype(np.float32)
    r[1::2, 0] = r[0::2, 0] + torch.from_numpy(dist * np.cos(ang)).to(r.device)
    r[1::2, 1] = r[0::2, 1] + torch.from_numpy(dist * np.sin(ang)).to(r.device)
    r[1::2, 2] = 0.3


def make_noise(case, step, seed):
    """Deterministic stand-in for a physics step between env steps (new dof / contact / root velocities so that
    consecutive steps differ).  Returned as arrays so fixtures can store them."""
    g = np.random.default_rng(seed * 1000003 + step)
    st = case["state"]
    return dict(dof_state=g.normal(0., 0.05, st["dof_state"].shape).astype(np.float32),
                contact_forces=g.normal(0., 0.05, st["contact_forces"].shape).astype(np.float32),
                root_vel=g.normal(0., 0.05, (st["root_states"].shape[0], 6)).astype(np.float32),
                root_xy=g.normal(0., 0.02, (st["root_states"].shape[0], 2)).astype(np.float32))


def apply_noise(state, noise):
    """state: dict of torch tensors (any device), mutated in place.  Same arithmetic for every implementation."""
    for k in ("dof_state", "contact_forces"):
        t = state[k]
        t += torch.from_numpy(noise[k]).to(t.device) * (t != 0)
    r = state["root_states"]
    r[:, 7:13] += torch.from_numpy(noise["root_vel"]).to(r.device)
    r[:, 0:2] += torch.from_numpy(noise["root_xy"]).to(r.device)


def perturb_state(state, step, seed):
    g = np.random.default_rng(seed * 1000003 + step)
    noise = dict(dof_state=g.normal(0., 0.05, tuple(state["dof_state"].shape)).astype(np.float32),
                 contact_forces=g.normal(0., 0.05, tuple(state["contact_forces"].shape)).astype(np.float32),
                 root_vel=g.normal(0., 0.05, (state["root_states"].shape[0], 6)).astype(np.float32),
                 root_xy=g.normal(0., 0.02, (state["root_states"].shape[0], 2)).astype(np.float32))
    apply_noise(state, noise)


def snapshot(env):
    """Comparable view of an env (reference instance, OracleEnv, or the product LeggedRobot): name -> CPU tensor."""
    names = ["obs_buf", "rew_buf", "reset_buf", "time_out_buf", "episode_length_buf", "commands", "base_lin_vel",
             "base_ang_vel", "projected_gravity", "last_actions", "last_dof_vel", "last_root_vel", "feet_air_time",
             "last_contacts", "root_states", "dof_state", "torques", "env_origins"]
    d = {n: getattr(env, n).detach().cpu().clone() for n in names}
    if isinstance(env.measured_heights, torch.Tensor):
        d["measured_heights"] = env.measured_heights.detach().cpu().clone()
    for k, v in env.episode_sums.items():
        d["sum_" + k] = v.detach().cpu().clone()
    for k, v in env.extras.get("episode", {}).items():
        d["ex_" + k] = torch.as_tensor(v).detach().cpu().clone().float()
    if "time_outs" in env.extras:
        d["ex_time_outs"] = env.extras["time_outs"].detach().cpu().clone()
    if hasattr(env, "terrain_levels"):
        d["terrain_levels"] = env.terrain_levels.detach().cpu().clone()
    if hasattr(env, "sea_hidden_state") and env.cfg.control.use_actuator_network:
        d["sea_h"] = env.sea_hidden_state.detach().cpu().clone()
        d["sea_c"] = env.sea_cell_state.detach().cpu().clone()
    return d
