import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import harness, game_oracle
from oracle.legged_oracle import yaw_only_apply
from legged_games_gym_b200.envs import task_registry
variant, n = "hl", 96
case = harness.build_case("low_level_game", n, seed=9, overrides={"env.episode_length_s": 0.1})
st_or = harness.torch_state(case); harness.place_predators(st_or, 9)
cfg = copy.deepcopy(task_registry.env_cfgs["high_level_game"]); cfg.env.num_envs = n
# mimic the product's ll cfg derivation
ll_cfg = copy.deepcopy(case["cfg"]); ll_cfg.rewards.scales.torques = -5.
case_or = dict(case, cfg=ll_cfg)
ll_or = harness.make_oracle(case_or, st_or)
orc = game_oracle.GameOracle(cfg, ll_or, variant)
for step in range(1, 7):
    tables = harness.step_tables(case["seed"], step, n, ll_or.num_obs)
    prey, pred, acts = harness.game_inputs(case, step, variant)
    orc.step(torch.cat((prey, pred), dim=1).clone(), acts.clone(), tables)
    e = 0
    rel = orc.predator_pos - orc.prey_states[:, :3]
    fwd = yaw_only_apply(ll_or.base_quat, ll_or.forward_vec)
    dot = torch.sum(fwd * rel, dim=-1); den = torch.norm(fwd, dim=-1) * torch.norm(rel, dim=-1)
    ang = torch.acos(dot / den)
    print(step, "reset", bool(orc.reset_buf[e]), "ll reset", bool(ll_or.reset_buf[e]), "quat", ll_or.base_quat[e].tolist(), "root quat", ll_or.root_states[0, 3:7].tolist(),
          "rel", rel[e].tolist(), "ang", float(ang[e]), "obs", orc.obs_buf[e, 9:16].tolist())
    harness.apply_noise(st_or, harness.make_noise(case, step, 5))
