"""Generate tests/golden/*.npz from the REFERENCE's own outputs (container only: needs /root/reference).

    python -m oracle.make_golden

Each fixture stores the complete inputs of a short multi-step run (synthetic state, per-step actions and
inter-step noise, height field, terrain levels, uniform tables are re-derived from oracle/philox.py by seed) and the
outputs the UNMODIFIED reference code produced on them through oracle/ref_loader.py (CPU, fp32).
Cases include N not a multiple of 32 (partial tile), pushes, terrain/command curricula and time-outs.
"""
import os

import numpy as np
import torch

from . import harness, ref_loader

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SMALL_TERRAIN = {"terrain.num_rows": 3, "terrain.num_cols": 4, "terrain.terrain_length": 4., "terrain.terrain_width": 4.,
                 "terrain.border_size": 5, "terrain.max_init_terrain_level": 2}

CASES = {
    "anymal_c_flat_lstm_n64": dict(task="anymal_c_flat", n=64, seed=11, steps=3, overrides={}),
    "anymal_c_flat_pd_n64": dict(task="anymal_c_flat", n=64, seed=12, steps=3,
                                 overrides={"control.use_actuator_network": False}),
    "anymal_c_rough_n96": dict(task="anymal_c_rough", n=96, seed=13, steps=3, overrides=dict(SMALL_TERRAIN), xy_max=(18., 22.)),
    "a1_push_curriculum_n72": dict(task="a1", n=72, seed=14, steps=5, xy_max=(18., 22.),
                                   overrides=dict(SMALL_TERRAIN, **{"commands.curriculum": True,
                                                                    "domain_rand.push_interval_s": 0.04,
                                                                    "env.episode_length_s": 0.06})),
    "cassie_n40": dict(task="cassie", n=40, seed=15, steps=3, overrides=dict(SMALL_TERRAIN), xy_max=(18., 22.)),
    # two actors per env, 18-body contact view, predator re-spawn on reset, pushes, time-outs
    "low_level_game_n56": dict(task="low_level_game", n=56, seed=16, steps=5, xy_max=(18., 22.),
                               overrides=dict(SMALL_TERRAIN, **{"domain_rand.push_interval_s": 0.04,
                                                                "env.episode_length_s": 0.06})),
}


# the two hierarchical games on top of low_level_game (reference classes built by ref_loader.make_ref_game)
GAME_CASES = {
    "high_level_game_n48": dict(variant="hl", n=48, seed=21, steps=5, xy_max=(18., 22.),
                                ll_overrides=dict(SMALL_TERRAIN, **{"env.episode_length_s": 0.1}),
                                overrides={"terrain.num_rows": 3, "terrain.num_cols": 4, "env.env_radius": 12.0,
                                           "rewards.scales.termination": -2.0}),
    "dec_high_level_game_n48": dict(variant="dec", n=48, seed=22, steps=5, ll_overrides={"env.episode_length_s": 0.1},
                                    overrides={"env.episode_length_s": 0.08}),
}


def build_game(spec):
    """(case of the low-level env, game cfg): the low-level cfg carries what the game constructor changes in it."""
    gcfg = harness.product_game_cfg(spec["variant"], spec["n"], spec["overrides"])
    ll_ov = dict(spec["ll_overrides"])
    ll_ov.update(harness.game_ll_overrides(gcfg))
    case = harness.build_case("low_level_game", spec["n"], seed=spec["seed"], overrides=ll_ov, xy_max=spec.get("xy_max", (83., 163.)))
    st = harness.torch_state(case)
    harness.place_predators(st, spec["seed"])
    case["state"]["root_states"] = st["root_states"].numpy().copy()
    return case, gcfg, ll_ov


def run_reference_game(name, spec):
    import contextlib
    import io
    from . import game_oracle
    case, gcfg, ll_ov = build_game(spec)
    n, variant = spec["n"], spec["variant"]
    st = harness.torch_state(case)
    hs = None if case["height_samples"] is None else torch.from_numpy(case["height_samples"].copy())
    ll = ref_loader.make_ref_env("low_level_game", n, case["consts"], st, height_samples=hs, cfg_overrides=ll_ov,
                                 init_levels=case["init_levels"])
    ll.episode_length_buf[:] = torch.from_numpy(case["state"]["episode_length_buf"])
    tap = ref_loader.attach_tap(ll)
    ref = ref_loader.make_ref_game(variant, ll, tap, cfg_overrides=spec["overrides"])
    ref.predator_pos = st["root_states"][1::2, :3].clone()
    out = {}
    for k in ("root_states", "dof_state", "contact_forces", "episode_length_buf"):
        out["in_" + k] = case["state"][k]
    if case["height_samples"] is not None:
        out["in_height_samples"] = case["height_samples"]
        out["in_init_levels"] = case["init_levels"]
    for step in range(1, spec["steps"] + 1):
        prey, pred, acts = harness.game_inputs(case, step, variant)
        out[f"s{step}_prey"], out[f"s{step}_pred"], out[f"s{step}_actions"] = prey.numpy().copy(), pred.numpy().copy(), acts.numpy().copy()
        tap.set_tables(harness.step_tables(case["seed"], step, n, ll.num_obs))
        ref.ll_policy = lambda obs: acts.clone()
        with tap.active(), contextlib.redirect_stdout(io.StringIO()):
            if variant == "hl":
                ref.step(torch.cat((prey, pred), dim=1).clone())
            else:
                ref.step(pred.clone(), prey.clone())
        for k, v in game_oracle.snapshot(ref).items():
            out[f"s{step}_{k}"] = v.numpy()
        for k, v in game_oracle.game_sums(ref).items():
            out[f"s{step}_sum_{k}"] = v.numpy()
        out[f"s{step}_ll_obs_buf"] = ll.obs_buf.numpy().copy()
        out[f"s{step}_ll_rew_buf"] = ll.rew_buf.numpy().copy()
        noise = harness.make_noise(case, step, spec["seed"])
        for k, v in noise.items():
            out[f"s{step}_noise_{k}"] = v
        harness.apply_noise(st, noise)
    return out


def build(spec):
    return harness.build_case(spec["task"], spec["n"], seed=spec["seed"], overrides=spec["overrides"],
                              xy_max=spec.get("xy_max", (83., 163.)))


def actions_for(spec, step):
    return np.random.default_rng(spec["seed"] * 7919 + step).normal(0, 1, (spec["n"], 12)).astype(np.float32)


def run_reference(name, spec):
    case = build(spec)
    st = harness.torch_state(case)
    hs = None if case["height_samples"] is None else torch.from_numpy(case["height_samples"].copy())
    ref = ref_loader.make_ref_env(spec["task"], spec["n"], case["consts"], st, height_samples=hs,
                                  cfg_overrides=spec["overrides"], init_levels=case["init_levels"])
    ref.episode_length_buf[:] = torch.from_numpy(case["state"]["episode_length_buf"])
    tap = ref_loader.attach_tap(ref)
    out = {}
    for k in ("root_states", "dof_state", "contact_forces", "episode_length_buf"):
        out["in_" + k] = case["state"][k]
    if case["height_samples"] is not None:
        out["in_height_samples"] = case["height_samples"]
        out["in_init_levels"] = case["init_levels"]
    for step in range(1, spec["steps"] + 1):
        acts = actions_for(spec, step)
        out[f"s{step}_actions"] = acts
        tables = harness.step_tables(case["seed"], step, spec["n"], ref.num_obs)
        tap.set_tables(tables)
        with tap.active():
            ref.step(torch.from_numpy(acts.copy()))
        for k, v in harness.snapshot(ref).items():
            out[f"s{step}_{k}"] = v.numpy()
        noise = harness.make_noise(case, step, spec["seed"])
        for k, v in noise.items():
            out[f"s{step}_noise_{k}"] = v
        harness.apply_noise(st, noise)
    return out


def main():
    import sys
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(1)
    only = set(sys.argv[1:])
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        out = run_reference(name, spec)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        nres = [int(out[f"s{s}_reset_buf"].sum()) for s in range(1, spec["steps"] + 1)]
        print(f"{name}: {len(out)} arrays, resets/step {nres}, {os.path.getsize(path) / 1024:.0f} KiB")
    for name, spec in GAME_CASES.items():
        if only and name not in only:
            continue
        out = run_reference_game(name, spec)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        nres = [int(out[f"s{s}_reset_buf"].sum()) for s in range(1, spec["steps"] + 1)]
        nvis = [int(out[f"s{s}_" + ("obs_buf" if spec["variant"] == "hl" else "obs_buf_prey")][:, 15].sum()) for s in range(1, spec["steps"] + 1)]
        print(f"{name}: {len(out)} arrays, resets/step {nres}, visible/step {nvis}, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
