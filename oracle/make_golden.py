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


if __name__ == "__main__":
    main()
