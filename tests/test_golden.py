"""Oracle vs the committed golden fixtures (tests/golden/*.npz = outputs of the UNMODIFIED reference, made by
oracle/make_golden.py).  Runs anywhere (no /root/reference needed).  Integer / mask outputs must be identical; fp32
outputs are compared at 1e-6 because torch's vectorised CPU reductions may round differently on another CPU model."""
import os

import numpy as np
import pytest
import torch

from oracle import harness
from oracle.make_golden import CASES, GAME_CASES, GOLDEN_DIR, build, build_game


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_fixture(name):
    spec = CASES[name]
    fx = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case = build(spec)
    for k in ("root_states", "dof_state", "contact_forces", "episode_length_buf"):
        assert np.array_equal(case["state"][k], fx["in_" + k]), f"synthetic input generator drifted: {k}"
        case["state"][k] = fx["in_" + k]
    st = harness.torch_state(case)
    orc = harness.make_oracle(case, st)
    for step in range(1, spec["steps"] + 1):
        tables = harness.step_tables(case["seed"], step, spec["n"], orc.num_obs)
        orc.step(torch.from_numpy(fx[f"s{step}_actions"].copy()), tables)
        snap = harness.snapshot(orc)
        for k, v in snap.items():
            want = fx[f"s{step}_{k}"]
            got = v.numpy()
            if want.dtype == np.bool_ or np.issubdtype(want.dtype, np.integer):
                assert np.array_equal(got, want), f"{name} step {step}: {k}"
            else:
                assert np.allclose(got, want, rtol=1e-6, atol=1e-6), f"{name} step {step}: {k}"
        harness.apply_noise(st, {k: fx[f"s{step}_noise_{k}"] for k in ("dof_state", "contact_forces", "root_vel", "root_xy")})


def check_game_snapshot(fx, step, snap, sums, name, rtol, atol):
    for k, v in list(snap.items()) + [("sum_" + k, v) for k, v in sums.items()]:
        want, got = fx[f"s{step}_{k}"], v.numpy()
        if want.dtype == np.bool_ or np.issubdtype(want.dtype, np.integer):
            assert np.array_equal(got.astype(np.int64), want.astype(np.int64)), f"{name} step {step}: {k}"
        else:
            assert np.allclose(got, want, rtol=rtol, atol=atol * max(1.0, float(np.abs(want).max()))), f"{name} step {step}: {k}"


@pytest.mark.parametrize("name", sorted(GAME_CASES))
def test_game_oracle_reproduces_reference_fixture(name):
    """oracle/game_oracle.py against the outputs of the reference HighLevelGame / DecHighLevelGame classes."""
    import copy
    from oracle import game_oracle
    spec = GAME_CASES[name]
    fx = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case, gcfg, _ = build_game(spec)
    for k in ("root_states", "dof_state", "contact_forces", "episode_length_buf"):
        assert np.array_equal(case["state"][k], fx["in_" + k]), f"synthetic input generator drifted: {k}"
    st = harness.torch_state(case)
    ll = harness.make_oracle(case, st)
    orc = game_oracle.GameOracle(copy.deepcopy(gcfg), ll, spec["variant"])
    for step in range(1, spec["steps"] + 1):
        tables = harness.step_tables(case["seed"], step, spec["n"], ll.num_obs)
        prey, pred, acts = (torch.from_numpy(fx[f"s{step}_{k}"].copy()) for k in ("prey", "pred", "actions"))
        if spec["variant"] == "hl":
            orc.step(torch.cat((prey, pred), dim=1), acts, tables)
        else:
            orc.step_dec(pred, prey, acts, tables)
        check_game_snapshot(fx, step, game_oracle.snapshot(orc), game_oracle.game_sums(orc), name, 1e-6, 1e-6)
        harness.apply_noise(st, {k: fx[f"s{step}_noise_{k}"] for k in ("dof_state", "contact_forces", "root_vel", "root_xy")})
