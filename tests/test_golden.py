"""Oracle vs the committed golden fixtures (tests/golden/*.npz = outputs of the UNMODIFIED reference, made by
oracle/make_golden.py).  Runs anywhere (no /root/reference needed).  Integer / mask outputs must be identical; fp32
outputs are compared at 1e-6 because torch's vectorised CPU reductions may round differently on another CPU model."""
import os

import numpy as np
import pytest
import torch

from oracle import harness
from oracle.make_golden import CASES, GOLDEN_DIR, build


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
