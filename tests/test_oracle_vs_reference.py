"""Pin the oracle restatement (oracle/legged_oracle.py) BIT-FOR-BIT against the unmodified reference executed through
oracle/ref_loader.py.  Needs /root/reference, so it runs in the build container only (skipped on the GPU box, where the
committed fixtures of tests/test_golden.py carry the pin)."""
import numpy as np
import pytest
import torch

from oracle import harness, ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not mounted")

CASES = [("anymal_c_flat", 64, None), ("anymal_c_rough", 128, None), ("a1", 128, None), ("cassie", 96, None),
         ("anymal_b", 64, None), ("anymal_c_flat", 64, {"control.use_actuator_network": False}),
         ("a1", 96, {"commands.curriculum": True, "domain_rand.push_interval_s": 0.04, "env.episode_length_s": 0.1}),
         ("a1", 64, {"control.control_type": "V"}), ("a1", 64, {"control.control_type": "T"}),
         # low_level_game: two actors per env (prey robot + predator sphere), 18-body contact view, predator re-spawn
         ("low_level_game", 96, None),
         ("low_level_game", 64, {"domain_rand.push_interval_s": 0.04, "env.episode_length_s": 0.1}),
         # configuration switches off the defaults: yaw-rate commands instead of heading, fixed terrain levels, no noise and
         # signed rewards, two decimation sub-steps, rapid command resampling without pushes, a rough-terrain robot without scan
         ("anymal_c_rough", 64, {"commands.heading_command": False, "env.episode_length_s": 0.1}),
         ("anymal_c_rough", 64, {"terrain.curriculum": False, "control.decimation": 2}),
         ("a1", 64, {"noise.add_noise": False, "rewards.only_positive_rewards": False}),
         ("a1", 64, {"domain_rand.push_robots": False, "commands.resampling_time": 0.04}),
         ("a1", 64, {"terrain.measure_heights": False, "env.num_observations": 48})]


@pytest.mark.parametrize("task,n,ov", CASES)
def test_oracle_is_bit_exact_with_reference(task, n, ov):
    case = harness.build_case(task, n, seed=3, overrides=ov)
    st_ref, st_or = harness.torch_state(case), harness.torch_state(case)
    hs = None if case["height_samples"] is None else torch.from_numpy(case["height_samples"].copy())
    ref = ref_loader.make_ref_env(task, n, case["consts"], st_ref, height_samples=hs, cfg_overrides=ov,
                                  init_levels=case["init_levels"])
    ref.episode_length_buf[:] = torch.from_numpy(case["state"]["episode_length_buf"])
    tap = ref_loader.attach_tap(ref)
    orc = harness.make_oracle(case, st_or)
    assert ref.reward_names == orc.reward_names
    for step in range(1, 6):
        tables = harness.step_tables(case["seed"], step, n, ref.num_obs)
        acts = torch.from_numpy(np.random.default_rng(step).normal(0, 1, (n, 12)).astype(np.float32))
        tap.set_tables(tables)
        with tap.active():
            ref.step(acts.clone())
        orc.step(acts.clone(), tables)
        a, b = harness.snapshot(ref), harness.snapshot(orc)
        assert a.keys() == b.keys()
        for k in a:
            assert torch.equal(a[k], b[k]), f"{task} step {step}: {k} differs from the reference"
        harness.perturb_state(st_ref, step, 5)
        harness.perturb_state(st_or, step, 5)


GAME_CASES = [("hl", 96, None), ("hl", 64, {"env.env_radius": 40.0, "rewards.scales.termination": -2.0, "rewards.only_positive_rewards": False}),
              ("dec", 96, {"env.episode_length_s": 0.08}), ("dec", 64, {"rewards_prey.scales.termination": -3.0, "rewards_prey.only_positive_rewards": False})]


@pytest.mark.parametrize("variant,n,ov", GAME_CASES)
def test_game_oracle_is_bit_exact_with_reference(variant, n, ov):
    """HighLevelGame / DecHighLevelGame (reference classes built around the reference LowLevelGame, low-level policy
    replaced by fixed actions) against oracle/game_oracle.py: every buffer after each of 6 steps."""
    import contextlib
    import io
    from oracle import game_oracle
    ll_ov = {"env.episode_length_s": 0.1}
    case = harness.build_case("low_level_game", n, seed=9, overrides=ll_ov)
    st_ref, st_or = harness.torch_state(case), harness.torch_state(case)
    harness.place_predators(st_ref, 9), harness.place_predators(st_or, 9)
    hs = torch.from_numpy(case["height_samples"].copy())
    ll_ref = ref_loader.make_ref_env("low_level_game", n, case["consts"], st_ref, height_samples=hs, cfg_overrides=ll_ov,
                                     init_levels=case["init_levels"])
    ll_ref.episode_length_buf[:] = torch.from_numpy(case["state"]["episode_length_buf"])
    tap = ref_loader.attach_tap(ll_ref)
    ref = ref_loader.make_ref_game(variant, ll_ref, tap, cfg_overrides=ov)
    ref.predator_pos = st_ref["root_states"][1::2, :3].clone()
    ll_or = harness.make_oracle(case, st_or)
    import copy
    cfg = copy.deepcopy(ref.cfg)
    orc = game_oracle.GameOracle(cfg, ll_or, variant)
    saw_reset = 0
    for step in range(1, 7):
        tables = harness.step_tables(case["seed"], step, n, ll_ref.num_obs)
        prey, pred, acts = harness.game_inputs(case, step, variant)
        tap.set_tables(tables)
        ref.ll_policy = lambda obs: acts.clone()
        with tap.active(), contextlib.redirect_stdout(io.StringIO()):
            if variant == "hl":
                ref.step(torch.cat((prey, pred), dim=1).clone())
            else:
                ref.step(pred.clone(), prey.clone())
        if variant == "hl":
            orc.step(torch.cat((prey, pred), dim=1).clone(), acts.clone(), tables)
        else:
            orc.step_dec(pred.clone(), prey.clone(), acts.clone(), tables)
        a, b = game_oracle.snapshot(ref), game_oracle.snapshot(orc)
        assert a.keys() == b.keys()
        for k in a:
            assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), f"{variant} step {step}: {k} differs from the reference"
        sa, sb = game_oracle.game_sums(ref), game_oracle.game_sums(orc)
        assert sa.keys() == sb.keys()
        for k in sa:
            assert torch.equal(sa[k], sb[k]), f"{variant} step {step}: episode sum {k}"
        la, lb = harness.snapshot(ll_ref), harness.snapshot(ll_or)
        for k in la:
            assert torch.equal(la[k], lb[k]), f"{variant} step {step}: low-level {k}"
        saw_reset += int(a["reset_buf"].sum())
        harness.perturb_state(st_ref, step, 5)
        harness.perturb_state(st_or, step, 5)
    assert saw_reset > 0


def test_product_cfgs_equal_reference_cfgs():
    """The cfg mirror (dict-spec built classes) must carry exactly the reference's values."""
    ref_loader.load_reference()
    from legged_gym.utils.task_registry import task_registry as ref_reg
    from legged_gym.utils.helpers import class_to_dict as ref_c2d
    from legged_games_gym_b200.envs import task_registry
    from legged_games_gym_b200.utils.helpers import class_to_dict
    for name in ("anymal_c_rough", "anymal_c_flat", "anymal_b", "a1", "cassie"):
        for mine, theirs in ((task_registry.env_cfgs[name], ref_reg.env_cfgs[name]),
                             (task_registry.train_cfgs[name], ref_reg.train_cfgs[name])):
            a, b = class_to_dict(mine), ref_c2d(theirs)
            a.pop("seed", None), b.pop("seed", None)
            assert a == b, f"cfg mismatch for {name}"
    from legged_games_gym_b200.envs import LowLevelGameCfg, LowLevelGamePPO
    assert class_to_dict(LowLevelGameCfg()) == {k: v for k, v in ref_c2d(ref_reg.env_cfgs["low_level_game"]).items() if k != "seed"}
    assert class_to_dict(LowLevelGamePPO()) == ref_c2d(ref_reg.train_cfgs["low_level_game"])
    for name in ("high_level_game", "dec_high_level_game"):
        a, b = class_to_dict(type(task_registry.env_cfgs[name])()), ref_c2d(ref_reg.env_cfgs[name])
        b.pop("seed", None)
        assert a == b, f"cfg mismatch for {name}"
        assert class_to_dict(type(task_registry.train_cfgs[name])()) == ref_c2d(ref_reg.train_cfgs[name])
    assert sorted(task_registry.task_classes) == sorted(ref_reg.task_classes)        # all eight tasks are registered


def test_lstm_plain_restatement_matches_aten_lstm():
    w = harness.lstm_weights()
    from oracle.legged_oracle import ActuatorLSTM
    net = ActuatorLSTM(w)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4096, 1, 2, generator=g)
    h, c = torch.randn(2, 4096, 8, generator=g) * 0.5, torch.randn(2, 4096, 8, generator=g) * 0.5
    a, ha, ca = net.forward(x, h.clone(), c.clone())
    b, hb, cb = net.forward_plain(x, h.clone(), c.clone())
    assert torch.allclose(a, b, rtol=1e-5, atol=5e-5) and torch.allclose(ha, hb, atol=1e-6) and torch.allclose(ca, cb, atol=1e-6)
    ref = torch.jit.load(ref_loader.REFERENCE_ROOT + "/resources/actuator_nets/anydrive_v3_lstm.pt")
    with torch.inference_mode():
        r, (hr, cr) = ref(x, (h.clone(), c.clone()))
    assert torch.equal(r, a) and torch.equal(hr, ha) and torch.equal(cr, ca)
