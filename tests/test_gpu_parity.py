"""GPU parity tests: the sm_100a kernels, called through the C ABI (ctypes -> liblgk.so), against the CPU oracle on
identical seeded inputs, and against the golden fixtures produced by the reference itself.

Bars (BASELINE.json north_star): bit-exact for reset / time-out masks, height-sample indices, episode counters,
terrain levels and command resampling under the shared Philox stream; rtol 1e-5 (atol 1e-5 x scale) for fp32
observations, rewards, torques; 1e-3 for policy outputs and GAE."""
import os

import numpy as np
import pytest
import torch

from oracle import harness, philox
from oracle.make_golden import CASES, GAME_CASES, GOLDEN_DIR, build, build_game
from tests.util import product_env, product_game, feeder_state, assert_snapshots_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def nat():
    from legged_games_gym_b200 import _native
    return _native


def stream():
    return torch.cuda.current_stream().cuda_stream


# ---------------------------------------------------------------------------------------------- RNG tap
@pytest.mark.parametrize("sid,count", [(philox.STREAM_CMD, 3), (philox.STREAM_RESET_DOF, 12), (philox.STREAM_RESET_ROOT, 8),
                                       (philox.STREAM_TERRAIN, 1), (philox.STREAM_OBS, 235), (philox.STREAM_OBS, 48)])
def test_rng_dump_bit_exact(sid, count):
    n, seed, step, off = 257, 0x1234567890ABCDEF, 77, 1000
    ids = np.arange(off, off + n)
    out = torch.empty(n, count, dtype=torch.int32, device=DEV)
    nat().check(nat().lib.lgk_rng_dump(seed, step, off, n, sid, count, 1, out.data_ptr(), stream()))
    if sid == philox.STREAM_OBS:
        u_want = philox.obs_uniforms(seed, step, ids, count)
    else:
        want = philox.raw_u32(seed, step, ids, sid, count)
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want)
        u_want = philox.uniforms(seed, step, ids, sid, count)
    u = torch.empty(n, count, dtype=torch.float32, device=DEV)
    nat().check(nat().lib.lgk_rng_dump(seed, step, off, n, sid, count, 0, u.data_ptr(), stream()))
    assert np.array_equal(u.cpu().numpy(), u_want)


def test_rng_normals_close():
    n, seed, step = 300, 5, 9
    z = torch.empty(n, 12, dtype=torch.float32, device=DEV)
    nat().check(nat().lib.lgk_rng_dump(seed, step, 0, n, philox.STREAM_ACT, 12, 2, z.data_ptr(), stream()))
    want = philox.normals(seed, step, np.arange(n), 12)
    assert np.allclose(z.cpu().numpy(), want, rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------------------------- height scan
@pytest.mark.parametrize("task,n", [("anymal_c_rough", 64), ("cassie", 100), ("anymal_c_rough", 4096)])
def test_height_scan_indices_bit_exact(task, n):
    case = harness.build_case(task, n, seed=21)
    # out-of-field roots exercise the index clip (LR:860-861): far negative, far positive, and huge
    r = case["state"]["root_states"]
    r[0, :2] = (-100., -100.)
    r[1, :2] = (500., 500.)
    r[2, :2] = (-24.96, 184.93)
    r[3, :2] = (3.0e9, -3.0e9)
    orc = harness.make_oracle(case)
    want_h = orc.get_heights()
    P = orc.num_height_points
    hs = torch.from_numpy(case["height_samples"]).to(DEV)
    m3 = torch.empty_like(hs)
    L = nat().lib
    nat().check(L.lgk_height_min3(hs.data_ptr(), m3.data_ptr(), hs.shape[0], hs.shape[1], stream()))
    # min3 against its definition (integer work: exact)
    h = case["height_samples"].astype(np.int32)
    ref3 = h.copy()
    ref3[:-1, :-1] = np.minimum(np.minimum(h[:-1, :-1], h[1:, :-1]), h[:-1, 1:])
    assert np.array_equal(m3.cpu().numpy().astype(np.int32), ref3)
    root = torch.from_numpy(case["state"]["root_states"]).to(DEV)
    pts = orc.height_points[0, :, :2].contiguous().to(DEV)
    out = torch.empty(n, P, device=DEV)
    px = torch.empty(n, P, dtype=torch.int32, device=DEV)
    py = torch.empty(n, P, dtype=torch.int32, device=DEV)
    t = case["cfg"].terrain
    nat().check(L.lgk_height_scan(root.data_ptr(), 1, 0, n, pts.data_ptr(), P, m3.data_ptr(), hs.shape[0], hs.shape[1],
                                  t.border_size, t.horizontal_scale, t.vertical_scale, out.data_ptr(), px.data_ptr(),
                                  py.data_ptr(), stream()))
    assert np.array_equal(px.cpu().numpy().reshape(-1), orc.last_px.numpy())
    assert np.array_equal(py.cpu().numpy().reshape(-1), orc.last_py.numpy())
    assert np.array_equal(out.cpu().numpy(), want_h.numpy())
    frac_clipped = float(((orc.last_px == 0) | (orc.last_px == hs.shape[0] - 2)).float().mean())
    assert 0.0 < frac_clipped < 0.5       # the clip path is exercised, but not only the clip path
    assert int(orc.last_py.max()) == hs.shape[1] - 2 and int(orc.last_py.min()) == 0


# ---------------------------------------------------------------------------------------------- torques
@pytest.mark.parametrize("ct", ["P", "V", "T"])
def test_pd_torques_bit_exact(ct):
    case = harness.build_case("a1", 200, seed=31, overrides={"control.control_type": ct})
    orc = harness.make_oracle(case)
    env, feeder = product_env(case)
    acts = torch.from_numpy(np.random.default_rng(1).normal(0, 40, (200, 12)).astype(np.float32))   # some beyond +-100
    acts[0, 0], acts[1, 1] = 250., -250.
    orc.last_dof_vel[:] = torch.from_numpy(np.random.default_rng(2).normal(0, 1, (200, 12)).astype(np.float32))
    env.last_dof_vel.copy_(orc.last_dof_vel.to(DEV))
    want = orc.compute_torques(torch.clip(acts, -100, 100))
    got = env._compute_torques(acts.to(DEV))
    assert np.array_equal(got.cpu().numpy(), want.numpy())


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_lstm_torques_and_state(variant):
    """the actuator-net kernels: 1 = one thread per (env, joint) sequence, 2 = role-split CTA (4 warps x 32 sequences),
    3 = one thread per sequence with the gate arithmetic on the packed fp32 pipe, 0 = auto (3 at this size)"""
    case = harness.build_case("anymal_c_flat", 323, seed=32)       # 3876 sequences: not a multiple of 32 or 128
    orc = harness.make_oracle(case)
    env, feeder = product_env(case)
    env._tq_params.lstm_variant = variant
    st_or = {"dof_state": orc.dof_state}
    worst = 0.0
    for k in range(6):
        acts = torch.from_numpy(np.random.default_rng(10 + k).normal(0, 1, (323, 12)).astype(np.float32))
        want = orc.compute_torques(acts).view(323, 12)
        got = env._compute_torques(acts.to(DEV))
        scale = float(want.abs().max())
        err = (got.cpu() - want).abs()
        worst = max(worst, float(err.max()))
        assert torch.all(err <= 1e-5 * scale + 1e-5 * want.abs()), f"substep {k}: max err {float(err.max()):.3e} (scale {scale:.1f})"
        assert torch.allclose(env.sea_hidden_state.cpu(), orc.sea_hidden_state, rtol=1e-5, atol=1e-5)
        assert torch.allclose(env.sea_cell_state.cpu(), orc.sea_cell_state, rtol=1e-5, atol=1e-5)
        noise = torch.from_numpy(np.random.default_rng(50 + k).normal(0, 0.3, tuple(orc.dof_state.shape)).astype(np.float32))
        orc.dof_state += noise
        feeder.dof_state += noise.to(DEV)
    print(f"LSTM torque max abs err over 6 sub-steps: {worst:.3e}")


# ---------------------------------------------------------------------------------------------- full step
STEP_CASES = [("anymal_c_flat", 64, None), ("anymal_c_rough", 256, None), ("anymal_c_rough", 1000, None),
              ("a1", 512, None), ("cassie", 130, None), ("anymal_b", 96, None),
              ("anymal_c_flat", 64, {"control.use_actuator_network": False}),
              ("a1", 160, {"commands.curriculum": True, "domain_rand.push_interval_s": 0.04, "env.episode_length_s": 0.1}),
              ("a1", 64, {"noise.add_noise": False, "rewards.only_positive_rewards": False}),
              # ragged / minimal batches: one env, a partial tile, N not a multiple of 4 (no 16-byte aligned rows -> the
              # non-TMA staging path), one env more than a tile
              ("a1", 1, None), ("anymal_c_rough", 33, None), ("anymal_c_flat", 3, None), ("cassie", 31, None),
              ("low_level_game", 200, None),
              ("low_level_game", 90, {"domain_rand.push_interval_s": 0.04, "env.episode_length_s": 0.1}),
              ("anymal_c_rough", 96, {"rewards.scales.base_height": -1.0, "rewards.scales.dof_vel": -1e-4,
                                      "rewards.scales.stand_still": -0.1, "rewards.scales.orientation": -1.0,
                                      "rewards.scales.termination": -5.0, "rewards.scales.feet_contact_forces": -0.01,
                                      "rewards.scales.dof_vel_limits": -0.5, "rewards.scales.torque_limits": -0.01,
                                      "rewards.scales.dof_pos_limits": -2.0, "rewards.max_contact_force": 1.5}),
              # configuration switches off the defaults (each pinned oracle == reference in tests/test_oracle_vs_reference.py)
              ("anymal_c_rough", 160, {"commands.heading_command": False, "env.episode_length_s": 0.1}),
              ("anymal_c_rough", 96, {"terrain.curriculum": False, "control.decimation": 2}),
              ("a1", 160, {"domain_rand.push_robots": False, "commands.resampling_time": 0.04}),
              ("a1", 96, {"terrain.measure_heights": False, "env.num_observations": 48})]


@pytest.mark.parametrize("task,n,ov", STEP_CASES)
def test_env_step_matches_oracle(task, n, ov):
    case = harness.build_case(task, n, seed=3, overrides=ov)
    st_or = harness.torch_state(case)
    orc = harness.make_oracle(case, st_or)
    env, feeder = product_env(case)
    st_gpu = feeder_state(feeder)
    assert env.reward_names == orc.reward_names
    for step in range(1, 7):
        tables = harness.step_tables(case["seed"], step, n, orc.num_obs)
        acts = torch.from_numpy(np.random.default_rng(step).normal(0, 1, (n, 12)).astype(np.float32))
        orc.step(acts.clone(), tables)
        env.step(acts.to(DEV))
        torch.cuda.synchronize()
        ids = env.reset_env_ids[: int(env.reset_count.item())].cpu().numpy()
        assert np.array_equal(ids, orc.reset_buf.nonzero().flatten().numpy()), "reset id list must equal nonzero() order"
        scales = {"torques": max(1.0, float(orc.torques.abs().max())), "sea_h": 1.0, "sea_c": 1.0}
        assert_snapshots_close(harness.snapshot(env), harness.snapshot(orc), step, atol_scale=scales,
                               heading_command=bool(env.cfg.commands.heading_command))
        noise = harness.make_noise(case, step, 5)
        harness.apply_noise(st_or, noise)
        harness.apply_noise(st_gpu, noise)


@pytest.mark.parametrize("task,n", [("anymal_c_rough", 300), ("anymal_c_rough", 8000), ("cassie", 77), ("a1", 1000),
                                    ("anymal_c_flat", 2100)])
def test_env_step_cuda_graph(task, n):
    """Same parity bar with the whole step replayed as one CUDA graph (device-side step counter); pushes every 2nd step so
    the device-derived push flag is exercised; 8000 envs = 250 tiles, more than one tile per persistent K1 CTA on a
    148-SM part only when few CTAs are resident -- LGK_K1_CTAS_PER_SM=1 in test_k1_persistent_loop covers that."""
    ov = {"domain_rand.push_interval_s": 0.04}
    case = harness.build_case(task, n, seed=6, overrides=ov)
    st_or = harness.torch_state(case)
    orc = harness.make_oracle(case, st_or)
    env, feeder = product_env(case, graph=True)
    assert env._graph_ok
    st_gpu = feeder_state(feeder)
    for step in range(1, 8):
        tables = harness.step_tables(case["seed"], step, n, orc.num_obs)
        acts = torch.from_numpy(np.random.default_rng(step).normal(0, 1, (n, 12)).astype(np.float32))
        orc.step(acts.clone(), tables)
        env.step(acts.to(DEV))
        torch.cuda.synchronize()
        assert int(env._step_counter_dev[0].item()) == step == env.common_step_counter
        assert_snapshots_close(harness.snapshot(env), harness.snapshot(orc), step,
                               atol_scale={"torques": max(1.0, float(orc.torques.abs().max()))},
                               heading_command=bool(env.cfg.commands.heading_command))
        noise = harness.make_noise(case, step, 5)
        harness.apply_noise(st_or, noise)
        harness.apply_noise(st_gpu, noise)
    assert env._graph is not None


GAME_STEP_CASES = [("hl", 96, None, None), ("hl", 200, {"env.env_radius": 40.0, "rewards.scales.termination": -2.0, "rewards.only_positive_rewards": False}, None),
              ("dec", 96, {"env.episode_length_s": 0.08}, {"terrain.mesh_type": "plane", "terrain.curriculum": False}),
              ("dec", 160, {"rewards_prey.scales.termination": -3.0, "rewards_prey.only_positive_rewards": False},
               {"terrain.mesh_type": "plane", "terrain.curriculum": False}),
              # the reference's own num_envs (high_level_game_flat_config.py:10) and a single partial warp
              ("hl", 2000, None, None), ("dec", 33, None, {"terrain.mesh_type": "plane", "terrain.curriculum": False}), ("hl", 1, None, None)]


@pytest.mark.parametrize("variant,n,ov,ll_ov", GAME_STEP_CASES)
def test_high_level_games_match_oracle(variant, n, ov, ll_ov):
    """HighLevelGame / DecHighLevelGame (lgk_game_step on top of the LowLevelGame kernels) against oracle/game_oracle.py,
    which tests/test_oracle_vs_reference.py pins bit-exactly to the reference classes: masks, counters and the sensing
    bits exact, fp32 within 1e-5, over 6 steps with captures, time-outs, low-level dones and occluded predators."""
    import copy
    from oracle import game_oracle
    ll_over = {"env.episode_length_s": 0.1}
    ll_over.update(ll_ov or {})
    case = harness.build_case("low_level_game", n, seed=9, overrides=ll_over)
    case["state"]["root_states"] = case["state"]["root_states"].copy()
    st_or = harness.torch_state(case)
    harness.place_predators(st_or, 9)
    case["state"]["root_states"] = st_or["root_states"].numpy().copy()
    acts_box = {}
    game, feeder = product_game(case, variant, ov, ll_policy=lambda obs: acts_box["a"])
    st_gpu = feeder_state(feeder)
    game.predator_pos.copy_(st_gpu["root_states"][1::2, :3])
    case_or = dict(case, cfg=copy.deepcopy(game.ll_env.cfg))
    ll_or = harness.make_oracle(case_or, st_or)
    orc = game_oracle.GameOracle(copy.deepcopy(game.cfg), ll_or, variant)
    saw_reset = saw_occluded = 0
    for step in range(1, 7):
        tables = harness.step_tables(case["seed"], step, n, ll_or.num_obs)
        prey, pred, acts = harness.game_inputs(case, step, variant)
        acts_box["a"] = acts.to(DEV)
        if variant == "hl":
            orc.step(torch.cat((prey, pred), dim=1).clone(), acts.clone(), tables)
            game.step(torch.cat((prey, pred), dim=1).to(DEV))
        else:
            orc.step_dec(pred.clone(), prey.clone(), acts.clone(), tables)
            game.step(pred.to(DEV), prey.to(DEV))
        torch.cuda.synchronize()
        a, b = game_oracle.snapshot(game), game_oracle.snapshot(orc)
        a = {k: v for k, v in a.items() if k in b}
        assert_snapshots_close(a, b, f"{variant}/{step}", exact=("reset_buf", "time_out_buf", "episode_length_buf", "curr_episode_step"))
        obs_a, obs_b = (a["obs_buf"], b["obs_buf"]) if variant == "hl" else (a["obs_buf_prey"], b["obs_buf_prey"])
        assert torch.equal(obs_a[:, 12:16], obs_b[:, 12:16]), "field-of-view sensing bits must be exact"
        sa, sb = game_oracle.game_sums(game), game_oracle.game_sums(orc)
        assert sa.keys() == sb.keys()
        assert_snapshots_close(sa, sb, f"{variant}/{step} sums")
        assert_snapshots_close(harness.snapshot(game.ll_env), harness.snapshot(ll_or), f"{variant}/{step} low-level",
                               atol_scale={"torques": max(1.0, float(ll_or.torques.abs().max()))})
        saw_reset += int(b["reset_buf"].sum())
        saw_occluded += int((obs_b[:, 15] == 0).sum())
        noise = harness.make_noise(case, step, 5)
        harness.apply_noise(st_or, noise)
        harness.apply_noise(st_gpu, noise)
    assert saw_reset > 0 and saw_occluded > 0


@pytest.mark.parametrize("name", sorted(GAME_CASES))
def test_high_level_games_match_reference_fixture(name):
    """The product games against tests/golden/*high_level_game*.npz = outputs of the UNMODIFIED reference classes."""
    from oracle import game_oracle
    from tests.test_golden import check_game_snapshot
    spec = GAME_CASES[name]
    fx = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case, gcfg, _ = build_game(spec)
    acts_box = {}
    game, feeder = product_game(case, spec["variant"], spec["overrides"], ll_policy=lambda obs: acts_box["a"])
    st_gpu = feeder_state(feeder)
    game.predator_pos.copy_(st_gpu["root_states"][1::2, :3])
    for step in range(1, spec["steps"] + 1):
        prey, pred, acts = (torch.from_numpy(fx[f"s{step}_{k}"].copy()).to(DEV) for k in ("prey", "pred", "actions"))
        acts_box["a"] = acts
        if spec["variant"] == "hl":
            game.step(torch.cat((prey, pred), dim=1))
        else:
            game.step(pred, prey)
        torch.cuda.synchronize()
        snap = {k: v for k, v in game_oracle.snapshot(game).items() if f"s{step}_{k}" in fx}
        assert {k for k in fx.files if k.startswith(f"s{step}_ex_")} <= {f"s{step}_{k}" for k in snap}
        check_game_snapshot(fx, step, snap, game_oracle.game_sums(game), name, 1e-5, 1e-5)
        assert np.allclose(game.ll_env.obs_buf.cpu().numpy(), fx[f"s{step}_ll_obs_buf"], rtol=1e-5, atol=1e-5 * 100)
        harness.apply_noise(st_gpu, {k: fx[f"s{step}_noise_{k}"] for k in ("dof_state", "contact_forces", "root_vel", "root_xy")})


@pytest.mark.parametrize("name", sorted(CASES))
def test_env_step_matches_reference_fixture(name):
    spec = CASES[name]
    fx = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case = build(spec)
    for k in ("root_states", "dof_state", "contact_forces", "episode_length_buf"):
        case["state"][k] = fx["in_" + k]
    env, feeder = product_env(case)
    st_gpu = feeder_state(feeder)
    for step in range(1, spec["steps"] + 1):
        env.step(torch.from_numpy(fx[f"s{step}_actions"].copy()).to(DEV))
        torch.cuda.synchronize()
        snap = harness.snapshot(env)
        want = {k: fx[f"s{step}_{k}"] for k in snap if f"s{step}_{k}" in fx.files}
        missing = [k[len(f"s{step}_"):] for k in fx.files if k.startswith(f"s{step}_") and "noise_" not in k
                   and k != f"s{step}_actions" and k[len(f"s{step}_"):] not in snap]
        assert not missing, f"product env lacks outputs the reference has: {missing}"
        scales = {"torques": max(1.0, float(np.abs(fx[f"s{step}_torques"]).max()))}
        assert_snapshots_close(snap, want, step, atol_scale=scales, heading_command=bool(env.cfg.commands.heading_command))
        harness.apply_noise(st_gpu, {k: fx[f"s{step}_noise_{k}"] for k in ("dof_state", "contact_forces", "root_vel", "root_xy")})


def test_explicit_reset_idx_and_reset():
    """BaseTask.reset(): reset_idx(arange(N)) then one zero-action step (base_task.py:114-118)."""
    case = harness.build_case("anymal_c_rough", 200, seed=8)
    st_or = harness.torch_state(case)
    orc = harness.make_oracle(case, st_or)
    env, feeder = product_env(case)
    # oracle: same sequence with explicit tables; the reset uses step counter 0, the step uses 1
    orc.tables = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in harness.step_tables(case["seed"], 0, 200, orc.num_obs).items()}
    orc.reset_buf = torch.ones(200, dtype=torch.bool)
    orc.reset_idx(torch.arange(200))
    obs_o, *_ = orc.step(torch.zeros(200, 12), harness.step_tables(case["seed"], 1, 200, orc.num_obs))
    obs_g, _ = env.reset()
    torch.cuda.synchronize()
    assert_snapshots_close(harness.snapshot(env), harness.snapshot(orc), 1,
                           atol_scale={"torques": max(1.0, float(orc.torques.abs().max()))})
    assert torch.allclose(obs_g.cpu(), obs_o, rtol=1e-5, atol=1e-5)


def test_user_reward_term_runs_split_phases():
    """A subclass adds a torch-written _reward_<name>: the step runs PRE, the Python term, POST (LR:199-203 order)."""
    from legged_games_gym_b200.envs import LeggedRobot, task_registry
    from legged_games_gym_b200.envs.a1.a1_config import A1RoughCfg

    class MyEnv(LeggedRobot):
        def _reward_zz_height_bonus(self):
            return torch.exp(-torch.square(self.root_states[:, 2] - 0.4))

    case = harness.build_case("a1", 96, seed=4)
    case["cfg"].rewards.scales.zz_height_bonus = 0.3
    task_registry.register("a1_custom", MyEnv, A1RoughCfg(), None)
    case["task"] = "a1_custom"
    st_or = harness.torch_state(case)
    orc = harness.make_oracle(case, st_or)
    orc._r_zz_height_bonus = lambda: torch.exp(-torch.square(orc.root_states[:, 2] - 0.4))
    env, feeder = product_env(case)
    assert env._python_reward_names == ["zz_height_bonus"]
    for step in range(1, 4):
        tables = harness.step_tables(case["seed"], step, 96, orc.num_obs)
        acts = torch.from_numpy(np.random.default_rng(step).normal(0, 1, (96, 12)).astype(np.float32))
        orc.step(acts.clone(), tables)
        env.step(acts.to(DEV))
        torch.cuda.synchronize()
        assert_snapshots_close(harness.snapshot(env), harness.snapshot(orc), step,
                               atol_scale={"torques": max(1.0, float(orc.torques.abs().max()))},
                               heading_command=bool(env.cfg.commands.heading_command))


def test_reset_idx_override_runs_where_the_reference_calls_it():
    """reset_idx is an extension point of the reference (README.md:56-66; its own Anymal overrides it, ANY:56-60) and
    LR:128-129 calls it every step between compute_reward and compute_observations.  A subclass that zeroes an extra
    buffer there must see exactly the reference's env_ids at exactly that point, and the step must still equal the oracle."""
    from legged_games_gym_b200.envs import LeggedRobot, task_registry
    from legged_games_gym_b200.envs.a1.a1_config import A1RoughCfg
    seen = []

    class MyEnv(LeggedRobot):
        def _init_buffers(self):
            super()._init_buffers()
            self.steps_since_reset = torch.zeros(self.num_envs, device=self.device)

        def reset_idx(self, env_ids):
            if self.init_done and len(env_ids) > 0:
                # the reference's ordering: rewards of this step are final, observations are not yet rebuilt
                seen.append((env_ids.clone(), self.rew_buf.clone(), self.episode_length_buf[env_ids].clone()))
            super().reset_idx(env_ids)
            self.steps_since_reset[env_ids] = 0.

    ov = {"env.episode_length_s": 0.1, "domain_rand.push_interval_s": 0.04, "commands.curriculum": True,
          "rewards.scales.termination": -2.0}
    case = harness.build_case("a1", 200, seed=12, overrides=ov)
    task_registry.register("a1_reset_override", MyEnv, A1RoughCfg(), None)
    case["task"] = "a1_reset_override"
    st_or = harness.torch_state(case)
    orc = harness.make_oracle(case, st_or)
    env, feeder = product_env(case, graph=True)
    assert env._reset_overridden and not env._graph_ok          # the split step is not graph-replayed
    st_gpu = feeder_state(feeder)
    want_since = torch.zeros(200)
    n_reset = 0
    for step in range(1, 9):
        tables = harness.step_tables(case["seed"], step, 200, orc.num_obs)
        acts = torch.from_numpy(np.random.default_rng(step).normal(0, 1, (200, 12)).astype(np.float32))
        orc.step(acts.clone(), tables)
        env.steps_since_reset += 1
        want_since += 1
        before = len(seen)
        env.step(acts.to(DEV))
        torch.cuda.synchronize()
        ids = orc.reset_buf.nonzero().flatten()
        want_since[ids] = 0
        assert_snapshots_close(harness.snapshot(env), harness.snapshot(orc), step,
                               atol_scale={"torques": max(1.0, float(orc.torques.abs().max()))},
                               heading_command=bool(env.cfg.commands.heading_command))
        assert torch.equal(env.steps_since_reset.cpu(), want_since)
        if len(ids) > 0:
            assert len(seen) == before + 1, "the override must be called once per step with resets"
            got_ids, rew_at_call, ep_at_call = seen[-1]
            assert torch.equal(got_ids.cpu(), ids)                                  # LR:128 nonzero() order
            assert torch.allclose(rew_at_call.cpu(), orc.rew_buf, rtol=1e-5, atol=1e-5)   # final (clip + termination term)
            assert bool((ep_at_call > 0).all())                                    # not yet zeroed when the override starts
            n_reset += len(ids)
        noise = harness.make_noise(case, step, 5)
        harness.apply_noise(st_or, noise)
        harness.apply_noise(st_gpu, noise)
    assert n_reset > 20


def test_step_is_deterministic_run_to_run():
    """compute-sanitizer is closed on the GPU pool, so the race evidence for K1 (shared-memory tiles re-used across the
    persistent loop, generic-proxy writes feeding async-proxy bulk stores, the warp-cooperative reset tail, named barriers)
    is behavioural: two envs fed the same state must agree bit for bit over steps with resets, pushes and resampling, on
    the fast path, the ragged-tile path and the two-actor path.  (extras["episode"] sums go through float atomics whose
    order is free: compared at 1e-6.)"""
    for task, n in (("anymal_c_rough", 148 * 32 + 96), ("a1", 4003), ("low_level_game", 1501)):
        ov = {"env.episode_length_s": 0.08, "domain_rand.push_interval_s": 0.04, "commands.resampling_time": 0.06}
        case = harness.build_case(task, n, seed=13, overrides=ov)
        runs = []
        for rep in range(2):
            env, feeder = product_env(case)
            st = feeder_state(feeder)
            snaps = []
            for step in range(1, 6):
                acts = torch.from_numpy(np.random.default_rng(step).normal(0, 1, (n, 12)).astype(np.float32)).to(DEV)
                env.step(acts)
                torch.cuda.synchronize()
                snaps.append(harness.snapshot(env))
                harness.apply_noise(st, harness.make_noise(case, step, 5))
            runs.append(snaps)
        assert sum(int(s["reset_buf"].sum()) for s in runs[0]) > n // 10
        for step, (a, b) in enumerate(zip(*runs), 1):
            for k in a:
                if k.startswith("ex_"):
                    assert torch.allclose(a[k], b[k], rtol=1e-5, atol=1e-6), f"{task} step {step}: {k}"
                else:
                    assert torch.equal(a[k], b[k]), f"{task} step {step}: {k} differs between two identical runs"


@pytest.mark.parametrize("graph", [False, True])
def test_fused_finalize_equals_the_two_calls(graph):
    """lgk_post_physics_finalize (the finalize pass riding in K2's grid, step handed over in word [1] of the counter)
    against lgk_post_physics + lgk_finalize_step on identical envs: state, reset id list and count, extras and the device
    step counter, bit for bit (extras go through float atomics: 1e-6).  4096 envs take the register-resident flag path,
    9000 (not a multiple of 16), 16 384 and 20 000 the longer sweeps (single vectors / one batch of eight / a batch plus
    singles per thread); flat tasks -- no K2 -- take the entry point's fallback to the two calls on both envs."""
    for task, n in (("anymal_c_rough", 4096), ("anymal_c_rough", 9000), ("a1", 16384), ("anymal_c_rough", 20000),
                    ("anymal_c_flat", 2100), ("cassie", 77)):
        ov = {"env.episode_length_s": 0.08, "domain_rand.push_interval_s": 0.04, "commands.resampling_time": 0.06}
        case = harness.build_case(task, n, seed=21, overrides=ov)
        runs = []
        for fuse in (True, False):
            env, feeder = product_env(case, graph=graph)
            env.fuse_finalize = fuse
            st = feeder_state(feeder)
            snaps = []
            for step in range(1, 7):
                acts = torch.from_numpy(np.random.default_rng(step).normal(0, 1, (n, 12)).astype(np.float32)).to(DEV)
                _, _, _, _, extras = env.step(acts)
                torch.cuda.synchronize()
                snap = harness.snapshot(env)
                cnt = int(env.reset_count.item())
                snap["reset_count"] = env.reset_count.clone()
                snap["reset_ids"] = env.reset_env_ids[:cnt].clone()
                snap["counter"] = env._step_counter_dev[:1].clone()
                snap["ex_means"] = env._episode_means.clone()
                snap["time_outs"] = extras["time_outs"].clone()
                assert torch.equal(snap["reset_ids"].long(), env.reset_buf.nonzero().flatten())
                snaps.append(snap)
                harness.apply_noise(st, harness.make_noise(case, step, 5))
            assert int(env._step_counter_dev[0].item()) == 6
            runs.append(snaps)
        assert sum(int(s["reset_count"]) for s in runs[0]) > n // 10
        for step, (a, b) in enumerate(zip(*runs), 1):
            for k in a:
                if k.startswith("ex_"):
                    assert torch.allclose(a[k], b[k], rtol=1e-5, atol=1e-6), f"{task} {n} step {step}: {k}"
                else:
                    assert torch.equal(a[k], b[k]), f"{task} {n} step {step}: {k} differs between the fused and the two-call step"


def test_k1_persistent_loop_many_tiles_per_cta():
    """K1's CTAs are persistent: with one resident CTA per SM (tuning knob of the launcher) 20 000 envs are 625 tiles on at
    most 148 CTAs, i.e. up to five tiles per CTA incl. a partial last tile -- shared-memory reuse, mbarrier phases and the
    L2 prefetch of the next tile must not change a bit against the default launch."""
    import subprocess, sys, json
    code = (
        "import sys, torch, numpy as np; sys.path.insert(0, %r)\n"
        "from oracle import harness\n"
        "from tests.util import product_env\n"
        "case = harness.build_case('anymal_c_rough', 20004, seed=5, overrides={'env.episode_length_s': 0.1})\n"
        "env, feeder = product_env(case)\n"
        "for step in range(1, 4):\n"
        "    a = torch.from_numpy(np.random.default_rng(step).normal(0, 1, (20004, 12)).astype(np.float32)).cuda()\n"
        "    env.step(a)\n"
        "torch.cuda.synchronize()\n"
        "snap = harness.snapshot(env)\n"
        "torch.save({k: v for k, v in snap.items() if not k.startswith('ex_')}, sys.argv[1])\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    outs = []
    for per_sm in ("1", "7"):
        path = f"/tmp/lgk_k1_persist_{per_sm}.pt"
        env = dict(os.environ, LGK_K1_CTAS_PER_SM=per_sm)
        r = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(torch.load(path))
    a, b = outs
    assert a.keys() == b.keys() and int(a["reset_buf"].sum()) > 100
    for k in a:
        assert torch.equal(a[k], b[k]), f"{k} differs between 1 and 7 resident K1 CTAs per SM"


# ---------------------------------------------------------------------------------------------- full size
FULL_SIZE_CASES = [
    # BASELINE configs[1]: 128 K1 tiles, one wave
    ("anymal_c_rough", 4096, {"env.episode_length_s": 0.1}, 4),
    # BASELINE configs[2]: a1, pushes every 2nd step, command resampling every 3rd, short episodes (time-outs + resets)
    ("a1", 16384, {"domain_rand.push_interval_s": 0.04, "commands.resampling_time": 0.06, "env.episode_length_s": 0.1}, 4),
    # BASELINE configs[4]: 2048 K1 tiles (several waves per SM), persistent K2 over 65 536 envs
    ("anymal_c_rough", 65536, {"env.episode_length_s": 0.06}, 3),
]


@pytest.mark.parametrize("task,n,ov,steps", FULL_SIZE_CASES)
def test_full_size_step_matches_oracle(task, n, ov, steps):
    """The full step at the BASELINE sizes against the CPU oracle (the same bars as the small cases: bit-exact masks,
    counters, reset rows and resampled commands, 1e-5 on fp32), the whole step replayed as one CUDA graph from the second
    step on like bench.py does, plus the size-independent invariants of the reset path."""
    case = harness.build_case(task, n, seed=2, overrides=ov)
    st_or = harness.torch_state(case)
    orc = harness.make_oracle(case, st_or)
    env, feeder = product_env(case, graph=True)
    st_gpu = feeder_state(feeder)
    n_reset = n_resampled = 0
    for step in range(1, steps + 1):
        tables = harness.step_tables(case["seed"], step, n, orc.num_obs)
        acts = torch.from_numpy(np.random.default_rng(step).normal(0, 1, (n, 12)).astype(np.float32))
        cmd_before = orc.commands.clone()
        orc.step(acts.clone(), tables)
        obs, _, rew, reset, extras = env.step(acts.to(DEV))
        torch.cuda.synchronize()
        scales = {"torques": max(1.0, float(orc.torques.abs().max())), "sea_h": 1.0, "sea_c": 1.0}
        assert_snapshots_close(harness.snapshot(env), harness.snapshot(orc), step, atol_scale=scales,
                               heading_command=bool(env.cfg.commands.heading_command))
        cnt = int(env.reset_count.item())
        ids = env.reset_env_ids[:cnt].long()
        assert torch.equal(ids, reset.nonzero().flatten()), "ids ascending == nonzero()"
        assert torch.all(env.episode_length_buf[ids] == 0) and torch.all(env.episode_length_buf[~reset] > 0)
        assert torch.equal(env.last_actions, env.actions)
        assert torch.equal(env.last_dof_vel, env.dof_vel) and torch.all(env.dof_vel[reset] == 0)
        assert torch.all(env.time_out_buf <= reset)
        for k, v in env.episode_sums.items():
            assert torch.all(v[reset] == 0), k
        n_reset += cnt
        n_resampled += int((orc.commands[:, :2] != cmd_before[:, :2]).any(dim=1).sum())
        noise = harness.make_noise(case, step, 5)
        harness.apply_noise(st_or, noise)
        harness.apply_noise(st_gpu, noise)
    assert env._graph is not None
    assert n_reset > n // 50 and n_resampled > 0, "the case must exercise resets and command resampling"
    assert torch.equal(env._get_heights().cpu(), orc.get_heights())


def test_domain_randomisation_props_reach_the_backend():
    """a1 (BASELINE configs[2]): friction buckets and added base mass are drawn at construction (LR:261-283, 316-327)."""
    case = harness.build_case("a1", 256, seed=4, overrides={"domain_rand.randomize_base_mass": True})
    env, feeder = product_env(case)
    lo, hi = env.cfg.domain_rand.friction_range
    assert tuple(env.friction_coeffs.shape) == (256, 1, 1)
    assert float(env.friction_coeffs.min()) >= lo and float(env.friction_coeffs.max()) <= hi
    assert env.friction_coeffs.unique().numel() <= 64                 # 64 buckets
    mlo, mhi = env.cfg.domain_rand.added_mass_range
    assert tuple(env.added_base_mass.shape) == (256,) and float(env.added_base_mass.min()) >= mlo and float(env.added_base_mass.max()) <= mhi
    assert feeder.friction_coeffs is env.friction_coeffs and feeder.added_base_mass is env.added_base_mass


@pytest.mark.parametrize("mode", ["unified", "explicit", "explicit_indexed_rows", "memcpy"])
def test_host_sim_state_step_replays_inside_the_graph(mode, monkeypatch):
    """bench.py's e2e leg: sim state in PINNED host memory, the whole step replayed as one CUDA graph.  `unified`: the
    kernels (K1's TMA bulk copies included) work directly on the pinned buffers over the unified address space;
    `explicit*`: device twins refreshed / written back by copy kernels, zero-copy torque sub-steps; `memcpy`: plain
    cudaMemcpyAsync nodes.  The host state changes between steps, pushes happen every second step, and everything must
    equal an env whose state lives on the device."""
    import bench
    bench.USE_GRAPH = True
    monkeypatch.setenv("LGK_HOST_UNIFIED", "1" if mode == "unified" else "0")
    monkeypatch.setenv("LGK_HOST_ZERO_COPY", "0" if mode == "memcpy" else "1")
    torch.manual_seed(0)                         # initial terrain levels come from torch's generator (LR:762)
    env_h, fh = bench.make_env(512, DEV, host_sim=True)
    torch.manual_seed(0)
    env_d, fd = bench.make_env(512, DEV, host_sim=False)
    assert fh.unified == (mode == "unified") and fh.zero_copy == (mode != "memcpy")
    fh.indexed_rows = mode == "explicit_indexed_rows"
    if mode == "memcpy":
        fh.KERNEL_COPY_MAX_BYTES = 0
    if mode == "unified":
        assert env_h.root_states.is_cuda and env_h.root_states.data_ptr() == fh.h_root.data_ptr()
    for e in (env_h, env_d):                     # pushes at steps 2 and 4 (whole-tile root write-back of K1)
        e.cfg.domain_rand.push_interval = 2
        e._params.push_interval = 2
    acts = fd.synthetic_actions
    g = torch.Generator().manual_seed(5)
    for step in range(5):
        env_h.step(acts)
        env_d.step(acts)
        torch.cuda.synchronize()
        assert torch.equal(env_h.obs_buf, env_d.obs_buf) and torch.equal(env_h.rew_buf, env_d.rew_buf), step
        assert torch.equal(env_h.reset_buf, env_d.reset_buf)
        assert torch.equal(fh.h_torques, env_d.torques.cpu())                       # the host's copy of the last sub-step's torques
        # the "simulator" moves: new joint velocities on the host side / on the device side
        dv = torch.randn(fh.h_dof.shape[0], generator=g) * 0.1
        fh.h_dof[:, 1] += dv
        fh.refresh_dof_state_tensor()            # a sim step ends with a refresh (LR:96); later refreshes replay in the graph
        fd.dof_state[:, 1] += dv.to(DEV)
        # resets and pushes were written back into the host copy of the sim state
        assert torch.equal(fh.h_root, env_d.root_states.cpu())
    assert env_h._graph is not None and env_d._graph is not None
    assert fh.h2d_bytes > 0 and fh.d2h_bytes > 0


def test_host_result_mirror_delivers_every_step_in_order():
    """bench.py's pipelined e2e loop: step k's observations / rewards / reset flags are downloaded on a copy stream while
    step k+1 runs.  Each wait(k) must return exactly what the env held after step k (the env rewrites its buffers in
    place, so a mirror that read them late would return step k+1's values)."""
    import bench
    from legged_games_gym_b200.sim.result_mirror import HostResultMirror
    bench.USE_GRAPH = True
    env, f = bench.make_env(1024, DEV, host_sim=True)
    mirror = HostResultMirror(env, depth=2)
    acts = f.synthetic_actions
    want = []
    g = torch.Generator().manual_seed(7)
    for k in range(7):
        env.step(acts)
        assert mirror.push() == k
        want.append((env.obs_buf.clone(), env.rew_buf.clone(), env.reset_buf.clone()))      # stream-ordered after the step
        f.h_dof[:, 1] += torch.randn(f.h_dof.shape[0], generator=g) * 0.1                  # the host "simulator" moves on
        if k > 0:
            got = mirror.wait(k - 1)
            o, r, z = want[k - 1]
            assert got["obs_buf"].is_pinned() and not got["obs_buf"].is_cuda
            assert torch.equal(got["obs_buf"], o.cpu()) and torch.equal(got["rew_buf"], r.cpu()) and torch.equal(got["reset_buf"], z.cpu()), k
    got = mirror.wait(6)
    assert torch.equal(got["obs_buf"], want[6][0].cpu())
    assert not torch.equal(want[5][0], want[6][0])                     # the steps did differ
    with pytest.raises(IndexError):
        mirror.wait(3)                                                 # slot long since reused
    assert mirror.bytes_per_push == env.obs_buf.numel() * 4 + env.rew_buf.numel() * 4 + env.reset_buf.numel()
    mirror.drain()


def feeder_actions(n, step):
    return torch.from_numpy(np.random.default_rng(100 + step).normal(0, 1, (n, 12)).astype(np.float32)).to(DEV)


# ---------------------------------------------------------------------------------------------- rsl_rl pieces
@pytest.mark.parametrize("T,N", [(24, 4096), (24, 100), (5, 33)])
def test_gae_matches_restated_rsl_rl(T, N):
    from oracle.rsl_oracle import compute_returns
    g = torch.Generator().manual_seed(0)
    rewards, values = torch.randn(T, N, 1, generator=g), torch.randn(T, N, 1, generator=g)
    dones = (torch.rand(T, N, 1, generator=g) < 0.1).to(torch.uint8)
    last = torch.randn(N, 1, generator=g)
    want_r, want_a = compute_returns(rewards, values, dones, last, 0.99, 0.95)
    d = lambda t: t.to(DEV).contiguous()
    r, v, dn, lv = d(rewards), d(values), d(dones), d(last)
    ret, adv = torch.empty_like(r), torch.empty_like(r)
    scratch = torch.zeros(4, dtype=torch.float64, device=DEV)
    nat().check(nat().lib.lgk_gae(r.data_ptr(), v.data_ptr(), dn.data_ptr(), lv.data_ptr(), T, N, 0.99, 0.95,
                                  ret.data_ptr(), adv.data_ptr(), scratch.data_ptr(), stream()))
    assert torch.allclose(ret.cpu(), want_r, rtol=1e-3, atol=1e-3)      # north_star tolerance for GAE returns
    assert torch.allclose(adv.cpu(), want_a, rtol=1e-3, atol=1e-3)
    assert torch.allclose(ret.cpu(), want_r, rtol=1e-5, atol=1e-5)      # and in fact fp32-tight


# variant 2 = tcgen05 kernel (TF32 operands, FP32 accumulation in TMEM), 1 = FP32 FFMA path; (96, 64, 32) does not fit the
# tensor-core kernel's tiling and must take the FP32 path on its own under variant 0 (auto)
@pytest.mark.parametrize("n,nobs,hidden,variant", [
    (4096, 235, (512, 256, 128), 2), (100, 48, (128, 64, 32), 2), (777, 169, (512, 256, 128), 2), (129, 235, (256, 128, 64), 2),
    (4096, 235, (512, 256, 128), 1), (100, 48, (128, 64, 32), 1), (300, 48, (96, 64, 32), 0)])
def test_policy_act_matches_restated_rsl_rl(n, nobs, hidden, variant):
    from oracle.rsl_oracle import ActorCriticOracle
    from legged_games_gym_b200.rsl_rl.modules import ActorCritic
    torch.manual_seed(0)
    nat().lib.lgk_policy_set_variant(variant)
    orc = ActorCriticOracle(nobs, nobs, 12, hidden, hidden)
    with torch.no_grad():
        orc.std.copy_(torch.linspace(0.5, 1.5, 12))
    ac = ActorCritic(nobs, nobs, 12, list(hidden), list(hidden)).to(DEV)
    ac.load_state_dict(orc.state_dict())
    obs = torch.randn(n, nobs) * 2
    seed, step = 17, 5
    eps = torch.from_numpy(philox.normals(seed, step, np.arange(n), 12))
    a, v, lp, mu, sg = orc.act(obs, obs, eps)
    ac.set_rng(seed, step)
    with torch.inference_mode():      # rollout context of rsl_rl's runner: the fused kernel path
        o = obs.to(DEV)
        out = ac.act_and_evaluate(o, o)      # PPO.act's call: actor + critic in one launch
        got_a, got_v, got_lp = out["actions"], out["values"], ac.get_actions_log_prob(out["actions"])
        assert torch.equal(ac.act(o), got_a)   # the actor-only launch (ActorCritic.act) draws the same actions
    torch.cuda.synchronize()
    tol = dict(rtol=1e-3, atol=1e-3)
    assert torch.allclose(ac.action_mean.cpu(), mu, **tol)
    assert torch.allclose(got_a.cpu(), a, **tol)
    assert torch.allclose(got_v.cpu(), v, **tol)
    assert torch.allclose(ac.action_std.cpu(), sg, **tol)
    # log-prob = sum over 12 actions: the per-action terms are checked too, so the bar is 1e-3 on the sum AND on its parts
    assert torch.allclose(got_lp.cpu(), lp, rtol=1e-3, atol=1e-3)
    sd = ac.action_std.cpu()
    per_action = -((got_a.cpu() - ac.action_mean.cpu()) ** 2) / (2 * sd ** 2) - sd.log() - 0.9189385332046727
    want_pa = -((a - mu) ** 2) / (2 * sg ** 2) - sg.log() - 0.9189385332046727
    assert torch.allclose(per_action, want_pa, rtol=1e-3, atol=1e-3)
    # a second call reuses the packed weight image (same torch version counters) and must reproduce the first
    with torch.inference_mode():
        again = ac.act(o)
    assert torch.equal(again, got_a)
    nat().lib.lgk_policy_set_variant(0)


@pytest.mark.parametrize("n,nobs,ncobs,nact,hidden", [
    (300, 48, 61, 5, (256, 64, 32)),        # 5 actions: rows not 16-byte aligned (scalar output stores), critic wider than actor
    (1000, 235, 187, 16, (512, 128, 64)),   # 16 actions: all four warps of a quadrant own an action quad
    (129, 19, 19, 3, (128, 64, 32)),        # the games' high-level agents: 19 observations, 3 commands
    (1, 235, 235, 12, (512, 256, 128)),     # a single environment: 127 empty rows in the tile
    (255, 256, 8, 12, (256, 256, 128))])    # widest observation (8 full chunks) next to the narrowest, hidden[1] = hidden[0]
def test_policy_kernel_action_counts_and_privileged_obs(n, nobs, ncobs, nact, hidden):
    """The tcgen05 kernel beyond the 12-action / shared-observation case: odd action counts, the 16-action maximum,
    privileged observations of another width, and its actor-only (act / act_inference) and PPO.act launches agreeing."""
    from oracle.rsl_oracle import ActorCriticOracle
    from legged_games_gym_b200.rsl_rl.modules import ActorCritic
    torch.manual_seed(1)
    nat().lib.lgk_policy_set_variant(2)
    orc = ActorCriticOracle(nobs, ncobs, nact, hidden, hidden)
    with torch.no_grad():
        orc.std.copy_(torch.linspace(0.3, 1.2, nact))
    ac = ActorCritic(nobs, ncobs, nact, list(hidden), list(hidden)).to(DEV)
    ac.load_state_dict(orc.state_dict())
    obs, cobs = torch.randn(n, nobs) * 2, torch.randn(n, ncobs) * 2
    seed, step = 23, 9
    eps = torch.from_numpy(philox.normals(seed, step, np.arange(n), nact))
    a, v, lp, mu, sg = orc.act(obs, cobs, eps)
    ac.set_rng(seed, step)
    tol = dict(rtol=1e-3, atol=1e-3)
    with torch.inference_mode():
        out = ac.act_and_evaluate(obs.to(DEV), cobs.to(DEV))
        got = {k: t.clone() for k, t in out.items()}
        assert torch.equal(ac.act(obs.to(DEV)), got["actions"])                  # actor-only launch: same draws
        assert torch.allclose(ac.act_inference(obs.to(DEV)).cpu(), mu, **tol)    # no sampling: the mean
        assert torch.allclose(ac.evaluate(cobs.to(DEV)).cpu(), v, **tol)         # torch critic on the same weights
    torch.cuda.synchronize()
    assert torch.allclose(got["mean"].cpu(), mu, **tol) and torch.allclose(got["actions"].cpu(), a, **tol)
    assert torch.allclose(got["values"].cpu(), v, **tol) and torch.allclose(got["sigma"].cpu(), sg, **tol)
    assert torch.allclose(got["logp"].cpu(), lp, **tol)
    nat().lib.lgk_policy_set_variant(0)


def test_episode_stats_kernel_matches_the_runner_bookkeeping():
    """lgk_episode_stats = rsl_rl OnPolicyRunner's cur_reward_sum / cur_episode_length bookkeeping with the finished
    episodes reduced to (sum of returns, sum of lengths, count)."""
    g = torch.Generator().manual_seed(11)
    n = 5000
    cur = torch.zeros(2, n, device=DEV)
    stats = torch.zeros(3, dtype=torch.float64, device=DEV)
    w_rew, w_len, w_stats = torch.zeros(n), torch.zeros(n), torch.zeros(3, dtype=torch.float64)
    for step in range(12):
        rew = torch.randn(n, generator=g)
        done = torch.rand(n, generator=g) < 0.1
        r_d, d_d = rew.to(DEV), done.to(DEV)
        assert nat().lib.lgk_episode_stats(r_d.data_ptr(), d_d.data_ptr(), cur[0].data_ptr(), cur[1].data_ptr(), stats.data_ptr(),
                                           n, stream()) == 0
        w_rew += rew
        w_len += 1
        w_stats += torch.stack([(w_rew[done]).double().sum(), (w_len[done]).double().sum(), done.double().sum()])
        w_rew[done] = 0
        w_len[done] = 0
    torch.cuda.synchronize()
    assert torch.equal(cur[0].cpu(), w_rew) and torch.equal(cur[1].cpu(), w_len)
    assert torch.allclose(stats.cpu(), w_stats, rtol=1e-12, atol=1e-9)
    assert nat().lib.lgk_episode_stats(None, None, None, None, None, 0, stream()) != 0


def test_pinned_copy_kernels():
    """lgk_copy_from_pinned / lgk_copy_to_pinned / lgk_copy_rows_to_pinned: exact copies over the unified address space,
    argument errors as codes."""
    lib = nat().lib
    g = torch.Generator().manual_seed(3)
    for n in (4, 1000, 98304, 262144 + 4):
        h = torch.randn(n, generator=g).pin_memory()
        d = torch.zeros(n, device=DEV)
        assert lib.lgk_copy_from_pinned(d.data_ptr(), h.data_ptr(), n * 4, stream()) == 0
        torch.cuda.synchronize()
        assert torch.equal(d.cpu(), h)
        h2 = torch.zeros(n).pin_memory()
        assert lib.lgk_copy_to_pinned(h2.data_ptr(), d.data_ptr(), n * 4, stream()) == 0
        torch.cuda.synchronize()
        assert torch.equal(h2, h)
    assert lib.lgk_copy_from_pinned(d.data_ptr(), h.data_ptr(), 24, stream()) == 1          # not a multiple of 16
    assert lib.lgk_copy_from_pinned(d.data_ptr() + 4, h.data_ptr(), 16, stream()) == 2      # LGK_ERR_ALIGN
    # indexed rows: actor rows 2 * id + 1 of a [2N, 13] tensor, first `count` ids only
    N = 300
    src = torch.randn(2 * N, 13, generator=g).to(DEV)
    dst = torch.zeros(2 * N, 13).pin_memory()
    ids = torch.tensor([5, 17, 299, 0, 42], dtype=torch.int32, device=DEV)
    count = torch.tensor([3], dtype=torch.int32, device=DEV)
    assert lib.lgk_copy_rows_to_pinned(dst.data_ptr(), src.data_ptr(), 13, ids.data_ptr(), count.data_ptr(), 2, 1, 5, stream()) == 0
    torch.cuda.synchronize()
    want = torch.zeros(2 * N, 13)
    for i in (5, 17, 299):
        want[2 * i + 1] = src[2 * i + 1].cpu()
    assert torch.equal(dst, want)


def test_game_prepare_clips_like_torch():
    """lgk_game_prepare = HLG:161-174: torch.clip semantics (NaN stays NaN), wrap_to_pi on the heading command, prey
    command handed to the low-level command buffer."""
    import ctypes as C
    from oracle.legged_oracle import wrap_pi
    lib = nat().lib
    n = 257
    g = torch.Generator().manual_seed(4)
    cmd = torch.randn(n, 6, generator=g) * 3.0
    cmd[3, 0] = float("nan")
    cmd[4, 2] = 9.5
    cmd[5, 2] = -7.0
    want = cmd.clone()
    want[:, 0] = torch.clip(want[:, 0], min=-1.0, max=1.0)
    want[:, 1] = torch.clip(want[:, 1], min=-0.5, max=0.75)
    want[:, 2] = wrap_pi(want[:, 2].clone())
    want[:, 4] = torch.clip(want[:, 4], min=-2.0, max=2.0)
    want[:, 5] = torch.clip(want[:, 5], min=-1.5, max=2.5)
    d = cmd.to(DEV)
    ll = torch.zeros(n, 4, device=DEV)
    ranges = (C.c_float * 8)(-1.0, 1.0, -0.5, 0.75, -2.0, 2.0, -1.5, 2.5)
    rc = lib.lgk_game_prepare(d.data_ptr(), 6, d[:, 4:].data_ptr(), 6, ll.data_ptr(), n, C.byref(ranges), 1, stream())
    assert rc == 0
    torch.cuda.synchronize()
    got = d.cpu()
    assert torch.allclose(got, want, rtol=0, atol=1e-6, equal_nan=True)
    assert torch.equal(got[:, [0, 1, 3, 4, 5]].nan_to_num(7.0), want[:, [0, 1, 3, 4, 5]].nan_to_num(7.0))      # clips are exact
    assert torch.equal(ll.cpu().nan_to_num(7.0), got[:, :4].nan_to_num(7.0))


def test_reference_error_behaviour():
    """The Python layer raises what the reference raises: NameError for an unknown control type (LR:394),
    ValueError for an unknown terrain mesh type (LR:250); the games:
    AttributeError for a predator termination scale (DHLG:358-359), ValueError when the low-level checkpoint directory is
    missing (helpers.py:109, reached from HLG:100)."""
    from legged_games_gym_b200.envs import task_registry
    from legged_games_gym_b200.utils import get_args
    # (the string 'none' never reaches the NameError of LR:847 in the reference either: create_sim rejects it first, LR:250)
    for ov, exc in (({"control.control_type": "X"}, NameError), ({"terrain.mesh_type": "none"}, ValueError),
                    ({"terrain.mesh_type": "blob"}, ValueError)):
        case = harness.build_case("a1", 32, seed=1)
        harness.apply_overrides(case["cfg"], ov)
        with pytest.raises(exc):
            env, _ = product_env(case)
            env.step(torch.zeros(32, 12, device=DEV))
    case = harness.build_case("low_level_game", 32, seed=1, overrides={"terrain.mesh_type": "plane", "terrain.curriculum": False})
    with pytest.raises(AttributeError, match="reward_scales"):
        product_game(case, "dec", {"rewards_predator.scales.termination": 1.0}, ll_policy=lambda o: o)
    a = get_args(["--task", "high_level_game", "--num_envs", "32", "--headless"])
    with pytest.raises(ValueError, match="No runs in this directory"):
        task_registry.make_env(name="high_level_game", args=a)
