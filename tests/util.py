"""Shared helpers for the parity tests: build the PRODUCT env (CUDA kernels through the C ABI) on the same seeded
inputs the oracle / the golden fixtures use."""
import copy
import types

import numpy as np
import torch

from oracle import harness


def product_env(case, device="cuda:0", graph=False, tile=0):
    from legged_games_gym_b200.envs import task_registry
    from legged_games_gym_b200.sim.state_feeder import SimBackend
    from legged_games_gym_b200.utils.helpers import SimParams

    class ExplicitFeeder(SimBackend):
        graph_safe = graph

        def __init__(self, st):
            self.root_states = torch.from_numpy(st["root_states"].copy()).to(device)
            self.dof_state = torch.from_numpy(st["dof_state"].copy()).to(device)
            self.contact_forces = torch.from_numpy(st["contact_forces"].copy()).to(device)

        def acquire_actor_root_state_tensor(self):
            return self.root_states

        def acquire_dof_state_tensor(self):
            return self.dof_state

        def acquire_net_contact_force_tensor(self):
            return self.contact_forces

    cfg = copy.deepcopy(case["cfg"])
    cfg.seed = case["seed"]
    terrain = None
    if case["height_samples"] is not None:
        hs = case["height_samples"]
        terrain = types.SimpleNamespace(cfg=cfg.terrain, env_length=cfg.terrain.terrain_length,
                                        env_width=cfg.terrain.terrain_width, heightsamples=hs, tot_rows=hs.shape[0],
                                        tot_cols=hs.shape[1], env_origins=case["terrain_origins"])
    cls = task_registry.get_task_class(case["task"])
    feeder = ExplicitFeeder(case["state"])
    env = cls(cfg=cfg, sim_params=SimParams(dt=cfg.sim.dt, use_gpu_pipeline=True), physics_engine="physx",
              sim_device=device, headless=True, sim_backend=feeder, terrain=terrain,
              init_terrain_levels=case["init_levels"])
    env.episode_length_buf[:] = torch.from_numpy(case["state"]["episode_length_buf"]).to(device)
    return env, feeder


def product_game(case, variant, overrides=None, ll_policy=None, device="cuda:0"):
    """The PRODUCT HighLevelGame ("hl") / DecHighLevelGame ("dec") around a LowLevelGame fed by the case's state."""
    from legged_games_gym_b200.envs import task_registry
    from legged_games_gym_b200.sim.state_feeder import SimBackend
    from legged_games_gym_b200.utils.helpers import SimParams

    class ExplicitFeeder(SimBackend):
        graph_safe = False

        def __init__(self, st):
            self.root_states = torch.from_numpy(st["root_states"].copy()).to(device)
            self.dof_state = torch.from_numpy(st["dof_state"].copy()).to(device)
            self.contact_forces = torch.from_numpy(st["contact_forces"].copy()).to(device)

        def acquire_actor_root_state_tensor(self):
            return self.root_states

        def acquire_dof_state_tensor(self):
            return self.dof_state

        def acquire_net_contact_force_tensor(self):
            return self.contact_forces

    name = "high_level_game" if variant == "hl" else "dec_high_level_game"
    cfg = copy.deepcopy(task_registry.env_cfgs[name])
    cfg.env.num_envs = case["cfg"].env.num_envs
    harness.apply_overrides(cfg, overrides)
    ll_cfg = copy.deepcopy(case["cfg"])
    ll_cfg.seed = case["seed"]
    terrain = None
    if case["height_samples"] is not None:
        hs = case["height_samples"]
        terrain = types.SimpleNamespace(cfg=ll_cfg.terrain, env_length=ll_cfg.terrain.terrain_length,
                                        env_width=ll_cfg.terrain.terrain_width, heightsamples=hs, tot_rows=hs.shape[0],
                                        tot_cols=hs.shape[1], env_origins=case["terrain_origins"])
    feeder = ExplicitFeeder(case["state"])
    cls = task_registry.get_task_class(name)
    game = cls(cfg=cfg, sim_params=SimParams(dt=ll_cfg.sim.dt, use_gpu_pipeline=True), physics_engine="physx",
               sim_device=device, headless=True, ll_policy=ll_policy, ll_env_cfg=ll_cfg, sim_backend=feeder,
               terrain=terrain, init_terrain_levels=case["init_levels"])
    game.ll_env.episode_length_buf[:] = torch.from_numpy(case["state"]["episode_length_buf"]).to(device)
    return game, feeder


def feeder_state(feeder):
    return dict(root_states=feeder.root_states, dof_state=feeder.dof_state, contact_forces=feeder.contact_forces)


FLOAT_RTOL, FLOAT_ATOL = 1e-5, 1e-5      # north_star: within 1e-5 relative for fp32 obs / rewards / torques

# north_star's bit-exact set: reset / termination masks, counters, height-sample indices (test_height_scan_*) and command
# resampling under fixed RNG.  Beyond the masks: everything reset_idx / _resample_commands / _push_robots WRITE is a
# uniform draw scaled by individually rounded fp32 ops (LR:353-366, 397-436, 438-444), and rows they do not touch are the
# inputs, so root_states, dof_state and env_origins are compared bit for bit as a whole; `commands` columns 0, 1 and 3
# (lin_vel_x, lin_vel_y, heading) always, column 2 (yaw rate) exactly wherever it is a resampled draw -- with
# heading_command it is recomputed every step through atan2 (LR:337-340) and falls under the fp32 tolerance.
EXACT = ("reset_buf", "time_out_buf", "episode_length_buf", "last_contacts", "terrain_levels", "ex_time_outs",
         "root_states", "dof_state", "env_origins")


def _excess(g, w, scale):
    err = np.abs(g.astype(np.float64) - w.astype(np.float64))
    tol = FLOAT_ATOL * scale + FLOAT_RTOL * np.abs(w)
    return err, tol


def assert_snapshots_close(got, want, step, exact=EXACT, atol_scale=None, skip=(), heading_command=True):
    """bit-exact on masks / counters / levels / sim-state rows / resampled commands; rtol 1e-5 (+ atol 1e-5 * scale of the
    quantity) on fp32.  The scale is one number per tensor for [N] / small [N, k] quantities and one number PER COLUMN for
    obs_buf, whose column groups differ by orders of magnitude (0.05-scaled joint velocities next to 5x heights)."""
    atol_scale = atol_scale or {}
    bad = []
    for k, w in want.items():
        if k in skip:
            continue
        assert k in got, f"step {step}: product snapshot lacks {k}"
        g = got[k]
        g = g.numpy() if isinstance(g, torch.Tensor) else np.asarray(g)
        w = w.numpy() if isinstance(w, torch.Tensor) else np.asarray(w)
        if k in exact or w.dtype == np.bool_ or np.issubdtype(w.dtype, np.integer):
            if np.issubdtype(w.dtype, np.floating):
                same = (g.view(np.uint32) == w.view(np.uint32)) | ((g == 0) & (w == 0))      # -0.0 == 0.0 like torch.equal
                if not same.all():
                    i = int(np.argmax(~same))
                    bad.append(f"{k}: {int((~same).sum())} of {w.size} differ (exact), first at flat {i}: got {g.flat[i]!r} want {w.flat[i]!r}")
            elif not np.array_equal(g.astype(np.int64), w.astype(np.int64)):
                bad.append(f"{k}: {int((g.astype(np.int64) != w.astype(np.int64)).sum())} of {w.size} differ (exact)")
            continue
        if k == "commands":
            cols = [0, 1, 3] if heading_command else [0, 1, 2, 3]
            cols = [c for c in cols if c < w.shape[1]]
            ge, we = g[:, cols], w[:, cols]
            same = (ge.view(np.uint32) == we.view(np.uint32)) | ((ge == 0) & (we == 0))
            if not same.all():
                r, c = np.argwhere(~same)[0]
                bad.append(f"commands[:, {cols}]: {int((~same).sum())} entries differ (exact), first env {r} col {cols[c]}: got {ge[r, c]!r} want {we[r, c]!r}")
        if k == "obs_buf" and w.ndim == 2 and k not in atol_scale:
            scale = np.maximum(np.abs(w).max(axis=0, keepdims=True), 1e-2) if w.size else 1.0
        else:
            scale = atol_scale.get(k, max(1.0, float(np.abs(w).max()) if w.size else 1.0))
        err, tol = _excess(g, w, scale)
        if not np.all(err <= tol):
            i = int(np.argmax(err - tol))
            bad.append(f"{k}: max excess at flat {i}: got {g.flat[i]!r} want {w.flat[i]!r} (|err| {err.flat[i]:.3e}, tol {tol.flat[i]:.3e})")
    assert not bad, f"step {step} mismatches:\n  " + "\n  ".join(bad)
