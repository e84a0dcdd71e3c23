"""TEST INFRASTRUCTURE ONLY -- run the UNMODIFIED reference hot path in this container.

/root/reference is pure Python but needs ``isaacgym`` (closed PhysX binding), ``rsl_rl`` and
``matplotlib``, none of which is installed or vendored.  This loader puts stub modules into
``sys.modules`` (only ``isaacgym.torch_utils`` carries real math -- see isaac_torch_utils.py),
imports ``legged_gym`` from /root/reference, and builds ``LeggedRobot`` / ``Anymal`` / ``Cassie``
instances with ``object.__new__`` + the reference's own ``_parse_cfg`` / ``_get_env_origins`` /
``_init_buffers`` / ``_prepare_reward_function`` (legged_robot.py:781, 752, 511, 583), fed by a
no-op gym whose ``acquire_*_tensor`` hand back caller-owned CPU tensors.

Used by ``oracle/make_golden.py`` (fixture generation) and by ``tests/test_oracle_vs_reference.py``
(skipped when /root/reference is absent, i.e. on the GPU box).  Never imported by the product.

RNG tap: the reference draws with ``torch_rand_float`` / ``torch.rand_like`` /
``torch.randint_like`` from the global generator in a data-dependent order
(SURVEY.md App. A.7).  ``RngTap`` replaces those three names for the duration of a call and
serves the values from explicit per-env tables ``U[stream][env, j]`` so that reference,
oracle and kernels all consume identical numbers.
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

from . import isaac_torch_utils
from . import philox

REFERENCE_ROOT = os.environ.get("LGK_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "legged_gym"))


class _Anything:
    """Attribute sink standing in for gymapi / gymutil symbols that are only touched at import."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__getattr__ = lambda attr: _Anything  # type: ignore[attr-defined]
    sys.modules[name] = m
    return m


_loaded = None


def load_reference():
    """Import the reference ``legged_gym`` package with stubs; returns the module namespace."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    iso = _stub("isaacgym")
    iso.gymapi = _stub("isaacgym.gymapi")
    iso.gymtorch = _stub("isaacgym.gymtorch", wrap_tensor=lambda t: t, unwrap_tensor=lambda t: t)
    iso.gymutil = _stub("isaacgym.gymutil")
    iso.terrain_utils = _stub("isaacgym.terrain_utils")
    tu = types.ModuleType("isaacgym.torch_utils")
    for n in isaac_torch_utils.__all__:
        setattr(tu, n, getattr(isaac_torch_utils, n))
    tu.__all__ = list(isaac_torch_utils.__all__)
    sys.modules["isaacgym.torch_utils"] = tu
    iso.torch_utils = tu
    rsl = _stub("rsl_rl")
    rsl.env = _stub("rsl_rl.env", VecEnv=object)
    rsl.runners = _stub("rsl_rl.runners", OnPolicyRunner=_Anything, DecGamePolicyRunner=_Anything)
    _stub("rsl_rl.runners.low_level_policy_runner", LLPolicyRunner=_Anything)
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import legged_gym.envs as envs  # registers the 8 tasks
    _loaded = envs
    return envs


class FakeGym:
    """Every Isaac Gym call the hot path makes (legged_robot.py:92-96, 111-112, 410, 434, 444,
    515-520) is a no-op; the three state tensors are owned by the caller."""

    def __init__(self, root_states, dof_state, contact_forces):
        self._t = (root_states, dof_state, contact_forces)

    def acquire_actor_root_state_tensor(self, sim):
        return self._t[0]

    def acquire_dof_state_tensor(self, sim):
        return self._t[1]

    def acquire_net_contact_force_tensor(self, sim):
        return self._t[2]

    def __getattr__(self, name):
        return lambda *a, **k: None


class _SimParams:
    def __init__(self, dt):
        self.dt = dt


class _Terrain:
    def __init__(self, cfg, env_origins):
        self.cfg = cfg
        self.env_length = cfg.terrain_length
        self.env_width = cfg.terrain_width
        self.env_origins = env_origins


def terrain_origins_grid(cfg_terrain):
    """Synthetic stand-in for Terrain.env_origins (legged_gym/utils/terrain.py:147-164): platform
    centre of sub-terrain (row i, col j); z from a deterministic pattern.  float64 like numpy."""
    rows, cols = cfg_terrain.num_rows, cfg_terrain.num_cols
    o = np.zeros((rows, cols, 3))
    for i in range(rows):
        for j in range(cols):
            o[i, j, 0] = (i + 0.5) * cfg_terrain.terrain_length
            o[i, j, 1] = (j + 0.5) * cfg_terrain.terrain_width
            o[i, j, 2] = 0.05 * ((3 * i + 7 * j) % 11)
    return o


def make_ref_env(task, num_envs, consts, state, height_samples=None, cfg_overrides=None,
                 init_levels=None):
    """Build the reference env class registered under ``task`` without Isaac Gym.

    consts: dict(dof_names, num_bodies, feet_indices, penalised_contact_indices,
                 termination_contact_indices, dof_lower, dof_upper, dof_vel_limits, torque_limits)
    state:  dict(root_states[N,13], dof_state[N*12,2], contact_forces[N*NB,3]) CPU float tensors
            (owned by the caller, mutated in place by the reference exactly like PhysX tensors).
    """
    envs = load_reference()
    from legged_gym.utils.task_registry import task_registry
    import copy
    cls = task_registry.get_task_class(task)
    cfg = copy.deepcopy(task_registry.env_cfgs[task])
    cfg.env.num_envs = num_envs
    if cfg_overrides:
        for path, val in cfg_overrides.items():
            obj = cfg
            parts = path.split(".")
            for p in parts[:-1]:
                obj = getattr(obj, p)
            setattr(obj, parts[-1], val)
    env = object.__new__(cls)
    dev = "cpu"
    env.cfg = cfg
    env.sim_params = _SimParams(cfg.sim.dt)
    env.height_samples = height_samples
    env.debug_viz = False
    env.init_done = False
    env._parse_cfg(cfg)
    # --- what BaseTask.__init__ would set (base_task.py:40-82)
    env.gym = FakeGym(state["root_states"], state["dof_state"], state["contact_forces"])
    env.sim = None
    env.viewer = None
    env.headless = True
    env.enable_viewer_sync = True
    env.device = dev
    env.num_envs = num_envs
    env.num_obs = cfg.env.num_observations
    env.num_privileged_obs = cfg.env.num_privileged_obs
    env.num_actions = cfg.env.num_actions
    env.obs_buf = torch.zeros(num_envs, env.num_obs, dtype=torch.float)
    env.rew_buf = torch.zeros(num_envs, dtype=torch.float)
    env.reset_buf = torch.ones(num_envs, dtype=torch.long)
    env.episode_length_buf = torch.zeros(num_envs, dtype=torch.long)
    env.time_out_buf = torch.zeros(num_envs, dtype=torch.bool)
    env.privileged_obs_buf = None
    env.extras = {}
    # --- what create_sim / _create_envs would set (legged_robot.py:232-251, 657-750)
    env.up_axis_idx = 2
    nd = len(consts["dof_names"])
    env.num_dof = nd
    env.num_dofs = nd
    env.num_bodies = consts["num_bodies"]
    env.dof_names = list(consts["dof_names"])
    env.feet_indices = torch.tensor(consts["feet_indices"], dtype=torch.long)
    env.penalised_contact_indices = torch.tensor(consts["penalised_contact_indices"], dtype=torch.long)
    env.termination_contact_indices = torch.tensor(consts["termination_contact_indices"], dtype=torch.long)
    # _process_dof_props semantics (legged_robot.py:299-313) through the reference's own method
    props = {"lower": np.asarray(consts["dof_lower"], dtype=np.float32),
             "upper": np.asarray(consts["dof_upper"], dtype=np.float32),
             "velocity": np.asarray(consts["dof_vel_limits"], dtype=np.float32),
             "effort": np.asarray(consts["torque_limits"], dtype=np.float32)}

    class _Props(dict):
        def __len__(self):
            return nd
    env._process_dof_props(_Props(props), 0)
    base_init = cfg.init_state.pos + cfg.init_state.rot + cfg.init_state.lin_vel + cfg.init_state.ang_vel
    env.base_init_state = torch.tensor(base_init, dtype=torch.float)
    if cfg.terrain.mesh_type in ("heightfield", "trimesh"):
        env.terrain = _Terrain(cfg.terrain, terrain_origins_grid(cfg.terrain))
    # _get_env_origins draws terrain_levels with torch.randint (legged_robot.py:762); make it explicit
    env._get_env_origins()
    if init_levels is not None and env.custom_origins:
        env.terrain_levels[:] = torch.as_tensor(init_levels, dtype=torch.long)
        env.env_origins[:] = env.terrain_origins[env.terrain_levels, env.terrain_types]
    if cls.__name__ == "LowLevelGame":        # what its _create_envs records (low_level_game.py:760-812): prey first
        env.prey_indices = torch.arange(num_envs, dtype=torch.long) * 2
        env.predator_indices = torch.arange(num_envs, dtype=torch.long) * 2 + 1
    if cls.__name__ == "Anymal" and cfg.control.use_actuator_network:
        path = cfg.control.actuator_net_file.format(LEGGED_GYM_ROOT_DIR=REFERENCE_ROOT)
        env.actuator_network = torch.jit.load(path)
    env._init_buffers()
    env._prepare_reward_function()
    env.init_done = True
    return env


class RngTap:
    """Serve the reference's random draws from explicit per-env tables.

    tables: dict stream_id -> float32 array [N, width] (uniforms in [0,1)), plus
            philox.STREAM_TERRAIN -> uint32 array [N, 1].
    The wrapped instance methods tell the tap which env_ids / stream the next draws belong to.
    """

    def __init__(self, env, mod):
        self.env = env
        self.mod = mod            # the reference module whose globals hold torch_rand_float
        self.tables = None
        self.ids = None
        self.stream = None
        self.col = 0
        self.in_reset = False
        self.game = False         # inside HighLevelGame.reset_idx / DecHighLevelGame.reset_idx: the GAME_* streams
        self._wrap()

    def set_tables(self, tables):
        self.tables = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in tables.items()}

    # -- patched draw functions
    def _rand_float(self, lower, upper, shape, device):
        n, w = shape
        t = self.tables[self.stream]
        ids = self.ids
        assert n == len(ids), (n, len(ids), self.stream)
        u = t[ids, self.col:self.col + w]
        self.col += w
        return (upper - lower) * u + lower

    def _rand_like(self, x, **k):
        t = self.tables[philox.STREAM_OBS]
        assert tuple(x.shape) == tuple(t.shape), (x.shape, t.shape)
        return t.clone()

    def _randint_like(self, x, high, **k):
        t = self.tables[philox.STREAM_TERRAIN][self.ids, 0].to(torch.int64)
        return (t % int(high)).to(x.dtype)

    def _begin(self, stream, ids):
        self.stream, self.ids, self.col = stream, ids, 0

    # low_level_game's predator spawn (low_level_game.py:421-422): Tensor.uniform_ on a [len(ids), 3] tensor, then
    # torch.rand(len(ids)) -- served from the PREDATOR table, columns 0:3 and 3
    def _uniform_(self, x, a=0.0, b=1.0):
        u = self.tables[philox.STREAM_GAME_PREDATOR if self.game else philox.STREAM_PREDATOR][self.ids, 0:x.shape[1]]
        assert tuple(u.shape) == tuple(x.shape), (u.shape, x.shape)
        x.copy_((b - a) * u + a)
        return x

    def _rand(self, n, **k):
        assert n == len(self.ids)
        return self.tables[philox.STREAM_GAME_PREDATOR if self.game else philox.STREAM_PREDATOR][self.ids, 3].clone()

    def _wrap(self):
        env, tap = self.env, self
        cls = type(env)
        o_resample = cls._resample_commands
        o_dofs = cls._reset_dofs
        o_roots = cls._reset_root_states
        o_push = cls._push_robots
        o_terr = cls._update_terrain_curriculum
        o_reset = cls.reset_idx

        def resample(self_, env_ids):
            tap._begin(philox.STREAM_RESET_CMD if tap.in_reset else philox.STREAM_CMD, env_ids)
            return o_resample(self_, env_ids)

        def dofs(self_, env_ids):
            tap._begin(philox.STREAM_GAME_DOF if tap.game else philox.STREAM_RESET_DOF, env_ids)
            return o_dofs(self_, env_ids)

        def roots(self_, env_ids):
            tap._begin(philox.STREAM_GAME_ROOT if tap.game else philox.STREAM_RESET_ROOT, env_ids)
            if not self_.custom_origins:
                tap.col = 2     # the xy draw is skipped on flat ground (legged_robot.py:427-429)
            return o_roots(self_, env_ids)

        def push(self_):
            tap._begin(philox.STREAM_PUSH, torch.arange(self_.num_envs))
            return o_push(self_)

        def terr(self_, env_ids):
            tap._begin(philox.STREAM_TERRAIN, env_ids)
            return o_terr(self_, env_ids)

        def reset(self_, env_ids):
            tap.in_reset = True
            try:
                return o_reset(self_, env_ids)
            finally:
                tap.in_reset = False

        env._resample_commands = types.MethodType(resample, env)
        env._reset_dofs = types.MethodType(dofs, env)
        env._reset_root_states = types.MethodType(roots, env)
        env._push_robots = types.MethodType(push, env)
        env._update_terrain_curriculum = types.MethodType(terr, env)
        env.reset_idx = types.MethodType(reset, env)

    @contextlib.contextmanager
    def active(self):
        g = self.mod.__dict__
        saved = (g["torch_rand_float"], torch.rand_like, torch.randint_like, torch.rand, torch.Tensor.uniform_)
        g["torch_rand_float"] = self._rand_float
        torch.rand_like = self._rand_like
        torch.randint_like = self._randint_like
        tap = self
        if philox.STREAM_PREDATOR in (self.tables or {}) and type(self.env).__name__ == "LowLevelGame":
            torch.rand = self._rand
            torch.Tensor.uniform_ = lambda x, a=0.0, b=1.0: tap._uniform_(x, a, b)
        try:
            yield
        finally:
            g["torch_rand_float"], torch.rand_like, torch.randint_like, torch.rand, torch.Tensor.uniform_ = saved


def attach_tap(env):
    import importlib
    # the draw helpers are module globals of the file that defines the class (LowLevelGame is a stand-alone copy of
    # LeggedRobot, not a subclass)
    mod = importlib.import_module(type(env).__module__) if type(env).__name__ == "LowLevelGame" else \
        importlib.import_module("legged_gym.envs.base.legged_robot")
    return RngTap(env, mod)


# ---------------------------------------------------------------------------------------------- high-level games
def make_ref_game(variant, ll_env, tap, cfg_overrides=None):
    """Build the reference HighLevelGame ("hl") / DecHighLevelGame ("dec") around an already built reference LowLevelGame
    instance, without its constructor (high_level_game.py:27-144 needs Isaac Gym, the forked rsl_rl LLPolicyRunner and a
    checkpoint that is not in the tree): object.__new__ + the attributes __init__ would set + the reference's own
    _parse_cfg / _init_buffers / _prepare_reward_function*.  The low-level policy is ``game.ll_policy`` (the caller sets
    it to a function returning the low-level actions of the step).  reset_idx is wrapped so that the second reset of the
    low-level root / dof state inside one step draws from the GAME_* streams of the tap."""
    import copy
    load_reference()
    from legged_gym.utils.task_registry import task_registry
    name = "high_level_game" if variant == "hl" else "dec_high_level_game"
    cls = task_registry.get_task_class(name)
    cfg = copy.deepcopy(task_registry.env_cfgs[name])
    n = ll_env.num_envs
    cfg.env.num_envs = n
    for path, val in (cfg_overrides or {}).items():
        obj = cfg
        parts = path.split(".")
        for p in parts[:-1]:
            obj = getattr(obj, p)
        setattr(obj, parts[-1], val)
    g = object.__new__(cls)
    g.cfg = cfg
    g.device = "cpu"
    g.headless = True
    g.capture_dist = cfg.env.capture_dist
    g.MAX_REL_POS = 100.
    g.ll_env = ll_env
    g.ll_policy = None
    g._parse_cfg(cfg)
    g.num_envs = n
    g.reset_buf = torch.ones(n, dtype=torch.long)
    g.episode_length_buf = torch.zeros(n, dtype=torch.long)
    g.time_out_buf = torch.zeros(n, dtype=torch.bool)
    g.curr_episode_step = torch.zeros(n, dtype=torch.long)
    g.extras = {}
    if variant == "hl":
        g.num_obs, g.num_privileged_obs, g.num_actions = cfg.env.num_observations, cfg.env.num_privileged_obs, cfg.env.num_actions
        g.obs_buf = g.MAX_REL_POS * torch.ones(n, g.num_obs, dtype=torch.float)
        g.rew_buf = torch.zeros(n, dtype=torch.float)
        g.privileged_obs_buf = None
        g._init_buffers()
        g._prepare_reward_function()
    else:
        g.num_obs_prey, g.num_obs_pred = cfg.env.num_observations_prey, cfg.env.num_observations_predator
        g.num_actions_prey, g.num_actions_pred = cfg.env.num_actions_prey, cfg.env.num_actions_predator
        g.num_privileged_obs_prey = g.num_privileged_obs_pred = None
        g.obs_buf_prey = g.MAX_REL_POS * torch.ones(n, g.num_obs_prey, dtype=torch.float)
        g.obs_buf_prey[:, 12:16] = 0
        g.rew_buf_prey = torch.zeros(n, dtype=torch.float)
        g.obs_buf_pred = g.MAX_REL_POS * torch.ones(n, g.num_obs_pred, dtype=torch.float)
        g.rew_buf_pred = torch.zeros(n, dtype=torch.float)
        g.privileged_obs_buf_pred = g.privileged_obs_buf_prey = None
        g._init_buffers()
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            g._prepare_reward_function_pred()
            g._prepare_reward_function_prey()
    g.init_done = True
    o_reset = cls.reset_idx

    def reset(self_, env_ids):
        tap.game = True
        try:
            return o_reset(self_, env_ids)
        finally:
            tap.game = False

    g.reset_idx = types.MethodType(reset, g)
    return g
