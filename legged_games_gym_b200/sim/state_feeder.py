"""Synthetic sim backend: stands where Isaac Gym / PhysX stands in the reference.

The reference env talks to PhysX through the tensor API on ``self.gym``
(legged_robot.py:92-96, 111-112, 410, 434, 444, 515-520).  PhysX is out of scope here ("fed by
state tensors"), so the env is constructed over a ``SimBackend`` exposing the same method names.
``StateFeeder`` owns the three state tensors (root_states, dof_state, net_contact_forces) on the
device and fills them with the seeded synthetic distributions of SURVEY.md section 8(d);
``HostStateFeeder`` keeps the sim state in pinned HOST memory and copies across PCIe at the same API
points where a CPU-pipeline PhysX would (the e2e leg of bench.py).
"""
import os

import numpy as np
import torch


def synth_state(num_envs, num_bodies, num_dof=12, seed=0, p_contact=0.3, actors_per_env=1, xy_max=(83., 163.),
                p_contact_body0=None):
    """numpy dict of seeded synthetic inputs (PCG64: stable across torch versions)."""
    g = np.random.default_rng(seed)
    N = num_envs
    na = N * actors_per_env
    root = np.zeros((na, 13), dtype=np.float32)
    root[:, 0] = g.uniform(-2., xy_max[0], na)
    root[:, 1] = g.uniform(-2., xy_max[1], na)
    root[:, 2] = g.normal(0.5, 0.1, na)
    q = g.normal(0., 1., (na, 4))
    q[:, :2] *= 0.3
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    root[:, 3:7] = q
    root[:, 7:13] = g.normal(0., 1., (na, 6))
    dof = g.normal(0., 1., (N * num_dof, 2)).astype(np.float32)
    force = g.normal(0., 2., (N * num_bodies, 3))          # draw order (force, then gate) is part of the fixtures
    gate = g.uniform(0., 1., (N * num_bodies, 1))
    thresh = np.full((N, num_bodies, 1), p_contact)
    if p_contact_body0 is not None:          # body 0 is the base/pelvis: its contact terminates the episode (LR:142)
        thresh[:, 0, :] = p_contact_body0
    contact = (force * (gate < thresh.reshape(-1, 1))).astype(np.float32)
    actions = g.normal(0., 1., (N, num_dof)).astype(np.float32)
    ep_len = g.integers(0, 1000, N).astype(np.int64)
    return dict(root_states=root, dof_state=dof, contact_forces=contact, actions=actions,
                episode_length_buf=ep_len)


def synth_height_field(rows=1300, cols=2100, seed=0):
    g = np.random.default_rng(seed + 7919)
    return g.integers(-300, 300, (rows, cols)).astype(np.int16)


def synth_terrain_origins(cfg_terrain):
    """Deterministic stand-in for Terrain.env_origins (reference utils/terrain.py:147-164): centre of
    sub-terrain (level i, type j) with a patterned platform height."""
    rows, cols = cfg_terrain.num_rows, cfg_terrain.num_cols
    o = np.zeros((rows, cols, 3))
    for i in range(rows):
        for j in range(cols):
            o[i, j, 0] = (i + 0.5) * cfg_terrain.terrain_length
            o[i, j, 1] = (j + 0.5) * cfg_terrain.terrain_width
            o[i, j, 2] = 0.05 * ((3 * i + 7 * j) % 11)
    return o


class SimBackend:
    """Method names = the Isaac Gym calls the hot path makes.  All are stream-ordered and non-blocking."""

    def acquire_actor_root_state_tensor(self):
        raise NotImplementedError

    def acquire_dof_state_tensor(self):
        raise NotImplementedError

    def acquire_net_contact_force_tensor(self):
        raise NotImplementedError

    def refresh_dof_state_tensor(self):
        pass

    def refresh_actor_root_state_tensor(self):
        pass

    def refresh_net_contact_force_tensor(self):
        pass

    def set_dof_actuation_force_tensor(self, torques):
        pass

    # optional zero-copy endpoints of a sim that lives in PINNED host memory: when they return a tensor, the torque kernel
    # reads the dof state from / mirrors its torques into that memory directly during the decimation loop, and the env
    # refreshes its device copy of the dof state once per step (after the last sub-step) instead of once per sub-step
    def dof_state_source(self):
        return None

    def actuation_force_sink(self):
        return None

    def simulate(self):
        pass

    def fetch_results(self):
        pass

    def set_dof_state_tensor_indexed(self, dof_state, env_ids_int32, count):
        pass

    def set_actor_root_state_tensor_indexed(self, root_states, env_ids_int32, count, actor_stride=1, actor_offset=0):
        """env_ids_int32[:count] (device int32, ascending) name the ENVS whose actor `actor_stride * id + actor_offset` was
        reset; the reference passes the actor ids themselves (LR:433-436, LLG:441-451)."""

    def set_actor_root_state_tensor(self, root_states):
        pass

    def set_dof_state_tensor(self, dof_state):
        pass

    # init-time domain randomisation (the reference writes these into PhysX shape / body properties, LR:261-283, 316-327)
    def set_rigid_shape_friction(self, friction_coeffs):
        self.friction_coeffs = friction_coeffs

    def set_base_mass_offsets(self, added_mass):
        self.added_base_mass = added_mass


class StateFeeder(SimBackend):
    """Device-resident synthetic state (inputs already in HBM when a step starts)."""
    graph_safe = True       # its hooks enqueue no work, so a whole env step can be captured into one CUDA graph

    def __init__(self, num_envs, num_bodies, num_dof=12, device="cuda", seed=0, p_contact=0.3,
                 actors_per_env=1, p_contact_body0=None):
        s = synth_state(num_envs, num_bodies, num_dof, seed, p_contact, actors_per_env, p_contact_body0=p_contact_body0)
        self.device = torch.device(device)
        self.num_envs, self.num_dof = num_envs, num_dof
        self.root_states = torch.from_numpy(s["root_states"]).to(self.device)
        self.dof_state = torch.from_numpy(s["dof_state"]).to(self.device)
        self.contact_forces = torch.from_numpy(s["contact_forces"]).to(self.device)
        self.synthetic_actions = torch.from_numpy(s["actions"]).to(self.device)
        self.synthetic_episode_length = torch.from_numpy(s["episode_length_buf"]).to(self.device)

    def acquire_actor_root_state_tensor(self):
        return self.root_states

    def acquire_dof_state_tensor(self):
        return self.dof_state

    def acquire_net_contact_force_tensor(self):
        return self.contact_forces


class _PinnedAlias:
    """__cuda_array_interface__ view of a pinned host tensor: with unified addressing the device reaches the allocation
    through the same pointer, so torch can wrap it as a CUDA tensor and every kernel works on it unchanged."""

    def __init__(self, t):
        self.__cuda_array_interface__ = {"shape": tuple(t.shape), "typestr": "<f4", "data": (t.data_ptr(), False),
                                         "version": 2, "strides": None}


class HostStateFeeder(StateFeeder):
    """Sim state lives in PINNED host memory (the reference's ``sim_device=cpu`` pipeline: PhysX results are host tensors).
    Three levels, each an A/B switch of the one above (bench.py's e2e leg; byte counters feed ``e2e.h2d_bytes_per_step`` /
    ``d2h_bytes_per_step``):

    * unified (default, ``LGK_HOST_UNIFIED=0`` turns it off): the tensors the env acquires ARE the pinned buffers, wrapped
      as CUDA tensors over the unified address space.  The step kernels pull what they read over PCIe themselves (K1's TMA
      bulk copies included) and write reset rows / pushed velocities straight back; refresh_* / set_* have nothing to do
      and a step has no copy at all.  Measured at 4096 envs: 110 us per step against 147 us with explicit transfers.
    * explicit transfers: device-resident twins, refresh_* = H2D, set_* = D2H, done by kernels over the unified address
      space for tensors up to 1 MB (``LGK_KERNEL_COPY_MAX``) and with zero-copy torque sub-steps (``LGK_HOST_ZERO_COPY``).
    * plain: cudaMemcpyAsync for everything (both switches off).

    Every hook is stream-ordered and capturable, so the whole step still replays as one CUDA graph."""
    graph_safe = True

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.h_root = self.root_states.cpu().pin_memory()
        self.h_dof = self.dof_state.cpu().pin_memory()
        self.h_contact = self.contact_forces.cpu().pin_memory()
        self.h_torques = None
        self.unified = os.environ.get("LGK_HOST_UNIFIED", "1") != "0"
        self.zero_copy = os.environ.get("LGK_HOST_ZERO_COPY", "1") != "0"
        # reset rows only (lgk_copy_rows_to_pinned) instead of whole tensors: 0.6 MB less D2H per step, measured neutral
        # for the step time at 4096 envs (two tiny launches against two 5-8 us copies), so off unless asked for
        self.indexed_rows = os.environ.get("LGK_HOST_INDEXED_ROWS", "0") == "1"
        self.state_in_host_memory = self.unified          # the env then launches the uncached-load (host_io) torque kernel
        if self.unified:
            dev = self.root_states.device
            self.root_states = torch.as_tensor(_PinnedAlias(self.h_root), device=dev)
            self.dof_state = torch.as_tensor(_PinnedAlias(self.h_dof), device=dev)
            self.contact_forces = torch.as_tensor(_PinnedAlias(self.h_contact), device=dev)
            for d, h in ((self.root_states, self.h_root), (self.dof_state, self.h_dof), (self.contact_forces, self.h_contact)):
                if not (d.is_cuda and d.data_ptr() == h.data_ptr()):
                    raise RuntimeError("pinned host memory is not reachable through the unified address space on this system")
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # tensors up to this size cross PCIe through a kernel over the unified address space (lgk_copy_*_pinned): inside a
    # captured graph that is 1.2-2x faster than a memcpy node for the 4 KB - 1 MB transfers of an env step
    # (scratch/pcie_probe.py); larger ones (the observation matrix) keep the copy engine
    KERNEL_COPY_MAX_BYTES = int(os.environ.get("LGK_KERNEL_COPY_MAX", 1 << 20))

    @staticmethod
    def _nbytes(t):
        return t.numel() * t.element_size()

    def _copy(self, dst, src, from_host):
        nbytes = self._nbytes(src)
        if (nbytes <= self.KERNEL_COPY_MAX_BYTES and nbytes % 16 == 0 and dst.is_contiguous() and src.is_contiguous()
                and dst.data_ptr() % 16 == 0 and src.data_ptr() % 16 == 0):
            from .. import _native as nat
            st = torch.cuda.current_stream().cuda_stream
            fn = nat.lib.lgk_copy_from_pinned if from_host else nat.lib.lgk_copy_to_pinned
            nat.check(fn(dst.data_ptr(), src.data_ptr(), nbytes, st), "lgk_copy_pinned")
        else:
            dst.copy_(src, non_blocking=True)
        return nbytes

    def _h2d(self, dst, src):
        self.h2d_bytes += self._copy(dst, src, True)

    def _d2h(self, dst, src):
        self.d2h_bytes += self._copy(dst, src, False)

    # unified mode: nothing to copy; the counters take what the kernels of the step pull / push over PCIe instead --
    # dof state: one pull per torque launch (counted at the refresh that follows it) plus K1's; root state: K1's tile loads
    # plus K2's pose reads; contact forces: K1's; torques: the mirror; reset rows: K1's direct writes
    def refresh_dof_state_tensor(self):
        if self.unified:
            self.h2d_bytes += self._nbytes(self.h_dof)
        else:
            self._h2d(self.dof_state, self.h_dof)

    def refresh_actor_root_state_tensor(self):
        if self.unified:
            self.h2d_bytes += 2 * self._nbytes(self.h_root) + self._nbytes(self.h_dof)
        else:
            self._h2d(self.root_states, self.h_root)

    def refresh_net_contact_force_tensor(self):
        if self.unified:
            self.h2d_bytes += self._nbytes(self.h_contact)
        else:
            self._h2d(self.contact_forces, self.h_contact)

    def set_dof_actuation_force_tensor(self, torques):
        if self.zero_copy or self.unified:
            return                      # the torque kernel has already written h_torques (actuation_force_sink)
        if self.h_torques is None:
            self.h_torques = torch.empty(torques.shape, dtype=torques.dtype).pin_memory()
        self._d2h(self.h_torques, torques)

    def dof_state_source(self):
        if self.unified or not self.zero_copy:
            return None                 # unified: the env's own dof_state tensor already is the pinned buffer
        self.h2d_bytes += self._nbytes(self.h_dof)           # one pull of the tensor per torque launch
        return self.h_dof

    def actuation_force_sink(self):
        if not (self.zero_copy or self.unified):
            return None
        if self.h_torques is None:
            self.h_torques = torch.empty(self.num_envs, self.num_dof, dtype=torch.float).pin_memory()
        self.d2h_bytes += self._nbytes(self.h_torques)
        return self.h_torques

    def _d2h_rows(self, dst, src, row_floats, env_ids_int32, count, stride, offset):
        """rows of the reset envs only (lgk_copy_rows_to_pinned); the byte counter takes the count of the step being run
        eagerly (the first one: bench.py reads the counters after it), a replayed graph moves what its step resets"""
        if not self.unified:
            from .. import _native as nat
            st = torch.cuda.current_stream().cuda_stream
            nat.check(nat.lib.lgk_copy_rows_to_pinned(dst.data_ptr(), src.data_ptr(), row_floats, env_ids_int32.data_ptr(),
                                                      count.data_ptr(), stride, offset, env_ids_int32.numel(), st),
                      "lgk_copy_rows_to_pinned")
        if not torch.cuda.is_current_stream_capturing():
            self.d2h_bytes += int(count.item()) * row_floats * 4

    def set_dof_state_tensor_indexed(self, dof_state, env_ids_int32, count):
        if (self.unified or self.indexed_rows) and torch.is_tensor(count):
            self._d2h_rows(self.h_dof, dof_state, 2 * self.num_dof, env_ids_int32, count, 1, 0)
        elif not self.unified:
            self._d2h(self.h_dof, dof_state)

    def set_actor_root_state_tensor_indexed(self, root_states, env_ids_int32, count, actor_stride=1, actor_offset=0):
        if (self.unified or self.indexed_rows) and torch.is_tensor(count):
            self._d2h_rows(self.h_root, root_states, 13, env_ids_int32, count, actor_stride, actor_offset)
        elif actor_offset == 0 and not self.unified:          # one copy of the whole tensor covers every actor of the env
            self._d2h(self.h_root, root_states)

    def set_actor_root_state_tensor(self, root_states):
        if self.unified:
            self.d2h_bytes += self._nbytes(self.h_root)       # the pushed tile rows were written by K1
        else:
            self._d2h(self.h_root, root_states)
