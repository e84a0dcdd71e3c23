"""Can the step kernels run directly on PINNED HOST memory (unified addressing), incl. K1's TMA bulk copies?
Builds a bench env whose sim-state tensors are CUDA-tensor aliases of pinned host buffers and compares / times it
against a device-resident env."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from legged_games_gym_b200.sim.state_feeder import StateFeeder

dev = "cuda:0"
torch.cuda.set_device(0)


class _CAI:
    def __init__(self, t):
        self.__cuda_array_interface__ = {"shape": tuple(t.shape), "typestr": "<f4", "data": (t.data_ptr(), False),
                                         "version": 2, "strides": None}


def alias(host_pinned):
    return torch.as_tensor(_CAI(host_pinned), device=dev)


h = torch.arange(16, dtype=torch.float32).pin_memory()
d = alias(h)
print("alias is_cuda", d.is_cuda, "ptr equal", d.data_ptr() == h.data_ptr())
d += 1
torch.cuda.synchronize()
print("host sees device write:", h[:4].tolist())
h[0] = 100.
print("device sees host write:", float((d * 1).cpu()[0]))


class UvaFeeder(StateFeeder):
    graph_safe = True

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.h_root, self.h_dof, self.h_contact = (t.cpu().pin_memory() for t in (self.root_states, self.dof_state, self.contact_forces))
        self.root_states, self.dof_state, self.contact_forces = alias(self.h_root), alias(self.h_dof), alias(self.h_contact)


N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
orig = bench.make_env.__globals__
import legged_games_gym_b200.sim.state_feeder as sf
sf_host = sf.HostStateFeeder
sf.HostStateFeeder = UvaFeeder                       # bench.make_env(host_sim=True) now builds the UVA feeder
torch.manual_seed(0)
env_u, fu = bench.make_env(N, dev, host_sim=True)
sf.HostStateFeeder = sf_host
torch.manual_seed(0)
env_d, fd = bench.make_env(N, dev, host_sim=False)
acts = fd.synthetic_actions
g = torch.Generator().manual_seed(5)
for step in range(4):
    env_u.step(acts); env_d.step(acts)
    torch.cuda.synchronize()
    ok = torch.equal(env_u.obs_buf, env_d.obs_buf) and torch.equal(env_u.rew_buf, env_d.rew_buf) and torch.equal(env_u.reset_buf, env_d.reset_buf)
    ok = ok and torch.equal(fu.h_root, env_d.root_states.cpu()) and torch.equal(fu.h_dof, env_d.dof_state.cpu())
    print("step", step, "identical to the device-resident env:", ok, "graph:", env_u._graph is not None, flush=True)
    dv = torch.randn(fu.h_dof.shape[0], generator=g) * 0.1
    fu.h_dof[:, 1] += dv                                  # the host "simulator" moves
    fd.dof_state[:, 1] += dv.to(dev)
for env, label in ((env_d, "device-resident"), (env_u, "pinned-host-resident (UVA)")):
    for _ in range(10):
        env.step(acts)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200):
        env.step(acts)
    b.record(); torch.cuda.synchronize()
    print(f"{label}: {a.elapsed_time(b) / 200 * 1e3:.1f} us per step", flush=True)
