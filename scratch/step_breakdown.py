"""Time sub-sequences of the env step as CUDA graphs (replica rotation) to see where a 4096-env step goes."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from legged_games_gym_b200 import _native as nat
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bench.USE_GRAPH = False
envs, feeders, per = bench.make_replicas(N, "cuda:0", 0, "rotate")
for e, f in zip(envs, feeders):
    e.step(f.synthetic_actions); e.step(f.synthetic_actions)
torch.cuda.synchronize()
st_obj = torch.cuda.Stream()
def graphs(fn):
    gs = []
    with torch.cuda.stream(st_obj):
        for e, f in zip(envs, feeders):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st_obj):
                fn(e, f, st_obj.cuda_stream)
            gs.append(g)
    return gs
def tq(e, f, st, n=4):
    e._tq_params.actions_in = f.synthetic_actions.data_ptr(); e._tq_params.actions_clipped = None
    for _ in range(n): nat.check(nat.lib.lgk_compute_torques(C.byref(e._tq_params), st))
def pp(e, f, st):
    e._params.phase_mask = 3
    nat.check(nat.lib.lgk_post_physics(C.byref(e._params), st))
def fin(e, f, st):
    e._finalize(st, advance=1)
def full(e, f, st):
    tq(e, f, st); pp(e, f, st); fin(e, f, st)
def timeit(name, gs, reps=300):
    with torch.cuda.stream(st_obj):
        for i in range(3 * len(gs)): gs[i % len(gs)].replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st_obj)
        for i in range(reps): gs[i % len(gs)].replay()
        b.record(st_obj); torch.cuda.synchronize()
    print(f"{name:28s} {a.elapsed_time(b) / reps * 1e3:8.2f} us", flush=True)
for pdl in (1, 0):
    nat.lib.lgk_set_pdl(pdl)
    print("PDL", pdl, "N", N)
    timeit("empty-ish (finalize only)", graphs(fin))
    timeit("1 x torque", graphs(lambda e, f, s: tq(e, f, s, 1)))
    timeit("4 x torque", graphs(tq))
    timeit("post_physics (K1+K2)", graphs(pp))
    timeit("full step (7 kernels)", graphs(full))
