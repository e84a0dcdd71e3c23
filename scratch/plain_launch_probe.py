"""Graph replay per step vs the same 7 launches issued directly through the C ABI (PDL-chained across step boundaries)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from legged_games_gym_b200 import _native as nat
dev = "cuda:0"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
envs, feeders, per = bench.make_replicas(N, dev, 0, "rotate")
acts = [f.synthetic_actions for f in feeders]
for e, a in zip(envs, acts):
    for _ in range(3):
        e.step(a)
torch.cuda.synchronize()
def timed(fn, steps=600):
    for i in range(30): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for i in range(steps): fn(i)
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps * 1e3, (t1 - t0) / steps * 1e6
print("graph replay per step: %.2f us GPU, %.2f us CPU" % timed(lambda i: envs[i % len(envs)].step(acts[i % len(envs)])))
st = torch.cuda.current_stream().cuda_stream
lib = nat.lib
def plain(i):
    e = envs[i % len(envs)]
    tp, p = e._tq_params, e._params
    tp.actions_in = acts[i % len(envs)].data_ptr(); tp.actions_clipped = e.actions.data_ptr()
    lib.lgk_compute_torques(C.byref(tp), st)
    tp.actions_in = e.actions.data_ptr(); tp.actions_clipped = None
    lib.lgk_compute_torques(C.byref(tp), st); lib.lgk_compute_torques(C.byref(tp), st); lib.lgk_compute_torques(C.byref(tp), st)
    p.phase_mask = nat.PHASE_PRE | nat.PHASE_POST
    lib.lgk_post_physics(C.byref(p), st)
    lib.lgk_finalize_step(C.byref(p), e.reset_env_ids.data_ptr(), e.reset_count.data_ptr(), e._episode_means.data_ptr(),
                          e._time_outs_extras.data_ptr() if hasattr(e, "_time_outs_extras") else None, 1, st)
print("plain PDL-chained launches per step: %.2f us GPU, %.2f us CPU" % timed(plain))
