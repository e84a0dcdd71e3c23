"""K1+K2 time with and without resetting envs (is the out-of-line reset path what stretches phase B under load?)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from legged_games_gym_b200 import _native as nat
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for label, pterm in (("bench reset rate", bench.P_TERMINATE), ("no resets", 0.0)):
    bench.P_TERMINATE = pterm
    envs, feeders, per = bench.make_replicas(n, "cuda:0", 0, "rotate")
    if pterm == 0.0:
        for e in envs:
            e.episode_length_buf.zero_()
    st = torch.cuda.current_stream().cuda_stream
    fns = []
    for env in envs:
        env._params.phase_mask = nat.PHASE_PRE | nat.PHASE_POST
        fns.append(lambda e=env: nat.lib.lgk_post_physics(C.byref(e._params), st))
    mean_s, best_s = bench.time_kernel(fns, 60)
    torch.cuda.synchronize()
    print(label, f"K1+K2 {mean_s * 1e6:.1f} us", "resets/step", float(envs[0].reset_buf.float().mean()), flush=True)
    del envs, feeders, fns
    torch.cuda.empty_cache()
