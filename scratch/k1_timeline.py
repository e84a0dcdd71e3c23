import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from legged_games_gym_b200 import _native as nat
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bench.USE_GRAPH = False
envs, feeders, per = bench.make_replicas(N, "cuda:0", 0, "rotate")
tl = torch.zeros(32, dtype=torch.int64, device="cuda:0")
nat.check(nat.lib.lgk_step_debug_timeline(tl.data_ptr()))
for rep in range(3):
    for e, f in zip(envs, feeders):
        e.step(f.synthetic_actions)
    torch.cuda.synchronize()
    t = tl.cpu().tolist()
    print(N, "K1 stamps (ns from entry):", [t[i] - t[0] for i in range(9)], "last CTA (same stamps, ns from its own entry; entry rel. to CTA 0 after wait):", [t[16 + i] - t[16] for i in range(9)], t[16] - t[1], flush=True)
nat.lib.lgk_step_debug_timeline(None)
