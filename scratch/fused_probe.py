"""A/B of lgk_post_physics: two-kernel chain vs the fused kernel at several scan-warp counts (launches back to back on
replicas larger than L2, like bench.kernel_rooflines)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from legged_games_gym_b200 import _native as nat

dev = "cuda:0"
torch.cuda.set_device(0)
sizes = [int(x) for x in (sys.argv[1:] or ["4096", "16384", "65536"])]
for n in sizes:
    envs, feeders, per = bench.make_replicas(n, dev, 0, "rotate")
    st = torch.cuda.current_stream().cuda_stream
    fns = []
    for env, f in zip(envs, feeders):
        env._params.phase_mask = nat.PHASE_PRE | nat.PHASE_POST
        fns.append(lambda e=env: nat.lib.lgk_post_physics(C.byref(e._params), st))
    reps = 200 if n <= 16384 else 60
    res = {}
    for label, fused, sw in ((os.environ.get("LABEL", "chain"), 0, 0),):
        nat.lib.lgk_set_fused(fused)
        nat.lib.lgk_set_fused_scan_warps(sw)
        mean_s, best_s = bench.time_kernel(fns, reps)
        res[label] = (mean_s * 1e6, best_s * 1e6)
    print(n, {k: f"{v[0]:.1f}/{v[1]:.1f} us" for k, v in res.items()}, "frac of HBM at best:",
          {k: round(bench.BYTES_POST_ROUGH * n / (v[0] * 1e-6) / 1e9 / 6552, 3) for k, v in res.items()}, flush=True)
    nat.lib.lgk_set_fused(1); nat.lib.lgk_set_fused_scan_warps(0)
    del envs, feeders, fns
    torch.cuda.empty_cache()
