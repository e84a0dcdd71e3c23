"""Per-CTA entry / exit stamps of K1 (build with -DLGK_EXP_CTA_STAMPS): python profiles/k1_cta_probe.py 65536"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from legged_games_gym_b200 import _native as nat  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
bench.USE_GRAPH = False
envs, feeders, per = bench.make_replicas(N, "cuda:0", 0, "rotate")
grid = (N + 31) // 32
tl = torch.zeros(32 + 8 * grid, dtype=torch.int64, device="cuda:0")
st = torch.cuda.current_stream().cuda_stream
for e, f in zip(envs, feeders):
    e._params.phase_mask = 3
    for _ in range(2):
        e.step(f.synthetic_actions)
nat.check(nat.lib.lgk_step_debug_timeline(tl.data_ptr()))
for rep in range(3):
    for e in envs:
        nat.lib.lgk_post_physics(C.byref(e._params), st)
    torch.cuda.synchronize()
t = tl.cpu().numpy()[32:].reshape(grid, 8)
start, end, sm = t[:, 0], t[:, 1], t[:, 7]
t0 = start.min()
life = (end - start) / 1e3
print(f"lib {nat.LIB_PATH}\nN {N} grid {grid}: span {(end.max() - t0) / 1e3:.2f} us; CTA life mean {life.mean():.2f} p10 {np.percentile(life, 10):.2f} "
      f"p50 {np.percentile(life, 50):.2f} p90 {np.percentile(life, 90):.2f} max {life.max():.2f} us")
for name, a, b in (("entry->pdl_wait done", 0, 2), ("pdl_wait->staged", 2, 3), ("staged->phase C done", 3, 4), ("phase C done->exit", 4, 1)):
    d = (t[:, b] - t[:, a]) / 1e3
    print(f"  {name}: mean {d.mean():.2f} p50 {np.percentile(d, 50):.2f} p90 {np.percentile(d, 90):.2f} max {d.max():.2f} us")
rel = (start - t0) / 1e3
print("entry times us: p10 %.2f p50 %.2f p90 %.2f max %.2f" % tuple(np.percentile(rel, q) for q in (10, 50, 90, 100)))
per_sm = np.bincount(sm.astype(int))
print("CTAs per SM: min %d max %d; SMs used %d" % (per_sm[per_sm > 0].min(), per_sm.max(), (per_sm > 0).sum()))
# concurrency on one SM: sort its CTAs by entry
s0 = int(sm[0])
idx = np.where(sm == s0)[0]
o = idx[np.argsort(start[idx])]
print("SM", s0, "CTA (entry, exit) us:", [(round((start[i] - t0) / 1e3, 1), round((end[i] - t0) / 1e3, 1)) for i in o])
nat.lib.lgk_step_debug_timeline(None)
