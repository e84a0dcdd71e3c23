"""Profiling target: eager (non-graph) env steps so that ncu sees each kernel as its own launch.
    python profiles/prof_target.py --num-envs 65536 --steps 3 [--rotate]
--rotate steps several env replicas round-robin exactly like bench.py (their combined state exceeds L2), so every profiled
launch streams its state from HBM and its DRAM byte count is comparable with the bench's timed launches."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--num-envs", type=int, default=4096)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--task", default="anymal_c_rough")
ap.add_argument("--rotate", action="store_true")
a = ap.parse_args()
bench.USE_GRAPH = False
bench.TASK = a.task
if a.rotate:
    envs, feeders, per = bench.make_replicas(a.num_envs, "cuda:0", 0, "rotate")
else:
    e, f = bench.make_env(a.num_envs, "cuda:0")
    envs, feeders = [e], [f]
for s in range(a.steps):
    for env, feeder in zip(envs, feeders):
        env.step(feeder.synthetic_actions)
torch.cuda.synchronize()
print("ok", a.num_envs, len(envs), float(envs[0].rew_buf.mean()))
