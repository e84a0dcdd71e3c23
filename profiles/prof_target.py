"""Profiling target: a handful of eager (non-graph) env steps so that ncu sees each kernel as its own launch.
    python profiles/prof_target.py --num-envs 65536 --steps 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--num-envs", type=int, default=4096)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--tile", type=int, default=0)
a = ap.parse_args()
bench.USE_GRAPH = False
env, feeder = bench.make_env(a.num_envs, "cuda:0")
for _ in range(a.steps):
    env.step(feeder.synthetic_actions)
torch.cuda.synchronize()
print("ok", a.num_envs, float(env.rew_buf.mean()))
