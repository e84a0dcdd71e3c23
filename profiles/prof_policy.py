"""Profiling target for the tcgen05 policy kernel: a few lgk_policy_act calls through the Python ActorCritic.
    python profiles/prof_policy.py --num-envs 65536
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from legged_games_gym_b200.rsl_rl.modules import ActorCritic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--num-envs", type=int, default=4096)
ap.add_argument("--calls", type=int, default=4)
a = ap.parse_args()
torch.manual_seed(0)
ac = ActorCritic(235, 235, 12, [512, 256, 128], [512, 256, 128]).to("cuda:0")
obs = torch.randn(a.num_envs, 235, device="cuda:0")
with torch.inference_mode():
    for i in range(a.calls):
        ac.set_rng(1, i)
        out = ac.act_and_evaluate(obs, obs)["actions"]      # PPO.act's call: actor + critic in one launch
torch.cuda.synchronize()
print("ok", a.num_envs, float(out.mean()))
