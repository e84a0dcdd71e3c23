import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import harness
from tests.util import product_env, feeder_state
from legged_games_gym_b200 import _native as nat
lib = nat.lib
task, n, ov = "anymal_c_rough", 4096, {"env.episode_length_s": 0.3, "domain_rand.push_interval_s": 0.04}
case = harness.build_case(task, n, seed=11, overrides=ov)
envs = []
for fused in (1, 0):
    lib.lgk_set_fused(fused)
    env, feeder = product_env(case)
    envs.append((fused, env, feeder_state(feeder)))
for step in range(1, 5):
    acts = torch.from_numpy(np.random.default_rng(step).normal(0, 1, (n, 12)).astype(np.float32)).to("cuda:0")
    noise = harness.make_noise(case, step, 5)
    snaps = []
    for fused, env, st in envs:
        lib.lgk_set_fused(fused)
        env.step(acts.clone()); torch.cuda.synchronize()
        snaps.append(harness.snapshot(env))
        harness.apply_noise(st, noise)
    a, b = snaps
    for k in b:
        if not torch.equal(a[k], b[k]):
            d = (a[k] != b[k])
            if d.dim() == 2:
                rows = d.any(1).nonzero().flatten(); cols = d.any(0).nonzero().flatten()
                print(step, k, "rows differing", rows.numel(), rows[:10].tolist(), "cols", cols[:20].tolist(), cols.numel(),
                      "resets among rows:", int(a["reset_buf"][rows].sum()), "max abs diff", float((a[k] - b[k]).abs().max()))
            else:
                print(step, k, int(d.sum()))
    print("step", step, "resets", int(a["reset_buf"].sum()))
