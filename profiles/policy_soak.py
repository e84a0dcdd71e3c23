"""Soak test of the tcgen05 policy kernel: random batch sizes / network shapes / observation widths, every result against
torch fp32, plus alternating modules (the games' pattern) and weight updates between calls.
    python profiles/policy_soak.py [iterations]
"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from legged_games_gym_b200 import _native as nat  # noqa: E402
from legged_games_gym_b200.rsl_rl.modules import ActorCritic  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = "cuda:0"
random.seed(0)
torch.manual_seed(0)
nat.lib.lgk_policy_set_variant(2)
shapes = [((512, 256, 128), 235, 235, 12), ((128, 64, 32), 48, 48, 12), ((256, 128, 64), 169, 187, 16), ((512, 128, 32), 256, 33, 5),
          ((128, 64, 32), 19, 19, 3), ((256, 256, 128), 100, 100, 1)]
mods = [ActorCritic(o, c, a, list(h), list(h)).to(dev) for h, o, c, a in shapes]
worst = 0.0
for it in range(iters):
    i = random.randrange(len(mods))
    ac = mods[i]
    h, o, c, a = shapes[i]
    n = random.choice([1, 2, 31, 127, 128, 129, 1000, 4096, 5000, 20000, random.randrange(1, 70000)])
    obs, cobs = torch.randn(n, o, device=dev) * 2, torch.randn(n, c, device=dev) * 2
    if it % 7 == 3:                      # an "optimizer step" between calls: the packed weights must follow
        with torch.no_grad():
            for p in ac.parameters():
                p.add_(torch.randn_like(p) * 0.01)
    ac.set_rng(it, it)
    with torch.inference_mode():
        out = ac.act_and_evaluate(obs, cobs)
        mu = ac.actor(obs)
        v = ac.critic(cobs)
    e = max(float((out["mean"] - mu).abs().max()), float((out["values"] - v).abs().max()))
    worst = max(worst, e)
    tol = 1e-3 * (1.0 + max(float(mu.abs().max()), float(v.abs().max())))
    if not (e <= tol) or not torch.isfinite(out["actions"]).all() or not torch.isfinite(out["logp"]).all():
        print(f"FAIL at iteration {it}: shape {shapes[i]} n={n} err {e:.3e} tol {tol:.3e}")
        sys.exit(1)
torch.cuda.synchronize()
print(f"policy soak ok: {iters} calls over {len(mods)} modules, worst |err| {worst:.2e}")
