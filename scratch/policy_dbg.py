"""GPU debugging aid for the tcgen05 policy kernel: compares TC / FP32 / torch on several shapes and times them."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.rsl_oracle import ActorCriticOracle
from oracle import philox
from legged_games_gym_b200.rsl_rl.modules import ActorCritic
from legged_games_gym_b200 import _native as nat
DEV = "cuda:0"
def run(n, nobs, hidden, variant):
    torch.manual_seed(0)
    orc = ActorCriticOracle(nobs, nobs, 12, hidden, hidden)
    with torch.no_grad(): orc.std.copy_(torch.linspace(0.5, 1.5, 12))
    ac = ActorCritic(nobs, nobs, 12, list(hidden), list(hidden)).to(DEV)
    ac.load_state_dict(orc.state_dict())
    obs = torch.randn(n, nobs) * 2
    eps = torch.from_numpy(philox.normals(17, 5, np.arange(n), 12))
    a, v, lp, mu, sg = orc.act(obs, obs, eps)
    ac.set_rng(17, 5)
    nat.lib.lgk_policy_set_variant(variant)
    o = obs.to(DEV)
    with torch.inference_mode():
        got_a = ac.act(o); got_v = ac.evaluate(o); got_lp = ac.get_actions_log_prob(got_a)
    torch.cuda.synchronize()
    e = lambda x, y: float((x.cpu() - y).abs().max())
    print(f"n={n} O={nobs} hid={hidden} variant={variant}: mu {e(ac.action_mean, mu):.2e} act {e(got_a, a):.2e} val {e(got_v, v):.2e} logp {e(got_lp, lp):.2e}  |mu|max {float(mu.abs().max()):.2f} |v|max {float(v.abs().max()):.2f}", flush=True)
    # timing
    with torch.inference_mode():
        for _ in range(3): ac.act(o)
        torch.cuda.synchronize()
        a0, b0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(20): ac.act(o)
        b0.record(); torch.cuda.synchronize()
    print(f"    {a0.elapsed_time(b0) / 20 * 1e3:.1f} us per act()", flush=True)
for shape in [(128, 48, (128, 64, 32)), (100, 48, (128, 64, 32)), (4096, 235, (512, 256, 128)), (777, 169, (512, 256, 128)), (65536, 235, (512, 256, 128))]:
    for variant in (1, 2):
        run(*shape, variant)
