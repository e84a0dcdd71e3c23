"""act() timing + phase timeline of the tcgen05 policy kernel (CTA (0,0) stamps) and its error against torch fp32.
    python profiles/policy_probe.py [num_envs ...]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from legged_games_gym_b200 import _native as nat  # noqa: E402
from legged_games_gym_b200.rsl_rl.modules import ActorCritic  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [4096, 16384, 65536]
torch.manual_seed(0)
dev = "cuda:0"
FLOP = 2 * 2 * (235 * 512 + 512 * 256 + 256 * 128 + 128 * 6.5)
for hidden, nobs in (((512, 256, 128), 235), ((128, 64, 32), 48)):
    ac = ActorCritic(nobs, nobs, 12, list(hidden), list(hidden)).to(dev)
    for n in sizes:
        obs = torch.randn(n, nobs, device=dev)
        tl = torch.zeros(128, dtype=torch.int64, device=dev)
        with torch.inference_mode():
            ac.set_rng(1, 0)
            out = ac.act_and_evaluate(obs, obs)
            torch.cuda.synchronize()
            mu_ref = ac.actor(obs)
            v_ref = ac.critic(obs)
            err_mu = float((ac.action_mean - mu_ref).abs().max())
            err_v = float((out["values"] - v_ref).abs().max())
            p = ac._last_params                      # nets = 3: actor + critic, the launch PPO.act makes
            st = torch.cuda.current_stream().cuda_stream
            for _ in range(20):
                nat.lib.lgk_policy_act(C.byref(p), st)
            for fl in (1, 2, 3):                    # experiments: no weight copies / no MMAs / neither
                nat.lib.lgk_policy_debug_timeline(tl.data_ptr(), fl)
                nat.lib.lgk_policy_act(C.byref(p), st)
                torch.cuda.synchronize()
                t = tl.cpu().tolist()
                print(f"   flags={fl}: timeline(ns) {[x - t[0] for x in t[:14]]}  issuer cycles wait_act/wait_tile/issue/commit {t[100:104]}")
            nat.lib.lgk_policy_debug_timeline(tl.data_ptr(), 0)
            nat.lib.lgk_policy_act(C.byref(p), st)
            torch.cuda.synchronize()
            nat.lib.lgk_policy_debug_timeline(None, 0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 100
            a.record()
            for _ in range(reps):
                nat.lib.lgk_policy_act(C.byref(p), st)
            b.record()
            torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / reps
        t = tl.cpu().tolist()
        rel = [x - t[0] for x in t[:16]]
        print(f"   issuer cycles wait_act/wait_tile/issue/commit {t[100:104]}  tile starts (cycles) {[x - t[16] for x in t[16:16 + 40]]}")
        flop = n * (FLOP if hidden[0] == 512 else 2 * 2 * (48 * 128 + 128 * 64 + 64 * 32 + 32 * 6.5))
        print(f"hidden={hidden} n={n}: {us:7.1f} us/act  {flop / us / 1e6:7.1f} TFLOP/s  max|err| mu {err_mu:.2e} v {err_v:.2e}  timeline(ns) {rel}", flush=True)
