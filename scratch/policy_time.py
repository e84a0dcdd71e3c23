"""Kernel-level timing of lgk_policy_act through the C ABI (no Python wrapper work inside the timed loop) + phase timeline."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from legged_games_gym_b200.rsl_rl.modules import ActorCritic
from legged_games_gym_b200 import _native as nat
DEV = "cuda:0"
FLAGS = int(os.environ.get("TC_FLAGS", "0"))
def bench(n, nobs, hidden, variant, reps=50):
    torch.manual_seed(0)
    ac = ActorCritic(nobs, nobs, 12, list(hidden), list(hidden)).to(DEV)
    nat.lib.lgk_policy_set_variant(variant)
    obs = torch.randn(n, nobs, device=DEV)
    with torch.inference_mode():
        ac.act(obs)
    p = ac._last_params
    st = torch.cuda.current_stream().cuda_stream
    tl = torch.zeros(16, dtype=torch.int64, device=DEV)
    nat.lib.lgk_policy_debug_timeline(tl.data_ptr(), FLAGS)
    for _ in range(3): nat.lib.lgk_policy_act(C.byref(p), st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): nat.check(nat.lib.lgk_policy_act(C.byref(p), st))
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / reps * 1e3
    flop = 2.0 * n * 2 * (nobs * hidden[0] + hidden[0] * hidden[1] + hidden[1] * hidden[2]) + 2.0 * n * hidden[2] * 13
    t = tl.cpu().tolist()
    d = [t[i] - t[0] for i in range(10)] if variant != 1 else []
    print(f"n={n} O={nobs} hid={hidden} variant={variant}: {us:.1f} us/call  {flop / us / 1e6:.1f} TFLOP/s  timeline(ns from setup) {d}", flush=True)
    nat.lib.lgk_policy_debug_timeline(None, 0)
for shape in [(4096, 235, (512, 256, 128)), (16384, 235, (512, 256, 128)), (65536, 235, (512, 256, 128)), (4096, 48, (128, 64, 32))]:
    for variant in (1, 2):
        bench(*shape, variant)
