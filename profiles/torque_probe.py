"""LSTM torque kernel variants, launch time per size (replicas larger than L2, back-to-back launches through the C ABI).
    python profiles/torque_probe.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from legged_games_gym_b200 import _native as nat  # noqa: E402

dev = "cuda:0"
bench.USE_GRAPH = True
for n in (4096, 16384, 65536):
    envs, feeders, _ = bench.make_replicas(n, dev, 0, "rotate")
    st = torch.cuda.current_stream().cuda_stream
    for variant in [int(v) for v in os.environ.get("VARIANTS", "1,3,2").split(",")]:
        fns = []
        for env, f in zip(envs, feeders):
            env._tq_params.actions_in = f.synthetic_actions.data_ptr()
            env._tq_params.actions_clipped = None
            env._tq_params.lstm_variant = variant
            fns.append(lambda e=env: nat.lib.lgk_compute_torques(C.byref(e._tq_params), st))
        mean_s, best_s = bench.time_kernel(fns, 200)
        gbs = bench.BYTES_TORQUE_LSTM * n / mean_s / 1e9
        print(f"n={n:6d} variant {variant}: {mean_s * 1e6:7.2f} us (best batch {best_s * 1e6:7.2f})  {gbs:7.1f} GB/s", flush=True)
    del envs, feeders
    torch.cuda.empty_cache()
