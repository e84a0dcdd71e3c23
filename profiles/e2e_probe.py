"""Where does the pipelined e2e step spend its time?  A/B legs around bench.py's e2e loop (4096 envs):

  python profiles/e2e_probe.py

  serial            : upload, step, download, host sync (round 1's loop)
  pipe              : bench.py's pipelined loop (HostResultMirror)
  pipe_nodl         : the same without the device->host copies (snapshot only): GPU + host launch cost, no D2H traffic
  pipe_devstate     : pipelined, sim state resident in HBM (no sysmem reads inside the kernels), downloads on
  step_only         : upload + step, nothing read back (unified host state)
  host_launch       : CPU time per iteration of the pipelined loop when nothing is waited for
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from legged_games_gym_b200.sim.result_mirror import HostResultMirror  # noqa: E402

DEV = "cuda:0"
N = int(os.environ.get("N", 4096))
STEPS = 200
bench.USE_GRAPH = True


def timed(fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    fn()
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / STEPS, (t1 - t0) * 1e6 / STEPS


def run(host_sim, download=True, lag=2, label=""):
    env, f = bench.make_env(N, DEV, host_sim=host_sim)
    h_act = f.synthetic_actions.cpu().pin_memory()
    d_act = env.action_buffer
    mirror = HostResultMirror(env, depth=lag + 1)
    if not download:
        def push_nodl():
            k = mirror.pushed
            s = k % mirror.depth
            if not mirror.snap[s]:
                mirror._bind()
            for n in mirror.names:
                mirror.snap[s][n].copy_(getattr(env, n), non_blocking=True)
            mirror.pushed += 1
            return k
        mirror.push = push_nodl
        mirror.wait = lambda k: None

    def loop(steps, wait=True):
        k = -1
        for _ in range(steps):
            d_act.copy_(h_act, non_blocking=True)
            env.step(d_act)
            k = mirror.push()
            if wait and k >= lag:
                mirror.wait(k - lag)
        torch.cuda.synchronize()

    loop(8)
    gpu_us, _ = timed(lambda: loop(STEPS))
    _, host_us = timed(lambda: loop(STEPS, wait=False)) if download else (0, 0)
    print(f"{label:16s} gpu {gpu_us:7.1f} us/step   host launch-only {host_us:7.1f} us/iter", flush=True)
    del env, f, mirror
    torch.cuda.empty_cache()


def serial(host_sim=True, readback=True, label="serial"):
    env, f = bench.make_env(N, DEV, host_sim=host_sim)
    h_act = f.synthetic_actions.cpu().pin_memory()
    d_act = env.action_buffer
    h_obs = torch.empty(env.obs_buf.shape).pin_memory()

    def loop(steps):
        for _ in range(steps):
            d_act.copy_(h_act, non_blocking=True)
            obs = env.step(d_act)[0]
            if readback:
                h_obs.copy_(obs, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        torch.cuda.synchronize()

    loop(8)
    gpu_us, host_us = timed(lambda: loop(STEPS))
    print(f"{label:16s} gpu {gpu_us:7.1f} us/step   host {host_us:7.1f} us/iter", flush=True)
    del env, f
    torch.cuda.empty_cache()


if __name__ == "__main__":
    if os.environ.get("LEGS") == "memcpy":
        os.environ["LGK_HOST_UNIFIED"] = "0"
        os.environ["LGK_HOST_ZERO_COPY"] = "0"
        from legged_games_gym_b200.sim.state_feeder import HostStateFeeder
        HostStateFeeder.KERNEL_COPY_MAX_BYTES = 0
        serial(label="serial_memcpy")
        serial(readback=False, label="step_only_memcpy")
        run(True, label="pipe_memcpy")
        sys.exit(0)
    serial()
    serial(readback=False, label="step_only")
    serial(host_sim=False, readback=False, label="step_only_dev")
    run(True, label="pipe")
    run(True, lag=1, label="pipe_lag1")
    run(True, download=False, label="pipe_nodl")
    run(False, label="pipe_devstate")
    os.environ["LGK_HOST_UNIFIED"] = "0"
    run(True, label="pipe_explicit")
    serial(label="serial_explicit")
