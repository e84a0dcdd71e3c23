"""A/B timing of the post-physics kernels (K1 post_kernel, K2 scan_obs_fast_kernel) for one or more builds of liblgk.so.

    python profiles/pp_probe.py --libs legged_games_gym_b200/liblgk.so,/tmp/variant.so --sizes 4096,65536

Every (lib, size) pair runs in its own process (LGK_LIB_PATH is read at import).  Timing as in bench.py's
kernel_rooflines: back-to-back launches through the C ABI cycling over env replicas whose buffers exceed L2, CUDA events
on the launch stream.  LGK_PP_ONLY=1 / 2 (a timing aid of lgk_post_physics) launches K1 / K2 alone.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(n, reps, task):
    import ctypes as C
    import torch
    import bench
    from legged_games_gym_b200 import _native as nat
    bench.TASK = task
    if os.environ.get("LGK_NO_PDL"):
        nat.lib.lgk_set_pdl(0)
    envs, feeders, per = bench.make_replicas(n, "cuda:0", 0, "rotate")
    st = torch.cuda.current_stream().cuda_stream
    for e, f in zip(envs, feeders):
        e._tq_params.actions_in = f.synthetic_actions.data_ptr()
        e._tq_params.actions_clipped = None
        e._params.phase_mask = nat.PHASE_PRE | nat.PHASE_POST
        for _ in range(2):
            e.step(f.synthetic_actions)
    torch.cuda.synchronize()
    out = {"lib": nat.LIB_PATH, "n": n, "task": task, "only": os.environ.get("LGK_PP_ONLY", "0"), "pdl": 0 if os.environ.get("LGK_NO_PDL") else 1, "replicas": len(envs)}
    fns = [lambda e=e: nat.lib.lgk_post_physics(C.byref(e._params), st) for e in envs]
    mean_s, best_s = bench.time_kernel(fns, reps)
    out["us"] = round(mean_s * 1e6, 2)
    out["best_us"] = round(best_s * 1e6, 2)
    fns = [lambda e=e: nat.lib.lgk_compute_torques(C.byref(e._tq_params), st) for e in envs]
    mean_s, best_s = bench.time_kernel(fns, reps)
    out["torque_us"] = round(mean_s * 1e6, 2)
    print("PPROBE " + json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", default=os.path.join(ROOT, "legged_games_gym_b200", "liblgk.so"))
    ap.add_argument("--sizes", default="4096,65536")
    ap.add_argument("--task", default="anymal_c_rough")
    ap.add_argument("--only", default="0,1,2", help="LGK_PP_ONLY values: 0 = K1+K2, 1 = K1 alone, 2 = K2 alone")
    ap.add_argument("--reps", type=int, default=100)
    ap.add_argument("--worker", type=int, default=0)
    a = ap.parse_args()
    if a.worker:
        worker(a.worker, a.reps, a.task)
        sys.exit(0)
    for lib in a.libs.split(","):
        for n in a.sizes.split(","):
            for only in a.only.split(","):
                env = dict(os.environ, LGK_LIB_PATH=os.path.abspath(lib), LGK_PP_ONLY=only)
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", n, "--reps", str(a.reps), "--task", a.task],
                                   env=env, capture_output=True, text=True)
                lines = [l for l in r.stdout.splitlines() if l.startswith("PPROBE ")]
                print(lines[-1] if lines else f"FAILED {lib} {n} {only}: {r.stderr[-400:]}", flush=True)
