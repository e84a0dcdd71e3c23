"""Build liblgk.so for sm_100a in-tree (nvcc cross-compiles without a GPU).

    python -m legged_games_gym_b200.csrc.build [--force] [--verbose]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "liblgk.so")
SOURCES = ["lgk_abi.cu", "lgk_heights.cu", "lgk_torque.cu", "lgk_post_k1.cu", "lgk_post_physics.cu", "lgk_gae.cu", "lgk_policy.cu", "lgk_policy_tc.cu", "lgk_game.cu"]
HEADERS = ["lgk_common.cuh", "lgk_rng.cuh", "lgk_math.cuh", "lgk_step_device.cuh", "lgk_tile.cuh", "lgk_policy_common.cuh",
           "lgk_policy_tc_plan.h", "../../include/lgk.h"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


PER_FILE_FLAGS = {}


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        p = os.path.join(HERE, f)
        if os.path.exists(p):
            h.update(open(p, "rb").read())
    h.update((" ".join(FLAGS) + repr(sorted(PER_FILE_FLAGS.items()))).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    stamp = OUT + ".stamp"
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in srcs:
        o = os.path.join(HERE, "build", os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        extra = PER_FILE_FLAGS.get(os.path.basename(s), [])
        cmd = [nvcc] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(out)
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", OUT] + objs + ["-lcudart", "-lcuda"]
    subprocess.check_call(link)
    open(stamp, "w").write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
