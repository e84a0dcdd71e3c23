"""Build an A/B variant of liblgk.so with extra -D flags (kernel experiments): python profiles/build_variant.py NAME -DX=1 ...
The result lands in legged_games_gym_b200/csrc/build/variants/NAME.so (git-ignored, travels with gpurun); select it with
LGK_LIB_PATH."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from legged_games_gym_b200.csrc import build as B  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(B.HERE, "build", "variants")
os.makedirs(os.path.join(out_dir, name), exist_ok=True)
objs, procs = [], []
for s in B.SOURCES:
    o = os.path.join(out_dir, name, s[:-3] + ".o")
    objs.append(o)
    procs.append(subprocess.Popen(["/usr/local/cuda/bin/nvcc"] + B.FLAGS + extra + ["-c", os.path.join(B.HERE, s), "-o", o]))
assert all(p.wait() == 0 for p in procs)
out = os.path.join(out_dir, name + ".so")
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-shared", "-Wno-deprecated-gpu-targets", "-o", out] + objs + ["-lcudart", "-lcuda"])
print(out)
