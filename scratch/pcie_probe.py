import torch, time
dev = "cuda:0"
def t_copy(nbytes, h2d, reps=50, graph=False):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    src, dst = (h, d) if h2d else (d, h)
    for _ in range(5): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if graph:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(10): dst.copy_(src, non_blocking=True)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g.replay(); torch.cuda.synchronize()
            a.record(s)
            for _ in range(reps // 10): g.replay()
            b.record(s); torch.cuda.synchronize()
        return a.elapsed_time(b) / (reps // 10 * 10) * 1e3
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): dst.copy_(src, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for nb in (4096, 16384, 196608, 393216, 851968, 3850240, 16 << 20):
    print(f"{nb:>9} B  H2D {t_copy(nb, True):7.1f} us ({nb / t_copy(nb, True) / 1e3:5.1f} GB/s)  D2H {t_copy(nb, False):7.1f} us ({nb / t_copy(nb, False) / 1e3:5.1f} GB/s)   in-graph H2D {t_copy(nb, True, graph=True):7.1f}  D2H {t_copy(nb, False, graph=True):7.1f}")

# ---- kernel copies over the unified address space (lgk_copy_from_pinned / lgk_copy_to_pinned)
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legged_games_gym_b200 import _native as nat
def t_kcopy(nbytes, h2d, graph, reps=50):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream()
    fn = (lambda: nat.lib.lgk_copy_from_pinned(d.data_ptr(), h.data_ptr(), nbytes, s.cuda_stream)) if h2d else \
         (lambda: nat.lib.lgk_copy_to_pinned(h.data_ptr(), d.data_ptr(), nbytes, s.cuda_stream))
    with torch.cuda.stream(s):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(10): fn()
            g.replay(); torch.cuda.synchronize()
            a.record(s)
            for _ in range(reps // 10): g.replay()
            b.record(s); torch.cuda.synchronize()
        else:
            a.record(s)
            for _ in range(reps): fn()
            b.record(s); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for nb in (4096, 16384, 196608, 393216, 851968, 3850240):
    print(f"{nb:>9} B  kernel H2D {t_kcopy(nb, True, False):7.1f} us  in-graph {t_kcopy(nb, True, True):7.1f}   kernel D2H {t_kcopy(nb, False, False):7.1f}  in-graph {t_kcopy(nb, False, True):7.1f}")
