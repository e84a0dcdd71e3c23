"""Does this host's link run both directions at once?  64 MiB pinned copies: each direction alone, then both on two streams.
    python profiles/pcie_duplex_probe.py
"""
import torch

dev = "cuda:0"
n = 64 << 20
h1, h2 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d1, d2 = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(do_h2d, do_d2h, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_event(a); s2.wait_event(a)
        if do_h2d:
            with torch.cuda.stream(s1):
                d1.copy_(h1, non_blocking=True)
        if do_d2h:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
        cur = torch.cuda.current_stream()
        cur.wait_stream(s1); cur.wait_stream(s2)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 1e3)
    return best


t_h, t_d, t_b = run(True, False), run(False, True), run(True, True)
print(f"h2d alone {n / t_h / 1e9:.1f} GB/s   d2h alone {n / t_d / 1e9:.1f} GB/s   both at once: {t_b * 1e3:.2f} ms "
      f"= {n / t_b / 1e9:.1f} GB/s per direction ({2 * n / t_b / 1e9:.1f} GB/s total; alone back to back would take {(t_h + t_d) * 1e3:.2f} ms)")
