"""Bind a rank to the host cores (and thereby the NUMA node) next to its GPU.

With one process per GPU and simulator state in pinned host memory, the ranks of a node otherwise share whichever
socket the launcher started them on: pinned buffers land on one memory controller and half of the GPUs reach them
through the inter-socket link.  Pinning the process to the GPU's own cores before the first pinned allocation places
those buffers (first touch) on the GPU's NUMA node.
"""
import os


def gpu_cpu_affinity(cuda_index):
    """Ideal host cores of CUDA device `cuda_index` as NVML reports them ([] when NVML cannot tell)."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(cuda_index)
        handle = None
        uuid = getattr(props, "uuid", None)
        if uuid is not None:
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                handle = None
        if handle is None:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = cuda_index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if cuda_index < len(ids) and ids[cuda_index].isdigit():
                    index = int(ids[cuda_index])
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        return [c for c in cpus if c < ncpu]
    except Exception:
        return []


def bind_to_gpu(cuda_index):
    """Restrict this process to the cores next to its GPU.  Returns (previous affinity, new affinity); the new set is
    empty (and nothing changed) when NVML has no answer or the set is not usable from this cgroup."""
    prev = sorted(os.sched_getaffinity(0))
    want = sorted(set(gpu_cpu_affinity(cuda_index)) & set(prev))
    if not want or want == prev:
        return prev, []
    try:
        os.sched_setaffinity(0, want)
    except OSError:
        return prev, []
    return prev, want


def restore_affinity(prev):
    try:
        os.sched_setaffinity(0, prev)
    except OSError:
        pass
