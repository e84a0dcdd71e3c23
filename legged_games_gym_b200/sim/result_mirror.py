"""Pipelined host copies of a step's results.

The reference hands ``obs, rew, reset`` of ``LeggedRobot.step`` (legged_robot.py:80-104) to a policy that lives on the
same device; a host-side consumer (logger, recorder, a CPU-pipeline simulator) only needs them eventually.  Downloading
the 3.85 MB observation matrix of 4096 envs on the launch stream costs 70 us of PCIe time per step during which the GPU
idles.  ``HostResultMirror`` takes the download off the step's critical path:

* ``push()`` -- right after ``env.step``: a device-to-device snapshot of the result tensors into slot ``k % depth`` on the
  launch stream (3 us; the env rewrites its buffers in place every step), then the device-to-host copy of that slot on
  a side stream into pinned memory;
* ``wait(k)`` -- blocks the host until step k's results are in host memory and returns the pinned tensors.

Step k's download overlaps step k+1's kernels (PCIe is full duplex: the step pulls its inputs host-to-device while the
copy engine pushes the previous results device-to-host).  Every step's results still cross PCIe; nothing is skipped.
"""
import torch


class HostResultMirror:
    def __init__(self, env, depth=2, with_privileged=False):
        if depth < 2:
            raise ValueError("HostResultMirror needs at least two slots (one downloading, one being filled)")
        self.env = env
        self.depth = depth
        dev = env.obs_buf.device
        names = ["obs_buf", "rew_buf", "reset_buf"]
        if with_privileged and env.privileged_obs_buf is not None:
            names.append("privileged_obs_buf")
        self.names = names
        self._src = {n: getattr(env, n) for n in names}
        # reset_buf becomes the persistent bool buffer after the first step: bind lazily in push()
        self.snap = [dict() for _ in range(depth)]
        self.host = [dict() for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.filled = [torch.cuda.Event() for _ in range(depth)]     # snapshot of slot s complete (launch stream)
        self.done = [torch.cuda.Event() for _ in range(depth)]       # download of slot s complete (copy stream)
        self._pending = [False] * depth
        self.pushed = 0
        self.bytes_per_push = 0

    def _bind(self):
        self.bytes_per_push = 0
        for n in self.names:
            t = getattr(self.env, n)
            self._src[n] = t
            for s in range(self.depth):
                if n not in self.snap[s] or self.snap[s][n].dtype != t.dtype or self.snap[s][n].shape != t.shape:
                    self.snap[s][n] = torch.empty_like(t)
                    self.host[s][n] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            self.bytes_per_push += t.numel() * t.element_size()

    def push(self):
        """Snapshot the env's current results and start their download; returns the step index to pass to wait()."""
        k = self.pushed
        s = k % self.depth
        if any(getattr(self.env, n) is not self._src[n] for n in self.names) or not self.snap[s]:
            self._bind()
        cur = torch.cuda.current_stream()
        if self._pending[s]:
            cur.wait_event(self.done[s])          # the slot's previous download must have left the device buffer
        for n in self.names:
            self.snap[s][n].copy_(self._src[n], non_blocking=True)
        self.filled[s].record(cur)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.filled[s])
            for n in self.names:
                self.host[s][n].copy_(self.snap[s][n], non_blocking=True)
            self.done[s].record(self.copy_stream)
        self._pending[s] = True
        self.pushed += 1
        return k

    def wait(self, k):
        """Host-side wait for the results of push number k; returns {name: pinned host tensor} (valid until push k+depth)."""
        if not (self.pushed - self.depth <= k < self.pushed):
            raise IndexError(f"results of push {k} are no longer (or not yet) held: {self.pushed} pushes, depth {self.depth}")
        s = k % self.depth
        self.done[s].synchronize()
        return self.host[s]

    def drain(self):
        for s in range(self.depth):
            if self._pending[s]:
                self.done[s].synchronize()
