"""Overlap host<->device copies of one batch with the solve of another.

The reference solves one problem per call on the host, so its inputs and results never
move.  A batched device solve pays a host->device copy of the inputs and a device->host
copy of the trajectories per batch (about 44 MB for 4096 bicycle problems, N=100).  A
`SolverPipeline` keeps `depth` solver instances, each on its own CUDA stream, and hands
them out round-robin: everything enqueued on a slot (uploads, `update()`, downloads) is
ordered on that slot's stream, while consecutive batches run on different streams and so
overlap copy with compute.  Nothing here touches the kernels; it is stream plumbing over
`BatchedOptim` (whose calls run on the current torch stream).

    pipe = SolverPipeline(lambda: optimizers.trajectory_tracking_mpc_time(batch=4096), depth=2)
    for batch in batches:
        with pipe.next() as slot:            # slot.opt: the solver, slot.index: which one
            upload(slot.opt, batch)          # pinned host -> device, non_blocking=True
            slot.opt.update()
            download(slot.opt, out[slot.index])
    pipe.synchronize()
"""

from __future__ import annotations

import torch


class _Slot:
    def __init__(self, index, opt, stream):
        self.index = index
        self.opt = opt
        self.stream = stream
        self.done = torch.cuda.Event()
        self._ctx = None
        self.graph = None

    def __enter__(self):
        self._ctx = torch.cuda.stream(self.stream)
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        self.done.record(self.stream)
        ctx, self._ctx = self._ctx, None
        return ctx.__exit__(*exc)

    def capture(self, body):
        """Record everything ``body(slot)`` enqueues (uploads, ``update()``, downloads — fixed
        buffers only) into a CUDA graph of this slot; ``replay()`` then re-issues the whole step
        with one launch.  An ``update()`` is 50-80 kernel launches: enqueued one by one, the host
        limits how many batches can be in flight."""
        torch.cuda.synchronize(self.stream.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.stream):
            body(self)
        self.graph = g
        return g

    def replay(self):
        with self:
            self.graph.replay()

    def wait(self):
        """Block the host until everything enqueued on this slot has finished (needed before
        the host reads the slot's pinned output buffers)."""
        self.done.synchronize()


class SolverPipeline:
    def __init__(self, factory, depth=2, device=None):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        if not torch.cuda.is_available():
            raise RuntimeError("SolverPipeline needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.slots = []
        cur = torch.cuda.current_stream(self.device)
        for i in range(depth):
            stream = torch.cuda.Stream(self.device)
            self.slots.append(_Slot(i, factory(), stream))
            stream.wait_stream(cur)        # the constructor's fills ran on the current stream
        self._n = 0

    def __len__(self):
        return len(self.slots)

    def next(self):
        """The slot for the next batch; use as a context manager."""
        slot = self.slots[self._n % len(self.slots)]
        self._n += 1
        return slot

    def fork(self):
        """Make every slot stream wait for the work already enqueued on the current stream."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.slots:
            s.stream.wait_stream(cur)

    def join(self):
        """Make the current stream wait for every slot (e.g. before recording a timing event)."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.slots:
            cur.wait_stream(s.stream)

    def synchronize(self):
        for s in self.slots:
            s.stream.synchronize()
