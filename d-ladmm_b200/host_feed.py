"""Host -> device feed of observation batches for inference over data that lives in host memory.

The reference's drivers slice a host-side numpy array per mini-batch and upload it synchronously before each
forward (`torch.from_numpy(...).cuda()`, main_syn_l1l1_scalar.py:264-268, 319-323).  Batches of problem instances are
independent, so the upload of batch i+1 can run on a copy stream while the K-layer forward of batch i computes:
`HostFeed` is that double buffer.  Every batch is still copied host -> device once, from pinned memory; only the
waiting is hidden.
"""
import torch


class HostFeed(object):
    """Iterate over `batches` (an iterable of (m, B) float32 host tensors, pinned for asynchronous copies) and yield
    each one as a device tensor.  The copy of the next batch is enqueued on a side stream before the current one is
    handed out.  A yielded tensor is valid until the next `next()` call on the feed (its buffer is then recycled)."""

    def __init__(self, batches, device, depth=2):
        if depth < 2:
            raise ValueError("depth must be >= 2 (one buffer in use, one in flight)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostFeed needs a CUDA device: d-ladmm_b200 has no CPU path")
        self._it = iter(batches)
        self._depth = depth
        self._copy_stream = torch.cuda.Stream(self.device)
        self._bufs = [None] * depth
        self._ready = [None] * depth          # event: copy into buffer j finished
        self._released = [None] * depth       # event: consumer's work on buffer j was enqueued before this point
        self._queue = []                      # buffer indices holding uploaded batches, oldest first
        self._next_buf = 0
        self._current = None
        self.bytes_copied = 0

    def _enqueue_one(self):
        try:
            host = next(self._it)
        except StopIteration:
            return False
        if host.dtype != torch.float32 or host.dim() != 2:
            raise RuntimeError("HostFeed batches must be (m, B) float32 tensors")
        j = self._next_buf
        self._next_buf = (j + 1) % self._depth
        buf = self._bufs[j]
        if buf is None or buf.shape != host.shape:
            # A fresh block comes from the CONSUMER stream's pool: kernels enqueued there may still be using the memory it
            # was carved from, so the first copy must wait for that stream; and the block is written on the copy stream, so
            # the allocator must not hand it out again (after the feed drops it) before the copy stream is done with it.
            cur = torch.cuda.current_stream(self.device)
            buf = self._bufs[j] = torch.empty(host.shape, dtype=torch.float32, device=self.device)
            buf.record_stream(self._copy_stream)
            self._copy_stream.wait_stream(cur)
            self._released[j] = None
        with torch.cuda.stream(self._copy_stream):
            if self._released[j] is not None:
                self._copy_stream.wait_event(self._released[j])     # do not overwrite a buffer a forward still reads
            buf.copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._ready[j] = ev
        self._queue.append(j)
        self.bytes_copied += host.numel() * 4
        return True

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        if self._current is not None:         # everything the consumer enqueued on the last batch precedes this event
            ev = torch.cuda.Event()
            ev.record(cur)
            self._released[self._current] = ev
            self._current = None
        while len(self._queue) < self._depth - 1 and self._enqueue_one():
            pass
        if not self._queue:
            raise StopIteration
        j = self._queue.pop(0)
        self._enqueue_one()                   # start the next upload before handing this batch out
        cur.wait_event(self._ready[j])
        self._current = j
        return self._bufs[j]


class HostDrain(object):
    """Device -> host counterpart of HostFeed: results of step i are copied to pinned host memory on a side stream while step
    i+1 computes.

        drain = HostDrain(device)
        for x in HostFeed(batches, device):
            Z = model(x, last_only=True)[0][0]
            host_z = drain.push(Z)          # -> the pinned host tensor of the step `depth` pushes ago, complete, or None
        for host_z in drain.flush(): ...    # the last `depth` results

    `push` snapshots nothing: it enqueues the copy of `t` after the producer's work (an event on the current stream) and the
    caller must not overwrite `t` before the copy ran -- the library's outputs are fresh tensors every call, so that holds
    for them by construction (`record_stream` keeps the allocator from recycling the block early)."""

    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostDrain needs a CUDA device: d-ladmm_b200 has no CPU path")
        self._depth = max(1, int(depth))
        self._copy_stream = torch.cuda.Stream(self.device)
        self._host = [None] * self._depth
        self._done = [None] * self._depth
        self._n = 0
        self.bytes_copied = 0

    def push(self, t):
        j = self._n % self._depth
        out = None
        if self._done[j] is not None:           # the copy that used this pinned buffer `depth` pushes ago
            self._done[j].synchronize()
            out = self._host[j]
            self._host[j] = None
        buf = self._pool_take(t)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ev)
            buf.copy_(t, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        t.record_stream(self._copy_stream)
        self._host[j], self._done[j] = buf, done
        self._n += 1
        self.bytes_copied += t.numel() * t.element_size()
        return out

    def _pool_take(self, t):
        """A pinned host buffer of t's shape: recycled from `recycle()` when possible (pinning memory is slow)."""
        pool = self.__dict__.setdefault("_pool", [])
        for i, b in enumerate(pool):
            if b.shape == t.shape and b.dtype == t.dtype:
                return pool.pop(i)
        return torch.empty(t.shape, dtype=t.dtype, pin_memory=True)     # allocated pinned (no pageable copy first)

    def recycle(self, host_tensor):
        """Give a host tensor returned by push()/flush() back once it has been consumed."""
        self.__dict__.setdefault("_pool", []).append(host_tensor)

    def flush(self):
        """Wait for the outstanding copies; returns their host tensors, oldest first."""
        outs = []
        for i in range(self._depth):
            j = (self._n + i) % self._depth
            if self._done[j] is not None:
                self._done[j].synchronize()
                outs.append(self._host[j])
                self._host[j] = self._done[j] = None
        return outs
