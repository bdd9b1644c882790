"""Host-fed inference with several batches in flight (the loop of reference run.py:137-180 `run.val`, non-force branch).

`model(batch)` has one host synchronisation (the edge / triplet counts size the buffers of the interaction kernels).
In a plain loop the GPU idles from that point of batch n+1 back to the end of batch n's readback.  Here consecutive
batches alternate between `depth` CUDA streams: the H2D copy, the graph kernels and the count readback of batch n+1 are
enqueued on the other stream while batch n's interaction blocks still run, and the energies of batch n are read
(from a pinned buffer) only when they are needed.  Every batch is still copied from host memory and every result is
still read on the host; only the order of the waits changes.  Results are bit-identical to the plain loop (same
kernels, same inputs; the streams share nothing but the read-only weights).
"""
import torch

# Measured on the B200, SphereNet QM9-shape batches of 128 (profiles/r02_batches_in_flight_final.txt): 1 batch at a time
# 109.5 k molecules/s, 2 in flight 152.0 k, 3: 162.8 k, 4: 163.9 k, 6: 164.1 k -- the other streams fill one stream's launch
# gaps, partial waves and count-readback bubble; beyond four there is nothing left to fill.  (Before the host work per
# forward was cut, profiles/r02_batches_in_flight.txt, the loop was host-bound and three were enough: 133.4 k.)
DEFAULT_DEPTH = 4


class InferencePipeline:
    def __init__(self, model, device, depth=DEFAULT_DEPTH):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.model, self.device, self.depth = model, torch.device(device), depth
        self.streams = [torch.cuda.Stream(self.device) for _ in range(depth)]
        self._slots = [None] * depth          # (event, pinned host buffer, device tensors kept alive)
        self._host = [None] * depth           # pinned result buffers, reused while the output shape stays the same
        self._n = 0

    def submit(self, host_batch):
        """Enqueue one batch (pinned host tensors give a truly asynchronous copy); returns its ticket."""
        slot = self._n % self.depth
        if self._slots[slot] is not None and self._slots[slot][0] is not None:
            raise RuntimeError("InferencePipeline: result() of the batch submitted `depth` tickets ago was never taken")
        st = self.streams[slot]
        st.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(st), torch.no_grad():
            db = host_batch.to(self.device, non_blocking=True)
            out = self.model(db)
            host = self._host[slot]
            if host is None or host.shape != out.shape or host.dtype != out.dtype:
                host = self._host[slot] = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            host.copy_(out, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(st)
        self._slots[slot] = (ev, host, (db, out))
        self._n += 1
        return self._n - 1

    def result(self, ticket):
        """Host tensor with the energies of `ticket` (waits for that batch only)."""
        slot = ticket % self.depth
        ev, host, _keep = self._slots[slot]
        if ev is None or ticket < self._n - self.depth or ticket >= self._n:
            raise RuntimeError(f"InferencePipeline: ticket {ticket} is not in flight")
        ev.synchronize()
        self._slots[slot] = (None, None, None)
        return host                            # valid until this slot's next submit(): copy it to keep it longer

    def map(self, host_batches):
        """Energies (host tensors) of an iterable of batches, in order, with `depth` batches in flight."""
        pending = []
        for hb in host_batches:
            pending.append(self.submit(hb))
            if len(pending) == self.depth:
                yield self.result(pending.pop(0))
        while pending:
            yield self.result(pending.pop(0))
