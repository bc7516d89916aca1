"""Serving loop around a CompiledModel: host clips in, host logits out, copies overlapped with compute.

The reference's inference loop (``model(inputs.to(device))`` then ``preds.cpu()``) serialises
H2D -> forward -> D2H.  On a B200 the forward of a SlowFast batch and the PCIe copy of its fp32
clips cost about the same, so the public serving call double-buffers them: a copy stream moves
batch i+1 into a staging buffer while the graph of batch i runs; the logits of batch i come back
on the compute stream behind the graph.  Every batch is still copied in and read out - only the
ordering changes.  One process per GPU; nothing here is collective.
"""
import torch


class ClipPipeline:
    def __init__(self, compiled, depth=2):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.cm = compiled
        self.depth = depth
        dev = compiled.plan.device
        self.dev = dev
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.staging = [[torch.empty_like(s) for s in compiled.static_in] for _ in range(depth)]
        self.h2d_done = [torch.cuda.Event() for _ in range(depth)]
        self.staging_free = [torch.cuda.Event() for _ in range(depth)]
        self.out_done = [torch.cuda.Event() for _ in range(depth)]
        self.host_out = [torch.empty(tuple(compiled.out_shape), dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.n_submitted = 0
        self.post = None          # optional callable(device_logits) -> device tensor to read back (e.g. all-gather)

    def submit(self, host_inputs):
        """Enqueue one batch of pinned HOST clips; returns a ticket for ``result``."""
        cm = self.cm
        ins = list(host_inputs) if cm.multi else [host_inputs]
        if len(ins) != len(cm.static_in):
            raise RuntimeError("expected %d input tensors, got %d" % (len(cm.static_in), len(ins)))
        slot = self.n_submitted % self.depth
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy_stream):
            if self.n_submitted >= self.depth:
                self.copy_stream.wait_event(self.staging_free[slot])
            for dst, src in zip(self.staging[slot], ins):
                if tuple(src.shape) != tuple(dst.shape):
                    raise RuntimeError("input shape %s differs from the compiled shape %s" % (tuple(src.shape), tuple(dst.shape)))
                dst.copy_(src, non_blocking=True)
            self.h2d_done[slot].record(self.copy_stream)
        cur.wait_event(self.h2d_done[slot])
        # (the host buffer of this slot is reused: the caller has consumed result(ticket - depth) by now)
        out = cm(self.staging[slot] if cm.multi else self.staging[slot][0])     # D2D into the plan's input + graph
        self.staging_free[slot].record(cur)
        if self.post is not None:
            out = self.post(out)
        self._copy_out(slot, out)
        self.out_done[slot].record(cur)
        t = self.n_submitted
        self.n_submitted += 1
        return t

    def _copy_out(self, slot, out):
        if self.host_out[slot].shape != out.shape:
            self.host_out[slot] = torch.empty(tuple(out.shape), dtype=torch.float32).pin_memory()
        self.host_out[slot].copy_(out, non_blocking=True)

    def result(self, ticket):
        """Block until the logits of ``ticket`` are in host memory and return them (a pinned tensor
        that is reused ``depth`` submissions later)."""
        if ticket >= self.n_submitted or ticket < self.n_submitted - self.depth:
            raise RuntimeError("ticket %d is not in flight" % ticket)
        slot = ticket % self.depth
        self.out_done[slot].synchronize()
        return self.host_out[slot]

    def run(self, batches):
        """Generator: feeds host batches through the pipeline, yields host logits in order."""
        pending = []
        for b in batches:
            pending.append(self.submit(b))
            if len(pending) == self.depth:
                yield self.result(pending.pop(0)).clone()
        while pending:
            yield self.result(pending.pop(0)).clone()
