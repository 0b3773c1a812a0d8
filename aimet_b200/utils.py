"""Small host-side helpers for feeding the calibration loop."""
import torch


class DevicePrefetcher:
    """Iterate over host batches as device tensors, copying batch i+1 on a separate CUDA stream into a second staging
    buffer while the caller is still computing on batch i (classic double buffering; `depth` batches ahead need
    `depth + 1` buffers). The host tensors should be pinned, otherwise the copies are synchronous.

        for x in DevicePrefetcher(pinned_batches, device):
            model(x)

    The yielded tensor is one of the staging buffers: it is valid on the stream that is current when it is yielded and
    until the loop asks for the next batch -- which is what a calibration forward needs. The buffers are allocated once
    (no allocator traffic per step); a batch whose shape or dtype differs from the first one is copied the plain way."""

    def __init__(self, batches, device, depth: int = 1):
        self.batches = batches
        self.device = torch.device(device)
        self.depth = max(1, int(depth))

    def __iter__(self):
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher stages batches onto a CUDA device")
        copy_stream = torch.cuda.Stream(self.device)
        source = iter(self.batches)
        slots = self.depth + 1
        buffers = [None] * slots
        ready = [torch.cuda.Event() for _ in range(slots)]
        released = [None] * slots          # recorded on the consumer's stream once it has moved past the buffer
        pending = []                       # (slot or None, tensor) in issue order
        issued = 0

        def enqueue():
            nonlocal issued
            try:
                host = next(source)
            except StopIteration:
                return False
            slot = issued % slots
            if buffers[slot] is None and (issued < slots):
                first = buffers[0]
                if first is None or (first.shape == host.shape and first.dtype == host.dtype):
                    buffers[slot] = torch.empty(host.shape, dtype=host.dtype, device=self.device)
            buf = buffers[slot]
            if buf is None or buf.shape != host.shape or buf.dtype != host.dtype:
                pending.append((None, host))          # odd one out: copied on the consumer's stream when its turn comes
            else:
                with torch.cuda.stream(copy_stream):
                    if released[slot] is not None:
                        copy_stream.wait_event(released[slot])
                    buf.copy_(host, non_blocking=True)
                    ready[slot].record(copy_stream)
                pending.append((slot, buf))
            issued += 1
            return True

        for _ in range(self.depth):
            if not enqueue():
                break
        while pending:
            slot, tensor = pending.pop(0)
            current = torch.cuda.current_stream(self.device)
            if slot is None:
                tensor = tensor.to(self.device, non_blocking=True)
            else:
                current.wait_event(ready[slot])
            enqueue()
            yield tensor
            if slot is not None:
                event = torch.cuda.Event()
                event.record(torch.cuda.current_stream(self.device))
                released[slot] = event
