"""Host -> device input pipeline: overlap the H2D copy of batch i+1 with the forward of batch i.

    feeder = DeviceFeeder(batch_shape, dtype=torch.bfloat16, device="cuda:0")
    for host_batch in loader:                    # pinned host tensors
        x = feeder.push(host_batch)              # device tensor, valid on the current stream
        logits = model(x)
        feeder.done(x)                           # REQUIRED: marks the buffer free once the forward has read it --
                                                 # without it the copy of batch i+2 may overwrite x under the forward

Two device buffers and a dedicated copy stream; the compute stream only ever waits on the event of the copy it is
about to consume, and a buffer is not overwritten before the forward that read it has finished.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


class DeviceFeeder:
    def __init__(self, shape: Tuple[int, ...], dtype: torch.dtype = torch.bfloat16,
                 device: Optional[torch.device] = None) -> None:
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.bufs = [torch.empty(shape, dtype=dtype, device=self.device) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.staged: Optional[int] = None      # index of the buffer holding a batch copied ahead of time
        self.i = 0

    def _start_copy(self, slot: int, host: torch.Tensor) -> None:
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])       # the forward that used this buffer is done
            self.bufs[slot].copy_(host, non_blocking=True)
            self.copied[slot].record(self.copy_stream)

    def push(self, host: torch.Tensor, next_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Return `host` on the device.  If `next_host` is given its copy is started now, so that it overlaps with
        whatever the caller enqueues on the compute stream before the next push()."""
        cur = torch.cuda.current_stream(self.device)
        slot = self.i & 1
        if self.staged != slot:
            self._start_copy(slot, host)
        cur.wait_event(self.copied[slot])
        self.staged = None
        if next_host is not None:
            self._start_copy(slot ^ 1, next_host)
            self.staged = slot ^ 1
        self.i += 1
        return self.bufs[slot]

    def done(self, slot_tensor: torch.Tensor) -> None:
        """Mark the buffer returned by the last push() as consumed by everything enqueued so far."""
        slot = 0 if slot_tensor.data_ptr() == self.bufs[0].data_ptr() else 1
        self.consumed[slot].record(torch.cuda.current_stream(self.device))
