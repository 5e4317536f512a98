"""Host-side staging of one step's inputs (the loader -> device hand-over of trainer.py:233-237).

The reference moves every entry of the batch dictionary with its own ``.to(device)``: ~25 small
cudaMemcpyAsync calls per step, each with its Python and driver overhead.  ``PinnedBatch`` packs the
dictionary into ONE page-locked arena when the batch is collated; a step then crosses PCIe as a single
asynchronous copy and the tensors are views of the device-side arena.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

_ALIGN = 256


class PinnedBatch:
    """``PinnedBatch(dict_of_cpu_tensors)``; ``.upload(device) -> (dict_of_device_views, device_arena)``.

    The copy is enqueued on the current stream of ``device`` with ``non_blocking=True``; the caller orders
    other streams after it with an event and keeps ``device_arena`` alive (``record_stream``) like any tensor
    handed from a copy stream to a compute stream."""

    def __init__(self, tensors: Dict):
        self.layout = []
        off = 0
        for k, t in tensors.items():
            if not isinstance(t, torch.Tensor):
                raise TypeError("PinnedBatch holds tensors only (key %r)" % (k,))
            t = t.detach().contiguous()
            n = t.numel() * t.element_size()
            self.layout.append((k, off, n, t.dtype, tuple(t.shape)))
            off = (off + n + _ALIGN - 1) // _ALIGN * _ALIGN
        self.nbytes = sum(n for _, _, n, _, _ in self.layout)
        self.arena = torch.empty(max(off, 1), dtype=torch.uint8)
        if torch.cuda.is_available():
            self.arena = self.arena.pin_memory()
        for (k, o, n, dt, shape), t in zip(self.layout, tensors.values()):
            self.arena[o:o + n].view(dt).view(shape).copy_(t.detach())

    def upload(self, device) -> Tuple[Dict, torch.Tensor]:
        dev_arena = self.arena.to(device, non_blocking=True)
        return {k: dev_arena[o:o + n].view(dt).view(shape) for (k, o, n, dt, shape) in self.layout}, dev_arena

    def upload_into(self, dev_arena: torch.Tensor) -> None:
        """Refill a device arena obtained from an earlier :meth:`upload` in place (static slots of a CUDA graph)."""
        dev_arena.copy_(self.arena, non_blocking=True)
