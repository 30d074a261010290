"""Input staging for the photometric engine: a ring of device-resident batch slots refilled from pinned host memory on a copy
stream, one step ahead of the compute stream.

Replaces, for the B200 engine, what the reference does per iteration with a pinned-memory DataLoader and `.to(device)` of every
tensor of the sample (vhap/model/tracker.py:1352-1357, vhap/data/video_dataset.py:209-264).  The slots have FIXED device addresses, so
the CUDA graphs captured per slot (Engine.graph_begin) stay valid while their contents change.  Targets travel as the dataset decodes
them -- uint8 RGB, 3 bytes per pixel -- and are normalised inside the kernels (vhap_frame_batch::target_format = 1); fp16 RGBA
(8 bytes per pixel) is accepted as well."""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch


def pin_sample(rgb, lmk2d, timesteps, RT=None, K=None) -> Dict[str, torch.Tensor]:
    """host side of one batch in pinned memory.  rgb: [B,H,W,3] uint8 (numpy / tensor) or [B,H,W,4] fp16."""
    t = torch.as_tensor(rgb)
    if t.dtype not in (torch.uint8, torch.float16):
        raise ValueError("pin_sample: rgb must be uint8 [B,H,W,3] or fp16 [B,H,W,4]")
    out = dict(rgb=t.contiguous().pin_memory(), lmk2d=torch.as_tensor(np.asarray(lmk2d), dtype=torch.float32)[:, :68].contiguous().pin_memory(),
               ts=torch.as_tensor(np.asarray(timesteps), dtype=torch.int32).pin_memory(), RT=RT, K=K)
    # view sharing (Engine.stage_sample): distinct timesteps of the batch + the inverse map, staged with the sample
    uniq, inv = np.unique(np.asarray(timesteps).reshape(-1), return_inverse=True)
    out["geo_ts"] = torch.as_tensor(uniq.astype(np.int32)).pin_memory()
    out["geo"] = torch.as_tensor(inv.astype(np.int32)).pin_memory()
    return out


class InputRing:
    """`depth` device slots; slot j is refilled (prefetch) only after the step that used it has been released."""

    def __init__(self, engine, first_samples: List[dict]):
        self.e = engine
        self.slots = [engine.stage_sample(s["rgb"], s["lmk2d"], s["ts"].numpy(), RT=s.get("RT"), K=s.get("K")) for s in first_samples]
        self.depth = len(self.slots)
        self.copy_stream = torch.cuda.Stream(engine.dev)
        self.ready = [torch.cuda.Event() for _ in range(self.depth)]
        self.freed: List[Optional[torch.cuda.Event]] = [None] * self.depth
        torch.cuda.current_stream(engine.dev).synchronize()
        for ev in self.ready:                                # the initial contents are valid
            ev.record(self.copy_stream)

    @property
    def batches(self):
        return self.slots

    def bytes_per_step(self) -> int:
        b = self.slots[0]
        return b.target.numel() * b.target.element_size() + b.lmk2d.numel() * 4 + b.timesteps.numel() * 4

    def prefetch(self, j: int, sample: dict) -> None:
        """enqueue the host -> device copies of `sample` into slot j on the copy stream (non-blocking for the host)"""
        bt = self.slots[j]
        if sample["rgb"].dtype != bt.target.dtype or tuple(sample["rgb"].shape) != tuple(bt.target.shape):
            raise ValueError("InputRing.prefetch: sample does not match the slot's shape / target format")
        with torch.cuda.stream(self.copy_stream):
            if self.freed[j] is not None:
                self.copy_stream.wait_event(self.freed[j])
            bt.target.copy_(sample["rgb"], non_blocking=True)
            bt.lmk2d.copy_(sample["lmk2d"], non_blocking=True)
            bt.timesteps.copy_(sample["ts"], non_blocking=True)
            if bt.geo is not None:                       # view sharing: the slot's geometry map follows the new timesteps
                if sample["geo_ts"].numel() != bt.geo_ts.numel():
                    raise ValueError("InputRing.prefetch: the number of distinct timesteps differs from the slot's")
                bt.geo.copy_(sample["geo"], non_blocking=True)
                bt.geo_ts.copy_(sample["geo_ts"], non_blocking=True)
            self.ready[j].record(self.copy_stream)

    def acquire(self, j: int):
        """make the current stream wait for slot j's copies; returns the staged Batch"""
        torch.cuda.current_stream(self.e.dev).wait_event(self.ready[j])
        return self.slots[j]

    def release(self, j: int) -> None:
        """the step that used slot j has been enqueued on the current stream: the slot may be refilled after it"""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.e.dev))
        self.freed[j] = ev
