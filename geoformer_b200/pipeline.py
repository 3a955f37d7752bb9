"""Batch pipeline: several image-pair batches in flight on one GPU.

The forward has two unavoidable host interludes per batch (SURVEY.md hard part 4): the OpenCV RANSAC of
``geo_module.py:48`` and the read-back of match counts that size the fine stage.  Running ``depth``
batches concurrently — each on its own CUDA stream, driven by its own host thread — lets the GPU work
of one batch fill the host gaps of another, which is the reference's serial per-pair loop
(``hpatches_helper.py:160-177``) turned into a throughput pipeline.  Results keep the input order.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, Iterable, List, Optional

import torch


class MatchPipeline:
    def __init__(self, model, depth: int = 2, device: Optional[torch.device] = None):
        self.model = model
        self.depth = max(1, int(depth))
        self.device = device or next(model.parameters()).device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.depth)]
        self.pool = ThreadPoolExecutor(max_workers=self.depth)

    def _job(self, slot: int, data: Dict[str, torch.Tensor], post: Optional[Callable]):
        torch.cuda.set_device(self.device)
        s = self.streams[slot]
        with torch.cuda.stream(s):
            data = {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) and not v.is_cuda else v)
                    for k, v in data.items()}
            out = self.model(data)
            res = post(out) if post is not None else out
        s.synchronize()
        return res

    def run(self, batches: Iterable[Dict[str, torch.Tensor]], post: Optional[Callable] = None) -> List:
        """``batches``: dicts with 'image0'/'image1' (CUDA tensors, or pinned host tensors which are uploaded
        inside the pipeline).  ``post(data)`` runs on the batch's stream (e.g. device->host of the results)."""
        main = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(main)
        futs = [self.pool.submit(self._job, i % self.depth, b, post) for i, b in enumerate(batches)]
        out = [f.result() for f in futs]
        for s in self.streams:
            main.wait_stream(s)
        return out

    def close(self):
        self.pool.shutdown(wait=True)
