"""Batch pipeline: several image-pair batches in flight on one GPU.

The forward has two unavoidable host interludes per batch (SURVEY.md hard part 4): the OpenCV RANSAC of
``geo_module.py:48`` and the read-back of match counts that size the fine stage.  Running ``depth``
batches concurrently — each on its own CUDA stream, driven by its own host thread — lets the GPU work
of one batch fill the host gaps of another, which is the reference's serial per-pair loop
(``hpatches_helper.py:160-177``) turned into a throughput pipeline.  Results keep the input order.
"""
from __future__ import annotations

import gc
import threading
from typing import Callable, Dict, Iterable, List, Optional

import torch


class MatchPipeline:
    def __init__(self, model, depth: int = 2, device: Optional[torch.device] = None, freeze_gc: bool = False,
                 prepare: Optional[Callable] = None):
        """freeze_gc: move the objects alive at the first run() to the permanent GC generation (gc.freeze()) — a
        process-wide side effect, hence opt-in; bench.py uses it (see run()).
        prepare(batch) -> data dict: optional hook run by the worker ON THE BATCH'S STREAM before the forward (e.g. the
        GPU image ingest of geoformer_b200.hpatches: uint8 upload + resize land on the stream that consumes them)."""
        self.model = model
        self.prepare = prepare
        self.freeze_gc = freeze_gc
        self.depth = max(1, int(depth))
        self.device = device or next(model.parameters()).device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.depth)]
        self._frozen = False

    def _job(self, slot: int, data: Dict[str, torch.Tensor], post: Optional[Callable]):
        s = self.streams[slot]
        with torch.cuda.stream(s):
            if self.prepare is not None:
                data = self.prepare(data)
            data = {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) and not v.is_cuda else v)
                    for k, v in data.items()}
            out = self.model(data)
            res = post(out) if post is not None else out
        s.synchronize()
        return res

    def run(self, batches: Iterable[Dict[str, torch.Tensor]], post: Optional[Callable] = None) -> List:
        """``batches``: dicts with 'image0'/'image1' (CUDA tensors, or pinned host tensors which are uploaded
        inside the pipeline).  ``post(data)`` runs on the batch's stream (e.g. device->host of the results).
        ``depth`` worker threads, each owning one stream, pull batches from the shared iterator, so ``depth``
        batches stay in flight until the input is exhausted; results keep the input order."""
        return list(self.run_iter(batches, post))

    def run_iter(self, batches: Iterable[Dict[str, torch.Tensor]], post: Optional[Callable] = None):
        """Generator form of ``run``: yields each batch's result IN INPUT ORDER as soon as it (and all earlier ones) is
        done, while the workers keep going.  The consumer runs on the calling thread - that is where per-batch
        collectives belong (every rank then issues them in the same, deterministic order; bench.py gathers the match
        lists there)."""
        main = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(main)
        with torch.cuda.device(self.device):
            self.model._weights(self.device)          # pack once, before the workers start (they would race to do it)
        if self.freeze_gc and not self._frozen:
            # a full (generation-2) collection walks every tracked object of the process (model, torch, cv2 ...) while
            # holding the GIL: measured as 200+ ms stalls of ALL launch threads.  Moving what is alive now to the
            # permanent generation keeps later collections proportional to the garbage the pipeline itself creates.
            gc.collect()
            gc.freeze()
            self._frozen = True
        it = enumerate(batches)
        lock = threading.Lock()
        done = threading.Condition()
        results: Dict[int, object] = {}
        errors: List[BaseException] = []
        state = {"issued": 0, "exhausted": False}

        def worker(slot: int):
            torch.cuda.set_device(self.device)
            while not errors:
                with lock:
                    nxt = next(it, None)
                    if nxt is None:
                        state["exhausted"] = True
                    else:
                        state["issued"] += 1
                if nxt is None:
                    break
                idx, batch = nxt
                try:
                    res = self._job(slot, batch, post)
                except BaseException as e:      # surfaced to the caller below (the helpers count match failures)
                    errors.append(e)
                    break
                with done:
                    results[idx] = res
                    done.notify_all()
            with done:
                done.notify_all()

        threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(self.depth)]
        for t in threads:
            t.start()
        nxt_idx = 0
        try:
            while True:
                with done:
                    while nxt_idx not in results and not errors and \
                            not (state["exhausted"] and nxt_idx >= state["issued"] and not any(t.is_alive() for t in threads)):
                        done.wait(timeout=0.05)
                    if errors:
                        break
                    if nxt_idx in results:
                        res = results.pop(nxt_idx)
                    else:
                        break
                yield res
                nxt_idx += 1
        finally:
            for t in threads:
                t.join()
            for s in self.streams:
                main.wait_stream(s)
        if errors:
            raise errors[0]
        while nxt_idx in results:               # results that completed while the last wait timed out
            yield results.pop(nxt_idx)
            nxt_idx += 1

    def close(self):
        pass
