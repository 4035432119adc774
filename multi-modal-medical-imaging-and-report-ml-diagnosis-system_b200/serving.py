"""Dynamic batching in front of `inference_batch` (second half).

The reference serves one study per HTTP request (`predict_view`, backend/api/views.py:62-103, calls `inference()`
synchronously); at B = 1 the GPU path is launch-latency bound (~0.6 ms per study, 1.6 k studies/s) while a batch of 256
takes 9 ms (28 k studies/s).  `BatchingQueue` sits between concurrent callers and the engine: requests that arrive
within `max_delay_ms` of each other (or until `max_batch` are waiting) are run as ONE `inference_batch` call and every
caller gets exactly the dict `inference()` would have returned.  It is a library-level batcher - the Django view itself
stays out of scope; INTEGRATION.md shows the two-line change in `predict_view`.

The queue is generic over the batch function so that its logic is testable without a GPU."""
from __future__ import annotations

import threading
import time
from concurrent.futures import Future


class BatchingQueue:
    def __init__(self, run_batch, max_batch: int = 256, max_delay_ms: float = 2.0):
        """run_batch(images: list, details: list[str]) -> list of per-study results, in order."""
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        self._run = run_batch
        self.max_batch = int(max_batch)
        self.max_delay = float(max_delay_ms) * 1e-3
        self._cv = threading.Condition()
        self._pending: list = []           # (image, details, future, t_arrival)
        self._closed = False
        self.batches = 0                   # telemetry: batches run / studies served / largest batch
        self.studies = 0
        self.largest = 0
        self._worker = threading.Thread(target=self._loop, name="mmdx-batcher", daemon=True)
        self._worker.start()

    @classmethod
    def for_bundle(cls, model_bundle, device=None, max_batch: int = 256, max_delay_ms: float = 2.0, max_len: int = 96,
                   gen_kwargs=False):
        """The queue in front of this package's `inference_batch` for one bundle / device."""
        from .inference_pipeline import inference_batch

        def run(images, details):
            return inference_batch(model_bundle, images, details, device=device, gen_kwargs=gen_kwargs, max_len=max_len)

        return cls(run, max_batch, max_delay_ms)

    def submit(self, image, patient_details: str) -> Future:
        """Enqueue one study; the Future resolves to the `inference()` result dict (or raises what the batch raised)."""
        f: Future = Future()
        with self._cv:
            if self._closed:
                raise RuntimeError("BatchingQueue is closed")
            self._pending.append((image, patient_details, f, time.monotonic()))
            self._cv.notify_all()
        return f

    def infer(self, image, patient_details: str, timeout: float | None = None):
        """Synchronous form - what `predict_view` calls instead of `inference()`."""
        return self.submit(image, patient_details).result(timeout)

    def close(self):
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._worker.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _take(self):
        """Blocks until a batch is due: `max_batch` studies are waiting, or the oldest has waited `max_delay`."""
        with self._cv:
            while True:
                if self._pending:
                    due = self._pending[0][3] + self.max_delay
                    now = time.monotonic()
                    if len(self._pending) >= self.max_batch or now >= due or self._closed:
                        batch, self._pending = self._pending[:self.max_batch], self._pending[self.max_batch:]
                        return batch
                    self._cv.wait(due - now)
                elif self._closed:
                    return None
                else:
                    self._cv.wait()

    def _loop(self):
        while True:
            batch = self._take()
            if batch is None:
                return
            try:
                res = self._run([b[0] for b in batch], [b[1] for b in batch])
                if len(res) != len(batch):
                    raise RuntimeError(f"run_batch returned {len(res)} results for {len(batch)} studies")
                for (_, _, f, _), r in zip(batch, res):
                    f.set_result(r)
            except BaseException as ex:      # noqa: BLE001 - every waiting caller must hear about it
                for _, _, f, _ in batch:
                    if not f.done():
                        f.set_exception(ex)
            self.batches += 1
            self.studies += len(batch)
            self.largest = max(self.largest, len(batch))
