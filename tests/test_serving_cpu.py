"""The dynamic batcher: host logic only, with a stub batch function - grouping, ordering, delay,
error propagation, shutdown."""
import threading
import time

import pytest

from mmdx_b200.serving import BatchingQueue


def test_concurrent_callers_are_batched_and_get_their_own_result():
    seen = []

    def run(images, details):
        seen.append(len(images))
        time.sleep(0.01)
        return [{"img": i, "txt": d} for i, d in zip(images, details)]

    with BatchingQueue(run, max_batch=16, max_delay_ms=20) as q:
        out = {}

        def worker(k):
            out[k] = q.infer(k, f"details {k}")

        ths = [threading.Thread(target=worker, args=(k,)) for k in range(40)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        assert all(out[k] == {"img": k, "txt": f"details {k}"} for k in range(40))
        assert max(seen) <= 16 and sum(seen) == 40 and len(seen) < 40       # requests were grouped
        assert q.studies == 40 and q.largest <= 16


def test_a_lone_request_waits_at_most_the_delay():
    with BatchingQueue(lambda im, de: [len(d) for d in de], max_batch=64, max_delay_ms=30) as q:
        t = time.monotonic()
        assert q.infer(None, "abc") == 3
        dt = time.monotonic() - t
        assert 0.02 < dt < 0.5


def test_full_batch_does_not_wait_for_the_delay():
    with BatchingQueue(lambda im, de: list(im), max_batch=4, max_delay_ms=5000) as q:
        t = time.monotonic()
        fs = [q.submit(i, "") for i in range(4)]
        assert [f.result(2) for f in fs] == [0, 1, 2, 3]
        assert time.monotonic() - t < 1.0


def test_errors_reach_every_caller_and_the_queue_survives():
    calls = []

    def run(images, details):
        calls.append(len(images))
        if any(i == "bad" for i in images):
            raise ValueError("boom")
        return list(images)

    with BatchingQueue(run, max_batch=8, max_delay_ms=20) as q:
        fs = [q.submit(x, "") for x in ("a", "bad", "c")]
        for f in fs:
            with pytest.raises(ValueError, match="boom"):
                f.result(2)
        assert q.infer("ok", "") == "ok"
    with pytest.raises(RuntimeError):
        q.submit("late", "")


def test_close_flushes_pending_requests():
    q = BatchingQueue(lambda im, de: list(im), max_batch=100, max_delay_ms=10000)
    fs = [q.submit(i, "") for i in range(5)]
    q.close()
    assert [f.result(1) for f in fs] == list(range(5))
