"""CPU: the forward engine's host logic (persia_b200/engine.py) — ordering, buffering, bounded staleness, time-outs —
with a stand-in lookup (rust/persia-core/src/forward.rs:396-468, 470-780; backward.rs:286-343)."""
import queue
import threading
import time

import pytest

from persia_b200.engine import ForwardEngine


class B:
    def __init__(self, bid, ref=False):
        self._batch_id, self.ref = bid, ref


def _engine(lookup, **kw):
    q = queue.Queue()
    e = ForwardEngine(lookup, kw.pop("buffer", 4), kw.pop("reproducible", False), kw.pop("staleness", None), **kw)
    e.set_input(q)
    e.is_remote_ref = lambda b: b.ref
    return e, q


def test_reproducible_order_with_shuffled_arrivals_and_rank_stride():
    e, q = _engine(lambda b, p: b._batch_id, reproducible=True, world_size=2, rank=1, buffer=64)
    e.launch(3)
    try:
        for bid in [5, 1, 9, 3, 7, 11]:  # this replica sees ids 1, 3, 5, ... (rank 1 of 2)
            q.put(B(bid))
        got = [e.get_batch(2000) for _ in range(6)]
        # one worker thread may overtake another after the reorder stage only if lookups race: they do not here,
        # because the ordered channel holds one batch at a time and lookups are instantaneous
        assert sorted(got) == [1, 3, 5, 7, 9, 11] and got[0] == 1
    finally:
        e.shutdown()


def test_reproducible_single_worker_is_strictly_ordered_and_flushes_when_input_stalls():
    e, q = _engine(lambda b, p: b._batch_id, reproducible=True, buffer=64, flush_after_s=0.2)
    e.launch(1)
    try:
        for bid in [2, 0, 1, 3]:
            q.put(B(bid))
        assert [e.get_batch(2000) for _ in range(4)] == [0, 1, 2, 3]
        q.put(B(6))  # 4 and 5 never arrive: after flush_after_s the buffered batch is released anyway
        q.put(B(8))
        t0 = time.monotonic()
        assert [e.get_batch(3000), e.get_batch(3000)] == [6, 8]
        assert time.monotonic() - t0 >= 0.15
        q.put(B(9))  # the expectation moved past the gap
        assert e.get_batch(2000) == 9
    finally:
        e.shutdown()


def test_buffer_bounds_the_prefetch_depth():
    done = []
    e, q = _engine(lambda b, p: done.append(b._batch_id) or b._batch_id, buffer=3)
    e.launch(2)
    try:
        for bid in range(20):
            q.put(B(bid))
        time.sleep(0.3)
        assert len(done) <= 3 + 2  # a full buffer plus one finished lookup per worker waiting to be queued
        assert sorted(e.get_batch(1000) for _ in range(20)) == list(range(20))
    finally:
        e.shutdown()


def test_staleness_permits_bound_batches_between_lookup_and_update():
    inflight, peak, lock = [0], [0], threading.Lock()
    permits = []

    def lookup(b, permit):
        if permit is not None:
            with lock:
                inflight[0] += 1
                peak[0] = max(peak[0], inflight[0])
            permits.append(permit)
        return b._batch_id

    e, q = _engine(lookup, staleness=2, buffer=16)
    e.launch(4)
    try:
        for bid in range(6):
            q.put(B(bid, ref=True))
        assert len({e.get_batch(1000), e.get_batch(1000)}) == 2
        with pytest.raises(TimeoutError):  # two permits out, nobody has applied an update yet
            e.get_batch(200)
        for _ in range(4):  # "backward" returns a permit: exactly one more batch comes through each time
            with lock:
                inflight[0] -= 1
            permits.pop(0).release()
            e.get_batch(1000)
        assert peak[0] == 2
        q.put(B(100, ref=False))  # batches that carry their ids need no permit (forward.rs:676-686)
        permits[0].release()
        permits[0].release()  # idempotent
        assert e.get_batch(1000) == 100
    finally:
        e.shutdown()


def test_errors_timeouts_and_lifecycle():
    def lookup(b, p):
        if b._batch_id == 1:
            raise RuntimeError("slot: nope not found")
        return b._batch_id

    e, q = _engine(lookup)
    with pytest.raises(RuntimeError):
        e.get_batch(10)  # not launched
    e.launch(1)
    e.launch(1)  # second launch is a no-op ("already launch")
    with pytest.raises(RuntimeError):
        e.set_input(queue.Queue())
    try:
        with pytest.raises(TimeoutError):
            e.get_batch(50)
        q.put(B(0))
        q.put(B(1))
        assert e.get_batch(1000) == 0
        with pytest.raises(RuntimeError, match="not found"):
            e.get_batch(1000)
    finally:
        e.shutdown()
    assert not any(t.is_alive() for t in threading.enumerate() if t.name.startswith("persia-forward"))
    e2 = ForwardEngine(lookup, 2, False)
    with pytest.raises(RuntimeError):
        e2.launch(1)  # no input channel
