"""Host logic of the stream batcher with a fake model — no GPU; plus one GPU test."""
import threading
import time

import numpy as np
import pytest

from infercam_onnx_b200.batcher import StreamBatcher


class FakeModel:
    """run_batch returns one detection per frame whose confidence encodes the frame's value."""

    def __init__(self, delay=0.0):
        self.batches = []
        self.delay = delay
        self.lock = threading.Lock()

    def run_batch(self, frames, cap):
        with self.lock:
            self.batches.append(len(frames))
        time.sleep(self.delay * (1 + (len(self.batches) % 2)))  # uneven latency: batches finish out of order
        dets = [np.float32([[0, 0, 1, 1, float(f)]]) for f in frames]
        return dets, [1] * len(frames)


def test_batches_respect_max_batch_and_keep_per_stream_order():
    model = FakeModel(delay=0.002)
    got = {}
    lock = threading.Lock()

    def on_result(stream, dets):
        with lock:
            got.setdefault(stream, []).append(dets[0][1])
    b = StreamBatcher(model, max_batch=16, max_delay=0.001, capacity=10_000, workers=3)
    for i in range(400):
        assert b.try_submit(i % 8, i, on_result)
    b.close()
    assert max(model.batches) <= 16 and sum(model.batches) == 400 and b.frames == 400
    for s in range(8):
        assert got[s] == [float(i) for i in range(400) if i % 8 == s]  # submission order per stream


def test_deadline_flushes_a_partial_batch():
    model = FakeModel()
    done = threading.Event()
    b = StreamBatcher(model, max_batch=256, max_delay=0.01, workers=1)
    t0 = time.monotonic()
    b.try_submit("cam", 7, lambda s, d: done.set())
    assert done.wait(2.0) and time.monotonic() - t0 < 1.0  # a lone frame is not held back for a full batch
    b.close()
    assert model.batches == [1]


def test_lossy_when_full_like_the_reference_channel():
    """router.rs:64-72: try_send_ref drops the frame when INFER_IMAGES_CHANNEL (capacity 10, lib.rs:37) is full."""
    gate = threading.Event()

    class Blocked(FakeModel):
        def run_batch(self, frames, cap):
            gate.wait(5.0)
            return super().run_batch(frames, cap)
    model = Blocked()
    b = StreamBatcher(model, max_batch=4, max_delay=0.0, capacity=10, workers=1)
    accepted = sum(b.try_submit("cam", i, lambda s, d: None) for i in range(100))
    assert accepted < 100 and b.dropped == 100 - accepted and accepted <= 10 + 4
    gate.set()
    b.close()
    assert b.frames == accepted
    with pytest.raises(RuntimeError):
        b.try_submit("cam", 0, lambda s, d: None)


def test_failed_batch_skips_its_frames():
    class Failing(FakeModel):
        def run_batch(self, frames, cap):
            if any(f < 0 for f in frames):
                raise RuntimeError("bad frame")
            return super().run_batch(frames, cap)
    seen = []
    b = StreamBatcher(Failing(), max_batch=1, max_delay=0.0, workers=1)
    for v in (1, -1, 2):
        b.try_submit("cam", v, lambda s, d: seen.append(d[0][1]))
    b.close()
    assert seen == [1.0, 2.0]  # inferer.rs:37 `if let Ok(..)`: the failed frame is skipped, the loop goes on


@pytest.mark.gpu
def test_batcher_with_the_real_model(make_onnx):
    from infercam_onnx_b200 import nn
    path = make_onnx(320, 240, cls_bias=-0.75)
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=32, lanes=2)
    try:
        rng = np.random.default_rng(5)
        frames = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(6)]
        expect = [m.run(f, cap=256) for f in frames]
        got = {}
        b = StreamBatcher(m, max_batch=32, max_delay=0.002, workers=2)
        for i in range(120):
            assert b.try_submit(i % 6, frames[i % 6], lambda s, d: got.setdefault(s, []).append(d))
        b.close()
        for s in range(6):
            assert len(got[s]) == 20 and all(d == expect[s] for d in got[s])
    finally:
        m.close()
