"""CPU tests of the oracle itself (no GPU): the C restatement against known answers, against an
independent pure-Python restatement of the same reference lines, and against the committed pins."""
import hashlib
import math

import numpy as np
import pytest

from oracle import hotpath


# ---- resize tap tables: values probed in SURVEY.md §8a-R ------------------------------------
def test_taps_640_to_320_known_values():
    left, nt, w = hotpath.axis_taps(640, 320)
    assert list(left[:3]) == [0, 1, 3] and list(nt[:3]) == [3, 4, 4]
    np.testing.assert_array_equal(w[1], np.float32([0.125, 0.375, 0.375, 0.125]))
    np.testing.assert_array_equal(w[0][:3], np.float32([3 / 7, 3 / 7, 1 / 7]))
    np.testing.assert_array_equal(w[319][:3], np.float32([1 / 7, 3 / 7, 3 / 7]))


def test_taps_720_to_240_and_upscale_known_values():
    _, nt, w = hotpath.axis_taps(720, 240)
    assert nt[5] == 7
    np.testing.assert_array_equal(w[5], np.float32([0, 0.11111111, 0.22222222, 0.33333337, 0.22222222, 0.11111111, 0]))
    _, _, w = hotpath.axis_taps(427, 480)
    np.testing.assert_array_equal(w[1], np.float32([0.16562498, 0.834375, 0]))
    # same length on an axis => taps [0,1,0]: exact pass-through
    left, nt, w = hotpath.axis_taps(50, 50)
    for o in range(1, 49):
        assert nt[o] == 3 and left[o] == o - 1
        np.testing.assert_array_equal(w[o], np.float32([0, 1, 0]))


def _py_resize(src, nw, nh):
    """Independent numpy-f32 restatement of image 0.24.5 resize (vertical then horizontal)."""
    f = np.float32
    h, w, _ = src.shape
    if (nw, nh) == (w, h):
        return src.copy()

    def taps(S, D):
        ratio = f(S) / f(D)
        sratio = ratio if ratio >= 1 else f(1)
        out = []
        for o in range(D):
            c = (f(o) + f(0.5)) * ratio
            left = min(max(int(math.floor(c - sratio)), 0), S - 1)
            right = min(max(int(math.ceil(c + sratio)), left + 1), S)
            c = c - f(0.5)
            ws = []
            for i in range(left, right):
                x = abs((f(i) - c) / sratio)
                ws.append(f(1) - x if x < 1 else f(0))
            s = f(0)
            for v in ws:
                s = f(s + v)
            out.append((left, [f(v / s) for v in ws]))
        return out

    tmp = np.zeros((nh, w, 3), np.float32)
    for oy, (l, ws) in enumerate(taps(h, nh)):
        acc = np.zeros((w, 3), np.float32)
        for i, wt in enumerate(ws):
            acc = (acc + (src[l + i].astype(np.float32) * wt).astype(np.float32)).astype(np.float32)
        tmp[oy] = acc
    out = np.zeros((nh, nw, 3), np.uint8)
    for ox, (l, ws) in enumerate(taps(w, nw)):
        acc = np.zeros((nh, 3), np.float32)
        for i, wt in enumerate(ws):
            acc = (acc + (tmp[:, l + i] * wt).astype(np.float32)).astype(np.float32)
        acc = np.clip(acc, 0, 255)
        out[:, ox] = np.where(acc - np.floor(acc) >= 0.5, np.floor(acc) + 1, np.floor(acc)).astype(np.uint8)
    return out


@pytest.mark.parametrize("shape,target", [((48, 64), (32, 24)), ((30, 40), (40, 30)), ((61, 37), (20, 24)),
                                           ((24, 32), (32, 24)), ((427 // 4, 160), (80, 60))])
def test_c_resize_matches_independent_restatement(shape, target):
    rng = np.random.default_rng(sum(shape))
    src = rng.integers(0, 256, (*shape, 3), dtype=np.uint8)
    np.testing.assert_array_equal(hotpath.resize_triangle(src, *target), _py_resize(src, *target))


def test_resize_properties():
    rng = np.random.default_rng(1)
    src = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    # identity size => copy (sample.rs early return)
    np.testing.assert_array_equal(hotpath.resize_triangle(src, 640, 480), src)
    # constant image stays constant (weights are normalised)
    const = np.full((90, 120, 3), 201, np.uint8)
    assert (hotpath.resize_triangle(const, 40, 30) == 201).all()
    # exact dyadic case: 2x downscale interior = ([1,3,3,1]/8) x ([1,3,3,1]/8)
    r = hotpath.resize_triangle(src, 320, 240)
    k = np.float32([1, 3, 3, 1]) / np.float32(8)
    patch = src[2 * 50 - 1:2 * 50 + 3, 2 * 70 - 1:2 * 70 + 3, 1].astype(np.float32)
    v = float((k[:, None] * patch).sum(0) @ k)
    assert r[50, 70, 1] == int(math.floor(v + 0.5))


def test_normalise_matches_formula():
    hwc = np.arange(256 * 3, dtype=np.uint8).reshape(16, 16, 3)
    out = hotpath.normalise_nchw(hwc, 0)
    mean = np.float32([0.485, 0.456, 0.406])
    std = np.float32([0.229, 0.224, 0.225])
    exp = ((hwc.astype(np.float32) / np.float32(255) - mean) / std).transpose(2, 0, 1)
    np.testing.assert_array_equal(out, exp.astype(np.float32))
    np.testing.assert_array_equal(hotpath.normalise_nchw(hwc, 1),
                                  ((hwc.astype(np.float32) - np.float32(127)) / np.float32(128)).transpose(2, 0, 1))


# ---- IoU / area: hand-computed (nn.rs:227-260) ----------------------------------------------
def test_bbox_area_and_iou_known_answers():
    assert hotpath.bbox_area([0.1, 0.2, 0.5, 0.6]) == pytest.approx(0.16, rel=1e-6)
    assert hotpath.bbox_area([0.5, 0.2, 0.1, 0.6]) == 0.0  # ill-defined box
    a, b = [0.0, 0.0, 2.0, 2.0], [1.0, 1.0, 3.0, 3.0]
    f = np.float32
    assert hotpath.iou(a, b) == float(f(1) / f(f(f(4) + f(4)) - f(1) + f(1e-7)))
    assert hotpath.iou(a, [5, 5, 6, 6]) == 0.0
    assert hotpath.iou(a, a) == float(f(4) / f(f(f(8) - f(4)) + f(1e-7)))
    assert hotpath.iou(a, b) == hotpath.iou(b, a)


def _py_postproc(scores, boxes, min_conf, max_iou):
    """Line-by-line Python restatement of nn.rs:109-140 + 198-224 (uses list.sort: stable)."""
    cands = [(boxes[k], scores[k, 1], k) for k in range(len(scores)) if scores[k, 1] > min_conf]
    cands.sort(key=lambda t: t[1])
    sel = []
    while cands:
        bb, c, k = cands.pop()
        if any(hotpath.iou(bb, s[0]) > max_iou for s in sel):
            continue
        sel.append((bb, c, k))
    return sel


def test_postproc_tie_break_and_strict_thresholds():
    # three identical scores: processing order must be HIGHER index first (stable asc sort + pop)
    boxes = np.float32([[0, 0, 1, 1], [0, 0, 1, 1.0001], [5, 5, 6, 6], [0, 0, 1, 1]])
    scores = np.float32([[0.3, 0.7], [0.3, 0.7], [0.5, 0.5], [0.3, 0.7]])
    dets, idx = hotpath.postproc(scores, boxes, 0.5, 0.5)
    assert list(idx) == [3]  # 0.5 is not > 0.5; 3 wins the tie and suppresses 1 and 0
    # iou exactly == max_iou does not suppress (strict >)
    b2 = np.float32([[0, 0, 2, 1], [0, 0, 1, 1]])
    s2 = np.float32([[0.1, 0.9], [0.2, 0.8]])
    v = hotpath.iou(b2[0], b2[1])
    _, idx = hotpath.postproc(s2, b2, 0.5, v)
    assert list(idx) == [0, 1]
    _, idx = hotpath.postproc(s2, b2, 0.5, np.nextafter(np.float32(v), np.float32(0)))
    assert list(idx) == [0]
    # NaN confidence is dropped
    s3 = np.float32([[0, np.nan], [0, 0.9]])
    _, idx = hotpath.postproc(s3, b2, 0.5, 0.5)
    assert list(idx) == [1]
    # empty
    dets, idx = hotpath.postproc(np.zeros((0, 2), np.float32), np.zeros((0, 4), np.float32), 0.5, 0.5)
    assert len(dets) == 0


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_postproc_matches_python_restatement(seed):
    rng = np.random.default_rng(seed)
    K = 300
    c = rng.random((K, 2)).astype(np.float32) * 0.6 + 0.2
    wh = rng.random((K, 2)).astype(np.float32) * 0.3
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    scores = rng.random((K, 2)).astype(np.float32)
    scores[rng.integers(0, K, 40), 1] = np.float32(0.75)  # plenty of exact ties
    dets, idx = hotpath.postproc(scores, boxes, 0.5, 0.4)
    ref = _py_postproc(scores, boxes, 0.5, 0.4)
    assert list(idx) == [r[2] for r in ref]
    np.testing.assert_array_equal(dets[:, 4], np.float32([r[1] for r in ref]))
    assert (np.diff(dets[:, 4]) <= 0).all()  # descending confidence (nn.rs:107-108)


# ---- committed pins (tests/golden/make_golden.py) -------------------------------------------
def test_oracle_reproduces_committed_pins(test_pics, oracle_pins):
    for k, im in test_pics.items():
        for (w, h) in ((320, 240), (640, 480)):
            r = hotpath.resize_triangle(im, w, h)
            assert hashlib.sha256(r.tobytes()).digest() == oracle_pins[f"{k}/resize{w}x{h}/sha256"].tobytes()
            np.testing.assert_array_equal(r[7], oracle_pins[f"{k}/resize{w}x{h}/row7"])


def test_oracle_cnn_reproduces_committed_pins(test_pics, oracle_pins, make_onnx):
    from oracle.ultraface_ref import UltrafaceOracle
    m = UltrafaceOracle(make_onnx(320, 240, seed=0), 320, 240, 0.5, 0.5)
    k = sorted(test_pics)[0]
    s, b = m.raw([test_pics[k]])
    np.testing.assert_allclose(s[0, :32], oracle_pins[f"{k}/scores_head"], atol=2e-6)
    np.testing.assert_allclose(b[0, :32], oracle_pins[f"{k}/boxes_head"], atol=2e-6)
    dets = m.postproc(s[0], b[0])
    assert len(dets) == int(oracle_pins[f"{k}/det_count"][0])


def test_oracle_graph_shapes(make_onnx):
    """K = 4420 / 17640 priors (README.md:121-125) and fp64 vs fp32 error budget of the oracle."""
    import torch
    from oracle.ultraface_ref import UltrafaceOracle
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    m32 = UltrafaceOracle(make_onnx(320, 240), 320, 240)
    m64 = UltrafaceOracle(make_onnx(320, 240), 320, 240, dtype=torch.float64)
    s32, b32 = m32.raw([img])
    s64, b64 = m64.raw([img])
    assert s32.shape == (1, 4420, 2) and b32.shape == (1, 4420, 4)
    np.testing.assert_allclose(s32.sum(-1), 1.0, atol=1e-6)
    assert np.abs(s32 - s64).max() < 2e-5 and np.abs(b32 - b64).max() < 2e-5
    m = UltrafaceOracle(make_onnx(640, 480, variant="slim"), 640, 480)
    s, b = m.raw([img])
    assert s.shape == (1, 17640, 2)


@pytest.mark.parametrize("cfg", [dict(width=320, height=240), dict(width=320, height=240, variant="slim"),
                                 dict(width=320, height=240, with_bn=True, seed=3), dict(width=640, height=480)])
def test_oracle_cnn_agrees_with_opencv_dnn(make_onnx, cfg):
    """Second opinion from an independent, third-party ONNX runtime (OpenCV's dnn module): it accepts
    the fixture file as a regular ONNX model and its outputs agree with the oracle's interpreter. tract itself
    cannot run here; this pins the operator semantics (Conv/BN/Relu/Concat/Transpose/Reshape/Softmax/Slice/...)."""
    cv2 = pytest.importorskip("cv2")
    from oracle.ultraface_ref import UltrafaceOracle
    w, h = cfg["width"], cfg["height"]
    path = make_onnx(w, h, variant=cfg.get("variant", "RFB"), with_bn=cfg.get("with_bn", False), seed=cfg.get("seed", 0))
    net = cv2.dnn.readNetFromONNX(path)
    m = UltrafaceOracle(path, w, h)
    x = m.preproc(np.random.default_rng(0).integers(0, 256, (480, 640, 3), dtype=np.uint8))
    net.setInput(x)
    names = net.getUnconnectedOutLayersNames()
    outs = dict(zip(names, net.forward(names)))
    s, b = m.raw_from_tensor(x)
    assert np.abs(outs["scores"].reshape(s.shape) - s).max() < 2e-5
    assert np.abs(outs["boxes"].reshape(b.shape) - b).max() < 2e-5


def test_oracle_runs_upstream_shaped_export_identically(make_onnx):
    """The interpreter handles the un-simplified export shape (Shape/Gather/Unsqueeze/Constant, attribute-form Slice,
    explicit BN) and gives exactly what the simplified file with the same weights gives."""
    from oracle.ultraface_ref import UltrafaceOracle
    frame = np.random.default_rng(3).integers(0, 256, (427, 640, 3), dtype=np.uint8)
    a = UltrafaceOracle(make_onnx(320, 240, seed=4, cls_bias=-0.75), 320, 240).raw([frame])
    b = UltrafaceOracle(make_onnx(320, 240, seed=4, cls_bias=-0.75, style="upstream"), 320, 240).raw([frame])
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


def test_reference_face_counts_through_the_oracle(real_rfb640_path, test_pics, reference_face_counts):
    """integration_tests.rs:20-34 through the CPU restatement: this is the test that PINS the oracle the moment the weight
    file the reference downloads is available (skipped until then: "parity unpinned")."""
    from oracle.ultraface_ref import UltrafaceOracle
    o = UltrafaceOracle(real_rfb640_path, 640, 480, 0.5, 0.5)
    assert {k: len(o.run(im)) for k, im in test_pics.items()} == reference_face_counts


def test_reference_photos_have_the_reference_geometry(test_pics, reference_face_counts):
    assert set(test_pics) == set(reference_face_counts)
    assert sorted(im.shape[:2] for im in test_pics.values()) == sorted([(427, 640)] * 5 + [(462, 640), (676, 640), (960, 640)])
