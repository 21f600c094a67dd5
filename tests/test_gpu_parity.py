"""GPU parity tests (run on the B200 box: pytest -m gpu). Everything goes through the C ABI
(infercam_onnx_b200.nn -> libultraface_b200.so) and is compared with the CPU oracle on the same
seeded inputs. Bars: u8 resize and normalised tensor bit-exact; raw scores/boxes within 1e-4 abs
(fp32 accumulate; BASELINE.json north_star); detection set identical except candidates whose
score / IoU lies within tolerance of a threshold; post-processing on identical raw tensors exact."""
import numpy as np
import pytest

from infercam_onnx_b200 import _capi, nn
from oracle import hotpath
from oracle.ultraface_ref import UltrafaceOracle

pytestmark = pytest.mark.gpu

TOL = 1e-4  # raw tensor tolerance (abs), fp32-accumulate mode


def _noise(n, h=480, w=640, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (n, h, w, 3), dtype=np.uint8)


def _smooth(n, h=480, w=640, seed=0):
    """low-pass noise so the resize is not tested on white noise only"""
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, (n, h // 16 + 2, w // 16 + 2, 3)).astype(np.float32)
    out = np.empty((n, h, w, 3), np.uint8)
    for i in range(n):
        out[i] = hotpath.resize_triangle(small[i].astype(np.uint8), w, h)
    return out


@pytest.fixture(scope="module")
def model320(make_onnx):
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=make_onnx(320, 240), max_batch=8)
    yield m
    m.close()


@pytest.fixture(scope="module")
def oracle320(make_onnx):
    return UltrafaceOracle(make_onnx(320, 240), 320, 240, 0.5, 0.5)


def _assert_detection_sets_match(gpu, ref, scores, boxes, min_conf=0.5, max_iou=0.5, tol=TOL):
    """Cross-implementation detection parity: `gpu` (post-processing of the GPU's raw tensors) against
    `ref` (post-processing of the ORACLE's raw tensors `scores`/`boxes`).

    The sets must be identical (every row within tol) unless a greedy decision was within tolerance:
    a candidate score within tol of min_conf, two overlapping candidates whose scores differ by less
    than tol (processing order may swap), or a candidate pair whose IoU is within 1e-3 of max_iou.
    Even then at most a few detections may differ (a flip can cascade to its neighbours)."""
    def unmatched(a, b):
        if len(a) == 0:
            return 0
        if len(b) == 0:
            return len(a)
        d = np.abs(a[:, None, :] - b[None, :, :]).max(-1)
        return int((d.min(1) > tol).sum())
    miss = unmatched(gpu, ref) + unmatched(ref, gpu)
    if miss == 0:
        return
    cand = np.nonzero(scores[:, 1] > min_conf - tol)[0]
    sc, bx = scores[cand, 1], boxes[cand]
    near_thr = bool((np.abs(sc - min_conf) <= tol).any())
    x0 = np.maximum(bx[:, None, 0], bx[None, :, 0]); y0 = np.maximum(bx[:, None, 1], bx[None, :, 1])
    x1 = np.minimum(bx[:, None, 2], bx[None, :, 2]); y1 = np.minimum(bx[:, None, 3], bx[None, :, 3])
    inter = np.clip(x1 - x0, 0, None) * np.clip(y1 - y0, 0, None)
    area = np.clip(bx[:, 2] - bx[:, 0], 0, None) * np.clip(bx[:, 3] - bx[:, 1], 0, None)
    iou = inter / (area[:, None] + area[None, :] - inter + 1e-7)
    off = ~np.eye(len(cand), dtype=bool)
    near_iou = bool((np.abs(iou - max_iou)[off] <= 1e-3).any())
    near_tie = bool(((np.abs(sc[:, None] - sc[None, :]) <= tol) & (iou > 0) & off).any())
    assert near_thr or near_iou or near_tie, f"{miss} detections differ without any decision near a threshold"
    assert miss <= max(4, 0.05 * (len(gpu) + len(ref))), f"{miss} of {len(gpu)}+{len(ref)} detections differ"


# ---------------------------------------------------------------------------------------------
# R1: resize, bit-exact
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("net", [(320, 240), (640, 480)])
def test_resize_bit_exact_on_reference_test_pics(make_onnx, test_pics, net):
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=make_onnx(*net), size=net)
    try:
        for k, im in test_pics.items():
            np.testing.assert_array_equal(m.preproc_u8(im), hotpath.resize_triangle(im, *net), err_msg=k)
    finally:
        m.close()


@pytest.mark.parametrize("shape", [(480, 640), (720, 1280), (240, 320), (427, 640), (960, 640), (1080, 1920),
                                   (100, 100), (241, 319), (33, 1000), (1000, 33), (7, 9), (479, 641), (3, 3)])
def test_resize_bit_exact_synthetic_and_odd_sizes(model320, shape):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    im = rng.integers(0, 256, (*shape, 3), dtype=np.uint8)
    np.testing.assert_array_equal(model320.preproc_u8(im), hotpath.resize_triangle(im, 320, 240))
    sm = _smooth(1, *shape, seed=1)[0] if min(shape) >= 32 else im
    np.testing.assert_array_equal(model320.preproc_u8(sm), hotpath.resize_triangle(sm, 320, 240))
    extreme = np.where(rng.random((*shape, 3)) < 0.5, 0, 255).astype(np.uint8)  # saturating input
    np.testing.assert_array_equal(model320.preproc_u8(extreme), hotpath.resize_triangle(extreme, 320, 240))


@pytest.mark.parametrize("shape", [(2160, 3840), (16, 4000), (3000, 24), (1200, 1600)])
def test_resize_bit_exact_large_and_extreme_aspect(model320, shape):
    """Maximum sizes: 4K frames (ratio 12 x 9: 20+ taps per axis) and extreme aspect ratios (stretch, nn.rs:74-80)."""
    rng = np.random.default_rng(shape[0] + shape[1])
    im = rng.integers(0, 256, (*shape, 3), dtype=np.uint8)
    np.testing.assert_array_equal(model320.preproc_u8(im), hotpath.resize_triangle(im, 320, 240))


def test_resize_pre024_switch(make_onnx):
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=make_onnx(320, 240),
                              resize_round_intermediate=True, max_batch=8)
    try:
        im = _noise(1, 427, 640, seed=5)[0]
        np.testing.assert_array_equal(m.preproc_u8(im), hotpath.resize_triangle(im, 320, 240, True))
        ims = _noise(8, 480, 640, seed=6)  # exact 2:1 batch: the integer kernel's (v + 4) >> 3 per pass
        got = m.preproc_u8_batch(ims)
        for i in range(len(ims)):
            np.testing.assert_array_equal(got[i], hotpath.resize_triangle(ims[i], 320, 240, True))
    finally:
        m.close()


def test_resize_exact_half_integer_kernel(make_onnx):
    """Exact 2:1 on both axes runs in packed integer arithmetic (interior) + f32 tap tables (border row / column):
    must equal the oracle's f32 restatement bit for bit, including saturated and constant images, on both nets."""
    for net, shape in (((320, 240), (480, 640)), ((640, 480), (960, 1280))):
        m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=make_onnx(*net), size=net, max_batch=16)
        try:
            rng = np.random.default_rng(net[0])
            ims = [rng.integers(0, 256, (*shape, 3), dtype=np.uint8),
                   np.where(rng.random((*shape, 3)) < 0.5, 0, 255).astype(np.uint8),
                   np.full((*shape, 3), 255, np.uint8), np.zeros((*shape, 3), np.uint8),
                   _smooth(1, *shape, seed=2)[0]]
            for im in ims:  # single frame: f32 kernel (too few whole-row CTAs to fill the GPU)
                np.testing.assert_array_equal(m.preproc_u8(im), hotpath.resize_triangle(im, *net))
            # batch of 10: the integer kernel (whole-row CTAs need >= 148 of them)
            batch = np.stack(ims + ims)
            got = m.preproc_u8_batch(batch)
            for i, im in enumerate(batch):
                np.testing.assert_array_equal(got[i], hotpath.resize_triangle(im, *net), err_msg=str(i))
        finally:
            m.close()


# ---------------------------------------------------------------------------------------------
# R2: normalise + HWC->NCHW, bit-exact (both presets)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("preset", [0, 1])
def test_normalised_tensor_bit_exact(make_onnx, preset):
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=make_onnx(320, 240), norm_preset=preset)
    try:
        im = _noise(1, seed=2)[0]
        exp = hotpath.normalise_nchw(hotpath.resize_triangle(im, 320, 240), preset)[None]
        np.testing.assert_array_equal(m.preproc(im), exp)
    finally:
        m.close()


# ---------------------------------------------------------------------------------------------
# M1: the conv stack, layer by layer and end to end
# ---------------------------------------------------------------------------------------------
def _layer_report(m, oracle, frame):
    import torch
    with torch.no_grad():
        _, env = oracle.net(torch.from_numpy(oracle.preproc(frame)), keep=True)
    rep = []
    for name, (idx, c, h, w) in m.tensors().items():
        if name not in env:
            continue
        got = m.tensor_read(idx, 0, (c, h, w))
        ref = env[name].numpy()[0]
        rep.append((name, (c, h, w), float(np.abs(got - ref).max()), float(np.abs(ref).max())))
    return rep


@pytest.mark.parametrize("flags", [_capi.UF_FLAG_FORCE_GENERIC, _capi.UF_FLAG_NO_FUSION, _capi.UF_FLAG_NO_TC,
                                   _capi.UF_FLAG_FUSE_DW_TC, _capi.UF_FLAG_TMA_SIMT_PW, _capi.UF_FLAG_DENSE3_TC, 0])
def test_every_materialised_layer_matches_oracle(make_onnx, oracle320, flags):
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=make_onnx(320, 240), flags=flags)
    try:
        frame = _smooth(1, seed=3)[0]
        m.run(frame)
        rep = _layer_report(m, oracle320, frame)
        assert len(rep) >= (40 if flags in (_capi.UF_FLAG_FORCE_GENERIC, _capi.UF_FLAG_NO_FUSION) else 25)
        bad = [r for r in rep if not r[2] <= 1e-4 * max(1.0, r[3])]
        assert not bad, f"layers off (name, chw, max_abs_diff, ref_max): {bad[:8]}"
    finally:
        m.close()


@pytest.mark.parametrize("cfg", [dict(wh=(320, 240), variant="RFB"), dict(wh=(320, 240), variant="slim"),
                                 dict(wh=(640, 480), variant="RFB"), dict(wh=(320, 240), variant="RFB", with_bn=True, seed=3),
                                 # maps 44x36 / 22x18 / 11x9 / 6x5: no kernel tile divides them (partial tiles everywhere)
                                 dict(wh=(352, 288), variant="RFB", seed=5), dict(wh=(352, 288), variant="slim", seed=6)])
def test_raw_outputs_within_1e4_and_detections_match(make_onnx, test_pics, cfg):
    wh = cfg["wh"]
    path = make_onnx(*wh, variant=cfg["variant"], with_bn=cfg.get("with_bn", False), seed=cfg.get("seed", 0), cls_bias=-0.75)
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, size=wh, max_batch=8)
    oracle = UltrafaceOracle(path, *wh, 0.5, 0.5)
    try:
        frames = [_noise(1, seed=11)[0], _smooth(1, seed=12)[0]] + [test_pics[k] for k in sorted(test_pics)[:2]]
        dets, counts = m.run_batch(frames, cap=512)
        s_gpu, b_gpu = m.raw_outputs(0, len(frames))
        s_ref, b_ref = oracle.raw(frames)
        assert np.abs(s_gpu - s_ref).max() <= TOL, np.abs(s_gpu - s_ref).max()
        assert np.abs(b_gpu - b_ref).max() <= TOL, np.abs(b_gpu - b_ref).max()
        for i in range(len(frames)):
            # post-processing on the GPU's own raw tensors must be reproduced exactly by the oracle
            ref, _ = hotpath.postproc(s_gpu[i], b_gpu[i], 0.5, 0.5)
            assert counts[i] == len(ref)
            np.testing.assert_array_equal(dets[i], ref[:512])
            # and against the oracle's raw tensors: identical set unless a decision is within tolerance
            ref2, _ = hotpath.postproc(s_ref[i], b_ref[i], 0.5, 0.5)
            _assert_detection_sets_match(ref, ref2, s_ref[i], b_ref[i])
    finally:
        m.close()


def test_single_frame_run_mirrors_reference_api(model320, oracle320, test_pics):
    """The reference's own test shape (integration_tests.rs:30-35): model.run(&image) -> Vec<(Bbox, f32)>."""
    for k in sorted(test_pics):
        got = model320.run(test_pics[k], cap=2048)
        s, b = model320.raw_outputs(0, 1)
        exp = oracle320.postproc(s[0], b[0])
        assert len(got) == len(exp), k
        for (gb, gc), (eb, ec) in zip(got, exp):
            assert gb == eb and gc == ec
        confs = [c for _, c in got]
        assert confs == sorted(confs, reverse=True)  # descending certainty (nn.rs:107-108)


# ---------------------------------------------------------------------------------------------
# P1-P6: post-processing alone on adversarial raw tensors (exact)
# ---------------------------------------------------------------------------------------------
def _random_raw(K, seed, tie_frac=0.2, spread=0.3):
    rng = np.random.default_rng(seed)
    c = rng.random((K, 2)).astype(np.float32) * 0.6 + 0.2
    wh = (rng.random((K, 2)).astype(np.float32) * spread).astype(np.float32)
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    scores = rng.random((K, 2)).astype(np.float32)
    ties = rng.integers(0, K, int(K * tie_frac))
    scores[ties, 1] = np.float32(0.75)
    dup = rng.integers(0, K, max(1, K // 20))
    boxes[dup] = boxes[(dup + 1) % K]  # exact duplicate boxes
    return scores, boxes


@pytest.mark.parametrize("K,seed,spread", [(1, 0, 0.3), (5, 1, 0.3), (257, 2, 0.3), (300, 3, 0.05), (4420, 4, 0.3),
                                           (4420, 5, 0.02), (10000, 6, 0.1), (17640, 7, 0.2), (20000, 8, 0.01)])
def test_postproc_exact_on_adversarial_inputs(model320, K, seed, spread):
    scores, boxes = _random_raw(K, seed, spread=spread)
    dets, idx = model320.postproc(scores, boxes)
    ref, ridx = hotpath.postproc(scores, boxes, 0.5, 0.5)
    np.testing.assert_array_equal(idx, ridx)
    np.testing.assert_array_equal(dets, ref)


@pytest.mark.parametrize("min_conf,max_iou", [(0.3, 0.3), (0.9, 0.7), (0.5, 0.0), (0.5, -1.0), (-1.0, 0.5), (0.5, 1.0)])
def test_postproc_other_thresholds(make_onnx, min_conf, max_iou):
    """Thresholds other than the reference's 0.5/0.5 (nn.rs:55 takes them as arguments): max_iou = 0 and < 0
    exercise the exact-division path, min_conf < 0 makes every prior a candidate (17 640 keys: the sort spills
    from shared memory to the global scratch)."""
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, max_iou, min_conf, onnx_path=make_onnx(320, 240))
    try:
        for K, seed, spread in [(4420, 11, 0.1), (17640, 12, 0.05)]:
            scores, boxes = _random_raw(K, seed, spread=spread)
            dets, idx = m.postproc(scores, boxes)
            ref, ridx = hotpath.postproc(scores, boxes, min_conf, max_iou)
            np.testing.assert_array_equal(idx, ridx)
            np.testing.assert_array_equal(dets, ref)
    finally:
        m.close()


def test_postproc_edge_cases(model320):
    K = 64
    boxes = np.tile(np.float32([[0.1, 0.1, 0.4, 0.4]]), (K, 1))
    for scores1 in (np.full(K, 0.5, np.float32),                       # nothing is > 0.5 (strict)
                    np.full(K, np.nan, np.float32),                    # NaN dropped
                    np.full(K, 0.9, np.float32),                       # all tied, all identical boxes
                    np.linspace(0.5, 1.0, K).astype(np.float32)):
        scores = np.stack([1 - scores1, scores1], 1).astype(np.float32)
        dets, idx = model320.postproc(scores, boxes)
        ref, ridx = hotpath.postproc(scores, boxes, 0.5, 0.5)
        np.testing.assert_array_equal(idx, ridx)
        np.testing.assert_array_equal(dets, ref)
    # ill-defined boxes (bottom-right above top-left) have zero area and never suppress
    bad = np.tile(np.float32([[0.5, 0.5, 0.1, 0.1]]), (8, 1))
    sc = np.stack([np.zeros(8), np.linspace(0.6, 0.9, 8)], 1).astype(np.float32)
    dets, idx = model320.postproc(sc, bad)
    ref, ridx = hotpath.postproc(sc, bad, 0.5, 0.5)
    np.testing.assert_array_equal(idx, ridx)
    assert len(idx) == 8


# ---------------------------------------------------------------------------------------------
# batching / API consistency
# ---------------------------------------------------------------------------------------------
def test_batch_chunking_and_mixed_sizes_equal_single_frame(make_onnx, test_pics):
    path = make_onnx(320, 240, cls_bias=-0.75)
    single = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path)
    batched = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=16, chunk=3, slots=2)
    try:
        pics = [test_pics[k] for k in sorted(test_pics)]
        frames = [_noise(1, seed=i)[0] for i in range(5)] + pics + [_noise(1, 240, 320, seed=9)[0]] + pics[:2] + \
                 [_smooth(1, 720, 1280, seed=4)[0]]
        dets, counts = batched.run_batch(frames, cap=256)
        for i, f in enumerate(frames):
            exp = single.run(f, cap=256)
            assert counts[i] == len(exp), i
            np.testing.assert_array_equal(dets[i], np.float32([[*b, c] for b, c in exp]).reshape(-1, 5))
        with pytest.raises(nn.UltrafaceError) as e:
            batched.run_batch(frames * 2)
        assert e.value.code == 7  # UF_ERR_CAPACITY
        assert batched.run_batch([], cap=4)[1] == []
    finally:
        single.close()
        batched.close()


def test_device_resident_input_equals_host_input(make_onnx):
    import torch
    path = make_onnx(320, 240, cls_bias=-0.75)
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=24, chunk=8)
    try:
        frames = _noise(24, seed=21)
        host, hc = m.run_batch(list(frames), cap=128)
        s_h, b_h = m.raw_outputs(0, 24)
        d = torch.from_numpy(frames).cuda()
        dev, dc = m.run_batch_device(d.data_ptr(), 640, 480, 24, cap=128)
        s_d, b_d = m.raw_outputs(0, 24)
        assert hc == dc
        np.testing.assert_array_equal(s_h, s_d)
        np.testing.assert_array_equal(b_h, b_d)
        for a, b in zip(host, dev):
            np.testing.assert_array_equal(a, b)
        pinned = nn.PinnedFrames(24, 480, 640)
        pinned.array[:] = frames
        pin, pc = m.run_batch_ptr(pinned.ptr, 640, 480, 24, cap=128)
        assert pc == hc
        pinned.free()
        # identity-size device input (RFB-640 style): the stem reads the caller's buffer directly
        small = torch.from_numpy(_noise(4, 240, 320, seed=22)).cuda()
        dev2, _ = m.run_batch_device(small.data_ptr(), 320, 240, 4, cap=128)
        host2, _ = m.run_batch(list(small.cpu().numpy()), cap=128)
        for a, b in zip(host2, dev2):
            np.testing.assert_array_equal(a, b)
    finally:
        m.close()


def test_programmatic_dependent_launch_gives_identical_results(make_onnx):
    """UF_FLAG_PDL: kernels start before their predecessor has finished and wait with griddepcontrol.wait."""
    path = make_onnx(320, 240, cls_bias=-0.75)
    frames = list(_noise(40, seed=77))
    a = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=40)
    ref = [a.run_batch(frames, cap=128) for _ in range(3)][-1]
    sa, ba = a.raw_outputs(0, 40)
    a.close()
    b = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=40, flags=_capi.UF_FLAG_PDL)
    try:
        got = [b.run_batch(frames, cap=128) for _ in range(3)][-1]  # third call replays the captured graph
        sb, bb = b.raw_outputs(0, 40)
        np.testing.assert_array_equal(sa, sb)
        np.testing.assert_array_equal(ba, bb)
        assert got[1] == ref[1]
    finally:
        b.close()


def test_concurrent_calls_on_one_handle(make_onnx):
    """`UltrafaceModel` is Send + Sync (inferer.rs:29-50): calls from several host threads on one handle run on
    separate lanes and must give exactly the single-threaded results."""
    import threading
    path = make_onnx(320, 240, cls_bias=-0.75)
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=48, lanes=3)
    try:
        sets = [_noise(48, seed=40 + i) for i in range(3)]
        expect = [m.run_batch(list(fs), cap=128) for fs in sets]
        got = [None] * 3

        def work(i):
            for _ in range(4):
                got[i] = m.run_batch(list(sets[i]), cap=128)
        ts = [threading.Thread(target=work, args=(i,)) for i in range(3)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        for i in range(3):
            assert got[i][1] == expect[i][1]
            for a, b in zip(got[i][0], expect[i][0]):
                np.testing.assert_array_equal(a, b)
    finally:
        m.close()


def test_full_size_batch_properties(make_onnx):
    """BASELINE config sizes (batch 256, 640x480): size-independent properties instead of the oracle —
    duplicated frames give identical results wherever they sit in the batch, detections are sorted,
    every detection is a raw (box, score) row above the threshold, no pair exceeds max_iou."""
    path = make_onnx(320, 240, cls_bias=-0.75)
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=256)
    try:
        base = _noise(8, seed=31)
        frames = np.concatenate([base] * 32)  # 256 frames, period 8
        dets, counts = m.run_batch(list(frames), cap=256)
        s, b = m.raw_outputs(0, 256)
        for i in range(256):
            assert counts[i] == counts[i % 8]
            np.testing.assert_array_equal(dets[i], dets[i % 8])
            np.testing.assert_array_equal(s[i], s[i % 8])
        for i in range(8):
            d = dets[i]
            assert (np.diff(d[:, 4]) <= 0).all() and (d[:, 4] > 0.5).all()
            rows = {tuple(r) for r in np.concatenate([b[i], s[i][:, 1:]], 1).tolist()}
            assert all(tuple(r) in rows for r in d.tolist())
            for x in range(len(d)):
                for y in range(x):
                    assert hotpath.iou(d[x, :4], d[y, :4]) <= 0.5
        # spot-check two of them against the oracle end to end
        oracle = UltrafaceOracle(path, 320, 240, 0.5, 0.5)
        s_ref, b_ref = oracle.raw(list(base[:2]))
        assert np.abs(s[:2] - s_ref).max() <= TOL and np.abs(b[:2] - b_ref).max() <= TOL
    finally:
        m.close()


def test_profile_counters(model320):
    model320.profile_enable(True)
    model320.profile_reset()
    n0 = model320.launch_count()
    model320.run(_noise(1, seed=1)[0])
    stats = model320.profile_read()
    model320.profile_enable(False)
    assert model320.launch_count() - n0 == sum(s["launches"] for s in stats)
    names = {s["name"].split("[")[0] for s in stats}
    assert {"resize_triangle", "stem_3x3s2_u8", "tail_post_softmax_decode_nms"} <= names
    assert any(n.startswith("fused_dw3x3_pw1x1") for n in names) and "pointwise1x1_tcgen05" in names
    assert all(s["device_ms"] > 0 for s in stats)
