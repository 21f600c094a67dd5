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


def _iou_matrix(bx):
    x0 = np.maximum(bx[:, None, 0], bx[None, :, 0]); y0 = np.maximum(bx[:, None, 1], bx[None, :, 1])
    x1 = np.minimum(bx[:, None, 2], bx[None, :, 2]); y1 = np.minimum(bx[:, None, 3], bx[None, :, 3])
    inter = np.clip(x1 - x0, 0, None) * np.clip(y1 - y0, 0, None)
    area = np.clip(bx[:, 2] - bx[:, 0], 0, None) * np.clip(bx[:, 3] - bx[:, 1], 0, None)
    return inter / (area[:, None] + area[None, :] - inter + 1e-7)


def _assert_detection_sets_match(gpu_idx, ref_idx, scores, boxes, min_conf=0.5, max_iou=0.5, tol=TOL, box_err=TOL):
    """Cross-implementation detection parity (BASELINE.json north_star: "the post-NMS detection set must be identical
    except for candidates whose score lies within that tolerance of the threshold"). `gpu_idx` / `ref_idx` are the
    prior indices selected from the GPU's and from the ORACLE's raw tensors; `scores` / `boxes` are the oracle's.

    Every prior in the symmetric difference must be EXPLAINED, one by one, by a greedy decision that was within
    tolerance: (a) its own score within tol of min_conf; (b) an overlapping candidate whose IoU with it is within
    the IoU tolerance of max_iou; (c) an overlapping candidate whose score is within tol of its own (processing order
    may swap); (d) cascade — it overlaps (IoU > max_iou - tolerance) a candidate that is itself explained. The IoU
    tolerance follows from the box error: d(IoU) <= ~8 * box_err / shorter side."""
    gpu_idx, ref_idx = [int(i) for i in gpu_idx], [int(i) for i in ref_idx]
    diff = set(gpu_idx) ^ set(ref_idx)
    if not diff:
        return 0
    cand = np.nonzero(scores[:, 1] > min_conf - tol)[0]
    pos = {int(k): i for i, k in enumerate(cand)}
    assert all(k in pos for k in diff), "a differing detection is not even a near-threshold candidate"
    sc, bx = scores[cand, 1], boxes[cand]
    side = np.minimum(bx[:, 2] - bx[:, 0], bx[:, 3] - bx[:, 1]).clip(1e-4, None)
    iou = _iou_matrix(bx)
    np.fill_diagonal(iou, 0.0)
    tol_iou = 1e-4 + 8.0 * box_err / np.minimum(side[:, None], side[None, :])
    fragile = np.abs(sc - min_conf) <= tol                                        # (a)
    fragile |= (np.abs(iou - max_iou) <= tol_iou).any(1)                          # (b)
    fragile |= ((np.abs(sc[:, None] - sc[None, :]) <= tol) & (iou > max_iou - tol_iou)).any(1)  # (c)
    explained = fragile.copy()
    near = iou > max_iou - tol_iou
    for _ in range(len(cand)):                                                    # (d) cascade to a fixpoint
        grown = explained | (near & explained[None, :]).any(1)
        if (grown == explained).all():
            break
        explained = grown
    unexplained = [k for k in diff if not explained[pos[k]]]
    assert not unexplained, f"{len(unexplained)} of {len(diff)} differing detections have no decision within tolerance: {unexplained[:8]}"
    return len(diff)


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


@pytest.mark.parametrize("net,round_intermediate", [((320, 240), False), ((640, 480), False), ((320, 240), True)])
def test_fused_resize_stem_kernel_sees_the_reference_pixels(make_onnx, net, round_intermediate):
    """Frames at exactly twice the network size never exist resized in memory: one kernel resamples (packed integer
    arithmetic for the interior, f32 tap tables for the border row / column), normalises and convolves. Its debug output —
    the u8 pixels it convolved — must equal the oracle's resize bit for bit, and the whole network must give exactly what
    the two-kernel path (UF_FLAG_NO_PRESTEM) gives."""
    w, h = net
    path = make_onnx(w, h, cls_bias=-0.75)
    kw = dict(onnx_path=path, size=net, max_batch=16, resize_round_intermediate=round_intermediate)
    fused = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, **kw)
    plain = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, flags=_capi.UF_FLAG_NO_PRESTEM, **kw)
    try:
        rng = np.random.default_rng(w)
        shape = (2 * h, 2 * w)
        ims = np.stack([rng.integers(0, 256, (*shape, 3), dtype=np.uint8),
                        np.where(rng.random((*shape, 3)) < 0.5, 0, 255).astype(np.uint8),
                        np.full((*shape, 3), 255, np.uint8), np.zeros((*shape, 3), np.uint8),
                        _smooth(1, *shape, seed=2)[0], rng.integers(0, 256, (*shape, 3), dtype=np.uint8)])
        got = fused.prestem_u8(ims)
        for i, im in enumerate(ims):
            np.testing.assert_array_equal(got[i], hotpath.resize_triangle(im, w, h, round_intermediate), err_msg=str(i))
        for n in (1, len(ims)):  # single frame (latency path) and a batch
            da, ca = fused.run_batch(list(ims[:n]), cap=512)
            sa, ba = fused.raw_outputs(0, n)
            db, cb = plain.run_batch(list(ims[:n]), cap=512)
            sb, bb = plain.raw_outputs(0, n)
            np.testing.assert_array_equal(sa, sb)
            np.testing.assert_array_equal(ba, bb)
            assert ca == cb
        names = {s_["name"].split("[")[0] for s_ in _profile_one(fused, ims[0])}
        assert "resize2x_norm_stem_u8" in names and "resize_triangle" not in names
        names = {s_["name"].split("[")[0] for s_ in _profile_one(plain, ims[0])}
        assert "resize_triangle" in names and "stem_3x3s2_u8" in names
    finally:
        fused.close()
        plain.close()


def _profile_one(m, frame):
    m.profile_enable(True)
    m.profile_reset()
    m.run(frame)
    stats = m.profile_read()
    m.profile_enable(False)
    return stats


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
        _check_end_to_end(dets, counts, s_gpu, b_gpu, s_ref, b_ref, 512)
    finally:
        m.close()


def _check_end_to_end(dets, counts, s_gpu, b_gpu, s_ref, b_ref, cap, min_conf=0.5, max_iou=0.5):
    """Raw tensors within TOL; post-processing of the GPU's own raw tensors reproduced EXACTLY by the oracle; and the
    selection from the GPU's raw tensors equals the selection from the oracle's, prior for prior, except explained ones."""
    es, eb = float(np.abs(s_gpu - s_ref).max()), float(np.abs(b_gpu - b_ref).max())
    assert es <= TOL and eb <= TOL, (es, eb)
    flips = 0
    for i in range(len(s_ref)):
        ref, ridx = hotpath.postproc(s_gpu[i], b_gpu[i], min_conf, max_iou)
        assert counts[i] == len(ref), i
        np.testing.assert_array_equal(dets[i], ref[:cap])
        ref2, ridx2 = hotpath.postproc(s_ref[i], b_ref[i], min_conf, max_iou)
        flips += _assert_detection_sets_match(ridx, ridx2, s_ref[i], b_ref[i], min_conf, max_iou, tol=max(4 * es, 1e-6),
                                              box_err=max(eb, 1e-7))
    return es, eb, flips


def test_reference_photos_at_reference_geometry(make_onnx, test_pics):
    """The reference's own inputs at its own geometry (integration_tests.rs:30-35): all eight 640-px-wide photos, whole,
    through RFB-640 (640 -> 640 identity horizontally, 427/462/676/960 -> 480 vertically), RFB-320 and slim-320
    (640 -> 320, x -> 240): resize bit-exact, raw tensors within 1e-4, detections as the oracle's."""
    assert len(test_pics) == 8 and all(im.shape[1] == 640 for im in test_pics.values())
    assert sorted({im.shape[0] for im in test_pics.values()}) == [427, 462, 676, 960]
    frames = [test_pics[k] for k in sorted(test_pics)]
    for wh, variant in (((640, 480), "RFB"), ((320, 240), "RFB"), ((320, 240), "slim")):
        path = make_onnx(*wh, variant=variant, cls_bias=-1.0)
        m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, size=wh, max_batch=8)
        oracle = UltrafaceOracle(path, *wh, 0.5, 0.5)
        try:
            for f in frames:
                np.testing.assert_array_equal(m.preproc_u8(f), hotpath.resize_triangle(f, *wh))
            dets, counts = m.run_batch(frames, cap=4096)
            s_gpu, b_gpu = m.raw_outputs(0, 8)
            s_ref, b_ref = oracle.raw(frames)
            _check_end_to_end(dets, counts, s_gpu, b_gpu, s_ref, b_ref, 4096)
            one = m.run(frames[0], cap=4096)  # the single-frame call gives what the batch gave
            assert len(one) == counts[0]
        finally:
            m.close()


def test_reference_face_counts(real_rfb640_path, test_pics, reference_face_counts):
    """The reference's own known answers (integration_tests.rs:20-34): RFB-640 at 0.5 / 0.5 finds 3,6,4,3,1,1,10,0 faces.
    Needs the weight file the reference downloads (nn.rs:21); skipped while it is absent. Photos are decoded by PIL
    (libjpeg-turbo) here, by the `jpeg-decoder` crate there: pixels may differ by an LSB before the path starts."""
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W640H480, 0.5, 0.5, onnx_path=real_rfb640_path)
    try:
        got = {k: len(m.run(im)) for k, im in test_pics.items()}
    finally:
        m.close()
    assert got == reference_face_counts


@pytest.mark.parametrize("wh,with_bn", [((320, 240), False), ((320, 240), True), ((640, 480), False)])
def test_upstream_shaped_export(make_onnx, wh, with_bn):
    """The files the reference downloads are un-simplified PyTorch-1.x opset-9 exports (nn.rs:21-22,165-172): computed
    reshape targets (Shape -> Gather -> Unsqueeze -> Concat), Constant-node priors sliced in the graph, attribute-form
    Slice, re-sliced centre-form boxes. Same weights written both ways must run identically on the GPU, and match the
    oracle's interpreter of the upstream-shaped file."""
    a = make_onnx(*wh, seed=4, cls_bias=-0.75, with_bn=with_bn)
    b = make_onnx(*wh, seed=4, cls_bias=-0.75, with_bn=with_bn, style="upstream")
    frames = [_noise(1, seed=51)[0], _smooth(1, seed=52)[0], _noise(1, 427, 640, seed=53)[0]]
    out = []
    for path in (a, b):
        m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, size=wh, max_batch=4)
        try:
            dets, counts = m.run_batch(frames, cap=2048)
            out.append((dets, counts, *m.raw_outputs(0, 3)))
        finally:
            m.close()
    np.testing.assert_array_equal(out[0][2], out[1][2])
    np.testing.assert_array_equal(out[0][3], out[1][3])
    assert out[0][1] == out[1][1]
    s_ref, b_ref = UltrafaceOracle(b, *wh, 0.5, 0.5).raw(frames)
    _check_end_to_end(out[1][0], out[1][1], out[1][2], out[1][3], s_ref, b_ref, 2048)


def test_graph_with_other_variances_follows_the_graph(make_onnx):
    """A graph whose decode uses centre 0.2 / size 0.1: the loader takes the constants from the graph (it does not
    assume 0.1 / 0.2) and the GPU follows the graph, as tract would."""
    path = make_onnx(320, 240, seed=2, tail="swapped_variances")
    assert nn.onnx_inspect(path, 320, 240)["center_variance"] == pytest.approx(0.2)
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path)
    try:
        f = _noise(1, seed=3)[0]
        m.run(f)
        s, b = m.raw_outputs(0, 1)
        s_ref, b_ref = UltrafaceOracle(path, 320, 240).raw([f])
        assert np.abs(b - b_ref).max() <= TOL and np.abs(s - s_ref).max() <= TOL
    finally:
        m.close()


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: RFB-640, batch 64, dense candidates (NMS-heavy)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cls_bias,p_lo,p_hi,n_oracle", [(-1.5, 0.004, 0.02, 64), (-0.75, 0.03, 0.08, 64), (0.65, 0.4, 0.6, 12)])
def test_config5_rfb640_batch64_nms_heavy(make_onnx, cls_bias, p_lo, p_hi, n_oracle):
    """p = 1 % / 5 % / 50 % of the 17 640 priors above min_confidence (seeded head bias): GPU against the oracle end to
    end on the whole batch (the 50 % case: the whole batch on the GPU, the first 12 frames against the oracle, whose NMS
    takes ~1 s per frame there; the rest through invariants)."""
    path = make_onnx(640, 480, cls_bias=cls_bias)
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W640H480, 0.5, 0.5, onnx_path=path, max_batch=64)
    oracle = UltrafaceOracle(path, 640, 480, 0.5, 0.5)
    try:
        frames = list(_noise(64, seed=64))
        cap = 17640
        dets, counts = m.run_batch(frames, cap=cap)
        s_gpu, b_gpu = m.raw_outputs(0, 64)
        p = float((s_gpu[..., 1] > 0.5).mean())
        assert p_lo <= p <= p_hi, p
        s_ref, b_ref = oracle.raw(frames[:n_oracle])
        _check_end_to_end(dets[:n_oracle], counts[:n_oracle], s_gpu[:n_oracle], b_gpu[:n_oracle], s_ref, b_ref, cap)
        for i in range(n_oracle, 64):  # invariants: sorted, above threshold, rows of the raw tensors
            d = dets[i]
            assert len(d) == counts[i] and (np.diff(d[:, 4]) <= 0).all() and (d[:, 4] > 0.5).all()
        # same frames device-resident (32-frame stages, other kernel path for the tail): identical
        import torch
        dev = torch.from_numpy(np.stack(frames)).cuda()
        dets2, counts2 = m.run_batch_device(dev.data_ptr(), 640, 480, 64, cap=cap)
        assert counts2 == counts
        for a, b in zip(dets, dets2):
            np.testing.assert_array_equal(a, b)
    finally:
        m.close()


def test_nms_heavy_rfb320_more_than_fast_path_detections(make_onnx):
    """cls_bias 3: every prior is a candidate, ~3 350 detections per frame — far more than the 128 that come back with the
    counts; the rest arrives by one strided copy per stage. Exact against the oracle's post-processing."""
    path = make_onnx(320, 240, cls_bias=3.0)
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=40)
    try:
        frames = list(_noise(40, seed=5))
        for cap in (4420, 1000, 100):
            dets, counts = m.run_batch(frames, cap=cap)
            s, b = m.raw_outputs(0, 40)
            for i in range(40):
                ref, _ = hotpath.postproc(s[i], b[i], 0.5, 0.5)
                assert counts[i] == len(ref) and counts[i] > 3000
                np.testing.assert_array_equal(dets[i], ref[:cap])
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


def test_postproc_signed_zero_scores_tie(make_onnx):
    """Rust's partial_cmp calls -0.0 and +0.0 equal (nn.rs:134), so among zero scores the prior index alone decides."""
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, -1.0, onnx_path=make_onnx(320, 240))
    try:
        K = 700
        rng = np.random.default_rng(0)
        scores, boxes = _random_raw(K, 9, spread=0.1)
        z = rng.integers(0, K, 300)
        scores[z, 1] = np.where(rng.random(300) < 0.5, np.float32(0.0), np.float32(-0.0))
        dets, idx = m.postproc(scores, boxes)
        ref, ridx = hotpath.postproc(scores, boxes, -1.0, 0.5)
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
    batched = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=24, chunk=3, slots=2)
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


def test_failed_call_leaves_no_stale_results(make_onnx):
    """A call that fails after some pipeline stages were queued (here: injected; in the field cudaMalloc / a bad size)
    must not leave pending slots behind: the next, smaller call would receive the failed call's results out of bounds."""
    path = make_onnx(320, 240, cls_bias=-0.75)
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=64, host_chunk=8, lanes=1)
    try:
        frames = list(_noise(64, seed=8))
        good, gc = m.run_batch(frames, cap=64)
        for stages in (1, 3, 6):
            m.debug_fail_after(stages)
            with pytest.raises(nn.UltrafaceError) as e:
                m.run_batch(frames, cap=64)
            assert e.value.code == 5
            m.debug_fail_after(-1)
            # a one-frame call right after: a guard page's worth of canaries around its tiny output arrays
            out = np.full((3, 4, 5), 7.0, np.float32)
            cnt = (_capi.C.c_uint32 * 3)(99, 99, 99)
            f = np.ascontiguousarray(frames[5])
            rc = _capi.load().uf_infer(m._h, f.ctypes.data_as(_capi.C.c_void_p), 640, 480,
                                       out[1].ctypes.data_as(_capi.C.POINTER(_capi.uf_det)), 4,
                                       _capi.C.cast(_capi.C.addressof(cnt) + 4, _capi.C.POINTER(_capi.C.c_uint32)))
            assert rc == 0
            assert cnt[0] == 99 and cnt[2] == 99 and cnt[1] == gc[5]
            assert (out[0] == 7.0).all() and (out[2] == 7.0).all()
            np.testing.assert_array_equal(out[1][: min(4, gc[5])], good[5][:4])
            again, ac = m.run_batch(frames, cap=64)
            assert ac == gc
    finally:
        m.close()


def test_resize_table_cache_is_bounded(model320):
    """Source sizes come from network peers: cycling through many must neither grow device memory without bound nor
    break frames whose table was evicted and rebuilt."""
    import torch
    rng = np.random.default_rng(3)
    first = rng.integers(0, 256, (50, 70, 3), dtype=np.uint8)
    exp = hotpath.resize_triangle(first, 320, 240)
    np.testing.assert_array_equal(model320.preproc_u8(first), exp)
    free0 = torch.cuda.mem_get_info()[0]
    for i in range(80):
        im = rng.integers(0, 256, (40 + i, 60 + 2 * i, 3), dtype=np.uint8)
        got = model320.preproc_u8(im)
        if i % 16 == 0:
            np.testing.assert_array_equal(got, hotpath.resize_triangle(im, 320, 240))
    np.testing.assert_array_equal(model320.preproc_u8(first), exp)  # evicted long ago, rebuilt
    assert free0 - torch.cuda.mem_get_info()[0] < 64 << 20


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
    assert {"resize2x_norm_stem_u8", "tail_post_softmax_decode_nms"} <= names
    assert any(n.startswith("fused_dw3x3_pw1x1") for n in names) and "pointwise1x1_tcgen05" in names
    assert all(s["device_ms"] > 0 for s in stats)
