"""N3, rectangles + JPEG encode after the path (inferer.rs:38-39, 58-92). The reference encodes with
libjpeg-turbo (`turbojpeg::compress_image(&frame, 95, Sub2x2)`); libjpeg-turbo is in this image behind PIL, so it is the
checker: the file PIL writes from the same pixels at the same quality must hold exactly the coefficients and tables ours
holds. The rectangle overlay is checked against the oracle's restatement of imageproc's `draw_hollow_rect` (unpinned)."""
import io

import numpy as np
import pytest
from PIL import Image

from infercam_onnx_b200 import nn
from oracle import draw as odraw


def _pil_jpeg(arr, quality=95):
    b = io.BytesIO()
    Image.fromarray(arr).save(b, "JPEG", quality=quality, subsampling=2)
    return b.getvalue()


def _to_plane_raster(info, coefs):
    """decode-order blocks (MCU-interleaved) -> per padded component plane, raster order (the encoder's layout)."""
    w, h = info["w"], info["h"]
    mx, my = (w + 15) // 16, (h + 15) // 16
    planes = [np.zeros((my * 2, mx * 2, 64), np.int16), np.zeros((my, mx, 64), np.int16), np.zeros((my, mx, 64), np.int16)]
    c = coefs.reshape(my, mx, 6, 64)
    planes[0][0::2, 0::2] = c[:, :, 0]
    planes[0][0::2, 1::2] = c[:, :, 1]
    planes[0][1::2, 0::2] = c[:, :, 2]
    planes[0][1::2, 1::2] = c[:, :, 3]
    planes[1][:] = c[:, :, 4]
    planes[2][:] = c[:, :, 5]
    return np.concatenate([p.reshape(-1, 64) for p in planes])


@pytest.mark.parametrize("shape", [(480, 640), (462, 640), (301, 333), (9, 17), (16, 8), (720, 1280)])
def test_huffman_writer_round_trips_libjpeg_turbo_files(test_pics, shape):
    """Host half alone: decode a libjpeg-turbo file to coefficients, write them back with our Huffman coder / file writer,
    and libjpeg-turbo must decode our file to exactly the pixels of its own (incl. frames whose last MCUs hold dummy blocks)."""
    pic = np.ascontiguousarray(np.resize(test_pics["omar-lopez-T6zu4jFhVwg"], (*shape, 3)) if shape[0] > 462 or shape[1] > 640
                               else test_pics["omar-lopez-T6zu4jFhVwg"][: shape[0], : shape[1]])
    for q in (95, 60):
        ref = _pil_jpeg(pic, q)
        info, coefs = nn.jpeg_coefficients(ref)
        lum, chr_ = nn.jpeg_quality_tables(q)
        np.testing.assert_array_equal(info["quant"][0], lum)  # jpeg_set_quality scaling = libjpeg-turbo's
        np.testing.assert_array_equal(info["quant"][1], chr_)
        ours = nn.jpeg_write_coefficients(shape[1], shape[0], q, _to_plane_raster(info, coefs))
        info2, coefs2 = nn.jpeg_coefficients(ours)
        np.testing.assert_array_equal(coefs2, coefs)
        np.testing.assert_array_equal(np.asarray(Image.open(io.BytesIO(ours)).convert("RGB")),
                                      np.asarray(Image.open(io.BytesIO(ref)).convert("RGB")))
        assert abs(len(ours) - len(ref)) <= 64  # same entropy-coded size, headers differ by a few bytes at most


def _boxes():
    return np.float32([[0.10, 0.20, 0.30, 0.60, 0.99], [0.5, 0.5, 0.5001, 0.9, 0.8],      # thin: width casts to 0 at small scales
                       [-0.2, -0.1, 0.25, 0.3, 0.7], [0.8, 0.7, 1.4, 1.3, 0.6],              # partly outside
                       [2.0, 2.0, 3.0, 3.0, 0.55], [0.3, 0.3, 0.2, 0.2, 0.5],                # wholly outside; inverted (skipped)
                       [0.0, 0.0, 1.0, 1.0, 0.51]])                                          # the whole frame


@pytest.mark.gpu
def test_gpu_rectangles_match_the_oracle(make_onnx, test_pics):
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=make_onnx(320, 240))
    try:
        for pic, (sw, sh) in ((test_pics["omar-lopez-T6zu4jFhVwg"], (640.0, 462.0)), (test_pics["bruce-mars-ZXq7xoo98b0"], (1280.0, 720.0)),
                              (test_pics["michael-dam-mEZ3PoFGs_k"][:100, :90], (90.0, 100.0))):
            got = m.draw_boxes(pic, _boxes(), sw, sh)
            np.testing.assert_array_equal(got, odraw.draw_boxes(pic, _boxes(), sw, sh))
            assert (got != pic).any()
        np.testing.assert_array_equal(m.draw_boxes(pic, np.zeros((0, 5), np.float32), 1.0, 1.0), pic)
    finally:
        m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(480, 640), (462, 640), (301, 333), (9, 17), (720, 1280)])
def test_gpu_encoder_writes_libjpeg_turbos_coefficients(make_onnx, test_pics, shape):
    """Colour conversion, 4:2:0 downsampling, forward DCT and quantisation on the GPU: for the same pixels and quality the
    coefficients in our file equal the ones in libjpeg-turbo's file (PIL), so both decode to the same picture."""
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=make_onnx(320, 240))
    try:
        rng = np.random.default_rng(shape[0])
        src = test_pics["omar-lopez-T6zu4jFhVwg"]
        pic = np.ascontiguousarray(np.resize(src, (*shape, 3)) if shape[0] > 462 or shape[1] > 640 else src[: shape[0], : shape[1]])
        for frame, q in ((pic, 95), (pic, 50), (rng.integers(0, 256, (*shape, 3), dtype=np.uint8), 95),
                         (np.where(rng.random((*shape, 3)) < 0.5, 0, 255).astype(np.uint8), 100)):
            ours = m.annotate_encode_jpeg(frame, np.zeros((0, 5), np.float32), 1.0, 1.0, quality=q)
            ref = _pil_jpeg(frame, q)
            ia, ca = nn.jpeg_coefficients(ours)
            ib, cb = nn.jpeg_coefficients(ref)
            np.testing.assert_array_equal(ia["quant"], ib["quant"])
            np.testing.assert_array_equal(ca, cb)
            np.testing.assert_array_equal(np.asarray(Image.open(io.BytesIO(ours)).convert("RGB")),
                                          np.asarray(Image.open(io.BytesIO(ref)).convert("RGB")))
    finally:
        m.close()


@pytest.mark.gpu
def test_gpu_annotate_reencode_from_jpeg(make_onnx, test_pics):
    """The production order of inferer.rs:35-39 in one call: JPEG in -> decode (GPU) -> rectangles -> encode -> JPEG out;
    equals decoding with libjpeg-turbo, drawing with the oracle and encoding with libjpeg-turbo."""
    path = make_onnx(320, 240, cls_bias=-0.75)
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path)
    try:
        b = io.BytesIO()
        Image.fromarray(test_pics["clarke-sanders-ybPJ47PMT_M"]).save(b, "JPEG", quality=90, subsampling=1)
        jpeg_in = b.getvalue()
        pixels = np.ascontiguousarray(np.asarray(Image.open(io.BytesIO(jpeg_in)).convert("RGB")))
        dets, counts = m.run_batch_jpeg([jpeg_in], cap=64)
        boxes = dets[0] if len(dets[0]) else _boxes()
        ours = m.annotate_encode_jpeg(jpeg_in, boxes, 640.0, 427.0, quality=95)
        ref = _pil_jpeg(odraw.draw_boxes(pixels, boxes, 640.0, 427.0), 95)
        np.testing.assert_array_equal(nn.jpeg_coefficients(ours)[1], nn.jpeg_coefficients(ref)[1])
        np.testing.assert_array_equal(np.asarray(Image.open(io.BytesIO(ours)).convert("RGB")),
                                      np.asarray(Image.open(io.BytesIO(ref)).convert("RGB")))
    finally:
        m.close()


# ---- text overlay (inferer.rs:80-88) ----
def _synthetic_atlas(seed=0, charset="0123456789.%", max_len=7):
    """Random glyph boxes and coverage in the shape a rusttype atlas has: 8 px advance, boxes that touch or overlap their
    neighbours by a column, coverage in [0, 1] with plenty of exact 0 and 1."""
    rng = np.random.default_rng(seed)
    glyphs, cov = [], []
    for pos in range(max_len):
        row = []
        for _ in charset:
            w, h = int(rng.integers(0, 11)), int(rng.integers(1, 14))
            x0, y0 = pos * 8 + int(rng.integers(-1, 3)), int(rng.integers(0, 6))
            v = rng.random(w * h).astype(np.float32)
            v[rng.random(w * h) < 0.2] = 0.0
            v[rng.random(w * h) < 0.2] = 1.0
            row.append((x0, y0, w, h, sum(len(c) for c in cov)))
            cov.append(v)
        glyphs.append(row)
    return charset, max_len, glyphs, np.concatenate(cov)


def test_confidence_text_is_rusts_two_decimal_format():
    """format!("{:.2}%", confidence * 100.0) on f32: the f32 product, its exact value rounded to two decimals."""
    rng = np.random.default_rng(0)
    for c in list(rng.random(200).astype(np.float32)) + [np.float32(v) for v in (0.5, 0.999999, 1.0, 0.70125, 0.9, 0.12345678)]:
        assert nn.confidence_text(c) == odraw.confidence_text(c)
    assert nn.confidence_text(0.5) == "50.00%" and nn.confidence_text(1.0) == "100.00%"
    assert nn.confidence_text(np.float32(0.9)) == "90.00%"      # 0.9f32 * 100f32 rounds to exactly 90
    assert nn.confidence_text(np.float32(0.97531)) == "97.53%"


@pytest.mark.gpu
def test_gpu_text_overlay_matches_the_oracle(make_onnx, test_pics):
    """Rectangles and text in the reference's order, glyphs blended as imageproc's draw_text_mut does, clipped at the frame's
    edges, over a random atlas (the real one comes from rusttype through the binding)."""
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=make_onnx(320, 240))
    try:
        atlas = _synthetic_atlas(3)
        m.text_atlas_set(*atlas)
        boxes = np.concatenate([_boxes(), np.float32([[0.11, 0.21, 0.33, 0.5, 0.6251],     # overlaps the first one's text
                                                      [0.93, 0.01, 0.99, 0.2, 0.87654],     # text runs off the right edge
                                                      [0.4, 0.985, 0.6, 0.999, 1.0],        # ... and off the bottom; 7 characters
                                                      [-0.01, -0.01, 0.2, 0.1, 0.5]])])     # origin above / left of the frame
        for pic, (sw, sh) in ((test_pics["omar-lopez-T6zu4jFhVwg"], (640.0, 462.0)), (test_pics["bruce-mars-ZXq7xoo98b0"], (1280.0, 720.0)),
                              (test_pics["michael-dam-mEZ3PoFGs_k"][:100, :90], (90.0, 100.0))):
            got = m.draw_boxes(pic, boxes, sw, sh)
            want = odraw.draw_boxes(pic, boxes, sw, sh, atlas)
            np.testing.assert_array_equal(got, want)
            assert (want != odraw.draw_boxes(pic, boxes, sw, sh)).any()  # the text shows
        # the encoded file carries the text too
        pic = test_pics["omar-lopez-T6zu4jFhVwg"]
        ours = m.annotate_encode_jpeg(pic, boxes, 640.0, 462.0, quality=95)
        ref = _pil_jpeg(odraw.draw_boxes(pic, boxes, 640.0, 462.0, atlas), 95)
        np.testing.assert_array_equal(nn.jpeg_coefficients(ours)[1], nn.jpeg_coefficients(ref)[1])
        m.text_atlas_set("", 0, [], [])  # removed: rectangles only again
        np.testing.assert_array_equal(m.draw_boxes(pic, boxes, 640.0, 462.0), odraw.draw_boxes(pic, boxes, 640.0, 462.0))
        with pytest.raises(nn.UltrafaceError):
            m.text_atlas_set("01", 1, [[(0, 0, 4, 4, 0), (0, 0, 4, 4, 10)]], np.zeros(20, np.float32))  # second glyph overruns
    finally:
        m.close()


@pytest.mark.gpu
def test_gpu_batch_reencode_writes_the_single_frame_files(make_onnx, test_pics):
    """uf_annotate_reencode_batch_jpeg (everything on the GPU incl. Huffman decoding AND coding, byte stuffing) must write, byte
    for byte, the files the per-frame call writes with the host Huffman coder — mixed sizes and samplings in one batch, odd
    sizes (dummy blocks), a restart-interval frame and a truncated frame (both decoded by the host decoder), more frames than
    one chunk, with and without the text overlay."""
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=make_onnx(320, 240), max_batch=64)
    try:
        rng = np.random.default_rng(11)
        pics = list(test_pics.values())

        def enc(a, q, ss, **kw):
            b = io.BytesIO()
            Image.fromarray(a).save(b, "JPEG", quality=q, subsampling=ss, **kw)
            return b.getvalue()
        jpegs = [enc(rng.integers(0, 256, (480, 640, 3), dtype=np.uint8), 92, 1) for _ in range(5)]          # dense: many 0xFF bytes
        jpegs += [enc(p, 85, 2) for p in pics] + [enc(pics[0][:301, :333], 80, 1), enc(pics[1][:9, :17], 90, 0), enc(pics[2][:100, :90], 75, 2)]
        jpegs += [enc(np.full((240, 320, 3), 128, np.uint8), 95, 2)]                                          # flat: one DC, all EOB
        jpegs += [jpegs[6][: len(jpegs[6]) * 2 // 3] + b"\xff\xd9"]                                           # truncated
        try:
            import cv2
            ok, buf = cv2.imencode(".jpg", pics[3][:, :, ::-1], [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, 5])
            jpegs.append(buf.tobytes())
        except ImportError:
            pass
        jpegs = jpegs * 3                                                                                     # > 32 frames: several chunks
        dets = [_boxes()[: (i % 5) + (i % 2)] for i in range(len(jpegs))]                                     # incl. frames without detections
        for atlas in (None, _synthetic_atlas(7)):
            if atlas:
                m.text_atlas_set(*atlas)
            got = m.annotate_reencode_batch_jpeg(jpegs, dets, 640.0, 480.0, quality=95)
            for i, (j, d) in enumerate(zip(jpegs, dets)):
                want = m.annotate_encode_jpeg(j, d, 640.0, 480.0, quality=95)
                assert got[i] == want, (i, len(got[i]), len(want), atlas is not None)
        low = m.annotate_reencode_batch_jpeg(jpegs[:3], dets[:3], 640.0, 480.0, quality=30)
        assert low[0] == m.annotate_encode_jpeg(jpegs[0], dets[0], 640.0, 480.0, quality=30)
        with pytest.raises(nn.UltrafaceError) as e:
            m.annotate_reencode_batch_jpeg(jpegs[:4], dets[:4], 640.0, 480.0, out_stride=2048)
        assert e.value.code == 7  # UF_ERR_CAPACITY
    finally:
        m.close()


@pytest.mark.gpu
def test_gpu_worker_batch_is_decode_detect_draw_encode(make_onnx, test_pics):
    """uf_worker_batch_jpeg = the body of the reference's worker loop (inferer.rs:35-46) for a batch: its detections are
    run_batch_jpeg's, its files are annotate_encode_jpeg's for those detections."""
    path = make_onnx(320, 240, cls_bias=-0.75)
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=64)
    try:
        rng = np.random.default_rng(12)
        pics = list(test_pics.values())

        def enc(a, q, ss):
            b = io.BytesIO()
            Image.fromarray(a).save(b, "JPEG", quality=q, subsampling=ss)
            return b.getvalue()
        jpegs = [enc(rng.integers(0, 256, (480, 640, 3), dtype=np.uint8), 90, 1) for _ in range(20)] + [enc(p, 85, 2) for p in pics]
        jpegs += [enc(pics[0][:240, :320], 90, 0), enc(pics[1][:301, :333], 80, 1)] + [enc(rng.integers(0, 256, (480, 640, 3), dtype=np.uint8), 90, 1) for _ in range(14)]
        m.text_atlas_set(*_synthetic_atlas(9))
        want_d, want_c = m.run_batch_jpeg(jpegs, cap=64)
        dets, counts, files = m.worker_batch_jpeg(jpegs, 1280.0, 720.0, quality=95, cap=64)
        assert counts == want_c and sum(counts) > 0
        for i in range(len(jpegs)):
            np.testing.assert_array_equal(dets[i], want_d[i])
            assert files[i] == m.annotate_encode_jpeg(jpegs[i], want_d[i], 1280.0, 720.0, quality=95), i
    finally:
        m.close()


@pytest.mark.gpu
def test_gpu_worker_calls_in_flight_do_not_share_scratch(make_onnx, test_pics):
    """Several host threads on one handle (one lane each): every call's detections and files are the single-threaded ones."""
    import threading

    path = make_onnx(320, 240, cls_bias=-0.75)
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=32, lanes=4)
    try:
        rng = np.random.default_rng(13)
        m.text_atlas_set(*_synthetic_atlas(4))

        def enc(a, q, ss):
            b = io.BytesIO()
            Image.fromarray(a).save(b, "JPEG", quality=q, subsampling=ss)
            return b.getvalue()
        sets = [[enc(rng.integers(0, 256, (480, 640, 3), dtype=np.uint8), 85 + t, 1) for _ in range(6 + 3 * t)] +
                [enc(p, 80 + t, 2) for p in list(test_pics.values())[t:t + 3]] for t in range(4)]
        want = [m.worker_batch_jpeg(js, 1280.0, 720.0, quality=95, cap=64) for js in sets]
        got = [None] * 4
        errs = []

        def work(t):
            try:
                for _ in range(6):
                    got[t] = m.worker_batch_jpeg(sets[t], 1280.0, 720.0, quality=95, cap=64)
                    assert got[t][1] == want[t][1] and got[t][2] == want[t][2]
            except Exception as e:  # noqa: BLE001
                errs.append((t, repr(e)))
        ts = [threading.Thread(target=work, args=(t,)) for t in range(4)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not errs, errs
    finally:
        m.close()


@pytest.fixture(scope="module")
def henc_sim_lib(tmp_path_factory):
    import ctypes as C
    import pathlib
    import subprocess

    root = pathlib.Path(__file__).resolve().parents[1]
    so = tmp_path_factory.mktemp("henc") / "henc_sim.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f"-I{root / 'include'}", "-o", str(so),
                    str(root / "tests/helpers/henc_sim.cc"), str(root / "infercam_onnx_b200/csrc/jpeg_encode.cc")], check=True)
    lib = C.CDLL(str(so))
    lib.henc_sim.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    return lib


@pytest.mark.parametrize("shape,quality", [((480, 640), 95), ((462, 640), 95), ((301, 333), 80), ((9, 17), 95), ((16, 8), 50), ((240, 320), 100)])
def test_device_huffman_coder_plan_on_the_host(test_pics, henc_sim_lib, shape, quality):
    """The device coder's plan — bits per block, prefix sum, bits OR-ed into a bit buffer at their offsets, padding, byte
    stuffing — walked serially on the CPU by tests/helpers/henc_sim.cc over the SAME per-block code the kernels compile
    (csrc/jpeg_henc_core.h: block source incl. libjpeg's dummy blocks, DC predictor, the bits of a coefficient), must write
    the file the sequential writer writes (itself checked against libjpeg-turbo above)."""
    import ctypes as C

    lib = henc_sim_lib
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    base = test_pics["omar-lopez-T6zu4jFhVwg"]
    pic = (np.resize(base, (h, w, 3)) if h > base.shape[0] or w > base.shape[1] else base[:h, :w]).astype(np.int16)
    for noise in (0, 40):  # the photo; the photo under heavy noise (long codes, ZRL runs, many 0xFF bytes)
        arr = (pic + rng.integers(-noise, noise + 1, pic.shape)).clip(0, 255).astype(np.uint8)
        info, coefs = nn.jpeg_coefficients(_pil_jpeg(arr, quality))
        planes = _to_plane_raster(info, coefs)
        want = nn.jpeg_write_coefficients(w, h, quality, planes)
        out = np.empty(len(want) + 4096, np.uint8)
        n = C.c_size_t()
        rc = lib.henc_sim(w, h, quality, planes.ctypes.data, out.ctypes.data, out.size, C.byref(n))
        assert rc == 0 and bytes(out[: n.value]) == want, (rc, n.value, len(want))
