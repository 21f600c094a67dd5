"""N2: JPEG in front of the path. The reference decodes with libjpeg-turbo (`turbojpeg::decompress_image`,
infer_server/src/inferer.rs:35); libjpeg-turbo itself is in this image behind PIL and OpenCV, so here the REFERENCE'S OWN
decoder is the checker: host Huffman decoding (product, csrc/jpeg_entropy.cc) + the oracle's restatement of the sample-domain
half (oracle/jpeg_oracle.c) must reproduce its pixels bit for bit on CPU, and the GPU kernels (csrc/kernels_jpeg.cu) must do
the same on the B200."""
import io

import numpy as np
import pytest
from PIL import Image

from infercam_onnx_b200 import nn
from oracle import jpeg as ojpeg


def _enc(arr, quality=90, subsampling=2, **kw):
    b = io.BytesIO()
    Image.fromarray(arr).save(b, "JPEG", quality=quality, subsampling=subsampling, **kw)
    return b.getvalue()


def _turbo(data):
    """libjpeg-turbo's pixels (default settings: ISLOW IDCT, fancy upsampling — what tjDecompress2 with flags 0 runs)."""
    return np.ascontiguousarray(np.asarray(Image.open(io.BytesIO(data)).convert("RGB")))


def _cases(test_pics):
    rng = np.random.default_rng(0)
    pic = test_pics["omar-lopez-T6zu4jFhVwg"]
    tall = test_pics["michael-dam-mEZ3PoFGs_k"]
    out = []
    for ss in (0, 1, 2):  # 4:4:4, 4:2:2 (what MJPG webcams send), 4:2:0
        for q in (35, 90, 100):
            out.append((f"photo ss{ss} q{q}", _enc(pic, q, ss)))
    out.append(("noise 4:2:0", _enc(rng.integers(0, 256, (480, 640, 3), dtype=np.uint8), 95, 2)))
    out.append(("saturated", _enc(np.where(rng.random((240, 320, 3)) < 0.5, 0, 255).astype(np.uint8), 100, 2)))
    out.append(("odd 333x301 4:2:0", _enc(pic[:301, :333], 85, 2)))
    out.append(("odd 333x301 4:2:2", _enc(pic[:301, :333], 85, 1)))
    out.append(("tiny 17x9", _enc(pic[:9, :17], 85, 2)))
    out.append(("tall 640x960", _enc(tall, 80, 2)))
    out.append(("optimised huffman", _enc(pic, 80, 2, optimize=True)))
    g = io.BytesIO()
    Image.fromarray(pic).convert("L").save(g, "JPEG", quality=90)
    out.append(("grey", g.getvalue()))
    try:
        import cv2
        ok, buf = cv2.imencode(".jpg", pic[:, :, ::-1], [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, 7])
        if ok:
            out.append(("restart interval 7", buf.tobytes()))
    except ImportError:
        pass
    return out


def test_huffman_decoder_and_oracle_reproduce_libjpeg_turbo(test_pics):
    for name, data in _cases(test_pics):
        info, coefs = nn.jpeg_coefficients(data)
        assert info["nonzero"] == int((coefs != 0).sum()), name
        got = ojpeg.reconstruct(coefs, info["w"], info["h"], info["hs"], info["vs"], info["quant"])
        np.testing.assert_array_equal(got, _turbo(data), err_msg=name)


def test_mjpeg_without_huffman_tables_uses_annex_k(test_pics):
    """Motion-JPEG frames from V4L2 webcams carry no DHT segment; the Annex K tables apply (PIL writes exactly those, so
    stripping the DHT segments of a PIL file must not change the result)."""
    data = _enc(test_pics["omar-lopez-T6zu4jFhVwg"][:240, :320], 85, 1)
    out, i = bytearray(data[:2]), 2
    while data[i + 1] != 0xDA:
        seg = 2 + int.from_bytes(data[i + 2:i + 4], "big")
        if data[i + 1] != 0xC4:
            out += data[i:i + seg]
        i += seg
    out += data[i:]
    assert len(out) < len(data) - 400
    a, ca = nn.jpeg_coefficients(data)
    b, cb = nn.jpeg_coefficients(bytes(out))
    np.testing.assert_array_equal(ca, cb)


def test_jpeg_errors_are_loud(test_pics):
    pic = test_pics["omar-lopez-T6zu4jFhVwg"][:64, :64]
    prog = _enc(pic, 90, 2, progressive=True)
    with pytest.raises(nn.UltrafaceError) as e:
        nn.jpeg_info(prog) and nn.jpeg_coefficients(prog)
    assert e.value.code == 4 and "progressive" in str(e.value)
    for bad in (b"", b"\x89PNG\r\n\x1a\n", b"\xff\xd8\xff\xd9", _enc(pic)[:40]):
        with pytest.raises(nn.UltrafaceError) as e:
            nn.jpeg_coefficients(bad)
        assert e.value.code in (1, 4)
    # a truncated entropy segment decodes (missing data reads as zero bits, as in libjpeg) instead of failing
    data = _enc(test_pics["omar-lopez-T6zu4jFhVwg"][:240, :320], 90, 2)
    info, coefs = nn.jpeg_coefficients(data[: len(data) * 2 // 3])
    assert info["w"] == 320 and coefs.shape == (info["nblocks"], 64)
    full = nn.jpeg_coefficients(data)[1]
    assert (coefs[: info["nblocks"] // 2] == full[: info["nblocks"] // 2]).all() and not coefs[-8:].any()
    # ... and gives the picture libjpeg-turbo gives for the same truncated file (grey below the point where the data ends)
    from PIL import ImageFile
    ImageFile.LOAD_TRUNCATED_IMAGES = True
    try:
        ref = np.asarray(Image.open(io.BytesIO(data[: len(data) * 2 // 3])).convert("RGB"))
    finally:
        ImageFile.LOAD_TRUNCATED_IMAGES = False
    got = ojpeg.reconstruct(coefs, info["w"], info["h"], info["hs"], info["vs"], info["quant"])
    assert (got != ref).any(2).mean() < 0.02  # identical but for the MCU row where PIL's streaming stops feeding data
    # hostile garbage after valid headers must not crash or hang
    rng = np.random.default_rng(1)
    sos = data.index(b"\xff\xda")
    for _ in range(20):
        junk = data[: sos + 14] + rng.integers(0, 256, 4000, dtype=np.uint8).tobytes()
        nn.jpeg_coefficients(junk)


@pytest.mark.gpu
def test_gpu_jpeg_decode_is_bit_exact_with_libjpeg_turbo(make_onnx, test_pics):
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=make_onnx(320, 240))
    try:
        for name, data in _cases(test_pics):
            np.testing.assert_array_equal(m.jpeg_decode_rgb(data), _turbo(data), err_msg=name)
    finally:
        m.close()


@pytest.mark.gpu
def test_gpu_jpeg_batch_equals_rgb_batch_of_the_same_pixels(make_onnx, test_pics):
    """uf_infer_batch_jpeg = uf_infer_batch on libjpeg-turbo's pixels: identical raw tensors and detections, for frames of
    mixed sizes / samplings in one batch, including 640x480 frames (decoded straight into the fused resize + stem kernel)
    and network-sized frames (decoded straight into the resized buffer)."""
    path = make_onnx(320, 240, cls_bias=-0.75)
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=48, host_chunk=8)
    try:
        rng = np.random.default_rng(3)
        pics = list(test_pics.values())
        jpegs = [_enc(rng.integers(0, 256, (480, 640, 3), dtype=np.uint8), 92, 1) for _ in range(10)]   # webcam-like: one whole stage
        jpegs += [_enc(p, 88, 2) for p in pics] + [_enc(pics[0][:240, :320], 90, 0), _enc(pics[1][:301, :333], 70, 1)]
        jpegs += [_enc(p, 95, 1) for p in pics[:3]]
        rgbs = [_turbo(j) for j in jpegs]
        dj, cj = m.run_batch_jpeg(jpegs, cap=256)
        sj, bj = m.raw_outputs(0, len(jpegs))
        dr, cr = m.run_batch(rgbs, cap=256)
        sr, br = m.raw_outputs(0, len(jpegs))
        np.testing.assert_array_equal(sj, sr)
        np.testing.assert_array_equal(bj, br)
        assert cj == cr
        for a, b in zip(dj, dr):
            np.testing.assert_array_equal(a, b)
        with pytest.raises(nn.UltrafaceError) as e:
            m.run_batch_jpeg(jpegs[:3] + [_enc(pics[0], progressive=True)])
        assert e.value.code == 4 and "frame 3" in str(e.value)
        again, ca = m.run_batch_jpeg(jpegs[:5], cap=256)  # the handle is fine after the refused batch
        assert ca == cj[:5]
    finally:
        m.close()


# ---- Huffman decoding on the device (kernels_jpeg_huff.cu) ----
def _damaged(data, rng, n):
    """n variants of one file: a flipped bit, a cut with EOI appended, three overwritten bytes — all inside the scan."""
    sos = data.index(b"\xff\xda") + 14
    out = []
    for t in range(n):
        e = bytearray(data)
        pos = int(rng.integers(sos, len(e) - 4))
        if t % 3 == 0:
            e[pos] ^= 1 << int(rng.integers(0, 8))
        elif t % 3 == 1:
            e = e[:pos] + b"\xff\xd9"
        else:
            e[pos:pos + 3] = bytes(rng.integers(0, 256, 3).tolist())
        out.append(bytes(e))
    return out


def test_huffman_synchronisation_rounds_on_the_host(test_pics, tmp_path):
    """The device decoder's round logic (guessed start states, in-CTA fixed point, one CTA of progress per launch, block
    prefix, write pass, DC scan, the status that sends a frame back to the host decoder), stepped serially on the CPU by
    tests/helpers/huff_sim.cc over the SAME symbol decoder the kernels compile (csrc/jpeg_huff_core.h), must reproduce the
    sequential decoder on intact files, and on damaged ones either reproduce it or decline (status != 0)."""
    import ctypes as C
    import pathlib
    import subprocess

    root = pathlib.Path(__file__).resolve().parents[1]
    so = tmp_path / "huff_sim.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f"-I{root / 'include'}", "-o", str(so),
                    str(root / "tests/helpers/huff_sim.cc"), str(root / "infercam_onnx_b200/csrc/jpeg_entropy.cc")], check=True)
    lib = C.CDLL(str(so))
    lib.huff_sim.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int]

    def sim(data, max_rounds=0, bits=0):
        info, ref = nn.jpeg_coefficients(data)
        out = np.zeros_like(ref)
        r, st = C.c_int(), C.c_int()
        rc = lib.huff_sim(data, len(data), out.ctypes.data, out.shape[0], C.byref(r), C.byref(st), max_rounds, bits)
        return rc, st.value, np.array_equal(out, ref), r.value

    for name, data in _cases(test_pics):
        for bits in (0, 128, 256, 2048):  # the engine picks the subsequence length per run
            rc, st, eq, r = sim(data, 0, bits)
            if name.startswith("restart"):
                assert rc == 1, name  # not for the device decoder
            elif bits >= 256 or len(data) < 3 * 4096:
                assert (rc, st, eq) == (0, 0, True), (name, bits, rc, st, eq, r)
            else:  # short subsequences on a long file: four launches may not settle it — then it must say so
                assert rc == 0 and (st & 2 or (st == 0 and eq)), (name, bits, rc, st, eq, r)
    # a frame of several CTAs' worth of data: the capped number of launches settles it; ONE launch leaves the CTAs after the
    # first on guessed start states, which the write pass' fixed-point check must notice (status bit 1) unless they happened
    # to be right
    rng = np.random.default_rng(4)
    big = np.asarray(Image.fromarray(test_pics["omar-lopez-T6zu4jFhVwg"]).resize((1280, 960), Image.BICUBIC)).astype(np.int16)
    big = _enc((big + rng.integers(-8, 8, big.shape)).clip(0, 255).astype(np.uint8), 92, 1)
    assert len(big) > 5 * 8192
    rc, st, eq, r = sim(big, 0, 256)
    assert (rc, st, eq) == (0, 0, True) and r // 1000 == 4, (rc, st, eq, r)
    rc, st, eq, r = sim(big, 1, 256)
    assert rc == 0 and (st & 2 or eq), (rc, st, eq, r)
    assert st & 2  # (with this picture the guesses are not all right)
    rng = np.random.default_rng(5)
    pic = test_pics["omar-lopez-T6zu4jFhVwg"][:240, :320]
    declined = 0
    for ss in (0, 1, 2):
        for e in _damaged(_enc(pic, 75, ss), rng, 60):
            rc, st, eq, _ = sim(e)
            assert rc in (0, 1)
            if rc == 0 and st == 0:
                assert eq
            else:
                declined += 1
    assert declined > 30  # every truncated file at least


@pytest.mark.gpu
def test_gpu_huffman_coefficients_equal_the_sequential_decoder(make_onnx, test_pics):
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=make_onnx(320, 240))
    try:
        rng = np.random.default_rng(2)
        cases = _cases(test_pics)
        big = np.asarray(Image.fromarray(test_pics["omar-lopez-T6zu4jFhVwg"]).resize((1920, 1080), Image.BICUBIC)).astype(np.int16)
        big = (big + rng.integers(-10, 10, big.shape)).clip(0, 255).astype(np.uint8)
        cases.append(("1080p 4:2:0", _enc(big, 90, 2)))
        cases.append(("1080p 4:2:2 q100", _enc(big, 100, 1)))
        for name, data in cases:
            _, ref = nn.jpeg_coefficients(data)
            got, launches = m.jpeg_coefficients_gpu(data)
            np.testing.assert_array_equal(got, ref, err_msg=name)
            assert (launches == 0) == name.startswith("restart"), (name, launches)
        pic = test_pics["omar-lopez-T6zu4jFhVwg"][:240, :320]
        on_device = 0
        for ss in (0, 1, 2):
            for e in _damaged(_enc(pic, 75, ss), rng, 30):
                _, ref = nn.jpeg_coefficients(e)
                got, launches = m.jpeg_coefficients_gpu(e)
                np.testing.assert_array_equal(got, ref)
                on_device += launches > 0
                np.testing.assert_array_equal(m.jpeg_decode_rgb(e), _turbo_lenient(e, m))
        assert on_device > 10  # flipped bits usually still decode to whole blocks: those stay on the device
    finally:
        m.close()


def _turbo_lenient(data, m):
    """Pixels for a damaged file: PIL refuses some of them, so the checker is the host-Huffman path of the same library
    (itself pinned to libjpeg-turbo on intact and truncated files above)."""
    info, coefs = nn.jpeg_coefficients(data)
    return ojpeg.reconstruct(coefs, info["w"], info["h"], info["hs"], info["vs"], info["quant"])


@pytest.mark.gpu
def test_gpu_huffman_batch_equals_host_huffman_batch(make_onnx, test_pics):
    """A mixed batch — sizes, samplings, a restart-interval frame, truncated and bit-flipped frames in the middle — gives the
    same detections whether the entropy decoding runs on the device (default) or on host threads (UF_FLAG_JPEG_HOST_HUFFMAN)."""
    from infercam_onnx_b200 import _capi

    path = make_onnx(320, 240, cls_bias=-0.75)
    rng = np.random.default_rng(8)
    pics = list(test_pics.values())
    jpegs = [_enc(rng.integers(0, 256, (480, 640, 3), dtype=np.uint8), 92, 1) for _ in range(9)]
    jpegs += [_enc(p, 88, 2) for p in pics] + [_enc(pics[1][:301, :333], 70, 1)]
    jpegs[4:4] = _damaged(jpegs[0], rng, 3)          # same size as their neighbours: inside a run
    jpegs += _damaged(_enc(pics[2], 80, 2), rng, 3)
    jpegs += [d for n, d in _cases(test_pics) if n.startswith("restart")]
    res = []
    for flags in (0, _capi.UF_FLAG_JPEG_HOST_HUFFMAN):
        m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=48, host_chunk=8, flags=flags)
        try:
            res.append(m.run_batch_jpeg(jpegs, cap=256))
            again = m.run_batch_jpeg(jpegs[:6], cap=256)   # storage reuse across calls
            assert again[1] == res[-1][1][:6]
        finally:
            m.close()
    assert res[0][1] == res[1][1]
    assert sum(res[0][1]) > 0
    for a, b in zip(res[0][0], res[1][0]):
        np.testing.assert_array_equal(a, b)
