"""Host-only tests of the ingest helpers behind the C ABI: the stream key `hashed(&id)`
(infer_server/src/lib.rs:39-46) and the bincode `ProtoMsg` wire message (common/src/protocol.rs)."""
import ctypes as C

import numpy as np
import pytest

from infercam_onnx_b200 import _capi, batcher, nn
from oracle import ingest


def _c_siphash(c, d, k0, k1, data):
    out = C.c_uint64()
    assert _capi.load().uf_debug_siphash(c, d, k0, k1, bytes(data), len(data), C.byref(out)) == 0
    return out.value


def test_siphash_published_vectors():
    """SipHash-2-4 test vectors of the reference implementation (key 00..0f, message 00..len-1): pins the generic c/d
    implementation in the library and in the oracle; uf_stream_hash is the same code with c = 1, d = 3."""
    k0, k1 = int.from_bytes(bytes(range(8)), "little"), int.from_bytes(bytes(range(8, 16)), "little")
    vectors = {0: 0x726FDB47DD0E0E31, 1: 0x74F839C593DC67FD, 2: 0x0D6C8009D9A94F5A, 3: 0x85676696D7FB7E2D,
               7: 0xAB0200F58B01D137, 8: 0x93F5F5799A932462, 15: 0xA129CA6149BE45E5}
    for n, want in vectors.items():
        msg = bytes(range(n))
        assert ingest.siphash(2, 4, k0, k1, msg) == want, n
        assert _c_siphash(2, 4, k0, k1, msg) == want, n


def test_stream_hash_matches_oracle_restatement():
    rng = np.random.default_rng(0)
    names = ["", "a", "cam0", "bla", "webcam-living-room", "x" * 7, "y" * 8, "z" * 9, "ünïcode-名前"] + \
            ["".join(chr(rng.integers(32, 127)) for _ in range(int(n))) for n in rng.integers(1, 40, 50)]
    for n in names:
        assert batcher.stream_hash(n) == ingest.hashed(n), n
    assert len({batcher.stream_hash(n) for n in names}) == len(set(names))
    # the 0xff terminator of `impl Hash for str` is part of the hash
    assert batcher.stream_hash("ab") != _c_siphash(1, 3, 0, 0, b"ab")
    assert batcher.stream_hash("ab") == _c_siphash(1, 3, 0, 0, b"ab\xff")


def test_protomsg_golden_vector_and_round_trip():
    """protocol.rs:36-50: FrameMsg{id: "bla", data: [1,2,3]} — bincode 1.3 = u64 LE lengths, raw bytes."""
    golden = bytes([3, 0, 0, 0, 0, 0, 0, 0]) + b"bla" + bytes([3, 0, 0, 0, 0, 0, 0, 0, 1, 2, 3])
    assert ingest.framemsg_bytes("bla", bytes([1, 2, 3])) == golden
    msg = ingest.protomsg_frame("bla", bytes([1, 2, 3]))
    assert msg == bytes([1, 0, 0, 0]) + golden
    assert batcher.protomsg_parse(msg) == ("FrameMsg", "bla", bytes([1, 2, 3]))
    assert batcher.protomsg_parse(ingest.protomsg_connect("cam7")) == ("ConnectReq", "cam7", b"")
    rng = np.random.default_rng(1)
    for n in (0, 1, 1000, 70000):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        m = ingest.protomsg_frame("stream-%d" % n, data)
        assert batcher.protomsg_parse(m) == ingest.protomsg_parse(m) == ("FrameMsg", "stream-%d" % n, data)


@pytest.mark.parametrize("bad", [b"", b"\x01\x00\x00", b"\x02\x00\x00\x00" + b"\x00" * 8,            # truncated / unknown variant
                                 b"\x01\x00\x00\x00" + (1 << 40).to_bytes(8, "little") + b"abc",        # length beyond the message
                                 b"\x01\x00\x00\x00" + (2).to_bytes(8, "little") + b"ab" + (9).to_bytes(8, "little") + b"x",
                                 b"\x00\x00\x00\x00" + (1).to_bytes(8, "little") + b"a" + b"trailing"])
def test_protomsg_malformed_is_an_error_not_a_crash(bad):
    with pytest.raises(nn.UltrafaceError) as e:
        batcher.protomsg_parse(bad)
    assert e.value.code == 1
