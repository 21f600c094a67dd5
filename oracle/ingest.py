"""CPU restatement of the stream key and the wire message on either side of the hot path.

TEST INFRASTRUCTURE — see oracle/__init__.py.

* `hashed` — infer_server/src/lib.rs:39-46: `DefaultHasher::new()` + `data.hash(&mut hasher)` + `finish()`. For the
  `&String` the router passes (router.rs:58) `impl Hash for str` writes the UTF-8 bytes followed by one 0xff byte;
  `DefaultHasher` is SipHash-1-3 with both keys zero (std, `SipHasher13::new_with_keys(0, 0)`). Restated from the
  published SipHash algorithm (Aumasson & Bernstein); the generic c/d implementation below reproduces the SipHash-2-4
  vectors of the reference implementation (tests/test_ingest.py), 1-3 only changes the round counts.
  PARITY STATUS: pinned to the published algorithm; not compared with a Rust build (no toolchain here).
* `protomsg_*` — common/src/protocol.rs:7-28 through bincode 1.3 (`bincode::serialize` / `deserialize`: little-endian,
  fixed-width integers, u32 enum variant index, u64 length prefixes); the reference's own round-trip test
  (protocol.rs:36-50: FrameMsg{id: "bla", data: [1,2,3]}) is the golden vector.
"""
from __future__ import annotations

import struct

M64 = (1 << 64) - 1


def _rotl(x: int, b: int) -> int:
    return ((x << b) | (x >> (64 - b))) & M64


def siphash(c: int, d: int, k0: int, k1: int, data: bytes) -> int:
    v0, v1, v2, v3 = k0 ^ 0x736F6D6570736575, k1 ^ 0x646F72616E646F6D, k0 ^ 0x6C7967656E657261, k1 ^ 0x7465646279746573

    def rnd(v0, v1, v2, v3):
        v0 = (v0 + v1) & M64; v1 = _rotl(v1, 13); v1 ^= v0; v0 = _rotl(v0, 32)
        v2 = (v2 + v3) & M64; v3 = _rotl(v3, 16); v3 ^= v2
        v0 = (v0 + v3) & M64; v3 = _rotl(v3, 21); v3 ^= v0
        v2 = (v2 + v1) & M64; v1 = _rotl(v1, 17); v1 ^= v2; v2 = _rotl(v2, 32)
        return v0, v1, v2, v3
    full = len(data) // 8 * 8
    for i in range(0, full, 8):
        m = int.from_bytes(data[i:i + 8], "little")
        v3 ^= m
        for _ in range(c):
            v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
        v0 ^= m
    last = int.from_bytes(data[full:], "little") | ((len(data) & 0xFF) << 56)
    v3 ^= last
    for _ in range(c):
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    v0 ^= last
    v2 ^= 0xFF
    for _ in range(d):
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    return v0 ^ v1 ^ v2 ^ v3


def hashed(name: str) -> int:
    """lib.rs:39-46 applied to `&String` (router.rs:58)."""
    return siphash(1, 3, 0, 0, name.encode() + b"\xff")


def protomsg_frame(stream_id: str, data: bytes) -> bytes:
    """bincode::serialize(&ProtoMsg::FrameMsg(FrameMsg{id, data})) — socket_sender.rs:80-85."""
    sid = stream_id.encode()
    return struct.pack("<I", 1) + struct.pack("<Q", len(sid)) + sid + struct.pack("<Q", len(data)) + bytes(data)


def protomsg_connect(name: str) -> bytes:
    raw = name.encode()
    return struct.pack("<I", 0) + struct.pack("<Q", len(raw)) + raw


def framemsg_bytes(stream_id: str, data: bytes) -> bytes:
    """bincode::serialize(&FrameMsg{..}) (the struct alone, as in the reference's unit test protocol.rs:36-50)."""
    return protomsg_frame(stream_id, data)[4:]


def protomsg_parse(msg: bytes):
    """ProtoMsg::deserialize (protocol.rs:24-28)."""
    (variant,) = struct.unpack_from("<I", msg, 0)
    pos = 4
    (n,) = struct.unpack_from("<Q", msg, pos)
    pos += 8
    sid = msg[pos:pos + n].decode()
    if len(msg) < pos + n:
        raise ValueError("truncated")
    pos += n
    if variant == 0:
        return "ConnectReq", sid, b""
    if variant != 1:
        raise ValueError("unknown variant")
    (m,) = struct.unpack_from("<Q", msg, pos)
    pos += 8
    if len(msg) < pos + m:
        raise ValueError("truncated")
    return "FrameMsg", sid, msg[pos:pos + m]
