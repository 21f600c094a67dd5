"""Minimal ONNX protobuf reader for the oracle (TEST INFRASTRUCTURE — see oracle/__init__.py).

The `onnx` Python package is not installed in this image; this decodes the wire
format directly (onnx.proto field numbers). It is deliberately independent of
the product's C++ loader (infercam_onnx_b200/csrc/onnx_graph.cc) and of the
fixture writer, so a bug in either shows up as a parity failure.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np


def _read_varint(buf: bytes, pos: int) -> tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: bytes):
    """Yield (field_number, wire_type, value) for one message."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        fnum, wire = key >> 3, key & 7
        if wire == 0:
            v, pos = _read_varint(buf, pos)
        elif wire == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wire == 2:
            ln, pos = _read_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wire == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wire}")
        yield fnum, wire, v


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_or_single_ints(wire: int, v) -> list[int]:
    if wire == 0:
        return [_signed(v)]
    out, pos = [], 0
    while pos < len(v):
        x, pos = _read_varint(v, pos)
        out.append(_signed(x))
    return out


_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64, 9: np.bool_, 11: np.float64}


def parse_tensor(buf: bytes) -> tuple[str, np.ndarray]:
    dims: list[int] = []
    dtype = 1
    name = ""
    raw = None
    floats: list[float] = []
    int32s: list[int] = []
    int64s: list[int] = []
    for f, w, v in _fields(buf):
        if f == 1:
            dims += _packed_or_single_ints(w, v)
        elif f == 2:
            dtype = v
        elif f == 4:
            if w == 5:
                floats.append(struct.unpack("<f", v)[0])
            else:
                floats += list(np.frombuffer(v, "<f4"))
        elif f == 5:
            int32s += _packed_or_single_ints(w, v)
        elif f == 7:
            int64s += _packed_or_single_ints(w, v)
        elif f == 8:
            name = bytes(v).decode()
        elif f == 9:
            raw = bytes(v)
    np_dt = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(np_dt).newbyteorder("<")).astype(np_dt)
    elif dtype == 1:
        arr = np.asarray(floats, np.float32)
    elif dtype == 7:
        arr = np.asarray(int64s, np.int64)
    else:
        arr = np.asarray(int32s).astype(np_dt)
    return name, arr.reshape(dims)


@dataclass
class Node:
    op: str
    inputs: list[str]
    outputs: list[str]
    name: str = ""
    attrs: dict = field(default_factory=dict)


def _parse_attr(buf: bytes):
    name = ""
    val = {}
    ints: list[int] = []
    floats: list[float] = []
    atype = 0
    for f, w, v in _fields(buf):
        if f == 1:
            name = bytes(v).decode()
        elif f == 2:
            val["f"] = struct.unpack("<f", v)[0]
        elif f == 3:
            val["i"] = _signed(v)
        elif f == 4:
            val["s"] = bytes(v)
        elif f == 5:
            val["t"] = parse_tensor(v)[1]
        elif f == 7:
            if w == 5:
                floats.append(struct.unpack("<f", v)[0])
            else:
                floats += list(np.frombuffer(v, "<f4"))
        elif f == 8:
            ints += _packed_or_single_ints(w, v)
        elif f == 20:
            atype = v
    if atype == 7 or (ints and atype == 0):
        return name, ints
    if atype == 6 or (floats and atype == 0):
        return name, floats
    for k in ("t", "i", "f", "s"):
        if k in val:
            return name, val[k]
    return name, ints  # empty INTS


def _parse_node(buf: bytes) -> Node:
    n = Node("", [], [])
    for f, w, v in _fields(buf):
        if f == 1:
            n.inputs.append(bytes(v).decode())
        elif f == 2:
            n.outputs.append(bytes(v).decode())
        elif f == 3:
            n.name = bytes(v).decode()
        elif f == 4:
            n.op = bytes(v).decode()
        elif f == 5:
            k, a = _parse_attr(v)
            n.attrs[k] = a
    return n


def _parse_value_info(buf: bytes) -> tuple[str, list[int]]:
    name, shape = "", []
    for f, w, v in _fields(buf):
        if f == 1:
            name = bytes(v).decode()
        elif f == 2:
            for f2, _, v2 in _fields(v):
                if f2 != 1:
                    continue
                for f3, _, v3 in _fields(v2):
                    if f3 != 2:
                        continue
                    for f4, _, v4 in _fields(v3):
                        if f4 != 1:
                            continue
                        d = -1
                        for f5, _, v5 in _fields(v4):
                            if f5 == 1:
                                d = _signed(v5)
                        shape.append(d)
    return name, shape


@dataclass
class Graph:
    nodes: list[Node]
    initializers: dict[str, np.ndarray]
    inputs: list[tuple[str, list[int]]]
    outputs: list[tuple[str, list[int]]]
    opset: int = 9


def load_onnx(path_or_bytes) -> Graph:
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    graph_buf = None
    opset = 9
    for f, w, v in _fields(data):
        if f == 7:
            graph_buf = v
        elif f == 8:
            dom, ver = "", 0
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    dom = bytes(v2).decode()
                elif f2 == 2:
                    ver = v2
            if dom in ("", "ai.onnx"):
                opset = ver
    if graph_buf is None:
        raise ValueError("no graph in model")
    g = Graph([], {}, [], [], opset)
    for f, w, v in _fields(graph_buf):
        if f == 1:
            g.nodes.append(_parse_node(v))
        elif f == 5:
            name, arr = parse_tensor(v)
            g.initializers[name] = arr
        elif f == 11:
            g.inputs.append(_parse_value_info(v))
        elif f == 12:
            g.outputs.append(_parse_value_info(v))
    g.inputs = [i for i in g.inputs if i[0] not in g.initializers]
    return g
