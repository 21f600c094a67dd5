"""Seeded random-init UltraFace ONNX files (numpy only, no `onnx` package needed).

The reference downloads `version-RFB-{320,640}.onnx` at run time
(/root/reference/infer_server/src/nn.rs:21-22,156-162); neither file exists in
this environment and there is no network. BASELINE.json's north_star allows
"random-init weights of the same graph", so this module writes that graph —
the upstream UltraFace RFB / slim topology — as a real
ONNX protobuf. Both the product loader (csrc/onnx_graph.cc) and the oracle
(oracle/ultraface_ref.py) read the *file*, so they share weights without sharing
code.

Op vocabulary matches the published (simplified) export: Conv, Relu, Concat,
Mul, Add, Transpose, Reshape, Softmax, Slice, Exp, Div, Sub, Constant-free
initialisers; `with_bn=True` additionally emits explicit BatchNormalization
nodes to exercise the loader's BN folding.
"""
from __future__ import annotations

import math
import struct
from typing import Iterable, Sequence

import numpy as np

# --------------------------------------------------------------------------
# protobuf wire-format writer (proto2 encoding rules; onnx.proto field numbers)
# --------------------------------------------------------------------------


def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field: int, wire: int) -> bytes:
    return _varint((field << 3) | wire)


def _f_varint(field: int, v: int) -> bytes:
    return _key(field, 0) + _varint(v)


def _f_bytes(field: int, b: bytes) -> bytes:
    return _key(field, 2) + _varint(len(b)) + b


def _f_str(field: int, s: str) -> bytes:
    return _f_bytes(field, s.encode())


def _f_float(field: int, v: float) -> bytes:
    return _key(field, 5) + struct.pack("<f", v)


def tensor_proto(name: str, arr: np.ndarray) -> bytes:
    """TensorProto: dims=1, data_type=2, name=8, raw_data=9."""
    arr = np.asarray(arr)
    arr = arr if arr.ndim == 0 else np.ascontiguousarray(arr)  # (ascontiguousarray would turn a scalar into [1])
    dt = {np.dtype("float32"): 1, np.dtype("int64"): 7}[arr.dtype]
    out = b"".join(_f_varint(1, int(d)) for d in arr.shape)
    out += _f_varint(2, dt)
    out += _f_str(8, name)
    out += _f_bytes(9, arr.tobytes())
    return out


def _attr(name: str, value) -> bytes:
    """AttributeProto: name=1, f=2, i=3, s=4, t=5, floats=7, ints=8, type=20."""
    out = _f_str(1, name)
    if isinstance(value, float):
        out += _f_float(2, value) + _f_varint(20, 1)
    elif isinstance(value, int):
        out += _f_varint(3, value) + _f_varint(20, 2)
    elif isinstance(value, (list, tuple)):
        out += b"".join(_f_varint(8, int(v)) for v in value) + _f_varint(20, 7)
    elif isinstance(value, np.ndarray):
        out += _f_bytes(5, tensor_proto("", value)) + _f_varint(20, 4)
    else:  # pragma: no cover
        raise TypeError(type(value))
    return out


def node_proto(op: str, inputs: Sequence[str], outputs: Sequence[str], name: str = "", **attrs) -> bytes:
    """NodeProto: input=1, output=2, name=3, op_type=4, attribute=5."""
    out = b"".join(_f_str(1, i) for i in inputs)
    out += b"".join(_f_str(2, o) for o in outputs)
    if name:
        out += _f_str(3, name)
    out += _f_str(4, op)
    for k, v in attrs.items():
        out += _f_bytes(5, _attr(k, v))
    return out


def value_info(name: str, shape: Iterable[int]) -> bytes:
    """ValueInfoProto{name=1,type=2{tensor_type=1{elem_type=1,shape=2{dim=1{dim_value=1}}}}}."""
    dims = b"".join(_f_bytes(1, _f_varint(1, int(d))) for d in shape)
    ttype = _f_varint(1, 1) + _f_bytes(2, dims)
    return _f_str(1, name) + _f_bytes(2, _f_bytes(1, ttype))


def model_proto(nodes: list[bytes], inits: list[bytes], inputs: list[bytes], outputs: list[bytes],
                opset: int = 9, graph_name: str = "ultraface") -> bytes:
    graph = b"".join(_f_bytes(1, n) for n in nodes)
    graph += _f_str(2, graph_name)
    graph += b"".join(_f_bytes(5, t) for t in inits)
    graph += b"".join(_f_bytes(11, i) for i in inputs)
    graph += b"".join(_f_bytes(12, o) for o in outputs)
    out = _f_varint(1, 4)  # ir_version
    out += _f_str(2, "tools.onnx_fixture")
    out += _f_bytes(7, graph)
    out += _f_bytes(8, _f_str(1, "") + _f_varint(2, opset))
    return out


# --------------------------------------------------------------------------
# UltraFace graph (upstream vision/nn/mb_tiny_RFB.py, vision/ssd/ssd.py,
# vision/utils/box_utils.py — restated in SURVEY.md §8a-graph)
# --------------------------------------------------------------------------

MIN_BOXES = [[10.0, 16.0, 24.0], [32.0, 48.0], [64.0, 96.0], [128.0, 192.0, 256.0]]
CENTER_VARIANCE = 0.1
SIZE_VARIANCE = 0.2


def feature_map_sizes(width: int, height: int) -> list[tuple[int, int]]:
    """(w, h) of the four SSD source maps: strides 8,16,32,64 with ceil (3x3 s2 pad 1)."""
    def down(v: int, times: int) -> int:
        for _ in range(times):
            v = (v + 1) // 2
        return v
    return [(down(width, t), down(height, t)) for t in (3, 4, 5, 6)]


def generate_priors(width: int, height: int) -> np.ndarray:
    """[K,4] centre-form priors, clamped to [0,1]; order (map, y, x, anchor)."""
    out = []
    for (fw, fh), boxes in zip(feature_map_sizes(width, height), MIN_BOXES):
        # upstream uses scale = image_size / shrinkage with shrinkage = image_size / fm
        # i.e. scale == fm exactly for these sizes.
        for j in range(fh):
            for i in range(fw):
                cx = (i + 0.5) / fw
                cy = (j + 0.5) / fh
                for mb in boxes:
                    out.append([cx, cy, mb / width, mb / height])
    p = np.asarray(out, dtype=np.float32)
    return np.clip(p, 0.0, 1.0)


class _Builder:
    def __init__(self, seed: int, with_bn: bool):
        self.rng = np.random.default_rng(seed)
        self.with_bn = with_bn
        self.nodes: list[bytes] = []
        self.inits: list[bytes] = []
        self.n = 0

    def fresh(self, hint: str = "t") -> str:
        self.n += 1
        return f"{hint}_{self.n}"

    def init(self, arr: np.ndarray, hint: str = "w") -> str:
        name = self.fresh(hint)
        self.inits.append(tensor_proto(name, arr))
        return name

    def conv(self, x: str, cin: int, cout: int, k: int = 1, stride: int = 1, pad: int = 0,
             dil: int = 1, groups: int = 1, relu: bool = True, bn: bool = True,
             bias_shift: np.ndarray | None = None, gain: float = 1.0) -> str:
        """Conv (+BatchNormalization if requested) (+Relu). He-normal weights.

        `bn` mirrors upstream: backbone/RFB convs are conv(no bias)+BN; head and
        extras convs carry a bias and no BN. With with_bn=False the BN is
        pre-folded (identity statistics -> small random bias).
        """
        fan_in = (cin // groups) * k * k
        std = gain * math.sqrt(2.0 / fan_in)
        w = (self.rng.standard_normal((cout, cin // groups, k, k)) * std).astype(np.float32)
        b = (self.rng.standard_normal(cout) * 0.05).astype(np.float32)
        if bias_shift is not None:
            b = (b + bias_shift).astype(np.float32)
        ins = [x, self.init(w, "W")]
        emit_bn = bn and self.with_bn
        if not emit_bn:
            ins.append(self.init(b, "B"))
        y = self.fresh("conv")
        self.nodes.append(node_proto("Conv", ins, [y], dilations=[dil, dil], group=groups,
                                     kernel_shape=[k, k], pads=[pad, pad, pad, pad],
                                     strides=[stride, stride]))
        if emit_bn:
            gamma = (1.0 + 0.1 * self.rng.standard_normal(cout)).astype(np.float32)
            mean = (0.1 * self.rng.standard_normal(cout)).astype(np.float32)
            var = (1.0 + 0.1 * self.rng.random(cout)).astype(np.float32)
            z = self.fresh("bn")
            self.nodes.append(node_proto(
                "BatchNormalization",
                [y, self.init(gamma, "gamma"), self.init(b, "beta"), self.init(mean, "mean"),
                 self.init(var, "var")], [z], epsilon=1e-5, momentum=0.9))
            y = z
        if relu:
            z = self.fresh("relu")
            self.nodes.append(node_proto("Relu", [y], [z]))
            y = z
        return y

    def dw_sep(self, x: str, cin: int, cout: int, stride: int) -> str:
        """upstream conv_dw: dw3x3(s)+BN+ReLU -> 1x1+BN+ReLU."""
        y = self.conv(x, cin, cin, 3, stride, 1, groups=cin)
        return self.conv(y, cin, cout, 1)

    def basic_rfb(self, x: str, cin: int, cout: int, scale: float = 1.0, emit_mul: bool = True) -> str:
        inter = cin // 8
        def branch(mid: Sequence[tuple[int, int]], dil: int) -> str:
            y = self.conv(x, cin, inter, 1, relu=False)
            c = inter
            for co, _ in mid:
                y = self.conv(y, c, co, 3, 1, 1, relu=True)
                c = co
            return self.conv(y, c, 2 * inter, 3, 1, dil, dil=dil, relu=False)
        b0 = branch([(2 * inter, 0)], 2)
        b1 = branch([(2 * inter, 0)], 3)
        b2 = branch([((inter // 2) * 3, 0), (2 * inter, 0)], 5)
        cat = self.fresh("cat")
        self.nodes.append(node_proto("Concat", [b0, b1, b2], [cat], axis=1))
        lin = self.conv(cat, 6 * inter, cout, 1, relu=False)
        short = self.conv(x, cin, cout, 1, relu=False)
        if emit_mul:
            m = self.fresh("mul")
            self.nodes.append(node_proto("Mul", [lin, self.init(np.asarray(scale, np.float32), "scale")], [m]))
            lin = m
        s = self.fresh("add")
        self.nodes.append(node_proto("Add", [lin, short], [s]))
        r = self.fresh("relu")
        self.nodes.append(node_proto("Relu", [s], [r]))
        return r


def build_ultraface_onnx(width: int = 320, height: int = 240, variant: str = "RFB", seed: int = 0,
                         with_bn: bool = False, cls_bias: float = 0.0, head_gain: float = 1.0,
                         style: str = "simplified", tail: str = "standard") -> bytes:
    """Return the serialized ModelProto.

    style="upstream" writes the graph the way the files the reference downloads
    (`version-RFB-{320,640}.onnx`, nn.rs:21-22: a PyTorch-1.x opset-9 export, not simplified) are shaped:
    every head reshape is fed by a computed shape (Shape -> Gather -> Unsqueeze -> Concat -> Reshape),
    all constants are `Constant` NODES (priors as one [1,K,4] tensor that is sliced in the graph, the
    variances, the reshape dims, the divisor 2), `Slice` carries starts/ends/axes as attributes, and
    center_form_to_corner_form re-slices the concatenated centre-form boxes. Same seed => same weights
    as style="simplified", so both files must produce identical tensors.
    tail="swapped_variances" / "no_exp" / "corner_only": deliberately non-UltraFace decode tails the
    loader must refuse (tail_check.cc).

    cls_bias shifts the face-class logit of every classification head: it sets
    the fraction of priors above min_confidence on synthetic frames (SURVEY.md
    §8d config 5). head_gain scales head weights (spread of logits / offsets).
    """
    assert variant in ("RFB", "slim") and style in ("simplified", "upstream")
    b = _Builder(seed, with_bn)
    upstream = style == "upstream"

    def const(arr: np.ndarray, hint: str = "c") -> str:
        """initialiser (simplified) or Constant node (upstream)"""
        if not upstream:
            return b.init(arr, hint)
        y = b.fresh(hint)
        b.nodes.append(node_proto("Constant", [], [y], value=np.asarray(arr)))
        return y

    c = 16
    x = b.conv("input", 3, c, 3, 2, 1)                       # 0 conv_bn(3,16,2)
    x = b.dw_sep(x, c, 2 * c, 1)                              # 1
    x = b.dw_sep(x, 2 * c, 2 * c, 2)                          # 2
    x = b.dw_sep(x, 2 * c, 2 * c, 1)                          # 3
    x = b.dw_sep(x, 2 * c, 4 * c, 2)                          # 4
    x = b.dw_sep(x, 4 * c, 4 * c, 1)                          # 5
    x = b.dw_sep(x, 4 * c, 4 * c, 1)                          # 6
    x = b.basic_rfb(x, 4 * c, 4 * c) if variant == "RFB" else b.dw_sep(x, 4 * c, 4 * c, 1)  # 7
    src0 = x
    x = b.dw_sep(x, 4 * c, 8 * c, 2)                          # 8
    x = b.dw_sep(x, 8 * c, 8 * c, 1)                          # 9
    x = b.dw_sep(x, 8 * c, 8 * c, 1)                          # 10
    src1 = x
    x = b.dw_sep(x, 8 * c, 16 * c, 2)                         # 11
    x = b.dw_sep(x, 16 * c, 16 * c, 1)                        # 12
    src2 = x
    # extras: 1x1 256->64 ReLU, dw3x3 s2 ReLU, 1x1 64->256 ReLU (bias, no BN)
    x = b.conv(x, 16 * c, 4 * c, 1, bn=False)
    x = b.conv(x, 4 * c, 4 * c, 3, 2, 1, groups=4 * c, bn=False)
    x = b.conv(x, 4 * c, 16 * c, 1, bn=False)
    src3 = x

    anchors = [len(m) for m in MIN_BOXES]
    srcs = [(src0, 4 * c), (src1, 8 * c), (src2, 16 * c), (src3, 16 * c)]

    def head(src: str, cin: int, cout: int, last: bool, shift: np.ndarray | None) -> str:
        if last:
            return b.conv(src, cin, cout, 3, 1, 1, relu=False, bn=False, gain=head_gain, bias_shift=shift)
        y = b.conv(src, cin, cin, 3, 1, 1, groups=cin, relu=True, bn=False)
        return b.conv(y, cin, cout, 1, relu=False, bn=False, gain=head_gain, bias_shift=shift)

    cls_parts, reg_parts = [], []
    for hi, ((src, cin), a) in enumerate(zip(srcs, anchors)):
        last = hi == 3
        for parts, per in ((cls_parts, 2), (reg_parts, 4)):
            shift = None
            if per == 2 and cls_bias != 0.0:
                shift = np.zeros(a * 2, np.float32)
                shift[1::2] = cls_bias  # channel 2*anchor+1 = face logit
            y = head(src, cin, a * per, last, shift)
            t = b.fresh("tr")
            b.nodes.append(node_proto("Transpose", [y], [t], perm=[0, 2, 3, 1]))
            r = b.fresh("rs")
            if upstream:  # x.view(x.size(0), -1, per): the batch dimension is read back from the tensor
                shp = b.fresh("shape")
                b.nodes.append(node_proto("Shape", [t], [shp]))
                g = b.fresh("gather")
                b.nodes.append(node_proto("Gather", [shp, const(np.asarray(0, np.int64), "idx")], [g], axis=0))
                dims = []
                for src_ in (g, const(np.asarray(-1, np.int64), "m1"), const(np.asarray(per, np.int64), "per")):
                    u = b.fresh("unsq")
                    b.nodes.append(node_proto("Unsqueeze", [src_], [u], axes=[0]))
                    dims.append(u)
                tgt = b.fresh("tgt")
                b.nodes.append(node_proto("Concat", dims, [tgt], axis=0))
                b.nodes.append(node_proto("Reshape", [t, tgt], [r]))
            else:
                b.nodes.append(node_proto("Reshape", [t, b.init(np.asarray([1, -1, per], np.int64), "shape")], [r]))
            parts.append(r)

    conf = b.fresh("conf")
    b.nodes.append(node_proto("Concat", cls_parts, [conf], axis=1))
    b.nodes.append(node_proto("Softmax", [conf], ["scores"], axis=2))
    loc = b.fresh("loc")
    b.nodes.append(node_proto("Concat", reg_parts, [loc], axis=1))

    pri = generate_priors(width, height)
    K = pri.shape[0]

    def op(kind: str, ins: list[str], **attrs) -> str:
        y = b.fresh(kind.lower())
        b.nodes.append(node_proto(kind, ins, [y], **attrs))
        return y

    def sl(x: str, lo: int, hi: int) -> str:
        return op("Slice", [x], axes=[2], starts=[lo], ends=[hi])

    if upstream:
        p_all = const(pri[None].copy(), "priors")  # one [1,K,4] Constant, sliced in the graph
        p_xy, p_wh = sl(p_all, 0, 2), sl(p_all, 2, 4)
        p_wh2 = sl(p_all, 2, 4)
    else:
        p_xy = b.init(pri[None, :, :2].copy(), "prior_xy")
        p_wh = p_wh2 = b.init(pri[None, :, 2:].copy(), "prior_wh")
    cv, sv = CENTER_VARIANCE, SIZE_VARIANCE
    if tail == "swapped_variances":
        cv, sv = sv, cv

    # box_utils.convert_locations_to_boxes
    l_xy = sl(loc, 0, 2)
    l_wh = sl(loc, 2, 4)
    cxy = op("Mul", [l_xy, const(np.asarray(cv, np.float32), "cv")])
    cxy = op("Mul", [cxy, p_wh])
    cxy = op("Add", [cxy, p_xy])
    wh = op("Mul", [l_wh, const(np.asarray(sv, np.float32), "sv")])
    if tail != "no_exp":
        wh = op("Exp", [wh])
    wh = op("Mul", [wh, p_wh2])
    if upstream:  # torch.cat([...], dim=-1) then center_form_to_corner_form slices it again
        centre = op("Concat", [cxy, wh], axis=2)
        cxy, wh = sl(centre, 0, 2), sl(centre, 2, 4)
        cxy2 = sl(centre, 0, 2)
        wh2 = sl(centre, 2, 4)
    else:
        cxy2, wh2 = cxy, wh
    if tail == "corner_only":  # a graph that outputs centre-form boxes
        b.nodes.append(node_proto("Concat", [cxy, wh], ["boxes"], axis=2))
    else:
        # box_utils.center_form_to_corner_form
        two = np.asarray(2.0, np.float32)
        tl = op("Sub", [cxy, op("Div", [wh, const(two, "two")])])
        br = op("Add", [cxy2, op("Div", [wh2, const(two, "two")])])
        b.nodes.append(node_proto("Concat", [tl, br], ["boxes"], axis=2))

    return model_proto(
        b.nodes, b.inits,
        inputs=[value_info("input", [1, 3, height, width])],
        outputs=[value_info("scores", [1, K, 2]), value_info("boxes", [1, K, 4])],
    )


def write_ultraface_onnx(path: str, **kw) -> str:
    data = build_ultraface_onnx(**kw)
    with open(path, "wb") as f:
        f.write(data)
    return path
