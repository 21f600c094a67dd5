"""CPU restatement of `self.model.run(...)` (infer_server/src/nn.rs:181) and of the
whole `UltrafaceModel::run` (nn.rs:179-185).

TEST INFRASTRUCTURE — see oracle/__init__.py. PARITY STATUS: "parity unpinned":
the CNN runs inside the un-vendored crate tract-onnx 0.19.2 (Cargo.toml:32,
Cargo.lock:2561-2562) on weights downloaded at run time (nn.rs:21-22); neither the
crate, cargo, nor the .onnx files exist here. What is restated is the published
semantics of the ONNX operators the UltraFace export uses — tract executes the
same operator definitions — evaluated node by node in fp32 (and optionally fp64
for the error budget) with PyTorch-CPU kernels. Summation order inside a Conv
differs from tract's im2col+FMA micro-kernels, which is why the contract is
"raw tensors within 1e-4 abs" (BASELINE.json north_star), not bit-exact.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import hotpath
from .onnx_reader import Graph, load_onnx


class OnnxInterpreter:
    """Evaluates the graph in file order (ONNX guarantees topological order)."""

    def __init__(self, graph: Graph, dtype=torch.float32):
        self.g = graph
        self.dtype = dtype
        self.consts = {}
        for k, v in graph.initializers.items():
            t = torch.from_numpy(np.array(v))
            self.consts[k] = t.to(dtype) if t.is_floating_point() else t

    def __call__(self, x: torch.Tensor, keep: bool = False):
        env = dict(self.consts)
        env[self.g.inputs[0][0]] = x.to(self.dtype)
        for n in self.g.nodes:
            ins = [env[i] if i else None for i in n.inputs]
            outs = getattr(self, "op_" + n.op)(n, *ins)
            if not isinstance(outs, tuple):
                outs = (outs,)
            for name, val in zip(n.outputs, outs):
                env[name] = val
        result = [env[o[0]] for o in self.g.outputs]
        return (result, env) if keep else result

    # ---- operators (ONNX operator spec, opset 9..13 forms used by UltraFace exports)
    def op_Conv(self, n, x, w, b=None):
        a = n.attrs
        k = a.get("kernel_shape", list(w.shape[2:]))
        pads = a.get("pads", [0, 0, 0, 0])
        assert pads[0] == pads[2] and pads[1] == pads[3], "asymmetric pads unsupported"
        return F.conv2d(x, w, b, stride=tuple(a.get("strides", [1, 1])), padding=(pads[0], pads[1]),
                        dilation=tuple(a.get("dilations", [1, 1])), groups=a.get("group", 1))

    def op_BatchNormalization(self, n, x, gamma, beta, mean, var):
        eps = n.attrs.get("epsilon", 1e-5)
        shape = (1, -1, 1, 1)
        return (x - mean.view(shape)) / torch.sqrt(var.view(shape) + eps) * gamma.view(shape) + beta.view(shape)

    def op_Relu(self, n, x):
        return torch.relu(x)

    def op_Concat(self, n, *xs):
        return torch.cat(xs, dim=n.attrs["axis"])

    def op_Add(self, n, a, b):
        return a + b

    def op_Sub(self, n, a, b):
        return a - b

    def op_Mul(self, n, a, b):
        return a * b

    def op_Div(self, n, a, b):
        return a / b

    def op_Exp(self, n, x):
        return torch.exp(x)

    def op_Transpose(self, n, x):
        return x.permute(*n.attrs["perm"]).contiguous()

    def op_Reshape(self, n, x, shape):
        shp = [int(s) for s in shape.tolist()]
        shp = [x.shape[i] if s == 0 else s for i, s in enumerate(shp)]
        return x.reshape(shp)

    def op_Softmax(self, n, x):
        axis = n.attrs.get("axis", 1 if self.g.opset < 13 else -1)
        if self.g.opset < 13:
            # opset<13: coerce to 2-D at `axis`, softmax over the flattened tail
            axis = axis % x.dim()
            lead = int(np.prod(x.shape[:axis])) if axis > 0 else 1
            return torch.softmax(x.reshape(lead, -1), dim=1).reshape(x.shape)
        return torch.softmax(x, dim=axis)

    def op_Slice(self, n, x, starts=None, ends=None, axes=None, steps=None):
        if starts is None:
            starts, ends = n.attrs["starts"], n.attrs["ends"]
            axes = n.attrs.get("axes", list(range(len(starts))))
            steps = [1] * len(starts)
        else:
            starts, ends = starts.tolist(), ends.tolist()
            axes = axes.tolist() if axes is not None else list(range(len(starts)))
            steps = steps.tolist() if steps is not None else [1] * len(starts)
        idx = [slice(None)] * x.dim()
        for s, e, ax, st in zip(starts, ends, axes, steps):
            dim = x.shape[ax]
            s = max(0, min(dim, s + dim if s < 0 else s))
            e = max(0, min(dim, e + dim if e < 0 else e))
            idx[ax] = slice(s, e, st)
        return x[tuple(idx)]

    def op_Constant(self, n):
        t = torch.from_numpy(np.array(n.attrs["value"]))
        return t.to(self.dtype) if t.is_floating_point() else t

    def op_Shape(self, n, x):
        return torch.tensor(list(x.shape), dtype=torch.int64)

    def op_Gather(self, n, x, idx):
        return torch.index_select(x, n.attrs.get("axis", 0), idx.reshape(-1)).reshape(
            list(x.shape[:n.attrs.get("axis", 0)]) + list(idx.shape) + list(x.shape[n.attrs.get("axis", 0) + 1:]))

    def op_Unsqueeze(self, n, x, axes=None):
        axes = n.attrs["axes"] if axes is None else axes.tolist()
        for ax in sorted(axes):
            x = x.unsqueeze(ax)
        return x

    def op_Cast(self, n, x):
        return x.to({1: self.dtype, 6: torch.int32, 7: torch.int64}[n.attrs["to"]])


class UltrafaceOracle:
    """Mirror of `UltrafaceModel` (nn.rs:45-67,178-186) on the CPU.

    new(variant,max_iou,min_confidence) -> __init__(onnx, width, height, max_iou, min_confidence)
    run(&RgbImage) -> run(rgb HxWx3 u8) -> list[((x0,y0,x1,y1), conf)]
    """

    def __init__(self, onnx, width: int, height: int, max_iou: float = 0.5, min_confidence: float = 0.5,
                 norm_preset: int = 0, dtype=torch.float32, round_intermediate: bool = False):
        self.graph = onnx if isinstance(onnx, Graph) else load_onnx(onnx)
        self.net = OnnxInterpreter(self.graph, dtype)
        self.width, self.height = width, height
        self.max_iou, self.min_confidence = float(max_iou), float(min_confidence)
        self.norm_preset = norm_preset
        self.round_intermediate = round_intermediate

    # nn.rs:70-94
    def preproc_u8(self, rgb: np.ndarray) -> np.ndarray:
        return hotpath.resize_triangle(rgb, self.width, self.height, self.round_intermediate)

    def preproc(self, rgb: np.ndarray) -> np.ndarray:
        return hotpath.normalise_nchw(self.preproc_u8(rgb), self.norm_preset)[None]

    # nn.rs:181
    def raw(self, rgb_batch) -> tuple[np.ndarray, np.ndarray]:
        """scores [B,K,2], boxes [B,K,4] for a list of HxWx3 u8 frames (graph is batch-1: loop)."""
        scores, boxes = [], []
        with torch.no_grad():
            for rgb in rgb_batch:
                s, b = self.net(torch.from_numpy(self.preproc(rgb)))
                scores.append(s.to(torch.float32).numpy()[0])
                boxes.append(b.to(torch.float32).numpy()[0])
        return np.stack(scores), np.stack(boxes)

    def raw_from_tensor(self, x: np.ndarray):
        with torch.no_grad():
            s, b = self.net(torch.from_numpy(x))
        return s.to(torch.float32).numpy(), b.to(torch.float32).numpy()

    # nn.rs:109-140
    def postproc(self, scores: np.ndarray, boxes: np.ndarray):
        dets, _ = hotpath.postproc(scores, boxes, self.min_confidence, self.max_iou)
        return [((float(d[0]), float(d[1]), float(d[2]), float(d[3])), float(d[4])) for d in dets]

    # nn.rs:179-185
    def run(self, rgb: np.ndarray):
        s, b = self.raw([rgb])
        return self.postproc(s[0], b[0])
