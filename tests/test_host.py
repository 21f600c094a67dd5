"""CPU tests of the host side of the product: the C-ABI library loads and exports every declared
symbol, the ONNX loader/lowering (no GPU needed) agrees with the oracle's independent reader, the
resize tap tables are bit-identical to the oracle's, and errors are loud (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from infercam_onnx_b200 import _capi, nn
from tools.onnx_fixture import generate_priors
from oracle import hotpath
from oracle.onnx_reader import load_onnx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ultraface_b200.h")).read()
    declared = set(re.findall(r"UF_API\s+[\w\s\*]+?\b(uf_\w+)\s*\(", header))
    assert len(declared) >= 24
    lib = ctypes.CDLL(_capi.lib_path() if os.path.exists(_capi.lib_path()) else _capi._build.build())
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ultraface_b200.h but not exported"
    assert declared == set(_capi.SIGNATURES), "ctypes binding and header disagree"
    assert b"sm_100a" in _capi.load().uf_version()


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_capi.uf_det) == 20
    assert ctypes.sizeof(_capi.uf_kernel_stat) == 64 + 8 * 5


@pytest.mark.parametrize("wh,variant,K,macs", [((320, 240), "RFB", 4420, 100.4e6), ((640, 480), "RFB", 17640, 399.2e6),
                                               ((320, 240), "slim", 4420, 81.4e6)])
def test_lowering_matches_survey_figures(make_onnx, wh, variant, K, macs):
    d = nn.onnx_inspect(make_onnx(*wh, variant=variant), *wh)
    assert d["num_priors"] == K
    assert abs(d["macs_per_frame"] - macs) / macs < 0.005
    assert d["priors_from_graph"] and d["warnings"] == ""
    assert [h["anchors"] for h in d["heads"]] == [3, 2, 2, 3]
    assert abs(d["priors_sum"] - float(generate_priors(*wh).astype(np.float64).sum())) < 1e-3
    convs = [o for o in d["ops"] if o["kind"] == 0]
    # 52 / 42 Conv nodes; the three sibling 64->8 1x1 convs of BasicRFB are merged into one 64->24 launch
    assert len(convs) == (50 if variant == "RFB" else 42) and len(convs) == len(d["ops"])  # everything fused
    assert sum(o["out"].endswith("#merged") and o["cout"] == 24 for o in convs) == (1 if variant == "RFB" else 0)
    if wh == (320, 240) and variant == "RFB":
        assert abs(d["conv_bytes_per_frame"] - 27.90e6) < 0.02e6
        assert sum(o["residual"] for o in convs) == 1  # RFB shortcut add fused into a conv epilogue
        assert sorted({o["pix_stride"] for o in convs if o["cout"] == 16 and o["dil"] > 1}) == [48]  # concat-free


def test_bn_folding_matches_numpy(make_onnx):
    path = make_onnx(320, 240, with_bn=True, seed=3)
    d = nn.onnx_inspect(path, 320, 240)
    g = load_onnx(path)
    nodes = {n.outputs[0]: n for n in g.nodes}
    consumers = {}
    for n in g.nodes:
        for i in n.inputs:
            consumers.setdefault(i, []).append(n)
    checked = 0
    convs = [n for n in g.nodes if n.op == "Conv"]
    expected = []  # (input name, kernel, w_abs, b_sum) per Conv node after numpy BN folding
    for n in convs:
        w = g.initializers[n.inputs[1]].astype(np.float64)
        b = g.initializers[n.inputs[2]].astype(np.float64) if len(n.inputs) > 2 else np.zeros(w.shape[0])
        nxt = consumers.get(n.outputs[0], [])
        if len(nxt) == 1 and nxt[0].op == "BatchNormalization":
            gam, bet, mu, var = (g.initializers[k].astype(np.float64) for k in nxt[0].inputs[1:5])
            s = gam / np.sqrt(var + nxt[0].attrs.get("epsilon", 1e-5))
            w = w * s[:, None, None, None]
            b = (b - mu) * s + bet
            checked += 1
        expected.append([n.inputs[0], w.shape[2], np.abs(w).sum(), b.sum(), w.shape[0]])
    # sibling 1x1 convs on the same input are merged by the lowering: fold their sums into the first one
    merged, seen = [], {}
    for e in expected:
        key = (e[0], e[1])
        if e[1] == 1 and e[4] == 8 and key in seen:
            merged[seen[key]][2] += e[2]
            merged[seen[key]][3] += e[3]
            continue
        if e[1] == 1 and e[4] == 8:
            seen[key] = len(merged)
        merged.append(list(e))
    assert len(merged) == len(d["ops"])
    for e, o in zip(merged, d["ops"]):
        assert o["w_abs"] == pytest.approx(e[2], rel=1e-4)
        assert o["b_sum"] == pytest.approx(e[3], abs=1e-3)
    assert checked >= 30
    assert nodes  # silence linters


def test_resize_taps_bit_identical_to_oracle():
    for s, dlen in [(640, 320), (480, 240), (427, 240), (427, 480), (462, 240), (676, 480), (960, 240), (1280, 320),
                    (720, 240), (1920, 320), (1080, 240), (100, 320), (320, 320), (641, 320), (7, 320)]:
        l, n, w = nn.resize_taps(s, dlen)
        ol, on, ow = hotpath.axis_taps(s, dlen)
        np.testing.assert_array_equal(l, ol)
        np.testing.assert_array_equal(n, on)
        np.testing.assert_array_equal(w, ow)


def test_errors_are_loud(tmp_path, make_onnx):
    with pytest.raises(nn.UltrafaceError) as e:
        nn.onnx_inspect(str(tmp_path / "missing.onnx"), 320, 240)
    assert e.value.code == 2  # UF_ERR_IO (the reference would try to download: nn.rs:156-162)
    bad = tmp_path / "bad.onnx"
    bad.write_bytes(b"\x3a\xff\xff\xff\xff\x0f garbage")
    with pytest.raises(nn.UltrafaceError) as e:
        nn.onnx_inspect(str(bad), 320, 240)
    assert e.value.code == 3  # UF_ERR_ONNX
    with pytest.raises(nn.UltrafaceError) as e:  # declared 320x240 graph, asked for 640x480
        nn.onnx_inspect(make_onnx(320, 240), 640, 480)
    assert e.value.code == 4
    with pytest.raises(nn.UltrafaceError):
        nn._as_rgb(np.zeros((4, 4), np.uint8))


def test_no_cpu_fallback(make_onnx):
    """Without a CUDA device the product refuses to load: nothing routes through the oracle."""
    if nn.device_count() > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(nn.UltrafaceError) as e:
        nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=make_onnx(320, 240))
    assert e.value.code == 6
    src = ""
    pkg = os.path.join(ROOT, "infercam_onnx_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cc", ".h")):
                src += open(os.path.join(dp, f)).read()
    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), "product must not import the oracle"


def test_variant_mirror():
    assert nn.UltrafaceVariant.W640H480.width_height() == (640, 480)  # nn.rs:36-41
    assert nn.UltrafaceVariant.W320H240.width_height() == (320, 240)
    assert nn.default_model_path(nn.UltrafaceVariant.W320H240).endswith("infercam_onnx/ultraface-RFB-320.onnx")


def test_upstream_shaped_export_lowers_to_the_same_plan(make_onnx):
    """nn.rs:165-172 loads `version-RFB-*.onnx`, an un-simplified opset-9 export: Shape/Gather/Unsqueeze/Concat reshape
    targets, Constant-node priors sliced in the graph, attribute-form Slice, explicit BatchNormalization. The lowering
    must produce the plan of the simplified file with the same weights, with the priors taken from the graph."""
    for wh in ((320, 240), (640, 480)):
        for with_bn in (False, True):
            a = nn.onnx_inspect(make_onnx(*wh, seed=4, with_bn=with_bn), *wh)
            b = nn.onnx_inspect(make_onnx(*wh, seed=4, with_bn=with_bn, style="upstream"), *wh)
            assert b["priors_from_graph"] and b["warnings"] == ""
            assert (a["num_priors"], a["priors_sum"], a["center_variance"], a["size_variance"]) == \
                   (b["num_priors"], b["priors_sum"], b["center_variance"], b["size_variance"])
            assert len(a["ops"]) == len(b["ops"])
            for x, y in zip(a["ops"], b["ops"]):
                for k in ("kind", "cin", "cout", "k", "stride", "pad", "dil", "groups", "relu", "residual", "h", "w",
                          "pix_stride", "base_off", "w_sum", "w_abs", "b_sum"):
                    assert x[k] == y[k], (k, x["out"], y["out"])


@pytest.mark.parametrize("style", ["simplified", "upstream"])
@pytest.mark.parametrize("tail", ["no_exp", "corner_only"])
def test_decode_tail_is_verified_by_evaluation(make_onnx, style, tail):
    """The post kernel hard-codes the UltraFace decode; a graph whose tail computes anything else must be refused at load
    (tail_check.cc runs the tail sub-graph on the host against the formula), not run with silently wrong boxes."""
    with pytest.raises(nn.UltrafaceError) as e:
        nn.onnx_inspect(make_onnx(320, 240, style=style, tail=tail), 320, 240)
    assert e.value.code == 4 and "decode" in str(e.value)


def _patched_onnx(tmp_path, name, mutate):
    """Writes a small fixture graph after `mutate(builder_module)` has monkey-patched the writer."""
    import tools.onnx_fixture as fx
    path = str(tmp_path / name)
    undo = mutate(fx)
    try:
        fx.write_ultraface_onnx(path, width=320, height=240, seed=1)
    finally:
        undo()
    return path


def test_hostile_conv_attributes_do_not_crash(tmp_path):
    """Entry points never abort (header contract): malformed strides / dilations / group are errors, not SIGFPE / OOB."""
    import tools.onnx_fixture as fx
    orig = fx.node_proto
    for bad in (dict(strides=[2]), dict(dilations=[]), dict(group=0), dict(strides=[0, 0]), dict(pads=[1, 1])):
        def patched(op, inputs, outputs, name="", **attrs):
            if op == "Conv":
                attrs.update(bad)
            return orig(op, inputs, outputs, name, **attrs)

        def mutate(mod):
            mod.node_proto = patched
            return lambda: setattr(mod, "node_proto", orig)
        path = _patched_onnx(tmp_path, "hostile.onnx", mutate)
        with pytest.raises(nn.UltrafaceError) as e:
            nn.onnx_inspect(path, 320, 240)
        assert e.value.code in (3, 4), bad


def test_scale_after_residual_add_is_not_folded(tmp_path):
    """conv -> Add(skip) -> Mul(v): folding v into the conv alone would compute conv*v + skip. The lowering must refuse
    (or keep the Mul), never fold."""
    import tools.onnx_fixture as fx
    orig = fx.node_proto
    state = {"after_add": None}

    def patched(op, inputs, outputs, name="", **attrs):
        if op == "Relu" and state["after_add"] == inputs[0]:
            # insert a per-tensor scale between the residual Add and its Relu
            scale = orig("Constant", [], ["rfb_scale"], value=np.asarray(2.0, np.float32))
            mul = orig("Mul", [inputs[0], "rfb_scale"], ["rfb_scaled"])
            state["extra"] = [scale, mul]
            return orig(op, ["rfb_scaled"], outputs, name, **attrs)
        if op == "Add" and len(inputs) == 2 and inputs[0].startswith(("mul", "conv")) and inputs[1].startswith("conv"):
            state["after_add"] = outputs[0]
        return orig(op, inputs, outputs, name, **attrs)

    class ListProxy(list):
        def append(self, item):
            extra = state.pop("extra", None)
            if extra:
                for x in extra:
                    super().append(x)
            super().append(item)

    orig_builder_init = fx._Builder.__init__

    def builder_init(self, seed, with_bn):
        orig_builder_init(self, seed, with_bn)
        self.nodes = ListProxy()

    def mutate(mod):
        mod.node_proto = patched
        mod._Builder.__init__ = builder_init

        def undo():
            mod.node_proto = orig
            mod._Builder.__init__ = orig_builder_init
        return undo
    path = _patched_onnx(tmp_path, "scaled_residual.onnx", mutate)
    g = load_onnx(path)
    assert any(n.op == "Mul" and "rfb_scale" in n.inputs for n in g.nodes), "fixture patch did not take"
    with pytest.raises(nn.UltrafaceError) as e:
        nn.onnx_inspect(path, 320, 240)
    assert e.value.code == 4


def test_rust_sys_crate_declares_every_header_symbol():
    """rust/ultraface-sys cannot be compiled here (no cargo), so at least keep it complete against the header."""
    header = open(os.path.join(ROOT, "include", "ultraface_b200.h")).read()
    declared = set(re.findall(r"UF_API\s+[\w\s\*]+?\b(uf_\w+)\s*\(", header))
    rust = open(os.path.join(ROOT, "rust", "ultraface-sys", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (uf_\w+)\s*\(", rust))
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))
    # struct field lists follow the header order
    for struct, cls in (("uf_config", _capi.uf_config), ("uf_info", _capi.uf_info), ("uf_result", _capi.uf_result),
                        ("uf_batcher_config", _capi.uf_batcher_config), ("uf_batcher_stats", _capi.uf_batcher_stats),
                        ("uf_jpeg_info", _capi.uf_jpeg_info)):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % struct, rust, re.S).group(1)
        assert re.findall(r"pub (\w+):", body) == [f for f, _ in cls._fields_], struct
