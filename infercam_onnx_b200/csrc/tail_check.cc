// tail_check.cc — structural check of the in-graph SSD decode tail, by evaluation.
//
// The post kernel (kernels_post.cu `tail_one`) hard-codes the UltraFace decode
//   c = loc[:2] * center_variance * prior[2:] + prior[:2];  wh = exp(loc[2:] * size_variance) * prior[2:]
//   boxes = [c - wh/2, c + wh/2]
// while tract executes whatever the ONNX tail says (/root/reference/infer_server/src/nn.rs:181). Exports differ
// in shape (constant-folded or not, Slice in attribute or input form, an extra Concat + re-Slice between
// convert_locations_to_boxes and center_form_to_corner_form), so instead of matching node patterns the loader RUNS
// the tail sub-graph on the host — a tiny interpreter over the dozen element-wise / shape operators it may contain —
// on a seeded probe tensor in place of the concatenated regression heads, and compares the result with the
// hard-coded formula applied to the priors / variances the lowering extracted. Any mismatch (another formula, wrong
// priors, swapped variances) is UF_ERR_UNSUPPORTED at load time instead of silently wrong boxes.
#include <algorithm>
#include <cmath>
#include <functional>
#include <map>
#include <set>
#include <stdexcept>

#include "plan.h"

namespace uf {
namespace {

struct HT {  // host tensor: float or int64
    std::vector<int64_t> dims;
    std::vector<float> f;
    std::vector<int64_t> i;
    bool is_int = false;
    int64_t numel() const {
        int64_t n = 1;
        for (auto d : dims) n *= d;
        return n;
    }
};

[[noreturn]] void bad(const std::string& why) { throw UnsupportedError("decode tail: " + why); }

HT from_onnx(const OnnxTensor& t) {
    HT h;
    h.dims = t.dims;
    if (t.dtype == 1) {
        h.f = t.f;
    } else if (t.dtype == 7) {
        h.i = t.i;
        h.is_int = true;
    } else {
        bad("constant '" + t.name + "' has an unsupported element type");
    }
    if ((int64_t)(h.is_int ? h.i.size() : h.f.size()) != h.numel()) bad("constant '" + t.name + "' carries no data");
    return h;
}

std::vector<int64_t> strides_of(const std::vector<int64_t>& d) {
    std::vector<int64_t> s(d.size(), 1);
    for (int k = (int)d.size() - 2; k >= 0; --k) s[k] = s[k + 1] * d[k + 1];
    return s;
}

HT binary(const std::string& op, const HT& a, const HT& b) {
    if (a.is_int || b.is_int) bad(op + " on integer tensors");
    const size_t nd = std::max(a.dims.size(), b.dims.size());
    std::vector<int64_t> da(nd, 1), db(nd, 1), dout(nd);
    std::copy(a.dims.begin(), a.dims.end(), da.begin() + (nd - a.dims.size()));
    std::copy(b.dims.begin(), b.dims.end(), db.begin() + (nd - b.dims.size()));
    for (size_t k = 0; k < nd; ++k) {
        if (da[k] != db[k] && da[k] != 1 && db[k] != 1) bad(op + " operands do not broadcast");
        dout[k] = std::max(da[k], db[k]);
    }
    HT o;
    o.dims = dout;
    o.f.resize((size_t)o.numel());
    const auto sa = strides_of(da), sb = strides_of(db), so = strides_of(dout);
    for (int64_t idx = 0; idx < o.numel(); ++idx) {
        int64_t ia = 0, ib = 0, r = idx;
        for (size_t k = 0; k < nd; ++k) {
            const int64_t c = r / so[k];
            r -= c * so[k];
            if (da[k] != 1) ia += c * sa[k];
            if (db[k] != 1) ib += c * sb[k];
        }
        const float x = a.f[(size_t)ia], y = b.f[(size_t)ib];
        o.f[(size_t)idx] = op == "Mul" ? x * y : op == "Add" ? x + y : op == "Sub" ? x - y : x / y;
    }
    return o;
}

HT slice(const HT& x, std::vector<int64_t> starts, std::vector<int64_t> ends, std::vector<int64_t> axes,
         std::vector<int64_t> steps) {
    const int nd = (int)x.dims.size();
    if (axes.empty())
        for (size_t k = 0; k < starts.size(); ++k) axes.push_back((int64_t)k);
    if (steps.empty()) steps.assign(starts.size(), 1);
    if (starts.size() != ends.size() || starts.size() != axes.size() || starts.size() != steps.size()) bad("malformed Slice");
    std::vector<int64_t> lo(nd, 0), hi(x.dims);
    for (size_t k = 0; k < starts.size(); ++k) {
        int64_t ax = axes[k] < 0 ? axes[k] + nd : axes[k];
        if (ax < 0 || ax >= nd || steps[k] != 1) bad("Slice axis / step unsupported");
        const int64_t dim = x.dims[ax];
        int64_t s = starts[k] < 0 ? starts[k] + dim : starts[k];
        int64_t e = ends[k] < 0 ? ends[k] + dim : ends[k];
        lo[ax] = std::max<int64_t>(0, std::min(dim, s));
        hi[ax] = std::max<int64_t>(lo[ax], std::min(dim, e));
    }
    HT o;
    o.is_int = x.is_int;
    o.dims.resize(nd);
    for (int k = 0; k < nd; ++k) o.dims[k] = hi[k] - lo[k];
    const auto sx = strides_of(x.dims), so = strides_of(o.dims);
    const int64_t n = o.numel();
    if (x.is_int) o.i.resize((size_t)n); else o.f.resize((size_t)n);
    for (int64_t idx = 0; idx < n; ++idx) {
        int64_t src = 0, r = idx;
        for (int k = 0; k < nd; ++k) {
            const int64_t c = r / so[k];
            r -= c * so[k];
            src += (c + lo[k]) * sx[k];
        }
        if (x.is_int) o.i[(size_t)idx] = x.i[(size_t)src]; else o.f[(size_t)idx] = x.f[(size_t)src];
    }
    return o;
}

HT concat(const std::vector<const HT*>& xs, int64_t axis) {
    if (xs.empty()) bad("empty Concat");
    const int nd = (int)xs[0]->dims.size();
    if (axis < 0) axis += nd;
    if (axis < 0 || axis >= nd) bad("Concat axis out of range");
    HT o;
    o.is_int = xs[0]->is_int;
    o.dims = xs[0]->dims;
    o.dims[axis] = 0;
    for (auto* x : xs) {
        if ((int)x->dims.size() != nd || x->is_int != o.is_int) bad("Concat operand mismatch");
        for (int k = 0; k < nd; ++k)
            if (k != axis && x->dims[k] != xs[0]->dims[k]) bad("Concat shape mismatch");
        o.dims[axis] += x->dims[axis];
    }
    int64_t outer = 1, inner = 1;
    for (int k = 0; k < axis; ++k) outer *= o.dims[k];
    for (int k = (int)axis + 1; k < nd; ++k) inner *= o.dims[k];
    if (o.is_int) o.i.reserve((size_t)o.numel()); else o.f.reserve((size_t)o.numel());
    for (int64_t a = 0; a < outer; ++a)
        for (auto* x : xs) {
            const int64_t len = x->dims[axis] * inner;
            if (o.is_int) o.i.insert(o.i.end(), x->i.begin() + a * len, x->i.begin() + (a + 1) * len);
            else o.f.insert(o.f.end(), x->f.begin() + a * len, x->f.begin() + (a + 1) * len);
        }
    return o;
}

std::vector<int64_t> ints_of(const HT& t) {
    if (!t.is_int) bad("expected an integer tensor");
    return t.i;
}

}  // namespace

void verify_decode_tail(const OnnxModel& m, const std::string& loc_value, const std::string& boxes_value, const Plan& plan) {
    const int K = plan.num_priors;
    std::map<std::string, int> producer;
    for (size_t i = 0; i < m.nodes.size(); ++i)
        for (auto& o : m.nodes[i].outputs) producer[o] = (int)i;
    // nodes between the regression-head Concat and the boxes output (reverse reachability)
    std::set<int> need;
    std::function<void(const std::string&)> visit = [&](const std::string& v) {
        if (v.empty() || v == loc_value || m.initializers.count(v)) return;
        auto it = producer.find(v);
        if (it == producer.end()) bad("value '" + v + "' has no producer");
        if (!need.insert(it->second).second) return;
        if (need.size() > 256) bad("more than 256 nodes between the regression heads and the boxes output");
        for (auto& in : m.nodes[it->second].inputs) visit(in);
    };
    visit(boxes_value);

    std::map<std::string, HT> env;
    HT probe;
    probe.dims = {1, K, 4};
    probe.f.resize((size_t)K * 4);
    uint32_t lcg = 12345u;
    for (auto& v : probe.f) {
        lcg = lcg * 1664525u + 1013904223u;
        v = ((lcg >> 8) * (1.0f / 16777216.0f) - 0.5f) * 6.0f;  // offsets in [-3, 3)
    }
    env[loc_value] = probe;
    auto get = [&](const std::string& v) -> const HT& {
        auto it = env.find(v);
        if (it != env.end()) return it->second;
        auto ci = m.initializers.find(v);
        if (ci == m.initializers.end()) bad("value '" + v + "' is not computable from the regression heads and constants");
        return env[v] = from_onnx(ci->second);
    };
    for (int idx : need) {  // std::set iterates in file order = topological order
        const OnnxNode& n = m.nodes[(size_t)idx];
        if (n.outputs.size() != 1) bad("operator '" + n.op + "' with several outputs");
        HT out;
        if (n.op == "Mul" || n.op == "Add" || n.op == "Sub" || n.op == "Div") {
            if (n.inputs.size() != 2) bad(n.op + " arity");
            out = binary(n.op, get(n.inputs[0]), get(n.inputs[1]));
        } else if (n.op == "Exp") {
            out = get(n.inputs[0]);
            if (out.is_int) bad("Exp on integers");
            for (auto& v : out.f) v = std::exp(v);
        } else if (n.op == "Slice") {
            const HT& x = get(n.inputs[0]);
            if (n.inputs.size() >= 3) {  // opset >= 10: starts / ends / axes / steps are inputs
                out = slice(x, ints_of(get(n.inputs[1])), ints_of(get(n.inputs[2])),
                            n.inputs.size() > 3 && !n.inputs[3].empty() ? ints_of(get(n.inputs[3])) : std::vector<int64_t>{},
                            n.inputs.size() > 4 && !n.inputs[4].empty() ? ints_of(get(n.inputs[4])) : std::vector<int64_t>{});
            } else {
                out = slice(x, n.attr_ints("starts", {}), n.attr_ints("ends", {}), n.attr_ints("axes", {}), {});
            }
        } else if (n.op == "Concat") {
            std::vector<const HT*> xs;
            for (auto& in : n.inputs) xs.push_back(&get(in));
            out = concat(xs, n.attr_i("axis", 0));
        } else if (n.op == "Identity" || n.op == "Cast") {
            out = get(n.inputs[0]);
            if (n.op == "Cast" && n.attr_i("to", 1) != (out.is_int ? 7 : 1)) bad("Cast between element types");
        } else if (n.op == "Unsqueeze" || n.op == "Squeeze") {
            out = get(n.inputs[0]);
            std::vector<int64_t> axes = n.inputs.size() > 1 ? ints_of(get(n.inputs[1])) : n.attr_ints("axes", {});
            std::sort(axes.begin(), axes.end());
            if (n.op == "Unsqueeze") {
                for (int64_t ax : axes) {
                    if (ax < 0) ax += (int64_t)out.dims.size() + 1;
                    if (ax < 0 || ax > (int64_t)out.dims.size()) bad("Unsqueeze axis");
                    out.dims.insert(out.dims.begin() + ax, 1);
                }
            } else {
                for (size_t k = axes.size(); k-- > 0;) {
                    int64_t ax = axes[k] < 0 ? axes[k] + (int64_t)out.dims.size() : axes[k];
                    if (ax < 0 || ax >= (int64_t)out.dims.size() || out.dims[ax] != 1) bad("Squeeze axis");
                    out.dims.erase(out.dims.begin() + ax);
                }
            }
        } else if (n.op == "Reshape") {
            out = get(n.inputs[0]);
            std::vector<int64_t> shp = n.inputs.size() > 1 ? ints_of(get(n.inputs[1])) : n.attr_ints("shape", {});
            int64_t known = 1, wild = -1;
            for (size_t k = 0; k < shp.size(); ++k) {
                if (shp[k] == 0 && k < out.dims.size()) shp[k] = out.dims[k];
                if (shp[k] == -1) wild = (int64_t)k; else known *= shp[k];
            }
            if (wild >= 0) shp[(size_t)wild] = known ? out.numel() / known : 0;
            int64_t total = 1;
            for (auto d : shp) total *= d;
            if (total != out.numel()) bad("Reshape element count");
            out.dims = shp;
        } else if (n.op == "Shape") {
            const HT& x = get(n.inputs[0]);
            out.is_int = true;
            out.dims = {(int64_t)x.dims.size()};
            out.i = x.dims;
        } else if (n.op == "Gather") {
            const HT& x = get(n.inputs[0]);
            const HT& ix = get(n.inputs[1]);
            if (n.attr_i("axis", 0) != 0 || x.dims.size() != 1 || !ix.is_int) bad("Gather other than on a 1-D tensor");
            out.is_int = x.is_int;
            out.dims = ix.dims;
            for (int64_t j : ix.i) {
                if (j < 0) j += x.dims[0];
                if (j < 0 || j >= x.dims[0]) bad("Gather index");
                if (x.is_int) out.i.push_back(x.i[(size_t)j]); else out.f.push_back(x.f[(size_t)j]);
            }
        } else {
            bad("operator '" + n.op + "' is not part of an SSD decode tail");
        }
        env[n.outputs[0]] = std::move(out);
    }
    const HT& got = get(boxes_value);
    if (got.is_int || got.numel() != (int64_t)K * 4 || got.dims.empty() || got.dims.back() != 4)
        bad("boxes output is not [.., K, 4]");
    // the formula the kernel implements, on the extracted priors / variances
    double worst = 0;
    for (int k = 0; k < K; ++k) {
        const float* l = &probe.f[(size_t)k * 4];
        const float* p = &plan.priors[(size_t)k * 4];
        const float cx = l[0] * plan.center_variance * p[2] + p[0];
        const float cy = l[1] * plan.center_variance * p[3] + p[1];
        const float w = std::exp(l[2] * plan.size_variance) * p[2];
        const float h = std::exp(l[3] * plan.size_variance) * p[3];
        const float ref[4] = {cx - w / 2.0f, cy - h / 2.0f, cx + w / 2.0f, cy + h / 2.0f};
        for (int c = 0; c < 4; ++c) worst = std::max(worst, (double)std::fabs(ref[c] - got.f[(size_t)k * 4 + c]));
    }
    if (!(worst <= 2e-5))
        bad("the graph's box decode differs from the UltraFace formula with the extracted priors / variances (max abs " +
            std::to_string(worst) + "); refusing to run a decode this library does not implement");
}

}  // namespace uf
