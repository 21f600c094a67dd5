// plan.cc — see plan.h.
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <map>
#include <set>
#include <stdexcept>

namespace uf {
void verify_decode_tail(const OnnxModel& m, const std::string& loc_value, const std::string& boxes_value, const Plan& plan);  // tail_check.cc
namespace {

struct HeadRef {
    int tensor = -1;
    int last_dim = -1;  // n per anchor from the Reshape target (2 for cls, 4 for reg), -1 unknown
};

const std::set<std::string> kTailOps = {"Slice", "Mul",   "Add",    "Sub",       "Div",     "Exp",
                                        "Concat", "Shape", "Gather", "Unsqueeze", "Squeeze", "Cast",
                                        "Reshape", "Constant", "Identity"};

struct Lowerer {
    const OnnxModel& m;
    Plan plan;
    std::map<std::string, std::vector<int>> consumers;  // value name -> node indices
    std::set<std::string> graph_outputs;
    std::map<std::string, int> tid;                     // feature tensors
    std::map<std::string, HeadRef> headref;             // Transpose/Reshape outputs of head convs
    std::map<std::string, std::vector<HeadRef>> headlist;  // Concat(axis=1) of headrefs
    std::set<std::string> tail;                         // values of the decode tail (not materialised)
    std::vector<bool> consumed;
    std::vector<HeadRef> cls_list, reg_list;
    std::string reg_value, scores_value;  // ONNX names: concatenated regression heads; Softmax output

    explicit Lowerer(const OnnxModel& mm) : m(mm) {}

    bool is_const(const std::string& n) const { return m.initializers.count(n) > 0; }
    const OnnxTensor& cst(const std::string& n) const { return m.initializers.at(n); }

    int uses(const std::string& n) const {
        auto it = consumers.find(n);
        int u = it == consumers.end() ? 0 : (int)it->second.size();
        return u + (graph_outputs.count(n) ? 1 : 0);
    }
    // index of the only consumer node, or -1
    int sole_consumer(const std::string& n) const {
        if (uses(n) != 1 || graph_outputs.count(n)) return -1;
        return consumers.at(n)[0];
    }

    int new_tensor(const std::string& name, int C, int H, int W) {
        TensorDesc t;
        t.name = name;
        t.C = C; t.H = H; t.W = W;
        t.buf = (int)plan.buffers.size();
        t.pix_stride = C;
        BufferDesc b;
        b.frame_floats = (int64_t)C * H * W;
        plan.buffers.push_back(b);
        plan.tensors.push_back(t);
        tid[name] = (int)plan.tensors.size() - 1;
        return (int)plan.tensors.size() - 1;
    }

    // per-channel constant (scalar, [C], [C,1,1] or [1,C,1,1]) -> vector of C values
    bool channel_const(const std::string& name, int C, std::vector<float>& out) const {
        if (!is_const(name)) return false;
        const OnnxTensor& t = cst(name);
        if (t.dtype != 1) return false;
        int64_t n = t.numel();
        if ((int64_t)t.f.size() != n) return false;
        if (n == 1) { out.assign(C, t.f[0]); return true; }
        if (n != C) return false;
        // the channel axis must be the one carrying C
        size_t nd = t.dims.size();
        bool ok = (nd == 1) || (nd == 3 && t.dims[0] == C) || (nd == 4 && t.dims[1] == C);
        if (!ok) return false;
        out = t.f;
        return true;
    }

    void lower_conv(int idx) {
        const OnnxNode& n = m.nodes[idx];
        if (n.inputs.size() < 2) throw std::runtime_error("onnx: Conv without weights");
        if (!is_const(n.inputs[1])) throw UnsupportedError("Conv weights of '" + n.outputs[0] + "' are not an initialiser");
        const OnnxTensor& wt = cst(n.inputs[1]);
        if (wt.dtype != 1 || wt.dims.size() != 4) throw UnsupportedError("Conv weights must be 4-D float32");
        for (int64_t d : wt.dims)
            if (d < 1 || d > 65536) throw std::runtime_error("onnx: Conv '" + n.outputs[0] + "' weight dimension out of range");
        const TensorDesc in = plan.tensors[tid.at(n.inputs[0])];
        Op op;
        op.kind = OpKind::Conv;
        op.in = tid.at(n.inputs[0]);
        op.onnx_node = n.outputs[0];
        op.cout = (int)wt.dims[0];
        const int64_t grp = n.attr_i("group", 1);
        if (grp < 1 || grp > 65536) throw std::runtime_error("onnx: Conv '" + n.outputs[0] + "' has a bad group count");
        op.groups = (int)grp;
        op.cin = (int)wt.dims[1] * op.groups;
        if (op.cin != in.C) throw std::runtime_error("onnx: Conv '" + n.outputs[0] + "' channel mismatch");
        if (wt.dims[2] != wt.dims[3]) throw UnsupportedError("non-square Conv kernel");
        op.k = (int)wt.dims[2];
        auto ks = n.attr_ints("kernel_shape", {op.k, op.k});
        if (ks.size() != 2 || ks[0] != op.k || ks[1] != op.k) throw std::runtime_error("onnx: kernel_shape/weight mismatch");
        auto st = n.attr_ints("strides", {1, 1});
        auto dl = n.attr_ints("dilations", {1, 1});
        auto pd = n.attr_ints("pads", {0, 0, 0, 0});
        if (st.size() != 2 || dl.size() != 2 || st[0] < 1 || dl[0] < 1 || st[0] > 64 || dl[0] > 64 || pd.size() != 4 || pd[0] < 0 || pd[0] > 1024)
            throw UnsupportedError("strides / dilations / pads of Conv '" + n.outputs[0] + "' are malformed");
        if (st[0] != st[1] || dl[0] != dl[1] || pd.size() != 4 || pd[0] != pd[1] || pd[0] != pd[2] || pd[0] != pd[3])
            throw UnsupportedError("anisotropic stride/dilation/pad in Conv '" + n.outputs[0] + "'");
        auto ap = n.attrs.find("auto_pad");
        if (ap != n.attrs.end() && !ap->second.s.empty() && ap->second.s != "NOTSET")
            throw UnsupportedError("auto_pad in Conv");
        op.stride = (int)st[0]; op.dil = (int)dl[0]; op.pad = (int)pd[0];
        if (op.cout % op.groups || op.cin % op.groups) throw std::runtime_error("onnx: bad group count");
        if ((int64_t)wt.f.size() != wt.numel()) throw std::runtime_error("onnx: Conv weights carry no data");
        op.w = wt.f;
        op.b.assign(op.cout, 0.f);
        if (n.inputs.size() >= 3 && !n.inputs[2].empty()) {
            if (!is_const(n.inputs[2])) throw UnsupportedError("Conv bias is not an initialiser");
            const OnnxTensor& bt = cst(n.inputs[2]);
            if ((int)bt.f.size() != op.cout) throw std::runtime_error("onnx: Conv bias size mismatch");
            op.b = bt.f;
        }
        int eff = op.dil * (op.k - 1) + 1;
        int Ho = (in.H + 2 * op.pad - eff) / op.stride + 1;
        int Wo = (in.W + 2 * op.pad - eff) / op.stride + 1;
        if (Ho <= 0 || Wo <= 0) throw std::runtime_error("onnx: Conv output is empty");

        // ---- fold / fuse the chain hanging off the conv output
        std::string cur = n.outputs[0];
        const size_t per_out = (size_t)(op.cin / op.groups) * op.k * op.k;
        while (true) {
            int ci = sole_consumer(cur);
            if (ci < 0 || consumed[ci]) break;
            const OnnxNode& c = m.nodes[ci];
            // once the skip connection is fused, (conv + res) * v is no longer conv * v + res: only Relu may follow
            if (op.in2 >= 0 && c.op != "Relu") break;
            if (c.op == "BatchNormalization" && c.inputs.size() == 5 && c.inputs[0] == cur) {
                std::vector<float> g, be, mu, var;
                if (!channel_const(c.inputs[1], op.cout, g) || !channel_const(c.inputs[2], op.cout, be) ||
                    !channel_const(c.inputs[3], op.cout, mu) || !channel_const(c.inputs[4], op.cout, var))
                    throw UnsupportedError("BatchNormalization with non-constant statistics");
                float eps = c.attr_f("epsilon", 1e-5f);
                for (int o = 0; o < op.cout; ++o) {
                    float s = g[o] / std::sqrt(var[o] + eps);
                    for (size_t j = 0; j < per_out; ++j) op.w[o * per_out + j] *= s;
                    op.b[o] = (op.b[o] - mu[o]) * s + be[o];
                }
            } else if ((c.op == "Mul" || c.op == "Add") && c.inputs.size() == 2) {
                const std::string& other = c.inputs[0] == cur ? c.inputs[1] : c.inputs[0];
                std::vector<float> v;
                if (channel_const(other, op.cout, v)) {
                    for (int o = 0; o < op.cout; ++o) {
                        if (c.op == "Mul") {
                            for (size_t j = 0; j < per_out; ++j) op.w[o * per_out + j] *= v[o];
                            op.b[o] *= v[o];
                        } else {
                            op.b[o] += v[o];
                        }
                    }
                } else if (c.op == "Add" && op.in2 < 0 && tid.count(other)) {
                    const TensorDesc& r = plan.tensors[tid.at(other)];
                    if (r.C != op.cout || r.H != Ho || r.W != Wo) break;  // broadcasting add: leave to Add op
                    op.in2 = tid.at(other);
                } else {
                    break;
                }
            } else if (c.op == "Relu") {
                op.relu = true;
                consumed[ci] = true;
                cur = c.outputs[0];
                break;
            } else {
                break;
            }
            consumed[ci] = true;
            cur = c.outputs[0];
        }
        op.out = new_tensor(cur, op.cout, Ho, Wo);
        plan.macs_per_frame += (uint64_t)Ho * Wo * op.cout * per_out;
        plan.conv_bytes_per_frame += 4ull * ((uint64_t)in.C * in.H * in.W + (uint64_t)op.cout * Ho * Wo);
        plan.ops.push_back(std::move(op));
    }

    void lower_concat(int idx) {
        const OnnxNode& n = m.nodes[idx];
        int C = 0, H = -1, W = -1;
        for (auto& i : n.inputs) {
            const TensorDesc& t = plan.tensors[tid.at(i)];
            if (H < 0) { H = t.H; W = t.W; }
            if (t.H != H || t.W != W) throw std::runtime_error("onnx: Concat spatial mismatch");
            C += t.C;
        }
        int out = new_tensor(n.outputs[0], C, H, W);
        int obuf = plan.tensors[out].buf;
        int off = 0;
        for (auto& i : n.inputs) {
            int t = tid.at(i);
            TensorDesc& td = plan.tensors[t];
            if (!td.is_input && !td.in_concat && td.pix_stride == td.C && td.base_off == 0) {
                // producer writes straight into the concat buffer: Concat costs nothing
                plan.buffers[td.buf].frame_floats = 0;
                td.buf = obuf; td.base_off = off; td.pix_stride = C; td.in_concat = true;
            } else {
                Op cp;
                cp.kind = OpKind::Copy;
                cp.in = t;
                // destination slice as its own tensor view
                TensorDesc v = plan.tensors[out];
                v.name = n.outputs[0] + "#slice" + std::to_string(off);
                v.C = td.C; v.base_off = off; v.in_concat = true;
                plan.tensors.push_back(v);
                cp.out = (int)plan.tensors.size() - 1;
                plan.ops.push_back(cp);
            }
            off += td.C;
        }
    }

    void run(int net_w, int net_h) {
        plan.net_w = net_w; plan.net_h = net_h;
        if (m.inputs.size() != 1) throw UnsupportedError("expected exactly one graph input");
        if (m.outputs.size() != 2) throw UnsupportedError("expected two graph outputs (scores, boxes)");
        const auto& gi = m.inputs[0];
        if (gi.dims.size() == 4) {
            if (gi.dims[1] > 0 && gi.dims[1] != 3) throw UnsupportedError("graph input is not 3-channel");
            // tract pins the input fact to f32[1,3,H,W] (nn.rs:165-169); a file that declares another
            // fixed size would fail shape inference there, so reject it here too.
            if ((gi.dims[2] > 0 && gi.dims[2] != net_h) || (gi.dims[3] > 0 && gi.dims[3] != net_w))
                throw UnsupportedError("graph input size differs from the requested variant");
        }
        for (auto& o : m.outputs) graph_outputs.insert(o.name);
        for (size_t i = 0; i < m.nodes.size(); ++i)
            for (auto& in : m.nodes[i].inputs)
                if (!in.empty()) consumers[in].push_back((int)i);
        consumed.assign(m.nodes.size(), false);

        TensorDesc in;
        in.name = gi.name; in.C = 3; in.H = net_h; in.W = net_w; in.buf = -1; in.pix_stride = 3; in.is_input = true;
        plan.tensors.push_back(in);
        tid[gi.name] = 0;

        for (size_t idx = 0; idx < m.nodes.size(); ++idx) {
            if (consumed[idx]) continue;
            const OnnxNode& n = m.nodes[idx];
            if (n.op != "Constant" && (n.inputs.empty() || n.outputs.empty()))
                throw std::runtime_error("onnx: node '" + n.op + "' without inputs or outputs");
            auto feat = [&](const std::string& s) { return tid.count(s) > 0; };
            auto tailish = [&](const std::string& s) {
                return s.empty() || is_const(s) || tail.count(s) || headref.count(s) || headlist.count(s);
            };
            if (n.op == "Constant") continue;
            if (n.op == "Conv" && !n.inputs.empty() && feat(n.inputs[0])) { lower_conv((int)idx); continue; }
            if (n.op == "Concat" && n.attr_i("axis", 1) == 1 && !n.inputs.empty() &&
                std::all_of(n.inputs.begin(), n.inputs.end(), feat)) { lower_concat((int)idx); continue; }
            if (n.op == "Relu" && feat(n.inputs[0])) {
                const TensorDesc t = plan.tensors[tid.at(n.inputs[0])];
                Op op; op.kind = OpKind::Relu; op.in = tid.at(n.inputs[0]);
                op.out = new_tensor(n.outputs[0], t.C, t.H, t.W);
                plan.ops.push_back(op);
                continue;
            }
            if (n.op == "Add" && n.inputs.size() == 2 && feat(n.inputs[0]) && feat(n.inputs[1])) {
                const TensorDesc a = plan.tensors[tid.at(n.inputs[0])], b = plan.tensors[tid.at(n.inputs[1])];
                if (a.C != b.C || a.H != b.H || a.W != b.W) throw UnsupportedError("broadcasting Add on feature maps");
                Op op; op.kind = OpKind::Add; op.in = tid.at(n.inputs[0]); op.in2 = tid.at(n.inputs[1]);
                int ci = sole_consumer(n.outputs[0]);
                std::string outn = n.outputs[0];
                if (ci >= 0 && m.nodes[ci].op == "Relu") { op.relu = true; consumed[ci] = true; outn = m.nodes[ci].outputs[0]; }
                op.out = new_tensor(outn, a.C, a.H, a.W);
                plan.ops.push_back(op);
                continue;
            }
            if (n.op == "Transpose" && feat(n.inputs[0])) {
                auto perm = n.attr_ints("perm", {});
                if (perm != std::vector<int64_t>{0, 2, 3, 1}) throw UnsupportedError("Transpose perm other than (0,2,3,1)");
                headref[n.outputs[0]] = HeadRef{tid.at(n.inputs[0]), -1};
                continue;
            }
            if (n.op == "Reshape" && headref.count(n.inputs[0])) {
                HeadRef h = headref.at(n.inputs[0]);
                if (n.inputs.size() > 1 && is_const(n.inputs[1])) {
                    const OnnxTensor& s = cst(n.inputs[1]);
                    if (s.i.size() == 3) h.last_dim = (int)s.i[2];
                }
                headref[n.outputs[0]] = h;
                continue;
            }
            if (n.op == "Concat" && !n.inputs.empty() &&
                std::all_of(n.inputs.begin(), n.inputs.end(), [&](const std::string& s) { return headref.count(s) > 0; })) {
                if (n.attr_i("axis", 1) != 1) throw UnsupportedError("head Concat on axis != 1");
                std::vector<HeadRef> l;
                for (auto& i : n.inputs) l.push_back(headref.at(i));
                headlist[n.outputs[0]] = l;
                continue;
            }
            if (n.op == "Softmax" && headlist.count(n.inputs[0])) {
                int64_t axis = n.attr_i("axis", m.opset < 13 ? 1 : -1);
                if (axis != 2 && axis != -1) throw UnsupportedError("Softmax axis must be the class axis");
                if (!cls_list.empty()) throw UnsupportedError("more than one Softmax over head outputs");
                cls_list = headlist.at(n.inputs[0]);
                scores_value = n.outputs[0];
                tail.insert(n.outputs[0]);
                continue;
            }
            bool touches_headlist = false, touches_feat = false;
            for (auto& i : n.inputs) {
                if (headlist.count(i)) touches_headlist = true;
                if (feat(i)) touches_feat = true;
            }
            if (kTailOps.count(n.op) && touches_headlist) {
                for (auto& i : n.inputs)
                    if (headlist.count(i)) {
                        if (reg_list.empty()) { reg_list = headlist.at(i); reg_value = i; }
                        else if (i != reg_value) throw UnsupportedError("more than one head list feeds the decode tail");
                    }
                for (auto& o : n.outputs) tail.insert(o);
                continue;
            }
            if (n.op == "Shape" && (touches_feat || headref.count(n.inputs[0]))) {  // un-simplified exports
                for (auto& o : n.outputs) tail.insert(o);
                continue;
            }
            if (kTailOps.count(n.op) && std::all_of(n.inputs.begin(), n.inputs.end(), tailish)) {
                // Reshape fed by a computed shape (Shape/Gather/Unsqueeze/Concat chain)
                if (n.op == "Reshape" && headref.count(n.inputs[0])) { headref[n.outputs[0]] = headref.at(n.inputs[0]); continue; }
                for (auto& o : n.outputs) tail.insert(o);
                continue;
            }
            throw UnsupportedError("operator '" + n.op + "' (output '" + (n.outputs.empty() ? "" : n.outputs[0]) +
                                   "') is outside the UltraFace graph family");
        }
        finish_heads();
        merge_sibling_pointwise();
        finalize_buffers();
    }

    void finish_heads() {
        if (cls_list.empty() || reg_list.empty() || cls_list.size() != reg_list.size())
            throw UnsupportedError("could not find matching classification / regression head lists");
        for (auto& o : m.outputs)
            if (!tail.count(o.name)) throw UnsupportedError("graph output '" + o.name + "' is not produced by the SSD tail");
        int K = 0;
        for (size_t i = 0; i < cls_list.size(); ++i) {
            const TensorDesc& c = plan.tensors[cls_list[i].tensor];
            const TensorDesc& r = plan.tensors[reg_list[i].tensor];
            if (c.H != r.H || c.W != r.W || c.C % 2 || r.C != 2 * c.C)
                throw UnsupportedError("head " + std::to_string(i) + ": expected 2 classes and 4 box offsets per anchor");
            if ((cls_list[i].last_dim > 0 && cls_list[i].last_dim != 2) || (reg_list[i].last_dim > 0 && reg_list[i].last_dim != 4))
                throw UnsupportedError("head Reshape target is not [1,-1,2] / [1,-1,4]");
            Head h;
            h.cls = cls_list[i].tensor; h.reg = reg_list[i].tensor;
            h.fm_w = c.W; h.fm_h = c.H; h.anchors = c.C / 2; h.prior_off = K;
            K += c.W * c.H * h.anchors;
            plan.heads.push_back(h);
        }
        plan.num_priors = K;
        BufferDesc cb; cb.frame_floats = (int64_t)K * 2;
        BufferDesc lb; lb.frame_floats = (int64_t)K * 4;
        plan.conf_buf = (int)plan.buffers.size(); plan.buffers.push_back(cb);
        plan.loc_buf = (int)plan.buffers.size(); plan.buffers.push_back(lb);
        for (auto& h : plan.heads) {
            for (int which = 0; which < 2; ++which) {
                TensorDesc& t = plan.tensors[which ? h.reg : h.cls];
                if (t.is_input || t.in_concat) throw UnsupportedError("head tensor is shared with a Concat");
                plan.buffers[t.buf].frame_floats = 0;
                t.buf = which ? plan.loc_buf : plan.conf_buf;
                t.base_off = (int64_t)h.prior_off * (which ? 4 : 2);
                t.pix_stride = t.C;
                t.in_concat = true;
            }
        }
        resolve_priors();
        // outputs[0] must be the Softmax itself (nn.rs:111 reads face probabilities from it), outputs[1] the decoded
        // boxes: verified by running the tail on the host against the formula the post kernel implements
        std::string boxes_value;
        bool scores_ok = false;
        for (auto& o : m.outputs) {
            if (o.name == scores_value) scores_ok = true;
            else boxes_value = o.name;
        }
        if (!scores_ok || boxes_value.empty()) throw UnsupportedError("graph outputs are not (Softmax scores, decoded boxes)");
        if (m.outputs[0].name != scores_value) throw UnsupportedError("graph output 0 is not the Softmax scores (nn.rs:111 expects scores first)");
        verify_decode_tail(m, reg_value, boxes_value, plan);
    }

    void resolve_priors() {
        const int K = plan.num_priors;
        bool can_generate = plan.heads.size() == 4 && plan.heads[0].anchors == 3 && plan.heads[1].anchors == 2 &&
                            plan.heads[2].anchors == 2 && plan.heads[3].anchors == 3;
        if (can_generate) plan.priors = generate_priors(plan.net_w, plan.net_h, plan.heads);
        // graph constants win when they can be identified
        const OnnxTensor *full = nullptr, *xy = nullptr, *wh = nullptr;
        int n_xy = 0, n_wh = 0, n_full = 0;
        for (auto& kv : m.initializers) {
            const OnnxTensor& t = kv.second;
            if (t.dtype != 1 || (int64_t)t.f.size() != t.numel()) continue;
            if (t.numel() == (int64_t)K * 4 && !t.dims.empty() && t.dims.back() == 4) { full = &t; ++n_full; }
            if (t.numel() == (int64_t)K * 2 && !t.dims.empty() && t.dims.back() == 2) {
                bool mul = false, add = false;
                auto it = consumers.find(kv.first);
                if (it != consumers.end())
                    for (int ci : it->second) {
                        if (m.nodes[ci].op == "Mul") mul = true;
                        if (m.nodes[ci].op == "Add") add = true;
                    }
                if (mul && !add) { wh = &t; ++n_wh; }
                if (add && !mul) { xy = &t; ++n_xy; }
            }
        }
        if (n_full == 1) {
            plan.priors = full->f; plan.priors_from_graph = true;
        } else if (n_xy == 1 && n_wh == 1) {
            plan.priors.resize((size_t)K * 4);
            for (int k = 0; k < K; ++k) {
                plan.priors[k * 4 + 0] = xy->f[k * 2]; plan.priors[k * 4 + 1] = xy->f[k * 2 + 1];
                plan.priors[k * 4 + 2] = wh->f[k * 2]; plan.priors[k * 4 + 3] = wh->f[k * 2 + 1];
            }
            plan.priors_from_graph = true;
        } else if (!can_generate) {
            throw UnsupportedError("prior boxes are neither identifiable in the graph nor generatable (non-standard anchors)");
        } else {
            plan.warnings += "priors regenerated from the UltraFace formula (no prior constant identified in the graph); ";
        }
        // variances: scalar float constants feeding Mul in the tail
        std::vector<float> scal;
        for (auto& n : m.nodes)
            if (n.op == "Mul" && n.outputs.size() == 1 && tail.count(n.outputs[0]))
                for (auto& i : n.inputs)
                    if (is_const(i) && cst(i).dtype == 1 && cst(i).numel() == 1 && cst(i).f.size() == 1) scal.push_back(cst(i).f[0]);
        if (scal.size() == 2) {
            plan.center_variance = scal[0];
            plan.size_variance = scal[1];
        } else if (!scal.empty()) {
            plan.warnings += "unexpected number of scalar Mul constants in the decode tail; using variances 0.1/0.2; ";
        }
    }

    // Sibling 1x1 convs that read the same tensor (the three 64->8 branch heads of BasicRFB) become ONE conv with
    // concatenated output channels: the input is read once and two launches disappear. The original outputs stay
    // addressable as channel slices (strided views) of the merged tensor, so their consumers are unchanged.
    void merge_sibling_pointwise() {
        for (size_t i = 0; i < plan.ops.size(); ++i) {
            Op& a = plan.ops[i];
            auto mergeable = [&](const Op& o) {
                if (o.kind != OpKind::Conv || o.k != 1 || o.groups != 1 || o.stride != 1 || o.pad != 0 || o.in2 >= 0) return false;
                const TensorDesc& t = plan.tensors[o.out];
                return !t.in_concat && !t.is_input && t.base_off == 0 && t.pix_stride == t.C && t.C % 4 == 0;
            };
            if (!mergeable(a)) continue;
            std::vector<size_t> sib;
            for (size_t j = i + 1; j < plan.ops.size(); ++j) {
                const Op& b = plan.ops[j];
                if (mergeable(b) && b.in == a.in && b.relu == a.relu && b.cin == a.cin) sib.push_back(j);
            }
            if (sib.empty()) continue;
            const TensorDesc first = plan.tensors[a.out];
            int ctot = a.cout;
            for (size_t j : sib) ctot += plan.ops[j].cout;
            // merged tensor + buffer
            TensorDesc mt;
            mt.name = first.name + "#merged";
            mt.C = ctot; mt.H = first.H; mt.W = first.W;
            mt.buf = (int)plan.buffers.size();
            mt.pix_stride = ctot;
            BufferDesc mb;
            mb.frame_floats = (int64_t)ctot * first.H * first.W;
            plan.buffers.push_back(mb);
            plan.tensors.push_back(mt);
            const int merged_id = (int)plan.tensors.size() - 1;
            int off = 0;
            auto retarget = [&](int tid_) {
                TensorDesc& t = plan.tensors[tid_];
                plan.buffers[t.buf].frame_floats = 0;
                t.buf = mt.buf; t.base_off = off; t.pix_stride = ctot; t.in_concat = true;
                off += t.C;
            };
            retarget(a.out);
            for (size_t j : sib) {
                const Op& b = plan.ops[j];
                a.w.insert(a.w.end(), b.w.begin(), b.w.end());  // [cout][cin] rows simply concatenate
                a.b.insert(a.b.end(), b.b.begin(), b.b.end());
                retarget(b.out);
            }
            a.cout = ctot;
            a.out = merged_id;
            for (size_t k = sib.size(); k-- > 0;) plan.ops.erase(plan.ops.begin() + (long)sib[k]);
        }
    }

    void finalize_buffers() {
        int64_t off = 0;
        for (auto& b : plan.buffers) {
            b.arena_off = off;
            // keep every buffer 16-byte aligned for float4 access at any chunk size
            off += (b.frame_floats + 3) / 4 * 4;
        }
        plan.arena_frame_floats = off;
    }
};

}  // namespace

std::vector<float> generate_priors(int net_w, int net_h, const std::vector<Head>& heads) {
    static const double min_boxes[4][3] = {{10, 16, 24}, {32, 48, 0}, {64, 96, 0}, {128, 192, 256}};
    std::vector<float> p;
    for (size_t hi = 0; hi < heads.size() && hi < 4; ++hi) {
        const Head& h = heads[hi];
        for (int j = 0; j < h.fm_h; ++j)
            for (int i = 0; i < h.fm_w; ++i)
                for (int a = 0; a < h.anchors; ++a) {
                    double v[4] = {(i + 0.5) / h.fm_w, (j + 0.5) / h.fm_h, min_boxes[hi][a] / net_w, min_boxes[hi][a] / net_h};
                    for (double x : v) {
                        float f = (float)x;
                        p.push_back(std::min(1.0f, std::max(0.0f, f)));
                    }
                }
    }
    return p;
}

Plan lower_ultraface(const OnnxModel& m, int net_w, int net_h) {
    Lowerer l(m);
    l.run(net_w, net_h);
    return std::move(l.plan);
}

}  // namespace uf
