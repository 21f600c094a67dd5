// onnx_graph.cc — see onnx_graph.h. Field numbers follow onnx.proto (ModelProto.graph = 7,
// GraphProto.node = 1 / initializer = 5 / input = 11 / output = 12, NodeProto.input = 1 /
// output = 2 / name = 3 / op_type = 4 / attribute = 5, TensorProto.dims = 1 / data_type = 2 /
// float_data = 4 / int64_data = 7 / name = 8 / raw_data = 9, AttributeProto.name = 1 / f = 2 /
// i = 3 / s = 4 / t = 5 / floats = 7 / ints = 8 / type = 20).
#include "onnx_graph.h"

#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace uf {
namespace {

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    bool done() const { return p >= end; }
    uint64_t varint() {
        uint64_t v = 0;
        int shift = 0;
        while (true) {
            if (p >= end) throw std::runtime_error("onnx: truncated varint");
            uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7F) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
            if (shift > 63) throw std::runtime_error("onnx: varint too long");
        }
    }
    Reader sub() {
        uint64_t n = varint();
        if ((uint64_t)(end - p) < n) throw std::runtime_error("onnx: truncated length-delimited field");
        Reader r{p, p + n};
        p += n;
        return r;
    }
    void skip(int wire) {
        switch (wire) {
            case 0: varint(); break;
            case 1: need(8); p += 8; break;
            case 2: sub(); break;
            case 5: need(4); p += 4; break;
            default: throw std::runtime_error("onnx: unsupported wire type");
        }
    }
    void need(size_t n) {
        if ((size_t)(end - p) < n) throw std::runtime_error("onnx: truncated fixed field");
    }
    float f32() {
        need(4);
        float v;
        memcpy(&v, p, 4);
        p += 4;
        return v;
    }
    std::string str() {
        Reader r = sub();
        return std::string((const char*)r.p, (size_t)(r.end - r.p));
    }
};

void read_ints(Reader& r, int wire, std::vector<int64_t>& out) {
    if (wire == 0) {
        out.push_back((int64_t)r.varint());
    } else if (wire == 2) {
        Reader s = r.sub();
        while (!s.done()) out.push_back((int64_t)s.varint());
    } else {
        throw std::runtime_error("onnx: bad wire type for repeated int");
    }
}

void read_floats(Reader& r, int wire, std::vector<float>& out) {
    if (wire == 5) {
        out.push_back(r.f32());
    } else if (wire == 2) {
        Reader s = r.sub();
        while (!s.done()) out.push_back(s.f32());
    } else {
        throw std::runtime_error("onnx: bad wire type for repeated float");
    }
}

OnnxTensor parse_tensor(Reader r) {
    OnnxTensor t;
    const uint8_t* raw = nullptr;
    size_t raw_n = 0;
    while (!r.done()) {
        uint64_t key = r.varint();
        int f = (int)(key >> 3), w = (int)(key & 7);
        if (f == 1) read_ints(r, w, t.dims);
        else if (f == 2 && w == 0) t.dtype = (int)r.varint();
        else if (f == 4) read_floats(r, w, t.f);
        else if (f == 7) read_ints(r, w, t.i);
        else if (f == 8 && w == 2) t.name = r.str();
        else if (f == 9 && w == 2) { Reader s = r.sub(); raw = s.p; raw_n = (size_t)(s.end - s.p); }
        else r.skip(w);
    }
    int64_t n = t.numel();
    if (raw) {
        if (t.dtype == 1) {
            if (raw_n != (size_t)n * 4) throw std::runtime_error("onnx: raw_data size mismatch in tensor " + t.name);
            t.f.resize((size_t)n);
            memcpy(t.f.data(), raw, raw_n);
        } else if (t.dtype == 7) {
            if (raw_n != (size_t)n * 8) throw std::runtime_error("onnx: raw_data size mismatch in tensor " + t.name);
            t.i.resize((size_t)n);
            memcpy(t.i.data(), raw, raw_n);
        }
    }
    return t;
}

OnnxAttr parse_attr(Reader r, std::string& name) {
    OnnxAttr a;
    bool has_f = false, has_i = false, has_t = false, has_s = false;
    while (!r.done()) {
        uint64_t key = r.varint();
        int f = (int)(key >> 3), w = (int)(key & 7);
        if (f == 1 && w == 2) name = r.str();
        else if (f == 2 && w == 5) { a.f = r.f32(); has_f = true; }
        else if (f == 3 && w == 0) { a.i = (int64_t)r.varint(); has_i = true; }
        else if (f == 4 && w == 2) { a.s = r.str(); has_s = true; }
        else if (f == 5 && w == 2) { a.t = parse_tensor(r.sub()); has_t = true; }
        else if (f == 7) read_floats(r, w, a.floats);
        else if (f == 8) read_ints(r, w, a.ints);
        else if (f == 20 && w == 0) a.type = (int)r.varint();
        else r.skip(w);
    }
    if (a.type == 0) {  // IR < 3 files omit `type`
        if (!a.ints.empty()) a.type = 7;
        else if (!a.floats.empty()) a.type = 6;
        else if (has_t) a.type = 4;
        else if (has_i) a.type = 2;
        else if (has_f) a.type = 1;
        else if (has_s) a.type = 3;
    }
    return a;
}

OnnxNode parse_node(Reader r) {
    OnnxNode n;
    while (!r.done()) {
        uint64_t key = r.varint();
        int f = (int)(key >> 3), w = (int)(key & 7);
        if (f == 1 && w == 2) n.inputs.push_back(r.str());
        else if (f == 2 && w == 2) n.outputs.push_back(r.str());
        else if (f == 3 && w == 2) n.name = r.str();
        else if (f == 4 && w == 2) n.op = r.str();
        else if (f == 5 && w == 2) {
            std::string name;
            OnnxAttr a = parse_attr(r.sub(), name);
            n.attrs[name] = std::move(a);
        } else r.skip(w);
    }
    return n;
}

OnnxValueInfo parse_value_info(Reader r) {
    OnnxValueInfo v;
    while (!r.done()) {
        uint64_t key = r.varint();
        int f = (int)(key >> 3), w = (int)(key & 7);
        if (f == 1 && w == 2) v.name = r.str();
        else if (f == 2 && w == 2) {
            Reader type = r.sub();
            while (!type.done()) {
                uint64_t k2 = type.varint();
                if ((k2 >> 3) == 1 && (k2 & 7) == 2) {  // tensor_type
                    Reader tt = type.sub();
                    while (!tt.done()) {
                        uint64_t k3 = tt.varint();
                        if ((k3 >> 3) == 2 && (k3 & 7) == 2) {  // shape
                            Reader sh = tt.sub();
                            while (!sh.done()) {
                                uint64_t k4 = sh.varint();
                                if ((k4 >> 3) == 1 && (k4 & 7) == 2) {  // dim
                                    Reader d = sh.sub();
                                    int64_t val = -1;
                                    while (!d.done()) {
                                        uint64_t k5 = d.varint();
                                        if ((k5 >> 3) == 1 && (k5 & 7) == 0) val = (int64_t)d.varint();
                                        else d.skip((int)(k5 & 7));
                                    }
                                    v.dims.push_back(val);
                                } else sh.skip((int)(k4 & 7));
                            }
                        } else tt.skip((int)(k3 & 7));
                    }
                } else type.skip((int)(k2 & 7));
            }
        } else r.skip(w);
    }
    return v;
}

}  // namespace

OnnxModel parse_onnx(const uint8_t* data, size_t size) {
    OnnxModel m;
    Reader r{data, data + size};
    bool have_graph = false;
    while (!r.done()) {
        uint64_t key = r.varint();
        int f = (int)(key >> 3), w = (int)(key & 7);
        if (f == 7 && w == 2) {
            have_graph = true;
            Reader g = r.sub();
            while (!g.done()) {
                uint64_t k2 = g.varint();
                int f2 = (int)(k2 >> 3), w2 = (int)(k2 & 7);
                if (f2 == 1 && w2 == 2) m.nodes.push_back(parse_node(g.sub()));
                else if (f2 == 5 && w2 == 2) {
                    OnnxTensor t = parse_tensor(g.sub());
                    std::string name = t.name;
                    m.initializers[name] = std::move(t);
                } else if (f2 == 11 && w2 == 2) m.inputs.push_back(parse_value_info(g.sub()));
                else if (f2 == 12 && w2 == 2) m.outputs.push_back(parse_value_info(g.sub()));
                else g.skip(w2);
            }
        } else if (f == 8 && w == 2) {
            Reader o = r.sub();
            std::string domain;
            int64_t version = 0;
            while (!o.done()) {
                uint64_t k2 = o.varint();
                if ((k2 >> 3) == 1 && (k2 & 7) == 2) domain = o.str();
                else if ((k2 >> 3) == 2 && (k2 & 7) == 0) version = (int64_t)o.varint();
                else o.skip((int)(k2 & 7));
            }
            if (domain.empty() || domain == "ai.onnx") m.opset = version;
        } else r.skip(w);
    }
    if (!have_graph) throw std::runtime_error("onnx: ModelProto has no graph");
    // graph inputs that are initialisers are weights, not runtime inputs (IR < 4 lists both)
    std::vector<OnnxValueInfo> real;
    for (auto& in : m.inputs)
        if (!m.initializers.count(in.name)) real.push_back(in);
    m.inputs = real;
    // Constant nodes become initialisers so the lowering sees one kind of constant
    for (auto& n : m.nodes) {
        if (n.op == "Constant" && n.outputs.size() == 1) {
            auto it = n.attrs.find("value");
            if (it != n.attrs.end() && it->second.type == 4) {
                OnnxTensor t = it->second.t;
                t.name = n.outputs[0];
                m.initializers[t.name] = std::move(t);
            }
        }
    }
    return m;
}

OnnxModel load_onnx_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw IoError("cannot open model file '" + path + "'");
    std::vector<uint8_t> buf;
    uint8_t tmp[1 << 16];
    size_t n;
    while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    fclose(f);
    if (buf.empty()) throw IoError("model file '" + path + "' is empty");
    return parse_onnx(buf.data(), buf.size());
}

}  // namespace uf
