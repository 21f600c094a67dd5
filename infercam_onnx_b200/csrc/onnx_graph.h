// onnx_graph.h — protobuf wire-format reader for ONNX ModelProto (no protobuf/onnx dependency).
//
// Replaces the parsing half of `tract_onnx::onnx().model_for_path(..)`
// (/root/reference/infer_server/src/nn.rs:168-169). Only what UltraFace exports
// need: nodes, attributes (i, f, ints, floats, t), float/int64 initialisers.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace uf {

struct OnnxTensor {
    std::string name;
    std::vector<int64_t> dims;
    int dtype = 0;  // 1 = float32, 7 = int64 (others are rejected)
    std::vector<float> f;
    std::vector<int64_t> i;
    int64_t numel() const {
        int64_t n = 1;
        for (auto d : dims) n *= d;
        return n;
    }
};

struct OnnxAttr {
    int type = 0;  // 1 f, 2 i, 3 s, 4 t, 6 floats, 7 ints
    float f = 0.f;
    int64_t i = 0;
    std::string s;
    std::vector<int64_t> ints;
    std::vector<float> floats;
    OnnxTensor t;
};

struct OnnxNode {
    std::string op, name;
    std::vector<std::string> inputs, outputs;
    std::map<std::string, OnnxAttr> attrs;
    int64_t attr_i(const std::string& k, int64_t dflt) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? dflt : it->second.i;
    }
    float attr_f(const std::string& k, float dflt) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? dflt : it->second.f;
    }
    std::vector<int64_t> attr_ints(const std::string& k, std::vector<int64_t> dflt) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? dflt : it->second.ints;
    }
};

struct OnnxValueInfo {
    std::string name;
    std::vector<int64_t> dims;
};

struct OnnxModel {
    int64_t opset = 9;
    std::vector<OnnxNode> nodes;
    std::map<std::string, OnnxTensor> initializers;
    std::vector<OnnxValueInfo> inputs, outputs;  // inputs exclude initialisers
};

// Throws std::runtime_error with a message on malformed input.
OnnxModel parse_onnx(const uint8_t* data, size_t size);
OnnxModel load_onnx_file(const std::string& path);  // throws IoError on open failure

struct IoError : public std::exception {
    std::string msg;
    explicit IoError(std::string m) : msg(std::move(m)) {}
    const char* what() const noexcept override { return msg.c_str(); }
};

}  // namespace uf
