// plan.h — lowering of an UltraFace-family ONNX graph to a launch plan.
//
// Replaces the optimising half of tract's `into_optimized().into_runnable()`
// (/root/reference/infer_server/src/nn.rs:170-172): BatchNormalization / scalar Mul /
// per-channel Add are folded into the preceding Conv, Relu and residual Add are fused
// into the Conv epilogue, channel Concat becomes "producers write at a channel offset",
// and the Transpose/Reshape/Concat head plumbing becomes "head convs write NHWC straight
// into the [K,2] / [K,4] buffers" (activations are NHWC, so the ONNX transpose is free).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "onnx_graph.h"

namespace uf {

// NHWC view of one activation tensor inside a per-frame buffer:
// addr(n,y,x,c) = buffer(n) + base_off + (y*W + x)*pix_stride + c
struct TensorDesc {
    std::string name;  // ONNX value name (last name of a fused chain)
    int C = 0, H = 0, W = 0;
    int buf = -1;          // index into Plan::buffers; -1 = graph input (u8 HWC frame)
    int64_t base_off = 0;  // floats
    int pix_stride = 0;    // floats
    bool is_input = false;
    bool in_concat = false;
};

struct BufferDesc {
    int64_t frame_floats = 0;  // 0 = dead (tensor was retargeted into another buffer)
    int64_t arena_off = 0;     // floats per frame before this buffer (filled by finalize)
};

enum class OpKind { Conv, Add, Relu, Copy };

struct Op {
    OpKind kind = OpKind::Conv;
    int in = -1, in2 = -1, out = -1;  // tensor ids (in2: residual for Conv, 2nd operand for Add)
    // Conv attributes
    int cin = 0, cout = 0, k = 1, stride = 1, pad = 0, dil = 1, groups = 1;
    bool relu = false;
    std::vector<float> w;  // [cout][cin/groups][k][k], folded
    std::vector<float> b;  // [cout]
    std::string onnx_node;  // first ONNX output name of the chain (diagnostics)
};

struct Head {
    int cls = -1, reg = -1;  // tensor ids
    int fm_w = 0, fm_h = 0, anchors = 0;
    int prior_off = 0;
};

struct Plan {
    int net_w = 0, net_h = 0;
    std::vector<TensorDesc> tensors;
    std::vector<BufferDesc> buffers;
    std::vector<Op> ops;
    std::vector<Head> heads;
    int num_priors = 0;
    int conf_buf = -1, loc_buf = -1;  // [K,2] raw logits, [K,4] raw offsets (NHWC head outputs)
    std::vector<float> priors;        // [K,4] centre form
    bool priors_from_graph = false;
    float center_variance = 0.1f, size_variance = 0.2f;
    int64_t arena_frame_floats = 0;
    uint64_t macs_per_frame = 0;
    uint64_t conv_bytes_per_frame = 0;  // sum over Conv nodes of (in+out)*4
    std::string warnings;
};

struct UnsupportedError : public std::exception {
    std::string msg;
    explicit UnsupportedError(std::string m) : msg(std::move(m)) {}
    const char* what() const noexcept override { return msg.c_str(); }
};

// Throws UnsupportedError / std::runtime_error.
Plan lower_ultraface(const OnnxModel& m, int net_w, int net_h);

// Upstream UltraFace prior generator (vision/ssd/config/fd_config.py + box_utils.generate_priors):
// per map (fm_w x fm_h, anchors from min_boxes), order (y, x, anchor), clamped to [0,1].
std::vector<float> generate_priors(int net_w, int net_h, const std::vector<Head>& heads);

}  // namespace uf
