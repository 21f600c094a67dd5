// resize_taps.h — host-side tap tables of `image::imageops::resize(.., FilterType::Triangle)`
// (call site /root/reference/infer_server/src/nn.rs:74-80; algorithm: image 0.24.5
// src/imageops/sample.rs horizontal_sample / vertical_sample). The weights depend only on
// (src_len, dst_len), so they are computed once per size pair with exactly the crate's f32
// operation order and uploaded; the kernel (kernels_preproc.cu) only multiplies and adds.
#pragma once
#include <cstdint>
#include <vector>

namespace uf {

struct AxisTaps {
    int src_len = 0, dst_len = 0, max_taps = 0;
    std::vector<int32_t> left, ntaps;  // [dst_len]
    std::vector<float> w;              // [dst_len][max_taps], normalised, zero padded
};

AxisTaps build_axis_taps(int src_len, int dst_len);

// widest span of source indices needed by any tile of `tile` consecutive outputs
int max_tile_span(const AxisTaps& t, int tile);

}  // namespace uf
