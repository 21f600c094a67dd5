// resize_taps.cc — see resize_taps.h. Compiled with -ffp-contract=off: every f32 operation
// below must round exactly like the Rust original (Rust never contracts to FMA).
#include "resize_taps.h"

#include <algorithm>
#include <cmath>

namespace uf {

static inline float triangle_kernel(float x) {
    const float a = std::fabs(x);
    return a < 1.0f ? 1.0f - a : 0.0f;
}

AxisTaps build_axis_taps(int src_len, int dst_len) {
    AxisTaps t;
    t.src_len = src_len;
    t.dst_len = dst_len;
    t.left.resize(dst_len);
    t.ntaps.resize(dst_len);
    const float ratio = (float)src_len / (float)dst_len;
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float support = 1.0f * sratio;  // Triangle filter support is 1.0
    std::vector<std::vector<float>> rows(dst_len);
    for (int o = 0; o < dst_len; ++o) {
        float c = ((float)o + 0.5f) * ratio;
        int64_t l = (int64_t)std::floor(c - support);
        l = std::min<int64_t>(std::max<int64_t>(l, 0), (int64_t)src_len - 1);
        int64_t r = (int64_t)std::ceil(c + support);
        r = std::min<int64_t>(std::max<int64_t>(r, l + 1), (int64_t)src_len);
        c = c - 0.5f;
        const int n = (int)(r - l);
        std::vector<float>& ws = rows[o];
        ws.resize(n);
        float sum = 0.0f;
        for (int i = 0; i < n; ++i) {
            const float w = triangle_kernel(((float)(l + i) - c) / sratio);
            ws[i] = w;
            sum += w;
        }
        for (int i = 0; i < n; ++i) ws[i] /= sum;
        t.left[o] = (int32_t)l;
        t.ntaps[o] = n;
        t.max_taps = std::max(t.max_taps, n);
    }
    t.w.assign((size_t)dst_len * t.max_taps, 0.0f);
    for (int o = 0; o < dst_len; ++o) std::copy(rows[o].begin(), rows[o].end(), t.w.begin() + (size_t)o * t.max_taps);
    return t;
}

int max_tile_span(const AxisTaps& t, int tile) {
    int worst = 0;
    for (int o0 = 0; o0 < t.dst_len; o0 += tile) {
        const int o1 = std::min(o0 + tile, t.dst_len) - 1;
        worst = std::max(worst, t.left[o1] + t.ntaps[o1] - t.left[o0]);
    }
    return worst;
}

}  // namespace uf
