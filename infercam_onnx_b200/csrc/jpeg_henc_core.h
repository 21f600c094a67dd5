// jpeg_henc_core.h — the per-block part of the device Huffman CODER (kernels_jpeg_henc.cu), written so that the same lines
// compile for the device and, for the CPU-only test that walks the blocks serially (tests/helpers/henc_sim.cc), for the
// host: where a block's coefficients come from (MCU order -> plane raster, libjpeg's dummy blocks), its DC predictor, the
// bits of one coefficient, and how bits are OR-ed into the big-endian bit buffer. Nothing in the library calls the host build.
#pragma once
#include <cstddef>
#include <cstdint>

#include "jpeg_decode.h"

#ifdef __CUDACC__
#define HE_FN __device__ __forceinline__
#else
#define HE_FN inline
#ifndef __restrict__
#define __restrict__
#endif
#endif

namespace uf {
namespace he {

struct EncTabs {          // per CTA, in shared memory
    uint32_t dc[2][16];   // (length << 16) | code, [0] luma [1] chroma
    uint32_t ac[2][256];
    uint8_t zz[64];       // zigzag position -> natural index
};

HE_FN uint32_t bit_length(uint32_t a) {  // 0 for 0
#ifdef __CUDA_ARCH__
    return 32u - (uint32_t)__clz((int)a);
#else
    return a ? 32u - (uint32_t)__builtin_clz(a) : 0u;
#endif
}
HE_FN uint32_t top_bit64(unsigned long long m) {  // index of the highest set bit (m != 0)
#ifdef __CUDA_ARCH__
    return 63u - (uint32_t)__clzll((long long)m);
#else
    return 63u - (uint32_t)__builtin_clzll(m);
#endif
}

HE_FN void or_word(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    atomicOr(p, v);
#else
    *p |= v;
#endif
}

// n bits (1..32) of v at bit position pos of the big-endian bit buffer P
HE_FN void put32(uint32_t* __restrict__ P, uint32_t v, uint32_t n, uint32_t pos) {
    const uint32_t sh = pos & 31u, w = pos >> 5, avail = 32u - sh;
    if (n <= avail) {
        or_word(P + w, v << (avail - n));
    } else {
        or_word(P + w, v >> (n - avail));
        or_word(P + w + 1, v << (32u - (n - avail)));
    }
}

HE_FN void put64(uint32_t* __restrict__ P, unsigned long long bits, uint32_t len, uint32_t pos, uint32_t cap_bits) {
    if (len == 0 || pos + len > cap_bits) return;  // (an overflowing frame is flagged by the scan kernel and redone on the host)
    if (len > 32) {
        put32(P, (uint32_t)(bits >> 32), len - 32, pos);
        put32(P, (uint32_t)bits, 32, pos + len - 32);
    } else {
        put32(P, (uint32_t)bits & (len == 32 ? 0xffffffffu : ((1u << len) - 1u)), len, pos);
    }
}

// the bits of one coefficient: AC at zigzag position k (k >= 1) with the nonzero mask M of the block, or the DC difference
HE_FN void coef_bits(const EncTabs& T, int chroma, uint32_t k, int v, unsigned long long M, unsigned long long& bits, uint32_t& len) {
    bits = 0;
    len = 0;
    if (k != 0 && v == 0) return;
    int a = v, m = v;
    if (v < 0) { a = -v; m = v - 1; }
    const uint32_t size = bit_length((uint32_t)a);
    const uint32_t mag = (uint32_t)m & ((1u << size) - 1u);
    if (k == 0) {
        const uint32_t e = T.dc[chroma][size];
        bits = ((unsigned long long)(e & 0xffffu) << size) | mag;
        len = (e >> 16) + size;
        return;
    }
    const unsigned long long below = M & ((1ull << k) - 1ull);
    const uint32_t pk = below ? top_bit64(below) : 0u;  // the nonzero before it (0: the DC position)
    uint32_t run = k - 1u - pk;
    const uint32_t zrl = T.ac[chroma][0xf0];
    for (; run > 15; run -= 16) {
        bits = (bits << (zrl >> 16)) | (zrl & 0xffffu);
        len += zrl >> 16;
    }
    const uint32_t e = T.ac[chroma][(run << 4) | size];
    bits = (((bits << (e >> 16)) | (e & 0xffffu)) << size) | mag;
    len += (e >> 16) + size;
}

// Block b of the frame in MCU order (YCbCr 4:2:0: slots Y00 Y10 Y01 Y11 Cb Cr): which block of the plane-raster coefficient
// buffer it is, whether it is one of libjpeg's dummy blocks (a luma block of an edge MCU wholly outside the image: AC zero,
// DC = the DC of the block coded before it in the MCU), and its DC difference against the block coded before it in the
// same component.
struct BlockSrc {
    uint32_t src;
    bool dummy;
    int chroma;
    int dc_diff;
};
HE_FN BlockSrc block_source(const JpegEncFrame& F, const int16_t* __restrict__ coefs, uint32_t b) {
    const uint32_t mcu = b / 6, s = b - mcu * 6, my = mcu / F.mcus_x, mx = mcu - my * F.mcus_x;
    BlockSrc R;
    R.chroma = s >= 4;
    R.dummy = false;
    auto y_blk = [&](uint32_t mx_, uint32_t my_, uint32_t s_, bool& dummy) -> uint32_t {
        const uint32_t bx = 2 * mx_ + (s_ & 1u), by = 2 * my_ + (s_ >> 1);
        dummy = bx >= F.wib0 || by >= F.hib0;
        return by * F.y_bw + bx;
    };
    auto y_eff_dc = [&](uint32_t mx_, uint32_t my_, uint32_t s_) -> int {  // DC as coded: a dummy block repeats the block before it
        bool d;
        uint32_t blk = y_blk(mx_, my_, s_, d);
        while (d && s_ > 0) blk = y_blk(mx_, my_, --s_, d);
        return coefs[(size_t)blk * 64];
    };
    int dc, pred = 0;
    if (!R.chroma) {
        R.src = y_blk(mx, my, s, R.dummy);
        dc = y_eff_dc(mx, my, s);
        if (s > 0) pred = y_eff_dc(mx, my, s - 1);
        else if (mcu > 0) { const uint32_t pm = mcu - 1, py = pm / F.mcus_x; pred = y_eff_dc(pm - py * F.mcus_x, py, 3); }
    } else {
        const uint32_t base = s == 4 ? F.cb_off : F.cr_off;
        R.src = base + my * F.c_bw + mx;
        dc = coefs[(size_t)R.src * 64];
        if (mcu > 0) { const uint32_t pm = mcu - 1, py = pm / F.mcus_x; pred = coefs[(size_t)(base + py * F.c_bw + (pm - py * F.mcus_x)) * 64]; }
    }
    R.dc_diff = dc - pred;
    return R;
}

// byte i of the bit buffer of `bits` bits, the last byte padded with 1-bits (libjpeg's flush)
HE_FN uint32_t packed_byte(const uint32_t* __restrict__ P, uint32_t i, uint32_t bits) {
    const uint32_t nbytes = (bits + 7) / 8, rem = bits & 7u;
    uint32_t v = (P[i >> 2] >> (24u - 8u * (i & 3u))) & 0xffu;
    if (i == nbytes - 1 && rem) v |= (1u << (8u - rem)) - 1u;
    return v;
}

}  // namespace he
}  // namespace uf
