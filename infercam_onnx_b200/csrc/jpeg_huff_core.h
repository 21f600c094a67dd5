// jpeg_huff_core.h — the symbol-level Huffman decoder of kernels_jpeg_huff.cu, written so that the same lines compile for
// the device (nvcc) and, for the CPU-only test that steps the synchronisation rounds serially (tests/helpers/huff_sim.cc),
// for the host. Nothing in the library calls the host build.
#pragma once
#include <cstddef>
#include <cstdint>

#include "jpeg_decode.h"

#ifdef __CUDACC__
#define JH_FN __device__ __forceinline__
#define JH_CONST __constant__
#else
#define JH_FN inline
#define JH_CONST static const
#ifndef __restrict__
#define __restrict__
#endif
#endif

namespace uf {
namespace jh {

JH_CONST uint8_t c_zigzag[80] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                                     6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                                     39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct Tabs {  // the frame's six tables in shared memory
    JpegHuffTab t[6];  // dc[c] = t[c], ac[c] = t[3 + c]
};

JH_FN unsigned long long pack_state(uint32_t p, uint32_t slot, uint32_t k) {
    return ((unsigned long long)p << 32) | (slot << 8) | k;
}

// 32 bits of the stream starting at bit p (big-endian bit order); d is 4-byte aligned and zero-padded past the end
JH_FN uint32_t window(const uint32_t* __restrict__ d, uint32_t p) {
#ifdef __CUDA_ARCH__
    const uint32_t w0 = __byte_perm(d[p >> 5], 0, 0x0123), w1 = __byte_perm(d[(p >> 5) + 1], 0, 0x0123);
    return __funnelshift_l(w1, w0, p & 31);
#else
    const uint8_t* q = reinterpret_cast<const uint8_t*>(d) + (size_t)(p >> 5) * 4;
    uint64_t v = 0;
    for (int i = 0; i < 8; ++i) v = (v << 8) | q[i];
    return (uint32_t)(v >> (32 - (p & 31)));
#endif
}

JH_FN int huff_symbol(const JpegHuffTab& T, uint32_t win, uint32_t& p) {
    const uint32_t e = T.look[win >> (32 - JH_LOOK)];
    if (e) {
        p += e >> 8;
        return (int)(e & 0xff);
    }
    int l = JH_LOOK + 1;
    int code = (int)(win >> (32 - l));
    while (l <= 16 && code > T.maxcode[l]) {
        ++l;
        code = (int)(win >> (32 - l));
    }
    if (l > 16) {  // not a code of this table (wrong guess or damaged data): as the host decoder, 16 bits, symbol 0
        p += 16;
        return 0;
    }
    p += l;
    return T.vals[(code + T.valoff[l]) & 0xff];
}

JH_FN int extend(uint32_t v, int s) { return v < (1u << (s - 1)) ? (int)v - (int)((1u << s) - 1) : (int)v; }

// Decodes from state (p, slot, k) while p < p_end. WRITE: coefficients go to `coefs` (dense blocks of the frame; DC as the
// difference), `next_block` = index of the next block to start, and decoding stops for good once `nblocks` blocks are done
// (the bits after the last block are padding). Returns the number of blocks started.
template <bool WRITE>
JH_FN uint32_t huff_run(const Tabs& tabs, const JpegHuffFrame& fr, const uint32_t* __restrict__ d, uint32_t& p,
                                             uint32_t& slot, uint32_t& k, uint32_t p_end, int16_t* __restrict__ coefs, uint32_t next_block) {
    uint32_t started = 0;
    const uint32_t bpm = fr.blocks_per_mcu;
    int16_t* blk = WRITE && k > 0 && next_block > 0 && next_block - 1 < fr.nblocks ? coefs + (size_t)(next_block - 1) * 64 : nullptr;
    while (p < p_end) {
        const int c = fr.slot_comp[slot];
        uint32_t win = window(d, p);
        if (k == 0) {  // DC
            if (WRITE && next_block >= fr.nblocks) break;  // every block of the frame is done: the rest is padding
            const int s = huff_symbol(tabs.t[c], win, p) & 15;
            int diff = 0;
            if (s) {
                diff = extend(window(d, p) >> (32 - s), s);
                p += s;
            }
            if (WRITE) {
                blk = coefs + (size_t)next_block * 64;
                blk[0] = (int16_t)diff;
                ++next_block;
            }
            ++started;
            k = 1;
        } else {
            const int rs = huff_symbol(tabs.t[3 + c], win, p);
            const int r = rs >> 4, s = rs & 15;
            if (s == 0) {
                k = r == 15 ? k + 16 : 64;  // ZRL / EOB
            } else {
                k += r;
                const int v = extend(window(d, p) >> (32 - s), s);
                p += s;
                if (WRITE && blk) blk[c_zigzag[k < 80 ? k : 79]] = (int16_t)v;
                ++k;
            }
            if (k >= 64) {  // block complete
                k = 0;
                slot = slot + 1 == bpm ? 0 : slot + 1;
            }
        }
    }
    return started;
}

}  // namespace jh
}  // namespace uf
