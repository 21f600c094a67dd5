// jpeg_huff_core.h — the symbol-level Huffman decoder of kernels_jpeg_huff.cu, written so that the same lines compile for
// the device (nvcc) and, for the CPU-only test that steps the synchronisation rounds serially (tests/helpers/huff_sim.cc),
// for the host. Nothing in the library calls the host build.
//
// Shaped for a GPU thread walking 128 bytes on its own: the upcoming bits live in a 64-bit register that is topped up one
// 32-bit word at a time from a word fetched a refill EARLIER (the load is off the dependent chain), one shared-memory lookup
// resolves a code of up to JH_LOOK bits, longer codes take a branch-free count over left-aligned length limits, and DC and
// AC symbols run through the same instructions (a warp's lanes are at different places in their blocks).
#pragma once
#include <cstddef>
#include <cstdint>

#include "jpeg_decode.h"

#ifdef __CUDACC__
#define JH_FN __device__ __forceinline__
#else
#define JH_FN inline
#ifndef __restrict__
#define __restrict__
#endif
#endif

namespace uf {
namespace jh {

struct Tabs {          // per CTA, in shared memory
    JpegHuffTabSet set;  // dc[c], ac[c]
    uint8_t zz[80];    // zigzag position -> natural index; 64..79 (overrun on damaged data) -> 63, as libjpeg's table
};

JH_FN uint8_t zigzag_natural(int i) {  // used to fill Tabs::zz
    const uint8_t z[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                           41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                           30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    return i < 64 ? z[i] : 63;
}

JH_FN unsigned long long pack_state(uint32_t p, uint32_t slot, uint32_t k) {
    return ((unsigned long long)p << 32) | (slot << 8) | k;
}

// word i of the stream, most significant bit first (d is 4-byte aligned and zero-padded past the end)
JH_FN uint32_t be32(const uint32_t* __restrict__ d, uint32_t i) {
#ifdef __CUDA_ARCH__
    return __byte_perm(d[i], 0, 0x0123);
#else
    const uint8_t* q = reinterpret_cast<const uint8_t*>(d) + (size_t)i * 4;
    return ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
#endif
}

JH_FN int extend(uint32_t v, uint32_t s) { return v < (1u << (s - 1)) ? (int)v - (int)((1u << s) - 1) : (int)v; }

// A code longer than JH_LOOK bits (x = the next 16 bits): its length = 1 + the number of length limits x reaches. Returns
// the symbol's look entry (jpeg_decode.h). Not a code of this table (wrong guess or damaged data): as the host decoder,
// 16 bits, symbol 0.
JH_FN uint32_t long_symbol(const JpegHuffTab& T, uint32_t x, bool dc) {
    uint32_t len = JH_LOOK + 1;
#pragma unroll
    for (int l = JH_LOOK + 1; l < 16; ++l) len += x >= T.lim[l];
    uint32_t sym = 0;
    if (x >= T.lim[16]) len = 16;
    else sym = T.vals[((x >> (16 - len)) + (uint32_t)T.valoff[len]) & 0xffu];
    const uint32_t s = sym & 15u, r = sym >> 4;
    const uint32_t kadv = dc ? 1u : (s ? r + 1 : (r == 15 ? 16u : 64u));
    return (len + s) | (kadv << 5) | (s << 12);
}

// What the write pass produces, per frame (the layout the IDCT kernel reads, kernels_jpeg.cu):
struct HuffOut {
    uint32_t* offs;     // [nblocks + 1] index of each block's first AC entry
    uint32_t* entries;  // (natural index << 16) | (uint16) value of the nonzero AC coefficients, in decode order
    int16_t* dcv;       // [nblocks] DC difference of each block (summed into DC values by jhuff_dc_kernel)
    uint32_t ent_cap;   // entries the frame's list can hold (an unsettled frame must not write past it)
};

// Decodes from state (p, slot, k) while p < p_end, counting the blocks started (low 16 bits of the result) and the nonzero
// AC coefficients (high 16 bits). WRITE: those go to `out`; next_block / next_entry = where this subsequence's first block
// start / first entry land, and decoding stops for good once `nblocks` blocks are done (the bits after the last block are
// padding). slotmap: 2 bits per block slot of the MCU = its component.
template <bool WRITE>
JH_FN uint32_t huff_run(const Tabs& tabs, uint32_t slotmap, uint32_t bpm, uint32_t nblocks, const uint32_t* __restrict__ d, uint32_t& p_io,
                        uint32_t& slot_io, uint32_t& k_io, uint32_t p_end, const HuffOut& out, uint32_t next_block, uint32_t next_entry) {
    uint32_t p = p_io, slot = slot_io, k = k_io, counts = 0;
    bool in_block = WRITE && k > 0 && next_block > 0 && next_block - 1 < nblocks;  // the block in progress is one of the frame's
    // acc: the next nb bits of the stream, left-aligned; `ahead` = word wi - 1, already fetched
    uint32_t wi = p >> 5;
    unsigned long long acc = (((unsigned long long)be32(d, wi) << 32) | be32(d, wi + 1)) << (p & 31);
    int nb = 64 - (int)(p & 31);
    uint32_t ahead = be32(d, wi + 2);
    wi += 3;
    uint32_t comp = (slotmap >> (2 * slot)) & 3u;
    const JpegHuffTab* Tdc = &tabs.set.dc[comp];
    const JpegHuffTab* Tac = &tabs.set.ac[comp];
    while (p < p_end) {
        if (nb <= 32) {
            acc |= (unsigned long long)ahead << (32 - nb);
            nb += 32;
            ahead = be32(d, wi);
            ++wi;
        }
        const bool dc = k == 0;
        if (WRITE && dc && next_block >= nblocks) break;  // every block of the frame is done: the rest is padding
        const JpegHuffTab& T = dc ? *Tdc : *Tac;
        const uint32_t hi = (uint32_t)(acc >> 32);
        uint32_t e = T.look[hi >> (32 - JH_LOOK)];
        if (e == 0) e = long_symbol(T, hi >> 16, dc);
        const uint32_t total = e & 31u, kadv = (e >> 5) & 127u, s = e >> 12;
        if (WRITE) {
            const int v = s ? extend((uint32_t)((acc << (total - s)) >> (64 - s)), s) : 0;
            if (dc) {
                out.offs[next_block] = next_entry;
                out.dcv[next_block] = (int16_t)v;
                ++next_block;
                in_block = true;
            } else if (s != 0) {
                const uint32_t at = k + kadv - 1;
                if (in_block && next_entry < out.ent_cap) out.entries[next_entry] = ((uint32_t)tabs.zz[at < 80 ? at : 79] << 16) | (uint32_t)(uint16_t)(int16_t)v;
                ++next_entry;
            }
        }
        counts += dc ? 1u : (s != 0 ? 0x10000u : 0u);
        acc <<= total;
        nb -= (int)total;
        p += total;
        k += kadv;
        if (k >= 64) {  // block complete
            k = 0;
            slot = slot + 1 == bpm ? 0 : slot + 1;
            comp = (slotmap >> (2 * slot)) & 3u;
            Tdc = &tabs.set.dc[comp];
            Tac = &tabs.set.ac[comp];
        }
    }
    p_io = p;
    slot_io = slot;
    k_io = k;
    return counts;
}

}  // namespace jh
}  // namespace uf
