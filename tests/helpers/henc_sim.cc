// henc_sim.cc — CPU-only test helper (never part of the library): walks the blocks of one frame through the device Huffman
// coder's plan serially — bit count per block, prefix sum, bits OR-ed into the big-endian bit buffer at their offsets, last
// byte padded, 0x00 after 0xFF — over the SAME per-block code the kernels compile (csrc/jpeg_henc_core.h: block source and
// dummy-block rule, DC predictor, the bits of a coefficient, put64), so that it can be checked against the sequential
// writer (jpeg_write_file) without a GPU. "Lane l" is a loop index; the ballot is a loop over the 64 positions.
#include <cstring>
#include <vector>

#include "../../infercam_onnx_b200/csrc/jpeg_decode.h"
#include "../../infercam_onnx_b200/csrc/jpeg_henc_core.h"

using namespace uf;
using namespace uf::he;

static const uint8_t kZz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// coefs: quantised blocks per component plane in raster order, as uf_jpeg_write_coefficients takes them
extern "C" int henc_sim(uint32_t w, uint32_t h, uint32_t quality, const int16_t* coefs, uint8_t* out, size_t cap, size_t* out_len) {
    const JpegPlan p = jpeg_encode_plan(w, h, (int)quality);
    JpegEncTables tabs;
    jpeg_std_enc_tables(tabs);
    EncTabs T;
    memcpy(&T, &tabs, sizeof(tabs));
    memcpy(T.zz, kZz, 64);
    JpegEncFrame F{};  // (as engine.cu fills it)
    F.coef_base = 0;
    F.mcus_x = p.mcus_x; F.mcus_y = p.mcus_y;
    F.y_bw = p.plane_w[0] / 8; F.c_bw = p.plane_w[1] / 8;
    F.cb_off = p.plane_off[1] / 64; F.cr_off = p.plane_off[2] / 64;
    F.wib0 = (p.real_w[0] + 7) / 8; F.hib0 = (p.real_h[0] + 7) / 8;
    F.nblocks = p.nblocks;
    F.pack_cap_bits = p.nblocks * 24 * 32;
    std::vector<uint32_t> bitlen(F.nblocks), bitoff(F.nblocks), packed((size_t)F.nblocks * 24 + 2, 0);
    for (int pass = 0; pass < 2; ++pass) {
        for (uint32_t b = 0; b < F.nblocks; ++b) {
            const BlockSrc S = block_source(F, coefs, b);
            const int16_t* blk = coefs + (size_t)S.src * 64;
            int v[64];
            unsigned long long M = 0;
            for (uint32_t k = 0; k < 64; ++k) {
                v[k] = k == 0 ? S.dc_diff : (S.dummy ? 0 : (int)blk[T.zz[k]]);
                if (k > 0 && v[k] != 0) M |= 1ull << k;
            }
            uint32_t at = pass ? bitoff[b] : 0, total = 0;
            for (uint32_t k = 0; k < 64; ++k) {  // lanes 0..31 first half, then second half: position order either way
                unsigned long long bits;
                uint32_t len;
                coef_bits(T, S.chroma, k, v[k], M, bits, len);
                if (pass) put64(packed.data(), bits, len, at + total, F.pack_cap_bits);
                total += len;
            }
            const uint32_t last = M ? top_bit64(M) : 0u;
            const uint32_t eob = last < 63 ? T.ac[S.chroma][0] : 0u;
            if (pass && eob) put64(packed.data(), eob & 0xffffu, eob >> 16, at + total, F.pack_cap_bits);
            total += eob >> 16;
            if (!pass) bitlen[b] = total;
        }
        if (!pass) {
            uint32_t run = 0;
            for (uint32_t b = 0; b < F.nblocks; ++b) { bitoff[b] = run; run += bitlen[b]; }
        }
    }
    const uint32_t bits = F.nblocks ? bitoff[F.nblocks - 1] + bitlen[F.nblocks - 1] : 0;
    if (bits > F.pack_cap_bits) return 2;
    std::vector<uint8_t> file;
    jpeg_write_headers(p, file);
    const uint32_t nbytes = (bits + 7) / 8;
    for (uint32_t i = 0; i < nbytes; ++i) {
        const uint32_t v = packed_byte(packed.data(), i, bits);
        file.push_back((uint8_t)v);
        if (v == 0xffu) file.push_back(0);
    }
    file.push_back(0xff);
    file.push_back(0xd9);
    *out_len = file.size();
    if (file.size() > cap) return 1;
    memcpy(out, file.data(), file.size());
    return 0;
}
