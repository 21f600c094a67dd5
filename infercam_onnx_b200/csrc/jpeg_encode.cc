// jpeg_encode.cc — host half of the JPEG encoder behind uf_annotate_encode_jpeg: quantisation
// tables for a quality setting, Huffman coding of the quantised blocks the GPU produced, and the JFIF file around them.
// Replaces `turbojpeg::compress_image(&frame, 95, Subsamp::Sub2x2)` (/root/reference/infer_server/src/inferer.rs:39):
// baseline, YCbCr 4:2:0, the Annex K Huffman tables, `jpeg_set_quality(q, force_baseline)` scaling — what libjpeg-turbo's
// compressor does with default settings. Written from the JPEG standard (T.81 Annex F.1.2 Huffman encoding, Annex K tables,
// B.2 marker syntax) and libjpeg's documented conventions (dummy blocks at the right / bottom edge repeat the previous
// block's DC, jccoefct.c); the sample-domain half (colour conversion, downsampling, forward DCT, quantisation) is in
// kernels_jpeg_enc.cu. Checked against libjpeg-turbo itself: the coefficients and tables of the file PIL writes from the
// same pixels must equal the ones in ours (tests/test_jpeg_encode.py).
#include <cstring>

#include "../../include/ultraface_b200.h"
#include "jpeg_decode.h"

namespace uf {
namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// Annex K.1 / K.2, natural (row-major) order
const uint8_t kStdLumQuant[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,  14, 13, 16, 24, 40,  57,
                                  69, 56, 14, 17, 22,  29,  51,  87,  80, 62, 18, 22, 37,  56,  68,  109, 103, 77, 24, 35, 55,  64,
                                  81, 104, 113, 92, 49, 64,  78,  87,  103, 121, 120, 101, 72,  92,  95,  98,  112, 100, 103, 99};
const uint8_t kStdChrQuant[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                  99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                  99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

// Annex K.3 (same tables as the decoder's MJPG defaults)
const uint8_t kDcLumBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChrBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125};
const uint8_t kAcLumVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1,
    0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56,
    0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85,
    0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa,
    0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
    0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};
const uint8_t kAcChrBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119};
const uint8_t kAcChrVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42,
    0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19,
    0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55,
    0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8,
    0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4,
    0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};

struct EncTable {
    uint16_t code[256];
    uint8_t len[256];
    EncTable(const uint8_t* bits, const uint8_t* vals) {
        memset(code, 0, sizeof(code));
        memset(len, 0, sizeof(len));
        int c = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            for (int i = 0; i < bits[l - 1]; ++i, ++k, ++c) {
                code[vals[k]] = (uint16_t)c;
                len[vals[k]] = (uint8_t)l;
            }
            c <<= 1;
        }
    }
};

struct BitWriter {
    std::vector<uint8_t>& out;
    uint64_t acc = 0;
    int n = 0;
    explicit BitWriter(std::vector<uint8_t>& o) : out(o) {}
    inline void put(uint32_t bits, int len) {
        acc = (acc << len) | (bits & ((1u << len) - 1u));
        n += len;
        while (n >= 8) {
            const uint8_t b = (uint8_t)(acc >> (n - 8));
            out.push_back(b);
            if (b == 0xff) out.push_back(0x00);  // byte stuffing
            n -= 8;
        }
    }
    void flush() {  // pad the last byte with 1-bits
        if (n > 0) put(0x7f, 8 - n);
    }
};

inline int bit_length(int v) {
    int n = 0;
    while (v) { ++n; v >>= 1; }
    return n;
}

void encode_block(BitWriter& bw, const int16_t* blk, int& last_dc, const EncTable& dc, const EncTable& ac) {
    int temp = blk[0] - last_dc, temp2 = temp;
    last_dc = blk[0];
    if (temp < 0) { temp = -temp; --temp2; }
    int nbits = bit_length(temp);
    bw.put(dc.code[nbits], dc.len[nbits]);
    if (nbits) bw.put((uint32_t)temp2, nbits);
    int r = 0;
    for (int k = 1; k < 64; ++k) {
        temp = blk[kZigzag[k]];
        if (temp == 0) { ++r; continue; }
        while (r > 15) { bw.put(ac.code[0xf0], ac.len[0xf0]); r -= 16; }
        temp2 = temp;
        if (temp < 0) { temp = -temp; --temp2; }
        nbits = bit_length(temp);
        const int sym = (r << 4) | nbits;
        bw.put(ac.code[sym], ac.len[sym]);
        bw.put((uint32_t)temp2, nbits);
        r = 0;
    }
    if (r > 0) bw.put(ac.code[0], ac.len[0]);
}

void put16(std::vector<uint8_t>& o, uint32_t v) { o.push_back((uint8_t)(v >> 8)); o.push_back((uint8_t)v); }

void put_dht(std::vector<uint8_t>& o, int index, const uint8_t* bits, const uint8_t* vals, int n) {
    o.push_back(0xff); o.push_back(0xc4);
    put16(o, 2 + 1 + 16 + n);
    o.push_back((uint8_t)index);
    o.insert(o.end(), bits, bits + 16);
    o.insert(o.end(), vals, vals + n);
}

}  // namespace

// jpeg_set_quality(quality, force_baseline = TRUE): scale factor 5000/q below 50, 200 - 2q from 50 up; entries clamped to 1..255
void jpeg_quality_tables(int quality, uint16_t lum[64], uint16_t chr[64]) {
    quality = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
    const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    for (int i = 0; i < 64; ++i) {
        long l = ((long)kStdLumQuant[i] * scale + 50) / 100, c = ((long)kStdChrQuant[i] * scale + 50) / 100;
        lum[i] = (uint16_t)(l < 1 ? 1 : (l > 255 ? 255 : l));
        chr[i] = (uint16_t)(c < 1 ? 1 : (c > 255 ? 255 : c));
    }
}

// Geometry of a w x h YCbCr 4:2:0 frame as the encoder lays it out (same struct as the decoder's).
JpegPlan jpeg_encode_plan(uint32_t w, uint32_t h, int quality) {
    JpegPlan p{};
    p.w = w; p.h = h; p.ncomp = 3; p.hmax = p.vmax = 2;
    p.hs[0] = p.vs[0] = 2; p.hs[1] = p.vs[1] = p.hs[2] = p.vs[2] = 1;
    p.mcus_x = (w + 15) / 16; p.mcus_y = (h + 15) / 16;
    uint32_t off = 0;
    for (uint32_t c = 0; c < 3; ++c) {
        for (uint32_t v = 0; v < p.vs[c]; ++v)
            for (uint32_t hh = 0; hh < p.hs[c]; ++hh) {
                p.slot_comp[p.blocks_per_mcu] = (uint8_t)c; p.slot_h[p.blocks_per_mcu] = (uint8_t)hh; p.slot_v[p.blocks_per_mcu] = (uint8_t)v;
                ++p.blocks_per_mcu;
            }
        p.plane_w[c] = p.mcus_x * p.hs[c] * 8; p.plane_h[c] = p.mcus_y * p.vs[c] * 8;
        p.real_w[c] = (w * p.hs[c] + 1) / 2; p.real_h[c] = (h * p.vs[c] + 1) / 2;
        p.plane_off[c] = off;
        off += p.plane_w[c] * p.plane_h[c];
    }
    p.plane_bytes = (off + 255) / 256 * 256;
    p.nblocks = p.mcus_x * p.mcus_y * p.blocks_per_mcu;
    jpeg_quality_tables(quality, p.quant[0], p.quant[1]);
    memcpy(p.quant[2], p.quant[1], sizeof(p.quant[1]));
    return p;
}

// SOI, JFIF APP0, DQT x 2, SOF0, DHT x 4, SOS
void jpeg_write_headers(const JpegPlan& p, std::vector<uint8_t>& out) {
    out.clear();
    out.reserve((size_t)p.w * p.h / 2 + 1024);
    const uint8_t head[] = {0xff, 0xd8, 0xff, 0xe0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};  // JFIF 1.01, aspect 1:1
    out.insert(out.end(), head, head + sizeof(head));
    for (int t = 0; t < 2; ++t) {
        out.push_back(0xff); out.push_back(0xdb);
        put16(out, 67);
        out.push_back((uint8_t)t);
        for (int k = 0; k < 64; ++k) out.push_back((uint8_t)p.quant[t][kZigzag[k]]);
    }
    out.push_back(0xff); out.push_back(0xc0);
    put16(out, 17);
    out.push_back(8);
    put16(out, p.h); put16(out, p.w);
    out.push_back(3);
    const uint8_t comps[9] = {1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1};
    out.insert(out.end(), comps, comps + 9);
    put_dht(out, 0x00, kDcLumBits, kDcVals, 12);
    put_dht(out, 0x10, kAcLumBits, kAcLumVals, 162);
    put_dht(out, 0x01, kDcChrBits, kDcVals, 12);
    put_dht(out, 0x11, kAcChrBits, kAcChrVals, 162);
    const uint8_t sos[] = {0xff, 0xda, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
    out.insert(out.end(), sos, sos + sizeof(sos));
}

void jpeg_std_enc_tables(JpegEncTables& t) {
    static const EncTable dc_l(kDcLumBits, kDcVals), dc_c(kDcChrBits, kDcVals), ac_l(kAcLumBits, kAcLumVals), ac_c(kAcChrBits, kAcChrVals);
    memset(&t, 0, sizeof(t));
    for (int i = 0; i < 12; ++i) {
        t.dc[0][i] = ((uint32_t)dc_l.len[i] << 16) | dc_l.code[i];
        t.dc[1][i] = ((uint32_t)dc_c.len[i] << 16) | dc_c.code[i];
    }
    for (int i = 0; i < 256; ++i) {
        t.ac[0][i] = ((uint32_t)ac_l.len[i] << 16) | ac_l.code[i];
        t.ac[1][i] = ((uint32_t)ac_c.len[i] << 16) | ac_c.code[i];
    }
}

// coefs: quantised blocks per component PLANE in raster order ([plane_off[c] / 64 + by * (plane_w[c] / 8) + bx][64], natural
// order inside a block), every block of the padded planes computed from edge-replicated samples. Blocks wholly outside the
// component (libjpeg's dummy blocks: beyond ceil(real_w / 8) in a row, beyond ceil(real_h / 8) rows) are replaced here by
// libjpeg's rule: all AC zero, DC = the DC of the block before it in the MCU.
void jpeg_write_file(const JpegPlan& p, const int16_t* coefs, std::vector<uint8_t>& out) {
    static const EncTable dc_l(kDcLumBits, kDcVals), dc_c(kDcChrBits, kDcVals), ac_l(kAcLumBits, kAcLumVals), ac_c(kAcChrBits, kAcChrVals);
    jpeg_write_headers(p, out);
    BitWriter bw(out);
    int last_dc[3] = {0, 0, 0};
    uint32_t wib[3], hib[3];
    for (int c = 0; c < 3; ++c) { wib[c] = (p.real_w[c] + 7) / 8; hib[c] = (p.real_h[c] + 7) / 8; }
    int16_t dummy[64];
    for (uint32_t my = 0; my < p.mcus_y; ++my)
        for (uint32_t mx = 0; mx < p.mcus_x; ++mx) {
            int prev_dc_in_mcu = 0;  // DC of the block coded just before (libjpeg: MCU_buffer[blkn - 1][0][0])
            for (uint32_t sl = 0; sl < p.blocks_per_mcu; ++sl) {
                const int c = p.slot_comp[sl];
                const uint32_t bx = mx * p.hs[c] + p.slot_h[sl], by = my * p.vs[c] + p.slot_v[sl];
                const int16_t* blk = coefs + ((size_t)p.plane_off[c] / 64 + (size_t)by * (p.plane_w[c] / 8) + bx) * 64;
                if (bx >= wib[c] || by >= hib[c]) {
                    memset(dummy, 0, sizeof(dummy));
                    dummy[0] = (int16_t)prev_dc_in_mcu;
                    blk = dummy;
                }
                prev_dc_in_mcu = blk[0];
                encode_block(bw, blk, last_dc[c], c ? dc_c : dc_l, c ? ac_c : ac_l);
            }
        }
    bw.flush();
    out.push_back(0xff); out.push_back(0xd9);
}

}  // namespace uf
