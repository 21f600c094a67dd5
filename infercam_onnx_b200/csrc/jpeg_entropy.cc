// jpeg_entropy.cc — header parsing and Huffman decoding of baseline JPEG on the host (see jpeg_decode.h).
// Written from the JPEG standard (ITU-T T.81: marker syntax B.2, Huffman decoding F.2.2, the Annex K.3 tables that
// Motion-JPEG streams omit) to produce exactly the coefficient blocks libjpeg-turbo's decoder produces, including its
// conventions for damaged data (zero bits after a marker, coefficient indices beyond 63 land on 63).
#include <cstring>

#include "../../include/ultraface_b200.h"
#include "jpeg_decode.h"

namespace uf {
namespace {

const uint8_t kZigzag[64 + 16] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                                  6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                                  39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
                                  63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};  // a run past the block ends on 63

// Annex K.3 (the tables every encoder writes unless it optimises; MJPG webcams leave the DHT segment out altogether)
const uint8_t kStdDcLumBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kStdDcChrBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kStdDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kStdAcLumBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125};
const uint8_t kStdAcLumVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1,
    0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56,
    0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85,
    0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa,
    0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
    0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};
const uint8_t kStdAcChrBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119};
const uint8_t kStdAcChrVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42,
    0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19,
    0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55,
    0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8,
    0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4,
    0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};

constexpr int LOOK = 9;  // bits resolved by one table lookup

struct HuffTable {
    bool defined = false;
    uint8_t bits[16] = {};
    uint8_t vals[256] = {};
    // derived
    uint16_t look[1 << LOOK];   // (length << 8) | symbol, 0 = longer than LOOK bits
    int16_t fast_ac[1 << LOOK]; // AC tables: code AND value bits inside LOOK bits -> (value << 8) | (run << 4) | total bits; 0 = no
    int32_t maxcode[18];        // largest code of each length (-1: none), [17] = sentinel
    int32_t valoff[17];         // vals index = code + valoff[length]

    void derive() {
        int code = 0, k = 0;
        memset(look, 0, sizeof(look));
        for (int l = 1; l <= 16; ++l) {
            valoff[l] = k - code;
            for (int i = 0; i < bits[l - 1]; ++i, ++k, ++code) {
                if (l <= LOOK) {
                    const int first = code << (LOOK - l);
                    for (int f = 0; f < (1 << (LOOK - l)); ++f) look[first + f] = (uint16_t)((l << 8) | vals[k]);
                }
            }
            maxcode[l] = bits[l - 1] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        // one lookup decodes a whole (run, value) pair when the code and its magnitude bits fit in LOOK bits and the value in a byte
        for (int i = 0; i < (1 << LOOK); ++i) {
            fast_ac[i] = 0;
            const uint16_t e = look[i];
            if (!e) continue;
            const int len = e >> 8, rs = e & 0xff, run = rs >> 4, mag = rs & 15;
            if (mag == 0 || len + mag > LOOK) continue;
            int k = ((i << len) & ((1 << LOOK) - 1)) >> (LOOK - mag);
            if (k < (1 << (mag - 1))) k -= (1 << mag) - 1;
            if (k >= -128 && k <= 127) fast_ac[i] = (int16_t)((k * 256) + (run * 16) + (len + mag));
        }
    }
    // derive_now = false: only bits / vals are wanted (the device decoder builds its own tables from them)
    void set(const uint8_t* b, const uint8_t* v, int n, bool derive_now = true) {
        memcpy(bits, b, 16);
        memset(vals, 0, sizeof(vals));
        memcpy(vals, v, (size_t)n);
        defined = true;
        if (derive_now) derive();
    }
};

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc = 0;  // bits are consumed from the top
    int nbits = 0;
    int real = 0;             // how many of the nbits are data (the rest is the zero fill behind a marker / the end)
    bool hit_marker = false;  // data exhausted or a marker met: zero bits from here on (libjpeg's behaviour)
    bool insufficient = false;  // zero-fill bits have been CONSUMED: libjpeg's `insufficient_data` — the MCU in progress is
                                // finished with zeros, every following MCU of the segment is left empty (uniform grey)

    void refill() {
        while (nbits <= 56) {
            uint32_t b = 0;
            if (!hit_marker) {
                if (p >= end) {
                    hit_marker = true;
                } else if (*p != 0xff) {
                    b = *p++;
                    real += 8;
                } else if (p + 1 < end && p[1] == 0x00) {  // stuffed zero
                    b = 0xff;
                    p += 2;
                    real += 8;
                } else if (p + 1 < end && p[1] == 0xff) {  // fill byte before a marker
                    ++p;
                    continue;
                } else {
                    hit_marker = true;  // a real marker (RSTn / EOI / ...) stays unread
                }
            }
            acc |= (uint64_t)b << (56 - nbits);
            nbits += 8;
        }
    }
    inline uint32_t peek(int n) { return (uint32_t)(acc >> (64 - n)); }
    inline void skip(int n) {
        acc <<= n;
        nbits -= n;
        real -= n;
        if (real < 0) { real = 0; insufficient = true; }
    }
    inline uint32_t get(int n) {
        const uint32_t v = peek(n);
        skip(n);
        return v;
    }
    void reset_at_restart() {
        acc = 0;
        nbits = 0;
        real = 0;
        hit_marker = false;
        insufficient = false;
    }
};

inline int decode_symbol(BitReader& br, const HuffTable& t) {
    if (br.nbits < 32) br.refill();
    const uint16_t e = t.look[br.peek(LOOK)];
    if (e) {
        br.skip(e >> 8);
        return e & 0xff;
    }
    int l = LOOK + 1;
    int32_t code = (int32_t)br.peek(l);
    while (l <= 16 && code > t.maxcode[l]) {
        ++l;
        code = (int32_t)br.peek(l);
    }
    if (l > 16) {  // not a code of this table (corrupt data): libjpeg warns and returns 0
        br.skip(16);
        return 0;
    }
    br.skip(l);
    return t.vals[(code + t.valoff[l]) & 0xff];
}

inline int extend(uint32_t v, int s) { return v < (1u << (s - 1)) ? (int)v - (int)((1u << s) - 1) : (int)v; }

[[noreturn]] void fail(int code, const std::string& msg) { throw JpegError{code, "jpeg: " + msg}; }

struct Parsed {
    JpegPlan plan{};
    HuffTable dc[4], ac[4];
    uint16_t qt[4][64];  // natural order
    bool qt_defined[4] = {false, false, false, false};
    int tq[3] = {0, 0, 0}, td[3] = {0, 0, 0}, ta[3] = {0, 0, 0};
    int comp_id[3] = {0, 0, 0};
    uint32_t restart_interval = 0;
    size_t scan_start = 0;  // first byte of the entropy-coded segment
    bool have_sof = false;
};

uint32_t be16(const uint8_t* p) { return ((uint32_t)p[0] << 8) | p[1]; }

void parse(const uint8_t* d, size_t len, Parsed& P, bool want_scan, bool derive_tables = true) {
    if (len < 4 || d[0] != 0xff || d[1] != 0xd8) fail(UF_ERR_INVALID_ARG, "no SOI marker");
    size_t i = 2;
    for (;;) {
        while (i < len && d[i] != 0xff) ++i;  // garbage between segments is skipped, as libjpeg does
        while (i < len && d[i] == 0xff) ++i;
        if (i >= len) fail(UF_ERR_INVALID_ARG, "no SOS marker before the end of the data");
        const uint8_t m = d[i++];
        if (m == 0xd8 || (m >= 0xd0 && m <= 0xd7) || m == 0x01) continue;  // stand-alone markers
        if (m == 0xd9) fail(UF_ERR_INVALID_ARG, "EOI before any scan");
        if (i + 2 > len) fail(UF_ERR_INVALID_ARG, "truncated segment");
        const uint32_t L = be16(d + i);
        if (L < 2 || i + L > len) fail(UF_ERR_INVALID_ARG, "segment length exceeds the data");
        const uint8_t* s = d + i + 2;
        const uint32_t n = L - 2;
        if (m == 0xc0 || m == 0xc1) {  // SOF0 baseline / SOF1 extended sequential, Huffman
            if (n < 6) fail(UF_ERR_INVALID_ARG, "short SOF");
            if (s[0] != 8) fail(UF_ERR_UNSUPPORTED, "only 8-bit samples are supported");
            JpegPlan& p = P.plan;
            p.h = be16(s + 1);
            p.w = be16(s + 3);
            p.ncomp = s[5];
            if (p.w == 0 || p.h == 0) fail(UF_ERR_UNSUPPORTED, "image size 0 (DNL marker) is not supported");
            if (p.w > 16384 || p.h > 16384) fail(UF_ERR_UNSUPPORTED, "image larger than 16384 x 16384");
            if (p.ncomp != 1 && p.ncomp != 3) fail(UF_ERR_UNSUPPORTED, "only 1 or 3 components are supported");
            if (n < 6 + 3 * p.ncomp) fail(UF_ERR_INVALID_ARG, "short SOF");
            for (uint32_t c = 0; c < p.ncomp; ++c) {
                P.comp_id[c] = s[6 + 3 * c];
                p.hs[c] = s[7 + 3 * c] >> 4;
                p.vs[c] = s[7 + 3 * c] & 15;
                P.tq[c] = s[8 + 3 * c];
                if (p.hs[c] < 1 || p.vs[c] < 1 || p.hs[c] > 2 || p.vs[c] > 2 || P.tq[c] > 3) fail(UF_ERR_UNSUPPORTED, "sampling factors beyond 2");
            }
            P.have_sof = true;
        } else if (m == 0xc2 || (m >= 0xc5 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc) || m == 0xc3) {
            fail(UF_ERR_UNSUPPORTED, m == 0xc2 ? "progressive JPEG (the decode path is for baseline MJPG frames)" : "unsupported JPEG process (lossless / arithmetic / hierarchical)");
        } else if (m == 0xc4) {  // DHT
            uint32_t q = 0;
            while (q < n) {
                if (q + 17 > n) fail(UF_ERR_INVALID_ARG, "short DHT");
                const int tc = s[q] >> 4, th = s[q] & 15;
                if (tc > 1 || th > 3) fail(UF_ERR_INVALID_ARG, "bad DHT table id");
                int cnt = 0;
                for (int k = 0; k < 16; ++k) cnt += s[q + 1 + k];
                if (cnt > 256 || q + 17 + cnt > n) fail(UF_ERR_INVALID_ARG, "bad DHT counts");
                // a code space that overflows is malformed
                int code = 0;
                for (int k = 0; k < 16; ++k) {
                    code += s[q + 1 + k];
                    if (code > (1 << (k + 1))) fail(UF_ERR_INVALID_ARG, "DHT is not a prefix code");
                    code <<= 1;
                }
                (tc ? P.ac[th] : P.dc[th]).set(s + q + 1, s + q + 17, cnt, derive_tables);
                q += 17 + cnt;
            }
        } else if (m == 0xdb) {  // DQT (zigzag order in the file)
            uint32_t q = 0;
            while (q < n) {
                const int pq = s[q] >> 4, tq = s[q] & 15;
                if (tq > 3 || pq > 1) fail(UF_ERR_INVALID_ARG, "bad DQT table id");
                const uint32_t need = 1 + (pq ? 128 : 64);
                if (q + need > n) fail(UF_ERR_INVALID_ARG, "short DQT");
                for (int k = 0; k < 64; ++k) P.qt[tq][kZigzag[k]] = pq ? (uint16_t)be16(s + q + 1 + 2 * k) : s[q + 1 + k];
                P.qt_defined[tq] = true;
                q += need;
            }
        } else if (m == 0xdd) {  // DRI
            if (n < 2) fail(UF_ERR_INVALID_ARG, "short DRI");
            P.restart_interval = be16(s);
        } else if (m == 0xda) {  // SOS
            if (!P.have_sof) fail(UF_ERR_INVALID_ARG, "SOS before SOF");
            if (n < 1) fail(UF_ERR_INVALID_ARG, "short SOS");
            const uint32_t ns = s[0];
            if (ns != P.plan.ncomp) fail(UF_ERR_UNSUPPORTED, "non-interleaved (multi-scan) JPEG");
            if (n < 1 + 2 * ns + 3) fail(UF_ERR_INVALID_ARG, "short SOS");
            for (uint32_t k = 0; k < ns; ++k) {
                if (s[1 + 2 * k] != P.comp_id[k]) fail(UF_ERR_UNSUPPORTED, "scan components out of frame order");
                P.td[k] = s[2 + 2 * k] >> 4;
                P.ta[k] = s[2 + 2 * k] & 15;
                if (P.td[k] > 3 || P.ta[k] > 3) fail(UF_ERR_INVALID_ARG, "bad table selector in SOS");
            }
            if (s[1 + 2 * ns] != 0 || s[2 + 2 * ns] != 63) fail(UF_ERR_UNSUPPORTED, "spectral selection in a sequential scan");
            P.scan_start = i + L;
            break;
        }
        i += L;
        if (!want_scan && P.have_sof) break;
    }
    // geometry
    JpegPlan& p = P.plan;
    if (!P.have_sof) fail(UF_ERR_INVALID_ARG, "no SOF marker");
    p.hmax = p.vmax = 1;
    for (uint32_t c = 0; c < p.ncomp; ++c) {
        if (p.hs[c] > p.hmax) p.hmax = p.hs[c];
        if (p.vs[c] > p.vmax) p.vmax = p.vs[c];
    }
    if (p.ncomp == 1) p.hs[0] = p.vs[0] = p.hmax = p.vmax = 1;  // a single-component scan is never interleaved
    if (p.ncomp == 3 && (p.hs[1] != 1 || p.vs[1] != 1 || p.hs[2] != 1 || p.vs[2] != 1))
        fail(UF_ERR_UNSUPPORTED, "subsampling other than 4:4:4 / 4:2:2 / 4:4:0 / 4:2:0");
    if (p.ncomp == 3 && p.hs[0] == 1 && p.vs[0] == 2) fail(UF_ERR_UNSUPPORTED, "4:4:0 subsampling");
    p.mcus_x = (p.w + 8 * p.hmax - 1) / (8 * p.hmax);
    p.mcus_y = (p.h + 8 * p.vmax - 1) / (8 * p.vmax);
    p.blocks_per_mcu = 0;
    uint32_t off = 0;
    for (uint32_t c = 0; c < p.ncomp; ++c) {
        for (uint32_t v = 0; v < p.vs[c]; ++v)
            for (uint32_t h = 0; h < p.hs[c]; ++h) {
                if (p.blocks_per_mcu >= JPEG_MAX_SLOTS) fail(UF_ERR_INVALID_ARG, "more than 10 blocks per MCU");
                p.slot_comp[p.blocks_per_mcu] = (uint8_t)c;
                p.slot_h[p.blocks_per_mcu] = (uint8_t)h;
                p.slot_v[p.blocks_per_mcu] = (uint8_t)v;
                ++p.blocks_per_mcu;
            }
        p.plane_w[c] = p.mcus_x * p.hs[c] * 8;
        p.plane_h[c] = p.mcus_y * p.vs[c] * 8;
        p.real_w[c] = (p.w * p.hs[c] + p.hmax - 1) / p.hmax;
        p.real_h[c] = (p.h * p.vs[c] + p.vmax - 1) / p.vmax;
        p.plane_off[c] = off;
        off += p.plane_w[c] * p.plane_h[c];
        if (want_scan) {
            if (!P.qt_defined[P.tq[c]]) fail(UF_ERR_INVALID_ARG, "quantisation table not defined");
            memcpy(p.quant[c], P.qt[P.tq[c]], sizeof(p.quant[c]));
        }
    }
    p.plane_bytes = (off + 255) / 256 * 256;
    p.nblocks = p.mcus_x * p.mcus_y * p.blocks_per_mcu;
    if (p.ncomp == 3 && p.hmax == 2 && p.real_w[1] <= 2) fail(UF_ERR_UNSUPPORTED, "image too narrow for libjpeg's fancy upsampling");
}

}  // namespace

bool JpegHuffKey::operator==(const JpegHuffKey& o) const { return memcmp(this, &o, sizeof(*this)) == 0; }

// The device decoder's tables (jpeg_decode.h: JpegHuffTab) from the DHT content. `dc`: the symbol is a size category.
static void build_device_table(const uint8_t bits[16], const uint8_t vals[256], bool dc, JpegHuffTab& T) {
    memset(&T, 0, sizeof(T));
    memcpy(T.vals, vals, 256);
    auto entry = [&](int len, int sym) -> uint16_t {
        const int s = sym & 15, r = sym >> 4;
        const int kadv = dc ? 1 : (s ? r + 1 : (r == 15 ? 16 : 64));
        return (uint16_t)((len + s) | (kadv << 5) | (s << 12));
    };
    int code = 0, k = 0;
    T.lim[0] = 0;
    for (int l = 1; l <= 16; ++l) {
        T.valoff[l] = k - code;
        for (int i = 0; i < bits[l - 1]; ++i, ++k, ++code) {
            if (l <= JH_LOOK && k < 256) {
                const int first = code << (JH_LOOK - l);
                for (int f = 0; f < (1 << (JH_LOOK - l)) && first + f < (1 << JH_LOOK); ++f) T.look[first + f] = entry(l, vals[k]);
            }
        }
        T.lim[l] = bits[l - 1] ? (uint32_t)code << (16 - l) : T.lim[l - 1];
        code <<= 1;
    }
}

void jpeg_build_tabset(const JpegHuffKey& key, JpegHuffTabSet& out) {
    for (int c = 0; c < 3; ++c) {
        build_device_table(key.bits[c], key.vals[c], true, out.dc[c]);
        build_device_table(key.bits[3 + c], key.vals[3 + c], false, out.ac[c]);
    }
}

void jpeg_prepare_bitstream(const uint8_t* data, size_t len, JpegBitstream& out) {
    Parsed P;
    parse(data, len, P, true, false);  // (the decoder tables are not derived here: the device decoder wants the DHT content only)
    if (!P.dc[0].defined) P.dc[0].set(kStdDcLumBits, kStdDcVals, 12, false);
    if (!P.dc[1].defined) P.dc[1].set(kStdDcChrBits, kStdDcVals, 12, false);
    if (!P.ac[0].defined) P.ac[0].set(kStdAcLumBits, kStdAcLumVals, 162, false);
    if (!P.ac[1].defined) P.ac[1].set(kStdAcChrBits, kStdAcChrVals, 162, false);
    const JpegPlan& p = P.plan;
    out.plan = p;
    JpegHuffFrame& h = out.huff;
    memset(&h, 0, sizeof(h));
    memset(&out.key, 0, sizeof(out.key));
    for (uint32_t c = 0; c < p.ncomp; ++c) {
        const HuffTable* src[2] = {&P.dc[P.td[c]], &P.ac[P.ta[c]]};
        if (!src[0]->defined || !src[1]->defined) fail(UF_ERR_INVALID_ARG, "Huffman table not defined");
        for (int k = 0; k < 2; ++k) {
            memcpy(out.key.bits[3 * k + c], src[k]->bits, 16);
            memcpy(out.key.vals[3 * k + c], src[k]->vals, 256);
        }
    }
    h.nblocks = p.nblocks;
    h.blocks_per_mcu = p.blocks_per_mcu;
    for (uint32_t sl = 0; sl < (uint32_t)JPEG_MAX_SLOTS; ++sl) h.slotmap |= (uint32_t)(p.slot_comp[sl] & 3u) << (2 * sl);
    // remove the byte stuffing; the segment ends at the first real marker
    out.data.clear();
    out.data.reserve(len - P.scan_start + 8);
    const uint8_t* q = data + P.scan_start;
    const uint8_t* end = data + len;
    bool marker_inside = false;
    while (q < end) {
        const uint8_t* f = (const uint8_t*)memchr(q, 0xff, (size_t)(end - q));
        if (!f) { out.data.insert(out.data.end(), q, end); break; }
        out.data.insert(out.data.end(), q, f);
        if (f + 1 >= end) break;
        if (f[1] == 0x00) { out.data.push_back(0xff); q = f + 2; continue; }
        if (f[1] == 0xff) { q = f + 1; continue; }  // fill byte
        marker_inside = f[1] != 0xd9;               // EOI ends the segment; anything else (RSTn, DNL ...) is for the host decoder
        break;
    }
    h.data_bits = (uint32_t)out.data.size() * 8;
    h.sub_bits = JH_DEFAULT_SUBSEQ_BITS;
    h.nsub = (h.data_bits + h.sub_bits - 1) / h.sub_bits;
    out.data.resize((out.data.size() + 15) / 16 * 16 + 32, 0);  // zero tail: the reader fetches up to 5 words past the end
    out.gpu_ok = P.restart_interval == 0 && !marker_inside && h.nsub > 0 && h.data_bits <= JH_MAX_DATA_BITS;
}

JpegPlan jpeg_parse_header(const uint8_t* data, size_t len) {
    Parsed P;
    parse(data, len, P, false);
    return P.plan;
}

void jpeg_entropy_decode(const uint8_t* data, size_t len, JpegCoefs& out) {
    Parsed P;
    parse(data, len, P, true);
    // Motion-JPEG frames carry no DHT: the Annex K tables apply (libjpeg-turbo does the same)
    if (!P.dc[0].defined) P.dc[0].set(kStdDcLumBits, kStdDcVals, 12);
    if (!P.dc[1].defined) P.dc[1].set(kStdDcChrBits, kStdDcVals, 12);
    if (!P.ac[0].defined) P.ac[0].set(kStdAcLumBits, kStdAcLumVals, 162);
    if (!P.ac[1].defined) P.ac[1].set(kStdAcChrBits, kStdAcChrVals, 162);
    const JpegPlan& p = P.plan;
    for (uint32_t c = 0; c < p.ncomp; ++c)
        if (!P.dc[P.td[c]].defined || !P.ac[P.ta[c]].defined) fail(UF_ERR_INVALID_ARG, "Huffman table not defined");
    out.plan = p;
    out.block_off.resize((size_t)p.nblocks + 1);
    out.entries.clear();  // (callers reuse JpegCoefs objects: the capacity survives from frame to frame)
    out.entries.reserve(std::min<size_t>((size_t)p.nblocks * 64, (len - P.scan_start) * 2 + 64));
    BitReader br{data + P.scan_start, data + len};
    int pred[3] = {0, 0, 0};
    const uint32_t n_mcu = p.mcus_x * p.mcus_y;
    uint32_t to_restart = P.restart_interval;
    uint32_t blk = 0;
    for (uint32_t mcu = 0; mcu < n_mcu; ++mcu) {
        if (P.restart_interval && to_restart == 0) {
            // byte-align, expect RSTn; whatever is there, predictions restart (libjpeg resynchronises more cleverly on
            // damaged data; on intact data this is identical)
            br.reset_at_restart();
            while (br.p + 1 < br.end && !(br.p[0] == 0xff && br.p[1] >= 0xd0 && br.p[1] <= 0xd7)) {
                if (br.p[0] == 0xff && br.p[1] != 0x00 && br.p[1] != 0xff) break;  // another marker: leave it
                ++br.p;
            }
            if (br.p + 1 < br.end && br.p[0] == 0xff && br.p[1] >= 0xd0 && br.p[1] <= 0xd7) br.p += 2;
            pred[0] = pred[1] = pred[2] = 0;
            to_restart = P.restart_interval;
        }
        if (br.insufficient) {  // data ran out in an earlier MCU: libjpeg leaves the rest of the segment zero
            for (uint32_t sl = 0; sl < p.blocks_per_mcu; ++sl, ++blk) out.block_off[blk] = (uint32_t)out.entries.size();
            if (P.restart_interval) --to_restart;
            continue;
        }
        for (uint32_t sl = 0; sl < p.blocks_per_mcu; ++sl, ++blk) {
            const int c = p.slot_comp[sl];
            const HuffTable& dct = P.dc[P.td[c]];
            const HuffTable& act = P.ac[P.ta[c]];
            out.block_off[blk] = (uint32_t)out.entries.size();
            int s = decode_symbol(br, dct) & 15;
            if (s) {
                if (br.nbits < 32) br.refill();
                pred[c] += extend(br.get(s), s);
            }
            const int16_t dc = (int16_t)pred[c];  // JCOEF is a short
            if (dc) out.entries.push_back((uint32_t)(uint16_t)dc);  // natural index 0
            for (int k = 1; k < 64;) {
                if (br.nbits < 32) br.refill();
                const int16_t fa = act.fast_ac[br.peek(LOOK)];
                if (fa) {  // code + magnitude bits resolved by one lookup
                    k += (fa >> 4) & 15;
                    br.skip(fa & 15);
                    out.entries.push_back(((uint32_t)kZigzag[k] << 16) | (uint16_t)(int16_t)(fa >> 8));
                    ++k;
                    continue;
                }
                const int rs = decode_symbol(br, act);
                const int r = rs >> 4;
                s = rs & 15;
                if (s == 0) {
                    if (r != 15) break;  // EOB
                    k += 16;
                    continue;
                }
                k += r;
                if (br.nbits < 32) br.refill();
                const int16_t v = (int16_t)extend(br.get(s), s);
                // two entries may target index 63 on damaged data; the later one wins, as in libjpeg's block array
                out.entries.push_back(((uint32_t)kZigzag[k < 80 ? k : 79] << 16) | (uint16_t)v);
                ++k;
            }
        }
        if (P.restart_interval) --to_restart;
    }
    out.block_off[p.nblocks] = (uint32_t)out.entries.size();
}

}  // namespace uf
