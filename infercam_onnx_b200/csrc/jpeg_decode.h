// jpeg_decode.h — N2: JPEG in front of the hot path.
//
// The reference decodes every frame on the CPU with libjpeg-turbo before the path starts
// (`turbojpeg::decompress_image`, /root/reference/infer_server/src/inferer.rs:35; README.md:62-64: ~15 ms per frame for
// decode + encode). Here the host parses the headers and removes the 0xFF00 byte stuffing (jpeg_entropy.cc:
// jpeg_prepare_bitstream); the entropy-coded bytes themselves cross PCIe and are Huffman-decoded on the GPU
// (kernels_jpeg_huff.cu) into per-block lists of nonzero coefficients. Dequantisation, the inverse DCT, chroma upsampling and
// YCbCr -> RGB (kernels_jpeg.cu) restate libjpeg-turbo's DEFAULT decoder bit for bit: `jpeg_idct_islow` (jidctint.c),
// `h2v1_fancy_upsample` / `h2v2_fancy_upsample` (jdsample.c) and `ycc_rgb_convert` (jdcolor.c), the algorithms tjDecompress2
// runs with flags = 0. A sequential host Huffman decoder (jpeg_entropy.cc: jpeg_entropy_decode, one frame per worker thread,
// same list format) takes the frames the device decoder declines — restart intervals, markers inside the scan, truncated or
// damaged data — and everything under UF_FLAG_JPEG_HOST_HUFFMAN. Scope: baseline sequential JPEG (SOF0 / SOF1 Huffman,
// 8 bit), one interleaved scan, 1 or 3 components, 4:4:4 / 4:2:2 / 4:2:0 — what a V4L2 MJPG webcam sends
// (cam_sender/src/sensors.rs). Progressive files are refused with UF_ERR_UNSUPPORTED.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace uf {

constexpr int JPEG_MAX_SLOTS = 10;  // blocks per MCU (JPEG limit)

// Geometry + tables of one frame: plain data, copied to the device as is.
struct JpegPlan {
    uint32_t w, h;                 // image size
    uint32_t ncomp;                // 1 or 3
    uint32_t hmax, vmax;           // largest sampling factors
    uint32_t mcus_x, mcus_y;       // MCUs per row / column
    uint32_t blocks_per_mcu;
    uint32_t nblocks;              // mcus_x * mcus_y * blocks_per_mcu (decode order)
    uint32_t hs[3], vs[3];         // sampling factors
    uint32_t plane_w[3], plane_h[3];   // component planes padded to whole MCUs (samples)
    uint32_t real_w[3], real_h[3];     // downsampled_width / height: ceil(w * hs / hmax) ...
    uint32_t plane_off[3];         // byte offset of the plane inside the frame's plane buffer
    uint32_t plane_bytes;          // all planes
    uint32_t offs_base;            // index of this frame's first block offset in the batch's offset array
    uint32_t entries_base;         // index of this frame's first entry in the batch's entry array
    uint32_t rgb_off_lo, rgb_off_hi;   // byte offset of the decoded RGB frame in the destination buffer (64 bit)
    uint32_t planes_off_lo, planes_off_hi;  // byte offset of the frame's planes in the batch's plane buffer
    uint8_t slot_comp[JPEG_MAX_SLOTS], slot_h[JPEG_MAX_SLOTS], slot_v[JPEG_MAX_SLOTS];  // block slot in the MCU -> component, offsets
    uint8_t pad_[2];
    uint16_t quant[3][64];         // per component, NATURAL (row-major) order
};

// One frame after Huffman decoding. entry = (natural index << 16) | (uint16) quantised value.
struct JpegCoefs {
    JpegPlan plan{};
    std::vector<uint32_t> block_off;  // nblocks + 1, offsets into `entries` (decode order)
    std::vector<uint32_t> entries;
};

struct JpegError {
    int code;  // uf_status
    std::string msg;
};

// ---- Huffman decoding on the GPU (kernels_jpeg_huff.cu): what the host prepares
constexpr int JH_LOOK = 10;         // bits resolved by one table lookup on the device
// One Huffman table in the device decoder's form. A look entry says everything the decoder's loop needs about a symbol:
//   bits 0-4   bits to consume: code length + magnitude bits
//   bits 5-11  how far the coefficient index moves: run + 1 for a coefficient, 16 for ZRL, 64 for EOB (= block done), 1 for DC
//   bits 12-15 magnitude bits (0: no value follows — EOB, ZRL, or a zero DC difference)
// 0 = the code is longer than JH_LOOK bits (lim / valoff / vals resolve it).
struct JpegHuffTab {
    uint16_t look[1 << JH_LOOK];
    uint32_t lim[17];               // [l]: 16-bit left-aligned windows below it hold a code of <= l bits (lengths without codes repeat)
    int32_t valoff[17];             // vals index = code + valoff[length]
    uint8_t vals[256];
};
struct JpegHuffTabSet {             // the tables of a frame, per component; frames with the same DHT content share one
    JpegHuffTab dc[3], ac[3];
};
struct JpegHuffKey {                // what a table set is built from (and compared by): the DHT content as the frame uses it
    uint8_t bits[6][16];            // dc of component 0..2, ac of component 0..2
    uint8_t vals[6][256];
    bool operator==(const JpegHuffKey& o) const;
};
struct JpegHuffFrame {              // plain data, copied to the device as is
    uint32_t tabset;                // index of the frame's table set in the batch
    uint32_t data_off;              // byte offset of the frame's UNSTUFFED entropy-coded segment in the batch's byte buffer
    uint32_t data_bits;             // its length in bits
    uint32_t sub_bits;              // bits per subsequence (= per GPU thread), chosen when the batch is laid out
    uint32_t nsub;                  // subsequences: ceil(data_bits / sub_bits)
    uint32_t sub_base;              // index of the frame's first subsequence in the batch's state arrays
    uint32_t offs_base;             // index of the frame's first block in the batch's block-offset and DC arrays (nblocks + 1 each)
    uint32_t ent_base;              // index of the frame's first entry in the batch's entry array
    uint32_t ent_cap;               // entries reserved for the frame: data_bits / 2 + 16 (an AC entry takes at least 2 bits)
    uint32_t nblocks, blocks_per_mcu;
    uint32_t slotmap;               // 2 bits per block slot of the MCU: its component
};
// Bits per GPU thread. Short subsequences give more threads and shorter passes, long ones fewer passes over the data (a
// decoder started in a guessed state needs a few thousand bits to fall into step, whatever the subsequence length: every
// thread decodes about 1 + that distance / sub_bits subsequences' worth of bits). The engine picks per run from how many
// bits the run holds (engine.cu: huffman_run_gpu).
constexpr uint32_t JH_MIN_SUBSEQ_BITS = 128, JH_MAX_SUBSEQ_BITS = 2048, JH_DEFAULT_SUBSEQ_BITS = 512;
constexpr uint32_t JH_MAX_DATA_BITS = 1u << 25;  // beyond (4 MB of entropy-coded data): host decoder
struct JpegBitstream {              // one frame, host side
    JpegPlan plan{};
    JpegHuffFrame huff{};           // geometry (offsets and the table set index are filled in when the batch is laid out)
    JpegHuffKey key{};
    std::vector<uint8_t> data;      // entropy-coded segment with the 0xFF00 stuffing removed, up to the first marker
    bool gpu_ok = false;            // false: restart intervals / markers inside the segment / oversized -> host entropy decoder
};
void jpeg_build_tabset(const JpegHuffKey& key, JpegHuffTabSet& out);
// Parses the headers, derives the decoder tables, removes the byte stuffing. Throws JpegError like jpeg_entropy_decode.
void jpeg_prepare_bitstream(const uint8_t* data, size_t len, JpegBitstream& out);

// ---- encoder (N3), jpeg_encode.cc
void jpeg_quality_tables(int quality, uint16_t lum[64], uint16_t chr[64]);   // jpeg_set_quality(q, force_baseline)
JpegPlan jpeg_encode_plan(uint32_t w, uint32_t h, int quality);              // YCbCr 4:2:0 geometry + tables
// coefs: quantised blocks per component plane in raster order (see jpeg_encode.cc); writes a complete JFIF file
void jpeg_write_file(const JpegPlan& p, const int16_t* coefs, std::vector<uint8_t>& out);
void jpeg_write_headers(const JpegPlan& p, std::vector<uint8_t>& out);  // SOI .. SOS (the entropy-coded segment and EOI follow)
struct JpegEncJob {        // one frame of a batch for the encoder's sample-domain kernels (kernels_jpeg_enc.cu)
    JpegPlan plan;
    unsigned long long rgb_off, planes_off;  // byte offsets of the frame's pixels / component planes
    unsigned long long coef_off;             // offset of its coefficients, in int16 elements
};
struct JpegEncFrame {      // plain data, copied to the device as is
    uint32_t coef_base;    // index of the frame's first block in the coefficient buffer (blocks per plane in raster order)
    uint32_t mcus_x, mcus_y;
    uint32_t y_bw, c_bw;   // blocks per row of the padded luma / chroma planes
    uint32_t cb_off, cr_off;  // first block of the Cb / Cr plane, relative to coef_base
    uint32_t wib0, hib0;   // luma blocks per row / column that hold image samples (beyond: libjpeg's dummy blocks)
    uint32_t nblocks;      // mcus * 6, in MCU order
    uint32_t len_base;     // index of the frame's first block in the bit-count / bit-offset arrays
    uint32_t pack_off;     // word offset of the frame's bit buffer
    uint32_t pack_cap_bits;
    uint32_t out_off;      // byte offset of the frame's entropy-coded segment in the output buffer
    uint32_t out_cap;
    uint32_t pad_;
};
struct JpegEncTables {     // Annex K tables in encoder form: (length << 16) | code; [0] luma, [1] chroma
    uint32_t dc[2][16];
    uint32_t ac[2][256];
};
void jpeg_std_enc_tables(JpegEncTables& t);                             // the Annex K tables as the device encoder reads them

// Parses the headers only. Throws JpegError.
JpegPlan jpeg_parse_header(const uint8_t* data, size_t len);
// Parses and Huffman-decodes. Throws JpegError (UF_ERR_UNSUPPORTED: progressive / arithmetic / 12 bit / multi-scan ...;
// UF_ERR_INVALID_ARG: not a JPEG / truncated headers). A truncated or corrupt entropy segment decodes like libjpeg does:
// missing data reads as zero bits.
void jpeg_entropy_decode(const uint8_t* data, size_t len, JpegCoefs& out);

}  // namespace uf
