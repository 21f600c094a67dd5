// kernels_jpeg_henc.cu — N3: the entropy-coding half of the JPEG encoder on the GPU, for batches. Replaces the Huffman coder
// of `turbojpeg::compress_image(&frame, 95, Subsamp::Sub2x2)` (/root/reference/infer_server/src/inferer.rs:39) after the forward
// DCT and quantisation (kernels_jpeg_enc.cu): baseline sequential Huffman coding (T.81 F.1.2) of a YCbCr 4:2:0 frame with
// the Annex K tables, byte for byte what the host writer (jpeg_encode.cc) — hence libjpeg-turbo — produces.
//
// Encoding, unlike decoding, is parallel by construction: the bits of a block depend only on its coefficients and on the DC
// of the block coded before it in the same component.
//   jenc_block_kernel<false>  warp per block (MCU order): DC difference, run lengths from a ballot of the nonzero
//                             coefficients in zigzag order, code lengths from the tables -> bits of the block
//   jenc_scan_kernel          CTA per frame: exclusive prefix sum -> bit offset of every block, bits of the frame
//   jenc_block_kernel<true>   the same walk, now assembling every coefficient's bits ([ZRL ...] code magnitude, <= 59 bits)
//                             and OR-ing them into the zeroed bit buffer at (block offset + lane offset)
//   jenc_stuff_kernel         8 CTAs per frame: pad the last byte with 1-bits, insert 0x00 after every 0xFF byte (count per
//                             thread, prefix sum, copy) -> the entropy-coded segment as it stands in the file
// The host prepends the headers and appends EOI (jpeg_write_headers). libjpeg's dummy blocks (luma blocks of an edge MCU that
// lie wholly outside the image) are coded as it codes them: AC zero, DC = the DC of the block before in the MCU.
#include "jpeg_decode.h"
#include "jpeg_henc_core.h"
#include "kernels.h"

namespace uf {

using namespace he;

namespace {

__constant__ uint8_t c_enc_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                         41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                         30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total) {
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - v;
}

}  // namespace

template <bool WRITE>
__global__ void __launch_bounds__(256)
jenc_block_kernel(JpegEncBatch B) {
    __shared__ EncTabs T;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(B.tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&T);
        for (int i = threadIdx.x; i < (int)(sizeof(JpegEncTables) / 4); i += 256) dst[i] = src[i];
        if (threadIdx.x < 64) T.zz[threadIdx.x] = c_enc_zigzag[threadIdx.x];
    }
    __syncthreads();
    const JpegEncFrame& F = B.frames[blockIdx.y];
    const uint32_t b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= F.nblocks) return;
    const int16_t* coefs = B.coefs + (size_t)F.coef_base * 64;
    const BlockSrc S = block_source(F, coefs, b);  // (all lanes compute the same few values)
    const int chroma = S.chroma;
    const int16_t* blk = coefs + (size_t)S.src * 64;
    const int v1 = lane == 0 ? S.dc_diff : (S.dummy ? 0 : (int)blk[T.zz[lane]]);
    const int v2 = S.dummy ? 0 : (int)blk[T.zz[lane + 32]];
    const uint32_t m1 = __ballot_sync(0xffffffffu, v1 != 0) & ~1u, m2 = __ballot_sync(0xffffffffu, v2 != 0);
    const unsigned long long M = (unsigned long long)m1 | ((unsigned long long)m2 << 32);
    unsigned long long bits1, bits2;
    uint32_t len1, len2;
    coef_bits(T, chroma, lane, v1, M, bits1, len1);
    coef_bits(T, chroma, lane + 32, v2, M, bits2, len2);
    int tot1, tot2;
    const int off1 = warp_excl_scan((int)len1, (int)lane, tot1);
    const int off2 = warp_excl_scan((int)len2, (int)lane, tot2);
    const uint32_t last = M ? top_bit64(M) : 0u;
    const uint32_t eob = last < 63 ? T.ac[chroma][0] : 0u;  // trailing zeros: EOB
    const uint32_t total = (uint32_t)(tot1 + tot2) + (eob >> 16);
    if (!WRITE) {
        if (lane == 0) B.bitlen[F.len_base + b] = total;
        return;
    }
    uint32_t* P = B.packed + F.pack_off;
    const uint32_t at = B.bitoff[F.len_base + b];
    put64(P, bits1, len1, at + (uint32_t)off1, F.pack_cap_bits);
    put64(P, bits2, len2, at + (uint32_t)(tot1 + off2), F.pack_cap_bits);
    if (lane == 0 && eob) put64(P, eob & 0xffffu, eob >> 16, at + (uint32_t)(tot1 + tot2), F.pack_cap_bits);
}

// exclusive prefix sum of the blocks' bit counts, one CTA per frame
__global__ void __launch_bounds__(1024)
jenc_scan_kernel(JpegEncBatch B) {
    __shared__ uint32_t s_warp[32];
    const JpegEncFrame& F = B.frames[blockIdx.x];
    const uint32_t n = F.nblocks, per = (n + 1023) / 1024, i0 = min(threadIdx.x * per, n), i1 = min(i0 + per, n);
    const uint32_t* len = B.bitlen + F.len_base;
    uint32_t* off = B.bitoff + F.len_base;
    uint32_t sum = 0;
    for (uint32_t i = i0; i < i1; ++i) sum += len[i];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t run = incl - sum, all = 0;
    for (uint32_t w = 0; w < 32; ++w) {
        if (w < warp) run += s_warp[w];
        all += s_warp[w];
    }
    for (uint32_t i = i0; i < i1; ++i) {
        off[i] = run;
        run += len[i];
    }
    if (threadIdx.x == 0) B.frame_bits[blockIdx.x] = all;
}

// bit buffer -> the entropy-coded segment: last byte padded with 1-bits, 0x00 after every 0xFF. A frame is cut into
// JENC_STUFF_SPLIT pieces, one CTA each: a CTA counts the 0xFF bytes of everything before its piece (whole words, straight
// from L2), then those of its own piece per thread, scans, and copies its bytes to where they land.
constexpr int JENC_STUFF_SPLIT = 8;

__global__ void __launch_bounds__(1024)
jenc_stuff_kernel(JpegEncBatch B) {
    __shared__ uint32_t s_warp[32];
    const JpegEncFrame& F = B.frames[blockIdx.y];
    const uint32_t bits = B.frame_bits[blockIdx.y];
    if (bits > F.pack_cap_bits) {  // does not fit the bit buffer: the host encoder takes the frame
        if (blockIdx.x == 0 && threadIdx.x == 0) B.out_len[blockIdx.y] = 0xffffffffu;
        return;
    }
    const uint32_t nbytes = (bits + 7) / 8;
    const uint32_t seg = ((nbytes + JENC_STUFF_SPLIT - 1) / JENC_STUFF_SPLIT + 3u) & ~3u;  // whole words
    const uint32_t s0 = min(blockIdx.x * seg, nbytes), s1 = min(s0 + seg, nbytes);
    if (s0 >= s1) {
        if (nbytes == 0 && blockIdx.x == 0 && threadIdx.x == 0) B.out_len[blockIdx.y] = 0;
        return;
    }
    const uint32_t* P = B.packed + F.pack_off;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto cta_scan = [&](uint32_t v, uint32_t& all) -> uint32_t {  // exclusive prefix over the CTA's threads, and the total
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += y;
        }
        __syncthreads();  // (s_warp may still be read from the previous use)
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = incl - v;
        all = 0;
        for (uint32_t w = 0; w < 32; ++w) {
            if (w < warp) before += s_warp[w];
            all += s_warp[w];
        }
        return before;
    };
    // 0xFF bytes in front of the piece (s0 is a multiple of 4 and lies before the padded last byte)
    uint32_t cnt = 0;
    for (uint32_t w = threadIdx.x; w < s0 / 4; w += 1024) {
        const uint32_t x = P[w];
        cnt += ((x >> 24) == 0xffu) + (((x >> 16) & 0xffu) == 0xffu) + (((x >> 8) & 0xffu) == 0xffu) + ((x & 0xffu) == 0xffu);
    }
    uint32_t front = 0;
    cta_scan(cnt, front);
    // the piece itself
    const uint32_t per = (s1 - s0 + 1023) / 1024, i0 = min(s0 + threadIdx.x * per, s1), i1 = min(i0 + per, s1);
    uint32_t ff = 0;
    for (uint32_t i = i0; i < i1; ++i) ff += packed_byte(P, i, bits) == 0xffu;
    uint32_t inside = 0;
    const uint32_t before = cta_scan(ff, inside);
    uint8_t* out = B.out + F.out_off;
    uint32_t pos = i0 + front + before;
    for (uint32_t i = i0; i < i1; ++i) {
        const uint32_t v = packed_byte(P, i, bits);
        if (pos < F.out_cap) out[pos] = (uint8_t)v;
        ++pos;
        if (v == 0xffu) {
            if (pos < F.out_cap) out[pos] = 0;
            ++pos;
        }
    }
    if (s1 == nbytes && threadIdx.x == 0) {  // the piece that holds the last byte knows the length of the whole segment
        const uint32_t out_len = nbytes + front + inside;
        B.out_len[blockIdx.y] = out_len > F.out_cap ? 0xffffffffu : out_len;
    }
}

void launch_jpeg_huffman_encode(const JpegEncBatch& B, int frames, uint32_t max_nblocks, cudaStream_t s) {
    if (frames <= 0) return;
    const dim3 grid((max_nblocks + 7) / 8, frames);
    jenc_block_kernel<false><<<grid, 256, 0, s>>>(B);
    jenc_scan_kernel<<<frames, 1024, 0, s>>>(B);
    jenc_block_kernel<true><<<grid, 256, 0, s>>>(B);
    jenc_stuff_kernel<<<dim3(JENC_STUFF_SPLIT, frames), 1024, 0, s>>>(B);
}

}  // namespace uf
