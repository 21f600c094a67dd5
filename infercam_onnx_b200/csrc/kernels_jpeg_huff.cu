// kernels_jpeg_huff.cu — N2: Huffman decoding of baseline JPEG on the GPU. Replaces the last host-side piece of
// `turbojpeg::decompress_image` (/root/reference/infer_server/src/inferer.rs:35), the bit-serial entropy decoder, for frames
// without restart intervals (what MJPG webcams send); frames with restart markers, and frames that fail the checks below,
// take the host decoder (jpeg_entropy.cc). What crosses PCIe is then the JPEG's own entropy-coded bytes.
//
// A Huffman stream can only be decoded from a known state (bit position, block slot in the MCU, coefficient index), and the
// state at any point depends on everything before it. The way round it (self-synchronising parallel decoding, Klein & Wiseman;
// Weissenberger & Schmidt's GPU formulation) is to cut the stream into subsequences of JH_SUBSEQ_BITS, let one thread decode
// each from a GUESSED state, and iterate: in every round thread t restarts from the END state thread t-1 reached in the
// round before. Thread 0 starts from the true state, so the correct prefix grows by at least one subsequence per round —
// and usually by many, because a decoder started in a wrong state falls into step with the true one after a few dozen
// symbols. The fixed point (no end state changes) IS the sequential decode: end[t] = decode(end[t-1]) for all t, end[-1] true.
// The rounds run inside a CTA over its 128 subsequences, and once per launch across CTAs (jhuff_sync_kernel).
// Then a prefix sum of the blocks started per subsequence gives every thread its output position, a second pass writes the
// coefficients (DC as differences), a per-component scan turns DC differences into values, and the dense blocks go to the
// IDCT kernel. The coefficients are the host decoder's, hence libjpeg-turbo's, bit for bit (tests/test_jpeg.py).
#include "jpeg_decode.h"
#include "jpeg_huff_core.h"
#include "kernels.h"

namespace uf {

using namespace jh;

namespace {

constexpr int JHT = 128;  // threads per CTA (one subsequence each)

__device__ __forceinline__ void load_tabs(Tabs& tabs, const JpegHuffFrame& fr) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&fr.dc[0]);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&tabs);
    for (int i = threadIdx.x; i < (int)(sizeof(Tabs) / 4); i += JHT) dst[i] = src[i];
}

}  // namespace

// One synchronisation round. A CTA owns JHT consecutive subsequences and iterates INSIDE the kernel until they agree with
// one another: thread i restarts from the end state thread i-1 reached, until no start state changes (<= JHT iterations;
// in practice a handful, a decoder started in a wrong state falls into step with the true one within a few dozen symbols).
// What a CTA cannot know is the end state of the CTA before it: that comes from the previous launch (`in`), so after launch
// r the first r+1 CTAs of every frame are exact, and ceil(nsub / JHT) launches make the whole frame the sequential decode —
// a fixed number, no flag to read back, nothing for the host to wait on. A CTA whose incoming state did not change since
// its last run copies its end states and leaves.
__global__ void __launch_bounds__(JHT)
jhuff_sync_kernel(JpegHuffBatch b, int first, const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out) {
    __shared__ Tabs tabs;
    __shared__ unsigned long long s_end[JHT];
    const JpegHuffFrame& fr = b.frames[blockIdx.y];
    const uint32_t t0 = blockIdx.x * JHT;
    if (t0 >= fr.nsub) return;
    const uint32_t t = t0 + threadIdx.x;
    const bool active = t < fr.nsub;
    const size_t gi = (size_t)fr.sub_base + t;
    const unsigned long long cta_start =
        blockIdx.x == 0 ? 0ull : (first ? pack_state(t0 * JH_SUBSEQ_BITS, 0, 0) : in[(size_t)fr.sub_base + t0 - 1]);
    if (!first && cta_start == b.start_used[(size_t)fr.sub_base + t0]) {  // (uniform over the CTA)
        if (active) out[gi] = in[gi];
        return;
    }
    load_tabs(tabs, fr);
    unsigned long long my_start = ~0ull, my_end = 0;
    uint32_t my_n = 0;
    bool dirty = false;
    if (first) {
        my_end = pack_state((t + 1) * JH_SUBSEQ_BITS, 0, 0);  // the next thread's first guess: its own beginning, slot 0, DC
    } else if (active) {
        my_start = b.start_used[gi];
        my_end = in[gi];
    }
    s_end[threadIdx.x] = my_end;
    __syncthreads();
    const uint32_t* data = reinterpret_cast<const uint32_t*>(b.bytes + fr.data_off);
    const uint32_t p_end = min((t + 1) * JH_SUBSEQ_BITS, fr.data_bits);
    for (int iter = 0; iter <= JHT; ++iter) {
        const unsigned long long ns = threadIdx.x == 0 ? cta_start : s_end[threadIdx.x - 1];
        bool ch = false;
        if (active && ns != my_start) {
            uint32_t p = (uint32_t)(ns >> 32), slot = (uint32_t)(ns >> 8) & 0xff, k = (uint32_t)ns & 0xff;
            if (slot >= fr.blocks_per_mcu) slot = 0;
            my_n = huff_run<false>(tabs, fr, data, p, slot, k, p_end, nullptr, 0);
            my_end = pack_state(p, slot, k);
            my_start = ns;
            ch = dirty = true;
        }
        __syncthreads();  // every thread has read its neighbour's state
        if (ch) s_end[threadIdx.x] = my_end;
        if (!__syncthreads_or(ch)) break;
    }
    if (active) {
        out[gi] = my_end;
        if (dirty) {
            b.start_used[gi] = my_start;
            b.nblk[gi] = my_n;
        }
    }
}

// exclusive prefix sum of the blocks started per subsequence, one CTA per frame
__global__ void __launch_bounds__(256)
jhuff_scan_kernel(JpegHuffBatch b) {
    __shared__ uint32_t part[256];
    const JpegHuffFrame& fr = b.frames[blockIdx.x];
    const uint32_t n = fr.nsub, per = (n + 255) / 256, t0 = threadIdx.x * per, t1 = min(t0 + per, n);
    uint32_t s = 0;
    for (uint32_t t = t0; t < t1; ++t) s += b.nblk[fr.sub_base + t];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int i = 0; i < 256; ++i) { const uint32_t v = part[i]; part[i] = acc; acc += v; }
    }
    __syncthreads();
    uint32_t acc = part[threadIdx.x];
    for (uint32_t t = t0; t < t1; ++t) {
        b.blk_base[fr.sub_base + t] = acc;
        acc += b.nblk[fr.sub_base + t];
    }
}

// second pass: every subsequence again, from its settled start state, writing the coefficients
__global__ void __launch_bounds__(JHT)
jhuff_write_kernel(JpegHuffBatch b, const unsigned long long* __restrict__ fin) {
    __shared__ Tabs tabs;
    const JpegHuffFrame& fr = b.frames[blockIdx.y];
    if (blockIdx.x * JHT >= fr.nsub) return;
    load_tabs(tabs, fr);
    __syncthreads();
    const uint32_t t = blockIdx.x * JHT + threadIdx.x;
    if (t >= fr.nsub) return;
    const size_t gi = (size_t)fr.sub_base + t;
    const unsigned long long start = t == 0 ? 0ull : fin[gi - 1];
    uint32_t p = (uint32_t)(start >> 32), slot = (uint32_t)(start >> 8) & 0xff, k = (uint32_t)start & 0xff;
    const uint32_t p_end = min((t + 1) * JH_SUBSEQ_BITS, fr.data_bits);
    const uint32_t base = b.blk_base[gi];
    const uint32_t n = huff_run<true>(tabs, fr, reinterpret_cast<const uint32_t*>(b.bytes + fr.data_off), p, slot, k, p_end,
                                      b.coefs + (size_t)fr.coef_base * 64, base);
    if (t == fr.nsub - 1)  // the frame decoded to exactly its blocks, ending on a block boundary: else the host decoder takes it
        b.status[blockIdx.y] = (base + n == fr.nblocks && k == 0) ? 0 : 1;
}

// DC differences -> DC values: per component a running sum over its blocks in decode order (JCOEF wraps at 16 bits)
__global__ void __launch_bounds__(96)
jhuff_dc_kernel(JpegHuffBatch b) {
    const JpegHuffFrame& fr = b.frames[blockIdx.x];
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t slots[JPEG_MAX_SLOTS], cnt = 0;
    for (uint32_t s = 0; s < fr.blocks_per_mcu; ++s)
        if (fr.slot_comp[s] == c) slots[cnt++] = s;
    if (cnt == 0) return;
    const uint32_t total = fr.nblocks / fr.blocks_per_mcu * cnt;
    int16_t* coefs = b.coefs + (size_t)fr.coef_base * 64;
    int carry = 0;
    constexpr int U = 4;  // chunks of 32 blocks whose loads are issued together (only the carry is serial)
    for (uint32_t i0 = 0; i0 < total; i0 += 32 * U) {
        size_t idx[U];
        int v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = i0 + u * 32 + lane;
            idx[u] = 0;
            v[u] = 0;
            if (i < total) {
                const uint32_t mcu = i / cnt, j = i - mcu * cnt;
                uint32_t sl = slots[0];
#pragma unroll
                for (int q = 1; q < JPEG_MAX_SLOTS; ++q)
                    if ((uint32_t)q == j) sl = slots[q];
                idx[u] = ((size_t)mcu * fr.blocks_per_mcu + sl) * 64;
                v[u] = coefs[idx[u]];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int x = v[u];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            x += carry;
            if (i0 + u * 32 + lane < total) coefs[idx[u]] = (int16_t)x;
            carry = __shfl_sync(0xffffffffu, x, 31);
        }
    }
}

void launch_jhuff_sync(const JpegHuffBatch& b, int frames, uint32_t max_nsub, int first, const unsigned long long* in,
                       unsigned long long* out, cudaStream_t s) {
    jhuff_sync_kernel<<<dim3((max_nsub + JHT - 1) / JHT, frames), JHT, 0, s>>>(b, first, in, out);
}

int jhuff_rounds(uint32_t max_nsub) { return (int)((max_nsub + JHT - 1) / JHT); }

void launch_jhuff_finish(const JpegHuffBatch& b, int frames, uint32_t max_nsub, const unsigned long long* fin, cudaStream_t s) {
    jhuff_scan_kernel<<<frames, 256, 0, s>>>(b);
    jhuff_write_kernel<<<dim3((max_nsub + JHT - 1) / JHT, frames), JHT, 0, s>>>(b, fin);
    jhuff_dc_kernel<<<frames, 96, 0, s>>>(b);
}

}  // namespace uf
