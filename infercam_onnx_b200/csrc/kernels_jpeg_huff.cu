// kernels_jpeg_huff.cu — N2: Huffman decoding of baseline JPEG on the GPU. Replaces the last host-side piece of
// `turbojpeg::decompress_image` (/root/reference/infer_server/src/inferer.rs:35), the bit-serial entropy decoder, for frames
// without restart intervals (what MJPG webcams send); frames with restart markers, and frames that fail the checks below,
// take the host decoder (jpeg_entropy.cc). What crosses PCIe is then the JPEG's own entropy-coded bytes.
//
// A Huffman stream can only be decoded from a known state (bit position, block slot in the MCU, coefficient index), and the
// state at any point depends on everything before it. The way round it (self-synchronising parallel decoding, Klein & Wiseman;
// Weissenberger & Schmidt's GPU formulation) is to cut the stream into subsequences of 256-2048 bits (picked per run, engine.cu), let one thread decode
// each from a GUESSED state, and iterate: in every round thread t restarts from the END state thread t-1 reached in the
// round before. Thread 0 starts from the true state, so the correct prefix grows by at least one subsequence per round —
// and usually by many, because a decoder started in a wrong state falls into step with the true one after a few dozen
// symbols. The fixed point (no end state changes) IS the sequential decode: end[t] = decode(end[t-1]) for all t, end[-1] true.
// The rounds run inside a CTA over its 256 subsequences, and once per launch across CTAs (jhuff_sync_kernel).
// Then a prefix sum of the blocks started and nonzero coefficients met per subsequence gives every thread its output
// position, a second pass writes the coefficients as the compact per-block lists the IDCT kernel reads (DC differences
// apart), and a per-component scan turns DC differences into values. The coefficients are the host decoder's, hence
// libjpeg-turbo's, bit for bit (tests/test_jpeg.py).
#include "jpeg_decode.h"
#include "jpeg_huff_core.h"
#include "kernels.h"

namespace uf {

using namespace jh;

namespace {

constexpr int JHT = 256;         // threads per CTA (one subsequence each): a CTA spans 8-64 KB of the stream
constexpr int JH_MAX_ROUNDS = 4;  // launches of the synchronisation kernel (see jhuff_sync_kernel)

__constant__ uint8_t c_zigzag[80] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                                     6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                                     39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

__device__ __forceinline__ void load_tabs(Tabs& tabs, const JpegHuffTabSet& set) {
    static_assert(sizeof(JpegHuffTabSet) % 16 == 0, "copied as uint4");
    const uint4* src = reinterpret_cast<const uint4*>(&set);
    uint4* dst = reinterpret_cast<uint4*>(&tabs.set);
    for (int i = threadIdx.x; i < (int)(sizeof(JpegHuffTabSet) / 16); i += JHT) dst[i] = src[i];
    if (threadIdx.x < 80) tabs.zz[threadIdx.x] = c_zigzag[threadIdx.x];
}

// blocks | entries << 32 of a packed per-subsequence count
__device__ __forceinline__ unsigned long long widen(uint32_t c) { return (unsigned long long)(c & 0xffffu) | ((unsigned long long)(c >> 16) << 32); }

// sum over the CTA, the same value in every thread (blockDim.x = JHT)
__device__ __forceinline__ unsigned long long cta_sum(unsigned long long v, unsigned long long* s_part) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < JHT / 32; ++w) t += s_part[w];
    __syncthreads();
    return t;
}

}  // namespace

// One synchronisation round. A CTA owns JHT consecutive subsequences and iterates INSIDE the kernel until they agree with
// one another: thread i restarts from the end state thread i-1 reached, until no start state changes (<= JHT iterations;
// in practice a handful, a decoder started in a wrong state falls into step with the true one within a few dozen symbols).
// What a CTA cannot know is the end state of the CTA before it: that comes from the previous launch (`in`), so after launch
// r the first r+1 CTAs of every frame are exact, and ceil(nsub / JHT) launches would make any frame the sequential decode.
// In practice launch 1 already hands every CTA the true state (the CTA before it fell into step long before its end), so
// min(ceil(nsub / JHT), JH_MAX_ROUNDS) launches are made — a fixed number, no flag to read back, nothing for the host to
// wait on — and the write pass CHECKS the fixed point (every thread's start state is its predecessor's end state); a frame
// that fails the check is handed to the host decoder like any other the device declines. A CTA whose incoming state did
// not change since its last run copies its end states and leaves. In every iteration the subsequences that have to be
// decoded again are packed to the front of the CTA (ballot + prefix), so that however few and scattered they are, the
// decode runs in full warps.
__global__ void __launch_bounds__(JHT)
jhuff_sync_kernel(JpegHuffBatch b, int first, const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out) {
    __shared__ __align__(16) Tabs tabs;
    // per subsequence of the CTA: the start state of its last decode, the end state that gave, what it counted
    __shared__ unsigned long long s_start[JHT], s_end[JHT], s_ns[JHT];
    __shared__ uint32_t s_cnt[JHT], s_woff[JHT / 32];
    __shared__ uint16_t s_list[JHT];
    __shared__ uint8_t s_dirty[JHT];
    const JpegHuffFrame& fr = b.frames[blockIdx.y];
    const uint32_t nsub = fr.nsub, t0 = blockIdx.x * JHT, sub_bits = fr.sub_bits;
    if (t0 >= nsub) return;
    const uint32_t tid = threadIdx.x, t = t0 + tid, lane = tid & 31, warp = tid >> 5;
    const bool active = t < nsub;
    const size_t gi = (size_t)fr.sub_base + t;
    const unsigned long long cta_start =
        blockIdx.x == 0 ? 0ull : (first ? pack_state(t0 * sub_bits, 0, 0) : in[(size_t)fr.sub_base + t0 - 1]);
    if (!first && cta_start == b.start_used[(size_t)fr.sub_base + t0]) {  // (uniform over the CTA)
        if (active) out[gi] = in[gi];
        return;
    }
    load_tabs(tabs, b.tabsets[fr.tabset]);
    s_start[tid] = (!first && active) ? b.start_used[gi] : ~0ull;
    // first launch: a thread's "end state" is the next thread's first guess — its own beginning, slot 0, DC
    s_end[tid] = first ? pack_state((t + 1) * sub_bits, 0, 0) : (active ? in[gi] : 0ull);
    s_cnt[tid] = 0;
    s_dirty[tid] = 0;
    __syncthreads();
    const uint32_t* data = reinterpret_cast<const uint32_t*>(b.bytes + fr.data_off);
    const uint32_t data_bits = fr.data_bits, bpm = fr.blocks_per_mcu, slotmap = fr.slotmap;
    for (int iter = 0; iter <= JHT; ++iter) {
        // who has to decode again: every subsequence whose predecessor's end state is not the state it started from last time
        const unsigned long long ns = tid == 0 ? cta_start : s_end[tid - 1];
        const bool need = active && ns != s_start[tid];
        const uint32_t bal = __ballot_sync(0xffffffffu, need);
        if (lane == 0) s_woff[warp] = __popc(bal);
        __syncthreads();
        uint32_t base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < JHT / 32; ++w) {
            const uint32_t c = s_woff[w];
            if (w < (int)warp) base += c;
            total += c;
        }
        if (total == 0) break;  // (uniform) the CTA agrees with itself
        if (need) {  // those are packed to the front, so that the decode below runs in full warps however few they are
            const uint32_t pos = base + __popc(bal & ((1u << lane) - 1u));
            s_list[pos] = (uint16_t)tid;
            s_ns[pos] = ns;
        }
        __syncthreads();
        if (tid < total) {
            const uint32_t i = s_list[tid];
            const unsigned long long st = s_ns[tid];
            uint32_t p = (uint32_t)(st >> 32), slot = (uint32_t)(st >> 8) & 0xff, k = (uint32_t)st & 0xff;
            if (slot >= bpm) slot = 0;
            const uint32_t p_end = min((t0 + i + 1) * sub_bits, data_bits);
            const uint32_t n = huff_run<false>(tabs, slotmap, bpm, 0, data, p, slot, k, p_end, HuffOut{}, 0, 0);
            s_start[i] = st;
            s_end[i] = pack_state(p, slot, k);
            s_cnt[i] = n;
            s_dirty[i] = 1;
        }
        __syncthreads();
    }
    if (active) {
        out[gi] = s_end[tid];
        if (s_dirty[tid]) {
            b.start_used[gi] = s_start[tid];
            b.counts[gi] = s_cnt[tid];
        }
    }
}

// Second pass: every subsequence again, from its settled start state, writing the block offsets, the AC entries and the DC
// differences. Where a thread's first block start and first entry land = the blocks and entries counted by every
// subsequence before it: summed over the CTAs before this one, scanned inside.
__global__ void __launch_bounds__(JHT)
jhuff_write_kernel(JpegHuffBatch b, const unsigned long long* __restrict__ fin) {
    __shared__ __align__(16) Tabs tabs;
    __shared__ unsigned long long s_part[JHT / 32], s_warp[JHT / 32];
    const JpegHuffFrame& fr = b.frames[blockIdx.y];
    const uint32_t nsub = fr.nsub, t0 = blockIdx.x * JHT;
    if (t0 >= nsub) return;
    load_tabs(tabs, b.tabsets[fr.tabset]);
    const uint32_t* counts = b.counts + fr.sub_base;
    unsigned long long before = 0;
    for (uint32_t u = threadIdx.x; u < t0; u += JHT) before += widen(counts[u]);
    before = cta_sum(before, s_part);  // (also the barrier after load_tabs)
    const uint32_t t = t0 + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool active = t < nsub;
    const unsigned long long mine = active ? widen(counts[t]) : 0;
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long base = before + incl - mine;
    for (uint32_t w = 0; w < warp; ++w) base += s_warp[w];
    if (!active) return;
    const uint32_t base_blk = (uint32_t)base, base_ent = (uint32_t)(base >> 32);
    const size_t gi = (size_t)fr.sub_base + t;
    const unsigned long long start = t == 0 ? 0ull : fin[gi - 1];
    if (start != b.start_used[gi]) atomicOr(b.status + blockIdx.y, 2);  // not the fixed point: the frame is not settled
    uint32_t p = (uint32_t)(start >> 32), slot = (uint32_t)(start >> 8) & 0xff, k = (uint32_t)start & 0xff;
    if (slot >= fr.blocks_per_mcu) slot = 0;
    const uint32_t p_end = min((t + 1) * fr.sub_bits, fr.data_bits), nblocks = fr.nblocks;
    const HuffOut out{b.offs + fr.offs_base, b.entries + fr.ent_base, b.dcv + fr.offs_base, fr.ent_cap};
    const uint32_t n = huff_run<true>(tabs, fr.slotmap, fr.blocks_per_mcu, nblocks, reinterpret_cast<const uint32_t*>(b.bytes + fr.data_off),
                                      p, slot, k, p_end, out, base_blk, base_ent);
    if (t == nsub - 1) {
        out.offs[nblocks] = base_ent + (n >> 16);  // end of the last block's entries
        if (!(base_blk + (n & 0xffffu) == nblocks && k == 0))  // not exactly the frame's blocks, ending on a block boundary
            atomicOr(b.status + blockIdx.y, 1);
    }
}

// DC differences -> DC values: per component a running sum over its blocks in decode order (the predictor is an int, the
// stored JCOEF its low 16 bits). One CTA per (frame, component): a contiguous piece per thread, summed, scanned, re-walked.
__global__ void __launch_bounds__(JHT)
jhuff_dc_kernel(JpegHuffBatch b) {
    __shared__ uint32_t s_slots[JPEG_MAX_SLOTS], s_cnt;
    __shared__ int s_warp[JHT / 32];
    const JpegHuffFrame& fr = b.frames[blockIdx.x];
    if (b.status[blockIdx.x] != 0) return;  // not every block has a DC difference: the host decoder redoes the frame
    const uint32_t c = blockIdx.y, bpm = fr.blocks_per_mcu;
    if (threadIdx.x == 0) {
        uint32_t cnt = 0;
        for (uint32_t s = 0; s < bpm; ++s)
            if (((fr.slotmap >> (2 * s)) & 3u) == c) s_slots[cnt++] = s;
        s_cnt = cnt;
    }
    __syncthreads();
    const uint32_t cnt = s_cnt;
    if (cnt == 0) return;
    const uint32_t total = fr.nblocks / bpm * cnt, per = (total + JHT - 1) / JHT;
    const uint32_t i0 = min(threadIdx.x * per, total), i1 = min(i0 + per, total);
    int16_t* dcv = b.dcv + fr.offs_base;
    auto at = [&](uint32_t i) -> uint32_t {
        const uint32_t mcu = i / cnt;
        return mcu * bpm + s_slots[i - mcu * cnt];
    };
    int sum = 0;
    for (uint32_t i = i0; i < i1; ++i) sum += dcv[at(i)];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int run = incl - sum;
    for (uint32_t w = 0; w < warp; ++w) run += s_warp[w];
    for (uint32_t i = i0; i < i1; i += 8) {  // loads of a group first: a store in between would order them
        uint32_t idx[8];
        int v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            idx[u] = i + u < i1 ? at(i + u) : 0;
            v[u] = i + u < i1 ? dcv[idx[u]] : 0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            run += v[u];
            if (i + u < i1) dcv[idx[u]] = (int16_t)run;
        }
    }
}

void launch_jhuff_sync(const JpegHuffBatch& b, int frames, uint32_t max_nsub, int first, const unsigned long long* in,
                       unsigned long long* out, cudaStream_t s) {
    jhuff_sync_kernel<<<dim3((max_nsub + JHT - 1) / JHT, frames), JHT, 0, s>>>(b, first, in, out);
}

int jhuff_rounds(uint32_t max_nsub) { return (int)min((max_nsub + JHT - 1) / JHT, (uint32_t)JH_MAX_ROUNDS); }

void launch_jhuff_finish(const JpegHuffBatch& b, int frames, uint32_t max_nsub, const unsigned long long* fin, cudaStream_t s) {
    jhuff_write_kernel<<<dim3((max_nsub + JHT - 1) / JHT, frames), JHT, 0, s>>>(b, fin);
    jhuff_dc_kernel<<<dim3(frames, 3), JHT, 0, s>>>(b);
}

}  // namespace uf
