// huff_sim.cc — CPU-only test helper (never part of the library): steps the synchronisation rounds of the device Huffman
// decoder (infercam_onnx_b200/csrc/kernels_jpeg_huff.cu) serially on the host, running the SAME symbol-level decoder
// (jpeg_huff_core.h), so that the round logic — guessed states, fixed point, block prefix, write pass, DC scan, status —
// can be checked against the sequential decoder without a GPU. Mirrors the kernels one to one; "thread t" is a loop index.
#include <cstring>
#include <vector>

#include "../../infercam_onnx_b200/csrc/jpeg_decode.h"
#include "../../infercam_onnx_b200/csrc/jpeg_huff_core.h"

using namespace uf;
using namespace uf::jh;

extern "C" int huff_sim(const uint8_t* jpeg, size_t len, int16_t* coefs, size_t cap_blocks, int* rounds, int* status, int max_rounds,
                        int sub_bits) {
    JpegBitstream jb;
    try {
        jpeg_prepare_bitstream(jpeg, len, jb);
    } catch (const JpegError& e) {
        return -e.code;
    }
    if (!jb.gpu_ok) return 1;
    if (cap_blocks < jb.plan.nblocks) return 2;
    if (sub_bits > 0) {  // (the engine picks this per run)
        jb.huff.sub_bits = (uint32_t)sub_bits;
        jb.huff.nsub = (jb.huff.data_bits + jb.huff.sub_bits - 1) / jb.huff.sub_bits;
    }
    const JpegHuffFrame& fr = jb.huff;
    const uint32_t SUB = fr.sub_bits;
    Tabs* tabs = new Tabs;
    jpeg_build_tabset(jb.key, tabs->set);
    for (int i = 0; i < 80; ++i) tabs->zz[i] = zigzag_natural(i);
    const uint32_t slotmap = fr.slotmap, bpm = fr.blocks_per_mcu;
    const uint32_t* d = reinterpret_cast<const uint32_t*>(jb.data.data());
    const uint32_t n = fr.nsub;
    std::vector<unsigned long long> in(n), out(n), start_used(n);
    std::vector<uint32_t> nblk(n), base(n);
    constexpr uint32_t JHT = 256;  // as the kernel
    int iters_max = 0;
    auto launch = [&](bool first) {  // one jhuff_sync_kernel launch: CTAs in any order, they only read `in` of the launch before
        for (uint32_t t0 = 0; t0 < n; t0 += JHT) {
            const uint32_t lanes = std::min(JHT, n - t0);
            const unsigned long long cta_start = t0 == 0 ? 0ull : (first ? pack_state(t0 * SUB, 0, 0) : in[t0 - 1]);
            if (!first && cta_start == start_used[t0]) {
                for (uint32_t i = 0; i < lanes; ++i) out[t0 + i] = in[t0 + i];
                continue;
            }
            unsigned long long my_start[JHT], my_end[JHT], s_end[JHT];
            uint32_t my_n[JHT];
            bool dirty[JHT];
            for (uint32_t i = 0; i < JHT; ++i) {
                my_start[i] = ~0ull; my_end[i] = 0; my_n[i] = 0; dirty[i] = false;
                if (first) my_end[i] = pack_state((t0 + i + 1) * SUB, 0, 0);
                else if (i < lanes) { my_start[i] = start_used[t0 + i]; my_end[i] = in[t0 + i]; }
                s_end[i] = my_end[i];
            }
            for (int iter = 0; iter <= (int)JHT; ++iter) {
                bool any = false, ch[JHT];
                for (uint32_t i = 0; i < JHT; ++i) {  // "threads", all reading the state of before this iteration
                    ch[i] = false;
                    const unsigned long long ns = i == 0 ? cta_start : s_end[i - 1];
                    if (i < lanes && ns != my_start[i]) {
                        uint32_t p = (uint32_t)(ns >> 32), slot = (uint32_t)(ns >> 8) & 0xff, k = (uint32_t)ns & 0xff;
                        if (slot >= fr.blocks_per_mcu) slot = 0;
                        const uint32_t p_end = std::min((t0 + i + 1) * SUB, fr.data_bits);
                        my_n[i] = huff_run<false>(*tabs, slotmap, bpm, 0, d, p, slot, k, p_end, HuffOut{}, 0, 0);
                        my_end[i] = pack_state(p, slot, k);
                        my_start[i] = ns;
                        ch[i] = dirty[i] = true;
                        any = true;
                    }
                }
                // (the loop above read s_end[i - 1] after iteration-local writes would have happened on a GPU only past the barrier:
                //  here ch[] delays the writes)
                for (uint32_t i = 0; i < JHT; ++i)
                    if (ch[i]) s_end[i] = my_end[i];
                iters_max = std::max(iters_max, iter + 1);
                if (!any) break;
            }
            for (uint32_t i = 0; i < lanes; ++i) {
                out[t0 + i] = my_end[i];
                if (dirty[i]) { start_used[t0 + i] = my_start[i]; nblk[t0 + i] = my_n[i]; }
            }
        }
        in.swap(out);
    };
    *rounds = (int)std::min<uint32_t>((n + JHT - 1) / JHT, max_rounds > 0 ? (uint32_t)max_rounds : 4u);  // JH_MAX_ROUNDS
    for (int r = 0; r < *rounds; ++r) launch(r == 0);
    *rounds = *rounds * 1000 + iters_max;  // launches * 1000 + the most in-CTA iterations any CTA took
    *status = 0;
    {
        const uint32_t nb = fr.nblocks, cap = fr.data_bits / 2 + 16;
        std::vector<uint32_t> offs((size_t)nb + 1, 0), ent(cap, 0);
        std::vector<int16_t> dcv(nb, 0);
        const HuffOut ho{offs.data(), ent.data(), dcv.data(), cap};
        unsigned long long acc = 0;  // blocks | entries << 32
        for (uint32_t t = 0; t < n; ++t) {
            base[t] = (uint32_t)acc;
            const uint32_t base_ent = (uint32_t)(acc >> 32);
            acc += (unsigned long long)(nblk[t] & 0xffffu) | ((unsigned long long)(nblk[t] >> 16) << 32);
            const unsigned long long start = t == 0 ? 0ull : in[t - 1];
            if (start != start_used[t]) *status |= 2;
            uint32_t p = (uint32_t)(start >> 32), slot = (uint32_t)(start >> 8) & 0xff, k = (uint32_t)start & 0xff;
            if (slot >= bpm) slot = 0;
            const uint32_t p_end = std::min((t + 1) * SUB, fr.data_bits);
            const uint32_t c = huff_run<true>(*tabs, slotmap, bpm, nb, d, p, slot, k, p_end, ho, base[t], base_ent);
            if (t == n - 1) {
                offs[nb] = base_ent + (c >> 16);
                if (!(base[t] + (c & 0xffffu) == nb && k == 0)) *status |= 1;
            }
        }
        memset(coefs, 0, (size_t)nb * 128);
        if (*status == 0) {
            int pred[3] = {0, 0, 0};
            for (uint32_t b = 0; b < nb; ++b) {
                const int c = (int)((fr.slotmap >> (2 * (b % fr.blocks_per_mcu))) & 3u);
                pred[c] += dcv[b];
                coefs[(size_t)b * 64] = (int16_t)pred[c];
                for (uint32_t e = offs[b]; e < offs[b + 1] && e < cap; ++e) coefs[(size_t)b * 64 + ((ent[e] >> 16) & 63)] = (int16_t)(ent[e] & 0xffffu);
            }
        }
    }
    delete tabs;
    return 0;
}
