// kernels_post.cu — K8 (softmax + prior decode) and K9-K11 (threshold, sort, greedy NMS).
//
// K8 restates the in-graph tail of the UltraFace ONNX that tract executes
// (/root/reference/infer_server/src/nn.rs:181):
//   scores = softmax(conf, axis=2)
//   c  = loc[:2] * center_variance * prior[2:] + prior[:2]
//   wh = exp(loc[2:] * size_variance) * prior[2:]
//   boxes = [c - wh/2, c + wh/2]
// K9-K11 restate `postproc` + `non_maximum_suppression` + `iou` + `bbox_area`
// (nn.rs:109-140, 198-260): strict `conf > min_confidence`, processing order = confidence
// descending with ties broken towards the HIGHER prior index (stable ascending sort + pop()
// from the back), suppression on strict `iou > max_iou`, EPS = 1e-7 inside the denominator,
// every f32 operation rounded separately (no FMA) in the reference's left-to-right order.
#include "kernels.h"
#include "pdl.cuh"

namespace uf {

__device__ __forceinline__ void tail_one(const float2 c, const float4 l, const float4 p, float cv, float sv,
                                         float2& score, float4& box) {
    const float m = fmaxf(c.x, c.y);
    const float e0 = expf(c.x - m), e1 = expf(c.y - m);
    const float s = e0 + e1;
    score = make_float2(e0 / s, e1 / s);
    const float cx = __fadd_rn(__fmul_rn(__fmul_rn(l.x, cv), p.z), p.x);
    const float cy = __fadd_rn(__fmul_rn(__fmul_rn(l.y, cv), p.w), p.y);
    const float w = __fmul_rn(expf(__fmul_rn(l.z, sv)), p.z);
    const float h = __fmul_rn(expf(__fmul_rn(l.w, sv)), p.w);
    const float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);
    box = make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
}

__global__ void __launch_bounds__(256)
tail_kernel(const float* __restrict__ conf, const float* __restrict__ loc, long long conf_fs, long long loc_fs,
            const float* __restrict__ priors, int K, float cv, float sv, float* __restrict__ scores,
            float* __restrict__ boxes, long long total) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long f = idx / K;
        const int k = (int)(idx - f * K);
        const float2 c = *reinterpret_cast<const float2*>(conf + f * conf_fs + 2 * (size_t)k);
        const float4 l = *reinterpret_cast<const float4*>(loc + f * loc_fs + 4 * (size_t)k);
        const float4 p = __ldg(reinterpret_cast<const float4*>(priors) + k);
        float2 so;
        float4 bo;
        tail_one(c, l, p, cv, sv, so, bo);
        *reinterpret_cast<float2*>(scores + (size_t)idx * 2) = so;
        *reinterpret_cast<float4*>(boxes + (size_t)idx * 4) = bo;
    }
}

void launch_tail(const float* conf, const float* loc, long long conf_frame_stride, long long loc_frame_stride,
                 const float* priors, int K, float center_var, float size_var, float* scores, float* boxes,
                 int frames, cudaStream_t s) {
    long long total = (long long)frames * K;
    long long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    launch_pdl(tail_kernel, dim3((unsigned)g), dim3(256), 0, s, conf, loc, conf_frame_stride, loc_frame_stride, priors, K, center_var,
               size_var, scores, boxes, total);
}

// ---------------------------------------------------------------------------------------------
// post_kernel: one CTA per frame.
//   1. count candidates (score > min_conf), pick the key array (shared memory if it fits)
//   2. fill keys = (orderable(score) << 32 | prior index), pad to a power of two with 0
//   3. bitonic sort, descending  => position order == the reference's pop() order
//   4. greedy NMS. Few candidates (n <= NMS_SMALL, a real scene): in this CTA, in chunks of PT candidates —
//      (A) every candidate of the chunk is tested against all boxes selected so far, (B) survivors are
//      resolved against each other with a PT x PT bit matrix built by warp ballots and swept by one warp.
//      Many candidates (the NMS-heavy configuration, thousands per frame): the CTA only publishes the sorted
//      candidates; nms_mask_kernel builds the upper-triangular suppression bit matrix of every such frame with
//      the whole GPU (all pairs are independent), nms_sweep_kernel then walks it in order, one CTA per frame,
//      32 candidates per step (the 32x32 diagonal block is resolved in registers with shuffles, the rows of
//      the survivors are OR-ed into the frame's `removed` bitset with all loads in flight at once).
//      Both produce exactly the reference's greedy selection.
// ---------------------------------------------------------------------------------------------
constexpr int PT = 256;      // NMS chunk (candidates resolved together)
constexpr int PG = 4;        // thread groups: group g tests the chunk against the selected boxes s == g (mod PG)
constexpr int PTHR = PT * PG;  // threads per CTA
constexpr int PW = PT / 32;  // mask words per row
constexpr int POST_SMEM_KEYS = 8192;   // sort keys held in shared memory (K <= 8192) ...
constexpr int POST_SMEM_KEYS_BIG = 16384;  // ... or 16384 (larger K: 128 KB of the 227 KB an sm_100 CTA may use)
constexpr int NMS_SMALL = 512;       // more candidates than this: bit-matrix path (needs PostBuffers::mask)
constexpr int MROWS = 64;            // nms_mask_kernel: rows staged per work unit
constexpr int MCOLS = 256;           // ... columns per pass (one per thread; 8 mask words)

// order-preserving map f32 -> u32; -0.0 and +0.0 map to the same key (Rust's partial_cmp calls them equal, so the
// prior index decides, nn.rs:134)
__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = f == 0.0f ? 0u : __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// nn.rs:251-260 (the reference names the y-extent `width` and the x-extent `height`; same product)
__device__ __forceinline__ float bbox_area(float x0, float y0, float x1, float y1) {
    const float width = __fsub_rn(y1, y0);
    const float height = __fsub_rn(x1, x0);
    return (width < 0.0f || height < 0.0f) ? 0.0f : __fmul_rn(width, height);
}

// nn.rs:227-243
__device__ __forceinline__ float iou(const float4 a, const float4 b) {
    const float o = bbox_area(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fminf(a.z, b.z), fminf(a.w, b.w));
    const float d = __fadd_rn(__fsub_rn(__fadd_rn(bbox_area(a.x, a.y, a.z, a.w), bbox_area(b.x, b.y, b.z, b.w)), o), 1.0e-7f);
    return __fdiv_rn(o, d);
}

// `iou(a, b) > max_iou` with the reference's exact arithmetic, skipping the division when the overlap is empty:
// then iou = 0 / (positive) = 0, which cannot exceed a non-negative threshold (exact is false: max_iou < 0).
__device__ __forceinline__ bool iou_exceeds(const float4 a, const float4 b, float max_iou, bool exact) {
    if (!exact) {
        const float w = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
        const float h = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
        if (w < 0.0f || h < 0.0f || __fmul_rn(w, h) == 0.0f) return false;
    }
    return iou(a, b) > max_iou;
}

// The same decision for the bit-matrix kernel, where both areas are known beforehand (area_a = bbox_area of the candidate,
// area_b of the earlier box) and max_iou >= 0: the division is replaced by a comparison against max_iou * denominator
// with a guard band of 1e-6 (8 ulp) on either side — inside the band (one pair in millions) the reference's division decides,
// so the result is the reference's, bit for bit. thr_hi / thr_lo = max_iou * (1 +- 1e-6).
__device__ __forceinline__ bool iou_exceeds_fast(const float4 a, float area_a, const float4 b, float area_b, float max_iou,
                                                 float thr_hi, float thr_lo) {
    const float w = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    const float h = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    const float o = (w < 0.0f || h < 0.0f) ? 0.0f : __fmul_rn(w, h);
    if (o == 0.0f) return false;  // 0 / positive = 0 cannot exceed a non-negative threshold
    const float d = __fadd_rn(__fsub_rn(__fadd_rn(area_a, area_b), o), 1.0e-7f);
    if (o > __fmul_rn(thr_hi, d)) return true;
    if (o < __fmul_rn(thr_lo, d)) return false;
    return __fdiv_rn(o, d) > max_iou;
}

// Raw head outputs for the fused tail (TAIL = true): the CTA first turns its frame's conf / loc rows into
// scores / boxes (same arithmetic as tail_kernel), counting candidates on the way, then carries on with them.
struct TailIn {
    const float* conf;
    const float* loc;
    long long conf_fs, loc_fs;
    const float* priors;
    float cv, sv;
};

template <bool TAIL>
__global__ void __launch_bounds__(PTHR)
post_kernel(float* __restrict__ scores, float* __restrict__ boxes, int K, float min_conf,
            float max_iou, PostBuffers pb, TailIn tin, int smem_keys) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(smem_raw);  // smem_keys
    float4* cbox = reinterpret_cast<float4*>(skeys + smem_keys);                  // PT
    unsigned* mask = reinterpret_cast<unsigned*>(cbox + PT);                      // PT * PW
    unsigned* alive_w = mask + PT * PW;                                           // PW
    int* kept = reinterpret_cast<int*>(alive_w + PW);                             // PT
    float* cscore = reinterpret_cast<float*>(kept + PT);                          // PT
    int* cidx = reinterpret_cast<int*>(cscore + PT);                              // PT
    unsigned* dead = reinterpret_cast<unsigned*>(cidx + PT);                      // PT flags (phase A, OR over groups)
    __shared__ int s_cnt, s_nk;
    const bool exact = !(max_iou >= 0.0f);  // negative / NaN threshold: no shortcut

    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.x;
    float* sc = scores + (size_t)f * K * 2;
    float4* bx = reinterpret_cast<float4*>(boxes + (size_t)f * K * 4);
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    // 1. count (strict >, NaN compares false: nn.rs:127)
    int local = 0;
    if (TAIL) {
        const float* cf = tin.conf + (size_t)f * tin.conf_fs;
        const float* lf = tin.loc + (size_t)f * tin.loc_fs;
        for (int k = tid; k < K; k += PTHR) {
            const float2 c = *reinterpret_cast<const float2*>(cf + 2 * (size_t)k);
            const float4 l = *reinterpret_cast<const float4*>(lf + 4 * (size_t)k);
            const float4 p = __ldg(reinterpret_cast<const float4*>(tin.priors) + k);
            float2 so;
            float4 bo;
            tail_one(c, l, p, tin.cv, tin.sv, so, bo);
            *reinterpret_cast<float2*>(sc + 2 * (size_t)k) = so;
            bx[k] = bo;
            local += so.y > min_conf ? 1 : 0;
        }
    } else {
        for (int k = tid; k < K; k += PTHR) local += sc[2 * (size_t)k + 1] > min_conf ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane == 0 && local) atomicAdd(&s_cnt, local);
    __syncthreads();
    const int n = s_cnt;
    int n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    unsigned long long* keys = (n_pad <= smem_keys) ? skeys : pb.sort_scratch + (size_t)f * pb.sort_cap;
    __syncthreads();
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    // 2. fill (order irrelevant: keys are unique, the sort fixes the order)
    for (int k0 = 0; k0 < K; k0 += PTHR) {
        const int k = k0 + tid;
        float v = 0.f;
        bool c = false;
        if (k < K) { v = sc[2 * (size_t)k + 1]; c = v > min_conf; }
        const unsigned bal = __ballot_sync(0xffffffffu, c);
        int basepos = 0;
        if (lane == 0 && bal) basepos = atomicAdd(&s_cnt, __popc(bal));
        basepos = __shfl_sync(0xffffffffu, basepos, 0);
        if (c) keys[basepos + __popc(bal & ((1u << lane) - 1))] = ((unsigned long long)f2ord(v) << 32) | (unsigned)k;
    }
    for (int i = n + tid; i < n_pad; i += PTHR) keys[i] = 0ull;
    __syncthreads();
    // 3. bitonic sort, descending
    for (int k2 = 2; k2 <= n_pad; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n_pad; i += PTHR) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool desc = (i & k2) == 0;
                    if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    // 4. greedy NMS
    float4* sel = reinterpret_cast<float4*>(pb.sel_boxes) + (size_t)f * K;
    float* dets = pb.dets + (size_t)f * K * 5;
    int* didx = pb.det_idx ? pb.det_idx + (size_t)f * K : nullptr;
    if (pb.mask != nullptr) {
        const bool big = n > NMS_SMALL;
        if (tid == 0) {
            pb.big_n[f] = big ? n : 0;
            if (big) atomicOr(pb.any_big, 1);
        }
        if (big) {  // publish the sorted candidates (boxes in processing order, keys = score | prior) and leave
            unsigned long long* gkeys = pb.sort_scratch + (size_t)f * pb.sort_cap;
            for (int i = tid; i < n; i += PTHR) {
                const unsigned long long key = keys[i];
                if (keys != gkeys) gkeys[i] = key;
                sel[i] = bx[(unsigned)(key & 0xffffffffull)];
            }
            return;
        }
    }
    int n_sel = 0;
    const int grp = tid / PT, ct = tid % PT;  // group, candidate slot within the chunk
    for (int base = 0; base < n; base += PT) {
        const int cnt = min(PT, n - base);
        const bool valid = ct < cnt;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            const unsigned k = (unsigned)(keys[base + ct] & 0xffffffffull);
            b = bx[k];
            if (grp == 0) {
                cscore[ct] = sc[2 * (size_t)k + 1];
                cidx[ct] = (int)k;
                cbox[ct] = b;
            }
        }
        if (grp == 0) dead[ct] = 0u;
        __syncthreads();
        // (A) against everything selected so far: PG groups split the selected list, flags are OR-ed in smem
        bool hit = false;
        if (valid) {
            for (int s = grp; s < n_sel && !hit; s += PG) hit = iou_exceeds(b, sel[s], max_iou, exact);
        }
        if (hit) dead[ct] = 1u;
        __syncthreads();
        bool alive = false;
        if (grp == 0) {
            alive = valid && dead[ct] == 0u;
            const unsigned bal = __ballot_sync(0xffffffffu, alive);
            if (lane == 0) alive_w[warp] = bal;
        }
        __syncthreads();
        // (B) bit matrix: row j, bit i  <=>  j alive, i alive, i after j, iou(i, j) > max_iou; rows split over groups
        if (true) {
            const bool my_alive = valid && ((alive_w[ct >> 5] >> (ct & 31)) & 1u);
            const int cw = ct >> 5;  // mask word this warp contributes to
            for (int j = grp; j < cnt; j += PG) {
                const bool aj = (alive_w[j >> 5] >> (j & 31)) & 1u;  // warp-uniform
                unsigned word = 0;
                if (aj) {
                    const bool pred = my_alive && ct > j && iou_exceeds(b, cbox[j], max_iou, exact);
                    word = __ballot_sync(0xffffffffu, pred);
                }
                if (lane == 0) mask[j * PW + cw] = word;
            }
        }
        __syncthreads();
        if (warp == 0) {
            unsigned rem = lane < PW ? ~alive_w[lane] : 0xffffffffu;
            int nk = 0;
            for (int j = 0; j < cnt; ++j) {
                const unsigned wj = __shfl_sync(0xffffffffu, rem, j >> 5);
                if ((wj >> (j & 31)) & 1u) continue;  // suppressed or dead: warp-uniform
                if (lane < PW) rem |= mask[j * PW + lane];
                if (lane == 0) kept[nk] = j;
                ++nk;
            }
            if (lane == 0) s_nk = nk;
        }
        __syncthreads();
        const int nk = s_nk;
        if (tid < nk) {
            const int j = kept[tid];
            const float4 kb = cbox[j];
            sel[n_sel + tid] = kb;
            float* d = dets + (size_t)(n_sel + tid) * 5;
            d[0] = kb.x; d[1] = kb.y; d[2] = kb.z; d[3] = kb.w; d[4] = cscore[j];
            if (didx) didx[n_sel + tid] = cidx[j];
        }
        n_sel += nk;
        __syncthreads();  // sel[] visible to the next chunk's phase A; cbox/mask free for reuse
    }
    if (tid == 0) pb.counts[f] = n_sel;
}

// ---------------------------------------------------------------------------------------------
// nms_mask_kernel: bit (i, j), j > i, of frame f  <=>  iou(cand_j, cand_i) > max_iou  (candidates in processing
// order). Layout: [frame][row block i/32][column block j/32][j%32] -> u32 whose bit i%32 is (i, j): each COLUMN keeps its 32
// decisions against a block of rows in one word (no transposition, no warp-wide operation in the producer), the 32
// columns of a block are adjacent, so this kernel stores, and the sweep loads, whole 128-byte lines. Work unit = (frame,
// MROWS rows): the rows (and their areas) are staged in shared memory once and every thread walks them for its own column,
// MCOLS columns per pass, starting at the diagonal. Persistent grid: units are dealt round-robin, frames without a
// published candidate list cost one shared-memory read.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MCOLS)
nms_mask_kernel(PostBuffers pb, int K, int frames, float max_iou) {
    __shared__ float4 rows[MROWS];
    __shared__ float rarea[MROWS];
    __shared__ int s_n[MCOLS];
    const bool exact = !(max_iou >= 0.0f);
    const float thr_hi = max_iou * 1.000001f, thr_lo = max_iou * 0.999999f;
    pdl_launch_dependents();
    pdl_wait();
    if (*pb.any_big == 0) return;  // the usual case: no frame of this stage has that many candidates
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = pb.mask_pitch, K32 = (K + 31) >> 5;
    long long unit0 = 0;  // global index of the frame's first unit
    for (int f0 = 0; f0 < frames; f0 += MCOLS) {
        __syncthreads();
        s_n[tid] = f0 + tid < frames ? pb.big_n[f0 + tid] : 0;
        __syncthreads();
        for (int ff = 0; ff < MCOLS && f0 + ff < frames; ++ff) {
            const int n = s_n[ff];
            if (n == 0) continue;
            const int f = f0 + ff;
            const int RT = (n + MROWS - 1) / MROWS;
            const float4* sel = reinterpret_cast<const float4*>(pb.sel_boxes) + (size_t)f * K;
            unsigned* mask = pb.mask + (size_t)f * K32 * W * 32;
            // first unit of this frame that is mine: unit0 + r == blockIdx.x (mod gridDim.x)
            int r = (int)(((long long)blockIdx.x - unit0 % gridDim.x + gridDim.x) % gridDim.x);
            for (; r < RT; r += gridDim.x) {
                const int row0 = r * MROWS;
                const int nrows = min(MROWS, n - row0);
                __syncthreads();
                if (tid < MROWS) {
                    const float4 rb4 = tid < nrows ? sel[row0 + tid] : make_float4(0.f, 0.f, 0.f, 0.f);
                    rows[tid] = rb4;
                    rarea[tid] = bbox_area(rb4.x, rb4.y, rb4.z, rb4.w);
                }
                __syncthreads();
                for (int col0 = row0 / MCOLS * MCOLS; col0 < n; col0 += MCOLS) {
                    const int j = col0 + tid;
                    const bool valid = j < n;
                    const float4 b = valid ? sel[j] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float barea = bbox_area(b.x, b.y, b.z, b.w);
                    const int word = (col0 >> 5) + warp;
                    if (word * 32 >= n) continue;  // warp-uniform
                    // rows this column has to be tested against: the earlier candidates (i < j) that exist (ri < nrows)
                    const int lim = valid ? min(j - row0, nrows) : 0;
#pragma unroll
                    for (int h = 0; h < MROWS / 32; ++h) {
                        const int rb = (row0 >> 5) + h;
                        if (rb * 32 >= n || word < rb) continue;  // past the end / below the diagonal: never read
                        // this column's 32 decisions against the rows of block rb, one bit per row — no warp-wide
                        // operation in the loop: the word is stored column-major (see the layout note above)
                        unsigned bits = 0u;
                        if (exact) {  // negative / NaN threshold: the reference's division, always
#pragma unroll 8
                            for (int i = 0; i < 32; ++i)
                                if (iou_exceeds(b, rows[32 * h + i], max_iou, true)) bits |= 1u << i;
                        } else {
                            // decision o > max_iou * d outside a guard band of +-1e-6 around the threshold; pairs inside the
                            // band (one in millions) are collected in `amb` and settled by the reference's division below
                            unsigned amb = 0u;
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const int ri = 32 * h + i;
                                const float4 r4 = rows[ri];
                                const float ww = __fsub_rn(fminf(b.w, r4.w), fmaxf(b.y, r4.y));
                                const float hh = __fsub_rn(fminf(b.z, r4.z), fmaxf(b.x, r4.x));
                                const float o = __fmul_rn(ww, hh);
                                const float d = __fadd_rn(__fsub_rn(__fadd_rn(barea, rarea[ri]), o), 1.0e-7f);
                                const bool overlap = !(ww < 0.0f) && !(hh < 0.0f) && o != 0.0f;  // NaN sides count as overlap, as in the reference
                                const bool sure = o > __fmul_rn(thr_hi, d);
                                if (overlap && sure) bits |= 1u << i;
                                if (overlap && !sure && !(o < __fmul_rn(thr_lo, d))) amb |= 1u << i;
                            }
                            while (amb) {  // rare, per lane
                                const int i = __ffs(amb) - 1;
                                amb &= amb - 1;
                                if (iou(b, rows[32 * h + i]) > max_iou) bits |= 1u << i;
                            }
                        }
                        const int nvalid = lim - 32 * h;  // rows of this block that exist and precede the column
                        bits &= nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : (1u << nvalid) - 1u);
                        mask[((size_t)rb * W + word) * 32 + lane] = bits;
                    }
                }
            }
            unit0 += RT;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// nms_sweep_kernel: one CTA per frame with a published candidate list; `removed` bitset in shared memory.
// Two levels: a super-chunk of 256 candidates (8 blocks of 32) is resolved by one warp from its 8x8-block diagonal part of
// the matrix held in shared memory — 32 candidates per step in registers (one shuffle each), the survivors' effect on the
// super-chunk's later blocks by one ballot per block — then all warps push the survivors' effect on every candidate beyond
// the super-chunk (whole lines, sixteen loads in flight per lane, one ballot per 32 candidates), so the serial chain per
// 256 candidates is one block load, eight register-resident steps and one parallel push.
// ---------------------------------------------------------------------------------------------
constexpr int SWEEP_THR = 512;
constexpr int SWEEP_WARPS = SWEEP_THR / 32;

__global__ void __launch_bounds__(SWEEP_THR)
nms_sweep_kernel(const float* __restrict__ scores, PostBuffers pb, int K) {
    extern __shared__ unsigned removed[];  // mask_pitch words
    __shared__ unsigned blk[8][8][32];     // [row block q][column block k][column in block] -> bits over the rows of block q
    __shared__ unsigned s_kept[8];
    pdl_launch_dependents();
    pdl_wait();
    const int f = blockIdx.x;
    const int n = pb.big_n[f];
    if (f == 0 && threadIdx.x == 0) *pb.any_big = 0;  // consumed (nms_mask_kernel ran before this kernel)
    if (n == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = pb.mask_pitch, K32 = (K + 31) >> 5;
    const int nw = (n + 31) >> 5;
    const unsigned* __restrict__ mask = pb.mask + (size_t)f * K32 * W * 32;
    const float4* sel = reinterpret_cast<const float4*>(pb.sel_boxes) + (size_t)f * K;
    const unsigned long long* keys = pb.sort_scratch + (size_t)f * pb.sort_cap;
    const float* sc = scores + (size_t)f * K * 2;
    float* dets = pb.dets + (size_t)f * K * 5;
    int* didx = pb.det_idx ? pb.det_idx + (size_t)f * K : nullptr;
    for (int w = tid; w < nw; w += SWEEP_THR) removed[w] = 0u;
    int n_sel = 0;  // tracked by warp 0
    for (int word0 = 0; word0 < nw; word0 += 8) {
        const int nq = min(8, nw - word0);  // row blocks (= words) of this super-chunk
        __syncthreads();
        // 1. diagonal super-block (upper triangle: word k >= row block q)
        for (int p = warp; p < 64; p += SWEEP_WARPS) {
            const int q = p >> 3, k = p & 7;
            if (q < nq && k < nq && k >= q) blk[q][k][lane] = mask[((size_t)(word0 + q) * W + word0 + k) * 32 + lane];
        }
        __syncthreads();
        // 2. eight serial steps of 32 candidates, warp 0. Lane = candidate j of the block; its word of the diagonal block
        //    holds, bit by bit, which earlier candidates i of the same block would suppress it.
        if (warp == 0) {
            for (int q = 0; q < 8; ++q) {
                unsigned kept = 0u;
                if (q < nq) {
                    const int row = 32 * (word0 + q) + lane;
                    const bool in = row < n;
                    // prefetch what the survivors will write (independent of the resolution below)
                    const unsigned long long key = in ? keys[row] : 0ull;
                    const float4 kb = in ? sel[row] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const unsigned col = blk[q][q][lane];
                    unsigned rem = removed[word0 + q];
                    if (32 * (word0 + q) + 32 > n) rem |= 0xffffffffu << (n - 32 * (word0 + q));  // past the end: never kept
#pragma unroll
                    for (int b = 0; b < 32; ++b) {
                        const unsigned cb = __shfl_sync(0xffffffffu, col, b);  // who suppresses candidate b
                        if (!((rem >> b) & 1u) && !(cb & kept)) kept |= 1u << b;
                    }
                    // survivors of this block suppress candidates of the super-chunk's later blocks
                    for (int k = q + 1; k < nq; ++k) {
                        const unsigned r = __ballot_sync(0xffffffffu, (blk[q][k][lane] & kept) != 0u);
                        if (lane == 0) removed[word0 + k] |= r;
                    }
                    __syncwarp();
                    if ((kept >> lane) & 1u) {
                        const int pos = n_sel + __popc(kept & ((1u << lane) - 1u));
                        const unsigned k = (unsigned)(key & 0xffffffffull);
                        float* d = dets + (size_t)pos * 5;
                        d[0] = kb.x; d[1] = kb.y; d[2] = kb.z; d[3] = kb.w; d[4] = sc[2 * (size_t)k + 1];
                        if (didx) didx[pos] = (int)k;
                    }
                    n_sel += __popc(kept);
                }
                if (lane == 0) s_kept[q] = kept;
            }
        }
        __syncthreads();
        // 3. push: a warp owns a word (32 candidates beyond the super-chunk), loads their words against each of the 8 row
        //    blocks (two words per trip: sixteen loads in flight), masks them with the blocks' survivors, one ballot
        const int wfar = word0 + 8;
        unsigned km[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) km[q] = s_kept[q];
        for (int w = wfar + warp; w < nw; w += 2 * SWEEP_WARPS) {
            const int w2 = w + SWEEP_WARPS;
            const bool two = w2 < nw;
            unsigned v[8], u[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                v[q] = q < nq ? mask[((size_t)(word0 + q) * W + w) * 32 + lane] : 0u;
                u[q] = (two && q < nq) ? mask[((size_t)(word0 + q) * W + w2) * 32 + lane] : 0u;
            }
            unsigned a = 0u, b2 = 0u;
#pragma unroll
            for (int q = 0; q < 8; ++q) { a |= v[q] & km[q]; b2 |= u[q] & km[q]; }
            a = __ballot_sync(0xffffffffu, a != 0u);
            b2 = __ballot_sync(0xffffffffu, b2 != 0u);
            if (lane == 0) {
                if (a) removed[w] |= a;
                if (two && b2) removed[w2] |= b2;
            }
        }
    }
    if (tid == 0) pb.counts[f] = n_sel;
}

static int smem_keys_for(int K) { return K <= POST_SMEM_KEYS ? POST_SMEM_KEYS : POST_SMEM_KEYS_BIG; }

static size_t post_smem_bytes(int smem_keys) {
    return (size_t)smem_keys * sizeof(unsigned long long) + PT * sizeof(float4) + PT * PW * sizeof(unsigned) +
           PW * sizeof(unsigned) + PT * sizeof(int) + PT * sizeof(float) + PT * sizeof(int) + PT * sizeof(unsigned);
}

int post_configure() {
    int e = (int)cudaFuncSetAttribute(post_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_smem_bytes(POST_SMEM_KEYS_BIG));
    if (e) return e;
    e = (int)cudaFuncSetAttribute(post_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_smem_bytes(POST_SMEM_KEYS_BIG));
    if (e) return e;
    return (int)cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
}

size_t post_sort_scratch_elems(int K) {
    size_t c = 1;
    while (c < (size_t)K) c <<= 1;
    return c;
}

int post_mask_pitch(int K) { return ((K + 31) / 32 + 3) / 4 * 4; }  // words per row, 16-byte multiple
size_t post_mask_words(int K) { return (size_t)((K + 31) / 32) * post_mask_pitch(K) * 32; }  // per frame
bool post_mask_supported(int K) { return K <= 32768; }             // sweep bitset <= 4 KB, matrix <= 128 MB per frame

void launch_nms_mask(int K, float max_iou, const PostBuffers& pb, int frames, cudaStream_t s) {
    // the whole GPU, as many CTAs as fit (the kernel is bound by issue and shared-memory broadcast latency)
    launch_pdl(nms_mask_kernel, dim3(148 * 4), dim3(MCOLS), 0, s, pb, K, frames, max_iou);
}

void launch_nms_sweep(const float* scores, int K, const PostBuffers& pb, int frames, cudaStream_t s) {
    launch_pdl(nms_sweep_kernel, dim3(frames), dim3(SWEEP_THR), (size_t)pb.mask_pitch * sizeof(unsigned), s, scores, pb, K);
}

void launch_post(const float* scores, const float* boxes, int K, float min_conf, float max_iou,
                 const PostBuffers& pb, int frames, cudaStream_t s) {
    const int sk = smem_keys_for(K);
    launch_pdl(post_kernel<false>, dim3(frames), dim3(PTHR), post_smem_bytes(sk), s, const_cast<float*>(scores),
               const_cast<float*>(boxes), K, min_conf, max_iou, pb, TailIn{}, sk);
}

// tail + post in one launch: scores / boxes are OUTPUTS here (kept for uf_raw_outputs and the hooks)
void launch_tail_post(const float* conf, const float* loc, long long conf_frame_stride, long long loc_frame_stride,
                      const float* priors, float center_var, float size_var, float* scores, float* boxes, int K,
                      float min_conf, float max_iou, const PostBuffers& pb, int frames, cudaStream_t s) {
    TailIn tin{conf, loc, conf_frame_stride, loc_frame_stride, priors, center_var, size_var};
    const int sk = smem_keys_for(K);
    launch_pdl(post_kernel<true>, dim3(frames), dim3(PTHR), post_smem_bytes(sk), s, scores, boxes, K, min_conf, max_iou, pb,
               tin, sk);
}

}  // namespace uf
