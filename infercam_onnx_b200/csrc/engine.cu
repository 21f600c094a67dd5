// engine.cu — model handle, weight prepack, chunked stream pipeline and the C ABI
// (include/ultraface_b200.h). Host-side replacement for `UltrafaceModel::{new,run}`
// (/root/reference/infer_server/src/nn.rs:55-67,178-186): the reference runs one frame at a
// time through tract on one CPU thread; here a batch is cut into chunks, each chunk flows
// H2D -> K1 resize -> conv stack -> K8 tail -> K9-11 post -> D2H on its own CUDA stream
// ("slot"), so the copy of chunk i+1 overlaps the kernels of chunk i. No CPU fallback exists:
// without a CUDA device uf_model_load fails with UF_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/ultraface_b200.h"
#include "jpeg_decode.h"
#include "kernels.h"
#include "onnx_graph.h"
#include "plan.h"
#include "resize_taps.h"

namespace uf {

struct CudaError : public std::exception {
    std::string msg;
    explicit CudaError(std::string m) : msg(std::move(m)) {}
    const char* what() const noexcept override { return msg.c_str(); }
};
struct ArgError : public std::exception {
    std::string msg;
    int code;
    ArgError(int c, std::string m) : msg(std::move(m)), code(c) {}
    const char* what() const noexcept override { return msg.c_str(); }
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                            std::to_string(__LINE__) + ")");                                       \
    } while (0)

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }  // batcher.cc

enum class Impl { Generic, Stem, Depthwise, Pointwise, PointwiseTC, FusedDwTC, FusedDwPw, FusedPix, FusedHead, FusedHead2, FusedTma, SmallDense, SmallDenseTC, Conv3x3Warp, Add, Relu, Copy };

static const char* impl_name(Impl i) {
    switch (i) {
        case Impl::Generic: return "conv_generic";
        case Impl::Stem: return "stem_3x3s2_u8";
        case Impl::Depthwise: return "depthwise3x3";
        case Impl::Pointwise: return "pointwise1x1";
        case Impl::PointwiseTC: return "pointwise1x1_tcgen05";
        case Impl::FusedDwTC: return "fused_dw3x3_pw1x1_tcgen05";
        case Impl::FusedDwPw: return "fused_dw3x3_pw1x1";
        case Impl::FusedPix: return "fused_dw3x3_pw1x1_pix";
        case Impl::FusedHead: return "fused_dw3x3_pw1x1_head";
        case Impl::FusedHead2: return "fused_dw3x3_pw1x1_head_pair";
        case Impl::FusedTma: return "fused_dw3x3_pw1x1_tma";
        case Impl::Conv3x3Warp: return "conv3x3_warp";
        case Impl::SmallDense: return "small_dense3x3";
        case Impl::SmallDenseTC: return "dense3x3_tcgen05";
        case Impl::Add: return "eltwise_add";
        case Impl::Relu: return "eltwise_relu";
        case Impl::Copy: return "eltwise_copy";
    }
    return "?";
}

struct Step {
    Impl impl = Impl::Generic;
    int op = -1, op2 = -1;  // plan op indices (op2: the pointwise of a fused pair)
    int op3 = -1, op4 = -1;  // FusedHead2: depthwise / pointwise of the second head on the same input
    uint64_t alg_bytes = 0;  // SURVEY.md §8(d) convention: sum over Conv nodes of (in+out)*4, per frame
    uint64_t min_bytes = 0;  // compulsory traffic of this launch (fusion removes the intermediate), per frame
    uint64_t flops = 0;      // per frame
    std::string label;       // "<kernel family>[<shape>]" for the per-launch profile
    int tc = -1;             // index into uf_model::tc_weights for PointwiseTC steps
    std::vector<float> host_w;  // FusedTma / Stem / SmallDense: weights in kernel-parameter layout (host copy)
};

struct TcWeights {           // 3xTF32 split of one 1x1 conv's weights, [N][K] K-major, plus their tensor maps
    float *d_hi = nullptr, *d_lo = nullptr;
    TmaMap tm_hi, tm_lo;
};

struct TapsEntry {
    AxisTaps v, h;
    int *d_vleft = nullptr, *d_vn = nullptr, *d_hleft = nullptr, *d_hn = nullptr;
    float *d_vw = nullptr, *d_hw = nullptr;
    ResizeTapsDev dev{};
    uint64_t last_use = 0;
    TapsEntry() = default;
    TapsEntry(const TapsEntry&) = delete;
    TapsEntry& operator=(const TapsEntry&) = delete;
    // cudaFree waits for the device, so kernels already queued with these tables finish first
    ~TapsEntry() { cudaFree(d_vleft); cudaFree(d_vn); cudaFree(d_vw); cudaFree(d_hleft); cudaFree(d_hn); cudaFree(d_hw); }
};
using TapsRef = std::shared_ptr<TapsEntry>;
constexpr size_t TAPS_CACHE_MAX = 32;  // distinct source sizes kept on the device (sizes come from network peers)

constexpr int DET_FAST = 128;  // detections per frame copied back with the counts in one D2H

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done_ev = nullptr;  // blocking-sync event at the end of a host-input stage (see wait_slot)
    bool sleep_wait = false;        // the pending stage recorded done_ev
    uint8_t* d_in = nullptr;
    size_t d_in_cap = 0;
    uint8_t* d_resized = nullptr;
    float* d_arena = nullptr;
    float *d_dets = nullptr, *d_sel = nullptr;
    int *d_counts = nullptr, *d_det_idx = nullptr, *d_big_n = nullptr;
    unsigned long long* d_sort = nullptr;
    unsigned* d_mask = nullptr;  // NMS suppression bit matrix [chunk][K][pitch] (frames with many candidates)
    int* h_counts = nullptr;   // pinned [chunk]
    float* h_dets = nullptr;   // pinned [chunk][DET_FAST][5]
    // N2: per-stage JPEG staging (pinned host side, device side), grown on demand
    uint8_t* h_jpeg = nullptr;
    uint8_t* d_jpeg = nullptr;
    size_t jpeg_cap = 0;
    uint8_t* d_planes = nullptr;
    size_t planes_cap = 0;
    uint8_t* d_ovl = nullptr;    // overlay lists of the frames being annotated (per slot: calls on different lanes run concurrently)
    size_t ovl_cap = 0;
    uint8_t* d_huff = nullptr;   // GPU Huffman scratch: subsequence states, block counts, dense coefficient blocks
    size_t huff_cap = 0;
    int* h_jstatus = nullptr;    // pinned [chunk + 1]: per frame of the stage 0 = decoded on the GPU, else redo on the host; [chunk] = round flag
    uint32_t jstatus_n = 0;      // frames of the stage whose status is meaningful (GPU Huffman path used)
    std::vector<uint32_t> redo;  // global frame indices the GPU Huffman path handed back (collected by harvest)
    // pending work description
    bool pending = false;
    uint32_t first = 0, n = 0;
    uint64_t batch_id = 0;  // which uf_infer_batch* call the slot's activations belong to
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;  // profiling pairs
    std::vector<int> ev_stat;                             // stat index per pair
    size_t ev_used = 0;
    std::vector<TmaMap> tm_a;                             // per step: activation tensor map (PointwiseTC / FusedTma input)
    std::vector<TmaMap> tm_o;                             // per step: output tensor map (FusedTma, PointwiseTC)
    std::vector<uint8_t> tm_o_ok;                         // per step: tm_o is valid
    std::vector<TmaMap> tm_r;                             // per step: residual tensor map (PointwiseTC with Add)
    std::vector<uint8_t> tm_r_ok;
    // CUDA graphs of the kernel chain (conv stack + tail + post + D2H), keyed by (first frame, frames, stem inside?)
    std::map<std::tuple<uint32_t, int, int>, cudaGraphExec_t> graphs;
    std::map<std::tuple<uint32_t, int, int>, int> graph_seen, graph_nodes;
};

// One lane = everything a call mutates: its pipeline slots and the raw outputs of its last batch. Calls from
// different host threads on one handle run on different lanes (weights, plan, tap tables are shared), so the
// kernels of one batch overlap the H2D copies of the next: that is how a stream batcher keeps PCIe busy.
struct Lane {
    std::mutex mu;
    std::vector<Slot> slots;
    float *d_scores = nullptr, *d_boxes = nullptr;  // raw outputs of the lane's last batch [max_batch][K][2|4]
    std::vector<JpegCoefs> jpeg_coefs;              // N2: Huffman-decoded frames of the call in progress (storage reused)
    std::vector<JpegBitstream> jpeg_streams;        // N2: frames prepared for the device Huffman decoder (storage reused)
    uint32_t last_n = 0;
    uint64_t batch_id = 0;
};

// A few host threads for the per-frame serial work in front of the GPU (Huffman decoding): one job at a time, the
// caller works too.
class HostPool {
public:
    HostPool() {
        unsigned n = std::thread::hardware_concurrency();
        n = std::max(1u, std::min(n ? n - 1 : 1u, 31u));
        // several ranks on one host: every rank's pool sized for the whole machine only makes them fight (UF_HOST_THREADS)
        if (const char* e = getenv("UF_HOST_THREADS")) n = std::max(1u, std::min(n, (unsigned)std::max(1, atoi(e))));
        for (unsigned i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void parallel_for(uint32_t n, const std::function<void(uint32_t)>& fn) {
        if (n == 0) return;
        std::unique_lock<std::mutex> job(job_mu_);  // one parallel_for at a time
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; n_ = n; next_ = 0; done_ = 0; ++epoch_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return done_ == n_; });
        fn_ = nullptr;
    }

private:
    void work() {
        for (;;) {
            uint32_t i;
            const std::function<void(uint32_t)>* fn;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (!fn_ || next_ >= n_) return;
                i = next_++;
                fn = fn_;
            }
            (*fn)(i);
            std::lock_guard<std::mutex> lk(mu_);
            if (++done_ == n_) done_cv_.notify_all();
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || epoch_ != seen; });
                if (stop_) return;
                seen = epoch_;
            }
            work();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, job_mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(uint32_t)>* fn_ = nullptr;
    uint32_t n_ = 0, next_ = 0, done_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

struct KernelStat {
    std::string name;
    uint64_t launches = 0;
    double ms = 0;
    uint64_t alg_bytes = 0, min_bytes = 0, flops = 0;
};

}  // namespace uf

using namespace uf;

struct uf_model {
    std::mutex aux_mu;  // tap-table cache, profile statistics, hook scratch
    uf_config cfg{};
    std::string onnx_path;
    Plan plan;
    int K = 0;
    uint32_t chunk = 0, host_chunk = 0, jpeg_chunk = 0, nslots = 0;
    // frames in a host-input stage from which its waiter sleeps on a blocking event instead of spinning (UF_SLEEP_WAIT_FROM).
    // 0 = never, the default: measured on one GPU, a sleeping waiter wakes late enough to cost the in-flight path 40 %, and
    // with 8 ranks on 32 cores it changed nothing (+-1 %). Kept as a knob for hosts with fewer cores than waiting threads.
    uint32_t sleep_wait_from = 0;
    uint32_t jh_threads = 150000, jh_bits = 0;  // device Huffman: threads a run should have / forced bits per thread (tuning knobs)
    std::vector<Step> steps;
    std::vector<size_t> w_off, b_off;  // per plan op, floats into d_weights
    float* d_weights = nullptr;
    uint64_t weight_bytes = 0, workspace_bytes = 0;
    float* d_lut = nullptr;
    float* d_priors = nullptr;
    std::vector<std::unique_ptr<Lane>> lanes;
    std::atomic<uint32_t> lane_rr{0};
    std::map<std::pair<int, int>, TapsRef> taps;
    uint64_t taps_clock = 0;
    std::atomic<uint64_t> launches{0};
    std::atomic<bool> profiling{false};
    std::atomic<uint64_t> jpeg_redone{0};  // frames the device Huffman decoder handed back to the host decoder
    std::atomic<int> fail_after_stages{-1};  // fault injection (uf_debug_fail_after): throw after that many submits
    std::vector<KernelStat> stats;
    std::map<std::string, int> stat_index;
    // hook scratch (uf_postproc / uf_preproc_*), grown on demand
    void* d_hook = nullptr;
    size_t d_hook_cap = 0;
    // N3 text overlay: the glyph atlas handed over by the binding (uf_text_atlas_set)
    std::mutex atlas_mu;
    std::string atlas_chars;
    uint32_t atlas_max_len = 0;
    std::vector<uf_glyph> atlas_glyphs;   // [max_len][chars]
    float* d_atlas = nullptr;             // coverage values
    size_t atlas_n = 0;
    std::vector<uint8_t> tensor_readable;  // per plan tensor: materialised in the arena
    std::vector<TcWeights> tc_weights;
    std::unique_ptr<HostPool> host_pool;  // Huffman decoding workers, created on first use
    HostPool& pool() {
        std::lock_guard<std::mutex> lk(aux_mu);
        if (!host_pool) host_pool.reset(new HostPool());
        return *host_pool;
    }
    bool prestem_ok = false;     // first step is the 3 -> 16 stem kernel and UF_FLAG_NO_PRESTEM is not set
    std::string prestem_label;

    ~uf_model();
};

namespace uf {

static inline int64_t align4(int64_t v) { return (v + 3) / 4 * 4; }

// low part of a 3xTF32 weight split, rounded to nearest tf32 (the tensor core would truncate it: a bias, see tf32_lo)
static inline float tf32_round(float v) {
    uint32_t u;
    memcpy(&u, &v, 4);
    u = (u + 0x1000u) & 0xffffe000u;
    memcpy(&v, &u, 4);
    return v;
}

static TView make_view(const uf_model& m, const Slot& s, int t) {
    const TensorDesc& td = m.plan.tensors[t];
    const BufferDesc& b = m.plan.buffers[td.buf];
    TView v;
    v.p = s.d_arena + b.arena_off * (int64_t)m.chunk + td.base_off;
    v.frame_stride = align4(b.frame_floats);
    v.pix_stride = td.pix_stride;
    v.C = td.C; v.H = td.H; v.W = td.W;
    return v;
}

static bool view_vec_ok(const TensorDesc& td) { return td.C % 4 == 0 && td.pix_stride % 4 == 0 && td.base_off % 4 == 0; }

static int stat_id(uf_model& m, const std::string& name) {
    auto it = m.stat_index.find(name);
    if (it != m.stat_index.end()) return it->second;
    KernelStat k;
    k.name = name;
    m.stats.push_back(k);
    m.stat_index[name] = (int)m.stats.size() - 1;
    return (int)m.stats.size() - 1;
}

// ---- profiling helpers: an event pair around one launch on the slot's stream
struct ProfScope {
    uf_model& m;
    Slot& s;
    bool on;
    ProfScope(uf_model& mm, Slot& ss, const std::string& name, uint64_t alg, uint64_t minb, uint64_t flops, int n_launches = 1)
        : m(mm), s(ss), on(mm.profiling) {
        m.launches += (uint64_t)n_launches;
        if (!on) return;
        std::lock_guard<std::mutex> lk(m.aux_mu);
        int id = stat_id(m, name);
        m.stats[id].launches += (uint64_t)n_launches;
        m.stats[id].alg_bytes += alg;
        m.stats[id].min_bytes += minb;
        m.stats[id].flops += flops;
        if (s.ev_used == s.ev.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            s.ev.emplace_back(a, b);
            s.ev_stat.push_back(0);
        }
        s.ev_stat[s.ev_used] = id;
        cudaEventRecord(s.ev[s.ev_used].first, s.stream);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(s.ev[s.ev_used].second, s.stream);
        s.ev_used++;
    }
};

static void collect_profile(uf_model& m, Slot& s) {
    std::lock_guard<std::mutex> lk(m.aux_mu);
    for (size_t i = 0; i < s.ev_used; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.ev[i].first, s.ev[i].second) == cudaSuccess) m.stats[s.ev_stat[i]].ms += ms;
    }
    s.ev_used = 0;
}

// ---- build: weights, steps, buffers -------------------------------------------------------
static void pack_weights(uf_model& m) {
    std::vector<float> blob;
    m.w_off.assign(m.plan.ops.size(), 0);
    m.b_off.assign(m.plan.ops.size(), 0);
    for (size_t i = 0; i < m.plan.ops.size(); ++i) {
        const Op& op = m.plan.ops[i];
        if (op.kind != OpKind::Conv) continue;
        const int cin_g = op.cin / op.groups, k = op.k;
        while (blob.size() % 4) blob.push_back(0.f);
        m.w_off[i] = blob.size();
        // ONNX [co][ci_g][ky][kx] -> [ky][kx][ci_g][co]: the one layout every conv kernel reads
        // (depthwise: [tap][C]; pointwise: [Cin][Cout]; dense: [ky][kx][ci][co])
        blob.resize(blob.size() + (size_t)k * k * cin_g * op.cout);
        float* dst = blob.data() + m.w_off[i];
        for (int co = 0; co < op.cout; ++co)
            for (int ci = 0; ci < cin_g; ++ci)
                for (int ky = 0; ky < k; ++ky)
                    for (int kx = 0; kx < k; ++kx)
                        dst[((size_t)(ky * k + kx) * cin_g + ci) * op.cout + co] =
                            op.w[(((size_t)co * cin_g + ci) * k + ky) * k + kx];
        while (blob.size() % 4) blob.push_back(0.f);
        m.b_off[i] = blob.size();
        blob.insert(blob.end(), op.b.begin(), op.b.end());
    }
    while (blob.size() % 4) blob.push_back(0.f);
    m.weight_bytes = blob.size() * sizeof(float);
    CK(cudaMalloc(&m.d_weights, std::max<size_t>(m.weight_bytes, 16)));
    CK(cudaMemcpy(m.d_weights, blob.data(), m.weight_bytes, cudaMemcpyHostToDevice));
}

static void build_steps(uf_model& m) {
    const Plan& p = m.plan;
    const bool force_generic = m.cfg.flags & UF_FLAG_FORCE_GENERIC;
    const bool no_fusion = m.cfg.flags & UF_FLAG_NO_FUSION;
    std::vector<int> uses(p.tensors.size(), 0);
    for (auto& op : p.ops) {
        if (op.in >= 0) uses[op.in]++;
        if (op.in2 >= 0) uses[op.in2]++;
    }
    m.tensor_readable.assign(p.tensors.size(), 1);
    m.tensor_readable[0] = 0;  // graph input is u8
    auto bytes_of = [&](int t) { const TensorDesc& d = p.tensors[t]; return 4ull * d.C * d.H * d.W; };
    auto macs_of = [&](const Op& op) {
        const TensorDesc& o = p.tensors[op.out];
        return (uint64_t)o.H * o.W * op.cout * (op.cin / op.groups) * op.k * op.k;
    };
    for (size_t i = 0; i < p.ops.size(); ++i) {
        const Op& op = p.ops[i];
        Step st;
        st.op = (int)i;
        if (op.kind != OpKind::Conv) {
            st.impl = op.kind == OpKind::Add ? Impl::Add : op.kind == OpKind::Relu ? Impl::Relu : Impl::Copy;
            st.min_bytes = bytes_of(op.in) + bytes_of(op.out) + (op.in2 >= 0 ? bytes_of(op.in2) : 0);
            m.steps.push_back(st);
            continue;
        }
        const TensorDesc& in = p.tensors[op.in];
        const TensorDesc& out = p.tensors[op.out];
        st.alg_bytes = bytes_of(op.in) + bytes_of(op.out);
        st.min_bytes = (in.is_input ? 3ull * in.H * in.W : bytes_of(op.in)) + bytes_of(op.out) +
                       (op.in2 >= 0 ? bytes_of(op.in2) : 0);
        st.flops = 2 * macs_of(op);
        st.impl = Impl::Generic;
        const bool dw = op.k == 3 && op.groups == op.cin && op.cin == op.cout && op.pad == 1 && op.dil == 1 &&
                        (op.stride == 1 || op.stride == 2) && op.in2 < 0 && !in.is_input && view_vec_ok(in) && view_vec_ok(out) &&
                        op.cout <= 512;  // depthwise3x3_kernel: one CTA row of 128 threads x 4 channels
        const bool pw = op.k == 1 && op.groups == 1 && op.stride == 1 && op.pad == 0 && !in.is_input && view_vec_ok(in);
        const bool no_tc = m.cfg.flags & UF_FLAG_NO_TC;
        // the TMA view of the activations is 2-D [frames*H*W][K]: frames must be densely packed
        auto tc_ok = [&](const Op& o) {
            const TensorDesc& ti = p.tensors[o.in];
            return !no_tc && o.k == 1 && o.groups == 1 && o.stride == 1 && o.pad == 0 && !ti.is_input && view_vec_ok(ti) &&
                   pointwise_tc_supported(o.cin, o.cout) &&
                   align4(p.buffers[ti.buf].frame_floats) == (int64_t)ti.H * ti.W * ti.pix_stride;
        };
        if (!force_generic) {
            if (in.is_input) {
                if (op.k == 3 && op.stride == 2 && op.pad == 1 && op.dil == 1 && op.groups == 1 && op.cin == 3 &&
                    op.cout == 16 && op.in2 < 0 && view_vec_ok(out))
                    st.impl = Impl::Stem;
            } else if (dw) {
                st.impl = Impl::Depthwise;
                if (!no_fusion && i + 1 < p.ops.size()) {
                    const Op& nx = p.ops[i + 1];
                    const bool nx_pw = nx.kind == OpKind::Conv && nx.k == 1 && nx.groups == 1 && nx.stride == 1 &&
                                       nx.pad == 0 && nx.in == op.out && nx.in2 < 0;
                    const bool pix = fused_dwpw_pix_supported(op.cout, nx.cout, op.stride);
                    const TensorDesc& o2 = p.tensors[nx.out];
                    const bool fusable = nx_pw && uses[op.out] == 1 && !out.in_concat && fused_dwpw_supported(op.cout, nx.cout);
                    // 16/32-channel pairs on the big maps: TMA-pipelined fused kernel (memory-bound)
                    const bool simt_pw = m.cfg.flags & UF_FLAG_TMA_SIMT_PW;
                    const bool tma = fusable && !no_tc &&
                                     (simt_pw ? fused_dwpw_tma_supported(op.cout, nx.cout, op.stride)
                                              : fused_dwpw_tc_supported(op.cout, nx.cout, op.stride)) &&
                                     o2.pix_stride == o2.C && o2.base_off % 4 == 0 && !o2.in_concat;
                    // wide pairs: depthwise kernel + tensor-core GEMM beats the SIMT fusion (those maps are L2-resident)
                    const bool split_tc = nx_pw && !tma && tc_ok(nx) && (op.cout >= 128 || (op.cout == 64 && nx.cout >= 32));
                    // depthwise computed inside the tensor-core kernel's converter warps: correct, but measured ~10 %
                    // slower than depthwise kernel + GEMM (8 converter warps cannot hide the L2 latency of the taps),
                    // so it is opt-in (UF_FLAG_FUSE_DW_TC) until the taps arrive by TMA
                    const bool dw_tc = split_tc && uses[op.out] == 1 && !out.in_concat && op.cout % 32 == 0 &&
                                       (m.cfg.flags & UF_FLAG_FUSE_DW_TC) && !(m.cfg.flags & UF_FLAG_NO_FUSION);
                    // SSD heads on the wide, tiny maps (128 / 256 channels -> <= 16): one constant-weight SIMT kernel instead
                    // of depthwise + GEMM, whose two launches are all fixed cost at that size
                    const bool head_wide = nx_pw && uses[op.out] == 1 && !out.in_concat && !no_tc && op.cout > 64 &&
                                           head_dwpw_supported(op.cout, nx.cout, op.stride) && !(m.cfg.flags & UF_FLAG_FUSE_DW_TC);
                    if ((fusable && !split_tc) || dw_tc || head_wide) {
                        const bool head = head_wide || (pix && !no_tc && head_dwpw_supported(op.cout, nx.cout, op.stride));
                        st.impl = dw_tc ? Impl::FusedDwTC : head_wide ? Impl::FusedHead : tma ? Impl::FusedTma : head ? Impl::FusedHead : pix ? Impl::FusedPix : Impl::FusedDwPw;
                        st.op2 = (int)i + 1;
                        st.alg_bytes += bytes_of(nx.in) + bytes_of(nx.out);
                        st.min_bytes = bytes_of(op.in) + bytes_of(nx.out);
                        st.flops += 2 * macs_of(nx);
                        m.tensor_readable[op.out] = 0;
                        // class + box heads of one map (same input, <= 8 and <= 16 outputs): one launch, one staged tile
                        if (st.impl == Impl::FusedHead && !m.steps.empty() && m.steps.back().impl == Impl::FusedHead) {
                            Step& pv = m.steps.back();
                            const Op& pdw = p.ops[pv.op];
                            const Op& ppw = p.ops[pv.op2];
                            if (pdw.in == op.in && head2_dwpw_supported(op.cout, std::min(ppw.cout, nx.cout), std::max(ppw.cout, nx.cout)) &&
                                std::min(ppw.cout, nx.cout) <= 8) {
                                pv.impl = Impl::FusedHead2;
                                pv.op3 = (int)i;
                                pv.op4 = (int)i + 1;
                                if (nx.cout < ppw.cout) { std::swap(pv.op, pv.op3); std::swap(pv.op2, pv.op4); }  // head A = the narrow one
                                pv.alg_bytes += st.alg_bytes;
                                pv.min_bytes += bytes_of(nx.out);
                                pv.flops += st.flops;
                                ++i;
                                continue;
                            }
                        }
                        m.steps.push_back(st);
                        ++i;
                        continue;
                    }
                }
            } else if (pw) {
                st.impl = tc_ok(op) ? Impl::PointwiseTC : Impl::Pointwise;
            } else if (op.k == 3 && op.stride == 1 && op.pad == op.dil && op.groups == 1 && op.in2 < 0 && !no_tc &&
                       (m.cfg.flags & UF_FLAG_DENSE3_TC) && dense3x3_tc_supported(op.cin, op.cout, op.dil) && view_vec_ok(in)) {
                st.impl = Impl::SmallDenseTC;  // zero-copy im2col on the tensor cores: correct, measured 20 % slower, opt-in
            } else if (op.k == 3 && op.stride == 1 && op.pad == op.dil && op.dil <= 8 && op.groups == 1 && op.in2 < 0 &&
                       small_dense_supported(op.cin, op.cout) && view_vec_ok(in) && view_vec_ok(out)) {
                st.impl = Impl::SmallDense;
            } else if (op.k == 3 && op.stride == 1 && op.pad == op.dil && op.groups == 1 && op.in2 < 0 &&
                       conv3x3_warp_supported(op.cin, op.cout) && view_vec_ok(in)) {
                st.impl = Impl::Conv3x3Warp;
            }
        }
        m.steps.push_back(st);
    }
}

static void build_tc_weights(uf_model& m) {
    for (Step& st : m.steps) {
        if (st.impl != Impl::PointwiseTC && st.impl != Impl::FusedDwTC && st.impl != Impl::FusedTma) continue;
        const Op& op = m.plan.ops[st.impl == Impl::PointwiseTC ? st.op : st.op2];
        const int N = op.cout, K = op.cin;
        std::vector<float> hi((size_t)N * K), lo((size_t)N * K);
        for (size_t i = 0; i < hi.size(); ++i) {  // op.w is [cout][cin][1][1] = [N][K], K contiguous
            uint32_t u;
            memcpy(&u, &op.w[i], 4);
            u &= 0xffffe000u;
            float h;
            memcpy(&h, &u, 4);
            hi[i] = h;
            lo[i] = tf32_round(op.w[i] - h);
        }
        TcWeights t;
        CK(cudaMalloc(&t.d_hi, hi.size() * sizeof(float)));
        CK(cudaMalloc(&t.d_lo, lo.size() * sizeof(float)));
        CK(cudaMemcpy(t.d_hi, hi.data(), hi.size() * sizeof(float), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(t.d_lo, lo.data(), lo.size() * sizeof(float), cudaMemcpyHostToDevice));
        const uint32_t box = (uint32_t)pointwise_tc_n_umma(N);
        const bool fused = st.impl == Impl::FusedTma;  // [N][<= 32] boxes, swizzle span = box width (64 / 128 B)
        const uint32_t kc = (uint32_t)std::min(K, 32);  // one K block = one swizzle span (64 / 128 B) per box
        if (fused ? (!make_tmap_f32_2d_sw(&t.tm_hi, t.d_hi, N, K, (uint64_t)K * 4, kc, (uint32_t)N) ||
                     !make_tmap_f32_2d_sw(&t.tm_lo, t.d_lo, N, K, (uint64_t)K * 4, kc, (uint32_t)N))
                  : (!make_tmap_f32_2d(&t.tm_hi, t.d_hi, N, K, (uint64_t)K * 4, box) ||
                     !make_tmap_f32_2d(&t.tm_lo, t.d_lo, N, K, (uint64_t)K * 4, box)))
            throw CudaError("cuTensorMapEncodeTiled failed for 1x1 weights of '" + m.plan.tensors[op.out].name + "'");
        st.tc = (int)m.tc_weights.size();
        m.tc_weights.push_back(t);
        m.weight_bytes += 2 * hi.size() * sizeof(float);
    }
}

// SmallDenseTC: W[n][c][tap] as [tap][chunk q][n (16)][4 floats] (the no-swizzle K-major operand layout), tf32 hi / lo
static void build_dense3_weights(uf_model& m) {
    for (Step& st : m.steps) {
        if (st.impl != Impl::SmallDenseTC) continue;
        const Op& op = m.plan.ops[st.op];
        const int CQ = ((op.cin + 7) / 8 * 8) / 4;
        std::vector<float> hi(dense3x3_tc_weight_floats(op.cin), 0.f), lo(hi.size(), 0.f);
        for (int n = 0; n < op.cout; ++n)
            for (int c = 0; c < op.cin; ++c)
                for (int tap = 0; tap < 9; ++tap) {
                    const float w = op.w[((size_t)n * op.cin + c) * 9 + tap];  // ONNX [n][c][ky][kx]
                    uint32_t u;
                    memcpy(&u, &w, 4);
                    u &= 0xffffe000u;
                    float h;
                    memcpy(&h, &u, 4);
                    const size_t i = (((size_t)tap * CQ + c / 4) * 16 + n) * 4 + c % 4;
                    hi[i] = h;
                    lo[i] = tf32_round(w - h);
                }
        TcWeights t;
        CK(cudaMalloc(&t.d_hi, hi.size() * sizeof(float)));
        CK(cudaMalloc(&t.d_lo, lo.size() * sizeof(float)));
        CK(cudaMemcpy(t.d_hi, hi.data(), hi.size() * sizeof(float), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(t.d_lo, lo.data(), lo.size() * sizeof(float), cudaMemcpyHostToDevice));
        st.tc = (int)m.tc_weights.size();
        m.tc_weights.push_back(t);
        m.weight_bytes += 2 * hi.size() * sizeof(float);
    }
}

// weights of the FusedTma steps in the layout of FusedWeights<C,N>: [dw 9*C][dw bias][pw C*N (ci major)][pw bias]
static void build_param_weights(uf_model& m) {
    for (Step& st : m.steps) {
        if (st.impl == Impl::Stem || st.impl == Impl::SmallDense) {
            // [ky][kx][ci][co] weights followed by the biases
            const Op& op = m.plan.ops[st.op];
            const int k = op.k, ci_n = op.cin, co_n = op.cout;
            st.host_w.assign((size_t)k * k * ci_n * co_n + co_n, 0.f);
            for (int co = 0; co < co_n; ++co)
                for (int ci = 0; ci < ci_n; ++ci)
                    for (int ky = 0; ky < k; ++ky)
                        for (int kx = 0; kx < k; ++kx)
                            st.host_w[((size_t)(ky * k + kx) * ci_n + ci) * co_n + co] = op.w[(((size_t)co * ci_n + ci) * k + ky) * k + kx];
            for (int co = 0; co < co_n; ++co) st.host_w[(size_t)k * k * ci_n * co_n + co] = op.b[co];
            continue;
        }
        if (st.impl != Impl::FusedTma && st.impl != Impl::FusedHead && st.impl != Impl::FusedHead2) continue;
        for (int half = 0; half < (st.impl == Impl::FusedHead2 ? 2 : 1); ++half) {
        const Op& dw = m.plan.ops[half ? st.op3 : st.op];
        const Op& pw = m.plan.ops[half ? st.op4 : st.op2];
        const int C = dw.cout, Nreal = pw.cout;
        // heads: outputs zero-padded to 8 / 16 (a pair: 8 for the first head, 16 for the second)
        const int N = st.impl == Impl::FusedHead2 ? (half ? 16 : 8) : st.impl == Impl::FusedHead ? (Nreal <= 8 ? 8 : 16) : Nreal;
        const size_t base = st.host_w.size();
        st.host_w.resize(base + fused_dwpw_tma_weight_floats(C, N), 0.f);
        float* p = st.host_w.data() + base;
        for (int t = 0; t < 9; ++t)
            for (int c = 0; c < C; ++c) p[t * C + c] = dw.w[(size_t)c * 9 + t];  // ONNX [c][1][ky][kx]
        for (int c = 0; c < C; ++c) p[9 * C + c] = dw.b[c];
        float* q = p + 10 * C;
        for (int ci = 0; ci < C; ++ci)
            for (int n = 0; n < Nreal; ++n) q[ci * N + n] = pw.w[(size_t)n * C + ci];  // ONNX [n][ci][1][1]
        for (int n = 0; n < Nreal; ++n) q[C * N + n] = pw.b[n];
        }
    }
}

static void label_steps(uf_model& m) {
    const Plan& p = m.plan;
    for (Step& st : m.steps) {
        const Op& op = p.ops[st.op];
        const TensorDesc& in = p.tensors[op.in];
        const TensorDesc& out = p.tensors[st.op2 >= 0 ? p.ops[st.op2].out : op.out];
        char buf[96];
        snprintf(buf, sizeof(buf), "%s[%d>%d k%d s%d d%d %dx%d]", impl_name(st.impl), in.C, out.C, op.k, op.stride, op.dil,
                 out.W, out.H);
        st.label = buf;
    }
}

static void build_lut(uf_model& m) {
    // nn.rs:85-88: (px as f32 / 255.0 - mean[c]) / std[c]; f32 operations in that order
    static const float mean[3] = {0.485f, 0.456f, 0.406f};
    static const float stdv[3] = {0.229f, 0.224f, 0.225f};
    std::vector<float> lut(768);
    for (int c = 0; c < 3; ++c)
        for (int v = 0; v < 256; ++v) {
            float r;
            if (m.cfg.norm_preset == UF_NORM_127_128) {
                volatile float d = (float)v - 127.0f;
                r = d / 128.0f;
            } else {
                volatile float q = (float)v / 255.0f;
                volatile float d = q - mean[c];
                r = d / stdv[c];
            }
            lut[c * 256 + v] = r;
        }
    CK(cudaMalloc(&m.d_lut, 768 * sizeof(float)));
    CK(cudaMemcpy(m.d_lut, lut.data(), 768 * sizeof(float), cudaMemcpyHostToDevice));
}

static void alloc_lane(uf_model& m, Lane& ln) {
    const int K = m.K;
    const size_t H = m.plan.net_h, W = m.plan.net_w;
    const size_t sort_cap = post_sort_scratch_elems(K);
    ln.slots.resize(m.nslots);
    uint64_t ws = 0;
    for (auto& s : ln.slots) {
        CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&s.done_ev, cudaEventBlockingSync | cudaEventDisableTiming));
        s.d_in_cap = (size_t)m.chunk * 640 * 480 * 3;
        CK(cudaMalloc(&s.d_in, s.d_in_cap));
        CK(cudaMalloc(&s.d_resized, (size_t)m.chunk * H * W * 3));
        const size_t arena = (size_t)m.plan.arena_frame_floats * m.chunk * sizeof(float);
        CK(cudaMalloc(&s.d_arena, arena));
        CK(cudaMemsetAsync(s.d_arena, 0, arena, s.stream));
        CK(cudaMalloc(&s.d_dets, (size_t)m.chunk * K * 5 * sizeof(float)));
        CK(cudaMalloc(&s.d_sel, (size_t)m.chunk * K * 4 * sizeof(float)));
        CK(cudaMalloc(&s.d_det_idx, (size_t)m.chunk * K * sizeof(int)));
        CK(cudaMalloc(&s.d_counts, (size_t)m.chunk * sizeof(int)));
        CK(cudaMalloc(&s.d_sort, (size_t)m.chunk * sort_cap * sizeof(unsigned long long)));
        CK(cudaMalloc(&s.d_big_n, ((size_t)m.chunk + 1) * sizeof(int)));  // [chunk] + the stage's any_big flag
        CK(cudaMemsetAsync(s.d_big_n, 0, ((size_t)m.chunk + 1) * sizeof(int), s.stream));
        if (post_mask_supported(K)) {
            const size_t mb = (size_t)m.chunk * post_mask_words(K) * sizeof(unsigned);
            CK(cudaMalloc(&s.d_mask, mb));
            ws += mb;
        }
        CK(cudaMallocHost(&s.h_counts, (size_t)m.chunk * sizeof(int)));
        CK(cudaMallocHost(&s.h_dets, (size_t)m.chunk * DET_FAST * 5 * sizeof(float)));
        CK(cudaMallocHost(&s.h_jstatus, ((size_t)m.chunk + 1) * sizeof(int)));
        ws += s.d_in_cap + (size_t)m.chunk * H * W * 3 + arena + (size_t)m.chunk * K * (5 + 4 + 1) * 4 +
              (size_t)m.chunk * sort_cap * 8;
        CK(cudaStreamSynchronize(s.stream));
    }
    for (auto& s : ln.slots) {
        s.tm_a.resize(m.steps.size());
        s.tm_o.resize(m.steps.size());
        s.tm_o_ok.assign(m.steps.size(), 0);
        s.tm_r.resize(m.steps.size());
        s.tm_r_ok.assign(m.steps.size(), 0);
        for (size_t i = 0; i < m.steps.size(); ++i) {
            if (m.steps[i].impl == Impl::FusedTma) {
                const Op& dw = m.plan.ops[m.steps[i].op];
                const Op& pw = m.plan.ops[m.steps[i].op2];
                int iw, ih, ow, oh;
                if (m.cfg.flags & UF_FLAG_TMA_SIMT_PW) fused_dwpw_tma_boxes(dw.stride, &iw, &ih, &ow, &oh);
                else fused_dwpw_tc_boxes(dw.stride, &iw, &ih, &ow, &oh);
                const int sc = (m.cfg.flags & UF_FLAG_TMA_SIMT_PW) ? 16 : fused_dwpw_tc_slice_channels(dw.cout);
                if (!make_tmap_nhwc(&s.tm_a[i], make_view(m, s, dw.in), (int)m.chunk, sc, iw, ih, sc * 4) ||
                    !make_tmap_nhwc(&s.tm_o[i], make_view(m, s, pw.out), (int)m.chunk, 32, ow, oh, 128))
                    throw CudaError("cuTensorMapEncodeTiled failed for fused layer '" + m.plan.tensors[pw.out].name + "'");
                continue;
            }
            if (m.steps[i].impl != Impl::PointwiseTC && m.steps[i].impl != Impl::FusedDwTC) continue;
            const bool dwtc = m.steps[i].impl == Impl::FusedDwTC;
            const Op& op = m.plan.ops[dwtc ? m.steps[i].op2 : m.steps[i].op];
            if (!dwtc) {
                TView v = make_view(m, s, op.in);
                if (!make_tmap_f32_2d(&s.tm_a[i], v.p, (uint64_t)m.chunk * v.H * v.W, (uint64_t)v.C, (uint64_t)v.pix_stride * 4, 128))
                    throw CudaError("cuTensorMapEncodeTiled failed for the activations of '" + m.plan.tensors[op.out].name + "'");
            }
            // TMA-store epilogue when the output rows are 16-byte aligned and frames are densely packed
            TView o = make_view(m, s, op.out);
            const TensorDesc& od = m.plan.tensors[op.out];
            auto storable = [&](const TView& t, const TensorDesc& td) {
                return t.pix_stride % 4 == 0 && td.base_off % 4 == 0 && t.C % 4 == 0 &&
                       t.frame_stride == (long long)t.H * t.W * t.pix_stride;
            };
            if (storable(o, od))
                s.tm_o_ok[i] = make_tmap_f32_2d_store(&s.tm_o[i], o.p, (uint64_t)m.chunk * o.H * o.W, (uint64_t)o.C,
                                                      (uint64_t)o.pix_stride * 4) ? 1 : 0;
            if (op.in2 >= 0) {
                TView r = make_view(m, s, op.in2);
                if (storable(r, m.plan.tensors[op.in2]))
                    s.tm_r_ok[i] = make_tmap_f32_2d_store(&s.tm_r[i], r.p, (uint64_t)m.chunk * r.H * r.W, (uint64_t)r.C,
                                                          (uint64_t)r.pix_stride * 4) ? 1 : 0;
            }
        }
    }
    CK(cudaMalloc(&ln.d_scores, (size_t)m.cfg.max_batch * K * 2 * sizeof(float)));
    CK(cudaMalloc(&ln.d_boxes, (size_t)m.cfg.max_batch * K * 4 * sizeof(float)));
    ws += (size_t)m.cfg.max_batch * K * 6 * sizeof(float);
    m.workspace_bytes += ws;
}

static TapsRef get_taps(uf_model& m, int sw, int sh) {
    std::lock_guard<std::mutex> lk(m.aux_mu);
    auto key = std::make_pair(sw, sh);
    auto it = m.taps.find(key);
    if (it != m.taps.end()) { it->second->last_use = ++m.taps_clock; return it->second; }
    TapsRef ep = std::make_shared<TapsEntry>();  // a CK throw below frees what was already allocated
    TapsEntry& e = *ep;
    e.v = build_axis_taps(sh, m.plan.net_h);
    e.h = build_axis_taps(sw, m.plan.net_w);
    auto up_i = [](const std::vector<int32_t>& v, int** d) {
        CK(cudaMalloc(d, v.size() * sizeof(int)));
        CK(cudaMemcpy(*d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
    };
    auto up_f = [](const std::vector<float>& v, float** d) {
        CK(cudaMalloc(d, v.size() * sizeof(float)));
        CK(cudaMemcpy(*d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    };
    // device tables: zero-padded to a pitch that is a multiple of 4 floats (16-byte rows; the <= 4-tap fast path)
    auto padded = [](const AxisTaps& a, int* pitch) {
        *pitch = (a.max_taps + 3) / 4 * 4;
        std::vector<float> w((size_t)a.dst_len * *pitch, 0.f);
        for (int o = 0; o < a.dst_len; ++o)
            for (int i = 0; i < a.max_taps; ++i) w[(size_t)o * *pitch + i] = a.w[(size_t)o * a.max_taps + i];
        return w;
    };
    int vpitch = 0, hpitch = 0;
    up_i(e.v.left, &e.d_vleft); up_i(e.v.ntaps, &e.d_vn); up_f(padded(e.v, &vpitch), &e.d_vw);
    up_i(e.h.left, &e.d_hleft); up_i(e.h.ntaps, &e.d_hn); up_f(padded(e.h, &hpitch), &e.d_hw);
    // CTA tile: up to 320 x 8 destination pixels (a whole row of the 320-wide net: the source span is then the
    // whole source row, every 32-lane trip of the vertical pass is full and nothing is read twice horizontally),
    // narrowed until the f32 intermediate fits ~64 KB of shared memory (3 CTAs per SM)
    int tw = std::min(m.plan.net_w, 320), th = 8;
    while (tw % 4) ++tw;
    int cols = max_tile_span(e.h, tw);
    while ((size_t)th * cols * 3 * sizeof(float) > 64 * 1024 && (tw > 8 || th > 1)) {
        if (tw > 64) tw = (tw / 2 + 3) / 4 * 4;  // keep th = 8 (the fast kernel's row count) as long as possible
        else if (th > 1) th /= 2;
        else tw /= 2;
        cols = max_tile_span(e.h, tw);
    }
    if ((size_t)th * cols * 3 * sizeof(float) > 200 * 1024)
        throw ArgError(UF_ERR_UNSUPPORTED, "resize ratio too large for the shared-memory tile (source " +
                                               std::to_string(sw) + "x" + std::to_string(sh) + ")");
    e.dev = ResizeTapsDev{e.d_vleft, e.d_vn, e.d_vw, vpitch, e.d_hleft, e.d_hn, e.d_hw, hpitch, tw, th, cols,
                          tw > 64 ? 64 : 0, tw > 64 ? max_tile_span(e.h, 64) : 0};
    if (m.taps.size() >= TAPS_CACHE_MAX) {  // evict the least recently used size (callers still using it hold a reference)
        auto old = m.taps.begin();
        for (auto jt = m.taps.begin(); jt != m.taps.end(); ++jt)
            if (jt->second->last_use < old->second->last_use) old = jt;
        m.taps.erase(old);
    }
    e.last_use = ++m.taps_clock;
    m.taps.emplace(key, ep);
    return ep;
}

// ---- execution -----------------------------------------------------------------------------
static void run_cnn(uf_model& m, Slot& s, const U8View& input, int frames, size_t first_step = 0,
                    size_t last_step = (size_t)-1) {
    const Plan& p = m.plan;
    for (size_t si = first_step; si < m.steps.size() && si < last_step; ++si) {
        const Step& st = m.steps[si];
        const Op& op = p.ops[st.op];
        const bool fused_pre = st.impl == Impl::Stem && input.W == 2 * p.net_w && input.H == 2 * p.net_h;
        // fused resize + normalise + stem: SURVEY.md 8(d) preproc bytes (u8 source in, f32 NCHW out) on top of the conv's
        const uint64_t pre_alg = fused_pre ? (uint64_t)input.W * input.H * 3 + 12ull * p.net_w * p.net_h : 0;
        const uint64_t pre_min = fused_pre ? (uint64_t)input.W * input.H * 3 - 3ull * p.net_w * p.net_h : 0;
        ProfScope ps(m, s, fused_pre ? m.prestem_label : st.label, (st.alg_bytes + pre_alg) * frames, (st.min_bytes + pre_min) * frames,
                     st.flops * frames);
        const float* w = m.d_weights + m.w_off[st.op];
        const float* b = m.d_weights + m.b_off[st.op];
        TView out = make_view(m, s, op.out);
        TView in = p.tensors[op.in].is_input ? TView{} : make_view(m, s, op.in);
        TView res;
        if (op.in2 >= 0) res = make_view(m, s, op.in2);
        switch (st.impl) {
            case Impl::Generic: {
                ConvParams cp{op.cin, op.cout, op.k, op.stride, op.pad, op.dil, op.groups, op.relu ? 1 : 0};
                launch_conv_generic(in, p.tensors[op.in].is_input ? &input : nullptr, m.d_lut, out,
                                    op.in2 >= 0 ? &res : nullptr, w, b, cp, frames, s.stream);
                break;
            }
            case Impl::Stem:
                if (input.W == 2 * p.net_w && input.H == 2 * p.net_h) {  // the frames themselves: resize + normalise + stem in one kernel
                    TapsRef t = get_taps(m, input.W, input.H);
                    launch_prestem(input, m.d_lut, t->dev, out, st.host_w.data(), op.relu, (int)m.cfg.resize_round_intermediate, frames,
                                   nullptr, s.stream);
                } else {
                    launch_stem(input, m.d_lut, out, st.host_w.data(), op.relu, frames, s.stream);
                }
                break;
            case Impl::Depthwise: launch_depthwise(in, out, w, b, op.stride, op.relu, frames, s.stream); break;
            case Impl::Pointwise:
                launch_pointwise(in, out, op.in2 >= 0 ? &res : nullptr, w, b, op.relu, frames, s.stream);
                break;
            case Impl::PointwiseTC: {
                const TcWeights& tw = m.tc_weights[st.tc];
                launch_pointwise_tc(s.tm_a[si], tw.tm_hi, tw.tm_lo, s.tm_o_ok[si] ? &s.tm_o[si] : nullptr,
                                    s.tm_r_ok[si] ? &s.tm_r[si] : nullptr, in, out, op.in2 >= 0 ? &res : nullptr, op.b.data(),
                                    op.relu, frames, s.stream);
                break;
            }
            case Impl::FusedDwTC: {
                const Op& pw = p.ops[st.op2];
                const TcWeights& tw = m.tc_weights[st.tc];
                TView o2 = make_view(m, s, pw.out);
                TView dummy = o2;
                dummy.C = pw.cin;
                TcDepthwise dwp{in, w, b, op.stride, op.relu ? 1 : 0};
                launch_pointwise_tc(tw.tm_hi, tw.tm_hi, tw.tm_lo, s.tm_o_ok[si] ? &s.tm_o[si] : nullptr, nullptr, dummy, o2, nullptr,
                                    pw.b.data(), pw.relu, frames, s.stream, &dwp);
                break;
            }
            case Impl::FusedDwPw: {
                const Op& pw = p.ops[st.op2];
                TView o2 = make_view(m, s, pw.out);
                launch_fused_dwpw(in, o2, w, b, op.stride, op.relu, m.d_weights + m.w_off[st.op2],
                                  m.d_weights + m.b_off[st.op2], pw.relu, frames, s.stream);
                break;
            }
            case Impl::FusedPix: {
                const Op& pw = p.ops[st.op2];
                TView o2 = make_view(m, s, pw.out);
                launch_fused_dwpw_pix(in, o2, w, b, op.stride, op.relu, m.d_weights + m.w_off[st.op2],
                                      m.d_weights + m.b_off[st.op2], pw.relu, frames, s.stream);
                break;
            }
            case Impl::FusedHead: {
                const Op& pw = p.ops[st.op2];
                TView o2 = make_view(m, s, pw.out);
                launch_head_dwpw(in, o2, st.host_w.data(), op.relu, pw.relu, frames, s.stream);
                break;
            }
            case Impl::FusedHead2: {
                const Op& pwa = p.ops[st.op2];
                const Op& dwb = p.ops[st.op3];
                const Op& pwb = p.ops[st.op4];
                const int relu_bits = (op.relu ? 1 : 0) | (pwa.relu ? 2 : 0) | (dwb.relu ? 4 : 0) | (pwb.relu ? 8 : 0);
                launch_head2_dwpw(in, make_view(m, s, pwa.out), make_view(m, s, pwb.out), st.host_w.data(), relu_bits, frames, s.stream);
                break;
            }
            case Impl::FusedTma: {
                const Op& pw = p.ops[st.op2];
                TView o2 = make_view(m, s, pw.out);
                if (m.cfg.flags & UF_FLAG_TMA_SIMT_PW) {
                    launch_fused_dwpw_tma(s.tm_a[si], s.tm_o[si], in, o2, st.host_w.data(), op.stride, op.relu, pw.relu, frames, s.stream);
                } else {
                    const TcWeights& tw = m.tc_weights[st.tc];
                    launch_fused_dwpw_tc(s.tm_a[si], s.tm_o[si], tw.tm_hi, tw.tm_lo, in, o2, st.host_w.data(), op.stride, op.relu,
                                         pw.relu, frames, s.stream);
                }
                break;
            }
            case Impl::SmallDenseTC: {
                const TcWeights& tw = m.tc_weights[st.tc];
                launch_dense3x3_tc(in, out, tw.d_hi, tw.d_lo, op.b.data(), op.dil, op.relu, frames, s.stream);
                break;
            }
            case Impl::SmallDense: launch_small_dense(in, out, st.host_w.data(), op.dil, op.relu, frames, s.stream); break;
            case Impl::Conv3x3Warp: launch_conv3x3_warp(in, out, w, b, op.dil, op.relu, frames, s.stream); break;
            case Impl::Add: launch_add(in, res, out, op.relu, frames, s.stream); break;
            case Impl::Relu: launch_relu(in, out, frames, s.stream); break;
            case Impl::Copy: launch_copy(in, out, frames, s.stream); break;
        }
    }
}

static void wait_slot(uf_model& m, Slot& s) {
    if (s.sleep_wait) {
        s.sleep_wait = false;
        CK(cudaEventSynchronize(s.done_ev));
    }
    CK(cudaStreamSynchronize(s.stream));
    if (m.profiling) collect_profile(m, s);
}

// copy results of a finished slot into the caller's arrays
static void harvest(uf_model& m, Slot& s, uf_det* out, uint32_t cap, uint32_t* n_out) {
    if (!s.pending) return;
    wait_slot(m, s);
    s.pending = false;
    const int K = m.K;
    uint32_t max_take = 0;
    for (uint32_t i = 0; i < s.jstatus_n && i < s.n; ++i)
        if (s.h_jstatus[i] != 0) s.redo.push_back(s.first + i);  // GPU Huffman declined this frame: its detections are void
    s.jstatus_n = 0;
    for (uint32_t i = 0; i < s.n; ++i) {
        const uint32_t cnt = (uint32_t)s.h_counts[i];
        const uint32_t g = s.first + i;
        if (n_out) n_out[g] = cnt;
        if (!out || cap == 0) continue;
        const uint32_t take = std::min(cnt, cap);
        max_take = std::max(max_take, take);
        memcpy(out + (size_t)g * cap, s.h_dets + (size_t)i * DET_FAST * 5, (size_t)std::min<uint32_t>(take, DET_FAST) * sizeof(uf_det));
    }
    // frames with more than DET_FAST faces (the NMS-heavy configuration): ONE strided copy for the whole stage,
    // sized from the largest count, straight into the caller's rows (entries past n_out[i] are unspecified)
    if (max_take > DET_FAST) {
        CK(cudaMemcpy2DAsync(out + (size_t)s.first * cap + DET_FAST, (size_t)cap * sizeof(uf_det), s.d_dets + (size_t)DET_FAST * 5,
                             (size_t)K * sizeof(uf_det), (size_t)(max_take - DET_FAST) * sizeof(uf_det), s.n,
                             cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
    }
}

// tail + post + D2H for `frames` frames whose conv outputs sit in slot s; global frame offset `first`
static void run_tail_post(uf_model& m, Lane& ln, Slot& s, uint32_t first, int frames) {
    const int K = m.K;
    TView conf, loc;
    {
        const BufferDesc& cb = m.plan.buffers[m.plan.conf_buf];
        const BufferDesc& lb = m.plan.buffers[m.plan.loc_buf];
        conf.p = s.d_arena + cb.arena_off * (int64_t)m.chunk; conf.frame_stride = align4(cb.frame_floats);
        loc.p = s.d_arena + lb.arena_off * (int64_t)m.chunk; loc.frame_stride = align4(lb.frame_floats);
    }
    float* scores = ln.d_scores + (size_t)first * K * 2;
    float* boxes = ln.d_boxes + (size_t)first * K * 4;
    PostBuffers pb{s.d_sort, (int)post_sort_scratch_elems(K), s.d_sel, s.d_dets, s.d_det_idx, s.d_counts,
                   s.d_mask, post_mask_pitch(K), s.d_big_n, s.d_big_n + m.chunk};
    // one CTA per frame decodes its own priors before the NMS: worth a launch when there are enough frames to fill the
    // GPU (or just one or two: batch-1 latency); a 32-frame stage of the 640x480 net (K = 17 640) decodes faster spread
    // over all SMs by the separate kernel
    const bool fuse_tail = !(m.cfg.flags & UF_FLAG_NO_FUSION) && (frames >= 96 || frames <= 2);
    if (!fuse_tail) {
        {
            ProfScope ps(m, s, "tail_softmax_decode", (uint64_t)frames * K * 6 * 4 * 2, (uint64_t)frames * K * 6 * 4 * 2, 0);
            launch_tail(conf.p, loc.p, conf.frame_stride, loc.frame_stride, m.d_priors, K, m.plan.center_variance,
                        m.plan.size_variance, scores, boxes, frames, s.stream);
        }
        ProfScope ps(m, s, "post_threshold_sort_nms", (uint64_t)frames * K * 6 * 4, (uint64_t)frames * K * 6 * 4, 0);
        launch_post(scores, boxes, K, m.cfg.min_confidence, m.cfg.max_iou, pb, frames, s.stream);
    } else {
        // one launch: each frame's CTA decodes its own priors, then thresholds / sorts / suppresses them
        ProfScope ps(m, s, "tail_post_softmax_decode_nms", (uint64_t)frames * K * 6 * 4 * 2, (uint64_t)frames * K * 6 * 4 * 2, 0);
        launch_tail_post(conf.p, loc.p, conf.frame_stride, loc.frame_stride, m.d_priors, m.plan.center_variance,
                         m.plan.size_variance, scores, boxes, K, m.cfg.min_confidence, m.cfg.max_iou, pb, frames, s.stream);
    }
    if (pb.mask) {  // frames with more than a few hundred candidates: suppression bit matrix, then the ordered sweep
        {
            ProfScope ps(m, s, "nms_bitmatrix", 0, 0, 0);
            launch_nms_mask(K, m.cfg.max_iou, pb, frames, s.stream);
        }
        ProfScope ps(m, s, "nms_sweep", 0, 0, 0);
        launch_nms_sweep(scores, K, pb, frames, s.stream);
    }
    CK(cudaMemcpyAsync(s.h_counts, s.d_counts, (size_t)frames * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpy2DAsync(s.h_dets, (size_t)DET_FAST * 5 * sizeof(float), s.d_dets, (size_t)K * 5 * sizeof(float),
                         (size_t)std::min(DET_FAST, K) * 5 * sizeof(float), frames, cudaMemcpyDeviceToHost, s.stream));
}

// conv stack + tail + post + D2H for one chunk. The chain is ~45 short kernels; replaying it as a CUDA graph
// removes the per-launch CPU cost and most of the inter-kernel gaps (this is what bounds batch-1 latency).
// The first step is kept outside the graph when it reads the caller's device buffer (pointer changes per call).
static void run_body(uf_model& m, Lane& ln, Slot& s, const U8View& input, uint32_t first, int frames) {
    const bool use_graph = !m.profiling && !(m.cfg.flags & UF_FLAG_NO_GRAPH);
    const bool stem_inside = input.p == s.d_resized;
    if (!use_graph) {
        run_cnn(m, s, input, frames);
        run_tail_post(m, ln, s, first, frames);
        CK(cudaGetLastError());  // launch-configuration errors are not sticky: ask for them here
        return;
    }
    const auto key = std::make_tuple(first, frames, stem_inside ? 1 : 0);
    auto it = s.graphs.find(key);
    if (it == s.graphs.end() && s.graphs.size() >= 64) {  // a caller with ever-changing batch shapes: stop caching
        run_cnn(m, s, input, frames);
        run_tail_post(m, ln, s, first, frames);
        CK(cudaGetLastError());
        return;
    }
    if (it == s.graphs.end()) {
        // first sighting: run eagerly (kernels set their function attributes on first use); capture on the second
        if (s.graph_seen[key]++ == 0) {
            run_cnn(m, s, input, frames);
            run_tail_post(m, ln, s, first, frames);
            CK(cudaGetLastError());
            return;
        }
        if (!stem_inside) run_cnn(m, s, input, frames, 0, 1);
        const uint64_t launches_before = m.launches.load();
        cudaGraph_t g = nullptr;
        CK(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
        try {
            run_cnn(m, s, input, frames, stem_inside ? 0 : 1);
            run_tail_post(m, ln, s, first, frames);
        } catch (...) {
            cudaStreamEndCapture(s.stream, &g);
            if (g) cudaGraphDestroy(g);
            throw;
        }
        CK(cudaStreamEndCapture(s.stream, &g));
        CK(cudaGetLastError());
        cudaGraphExec_t ge = nullptr;
        cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) throw CudaError(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        s.graphs[key] = ge;
        const uint64_t recorded = m.launches.load() - launches_before;  // (other lanes may add real launches meanwhile:
        s.graph_nodes[key] = (int)recorded;                             //  the count is only used for the launch statistic)
        m.launches -= recorded;  // capture recorded the kernels, it did not launch them
        it = s.graphs.find(key);
    } else if (!stem_inside) {
        run_cnn(m, s, input, frames, 0, 1);
    }
    CK(cudaGraphLaunch(it->second, s.stream));
    m.launches += (uint64_t)s.graph_nodes[key];
}

struct FrameSrc {
    const uint8_t* p;            // RGB8 pixels in host memory, or NULL when the frame arrives as JPEG
    uint32_t w, h;
    const JpegCoefs* jc = nullptr;      // JPEG, Huffman-decoded on the host
    const JpegBitstream* jb = nullptr;  // JPEG, to be Huffman-decoded on the GPU
};

// staging buffer for frames that need a resize; a failed re-allocation leaves the slot empty, not broken
static void grow_input(Slot& s, size_t need) {
    if (need <= s.d_in_cap) return;
    CK(cudaStreamSynchronize(s.stream));
    cudaFree(s.d_in);
    s.d_in = nullptr;
    s.d_in_cap = 0;
    CK(cudaMalloc(&s.d_in, need));
    s.d_in_cap = need;
}

// ---- N2: JPEG frames. Layout of a run in the stage's staging buffer (host pinned = device image):
//   [JpegPlan x frames][block offsets of every frame][entries of every frame], each part 16-byte aligned
static size_t jpeg_stage_bytes(const JpegCoefs& jc) {
    return sizeof(JpegPlan) + (jc.block_off.size() + jc.entries.size()) * sizeof(uint32_t) + 64;
}

static void grow_jpeg(Slot& s, size_t need, size_t planes_need) {
    if (need > s.jpeg_cap) {
        CK(cudaStreamSynchronize(s.stream));
        cudaFreeHost(s.h_jpeg); cudaFree(s.d_jpeg);
        s.h_jpeg = s.d_jpeg = nullptr;
        s.jpeg_cap = 0;
        const size_t cap = need + need / 4;
        CK(cudaMallocHost(&s.h_jpeg, cap));
        CK(cudaMalloc(&s.d_jpeg, cap));
        s.jpeg_cap = cap;
    }
    if (planes_need > s.planes_cap) {
        CK(cudaStreamSynchronize(s.stream));
        cudaFree(s.d_planes);
        s.d_planes = nullptr;
        s.planes_cap = 0;
        CK(cudaMalloc(&s.d_planes, planes_need + planes_need / 4));
        s.planes_cap = planes_need + planes_need / 4;
    }
}

// `cnt` same-size JPEG frames -> RGB8 at dst (frame k at dst + k * w * h * 3): one H2D of plans + nonzero coefficients, then
// dequantise + IDCT + upsample + colour on the device
static void decode_jpeg_run(uf_model& m, Slot& s, const FrameSrc* fr, uint32_t cnt, uint8_t* dst, size_t& jpeg_used, size_t& planes_used) {
    auto a16 = [](size_t v) { return (v + 15) / 16 * 16; };
    size_t n_offs = 0, n_ent = 0;
    for (uint32_t k = 0; k < cnt; ++k) { n_offs += fr[k].jc->block_off.size(); n_ent += fr[k].jc->entries.size(); }
    const size_t base = a16(jpeg_used);
    const size_t o_plans = base, o_offs = a16(o_plans + cnt * sizeof(JpegPlan)), o_ent = a16(o_offs + n_offs * 4), end = o_ent + n_ent * 4;
    if (end > s.jpeg_cap) throw CudaError("internal: JPEG staging buffer undersized");
    JpegPlan* plans = reinterpret_cast<JpegPlan*>(s.h_jpeg + o_plans);
    uint32_t* offs = reinterpret_cast<uint32_t*>(s.h_jpeg + o_offs);
    uint32_t* ent = reinterpret_cast<uint32_t*>(s.h_jpeg + o_ent);
    size_t io = 0, ie = 0;
    uint32_t max_blocks = 0;
    const size_t fb = (size_t)fr[0].w * fr[0].h * 3;
    for (uint32_t k = 0; k < cnt; ++k) {
        const JpegCoefs& jc = *fr[k].jc;
        JpegPlan p = jc.plan;
        p.offs_base = (uint32_t)io;
        p.entries_base = (uint32_t)ie;
        const size_t ro = (size_t)k * fb, po = planes_used;
        p.rgb_off_lo = (uint32_t)ro; p.rgb_off_hi = (uint32_t)(ro >> 32);
        p.planes_off_lo = (uint32_t)po; p.planes_off_hi = (uint32_t)(po >> 32);
        planes_used += p.plane_bytes;
        plans[k] = p;
        memcpy(offs + io, jc.block_off.data(), jc.block_off.size() * 4);
        memcpy(ent + ie, jc.entries.data(), jc.entries.size() * 4);
        io += jc.block_off.size();
        ie += jc.entries.size();
        max_blocks = std::max(max_blocks, p.nblocks);
    }
    if (planes_used > s.planes_cap) throw CudaError("internal: JPEG plane buffer undersized");
    CK(cudaMemcpyAsync(s.d_jpeg + base, s.h_jpeg + base, end - base, cudaMemcpyHostToDevice, s.stream));
    jpeg_used = end;
    JpegBatchDev b{reinterpret_cast<const JpegPlan*>(s.d_jpeg + o_plans), reinterpret_cast<const uint32_t*>(s.d_jpeg + o_offs),
                   reinterpret_cast<const uint32_t*>(s.d_jpeg + o_ent), s.d_planes, dst};
    // SURVEY.md 8(d)-style accounting: coefficients in, planes written + read, RGB out
    const uint64_t bytes = (uint64_t)(end - base) + 2ull * (planes_used) + (uint64_t)cnt * fb;
    ProfScope ps(m, s, "jpeg_idct_upsample_rgb", bytes, (uint64_t)(end - base) + (uint64_t)cnt * fb, 0, 2);
    launch_jpeg_decode(b, (int)cnt, max_blocks, fr[0].w, fr[0].h, s.stream);
}

// ---- N2, Huffman decoding on the device too. Scratch per frame: 3 state arrays + 2 count arrays per subsequence, the dense blocks.
static size_t huff_ent_cap(const JpegBitstream& jb) { return (size_t)jb.huff.data_bits / 2 + 16; }
static size_t huff_scratch_bytes(const JpegBitstream& jb) {  // (+ slack for the alignment of a run's sections)
    const size_t nsub_max = (jb.huff.data_bits + JH_MIN_SUBSEQ_BITS - 1) / JH_MIN_SUBSEQ_BITS;  // (whatever length the run picks)
    return nsub_max * (3 * 8 + 4) + ((size_t)jb.plan.nblocks + 1) * (4 + 2) + huff_ent_cap(jb) * 4 + 1024;
}
// (staging: a frame may bring its own table set — frames of one camera share one, but the bound cannot assume it)
static size_t huff_stage_bytes(const JpegBitstream& jb) {
    return sizeof(JpegPlan) + sizeof(JpegHuffFrame) + sizeof(JpegHuffTabSet) + jb.data.size() + 96;
}
static void grow_huff(Slot& s, size_t need) {
    if (need <= s.huff_cap) return;
    CK(cudaStreamSynchronize(s.stream));
    cudaFree(s.d_huff);
    s.d_huff = nullptr;
    s.huff_cap = 0;
    CK(cudaMalloc(&s.d_huff, need + need / 4));
    s.huff_cap = need + need / 4;
}


struct HuffRun {
    const JpegPlan* d_plans = nullptr;   // device: the run's plans (offs_base = first block of the frame in d_coefs)
    const uint32_t* d_offs = nullptr;    // device: per frame nblocks + 1 block offsets into its AC entries
    const uint32_t* d_entries = nullptr; // device: AC entries (natural index << 16 | value)
    const int16_t* d_dcv = nullptr;      // device: DC values
    const int* d_status = nullptr;       // device: per frame, 0 = decoded to exactly its blocks
    size_t n_blocks = 0;
    uint32_t max_blocks = 0;
    int rounds = 0;
};

// Entropy-decodes `cnt` prepared frames on the device, asynchronously on the slot's stream. rgb_frame_bytes: stride of the frames in the RGB destination of the later stages.
static HuffRun huffman_run_gpu(uf_model& m, Slot& s, const FrameSrc* fr, uint32_t cnt, size_t rgb_frame_bytes, size_t& jpeg_used,
                               size_t& planes_used, size_t& huff_used) {
    auto a16 = [](size_t v) { return (v + 15) / 16 * 16; };
    HuffRun R;
    size_t n_bytes = 0, n_sub = 0, n_ent = 0, n_bits = 0;
    uint32_t max_nsub = 0;
    for (uint32_t k = 0; k < cnt; ++k) n_bits += fr[k].jb->huff.data_bits;
    // Bits per thread: as long as the run still gives the GPU ~150 k threads (a thread that starts from a guess needs a few
    // thousand bits to fall into step, so the work is about 1 + that / sub_bits passes over the data: long is cheap, but a
    // small run needs short subsequences to have threads at all).
    uint32_t sub_bits = JH_MAX_SUBSEQ_BITS;
    while (sub_bits > 256 && n_bits / sub_bits < m.jh_threads) sub_bits /= 2;  // (measured: 256 for 16 frames of 75 KB, 512-1024 for 128)
    if (m.jh_bits) sub_bits = m.jh_bits;
    auto nsub_of = [&](const JpegBitstream& jb) { return (jb.huff.data_bits + sub_bits - 1) / sub_bits; };
    for (uint32_t k = 0; k < cnt; ++k) {
        n_bytes += a16(fr[k].jb->data.size());
        n_sub += nsub_of(*fr[k].jb);
        n_ent += huff_ent_cap(*fr[k].jb);
        R.n_blocks += fr[k].jb->plan.nblocks;
        max_nsub = std::max(max_nsub, nsub_of(*fr[k].jb));
        R.max_blocks = std::max(R.max_blocks, fr[k].jb->plan.nblocks);
    }
    // the run's distinct table sets (normally one: every frame of a camera carries the same DHT, or none)
    std::vector<uint32_t> set_of(cnt);
    std::vector<const JpegHuffKey*> keys;
    for (uint32_t k = 0; k < cnt; ++k) {
        uint32_t u = 0;
        while (u < keys.size() && !(*keys[u] == fr[k].jb->key)) ++u;
        if (u == keys.size()) keys.push_back(&fr[k].jb->key);
        set_of[k] = u;
    }
    // staging (pinned = device image): [JpegPlan x cnt][JpegHuffFrame x cnt][JpegHuffTabSet x distinct][bytes]
    const size_t base = a16(jpeg_used);
    const size_t o_plans = base, o_hf = a16(o_plans + cnt * sizeof(JpegPlan)), o_ts = a16(o_hf + cnt * sizeof(JpegHuffFrame)),
                 o_bytes = a16(o_ts + keys.size() * sizeof(JpegHuffTabSet)), end = o_bytes + n_bytes;
    if (end > s.jpeg_cap) throw CudaError("internal: JPEG staging buffer undersized");
    // device scratch: [state A][state B][start_used][counts][status][block offsets][DC][AC entries]
    const size_t hb = (huff_used + 255) / 256 * 256;
    const size_t h_a = hb, h_b = h_a + n_sub * 8, h_su = h_b + n_sub * 8, h_nb = h_su + n_sub * 8, h_fl = a16(h_nb + n_sub * 4),
                 h_of = a16(h_fl + ((size_t)cnt + 1) * 4), h_dc = a16(h_of + (R.n_blocks + cnt) * 4), h_en = a16(h_dc + (R.n_blocks + cnt) * 2),
                 h_end = h_en + n_ent * 4;
    if (h_end > s.huff_cap) throw CudaError("internal: GPU Huffman scratch undersized");
    if (n_ent >= (1ull << 32) || R.n_blocks + cnt >= (1ull << 32)) throw CudaError("internal: JPEG run too large for 32-bit indices");
    huff_used = h_end;
    JpegPlan* plans = reinterpret_cast<JpegPlan*>(s.h_jpeg + o_plans);
    JpegHuffFrame* hfs = reinterpret_cast<JpegHuffFrame*>(s.h_jpeg + o_hf);
    for (size_t u = 0; u < keys.size(); ++u) jpeg_build_tabset(*keys[u], reinterpret_cast<JpegHuffTabSet*>(s.h_jpeg + o_ts)[u]);
    size_t ib = 0, isub = 0, iblk = 0, ient = 0;
    std::vector<size_t> byte_off(cnt);
    for (uint32_t k = 0; k < cnt; ++k) {
        const JpegBitstream& jb = *fr[k].jb;
        JpegPlan p = jb.plan;
        p.offs_base = (uint32_t)(iblk + k);  // (nblocks + 1 offsets per frame)
        p.entries_base = (uint32_t)ient;
        const size_t ro = (size_t)k * rgb_frame_bytes, po = planes_used;
        p.rgb_off_lo = (uint32_t)ro; p.rgb_off_hi = (uint32_t)(ro >> 32);
        p.planes_off_lo = (uint32_t)po; p.planes_off_hi = (uint32_t)(po >> 32);
        planes_used += p.plane_bytes;
        plans[k] = p;
        JpegHuffFrame h = jb.huff;
        h.sub_bits = sub_bits; h.nsub = nsub_of(jb);
        h.tabset = set_of[k]; h.data_off = (uint32_t)ib; h.sub_base = (uint32_t)isub;
        h.offs_base = (uint32_t)(iblk + k); h.ent_base = (uint32_t)ient; h.ent_cap = (uint32_t)huff_ent_cap(jb);
        hfs[k] = h;
        byte_off[k] = o_bytes + ib;
        ib += a16(jb.data.size());
        isub += h.nsub;
        iblk += jb.plan.nblocks;
        ient += huff_ent_cap(jb);
    }
    m.pool().parallel_for(cnt, [&](uint32_t k) { memcpy(s.h_jpeg + byte_off[k], fr[k].jb->data.data(), fr[k].jb->data.size()); });
    CK(cudaMemcpyAsync(s.d_jpeg + base, s.h_jpeg + base, end - base, cudaMemcpyHostToDevice, s.stream));
    jpeg_used = end;
    int* d_flags = reinterpret_cast<int*>(s.d_huff + h_fl);  // [1 + k] = status of frame k
    JpegHuffBatch hbt{reinterpret_cast<const JpegHuffFrame*>(s.d_jpeg + o_hf), reinterpret_cast<const JpegHuffTabSet*>(s.d_jpeg + o_ts),
                      s.d_jpeg + o_bytes,
                      reinterpret_cast<unsigned long long*>(s.d_huff + h_su), reinterpret_cast<uint32_t*>(s.d_huff + h_nb),
                      reinterpret_cast<uint32_t*>(s.d_huff + h_of), reinterpret_cast<uint32_t*>(s.d_huff + h_en),
                      reinterpret_cast<int16_t*>(s.d_huff + h_dc), d_flags + 1};
    unsigned long long* st[2] = {reinterpret_cast<unsigned long long*>(s.d_huff + h_a), reinterpret_cast<unsigned long long*>(s.d_huff + h_b)};
    R.d_plans = reinterpret_cast<const JpegPlan*>(s.d_jpeg + o_plans);
    R.d_offs = reinterpret_cast<const uint32_t*>(s.d_huff + h_of);
    R.d_entries = reinterpret_cast<const uint32_t*>(s.d_huff + h_en);
    R.d_dcv = reinterpret_cast<const int16_t*>(s.d_huff + h_dc);
    R.d_status = d_flags + 1;
    R.rounds = jhuff_rounds(max_nsub);  // a fixed count, no read-back: the write pass verifies the result
    ProfScope ps(m, s, "jpeg_huffman_gpu", 2 * (uint64_t)n_bytes + R.n_blocks * 8, (uint64_t)n_bytes + R.n_blocks * 6, 0, R.rounds + 2);
    CK(cudaMemsetAsync(s.d_huff + h_fl, 0, ((size_t)cnt + 1) * 4, s.stream));  // status
    int cur = 1;
    for (int r = 0; r < R.rounds; ++r) {
        launch_jhuff_sync(hbt, (int)cnt, max_nsub, r == 0, st[cur], st[cur ^ 1], s.stream);
        cur ^= 1;  // st[cur] holds the latest end states
    }
    launch_jhuff_finish(hbt, (int)cnt, max_nsub, st[cur], s.stream);
    return R;
}

// `cnt` same-size JPEG frames (frames [first_in_stage, +cnt) of the stage) -> RGB8 at dst, entropy decoding included. A
// frame that does not decode to exactly its blocks is flagged in s.h_jstatus and redone by the caller on the host decoder.
static void decode_jpeg_run_gpu(uf_model& m, Slot& s, const FrameSrc* fr, uint32_t first_in_stage, uint32_t cnt, uint8_t* dst,
                                size_t& jpeg_used, size_t& planes_used, size_t& huff_used) {
    const size_t fb = (size_t)fr[0].w * fr[0].h * 3;
    const HuffRun R = huffman_run_gpu(m, s, fr, cnt, fb, jpeg_used, planes_used, huff_used);
    if (planes_used > s.planes_cap) throw CudaError("internal: JPEG plane buffer undersized");
    s.jstatus_n = std::max(s.jstatus_n, first_in_stage + cnt);
    CK(cudaMemcpyAsync(s.h_jstatus + first_in_stage, R.d_status, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s.stream));
    JpegBatchDev b{R.d_plans, R.d_offs, R.d_entries, s.d_planes, dst, R.d_dcv, R.d_status};
    ProfScope ps(m, s, "jpeg_idct_upsample_rgb", (uint64_t)R.n_blocks * 16 + 2ull * planes_used + (uint64_t)cnt * fb,
                 (uint64_t)R.n_blocks * 16 + (uint64_t)cnt * fb, 0, 2);
    launch_jpeg_decode(b, (int)cnt, R.max_blocks, fr[0].w, fr[0].h, s.stream);
}

// One chunk of host frames on slot s.
static void run_chunk_host(uf_model& m, Lane& ln, Slot& s, const FrameSrc* fr, uint32_t first, uint32_t n) {
    const int W = m.plan.net_w, H = m.plan.net_h;
    const size_t out_frame = (size_t)W * H * 3;
    // total staging bytes for frames that need a resize
    size_t need = 0;
    for (uint32_t i = 0; i < n; ++i)
        if ((int)fr[i].w != W || (int)fr[i].h != H) need += (size_t)fr[i].w * fr[i].h * 3 + 16;
    grow_input(s, need);
    // every frame of the stage at exactly twice the network size (the benchmark's 640x480 webcam frames into RFB-320): they
    // are only copied; the first kernel of the chain resamples, normalises and convolves them
    bool all_double = m.prestem_ok && n > 0 && (int)fr[0].w == 2 * W && (int)fr[0].h == 2 * H;
    for (uint32_t k = 1; k < n && all_double; ++k)
        all_double = fr[k].w == fr[0].w && fr[k].h == fr[0].h && (fr[k].jc != nullptr) == (fr[0].jc != nullptr);
    if (all_double) {
        TapsRef t = get_taps(m, fr[0].w, fr[0].h);
        all_double = prestem_supported(s.d_in, (long long)fr[0].w * fr[0].h * 3, fr[0].w, fr[0].h, W, H, t->dev);
    }
    // JPEG frames of the stage: size the staging buffers once (growing them mid-stage would strand copies in flight)
    size_t jpeg_need = 0, planes_need = 0, jpeg_used = 0, planes_used = 0, huff_need = 0, huff_used = 0;
    s.jstatus_n = 0;
    for (uint32_t k = 0; k < n; ++k) {
        s.h_jstatus[k] = 0;
        if (fr[k].jc) {
            jpeg_need += jpeg_stage_bytes(*fr[k].jc);
            planes_need += fr[k].jc->plan.plane_bytes;
        } else if (fr[k].jb) {
            jpeg_need += huff_stage_bytes(*fr[k].jb);
            planes_need += fr[k].jb->plan.plane_bytes;
            huff_need += huff_scratch_bytes(*fr[k].jb);
        }
    }
    if (jpeg_need) grow_jpeg(s, jpeg_need, planes_need);
    grow_huff(s, huff_need);
    size_t off = 0;
    uint32_t i = 0;
    while (i < n) {
        // run of frames with identical size (and, for the copy, contiguous host addresses)
        uint32_t j = i + 1;
        const size_t fb = (size_t)fr[i].w * fr[i].h * 3;
        while (j < n && fr[j].w == fr[i].w && fr[j].h == fr[i].h && (fr[j].jc != nullptr) == (fr[i].jc != nullptr) &&
               (fr[j].jb != nullptr) == (fr[i].jb != nullptr)) ++j;
        const bool ident = (int)fr[i].w == W && (int)fr[i].h == H;
        if (!ident) off = (off + 15) / 16 * 16;  // keep every run 16-byte aligned for the fast resize path
        uint8_t* dst = ident ? s.d_resized + (size_t)i * out_frame : s.d_in + off;
        uint32_t a = i;
        if (fr[i].jc) {  // JPEG: coefficients up, pixels made on the device (the runs of a stage share the staging buffer)
            decode_jpeg_run(m, s, fr + i, j - i, dst, jpeg_used, planes_used);
            a = j;
        } else if (fr[i].jb) {  // JPEG: the entropy-coded bytes up, Huffman decoding on the device too
            decode_jpeg_run_gpu(m, s, fr + i, i, j - i, dst, jpeg_used, planes_used, huff_used);
            a = j;
        }
        while (a < j) {  // merge host-contiguous frames into one cudaMemcpyAsync
            uint32_t b = a + 1;
            while (b < j && fr[b].p == fr[b - 1].p + fb) ++b;
            CK(cudaMemcpyAsync(dst + (size_t)(a - i) * fb, fr[a].p, (size_t)(b - a) * fb, cudaMemcpyHostToDevice, s.stream));
            a = b;
        }
        if (!ident && !all_double) {
            TapsRef t = get_taps(m, fr[i].w, fr[i].h);
            ProfScope ps(m, s, "resize_triangle", (uint64_t)(j - i) * (fb + out_frame), (uint64_t)(j - i) * (fb + out_frame), 0);
            launch_resize(s.d_in + off, (long long)fb, fr[i].w, fr[i].h, s.d_resized + (size_t)i * out_frame,
                          (long long)out_frame, W, H, (int)(j - i), t->dev, m.cfg.resize_round_intermediate, s.stream);
            off += (size_t)(j - i) * fb;
        }
        i = j;
    }
    U8View input{s.d_resized, (long long)out_frame, H, W};
    if (all_double) input = U8View{s.d_in, (long long)fr[0].w * fr[0].h * 3, (int)fr[0].h, (int)fr[0].w};
    run_body(m, ln, s, input, first, (int)n);
    s.pending = true; s.first = first; s.n = n;
    // (optional, see uf_model::sleep_wait_from)
    s.sleep_wait = m.sleep_wait_from > 0 && n >= m.sleep_wait_from;
    if (s.sleep_wait) CK(cudaEventRecord(s.done_ev, s.stream));
}

// One chunk of device-resident frames (identical size, contiguous) on slot s.
static void run_chunk_device(uf_model& m, Lane& ln, Slot& s, const uint8_t* d_rgb, uint32_t w, uint32_t h, uint32_t first, uint32_t n) {
    const int W = m.plan.net_w, H = m.plan.net_h;
    const size_t out_frame = (size_t)W * H * 3, fb = (size_t)w * h * 3;
    U8View input{s.d_resized, (long long)out_frame, H, W};
    if ((int)w == W && (int)h == H) {
        input.p = d_rgb;  // identity resize (sample.rs early return): the stem reads the caller's frames
        input.frame_stride = (long long)fb;
    } else {
        TapsRef t = get_taps(m, w, h);
        if (m.prestem_ok && prestem_supported(d_rgb, (long long)fb, w, h, W, H, t->dev)) {
            input = U8View{d_rgb, (long long)fb, (int)h, (int)w};  // resampled inside the first kernel of the chain
        } else {
            ProfScope ps(m, s, "resize_triangle", (uint64_t)n * (fb + out_frame), (uint64_t)n * (fb + out_frame), 0);
            launch_resize(d_rgb, (long long)fb, w, h, s.d_resized, (long long)out_frame, W, H, (int)n, t->dev,
                          m.cfg.resize_round_intermediate, s.stream);
        }
    }
    run_body(m, ln, s, input, first, (int)n);
    s.pending = true; s.first = first; s.n = n;
}

// `taper`: shrink the last pipeline stages (host input). The copy of stage i+1 hides behind the kernels of stage
// i, but nothing hides the kernels of the LAST stage, so the tail of the batch is cut into smaller stages.
// after a failed call: nothing of it may survive into the next one (a stale `pending` slot would make the next
// call's harvest write the failed call's results into the new caller's, possibly much smaller, arrays)
static void abandon_lane(uf_model& m, Lane& ln) {
    for (auto& s : ln.slots) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s.stream, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) {
            cudaGraph_t g = nullptr;
            cudaStreamEndCapture(s.stream, &g);
            if (g) cudaGraphDestroy(g);
        }
        cudaStreamSynchronize(s.stream);
        s.pending = false;
        s.sleep_wait = false;
        s.ev_used = 0;
        s.n = 0;
        s.jstatus_n = 0;
        s.redo.clear();
    }
    ln.last_n = 0;
    cudaGetLastError();
}

template <typename F>
static void run_pipeline(uf_model& m, Lane& ln, uint32_t n, uint32_t step, bool taper, uf_det* out, uint32_t cap,
                         uint32_t* n_out, F&& submit) {
    if (n > m.cfg.max_batch) throw ArgError(UF_ERR_CAPACITY, "batch of " + std::to_string(n) + " exceeds max_batch " + std::to_string(m.cfg.max_batch));
    CK(cudaSetDevice(m.cfg.device));
    const uint32_t nslots = m.profiling ? 1 : m.nslots;  // profiling: one stream, so event pairs time kernels alone
    uint32_t c = 0;
    ++ln.batch_id;
    try {
        for (uint32_t first = 0; first < n; ++c) {
            uint32_t cnt = std::min(step, n - first);
            if (taper && !m.profiling) {
                const uint32_t rem = n - first;
                if (rem <= step && rem > 16) cnt = std::max<uint32_t>(16, (rem / 2 + 7) / 8 * 8);
            }
            Slot& s = ln.slots[c % nslots];
            harvest(m, s, out, cap, n_out);
            if (m.fail_after_stages.load() >= 0 && (int)c >= m.fail_after_stages.load())
                throw CudaError("injected fault after " + std::to_string(c) + " pipeline stages (uf_debug_fail_after)");
            submit(s, first, cnt);
            s.batch_id = ln.batch_id;
            first += cnt;
        }
        for (uint32_t k = 0; k < nslots; ++k) harvest(m, ln.slots[(c + k) % nslots], out, cap, n_out);
    } catch (...) {
        abandon_lane(m, ln);
        throw;
    }
    ln.last_n = n;
}

// Picks a lane for a call: the first one that is free, else waits for one round-robin. Profiling and the
// parity hooks always use lane 0 (a single-threaded caller therefore always sees its own last batch there).
struct LaneLock {
    Lane* lane = nullptr;
    std::unique_lock<std::mutex> lk;
    LaneLock(uf_model& m, bool force0) {
        if (!force0 && !m.profiling) {
            for (auto& l : m.lanes) {
                std::unique_lock<std::mutex> t(l->mu, std::try_to_lock);
                if (t.owns_lock()) { lane = l.get(); lk = std::move(t); return; }
            }
            lane = m.lanes[m.lane_rr++ % m.lanes.size()].get();
        } else {
            lane = m.lanes[0].get();
        }
        lk = std::unique_lock<std::mutex>(lane->mu);
    }
};

static void* hook_scratch(uf_model& m, size_t bytes) {
    if (bytes > m.d_hook_cap) {
        if (m.d_hook) CK(cudaFree(m.d_hook));
        m.d_hook = nullptr;
        CK(cudaMalloc(&m.d_hook, bytes));
        m.d_hook_cap = bytes;
    }
    return m.d_hook;
}

static uf_model* load_model(const uf_config& cfg_in) {
    uf_config cfg = cfg_in;
    if (!cfg.onnx_path) throw ArgError(UF_ERR_INVALID_ARG, "onnx_path is NULL");
    if (cfg.net_w == 0 || cfg.net_h == 0 || cfg.net_w > 8192 || cfg.net_h > 8192) throw ArgError(UF_ERR_INVALID_ARG, "bad network input size");
    if (cfg.max_batch == 0) cfg.max_batch = 1;
    if (cfg.norm_preset > UF_NORM_127_128) throw ArgError(UF_ERR_INVALID_ARG, "unknown norm_preset");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        throw ArgError(UF_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
    }
    if (cfg.device < 0 || cfg.device >= ndev) throw ArgError(UF_ERR_INVALID_ARG, "device ordinal out of range");
    OnnxModel om = load_onnx_file(cfg.onnx_path);
    std::unique_ptr<uf_model> m(new uf_model());
    m->onnx_path = cfg.onnx_path;
    m->cfg = cfg;
    m->cfg.onnx_path = m->onnx_path.c_str();
    m->plan = lower_ultraface(om, (int)cfg.net_w, (int)cfg.net_h);
    m->K = m->plan.num_priors;
    CK(cudaSetDevice(cfg.device));
    // chunk: big enough to amortise launches, small enough that a chunk's activations
    // (~arena_frame_floats*4 B per frame) do not dwarf the 126 MB L2
    // chunk = frames per launch for device-resident input (big: fewer, fuller launches);
    // host_chunk = frames per pipeline stage for host input (small: the H2D copy overlaps the kernels)
    uint32_t chunk = cfg.chunk;
    if (chunk == 0) chunk = (uint64_t)cfg.net_w * cfg.net_h <= 320 * 240 ? 128 : 32;
    chunk = std::min(chunk, cfg.max_batch);
    m->chunk = chunk;
    m->host_chunk = std::max<uint32_t>(1, std::min<uint32_t>(chunk, (uint64_t)cfg.net_w * cfg.net_h <= 320 * 240 ? 64 : 16));
    if (cfg.host_chunk) m->host_chunk = std::max<uint32_t>(1, std::min<uint32_t>(chunk, cfg.host_chunk));
    m->jpeg_chunk = std::max(m->host_chunk, std::min<uint32_t>(chunk, 128));
    if (const char* e = getenv("UF_JPEG_CHUNK")) m->jpeg_chunk = std::max<uint32_t>(1, std::min<uint32_t>(chunk, (uint32_t)atoi(e)));  // tuning knob
    if (const char* e = getenv("UF_SLEEP_WAIT_FROM")) m->sleep_wait_from = (uint32_t)std::max(0, atoi(e));
    if (const char* e = getenv("UF_JH_THREADS")) m->jh_threads = (uint32_t)std::max(1, atoi(e));
    if (const char* e = getenv("UF_JH_BITS"); e && atoi(e) > 0) m->jh_bits = std::max(JH_MIN_SUBSEQ_BITS, std::min(JH_MAX_SUBSEQ_BITS, (uint32_t)atoi(e) / 32 * 32));
    uint32_t nslots = cfg.slots ? cfg.slots : 4;
    const uint32_t nchunks = (cfg.max_batch + m->host_chunk - 1) / m->host_chunk;
    m->nslots = std::max<uint32_t>(1, std::min(nslots, nchunks));
    pack_weights(*m);
    build_steps(*m);
    build_tc_weights(*m);
    build_dense3_weights(*m);
    build_param_weights(*m);
    label_steps(*m);
    m->prestem_ok = !m->steps.empty() && m->steps[0].impl == Impl::Stem && !(cfg.flags & UF_FLAG_NO_PRESTEM);
    if (m->prestem_ok) {
        const std::string& l = m->steps[0].label;
        m->prestem_label = "resize2x_norm_stem_u8" + l.substr(l.find('['));
    }
    build_lut(*m);
    CK(cudaMalloc(&m->d_priors, (size_t)m->K * 4 * sizeof(float)));
    CK(cudaMemcpy(m->d_priors, m->plan.priors.data(), (size_t)m->K * 4 * sizeof(float), cudaMemcpyHostToDevice));
    cudaError_t pe = (cudaError_t)post_configure();
    if (pe != cudaSuccess) throw CudaError(std::string("post_configure: ") + cudaGetErrorString(pe));
    uint32_t nlanes = cfg.lanes ? cfg.lanes : 2;
    nlanes = std::max<uint32_t>(1, std::min<uint32_t>(nlanes, 8));
    for (uint32_t i = 0; i < nlanes; ++i) {
        m->lanes.emplace_back(new Lane());
        alloc_lane(*m, *m->lanes.back());
    }
    return m.release();
}

}  // namespace uf

uf_model::~uf_model() {
    cudaSetDevice(cfg.device);
    for (auto& lane : lanes)
    for (auto& s : lane->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        for (auto& e : s.ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        for (auto& g : s.graphs) cudaGraphExecDestroy(g.second);
        cudaFree(s.d_in); cudaFree(s.d_resized); cudaFree(s.d_arena); cudaFree(s.d_dets); cudaFree(s.d_sel);
        cudaFree(s.d_det_idx); cudaFree(s.d_counts); cudaFree(s.d_sort); cudaFree(s.d_big_n); cudaFree(s.d_mask);
        cudaFreeHost(s.h_jpeg); cudaFree(s.d_jpeg); cudaFree(s.d_planes); cudaFree(s.d_huff); cudaFree(s.d_ovl); cudaFreeHost(s.h_jstatus);
        cudaFreeHost(s.h_counts); cudaFreeHost(s.h_dets);
        if (s.done_ev) cudaEventDestroy(s.done_ev);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    taps.clear();
    for (auto& t : tc_weights) { cudaFree(t.d_hi); cudaFree(t.d_lo); }
    for (auto& lane : lanes) { cudaFree(lane->d_scores); cudaFree(lane->d_boxes); }
    cudaFree(d_weights); cudaFree(d_lut); cudaFree(d_priors); cudaFree(d_hook); cudaFree(d_atlas);
}

// ---- C ABI -----------------------------------------------------------------------------------
template <typename F>
static int guarded(F&& f) {
    try {
        f();
        return UF_OK;
    } catch (const ArgError& e) { g_last_error = e.msg; return e.code;
    } catch (const IoError& e) { g_last_error = e.msg; return UF_ERR_IO;
    } catch (const UnsupportedError& e) { g_last_error = e.msg; return UF_ERR_UNSUPPORTED;
    } catch (const JpegError& e) { g_last_error = e.msg; return e.code;
    } catch (const CudaError& e) { g_last_error = e.msg; cudaGetLastError(); return UF_ERR_CUDA;
    } catch (const std::bad_alloc&) { g_last_error = "out of host memory"; return UF_ERR_INVALID_ARG;
    } catch (const std::exception& e) { g_last_error = e.what(); return UF_ERR_ONNX; }
}

#define REQUIRE(cond, msg) do { if (!(cond)) throw ArgError(UF_ERR_INVALID_ARG, msg); } while (0)

extern "C" {

int uf_model_load_ex(const uf_config* cfg, uf_model** out) {
    return guarded([&] {
        REQUIRE(cfg && out, "null argument");
        REQUIRE(cfg->struct_size == sizeof(uf_config), "uf_config.struct_size mismatch");
        *out = nullptr;
        *out = load_model(*cfg);
    });
}

int uf_model_load(const char* onnx_path, uint32_t net_w, uint32_t net_h, float max_iou, float min_confidence,
                  int32_t device, uint32_t max_batch, uf_model** out) {
    uf_config c;
    memset(&c, 0, sizeof(c));
    c.struct_size = sizeof(c);
    c.onnx_path = onnx_path;
    c.net_w = net_w; c.net_h = net_h;
    c.max_iou = max_iou; c.min_confidence = min_confidence;
    c.device = device; c.max_batch = max_batch;
    return uf_model_load_ex(&c, out);
}

void uf_model_free(uf_model* m) { delete m; }

int uf_model_info(const uf_model* m, uf_info* o) {
    return guarded([&] {
        REQUIRE(m && o, "null argument");
        memset(o, 0, sizeof(*o));
        o->net_w = m->plan.net_w; o->net_h = m->plan.net_h;
        o->num_priors = m->K;
        o->num_layers = (uint32_t)m->steps.size();
        uint32_t nt = 0;
        for (auto r : m->tensor_readable) nt += r;
        o->num_tensors = nt;
        o->max_batch = m->cfg.max_batch; o->chunk = m->chunk; o->slots = m->nslots;
        o->weight_bytes = m->weight_bytes; o->workspace_bytes = m->workspace_bytes;
        // preproc (640x480 u8 in + f32 NCHW out) + conv nodes + post (K*6*4)
        const uint64_t pre = 640ull * 480 * 3 + 4ull * 3 * m->plan.net_w * m->plan.net_h;
        o->algorithmic_bytes_per_frame = pre + m->plan.conv_bytes_per_frame + 24ull * m->K;
        o->macs_per_frame = m->plan.macs_per_frame;
    });
}

int uf_infer_batch(uf_model* m, const uint8_t* const* rgb, const uint32_t* w, const uint32_t* h, uint32_t n,
                   uf_det* out, uint32_t cap, uint32_t* n_out) {
    return guarded([&] {
        REQUIRE(m && (n == 0 || (rgb && w && h)) && n_out, "null argument");
        REQUIRE(cap == 0 || out, "out is NULL with cap > 0");
        std::vector<FrameSrc> fr(n);
        for (uint32_t i = 0; i < n; ++i) {
            REQUIRE(rgb[i] && w[i] > 0 && h[i] > 0 && w[i] <= 16384 && h[i] <= 16384, "bad frame " + std::to_string(i));
            fr[i] = FrameSrc{rgb[i], w[i], h[i]};
        }
        LaneLock ll(*m, false);
        Lane& ln = *ll.lane;
        // host frames: small pipeline stages so the H2D copy of stage i+1 hides behind the kernels of stage i
        run_pipeline(*m, ln, n, m->host_chunk, true, out, cap, n_out, [&](Slot& s, uint32_t first, uint32_t cnt) {
            run_chunk_host(*m, ln, s, fr.data() + first, first, cnt);
        });
    });
}

// Huffman decoding of n frames on the library's host worker threads (the only serial part of JPEG decoding)
static void entropy_decode_all(uf_model& m, const uint8_t* const* jpeg, const size_t* len, uint32_t n, std::vector<JpegCoefs>& out) {
    if (out.size() < n) out.resize(n);
    std::vector<JpegError> errs(n, JpegError{UF_OK, ""});
    m.pool().parallel_for(n, [&](uint32_t i) {
        try {
            jpeg_entropy_decode(jpeg[i], len[i], out[i]);
        } catch (const JpegError& e) {
            errs[i] = e;
        } catch (const std::exception& e) {
            errs[i] = JpegError{UF_ERR_INVALID_ARG, e.what()};
        }
    });
    for (uint32_t i = 0; i < n; ++i)
        if (errs[i].code != UF_OK) throw JpegError{errs[i].code, "frame " + std::to_string(i) + ": " + errs[i].msg};
}

int uf_infer_batch_jpeg(uf_model* m, const uint8_t* const* jpeg, const size_t* len, uint32_t n, uf_det* out, uint32_t cap,
                        uint32_t* n_out) {
    return guarded([&] {
        REQUIRE(m && (n == 0 || (jpeg && len)) && n_out, "null argument");
        REQUIRE(cap == 0 || out, "out is NULL with cap > 0");
        for (uint32_t i = 0; i < n; ++i) REQUIRE(jpeg[i] && len[i] >= 4, "bad frame " + std::to_string(i));
        if (n > m->cfg.max_batch) throw ArgError(UF_ERR_CAPACITY, "batch of " + std::to_string(n) + " exceeds max_batch " + std::to_string(m->cfg.max_batch));
        LaneLock ll(*m, false);
        Lane& ln = *ll.lane;
        std::vector<FrameSrc> fr(n);
        const bool gpu_huffman = !(m->cfg.flags & UF_FLAG_JPEG_HOST_HUFFMAN);
        if (gpu_huffman) {
            // host: headers, tables, byte unstuffing (a memchr pass); device: everything else
            std::vector<JpegBitstream>& bs = ln.jpeg_streams;
            if (bs.size() < n) bs.resize(n);
            std::vector<JpegError> errs(n, JpegError{UF_OK, ""});
            m->pool().parallel_for(n, [&](uint32_t i) {
                try { jpeg_prepare_bitstream(jpeg[i], len[i], bs[i]); }
                catch (const JpegError& e) { errs[i] = e; }
                catch (const std::exception& e) { errs[i] = JpegError{UF_ERR_INVALID_ARG, e.what()}; }
            });
            for (uint32_t i = 0; i < n; ++i)
                if (errs[i].code != UF_OK) throw JpegError{errs[i].code, "frame " + std::to_string(i) + ": " + errs[i].msg};
            std::vector<JpegCoefs>& coefs = ln.jpeg_coefs;
            if (coefs.size() < n) coefs.resize(n);
            for (uint32_t i = 0; i < n; ++i) {
                if (bs[i].gpu_ok) {
                    fr[i] = FrameSrc{nullptr, bs[i].plan.w, bs[i].plan.h, nullptr, &bs[i]};
                } else {  // restart intervals / markers inside the segment: the host decoder
                    jpeg_entropy_decode(jpeg[i], len[i], coefs[i]);
                    fr[i] = FrameSrc{nullptr, coefs[i].plan.w, coefs[i].plan.h, &coefs[i], nullptr};
                }
            }
        } else {
            std::vector<JpegCoefs>& coefs = ln.jpeg_coefs;
            entropy_decode_all(*m, jpeg, len, n, coefs);
            for (uint32_t i = 0; i < n; ++i) fr[i] = FrameSrc{nullptr, coefs[i].plan.w, coefs[i].plan.h, &coefs[i], nullptr};
        }
        // (compressed frames are small: the stage can be as large as one for device-resident input, which fills the GPU better)
        run_pipeline(*m, ln, n, gpu_huffman ? m->jpeg_chunk : m->host_chunk, true, out, cap, n_out, [&](Slot& s, uint32_t first, uint32_t cnt) {
            run_chunk_host(*m, ln, s, fr.data() + first, first, cnt);
        });
        // frames the device decoder handed back (truncated / damaged streams): host decoder, one by one
        std::vector<uint32_t> redo;
        for (auto& s : ln.slots) {
            redo.insert(redo.end(), s.redo.begin(), s.redo.end());
            s.redo.clear();
        }
        for (uint32_t g : redo) {
            JpegCoefs jc;
            jpeg_entropy_decode(jpeg[g], len[g], jc);
            FrameSrc one{nullptr, jc.plan.w, jc.plan.h, &jc, nullptr};
            run_pipeline(*m, ln, 1, m->host_chunk, true, out ? out + (size_t)g * cap : nullptr, cap, n_out + g,
                         [&](Slot& s, uint32_t first, uint32_t cnt) { run_chunk_host(*m, ln, s, &one, first, cnt); });
        }
        m->jpeg_redone += redo.size();
        if (!redo.empty()) ln.last_n = 0;  // the raw-output hooks would mix two calls
    });
}

int uf_jpeg_decode_rgb(uf_model* m, const uint8_t* jpeg, size_t len, uint8_t* out_rgb, size_t cap_bytes, uint32_t* w, uint32_t* h) {
    return guarded([&] {
        REQUIRE(m && jpeg && w && h, "null argument");
        JpegBitstream jb;
        const bool gpu_huffman = !(m->cfg.flags & UF_FLAG_JPEG_HOST_HUFFMAN);
        if (gpu_huffman) {
            jpeg_prepare_bitstream(jpeg, len, jb);
            *w = jb.plan.w;
            *h = jb.plan.h;
            const size_t fb = (size_t)jb.plan.w * jb.plan.h * 3;
            if (!out_rgb || cap_bytes < fb) throw ArgError(UF_ERR_CAPACITY, "output buffer smaller than w * h * 3");
            if (jb.gpu_ok) {
                LaneLock ll(*m, true);
                Slot& s = ll.lane->slots[0];
                CK(cudaSetDevice(m->cfg.device));
                grow_input(s, fb);
                grow_jpeg(s, huff_stage_bytes(jb), jb.plan.plane_bytes);
                grow_huff(s, huff_scratch_bytes(jb));
                FrameSrc fr{nullptr, jb.plan.w, jb.plan.h, nullptr, &jb};
                size_t ju = 0, pu = 0, hu = 0;
                s.h_jstatus[0] = 0;
                decode_jpeg_run_gpu(*m, s, &fr, 0, 1, s.d_in, ju, pu, hu);
                CK(cudaMemcpyAsync(out_rgb, s.d_in, fb, cudaMemcpyDeviceToHost, s.stream));
                CK(cudaStreamSynchronize(s.stream));
                CK(cudaGetLastError());
                s.jstatus_n = 0;
                if (s.h_jstatus[0] == 0) return;
                m->jpeg_redone++;  // damaged / truncated: the host decoder below
            }
        }
        JpegCoefs jc;
        jpeg_entropy_decode(jpeg, len, jc);
        *w = jc.plan.w;
        *h = jc.plan.h;
        const size_t fb = (size_t)jc.plan.w * jc.plan.h * 3;
        if (!out_rgb || cap_bytes < fb) throw ArgError(UF_ERR_CAPACITY, "output buffer smaller than w * h * 3");
        LaneLock ll(*m, true);
        Slot& s = ll.lane->slots[0];
        CK(cudaSetDevice(m->cfg.device));
        grow_input(s, fb);
        grow_jpeg(s, jpeg_stage_bytes(jc), jc.plan.plane_bytes);
        FrameSrc fr{nullptr, jc.plan.w, jc.plan.h, &jc};
        size_t ju = 0, pu = 0;
        decode_jpeg_run(*m, s, &fr, 1, s.d_in, ju, pu);
        CK(cudaMemcpyAsync(out_rgb, s.d_in, fb, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaGetLastError());
    });
}

int uf_jpeg_coefficients_gpu(uf_model* m, const uint8_t* jpeg, size_t len, int16_t* coefs, size_t cap_blocks, int32_t* on_device) {
    return guarded([&] {
        REQUIRE(m && jpeg && coefs && on_device, "null argument");
        JpegBitstream jb;
        jpeg_prepare_bitstream(jpeg, len, jb);
        if (cap_blocks < jb.plan.nblocks) throw ArgError(UF_ERR_CAPACITY, "coefficient buffer smaller than nblocks");
        *on_device = 0;
        if (jb.gpu_ok) {
            LaneLock ll(*m, true);
            Slot& s = ll.lane->slots[0];
            CK(cudaSetDevice(m->cfg.device));
            grow_jpeg(s, huff_stage_bytes(jb), 0);
            grow_huff(s, huff_scratch_bytes(jb));
            FrameSrc fr{nullptr, jb.plan.w, jb.plan.h, nullptr, &jb};
            size_t ju = 0, pu = 0, hu = 0;
            const HuffRun R = huffman_run_gpu(*m, s, &fr, 1, 0, ju, pu, hu);
            const uint32_t nb = jb.plan.nblocks;
            std::vector<uint32_t> offs((size_t)nb + 1);
            std::vector<int16_t> dcv(nb);
            CK(cudaMemcpyAsync(offs.data(), R.d_offs, offs.size() * 4, cudaMemcpyDeviceToHost, s.stream));
            CK(cudaMemcpyAsync(dcv.data(), R.d_dcv, dcv.size() * 2, cudaMemcpyDeviceToHost, s.stream));
            CK(cudaMemcpyAsync(s.h_jstatus, R.d_status, 4, cudaMemcpyDeviceToHost, s.stream));
            CK(cudaStreamSynchronize(s.stream));
            CK(cudaGetLastError());
            if (s.h_jstatus[0] == 0) {
                if (offs[nb] > huff_ent_cap(jb)) throw CudaError("internal: device Huffman entry count out of range");
                std::vector<uint32_t> ent(offs[nb]);
                CK(cudaMemcpy(ent.data(), R.d_entries, ent.size() * 4, cudaMemcpyDeviceToHost));
                memset(coefs, 0, (size_t)nb * 128);
                for (uint32_t bk = 0; bk < nb; ++bk) {
                    coefs[(size_t)bk * 64] = dcv[bk];
                    for (uint32_t e = offs[bk]; e < offs[bk + 1] && e < ent.size(); ++e)
                        coefs[(size_t)bk * 64 + ((ent[e] >> 16) & 63)] = (int16_t)(ent[e] & 0xffffu);
                }
                *on_device = 1 + R.rounds;
                return;
            }
        }
        uf_jpeg_info info;
        if (uf_jpeg_coefficients(jpeg, len, &info, coefs, cap_blocks) != UF_OK) throw ArgError(UF_ERR_INVALID_ARG, g_last_error);
    });
}

// ---- N3: rectangles as the reference computes them (inferer.rs:66-75) and imageproc draws them ----
static int32_t sat_i32(float v) {  // Rust `as i32`: saturating, NaN -> 0
    if (!(v == v)) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}
static uint32_t sat_u32(float v) {  // Rust `as u32`
    if (!(v == v) || v <= 0.0f) return 0;
    if (v >= 4294967296.0f) return UINT32_MAX;
    return (uint32_t)v;
}

static std::string confidence_text(float confidence) {
    volatile float pct = confidence * 100.0f;  // f32 product, as the reference's `confidence * 100.0`
    char buf[64];
    snprintf(buf, sizeof(buf), "%.2f%%", (double)pct);  // the exact value of the f32, two decimals (Rust: "{:.2}%")
    return buf;
}

static std::vector<int4> reference_rects(const uf_det* dets, uint32_t n, float width, float height, std::vector<uint32_t>* which = nullptr) {
    std::vector<int4> r;
    for (uint32_t i = 0; i < n; ++i) {
        // inferer.rs:69-75, f32 arithmetic, one rounding per operation
        volatile float x_tl = dets[i].x0 * width, y_tl = dets[i].y0 * height;
        volatile float x_br = dets[i].x1 * width, y_br = dets[i].y1 * height;
        volatile float rect_width = x_br - x_tl, rect_height = y_br - y_tl;
        const uint32_t rw = sat_u32(rect_width), rh = sat_u32(rect_height);
        if (rw == 0 || rh == 0) continue;  // Rect::of_size asserts on these: the reference would panic, we skip the box
        const int64_t left = sat_i32(x_tl), top = sat_i32(y_tl);
        const int64_t right = left + (int64_t)rw - 1, bottom = top + (int64_t)rh - 1;  // imageproc Rect::right / bottom
        auto cl = [](int64_t v) { return (int)std::max<int64_t>(-(1 << 30), std::min<int64_t>(1 << 30, v)); };  // clipping happens per pixel anyway
        r.push_back(make_int4(cl(left), cl(top), cl(right), cl(bottom)));
        if (which) which->push_back(i);
    }
    return r;
}

// What the overlay of one frame consists of: rectangles, and — when the model holds a glyph atlas — the placed glyphs of
// every detection's text (glyphs of detection k: [gstart[k], gstart[k + 1])). text = false: rectangles only, order-free.
struct OverlayLists {
    std::vector<int4> rects;
    std::vector<uint32_t> gstart;
    std::vector<OverlayGlyph> glyphs;
    bool text = false;
};

// (m.atlas_mu held by the caller)
static void build_overlay_lists(uf_model& m, uint32_t w, uint32_t h, const uf_det* dets, uint32_t n_dets, float scale_w, float scale_h,
                                OverlayLists& L) {
    std::vector<uint32_t> which;
    std::vector<int4> rects = reference_rects(dets, n_dets, scale_w, scale_h, &which);
    L.rects.clear(); L.gstart.clear(); L.glyphs.clear();
    L.text = !m.atlas_chars.empty();
    if (L.text) {
        // rectangles and text in the reference's order: one list, walked by one CTA (kernels_jpeg_enc.cu)
        const uint32_t nc = (uint32_t)m.atlas_chars.size();
        for (size_t k = 0; k < rects.size(); ++k) {
            L.gstart.push_back((uint32_t)L.glyphs.size());
            const int ox = rects[k].x, oy = rects[k].y;  // x_tl as i32, y_tl as i32
            const std::string text = confidence_text(dets[which[k]].conf);
            for (uint32_t pos = 0; pos < text.size() && pos < m.atlas_max_len; ++pos) {
                const size_t ci = m.atlas_chars.find(text[pos]);
                if (ci == std::string::npos) continue;
                const uf_glyph& g = m.atlas_glyphs[(size_t)pos * nc + ci];
                if (g.w == 0 || g.h == 0) continue;
                const int64_t gx = (int64_t)ox + g.x0, gy = (int64_t)oy + g.y0;
                if (gx >= (int64_t)w || gy >= (int64_t)h || gx + (int64_t)g.w <= 0 || gy + (int64_t)g.h <= 0) continue;  // wholly outside
                L.glyphs.push_back(OverlayGlyph{(int32_t)gx, (int32_t)gy, g.w, g.h, g.offset});
            }
            // (clamping a rectangle's far-off corners changes nothing: pixels outside the frame are skipped one by one)
            int4 r = make_int4(std::max(rects[k].x, -1), std::max(rects[k].y, -1), std::min(rects[k].z, (int)w), std::min(rects[k].w, (int)h));
            if (r.z < r.x || r.w < r.y) r = make_int4(-1, -1, -2, -2);  // wholly outside: empty loops
            L.rects.push_back(r);
        }
        L.gstart.push_back((uint32_t)L.glyphs.size());
        return;
    }
    // no atlas: rectangles only, all at once (same colour: their order does not show). Clip rectangles that lie wholly
    // outside early (a box far off-frame would otherwise be a long empty loop)
    for (const int4& r : rects)
        if (r.z >= 0 && r.w >= 0 && r.x < (int)w && r.y < (int)h)
            L.rects.push_back(make_int4(std::max(r.x, -1), std::max(r.y, -1), std::min(r.z, (int)w), std::min(r.w, (int)h)));
}

// uploads the lists of the frames in one piece and draws every frame's overlay (one launch, CTA = frame) on its RGB image at
// rgb_base + rgb_off[k], w[k] x h[k]. Synchronises the stream before returning (the lists are pageable host memory).
static void draw_overlays(uf_model& m, Slot& s, const std::vector<OverlayLists>& L, uint8_t* rgb_base, const size_t* rgb_off, const uint32_t* w,
                          const uint32_t* h) {
    size_t n_r = 0, n_s = 0, n_g = 0;
    for (const auto& l : L) { n_r += l.rects.size(); n_s += l.gstart.size(); n_g += l.glyphs.size(); }
    if (n_r == 0) return;
    const size_t b_f = (L.size() * sizeof(OverlayFrame) + 15) / 16 * 16, b_r = n_r * sizeof(int4), b_s = (n_s * 4 + 15) / 16 * 16,
                 b_g = n_g * sizeof(OverlayGlyph);
    std::vector<uint8_t> host(b_f + b_r + b_s + b_g + 16);
    OverlayFrame* hf = reinterpret_cast<OverlayFrame*>(host.data());
    int4* hr = reinterpret_cast<int4*>(host.data() + b_f);
    uint32_t* hs = reinterpret_cast<uint32_t*>(host.data() + b_f + b_r);
    OverlayGlyph* hg = reinterpret_cast<OverlayGlyph*>(host.data() + b_f + b_r + b_s);
    if (host.size() > s.ovl_cap) {
        CK(cudaStreamSynchronize(s.stream));
        cudaFree(s.d_ovl);
        s.d_ovl = nullptr;
        s.ovl_cap = 0;
        CK(cudaMalloc(&s.d_ovl, host.size() * 2));
        s.ovl_cap = host.size() * 2;
    }
    uint8_t* d = s.d_ovl;
    size_t ir = 0, is = 0, ig = 0;
    for (size_t k = 0; k < L.size(); ++k) {
        const auto& l = L[k];
        hf[k] = OverlayFrame{(unsigned long long)rgb_off[k], (int32_t)w[k], (int32_t)h[k], (uint32_t)ir, (uint32_t)l.rects.size(), (uint32_t)is,
                             (uint32_t)ig, l.text ? 1u : 0u, 0u};
        if (!l.rects.empty()) memcpy(hr + ir, l.rects.data(), l.rects.size() * sizeof(int4));
        for (size_t q = 0; q < l.gstart.size(); ++q) hs[is + q] = l.gstart[q];  // (relative to the frame's first glyph)
        if (!l.glyphs.empty()) memcpy(hg + ig, l.glyphs.data(), l.glyphs.size() * sizeof(OverlayGlyph));
        ir += l.rects.size(); is += l.gstart.size(); ig += l.glyphs.size();
    }
    CK(cudaMemcpyAsync(d, host.data(), host.size(), cudaMemcpyHostToDevice, s.stream));
    m.launches++;
    launch_draw_overlay_batch(rgb_base, reinterpret_cast<const OverlayFrame*>(d), (int)L.size(), reinterpret_cast<const int4*>(d + b_f),
                              reinterpret_cast<const uint32_t*>(d + b_f + b_r), reinterpret_cast<const OverlayGlyph*>(d + b_f + b_r + b_s),
                              m.d_atlas, s.stream);
    CK(cudaStreamSynchronize(s.stream));
}

// draws on the RGB frame at s.d_in (device) and, if `file` is given, encodes it; everything on slot 0 of the locked lane
static void annotate_on_device(uf_model& m, Slot& s, uint32_t w, uint32_t h, const uf_det* dets, uint32_t n_dets, float scale_w, float scale_h,
                               int quality, std::vector<uint8_t>* file) {
    {
        std::lock_guard<std::mutex> atlas_lk(m.atlas_mu);  // (held until the overlay has been drawn: the atlas is in use)
        std::vector<OverlayLists> L(1);
        build_overlay_lists(m, w, h, dets, n_dets, scale_w, scale_h, L[0]);
        const size_t zero = 0;
        draw_overlays(m, s, L, s.d_in, &zero, &w, &h);
    }
    if (!file) return;
    const JpegPlan plan = jpeg_encode_plan(w, h, quality);
    const size_t coef_bytes = (size_t)plan.plane_bytes * sizeof(int16_t);
    grow_jpeg(s, coef_bytes, plan.plane_bytes);
    m.launches += 2;
    launch_jpeg_encode(s.d_in, plan, s.d_planes, reinterpret_cast<int16_t*>(s.d_jpeg), s.stream);
    CK(cudaMemcpyAsync(s.h_jpeg, s.d_jpeg, coef_bytes, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    CK(cudaGetLastError());
    jpeg_write_file(plan, reinterpret_cast<const int16_t*>(s.h_jpeg), *file);
}

static void deliver_file(const std::vector<uint8_t>& file, uint8_t* out, size_t cap, size_t* out_len) {
    *out_len = file.size();
    if (file.size() > cap || !out) throw ArgError(UF_ERR_CAPACITY, "output buffer smaller than the encoded file (" + std::to_string(file.size()) + " bytes)");
    memcpy(out, file.data(), file.size());
}

int uf_text_atlas_set(uf_model* m, const char* charset, uint32_t n_chars, uint32_t max_len, const uf_glyph* glyphs, const float* coverage,
                      size_t n_coverage) {
    return guarded([&] {
        REQUIRE(m, "null argument");
        std::lock_guard<std::mutex> lk(m->atlas_mu);
        CK(cudaSetDevice(m->cfg.device));
        if (n_chars == 0) {
            m->atlas_chars.clear();
            m->atlas_glyphs.clear();
            m->atlas_max_len = 0;
            return;
        }
        REQUIRE(charset && glyphs && max_len > 0 && max_len <= 64 && n_chars <= 128 && (coverage || n_coverage == 0), "bad argument");
        const size_t ng = (size_t)n_chars * max_len;
        for (size_t i = 0; i < ng; ++i) {
            const uf_glyph& g = glyphs[i];
            if (g.w == 0 || g.h == 0) continue;
            if (g.w > 4096 || g.h > 4096 || (size_t)g.offset + (size_t)g.w * g.h > n_coverage)
                throw ArgError(UF_ERR_INVALID_ARG, "glyph " + std::to_string(i) + " points outside the coverage array");
        }
        float* d = nullptr;
        CK(cudaMalloc(&d, std::max<size_t>(n_coverage, 1) * sizeof(float)));
        if (n_coverage) {
            cudaError_t e = cudaMemcpy(d, coverage, n_coverage * sizeof(float), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { cudaFree(d); throw CudaError(std::string("cudaMemcpy: ") + cudaGetErrorString(e)); }
        }
        cudaFree(m->d_atlas);
        m->d_atlas = d;
        m->atlas_n = n_coverage;
        m->atlas_chars.assign(charset, n_chars);
        m->atlas_max_len = max_len;
        m->atlas_glyphs.assign(glyphs, glyphs + ng);
    });
}

int uf_confidence_text(float confidence, char* out, size_t cap) {
    return guarded([&] {
        REQUIRE(out && cap >= 16, "bad argument");
        const std::string t = confidence_text(confidence);
        if (t.size() + 1 > cap) throw ArgError(UF_ERR_CAPACITY, "text buffer too small");
        memcpy(out, t.c_str(), t.size() + 1);
    });
}

int uf_annotate_encode_jpeg(uf_model* m, const uint8_t* rgb, uint32_t w, uint32_t h, const uf_det* dets, uint32_t n_dets, float scale_w,
                            float scale_h, uint32_t quality, uint8_t* out, size_t cap, size_t* out_len) {
    return guarded([&] {
        REQUIRE(m && rgb && out_len && w > 0 && h > 0 && w <= 16384 && h <= 16384 && (n_dets == 0 || dets), "bad argument");
        LaneLock ll(*m, true);
        Slot& s = ll.lane->slots[0];
        CK(cudaSetDevice(m->cfg.device));
        const size_t fb = (size_t)w * h * 3;
        grow_input(s, fb);
        CK(cudaMemcpyAsync(s.d_in, rgb, fb, cudaMemcpyHostToDevice, s.stream));
        std::vector<uint8_t> file;
        annotate_on_device(*m, s, w, h, dets, n_dets, scale_w, scale_h, (int)quality, &file);
        deliver_file(file, out, cap, out_len);
    });
}

int uf_annotate_reencode_jpeg(uf_model* m, const uint8_t* jpeg, size_t len, const uf_det* dets, uint32_t n_dets, float scale_w,
                              float scale_h, uint32_t quality, uint8_t* out, size_t cap, size_t* out_len) {
    return guarded([&] {
        REQUIRE(m && jpeg && out_len && (n_dets == 0 || dets), "bad argument");
        JpegCoefs jc;
        jpeg_entropy_decode(jpeg, len, jc);
        LaneLock ll(*m, true);
        Slot& s = ll.lane->slots[0];
        CK(cudaSetDevice(m->cfg.device));
        const size_t fb = (size_t)jc.plan.w * jc.plan.h * 3;
        grow_input(s, fb);
        grow_jpeg(s, jpeg_stage_bytes(jc), jc.plan.plane_bytes);
        FrameSrc fr{nullptr, jc.plan.w, jc.plan.h, &jc};
        size_t ju = 0, pu = 0;
        decode_jpeg_run(*m, s, &fr, 1, s.d_in, ju, pu);
        std::vector<uint8_t> file;
        annotate_on_device(*m, s, jc.plan.w, jc.plan.h, dets, n_dets, scale_w, scale_h, (int)quality, &file);
        deliver_file(file, out, cap, out_len);
    });
}

// waits for everything queued on the slot's stream, asleep (the waits of the annotate path are long: a stage of frames each)
static void sleep_sync(uf_model& m, Slot& s) {
    (void)m;
    CK(cudaEventRecord(s.done_ev, s.stream));
    CK(cudaEventSynchronize(s.done_ev));
    CK(cudaStreamSynchronize(s.stream));
}

// One chunk of the batch form: frames [0, cnt) of jpeg/len, their detections at dets + det_first[k]; files to out + k * stride.
// Everything runs on slot 0 of the caller's lane, stage after stage, with three waits for the device (decode status, the
// sizes of the entropy-coded segments, the segments themselves).
// det_out != NULL: the detections are not given but FOUND here, on the decoded frames (the hot path, device-resident input),
// returned in det_out / n_out and drawn.
static void reencode_chunk(uf_model& m, Lane& ln, Slot& s, const uint8_t* const* jpeg, const size_t* len, uint32_t cnt, const uf_det* dets,
                           const uint32_t* det_first, const uint32_t* det_counts, float scale_w, float scale_h, int quality, uint8_t* out,
                           size_t stride, size_t* out_len, uint32_t first_index, uf_det* det_out = nullptr, uint32_t det_cap = 0,
                           uint32_t* n_out = nullptr) {
    auto a256 = [](size_t v) { return (v + 255) / 256 * 256; };
    // 1. headers, byte unstuffing (host threads)
    std::vector<JpegBitstream> bs(cnt);
    std::vector<JpegCoefs> hostc(cnt);
    std::vector<JpegError> errs(cnt, JpegError{UF_OK, ""});
    m.pool().parallel_for(cnt, [&](uint32_t i) {
        try {
            jpeg_prepare_bitstream(jpeg[i], len[i], bs[i]);
            if (!bs[i].gpu_ok) jpeg_entropy_decode(jpeg[i], len[i], hostc[i]);
        } catch (const JpegError& e) { errs[i] = e; }
        catch (const std::exception& e) { errs[i] = JpegError{UF_ERR_INVALID_ARG, e.what()}; }
    });
    for (uint32_t i = 0; i < cnt; ++i)
        if (errs[i].code != UF_OK) throw JpegError{errs[i].code, "frame " + std::to_string(first_index + i) + ": " + errs[i].msg};
    // 2. sizes: RGB frames, decoder staging / scratch, encoder planes / coefficients / bit buffers
    std::vector<size_t> rgb_off(cnt), pl_off(cnt), co_off(cnt);
    std::vector<JpegPlan> eplan(cnt);
    std::vector<FrameSrc> fr(cnt);
    size_t rgb_bytes = 0, jpeg_need = 0, planes_dec = 0, huff_dec = 0, planes_enc = 0, coef_bytes = 0, nblk_all = 0, pack_words = 0, out_bytes = 0;
    uint32_t max_nblk = 0;
    for (uint32_t k = 0; k < cnt; ++k) {
        const JpegPlan& dp = bs[k].plan;
        rgb_off[k] = rgb_bytes;
        rgb_bytes += (size_t)dp.w * dp.h * 3;
        if (k + 1 < cnt && (bs[k + 1].plan.w != dp.w || bs[k + 1].plan.h != dp.h || bs[k + 1].gpu_ok != bs[k].gpu_ok)) rgb_bytes = a256(rgb_bytes);
        if (bs[k].gpu_ok) {
            fr[k] = FrameSrc{nullptr, dp.w, dp.h, nullptr, &bs[k]};
            jpeg_need += huff_stage_bytes(bs[k]);
            huff_dec += huff_scratch_bytes(bs[k]);
        } else {
            fr[k] = FrameSrc{nullptr, dp.w, dp.h, &hostc[k], nullptr};
            jpeg_need += jpeg_stage_bytes(hostc[k]);
        }
        // (a frame the device decoder hands back is redone through the host decoder's staging: reserve for it too)
        planes_dec += dp.plane_bytes;
        eplan[k] = jpeg_encode_plan(dp.w, dp.h, quality);
        pl_off[k] = planes_enc;
        planes_enc += a256(eplan[k].plane_bytes);
        co_off[k] = coef_bytes;
        coef_bytes += a256((size_t)eplan[k].plane_bytes * sizeof(int16_t));
        nblk_all += eplan[k].nblocks;
        max_nblk = std::max(max_nblk, eplan[k].nblocks);
        pack_words += (size_t)eplan[k].nblocks * 24;          // 96 bytes of bit buffer per block (a q95 block takes ~20)
        out_bytes += a256((size_t)eplan[k].nblocks * 120);    // + room for the stuffed zeros
    }
    const size_t e_frames = a256(coef_bytes), e_tab = a256(e_frames + cnt * sizeof(JpegEncFrame)), e_len = a256(e_tab + sizeof(JpegEncTables)),
                 e_off = a256(e_len + nblk_all * 4), e_fb = a256(e_off + nblk_all * 4), e_ol = a256(e_fb + cnt * 4), e_pack = a256(e_ol + cnt * 4),
                 e_out = a256(e_pack + pack_words * 4), e_end = e_out + out_bytes;
    grow_input(s, rgb_bytes + 256);
    grow_jpeg(s, std::max(jpeg_need, a256(cnt * sizeof(JpegEncFrame)) + sizeof(JpegEncTables) + cnt * sizeof(JpegEncJob) + 1024),
              std::max(planes_dec, planes_enc));
    grow_huff(s, std::max(huff_dec, e_end));
    // 3. decode: runs of same-size frames; then the frames the device decoder handed back, one by one
    size_t ju = 0, pu = 0, hu = 0;
    for (uint32_t k = 0; k < cnt; ++k) {
        s.h_jstatus[k] = 0;
        if (fr[k].w > 8192 || fr[k].h > 8192) throw ArgError(UF_ERR_UNSUPPORTED, "frame " + std::to_string(first_index + k) + ": larger than 8192 x 8192");
    }
    for (uint32_t i = 0; i < cnt;) {
        uint32_t j = i + 1;
        while (j < cnt && fr[j].w == fr[i].w && fr[j].h == fr[i].h && (fr[j].jb != nullptr) == (fr[i].jb != nullptr)) ++j;
        if (fr[i].jb) decode_jpeg_run_gpu(m, s, fr.data() + i, i, j - i, s.d_in + rgb_off[i], ju, pu, hu);
        else decode_jpeg_run(m, s, fr.data() + i, j - i, s.d_in + rgb_off[i], ju, pu);
        i = j;
    }
    sleep_sync(m, s);
    s.jstatus_n = 0;
    for (uint32_t k = 0; k < cnt; ++k) {
        if (!fr[k].jb || s.h_jstatus[k] == 0) continue;
        jpeg_entropy_decode(jpeg[k], len[k], hostc[k]);
        grow_jpeg(s, jpeg_stage_bytes(hostc[k]), hostc[k].plan.plane_bytes);
        FrameSrc one{nullptr, hostc[k].plan.w, hostc[k].plan.h, &hostc[k], nullptr};
        size_t ju1 = 0, pu1 = 0;
        decode_jpeg_run(m, s, &one, 1, s.d_in + rgb_off[k], ju1, pu1);
        sleep_sync(m, s);
        m.jpeg_redone++;
    }
    // 3b. the hot path on the decoded frames, run of same-size frames by run (they only read the pixels)
    std::vector<uint32_t> found_first, found_counts;
    if (det_out) {
        for (uint32_t i = 0; i < cnt;) {
            uint32_t j = i + 1;
            while (j < cnt && fr[j].w == fr[i].w && fr[j].h == fr[i].h && rgb_off[j] == rgb_off[j - 1] + (size_t)fr[i].w * fr[i].h * 3) ++j;
            const uint8_t* d_rgb = s.d_in + rgb_off[i];
            const uint32_t w = fr[i].w, h = fr[i].h;
            const size_t fb = (size_t)w * h * 3;
            run_pipeline(m, ln, j - i, m.chunk, false, det_out + (size_t)i * det_cap, det_cap, n_out + i,
                         [&](Slot& sl, uint32_t first, uint32_t c) { run_chunk_device(m, ln, sl, d_rgb + (size_t)first * fb, w, h, first, c); });
            i = j;
        }
        found_first.resize(cnt);
        found_counts.resize(cnt);
        for (uint32_t k = 0; k < cnt; ++k) {
            found_first[k] = k * det_cap;
            found_counts[k] = std::min(n_out[k], det_cap);  // (the overlay shows what the caller gets)
        }
        dets = det_out;
        det_first = found_first.data();
        det_counts = found_counts.data();
    }
    // 4. overlay
    {
        std::lock_guard<std::mutex> atlas_lk(m.atlas_mu);
        std::vector<OverlayLists> L(cnt);
        std::vector<uint32_t> ws(cnt), hs(cnt);
        for (uint32_t k = 0; k < cnt; ++k) {
            ws[k] = fr[k].w; hs[k] = fr[k].h;
            build_overlay_lists(m, ws[k], hs[k], dets + det_first[k], det_counts[k], scale_w, scale_h, L[k]);
        }
        draw_overlays(m, s, L, s.d_in, rgb_off.data(), ws.data(), hs.data());
    }
    // 5. colour conversion, downsampling, forward DCT, quantisation — then Huffman coding and byte stuffing, all frames at once
    int16_t* d_coefs = reinterpret_cast<int16_t*>(s.d_huff);
    JpegEncFrame* hf = reinterpret_cast<JpegEncFrame*>(s.h_jpeg);
    const size_t o_jobs = a256(a256(cnt * sizeof(JpegEncFrame)) + sizeof(JpegEncTables));
    JpegEncJob* hj = reinterpret_cast<JpegEncJob*>(s.h_jpeg + o_jobs);
    size_t lb = 0, pw = 0, ob = 0;
    uint32_t max_cw = 0, max_ch = 0, max_pblocks = 0;
    for (uint32_t k = 0; k < cnt; ++k) {
        const JpegPlan& p = eplan[k];
        hj[k] = JpegEncJob{p, (unsigned long long)rgb_off[k], (unsigned long long)pl_off[k], (unsigned long long)(co_off[k] / sizeof(int16_t))};
        max_cw = std::max(max_cw, p.plane_w[1]);
        max_ch = std::max(max_ch, p.plane_h[1]);
        max_pblocks = std::max(max_pblocks, (p.plane_w[0] * p.plane_h[0] + 2 * p.plane_w[1] * p.plane_h[1]) / 64);
        JpegEncFrame f{};
        f.coef_base = (uint32_t)(co_off[k] / 128);
        f.mcus_x = p.mcus_x; f.mcus_y = p.mcus_y;
        f.y_bw = p.plane_w[0] / 8; f.c_bw = p.plane_w[1] / 8;
        f.cb_off = p.plane_off[1] / 64; f.cr_off = p.plane_off[2] / 64;
        f.wib0 = (p.real_w[0] + 7) / 8; f.hib0 = (p.real_h[0] + 7) / 8;
        f.nblocks = p.nblocks;
        f.len_base = (uint32_t)lb; lb += p.nblocks;
        f.pack_off = (uint32_t)pw; f.pack_cap_bits = p.nblocks * 24 * 32; pw += (size_t)p.nblocks * 24;
        f.out_off = (uint32_t)ob; f.out_cap = (uint32_t)a256((size_t)p.nblocks * 120); ob += f.out_cap;
        hf[k] = f;
    }
    JpegEncTables* ht = reinterpret_cast<JpegEncTables*>(s.h_jpeg + a256(cnt * sizeof(JpegEncFrame)));
    jpeg_std_enc_tables(*ht);
    CK(cudaMemcpyAsync(s.d_jpeg + o_jobs, hj, cnt * sizeof(JpegEncJob), cudaMemcpyHostToDevice, s.stream));
    m.launches += 2;
    launch_jpeg_encode_batch(s.d_in, reinterpret_cast<const JpegEncJob*>(s.d_jpeg + o_jobs), (int)cnt, max_cw, max_ch, max_pblocks, s.d_planes,
                             d_coefs, s.stream);
    CK(cudaMemcpyAsync(s.d_huff + e_frames, hf, cnt * sizeof(JpegEncFrame), cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.d_huff + e_tab, ht, sizeof(JpegEncTables), cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemsetAsync(s.d_huff + e_pack, 0, pack_words * 4, s.stream));
    JpegEncBatch B{reinterpret_cast<const JpegEncFrame*>(s.d_huff + e_frames), d_coefs, reinterpret_cast<const JpegEncTables*>(s.d_huff + e_tab),
                   reinterpret_cast<uint32_t*>(s.d_huff + e_len), reinterpret_cast<uint32_t*>(s.d_huff + e_off),
                   reinterpret_cast<uint32_t*>(s.d_huff + e_fb), reinterpret_cast<uint32_t*>(s.d_huff + e_pack), s.d_huff + e_out,
                   reinterpret_cast<uint32_t*>(s.d_huff + e_ol)};
    m.launches += 4;
    launch_jpeg_huffman_encode(B, (int)cnt, max_nblk, s.stream);
    uint32_t* h_len = reinterpret_cast<uint32_t*>(s.h_jstatus);
    CK(cudaMemcpyAsync(h_len, s.d_huff + e_ol, cnt * 4, cudaMemcpyDeviceToHost, s.stream));
    sleep_sync(m, s);
    // 6. files: the segments come back through the pinned staging buffer in one burst of copies, host threads then write
    //    headers + segment + EOI into the caller's buffer
    std::vector<size_t> seg_off(cnt, 0);
    size_t seg_bytes = 0;
    for (uint32_t k = 0; k < cnt; ++k)
        if (h_len[k] != 0xffffffffu) { seg_off[k] = seg_bytes; seg_bytes += ((size_t)h_len[k] + 15) / 16 * 16; }
    std::vector<uint32_t> seg_len(h_len, h_len + cnt);  // (h_len lives in the slot's status array)
    std::vector<JpegEncFrame> encf(hf, hf + cnt);       // (and hf in the staging buffer that may be regrown now)
    grow_jpeg(s, seg_bytes + 16, 0);
    for (uint32_t k = 0; k < cnt; ++k)
        if (seg_len[k] != 0xffffffffu && seg_len[k] > 0)
            CK(cudaMemcpyAsync(s.h_jpeg + seg_off[k], s.d_huff + e_out + encf[k].out_off, seg_len[k], cudaMemcpyDeviceToHost, s.stream));
    sleep_sync(m, s);
    CK(cudaGetLastError());
    std::atomic<bool> too_small{false};
    std::vector<std::string> fails(cnt);
    m.pool().parallel_for(cnt, [&](uint32_t k) {
        try {
            uint8_t* dst = out + (size_t)k * stride;
            std::vector<uint8_t> head;
            if (seg_len[k] == 0xffffffffu) return;  // (below)
            jpeg_write_headers(eplan[k], head);
            out_len[k] = head.size() + seg_len[k] + 2;
            if (out_len[k] > stride) { too_small = true; return; }
            memcpy(dst, head.data(), head.size());
            memcpy(dst + head.size(), s.h_jpeg + seg_off[k], seg_len[k]);
            dst[head.size() + seg_len[k]] = 0xff;
            dst[head.size() + seg_len[k] + 1] = 0xd9;
        } catch (const std::exception& e) { fails[k] = e.what(); }
    });
    for (uint32_t k = 0; k < cnt; ++k) {
        if (!fails[k].empty()) throw CudaError("frame " + std::to_string(first_index + k) + ": " + fails[k]);
        if (seg_len[k] != 0xffffffffu) continue;
        // did not fit the device buffers (far denser than any camera frame): the host encoder
        std::vector<int16_t> co((size_t)eplan[k].plane_bytes);
        std::vector<uint8_t> file;
        CK(cudaMemcpy(co.data(), d_coefs + co_off[k] / sizeof(int16_t), co.size() * sizeof(int16_t), cudaMemcpyDeviceToHost));
        jpeg_write_file(eplan[k], co.data(), file);
        out_len[k] = file.size();
        if (file.size() > stride) too_small = true;
        else memcpy(out + (size_t)k * stride, file.data(), file.size());
    }
    if (too_small.load()) throw ArgError(UF_ERR_CAPACITY, "an encoded file is larger than out_stride (out_len holds the sizes)");
}

// The chunks of a batch for the annotate path, several at a time: a chunk walks its stages one after another with waits in
// between, so up to one host thread per lane takes chunks from the list and runs each on a lane of its own (a caller with
// one thread then gets what a caller with several calls in flight gets). fn(lane, slot, first frame, count).
static void run_annotate_chunks(uf_model& m, uint32_t n, const std::function<void(Lane&, Slot&, uint32_t, uint32_t)>& fn) {
    const uint32_t step = std::max<uint32_t>(1, std::min<uint32_t>(64, m.chunk));
    const uint32_t nchunks = (n + step - 1) / step;
    std::atomic<uint32_t> next{0};
    std::atomic<bool> too_small{false}, stop{false};
    std::exception_ptr err;
    std::mutex err_mu;
    auto work = [&] {
        cudaSetDevice(m.cfg.device);
        for (;;) {
            const uint32_t c = next.fetch_add(1);
            if (c >= nchunks || stop.load()) return;
            try {
                LaneLock ll(m, false);
                fn(*ll.lane, ll.lane->slots[0], c * step, std::min(step, n - c * step));
                ll.lane->last_n = 0;
            } catch (const ArgError& e) {
                if (e.code == UF_ERR_CAPACITY) { too_small = true; continue; }  // the other chunks still report their sizes
                std::lock_guard<std::mutex> lk(err_mu);
                if (!err) err = std::current_exception();
                stop = true;
            } catch (...) {
                std::lock_guard<std::mutex> lk(err_mu);
                if (!err) err = std::current_exception();
                stop = true;
            }
        }
    };
    const uint32_t nthreads = m.profiling ? 1 : std::min<uint32_t>(nchunks, (uint32_t)std::min<size_t>(m.lanes.size(), 4));
    std::vector<std::thread> ts;
    for (uint32_t t = 1; t < nthreads; ++t) ts.emplace_back(work);
    work();
    for (auto& t : ts) t.join();
    if (err) std::rethrow_exception(err);
    if (too_small.load()) throw ArgError(UF_ERR_CAPACITY, "an encoded file is larger than out_stride (out_len holds the sizes)");
}

int uf_annotate_reencode_batch_jpeg(uf_model* m, const uint8_t* const* jpeg, const size_t* len, uint32_t n, const uf_det* dets,
                                    const uint32_t* det_counts, float scale_w, float scale_h, uint32_t quality, uint8_t* out, size_t out_stride,
                                    size_t* out_len) {
    return guarded([&] {
        REQUIRE(m && out_len && (n == 0 || (jpeg && len && det_counts && out)) && out_stride >= 1024, "bad argument");
        for (uint32_t i = 0; i < n; ++i) REQUIRE(jpeg[i] && len[i] > 0, "null frame");
        std::vector<uint32_t> first(n + 1, 0);
        for (uint32_t i = 0; i < n; ++i) first[i + 1] = first[i] + det_counts[i];
        REQUIRE(first[n] == 0 || dets, "null detections");
        CK(cudaSetDevice(m->cfg.device));
        run_annotate_chunks(*m, n, [&](Lane& ln, Slot& s, uint32_t f0, uint32_t cnt) {
            reencode_chunk(*m, ln, s, jpeg + f0, len + f0, cnt, dets, first.data() + f0, det_counts + f0, scale_w, scale_h, (int)quality,
                           out + (size_t)f0 * out_stride, out_stride, out_len + f0, f0);
        });
    });
}

int uf_worker_batch_jpeg(uf_model* m, const uint8_t* const* jpeg, const size_t* len, uint32_t n, float scale_w, float scale_h, uint32_t quality,
                         uf_det* dets, uint32_t cap, uint32_t* n_dets, uint8_t* out, size_t out_stride, size_t* out_len) {
    return guarded([&] {
        REQUIRE(m && out_len && n_dets && cap > 0 && dets && (n == 0 || (jpeg && len && out)) && out_stride >= 1024, "bad argument");
        for (uint32_t i = 0; i < n; ++i) REQUIRE(jpeg[i] && len[i] > 0, "null frame");
        CK(cudaSetDevice(m->cfg.device));
        run_annotate_chunks(*m, n, [&](Lane& ln, Slot& s, uint32_t f0, uint32_t cnt) {
            reencode_chunk(*m, ln, s, jpeg + f0, len + f0, cnt, nullptr, nullptr, nullptr, scale_w, scale_h, (int)quality,
                           out + (size_t)f0 * out_stride, out_stride, out_len + f0, f0, dets + (size_t)f0 * cap, cap, n_dets + f0);
        });
    });
}

int uf_jpeg_write_coefficients(uint32_t w, uint32_t h, uint32_t quality, const int16_t* coefs, size_t n_blocks, uint8_t* out, size_t cap,
                               size_t* out_len) {
    return guarded([&] {
        REQUIRE(coefs && out_len && w > 0 && h > 0 && w <= 16384 && h <= 16384, "bad argument");
        const JpegPlan plan = jpeg_encode_plan(w, h, (int)quality);
        const size_t need = ((size_t)plan.plane_w[0] * plan.plane_h[0] + 2 * (size_t)plan.plane_w[1] * plan.plane_h[1]) / 64;
        REQUIRE(n_blocks == need, "n_blocks does not match the padded planes of a 4:2:0 frame of this size (" + std::to_string(need) + ")");
        std::vector<uint8_t> file;
        jpeg_write_file(plan, coefs, file);
        deliver_file(file, out, cap, out_len);
    });
}

int uf_jpeg_quality_tables(uint32_t quality, uint16_t* lum64, uint16_t* chr64) {
    return guarded([&] {
        REQUIRE(lum64 && chr64, "null argument");
        jpeg_quality_tables((int)quality, lum64, chr64);
    });
}

int uf_draw_boxes_rgb(uf_model* m, const uint8_t* rgb, uint32_t w, uint32_t h, const uf_det* dets, uint32_t n_dets, float scale_w,
                      float scale_h, uint8_t* out_rgb) {
    return guarded([&] {
        REQUIRE(m && rgb && out_rgb && w > 0 && h > 0 && w <= 16384 && h <= 16384 && (n_dets == 0 || dets), "bad argument");
        LaneLock ll(*m, true);
        Slot& s = ll.lane->slots[0];
        CK(cudaSetDevice(m->cfg.device));
        const size_t fb = (size_t)w * h * 3;
        grow_input(s, fb);
        CK(cudaMemcpyAsync(s.d_in, rgb, fb, cudaMemcpyHostToDevice, s.stream));
        annotate_on_device(*m, s, w, h, dets, n_dets, scale_w, scale_h, 0, nullptr);
        CK(cudaMemcpyAsync(out_rgb, s.d_in, fb, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaGetLastError());
    });
}

int uf_infer(uf_model* m, const uint8_t* rgb, uint32_t w, uint32_t h, uf_det* out, uint32_t cap, uint32_t* n_out) {
    const uint8_t* ptrs[1] = {rgb};
    return uf_infer_batch(m, ptrs, &w, &h, 1, out, cap, n_out);
}

int uf_infer_batch_device(uf_model* m, const uint8_t* d_rgb, uint32_t w, uint32_t h, uint32_t n, uf_det* out,
                          uint32_t cap, uint32_t* n_out) {
    return guarded([&] {
        REQUIRE(m && (n == 0 || d_rgb) && n_out && w > 0 && h > 0, "bad argument");
        REQUIRE(cap == 0 || out, "out is NULL with cap > 0");
        LaneLock ll(*m, false);
        Lane& ln = *ll.lane;
        const size_t fb = (size_t)w * h * 3;
        run_pipeline(*m, ln, n, m->chunk, false, out, cap, n_out, [&](Slot& s, uint32_t first, uint32_t cnt) {
            run_chunk_device(*m, ln, s, d_rgb + (size_t)first * fb, w, h, first, cnt);
        });
    });
}

int uf_raw_outputs(uf_model* m, uint32_t first, uint32_t n, float* scores, float* boxes) {
    return guarded([&] {
        REQUIRE(m, "null model");
        LaneLock ll(*m, true);
        Lane& ln = *ll.lane;
        REQUIRE((uint64_t)first + n <= ln.last_n, "frame range outside the last batch");
        CK(cudaSetDevice(m->cfg.device));
        if (scores) CK(cudaMemcpy(scores, ln.d_scores + (size_t)first * m->K * 2, (size_t)n * m->K * 2 * sizeof(float), cudaMemcpyDeviceToHost));
        if (boxes) CK(cudaMemcpy(boxes, ln.d_boxes + (size_t)first * m->K * 4, (size_t)n * m->K * 4 * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

static void preproc_to_slot0(uf_model* m, Lane& ln, const uint8_t* rgb, uint32_t w, uint32_t h) {
    Slot& s = ln.slots[0];
    const int W = m->plan.net_w, H = m->plan.net_h;
    const size_t fb = (size_t)w * h * 3;
    if ((int)w == W && (int)h == H) {
        CK(cudaMemcpyAsync(s.d_resized, rgb, fb, cudaMemcpyHostToDevice, s.stream));
    } else {
        grow_input(s, fb);
        CK(cudaMemcpyAsync(s.d_in, rgb, fb, cudaMemcpyHostToDevice, s.stream));
        TapsRef t = get_taps(*m, w, h);
        m->launches++;
        launch_resize(s.d_in, (long long)fb, w, h, s.d_resized, (long long)W * H * 3, W, H, 1, t->dev,
                      m->cfg.resize_round_intermediate, s.stream);
    }
}

int uf_preproc_u8(uf_model* m, const uint8_t* rgb, uint32_t w, uint32_t h, uint8_t* out_u8) {
    return guarded([&] {
        REQUIRE(m && rgb && out_u8 && w > 0 && h > 0, "bad argument");
        LaneLock ll(*m, true);
        Lane& ln = *ll.lane;
        CK(cudaSetDevice(m->cfg.device));
        preproc_to_slot0(m, ln, rgb, w, h);
        Slot& s = ln.slots[0];
        CK(cudaMemcpyAsync(out_u8, s.d_resized, (size_t)m->plan.net_w * m->plan.net_h * 3, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaGetLastError());
    });
}

int uf_preproc_u8_batch(uf_model* m, const uint8_t* rgb, uint32_t w, uint32_t h, uint32_t n, uint8_t* out_u8) {
    return guarded([&] {
        REQUIRE(m && rgb && out_u8 && w > 0 && h > 0 && n > 0, "bad argument");
        if (n > m->chunk) throw ArgError(UF_ERR_CAPACITY, "uf_preproc_u8_batch: n exceeds the stage size (uf_config.chunk)");
        LaneLock ll(*m, true);
        Slot& s = ll.lane->slots[0];
        CK(cudaSetDevice(m->cfg.device));
        const int W = m->plan.net_w, H = m->plan.net_h;
        const size_t fb = (size_t)w * h * 3, ob = (size_t)W * H * 3;
        if ((int)w == W && (int)h == H) {
            CK(cudaMemcpyAsync(s.d_resized, rgb, fb * n, cudaMemcpyHostToDevice, s.stream));
        } else {
            grow_input(s, fb * n);
            CK(cudaMemcpyAsync(s.d_in, rgb, fb * n, cudaMemcpyHostToDevice, s.stream));
            TapsRef t = get_taps(*m, w, h);
            m->launches++;
            launch_resize(s.d_in, (long long)fb, w, h, s.d_resized, (long long)ob, W, H, (int)n, t->dev,
                          m->cfg.resize_round_intermediate, s.stream);
        }
        CK(cudaMemcpyAsync(out_u8, s.d_resized, ob * n, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaGetLastError());
    });
}

int uf_debug_prestem_u8(uf_model* m, const uint8_t* rgb, uint32_t w, uint32_t h, uint32_t n, uint8_t* out_u8) {
    return guarded([&] {
        REQUIRE(m && rgb && out_u8 && w > 0 && h > 0 && n > 0, "bad argument");
        if (n > m->chunk) throw ArgError(UF_ERR_CAPACITY, "uf_debug_prestem_u8: n exceeds the stage size (uf_config.chunk)");
        LaneLock ll(*m, true);
        Slot& s = ll.lane->slots[0];
        CK(cudaSetDevice(m->cfg.device));
        const int W = m->plan.net_w, H = m->plan.net_h;
        const size_t fb = (size_t)w * h * 3, ob = (size_t)W * H * 3;
        if (!m->prestem_ok) throw ArgError(UF_ERR_UNSUPPORTED, "this model's first layer does not run the fused resize + stem kernel");
        grow_input(s, fb * n);
        TapsRef t = get_taps(*m, w, h);
        if (!prestem_supported(s.d_in, (long long)fb, w, h, W, H, t->dev))
            throw ArgError(UF_ERR_UNSUPPORTED, "frames are not at exactly twice the network size");
        CK(cudaMemcpyAsync(s.d_in, rgb, fb * n, cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemsetAsync(s.d_resized, 0xA5, ob * n, s.stream));
        const Step& st = m->steps[0];
        const Op& op = m->plan.ops[st.op];
        m->launches++;
        launch_prestem(U8View{s.d_in, (long long)fb, (int)h, (int)w}, m->d_lut, t->dev, make_view(*m, s, op.out), st.host_w.data(), op.relu,
                       (int)m->cfg.resize_round_intermediate, (int)n, s.d_resized, s.stream);
        CK(cudaMemcpyAsync(out_u8, s.d_resized, ob * n, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaGetLastError());
    });
}

int uf_preproc_f32(uf_model* m, const uint8_t* rgb, uint32_t w, uint32_t h, float* out) {
    return guarded([&] {
        REQUIRE(m && rgb && out && w > 0 && h > 0, "bad argument");
        LaneLock ll(*m, true);
        Lane& ln = *ll.lane;
        CK(cudaSetDevice(m->cfg.device));
        preproc_to_slot0(m, ln, rgb, w, h);
        Slot& s = ln.slots[0];
        const size_t nfl = (size_t)3 * m->plan.net_w * m->plan.net_h;
        float* d = (float*)hook_scratch(*m, nfl * sizeof(float));
        m->launches++;
        launch_normalise_nchw(s.d_resized, m->plan.net_w, m->plan.net_h, m->d_lut, d, s.stream);
        CK(cudaMemcpyAsync(out, d, nfl * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaGetLastError());
    });
}

int uf_postproc(uf_model* m, const float* scores, const float* boxes, uint32_t K, uf_det* out, uint32_t cap,
                uint32_t* n_out, int32_t* out_prior_idx) {
    return guarded([&] {
        REQUIRE(m && n_out && (K == 0 || (scores && boxes)), "null argument");
        REQUIRE(cap == 0 || out, "out is NULL with cap > 0");
        REQUIRE(K <= (1u << 20), "K too large");
        *n_out = 0;
        if (K == 0) return;
        LaneLock ll(*m, true);
        Lane& ln = *ll.lane;
        CK(cudaSetDevice(m->cfg.device));
        Slot& s = ln.slots[0];
        const size_t sort_cap = post_sort_scratch_elems((int)K);
        // layout: scores | boxes | sel | dets | idx | count | sort keys
        auto a16 = [](size_t v) { return (v + 15) / 16 * 16; };
        const size_t o_scores = 0, o_boxes = a16(o_scores + (size_t)K * 2 * 4), o_sel = a16(o_boxes + (size_t)K * 4 * 4),
                     o_dets = a16(o_sel + (size_t)K * 4 * 4), o_idx = a16(o_dets + (size_t)K * 5 * 4),
                     o_cnt = a16(o_idx + (size_t)K * 4), o_big = a16(o_cnt + 4), o_sort = a16(o_big + 8),
                     o_mask = a16(o_sort + sort_cap * 8),
                     total = o_mask + (post_mask_supported((int)K) ? post_mask_words((int)K) * 4 : 0);
        char* d = (char*)hook_scratch(*m, total);
        CK(cudaMemcpyAsync(d + o_scores, scores, (size_t)K * 2 * 4, cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(d + o_boxes, boxes, (size_t)K * 4 * 4, cudaMemcpyHostToDevice, s.stream));
        PostBuffers pb{(unsigned long long*)(d + o_sort), (int)sort_cap, (float*)(d + o_sel), (float*)(d + o_dets),
                       (int*)(d + o_idx), (int*)(d + o_cnt),
                       post_mask_supported((int)K) ? (unsigned*)(d + o_mask) : nullptr, post_mask_pitch((int)K), (int*)(d + o_big),
                       (int*)(d + o_big) + 1};
        m->launches += pb.mask ? 3 : 1;
        CK(cudaMemsetAsync(d + o_big, 0, 8, s.stream));
        launch_post((const float*)(d + o_scores), (const float*)(d + o_boxes), (int)K, m->cfg.min_confidence,
                    m->cfg.max_iou, pb, 1, s.stream);
        if (pb.mask) {
            launch_nms_mask((int)K, m->cfg.max_iou, pb, 1, s.stream);
            launch_nms_sweep((const float*)(d + o_scores), (int)K, pb, 1, s.stream);
        }
        int cnt = 0;
        CK(cudaMemcpyAsync(&cnt, d + o_cnt, 4, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaGetLastError());
        *n_out = (uint32_t)cnt;
        const uint32_t take = std::min<uint32_t>((uint32_t)cnt, cap);
        if (take) {
            CK(cudaMemcpy(out, d + o_dets, (size_t)take * sizeof(uf_det), cudaMemcpyDeviceToHost));
            if (out_prior_idx) CK(cudaMemcpy(out_prior_idx, d + o_idx, (size_t)take * 4, cudaMemcpyDeviceToHost));
        }
    });
}

int uf_tensor_count(const uf_model* m, uint32_t* n) {
    return guarded([&] {
        REQUIRE(m && n, "null argument");
        *n = (uint32_t)m->plan.tensors.size();
    });
}

int uf_tensor_info(const uf_model* m, uint32_t i, const char** onnx_name, uint32_t* c, uint32_t* h, uint32_t* w) {
    return guarded([&] {
        REQUIRE(m && i < m->plan.tensors.size(), "tensor index out of range");
        const TensorDesc& t = m->plan.tensors[i];
        if (onnx_name) *onnx_name = m->tensor_readable[i] ? t.name.c_str() : "";
        if (c) *c = t.C;
        if (h) *h = t.H;
        if (w) *w = t.W;
    });
}

int uf_tensor_read(uf_model* m, uint32_t i, uint32_t frame, float* out_nchw) {
    return guarded([&] {
        REQUIRE(m && out_nchw && i < m->plan.tensors.size(), "bad argument");
        REQUIRE(m->tensor_readable[i], "tensor is not materialised (graph input, or fused away)");
        LaneLock ll(*m, true);
        Lane& ln = *ll.lane;
        REQUIRE(frame < ln.last_n, "frame outside the last batch");
        Slot* found = nullptr;
        for (auto& sl : ln.slots)
            if (sl.batch_id == ln.batch_id && frame >= sl.first && frame < sl.first + sl.n) found = &sl;
        REQUIRE(found != nullptr, "that frame's workspace has been reused by a later pipeline stage");
        CK(cudaSetDevice(m->cfg.device));
        Slot& s = *found;
        TView v = make_view(*m, s, (int)i);
        const size_t nfl = (size_t)v.C * v.H * v.W;
        float* d = (float*)hook_scratch(*m, nfl * sizeof(float));
        launch_nhwc_to_nchw(v, (int)(frame - s.first), d, s.stream);
        CK(cudaMemcpyAsync(out_nchw, d, nfl * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaGetLastError());
    });
}

int uf_host_alloc(size_t bytes, void** out) {
    return guarded([&] {
        REQUIRE(out, "null argument");
        *out = nullptr;
        CK(cudaMallocHost(out, bytes ? bytes : 1));
    });
}

void uf_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int uf_profile_enable(uf_model* m, int on) {
    return guarded([&] {
        REQUIRE(m, "null model");
        LaneLock ll(*m, true);
        m->profiling = on != 0;
    });
}

int uf_profile_reset(uf_model* m) {
    return guarded([&] {
        REQUIRE(m, "null model");
        std::lock_guard<std::mutex> lk(m->aux_mu);
        for (auto& s : m->stats) { s.launches = 0; s.ms = 0; s.alg_bytes = 0; s.min_bytes = 0; s.flops = 0; }
    });
}

int uf_profile_read(uf_model* m, uf_kernel_stat* out, uint32_t cap, uint32_t* n_out) {
    return guarded([&] {
        REQUIRE(m && n_out, "null argument");
        std::lock_guard<std::mutex> lk(m->aux_mu);
        *n_out = (uint32_t)m->stats.size();
        for (uint32_t i = 0; i < cap && i < m->stats.size(); ++i) {
            memset(&out[i], 0, sizeof(out[i]));
            snprintf(out[i].name, sizeof(out[i].name), "%s", m->stats[i].name.c_str());
            out[i].launches = m->stats[i].launches;
            out[i].device_ms = m->stats[i].ms;
            out[i].algorithmic_bytes = m->stats[i].alg_bytes;
            out[i].compulsory_bytes = m->stats[i].min_bytes;
            out[i].flops = m->stats[i].flops;
        }
    });
}

int uf_debug_fail_after(uf_model* m, int32_t stages) {
    return guarded([&] {
        REQUIRE(m, "null model");
        m->fail_after_stages = stages;
    });
}

int uf_launch_count(const uf_model* m, uint64_t* n) {
    return guarded([&] {
        REQUIRE(m && n, "null argument");
        *n = m->launches.load();
    });
}

int uf_onnx_inspect(const char* onnx_path, uint32_t net_w, uint32_t net_h, char* out, size_t cap, size_t* needed) {
    return guarded([&] {
        REQUIRE(onnx_path && needed, "null argument");
        OnnxModel om = load_onnx_file(onnx_path);
        Plan p = lower_ultraface(om, (int)net_w, (int)net_h);
        std::string j = "{";
        auto kv = [&](const std::string& k, const std::string& v, bool last = false) { j += "\"" + k + "\": " + v + (last ? "" : ", "); };
        kv("num_priors", std::to_string(p.num_priors));
        kv("macs_per_frame", std::to_string(p.macs_per_frame));
        kv("conv_bytes_per_frame", std::to_string(p.conv_bytes_per_frame));
        kv("arena_frame_floats", std::to_string(p.arena_frame_floats));
        kv("priors_from_graph", p.priors_from_graph ? "true" : "false");
        kv("center_variance", std::to_string(p.center_variance));
        kv("size_variance", std::to_string(p.size_variance));
        double ps = 0;
        for (float v : p.priors) ps += v;
        kv("priors_sum", std::to_string(ps));
        std::string w = p.warnings;
        for (auto& c : w) if (c == '"') c = '\'';
        kv("warnings", "\"" + w + "\"");
        std::string heads = "[";
        for (size_t i = 0; i < p.heads.size(); ++i) {
            const Head& h = p.heads[i];
            heads += std::string(i ? ", " : "") + "{\"fm_w\": " + std::to_string(h.fm_w) + ", \"fm_h\": " + std::to_string(h.fm_h) +
                     ", \"anchors\": " + std::to_string(h.anchors) + ", \"prior_off\": " + std::to_string(h.prior_off) + "}";
        }
        kv("heads", heads + "]");
        std::string ops = "[";
        for (size_t i = 0; i < p.ops.size(); ++i) {
            const Op& o = p.ops[i];
            double ws = 0, wa = 0, bs = 0;
            for (float v : o.w) { ws += v; wa += std::fabs(v); }
            for (float v : o.b) bs += v;
            const TensorDesc& t = p.tensors[o.out];
            char buf[640];
            snprintf(buf, sizeof(buf),
                     "%s{\"kind\": %d, \"out\": \"%s\", \"cin\": %d, \"cout\": %d, \"k\": %d, \"stride\": %d, \"pad\": %d, "
                     "\"dil\": %d, \"groups\": %d, \"relu\": %d, \"residual\": %d, \"h\": %d, \"w\": %d, \"pix_stride\": %d, "
                     "\"base_off\": %lld, \"w_sum\": %.9g, \"w_abs\": %.9g, \"b_sum\": %.9g}",
                     i ? ", " : "", (int)o.kind, t.name.c_str(), o.cin, o.cout, o.k, o.stride, o.pad, o.dil, o.groups, o.relu ? 1 : 0,
                     o.in2 >= 0 ? 1 : 0, t.H, t.W, t.pix_stride, (long long)t.base_off, ws, wa, bs);
            ops += buf;
        }
        kv("ops", ops + "]", true);
        j += "}";
        *needed = j.size() + 1;
        if (out && cap) {
            size_t n = std::min(cap - 1, j.size());
            memcpy(out, j.data(), n);
            out[n] = 0;
        }
    });
}

static void fill_jpeg_info(const JpegPlan& p, uf_jpeg_info* o) {
    memset(o, 0, sizeof(*o));
    o->w = p.w; o->h = p.h; o->ncomp = p.ncomp; o->nblocks = p.nblocks;
    for (uint32_t c = 0; c < p.ncomp; ++c) { o->hs[c] = p.hs[c]; o->vs[c] = p.vs[c]; }
}

int uf_jpeg_info_read(const uint8_t* jpeg, size_t len, uf_jpeg_info* out) {
    return guarded([&] {
        REQUIRE(jpeg && out, "null argument");
        fill_jpeg_info(jpeg_parse_header(jpeg, len), out);
    });
}

int uf_jpeg_coefficients(const uint8_t* jpeg, size_t len, uf_jpeg_info* info, int16_t* coefs, size_t cap_blocks) {
    return guarded([&] {
        REQUIRE(jpeg && info, "null argument");
        JpegCoefs jc;
        jpeg_entropy_decode(jpeg, len, jc);
        fill_jpeg_info(jc.plan, info);
        info->nonzero = (uint32_t)jc.entries.size();
        memcpy(info->quant, jc.plan.quant, sizeof(info->quant));
        if (!coefs) return;
        if (cap_blocks < jc.plan.nblocks) throw ArgError(UF_ERR_CAPACITY, "coefficient buffer smaller than nblocks");
        memset(coefs, 0, (size_t)jc.plan.nblocks * 64 * sizeof(int16_t));
        for (uint32_t b = 0; b < jc.plan.nblocks; ++b)
            for (uint32_t e = jc.block_off[b]; e < jc.block_off[b + 1]; ++e)
                coefs[(size_t)b * 64 + (jc.entries[e] >> 16)] = (int16_t)(jc.entries[e] & 0xffffu);
    });
}

int uf_resize_taps(uint32_t src_len, uint32_t dst_len, int32_t* left, int32_t* ntaps, float* w, uint32_t w_pitch,
                   uint32_t* max_taps) {
    return guarded([&] {
        REQUIRE(src_len > 0 && dst_len > 0 && max_taps, "bad argument");
        AxisTaps t = build_axis_taps((int)src_len, (int)dst_len);
        *max_taps = (uint32_t)t.max_taps;
        if (left && ntaps && w) {
            REQUIRE(w_pitch >= (uint32_t)t.max_taps, "w_pitch smaller than max taps");
            for (uint32_t o = 0; o < dst_len; ++o) {
                left[o] = t.left[o];
                ntaps[o] = t.ntaps[o];
                for (uint32_t i = 0; i < w_pitch; ++i) w[(size_t)o * w_pitch + i] = i < (uint32_t)t.max_taps ? t.w[(size_t)o * t.max_taps + i] : 0.f;
            }
        }
    });
}

const char* uf_last_error(void) { return g_last_error.c_str(); }
const char* uf_version(void) { return "ultraface_b200 0.1 (sm_100a)"; }

int uf_device_count(int32_t* n) {
    return guarded([&] {
        REQUIRE(n, "null argument");
        int c = 0;
        if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); c = 0; }
        *n = c;
    });
}

}  // extern "C"
