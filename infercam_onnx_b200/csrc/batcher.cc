// batcher.cc — stream batcher and stream -> GPU router behind the C ABI (uf_batcher_*), plus the ingest helpers
// uf_stream_hash / uf_protomsg_parse. SURVEY.md §8f rows N1 and N4.
//
// What it replaces in the reference (/root/reference/infer_server/src): the single inference task that handles one
// frame at a time (`Inferer::run`, inferer.rs:29-50), the bounded lossy queue in front of it (`INFER_IMAGES_CHANNEL`,
// lib.rs:32-37, filled with `try_send_ref` in router.rs:64-72) and the `hashed(&id)` stream key (lib.rs:39-46,
// router.rs:58) — which here also picks the GPU. Host plumbing only: no arithmetic on pixels happens in this file; the
// batches go through the same entry point an external caller would use (uf_infer_batch).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ultraface_b200.h"

namespace uf {
void set_last_error(const std::string& msg);  // engine.cu (thread-local message behind uf_last_error)
}

namespace {

using Clock = std::chrono::steady_clock;

struct Ticket {            // one pinned slot
    uint8_t* buf = nullptr;
    uint64_t stream = 0, tag = 0;
    uint32_t w = 0, h = 0;
    size_t jpeg_len = 0;   // > 0: the slot holds a JPEG file of that many bytes, not RGB8 pixels (N2)
    Clock::time_point t_commit;
    int state = 0;         // 0 free, 1 acquired, 2 queued, 3 in a batch
};

struct Done {
    uf_result res;
    std::vector<uf_det> dets;
    std::vector<uint8_t> file;  // annotate mode: the annotated JPEG
};

struct Device {
    int32_t ordinal = 0;
    uf_model* model = nullptr;
    bool pinned = true;
    uint8_t* pool = nullptr;           // pinned, nslots * slot_bytes
    std::vector<Ticket> slots;
    std::vector<uint32_t> free_slots;
    std::deque<uint32_t> queue;        // committed slots, submission order
    std::mutex mu;
    std::condition_variable cv;
    uint64_t next_batch = 0;           // sequence number of the next batch formed
    uint64_t next_deliver = 0;         // sequence number allowed to publish its results
    std::condition_variable deliver_cv;
    std::vector<std::thread> workers;
};

}  // namespace

struct uf_batcher {
    uf_batcher_config cfg{};
    uf_batch_fn backend = nullptr;  // NULL: uf_infer_batch on the device's own handle
    void* backend_user = nullptr;
    std::string onnx_path;
    std::vector<std::unique_ptr<Device>> devs;
    size_t slot_bytes = 0;
    uint32_t nslots = 0;
    std::atomic<bool> closing{false};
    // results
    std::mutex done_mu;
    std::condition_variable done_cv;
    std::deque<Done> done;
    std::atomic<uint64_t> submitted{0}, dropped{0}, completed{0}, failed{0}, batches{0}, inflight{0};
};

namespace {

inline uint64_t ticket_of(uint32_t dev, uint32_t slot) { return ((uint64_t)dev << 32) | slot; }

void worker_loop(uf_batcher* b, Device* d) {
    const uint32_t max_batch = b->cfg.max_batch, det_cap = b->cfg.det_cap;
    std::vector<uint32_t> take;
    std::vector<const uint8_t*> ptrs;
    std::vector<uint32_t> ws, hs, counts;
    std::vector<size_t> lens;
    std::vector<uf_det> dets;
    // annotate mode (frames submitted as JPEG): the whole loop body of inferer.rs:35-46 in one call per batch
    const bool annotate = !b->backend && b->cfg.annotate_quality > 0;
    const size_t file_stride = b->cfg.annotate_max_bytes;
    std::vector<uint8_t> files;
    std::vector<size_t> file_lens;
    auto jpeg_call = [&](uint32_t first, uint32_t cnt) -> int {
        if (!annotate)
            return uf_infer_batch_jpeg(d->model, ptrs.data() + first, lens.data() + first, cnt, dets.data() + (size_t)first * det_cap, det_cap,
                                       counts.data() + first);
        return uf_worker_batch_jpeg(d->model, ptrs.data() + first, lens.data() + first, cnt, b->cfg.annotate_scale_w, b->cfg.annotate_scale_h,
                                    b->cfg.annotate_quality, dets.data() + (size_t)first * det_cap, det_cap, counts.data() + first,
                                    files.data() + (size_t)first * file_stride, file_stride, file_lens.data() + first);
    };
    if (!b->backend) cudaSetDevice(d->ordinal);
    for (;;) {
        uint64_t seq;
        {
            std::unique_lock<std::mutex> lk(d->mu);
            d->cv.wait(lk, [&] { return !d->queue.empty() || b->closing.load(); });
            if (d->queue.empty()) return;  // closing and drained
            // a batch is worth waiting for, but not for long: a lone stream must not be held back
            const auto deadline = Clock::now() + std::chrono::microseconds(b->cfg.max_delay_us);
            while (d->queue.size() < max_batch && !b->closing.load())
                if (d->cv.wait_until(lk, deadline) == std::cv_status::timeout) break;
            take.clear();
            while (!d->queue.empty() && take.size() < max_batch) {
                take.push_back(d->queue.front());
                d->queue.pop_front();
                d->slots[take.back()].state = 3;
            }
            seq = d->next_batch++;
        }
        const uint32_t n = (uint32_t)take.size();
        // RGB frames first, JPEG frames after them (stable), so that each kind goes through its own batched entry point;
        // results are published per frame, so the order inside the batch does not matter
        std::stable_partition(take.begin(), take.end(), [&](uint32_t s) { return d->slots[s].jpeg_len == 0; });
        uint32_t n_rgb = 0;
        while (n_rgb < n && d->slots[take[n_rgb]].jpeg_len == 0) ++n_rgb;
        ptrs.resize(n); ws.resize(n); hs.resize(n); counts.assign(n, 0);
        lens.resize(n);
        dets.resize((size_t)n * det_cap);
        for (uint32_t i = 0; i < n; ++i) {
            const Ticket& t = d->slots[take[i]];
            ptrs[i] = t.buf; ws[i] = t.w; hs[i] = t.h; lens[i] = t.jpeg_len;
        }
        int rc_rgb = UF_OK, rc_jpeg = UF_OK;
        if (b->backend) {
            rc_rgb = rc_jpeg = b->backend(b->backend_user, d->ordinal, ptrs.data(), ws.data(), hs.data(), n, dets.data(), det_cap, counts.data());
        } else {
            if (n_rgb) rc_rgb = uf_infer_batch(d->model, ptrs.data(), ws.data(), hs.data(), n_rgb, dets.data(), det_cap, counts.data());
            if (annotate) {
                files.resize((size_t)n * file_stride);
                file_lens.assign(n, 0);
            }
            if (n > n_rgb) rc_jpeg = jpeg_call(n_rgb, n - n_rgb);
        }
        // one undecodable file fails the whole JPEG call: isolate it so that its batch-mates are not skipped with it
        std::vector<int> rc_each;
        if (!b->backend && rc_jpeg != UF_OK && n - n_rgb > 1) {
            rc_each.assign(n, UF_OK);
            for (uint32_t i = n_rgb; i < n; ++i) rc_each[i] = jpeg_call(i, 1);
        }
        const auto now = Clock::now();
        std::vector<Done> out(n);
        for (uint32_t i = 0; i < n; ++i) {
            const Ticket& t = d->slots[take[i]];
            Done& o = out[i];
            const int rc = i < n_rgb ? rc_rgb : rc_each.empty() ? rc_jpeg : rc_each[i];
            o.res.stream = t.stream; o.res.user_tag = t.tag; o.res.device = d->ordinal; o.res.status = rc;
            o.res.n_dets = rc == UF_OK ? counts[i] : 0; o.res.batch_size = n;
            o.res.latency_us = (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(now - t.t_commit).count();
            if (rc == UF_OK) {
                const uint32_t k = counts[i] < det_cap ? counts[i] : det_cap;
                o.dets.assign(dets.begin() + (size_t)i * det_cap, dets.begin() + (size_t)i * det_cap + k);
                if (annotate && i >= n_rgb) {
                    o.file.assign(files.begin() + (size_t)i * file_stride, files.begin() + (size_t)i * file_stride + file_lens[i]);
                    o.res.file_bytes = (uint32_t)file_lens[i];
                }
            }
        }
        {
            // publish batches in the order they were formed: a stream lives on one device, so this keeps every
            // stream's results in submission order even with several batches in flight
            std::unique_lock<std::mutex> lk(d->mu);
            d->deliver_cv.wait(lk, [&] { return d->next_deliver == seq; });
            {
                std::lock_guard<std::mutex> dl(b->done_mu);
                for (auto& o : out) b->done.push_back(std::move(o));
            }
            for (uint32_t s : take) {
                d->slots[s].state = 0;
                d->free_slots.push_back(s);
            }
            d->next_deliver++;
            d->deliver_cv.notify_all();
        }
        b->batches++;
        (rc_rgb == UF_OK ? b->completed : b->failed) += n_rgb;
        if (rc_each.empty()) {
            (rc_jpeg == UF_OK ? b->completed : b->failed) += n - n_rgb;
        } else {
            for (uint32_t i = n_rgb; i < n; ++i) (rc_each[i] == UF_OK ? b->completed : b->failed) += 1;
        }
        b->inflight -= n;
        b->done_cv.notify_all();
    }
}

struct Fail {
    int code;
    std::string msg;
};

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return UF_OK;
    } catch (const Fail& e) {
        uf::set_last_error(e.msg);
        return e.code;
    } catch (const std::exception& e) {
        uf::set_last_error(e.what());
        return UF_ERR_INVALID_ARG;
    }
}

#define NEED(cond, msg) do { if (!(cond)) throw Fail{UF_ERR_INVALID_ARG, msg}; } while (0)

// SipHash-c-d (Aumasson & Bernstein), streaming over one buffer plus a trailing byte; Rust's DefaultHasher is c=1, d=3
inline uint64_t rotl(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }

uint64_t siphash(int c_rounds, int d_rounds, uint64_t k0, uint64_t k1, const uint8_t* in, size_t len) {
    uint64_t v0 = k0 ^ 0x736f6d6570736575ull, v1 = k1 ^ 0x646f72616e646f6dull, v2 = k0 ^ 0x6c7967656e657261ull,
             v3 = k1 ^ 0x7465646279746573ull;
    auto round = [&] {
        v0 += v1; v1 = rotl(v1, 13); v1 ^= v0; v0 = rotl(v0, 32);
        v2 += v3; v3 = rotl(v3, 16); v3 ^= v2;
        v0 += v3; v3 = rotl(v3, 21); v3 ^= v0;
        v2 += v1; v1 = rotl(v1, 17); v1 ^= v2; v2 = rotl(v2, 32);
    };
    const size_t full = len / 8 * 8;
    for (size_t i = 0; i < full; i += 8) {
        uint64_t m = 0;
        for (int k = 0; k < 8; ++k) m |= (uint64_t)in[i + k] << (8 * k);
        v3 ^= m;
        for (int r = 0; r < c_rounds; ++r) round();
        v0 ^= m;
    }
    uint64_t last = (uint64_t)(len & 0xff) << 56;
    for (size_t k = 0; k < len - full; ++k) last |= (uint64_t)in[full + k] << (8 * k);
    v3 ^= last;
    for (int r = 0; r < c_rounds; ++r) round();
    v0 ^= last;
    v2 ^= 0xff;
    for (int r = 0; r < d_rounds; ++r) round();
    return v0 ^ v1 ^ v2 ^ v3;
}

}  // namespace

extern "C" {

static void free_device(Device& d) {
    if (d.model) uf_model_free(d.model);
    if (d.pinned) uf_host_free(d.pool); else free(d.pool);
}

int uf_batcher_create(const uf_batcher_config* cfg, uf_batcher** out) { return uf_batcher_create_ex(cfg, nullptr, nullptr, out); }

int uf_batcher_create_ex(const uf_batcher_config* cfg, uf_batch_fn backend, void* user, uf_batcher** out) {
    return guarded([&] {
        NEED(cfg && out, "null argument");
        NEED(cfg->struct_size == sizeof(uf_batcher_config), "uf_batcher_config.struct_size mismatch");
        NEED(backend || (cfg->model.struct_size == sizeof(uf_config) && cfg->model.onnx_path), "uf_batcher_config.model is not a filled uf_config");
        *out = nullptr;
        std::unique_ptr<uf_batcher> b(new uf_batcher());
        b->cfg = *cfg;
        b->backend = backend;
        b->backend_user = user;
        b->onnx_path = cfg->model.onnx_path ? cfg->model.onnx_path : "";
        b->cfg.model.onnx_path = b->onnx_path.c_str();
        uf_batcher_config& c = b->cfg;
        if (c.max_batch == 0) c.max_batch = 64;
        if (c.max_delay_us == 0) c.max_delay_us = 2000;
        if (c.capacity == 0) c.capacity = 2 * c.max_batch;
        if (c.workers == 0) c.workers = 2;
        if (c.det_cap == 0) c.det_cap = 64;
        if (c.max_frame_bytes == 0) c.max_frame_bytes = 1280 * 720 * 3;
        if (c.annotate_scale_w == 0.0f) c.annotate_scale_w = 1280.0f;
        if (c.annotate_scale_h == 0.0f) c.annotate_scale_h = 720.0f;
        if (c.annotate_max_bytes == 0) c.annotate_max_bytes = 1u << 20;
        NEED(c.annotate_quality <= 100 && c.annotate_max_bytes >= 1024, "annotate_quality / annotate_max_bytes out of range");
        NEED(c.workers <= 8 && c.max_batch <= 4096 && c.capacity <= (1u << 20), "workers / max_batch / capacity out of range");
        std::vector<int32_t> ords(cfg->devices ? cfg->devices : nullptr, cfg->devices ? cfg->devices + cfg->n_devices : nullptr);
        if (ords.empty()) ords.push_back(0);
        NEED(ords.size() <= 64, "too many devices");
        b->slot_bytes = ((size_t)c.max_frame_bytes + 255) / 256 * 256;
        b->nslots = c.capacity + c.workers * c.max_batch;  // queued + riding in a batch
        for (size_t i = 0; i < ords.size(); ++i) {
            std::unique_ptr<Device> d(new Device());
            d->ordinal = ords[i];
            void* p = nullptr;
            if (backend) {  // injected backend (tests): no handle, no CUDA, pageable pool
                d->pinned = false;
                p = malloc(b->slot_bytes * b->nslots);
                if (!p) {
                    for (auto& e : b->devs) free_device(*e);
                    throw Fail{UF_ERR_INVALID_ARG, "out of host memory for the frame pool"};
                }
            } else {
                uf_config mc = c.model;
                mc.device = ords[i];
                if (mc.max_batch < c.max_batch) mc.max_batch = c.max_batch;
                if (mc.lanes == 0) mc.lanes = c.workers;
                int rc = uf_model_load_ex(&mc, &d->model);
                if (rc != UF_OK) {
                    for (auto& e : b->devs) free_device(*e);
                    throw Fail{rc, std::string("device ") + std::to_string(ords[i]) + ": " + uf_last_error()};
                }
                cudaSetDevice(ords[i]);
                rc = uf_host_alloc(b->slot_bytes * b->nslots, &p);
                if (rc != UF_OK) {
                    uf_model_free(d->model);
                    for (auto& e : b->devs) free_device(*e);
                    throw Fail{rc, std::string("pinned frame pool: ") + uf_last_error()};
                }
            }
            d->pool = (uint8_t*)p;
            d->slots.resize(b->nslots);
            for (uint32_t s = 0; s < b->nslots; ++s) {
                d->slots[s].buf = d->pool + (size_t)s * b->slot_bytes;
                d->free_slots.push_back(b->nslots - 1 - s);
            }
            b->devs.push_back(std::move(d));
        }
        for (auto& d : b->devs)
            for (uint32_t w = 0; w < c.workers; ++w) d->workers.emplace_back(worker_loop, b.get(), d.get());
        *out = b.release();
    });
}

void uf_batcher_destroy(uf_batcher* b) {
    if (!b) return;
    b->closing = true;
    for (auto& d : b->devs) {
        { std::lock_guard<std::mutex> lk(d->mu); }
        d->cv.notify_all();
    }
    for (auto& d : b->devs)
        for (auto& t : d->workers) t.join();
    for (auto& d : b->devs) free_device(*d);
    delete b;
}

int uf_batcher_acquire(uf_batcher* b, uint64_t stream, size_t bytes, uint8_t** buf, uint64_t* ticket) {
    return guarded([&] {
        NEED(b && buf && ticket, "null argument");
        *buf = nullptr;
        *ticket = 0;
        NEED(!b->closing.load(), "batcher is closing");
        if (bytes > b->slot_bytes) throw Fail{UF_ERR_CAPACITY, "frame larger than uf_batcher_config.max_frame_bytes"};
        const uint32_t di = (uint32_t)(stream % b->devs.size());
        Device& d = *b->devs[di];
        std::lock_guard<std::mutex> lk(d.mu);
        // lossy, like try_send_ref on a full INFER_IMAGES_CHANNEL (router.rs:64-72): the frame is dropped, the caller goes on
        if (d.queue.size() >= b->cfg.capacity || d.free_slots.empty()) {
            b->dropped++;
            return;
        }
        const uint32_t s = d.free_slots.back();
        d.free_slots.pop_back();
        Ticket& t = d.slots[s];
        t.state = 1;
        t.stream = stream;
        *buf = t.buf;
        *ticket = ticket_of(di, s);
    });
}

static Ticket& ticket_ref(uf_batcher* b, uint64_t ticket, Device** dev) {
    const uint32_t di = (uint32_t)(ticket >> 32), s = (uint32_t)ticket;
    if (di >= b->devs.size() || s >= b->nslots) throw Fail{UF_ERR_INVALID_ARG, "bad ticket"};
    *dev = b->devs[di].get();
    return (*dev)->slots[s];
}

int uf_batcher_commit(uf_batcher* b, uint64_t ticket, uint32_t w, uint32_t h, uint64_t user_tag) {
    return guarded([&] {
        NEED(b, "null argument");
        Device* d = nullptr;
        Ticket& t = ticket_ref(b, ticket, &d);
        NEED(w > 0 && h > 0 && (size_t)w * h * 3 <= b->slot_bytes, "frame size does not fit the slot");
        {
            std::lock_guard<std::mutex> lk(d->mu);
            NEED(t.state == 1, "ticket is not in the acquired state");
            t.w = w; t.h = h; t.jpeg_len = 0; t.tag = user_tag; t.t_commit = Clock::now(); t.state = 2;
            d->queue.push_back((uint32_t)ticket);
            b->submitted++;
            b->inflight++;
        }
        d->cv.notify_one();
    });
}

int uf_batcher_abort(uf_batcher* b, uint64_t ticket) {
    return guarded([&] {
        NEED(b, "null argument");
        Device* d = nullptr;
        Ticket& t = ticket_ref(b, ticket, &d);
        std::lock_guard<std::mutex> lk(d->mu);
        NEED(t.state == 1, "ticket is not in the acquired state");
        t.state = 0;
        d->free_slots.push_back((uint32_t)ticket);
    });
}

int uf_batcher_try_submit(uf_batcher* b, uint64_t stream, const uint8_t* rgb, uint32_t w, uint32_t h, uint64_t user_tag,
                          int32_t* accepted) {
    if (accepted) *accepted = 0;
    if (!b || !rgb || !accepted || w == 0 || h == 0) {
        uf::set_last_error("bad argument");
        return UF_ERR_INVALID_ARG;
    }
    uint8_t* buf = nullptr;
    uint64_t ticket = 0;
    int rc = uf_batcher_acquire(b, stream, (size_t)w * h * 3, &buf, &ticket);
    if (rc != UF_OK || !buf) return rc;
    memcpy(buf, rgb, (size_t)w * h * 3);
    rc = uf_batcher_commit(b, ticket, w, h, user_tag);
    if (rc == UF_OK) *accepted = 1;
    else uf_batcher_abort(b, ticket);
    return rc;
}

int uf_batcher_commit_jpeg(uf_batcher* b, uint64_t ticket, size_t jpeg_len, uint64_t user_tag) {
    return guarded([&] {
        NEED(b, "null argument");
        Device* d = nullptr;
        Ticket& t = ticket_ref(b, ticket, &d);
        NEED(jpeg_len >= 4 && jpeg_len <= b->slot_bytes, "JPEG length does not fit the slot");
        {
            std::lock_guard<std::mutex> lk(d->mu);
            NEED(t.state == 1, "ticket is not in the acquired state");
            t.w = t.h = 0; t.jpeg_len = jpeg_len; t.tag = user_tag; t.t_commit = Clock::now(); t.state = 2;
            d->queue.push_back((uint32_t)ticket);
            b->submitted++;
            b->inflight++;
        }
        d->cv.notify_one();
    });
}

int uf_batcher_try_submit_jpeg(uf_batcher* b, uint64_t stream, const uint8_t* jpeg, size_t len, uint64_t user_tag, int32_t* accepted) {
    if (accepted) *accepted = 0;
    if (!b || !jpeg || !accepted || len < 4) {
        uf::set_last_error("bad argument");
        return UF_ERR_INVALID_ARG;
    }
    uint8_t* buf = nullptr;
    uint64_t ticket = 0;
    int rc = uf_batcher_acquire(b, stream, len, &buf, &ticket);
    if (rc != UF_OK || !buf) return rc;
    memcpy(buf, jpeg, len);
    rc = uf_batcher_commit_jpeg(b, ticket, len, user_tag);
    if (rc == UF_OK) *accepted = 1;
    else uf_batcher_abort(b, ticket);
    return rc;
}

int uf_batcher_ingest(uf_batcher* b, const uint8_t* msg, size_t len, uint64_t user_tag, int32_t* accepted, uint64_t* stream) {
    if (accepted) *accepted = 0;
    uint32_t kind = 0;
    const uint8_t *id = nullptr, *data = nullptr;
    size_t id_len = 0, data_len = 0;
    int rc = uf_protomsg_parse(msg, len, &kind, &id, &id_len, &data, &data_len);
    if (rc != UF_OK) return rc;
    uint64_t key = 0;
    rc = uf_stream_hash(id, id_len, &key);
    if (rc != UF_OK) return rc;
    if (stream) *stream = key;
    if (kind != 1) return UF_OK;  // ConnectReq: nothing to infer
    return uf_batcher_try_submit_jpeg(b, key, data, data_len, user_tag, accepted);
}

int uf_batcher_poll(uf_batcher* b, uf_result* res, uf_det* dets, uint32_t cap, uint32_t timeout_ms, uint32_t* n_out) {
    return guarded([&] {
        NEED(b && n_out && (cap == 0 || (res && dets)), "null argument");
        *n_out = 0;
        std::unique_lock<std::mutex> lk(b->done_mu);
        if (b->done.empty() && timeout_ms)
            b->done_cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return !b->done.empty(); });
        uint32_t n = 0;
        while (n < cap && !b->done.empty()) {
            Done& o = b->done.front();
            res[n] = o.res;
            if (!o.dets.empty()) memcpy(dets + (size_t)n * b->cfg.det_cap, o.dets.data(), o.dets.size() * sizeof(uf_det));
            b->done.pop_front();
            ++n;
        }
        *n_out = n;
    });
}

int uf_batcher_poll_frames(uf_batcher* b, uf_result* res, uf_det* dets, uint8_t* files, size_t file_stride, uint32_t cap, uint32_t timeout_ms,
                           uint32_t* n_out) {
    return guarded([&] {
        NEED(b && n_out && (cap == 0 || (res && dets && files)), "null argument");
        NEED(file_stride >= b->cfg.annotate_max_bytes, "file_stride smaller than uf_batcher_config.annotate_max_bytes");
        *n_out = 0;
        std::unique_lock<std::mutex> lk(b->done_mu);
        if (b->done.empty() && timeout_ms)
            b->done_cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return !b->done.empty(); });
        uint32_t n = 0;
        while (n < cap && !b->done.empty()) {
            Done& o = b->done.front();
            res[n] = o.res;
            if (!o.dets.empty()) memcpy(dets + (size_t)n * b->cfg.det_cap, o.dets.data(), o.dets.size() * sizeof(uf_det));
            if (!o.file.empty()) memcpy(files + (size_t)n * file_stride, o.file.data(), o.file.size());
            b->done.pop_front();
            ++n;
        }
        *n_out = n;
    });
}

int uf_batcher_flush(uf_batcher* b, uint32_t timeout_ms) {
    return guarded([&] {
        NEED(b, "null argument");
        std::unique_lock<std::mutex> lk(b->done_mu);
        const bool ok = b->done_cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return b->inflight.load() == 0; });
        if (!ok) throw Fail{UF_ERR_CAPACITY, "uf_batcher_flush timed out with frames still in flight"};
    });
}

int uf_batcher_stats_read(const uf_batcher* b, uf_batcher_stats* out) {
    return guarded([&] {
        NEED(b && out, "null argument");
        out->submitted = b->submitted.load(); out->dropped = b->dropped.load(); out->completed = b->completed.load();
        out->failed = b->failed.load(); out->batches = b->batches.load();
    });
}

int uf_batcher_owner(const uf_batcher* b, uint64_t stream, int32_t* device) {
    return guarded([&] {
        NEED(b && device, "null argument");
        *device = b->devs[stream % b->devs.size()]->ordinal;
    });
}

int uf_batcher_model(uf_batcher* b, uint32_t device_slot, uf_model** out) {
    return guarded([&] {
        NEED(b && out && device_slot < b->devs.size(), "bad argument");
        NEED(b->devs[device_slot]->model, "this batcher runs an injected backend, it owns no model handle");
        *out = b->devs[device_slot]->model;
    });
}

// Measurement aid: drives the batcher from C++ — `producers` threads submit `total` frames (frame i = frames + (i % n_frames)
// * w * h * 3, stream = streams[i % n_streams]) with the lossy try_submit, retrying dropped frames so that every frame is
// counted, while the calling thread polls — and reports the wall time. What a Rust ingest task would do, minus Rust.
int uf_debug_batcher_drive(uf_batcher* b, const uint8_t* frames, uint32_t n_frames, uint32_t w, uint32_t h, const uint64_t* streams,
                           uint32_t n_streams, uint64_t total, uint32_t producers, double* seconds, uint64_t* detections) {
    return guarded([&] {
        NEED(b && frames && n_frames && w && h && streams && n_streams && producers && producers <= 64 && seconds, "bad argument");
        const size_t fb = (size_t)w * h * 3;
        std::atomic<uint64_t> next{0}, errors{0};
        const auto t0 = Clock::now();
        std::vector<std::thread> ts;
        for (uint32_t p = 0; p < producers; ++p)
            ts.emplace_back([&] {
                for (;;) {
                    const uint64_t i = next.fetch_add(1);
                    if (i >= total) return;
                    for (;;) {
                        int32_t ok = 0;
                        if (uf_batcher_try_submit(b, streams[i % n_streams], frames + (i % n_frames) * fb, w, h, i, &ok) != UF_OK) { errors++; break; }
                        if (ok) break;
                        std::this_thread::sleep_for(std::chrono::microseconds(50));
                    }
                }
            });
        std::vector<uf_result> res(1024);
        std::vector<uf_det> dets((size_t)1024 * b->cfg.det_cap);
        uint64_t got = 0, ndet = 0;
        while (got + errors.load() < total) {
            uint32_t n = 0;
            if (uf_batcher_poll(b, res.data(), dets.data(), 1024, 20, &n) != UF_OK) break;
            for (uint32_t k = 0; k < n; ++k) ndet += res[k].n_dets;
            got += n;
        }
        for (auto& t : ts) t.join();
        *seconds = std::chrono::duration<double>(Clock::now() - t0).count();
        if (detections) *detections = ndet;
        if (errors.load()) throw Fail{UF_ERR_INVALID_ARG, "uf_batcher_try_submit failed for " + std::to_string(errors.load()) + " frames"};
    });
}

int uf_debug_batcher_drive_msgs(uf_batcher* b, const uint8_t* const* msgs, const size_t* lens, uint32_t n_msgs, uint64_t total,
                                uint32_t producers, double* seconds, uint64_t* detections) {
    return guarded([&] {
        NEED(b && msgs && lens && n_msgs && producers && producers <= 64 && seconds, "bad argument");
        std::atomic<uint64_t> next{0}, errors{0};
        const auto t0 = Clock::now();
        std::vector<std::thread> ts;
        for (uint32_t p = 0; p < producers; ++p)
            ts.emplace_back([&] {
                for (;;) {
                    const uint64_t i = next.fetch_add(1);
                    if (i >= total) return;
                    for (;;) {
                        int32_t ok = 0;
                        if (uf_batcher_ingest(b, msgs[i % n_msgs], lens[i % n_msgs], i, &ok, nullptr) != UF_OK) { errors++; break; }
                        if (ok) break;
                        std::this_thread::sleep_for(std::chrono::microseconds(50));
                    }
                }
            });
        std::vector<uf_result> res(1024);
        std::vector<uf_det> dets((size_t)1024 * b->cfg.det_cap);
        uint64_t got = 0, ndet = 0, failed = 0;
        while (got + errors.load() < total) {
            uint32_t n = 0;
            if (uf_batcher_poll(b, res.data(), dets.data(), 1024, 20, &n) != UF_OK) break;
            for (uint32_t k = 0; k < n; ++k) {
                ndet += res[k].n_dets;
                failed += res[k].status != 0;
            }
            got += n;
        }
        for (auto& t : ts) t.join();
        *seconds = std::chrono::duration<double>(Clock::now() - t0).count();
        if (detections) *detections = ndet;
        if (errors.load()) throw Fail{UF_ERR_INVALID_ARG, "uf_batcher_ingest failed for " + std::to_string(errors.load()) + " messages"};
        if (failed) throw Fail{UF_ERR_INVALID_ARG, std::to_string(failed) + " frames came back with an error status"};
    });
}

int uf_stream_hash(const uint8_t* name, size_t len, uint64_t* out) {
    return guarded([&] {
        NEED(out && (name || len == 0), "null argument");
        // `impl Hash for str`: state.write(bytes); state.write_u8(0xff)  — SipHash is a byte-stream hash, so this is the
        // hash of name || 0xff
        std::vector<uint8_t> buf(name, name + len);
        buf.push_back(0xff);
        *out = siphash(1, 3, 0, 0, buf.data(), buf.size());
    });
}

// test hook: the generic SipHash-c-d with explicit keys (checked against the published SipHash-2-4 vectors)
int uf_debug_siphash(uint32_t c, uint32_t d, uint64_t k0, uint64_t k1, const uint8_t* in, size_t len, uint64_t* out) {
    return guarded([&] {
        NEED(out && (in || len == 0) && c >= 1 && c <= 8 && d >= 1 && d <= 8, "bad argument");
        *out = siphash((int)c, (int)d, k0, k1, in, len);
    });
}

int uf_protomsg_parse(const uint8_t* msg, size_t len, uint32_t* kind, const uint8_t** id, size_t* id_len,
                      const uint8_t** data, size_t* data_len) {
    return guarded([&] {
        NEED(msg && kind && id && id_len && data && data_len, "null argument");
        *id = *data = nullptr;
        *id_len = *data_len = 0;
        size_t pos = 0;
        auto u = [&](int bytes) -> uint64_t {
            if (len - pos < (size_t)bytes) throw Fail{UF_ERR_INVALID_ARG, "ProtoMsg: truncated"};
            uint64_t v = 0;
            for (int k = 0; k < bytes; ++k) v |= (uint64_t)msg[pos + k] << (8 * k);
            pos += bytes;
            return v;
        };
        auto blob = [&](const uint8_t** p, size_t* n) {
            const uint64_t l = u(8);
            if (l > len - pos) throw Fail{UF_ERR_INVALID_ARG, "ProtoMsg: length prefix exceeds the message"};
            *p = msg + pos;
            *n = (size_t)l;
            pos += (size_t)l;
        };
        const uint64_t variant = u(4);
        if (variant > 1) throw Fail{UF_ERR_INVALID_ARG, "ProtoMsg: unknown enum variant"};
        *kind = (uint32_t)variant;
        blob(id, id_len);
        if (variant == 1) blob(data, data_len);
        if (pos != len) throw Fail{UF_ERR_INVALID_ARG, "ProtoMsg: trailing bytes"};  // bincode::deserialize tolerates them; a frame never has any
    });
}

}  // extern "C"
