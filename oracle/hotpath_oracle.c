/*
 * hotpath_oracle.c — CPU restatement of the reference's face-detection hot path
 * (everything except the CNN, which lives in oracle/ultraface_ref.py).
 *
 * THIS IS TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load it, and only as the
 * checker. The product (infercam_onnx_b200/) never links or calls it.
 *
 * PARITY STATUS: "parity unpinned" for resize bytes and NMS semantics. The
 * reference holds no golden vectors for them: its only test on
 * this path (infer_server/tests/integration_tests.rs:20-34) asserts face counts
 * and needs the real ONNX weights, which are downloaded at run time
 * (infer_server/src/nn.rs:21-22,156-162) and are absent here. The reference's
 * Rust sources cannot be compiled in this image (no cargo/rustc). What pins this
 * file is (a) the first-party Rust it restates line by line (nn.rs) and (b) the
 * published algorithm of the un-vendored crate `image 0.24.5`
 * (Cargo.toml:18, Cargo.lock:946-947), `src/imageops/sample.rs`
 * (`resize`, `vertical_sample`, `horizontal_sample`, `triangle_kernel`),
 * anchored on the call site nn.rs:74-80.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: the Rust code never
 * contracts a*b+c into an FMA, so neither may this file).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* R1: image::imageops::resize(img, nw, nh, FilterType::Triangle)            */
/*     call site: infer_server/src/nn.rs:74-80                               */
/* ------------------------------------------------------------------------- */

/* image 0.24.5 sample.rs `triangle_kernel`: 1-|x| inside (-1,1), else 0. */
static float tri(float x) {
    float a = fabsf(x);
    return a < 1.0f ? 1.0f - a : 0.0f;
}

static int64_t clamp_i64(int64_t v, int64_t lo, int64_t hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}

/*
 * One axis of the separable filter: for each of `dst_len` outputs the first
 * source index `left[o]`, tap count `ntaps[o]` and normalised f32 weights
 * `w[o*max_taps + i]`. Follows sample.rs horizontal_sample/vertical_sample:
 *   ratio = S as f32 / D as f32; sratio = max(ratio, 1); support = 1.0 * sratio
 *   c = (o + 0.5) * ratio; left = clamp(floor(c - support), 0, S-1)
 *   right = clamp(ceil(c + support), left+1, S); c -= 0.5
 *   w_i = tri((i - c)/sratio); w_i /= sum(w)   (sum in ascending i, f32)
 * Returns the max tap count (call with w == NULL to size the table).
 */
int orc_resize_axis_taps(int src_len, int dst_len, int max_taps,
                         int32_t* left, int32_t* ntaps, float* w) {
    float ratio = (float)src_len / (float)dst_len;
    float sratio = ratio < 1.0f ? 1.0f : ratio;
    float support = 1.0f * sratio;
    int worst = 0;
    for (int o = 0; o < dst_len; ++o) {
        float c = ((float)o + 0.5f) * ratio;
        int64_t l = (int64_t)floorf(c - support);
        l = clamp_i64(l, 0, (int64_t)src_len - 1);
        int64_t r = (int64_t)ceilf(c + support);
        r = clamp_i64(r, l + 1, (int64_t)src_len);
        c = c - 0.5f;
        int n = (int)(r - l);
        if (n > worst) worst = n;
        if (!w) continue;
        if (n > max_taps) return -1;
        left[o] = (int32_t)l;
        ntaps[o] = n;
        float sum = 0.0f;
        for (int i = 0; i < n; ++i) {
            float wi = tri(((float)(l + i) - c) / sratio);
            w[(size_t)o * max_taps + i] = wi;
            sum += wi;
        }
        for (int i = 0; i < n; ++i) w[(size_t)o * max_taps + i] /= sum;
    }
    return worst;
}

/* f32::round — half away from zero (FloatNearest in sample.rs). */
static float round_half_away(float v) { return roundf(v); }

static float clampf(float v, float lo, float hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}

/*
 * src: h x w x 3 u8 (RgbImage, row-major HWC, no padding); dst: nh x nw x 3.
 * round_intermediate = 0 reproduces image 0.24.x (f32 image between the two
 * passes); 1 reproduces the pre-0.24 behaviour (u8 between passes) and exists
 * only as the one-line switch SURVEY.md §8a-R asks for.
 * Returns 0, or -1 on allocation failure.
 */
int orc_resize_triangle(const uint8_t* src, int w, int h, uint8_t* dst, int nw,
                        int nh, int round_intermediate) {
    if (nw == w && nh == h) { /* sample.rs resize(): identity → plain copy */
        memcpy(dst, src, (size_t)w * h * 3);
        return 0;
    }
    int vt = orc_resize_axis_taps(h, nh, 0, NULL, NULL, NULL);
    int ht = orc_resize_axis_taps(w, nw, 0, NULL, NULL, NULL);
    int32_t* vl = malloc(sizeof(int32_t) * nh);
    int32_t* vn = malloc(sizeof(int32_t) * nh);
    float* vw = malloc(sizeof(float) * (size_t)nh * vt);
    int32_t* hl = malloc(sizeof(int32_t) * nw);
    int32_t* hn = malloc(sizeof(int32_t) * nw);
    float* hw = malloc(sizeof(float) * (size_t)nw * ht);
    float* tmp = malloc(sizeof(float) * (size_t)w * nh * 3);
    if (!vl || !vn || !vw || !hl || !hn || !hw || !tmp) return -1;
    orc_resize_axis_taps(h, nh, vt, vl, vn, vw);
    orc_resize_axis_taps(w, nw, ht, hl, hn, hw);

    /* vertical_sample: w x nh f32 image, not clamped, not rounded */
    for (int oy = 0; oy < nh; ++oy) {
        const float* ws = vw + (size_t)oy * vt;
        for (int x = 0; x < w; ++x) {
            for (int c = 0; c < 3; ++c) {
                float t = 0.0f;
                for (int i = 0; i < vn[oy]; ++i) {
                    float p = (float)src[((size_t)(vl[oy] + i) * w + x) * 3 + c];
                    float m = p * ws[i];
                    t = t + m;
                }
                if (round_intermediate)
                    t = round_half_away(clampf(t, 0.0f, 255.0f));
                tmp[((size_t)oy * w + x) * 3 + c] = t;
            }
        }
    }
    /* horizontal_sample: clamp to [0,255], round half away, cast to u8 */
    for (int ox = 0; ox < nw; ++ox) {
        const float* ws = hw + (size_t)ox * ht;
        for (int y = 0; y < nh; ++y) {
            for (int c = 0; c < 3; ++c) {
                float t = 0.0f;
                for (int i = 0; i < hn[ox]; ++i) {
                    float p = tmp[((size_t)y * w + hl[ox] + i) * 3 + c];
                    float m = p * ws[i];
                    t = t + m;
                }
                t = round_half_away(clampf(t, 0.0f, 255.0f));
                dst[((size_t)y * nw + ox) * 3 + c] = (uint8_t)t;
            }
        }
    }
    free(vl); free(vn); free(vw); free(hl); free(hn); free(hw); free(tmp);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* R2: normalise + HWC→NCHW, nn.rs:82-91                                     */
/*     out[0,c,y,x] = (px as f32 / 255.0 - mean[c]) / std[c]                 */
/*     preset 0 = reference constants (nn.rs:86-87); preset 1 = the          */
/*     (x-127)/128 variant BASELINE.json's north_star mentions.              */
/* ------------------------------------------------------------------------- */
void orc_normalise_nchw(const uint8_t* hwc, int w, int h, int preset, float* out) {
    static const float mean[3] = {0.485f, 0.456f, 0.406f};
    static const float std[3] = {0.229f, 0.224f, 0.225f};
    for (int c = 0; c < 3; ++c)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                float p = (float)hwc[((size_t)y * w + x) * 3 + c];
                float v;
                if (preset == 0) {
                    float q = p / 255.0f;
                    float d = q - mean[c];
                    v = d / std[c];
                } else {
                    float d = p - 127.0f;
                    v = d / 128.0f;
                }
                out[((size_t)c * h + y) * w + x] = v;
            }
}

/* ------------------------------------------------------------------------- */
/* P5/P6: iou + bbox_area, nn.rs:227-260, EPS nn.rs:18                       */
/* ------------------------------------------------------------------------- */
static const float EPS = 1.0e-7f;

float orc_bbox_area(const float* b) {
    float width = b[3] - b[1];  /* names swapped in the reference; product identical */
    float height = b[2] - b[0];
    if (width < 0.0f || height < 0.0f) return 0.0f;
    return width * height;
}

float orc_iou(const float* a, const float* b) {
    float ov[4];
    ov[0] = fmaxf(a[0], b[0]);
    ov[1] = fmaxf(a[1], b[1]);
    ov[2] = fminf(a[2], b[2]);
    ov[3] = fminf(a[3], b[3]);
    float o = orc_bbox_area(ov);
    float s = orc_bbox_area(a) + orc_bbox_area(b);
    s = s - o;
    s = s + EPS;
    return o / s;
}

/* ------------------------------------------------------------------------- */
/* P1-P4: postproc, nn.rs:109-140 + non_maximum_suppression nn.rs:198-224    */
/* ------------------------------------------------------------------------- */
typedef struct { float conf; int32_t idx; } cand_t;

/* Stable merge sort ascending by conf — Rust's slice::sort_by is stable. */
static void merge_sort(cand_t* a, cand_t* t, int n) {
    if (n < 2) return;
    int m = n / 2;
    merge_sort(a, t, m);
    merge_sort(a + m, t, n - m);
    int i = 0, j = m, k = 0;
    while (i < m && j < n) {
        if (a[j].conf < a[i].conf) t[k++] = a[j++]; /* strict: equal keys keep order */
        else t[k++] = a[i++];
    }
    while (i < m) t[k++] = a[i++];
    while (j < n) t[k++] = a[j++];
    memcpy(a, t, sizeof(cand_t) * n);
}

/*
 * scores: K x 2 (face probability in column 1); boxes: K x 4.
 * out: up to cap rows of [x0,y0,x1,y1,conf] in selection order (descending
 * confidence); out_idx (optional): the prior index of each selected row.
 * Returns the number selected (may exceed cap; only cap rows are written),
 * or -1 on allocation failure.
 */
int orc_postproc(const float* scores, const float* boxes, int K, float min_conf,
                 float max_iou, float* out, int32_t* out_idx, int cap) {
    cand_t* c = malloc(sizeof(cand_t) * (K > 0 ? K : 1));
    cand_t* t = malloc(sizeof(cand_t) * (K > 0 ? K : 1));
    int32_t* sel = malloc(sizeof(int32_t) * (K > 0 ? K : 1));
    if (!c || !t || !sel) return -1;
    int n = 0;
    for (int k = 0; k < K; ++k) {
        float conf = scores[(size_t)k * 2 + 1];
        if (conf > min_conf) { c[n].conf = conf; c[n].idx = k; ++n; } /* strict; NaN dropped */
    }
    merge_sort(c, t, n);
    int ns = 0;
    for (int p = n - 1; p >= 0; --p) { /* pop() from the back */
        const float* bb = boxes + (size_t)c[p].idx * 4;
        int keep = 1;
        for (int s = 0; s < ns; ++s) {
            if (orc_iou(bb, boxes + (size_t)sel[s] * 4) > max_iou) { keep = 0; break; }
        }
        if (!keep) continue;
        if (ns < cap) {
            memcpy(out + (size_t)ns * 5, bb, sizeof(float) * 4);
            out[(size_t)ns * 5 + 4] = c[p].conf;
            if (out_idx) out_idx[ns] = c[p].idx;
        }
        sel[ns++] = c[p].idx;
    }
    free(c); free(t); free(sel);
    return ns;
}
