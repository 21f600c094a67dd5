/* jpeg_oracle.c — CPU restatement of the sample-domain half of libjpeg-turbo's default decoder, the decoder the
 * reference runs in front of the hot path (`turbojpeg::decompress_image`, infer_server/src/inferer.rs:35; turbojpeg 0.5.2
 * over libjpeg-turbo, Cargo.lock). TEST INFRASTRUCTURE (see oracle/__init__.py).
 *
 * Restated from libjpeg-turbo's published sources (not vendored under /root/reference): `jpeg_idct_islow` (jidctint.c),
 * `h2v1_fancy_upsample` / `h2v2_fancy_upsample` (jdsample.c), `ycc_rgb_convert` + `build_ycc_rgb_table` (jdcolor.c), the
 * range-limit table of jdmaster.c (`prepare_range_limit_table`) and the edge handling of jdmainct.c (context rows
 * replicate the first / last REAL sample row). PARITY STATUS: pinned — libjpeg-turbo itself is in this image behind PIL
 * and OpenCV, and tests/test_jpeg.py checks this restatement against its pixels bit for bit on every sampling mode.
 *
 * Input: quantised coefficient blocks in decode (MCU-interleaved) order, natural (row-major) order inside a block.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CONST_BITS 13
#define PASS1_BITS 2
#define FIX_0_298631336 2446
#define FIX_0_390180644 3196
#define FIX_0_541196100 4433
#define FIX_0_765366865 6270
#define FIX_0_899976223 7373
#define FIX_1_175875602 9633
#define FIX_1_501321110 12299
#define FIX_1_847759065 15137
#define FIX_1_961570560 16069
#define FIX_2_053119869 16819
#define FIX_2_562915447 20995
#define FIX_3_072711026 25172
#define DESCALE(x, n) (((x) + (1L << ((n)-1))) >> (n))

/* range_limit[x & 1023] of the IDCT: sample_range_limit + CENTERJSAMPLE, i.e. clamp(x + 128) with libjpeg's wrap-around
 * for wildly out-of-range values */
static uint8_t idct_limit(long x) {
    int v = (int)(x & 1023);
    if (v < 128) return (uint8_t)(v + 128);
    if (v < 512) return 255;
    if (v < 896) return 0;
    return (uint8_t)(v - 896);
}

static void idct_islow(const int16_t* coef, const uint16_t* q, uint8_t* out, int stride) {
    long tmp0, tmp1, tmp2, tmp3, tmp10, tmp11, tmp12, tmp13, z1, z2, z3, z4, z5;
    int ws[64];
    for (int c = 0; c < 8; ++c) {
        const int16_t* in = coef + c;
        const uint16_t* qq = q + c;
#define DQ(r) ((long)in[8 * (r)] * (long)qq[8 * (r)])
        z2 = DQ(2); z3 = DQ(6);
        z1 = (z2 + z3) * FIX_0_541196100;
        tmp2 = z1 + z3 * (-FIX_1_847759065);
        tmp3 = z1 + z2 * FIX_0_765366865;
        z2 = DQ(0); z3 = DQ(4);
        tmp0 = (z2 + z3) * (1L << CONST_BITS);
        tmp1 = (z2 - z3) * (1L << CONST_BITS);
        tmp10 = tmp0 + tmp3; tmp13 = tmp0 - tmp3; tmp11 = tmp1 + tmp2; tmp12 = tmp1 - tmp2;
        tmp0 = DQ(7); tmp1 = DQ(5); tmp2 = DQ(3); tmp3 = DQ(1);
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; z4 = tmp1 + tmp3;
        z5 = (z3 + z4) * FIX_1_175875602;
        tmp0 *= FIX_0_298631336; tmp1 *= FIX_2_053119869; tmp2 *= FIX_3_072711026; tmp3 *= FIX_1_501321110;
        z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        ws[c + 0] = (int)DESCALE(tmp10 + tmp3, CONST_BITS - PASS1_BITS);
        ws[c + 56] = (int)DESCALE(tmp10 - tmp3, CONST_BITS - PASS1_BITS);
        ws[c + 8] = (int)DESCALE(tmp11 + tmp2, CONST_BITS - PASS1_BITS);
        ws[c + 48] = (int)DESCALE(tmp11 - tmp2, CONST_BITS - PASS1_BITS);
        ws[c + 16] = (int)DESCALE(tmp12 + tmp1, CONST_BITS - PASS1_BITS);
        ws[c + 40] = (int)DESCALE(tmp12 - tmp1, CONST_BITS - PASS1_BITS);
        ws[c + 24] = (int)DESCALE(tmp13 + tmp0, CONST_BITS - PASS1_BITS);
        ws[c + 32] = (int)DESCALE(tmp13 - tmp0, CONST_BITS - PASS1_BITS);
#undef DQ
    }
    for (int r = 0; r < 8; ++r) {
        const int* w = ws + 8 * r;
        uint8_t* o = out + (size_t)r * stride;
        z2 = w[2]; z3 = w[6];
        z1 = (z2 + z3) * FIX_0_541196100;
        tmp2 = z1 + z3 * (-FIX_1_847759065);
        tmp3 = z1 + z2 * FIX_0_765366865;
        tmp0 = ((long)w[0] + (long)w[4]) * (1L << CONST_BITS);
        tmp1 = ((long)w[0] - (long)w[4]) * (1L << CONST_BITS);
        tmp10 = tmp0 + tmp3; tmp13 = tmp0 - tmp3; tmp11 = tmp1 + tmp2; tmp12 = tmp1 - tmp2;
        tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; z4 = tmp1 + tmp3;
        z5 = (z3 + z4) * FIX_1_175875602;
        tmp0 *= FIX_0_298631336; tmp1 *= FIX_2_053119869; tmp2 *= FIX_3_072711026; tmp3 *= FIX_1_501321110;
        z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        o[0] = idct_limit(DESCALE(tmp10 + tmp3, CONST_BITS + PASS1_BITS + 3));
        o[7] = idct_limit(DESCALE(tmp10 - tmp3, CONST_BITS + PASS1_BITS + 3));
        o[1] = idct_limit(DESCALE(tmp11 + tmp2, CONST_BITS + PASS1_BITS + 3));
        o[6] = idct_limit(DESCALE(tmp11 - tmp2, CONST_BITS + PASS1_BITS + 3));
        o[2] = idct_limit(DESCALE(tmp12 + tmp1, CONST_BITS + PASS1_BITS + 3));
        o[5] = idct_limit(DESCALE(tmp12 - tmp1, CONST_BITS + PASS1_BITS + 3));
        o[3] = idct_limit(DESCALE(tmp13 + tmp0, CONST_BITS + PASS1_BITS + 3));
        o[4] = idct_limit(DESCALE(tmp13 - tmp0, CONST_BITS + PASS1_BITS + 3));
    }
}

static uint8_t clamp255(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

/* one upsampled chroma sample at output (x, y); plane = downsampled component, rw x rh REAL samples, stride pw */
static int up_sample(const uint8_t* plane, int pw, int rw, int rh, int hs_ratio, int vs_ratio, int x, int y) {
    if (hs_ratio == 1 && vs_ratio == 1) return plane[(size_t)y * pw + x];
    if (hs_ratio == 2 && vs_ratio == 1) { /* h2v1_fancy_upsample */
        const uint8_t* in = plane + (size_t)y * pw;
        int cx = x >> 1;
        if (x & 1) return cx == rw - 1 ? in[cx] : (in[cx] * 3 + in[cx + 1] + 2) >> 2;
        return cx == 0 ? in[0] : (in[cx] * 3 + in[cx - 1] + 1) >> 2;
    }
    /* h2v2_fancy_upsample: rows: nearer input row weight 3, farther 1; context rows replicate the first / last real row */
    int cy = y >> 1;
    int oy = (y & 1) ? cy + 1 : cy - 1;
    if (oy < 0) oy = 0;
    if (oy > rh - 1) oy = rh - 1;
    const uint8_t* in0 = plane + (size_t)cy * pw;
    const uint8_t* in1 = plane + (size_t)oy * pw;
    int cx = x >> 1;
    int thiscol = in0[cx] * 3 + in1[cx];
    if (x & 1) {
        if (cx == rw - 1) return (thiscol * 4 + 7) >> 4;
        return (thiscol * 3 + (in0[cx + 1] * 3 + in1[cx + 1]) + 7) >> 4;
    }
    if (cx == 0) return (thiscol * 4 + 8) >> 4;
    return (thiscol * 3 + (in0[cx - 1] * 3 + in1[cx - 1]) + 8) >> 4;
}

/* geometry mirrors uf::JpegPlan (csrc/jpeg_decode.h) */
int orc_jpeg_reconstruct(const int16_t* coefs, uint32_t w, uint32_t h, uint32_t ncomp, const uint32_t* hs, const uint32_t* vs,
                         const uint16_t* quant /* [ncomp][64] natural order */, uint8_t* rgb_out /* h*w*3 */) {
    uint32_t hmax = 1, vmax = 1;
    for (uint32_t c = 0; c < ncomp; ++c) { if (hs[c] > hmax) hmax = hs[c]; if (vs[c] > vmax) vmax = vs[c]; }
    const uint32_t mx = (w + 8 * hmax - 1) / (8 * hmax), my = (h + 8 * vmax - 1) / (8 * vmax);
    uint8_t* planes[3] = {0, 0, 0};
    uint32_t pw[3], ph[3], rw[3], rh[3];
    for (uint32_t c = 0; c < ncomp; ++c) {
        pw[c] = mx * hs[c] * 8; ph[c] = my * vs[c] * 8;
        rw[c] = (w * hs[c] + hmax - 1) / hmax; rh[c] = (h * vs[c] + vmax - 1) / vmax;
        planes[c] = (uint8_t*)malloc((size_t)pw[c] * ph[c]);
        if (!planes[c]) return -1;
    }
    size_t blk = 0;
    for (uint32_t y = 0; y < my; ++y)
        for (uint32_t x = 0; x < mx; ++x)
            for (uint32_t c = 0; c < ncomp; ++c)
                for (uint32_t v = 0; v < vs[c]; ++v)
                    for (uint32_t hh = 0; hh < hs[c]; ++hh, ++blk) {
                        const uint32_t bx = x * hs[c] + hh, by = y * vs[c] + v;
                        idct_islow(coefs + blk * 64, quant + 64 * c, planes[c] + (size_t)by * 8 * pw[c] + bx * 8, (int)pw[c]);
                    }
    /* jdcolor.c build_ycc_rgb_table */
    int cr_r[256], cb_b[256];
    long cr_g[256], cb_g[256];
#define FIX(x) ((long)((x) * (1L << 16) + 0.5))
    for (int i = 0; i < 256; ++i) {
        long x = i - 128;
        cr_r[i] = (int)((FIX(1.40200) * x + (1L << 15)) >> 16);
        cb_b[i] = (int)((FIX(1.77200) * x + (1L << 15)) >> 16);
        cr_g[i] = (-FIX(0.71414)) * x;
        cb_g[i] = (-FIX(0.34414)) * x + (1L << 15);
    }
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            uint8_t* o = rgb_out + ((size_t)y * w + x) * 3;
            const int Y = planes[0][(size_t)y * pw[0] + x];
            if (ncomp == 1) { o[0] = o[1] = o[2] = (uint8_t)Y; continue; }
            const int hr = (int)(hmax / hs[1]), vr = (int)(vmax / vs[1]);
            const int cb = up_sample(planes[1], (int)pw[1], (int)rw[1], (int)rh[1], hr, vr, (int)x, (int)y);
            const int cr = up_sample(planes[2], (int)pw[2], (int)rw[2], (int)rh[2], hr, vr, (int)x, (int)y);
            o[0] = clamp255(Y + cr_r[cr]);
            o[1] = clamp255(Y + (int)((cb_g[cb] + cr_g[cr]) >> 16));
            o[2] = clamp255(Y + cb_b[cb]);
        }
    for (uint32_t c = 0; c < ncomp; ++c) free(planes[c]);
    return 0;
}
