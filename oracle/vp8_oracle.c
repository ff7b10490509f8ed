/*
 * oracle/vp8_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the reference's inter-frame device kernels; see vp8_oracle.h for
 * what pins it.  All integer arithmetic is 32-bit two's complement with arithmetic >> and
 * truncating / (compile with -fwrapv); 16-bit lanes of the reference's short/ushort vector
 * code are reproduced with explicit (int16_t)/(uint16_t) truncation.
 */
#include "vp8_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* src/GPU_kernels.cl:58-80 == src/vp8enc.h:17-41 (VP8 spec tables) */
static const int dc_qlookup[128] = {
    4,   5,   6,   7,   8,   9,   10,  10,  11,  12,  13,  14,  15,  16,  17,  17,  18,  19,  20,  20,  21,  21,
    22,  22,  23,  23,  24,  25,  25,  26,  27,  28,  29,  30,  31,  32,  33,  34,  35,  36,  37,  37,  38,  39,
    40,  41,  42,  43,  44,  45,  46,  46,  47,  48,  49,  50,  51,  52,  53,  54,  55,  56,  57,  58,  59,  60,
    61,  62,  63,  64,  65,  66,  67,  68,  69,  70,  71,  72,  73,  74,  75,  76,  76,  77,  78,  79,  80,  81,
    82,  83,  84,  85,  86,  87,  88,  89,  91,  93,  95,  96,  98,  100, 101, 102, 104, 106, 108, 110, 112, 114,
    116, 118, 122, 124, 126, 128, 130, 132, 134, 136, 138, 140, 143, 145, 148, 151, 154, 157};
static const int ac_qlookup[128] = {
    4,   5,   6,   7,   8,   9,   10,  11,  12,  13,  14,  15,  16,  17,  18,  19,  20,  21,  22,  23,  24,  25,
    26,  27,  28,  29,  30,  31,  32,  33,  34,  35,  36,  37,  38,  39,  40,  41,  42,  43,  44,  45,  46,  47,
    48,  49,  50,  51,  52,  53,  54,  55,  56,  57,  58,  60,  62,  64,  66,  68,  70,  72,  74,  76,  78,  80,
    82,  84,  86,  88,  90,  92,  94,  96,  98,  100, 102, 104, 106, 108, 110, 112, 114, 116, 119, 122, 125, 128,
    131, 134, 137, 140, 143, 146, 149, 152, 155, 158, 161, 164, 167, 170, 173, 177, 181, 185, 189, 193, 197, 201,
    205, 209, 213, 217, 221, 225, 229, 234, 239, 245, 249, 254, 259, 264, 269, 274, 279, 284};
/* RFC 6386 subpixel_filters == src/GPU_kernels.cl:563-572 */
static const int sixtap[8][6] = {{0, 0, 128, 0, 0, 0},     {0, -6, 123, 12, -1, 0}, {2, -11, 108, 36, -8, 1},
                                 {0, -9, 93, 50, -6, 0},   {3, -16, 77, 77, -16, 3}, {0, -6, 50, 93, -9, 0},
                                 {1, -8, 36, 108, -11, 2}, {0, -1, 12, 123, -6, 0}};
/* src/GPU_kernels.cl:1489 */
static const int inv_zigzag[16] = {0, 1, 5, 6, 2, 4, 7, 12, 3, 8, 11, 13, 9, 10, 14, 15};

static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int sat8(int v) { return clampi(v, 0, 255); }

/* ---------------------------------------------------------------------------------------- */
int vp8o_weight(const int r[16]) {
    /* pass 1, per column; lines 87-133.  Lines 105-108 overwrite s4..s7 (b1) with c1 and leave
     * s8..sB holding the raw third row: the "clobber" quirk Q1. */
    int o[16], f[16];
    for (int k = 0; k < 4; ++k) {
        const int r0 = r[k], r1 = r[4 + k], r2 = r[8 + k], r3 = r[12 + k];
        const int a1 = (r0 + r3) << 3;
        const int d1 = (r0 - r3) << 3;
        const int c1 = (r1 - r2) << 3;
        const int x = r2;
        o[k] = a1 + c1;
        o[8 + k] = a1 - c1;
        o[4 + k] = (x * 2217 + d1 * 5352 + 14500) >> 12;
        o[12 + k] = (d1 * 2217 - x * 5352 + 7500) >> 12;
    }
    /* pass 2, per row; lines 135-180 */
    for (int row = 0; row < 4; ++row) {
        const int *e = o + 4 * row;
        const int a = e[0] + e[3], d = e[0] - e[3], b = e[1] + e[2], c = e[1] - e[2];
        f[4 * row + 0] = (a + b + 7) >> 4;
        f[4 * row + 2] = (a - b + 7) >> 4;
        f[4 * row + 1] = ((c * 2217 + d * 5352 + 12000) >> 16) + (d != 0);
        f[4 * row + 3] = (d * 2217 - c * 5352 + 51000) >> 16;
    }
    /* lines 182-187 */
    int sum = iabs(f[0]) / 4;
    for (int i = 1; i < 16; ++i) sum += iabs(f[i]);
    return sum;
}

void vp8o_reset_vectors(int16_t *last1, int16_t *last2, int16_t *gold1, int16_t *gold2, int16_t *alt1,
                        int16_t *alt2, int32_t *last_Bdiff, int32_t *gold_Bdiff, int32_t *alt_Bdiff, int n) {
    int16_t *nets[6] = {last1, last2, gold1, gold2, alt1, alt2};
    int32_t *m[3] = {last_Bdiff, gold_Bdiff, alt_Bdiff};
    for (int k = 0; k < 6; ++k) memset(nets[k], 0, (size_t)n * 4);
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < n; ++i) m[k][i] = 0x7fffffff;
}

void vp8o_downsample_x2(const uint8_t *src, uint8_t *dst, int w, int h) {
    const int n = w * h / 4;
#pragma omp parallel for
    for (int b = 0; b < n; ++b) {
        const int x = (b % (w / 2)) * 2, y = (b / (w / 2)) * 2;
        const int i = y * w + x;
        dst[(y / 2) * (w / 2) + x / 2] = (uint8_t)((src[i] + src[i + 1] + src[i + w] + src[i + w + 1] + 2) / 4);
    }
}

/* ---------------------------------------------------------------------------------------- */
void vp8o_luma_search_1step(const uint8_t *cur, const uint8_t *prev, const int16_t *src_net, int16_t *dst_net,
                            int net_width, int width, int height, int rate) {
    const int cut_width = (width / 8) * 8;
    const int nblocks = (width / 8) * (height / 8);
    static const int dx1[4] = {0, 0, 4, 4}, dy1[4] = {0, 4, 0, 4}; /* lines 456-457 */
    /* dst_net is never src_net (ping-pong), but blocks of one launch read parents that other
     * blocks of the same launch never write, so a parallel loop is exact */
#pragma omp parallel for schedule(dynamic, 16)
    for (int n = 0; n < nblocks; ++n) {
        const int16_t cx = (int16_t)((n % (cut_width / 8)) * 8);
        const int16_t cy = (int16_t)((n / (cut_width / 8)) * 8);
        const int16_t px0 = (int16_t)(cx / 2), py0 = (int16_t)(cy / 2);
        const int parent = (py0 / 8) * net_width + (px0 / 8);
        /* short2 /= int, then forced to 0 above rate 8; lines 498-501 */
        int16_t v0x = (int16_t)(src_net[2 * parent] / (int16_t)rate);
        int16_t v0y = (int16_t)(src_net[2 * parent + 1] / (int16_t)rate);
        if (rate > 8) v0x = v0y = 0;
        int16_t bestx = v0x, besty = v0y; /* a displacement until a candidate wins (Q2) */
        const int out = (cy / 8) * net_width + (cx / 8);
        uint16_t min_diff = 0x7fff;
        for (int k = 0; k < 25; ++k) {
            const int16_t px = (int16_t)(cx + v0x + (k % 5 - 2));
            const int16_t py = (int16_t)(cy + v0y + (k / 5 - 2));
            /* candidates not fully inside the frame get Diff |= 0x7fff and can never pass the
             * strict "<" (lines 546-553); the reference reads out of bounds for them, we do not */
            if (px < 0 || px > width - 8 || py < 0 || py > height - 8) continue;
            uint16_t diff = 0;
            for (int j = 0; j < 4; ++j) {
                int r[16];
                for (int y = 0; y < 4; ++y)
                    for (int x = 0; x < 4; ++x)
                        r[4 * y + x] = (int)cur[(cy + dy1[j] + y) * width + cx + dx1[j] + x] -
                                       (int)prev[(py + dy1[j] + y) * width + px + dx1[j] + x];
                diff = (uint16_t)(diff + vp8o_weight(r));
            }
            /* lines 542-543: neighbour coherence, only at rates 2 and 1 (Q3) */
            diff = (uint16_t)(diff + (iabs(iabs(px - cx) - v0x) + iabs(iabs(py - cy) - v0y)) * (rate < 4) * 64 / 2);
            if (diff < min_diff) {
                bestx = px;
                besty = py;
                min_diff = diff;
            }
        }
        dst_net[2 * out] = (int16_t)((int16_t)(bestx - cx) * (int16_t)rate);
        dst_net[2 * out + 1] = (int16_t)((int16_t)(besty - cy) * (int16_t)rate);
    }
}

/* clamp-to-edge nearest read, the sampler at src/GPU_kernels.cl:562 */
static inline int img_px(const uint8_t *img, int w, int h, int x, int y) {
    return img[(size_t)clampi(y, 0, h - 1) * w + clampi(x, 0, w - 1)];
}
/* one horizontally filtered sample, before the uchar conversion */
static inline int hraw(const uint8_t *img, int w, int h, int x, int y, int fx) {
    int s = 64;
    for (int t = 0; t < 6; ++t) s += sixtap[fx][t] * img_px(img, w, h, x - 2 + t, y);
    return s / 128; /* truncating division */
}

/* "search flavour" 4x4 prediction (construct_opt1/2, lines 776-1066): every horizontal line
 * is saturated before the vertical pass */
static void predict4x4_search(const uint8_t *img, int w, int h, int ox, int oy, int fx, int fy, int out[16]) {
    int line[9][4];
    for (int l = 0; l < 9; ++l)
        for (int c = 0; c < 4; ++c) line[l][c] = sat8(hraw(img, w, h, ox + c, oy - 2 + l, fx));
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            int s = 64;
            for (int t = 0; t < 6; ++t) s += sixtap[fy][t] * line[r + t][c];
            out[4 * r + c] = sat8(s / 128);
        }
}

void vp8o_luma_search_2step(const uint8_t *cur, const uint8_t *ref, const int16_t *net, int16_t *ref_net,
                            int32_t *ref_Bdiff, int width, int height) {
    const int nblocks = width * height / 64;
    static const int dx4[4] = {0, 0, 16, 16}, dy4[4] = {0, 16, 0, 16}; /* lines 454-455 */
#pragma omp parallel for schedule(dynamic, 16)
    for (int n = 0; n < nblocks; ++n) {
        /* all of these live in short lanes in the reference (pdata1/pdata2) */
        const int16_t v0x = (int16_t)(net[2 * n] * 4), v0y = (int16_t)(net[2 * n + 1] * 4);
        const int bw = width / 8;
        const int16_t c4x = (int16_t)((n % bw) * 8 * 4), c4y = (int16_t)((n / bw) * 8 * 4);
        const int bx = (n % bw) * 8, by = (n / bw) * 8;
        int16_t bestx = (int16_t)(width * 4 - 32), besty = (int16_t)(height * 4 - 32);
        int min_diff = 0x7fff;
        for (int k = 0; k < 26; ++k) {
            int16_t qx = (int16_t)(c4x + v0x + (k % 5 - 2));
            int16_t qy = (int16_t)(c4y + v0y + (k / 5 - 2));
            if (k == 25) {
                qx = c4x;
                qy = c4y;
            }
            /* forbidden candidates: garbage cost | 0x7fff >= 0x7fff never passes "<" (lines 1180-1190) */
            if (qx < 0 || qx > width * 4 - 32 || qy < 0 || qy > height * 4 - 32) continue;
            const int fx = (qx % 4) * 2, fy = (qy % 4) * 2;
            int diff = 0;
            for (int j = 0; j < 4; ++j) {
                int pred[16], r[16];
                predict4x4_search(ref, width, height, (qx + dx4[j]) / 4, (qy + dy4[j]) / 4, fx, fy, pred);
                for (int y = 0; y < 4; ++y)
                    for (int x = 0; x < 4; ++x)
                        r[4 * y + x] = (int)cur[(by + dy4[j] / 4 + y) * width + bx + dx4[j] / 4 + x] - pred[4 * y + x];
                diff += vp8o_weight(r);
            }
            if (k != 25) diff += (iabs(qx - c4x - v0x) + iabs(qy - c4y - v0y)) * 64 / 2;
            if (diff < min_diff) {
                bestx = qx;
                besty = qy;
                min_diff = diff;
            }
        }
        const int16_t mvx = (int16_t)(bestx - c4x), mvy = (int16_t)(besty - c4y);
        if (mvx != 0 || mvy != 0) min_diff -= (iabs(mvx - v0x) + iabs(mvy - v0y)) * 64 / 2; /* lines 1195-1197 */
        ref_net[2 * n] = mvx;
        ref_net[2 * n + 1] = mvy;
        ref_Bdiff[n] = min_diff;
    }
}

void vp8o_select_reference(const int16_t *last_net, const int16_t *gold_net, const int16_t *alt_net,
                           const int32_t *last_Bdiff, const int32_t *gold_Bdiff, const int32_t *alt_Bdiff,
                           int32_t *MB_reference_frame, int16_t *MB_vectors, int width, int height,
                           int use_golden, int use_altref) {
    const int mb_width = width / 16, mb_count = mb_width * (height / 16), bw = mb_width * 2;
    for (int mb = 0; mb < mb_count; ++mb) {
        const int b = ((mb / mb_width) * 2) * bw + (mb % mb_width) * 2;
        const int idx[4] = {b, b + 1, b + bw, b + bw + 1};
        int d1 = 0, d2 = 0x7fffffff, ref;
        for (int i = 0; i < 4; ++i) d1 += last_Bdiff[idx[i]];
        if (use_altref == 1) {
            d2 = 0;
            for (int i = 0; i < 4; ++i) d2 += alt_Bdiff[idx[i]];
        }
        ref = (d1 <= d2) ? VP8O_LAST : VP8O_ALTREF;
        d1 = (d1 <= d2) ? d1 : d2;
        d2 = 0x7fffffff;
        if (use_golden == 1) {
            d2 = 0;
            for (int i = 0; i < 4; ++i) d2 += gold_Bdiff[idx[i]];
        }
        ref = (d1 <= d2) ? ref : VP8O_GOLDEN;
        const int16_t *net = ref == VP8O_LAST ? last_net : (ref == VP8O_GOLDEN ? gold_net : alt_net);
        MB_reference_frame[mb] = ref;
        for (int i = 0; i < 4; ++i) {
            MB_vectors[8 * mb + 2 * i] = net[2 * idx[i]];
            MB_vectors[8 * mb + 2 * i + 1] = net[2 * idx[i] + 1];
        }
    }
}

void vp8o_pack_8x8_into_16x16(const int16_t *MB_vectors, int32_t *MB_parts, float *MB_SSIM, int mb_count) {
    for (int mb = 0; mb < mb_count; ++mb) {
        const int16_t *v = MB_vectors + 8 * mb;
        int same = 1;
        for (int i = 1; i < 4; ++i) same &= (v[2 * i] == v[0]) && (v[2 * i + 1] == v[1]);
        MB_SSIM[mb] = -2.0f;
        MB_parts[mb] = same ? VP8O_ARE16x16 : VP8O_ARE8x8;
    }
}

/* "predictor flavour" (construct, lines 574-774): source lines Y-2..Y+3 saturate, lines
 * Y+4..Y+6 are stored with a wrapping (uchar) cast (Q5); vertical results saturate */
static int predict4x4_construct(const uint8_t *img, int w, int h, int ox, int oy, int fx, int fy, int out[16]) {
    int line[9][4], wrapped = 0;
    for (int l = 0; l < 9; ++l)
        for (int c = 0; c < 4; ++c) {
            const int v = hraw(img, w, h, ox + c, oy - 2 + l, fx);
            line[l][c] = (l < 6) ? sat8(v) : (v & 255);
            wrapped |= (l >= 6) && (v < 0 || v > 255);
        }
    int differs = 0;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            int s = 64;
            for (int t = 0; t < 6; ++t) s += sixtap[fy][t] * line[r + t][c];
            out[4 * r + c] = sat8(s / 128);
            if (wrapped) { /* test aid: what a conforming decoder predicts here (all lines saturate) */
                int s2 = 64;
                for (int t = 0; t < 6; ++t) {
                    const int l = r + t;
                    const int v = hraw(img, w, h, ox + c, oy - 2 + l, fx);
                    s2 += sixtap[fy][t] * sat8(v);
                }
                differs |= sat8(s2 / 128) != out[4 * r + c];
            }
        }
    return differs; /* Q5 changed the predictor of this block */
}

/* ---- test aid: where did the reference's non-conforming arithmetic change a result? -------------------------
 * Q5 (wrapping lines of `construct`) and Q7 (unclamped chaining in the loop filter) make the encoder's
 * reconstruction differ from what a VP8 decoder computes from the same bitstream.  The log records the first
 * events of each kind as (x, y, plane size tag) so that tests/test_decoder_pin.py can say exactly from which
 * frame on a libvpx decode may legitimately differ.  Not part of any result. */
#define VP8O_QUIRK_LOG_CAP 64
static int quirk_count[2];
static int quirk_pos[2][VP8O_QUIRK_LOG_CAP][3];
static void quirk_event(int which, int x, int y, int tag) {
#pragma omp critical(vp8o_quirk)
    {
        if (quirk_count[which] < VP8O_QUIRK_LOG_CAP) {
            quirk_pos[which][quirk_count[which]][0] = x;
            quirk_pos[which][quirk_count[which]][1] = y;
            quirk_pos[which][quirk_count[which]][2] = tag;
        }
        ++quirk_count[which];
    }
}
void vp8o_quirk_log_reset(void) { quirk_count[0] = quirk_count[1] = 0; }
int vp8o_quirk_log_get(int which, int *xyt, int cap) {
    const int n = quirk_count[which] < VP8O_QUIRK_LOG_CAP ? quirk_count[which] : VP8O_QUIRK_LOG_CAP;
    for (int i = 0; i < n && i < cap; ++i)
        for (int k = 0; k < 3; ++k) xyt[3 * i + k] = quirk_pos[which][i][k];
    return quirk_count[which];
}

void vp8o_prepare_predictors_and_residual(const uint8_t *cur, const uint8_t *ref, uint8_t *predictor,
                                          int16_t *residual, const int32_t *MB_reference_frame,
                                          const int16_t *MB_vectors, int width, int height, int plane, int ref_id) {
    const int mb_size = plane == 0 ? 16 : 8;
    const int g = plane == 0 ? 4 : 8;
    const int nblocks = (width / 4) * (height / 4);
#pragma omp parallel for schedule(dynamic, 64)
    for (int b = 0; b < nblocks; ++b) {
        const int x = (b % (width / 4)) * 4, y = (b / (width / 4)) * 4;
        const int mb = (y / mb_size) * (width / mb_size) + x / mb_size;
        if (MB_reference_frame[mb] != ref_id) continue;
        const int q = ((y % mb_size) / (mb_size / 2)) * 2 + (x % mb_size) / (mb_size / 2);
        const int vx = MB_vectors[8 * mb + 2 * q], vy = MB_vectors[8 * mb + 2 * q + 1];
        const int tx = x * g + vx, ty = y * g + vy;
        /* C remainder/division as in the kernel (lines 1315-1319).  tx,ty are >= 0 for every
         * vector the search can emit; a negative remainder would index the tap table out of
         * bounds in the reference, so we pin that unreachable case to phase 0 */
        int dx = (tx % g) * (plane == 0 ? 2 : 1), dy = (ty % g) * (plane == 0 ? 2 : 1);
        if (dx < 0) dx = 0;
        if (dy < 0) dy = 0;
        int pred[16];
        if (predict4x4_construct(ref, width, height, tx / g, ty / g, dx, dy, pred)) quirk_event(0, x, y, plane);
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                const int i = (y + r) * width + x + c;
                predictor[i] = (uint8_t)pred[4 * r + c];
                residual[i] = (int16_t)((int)cur[i] - pred[4 * r + c]);
            }
    }
}

/* ---------------------------------------------------------------------------------------- */
typedef struct { int y_dc, y_ac, uv_dc, uv_ac, y2_dc, y2_ac; } quants;
/* duplicated on the device in dct4x4 / wht4x4_iwht4x4 / idct4x4 (Q11): lines 1394-1408, 1515-1524 */
static quants derive_quants(const vp8o_segment_data *SD, int seg) {
    quants q;
    const int i = SD[seg].y_ac_i;
    q.y_ac = ac_qlookup[i];
    q.y_dc = dc_qlookup[clampi(i + SD[0].y_dc_idelta, 0, 127)];
    q.uv_dc = dc_qlookup[clampi(i + SD[0].uv_dc_idelta, 0, 127)];
    q.uv_ac = ac_qlookup[clampi(i + SD[0].uv_ac_idelta, 0, 127)];
    if (q.uv_dc > 132) q.uv_dc = 132;
    q.y2_dc = dc_qlookup[clampi(i + SD[0].y2_dc_idelta, 0, 127)] * 2;
    q.y2_ac = 31 * ac_qlookup[clampi(i + SD[0].y2_ac_idelta, 0, 127)] / 20;
    if (q.y2_ac < 8) q.y2_ac = 8;
    return q;
}

void vp8o_dct4x4(const int16_t *residual, int16_t *MB, int32_t *MB_segment_id, const int32_t *MB_parts,
                 const float *MB_SSIM, int width, int height, const vp8o_segment_data *SD, int segment_id,
                 float SSIM_target, int plane) {
    const int mb_size = plane == 0 ? 16 : 8;
    const int nblocks = (width / 4) * (height / 4);
    const quants Q = derive_quants(SD, segment_id);
#pragma omp parallel for schedule(dynamic, 64)
    for (int b = 0; b < nblocks; ++b) {
        const int x = (b % (width / 4)) * 4, y = (b / (width / 4)) * 4;
        const int mb = (y / mb_size) * (width / mb_size) + x / mb_size;
        if (MB_SSIM[mb] > SSIM_target) continue; /* line 1391 */
        MB_segment_id[mb] = segment_id;
        const int dc_q = plane == 0 ? (MB_parts[mb] == VP8O_ARE16x16 ? 1 : Q.y_dc) : Q.uv_dc;
        const int ac_q = plane == 0 ? Q.y_ac : Q.uv_ac;
        int L[16], o[16];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) L[4 * r + c] = residual[(y + r) * width + x + c];
        /* first pass down the columns (lines 1417-1429): the transpose of the libvpx order (Q8) */
        for (int c = 0; c < 4; ++c) {
            const int a1 = (L[c] + L[12 + c]) << 3, d1 = (L[c] - L[12 + c]) << 3;
            const int b1 = (L[4 + c] + L[8 + c]) << 3, c1 = (L[4 + c] - L[8 + c]) << 3;
            o[c] = a1 + b1;
            o[8 + c] = a1 - b1;
            o[4 + c] = (c1 * 2217 + d1 * 5352 + 14500) >> 12;
            o[12 + c] = (d1 * 2217 - c1 * 5352 + 7500) >> 12;
        }
        /* second pass along the rows (lines 1431-1476) */
        for (int r = 0; r < 4; ++r) {
            const int *e = o + 4 * r;
            const int a1 = e[0] + e[3], d1 = e[0] - e[3], b1 = e[1] + e[2], c1 = e[1] - e[2];
            L[4 * r + 0] = (a1 + b1 + 7) >> 4;
            L[4 * r + 2] = (a1 - b1 + 7) >> 4;
            L[4 * r + 1] = ((c1 * 2217 + d1 * 5352 + 12000) >> 16) + (d1 != 0);
            L[4 * r + 3] = (d1 * 2217 - c1 * 5352 + 51000) >> 16;
        }
        int blk = ((y % mb_size) / 4) * (mb_size / 4) + (x % mb_size) / 4;
        blk += plane == 1 ? 16 : (plane == 2 ? 20 : 0);
        int16_t *dst = MB + (size_t)mb * 400 + blk * 16;
        for (int k = 0; k < 16; ++k) dst[inv_zigzag[k]] = (int16_t)(L[k] / (k == 0 ? dc_q : ac_q)); /* truncating */
    }
}

/* the 1-D butterfly shared by WHT_and_quant / dequant_and_iWHT (lines 261-309, 349-395) */
static inline void wht_butterfly(int a0, int a1, int a2, int a3, int *o0, int *o1, int *o2, int *o3) {
    const int a = a0 + a3, b = a1 + a2, c = a1 - a2, d = a0 - a3;
    *o0 = a + b;
    *o1 = c + d;
    *o2 = a - b;
    *o3 = d - c;
}

void vp8o_wht4x4_iwht4x4(int16_t *MB, int32_t *MB_segment_id, const int32_t *MB_parts,
                         const vp8o_segment_data *SD, int segment_id, int mb_count) {
    const quants Q = derive_quants(SD, segment_id);
    for (int mb = 0; mb < mb_count; ++mb) {
        if (MB_segment_id[mb] != segment_id) continue;
        if (MB_parts[mb] != VP8O_ARE16x16) continue;
        int16_t *m = MB + (size_t)mb * 400;
        int L[16], t[16];
        for (int k = 0; k < 16; ++k) L[k] = m[k * 16]; /* DC of luma block k, 4x4 in block raster order */
        /* forward: down the columns, then along the rows */
        for (int c = 0; c < 4; ++c) wht_butterfly(L[c], L[4 + c], L[8 + c], L[12 + c], &t[c], &t[4 + c], &t[8 + c], &t[12 + c]);
        for (int r = 0; r < 4; ++r) wht_butterfly(t[4 * r], t[4 * r + 1], t[4 * r + 2], t[4 * r + 3], &L[4 * r], &L[4 * r + 1], &L[4 * r + 2], &L[4 * r + 3]);
        for (int k = 0; k < 16; ++k) {
            int v = L[k];
            v += (v > 0);
            v >>= 1;
            v /= (k == 0 ? Q.y2_dc : Q.y2_ac);
            L[k] = v;
            m[24 * 16 + inv_zigzag[k]] = (int16_t)v; /* Y2 block, lines 1532-1535 */
        }
        /* dequantise and invert: along the rows, then down the columns, (v+3)>>3 */
        for (int k = 0; k < 16; ++k) L[k] *= (k == 0 ? Q.y2_dc : Q.y2_ac);
        for (int r = 0; r < 4; ++r) wht_butterfly(L[4 * r], L[4 * r + 1], L[4 * r + 2], L[4 * r + 3], &t[4 * r], &t[4 * r + 1], &t[4 * r + 2], &t[4 * r + 3]);
        for (int c = 0; c < 4; ++c) wht_butterfly(t[c], t[4 + c], t[8 + c], t[12 + c], &L[c], &L[4 + c], &L[8 + c], &L[12 + c]);
        for (int k = 0; k < 16; ++k) m[k * 16] = (int16_t)((L[k] + 3) >> 3); /* lines 1537-1540 (Q9) */
    }
}

/* one 1-D inverse DCT (lines 201-216 / 225-242) */
static inline void idct_1d(int i0, int i1, int i2, int i3, int *o0, int *o1, int *o2, int *o3) {
    const int a1 = i0 + i2, b1 = i0 - i2;
    const int c1 = ((i1 * 35468) >> 16) - (i3 + ((i3 * 20091) >> 16));
    const int d1 = (i1 + ((i1 * 20091) >> 16)) + ((i3 * 35468) >> 16);
    *o0 = a1 + d1;
    *o3 = a1 - d1;
    *o1 = b1 + c1;
    *o2 = b1 - c1;
}

void vp8o_idct4x4(uint8_t *recon, const uint8_t *predictor, const int16_t *MB, const int32_t *MB_segment_id,
                  const int32_t *MB_parts, int width, int height, const vp8o_segment_data *SD, int segment_id,
                  int plane) {
    const int mb_size = plane == 0 ? 16 : 8;
    const int nblocks = (width / 4) * (height / 4);
    const quants Q = derive_quants(SD, segment_id);
#pragma omp parallel for schedule(dynamic, 64)
    for (int b = 0; b < nblocks; ++b) {
        const int x = (b % (width / 4)) * 4, y = (b / (width / 4)) * 4;
        const int mb = (y / mb_size) * (width / mb_size) + x / mb_size;
        if (MB_segment_id[mb] != segment_id) continue;
        const int dc_q = plane == 0 ? (MB_parts[mb] == VP8O_ARE16x16 ? 1 : Q.y_dc) : Q.uv_dc;
        const int ac_q = plane == 0 ? Q.y_ac : Q.uv_ac;
        int blk = ((y % mb_size) / 4) * (mb_size / 4) + (x % mb_size) / 4;
        blk += plane == 1 ? 16 : (plane == 2 ? 20 : 0);
        const int16_t *src = MB + (size_t)mb * 400 + blk * 16;
        int L[16], t[16];
        for (int k = 0; k < 16; ++k) L[k] = src[inv_zigzag[k]] * (k == 0 ? dc_q : ac_q);
        for (int c = 0; c < 4; ++c) idct_1d(L[c], L[4 + c], L[8 + c], L[12 + c], &t[c], &t[4 + c], &t[8 + c], &t[12 + c]);
        for (int r = 0; r < 4; ++r) {
            int o0, o1, o2, o3;
            idct_1d(t[4 * r], t[4 * r + 1], t[4 * r + 2], t[4 * r + 3], &o0, &o1, &o2, &o3);
            const int i = (y + r) * width + x;
            recon[i + 0] = (uint8_t)sat8(((o0 + 4) >> 3) + predictor[i + 0]);
            recon[i + 1] = (uint8_t)sat8(((o1 + 4) >> 3) + predictor[i + 1]);
            recon[i + 2] = (uint8_t)sat8(((o2 + 4) >> 3) + predictor[i + 2]);
            recon[i + 3] = (uint8_t)sat8(((o3 + 4) >> 3) + predictor[i + 3]);
        }
    }
}

/* ---------------------------------------------------------------------------------------- */
/* float4-lane structured sums of the SSIM kernels.  mad() == fmaf; everything
 * else rounds after every operation (-ffp-contract=off). */
void vp8o_count_SSIM(const uint8_t *f1, const uint8_t *f2, const int32_t *MB_segment_id, float *metric, int width,
                     int height, int segment_id, int n) {
    const int mbw = width / n, mb_count = mbw * (height / n);
    const float c1 = 0.01f * 0.01f * 255 * 255, c2 = 0.03f * 0.03f * 255 * 255;
    const float area = (float)(n * n);
#pragma omp parallel for schedule(dynamic, 16)
    for (int mb = 0; mb < mb_count; ++mb) {
        if (MB_segment_id[mb] != segment_id) continue;
        const uint8_t *a = f1 + (size_t)(mb / mbw) * n * width + (mb % mbw) * n;
        const uint8_t *b = f2 + (size_t)(mb / mbw) * n * width + (mb % mbw) * n;
        float l[4], M1, M2, D, C;
        /* means: per-lane integer-valued float sums, then s0+s1+s2+s3 */
        for (int k = 0; k < 4; ++k) l[k] = 0.0f;
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) l[x & 3] += (float)a[y * width + x];
        M1 = (l[0] + l[1] + l[2] + l[3]) / area;
        for (int k = 0; k < 4; ++k) l[k] = 0.0f;
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) l[x & 3] += (float)b[y * width + x];
        M2 = (l[0] + l[1] + l[2] + l[3]) / area;
        /* variances: first group d*d, every later group mad(d,d,acc); lines 1694-1697 */
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) {
                const float d = (float)a[y * width + x] - M1;
                l[x & 3] = (y == 0 && x < 4) ? d * d : fmaf(d, d, l[x & 3]);
            }
        D = (l[0] + l[1] + l[2] + l[3]) / area;
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) {
                const float d = (float)b[y * width + x] - M2;
                l[x & 3] = (y == 0 && x < 4) ? d * d : fmaf(d, d, l[x & 3]);
            }
        D += (l[0] + l[1] + l[2] + l[3]) / area;
        /* covariance: separately rounded multiply then add (IL += d1*d2) */
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) {
                const float p = ((float)a[y * width + x] - M1) * ((float)b[y * width + x] - M2);
                l[x & 3] = (y == 0 && x < 4) ? p : l[x & 3] + p;
            }
        C = (l[0] + l[1] + l[2] + l[3]) / area;
        C = fmaf(M1, M2 * 2, c1) * fmaf(C, 2, c2) / (fmaf(M1, M1, fmaf(M2, M2, c1)) * (D + c2));
        D = M1 - M2;
        D = D < 0 ? -D : D;
        D = D > 4 ? 0.02f * D : 0.0f;
        metric[mb] = C - D;
    }
}

void vp8o_gather_SSIM(const float *m1, const float *m2, const float *m3, float *MB_SSIM, int mb_count) {
    for (int mb = 0; mb < mb_count; ++mb) MB_SSIM[mb] = (m1[mb] + m2[mb] + m3[mb]) / 3;
}

/* ---------------------------------------------------------------------------------------- */
void vp8o_prepare_filter_mask(const int16_t *MB, int32_t *MB_non_zero_coeffs, const int32_t *MB_parts,
                              int32_t *mb_mask, int width, int height) {
    const int mb_count = (width / 16) * (height / 16);
#pragma omp parallel for
    for (int mb = 0; mb < mb_count; ++mb) {
        const int16_t *m = MB + (size_t)mb * 400;
        int coeffs = 0;
        for (int b = 0; b < 16; ++b)
            for (int i = 1; i < 16; ++i) coeffs += iabs(m[b * 16 + i]);
        for (int b = 16; b < 24; ++b)
            for (int i = 0; i < 16; ++i) coeffs += iabs(m[b * 16 + i]);
        if (MB_parts[mb] == VP8O_ARE16x16) {
            for (int i = 0; i < 16; ++i) coeffs += iabs(m[24 * 16 + i]);
        } else {
            for (int b = 0; b < 16; ++b) coeffs += iabs(m[b * 16]);
        }
        MB_non_zero_coeffs[mb] = coeffs;
        mb_mask[mb] = (MB_parts[mb] != VP8O_ARE16x16 || coeffs > 0) ? -1 : 0;
    }
}

/* int16 lane helpers: the reference filters in short8 vectors */
static inline int16_t s16(int v) { return (int16_t)v; }
static inline uint16_t uabs16(int16_t v) { return (uint16_t)(v < 0 ? -v : v); }
static inline int16_t c128(int16_t v) { return v < -128 ? -128 : (v > 127 ? 127 : v); }

static inline int lf_mask(int16_t p3, int16_t p2, int16_t p1, int16_t p0, int16_t q0, int16_t q1, int16_t q2,
                          int16_t q3, uint16_t e_lim, uint16_t i_lim) {
    int m = uabs16(s16(p3 - p2)) > i_lim;
    m |= uabs16(s16(p2 - p1)) > i_lim;
    m |= uabs16(s16(p1 - p0)) > i_lim;
    m |= uabs16(s16(q1 - q0)) > i_lim;
    m |= uabs16(s16(q2 - q1)) > i_lim;
    m |= uabs16(s16(q3 - q2)) > i_lim;
    m |= (uint16_t)(uabs16(s16(p0 - q0)) * 2 + uabs16(s16(p1 - q1)) / 2) > e_lim;
    return !m;
}

/* src/CPU_kernels.cl:829-883, one lane.  p3/q3 are read only */
static inline void filter_mb_edge(int16_t p3, int16_t *p2, int16_t *p1, int16_t *p0, int16_t *q0, int16_t *q1,
                                  int16_t *q2, int16_t q3, uint16_t mb_lim, uint16_t int_lim, uint16_t hev_thr) {
    const int mask = lf_mask(p3, *p2, *p1, *p0, *q0, *q1, *q2, q3, mb_lim, int_lim);
    const int hev = (uabs16(s16(*p1 - *p0)) > hev_thr) | (uabs16(s16(*q1 - *q0)) > hev_thr);
    int16_t w = c128(s16(*p1 - *q1));
    w = c128(s16(w + s16(s16(*q0 - *p0) * 3)));
    if (!mask) w = 0;
    int16_t a = hev ? w : 0;
    const int16_t b = s16(c128(s16(a + 3)) >> 3);
    a = s16(c128(s16(a + 4)) >> 3);
    *q0 = s16(*q0 - a);
    *p0 = s16(*p0 + b);
    if (hev) w = 0;
    a = c128(s16(s16(w * 27 + 63) >> 7));
    *q0 = s16(*q0 - a);
    *p0 = s16(*p0 + a);
    a = c128(s16(s16(w * 18 + 63) >> 7));
    *q1 = s16(*q1 - a);
    *p1 = s16(*p1 + a);
    a = c128(s16(s16(w * 9 + 63) >> 7));
    *q2 = s16(*q2 - a);
    *p2 = s16(*p2 + a);
}

/* src/CPU_kernels.cl:885-926, one lane */
static inline void filter_b_edge(int16_t p3, int16_t p2, int16_t *p1, int16_t *p0, int16_t *q0, int16_t *q1,
                                 int16_t q2, int16_t q3, uint16_t b_lim, uint16_t int_lim, uint16_t hev_thr) {
    const int mask = lf_mask(p3, p2, *p1, *p0, *q0, *q1, q2, q3, b_lim, int_lim);
    const int hev = (uabs16(s16(*p1 - *p0)) > hev_thr) | (uabs16(s16(*q1 - *q0)) > hev_thr);
    int16_t a = c128(s16(*p1 - *q1));
    if (!hev) a = 0;
    a = c128(s16(a + s16(s16(*q0 - *p0) * 3)));
    if (!mask) a = 0;
    const int16_t b = s16(c128(s16(a + 3)) >> 3);
    a = s16(c128(s16(a + 4)) >> 3);
    *q0 = s16(*q0 - a);
    *p0 = s16(*p0 + b);
    a = s16(s16(a + 1) >> 1);
    if (hev) a = 0;
    *q1 = s16(*q1 - a);
    *p1 = s16(*p1 + a);
}

static inline uint8_t lf_store(int16_t v) { return (uint8_t)sat8(v + 128); }

void vp8o_loop_filter_frame(uint8_t *frame, const int32_t *MB_segment_ids, const int32_t *mb_mask,
                            const vp8o_segment_data *SD, int width, int height, int n) {
    const int mb_width = width / n, mb_count = mb_width * (height / n);
    for (int mb = 0; mb < mb_count; ++mb) {
        const vp8o_segment_data *sd = &SD[MB_segment_ids[mb]];
        if (sd->loop_filter_level == 0) return; /* ends the whole plane (Q6) */
        const uint16_t int_lim = (uint16_t)(int16_t)sd->interior_limit, mb_lim = (uint16_t)(int16_t)sd->mbedge_limit;
        const uint16_t b_lim = (uint16_t)(int16_t)sd->sub_bedge_limit, hev_thr = (uint16_t)(int16_t)sd->hev_threshold;
        const int x0 = (mb % mb_width) * n, y0 = (mb / mb_width) * n;
        /* pass over the vertical edges, then over the horizontal ones; 8 lanes at a time in
         * the reference, lanes are independent so a lane loop is exact.  Within a lane the
         * unclamped q0..q3 of one edge become p3..p0 of the next (Q7) */
        for (int dir = 0; dir < 2; ++dir) {
            const int along = dir == 0 ? 1 : width;  /* step across the edge */
            const int lane = dir == 0 ? width : 1;   /* step from lane to lane */
            const int has_edge = dir == 0 ? (x0 > 0) : (y0 > 0);
            for (int k = 0; k < n; ++k) {
                uint8_t *base = frame + (size_t)y0 * width + x0 + (size_t)k * lane;
                int16_t p3 = 0, p2 = 0, p1 = 0, p0 = 0, q0, q1, q2, q3;
                q0 = s16(base[0] - 128);
                q1 = s16(base[along] - 128);
                q2 = s16(base[2 * along] - 128);
                q3 = s16(base[3 * along] - 128);
                if (has_edge) {
                    p3 = s16(base[-4 * along] - 128);
                    p2 = s16(base[-3 * along] - 128);
                    p1 = s16(base[-2 * along] - 128);
                    p0 = s16(base[-1 * along] - 128);
                    filter_mb_edge(p3, &p2, &p1, &p0, &q0, &q1, &q2, q3, mb_lim, int_lim, hev_thr);
                    base[-3 * along] = lf_store(p2);
                    base[-2 * along] = lf_store(p1);
                    base[-1 * along] = lf_store(p0);
                    base[0] = lf_store(q0);
                    base[along] = lf_store(q1);
                    base[2 * along] = lf_store(q2);
                }
                for (int e = 4; e < n && mb_mask[mb]; e += 4) {
                    uint8_t *eb = base + (size_t)e * along;
                    p3 = q0;
                    p2 = q1;
                    p1 = q2;
                    p0 = q3;
                    if (p3 < -128 || p3 > 127 || p2 < -128 || p2 > 127 || p1 < -128 || p1 > 127) /* (p0 = q3 is never modified) */
                        quirk_event(1, dir == 0 ? x0 + e : x0 + k, dir == 0 ? y0 + k : y0 + e, n);
                    q0 = s16(eb[0] - 128);
                    q1 = s16(eb[along] - 128);
                    q2 = s16(eb[2 * along] - 128);
                    q3 = s16(eb[3 * along] - 128);
                    filter_b_edge(p3, p2, &p1, &p0, &q0, &q1, q2, q3, b_lim, int_lim, hev_thr);
                    eb[-2 * along] = lf_store(p1);
                    eb[-1 * along] = lf_store(p0);
                    eb[0] = lf_store(q0);
                    eb[along] = lf_store(q1);
                }
            }
        }
    }
}

/* ---- the host's per-frame reductions ------------------------------------------------------------
 * get_loopfilter_strength(), src/vp8enc.cpp:96-127, and the chroma differences of scene_change(),
 * src/vp8enc.cpp:265-285.  The reference accumulates in `int`; at the largest frame sizes those sums pass 2^31.
 * They are formed here in unsigned 32-bit arithmetic (what the compiled reference does) and read back as signed.
 * out4 = { reductor, sharpness, Udiff, Vdiff }. */
void vp8o_frame_statistics(const uint8_t *cur_y, int width, int height, const uint8_t *last_u, const uint8_t *cur_u,
                           const uint8_t *last_v, const uint8_t *cur_v, int32_t *out4) {
    if (cur_y) {
        const int n = width * height;
        uint32_t acc = 0;
        for (int i = 0; i < n; ++i) acc += cur_y[i];
        int avg = (int)(acc + (uint32_t)(n / 2));
        avg /= n;
        out4[0] = avg * 5 / 255 + 3;
        uint32_t div = 0;
        for (int i = 1; i < height - 1; ++i)
            for (int j = 1; j < width - 1; ++j) {
                const int p = i * width + j;
                int a = cur_y[p - width - 1] + cur_y[p - width] + cur_y[p - width + 1] + cur_y[p - 1] + cur_y[p + 1] +
                        cur_y[p + width - 1] + cur_y[p + width] + cur_y[p + width + 1];
                a /= 8;
                div += (uint32_t)((cur_y[p] - a) * (cur_y[p] - a));
            }
        int d = (int)(div + (uint32_t)((height - 1) * (width - 1) / 2));
        d /= (height - 1) * (width - 1);
        const int sh = d / 8;
        out4[1] = sh > 7 ? 7 : sh;
    }
    if (last_u && cur_u && last_v && cur_v) {
        const int n = (width / 2) * (height / 2);
        uint32_t du = 0, dv = 0;
        for (int i = 0; i < n; ++i) {
            du += (uint32_t)abs((int)last_u[i] - (int)cur_u[i]);
            dv += (uint32_t)abs((int)last_v[i] - (int)cur_v[i]);
        }
        out4[2] = (int)du / n;
        out4[3] = (int)dv / n;
    }
}

/* ---------------------------------------------------------------------------------------- */
struct vp8o_frame_ctx {
    int w, h, mb_count;
    uint8_t *cur_pyr[5], *last_pyr[5], *gold_pyr[5], *alt_pyr[5]; /* index k: downsampled by 2^k; [0] of last = recon */
    uint8_t *img[3][3];                                           /* [ref][plane] image objects */
    int16_t *net[3][2];
    int32_t *metrics[3];
    uint8_t *pred[3];
    int16_t *res[3];
    int16_t *coeffs;
};

vp8o_frame_ctx *vp8o_ctx_create(int w, int h) {
    vp8o_frame_ctx *c = (vp8o_frame_ctx *)calloc(1, sizeof(*c));
    c->w = w;
    c->h = h;
    c->mb_count = (w / 16) * (h / 16);
    for (int k = 0; k < 5; ++k) {
        const size_t sz = (size_t)(w >> k) * (h >> k);
        c->cur_pyr[k] = (uint8_t *)calloc(sz, 1);
        c->last_pyr[k] = (uint8_t *)calloc(sz, 1);
        c->gold_pyr[k] = (uint8_t *)calloc(sz, 1);
        c->alt_pyr[k] = (uint8_t *)calloc(sz, 1);
    }
    for (int r = 0; r < 3; ++r) {
        for (int p = 0; p < 3; ++p) c->img[r][p] = (uint8_t *)calloc((size_t)w * h / (p ? 4 : 1), 1);
        for (int k = 0; k < 2; ++k) c->net[r][k] = (int16_t *)calloc((size_t)c->mb_count * 4, 4);
        c->metrics[r] = (int32_t *)calloc((size_t)c->mb_count * 4, 4);
    }
    for (int p = 0; p < 3; ++p) {
        c->pred[p] = (uint8_t *)calloc((size_t)w * h / (p ? 4 : 1), 1);
        c->res[p] = (int16_t *)calloc((size_t)w * h / (p ? 4 : 1), 2);
    }
    c->coeffs = (int16_t *)calloc((size_t)c->mb_count * 400, 2);
    return c;
}

void vp8o_ctx_destroy(vp8o_frame_ctx *c) {
    if (!c) return;
    for (int k = 0; k < 5; ++k) {
        free(c->cur_pyr[k]);
        free(c->last_pyr[k]);
        free(c->gold_pyr[k]);
        free(c->alt_pyr[k]);
    }
    for (int r = 0; r < 3; ++r) {
        for (int p = 0; p < 3; ++p) free(c->img[r][p]);
        for (int k = 0; k < 2; ++k) free(c->net[r][k]);
        free(c->metrics[r]);
    }
    for (int p = 0; p < 3; ++p) {
        free(c->pred[p]);
        free(c->res[p]);
    }
    free(c->coeffs);
    free(c);
}

void vp8o_inter_frame(vp8o_frame_ctx *c, const uint8_t *cur_y, const uint8_t *cur_u, const uint8_t *cur_v,
                      uint8_t *recon_y, uint8_t *recon_u, uint8_t *recon_v, const vp8o_segment_data *SD,
                      float SSIM_target, int prev_is_golden, int prev_is_altref, int altref_differs_from_golden,
                      int16_t *MB_coeffs, int16_t *MB_vectors, int32_t *MB_parts, int32_t *MB_reference_frame,
                      int32_t *MB_segment_id, float *MB_SSIM) {
    const int w = c->w, h = c->h, M = c->mb_count;
    const size_t ysz = (size_t)w * h, csz = ysz / 4;
    const int use_golden = !prev_is_golden;
    const int use_altref = !prev_is_altref && altref_differs_from_golden; /* src/inter_part.h:103-104 */
    const uint8_t *cur[3] = {cur_y, cur_u, cur_v};
    uint8_t *recon[3] = {recon_y, recon_u, recon_v};

    /* src/vp8enc.cpp:386-401: uploads; last_pyr[0] plays reconstructed_frame_Y */
    memcpy(c->cur_pyr[0], cur_y, ysz);
    memcpy(c->last_pyr[0], recon_y, ysz);
    memcpy(c->img[VP8O_LAST][0], recon_y, ysz);
    memcpy(c->img[VP8O_LAST][1], recon_u, csz);
    memcpy(c->img[VP8O_LAST][2], recon_v, csz);

    /* src/inter_part.h:1-94 */
    vp8o_reset_vectors(c->net[0][0], c->net[0][1], c->net[1][0], c->net[1][1], c->net[2][0], c->net[2][1],
                       c->metrics[0], c->metrics[1], c->metrics[2], M * 4);
    for (int k = 0; k < 4; ++k) {
        vp8o_downsample_x2(c->last_pyr[k], c->last_pyr[k + 1], w >> k, h >> k);
        vp8o_downsample_x2(c->cur_pyr[k], c->cur_pyr[k + 1], w >> k, h >> k);
    }
    if (prev_is_golden) {
        for (int k = 0; k < 5; ++k) memcpy(c->gold_pyr[k], c->last_pyr[k], (size_t)(w >> k) * (h >> k));
        for (int p = 0; p < 3; ++p) memcpy(c->img[VP8O_GOLDEN][p], c->img[VP8O_LAST][p], p ? csz : ysz);
    }
    if (prev_is_altref) {
        for (int k = 0; k < 5; ++k) memcpy(c->alt_pyr[k], c->last_pyr[k], (size_t)(w >> k) * (h >> k));
        for (int p = 0; p < 3; ++p) memcpy(c->img[VP8O_ALTREF][p], c->img[VP8O_LAST][p], p ? csz : ysz);
    }

    /* src/inter_part.h:109-236: pyramid search; nets ping-pong 1->2, 2->1, 1->2, 2->1, 1->2, then qpel 2->1 (Q4) */
    uint8_t **pyr[3] = {c->last_pyr, c->gold_pyr, c->alt_pyr};
    const int use[3] = {1, use_golden, use_altref};
    for (int k = 4; k >= 0; --k) {
        const int src = (k & 1) ? 1 : 0; /* k=4: net1->net2, k=3: 2->1, ... k=0: 1->2 */
        for (int r = 0; r < 3; ++r)
            if (use[r])
                vp8o_luma_search_1step(c->cur_pyr[k], pyr[r][k], c->net[r][src], c->net[r][src ^ 1], (w / 16) * 2,
                                       w >> k, h >> k, 1 << k);
    }
    for (int r = 0; r < 3; ++r)
        if (use[r]) vp8o_luma_search_2step(cur_y, c->img[r][0], c->net[r][1], c->net[r][0], c->metrics[r], w, h);

    /* src/inter_part.h:250-266 */
    vp8o_select_reference(c->net[0][0], c->net[1][0], c->net[2][0], c->metrics[0], c->metrics[1], c->metrics[2],
                          MB_reference_frame, MB_vectors, w, h, use_golden, use_altref);
    vp8o_pack_8x8_into_16x16(MB_vectors, MB_parts, MB_SSIM, M);

    /* src/inter_part.h:268-326 */
    for (int p = 0; p < 3; ++p)
        for (int r = 0; r < 3; ++r)
            if (use[r])
                vp8o_prepare_predictors_and_residual(cur[p], c->img[r][p], c->pred[p], c->res[p], MB_reference_frame,
                                                     MB_vectors, p ? w / 2 : w, p ? h / 2 : h, p, r);

    /* src/inter_part.h:329-378; metrics1..3 are reused as float scratch (src/init.h:1098,1111,1118) */
    for (int s = 3; s >= 0; --s) {
        for (int p = 0; p < 3; ++p)
            vp8o_dct4x4(c->res[p], c->coeffs, MB_segment_id, MB_parts, MB_SSIM, p ? w / 2 : w, p ? h / 2 : h, SD, s,
                        SSIM_target, p);
        vp8o_wht4x4_iwht4x4(c->coeffs, MB_segment_id, MB_parts, SD, s, M);
        for (int p = 0; p < 3; ++p)
            vp8o_idct4x4(recon[p], c->pred[p], c->coeffs, MB_segment_id, MB_parts, p ? w / 2 : w, p ? h / 2 : h, SD, s, p);
        for (int p = 0; p < 3; ++p)
            vp8o_count_SSIM(cur[p], recon[p], MB_segment_id, (float *)c->metrics[p], p ? w / 2 : w, p ? h / 2 : h, s,
                            p ? 8 : 16);
        vp8o_gather_SSIM((float *)c->metrics[0], (float *)c->metrics[1], (float *)c->metrics[2], MB_SSIM, M);
    }
    memcpy(MB_coeffs, c->coeffs, (size_t)M * 800);
}

void vp8o_loop_filter_planes(uint8_t *y, uint8_t *u, uint8_t *v, const int16_t *MB_coeffs,
                             const int32_t *MB_parts, const int32_t *MB_segment_id, const vp8o_segment_data *SD,
                             int32_t *MB_non_zero_coeffs, int width, int height) {
    const int M = (width / 16) * (height / 16);
    int32_t *mask = (int32_t *)malloc((size_t)M * 4);
    vp8o_prepare_filter_mask(MB_coeffs, MB_non_zero_coeffs, MB_parts, mask, width, height);
#pragma omp parallel sections
    {
#pragma omp section
        vp8o_loop_filter_frame(y, MB_segment_id, mask, SD, width, height, 16);
#pragma omp section
        vp8o_loop_filter_frame(u, MB_segment_id, mask, SD, width / 2, height / 2, 8);
#pragma omp section
        vp8o_loop_filter_frame(v, MB_segment_id, mask, SD, width / 2, height / 2, 8);
    }
    free(mask);
}
