/* TEST INFRASTRUCTURE -- CPU restatement of the reference's intra (key-frame) macroblock path, SURVEY.md 8f-4.
 *
 * What it restates: intra_transform() / predict_and_transform_mb() and their helpers in the reference host,
 * src/intra_part.h:37-515 (iDCT4x4, DCT4x4, weight, quant4x4, pick_luma_predictor), :517-741 (the macroblock walk)
 * and :1089-1128 (the frame loop).  Pinned against the reference's own code compiled from where it lies
 * (oracle/ref_intra.cpp -> oracle/_ref/libref_intra.so, tests/test_oracle_vs_ref.py).  Only tests/, smoke() and
 * bench.py's cpu_baseline leg may use this file.
 *
 * Behaviour, in the reference's terms:
 *   - every macroblock is coded B_PRED: each of the sixteen 4x4 luma sub-blocks, in raster order, takes the one of
 *     the ten sub-block modes (enum order DC, TM, VE, HE, LD, RD, VR, VL, HD, HU) whose residual has the smallest
 *     weight(); ties go to the earlier mode.  Chroma is TM_PRED for the whole 8x8 block.
 *   - a sub-block is predicted from the RECONSTRUCTED pixels above (8 of them: above and above-right), to the left
 *     and above-left; outside the frame the row above reads 127, the column to the left 129, the corner 127 in the
 *     top macroblock row and 129 in the left column below it.  Sub-blocks of column 3 in rows 1-3 take their
 *     above-right pixels from the macroblock ABOVE (the four pixels right of it, or its last pixel repeated in the
 *     last macroblock column), as every VP8 decoder does.
 *   - residual -> forward DCT (16-bit stores) -> quantise with round-half-away ... except coefficient 11, whose
 *     rounding takes the sign of (the already rounded) coefficient 10 -> dequantise + inverse DCT (16-bit stores)
 *     + predictor, saturated -> reconstruction; the stored coefficients are the quantised ones in zig-zag order.
 *   - results: 25 x 16 coefficients per macroblock (block 24 untouched), 16 sub-block modes, parts = 2 (are4x4),
 *     segment id 0.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const int zigzag_of[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};

static inline int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* forward transform of a 4x4 residual, both passes with the reference's 16-bit stores (:124-158) */
static void fdct_intra(const int16_t r[16], int16_t out[16]) {
    int16_t t[16];
    for (int i = 0; i < 4; ++i) {
        const int a = (r[4 * i] + r[4 * i + 3]) << 3, b = (r[4 * i + 1] + r[4 * i + 2]) << 3;
        const int c = (r[4 * i + 1] - r[4 * i + 2]) << 3, d = (r[4 * i] - r[4 * i + 3]) << 3;
        t[4 * i] = (int16_t)(a + b);
        t[4 * i + 2] = (int16_t)(a - b);
        t[4 * i + 1] = (int16_t)((c * 2217 + d * 5352 + 14500) >> 12);
        t[4 * i + 3] = (int16_t)((d * 2217 - c * 5352 + 7500) >> 12);
    }
    for (int i = 0; i < 4; ++i) {
        const int a = t[i] + t[12 + i], b = t[4 + i] + t[8 + i], c = t[4 + i] - t[8 + i], d = t[i] - t[12 + i];
        out[i] = (int16_t)((a + b + 7) >> 4);
        out[8 + i] = (int16_t)((a - b + 7) >> 4);
        out[4 + i] = (int16_t)(((c * 2217 + d * 5352 + 12000) >> 16) + (d != 0));
        out[12 + i] = (int16_t)((d * 2217 - c * 5352 + 51000) >> 16);
    }
}

/* cost of a residual: sum of the magnitudes of its transform, the DC quartered (:159-210) */
static int weight_intra(const int16_t r[16]) {
    int16_t f[16];
    fdct_intra(r, f);
    f[0] = (int16_t)(f[0] / 4);
    int s = 0;
    for (int i = 0; i < 16; ++i) s += f[i] < 0 ? -f[i] : f[i];
    return s;
}

/* :212-250 -- rounding by half a step away from zero, then the truncating division, all in 16 bits;
 * coefficient 11 rounds in the direction of coefficient 10 (after coefficient 10 has been rounded) */
static void quant_intra(int16_t c[16], int dc_q, int ac_q) {
    for (int i = 0; i < 16; ++i) {
        const int q = i == 0 ? dc_q : ac_q;
        const int16_t sign_of = i == 11 ? c[10] : c[i];
        c[i] = (int16_t)(c[i] + (sign_of < 0 ? -q / 2 : q / 2));
    }
    for (int i = 0; i < 16; ++i) c[i] = (int16_t)(c[i] / (int16_t)(i == 0 ? dc_q : ac_q));
}

/* :40-122 -- dequantise, inverse transform (16-bit intermediate), add the predictor, saturate */
static void idct_intra(const int16_t c[16], const uint8_t pred[16], uint8_t out[16], int dc_q, int ac_q) {
    int16_t t[16];
    for (int i = 0; i < 4; ++i) {
        const int i0 = c[i] * (i == 0 ? dc_q : ac_q), i4 = c[4 + i] * ac_q, i8 = c[8 + i] * ac_q, i12 = c[12 + i] * ac_q;
        const int a = i0 + i8, b = i0 - i8;
        const int cc = ((i4 * 35468) >> 16) - (i12 + ((i12 * 20091) >> 16));
        const int d = (i4 + ((i4 * 20091) >> 16)) + ((i12 * 35468) >> 16);
        t[i] = (int16_t)(a + d);
        t[12 + i] = (int16_t)(a - d);
        t[4 + i] = (int16_t)(b + cc);
        t[8 + i] = (int16_t)(b - cc);
    }
    for (int i = 0; i < 4; ++i) {
        const int16_t *p = t + 4 * i;
        const int a = p[0] + p[2], b = p[0] - p[2];
        const int cc = ((p[1] * 35468) >> 16) - (p[3] + ((p[3] * 20091) >> 16));
        const int d = (p[1] + ((p[1] * 20091) >> 16)) + ((p[3] * 35468) >> 16);
        out[4 * i] = (uint8_t)clamp255((int16_t)(((a + d + 4) >> 3) + pred[4 * i]));
        out[4 * i + 3] = (uint8_t)clamp255((int16_t)(((a - d + 4) >> 3) + pred[4 * i + 3]));
        out[4 * i + 1] = (uint8_t)clamp255((int16_t)(((b + cc + 4) >> 3) + pred[4 * i + 1]));
        out[4 * i + 2] = (uint8_t)clamp255((int16_t)(((b - cc + 4) >> 3) + pred[4 * i + 2]));
    }
}

/* the ten sub-block predictors of RFC 6386 12.3 from A[0..7] (above, above-right), L[0..3] (left), P (above-left) */
static void subblock_predictor(int mode, const int16_t *A, const int16_t *L, int P, uint8_t B[16]) {
#define AVG3(x, y, z) (uint8_t)(((x) + 2 * (y) + (z) + 2) >> 2)
#define AVG2(x, y) (uint8_t)(((x) + (y) + 1) >> 1)
    switch (mode) {
        case 0: { /* B_DC_PRED */
            int v = 4;
            for (int i = 0; i < 4; ++i) v += A[i] + L[i];
            memset(B, (uint8_t)(v >> 3), 16);
            break;
        }
        case 1: /* B_TM_PRED */
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) B[4 * r + c] = (uint8_t)clamp255(A[c] + L[r] - P);
            break;
        case 2: /* B_VE_PRED */
            for (int c = 0; c < 4; ++c) {
                const uint8_t v = AVG3(c ? A[c - 1] : P, A[c], A[c + 1]);
                B[c] = B[4 + c] = B[8 + c] = B[12 + c] = v;
            }
            break;
        case 3: /* B_HE_PRED */
            for (int r = 0; r < 4; ++r) {
                const uint8_t v = r < 3 ? AVG3(r ? L[r - 1] : P, L[r], L[r + 1]) : AVG3(L[2], L[3], L[3]);
                B[4 * r] = B[4 * r + 1] = B[4 * r + 2] = B[4 * r + 3] = v;
            }
            break;
        case 4: /* B_LD_PRED: down-left diagonals of the row above */
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) {
                    const int k = r + c;
                    B[4 * r + c] = k < 6 ? AVG3(A[k], A[k + 1], A[k + 2]) : AVG3(A[6], A[7], A[7]);
                }
            break;
        case 5: { /* B_RD_PRED: down-right diagonals of the edge L3..L0, P, A0..A3 */
            const int E[9] = {L[3], L[2], L[1], L[0], P, A[0], A[1], A[2], A[3]};
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) {
                    const int k = 3 - r + c; /* 0..6 */
                    B[4 * r + c] = AVG3(E[k], E[k + 1], E[k + 2]);
                }
            break;
        }
        case 6: { /* B_VR_PRED */
            const int E[9] = {L[3], L[2], L[1], L[0], P, A[0], A[1], A[2], A[3]};
            B[12] = AVG3(E[1], E[2], E[3]);
            B[8] = AVG3(E[2], E[3], E[4]);
            B[13] = B[4] = AVG3(E[3], E[4], E[5]);
            B[9] = B[0] = AVG2(E[4], E[5]);
            B[14] = B[5] = AVG3(E[4], E[5], E[6]);
            B[10] = B[1] = AVG2(E[5], E[6]);
            B[15] = B[6] = AVG3(E[5], E[6], E[7]);
            B[11] = B[2] = AVG2(E[6], E[7]);
            B[7] = AVG3(E[6], E[7], E[8]);
            B[3] = AVG2(E[7], E[8]);
            break;
        }
        case 7: /* B_VL_PRED */
            B[0] = AVG2(A[0], A[1]);
            B[4] = AVG3(A[0], A[1], A[2]);
            B[8] = B[1] = AVG2(A[1], A[2]);
            B[12] = B[5] = AVG3(A[1], A[2], A[3]);
            B[9] = B[2] = AVG2(A[2], A[3]);
            B[13] = B[6] = AVG3(A[2], A[3], A[4]);
            B[10] = B[3] = AVG2(A[3], A[4]);
            B[14] = B[7] = AVG3(A[3], A[4], A[5]);
            B[11] = AVG3(A[4], A[5], A[6]);
            B[15] = AVG3(A[5], A[6], A[7]);
            break;
        case 8: { /* B_HD_PRED */
            const int E[9] = {L[3], L[2], L[1], L[0], P, A[0], A[1], A[2], A[3]};
            B[12] = AVG2(E[0], E[1]);
            B[13] = AVG3(E[0], E[1], E[2]);
            B[8] = B[14] = AVG2(E[1], E[2]);
            B[9] = B[15] = AVG3(E[1], E[2], E[3]);
            B[4] = B[10] = AVG2(E[2], E[3]);
            B[5] = B[11] = AVG3(E[2], E[3], E[4]);
            B[0] = B[6] = AVG2(E[3], E[4]);
            B[1] = B[7] = AVG3(E[3], E[4], E[5]);
            B[2] = AVG3(E[4], E[5], E[6]);
            B[3] = AVG3(E[5], E[6], E[7]);
            break;
        }
        default: /* B_HU_PRED */
            B[0] = AVG2(L[0], L[1]);
            B[1] = AVG3(L[0], L[1], L[2]);
            B[2] = B[4] = AVG2(L[1], L[2]);
            B[3] = B[5] = AVG3(L[1], L[2], L[3]);
            B[6] = B[8] = AVG2(L[2], L[3]);
            B[7] = B[9] = AVG3(L[2], L[3], L[3]);
            B[10] = B[11] = B[12] = B[13] = B[14] = B[15] = (uint8_t)L[3];
            break;
    }
#undef AVG3
#undef AVG2
}

/* quants = { y_dc_q, y_ac_q, uv_dc_q, uv_ac_q } of the intra segment (frames.y_dc_q[intra_segment] ...).
 * MB: 25 x 16 int16 per macroblock (zig-zag order; block 24 is left alone), modes: 16 int32 per macroblock. */
void vp8o_intra_frame(int width, int height, const uint8_t *cur_y, const uint8_t *cur_u, const uint8_t *cur_v,
                      uint8_t *rec_y, uint8_t *rec_u, uint8_t *rec_v, int16_t *MB, int32_t *modes, int32_t *parts,
                      int32_t *segment_id, const int32_t *quants) {
    const int mbw = width / 16, mbh = height / 16, cw = width / 2;
    for (int mb = 0; mb < mbw * mbh; ++mb) {
        const int mr = mb / mbw, mc = mb % mbw;
        int16_t *rec = MB + (size_t)mb * 400;
        parts[mb] = 2;
        segment_id[mb] = 0;
        /* ---- luma: the row above the macroblock (20 pixels), its left column, the corner ---- */
        int16_t above[20], left[16];
        for (int i = 0; i < 20; ++i) {
            if (mr == 0) above[i] = 127;
            else if (i < 16 || mc < mbw - 1) above[i] = rec_y[(size_t)(16 * mr - 1) * width + 16 * mc + i];
            else above[i] = above[15];
        }
        for (int i = 0; i < 16; ++i) left[i] = mc == 0 ? 129 : rec_y[(size_t)(16 * mr + i) * width + 16 * mc - 1];
        const int corner = mr == 0 ? 127 : (mc == 0 ? 129 : rec_y[(size_t)(16 * mr - 1) * width + 16 * mc - 1]);
        for (int b = 0; b < 16; ++b) {
            const int br = b >> 2, bc = b & 3, x0 = 16 * mc + 4 * bc, y0 = 16 * mr + 4 * br;
            /* neighbours of the sub-block: reconstructed pixels of this macroblock where they exist */
            int16_t A[8], L[4];
            int P;
            for (int i = 0; i < 8; ++i) {
                const int col = 4 * bc + i;
                if (br == 0 || col >= 16) A[i] = above[col];
                else A[i] = rec_y[(size_t)(y0 - 1) * width + 16 * mc + col];
            }
            for (int i = 0; i < 4; ++i) L[i] = bc == 0 ? left[4 * br + i] : rec_y[(size_t)(y0 + i) * width + x0 - 1];
            if (br == 0) P = bc == 0 ? corner : above[4 * bc - 1];
            else if (bc == 0) P = left[4 * br - 1];
            else P = rec_y[(size_t)(y0 - 1) * width + x0 - 1];
            uint8_t src[16], best_pred[16], pred[16];
            int16_t best_res[16], res[16];
            for (int i = 0; i < 16; ++i) src[i] = cur_y[(size_t)(y0 + (i >> 2)) * width + x0 + (i & 3)];
            int best_mode = 0, best_w = 0;
            for (int m = 0; m < 10; ++m) {
                subblock_predictor(m, A, L, P, pred);
                for (int i = 0; i < 16; ++i) res[i] = (int16_t)(src[i] - pred[i]);
                const int wgt = (int16_t)weight_intra(res);
                if (m == 0 || wgt < best_w) {
                    best_mode = m;
                    best_w = wgt;
                    memcpy(best_pred, pred, 16);
                    memcpy(best_res, res, sizeof(res));
                }
            }
            modes[16 * mb + b] = best_mode;
            int16_t c[16];
            uint8_t out[16];
            fdct_intra(best_res, c);
            quant_intra(c, quants[0], quants[1]);
            idct_intra(c, best_pred, out, quants[0], quants[1]);
            for (int i = 0; i < 16; ++i) rec_y[(size_t)(y0 + (i >> 2)) * width + x0 + (i & 3)] = out[i];
            for (int i = 0; i < 16; ++i) rec[16 * b + i] = c[zigzag_of[i]];
        }
        /* ---- chroma: TM_PRED from the macroblock's own border, per plane ---- */
        for (int pl = 0; pl < 2; ++pl) {
            const uint8_t *cur = pl ? cur_v : cur_u;
            uint8_t *rp = pl ? rec_v : rec_u;
            int16_t ca[8], cl[8];
            for (int i = 0; i < 8; ++i) {
                ca[i] = mr == 0 ? 127 : rp[(size_t)(8 * mr - 1) * cw + 8 * mc + i];
                cl[i] = mc == 0 ? 129 : rp[(size_t)(8 * mr + i) * cw + 8 * mc - 1];
            }
            const int cp = mr == 0 ? 127 : (mc == 0 ? 129 : rp[(size_t)(8 * mr - 1) * cw + 8 * mc - 1]);
            for (int b = 0; b < 4; ++b) {
                const int br = b >> 1, bc = b & 1, x0 = 8 * mc + 4 * bc, y0 = 8 * mr + 4 * br;
                uint8_t pred[16], out[16];
                int16_t res[16], c[16];
                for (int i = 0; i < 16; ++i) {
                    pred[i] = (uint8_t)clamp255(ca[4 * bc + (i & 3)] + cl[4 * br + (i >> 2)] - cp);
                    res[i] = (int16_t)(cur[(size_t)(y0 + (i >> 2)) * cw + x0 + (i & 3)] - pred[i]);
                }
                fdct_intra(res, c);
                quant_intra(c, quants[2], quants[3]);
                idct_intra(c, pred, out, quants[2], quants[3]);
                for (int i = 0; i < 16; ++i) rp[(size_t)(y0 + (i >> 2)) * cw + x0 + (i & 3)] = out[i];
                for (int i = 0; i < 16; ++i) rec[16 * (16 + 4 * pl + b) + i] = c[zigzag_of[i]];
            }
        }
    }
}
