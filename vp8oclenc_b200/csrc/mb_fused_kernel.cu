// Fused macroblock kernel of the vp8oclenc_b200 engine (sm_100a).
//
// One launch replaces the tail of inter_transform() (src/inter_part.h:268-378):
//   prepare_predictors_and_residual (Y,U,V x up to 3 references)          9 launches
//   for each of the 4 segments: dct4x4 x3, wht4x4_iwht4x4, idct4x4 x3,
//                               count_SSIM_luma, count_SSIM_chroma x2, gather_SSIM   44 launches
// That is legal because the SSIM-driven re-quantisation ladder is per macroblock:
// a macroblock is redone at the next finer quantiser only while ITS OWN SSIM is <= the target.
// Predictors and residuals never leave the SM (the host never reads those buffers); the 800-byte
// coefficient record, the reconstruction and the per-MB scalars are the only stores.
//
// Mapping: one warp per macroblock, one lane per 4x4 block (lanes 0-15 Y, 16-19 U, 20-23 V);
// the Y2 transform runs on lane 24; the float SSIM runs as 36 strictly ordered accumulation
// chains (the reference's float4 lane structure) spread over lanes 0-23.
// Arithmetic identical to transform_kernels.cu / the oracle; compiled with -fmad=false.
#include "common.cuh"

namespace vp8 {

struct FusedRefs {
    const uint8_t *img[3][3];  // [LAST/GOLDEN/ALTREF][Y/U/V]
};

constexpr int FUSED_WARPS = 4;

__device__ __forceinline__ void wht_bf(int a0, int a1, int a2, int a3, int &o0, int &o1, int &o2, int &o3) {
    const int a = a0 + a3, b = a1 + a2, c = a1 - a2, d = a0 - a3;
    o0 = a + b; o1 = c + d; o2 = a - b; o3 = d - c;
}
// element r of wht_bf(a0, a1, a2, a3)
__device__ __forceinline__ int wht_pick(int a0, int a1, int a2, int a3, int r) {
    const int a = a0 + a3, b = a1 + a2, c = a1 - a2, d = a0 - a3;
    const int p = (r & 1) ? d : a, q = (r & 1) ? c : b;
    return (r & 2) ? p - q : p + q;
}
__device__ __forceinline__ void idct1d(int i0, int i1, int i2, int i3, int &o0, int &o1, int &o2, int &o3) {
    const int a1 = i0 + i2, b1 = i0 - i2;
    const int c1 = ((i1 * 35468) >> 16) - (i3 + ((i3 * 20091) >> 16));
    const int d1 = (i1 + ((i1 * 20091) >> 16)) + ((i3 * 35468) >> 16);
    o0 = a1 + d1; o3 = a1 - d1; o1 = b1 + c1; o2 = b1 - c1;
}

// six-tap prediction of one 4x4 block, "predictor flavour" (construct, src/GPU_kernels.cl:574-774):
// lines Y-2..Y+3 saturate, lines Y+4..Y+6 wrap (Q5)
__device__ __forceinline__ int dp4a_px_taps(uint32_t px, uint32_t taps, int acc) {  // unsigned pixels x signed taps
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(px), "r"(taps), "r"(acc));
    return d;
}
__device__ __forceinline__ void predict4x4(const uint8_t *__restrict__ ref, int w, int h, int ox, int oy, int fx, int fy,
                                           int (&pred)[16]) {
    int line[9][4];
    // Horizontal pass: the six taps of an output are two dp4a over funnel-shifted words of the window line.
    // The taps are packed as signed bytes; phase 0's single tap 128 wraps to -128, so its sum is taken negative.
    const uint32_t tlo = (uint32_t)(c_sixtap[fx][0] & 255) | ((uint32_t)(c_sixtap[fx][1] & 255) << 8) |
                         ((uint32_t)(c_sixtap[fx][2] & 255) << 16) | ((uint32_t)(c_sixtap[fx][3] & 255) << 24);
    const uint32_t thi = (uint32_t)(c_sixtap[fx][4] & 255) | ((uint32_t)(c_sixtap[fx][5] & 255) << 8);
    const int sgn = fx == 0 ? -1 : 1;
    // the 9x9 window needs no clamping and the three aligned words per line stay inside the row
    const bool inside = ox >= 2 && ox + 10 <= w && oy >= 2 && oy + 6 < h;
    const int off = (ox - 2) & 3;
    uint32_t wa[9], wb[9], wc[9];  // window bytes 0-3, 4-7, 8.. of every line
    if (inside) {
        const uint32_t *row = reinterpret_cast<const uint32_t *>(ref + (size_t)(oy - 2) * w + ((ox - 2) & ~3));
#pragma unroll
        for (int l = 0; l < 9; ++l) {  // 9 unaligned bytes out of three aligned words
            const uint32_t *rl = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(row) + (size_t)l * w);
            const uint32_t w0 = __ldg(rl), w1 = __ldg(rl + 1), w2 = __ldg(rl + 2);
            wa[l] = __funnelshift_r(w0, w1, 8 * off);
            wb[l] = __funnelshift_r(w1, w2, 8 * off);
            wc[l] = w2 >> (8 * off);
        }
    } else {
#pragma unroll
        for (int l = 0; l < 9; ++l) {
            const uint8_t *row = ref + (size_t)clampi(oy - 2 + l, 0, h - 1) * w;
            uint32_t p[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) p[c] = __ldg(row + clampi(ox - 2 + c, 0, w - 1));
            wa[l] = p[0] | (p[1] << 8) | (p[2] << 16) | (p[3] << 24);
            wb[l] = p[4] | (p[5] << 8) | (p[6] << 16) | (p[7] << 24);
            wc[l] = p[8];
        }
    }
#pragma unroll
    for (int l = 0; l < 9; ++l) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t lo = c ? __funnelshift_r(wa[l], wb[l], 8 * c) : wa[l];
            const uint32_t hi = c ? __funnelshift_r(wb[l], wc[l], 8 * c) : wb[l];
            const int s = 64 + sgn * dp4a_px_taps(hi, thi, dp4a_px_taps(lo, tlo, 0));
            // s / 128 truncates: under the saturation that is s >> 7 (both give 0 for negative s), under the
            // wrap of lines 6..8 it is not
            line[l][c] = (l < 6) ? sat8(s >> 7) : (((s + ((s >> 31) & 127)) >> 7) & 255);
        }
    }
    const int u0 = c_sixtap[fy][0], u1 = c_sixtap[fy][1], u2 = c_sixtap[fy][2], u3 = c_sixtap[fy][3],
              u4 = c_sixtap[fy][4], u5 = c_sixtap[fy][5];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int s = 64 + u0 * line[r][c] + u1 * line[r + 1][c] + u2 * line[r + 2][c] + u3 * line[r + 3][c] +
                          u4 * line[r + 4][c] + u5 * line[r + 5][c];
            pred[4 * r + c] = sat8(s >> 7);
        }
}

#ifndef VP8_FUSED_MINCTAS
#define VP8_FUSED_MINCTAS 5
#endif
__global__ void __launch_bounds__(FUSED_WARPS * 32, VP8_FUSED_MINCTAS)
k_mb_fused(const uint8_t *__restrict__ cur_y, const uint8_t *__restrict__ cur_u, const uint8_t *__restrict__ cur_v,
           FusedRefs refs, const int *__restrict__ MB_ref, const short2 *__restrict__ MB_vec,
           const int *__restrict__ MB_parts, short *__restrict__ MB, int *__restrict__ MB_seg,
           float *__restrict__ MB_SSIM, uint8_t *__restrict__ rec_y, uint8_t *__restrict__ rec_u,
           uint8_t *__restrict__ rec_v, const vp8b200_segment_data *__restrict__ SD, float SSIM_target, int width,
           int height, int mb_count) {
    __shared__ __align__(16) uint8_t s_cur[FUSED_WARPS][384], s_rec[FUSED_WARPS][384];  // Y 16x16 | U 8x8 | V 8x8, row-major each
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * FUSED_WARPS + warp;
    if (mb >= mb_count) return;  // whole warp
    const int mbw = width >> 4;
    const int mbx = mb % mbw, mby = mb / mbw;

    // ---- lane -> block ----
    const int plane = lane < 16 ? 0 : (lane < 20 ? 1 : 2);
    const bool has_block = lane < 24;
    const int k = lane < 16 ? lane : (lane - 16) & 3;              // block index inside its plane
    const int bpr = plane == 0 ? 4 : 2;                            // blocks per row of the plane's MB
    const int bx = k % bpr, by = k / bpr;
    const int pw = plane == 0 ? width : width >> 1, ph = plane == 0 ? height : height >> 1;
    const int n = plane == 0 ? 16 : 8;
    const int px = mbx * n + bx * 4, py = mby * n + by * 4;        // pixel position in the plane
    const uint8_t *cur = plane == 0 ? cur_y : (plane == 1 ? cur_u : cur_v);
    uint8_t *rec = plane == 0 ? rec_y : (plane == 1 ? rec_u : rec_v);
    const int tile0 = plane == 0 ? 0 : (plane == 1 ? 256 : 320);   // offset of the plane's tile in s_cur/s_rec
    const int parts = MB_parts[mb];

    // ---- prediction + residual (prepare_predictors_and_residual, src/GPU_kernels.cl:1285-1344) ----
    int res[16], pred[16];
    if (has_block) {
        const int ref = MB_ref[mb];
        const int q = plane == 0 ? (by >> 1) * 2 + (bx >> 1) : by * 2 + bx;
        const short2 v = MB_vec[4 * mb + q];
        const int g = plane == 0 ? 4 : 8;
        const int tx = px * g + v.x, ty = py * g + v.y;
        const int fx = max((tx % g) * (plane == 0 ? 2 : 1), 0), fy = max((ty % g) * (plane == 0 ? 2 : 1), 0);
        predict4x4(refs.img[ref][plane], pw, ph, tx / g, ty / g, fx, fy, pred);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t cw = *reinterpret_cast<const uint32_t *>(cur + (size_t)(py + r) * pw + px);
            *reinterpret_cast<uint32_t *>(&s_cur[warp][tile0 + (by * 4 + r) * n + bx * 4]) = cw;
#pragma unroll
            for (int c = 0; c < 4; ++c) res[4 * r + c] = (int)((cw >> (8 * c)) & 255) - pred[4 * r + c];
        }
    }
    __syncwarp();

    // ---- the SSIM ladder: LQ -> AQ -> HQ -> UQ while this macroblock's SSIM is <= the target ----
    float ssim = -2.0f;  // pack_8x8_into_16x16 left -2 in MB_SSIM
    int seg_written = -1;
    for (int seg = 3; seg >= 0; --seg) {
        if (ssim > SSIM_target) continue;  // dct4x4's early return, src/GPU_kernels.cl:1391
        seg_written = seg;
        const Quants Q = derive_quants(SD, seg);
        const int dc_q = plane == 0 ? (parts == ARE16x16 ? 1 : Q.y_dc) : Q.uv_dc;
        const int ac_q = plane == 0 ? Q.y_ac : Q.uv_ac;
        const uint32_t mdc = magic(dc_q), mac = magic(ac_q), my2_dc = magic(Q.y2_dc), my2_ac = magic(Q.y2_ac);
        int coef[16];  // quantised, raster order
        if (has_block) {
            int o[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int a1 = (res[c] + res[12 + c]) << 3, d1 = (res[c] - res[12 + c]) << 3;
                const int b1 = (res[4 + c] + res[8 + c]) << 3, c1 = (res[4 + c] - res[8 + c]) << 3;
                o[c] = a1 + b1;
                o[8 + c] = a1 - b1;
                o[4 + c] = (c1 * 2217 + d1 * 5352 + 14500) >> 12;
                o[12 + c] = (d1 * 2217 - c1 * 5352 + 7500) >> 12;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int a1 = o[4 * r] + o[4 * r + 3], d1 = o[4 * r] - o[4 * r + 3];
                const int b1 = o[4 * r + 1] + o[4 * r + 2], c1 = o[4 * r + 1] - o[4 * r + 2];
                coef[4 * r + 0] = div_magic((a1 + b1 + 7) >> 4, r == 0 ? mdc : mac);
                coef[4 * r + 2] = div_magic((a1 - b1 + 7) >> 4, mac);
                coef[4 * r + 1] = div_magic(((c1 * 2217 + d1 * 5352 + 12000) >> 16) + (d1 != 0), mac);
                coef[4 * r + 3] = div_magic((d1 * 2217 - c1 * 5352 + 51000) >> 16, mac);
            }
            coef[0] = (int)(short)coef[0];  // the record stores int16
        }
        // Y2: WHT of the 16 luma DCs, reconstructed DCs go back into the luma blocks (Q9).  Lane i < 16 holds the
        // DC of luma block i and computes element i of every butterfly pass from its row / column, fetched with
        // shuffles (parts is warp-uniform, all lanes take part; lanes >= 16 compute values nobody uses).
        if (parts == ARE16x16) {
            const int col = lane & 3, row4 = lane & 12, r_out = (lane >> 2) & 3;
            int v = lane < 16 ? coef[0] : 0;
            v = wht_pick(__shfl_sync(0xffffffffu, v, col), __shfl_sync(0xffffffffu, v, col + 4),
                         __shfl_sync(0xffffffffu, v, col + 8), __shfl_sync(0xffffffffu, v, col + 12), r_out);
            v = wht_pick(__shfl_sync(0xffffffffu, v, row4), __shfl_sync(0xffffffffu, v, row4 + 1),
                         __shfl_sync(0xffffffffu, v, row4 + 2), __shfl_sync(0xffffffffu, v, row4 + 3), col);
            const bool dc = (lane & 15) == 0;
            const int qq = dc ? Q.y2_dc : Q.y2_ac;
            v += (v > 0);
            v >>= 1;
            v = div_magic(v, dc ? my2_dc : my2_ac);
            if (lane < 16) MB[(size_t)mb * 400 + 24 * 16 + ((0xFEA9DB83C7426510ULL >> (4 * lane)) & 15)] = (short)v;
            v *= qq;
            v = wht_pick(__shfl_sync(0xffffffffu, v, row4), __shfl_sync(0xffffffffu, v, row4 + 1),
                         __shfl_sync(0xffffffffu, v, row4 + 2), __shfl_sync(0xffffffffu, v, row4 + 3), col);
            v = wht_pick(__shfl_sync(0xffffffffu, v, col), __shfl_sync(0xffffffffu, v, col + 4),
                         __shfl_sync(0xffffffffu, v, col + 8), __shfl_sync(0xffffffffu, v, col + 12), r_out);
            if (lane < 16) coef[0] = (int)(short)((v + 3) >> 3);
        }
        if (has_block) {
            // the coefficient block in zig-zag position order, two 16-byte stores
            __align__(16) short out[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) out[inv_zz(i)] = (short)coef[i];
            const int blk = plane == 0 ? k : (plane == 1 ? 16 + k : 20 + k);
            int4 *d4 = reinterpret_cast<int4 *>(MB + (size_t)mb * 400 + blk * 16);
            d4[0] = reinterpret_cast<const int4 *>(out)[0];
            d4[1] = reinterpret_cast<const int4 *>(out)[1];
            // dequantise + inverse DCT + prediction (idct4x4, src/GPU_kernels.cl:1545-1608)
            int L[16], t[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) L[i] = (int)(short)coef[i] * (i == 0 ? dc_q : ac_q);
#pragma unroll
            for (int c = 0; c < 4; ++c) idct1d(L[c], L[4 + c], L[8 + c], L[12 + c], t[c], t[4 + c], t[8 + c], t[12 + c]);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int o0, o1, o2, o3;
                idct1d(t[4 * r], t[4 * r + 1], t[4 * r + 2], t[4 * r + 3], o0, o1, o2, o3);
                const uint32_t wv = (uint32_t)sat8(((o0 + 4) >> 3) + pred[4 * r]) |
                                    ((uint32_t)sat8(((o1 + 4) >> 3) + pred[4 * r + 1]) << 8) |
                                    ((uint32_t)sat8(((o2 + 4) >> 3) + pred[4 * r + 2]) << 16) |
                                    ((uint32_t)sat8(((o3 + 4) >> 3) + pred[4 * r + 3]) << 24);
                *reinterpret_cast<uint32_t *>(rec + (size_t)(py + r) * pw + px) = wv;
                *reinterpret_cast<uint32_t *>(&s_rec[warp][tile0 + (by * 4 + r) * n + bx * 4]) = wv;
            }
        }
        __syncwarp();

        // ---- SSIM of the three planes (count_SSIM_luma/chroma + gather_SSIM) ----
        // exact integer sums for the means
        int sc = 0, sr = 0;
        if (has_block) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint32_t a = *reinterpret_cast<const uint32_t *>(&s_cur[warp][tile0 + (by * 4 + r) * n + bx * 4]);
                const uint32_t b = *reinterpret_cast<const uint32_t *>(&s_rec[warp][tile0 + (by * 4 + r) * n + bx * 4]);
                sc = (int)__dp4a(a, 0x01010101u, (unsigned)sc);  // sum of the four bytes, one instruction
                sr = (int)__dp4a(b, 0x01010101u, (unsigned)sr);
            }
        }
        // segmented reductions: lanes 0-15 | 16-19 | 20-23
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
            const int oc = __shfl_xor_sync(0xffffffffu, sc, o), orr = __shfl_xor_sync(0xffffffffu, sr, o);
            if (lane < 16 || o <= 2) {
                sc += oc;
                sr += orr;
            }
        }
        const int sumc[3] = {__shfl_sync(0xffffffffu, sc, 0), __shfl_sync(0xffffffffu, sc, 16), __shfl_sync(0xffffffffu, sc, 20)};
        const int sumr[3] = {__shfl_sync(0xffffffffu, sr, 0), __shfl_sync(0xffffffffu, sr, 16), __shfl_sync(0xffffffffu, sr, 20)};
        float M1[3], M2[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            // (x / 256 and x / 64 are exact scalings: the correctly rounded quotient is the product with 2^-8 / 2^-6)
            const float inv_area = p == 0 ? 0.00390625f : 0.015625f;
            M1[p] = __fmul_rn((float)sumc[p], inv_area);
            M2[p] = __fmul_rn((float)sumr[p], inv_area);
        }
        // 36 ordered chains: quantity qn (0 var(cur), 1 var(rec), 2 cov) x float4 lane kk, per plane.
        // lanes 0-11: luma (64 elements each); lanes 12-23: U then V (16 elements each).
        // Chain kk takes every fourth byte of the row-major tile (i = 4m + kk), so a 16-byte load brings four of
        // its elements.  pixel - mean is formed exactly in one FADD: as_float(0x47000000 | pixel << 8) is
        // 32768 + pixel, and 32768 + mean is representable (the mean is a multiple of 1/256 below 256), so the
        // difference equals the reference's fsub(float(pixel), mean), which is exact as well.  The chain itself is
        // one FFMA per element: fma(d, d, acc) where the reference has mad, fma(d1*d2, 1, acc) == acc + d1*d2 where
        // it multiplies and adds.
        float acc0 = 0.0f, acc1 = 0.0f;
        if (lane < 24) {
            const bool luma = lane < 12;
            const int idx = luma ? lane : lane - 12;
            const int qn = idx >> 2, kk = idx & 3;
            const bool nonfused = qn == 2;
            const uint32_t sel = 0x7604u | ((uint32_t)kk << 4);   // byte kk of the word -> bits 8..15 of 0x47000000
            const uint8_t *tx = (qn == 1 ? s_rec[warp] : s_cur[warp]) + (luma ? 0 : 256);
            const uint8_t *ty = (qn == 0 ? s_cur[warp] : s_rec[warp]) + (luma ? 0 : 256);
            const int p0 = luma ? 0 : 1, p1 = luma ? 0 : 2;       // plane of element groups 0-3 and 4-7
            float kx = __fadd_rn(32768.0f, qn == 1 ? M2[p0] : M1[p0]), ky = __fadd_rn(32768.0f, qn == 0 ? M1[p0] : M2[p0]);
            float acc = 0.0f;
            auto group = [&](int g, bool first) {
                const uint4 wx = *reinterpret_cast<const uint4 *>(tx + 16 * g);
                const uint4 wy = *reinterpret_cast<const uint4 *>(ty + 16 * g);
                const uint32_t ax[4] = {wx.x, wx.y, wx.z, wx.w}, ay[4] = {wy.x, wy.y, wy.z, wy.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float x = __fsub_rn(__uint_as_float(__byte_perm(ax[e], 0x47000000u, sel)), kx);
                    const float y = __fsub_rn(__uint_as_float(__byte_perm(ay[e], 0x47000000u, sel)), ky);
                    const float pr = __fmul_rn(x, y);
                    if (first && e == 0) acc = pr;
                    else acc = __fmaf_rn(nonfused ? pr : x, nonfused ? 1.0f : y, acc);
                }
            };
#pragma unroll
            for (int g = 0; g < 4; ++g) group(g, g == 0);
            if (!luma) {  // chroma lanes: U is done, start V
                acc0 = acc;
                kx = __fadd_rn(32768.0f, qn == 1 ? M2[p1] : M1[p1]);
                ky = __fadd_rn(32768.0f, qn == 0 ? M1[p1] : M2[p1]);
            }
            {   // group 4: the first element restarts the chain on the chroma lanes only
                const uint4 wx = *reinterpret_cast<const uint4 *>(tx + 64);
                const uint4 wy = *reinterpret_cast<const uint4 *>(ty + 64);
                const uint32_t ax[4] = {wx.x, wx.y, wx.z, wx.w}, ay[4] = {wy.x, wy.y, wy.z, wy.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float x = __fsub_rn(__uint_as_float(__byte_perm(ax[e], 0x47000000u, sel)), kx);
                    const float y = __fsub_rn(__uint_as_float(__byte_perm(ay[e], 0x47000000u, sel)), ky);
                    const float pr = __fmul_rn(x, y);
                    const float nx = __fmaf_rn(nonfused ? pr : x, nonfused ? 1.0f : y, acc);
                    acc = (e == 0 && !luma) ? pr : nx;
                }
            }
#pragma unroll
            for (int g = 5; g < 8; ++g) group(g, false);
            if (luma) {
#pragma unroll
                for (int g = 8; g < 16; ++g) group(g, false);
                acc0 = acc;
            } else {
                acc1 = acc;
            }
        }
        // (s0+s1)+s2)+s3 inside every group of four lanes
        float tot0, tot1;
        {
            const int b4 = lane & ~3;
            const float a0 = __shfl_sync(0xffffffffu, acc0, b4), a1 = __shfl_sync(0xffffffffu, acc0, b4 + 1),
                        a2 = __shfl_sync(0xffffffffu, acc0, b4 + 2), a3 = __shfl_sync(0xffffffffu, acc0, b4 + 3);
            tot0 = __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);
            const float c0 = __shfl_sync(0xffffffffu, acc1, b4), c1 = __shfl_sync(0xffffffffu, acc1, b4 + 1),
                        c2 = __shfl_sync(0xffffffffu, acc1, b4 + 2), c3 = __shfl_sync(0xffffffffu, acc1, b4 + 3);
            tot1 = __fadd_rn(__fadd_rn(__fadd_rn(c0, c1), c2), c3);
        }
        float metric[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const int base = p == 0 ? 0 : 12;
            const float inv_area = p == 0 ? 0.00390625f : 0.015625f;  // exact, as above
            const float v1 = p == 2 ? __shfl_sync(0xffffffffu, tot1, base) : __shfl_sync(0xffffffffu, tot0, base);
            const float v2 = p == 2 ? __shfl_sync(0xffffffffu, tot1, base + 4) : __shfl_sync(0xffffffffu, tot0, base + 4);
            const float cv = p == 2 ? __shfl_sync(0xffffffffu, tot1, base + 8) : __shfl_sync(0xffffffffu, tot0, base + 8);
            float D = __fmul_rn(v1, inv_area);
            D = __fadd_rn(D, __fmul_rn(v2, inv_area));
            float C = __fmul_rn(cv, inv_area);
            const float c1 = __fmul_rn(__fmul_rn(__fmul_rn(0.01f, 0.01f), 255.0f), 255.0f);
            const float c2 = __fmul_rn(__fmul_rn(__fmul_rn(0.03f, 0.03f), 255.0f), 255.0f);
            const float num = __fmul_rn(__fmaf_rn(M1[p], __fmul_rn(M2[p], 2.0f), c1), __fmaf_rn(C, 2.0f, c2));
            const float den = __fmul_rn(__fmaf_rn(M1[p], M1[p], __fmaf_rn(M2[p], M2[p], c1)), __fadd_rn(D, c2));
            C = __fdiv_rn(num, den);
            float dm = __fsub_rn(M1[p], M2[p]);
            dm = dm < 0.0f ? -dm : dm;
            dm = dm > 4.0f ? __fmul_rn(0.02f, dm) : 0.0f;
            metric[p] = __fsub_rn(C, dm);
        }
        ssim = __fdiv_rn(__fadd_rn(__fadd_rn(metric[0], metric[1]), metric[2]), 3.0f);
        __syncwarp();
    }
    if (lane == 0) {
        MB_SSIM[mb] = ssim;
        if (seg_written >= 0) MB_seg[mb] = seg_written;
    }
}

}  // namespace vp8

using namespace vp8;

// Fused replacement of the 53 launches of src/inter_part.h:268-378.  img[r][p] may be NULL for
// references that are not in use this frame.  Requires SSIM_target >= -2 (always true for the
// reference's command line: the default is -1, -SSIM-target gives 0.xx).
extern "C" int vp8b200_mb_predict_transform_fused(void *stream, const uint8_t *cur_y, const uint8_t *cur_u,
                                                  const uint8_t *cur_v, const uint8_t *const img[9],
                                                  const int32_t *MB_reference_frame, const int16_t *MB_vectors,
                                                  const int32_t *MB_parts, int16_t *MB, int32_t *MB_segment_id,
                                                  float *MB_SSIM, uint8_t *recon_y, uint8_t *recon_u, uint8_t *recon_v,
                                                  const vp8b200_segment_data *SD, float SSIM_target, int width,
                                                  int height) {
    const int M = (width / 16) * (height / 16);
    if (M <= 0) return 0;
    if (!(SSIM_target >= -2.0f)) return -(int)cudaErrorInvalidValue;
    FusedRefs r;
    for (int i = 0; i < 9; ++i) r.img[i / 3][i % 3] = img[i];
    k_mb_fused<<<(M + FUSED_WARPS - 1) / FUSED_WARPS, FUSED_WARPS * 32, 0, (cudaStream_t)stream>>>(
        cur_y, cur_u, cur_v, r, MB_reference_frame, (const short2 *)MB_vectors, MB_parts, MB, MB_segment_id, MB_SSIM,
        recon_y, recon_u, recon_v, SD, SSIM_target, width, height, M);
    VP8_LAUNCH_CHECK();
}
