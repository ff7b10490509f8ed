// Shared device helpers of the vp8oclenc_b200 CUDA engine (sm_100a).
//
// Arithmetic contract: 32-bit two's-complement integers,
// arithmetic >>, truncating /.  Every quirk of the reference kernels that is observable in
// motion vectors, coefficients or reconstructed pixels is reproduced and cited where it lives.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vp8b200.h"

#define VP8_LAUNCH_CHECK() \
    do {                   \
        cudaError_t e_ = cudaGetLastError(); \
        return e_ == cudaSuccess ? 0 : -(int)e_; \
    } while (0)

namespace vp8 {

enum { ARE16x16 = 0, ARE8x8 = 1, ARE4x4 = 2 };
enum { LAST = 0, GOLDEN = 1, ALTREF = 2 };

// VP8 quantiser tables (RFC 6386 14.1; reference copy: src/GPU_kernels.cl:58-80)
static __constant__ const short c_dc_q[128] = {
    4,   5,   6,   7,   8,   9,   10,  10,  11,  12,  13,  14,  15,  16,  17,  17,  18,  19,  20,  20,  21,  21,
    22,  22,  23,  23,  24,  25,  25,  26,  27,  28,  29,  30,  31,  32,  33,  34,  35,  36,  37,  37,  38,  39,
    40,  41,  42,  43,  44,  45,  46,  46,  47,  48,  49,  50,  51,  52,  53,  54,  55,  56,  57,  58,  59,  60,
    61,  62,  63,  64,  65,  66,  67,  68,  69,  70,  71,  72,  73,  74,  75,  76,  76,  77,  78,  79,  80,  81,
    82,  83,  84,  85,  86,  87,  88,  89,  91,  93,  95,  96,  98,  100, 101, 102, 104, 106, 108, 110, 112, 114,
    116, 118, 122, 124, 126, 128, 130, 132, 134, 136, 138, 140, 143, 145, 148, 151, 154, 157};
static __constant__ const short c_ac_q[128] = {
    4,   5,   6,   7,   8,   9,   10,  11,  12,  13,  14,  15,  16,  17,  18,  19,  20,  21,  22,  23,  24,  25,
    26,  27,  28,  29,  30,  31,  32,  33,  34,  35,  36,  37,  38,  39,  40,  41,  42,  43,  44,  45,  46,  47,
    48,  49,  50,  51,  52,  53,  54,  55,  56,  57,  58,  60,  62,  64,  66,  68,  70,  72,  74,  76,  78,  80,
    82,  84,  86,  88,  90,  92,  94,  96,  98,  100, 102, 104, 106, 108, 110, 112, 114, 116, 119, 122, 125, 128,
    131, 134, 137, 140, 143, 146, 149, 152, 155, 158, 161, 164, 167, 170, 173, 177, 181, 185, 189, 193, 197, 201,
    205, 209, 213, 217, 221, 225, 229, 234, 239, 245, 249, 254, 259, 264, 269, 274, 279, 284};

// six-tap sub-pixel filters indexed by eighth-pel phase (RFC 6386 subpixel_filters;
// reference copy: src/GPU_kernels.cl:563-572)
static __constant__ const short c_sixtap[8][6] = {
    {0, 0, 128, 0, 0, 0},     {0, -6, 123, 12, -1, 0}, {2, -11, 108, 36, -8, 1}, {0, -9, 93, 50, -6, 0},
    {3, -16, 77, 77, -16, 3}, {0, -6, 50, 93, -9, 0},  {1, -8, 36, 108, -11, 2}, {0, -1, 12, 123, -6, 0}};

// position in the zig-zag scan of raster coefficient k (src/GPU_kernels.cl:1489)
static __constant__ const unsigned char c_inv_zigzag[16] = {0, 1, 5, 6, 2, 4, 7, 12, 3, 8, 11, 13, 9, 10, 14, 15};

// the same table as a packed literal, so that unrolled loops index it at compile time
__device__ __forceinline__ constexpr int inv_zz(int k) { return (int)((0xFEA9DB83C7426510ULL >> (4 * k)) & 15); }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
// clamp to 0..255 in one instruction: VIMNMX.RELU = max(min(v, 255), 0)
__device__ __forceinline__ int sat8(int v) { return __vimin_s32_relu(v, 255); }

// Block cost of a 4x4 residual r[row*4+col] (weight_opt, src/GPU_kernels.cl:85-190).
// Pass 1 runs down the columns with the reference's register clobber (Q1): the "b1" term is
// overwritten by c1 and the odd outputs use the RAW third row in place of c1.  Pass 2 is the
// regular VP8 second pass along the rows.  Result = |DC|/4 + sum of the other magnitudes.
//
// The reference scales the pass-1 sums by 8 ((r0+r3)<<3 ...).  Here the factor is folded into what
// follows, which is exact: d1*5352 = (r0-r3)*42816, d1*2217 = (r0-r3)*17736; rows 0 and 2 of the
// intermediate are 8*p with p = (r0+r3) +- (r1-r2), and for them (8m+7)>>4 == m>>1 for every
// integer m, (8c*2217 + 8d*5352 + k)>>16 == (c*17736 + d*42816 + k)>>16, (8d != 0) == (d != 0).
__device__ __forceinline__ int weight4x4_rows(const int (&o)[16]);
__device__ __forceinline__ int weight4x4(const int (&r)[16]) {
    int o[16];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int s = r[k] + r[12 + k], t = r[k] - r[12 + k], u = r[4 + k] - r[8 + k];
        const int x = r[8 + k];
        o[k] = s + u;      // = reference value / 8
        o[8 + k] = s - u;  // = reference value / 8
        o[4 + k] = (x * 2217 + t * 42816 + 14500) >> 12;
        o[12 + k] = (t * 17736 - x * 5352 + 7500) >> 12;
    }
    return weight4x4_rows(o);
}

// The same cost from per-column features, for searches whose candidates are full-pel shifts of one window
// (luma_search_1step).  Pass 1 of weight4x4 is linear in the residual before its shift, so it splits into a part
// of the current block and a part of the predictor column, and a predictor column (4 vertically adjacent window
// pixels) is shared by every candidate that covers it:
//   o[k]      = (s+u)(cur) - (s+u)(pred)                 s = p0+p3, t = p0-p3, u = p1-p2, x = p2
//   o[8+k]    = (s-u)(cur) - (s-u)(pred)
//   o[4+k]    = ((x*2217 + t*42816 + 14500)(cur) - (x*2217 + t*42816)(pred)) >> 12
//   o[12+k]   = ((t*17736 - x*5352 + 7500)(cur) - (t*17736 - x*5352)(pred)) >> 12
// (exact integer identities; the magnitudes stay below 2^25).
struct ColFeat {
    int e0, e8, a4, a12;
};
__device__ __forceinline__ ColFeat col_features(int p0, int p1, int p2, int p3, int k4, int k12) {
    const int s = p0 + p3, t = p0 - p3, u = p1 - p2;
    ColFeat f;
    f.e0 = s + u;
    f.e8 = s - u;
    f.a4 = p2 * 2217 + t * 42816 + k4;
    f.a12 = t * 17736 - p2 * 5352 + k12;
    return f;
}
// pass 2 of weight4x4 on the intermediate o (rows 0 and 2 are the reference's values / 8)
#ifndef VP8_COST_SAD
#define VP8_COST_SAD 1
#endif
__device__ __forceinline__ int weight4x4_rows(const int (&o)[16]) {
    int sum = 0, sum2 = 0;
#pragma unroll
    for (int row = 0; row < 4; ++row) {
        const int a = o[4 * row] + o[4 * row + 3], d = o[4 * row] - o[4 * row + 3];
        const int b = o[4 * row + 1] + o[4 * row + 2], c = o[4 * row + 1] - o[4 * row + 2];
        int f0, f1, f2, f3;
#if VP8_COST_SAD
        // |x| + acc is one VABSDIFF (x - 0, accumulate) instead of IABS + a share of an IADD3; the "+ (d != 0)" of f1
        // rides in the VABSDIFF's second operand: |t + (d != 0)| = |t - m| with m = d ? -1 : 0
        if (row & 1) {
            f0 = (a + b + 7) >> 4;
            f2 = (a - b + 7) >> 4;
            f1 = (c * 2217 + d * 5352 + 12000) >> 16;
            f3 = (d * 2217 - c * 5352 + 51000) >> 16;
        } else {
            f0 = (a + b) >> 1;
            f2 = (a - b) >> 1;
            f1 = (c * 17736 + d * 42816 + 12000) >> 16;
            f3 = (d * 17736 - c * 42816 + 51000) >> 16;
        }
        if (row == 0) sum += abs(f0) >> 2; else sum = __sad(f0, 0, sum);
        sum2 = __sad(f1, d != 0 ? -1 : 0, sum2);
        sum = __sad(f2, 0, sum);
        sum2 = __sad(f3, 0, sum2);
    }
    return sum + sum2;
#else
        if (row & 1) {
            f0 = (a + b + 7) >> 4;
            f2 = (a - b + 7) >> 4;
            f1 = ((c * 2217 + d * 5352 + 12000) >> 16) + (d != 0);
            f3 = (d * 2217 - c * 5352 + 51000) >> 16;
        } else {
            f0 = (a + b) >> 1;
            f2 = (a - b) >> 1;
            f1 = ((c * 17736 + d * 42816 + 12000) >> 16) + (d != 0);
            f3 = (d * 17736 - c * 42816 + 51000) >> 16;
        }
        sum += (row == 0 ? (abs(f0) >> 2) : abs(f0)) + abs(f1) + abs(f2) + abs(f3);
    }
    return sum;
#endif
}

// Truncating x / q through m = magic(q): |x| * m >> 32 with the sign put back.  m * q = 2^32 + e with 0 <= e < q,
// so the quotient is exact while |x| * e < 2^32; quantisers are <= 440 and the dividends below 2^18.
// q == 1 (m would be 2^32) is encoded as m == 0.
__device__ __forceinline__ uint32_t magic(int q) { return q == 1 ? 0u : 0xffffffffu / (uint32_t)q + 1u; }
__device__ __forceinline__ int div_magic(int x, uint32_t m) {
    const int r = (int)__umulhi((uint32_t)abs(x), m);
    return m == 0u ? x : (x < 0 ? -r : r);
}

// quantisers of one segment, derived exactly as the device code of the reference does it
// in three places (Q11; src/GPU_kernels.cl:1394-1408, 1515-1524, 1568-1582)
struct Quants {
    int y_dc, y_ac, uv_dc, uv_ac, y2_dc, y2_ac;
};
__device__ __forceinline__ Quants derive_quants(const vp8b200_segment_data *SD, int seg) {
    Quants q;
    const int i = SD[seg].y_ac_i;
    q.y_ac = c_ac_q[i];
    q.y_dc = c_dc_q[clampi(i + SD[0].y_dc_idelta, 0, 127)];
    q.uv_dc = min((int)c_dc_q[clampi(i + SD[0].uv_dc_idelta, 0, 127)], 132);
    q.uv_ac = c_ac_q[clampi(i + SD[0].uv_ac_idelta, 0, 127)];
    q.y2_dc = c_dc_q[clampi(i + SD[0].y2_dc_idelta, 0, 127)] * 2;
    q.y2_ac = max(31 * c_ac_q[clampi(i + SD[0].y2_ac_idelta, 0, 127)] / 20, 8);
    return q;
}

}  // namespace vp8
