// Prediction / transform kernels of the vp8oclenc_b200 engine (sm_100a):
//   prepare_predictors_and_residual (six-tap "predictor flavour"), dct4x4, wht4x4_iwht4x4,
//   idct4x4, count_SSIM (luma/chroma), gather_SSIM.
// These are memory-bound byte/int16 kernels; at the BASELINE frame sizes the whole working set
// lives in the 126 MB L2, so they are launch-latency bound until fused.
//
// This file is compiled with -fmad=false and the SSIM code uses explicit rounding intrinsics:
// the float path must round exactly where the reference source says.
#include "common.cuh"

namespace vp8 {

// ------------------------------------------------------------------------------------------
// one thread per 4x4 block (src/GPU_kernels.cl:1285-1344).  construct() (lines 574-774) filters
// source lines Y-2..Y+3 with saturation but stores lines Y+4..Y+6 with a WRAPPING (uchar) cast
// of the truncating-divided sum (Q5); the vertical results are saturated.
__global__ void k_prepare_predictors_and_residual(const uint8_t *__restrict__ cur, const uint8_t *__restrict__ ref,
                                                  uint8_t *__restrict__ predictor, short *__restrict__ residual,
                                                  const int *__restrict__ MB_ref, const short2 *__restrict__ MB_vec,
                                                  int width, int height, int plane, int ref_id) {
    const int bwid = width >> 2;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= bwid * (height >> 2)) return;
    const int x = (b % bwid) * 4, y = (b / bwid) * 4;
    const int mb_size = plane == 0 ? 16 : 8, g = plane == 0 ? 4 : 8;
    const int mb = (y / mb_size) * (width / mb_size) + x / mb_size;
    if (MB_ref[mb] != ref_id) return;
    const int q = ((y % mb_size) / (mb_size / 2)) * 2 + (x % mb_size) / (mb_size / 2);
    const short2 v = MB_vec[4 * mb + q];
    const int tx = x * g + v.x, ty = y * g + v.y;  // luma: quarter pels, chroma: the same vector as eighth pels
    // tx, ty >= 0 for every vector the search can emit; the reference would index its tap table
    // out of bounds otherwise, so that unreachable case is pinned to phase 0 (as in the oracle)
    const int fx = max((tx % g) * (plane == 0 ? 2 : 1), 0), fy = max((ty % g) * (plane == 0 ? 2 : 1), 0);
    const int ox = tx / g, oy = ty / g;

    int line[9][4];
#pragma unroll
    for (int l = 0; l < 9; ++l) {
        const uint8_t *row = ref + (size_t)clampi(oy - 2 + l, 0, height - 1) * width;
        int p[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) p[c] = __ldg(row + clampi(ox - 2 + c, 0, width - 1));
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int s = 64;
#pragma unroll
            for (int t = 0; t < 6; ++t) s += (int)c_sixtap[fx][t] * p[c + t];
            s /= 128;  // truncating
            line[l][c] = (l < 6) ? sat8(s) : (s & 255);
        }
    }
    const uint32_t cw[4] = {*reinterpret_cast<const uint32_t *>(cur + (size_t)(y + 0) * width + x),
                            *reinterpret_cast<const uint32_t *>(cur + (size_t)(y + 1) * width + x),
                            *reinterpret_cast<const uint32_t *>(cur + (size_t)(y + 2) * width + x),
                            *reinterpret_cast<const uint32_t *>(cur + (size_t)(y + 3) * width + x)};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        uint32_t pw = 0;
        short res[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int s = 64;
#pragma unroll
            for (int t = 0; t < 6; ++t) s += (int)c_sixtap[fy][t] * line[r + t][c];
            const int pv = sat8(s >> 7);
            pw |= (uint32_t)pv << (8 * c);
            res[c] = (short)((int)((cw[r] >> (8 * c)) & 255) - pv);
        }
        const size_t i = (size_t)(y + r) * width + x;
        *reinterpret_cast<uint32_t *>(predictor + i) = pw;
        *reinterpret_cast<short4 *>(residual + i) = make_short4(res[0], res[1], res[2], res[3]);
    }
}

// ------------------------------------------------------------------------------------------
// one thread per 4x4 block (src/GPU_kernels.cl:1368-1496)
__global__ void k_dct4x4(const short *__restrict__ residual, short *__restrict__ MB, int *__restrict__ MB_seg,
                         const int *__restrict__ MB_parts, const float *__restrict__ MB_SSIM, int width, int height,
                         const vp8b200_segment_data *__restrict__ SD, int segment_id, float SSIM_target, int plane) {
    const int bwid = width >> 2;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= bwid * (height >> 2)) return;
    const int x = (b % bwid) * 4, y = (b / bwid) * 4;
    const int mb_size = plane == 0 ? 16 : 8;
    const int mb = (y / mb_size) * (width / mb_size) + x / mb_size;
    if (MB_SSIM[mb] > SSIM_target) return;  // the SSIM ladder (Q10)
    MB_seg[mb] = segment_id;
    const Quants Q = derive_quants(SD, segment_id);
    const int dc_q = plane == 0 ? (MB_parts[mb] == ARE16x16 ? 1 : Q.y_dc) : Q.uv_dc;  // Q9
    const int ac_q = plane == 0 ? Q.y_ac : Q.uv_ac;
    int L[16], o[16];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const short4 v = *reinterpret_cast<const short4 *>(residual + (size_t)(y + r) * width + x);
        L[4 * r] = v.x; L[4 * r + 1] = v.y; L[4 * r + 2] = v.z; L[4 * r + 3] = v.w;
    }
    // first pass DOWN THE COLUMNS, second along the rows: the transpose of libvpx's order (Q8)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int a1 = (L[c] + L[12 + c]) << 3, d1 = (L[c] - L[12 + c]) << 3;
        const int b1 = (L[4 + c] + L[8 + c]) << 3, c1 = (L[4 + c] - L[8 + c]) << 3;
        o[c] = a1 + b1;
        o[8 + c] = a1 - b1;
        o[4 + c] = (c1 * 2217 + d1 * 5352 + 14500) >> 12;
        o[12 + c] = (d1 * 2217 - c1 * 5352 + 7500) >> 12;
    }
    int blk = ((y % mb_size) / 4) * (mb_size / 4) + (x % mb_size) / 4;
    blk += plane == 1 ? 16 : (plane == 2 ? 20 : 0);
    short *dst = MB + (size_t)mb * 400 + blk * 16;
    __align__(16) short out[16];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int a1 = o[4 * r] + o[4 * r + 3], d1 = o[4 * r] - o[4 * r + 3];
        const int b1 = o[4 * r + 1] + o[4 * r + 2], c1 = o[4 * r + 1] - o[4 * r + 2];
        const int f0 = (a1 + b1 + 7) >> 4;
        const int f2 = (a1 - b1 + 7) >> 4;
        const int f1 = ((c1 * 2217 + d1 * 5352 + 12000) >> 16) + (d1 != 0);
        const int f3 = (d1 * 2217 - c1 * 5352 + 51000) >> 16;
        // plain truncating division: no rounding, no zero bin
        out[inv_zz(4 * r + 0)] = (short)(f0 / (r == 0 ? dc_q : ac_q));
        out[inv_zz(4 * r + 1)] = (short)(f1 / ac_q);
        out[inv_zz(4 * r + 2)] = (short)(f2 / ac_q);
        out[inv_zz(4 * r + 3)] = (short)(f3 / ac_q);
    }
    // 32-byte block record, two 16-byte stores
    int4 *d4 = reinterpret_cast<int4 *>(dst);
    const int4 *o4 = reinterpret_cast<const int4 *>(out);
    d4[0] = o4[0];
    d4[1] = o4[1];
}

__device__ __forceinline__ void wht_butterfly(int a0, int a1, int a2, int a3, int &o0, int &o1, int &o2, int &o3) {
    const int a = a0 + a3, b = a1 + a2, c = a1 - a2, d = a0 - a3;
    o0 = a + b; o1 = c + d; o2 = a - b; o3 = d - c;
}

// one thread per macroblock (src/GPU_kernels.cl:257-401, 1498-1543): Y2 = WHT of the 16 luma DCs;
// the reconstructed DCs are written back into the luma blocks (Q9)
__global__ void k_wht4x4_iwht4x4(short *__restrict__ MB, const int *__restrict__ MB_seg,
                                 const int *__restrict__ MB_parts, const vp8b200_segment_data *__restrict__ SD,
                                 int segment_id, int mb_count) {
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if (mb >= mb_count) return;
    if (MB_seg[mb] != segment_id || MB_parts[mb] != ARE16x16) return;
    const Quants Q = derive_quants(SD, segment_id);
    short *m = MB + (size_t)mb * 400;
    int L[16], t[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) L[k] = m[k * 16];
#pragma unroll
    for (int c = 0; c < 4; ++c) wht_butterfly(L[c], L[4 + c], L[8 + c], L[12 + c], t[c], t[4 + c], t[8 + c], t[12 + c]);
#pragma unroll
    for (int r = 0; r < 4; ++r)
        wht_butterfly(t[4 * r], t[4 * r + 1], t[4 * r + 2], t[4 * r + 3], L[4 * r], L[4 * r + 1], L[4 * r + 2], L[4 * r + 3]);
    __align__(16) short y2[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int q = k == 0 ? Q.y2_dc : Q.y2_ac;
        int v = L[k];
        v += (v > 0);
        v >>= 1;
        v /= q;
        y2[inv_zz(k)] = (short)v;
        L[k] = v * q;
    }
    int4 *d4 = reinterpret_cast<int4 *>(m + 24 * 16);
    d4[0] = reinterpret_cast<const int4 *>(y2)[0];
    d4[1] = reinterpret_cast<const int4 *>(y2)[1];
#pragma unroll
    for (int r = 0; r < 4; ++r)
        wht_butterfly(L[4 * r], L[4 * r + 1], L[4 * r + 2], L[4 * r + 3], t[4 * r], t[4 * r + 1], t[4 * r + 2], t[4 * r + 3]);
#pragma unroll
    for (int c = 0; c < 4; ++c) wht_butterfly(t[c], t[4 + c], t[8 + c], t[12 + c], L[c], L[4 + c], L[8 + c], L[12 + c]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m[k * 16] = (short)((L[k] + 3) >> 3);
}

__device__ __forceinline__ void idct_1d(int i0, int i1, int i2, int i3, int &o0, int &o1, int &o2, int &o3) {
    const int a1 = i0 + i2, b1 = i0 - i2;
    const int c1 = ((i1 * 35468) >> 16) - (i3 + ((i3 * 20091) >> 16));
    const int d1 = (i1 + ((i1 * 20091) >> 16)) + ((i3 * 35468) >> 16);
    o0 = a1 + d1; o3 = a1 - d1; o1 = b1 + c1; o2 = b1 - c1;
}

// one thread per 4x4 block (src/GPU_kernels.cl:192-255, 1545-1608)
__global__ void k_idct4x4(uint8_t *__restrict__ recon, const uint8_t *__restrict__ predictor,
                          const short *__restrict__ MB, const int *__restrict__ MB_seg,
                          const int *__restrict__ MB_parts, int width, int height,
                          const vp8b200_segment_data *__restrict__ SD, int segment_id, int plane) {
    const int bwid = width >> 2;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= bwid * (height >> 2)) return;
    const int x = (b % bwid) * 4, y = (b / bwid) * 4;
    const int mb_size = plane == 0 ? 16 : 8;
    const int mb = (y / mb_size) * (width / mb_size) + x / mb_size;
    if (MB_seg[mb] != segment_id) return;
    const Quants Q = derive_quants(SD, segment_id);
    const int dc_q = plane == 0 ? (MB_parts[mb] == ARE16x16 ? 1 : Q.y_dc) : Q.uv_dc;
    const int ac_q = plane == 0 ? Q.y_ac : Q.uv_ac;
    int blk = ((y % mb_size) / 4) * (mb_size / 4) + (x % mb_size) / 4;
    blk += plane == 1 ? 16 : (plane == 2 ? 20 : 0);
    __align__(16) short in[16];
    const int4 *s4 = reinterpret_cast<const int4 *>(MB + (size_t)mb * 400 + blk * 16);
    reinterpret_cast<int4 *>(in)[0] = s4[0];
    reinterpret_cast<int4 *>(in)[1] = s4[1];
    int L[16], t[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) L[k] = (int)in[inv_zz(k)] * (k == 0 ? dc_q : ac_q);
#pragma unroll
    for (int c = 0; c < 4; ++c) idct_1d(L[c], L[4 + c], L[8 + c], L[12 + c], t[c], t[4 + c], t[8 + c], t[12 + c]);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int o0, o1, o2, o3;
        idct_1d(t[4 * r], t[4 * r + 1], t[4 * r + 2], t[4 * r + 3], o0, o1, o2, o3);
        const size_t i = (size_t)(y + r) * width + x;
        const uint32_t p = *reinterpret_cast<const uint32_t *>(predictor + i);
        const uint32_t w = (uint32_t)sat8(((o0 + 4) >> 3) + (int)(p & 255)) |
                           ((uint32_t)sat8(((o1 + 4) >> 3) + (int)((p >> 8) & 255)) << 8) |
                           ((uint32_t)sat8(((o2 + 4) >> 3) + (int)((p >> 16) & 255)) << 16) |
                           ((uint32_t)sat8(((o3 + 4) >> 3) + (int)(p >> 24)) << 24);
        *reinterpret_cast<uint32_t *>(recon + i) = w;
    }
}

// ------------------------------------------------------------------------------------------
// SSIM of one N x N macroblock per thread (src/GPU_kernels.cl:1610-2095).  The reference sums in
// float4 lanes (lane = column mod 4) in raster order, mad() is a fused multiply-add, everything
// else rounds after each operation; reproduced with explicit intrinsics.
template <int N>
__global__ void k_count_SSIM(const uint8_t *__restrict__ f1, const uint8_t *__restrict__ f2,
                             const int *__restrict__ MB_seg, float *__restrict__ metric, int width, int mb_count,
                             int segment_id) {
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if (mb >= mb_count) return;
    if (MB_seg[mb] != segment_id) return;
    const int mbw = width / N;
    const uint8_t *a = f1 + (size_t)(mb / mbw) * N * width + (mb % mbw) * N;
    const uint8_t *b = f2 + (size_t)(mb / mbw) * N * width + (mb % mbw) * N;
    const float area = (float)(N * N);
    float l[4];
    float M[2], D = 0.0f;
    // means (integer-valued float sums are exact, the order does not matter)
    for (int f = 0; f < 2; ++f) {
        const uint8_t *p = f ? b : a;
        unsigned s = 0;
        for (int y = 0; y < N; ++y)
            for (int x = 0; x < N; x += 4) {
                const uint32_t w = *reinterpret_cast<const uint32_t *>(p + (size_t)y * width + x);
                s += (w & 255) + ((w >> 8) & 255) + ((w >> 16) & 255) + (w >> 24);
            }
        M[f] = __fdiv_rn((float)s, area);
    }
    // variances: first group d*d, then mad(d,d,acc) (lines 1694-1697)
    for (int f = 0; f < 2; ++f) {
        const uint8_t *p = f ? b : a;
        for (int y = 0; y < N; ++y)
            for (int x = 0; x < N; x += 4) {
                const uint32_t w = *reinterpret_cast<const uint32_t *>(p + (size_t)y * width + x);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float d = __fsub_rn((float)((w >> (8 * k)) & 255), M[f]);
                    l[k] = (y == 0 && x == 0) ? __fmul_rn(d, d) : __fmaf_rn(d, d, l[k]);
                }
            }
        const float var = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(l[0], l[1]), l[2]), l[3]), area);
        D = f ? __fadd_rn(D, var) : var;
    }
    // covariance: separately rounded multiply, then add
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; x += 4) {
            const uint32_t wa = *reinterpret_cast<const uint32_t *>(a + (size_t)y * width + x);
            const uint32_t wb = *reinterpret_cast<const uint32_t *>(b + (size_t)y * width + x);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float pr = __fmul_rn(__fsub_rn((float)((wa >> (8 * k)) & 255), M[0]),
                                           __fsub_rn((float)((wb >> (8 * k)) & 255), M[1]));
                l[k] = (y == 0 && x == 0) ? pr : __fadd_rn(l[k], pr);
            }
        }
    float C = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(l[0], l[1]), l[2]), l[3]), area);
    const float c1 = __fmul_rn(__fmul_rn(__fmul_rn(0.01f, 0.01f), 255.0f), 255.0f);
    const float c2 = __fmul_rn(__fmul_rn(__fmul_rn(0.03f, 0.03f), 255.0f), 255.0f);
    const float num = __fmul_rn(__fmaf_rn(M[0], __fmul_rn(M[1], 2.0f), c1), __fmaf_rn(C, 2.0f, c2));
    const float den = __fmul_rn(__fmaf_rn(M[0], M[0], __fmaf_rn(M[1], M[1], c1)), __fadd_rn(D, c2));
    C = __fdiv_rn(num, den);
    float dm = __fsub_rn(M[0], M[1]);
    dm = dm < 0.0f ? -dm : dm;
    dm = dm > 4.0f ? __fmul_rn(0.02f, dm) : 0.0f;
    metric[mb] = __fsub_rn(C, dm);
}

__global__ void k_gather_SSIM(const float *__restrict__ m1, const float *__restrict__ m2, const float *__restrict__ m3,
                              float *__restrict__ MB_SSIM, int mb_count) {
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if (mb >= mb_count) return;
    MB_SSIM[mb] = __fdiv_rn(__fadd_rn(__fadd_rn(m1[mb], m2[mb]), m3[mb]), 3.0f);
}

}  // namespace vp8

using namespace vp8;

extern "C" int vp8b200_prepare_predictors_and_residual(void *stream, const uint8_t *cur, const uint8_t *ref,
                                                       uint8_t *predictor, int16_t *residual, const int32_t *MB_ref,
                                                       const int16_t *MB_vec, int width, int height, int plane,
                                                       int ref_id) {
    const int n = (width / 4) * (height / 4);
    if (n <= 0) return 0;
    k_prepare_predictors_and_residual<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        cur, ref, predictor, residual, MB_ref, (const short2 *)MB_vec, width, height, plane, ref_id);
    VP8_LAUNCH_CHECK();
}

extern "C" int vp8b200_dct4x4(void *stream, const int16_t *residual, int16_t *MB, int32_t *MB_seg,
                              const int32_t *MB_parts, const float *MB_SSIM, int width, int height,
                              const vp8b200_segment_data *SD, int segment_id, float SSIM_target, int plane) {
    const int n = (width / 4) * (height / 4);
    if (n <= 0) return 0;
    k_dct4x4<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(residual, MB, MB_seg, MB_parts, MB_SSIM, width, height,
                                                               SD, segment_id, SSIM_target, plane);
    VP8_LAUNCH_CHECK();
}

extern "C" int vp8b200_wht4x4_iwht4x4(void *stream, int16_t *MB, int32_t *MB_seg, const int32_t *MB_parts,
                                      const vp8b200_segment_data *SD, int segment_id, int mb_count) {
    if (mb_count <= 0) return 0;
    k_wht4x4_iwht4x4<<<(mb_count + 63) / 64, 64, 0, (cudaStream_t)stream>>>(MB, MB_seg, MB_parts, SD, segment_id,
                                                                           mb_count);
    VP8_LAUNCH_CHECK();
}

extern "C" int vp8b200_idct4x4(void *stream, uint8_t *recon, const uint8_t *predictor, const int16_t *MB,
                               const int32_t *MB_seg, const int32_t *MB_parts, int width, int height,
                               const vp8b200_segment_data *SD, int segment_id, int plane) {
    const int n = (width / 4) * (height / 4);
    if (n <= 0) return 0;
    k_idct4x4<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(recon, predictor, MB, MB_seg, MB_parts, width, height,
                                                                SD, segment_id, plane);
    VP8_LAUNCH_CHECK();
}

extern "C" int vp8b200_count_SSIM(void *stream, const uint8_t *f1, const uint8_t *f2, const int32_t *MB_seg,
                                  float *metric, int width, int height, int segment_id, int mb_size) {
    const int M = (width / mb_size) * (height / mb_size);
    if (M <= 0) return 0;
    if (mb_size == 16)
        k_count_SSIM<16><<<(M + 63) / 64, 64, 0, (cudaStream_t)stream>>>(f1, f2, MB_seg, metric, width, M, segment_id);
    else
        k_count_SSIM<8><<<(M + 63) / 64, 64, 0, (cudaStream_t)stream>>>(f1, f2, MB_seg, metric, width, M, segment_id);
    VP8_LAUNCH_CHECK();
}

extern "C" int vp8b200_gather_SSIM(void *stream, const float *m1, const float *m2, const float *m3, float *MB_SSIM,
                                   int mb_count) {
    if (mb_count <= 0) return 0;
    k_gather_SSIM<<<(mb_count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(m1, m2, m3, MB_SSIM, mb_count);
    VP8_LAUNCH_CHECK();
}
