// Loop-filter kernels of the vp8oclenc_b200 engine (sm_100a): prepare_filter_mask and the VP8
// normal loop filter (src/CPU_kernels.cl:782-1075, 1333-1438).
//
// The reference filters a plane with ONE work-item walking the macroblocks in raster order.
// Raster order only constrains MB(r,c) to run after MB(r,c-1) (its left edge) and after
// MB(r-1,c+1) (whose left-edge filter rewrites the bottom-right pixels of MB(r-1,c), which the
// top-edge filter of MB(r,c) reads).  All macroblocks with the same c + 2r are therefore
// independent: the kernel runs the mb_w + 2(mb_h-1) anti-diagonal "stages" as a wavefront,
// 16 lanes per macroblock (one per pixel row, then one per pixel column), all three planes
// concurrently.  The honest bound of this kernel is stages x (L2 round trip + filter latency),
// not HBM bandwidth ("hard parts").
//
// Arithmetic: the reference uses short8 lanes.  Every intermediate stays far inside int16
// (pixels - 128 drift by at most a few dozen between the clamped stores, the largest product
// is 27 * 127), so plain int arithmetic is bit-identical.
#include "common.cuh"

namespace vp8 {

__device__ __forceinline__ int c128(int v) { return min(max(v, -128), 127); }

__device__ __forceinline__ bool lf_mask(int p3, int p2, int p1, int p0, int q0, int q1, int q2, int q3, int e_lim,
                                        int i_lim) {
    return !(abs(p3 - p2) > i_lim || abs(p2 - p1) > i_lim || abs(p1 - p0) > i_lim || abs(q1 - q0) > i_lim ||
             abs(q2 - q1) > i_lim || abs(q3 - q2) > i_lim || (abs(p0 - q0) * 2 + (abs(p1 - q1) >> 1)) > e_lim);
}

// macroblock edge (filter_mb_edge8, src/CPU_kernels.cl:829-883), one lane
__device__ __forceinline__ void filter_mb_edge(int p3, int &p2, int &p1, int &p0, int &q0, int &q1, int &q2, int q3,
                                               int mb_lim, int int_lim, int hev_thr) {
    const bool mask = lf_mask(p3, p2, p1, p0, q0, q1, q2, q3, mb_lim, int_lim);
    const bool hev = abs(p1 - p0) > hev_thr || abs(q1 - q0) > hev_thr;
    int w = c128(c128(p1 - q1) + 3 * (q0 - p0));
    if (!mask) w = 0;
    int a = hev ? w : 0;
    const int b = c128(a + 3) >> 3;
    a = c128(a + 4) >> 3;
    q0 -= a;
    p0 += b;
    if (hev) w = 0;
    a = c128((27 * w + 63) >> 7);
    q0 -= a;
    p0 += a;
    a = c128((18 * w + 63) >> 7);
    q1 -= a;
    p1 += a;
    a = c128((9 * w + 63) >> 7);
    q2 -= a;
    p2 += a;
}

// inner (sub-block) edge (filter_b_edge8, src/CPU_kernels.cl:885-926), one lane
__device__ __forceinline__ void filter_b_edge(int p3, int p2, int &p1, int &p0, int &q0, int &q1, int q2, int q3,
                                              int b_lim, int int_lim, int hev_thr) {
    const bool mask = lf_mask(p3, p2, p1, p0, q0, q1, q2, q3, b_lim, int_lim);
    const bool hev = abs(p1 - p0) > hev_thr || abs(q1 - q0) > hev_thr;
    int a = hev ? c128(p1 - q1) : 0;
    a = c128(a + 3 * (q0 - p0));
    if (!mask) a = 0;
    const int b = c128(a + 3) >> 3;
    a = c128(a + 4) >> 3;
    q0 -= a;
    p0 += b;
    a = (a + 1) >> 1;
    if (hev) a = 0;
    q1 -= a;
    p1 += a;
}

// All edges that cross one line of N pixels (v[4..4+N)) plus the 4 pixels before it (v[0..4)).
// v holds pixel-128 and keeps the UNCLAMPED results: within one macroblock the reference
// hands the unclamped q0..q3 of an edge on as p3..p0 of the next (Q7); memory gets clamped.
template <int N>
__device__ __forceinline__ void filter_line(int (&v)[N + 4], bool mb_edge, bool inner, int mb_lim, int b_lim,
                                            int int_lim, int hev_thr) {
    if (mb_edge) filter_mb_edge(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], mb_lim, int_lim, hev_thr);
    if (inner) {
#pragma unroll
        for (int e = 4; e < N; e += 4)
            filter_b_edge(v[e], v[e + 1], v[e + 2], v[e + 3], v[e + 4], v[e + 5], v[e + 6], v[e + 7], b_lim, int_lim,
                          hev_thr);
    }
}

struct LFPlanes {
    uint8_t *ptr[3];
};

// One CTA per plane, 64 macroblock slots x 16 lanes.
template <int N>
__device__ void lf_plane(uint8_t *__restrict__ frame, const int *__restrict__ seg, const int *__restrict__ mb_mask,
                         const vp8b200_segment_data *__restrict__ SD, int width, int height, int *s_stop) {
    const int mbw = width / N, mbh = height / N, mb_count = mbw * mbh;
    const int tid = threadIdx.x, slot = tid >> 4, lane = tid & 15;

    // "if (SD[i].loop_filter_level == 0) return;" ends the WHOLE plane at the first such
    // macroblock in raster order (Q6): find that index
    if (tid == 0) *s_stop = mb_count;
    __syncthreads();
    int first = mb_count;
    for (int mb = tid; mb < mb_count; mb += blockDim.x)
        if (SD[seg[mb]].loop_filter_level == 0) {
            first = mb;
            break;
        }
    if (first < mb_count) atomicMin(s_stop, first);
    __syncthreads();
    const int stop = *s_stop;

    const int stages = mbw + 2 * (mbh - 1);
    for (int s = 0; s < stages; ++s) {
        for (int r0 = 0; r0 < mbh; r0 += 64) {
            const int r = r0 + slot, c = s - 2 * r;
            const int mb = r * mbw + c;
            const bool active = r < mbh && c >= 0 && c < mbw && mb < stop && lane < N;
            int mb_lim = 0, b_lim = 0, int_lim = 0, hev_thr = 0;
            bool inner = false;
            const int x0 = c * N, y0 = r * N;
            if (active) {
                const vp8b200_segment_data *sd = SD + seg[mb];
                int_lim = (short)sd->interior_limit;
                mb_lim = (short)sd->mbedge_limit;
                b_lim = (short)sd->sub_bedge_limit;
                hev_thr = (short)sd->hev_threshold;
                inner = mb_mask[mb] != 0;
                // pass 1: vertical edges, one lane per pixel row
                uint8_t *row = frame + (size_t)(y0 + lane) * width + x0;
                int v[N + 4];
                uint32_t words[N / 4 + 1];
                words[0] = x0 > 0 ? *reinterpret_cast<const uint32_t *>(row - 4) : 0;
#pragma unroll
                for (int k = 0; k < N / 4; ++k) words[k + 1] = *reinterpret_cast<const uint32_t *>(row + 4 * k);
#pragma unroll
                for (int k = 0; k < N + 4; ++k) v[k] = (int)((words[k >> 2] >> (8 * (k & 3))) & 255) - 128;
                filter_line<N>(v, x0 > 0, inner, mb_lim, b_lim, int_lim, hev_thr);
#pragma unroll
                for (int k = 0; k < N / 4 + 1; ++k)
                    words[k] = (uint32_t)sat8(v[4 * k] + 128) | ((uint32_t)sat8(v[4 * k + 1] + 128) << 8) |
                               ((uint32_t)sat8(v[4 * k + 2] + 128) << 16) | ((uint32_t)sat8(v[4 * k + 3] + 128) << 24);
                if (x0 > 0) *reinterpret_cast<uint32_t *>(row - 4) = words[0];
#pragma unroll
                for (int k = 0; k < N / 4; ++k) *reinterpret_cast<uint32_t *>(row + 4 * k) = words[k + 1];
            }
            __syncwarp();  // both passes of a macroblock live in one warp (16 lanes)
            if (active) {
                // pass 2: horizontal edges, one lane per pixel column
                uint8_t *col = frame + (size_t)y0 * width + x0 + lane;
                int v[N + 4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = y0 > 0 ? (int)col[(ptrdiff_t)(k - 4) * width] - 128 : 0;
#pragma unroll
                for (int k = 0; k < N; ++k) v[k + 4] = (int)col[(size_t)k * width] - 128;
                filter_line<N>(v, y0 > 0, inner, mb_lim, b_lim, int_lim, hev_thr);
                if (y0 > 0) {
#pragma unroll
                    for (int k = 1; k < 4; ++k) col[(ptrdiff_t)(k - 4) * width] = (uint8_t)sat8(v[k] + 128);
                }
#pragma unroll
                for (int k = 0; k < N; ++k) col[(size_t)k * width] = (uint8_t)sat8(v[k + 4] + 128);
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024)
k_loop_filter(LFPlanes planes, int first_plane, const int *__restrict__ seg, const int *__restrict__ mb_mask,
              const vp8b200_segment_data *__restrict__ SD, int luma_width, int luma_height) {
    __shared__ int s_stop;
    const int plane = first_plane + blockIdx.x;
    if (plane == 0)
        lf_plane<16>(planes.ptr[0], seg, mb_mask, SD, luma_width, luma_height, &s_stop);
    else
        lf_plane<8>(planes.ptr[plane], seg, mb_mask, SD, luma_width / 2, luma_height / 2, &s_stop);
}

// one warp per macroblock: sum of |coefficient| over the positions the entropy coder will
// visit, and the inner-edge mask (prepare_filter_mask, src/CPU_kernels.cl:782-827)
__global__ void k_prepare_filter_mask(const int *__restrict__ MB, int *__restrict__ nz, const int *__restrict__ parts,
                                      int *__restrict__ mb_mask, int mb_count) {
    const int mb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (mb >= mb_count) return;
    const int split = parts[mb];
    const int *m = MB + (size_t)mb * 200;
    int sum = 0;
    for (int i = lane; i < 200; i += 32) {
        const int w = __ldg(m + i);
        const int blk = i >> 3, pos = (i & 7) * 2;  // two coefficients per word
        const int lo = abs((int)(short)(w & 0xffff)), hi = abs(w >> 16);
        bool take_lo, take_hi = true;
        if (blk < 16) {
            take_lo = pos != 0 || split != ARE16x16;  // luma DC only counts without a Y2 block
        } else if (blk < 24) {
            take_lo = true;
        } else {
            take_lo = take_hi = (split == ARE16x16);
        }
        sum += (take_lo ? lo : 0) + (take_hi ? hi : 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) {
        nz[mb] = sum;
        mb_mask[mb] = (split != ARE16x16 || sum > 0) ? -1 : 0;
    }
}

}  // namespace vp8

using namespace vp8;

extern "C" int vp8b200_prepare_filter_mask(void *stream, const int16_t *MB, int32_t *nz, const int32_t *parts,
                                           int32_t *mb_mask, int width, int height) {
    const int M = (width / 16) * (height / 16);
    if (M <= 0) return 0;
    k_prepare_filter_mask<<<(M * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const int *)MB, nz, parts, mb_mask, M);
    VP8_LAUNCH_CHECK();
}

extern "C" int vp8b200_loop_filter_frame(void *stream, uint8_t *frame, const int32_t *seg, const int32_t *mb_mask,
                                         const vp8b200_segment_data *SD, int width, int height, int mb_size) {
    if (width < mb_size || height < mb_size) return 0;
    LFPlanes p;
    p.ptr[0] = p.ptr[1] = p.ptr[2] = frame;
    // the chroma path takes the LUMA size and halves it
    if (mb_size == 16)
        k_loop_filter<<<1, 1024, 0, (cudaStream_t)stream>>>(p, 0, seg, mb_mask, SD, width, height);
    else
        k_loop_filter<<<1, 1024, 0, (cudaStream_t)stream>>>(p, 1, seg, mb_mask, SD, width * 2, height * 2);
    VP8_LAUNCH_CHECK();
}

extern "C" int vp8b200_loop_filter_planes(void *stream, uint8_t *y, uint8_t *u, uint8_t *v, const int32_t *seg,
                                          const int32_t *mb_mask, const vp8b200_segment_data *SD, int width,
                                          int height) {
    if (width < 16 || height < 16) return 0;
    LFPlanes p;
    p.ptr[0] = y;
    p.ptr[1] = u;
    p.ptr[2] = v;
    k_loop_filter<<<3, 1024, 0, (cudaStream_t)stream>>>(p, 0, seg, mb_mask, SD, width, height);
    VP8_LAUNCH_CHECK();
}
