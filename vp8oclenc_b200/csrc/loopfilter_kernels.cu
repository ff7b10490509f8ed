// Loop-filter kernels of the vp8oclenc_b200 engine (sm_100a): prepare_filter_mask and the VP8
// normal loop filter (src/CPU_kernels.cl:782-1075, 1333-1438).
//
// The reference filters a plane with ONE work-item walking the macroblocks in raster order.
// Raster order only constrains MB(r,c) to run after MB(r,c-1) (its left edge) and after
// MB(r-1,c+1) (whose left-edge filter rewrites the right-most pixels of MB(r-1,c), which the
// top-edge filter of MB(r,c) reads).  The kernel therefore runs ONE CTA PER MACROBLOCK ROW of
// every plane, all rows of all three planes concurrently, as a wavefront:
//   * the row's whole pixel strip (16 x width luma, 8 x width/2 chroma) is loaded into shared
//     memory once with coalesced 16-byte loads and written back once; the serial walk along the
//     row touches shared memory only;
//   * rows talk through global memory: a row publishes, per macroblock, its final bottom four
//     pixel lines and then a progress counter (release); the row below acquires the counter,
//     reads those four lines (bypassing L1), filters the shared edge and owns the three lines it
//     rewrites from then on;
//   * rows are handed to CTAs through an atomic ticket in launch order, so a waiting row's
//     predecessor is always already running: no deadlock whatever the residency.
// The bound of this kernel is the dependency chain ~ (mb_w + c*mb_h) macroblock steps of eight
// serial edge filters each, not HBM bandwidth ("hard parts").
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

__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) {
    return (uint32_t)sat8(a + 128) | ((uint32_t)sat8(b + 128) << 8) | ((uint32_t)sat8(c + 128) << 16) |
           ((uint32_t)sat8(d + 128) << 24);
}

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// loads that must observe another SM's stores: bypass the (non-coherent) L1
__device__ __forceinline__ uint32_t ld_cg_u8(const uint8_t *p) {
    uint32_t v;
    asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct LFPlanes {
    uint8_t *ptr[3];
};

constexpr int LF_THREADS = 128;

// one macroblock row of one plane; smem = strip of N rows, stride S = width + 4 (word stride
// odd -> the row-per-lane accesses of the horizontal pass are bank-conflict free), then the
// per-macroblock limits
template <int N>
__device__ void lf_row(uint8_t *__restrict__ frame, const int *__restrict__ seg, const int *__restrict__ mb_mask,
                       const vp8b200_segment_data *__restrict__ SD, int width, int height, int r, int stop,
                       int *progress, unsigned char *smem) {
    const int mbw = width / N, mbh = height / N;
    const int S = width + 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncols = max(0, min(mbw, stop - r * mbw));  // macroblocks of this row before the Q6 stop
    if (ncols == 0) return;
    uint8_t *strip = smem;
    int4 *lims = reinterpret_cast<int4 *>(smem + ((N * S + 15) & ~15));  // {mb_lim, b_lim, int_lim, hev_thr | inner<<16}
    const int y0 = r * N;

    // ---- stage the strip and the per-macroblock limits ----
    // 8-byte chunks: plane widths are multiples of 8 (chroma of a 16-aligned luma), rows 8-byte aligned
    const int chunks = width / 8;
    for (int i = tid; i < N * chunks; i += LF_THREADS) {
        const int row = i / chunks, xc = i % chunks;
        const uint2 v = *reinterpret_cast<const uint2 *>(frame + (size_t)(y0 + row) * width + 8 * xc);
        uint32_t *d = reinterpret_cast<uint32_t *>(strip + row * S + 4 + 8 * xc);
        d[0] = v.x; d[1] = v.y;
    }
    for (int c = tid; c < ncols; c += LF_THREADS) {
        const int mb = r * mbw + c;
        const vp8b200_segment_data *sd = SD + seg[mb];
        lims[c] = make_int4((short)sd->mbedge_limit, (short)sd->sub_bedge_limit, (short)sd->interior_limit,
                            ((short)sd->hev_threshold & 0xffff) | (mb_mask[mb] != 0 ? 0x10000 : 0));
    }
    __syncthreads();

    // ---- the serial walk, warp-specialised so that warp 0 never waits for global memory ----
    //   warp 0  filters: pass 1 (vertical edges) + pass 2 (horizontal edges) per macroblock, shared memory only
    //   warp 1  prefetches the four pixel lines above each macroblock once the row above has finalised them
    //   warp 2  publishes the bottom four lines of every finalised macroblock and the progress counter
    // hand-offs between the three are counters in shared memory.
    constexpr int TOPQ = 8;  // depth of the top-line ring
    uint32_t *topq = reinterpret_cast<uint32_t *>(lims + mbw);             // [TOPQ][4 lines][N/4 words]
    volatile int *flags = reinterpret_cast<volatile int *>(topq + TOPQ * N);  // h_done, v_done, top_ready
    if (tid < 3) flags[tid] = 0;
    __syncthreads();
    const int *above = progress - 1;  // progress counter of row r-1 (only read when r > 0)

    if (warp == 0) {
        for (int c = 0; c < ncols; ++c) {
            const int4 lm = lims[c];
            const int mb_lim = lm.x, b_lim = lm.y, int_lim = lm.z, hev_thr = lm.w & 0xffff;
            const bool inner = (lm.w >> 16) != 0;
            const int x0 = c * N;
            // pass 1: vertical edges, one lane per pixel row
            if (lane < N) {
                uint32_t *row = reinterpret_cast<uint32_t *>(strip + lane * S + 4 + x0);
                uint32_t words[N / 4 + 1];
                words[0] = c > 0 ? row[-1] : 0;
#pragma unroll
                for (int k = 0; k < N / 4; ++k) words[k + 1] = row[k];
                int v[N + 4];
#pragma unroll
                for (int k = 0; k < N + 4; ++k) v[k] = (int)((words[k >> 2] >> (8 * (k & 3))) & 255) - 128;
                filter_line<N>(v, c > 0, inner, mb_lim, b_lim, int_lim, hev_thr);
                if (c > 0) row[-1] = pack4(v[0], v[1], v[2], v[3]);
#pragma unroll
                for (int k = 0; k < N / 4; ++k) row[k] = pack4(v[4 * k + 4], v[4 * k + 5], v[4 * k + 6], v[4 * k + 7]);
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                flags[0] = c + 1;  // h_done: macroblock c-1 is final now
            }
            // pass 2: horizontal edges, one lane per pixel column
            if (r > 0) {
                while (flags[2] < c + 1) {}  // top lines of macroblock c are in the ring
                __threadfence_block();
            }
            if (lane < N) {
                int v[N + 4];
                if (r > 0) {
                    const uint8_t *tq = reinterpret_cast<const uint8_t *>(topq + (c % TOPQ) * N) + lane;
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = (int)tq[k * N] - 128;
                } else {
                    v[0] = v[1] = v[2] = v[3] = 0;
                }
                uint8_t *scol = strip + 4 + x0 + lane;
#pragma unroll
                for (int k = 0; k < N; ++k) v[k + 4] = (int)scol[k * S] - 128;
                filter_line<N>(v, r > 0, inner, mb_lim, b_lim, int_lim, hev_thr);
                if (r > 0) {  // three lines of the row above now belong to this row: straight to global memory
                    uint8_t *gcol = frame + (size_t)y0 * width + x0 + lane;
#pragma unroll
                    for (int k = 1; k < 4; ++k) gcol[(ptrdiff_t)(k - 4) * width] = (uint8_t)sat8(v[k] + 128);
                }
#pragma unroll
                for (int k = 0; k < N; ++k) scol[k * S] = (uint8_t)sat8(v[k + 4] + 128);
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                flags[1] = c + 1;  // v_done
            }
        }
    } else if (warp == 1) {
        if (r > 0) {
            for (int c = 0; c < ncols; ++c) {
                while (c - flags[1] >= TOPQ) __nanosleep(40);  // ring slot free again
                if (lane == 0)
                    while (ld_acquire(above) < c + 1) __nanosleep(20);  // MB(r-1,c) final <=> MB(r-1,c+1) edge-filtered
                __syncwarp();
                if (lane < N) {  // 4 lines x N/4 words
                    const int line = lane / (N / 4), wd = lane % (N / 4);
                    uint32_t w;
                    const uint8_t *g = frame + (size_t)(y0 - 4 + line) * width + c * N + 4 * wd;
                    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(w) : "l"(g) : "memory");
                    topq[(c % TOPQ) * N + line * (N / 4) + wd] = w;
                }
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    flags[2] = c + 1;  // top_ready
                }
            }
        }
    } else if (warp == 2) {
        for (int c = 0; c < ncols; ++c) {
            // macroblock c is final once pass 1 of macroblock c+1 ran, the last one after its own pass 2
            if (c + 1 < ncols) {
                while (flags[0] < c + 2) __nanosleep(40);
            } else {
                while (flags[1] < ncols) __nanosleep(40);
            }
            __threadfence_block();
            if (lane < N) {
                const int line = N - 4 + lane / (N / 4), wd = lane % (N / 4);
                const uint32_t w = *reinterpret_cast<const uint32_t *>(strip + line * S + 4 + c * N + 4 * wd);
                *reinterpret_cast<uint32_t *>(frame + (size_t)(y0 + line) * width + c * N + 4 * wd) = w;
            }
            __syncwarp();
            if (lane == 0) st_release(progress, c + 1 < ncols ? c + 1 : mbw + 1);
        }
    }
    __syncthreads();
    // ---- write back lines 0..N-5 of the filtered range (the last four lines are already out,
    // and the row below may have rewritten three of them since) ----
    const int out_chunks = ncols * N / 8;
    for (int i = tid; i < (N - 4) * out_chunks; i += LF_THREADS) {
        const int row = i / out_chunks, xc = i % out_chunks;
        const uint32_t *s = reinterpret_cast<const uint32_t *>(strip + row * S + 4 + 8 * xc);
        *reinterpret_cast<uint2 *>(frame + (size_t)(y0 + row) * width + 8 * xc) = make_uint2(s[0], s[1]);
    }
    (void)mbh;
}

// ctrl[0] = ticket counter, ctrl[1 + plane*max_rows + row] = progress of that row
__global__ void __launch_bounds__(LF_THREADS)
k_loop_filter(LFPlanes planes, int first_plane, int num_planes, const int *__restrict__ seg,
              const int *__restrict__ mb_mask, const vp8b200_segment_data *__restrict__ SD, int luma_width,
              int luma_height, int *ctrl, int max_rows) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_ticket, s_stop;
    const int mbw = luma_width / 16, mbh = luma_height / 16, mb_count = mbw * mbh;
    if (threadIdx.x == 0) {
        s_ticket = atomicAdd(&ctrl[0], 1);
        s_stop = mb_count;
    }
    __syncthreads();
    // "if (SD[i].loop_filter_level == 0) return;" ends the WHOLE plane at the first such macroblock
    // in raster order (Q6): find that index
    unsigned zero_mask = 0;  // segments whose filter level is 0
#pragma unroll
    for (int sgm = 0; sgm < 4; ++sgm) zero_mask |= (SD[sgm].loop_filter_level == 0) << sgm;
    if (zero_mask) {  // rare: scan the segment map for the first such macroblock
        int first = mb_count;
        for (int mb = threadIdx.x; mb < mb_count && first == mb_count; mb += LF_THREADS)
            if ((zero_mask >> seg[mb]) & 1) first = mb;
        if (first < mb_count) atomicMin(&s_stop, first);
    }
    __syncthreads();
    const int t = s_ticket;
    const int plane = first_plane + t % num_planes, r = t / num_planes;
    if (r >= mbh) return;
    int *progress = ctrl + 1 + (plane - first_plane) * max_rows + r;
    if (plane == 0)
        lf_row<16>(planes.ptr[0], seg, mb_mask, SD, luma_width, luma_height, r, s_stop, progress, smem);
    else
        lf_row<8>(planes.ptr[plane], seg, mb_mask, SD, luma_width / 2, luma_height / 2, r, s_stop, progress, smem);
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

// ticket + progress counters, one allocation per process (launches on different streams must
// not overlap; the engine and the shim use a single stream)
static int *g_lf_ctrl = nullptr;
static const int LF_MAX_ROWS = 2048;

static int launch_loop_filter(void *stream, LFPlanes p, int first_plane, int num_planes, const int32_t *seg,
                              const int32_t *mb_mask, const vp8b200_segment_data *SD, int luma_w, int luma_h) {
    cudaStream_t st = (cudaStream_t)stream;
    const int mbh = luma_h / 16;
    if (mbh > LF_MAX_ROWS) return -(int)cudaErrorInvalidValue;
    if (!g_lf_ctrl && cudaMalloc((void **)&g_lf_ctrl, sizeof(int) * (1 + 3 * LF_MAX_ROWS)) != cudaSuccess)
        return -(int)cudaErrorMemoryAllocation;
    cudaMemsetAsync(g_lf_ctrl, 0, sizeof(int) * (1 + (size_t)num_planes * LF_MAX_ROWS), st);
    const int strip_w = first_plane == 0 ? luma_w : luma_w / 2;
    const int strip_n = first_plane == 0 ? 16 : 8;
    const size_t smem = (((size_t)strip_n * (strip_w + 4) + 15) & ~(size_t)15) + (size_t)(luma_w / 16) * sizeof(int4) +
                        8 * 16 * sizeof(uint32_t) + 16;  // strip + limits + top-line ring + flags
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(k_loop_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return -(int)cudaGetLastError();
        configured = smem;
    }
    k_loop_filter<<<mbh * num_planes, LF_THREADS, smem, st>>>(p, first_plane, num_planes, seg, mb_mask, SD, luma_w, luma_h,
                                                             g_lf_ctrl, LF_MAX_ROWS);
    VP8_LAUNCH_CHECK();
}

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
    if (mb_size == 16) return launch_loop_filter(stream, p, 0, 1, seg, mb_mask, SD, width, height);
    return launch_loop_filter(stream, p, 1, 1, seg, mb_mask, SD, width * 2, height * 2);
}

extern "C" int vp8b200_loop_filter_planes(void *stream, uint8_t *y, uint8_t *u, uint8_t *v, const int32_t *seg,
                                          const int32_t *mb_mask, const vp8b200_segment_data *SD, int width,
                                          int height) {
    if (width < 16 || height < 16) return 0;
    LFPlanes p;
    p.ptr[0] = y;
    p.ptr[1] = u;
    p.ptr[2] = v;
    return launch_loop_filter(stream, p, 0, 3, seg, mb_mask, SD, width, height);
}
