// Motion-estimation kernels of the vp8oclenc_b200 engine (sm_100a):
//   reset_vectors, downsample_x2, luma_search_1step, luma_search_2step, select_reference,
//   pack_8x8_into_16x16.
//
// These are integer-ALU-bound kernels (the block metric is a 4x4 transform
// cost, not a SAD, so packed-byte SAD intrinsics cannot be used).  Design:
//   * the search window of every 8x8 block is staged ONCE in shared memory and shared by all
//     candidates (the reference re-reads global memory / the image per candidate);
//   * the work unit is (block, candidate, 4x4 sub-block): 100 units per block for the
//     full-pel levels, 104 for the quarter-pel level, so warps are (almost) fully populated
//     instead of 25/32 lanes;
//   * the four sub-block costs of a candidate are combined with two warp shuffles and the
//     winner is found with one shared-memory atomicMin on the key (cost << 8 | scan index),
//     which reproduces the reference's "first candidate in scan order wins ties";
//   * in the quarter-pel kernel the horizontal six-tap pass is computed once per x-phase
//     (5 variants) and shared by the 5 y-phases, instead of once per candidate.
#include "common.cuh"

namespace vp8 {

// ------------------------------------------------------------------------------------------
__global__ void k_reset_vectors(int *n0, int *n1, int *n2, int *n3, int *n4, int *n5, int *m0, int *m1, int *m2, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    n0[i] = 0; n1[i] = 0; n2[i] = 0; n3[i] = 0; n4[i] = 0; n5[i] = 0;
    m0[i] = 0x7fffffff; m1[i] = 0x7fffffff; m2[i] = 0x7fffffff;
}

// one thread per output pixel; (a+b+c+d+2)/4, src/GPU_kernels.cl:429-451
__global__ void k_downsample_x2(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int w, int h) {
    const int ow = w >> 1, oh = h >> 1;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= ow || y >= oh) return;
    const uint8_t *p = src + (size_t)(2 * y) * w + 2 * x;
    // w is even, so both 2-byte loads are aligned
    const unsigned a = *reinterpret_cast<const unsigned short *>(p);
    const unsigned b = *reinterpret_cast<const unsigned short *>(p + w);
    dst[(size_t)y * ow + x] = (uint8_t)(((a & 255) + (a >> 8) + (b & 255) + (b >> 8) + 2) >> 2);
}

// ------------------------------------------------------------------------------------------
// luma_search_1step (src/GPU_kernels.cl:459-560): +-2 full search around the parent vector.
constexpr int S1_BLOCKS = 8;             // 8x8 blocks per CTA
constexpr int S1_THREADS = S1_BLOCKS * 100;  // (25 candidates x 4 sub-blocks) per block = 25 full warps

// up to three references (LAST / GOLDEN / ALTREF) searched by one launch: blockIdx.y picks the set
struct Search1Refs {
    const uint8_t *prev[3];
    const short2 *src_net[3];
    short2 *dst_net[3];
};
struct Search2Refs {
    const uint8_t *ref[3];
    const short2 *net[3];
    short2 *ref_net[3];
    int *ref_Bdiff[3];
};

__global__ void __launch_bounds__(S1_THREADS)
k_luma_search_1step(const uint8_t *__restrict__ cur, Search1Refs refs, int net_width, int width, int height, int rate,
                    int nblocks) {
    const uint8_t *__restrict__ prev = refs.prev[blockIdx.y];
    const short2 *__restrict__ src_net = refs.src_net[blockIdx.y];
    short2 *__restrict__ dst_net = refs.dst_net[blockIdx.y];
    __shared__ uint8_t s_win[S1_BLOCKS][12][12];   // prev pixels [c+v0-2, c+v0+10) in both axes
    __shared__ uint8_t s_cur[S1_BLOCKS][8][8];
    __shared__ int s_cx[S1_BLOCKS], s_cy[S1_BLOCKS], s_vx[S1_BLOCKS], s_vy[S1_BLOCKS];
    __shared__ unsigned s_key[S1_BLOCKS];

    const int tid = threadIdx.x;
    const int bw = width >> 3;  // == cut_width/8
    if (tid < S1_BLOCKS) {
        const int n = blockIdx.x * S1_BLOCKS + tid;
        int cx = 0, cy = 0, vx = 0, vy = 0;
        if (n < nblocks) {
            cx = (n % bw) * 8;
            cy = (n / bw) * 8;
            // parent = the block that covers c/2 one level up; nets keep the full-resolution
            // stride at every level (Q4).  Stored vectors are in full-pel units x rate.
            const short2 pv = src_net[((cy / 2) / 8) * net_width + (cx / 2) / 8];
            vx = (short)((int)pv.x / (int)(short)rate);
            vy = (short)((int)pv.y / (int)(short)rate);
            if (rate > 8) vx = vy = 0;
        }
        s_cx[tid] = cx; s_cy[tid] = cy; s_vx[tid] = vx; s_vy[tid] = vy;
        s_key[tid] = 0xffffffffu;
    }
    __syncthreads();

    // stage windows and current blocks.  Pixels outside the plane are only ever used by
    // candidates that are not fully inside it, and those can never win (Q2), so the
    // coordinates are simply clamped to keep the loads in bounds.
    for (int i = tid; i < S1_BLOCKS * 144; i += S1_THREADS) {
        const int b = i / 144, r = (i % 144) / 12, c = i % 12;
        const int x = clampi(s_cx[b] + s_vx[b] - 2 + c, 0, width - 1);
        const int y = clampi(s_cy[b] + s_vy[b] - 2 + r, 0, height - 1);
        s_win[b][r][c] = __ldg(prev + (size_t)y * width + x);
    }
    for (int i = tid; i < S1_BLOCKS * 64; i += S1_THREADS) {
        const int b = i >> 6, r = (i >> 3) & 7, c = i & 7;
        const int n = blockIdx.x * S1_BLOCKS + b;
        s_cur[b][r][c] = (n < nblocks) ? __ldg(cur + (size_t)(s_cy[b] + r) * width + s_cx[b] + c) : 0;
    }
    __syncthreads();

    {
        const int b = tid / 100, u = tid % 100, k = u >> 2, j = u & 3;
        const int dx = k % 5, dy = k / 5;               // window offset of the candidate
        const int sx = (j >> 1) * 4, sy = (j & 1) * 4;  // sub-block order (0,0),(0,4),(4,0),(4,4) as (x,y)
        int r[16];
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x)
                r[4 * y + x] = (int)s_cur[b][sy + y][sx + x] - (int)s_win[b][dy + sy + y][dx + sx + x];
        int cost = weight4x4(r);
        cost += __shfl_xor_sync(0xffffffffu, cost, 1);
        cost += __shfl_xor_sync(0xffffffffu, cost, 2);
        if (j == 0) {
            const int cx = s_cx[b], cy = s_cy[b], vx = s_vx[b], vy = s_vy[b];
            const int px = (short)(cx + vx + dx - 2), py = (short)(cy + vy + dy - 2);
            const bool inside = px >= 0 && px <= width - 8 && py >= 0 && py <= height - 8;
            // the reference accumulates in an unsigned short (Q2) and adds the neighbour-coherence
            // term only at rates 2 and 1 (Q3): (| |p.x-c.x| - v0.x | + | |p.y-c.y| - v0.y |) * 32
            unsigned diff = (unsigned)cost & 0xffffu;
            if (rate < 4) diff = (diff + (unsigned)((abs(abs(px - cx) - vx) + abs(abs(py - cy) - vy)) * 32)) & 0xffffu;
            if (inside && diff < 0x7fffu && blockIdx.x * S1_BLOCKS + b < nblocks)
                atomicMin(&s_key[b], (diff << 8) | (unsigned)k);
        }
    }
    __syncthreads();

    if (tid < S1_BLOCKS) {
        const int n = blockIdx.x * S1_BLOCKS + tid;
        if (n < nblocks) {
            const unsigned key = s_key[tid];
            const int cx = s_cx[tid], cy = s_cy[tid];
            int bx, by;  // "vector" of the reference: a position once a candidate has won, else v0 itself
            if (key == 0xffffffffu) {
                bx = s_vx[tid];
                by = s_vy[tid];
            } else {
                const int k = key & 255;
                bx = (short)(cx + s_vx[tid] + (k % 5) - 2);
                by = (short)(cy + s_vy[tid] + (k / 5) - 2);
            }
            short2 out;
            out.x = (short)((int)(short)(bx - cx) * (int)(short)rate);
            out.y = (short)((int)(short)(by - cy) * (int)(short)rate);
            dst_net[(cy / 8) * net_width + cx / 8] = out;
        }
    }
}

// ------------------------------------------------------------------------------------------
// luma_search_2step (src/GPU_kernels.cl:776-1203): 25 quarter-pel candidates around 4*v plus
// the zero vector, six-tap interpolation of the "search flavour" (every horizontally filtered
// line is saturated before the vertical pass, Q5).  One CTA per 8x8 block.
constexpr int S2_THREADS = 128;

// x/y phase variants of the five offsets -2..+2 quarter pels around a full-pel position:
// integer origin offset (-1,-1,0,0,0) and eighth-pel filter phase (4,6,0,2,4)
__device__ __forceinline__ int s2_origin(int i) { return i < 2 ? -1 : 0; }
__device__ __forceinline__ int s2_phase(int i) { return (0x46024 >> (4 * (4 - i))) & 15; }

__global__ void __launch_bounds__(S2_THREADS)
k_luma_search_2step(const uint8_t *__restrict__ cur, Search2Refs refs, int width, int height) {
    const uint8_t *__restrict__ ref = refs.ref[blockIdx.y];
    const short2 *__restrict__ net = refs.net[blockIdx.y];
    short2 *__restrict__ ref_net = refs.ref_net[blockIdx.y];
    int *__restrict__ ref_Bdiff = refs.ref_Bdiff[blockIdx.y];
    __shared__ uint8_t s_win[14][16];      // ref pixels rows/cols [base-3, base+11), clamp-to-edge
    __shared__ uint32_t s_h[5][14][2];     // horizontally filtered + saturated lines, 8 px per row
    __shared__ uint32_t s_cur[8][2];
    __shared__ uint32_t s_zero[8][2];      // co-located reference block (candidate #25)
    __shared__ unsigned s_key;

    const int tid = threadIdx.x;
    const int n = blockIdx.x;
    const int bw = width >> 3;
    const int bx = (n % bw) * 8, by = (n / bw) * 8;
    const short2 v = net[n];
    const int v0x = (short)(v.x * 4), v0y = (short)(v.y * 4);  // short lanes in the reference
    const int basex = bx + (v0x >> 2), basey = by + (v0y >> 2);  // v0 is a multiple of 4

    if (tid == 0) s_key = 0xffffffffu;
    for (int i = tid; i < 14 * 14; i += S2_THREADS) {
        const int r = i / 14, c = i % 14;
        const int x = clampi(basex - 3 + c, 0, width - 1), y = clampi(basey - 3 + r, 0, height - 1);
        s_win[r][c] = __ldg(ref + (size_t)y * width + x);
    }
    if (tid < 16) {
        const int r = tid >> 1, h = tid & 1;
        s_cur[r][h] = __ldg(reinterpret_cast<const uint32_t *>(cur + (size_t)(by + r) * width + bx) + h);
        s_zero[r][h] = __ldg(reinterpret_cast<const uint32_t *>(ref + (size_t)(by + r) * width + bx) + h);
    }
    __syncthreads();

    // horizontal pass: 5 x-variants x 14 rows x 8 columns, four columns (one word) per item
    for (int i = tid; i < 5 * 14 * 2; i += S2_THREADS) {
        const int var = i / 28, r = (i % 28) >> 1, h = i & 1;
        const int ph = s2_phase(var), first = s2_origin(var) + 1 + 4 * h;  // window column of tap 0 of output 0
        uint32_t packed = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int s = 64;
#pragma unroll
            for (int t = 0; t < 6; ++t) s += (int)c_sixtap[ph][t] * (int)s_win[r][first + c + t];
            // (s/128 then saturate) == saturate(s>>7): the two only differ for -128<s<0, both give 0
            packed |= (uint32_t)sat8(s >> 7) << (8 * c);
        }
        s_h[var][r][h] = packed;
    }
    __syncthreads();

    {
        // 104 work units; the 24 spare lanes of the last warp run the zero-vector path with
        // valid=false so that the full-mask shuffles below are executed by whole warps
        const int k = min(tid >> 2, 25), j = tid & 3;
        const int sx = j >> 1, sy = (j & 1) * 4;  // sub-block (x word, y row offset); order as dx4/dy4
        int r[16];
        bool valid = tid < 104;
        int penalty = 0;
        if (k < 25) {
            const int xi = k % 5, yi = k / 5;
            const int qx = (short)(bx * 4 + v0x + xi - 2), qy = (short)(by * 4 + v0y + yi - 2);
            valid = qx >= 0 && qx <= width * 4 - 32 && qy >= 0 && qy <= height * 4 - 32;
            penalty = (abs(xi - 2) + abs(yi - 2)) * 32;
            const int ph = s2_phase(yi), row0 = s2_origin(yi) + 1 + sy;  // s_h row of tap 0 of output row 0
            int col[9][4];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const uint32_t w = s_h[xi][row0 + t][sx];
                col[t][0] = w & 255; col[t][1] = (w >> 8) & 255; col[t][2] = (w >> 16) & 255; col[t][3] = w >> 24;
            }
            const int f0 = c_sixtap[ph][0], f1 = c_sixtap[ph][1], f2 = c_sixtap[ph][2], f3 = c_sixtap[ph][3],
                      f4 = c_sixtap[ph][4], f5 = c_sixtap[ph][5];
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const uint32_t cw = s_cur[sy + y][sx];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int s = 64 + f0 * col[y][x] + f1 * col[y + 1][x] + f2 * col[y + 2][x] + f3 * col[y + 3][x] +
                                  f4 * col[y + 4][x] + f5 * col[y + 5][x];
                    r[4 * y + x] = (int)((cw >> (8 * x)) & 255) - sat8(s >> 7);
                }
            }
        } else {  // candidate #25: the zero vector, full-pel, no penalty
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const uint32_t cw = s_cur[sy + y][sx], zw = s_zero[sy + y][sx];
#pragma unroll
                for (int x = 0; x < 4; ++x) r[4 * y + x] = (int)((cw >> (8 * x)) & 255) - (int)((zw >> (8 * x)) & 255);
            }
        }
        int cost = weight4x4(r);
        cost += __shfl_xor_sync(0xffffffffu, cost, 1);
        cost += __shfl_xor_sync(0xffffffffu, cost, 2);
        cost += penalty;
        if (j == 0 && valid && cost < 0x7fff) atomicMin(&s_key, ((unsigned)cost << 5) | (unsigned)k);
    }
    __syncthreads();

    if (tid == 0) {
        const unsigned key = s_key;
        int qx = (short)(width * 4 - 32), qy = (short)(height * 4 - 32), best = 0x7fff;
        if (key != 0xffffffffu) {
            const int k = key & 31;
            best = (int)(key >> 5);
            qx = bx * 4;
            qy = by * 4;
            if (k < 25) {
                qx = (short)(qx + v0x + (k % 5) - 2);
                qy = (short)(qy + v0y + (k / 5) - 2);
            }
        }
        const int mvx = (short)(qx - bx * 4), mvy = (short)(qy - by * 4);
        // the coherence term is taken back out of the stored metric unless the winner is the zero MV
        if (mvx != 0 || mvy != 0) best -= (abs(mvx - v0x) + abs(mvy - v0y)) * 32;
        ref_net[n] = make_short2((short)mvx, (short)mvy);
        ref_Bdiff[n] = best;
    }
}

// ------------------------------------------------------------------------------------------
__global__ void k_select_reference(const short2 *__restrict__ last_net, const short2 *__restrict__ gold_net,
                                   const short2 *__restrict__ alt_net, const int *__restrict__ last_d,
                                   const int *__restrict__ gold_d, const int *__restrict__ alt_d,
                                   int *__restrict__ MB_ref, short2 *__restrict__ MB_vec, int mb_width, int mb_count,
                                   int use_golden, int use_altref) {
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if (mb >= mb_count) return;
    const int bw = mb_width * 2;
    const int b = ((mb / mb_width) * 2) * bw + (mb % mb_width) * 2;
    int d1 = last_d[b] + last_d[b + 1] + last_d[b + bw] + last_d[b + bw + 1];
    int d2 = 0x7fffffff;  // an unused reference counts as the constant, not as a sum (A6)
    if (use_altref == 1) d2 = alt_d[b] + alt_d[b + 1] + alt_d[b + bw] + alt_d[b + bw + 1];
    int ref = (d1 <= d2) ? LAST : ALTREF;
    d1 = (d1 <= d2) ? d1 : d2;
    d2 = 0x7fffffff;
    if (use_golden == 1) d2 = gold_d[b] + gold_d[b + 1] + gold_d[b + bw] + gold_d[b + bw + 1];
    ref = (d1 <= d2) ? ref : GOLDEN;
    const short2 *net = ref == LAST ? last_net : (ref == GOLDEN ? gold_net : alt_net);
    MB_ref[mb] = ref;
    MB_vec[4 * mb + 0] = net[b];
    MB_vec[4 * mb + 1] = net[b + 1];
    MB_vec[4 * mb + 2] = net[b + bw];
    MB_vec[4 * mb + 3] = net[b + bw + 1];
}

__global__ void k_pack_8x8_into_16x16(const int *__restrict__ MB_vec, int *__restrict__ MB_parts,
                                      float *__restrict__ MB_SSIM, int mb_count) {
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if (mb >= mb_count) return;
    const int4 v = *reinterpret_cast<const int4 *>(MB_vec + 4 * mb);  // 4 x short2
    MB_SSIM[mb] = -2.0f;
    MB_parts[mb] = (v.x == v.y && v.x == v.z && v.x == v.w) ? ARE16x16 : ARE8x8;
}

}  // namespace vp8

// ------------------------------------------------------------------------------------------
using namespace vp8;

extern "C" int vp8b200_reset_vectors(void *stream, int16_t *l1, int16_t *l2, int16_t *g1, int16_t *g2, int16_t *a1,
                                     int16_t *a2, int32_t *lb, int32_t *gb, int32_t *ab, int n) {
    if (n <= 0) return 0;
    k_reset_vectors<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((int *)l1, (int *)l2, (int *)g1, (int *)g2,
                                                                      (int *)a1, (int *)a2, lb, gb, ab, n);
    VP8_LAUNCH_CHECK();
}

extern "C" int vp8b200_downsample_x2(void *stream, const uint8_t *src, uint8_t *dst, int w, int h) {
    if (w < 2 || h < 2) return 0;
    dim3 block(32, 8), grid((w / 2 + 31) / 32, (h / 2 + 7) / 8);
    k_downsample_x2<<<grid, block, 0, (cudaStream_t)stream>>>(src, dst, w, h);
    VP8_LAUNCH_CHECK();
}

extern "C" int vp8b200_luma_search_1step(void *stream, const uint8_t *cur, const uint8_t *prev, const int16_t *src_net,
                                         int16_t *dst_net, int net_width, int width, int height, int rate) {
    return vp8b200_luma_search_1step_multi(stream, cur, 1, &prev, &src_net, &dst_net, net_width, width, height, rate);
}

extern "C" int vp8b200_luma_search_1step_multi(void *stream, const uint8_t *cur, int nrefs, const uint8_t *const *prev,
                                               const int16_t *const *src_net, int16_t *const *dst_net, int net_width,
                                               int width, int height, int rate) {
    const int nblocks = (width / 8) * (height / 8);
    if (nblocks <= 0 || nrefs <= 0) return 0;
    if (nrefs > 3) return -(int)cudaErrorInvalidValue;
    Search1Refs r = {};
    for (int i = 0; i < nrefs; ++i) {
        r.prev[i] = prev[i];
        r.src_net[i] = (const short2 *)src_net[i];
        r.dst_net[i] = (short2 *)dst_net[i];
    }
    dim3 grid((nblocks + S1_BLOCKS - 1) / S1_BLOCKS, nrefs);
    k_luma_search_1step<<<grid, S1_THREADS, 0, (cudaStream_t)stream>>>(cur, r, net_width, width, height, rate, nblocks);
    VP8_LAUNCH_CHECK();
}

extern "C" int vp8b200_luma_search_2step(void *stream, const uint8_t *cur, const uint8_t *ref, const int16_t *net,
                                         int16_t *ref_net, int32_t *ref_Bdiff, int width, int height) {
    return vp8b200_luma_search_2step_multi(stream, cur, 1, &ref, &net, &ref_net, &ref_Bdiff, width, height);
}

extern "C" int vp8b200_luma_search_2step_multi(void *stream, const uint8_t *cur, int nrefs, const uint8_t *const *ref,
                                               const int16_t *const *net, int16_t *const *ref_net,
                                               int32_t *const *ref_Bdiff, int width, int height) {
    const int nblocks = width * height / 64;
    if (nblocks <= 0 || nrefs <= 0) return 0;
    if (nrefs > 3) return -(int)cudaErrorInvalidValue;
    Search2Refs r = {};
    for (int i = 0; i < nrefs; ++i) {
        r.ref[i] = ref[i];
        r.net[i] = (const short2 *)net[i];
        r.ref_net[i] = (short2 *)ref_net[i];
        r.ref_Bdiff[i] = ref_Bdiff[i];
    }
    dim3 grid(nblocks, nrefs);
    k_luma_search_2step<<<grid, S2_THREADS, 0, (cudaStream_t)stream>>>(cur, r, width, height);
    VP8_LAUNCH_CHECK();
}

extern "C" int vp8b200_select_reference(void *stream, const int16_t *ln, const int16_t *gn, const int16_t *an,
                                        const int32_t *ld, const int32_t *gd, const int32_t *ad, int32_t *MB_ref,
                                        int16_t *MB_vec, int width, int height, int use_golden, int use_altref) {
    const int mbw = width / 16, M = mbw * (height / 16);
    if (M <= 0) return 0;
    k_select_reference<<<(M + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        (const short2 *)ln, (const short2 *)gn, (const short2 *)an, ld, gd, ad, MB_ref, (short2 *)MB_vec, mbw, M,
        use_golden, use_altref);
    VP8_LAUNCH_CHECK();
}

extern "C" int vp8b200_pack_8x8_into_16x16(void *stream, const int16_t *MB_vec, int32_t *MB_parts, float *MB_SSIM,
                                           int mb_count) {
    if (mb_count <= 0) return 0;
    k_pack_8x8_into_16x16<<<(mb_count + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const int *)MB_vec, MB_parts,
                                                                                   MB_SSIM, mb_count);
    VP8_LAUNCH_CHECK();
}
