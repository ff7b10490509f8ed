// Motion-estimation kernels of the vp8oclenc_b200 engine (sm_100a):
//   reset_vectors, downsample_x2, luma_search_1step, luma_search_2step, select_reference,
//   pack_8x8_into_16x16.
//
// These are integer-ALU-bound kernels (the block metric is a 4x4 transform
// cost, not a SAD, so packed-byte SAD intrinsics cannot be used).  Design:
//   * the search window of every 8x8 block is staged ONCE in shared memory and shared by all
//     candidates (the reference re-reads global memory / the image per candidate);
//   * a thread owns a 4x4 sub-block and a line of five candidates (20 threads per block plus, in the
//     quarter-pel kernel, 4 for the zero vector), so warps are fully populated instead of 25/32
//     lanes and everything the five candidates share is loaded and unpacked once;
//   * the four sub-block costs of a candidate are combined with two warp shuffles; a thread keeps
//     the minimum of its candidates' keys (cost << 8 | scan index) and the block winner is one
//     shared-memory atomicMin per thread, which reproduces "first candidate in scan order wins";
//   * in the quarter-pel kernel the horizontal six-tap pass is computed once per x-phase
//     (5 variants) and shared by the 5 y-phases, instead of once per candidate; a thread owns one
//     x-phase of one 4x4 sub-block and walks its five y-phases with the lines held in registers.
#include <cuda.h>

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
//
// S1_BLOCKS 8x8 blocks per CTA, 20 threads per block: a thread owns one row of candidates (dy) of
// one 4x4 sub-block (j) and walks the five dx.  The eight window bytes per line that those five
// candidates cover are two aligned words held in registers.  No residual is ever formed: the first
// pass of the cost transform is linear before its shift, so it is computed per COLUMN of the current
// sub-block (4 columns, once) and of the window (8 columns, each once) and a candidate only subtracts
// the two (col_features / weight4x4_rows in common.cuh); candidate dx uses window columns dx..dx+3.
#ifndef VP8_S1_BLOCKS
#define VP8_S1_BLOCKS 8
#endif
#ifndef VP8_S1_MINCTAS
#define VP8_S1_MINCTAS 7
#endif
constexpr int S1_BLOCKS = VP8_S1_BLOCKS;
static_assert(S1_BLOCKS % 8 == 0, "threads must fill whole warps (full-mask shuffles)");
constexpr int S1_THREADS = S1_BLOCKS * 20;

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

__global__ void __launch_bounds__(S1_THREADS, VP8_S1_MINCTAS)
k_luma_search_1step(const uint8_t *__restrict__ cur, Search1Refs refs, int net_width, int width, int height, int rate,
                    int nblocks) {
    // (selects, not refs.x[blockIdx.y]: indexing a kernel parameter dynamically copies it to local memory)
    const int ri = blockIdx.y;
    const uint8_t *__restrict__ prev = ri == 0 ? refs.prev[0] : ri == 1 ? refs.prev[1] : refs.prev[2];
    const short2 *__restrict__ src_net = ri == 0 ? refs.src_net[0] : ri == 1 ? refs.src_net[1] : refs.src_net[2];
    short2 *__restrict__ dst_net = ri == 0 ? refs.dst_net[0] : ri == 1 ? refs.dst_net[1] : refs.dst_net[2];
    __shared__ uint32_t s_win[S1_BLOCKS][12][4];   // prev pixels [c+v0-2, c+v0+10) in both axes, 16-byte lines
    __shared__ uint32_t s_cur[S1_BLOCKS][8][2];
    __shared__ int4 s_geo[S1_BLOCKS];              // cx, cy, vx, vy
    __shared__ unsigned s_key[S1_BLOCKS];

    const int tid = threadIdx.x;
    const int bw = width >> 3;  // == cut_width/8
    const int n0 = blockIdx.x * S1_BLOCKS;
    if (tid < S1_BLOCKS) {
        const int n = n0 + tid;
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
        s_geo[tid] = make_int4(cx, cy, vx, vy);
        s_key[tid] = 0xffffffffu;
    }
    __syncthreads();

    // stage windows and current blocks.  Pixels outside the plane are only ever used by
    // candidates that are not fully inside it, and those can never win (Q2), so the
    // coordinates are simply clamped to keep the loads in bounds.
    // One window line per thread: four aligned words cover its 12 pixels wherever they start; lines
    // that touch the left/right edge of the plane (or an unaligned plane) take the clamped byte path.
    for (int i = tid; i < S1_BLOCKS * 12; i += S1_THREADS) {
        const int b = i / 12, r = i % 12;
        const int4 g = s_geo[b];
        const int x0 = g.x + g.z - 2, y = clampi(g.y + g.w - 2 + r, 0, height - 1);
        const uint8_t *line = prev + (size_t)y * width;
        const int xa = x0 & ~3;
        if (x0 >= 0 && xa + 16 <= width && ((reinterpret_cast<uintptr_t>(prev) | (unsigned)width) & 3) == 0) {
            const uint32_t *wp = reinterpret_cast<const uint32_t *>(line + xa);
            const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3);
            const int sh = 8 * (x0 & 3);
            s_win[b][r][0] = __funnelshift_r(w0, w1, sh);
            s_win[b][r][1] = __funnelshift_r(w1, w2, sh);
            s_win[b][r][2] = __funnelshift_r(w2, w3, sh);
        } else {
            uint8_t *dst = reinterpret_cast<uint8_t *>(&s_win[b][r][0]);
            for (int c = 0; c < 12; ++c) dst[c] = __ldg(line + clampi(x0 + c, 0, width - 1));
        }
    }
    if ((width & 3) == 0) {
        if (tid < S1_BLOCKS * 16) {
            const int b = tid >> 4, r = (tid >> 1) & 7, h = tid & 1;
            const int4 g = s_geo[b];  // blocks past the end read block 0's pixels and are dropped below
            s_cur[b][r][h] = __ldg(reinterpret_cast<const uint32_t *>(cur + (size_t)(g.y + r) * width + g.x) + h);
        }
    } else {  // pyramid planes of frames whose width is not a multiple of 64: lines are not word aligned
        for (int i = tid; i < S1_BLOCKS * 64; i += S1_THREADS) {
            const int b = i >> 6, r = (i >> 3) & 7, c = i & 7;
            const int4 g = s_geo[b];
            reinterpret_cast<uint8_t *>(&s_cur[b][r][0])[c] = __ldg(cur + (size_t)(g.y + r) * width + g.x + c);
        }
    }
    __syncthreads();

    {
        const int b = tid / 20, u = tid % 20, dy = u >> 2, j = u & 3;  // 20 is a multiple of 4: j groups stay aligned
        const int sxw = j >> 1, sy = (j & 1) * 4;  // sub-block order (0,0),(0,4),(4,0),(4,4) as (x,y)
        const int4 g = s_geo[b];
        const int cx = g.x, cy = g.y, vx = g.z, vy = g.w;
        const bool live = n0 + b < nblocks;
        // column features of the current sub-block (once) and of the eight window columns the five candidates
        // cover (each once, when the first candidate needs it): see col_features in common.cuh
        uint32_t w0[4], w1[4], cw[4];
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            cw[y] = s_cur[b][sy + y][sxw];
            w0[y] = s_win[b][dy + sy + y][sxw];
            w1[y] = s_win[b][dy + sy + y][sxw + 1];
        }
        ColFeat cf[4], pf[8];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            cf[k] = col_features((int)__byte_perm(cw[0], 0, 0x4440 + k), (int)__byte_perm(cw[1], 0, 0x4440 + k),
                                 (int)__byte_perm(cw[2], 0, 0x4440 + k), (int)__byte_perm(cw[3], 0, 0x4440 + k), 14500, 7500);
        auto window_column = [&](int X) {  // X is a compile-time constant after unrolling
            const int sel = 0x4440 + (X & 3);
            return X < 4 ? col_features((int)__byte_perm(w0[0], 0, sel), (int)__byte_perm(w0[1], 0, sel),
                                        (int)__byte_perm(w0[2], 0, sel), (int)__byte_perm(w0[3], 0, sel), 0, 0)
                         : col_features((int)__byte_perm(w1[0], 0, sel), (int)__byte_perm(w1[1], 0, sel),
                                        (int)__byte_perm(w1[2], 0, sel), (int)__byte_perm(w1[3], 0, sel), 0, 0);
        };
#pragma unroll
        for (int X = 0; X < 3; ++X) pf[X] = window_column(X);
        const int py = (short)(cy + vy + dy - 2);
        const bool yok = py >= 0 && py <= height - 8;
        const int ypen = abs(abs(py - cy) - vy);
        unsigned best = 0xffffffffu;
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
            pf[dx + 3] = window_column(dx + 3);
            int o[16];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                o[k] = cf[k].e0 - pf[dx + k].e0;
                o[8 + k] = cf[k].e8 - pf[dx + k].e8;
                o[4 + k] = (cf[k].a4 - pf[dx + k].a4) >> 12;
                o[12 + k] = (cf[k].a12 - pf[dx + k].a12) >> 12;
            }
            int cost = weight4x4_rows(o);
            cost += __shfl_xor_sync(0xffffffffu, cost, 1);
            cost += __shfl_xor_sync(0xffffffffu, cost, 2);
            const int px = (short)(cx + vx + dx - 2);
            const bool inside = yok && px >= 0 && px <= width - 8;
            // the reference accumulates in an unsigned short (Q2) and adds the neighbour-coherence
            // term only at rates 2 and 1 (Q3): (| |p.x-c.x| - v0.x | + | |p.y-c.y| - v0.y |) * 32
            unsigned diff = (unsigned)cost & 0xffffu;
            if (rate < 4) diff = (diff + (unsigned)((abs(abs(px - cx) - vx) + ypen) * 32)) & 0xffffu;
            if (inside && diff < 0x7fffu) best = min(best, (diff << 8) | (unsigned)(dy * 5 + dx));
        }
        if (j == 0 && live && best != 0xffffffffu) atomicMin(&s_key[b], best);
    }
    __syncthreads();

    if (tid < S1_BLOCKS) {
        const int n = n0 + tid;
        if (n < nblocks) {
            const unsigned key = s_key[tid];
            const int4 g = s_geo[tid];
            const int cx = g.x, cy = g.y;
            int bx, by;  // "vector" of the reference: a position once a candidate has won, else v0 itself
            if (key == 0xffffffffu) {
                bx = g.z;
                by = g.w;
            } else {
                const int k = key & 255;
                bx = (short)(cx + g.z + (k % 5) - 2);
                by = (short)(cy + g.w + (k / 5) - 2);
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
// line is saturated before the vertical pass, Q5).
//
// S2_BLOCKS 8x8 blocks per CTA, 20 threads per block: a thread owns (x-phase xi, 4x4 sub-block j) and walks the
// five y-phases of its column of candidates.  The horizontally filtered lines are kept transposed (a word =
// four consecutive lines of one column), so the vertical six taps are two dp4a; the twelve lines a thread needs
// are loaded once for its five candidates (12 registers), the current sub-block once; the y-phase loop is
// unrolled, so taps and alignment shifts are immediates and the full-pel phase has no filter at all.
// The zero-vector candidate is evaluated by one warp during the second, partly filled round of the horizontal
// pass: no thread is set aside for it, the CTA is five full warps and five CTAs fit an SM.
#ifndef VP8_S2_BLOCKS
#define VP8_S2_BLOCKS 8
#endif
#ifndef VP8_S2_MINCTAS
#define VP8_S2_MINCTAS 5
#endif
#ifndef VP8_S2_UNROLL
#define VP8_S2_UNROLL 1
#endif
constexpr int S2_BLOCKS = VP8_S2_BLOCKS;
static_assert(S2_BLOCKS % 8 == 0, "the six-tap threads must fill whole warps (full-mask shuffles)");
constexpr int S2_THREADS = S2_BLOCKS * 20;
static_assert(S2_BLOCKS * 28 - S2_THREADS <= S2_THREADS - 32, "round 2 of the horizontal pass must leave a warp free");

// x/y phase variants of the five offsets -2..+2 quarter pels around a full-pel position:
// integer origin offset (-1,-1,0,0,0) and eighth-pel filter phase (4,6,0,2,4)
__device__ __forceinline__ constexpr int s2_origin(int i) { return i < 2 ? -1 : 0; }
__device__ __forceinline__ constexpr int s2_phase(int i) { return (0x46024 >> (4 * (4 - i))) & 15; }

// the six taps of eighth-pel phase PH as compile-time constants (c_sixtap, RFC 6386)
__device__ __forceinline__ constexpr int six_tap(int ph, int t) {
    return ph == 2   ? (t == 0 ? 2 : t == 1 ? -11 : t == 2 ? 108 : t == 3 ? 36 : t == 4 ? -8 : 1)
           : ph == 4 ? (t == 0 ? 3 : t == 1 ? -16 : t == 2 ? 77 : t == 3 ? 77 : t == 4 ? -16 : 3)
                     : (t == 0 ? 1 : t == 1 ? -8 : t == 2 ? 36 : t == 3 ? 108 : t == 4 ? -11 : 2);  // ph == 6
}

// unsigned pixels x signed taps
__device__ __forceinline__ int dp4a_u8s8(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ constexpr uint32_t pack_taps(int a, int b, int c, int d) {
    return (uint32_t)(a & 255) | ((uint32_t)(b & 255) << 8) | ((uint32_t)(c & 255) << 16) | ((uint32_t)(d & 255) << 24);
}

// Residual of one candidate: cur - saturate((64 + sum taps * lines) >> 7).  tl[x] = the twelve lines
// sy..sy+11 of column x as three words (line = byte), so that four taps are one dp4a; first = line of
// tap 0 of output row 0.  All shifts are compile-time constants after unrolling.
__device__ __forceinline__ constexpr uint32_t taps_lo(int ph) {
    return pack_taps(six_tap(ph, 0), six_tap(ph, 1), six_tap(ph, 2), six_tap(ph, 3));
}
__device__ __forceinline__ constexpr uint32_t taps_hi(int ph) { return pack_taps(six_tap(ph, 4), six_tap(ph, 5), 0, 0); }

// Vertical six taps without aligning the data: the twelve lines of a column sit in three words (line = byte), an
// output whose first line is byte k of word a multiplies the words a, a+1 (and a+2 when k = 3) by tap words that carry
// the taps at byte k onwards -- two or three dp4a, no funnel shift.  vtap_word(ph, k, j): the part of phase ph's taps
// that falls into word j when tap 0 sits at byte k.  All arguments are literals after unrolling.
__device__ __forceinline__ constexpr uint32_t vtap_word(int ph, int k, int j) {
    uint32_t w = 0;
    for (int i = 0; i < 6; ++i)
        if (((k + i) >> 2) == j) w |= (uint32_t)(six_tap(ph, i) & 255) << (8 * ((k + i) & 3));
    return w;
}
// four values -> four saturated bytes (a in byte 0): the six-tap's ">> 7 then clamp to 0..255" is the saturation
__device__ __forceinline__ uint32_t pack4_sat_u8(int a, int b, int c, int d) {
    uint32_t hi, r;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(hi));
    return r;
}
// the four predicted pixels (output rows 0..3, packed) of one column for y-phase ph; c = lines 0..11 of the column,
// first = line of tap 0 of output row 0
__device__ __forceinline__ uint32_t s2_pred_column(const uint32_t (&c)[3], int ph, int first) {
    int s[4];
#pragma unroll
    for (int y = 0; y < 4; ++y) {
        const int s0 = first + y, a = s0 >> 2, k = s0 & 3;
        s[y] = dp4a_u8s8(c[a + 1], vtap_word(ph, k, 1), dp4a_u8s8(c[a], vtap_word(ph, k, 0), 64));
        if (k == 3) s[y] = dp4a_u8s8(c[a + 2], vtap_word(ph, k, 2), s[y]);
    }
    return pack4_sat_u8(s[0] >> 7, s[1] >> 7, s[2] >> 7, s[3] >> 7);
}
// Pass 1 of the cost transform for one column from its packed predicted pixels and the features of the current
// sub-block's column (col_features, common.cuh): the transform is linear before its shifts, so it never sees a
// residual.  The four sums of the predicted column are dp4a with +-1 weights.
__device__ __forceinline__ void s2_column_pass1(uint32_t pred, const int4 cf, int &o0, int &o4, int &o8, int &o12) {
    const int e0 = dp4a_u8s8(pred, 0x01ff0101u, 0);  // p0 + p1 - p2 + p3   (s + u)
    const int e8 = dp4a_u8s8(pred, 0x0101ff01u, 0);  // p0 - p1 + p2 + p3   (s - u)
    const int t = dp4a_u8s8(pred, 0xff000001u, 0);   // p0 - p3
    const int x = (int)__byte_perm(pred, 0, 0x4442); // p2
    o0 = cf.x - e0;
    o8 = cf.y - e8;
    o4 = (cf.z - x * 2217 - t * 42816) >> 12;
    o12 = (cf.w - t * 17736 + x * 5352) >> 12;
}

// Staging of the reference windows, two ways (the north star names the second, so it is measured):
//   WindowsByLoads   one window line per thread: five aligned word loads + funnel shifts (the production path)
//   WindowsByTMA     one 2-D box copy (cp.async.bulk.tensor, 32 x 14 bytes from the 16-byte boundary below the window:
//                    a box cannot start at an arbitrary byte) per block out of a replicate-padded copy of the
//                    reference plane, completion through an mbarrier
struct WindowsByLoads {};
struct WindowsByTMA {
    CUtensorMap map;  // the padded plane, uint8, box 32 x 14
    int pad;          // replicated border on every side
};
template <class Staging>
__device__ __forceinline__ void luma_search_2step_body(const uint8_t *__restrict__ cur, const Search2Refs &refs, int width,
                                                       int height, int nblocks, const Staging &staging) {
    constexpr bool kTMA = sizeof(Staging) > 1;
    // words per window line.  A TMA box has to start on a 16-byte boundary of the plane, so the 14 bytes a window line
    // needs arrive somewhere inside a dense 32-byte box line and the horizontal pass shifts them out itself
    constexpr int WIN_PITCH = kTMA ? 8 : 5;
    const int ri = blockIdx.y;
    const uint8_t *__restrict__ ref = ri == 0 ? refs.ref[0] : ri == 1 ? refs.ref[1] : refs.ref[2];
    const short2 *__restrict__ net = ri == 0 ? refs.net[0] : ri == 1 ? refs.net[1] : refs.net[2];
    short2 *__restrict__ ref_net = ri == 0 ? refs.ref_net[0] : ri == 1 ? refs.ref_net[1] : refs.ref_net[2];
    int *__restrict__ ref_Bdiff = ri == 0 ? refs.ref_Bdiff[0] : ri == 1 ? refs.ref_Bdiff[1] : refs.ref_Bdiff[2];
    // ref pixels rows/cols [base-3, base+11), clamp-to-edge; 20-byte lines (TMA: 32-byte box lines, 512 bytes per block)
    __shared__ __align__(128) uint32_t s_win[S2_BLOCKS][kTMA ? 16 : 14][WIN_PITCH];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ int s_off[S2_BLOCKS];  // TMA: byte offset of the window inside its box lines (0..15)
    // horizontally filtered + saturated lines, TRANSPOSED: [x-phase][column][line], 16 lines (14 used) per column
    __shared__ __align__(4) uint8_t s_h[S2_BLOCKS][5][8][16];
    __shared__ uint32_t s_cur[S2_BLOCKS][8][2];
    __shared__ uint32_t s_zero[S2_BLOCKS][8][2];   // co-located reference block (candidate #25)
    __shared__ unsigned s_key[S2_BLOCKS];
    __shared__ int4 s_cf[S2_BLOCKS][4][4];         // pass-1 features of the current block: [sub-block][column]

    const int tid = threadIdx.x;
    const int bw = width >> 3;
    const int n0 = blockIdx.x * S2_BLOCKS;

    __shared__ int4 s_geo[S2_BLOCKS];              // bx, by, v0x, v0y of every block
    // Every staging thread derives the geometry of its block itself (one cached load and a division): no
    // separate "eight threads look up the vectors" phase with a barrier behind it.
    auto block_geometry = [&](int b) {
        const int n = min(n0 + b, nblocks - 1);
        const short2 v = __ldg(net + n);
        // short lanes in the reference; v0 is a multiple of 4
        return make_int4((n % bw) * 8, (n / bw) * 8, (short)(v.x * 4), (short)(v.y * 4));
    };
    auto geometry = [&](int b, int &bx, int &by, int &v0x, int &v0y) {
        const int4 g = s_geo[b];
        bx = g.x; by = g.y; v0x = g.z; v0y = g.w;
    };
    if (tid < S2_BLOCKS) s_key[tid] = 0xffffffffu;

    if constexpr (kTMA) {
        // one thread per block: geometry, then the box copy; everybody waits on the mbarrier below
        const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid < 32) {
            // lanes 0..7 work out where their block's window lies; ONE lane issues the copies (the TMA instruction
            // takes its operands from uniform registers: a warp cannot issue eight different ones at once)
            int x0 = 0, y0 = 0;
            if (tid < S2_BLOCKS) {
                const int4 g = block_geometry(tid);
                s_geo[tid] = g;
                x0 = g.x + (g.z >> 2) - 3 + staging.pad;
                y0 = g.y + (g.w >> 2) - 3 + staging.pad;
                s_off[tid] = x0 & 15;
                x0 &= ~15;
            }
            if (tid == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(S2_BLOCKS * 14 * 32) : "memory");
#pragma unroll 1
            for (int b = 0; b < S2_BLOCKS; ++b) {
                const int bx0 = __shfl_sync(0xffffffffu, x0, b), by0 = __shfl_sync(0xffffffffu, y0, b);
                if (tid == 0) {
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_win[b][0][0]);
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(dst), "l"(&staging.map), "r"(bx0), "r"(by0), "r"(mbar) : "memory");
                }
            }
        }
    } else {
    // one window line per thread: five aligned words cover its 14 pixels wherever they start; lines that
        // touch the left/right frame edge (or an unaligned plane) take the clamped byte path
        for (int i = tid; i < S2_BLOCKS * 14; i += S2_THREADS) {
            const int b = i / 14, r = i % 14;
            const int4 g = block_geometry(b);
            if (r == 0) s_geo[b] = g;
            const int x0 = g.x + (g.z >> 2) - 3, y = clampi(g.y + (g.w >> 2) - 3 + r, 0, height - 1);
            const uint8_t *line = ref + (size_t)y * width;
            const int xa = x0 & ~3;
            if (x0 >= 0 && xa + 20 <= width && ((reinterpret_cast<uintptr_t>(ref) | (unsigned)width) & 3) == 0) {
                const uint32_t *wp = reinterpret_cast<const uint32_t *>(line + xa);
                const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3), w4 = __ldg(wp + 4);
                const int sh = 8 * (x0 & 3);
                s_win[b][r][0] = __funnelshift_r(w0, w1, sh);
                s_win[b][r][1] = __funnelshift_r(w1, w2, sh);
                s_win[b][r][2] = __funnelshift_r(w2, w3, sh);
                s_win[b][r][3] = __funnelshift_r(w3, w4, sh);
            } else {
                uint8_t *dst = reinterpret_cast<uint8_t *>(&s_win[b][r][0]);
                for (int c = 0; c < 14; ++c) dst[c] = __ldg(line + clampi(x0 + c, 0, width - 1));
            }
        }
    }
    // (S2_THREADS - S2_BLOCKS * 16 = 32 < 112: going backwards spreads the two jobs over the threads)
    if (tid >= S2_THREADS - S2_BLOCKS * 16) {
        const int t = S2_THREADS - 1 - tid;
        const int b = t >> 4, r = (t >> 1) & 7, h = t & 1;
        const int n = min(n0 + b, nblocks - 1);
        const int gx = (n % bw) * 8, gy = (n / bw) * 8;
        s_cur[b][r][h] = __ldg(reinterpret_cast<const uint32_t *>(cur + (size_t)(gy + r) * width + gx) + h);
        s_zero[b][r][h] = __ldg(reinterpret_cast<const uint32_t *>(ref + (size_t)(gy + r) * width + gx) + h);
    }
    if constexpr (kTMA) {
        const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(mbar) : "memory");
    }
    __syncthreads();

    // horizontal pass: 5 x-variants x 14 rows x 8 columns per block.  An item is (block, row, half line): its
    // three window words give all five variants of four columns.  An output's six taps are two dp4a on
    // funnel-shifted words; the two variants with origin -1 start at window byte 0, the two with origin 0 at
    // byte 1, so six funnel shifts serve all sixteen filtered outputs.  The full-pel variant is a byte copy (its
    // tap 128 does not fit a signed byte).  (s/128 then saturate) == saturate(s>>7): the two only differ for
    // -128<s<0, both give 0.
    for (int i = tid; i < S2_BLOCKS * 28; i += S2_THREADS) {
        const int b = i / 28, q = i % 28, r = q >> 1, h = q & 1;
        uint32_t wa, wb, wc;
        if constexpr (kTMA) {
            const int o = s_off[b], q = (o >> 2) + h, sh = 8 * (o & 3);
            const uint32_t t0 = s_win[b][r][q], t1 = s_win[b][r][q + 1], t2 = s_win[b][r][q + 2], t3 = s_win[b][r][q + 3];
            wa = __funnelshift_r(t0, t1, sh);
            wb = __funnelshift_r(t1, t2, sh);
            wc = __funnelshift_r(t2, t3, sh);
        } else {
            wa = s_win[b][r][h];
            wb = s_win[b][r][h + 1];
            wc = s_win[b][r][h + 2];
        }
        const uint32_t lo[5] = {wa, __funnelshift_r(wa, wb, 8), __funnelshift_r(wa, wb, 16), __funnelshift_r(wa, wb, 24), wb};
        const uint32_t hi[5] = {wb, __funnelshift_r(wb, wc, 8), __funnelshift_r(wb, wc, 16), __funnelshift_r(wb, wc, 24), wc};
        uint8_t *dst = &s_h[b][0][4 * h][r];  // + 128 per variant, + 16 per column
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            dst[0 * 128 + 16 * c] = (uint8_t)sat8(dp4a_u8s8(hi[c], taps_hi(4), dp4a_u8s8(lo[c], taps_lo(4), 64)) >> 7);
            dst[1 * 128 + 16 * c] = (uint8_t)sat8(dp4a_u8s8(hi[c], taps_hi(6), dp4a_u8s8(lo[c], taps_lo(6), 64)) >> 7);
            dst[2 * 128 + 16 * c] = (uint8_t)(lo[3] >> (8 * c));  // tap 2 of outputs 0..3 = window bytes 3..6
            dst[3 * 128 + 16 * c] = (uint8_t)sat8(dp4a_u8s8(hi[c + 1], taps_hi(2), dp4a_u8s8(lo[c + 1], taps_lo(2), 64)) >> 7);
            dst[4 * 128 + 16 * c] = (uint8_t)sat8(dp4a_u8s8(hi[c + 1], taps_hi(4), dp4a_u8s8(lo[c + 1], taps_lo(4), 64)) >> 7);
        }
    }
    // pass-1 features of every column of the current sub-blocks (the candidates only subtract theirs)
    for (int i = tid; i < S2_BLOCKS * 16; i += S2_THREADS) {
        const int b = i >> 4, j = (i >> 2) & 3, k = i & 3, sx = j >> 1, sy = (j & 1) * 4;
        const uint32_t sel = 0x4440 + k;
        const ColFeat f = col_features((int)__byte_perm(s_cur[b][sy][sx], 0, sel), (int)__byte_perm(s_cur[b][sy + 1][sx], 0, sel),
                                       (int)__byte_perm(s_cur[b][sy + 2][sx], 0, sel), (int)__byte_perm(s_cur[b][sy + 3][sx], 0, sel),
                                       14500, 7500);
        s_cf[b][j][k] = make_int4(f.e0, f.e8, f.a4, f.a12);
    }
    // candidate #25, the zero vector (full-pel, no penalty), on the last warp, which has no item in round 2:
    // four threads per block, one 4x4 sub-block each
    if (tid >= S2_THREADS - 32) {
        const int t = tid - (S2_THREADS - 32), b = t >> 2, j = t & 3;
        const int sx = j >> 1, sy = (j & 1) * 4;
        int r[16];
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const uint32_t cw = s_cur[b][sy + y][sx], zw = s_zero[b][sy + y][sx];
#pragma unroll
            for (int x = 0; x < 4; ++x) r[4 * y + x] = (int)__byte_perm(cw, 0, 0x4440 + x) - (int)__byte_perm(zw, 0, 0x4440 + x);
        }
        int cost = weight4x4(r);
        cost += __shfl_xor_sync(0xffffffffu, cost, 1);
        cost += __shfl_xor_sync(0xffffffffu, cost, 2);
        if (j == 0 && n0 + b < nblocks && cost < 0x7fff) atomicMin(&s_key[b], ((unsigned)cost << 5) | 25u);
    }
    __syncthreads();

    {
        const int b = tid / 20, u = tid % 20;
        const int xi = u >> 2, j = tid & 3;             // 20 is a multiple of 4
        const int sx = j >> 1, sy = (j & 1) * 4;        // sub-block (x word, y row offset); order as dx4/dy4
        const bool live = n0 + b < nblocks;
        int bx, by, v0x, v0y;
        geometry(b, bx, by, v0x, v0y);
        unsigned best = 0xffffffffu;
        uint32_t tl[4][3];  // lines sy .. sy+11 of the four columns of this x-phase (lines >= 14 are never used)
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const uint32_t *cw = reinterpret_cast<const uint32_t *>(&s_h[b][xi][4 * sx + x][sy]);
            tl[x][0] = cw[0]; tl[x][1] = cw[1]; tl[x][2] = cw[2];
        }
        const int qx = (short)(bx * 4 + v0x + xi - 2);
        const bool xok = qx >= 0 && qx <= width * 4 - 32;
        // y-phases 4,6,0,2,4 with first line 0,0,1,1,1 (s2_origin + 1)
#if VP8_S2_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
        for (int yi = 0; yi < 5; ++yi) {
            int o[16];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                uint32_t pred;
                if (yi == 0) pred = s2_pred_column(tl[x], 4, 0);
                else if (yi == 1) pred = s2_pred_column(tl[x], 6, 0);
                else if (yi == 2) pred = __funnelshift_r(tl[x][0], tl[x][1], 24);  // full-pel phase: lines 3..6 themselves
                else if (yi == 3) pred = s2_pred_column(tl[x], 2, 1);
                else pred = s2_pred_column(tl[x], 4, 1);
                s2_column_pass1(pred, s_cf[b][j][x], o[x], o[4 + x], o[8 + x], o[12 + x]);
            }
            int cost = weight4x4_rows(o);
            cost += __shfl_xor_sync(0xffffffffu, cost, 1);
            cost += __shfl_xor_sync(0xffffffffu, cost, 2);
            cost += (abs(xi - 2) + abs(yi - 2)) * 32;
            const int qy = (short)(by * 4 + v0y + yi - 2);
            const bool valid = xok && qy >= 0 && qy <= height * 4 - 32;
            if (valid && cost < 0x7fff) best = min(best, ((unsigned)cost << 5) | (unsigned)(yi * 5 + xi));
        }
        if (j == 0 && live && best != 0xffffffffu) atomicMin(&s_key[b], best);
    }
    __syncthreads();

    if (tid < S2_BLOCKS && n0 + tid < nblocks) {
        const int n = n0 + tid;
        int bx, by, v0x, v0y;
        geometry(tid, bx, by, v0x, v0y);
        const unsigned key = s_key[tid];
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

__global__ void __launch_bounds__(S2_THREADS, VP8_S2_MINCTAS)
k_luma_search_2step(const uint8_t *__restrict__ cur, Search2Refs refs, int width, int height, int nblocks) {
    luma_search_2step_body(cur, refs, width, height, nblocks, WindowsByLoads());
}
__global__ void __launch_bounds__(S2_THREADS, VP8_S2_MINCTAS)
k_luma_search_2step_tma(const uint8_t *__restrict__ cur, Search2Refs refs, int width, int height, int nblocks,
                        const __grid_constant__ WindowsByTMA staging) {
    luma_search_2step_body(cur, refs, width, height, nblocks, staging);
}

// replicate-padded copy of a plane (the TMA variant reads its windows from it: a box copy fills what lies outside the
// tensor with zeros, the search wants the edge pixel; `pad` >= 4 covers every candidate that can win)
__global__ void k_pad_replicate(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int width, int height, int pad) {
    const int pw = width + 2 * pad, ph = height + 2 * pad;
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x4 >= pw || y >= ph) return;
    const uint8_t *line = src + (size_t)clampi(y - pad, 0, height - 1) * width;
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) w |= (uint32_t)line[clampi(x4 + k - pad, 0, width - 1)] << (8 * k);
    *reinterpret_cast<uint32_t *>(dst + (size_t)y * pw + x4) = w;
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
    dim3 grid((nblocks + S2_BLOCKS - 1) / S2_BLOCKS, nrefs);
    k_luma_search_2step<<<grid, S2_THREADS, 0, (cudaStream_t)stream>>>(cur, r, width, height, nblocks);
    VP8_LAUNCH_CHECK();
}

// ---- experiment: luma_search_2step with TMA-staged windows (see WindowsByTMA) ---------------------------------
// Pads `ref` (replicated border of 16 pixels) into `padded` (caller-provided, (width+32)*(height+32) bytes), builds
// the tensor map and runs the search.  Results are bit-identical to vp8b200_luma_search_2step (tested); the point is
// the A/B timing of the two staging methods (tools/me_tma_ab.py).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
extern "C" int vp8b200_experiment_search_2step_tma(void *stream, const uint8_t *cur, const uint8_t *ref, uint8_t *padded,
                                                   const int16_t *net, int16_t *ref_net, int32_t *ref_Bdiff, int width,
                                                   int height, int pad_only) {
    const int pad = 16;
    const int nblocks = width * height / 64;
    if (nblocks <= 0 || (width & 15)) return -(int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    const int pw = width + 2 * pad, ph = height + 2 * pad;
    k_pad_replicate<<<dim3((pw / 4 + 127) / 128, ph), 128, 0, st>>>(ref, padded, width, height, pad);
    if (pad_only) VP8_LAUNCH_CHECK();
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult q;
        void *fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -1;
        encode = (EncodeTiledFn)fn;
    }
    WindowsByTMA staging;
    staging.pad = pad;
    const cuuint64_t dims[2] = {(cuuint64_t)pw, (cuuint64_t)ph}, strides[1] = {(cuuint64_t)pw};
    const cuuint32_t box[2] = {32, 14}, estr[2] = {1, 1};
    if (encode(&staging.map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, padded, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -2;
    Search2Refs r = {};
    r.ref[0] = ref;  // (the co-located block of the zero-vector candidate still comes from the plane itself)
    r.net[0] = (const short2 *)net;
    r.ref_net[0] = (short2 *)ref_net;
    r.ref_Bdiff[0] = ref_Bdiff;
    k_luma_search_2step_tma<<<dim3((nblocks + S2_BLOCKS - 1) / S2_BLOCKS, 1), S2_THREADS, 0, st>>>(cur, r, width, height, nblocks, staging);
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
