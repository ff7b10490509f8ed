// Intra (key-frame) macroblock path of the vp8oclenc_b200 engine (sm_100a), SURVEY.md 8f-4:
// intra_transform() / predict_and_transform_mb() of the reference host (src/intra_part.h:37-741, 1089-1128), which
// the reference runs as plain C on one host thread, macroblock after macroblock in raster order.
//
// What constrains the order is only that a macroblock is predicted from RECONSTRUCTED neighbours: its left
// neighbour, the row above and -- through the above-right pixels of its last sub-block column -- the macroblock
// above-right.  So macroblock (r, c) may run once (r, c-1) and (r-1, c+1) are done: the same wavefront as the loop
// filter, c + 2r.  The kernel runs one CTA per macroblock row:
//   * rows are handed out through an atomic ticket in launch order (a row's predecessor is always running) and
//     publish the number of macroblocks they have finished (release store); the row below waits for it to be two
//     ahead (acquire load) and reads the pixels above it past the L1;
//   * warp 0 codes the luma: the sixteen 4x4 sub-blocks depend on each other the same way (left, above,
//     above-right), so they run as a wavefront of ten steps, two sub-blocks per step, one on each half of the warp.
//     Within a half, lane m evaluates sub-block mode m: predictor (every directional mode is a table of 3-tap
//     averages over the thirteen edge pixels), residual, forward DCT, weight; the lowest (weight, mode) wins by
//     shuffles; the winning lane already holds the transform, quantises, dequantises, inverts and stores;
//   * warp 1 codes the chroma (TM_PRED from the macroblock's border, eight independent 4x4 blocks on eight lanes).
// Arithmetic follows the reference: 16-bit stores inside both transforms, rounding of coefficient 11 by the sign of
// coefficient 10, ties between modes to the earlier one.  Pinned through oracle/vp8_oracle_intra.c against the
// reference's own code (oracle/ref_intra.cpp).
#include "common.cuh"

namespace vp8 {

// (a, b, c) edge indices of every predictor pixel of the eight directional modes + DC (index 13 = the DC value):
// pixel = (E[a] + 2 E[b] + E[c] + 2) >> 2, which also covers the two-tap averages (a == c) and plain copies.
// Edge order: 0..3 = L3 L2 L1 L0, 4 = P (above-left), 5..12 = A0..A7.  Mode 1 (TM) is computed separately.
struct IntraTaps { unsigned short t[10][16]; };
__host__ __device__ constexpr unsigned short tap(int a, int b, int c) { return (unsigned short)(a | (b << 4) | (c << 8)); }
__host__ __device__ constexpr unsigned short avg3(int k) { return tap(k, k + 1, k + 2); }
__host__ __device__ constexpr unsigned short avg2(int k) { return tap(k, k + 1, k); }
constexpr int A0 = 5, L0 = 3, L1 = 2, L2 = 1, L3 = 0;
__constant__ IntraTaps c_intra_taps = {{
    // B_DC_PRED
    {tap(13, 13, 13), tap(13, 13, 13), tap(13, 13, 13), tap(13, 13, 13), tap(13, 13, 13), tap(13, 13, 13), tap(13, 13, 13), tap(13, 13, 13),
     tap(13, 13, 13), tap(13, 13, 13), tap(13, 13, 13), tap(13, 13, 13), tap(13, 13, 13), tap(13, 13, 13), tap(13, 13, 13), tap(13, 13, 13)},
    // B_TM_PRED (placeholder, not used)
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    // B_VE_PRED: column c = avg3(P/A[c-1], A[c], A[c+1])
    {avg3(4), avg3(5), avg3(6), avg3(7), avg3(4), avg3(5), avg3(6), avg3(7), avg3(4), avg3(5), avg3(6), avg3(7), avg3(4), avg3(5), avg3(6), avg3(7)},
    // B_HE_PRED: row r = avg3(P/L[r-1], L[r], L[r+1]); last row (L2 + 3 L3 + 2) >> 2
    {tap(4, L0, L1), tap(4, L0, L1), tap(4, L0, L1), tap(4, L0, L1), tap(L0, L1, L2), tap(L0, L1, L2), tap(L0, L1, L2), tap(L0, L1, L2),
     tap(L1, L2, L3), tap(L1, L2, L3), tap(L1, L2, L3), tap(L1, L2, L3), tap(L2, L3, L3), tap(L2, L3, L3), tap(L2, L3, L3), tap(L2, L3, L3)},
    // B_LD_PRED: avg3 of A[r+c ..], last one (A6 + 3 A7 + 2) >> 2
    {avg3(A0 + 0), avg3(A0 + 1), avg3(A0 + 2), avg3(A0 + 3), avg3(A0 + 1), avg3(A0 + 2), avg3(A0 + 3), avg3(A0 + 4),
     avg3(A0 + 2), avg3(A0 + 3), avg3(A0 + 4), avg3(A0 + 5), avg3(A0 + 3), avg3(A0 + 4), avg3(A0 + 5), tap(A0 + 6, A0 + 7, A0 + 7)},
    // B_RD_PRED: avg3 of the edge at 3 - r + c
    {avg3(3), avg3(4), avg3(5), avg3(6), avg3(2), avg3(3), avg3(4), avg3(5), avg3(1), avg3(2), avg3(3), avg3(4), avg3(0), avg3(1), avg3(2), avg3(3)},
    // B_VR_PRED
    {avg2(4), avg2(5), avg2(6), avg2(7), avg3(3), avg3(4), avg3(5), avg3(6), avg3(2), avg2(4), avg2(5), avg2(6), avg3(1), avg3(3), avg3(4), avg3(5)},
    // B_VL_PRED
    {avg2(A0 + 0), avg2(A0 + 1), avg2(A0 + 2), avg2(A0 + 3), avg3(A0 + 0), avg3(A0 + 1), avg3(A0 + 2), avg3(A0 + 3),
     avg2(A0 + 1), avg2(A0 + 2), avg2(A0 + 3), avg3(A0 + 4), avg3(A0 + 1), avg3(A0 + 2), avg3(A0 + 3), avg3(A0 + 5)},
    // B_HD_PRED
    {avg2(3), avg3(3), avg3(4), avg3(5), avg2(2), avg3(2), avg2(3), avg3(3), avg2(1), avg3(1), avg2(2), avg3(2), avg2(0), avg3(0), avg2(1), avg3(1)},
    // B_HU_PRED (edge walks L0 -> L3, i.e. indices downwards; the tail repeats L3)
    {tap(L0, L1, L0), tap(L0, L1, L2), tap(L1, L2, L1), tap(L1, L2, L3), tap(L1, L2, L1), tap(L1, L2, L3), tap(L2, L3, L2), tap(L2, L3, L3),
     tap(L2, L3, L2), tap(L2, L3, L3), tap(L3, L3, L3), tap(L3, L3, L3), tap(L3, L3, L3), tap(L3, L3, L3), tap(L3, L3, L3), tap(L3, L3, L3)},
}};

__device__ __forceinline__ int s16(int v) { return (int)(short)v; }

// forward DCT of a 4x4 residual with the reference's 16-bit stores (src/intra_part.h:124-158)
__device__ __forceinline__ void intra_fdct(const int (&r)[16], int (&f)[16]) {
    int t[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int a = (r[4 * i] + r[4 * i + 3]) << 3, b = (r[4 * i + 1] + r[4 * i + 2]) << 3;
        const int c = (r[4 * i + 1] - r[4 * i + 2]) << 3, d = (r[4 * i] - r[4 * i + 3]) << 3;
        t[4 * i] = s16(a + b);
        t[4 * i + 2] = s16(a - b);
        t[4 * i + 1] = s16((c * 2217 + d * 5352 + 14500) >> 12);
        t[4 * i + 3] = s16((d * 2217 - c * 5352 + 7500) >> 12);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int a = t[i] + t[12 + i], b = t[4 + i] + t[8 + i], c = t[4 + i] - t[8 + i], d = t[i] - t[12 + i];
        f[i] = s16((a + b + 7) >> 4);
        f[8 + i] = s16((a - b + 7) >> 4);
        f[4 + i] = s16(((c * 2217 + d * 5352 + 12000) >> 16) + (d != 0));
        f[12 + i] = s16((d * 2217 - c * 5352 + 51000) >> 16);
    }
}

// quantise (:212-250), then dequantise + inverse DCT + predictor (:40-122); f becomes the quantised block
__device__ __forceinline__ void intra_quant_recon(int (&f)[16], const int (&pred)[16], int (&out)[16], int dc_q, int ac_q,
                                                  uint32_t m_dc, uint32_t m_ac) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int q = i == 0 ? dc_q : ac_q;
        const int sign_of = i == 11 ? f[10] : f[i];  // (coefficient 11 rounds the way coefficient 10 -- already rounded -- points)
        f[i] = s16(f[i] + (sign_of < 0 ? -q / 2 : q / 2));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = div_magic(f[i], i == 0 ? m_dc : m_ac);  // (exact: |f| < 2^15, see common.cuh)
    int t[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int i0 = f[i] * (i == 0 ? dc_q : ac_q), i4 = f[4 + i] * ac_q, i8 = f[8 + i] * ac_q, i12 = f[12 + i] * ac_q;
        const int a = i0 + i8, b = i0 - i8;
        const int c = ((i4 * 35468) >> 16) - (i12 + ((i12 * 20091) >> 16));
        const int d = (i4 + ((i4 * 20091) >> 16)) + ((i12 * 35468) >> 16);
        t[i] = s16(a + d);
        t[12 + i] = s16(a - d);
        t[4 + i] = s16(b + c);
        t[8 + i] = s16(b - c);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p0 = t[4 * i], p1 = t[4 * i + 1], p2 = t[4 * i + 2], p3 = t[4 * i + 3];
        const int a = p0 + p2, b = p0 - p2;
        const int c = ((p1 * 35468) >> 16) - (p3 + ((p3 * 20091) >> 16));
        const int d = (p1 + ((p1 * 20091) >> 16)) + ((p3 * 35468) >> 16);
        out[4 * i] = sat8(s16(((a + d + 4) >> 3) + pred[4 * i]));
        out[4 * i + 3] = sat8(s16(((a - d + 4) >> 3) + pred[4 * i + 3]));
        out[4 * i + 1] = sat8(s16(((b + c + 4) >> 3) + pred[4 * i + 1]));
        out[4 * i + 2] = sat8(s16(((b - c + 4) >> 3) + pred[4 * i + 2]));
    }
}

// The same on the sixteen lanes of a half-warp, lane i owning coefficient / pixel i (i = 4 row + column): returns the
// reconstructed pixel and stores the quantised coefficient at its zig-zag position of dst (if not null).  Every lane
// of the warp has to call it (shuffles in segments of 16).
__device__ __forceinline__ int intra_quant_recon_lane(int f, int pred, int i, int dc_q, int ac_q, uint32_t m_dc, uint32_t m_ac,
                                                      int16_t *dst) {
    const int q = i == 0 ? dc_q : ac_q, hq = q / 2;
    const int own = s16(f + (f < 0 ? -hq : hq));
    const int ten = __shfl_sync(0xffffffffu, own, 10, 16);  // coefficient 11 rounds the way the rounded coefficient 10 points
    const int r = i == 11 ? s16(f + (ten < 0 ? -hq : hq)) : own;
    const int qv = div_magic(r, i == 0 ? m_dc : m_ac);
    if (dst) dst[(int)((0xFEA9DB83C7426510ull >> (4 * i)) & 15)] = (int16_t)qv;  // inverse of the zig-zag scan
    const int dq = qv * q;
    const int col = i & 3, row = i >> 2;
    {   // vertical pass: the column's four coefficients
        const int i0 = __shfl_sync(0xffffffffu, dq, col, 16), i4 = __shfl_sync(0xffffffffu, dq, 4 + col, 16);
        const int i8 = __shfl_sync(0xffffffffu, dq, 8 + col, 16), i12 = __shfl_sync(0xffffffffu, dq, 12 + col, 16);
        const int a = i0 + i8, b = i0 - i8;
        const int c = ((i4 * 35468) >> 16) - (i12 + ((i12 * 20091) >> 16));
        const int d = (i4 + ((i4 * 20091) >> 16)) + ((i12 * 35468) >> 16);
        const int outer = (row == 0 || row == 3), p = outer ? a : b, s = outer ? d : c;
        f = s16(row < 2 ? (row == 0 ? p + s : p + s) : p - s);  // rows 0..3: a + d, b + c, b - c, a - d
    }
    {   // horizontal pass: the line's four values
        const int p0 = __shfl_sync(0xffffffffu, f, 4 * row, 16), p1 = __shfl_sync(0xffffffffu, f, 4 * row + 1, 16);
        const int p2 = __shfl_sync(0xffffffffu, f, 4 * row + 2, 16), p3 = __shfl_sync(0xffffffffu, f, 4 * row + 3, 16);
        const int a = p0 + p2, b = p0 - p2;
        const int c = ((p1 * 35468) >> 16) - (p3 + ((p3 * 20091) >> 16));
        const int d = (p1 + ((p1 * 20091) >> 16)) + ((p3 * 35468) >> 16);
        const int outer = (col == 0 || col == 3), p = outer ? a : b, s = outer ? d : c;
        const int v = col < 2 ? p + s : p - s;  // columns 0..3: a + d, b + c, b - c, a - d
        return sat8(s16(((v + 4) >> 3) + pred));
    }
}

// the quantised block in zig-zag order as eight packed words
__device__ __forceinline__ void store_zigzag(int16_t *dst, const int (&f)[16]) {
    constexpr int zz[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = (uint32_t)(f[zz[2 * i]] & 0xffff) | ((uint32_t)f[zz[2 * i + 1]] << 16);
    uint4 *d = reinterpret_cast<uint4 *>(dst);  // (records are 800 bytes, blocks 32: 16-byte aligned)
    d[0] = make_uint4(w[0], w[1], w[2], w[3]);
    d[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_cg_u8(const uint8_t *p) {
    uint32_t v;
    asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct IntraQuants { int y_dc, y_ac, uv_dc, uv_ac; };

// ctrl[0] = ticket, ctrl[1 + r] = luma macroblocks finished in row r, ctrl[1 + mbh + r] = chroma macroblocks finished
#if defined(INTRA_EXP_TIMELINE)
__device__ unsigned long long g_timeline[4][128][16];
__device__ long long g_phase[10][8];  // clock64 inside the ten sub-block steps of macroblock (0, 5)
#define PHASE_MARK(k) do { if (lane == 0 && r == 0 && c == 5) g_phase[t][k] = clock64(); } while (0)
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#else
#define PHASE_MARK(k) do { } while (0)
#endif
__global__ void __launch_bounds__(64)
k_intra_frame(const uint8_t *__restrict__ cur_y, const uint8_t *__restrict__ cur_u, const uint8_t *__restrict__ cur_v,
              uint8_t *rec_y, uint8_t *rec_u, uint8_t *rec_v, int16_t *__restrict__ MB, int *__restrict__ modes,
              int *__restrict__ parts, int *__restrict__ segment_id, int width, int height, IntraQuants q, int *ctrl) {
    const int mbw = width >> 4, mbh = height >> 4, cw = width >> 1;
    __shared__ int s_row;
    // luma tile of the macroblock with its border, pixel (x, y) at ty[1 + y][XO + x] (XO = 4 keeps the 4-pixel groups
    // word aligned): ty[0][XO - 1] the corner, ty[0][XO .. XO + 19] the row above (with above-right), ty[1 + y][XO - 1]
    // the column to the left; rows 4, 8, 12 carry the above-right pixels of the sub-block column 3 at [XO + 16 ..]
    // (copies of the row above the macroblock, as every VP8 decoder uses)
    constexpr int XO = 4;
    __shared__ __align__(16) uint8_t ty[17][28];
    __shared__ __align__(16) uint8_t sy[16][16];  // source luma
    __shared__ int s_edge[2][16];      // per half-warp: L3 L2 L1 L0 P A0..A7, DC value
    __shared__ __align__(16) int s_blk[2][32];  // per half-warp: the winning mode's coefficients and predictor
    // the tap table, transposed ([pixel][mode]): the ten lanes of a half-warp read ten neighbouring entries (from
    // constant memory their ten different addresses would be served one after the other)
    __shared__ uint32_t s_idx4[4][16];  // per mode: for each predicted pixel the entry of s_val it is (bytes, four pixels per word)
    __shared__ int s_val[2][28];        // per half-warp: the values a directional predictor can be, see below
    __shared__ uint8_t tc[2][9][12];   // chroma tiles with border, per plane
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s_row = atomicAdd(&ctrl[0], 1);
    // Every directional predictor pixel is (E[a] + 2 E[b] + E[c] + 2) >> 2 over the 13 edge pixels E (c_intra_taps), and
    // only 27 different values occur per sub-block: T3[k] = the three-tap average centred on E[k] with the ends
    // replicated (k = 0..12), T2[k] = (E[k] + E[k+1] + 1) >> 1 (k = 0..11, stored at 13 + k), the DC value (25) and E[0]
    // itself (26).  The half-warp computes them once per sub-block; a mode's pixel is then one indexed load.
    for (int i = tid; i < 160; i += 64) {
        const int mode = i >> 4, px = i & 15, t = c_intra_taps.t[mode][px];
        const int a = t & 15, b = (t >> 4) & 15, c3 = t >> 8;
        const int idx = (a == b && b == c3) ? (a == 13 ? 25 : 26) : (a == c3 ? 13 + min(a, b) : (b == c3 ? (a < b ? 12 : 0) : min(a, c3) + 1));
        reinterpret_cast<uint8_t *>(&s_idx4[px >> 2][mode])[px & 3] = (uint8_t)idx;
    }
    __syncthreads();
    const int r = s_row;
    if (r >= mbh) return;
    int *done_y = ctrl + 1, *done_c = ctrl + 1 + mbh;
    const uint32_t m_y_dc = magic(q.y_dc), m_y_ac = magic(q.y_ac), m_uv_dc = magic(q.uv_dc), m_uv_ac = magic(q.uv_ac);

    if (warp == 0) {
        // ------------------------------------------------ luma ------------------------------------------------
        const int h = lane >> 4, m = lane & 15;  // half-warp, sub-block mode evaluated by this lane
        // source pixels, 8 bytes per lane, fetched one macroblock ahead (the walk along the row is latency-bound)
        const uint8_t *src_lane = cur_y + (size_t)(16 * r + (lane >> 1)) * width + 8 * (lane & 1);
        uint2 src_next = *reinterpret_cast<const uint2 *>(src_lane);
        for (int c = 0; c < mbw; ++c) {
            const int mb = r * mbw + c;
#if defined(INTRA_EXP_TIMELINE)
            if (lane == 0 && r < 4 && c < 128) g_timeline[r][c][0] = gtime();
#endif
            if (r > 0) {
                const int need = min(c + 2, mbw);
                // every lane polls (one broadcast load): a single polling lane leaves the warp split for the rest of the
                // macroblock and every instruction issues twice
                while (ld_acquire_gpu(&done_y[r - 1]) < need) {}
                __syncwarp();
            }
#if defined(INTRA_EXP_TIMELINE)
            if (lane == 0 && r < 4 && c < 128) g_timeline[r][c][1] = gtime();
#endif
            // left column: the previous macroblock's last column (still in the tile), 129 at the frame edge
            if (lane < 16) ty[1 + lane][XO - 1] = c == 0 ? (uint8_t)129 : ty[1 + lane][XO + 15];
            __syncwarp();
            // the row above (20 pixels) and the corner
            if (lane < 21) {
                uint32_t v;
                if (r == 0) {
                    v = 127;
                } else if (lane == 0) {
                    v = c == 0 ? 129u : ld_cg_u8(rec_y + (size_t)(16 * r - 1) * width + 16 * c - 1);
                } else {
                    const int x = (c == mbw - 1 && lane > 16) ? 15 : lane - 1;  // last column: repeat the last pixel
                    v = ld_cg_u8(rec_y + (size_t)(16 * r - 1) * width + 16 * c + x);
                }
                ty[0][XO - 1 + lane] = (uint8_t)v;
            }
            *reinterpret_cast<uint2 *>(&sy[lane >> 1][8 * (lane & 1)]) = src_next;
            if (c + 1 < mbw) src_next = *reinterpret_cast<const uint2 *>(src_lane + 16 * (c + 1));
            __syncwarp();
            if (lane < 12) ty[4 + 4 * (lane >> 2)][XO + 16 + (lane & 3)] = ty[0][XO + 16 + (lane & 3)];
            __syncwarp();
#if defined(INTRA_EXP_TIMELINE)
            if (lane == 0 && r < 4 && c < 128) g_timeline[r][c][2] = gtime();
#endif

#pragma unroll 1
            for (int t = 0; t < 10; ++t) {
                // sub-blocks with bc + 2 br == t; the half-warp with the smaller br first
                const int br0 = max(0, (t - 2) >> 1) , nvalid = (min(3, t >> 1) - br0) + 1;
                const int br = br0 + h, bc = t - 2 * br;
                const bool valid = h < nvalid && bc >= 0 && bc <= 3;
                const int y0 = 4 * br, x0 = 4 * bc;
                PHASE_MARK(0);
                {
                    // edge pixels of the sub-block out of the tile: m = 0..3 -> L3..L0, 4 -> P, 5..12 -> A0..A7
                    int v = 0;
                    if (valid && m < 13) {
                        if (m < 4) v = ty[1 + y0 + (3 - m)][XO - 1 + x0];
                        else if (m == 4) v = ty[y0][XO - 1 + x0];
                        else v = ty[y0][XO + x0 + (m - 5)];
                    }
                    const int up = __shfl_sync(0xffffffffu, v, max(m - 1, 0), 16), dn = __shfl_sync(0xffffffffu, v, min(m + 1, 12), 16);
                    int dc = (m < 4 || (m >= 5 && m < 9)) ? v : 0;
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) dc += __shfl_xor_sync(0xffffffffu, dc, o);
                    if (valid && m < 13) {
                        s_edge[h][m] = v;  // (B_TM_PRED works on the pixels themselves)
                        s_val[h][m] = (up + 2 * v + dn + 2) >> 2;
                        if (m < 12) s_val[h][13 + m] = (v + dn + 1) >> 1;
                        if (m == 0) s_val[h][26] = v;
                    }
                    if (valid && m == 13) s_val[h][25] = (dc + 4) >> 3;
                }
                __syncwarp();
                PHASE_MARK(1);
                int pred[16], f[16];
                int key = 0x7fffffff;
                if (valid && m < 10) {
                    int res[16];
                    if (m == 1) {
                        const int P = s_edge[h][4];
#pragma unroll
                        for (int i = 0; i < 16; ++i) pred[i] = sat8(s_edge[h][5 + (i & 3)] + s_edge[h][3 - (i >> 2)] - P);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t ix = s_idx4[j][m];
#pragma unroll
                            for (int k = 0; k < 4; ++k) pred[4 * j + k] = s_val[h][(ix >> (8 * k)) & 255];
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) res[i] = (int)sy[y0 + (i >> 2)][x0 + (i & 3)] - pred[i];
                    intra_fdct(res, f);
                    int wgt = abs(s16(f[0] / 4));
#pragma unroll
                    for (int i = 1; i < 16; ++i) wgt = (int)__sad(f[i], 0, (unsigned)wgt);  // |f| + wgt: one VABSDIFF
                    key = (s16(wgt) << 4) | m;  // ties: the earlier mode
                }
                PHASE_MARK(2);
                int best = key;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
                PHASE_MARK(3);
                // The winning lane hands its coefficients and predictor to the sixteen lanes of its half-warp: quantising,
                // the inverse transform (two passes, each lane one output, inputs by shuffles) and the stores on sixteen
                // lanes instead of one (this tail was 58 % of a step when the winner did it alone).
                if (valid && key == best) {  // exactly one lane of the half-warp
                    int4 *blk = reinterpret_cast<int4 *>(s_blk[h]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        blk[i] = make_int4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                        blk[4 + i] = make_int4(pred[4 * i], pred[4 * i + 1], pred[4 * i + 2], pred[4 * i + 3]);
                    }
                    modes[16 * mb + 4 * br + bc] = m;
                }
                __syncwarp();
                {
                    const int out = intra_quant_recon_lane(s_blk[h][m], s_blk[h][16 + m], m, q.y_dc, q.y_ac, m_y_dc, m_y_ac,
                                                           valid ? MB + (size_t)mb * 400 + 16 * (4 * br + bc) : nullptr);
                    if (valid) ty[1 + y0 + (m >> 2)][XO + x0 + (m & 3)] = (uint8_t)out;
                }
                __syncwarp();
                PHASE_MARK(4);
#if defined(INTRA_EXP_TIMELINE)
                if (lane == 0 && r < 4 && c < 128) g_timeline[r][c][3 + t] = gtime();
#endif
            }
            // the macroblock's reconstruction to the frame; then tell the row below
            {
                const int row = lane >> 1, half = lane & 1;
                const uint32_t *src = reinterpret_cast<const uint32_t *>(&ty[1 + row][XO + 8 * half]);
                *reinterpret_cast<uint2 *>(rec_y + (size_t)(16 * r + row) * width + 16 * c + 8 * half) = make_uint2(src[0], src[1]);
            }
            if (lane == 0) {
                parts[mb] = ARE4x4;
                segment_id[mb] = 0;
            }
            __syncwarp();
#if defined(INTRA_EXP_TIMELINE)
            if (lane == 0 && r < 4 && c < 128) g_timeline[r][c][13] = gtime();
#endif
            // (the release below is cumulative over the lanes the __syncwarp ordered before it: no separate fence)
            if (lane == 0) st_release_gpu(&done_y[r], c + 1);
#if defined(INTRA_EXP_TIMELINE)
            if (lane == 0 && r < 4 && c < 128) g_timeline[r][c][14] = gtime();
#endif
        }
    } else {
        // ------------------------------------------------ chroma: TM_PRED, eight blocks on eight lanes ------------------------------------------------
        for (int c = 0; c < mbw; ++c) {
            const int mb = r * mbw + c;
            if (r > 0) {
                const int need = min(c + 1, mbw);
                while (ld_acquire_gpu(&done_c[r - 1]) < need) {}
                __syncwarp();
            }
            const int pl = lane >> 4, l = lane & 15;
            uint8_t *rp = pl ? rec_v : rec_u;
            const uint8_t *cp = pl ? cur_v : cur_u;
            if (l < 8) tc[pl][1 + l][0] = c == 0 ? (uint8_t)129 : tc[pl][1 + l][8];
            __syncwarp();
            if (l < 9) {
                uint32_t v;
                if (r == 0) v = 127;
                else if (l == 0) v = c == 0 ? 129u : ld_cg_u8(rp + (size_t)(8 * r - 1) * cw + 8 * c - 1);
                else v = ld_cg_u8(rp + (size_t)(8 * r - 1) * cw + 8 * c + l - 1);
                tc[pl][0][l] = (uint8_t)v;
            }
            __syncwarp();
            if (l < 4) {
                const int br = l >> 1, bc = l & 1, x0 = 4 * bc, y0 = 4 * br;
                const int P = tc[pl][0][0];
                int pred[16], res[16], f[16], out[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    pred[i] = sat8((int)tc[pl][0][1 + x0 + (i & 3)] + (int)tc[pl][1 + y0 + (i >> 2)][0] - P);
                    res[i] = (int)cp[(size_t)(8 * r + y0 + (i >> 2)) * cw + 8 * c + x0 + (i & 3)] - pred[i];
                }
                intra_fdct(res, f);
                intra_quant_recon(f, pred, out, q.uv_dc, q.uv_ac, m_uv_dc, m_uv_ac);
                // (the blocks of a plane only read the border: the tile's interior can be overwritten right away, except
                // column 8, which the next macroblock needs and which block column 1 writes last)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t w = (uint32_t)out[4 * i] | ((uint32_t)out[4 * i + 1] << 8) | ((uint32_t)out[4 * i + 2] << 16) | ((uint32_t)out[4 * i + 3] << 24);
                    *reinterpret_cast<uint32_t *>(rp + (size_t)(8 * r + y0 + i) * cw + 8 * c + x0) = w;
                    if (bc == 1) tc[pl][1 + y0 + i][8] = (uint8_t)out[4 * i + 3];
                }
                store_zigzag(MB + (size_t)mb * 400 + 16 * (16 + 4 * pl + l), f);
            }
            __syncwarp();
            // (cumulative release, as above)
            if (lane == 0) st_release_gpu(&done_c[r], c + 1);
        }
    }
}

}  // namespace vp8

using namespace vp8;

#if defined(INTRA_EXP_TIMELINE)
extern "C" int vp8b200_intra_debug_timeline(unsigned long long *out) {
    return (int)cudaMemcpyFromSymbol(out, g_timeline, sizeof(g_timeline));
}
extern "C" int vp8b200_intra_debug_phases(long long *out) {
    return (int)cudaMemcpyFromSymbol(out, g_phase, sizeof(g_phase));
}
#endif
extern "C" size_t vp8b200_intra_frame_scratch_bytes(int width, int height) {
    return (size_t)(1 + 2 * (height / 16)) * sizeof(int);
}

extern "C" int vp8b200_intra_frame(void *stream, const uint8_t *cur_y, const uint8_t *cur_u, const uint8_t *cur_v,
                                   uint8_t *rec_y, uint8_t *rec_u, uint8_t *rec_v, int16_t *MB, int32_t *modes,
                                   int32_t *MB_parts, int32_t *MB_segment_id, int width, int height, int y_dc_q, int y_ac_q,
                                   int uv_dc_q, int uv_ac_q, void *scratch) {
    if (width < 16 || height < 16 || (width & 15) || (height & 15) || !scratch) return -(int)cudaErrorInvalidValue;
    if (y_dc_q < 1 || y_ac_q < 1 || uv_dc_q < 1 || uv_ac_q < 1) return -(int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    const int mbh = height / 16;
    if (cudaMemsetAsync(scratch, 0, vp8b200_intra_frame_scratch_bytes(width, height), st) != cudaSuccess)
        return -(int)cudaGetLastError();
    const IntraQuants q = {y_dc_q, y_ac_q, uv_dc_q, uv_ac_q};
    k_intra_frame<<<mbh, 64, 0, st>>>(cur_y, cur_u, cur_v, rec_y, rec_u, rec_v, MB, modes, MB_parts, MB_segment_id, width, height, q,
                                      (int *)scratch);
    VP8_LAUNCH_CHECK();
}
