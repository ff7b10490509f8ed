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
#include <cstdlib>
#include <mutex>
#include <cooperative_groups.h>

#include "common.cuh"

namespace vp8 {

__device__ __forceinline__ int c128(int v) { return min(max(v, -128), 127); }

// branch-free on purpose: the lanes of the filtering warp take different decisions per pixel line,
// short-circuit evaluation would serialise them
// (|a - b| is one VABSDIFF through __sad; the six interior terms meet in one maximum before the single compare)
__device__ __forceinline__ int lf_ad(int a, int b) { return (int)__sad(a, b, 0u); }
__device__ __forceinline__ int lf_over(int a, int b, int lim) { return lf_ad(a, b) > lim; }

__device__ __forceinline__ int lf_filter_off(int p3, int p2, int p1, int p0, int q0, int q1, int q2, int q3, int e_lim,
                                             int i_lim) {
    const int inner = max(max(max(lf_ad(p3, p2), lf_ad(p2, p1)), lf_ad(p1, p0)),
                          max(max(lf_ad(q1, q0), lf_ad(q2, q1)), lf_ad(q3, q2)));
    return (inner > i_lim) | ((lf_ad(p0, q0) * 2 + (lf_ad(p1, q1) >> 1)) > e_lim);
}
__device__ __forceinline__ int lf_hev(int p1, int p0, int q0, int q1, int hev_thr) {
    return max(lf_ad(p1, p0), lf_ad(q1, q0)) > hev_thr;
}

// macroblock edge (filter_mb_edge8, src/CPU_kernels.cl:829-883), one lane
__device__ __forceinline__ void filter_mb_edge(int p3, int &p2, int &p1, int &p0, int &q0, int &q1, int &q2, int q3,
                                               int mb_lim, int int_lim, int hev_thr) {
    const int off = lf_filter_off(p3, p2, p1, p0, q0, q1, q2, q3, mb_lim, int_lim);
    const int hev = lf_hev(p1, p0, q0, q1, hev_thr);
    int w = c128(c128(p1 - q1) + 3 * (q0 - p0));
    w = off ? 0 : w;
    // (w is in -128..127 from here on: w + 3 and w + 4 can only leave the range at the top, and
    // (27 w + 63) >> 7 and its siblings stay within -27..27 -- the reference's clamps there never act)
    int a = hev ? w : 0;
    const int b = min(a + 3, 127) >> 3;
    a = min(a + 4, 127) >> 3;
    q0 -= a;
    p0 += b;
    w = hev ? 0 : w;
    a = (27 * w + 63) >> 7;
    q0 -= a;
    p0 += a;
    a = (18 * w + 63) >> 7;
    q1 -= a;
    p1 += a;
    a = (9 * w + 63) >> 7;
    q2 -= a;
    p2 += a;
}

// inner (sub-block) edge (filter_b_edge8, src/CPU_kernels.cl:885-926), one lane
__device__ __forceinline__ void filter_b_edge(int p3, int p2, int &p1, int &p0, int &q0, int &q1, int q2, int q3,
                                              int b_lim, int int_lim, int hev_thr) {
    const int off = lf_filter_off(p3, p2, p1, p0, q0, q1, q2, q3, b_lim, int_lim);
    const int hev = lf_hev(p1, p0, q0, q1, hev_thr);
    int a = hev ? c128(p1 - q1) : 0;
    a = c128(a + 3 * (q0 - p0));
    a = off ? 0 : a;
    const int b = min(a + 3, 127) >> 3;  // (a is in -128..127: only the upper clamp can act)
    a = min(a + 4, 127) >> 3;
    q0 -= a;
    p0 += b;
    a = (a + 1) >> 1;
    a = hev ? 0 : a;
    q1 -= a;
    p1 += a;
}

// Inner edges 8 and 12 of a line depend on the edge before them only through two terms of the filter mask (their
// p3, p2 are the q0, q1 the edge before has just rewritten; p1, p0 and the q side are still untouched).  So the
// filter is evaluated from the untouched pixels while the edges before it are still in flight (the lanes of a
// filter warp are latency-bound: one dependent chain per lane) and only the mask is resolved afterwards: a masked-off
// edge leaves all four pixels as they were (a = 0 gives b = 0 and (a + 1) >> 1 = 0).
struct SpecEdge {
    int p1, p0, q0, q1;  // what the four pixels become if the edge is filtered
    int off;             // mask terms that do not involve p3, p2
};
__device__ __forceinline__ SpecEdge spec_b_edge(int p1, int p0, int q0, int q1, int q2, int q3, int b_lim, int int_lim,
                                                int hev_thr) {
    SpecEdge s;
    s.off = lf_over(p1, p0, int_lim) | lf_over(q1, q0, int_lim) | lf_over(q2, q1, int_lim) | lf_over(q3, q2, int_lim) |
            ((lf_ad(p0, q0) * 2 + (lf_ad(p1, q1) >> 1)) > b_lim);
    const int hev = lf_hev(p1, p0, q0, q1, hev_thr);
    int a = hev ? c128(p1 - q1) : 0;
    a = c128(a + 3 * (q0 - p0));
    const int b = c128(a + 3) >> 3;
    a = c128(a + 4) >> 3;
    s.q0 = q0 - a;
    s.p0 = p0 + b;
    a = (a + 1) >> 1;
    a = hev ? 0 : a;
    s.q1 = q1 - a;
    s.p1 = p1 + a;
    return s;
}
__device__ __forceinline__ void resolve_b_edge(const SpecEdge &s, int p3, int p2, int &p1, int &p0, int &q0, int &q1,
                                               int int_lim) {
    const int off = s.off | lf_over(p3, p2, int_lim) | lf_over(p2, p1, int_lim);
    p1 = off ? p1 : s.p1;
    p0 = off ? p0 : s.p0;
    q0 = off ? q0 : s.q0;
    q1 = off ? q1 : s.q1;
}

// The edges of one line of N pixels of a macroblock, v[0..3] = the four pixels before it, v[4..4+N) the line, all as
// pixel - 128; results stay UNCLAMPED in v (within one macroblock the reference hands the unclamped q0..q3 of an edge
// on as p3..p0 of the next, Q7; memory gets the clamped values).  Same results as the edge-after-edge walk
// (filter_mb_edge, then filter_b_edge at 4, 8, 12), shorter dependent chain: see SpecEdge.
template <int N>
__device__ __forceinline__ void filter_line_spec(int (&v)[N + 4], bool mb_edge, bool inner, int mb_lim, int b_lim,
                                                 int int_lim, int hev_thr) {
    SpecEdge s8, s12;
    if (N == 16 && inner) {
        s8 = spec_b_edge(v[10], v[11], v[12], v[13], v[14], v[15], b_lim, int_lim, hev_thr);
        s12 = spec_b_edge(v[14], v[15], v[16], v[17], v[18], v[19], b_lim, int_lim, hev_thr);
    }
    if (mb_edge) filter_mb_edge(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], mb_lim, int_lim, hev_thr);
    if (inner) {
        filter_b_edge(v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], b_lim, int_lim, hev_thr);
        if (N == 16) {
            resolve_b_edge(s8, v[8], v[9], v[10], v[11], v[12], v[13], int_lim);
            resolve_b_edge(s12, v[12], v[13], v[14], v[15], v[16], v[17], int_lim);
        }
    }
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

// four (pixel-128) values -> four saturated pixels: sat_s8 then flip the sign bits
// (cvt.pack.sat.s8.s32: d = c<<16 | sat(a)<<8 | sat(b))
__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) {
    uint32_t hi, r;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(hi));
    return r ^ 0x80808080u;
}
// byte k of a word of pixels -> pixel-128 (flip the sign bit, sign-extend)
// (one PRMT: selector nibble k takes byte k, nibbles 8|k replicate its sign bit)
__device__ __forceinline__ int unpack1(uint32_t w_flipped, int k) {
    int v;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(v) : "r"(w_flipped), "r"(0x8880 + 0x1111 * k));
    return v;
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

// progress flags between the warps of a CTA (shared memory): release store / acquire load at CTA scope -- the
// sequentially consistent fence __threadfence_block() compiles to (MEMBAR.SC.CTA) is not needed for a hand-off
__device__ __forceinline__ void flag_release(volatile int *p, int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(const_cast<int *>(p))), "r"(v) : "memory");
}
__device__ __forceinline__ int flag_acquire(volatile int *p) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(const_cast<int *>(p))) : "memory");
    return v;
}

// Hand-off between two CTAs of a cluster through distributed shared memory.  A cluster-scope release/acquire pair
// compiles to MEMBAR.ALL.GPU / CCTL.IVALL on sm_100a -- far too heavy for a once-per-macroblock signal on the
// critical warp -- so the lines travel as st.async stores that count bytes on an mbarrier of the receiving CTA:
// the receiver waits for the barrier's phase (a CTA-local SYNCS.PHASECHK), no fence on either side.
__device__ __forceinline__ uint32_t smem_addr(const volatile void *p) { return (uint32_t)__cvta_generic_to_shared(const_cast<void *>(p)); }
__device__ __forceinline__ uint32_t cluster_addr(uint32_t local, int rank) {  // the same location in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_u32(uint32_t remote, uint32_t v, uint32_t remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(remote), "r"(v), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t mbar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t mbar, int parity) {
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ int ld_acquire_cluster(uint32_t remote) {
    int v;
    asm volatile("ld.acquire.cluster.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(remote) : "memory");
    return v;
}

struct LFPlanes {
    uint8_t *ptr[3];
};

// warps: 3 per plane slot (filter, receiver, sender); up to 3 plane slots per CTA.  The helper warps spin on flags
// and on the mailbox; a warp that spins on the scheduler of a filter warp takes issue slots from the one warp the
// whole wavefront waits for (28 % + 20 % of the kernel's executed instructions were those two loops).  So:
// twelve warps, the luma filter warp alone on its scheduler (warps 4 and 8 idle), and the helpers back off with
// nanosleep between polls.
#ifndef VP8_LF_SPREAD
#define VP8_LF_SPREAD 1
#endif
#ifndef VP8_LF_BACKOFF_NS
#define VP8_LF_BACKOFF_NS 40
#endif
constexpr int LF_THREADS = VP8_LF_SPREAD ? 384 : 288;
__device__ __forceinline__ void lf_backoff() {
#if VP8_LF_BACKOFF_NS > 0
    __nanosleep(VP8_LF_BACKOFF_NS);
#endif
}
constexpr int LF_TOPQ = 8;  // depth of the top-line ring

// shared-memory layout of one plane slot
struct LFSlot {
    uint8_t *strip;        // N rows, stride S = width + 4 (word stride odd: pass 1 is bank-conflict free)
    int4 *lims;            // per macroblock {mb_lim, b_lim, int_lim, hev_thr | inner<<16}
    uint32_t *topq;        // [LF_TOPQ][4 lines][N/4 words]: the four pixel lines above each macroblock
    volatile int *flags;   // h_done, v_done, top_ready
    uint64_t *mbar;        // [LF_TOPQ] one per ring slot: bytes of the lines above, sent by the CTA above (cluster hand-off)
};
__host__ __device__ inline size_t lf_slot_bytes(int n, int width) {
    return (((size_t)n * (width + 4) + 15) & ~(size_t)15) + (size_t)(width / n) * 16 + (size_t)LF_TOPQ * n * 4 + 16 + (size_t)LF_TOPQ * 8;
}
__device__ inline LFSlot lf_slot(unsigned char *base, int n, int width) {
    LFSlot s;
    s.strip = base;
    s.lims = reinterpret_cast<int4 *>(base + (((size_t)n * (width + 4) + 15) & ~(size_t)15));
    s.topq = reinterpret_cast<uint32_t *>(s.lims + width / n);
    s.flags = reinterpret_cast<volatile int *>(s.topq + LF_TOPQ * n);
    s.mbar = reinterpret_cast<uint64_t *>(const_cast<int *>(s.flags) + 4);
    return s;
}

// Rows hand their bottom pixel lines to the row below through a "tagged mailbox" in global memory:
// 32 words per macroblock, each carrying two pixels and a 16-bit launch tag.  A word is written
// atomically, so a reader that sees the current tag in a word also sees that word's pixels: one
// round trip, no fence, no separate flag.  Tags differ from launch to launch (the host wraps them
// safely), every slot that is read in a launch is written in the same launch.
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// VP8_LF_SPECULATIVE selects filter_line_spec (measured on B200, 1080p, three planes: 242 us against 227 us for the
// edge-after-edge walk: the filter warp is bound by the number of instructions it has to issue on its own, about 700
// per macroblock at an IPC of 0.3, not by the depth of the dependent chain, and speculation adds instructions)
#if defined(VP8_LF_SPECULATIVE)
#define LF_FILTER_LINE filter_line_spec
#else
#define LF_FILTER_LINE filter_line
#endif
template <int N>
__device__ void lf_plane_roles(uint8_t *__restrict__ frame, int width, int r, int ncols, bool row_below_exists,
                               int stop_below_cols, LFSlot sl, uint32_t *mail_row, const uint32_t *mail_above,
                               unsigned tag, int role, int lane, bool dsmem_up, int rank_below) {
    // dsmem_up: the row above runs in the previous CTA of this cluster and sends the lines above every macroblock
    // straight into this CTA's ring (topq), counted on the ring slot's mbarrier; rank_below >= 0: this row does the
    // same for that CTA of the cluster.  Rows at a cluster border use the global mailbox.
    const bool dsmem_down = rank_below >= 0;
    const int S = width + 4;
    const int y0 = r * N;
    uint8_t *strip = sl.strip;
    volatile int *flags = sl.flags;
    uint32_t *topq = sl.topq;
    constexpr int TOPQ = LF_TOPQ;

    if (role == 0) {
        // ---- filter warp: pass 1 (vertical edges) + pass 2 (horizontal edges), shared memory only ----
        for (int c = 0; c < ncols; ++c) {
            const int4 lm = sl.lims[c];
            const int mb_lim = lm.x, b_lim = lm.y, int_lim = lm.z, hev_thr = lm.w & 0xffff;
            const bool inner = (lm.w >> 16) != 0;
            const int x0 = c * N;
            // Both passes walk the N/4 four-pixel groups of a line with a sliding window p3..p0 | q0..q3
            // kept in registers: the unclamped q's of one edge are the p's of the next (Q7).
            if (lane < N) {  // pass 1: one lane per pixel row, the whole line in registers
                uint32_t *row = reinterpret_cast<uint32_t *>(strip + lane * S + 4 + x0);
                int v[N + 4];
                uint32_t w = c > 0 ? row[-1] ^ 0x80808080u : 0u;
                v[0] = unpack1(w, 0); v[1] = unpack1(w, 1); v[2] = unpack1(w, 2); v[3] = unpack1(w, 3);
#pragma unroll
                for (int e = 0; e < N / 4; ++e) {
                    w = row[e] ^ 0x80808080u;
                    v[4 + 4 * e] = unpack1(w, 0); v[5 + 4 * e] = unpack1(w, 1); v[6 + 4 * e] = unpack1(w, 2); v[7 + 4 * e] = unpack1(w, 3);
                }
                LF_FILTER_LINE<N>(v, c > 0, inner, mb_lim, b_lim, int_lim, hev_thr);
                if (c > 0) row[-1] = pack4(v[0], v[1], v[2], v[3]);
#pragma unroll
                for (int e = 0; e < N / 4; ++e) row[e] = pack4(v[4 + 4 * e], v[5 + 4 * e], v[6 + 4 * e], v[7 + 4 * e]);
            }
            __syncwarp();
            if (lane == 0) flag_release(&flags[0], c + 1);  // h_done: macroblock c-1 is final now
            if (r > 0) {  // the lines above macroblock c are in the ring
                if (dsmem_up) {
                    const uint32_t mb_slot = smem_addr(sl.mbar + (c % TOPQ));
                    if (lane == 0) mbar_expect(mb_slot, 4 * N);
                    while (!mbar_try_wait(mb_slot, (c / TOPQ) & 1)) {}
                } else {
                    while (flag_acquire(&flags[2]) < c + 1) {}
                }
            }
            if (lane < N) {  // pass 2: one lane per pixel column, the whole column in registers
                uint8_t *scol = strip + 4 + x0 + lane;
                int v[N + 4];
#pragma unroll
                for (int j = 0; j < N; ++j) v[4 + j] = (int)scol[j * S] - 128;
                v[0] = v[1] = v[2] = v[3] = 0;
                if (r > 0) {
                    const uint8_t *tq = reinterpret_cast<const uint8_t *>(topq + (c % TOPQ) * N) + lane;
                    v[0] = (int)tq[0] - 128; v[1] = (int)tq[N] - 128; v[2] = (int)tq[2 * N] - 128; v[3] = (int)tq[3 * N] - 128;
                }
                LF_FILTER_LINE<N>(v, r > 0, inner, mb_lim, b_lim, int_lim, hev_thr);
                if (r > 0) {
                    // the three lines above now belong to this row: straight to the frame
                    uint8_t *gcol = frame + (size_t)y0 * width + x0 + lane;
                    const uint32_t t = pack4(v[0], v[1], v[2], v[3]);
                    gcol[-3 * (ptrdiff_t)width] = (uint8_t)(t >> 8);
                    gcol[-2 * (ptrdiff_t)width] = (uint8_t)(t >> 16);
                    gcol[-1 * (ptrdiff_t)width] = (uint8_t)(t >> 24);
                }
#pragma unroll
                for (int e = 0; e < N / 4; ++e) {
                    uint8_t *g = scol + 4 * e * S;
                    const uint32_t t = pack4(v[4 + 4 * e], v[5 + 4 * e], v[6 + 4 * e], v[7 + 4 * e]);
                    g[0] = (uint8_t)t; g[S] = (uint8_t)(t >> 8); g[2 * S] = (uint8_t)(t >> 16); g[3 * S] = (uint8_t)(t >> 24);
                }
            }
            __syncwarp();
            // v_done.  (With dsmem_up the sender of the row above reads it to know a ring slot is free again: this
            // warp's loads of the slot have long returned -- pass 2 consumed them -- when the flag is written.)
            if (lane == 0) flag_release(&flags[1], c + 1);
        }
    } else if (role == 1) {
        // ---- receiver warp: the four lines above each macroblock, from the mailbox of the row above ----
        if (r > 0 && !dsmem_up) {
            for (int c = 0; c < ncols; ++c) {
                while (c - flag_acquire(&flags[1]) >= TOPQ) lf_backoff();  // ring slot free again
                const uint32_t *m = mail_above + (size_t)c * 32 + lane;
                uint32_t w;
                for (;;) {
                    w = ld_volatile_u32(m);
                    if (__all_sync(0xffffffffu, (w >> 16) == tag)) break;
                    lf_backoff();
                }
                // lane = 2*k + h holds pixels (2h, 2h+1) of word k of the 4 x N/4 word block (N=16),
                // see the sender; reassemble whole words
                const uint32_t other = __shfl_xor_sync(0xffffffffu, w, 1);
                if (N == 16) {
                    if ((lane & 1) == 0) topq[(c % TOPQ) * N + (lane >> 1)] = (w & 0xffffu) | (other << 16);
                } else {
                    if ((lane & 1) == 0 && lane < 16) topq[(c % TOPQ) * N + (lane >> 1)] = (w & 0xffffu) | (other << 16);
                }
                __syncwarp();
                if (lane == 0) flag_release(&flags[2], c + 1);  // top_ready
            }
        }
    } else {
        // ---- sender warp: bottom lines of every finalised macroblock ----
        int below_v_done = 0;  // last progress seen of the row below (ring flow control of the cluster hand-off)
        const uint32_t topq_below = dsmem_down ? cluster_addr(smem_addr(topq), rank_below) : 0u;
        const uint32_t mbar_below = dsmem_down ? cluster_addr(smem_addr(sl.mbar), rank_below) : 0u;
        const uint32_t flags_below_v_done = dsmem_down ? cluster_addr(smem_addr(flags + 1), rank_below) : 0u;
        for (int c = 0; c < ncols; ++c) {
            // macroblock c is final once pass 1 of macroblock c+1 ran, the last one after its own pass 2
            if (c + 1 < ncols) {
                while (flag_acquire(&flags[0]) < c + 2) lf_backoff();
            } else {
                while (flag_acquire(&flags[1]) < ncols) lf_backoff();
            }
            const bool consumer = row_below_exists && c < stop_below_cols;
            if (consumer && dsmem_down) {
                // straight into the ring of the CTA below: 4 lines x N/4 words, word k from lane k
                while (c - below_v_done >= LF_TOPQ) {
                    below_v_done = ld_acquire_cluster(flags_below_v_done);
                    if (c - below_v_done >= LF_TOPQ) lf_backoff();
                }
                if (lane < N) {
                    const int line = N - 4 + lane / (N / 4), wd = lane % (N / 4);
                    st_async_u32(topq_below + 4u * (uint32_t)((c % LF_TOPQ) * N + lane),
                                 *reinterpret_cast<const uint32_t *>(strip + line * S + 4 + c * N + 4 * wd),
                                 mbar_below + 8u * (uint32_t)(c % LF_TOPQ));
                }
            } else if (consumer) {
                // 4 lines x N/4 words, split in halves: word k -> lanes 2k, 2k+1 (N=8 uses lanes 0..15)
                const int k = lane >> 1, h = lane & 1;
                if (N == 16 || lane < 16) {
                    const int line = N - 4 + k / (N / 4), wd = k % (N / 4);
                    const uint32_t w = *reinterpret_cast<const uint32_t *>(strip + line * S + 4 + c * N + 4 * wd);
                    mail_row[(size_t)c * 32 + lane] = ((w >> (16 * h)) & 0xffffu) | (tag << 16);
                } else {
                    mail_row[(size_t)c * 32 + lane] = tag << 16;  // unused half of a chroma slot: tag only
                }
            } else if (lane < 3 * (N / 4)) {
                // nobody below will touch these lines: lines N-3..N-1 go to the frame from here
                const int line = N - 3 + lane / (N / 4), wd = lane % (N / 4);
                const uint32_t w = *reinterpret_cast<const uint32_t *>(strip + line * S + 4 + c * N + 4 * wd);
                *reinterpret_cast<uint32_t *>(frame + (size_t)(y0 + line) * width + c * N + 4 * wd) = w;
            }
        }
    }
}

// ctrl[0] = ticket counter.  One CTA = macroblock row r of every plane in the launch.
__global__ void __launch_bounds__(LF_THREADS)
k_loop_filter(LFPlanes planes, int first_plane, int num_planes, const int *__restrict__ seg,
              const int *__restrict__ mb_mask, const vp8b200_segment_data *__restrict__ SD, int luma_width,
              int luma_height, int *ctrl, uint32_t *mail, unsigned tag) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_ticket, s_stop;
    const int mbw = luma_width / 16, mbh = luma_height / 16, mb_count = mbw * mbh;
    const int tid = threadIdx.x;
    // A cluster of C CTAs takes C consecutive rows (one ticket per cluster): inside the cluster a row hands its bottom
    // lines to the next one through distributed shared memory instead of the global mailbox.
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
    // slot ps starts after the slots before it (no pointer table: the accesses stay LDS/STS)
    auto slot_of = [&](int ps) {
        size_t off = 0;
        for (int i = 0; i < ps; ++i) off += lf_slot_bytes(first_plane + i == 0 ? 16 : 8, first_plane + i == 0 ? luma_width : luma_width / 2);
        return lf_slot(smem + off, first_plane + ps == 0 ? 16 : 8, first_plane + ps == 0 ? luma_width : luma_width / 2);
    };
    if (tid == 0) {
        if (crank == 0) s_ticket = atomicAdd(&ctrl[0], 1);  // clusters start in ticket order: a row's predecessor is always running
        s_stop = mb_count;
    }
    // progress flags: cleared before the cluster barrier, because the CTA above may write top_ready of this one
    // as soon as it has passed that barrier
    for (int ps = 0; ps < num_planes; ++ps) {
        const LFSlot slot = slot_of(ps);
        if (tid < 3) slot.flags[tid] = 0;
        if (csize > 1 && tid >= 32 && tid < 32 + LF_TOPQ)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(slot.mbar + (tid - 32))) : "memory");
    }
    if (csize > 1) {
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        cluster.sync();
        if (tid == 0 && crank != 0) s_ticket = *cluster.map_shared_rank(&s_ticket, 0);
    }
    __syncthreads();
    // "if (SD[i].loop_filter_level == 0) return;" ends the WHOLE plane at the first such macroblock
    // in raster order (Q6): find that index
    unsigned zero_mask = 0;  // segments whose filter level is 0
#pragma unroll
    for (int sgm = 0; sgm < 4; ++sgm) zero_mask |= (SD[sgm].loop_filter_level == 0) << sgm;
    if (zero_mask) {  // rare: scan the segment map for the first such macroblock
        int first = mb_count;
        for (int mb = tid; mb < mb_count && first == mb_count; mb += LF_THREADS)
            if ((zero_mask >> seg[mb]) & 1) first = mb;
        if (first < mb_count) atomicMin(&s_stop, first);
    }
    __syncthreads();
    const int r = s_ticket * csize + crank, stop = s_stop;
    const int ncols = r < mbh ? max(0, min(mbw, stop - r * mbw)) : 0;  // macroblocks of this row before the Q6 stop
    if (ncols == 0) {  // nothing to filter: nobody writes into this CTA (its predecessor sees below_cols == 0)
        if (csize > 1) cluster.sync();
        return;
    }
    const int below_cols = max(0, min(mbw, stop - (r + 1) * mbw));
    const bool row_below = r + 1 < mbh;
    const bool dsmem_up = csize > 1 && crank > 0 && r > 0;
    const bool dsmem_down = csize > 1 && crank + 1 < csize && row_below;

    // ---- stage the strips and the per-macroblock limits of every plane ----
    for (int ps = 0; ps < num_planes; ++ps) {
        const int plane = first_plane + ps;
        const int n = plane == 0 ? 16 : 8;
        const int width = plane == 0 ? luma_width : luma_width / 2;
        const LFSlot slot = slot_of(ps);
        const uint8_t *frame = planes.ptr[plane];
        const int S = width + 4, chunks = width / 8, y0 = r * n;
        // 8-byte chunks: plane widths are multiples of 8 (chroma of a 16-aligned luma), rows 8-byte aligned
        for (int i = tid; i < n * chunks; i += LF_THREADS) {
            const int row = i / chunks, xc = i % chunks;
            const uint2 v = *reinterpret_cast<const uint2 *>(frame + (size_t)(y0 + row) * width + 8 * xc);
            uint32_t *d = reinterpret_cast<uint32_t *>(slot.strip + row * S + 4 + 8 * xc);
            d[0] = v.x; d[1] = v.y;
        }
        for (int c = tid; c < ncols; c += LF_THREADS) {
            const int mb = r * mbw + c;
            const vp8b200_segment_data *sd = SD + seg[mb];
            slot.lims[c] = make_int4((short)sd->mbedge_limit, (short)sd->sub_bedge_limit, (short)sd->interior_limit,
                                          ((short)sd->hev_threshold & 0xffff) | (mb_mask[mb] != 0 ? 0x10000 : 0));
        }
    }
    __syncthreads();

    // ---- the serial walk: warp w serves plane slot w/3 in role w%3 ----
    {
        const int warp = tid >> 5, lane = tid & 31;
#if VP8_LF_SPREAD
        // warp -> scheduler warp % 4:  0: luma filter alone | 1: U filter, U receiver, U sender | 2: the same for V |
        // 3: luma receiver, luma sender
        const int ps = (warp & 3) == 0 ? (warp == 0 ? 0 : 99) : ((warp & 3) == 3 ? (warp == 11 ? 99 : 0) : (warp & 3));
        const int role = (warp & 3) == 3 ? 1 + (warp >> 2) : (warp >> 2);
#else
        const int ps = warp / 3, role = warp % 3;
#endif
        if (ps < num_planes) {
            const int plane = first_plane + ps;
            uint32_t *mail_plane = mail + (size_t)ps * mbh * mbw * 32;
            uint32_t *mail_row = mail_plane + (size_t)r * mbw * 32;
            const uint32_t *mail_above = mail_plane + (size_t)(r - 1) * mbw * 32;  // only read when r > 0
            const LFSlot slot = slot_of(ps);
            const int rank_below = dsmem_down ? crank + 1 : -1;
            if (plane == 0)
                lf_plane_roles<16>(planes.ptr[0], luma_width, r, ncols, row_below, below_cols, slot, mail_row,
                                   mail_above, tag, role, lane, dsmem_up, rank_below);
            else
                lf_plane_roles<8>(planes.ptr[plane], luma_width / 2, r, ncols, row_below, below_cols, slot, mail_row,
                                  mail_above, tag, role, lane, dsmem_up, rank_below);
        }
    }
    __syncthreads();

    // ---- write back lines 0..N-4 of the filtered range.  Lines N-3..N-1 reach the frame through the
    // row below (which rewrites them) or through the sender warp when there is no such row ----
    for (int ps = 0; ps < num_planes; ++ps) {
        const int plane = first_plane + ps;
        const int n = plane == 0 ? 16 : 8;
        const int width = plane == 0 ? luma_width : luma_width / 2;
        uint8_t *frame = planes.ptr[plane];
        const int S = width + 4, y0 = r * n;
        const LFSlot slot = slot_of(ps);
        const int out_chunks = ncols * n / 8;
        for (int i = tid; i < (n - 3) * out_chunks; i += LF_THREADS) {
            const int row = i / out_chunks, xc = i % out_chunks;
            const uint32_t *s = reinterpret_cast<const uint32_t *>(slot.strip + row * S + 4 + 8 * xc);
            *reinterpret_cast<uint2 *>(frame + (size_t)(y0 + row) * width + 8 * xc) = make_uint2(s[0], s[1]);
        }
    }
    if (csize > 1) cluster.sync();  // no CTA of a cluster leaves while a neighbour may still touch its shared memory
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

// ticket counter and mailbox of the launches of one stream (launches on the same stream are ordered, launches
// on different streams may overlap and must not share them)
struct LFContext {
    bool used;
    cudaStream_t stream;
    int *ctrl;
    uint32_t *mail;
    size_t mail_words;
    unsigned tag;
    size_t smem_configured;  // dynamic shared memory the kernel has been allowed on this stream's device
};
static LFContext g_lf_ctx[256];
static std::mutex g_lf_mutex;

static LFContext *lf_context(cudaStream_t st) {
    std::lock_guard<std::mutex> lock(g_lf_mutex);
    LFContext *free_slot = nullptr;
    for (LFContext &c : g_lf_ctx) {
        if (c.used && c.stream == st) return &c;
        if (!c.used && !free_slot) free_slot = &c;
    }
    if (!free_slot) return nullptr;
    *free_slot = LFContext{false, st, nullptr, nullptr, 0, 0, 0};
    if (cudaMalloc((void **)&free_slot->ctrl, sizeof(int) * 4) != cudaSuccess) return nullptr;
    free_slot->used = true;
    return free_slot;
}

// the owner of a stream (vp8b200_engine_destroy) gives the stream's ticket counter and mailbox back; a stream handle
// the driver hands out again later starts from a clean state
extern "C" void vp8b200_loop_filter_release(void *stream) {
    std::lock_guard<std::mutex> lock(g_lf_mutex);
    for (LFContext &c : g_lf_ctx)
        if (c.used && c.stream == (cudaStream_t)stream) {
            cudaStreamSynchronize(c.stream);
            cudaFree(c.ctrl);
            if (c.mail) cudaFree(c.mail);
            c = LFContext{false, nullptr, nullptr, nullptr, 0, 0, 0};
        }
}

static int launch_loop_filter(void *stream, LFPlanes p, int first_plane, int num_planes, const int32_t *seg,
                              const int32_t *mb_mask, const vp8b200_segment_data *SD, int luma_w, int luma_h) {
    cudaStream_t st = (cudaStream_t)stream;
    const int mbw = luma_w / 16, mbh = luma_h / 16;
    LFContext *c = lf_context(st);
    if (!c) return -(int)cudaErrorMemoryAllocation;
    const size_t need = (size_t)num_planes * mbh * mbw * 32;
    if (need > c->mail_words) {
        cudaStreamSynchronize(st);
        if (c->mail) cudaFree(c->mail);
        c->mail = nullptr;
        if (cudaMalloc((void **)&c->mail, need * 4) != cudaSuccess) return -(int)cudaErrorMemoryAllocation;
        c->mail_words = need;
        c->tag = 0;
    }
    // 16-bit launch tags; when they wrap (or the mailbox is new) clear it so that no stale tag can match
    c->tag = (c->tag + 1) & 0xffffu;
    if (c->tag == 0 || c->tag == 1) {
        cudaMemsetAsync(c->mail, 0, c->mail_words * 4, st);
        c->tag = 1;
    }
    cudaMemsetAsync(c->ctrl, 0, sizeof(int) * 4, st);
    size_t smem = 0;
    for (int ps = 0; ps < num_planes; ++ps) {
        const int plane = first_plane + ps;
        smem += lf_slot_bytes(plane == 0 ? 16 : 8, plane == 0 ? luma_w : luma_w / 2);
    }
    if (smem > c->smem_configured) {  // (the attribute is per device: remembered per stream, not per process)
        if (cudaFuncSetAttribute(k_loop_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return -(int)cudaGetLastError();
        c->smem_configured = smem;
    }
    // clusters of LF_CLUSTER consecutive rows (VP8B200_LF_CLUSTER=1 switches the distributed-shared-memory hand-off off)
    static const int cluster_rows = [] {
        const char *e = getenv("VP8B200_LF_CLUSTER");
        const int v = e ? atoi(e) : 8;
        return v < 1 ? 1 : (v > 8 ? 8 : v);
    }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((mbh + cluster_rows - 1) / cluster_rows * cluster_rows));
    cfg.blockDim = dim3(LF_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster_rows;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int *ctrl = c->ctrl;
    uint32_t *mail = c->mail;
    unsigned tag = c->tag;
    if (cudaLaunchKernelEx(&cfg, k_loop_filter, p, first_plane, num_planes, (const int *)seg, (const int *)mb_mask, SD, luma_w,
                           luma_h, ctrl, mail, tag) != cudaSuccess)
        return -(int)cudaGetLastError();
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
