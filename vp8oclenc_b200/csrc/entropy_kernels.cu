// Token statistics and token streams of the coefficient partitions on the GPU.
//
// The reference codes the coefficients in three "CPU program" kernels (src/CPU_kernels.cl:347-778):
// count_probs (token statistics per partition + the neighbour contexts), num_div_denom and
// encode_coefficients (the boolean coder).  Only the boolean coder is inherently serial.  Everything
// before it -- which decisions are coded, in which order, against which probability slot -- is
// computed here for all macroblocks in parallel:
//
//   k_entropy_scan<false>  one warp per macroblock, one lane per 4x4 block: neighbour contexts
//                          (third_context), the per-partition statistics num/den [P][4][8][3][11]
//                          (shared-memory histograms, flushed with atomics), the reference's
//                          count-past-end-of-block habit as a histogram of tail starts, and the
//                          number of coded decisions of every macroblock;
//   k_entropy_offsets      exclusive scan of those numbers in partition coding order;
//   k_entropy_finish       partition bases, tail expansion into the statistics;
//   k_entropy_scan<true>   the same walk again, now writing every decision as a 16-bit entry
//                          (bit 15 = value, bits 0-10 = probability slot, or 1056 + p for a
//                          literal with fixed probability p) at its place in the partition stream.
//
// The host then runs only RFC 6386's bool coder over the streams (entropy_host.cpp,
// encode_token_streams).  Semantics are those of entropy_host.cpp, which is pinned against the
// reference's kernels; tests compare the two paths bit for bit.
#include "common.cuh"

namespace vp8 {

constexpr int ENT_WARPS = 8;  // macroblocks per CTA (all of one macroblock row -> one partition)

__device__ __forceinline__ int ent_band(int i) {
    // {0,1,2,3,6,4,5,6,6,6,6,6,6,6,6,7} packed in nibbles
    return (int)((0x7666666665463210ULL >> (4 * i)) & 15);
}
__device__ __forceinline__ int ent_ctx_index(int type, int band, int ctx, int slot) {
    return (((type << 3) + band) * 3 + ctx) * 11 + slot;
}

static __constant__ const unsigned char c_cat_prob[6][11] = {
    {159}, {165, 145}, {173, 148, 140}, {176, 155, 140, 135}, {180, 157, 141, 134, 130},
    {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129}};
static __constant__ const unsigned char c_cat_base[6] = {5, 7, 11, 19, 35, 67};
static __constant__ const unsigned char c_cat_bits[6] = {1, 2, 3, 4, 5, 11};

// what happens to a coded decision: statistics (shared-memory histograms), counting, or the stream
struct StatSink {
    unsigned *s_num, *s_den;
    int n;
    __device__ __forceinline__ void decision(int idx, int bit) {
        atomicAdd(&s_den[idx], 1u);
        if (!bit) atomicAdd(&s_num[idx], 1u);  // zeros are counted
        ++n;
    }
    __device__ __forceinline__ void literal(int, int) { ++n; }
};
struct CountSink {
    int n;
    __device__ __forceinline__ void decision(int, int) { ++n; }
    __device__ __forceinline__ void literal(int, int) { ++n; }
};
struct WriteSink {
    uint16_t *out;
    __device__ __forceinline__ void decision(int idx, int bit) { *out++ = (uint16_t)(idx | (bit << 15)); }
    __device__ __forceinline__ void literal(int prob, int bit) { *out++ = (uint16_t)((1056 + prob) | (bit << 15)); }
};

// one block, RFC 6386 13.2 token tree written out; returns the position of the end-of-block token (16: none)
template <class Sink>
__device__ int ent_walk_block(const int16_t *coef, int type, int ctx, unsigned mask, Sink &e) {
    const int first = (type == 0) ? 1 : 0;
    if (first) mask &= ~1u;
    const int last = mask ? 31 - __clz(mask) : -1;
    bool prev_zero = false;
    int i = first;
    for (; i <= last; ++i) {
        const int band = ent_band(i);
        const int v = coef[i];
        const int base = ent_ctx_index(type, band, ctx, 0);
        if (v == 0) {
            if (!prev_zero) e.decision(base, 1);
            e.decision(base + 1, 0);
            prev_zero = true;
            ctx = 0;
            continue;
        }
        const int mag = abs(v);
        if (!prev_zero) e.decision(base, 1);
        e.decision(base + 1, 1);
        if (mag == 1) {
            e.decision(base + 2, 0);
            ctx = 1;
        } else {
            e.decision(base + 2, 1);
            if (mag <= 4) {
                e.decision(base + 3, 0);
                if (mag == 2) {
                    e.decision(base + 4, 0);
                } else {
                    e.decision(base + 4, 1);
                    e.decision(base + 5, mag == 4);
                }
            } else {
                e.decision(base + 3, 1);
                int c;
                if (mag <= 10) {
                    e.decision(base + 6, 0);
                    e.decision(base + 7, mag > 6);
                    c = mag > 6;
                } else {
                    e.decision(base + 6, 1);
                    if (mag <= 34) {
                        e.decision(base + 8, 0);
                        e.decision(base + 9, mag > 18);
                        c = 2 + (mag > 18);
                    } else {
                        e.decision(base + 8, 1);
                        e.decision(base + 10, mag > 66);
                        c = 4 + (mag > 66);
                    }
                }
                const int extra = mag - c_cat_base[c], nb = c_cat_bits[c];
                for (int b = 0; b < nb; ++b) e.literal(c_cat_prob[c][b], (extra >> (nb - 1 - b)) & 1);
            }
            ctx = 2;
        }
        e.literal(128, v < 0);  // sign
        prev_zero = false;
    }
    if (i < 16) e.decision(ent_ctx_index(type, ent_band(i), ctx, 0), 0);  // end of block
    return i;
}

// "is anything non-zero in coefficients first..15 of this block" straight from global memory
__device__ __forceinline__ bool ent_block_nonzero(const int16_t *blk, int first) {
    const uint4 a = *reinterpret_cast<const uint4 *>(blk), b = *reinterpret_cast<const uint4 *>(blk + 8);
    const unsigned x0 = first ? (a.x & 0xffff0000u) : a.x;
    return (x0 | a.y | a.z | a.w | b.x | b.y | b.z | b.w) != 0;
}

template <bool WRITE>
__global__ void __launch_bounds__(ENT_WARPS * 32)
k_entropy_scan(const int16_t *__restrict__ MB, const int *__restrict__ nz, const int *__restrict__ parts, int mb_width,
               int mb_height, int P, unsigned *__restrict__ g_num, unsigned *__restrict__ g_den,
               unsigned *__restrict__ g_tail, uint8_t *__restrict__ third_context, int *__restrict__ mb_tokens,
               const int *__restrict__ mb_offset, const unsigned *__restrict__ part_info, uint16_t *__restrict__ tokens,
               unsigned capacity) {
    __shared__ __align__(16) int16_t s_coef[ENT_WARPS][400];
    __shared__ unsigned s_num[WRITE ? 1 : 1056], s_den[WRITE ? 1 : 1056], s_tail[WRITE ? 1 : 4 * 17];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.y, col = blockIdx.x * ENT_WARPS + warp;
    const int p = row % P;
    if (!WRITE) {
        for (int i = threadIdx.x; i < 1056; i += ENT_WARPS * 32) s_num[i] = s_den[i] = 0;
        if (threadIdx.x < 4 * 17) s_tail[threadIdx.x] = 0;
        __syncthreads();
    } else if (part_info[2 * P] > capacity) {
        return;  // the streams do not fit: the host codes this frame from the coefficients instead
    }
    const int mb = row * mb_width + col;
    const bool coded = col < mb_width && nz[mb] != 0;  // a skipped macroblock codes nothing, contexts stay as they were
    if (coded) {
        // stage the macroblock's 400 coefficients (coalesced 16-byte loads)
        const uint4 *src = reinterpret_cast<const uint4 *>(MB + (size_t)mb * 400);
        uint4 *dst = reinterpret_cast<uint4 *>(s_coef[warp]);
        for (int i = lane; i < 50; i += 32) dst[i] = src[i];
        __syncwarp();
        const bool has_y2 = parts[mb] == 0;
        // lanes in coding order: [Y2] Y0..Y15 U0..U3 V0..V3
        const int b = has_y2 ? (lane == 0 ? 24 : lane - 1) : lane;
        const bool active = has_y2 ? lane < 25 : lane < 24;
        // non-zero mask of the own block
        unsigned mask = 0;
        if (active) {
            const int16_t *c = s_coef[warp] + b * 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) mask |= (c[i] != 0 ? 1u : 0u) << i;
        }
        // "non-empty" flags of all 25 own blocks as seen by a neighbour (a Y block's DC lives in Y2 when there is one)
        const bool own_flag = active && ((b < 16 && has_y2) ? (mask & ~1u) != 0 : mask != 0);
        const unsigned own_bits_lane = __ballot_sync(0xffffffffu, own_flag);          // bit = lane
        const unsigned own = has_y2 ? ((own_bits_lane >> 1) | ((own_bits_lane & 1u) << 24)) : own_bits_lane;  // bit = block
        // the eight blocks of the macroblock above and the eight of the macroblock to the left that touch this one
        bool nb_flag = false;
        if (lane < 8 && row > 0) {
            const int pm = mb - mb_width;
            const int pb = lane < 4 ? 12 + lane : (lane < 6 ? 18 + (lane - 4) : 22 + (lane - 6));
            nb_flag = ent_block_nonzero(MB + (size_t)pm * 400 + pb * 16, (pb < 16 && parts[pm] == 0) ? 1 : 0);
        } else if (lane >= 8 && lane < 16 && col > 0) {
            const int pm = mb - 1, l = lane - 8;
            const int pb = l < 4 ? 4 * l + 3 : (l < 6 ? 17 + 2 * (l - 4) : 21 + 2 * (l - 6));
            nb_flag = ent_block_nonzero(MB + (size_t)pm * 400 + pb * 16, (pb < 16 && parts[pm] == 0) ? 1 : 0);
        }
        const unsigned nb = __ballot_sync(0xffffffffu, nb_flag);  // bits 0-3 above Y, 4-5 above U, 6-7 above V, 8-11 left Y, 12-13 left U, 14-15 left V
        int ctx = 0, type = 3;
        if (active) {
            if (b == 24) {
                // Y2: the nearest macroblocks above / to the left that have a Y2 block
                type = 1;
                if (row > 0) {
                    int q = mb - mb_width;
                    while (q >= 0 && parts[q] != 0) q -= mb_width;
                    if (q >= 0) ctx += ent_block_nonzero(MB + (size_t)q * 400 + 24 * 16, 0);
                }
                if (col > 0) {
                    int q = mb - 1;
                    while (q >= row * mb_width && parts[q] != 0) --q;
                    if (q >= row * mb_width) ctx += ent_block_nonzero(MB + (size_t)q * 400 + 24 * 16, 0);
                }
            } else if (b < 16) {
                type = has_y2 ? 0 : 3;
                const bool up = (b >> 2) > 0 ? ((own >> (b - 4)) & 1) : ((nb >> (b & 3)) & 1);
                const bool lf = (b & 3) > 0 ? ((own >> (b - 1)) & 1) : ((nb >> (8 + (b >> 2))) & 1);
                ctx = (int)up + (int)lf;
            } else {
                type = 2;
                const int pl = (b - 16) >> 2, k = (b - 16) & 3;
                const bool up = (k >> 1) > 0 ? ((own >> (b - 2)) & 1) : ((nb >> (4 + 2 * pl + (k & 1))) & 1);
                const bool lf = (k & 1) > 0 ? ((own >> (b - 1)) & 1) : ((nb >> (12 + 2 * pl + (k >> 1))) & 1);
                ctx = (int)up + (int)lf;
            }
        }
        const int16_t *blk = s_coef[warp] + b * 16;
        if (!WRITE) {
            StatSink e{s_num, s_den, 0};
            if (active) {
                third_context[(size_t)mb * 25 + b] = (uint8_t)ctx;
                const int eob = ent_walk_block(blk, type, ctx, mask, e);
                atomicAdd(&s_tail[type * 17 + (eob < 16 ? eob + 1 : 16)], 1u);
            }
            int n = e.n;  // total decisions of the macroblock
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
            if (lane == 0) mb_tokens[mb] = n;
        } else {
            // this lane's decisions start after those of the lanes before it (coding order = lane order):
            // count (the walk is short; recomputing beats keeping per-block counts in memory), scan, write
            CountSink c{0};
            if (active) ent_walk_block(blk, type, ctx, mask, c);
            int incl = c.n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (active) {
                WriteSink w{tokens + part_info[p] + (unsigned)mb_offset[mb] + (unsigned)(incl - c.n)};
                ent_walk_block(blk, type, ctx, mask, w);
            }
        }
    } else if (!WRITE && col < mb_width && lane == 0) {
        mb_tokens[mb] = 0;
    }
    if (!WRITE) {
        __syncthreads();
        unsigned *num = g_num + (size_t)p * 1056, *den = g_den + (size_t)p * 1056;
        for (int i = threadIdx.x; i < 1056; i += ENT_WARPS * 32) {
            const unsigned d = s_den[i];
            if (d) {
                atomicAdd(&den[i], d);
                const unsigned z = s_num[i];
                if (z) atomicAdd(&num[i], z);
            }
        }
        if (threadIdx.x < 4 * 17 && s_tail[threadIdx.x]) atomicAdd(&g_tail[p * 68 + threadIdx.x], s_tail[threadIdx.x]);
    }
}

// num = 0, den = 1, tails = 0 (count_probs starts every table at 0 / 1, src/CPU_kernels.cl:541-560)
__global__ void k_entropy_init(unsigned *num, unsigned *den, unsigned *tail, int P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P * 1056) {
        num[i] = 0;
        den[i] = 1;
    }
    if (i < P * 68) tail[i] = 0;
}

// one CTA per partition: exclusive scan of the macroblocks' decision counts in coding order
// (rows p, p+P, ... ; columns left to right)
__global__ void __launch_bounds__(1024)
k_entropy_offsets(const int *__restrict__ mb_tokens, int *__restrict__ mb_offset, unsigned *__restrict__ part_count,
                  int mb_width, int mb_height, int P) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rows = (mb_height - p + P - 1) / P;
    const int count = rows * mb_width;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < count; base += 1024) {
        const int i = base + tid;
        int mb = -1, v = 0;
        if (i < count) {
            mb = (p + (i / mb_width) * P) * mb_width + i % mb_width;
            v = mb_tokens[mb];
        }
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            s_warp[lane] = wi - w;  // exclusive
        }
        __syncthreads();
        const int carry = s_carry;
        if (mb >= 0) mb_offset[mb] = carry + s_warp[warp] + incl - v;
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_warp[31] + incl;
        __syncthreads();
    }
    if (tid == 0) part_count[p] = (unsigned)s_carry;
}

// part_info: [0..P) stream base of every partition, [P..2P) decision count, [2P] total.  Also expands the
// tail histograms: a block whose end-of-block token sat at position i-1 contributes one end-of-block
// decision with context 2 at every position i..15 (src/CPU_kernels.cl:503-538).
__global__ void k_entropy_finish(unsigned *part_info, unsigned *num, unsigned *den, const unsigned *tail, int P) {
    const int t = threadIdx.x;
    if (t == 0) {
        unsigned base = 0;
        for (int p = 0; p < P; ++p) {
            part_info[p] = base;
            base += part_info[P + p];
        }
        part_info[2 * P] = base;
    }
    if (t < P * 4) {
        const int p = t >> 2, type = t & 3;
        unsigned running = 0;
        for (int i = 1; i < 16; ++i) {
            running += tail[p * 68 + type * 17 + i];
            const int idx = ent_ctx_index(type, ent_band(i), 2, 0);
            num[(size_t)p * 1056 + idx] += running;
            den[(size_t)p * 1056 + idx] += running;
        }
    }
}

// ------------------------------------------------------------------------------------------
// RFC 6386 section 7.3 boolean entropy encoder over the decision streams: one warp per partition.  The coder is
// a serial recurrence (range, bottom, bit count), so one lane runs it; the warp resolves the probabilities of
// the next 256 decisions into shared memory meanwhile (coalesced stream reads, table lookups), which leaves the
// serial lane a shared-memory load per decision.  Same arithmetic as entropy_host.cpp's BoolSink (renormalisation
// in at most two strides, carries walked back through the bytes already written), hence the same bytes.
// Opt-in (VP8B200_GPU_BOOLCODER=1) and, as measured, not yet worth it: a lone lane retires an instruction every
// ~8 cycles, about 110 ns per decision, 7 ms per 1080p frame against 0.3 ms on eight host threads (one instance
// 101 instead of 257 frames/s, 32 instances 1494 instead of 2260).  It costs the host nothing, so it is the
// building block for the parallel formulation (range as a 128-state machine scanned over chunks).
constexpr int BC_CHUNK = 256;

struct BoolCoder {
    uint8_t *out;
    uint32_t range = 255, bottom = 0, count = 0;
    int bit_count = 24;
    __device__ __forceinline__ static void carry(uint8_t *q) {
        while (*--q == 255) *q = 0;
        ++*q;
    }
    // the renormalisation of entropy_host.cpp's BoolSink::put from the point where range and bottom have taken
    // the decision in: at most two strides, a byte leaves between them, carries walk back through the output
    __device__ __forceinline__ void renormalise_general(int s) {
        if (s >= bit_count) {
            const int c = bit_count;
            for (uint32_t t = bottom >> (32 - c); t; t &= t - 1) carry(out);
            bottom <<= c;
            *out++ = (uint8_t)(bottom >> 24);
            ++count;
            bottom &= (1u << 24) - 1;
            bit_count = 8;
            s -= c;
        }
        const unsigned long long wide = (unsigned long long)bottom << s;
        for (uint32_t t = (uint32_t)(wide >> 32); t; t &= t - 1) carry(out);
        bottom = (uint32_t)wide;
        bit_count -= s;
    }
    __device__ __forceinline__ void put(uint32_t prob, uint32_t bit) {
        const uint32_t split = 1 + (((range - 1) * prob) >> 8);
        const uint32_t r = bit ? range - split : split;
        bottom += bit ? split : 0u;
        const int s = __clz(r) - 24;
        range = r << s;
        // Common case: no byte leaves (s < bit_count) and none of the s bits shifted out of bottom is set (no carry):
        // plain shift.  __funnelshift_l(bottom, 0, s) is bottom >> (32 - s), 0 for s == 0.
        if (s >= bit_count || __funnelshift_l(bottom, 0u, s) != 0u) {
            renormalise_general(s);
        } else {
            bottom <<= s;
            bit_count -= s;
        }
    }
    __device__ __forceinline__ void finish() {
        int c = bit_count;
        uint32_t v = bottom;
        if (v & (1u << (32 - c))) carry(out);
        v <<= c & 7;
        c >>= 3;
        while (--c >= 0) v <<= 8;
        for (c = 0; c < 4; ++c) {
            *out++ = (uint8_t)(v >> 24);
            ++count;
            v <<= 8;
        }
    }
};

__global__ void __launch_bounds__(32) k_entropy_boolcode(const uint16_t *__restrict__ tokens, const uint32_t *__restrict__ part_info,
                                                         const uint32_t *__restrict__ coeff_probs, uint8_t *output,
                                                         int32_t *partition_sizes, int P, int partition_step) {
    __shared__ uint8_t s_tab[1056 + 256];          // one table for both kinds of entry (slot / fixed probability)
    __shared__ __align__(16) uint16_t s_dec[2][BC_CHUNK];  // probability | bit << 8 of a chunk of decisions, double buffered
    const int p = blockIdx.x, lane = threadIdx.x;
    for (int i = lane; i < 1056 + 256; i += 32) s_tab[i] = i < 1056 ? (uint8_t)coeff_probs[i] : (uint8_t)(i - 1056);
    const uint16_t *t = tokens + part_info[p];
    const uint32_t n = part_info[P + p];
    __syncwarp();
    auto resolve = [&](uint32_t base, int buf) {
#pragma unroll
        for (int k = 0; k < BC_CHUNK / 32; ++k) {
            const uint32_t i = base + k * 32 + lane;
            if (i < n) {
                const uint32_t e = t[i];
                s_dec[buf][k * 32 + lane] = (uint16_t)(s_tab[e & 0x7ff] | ((e >> 15) << 8));
            }
        }
    };
    BoolCoder bc;
    bc.out = output + (size_t)partition_step * p;
    if (n) resolve(0, 0);
    __syncwarp();
    int buf = 0;
    for (uint32_t base = 0; base < n; base += BC_CHUNK, buf ^= 1) {
        if (lane == 0) {
            const uint32_t m = min((uint32_t)BC_CHUNK, n - base);
            uint32_t j = 0;
            for (; j + 8 <= m; j += 8) {  // eight decisions per 16-byte shared load
                const uint4 q = *reinterpret_cast<const uint4 *>(&s_dec[buf][j]);
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    bc.put(w[e] & 255u, (w[e] >> 8) & 1u);
                    bc.put((w[e] >> 16) & 255u, w[e] >> 24);
                }
            }
            for (; j < m; ++j) {
                const uint32_t d = s_dec[buf][j];
                bc.put(d & 255u, d >> 8);
            }
        } else if (base + BC_CHUNK < n) {
            resolve(base + BC_CHUNK, buf ^ 1);  // the other 31 lanes prepare the next chunk (lane 0's share: below)
        }
        __syncwarp();
        if (base + BC_CHUNK < n) {  // entries of lane 0 of the next chunk
#pragma unroll
            for (int k = 0; k < BC_CHUNK / 32; ++k) {
                const uint32_t i = base + BC_CHUNK + k * 32;
                if (lane == 0 && i < n) {
                    const uint32_t e = t[i];
                    s_dec[buf ^ 1][k * 32] = (uint16_t)(s_tab[e & 0x7ff] | ((e >> 15) << 8));
                }
            }
        }
        __syncwarp();
    }
    if (lane == 0) {
        bc.finish();
        partition_sizes[p] = (int32_t)bc.count;
    }
}

}  // namespace vp8

using namespace vp8;

extern "C" int vp8b200_entropy_boolcode(void *stream, const uint16_t *tokens, const uint32_t *part_info,
                                        const uint32_t *coeff_probs, uint8_t *output, int32_t *partition_sizes,
                                        int num_partitions, int partition_step) {
    if (num_partitions < 1 || num_partitions > 8) return -(int)cudaErrorInvalidValue;
    k_entropy_boolcode<<<num_partitions, 32, 0, (cudaStream_t)stream>>>(tokens, part_info, coeff_probs, output, partition_sizes,
                                                                         num_partitions, partition_step);
    VP8_LAUNCH_CHECK();
}

extern "C" int vp8b200_entropy_tokens(void *stream, const int16_t *MB, const int32_t *MB_non_zero_coeffs,
                                      const int32_t *MB_parts, int mb_width, int mb_height, int num_partitions,
                                      uint32_t *coeff_probs, uint32_t *coeff_probs_denom, uint8_t *third_context,
                                      uint16_t *tokens, uint32_t capacity, int32_t *mb_tokens, int32_t *mb_offset,
                                      uint32_t *part_info, uint32_t *tail_scratch) {
    const int P = num_partitions;
    if (mb_width <= 0 || mb_height <= 0) return 0;
    if (P < 1 || P > 8) return -(int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    k_entropy_init<<<(P * 1056 + 255) / 256, 256, 0, st>>>(coeff_probs, coeff_probs_denom, tail_scratch, P);
    dim3 grid((mb_width + ENT_WARPS - 1) / ENT_WARPS, mb_height);
    k_entropy_scan<false><<<grid, ENT_WARPS * 32, 0, st>>>(MB, MB_non_zero_coeffs, MB_parts, mb_width, mb_height, P, coeff_probs,
                                                          coeff_probs_denom, tail_scratch, third_context, mb_tokens, nullptr,
                                                          nullptr, nullptr, 0);
    k_entropy_offsets<<<P, 1024, 0, st>>>(mb_tokens, mb_offset, part_info + P, mb_width, mb_height, P);
    k_entropy_finish<<<1, 32, 0, st>>>(part_info, coeff_probs, coeff_probs_denom, tail_scratch, P);
    k_entropy_scan<true><<<grid, ENT_WARPS * 32, 0, st>>>(MB, MB_non_zero_coeffs, MB_parts, mb_width, mb_height, P, nullptr,
                                                         nullptr, nullptr, nullptr, nullptr, mb_offset, part_info, tokens,
                                                         capacity);
    VP8_LAUNCH_CHECK();
}
