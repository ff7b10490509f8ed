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
        // (loads first, then the sums per coefficient band, then one update per band: the fifteen positions
        // share seven bands, and nothing here waits for a store it has just issued)
        unsigned tl[16], add[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int i = 1; i < 16; ++i) tl[i] = tail[p * 68 + type * 17 + i];
        unsigned running = 0;
#pragma unroll
        for (int i = 1; i < 16; ++i) {
            running += tl[i];
            add[ent_band(i)] += running;
        }
        unsigned n0[8], d0[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int idx = ent_ctx_index(type, b, 2, 0);
            n0[b] = num[(size_t)p * 1056 + idx];
            d0[b] = den[(size_t)p * 1056 + idx];
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int idx = ent_ctx_index(type, b, 2, 0);
            num[(size_t)p * 1056 + idx] = n0[b] + add[b];
            den[(size_t)p * 1056 + idx] = d0[b] + add[b];
        }
    }
}

// ------------------------------------------------------------------------------------------
// RFC 6386 section 7.3 boolean entropy encoder over the decision streams, in parallel.
//
// The coder looks serial (range, bottom, bit count), but its output is a plain number.  With R_i the range before
// decision i, split_i = 1 + ((R_i - 1) * prob_i >> 8), s_i the normalisation shift of decision i and
// T_i = s_0 + .. + s_(i-1), the bytes of a partition are the big-endian digits of
//     L = sum over the decisions with value 1 of  split_i << (T' - T_i),     T' = 24 + 8 m,
// m = bytes that leave during coding (the k with 24 + 8 k <= T_n), m + 4 bytes in all (the flush writes four).
// Carry propagation into bytes already written is just the carries of this sum.  (Checked against the serial
// coder of entropy_host.cpp, which is checked against the reference's kernel.)  So:
//   A  the range is a state machine with 128 states: a warp walks a chunk of BC_CHUNK decisions from all 128
//      possible ranges at once (4 per lane; after 32 decisions one per lane for the ranges that are still
//      distinct) and leaves the chunk's map: range in -> range out, sum of shifts;
//   B  one thread per partition chains the maps out of shared memory: range and T at the start of every chunk, T_n;
//   C  one thread per chunk walks it again from its now known start and adds split_i << (T' - T_i) into the
//      partition's 32-bit words (64-bit accumulators: a few threads touch a word);
//   D  per partition the words are reduced to 32 bits each, the remaining 0/1 carries resolved by a
//      generate/propagate scan, and the bytes written most significant first.
constexpr int BC_CHUNK = 128;

struct BoolcodeScratch {  // carved out of one allocation, see vp8b200_entropy_boolcode_scratch_bytes
    uint8_t *map_range;            // [chunks][128]
    uint16_t *map_shift;           // [chunks][128]
    uint8_t *start_range;          // [chunks]
    uint32_t *start_shift;         // [chunks]
    uint32_t *params;              // [P][4]: T', bytes, words, -
    unsigned long long *words;     // [P][words_per_partition]
    uint32_t words_per_partition;
};

__device__ __forceinline__ uint32_t bc_chunk_base(const uint32_t *part_info, int P, int p) {
    uint32_t base = 0;
    for (int q = 0; q < p; ++q) base += (part_info[P + q] + BC_CHUNK - 1) / BC_CHUNK;
    return base;
}
// one decision from range R: returns the new range, adds the shift to T, hands back split
__device__ __forceinline__ uint32_t bc_step(uint32_t R, uint32_t prob, bool bit, uint32_t &T, uint32_t &split) {
    split = 1 + (((R - 1) * prob) >> 8);
    const uint32_t r = bit ? R - split : split;
    const int s = __clz(r) - 24;
    T += s;
    return r << s;
}

// A: blockIdx.y = partition, one warp per chunk
constexpr int BC_A_WARPS = 4;
__global__ void __launch_bounds__(BC_A_WARPS * 32) k_boolcode_maps(const uint16_t *__restrict__ tokens,
                                                                   const uint32_t *__restrict__ part_info,
                                                                   const uint32_t *__restrict__ coeff_probs, int P,
                                                                   BoolcodeScratch sc) {
    __shared__ uint8_t s_tab[1056 + 256];  // one table for both kinds of entry (slot / fixed probability)
    for (int i = threadIdx.x; i < 1056 + 256; i += blockDim.x) s_tab[i] = i < 1056 ? (uint8_t)coeff_probs[i] : (uint8_t)(i - 1056);
    __syncthreads();
    const int p = blockIdx.y, lane = threadIdx.x & 31;
    const uint32_t k = blockIdx.x * BC_A_WARPS + (threadIdx.x >> 5);
    const uint32_t n = part_info[P + p];
    if (k * BC_CHUNK >= n) return;  // whole warp
    const uint16_t *t = tokens + part_info[p] + (size_t)k * BC_CHUNK;
    const uint32_t cnt = min((uint32_t)BC_CHUNK, n - k * BC_CHUNK);
    __shared__ uint32_t s_seen[BC_A_WARPS][4];
    const int warp = threadIdx.x >> 5;
    uint32_t R[4], T[4] = {0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < 4; ++q) R[q] = 128 + 32 * q + lane;
    // 32 decisions (one coalesced read of the stream) applied to NS ranges per lane
    auto batch = [&](uint32_t j0, uint32_t *Rs, uint32_t *Ts, auto ns) {
        constexpr int NS = decltype(ns)::value;
        uint32_t d = 0;
        if (j0 + lane < cnt) {
            const uint32_t e = t[j0 + lane];
            d = s_tab[e & 0x7ff] | ((e >> 15) << 8);
        }
        const int m = (int)min(32u, cnt - j0);
        auto one = [&](int j) {
            const uint32_t dj = __shfl_sync(0xffffffffu, d, j);
            const uint32_t prob = dj & 255u;
            const bool bit = (dj >> 8) != 0;  // warp-uniform
            uint32_t split;
#pragma unroll
            for (int q = 0; q < NS; ++q) Rs[q] = bc_step(Rs[q], prob, bit, Ts[q], split);
        };
        if (m == 32) {
#pragma unroll 8
            for (int j = 0; j < 32; ++j) one(j);
        } else {
            for (int j = 0; j < m; ++j) one(j);
        }
    };
    struct Four { enum { value = 4 }; };
    struct One { enum { value = 1 }; };
    batch(0, R, T, Four());
    if (cnt > 32) {
        // The 128 trajectories fall together quickly (about 15 distinct ranges are left after 32 decisions of a
        // typical stream).  If at most 32 are left, the rest of the chunk is walked once per distinct range, one per
        // lane, and every start range inherits the outcome of the one it has merged with.
        if (lane < 4) s_seen[warp][lane] = 0;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) atomicOr(&s_seen[warp][(R[q] - 128) >> 5], 1u << ((R[q] - 128) & 31));
        __syncwarp();
        const uint32_t m0 = s_seen[warp][0], m1 = s_seen[warp][1], m2 = s_seen[warp][2], m3 = s_seen[warp][3];
        const int c0 = __popc(m0), c1 = __popc(m1), c2 = __popc(m2), c3 = __popc(m3);
        if (c0 + c1 + c2 + c3 <= 32) {  // warp-uniform
            int rep[4];  // which lane carries the range this one has become
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = (int)R[q] - 128, w = idx >> 5, b = idx & 31;
                const uint32_t mw = w == 0 ? m0 : w == 1 ? m1 : w == 2 ? m2 : m3;
                rep[q] = (w > 0 ? c0 : 0) + (w > 1 ? c1 : 0) + (w > 2 ? c2 : 0) + __popc(mw & ((1u << b) - 1u));
            }
            // lane r carries the r-th distinct range (lanes past the count carry a dummy)
            uint32_t Rr = 128, Tr = 0;
            {
                int r = lane;
                if (r < c0) Rr = 128 + __fns(m0, 0, r + 1);
                else if ((r -= c0) < c1) Rr = 160 + __fns(m1, 0, r + 1);
                else if ((r -= c1) < c2) Rr = 192 + __fns(m2, 0, r + 1);
                else if ((r -= c2) < c3) Rr = 224 + __fns(m3, 0, r + 1);
            }
            for (uint32_t j0 = 32; j0 < cnt; j0 += 32) batch(j0, &Rr, &Tr, One());
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                R[q] = __shfl_sync(0xffffffffu, Rr, rep[q]);
                T[q] += __shfl_sync(0xffffffffu, Tr, rep[q]);
            }
        } else {
            for (uint32_t j0 = 32; j0 < cnt; j0 += 32) batch(j0, R, T, Four());
        }
    }
    const size_t c = (size_t)bc_chunk_base(part_info, P, p) + k;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        sc.map_range[c * 128 + 32 * q + lane] = (uint8_t)R[q];
        sc.map_shift[c * 128 + 32 * q + lane] = (uint16_t)T[q];
    }
}

// B: one CTA of eight warps per partition.  A round takes eight blocks of 32 chunks: all threads stage their
// maps into shared memory; warp w composes block w for all 128 ranges at once (4 per lane: the block's own map);
// one thread chains the eight block maps; then lane 0 of warp w walks block w from its now known start and leaves
// the start of every chunk.  Range and sum of shifts carry over from round to round.  At the end everybody clears
// the partition's words.
constexpr int BC_B_BLOCK = 32, BC_B_WARPS = 8;
constexpr size_t BC_B_SMEM = (size_t)BC_B_WARPS * BC_B_BLOCK * 128 * 3;  // ranges (1 byte) + shifts (2 bytes)
__global__ void __launch_bounds__(BC_B_WARPS * 32) k_boolcode_chain(const uint32_t *__restrict__ part_info, int P,
                                                                    BoolcodeScratch sc, int partition_step) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    uint16_t *s_s = reinterpret_cast<uint16_t *>(s_dyn);                           // [8][32][128]
    uint8_t *s_r = s_dyn + (size_t)BC_B_WARPS * BC_B_BLOCK * 128 * 2;              // [8][32][128]
    __shared__ uint8_t s_comp_r[BC_B_WARPS][128];
    __shared__ uint32_t s_comp_t[BC_B_WARPS][128];
    __shared__ uint32_t s_block_r[BC_B_WARPS], s_block_t[BC_B_WARPS], s_carry[2], s_words;
    const int p = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t n = part_info[P + p];
    const uint32_t K = (n + BC_CHUNK - 1) / BC_CHUNK;
    const size_t cb = bc_chunk_base(part_info, P, p);
    constexpr uint32_t ROUND = BC_B_WARPS * BC_B_BLOCK;  // chunks per round
    if (tid == 0) {
        s_carry[0] = 255;
        s_carry[1] = 0;
    }
    for (uint32_t k0 = 0; k0 < K; k0 += ROUND) {
        const uint32_t chunks = min(ROUND, K - k0);
        {   // stage (the scratch regions are 256-byte aligned and a chunk's map is 128 entries: 16-byte copies)
            const uint4 *gr = reinterpret_cast<const uint4 *>(sc.map_range + (cb + k0) * 128);
            const uint4 *gs = reinterpret_cast<const uint4 *>(sc.map_shift + (cb + k0) * 128);
            uint4 *dr = reinterpret_cast<uint4 *>(s_r), *ds = reinterpret_cast<uint4 *>(s_s);
            for (uint32_t i = tid; i < chunks * 8; i += blockDim.x) dr[i] = gr[i];
            for (uint32_t i = tid; i < chunks * 16; i += blockDim.x) ds[i] = gs[i];
        }
        __syncthreads();
        const uint32_t first = warp * BC_B_BLOCK;                       // this warp's block inside the round
        const uint32_t mine = first < chunks ? min((uint32_t)BC_B_BLOCK, chunks - first) : 0;
        {   // the block's own map
            uint32_t R[4], T[4] = {0, 0, 0, 0};
#pragma unroll
            for (int q = 0; q < 4; ++q) R[q] = 128 + 32 * q + lane;
            for (uint32_t k = 0; k < mine; ++k) {
                const uint8_t *mr = s_r + (size_t)(first + k) * 128 - 128;
                const uint16_t *ms = s_s + (size_t)(first + k) * 128 - 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    T[q] += ms[R[q]];
                    R[q] = mr[R[q]];
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                s_comp_r[warp][32 * q + lane] = (uint8_t)R[q];
                s_comp_t[warp][32 * q + lane] = T[q];
            }
        }
        __syncthreads();
        if (tid == 0) {  // where every block starts
            uint32_t state = s_carry[0], T = s_carry[1];
            for (int w = 0; w < BC_B_WARPS; ++w) {
                s_block_r[w] = state;
                s_block_t[w] = T;
                T += s_comp_t[w][state - 128];
                state = s_comp_r[w][state - 128];   // (blocks past the end are the identity with no shift)
            }
            s_carry[0] = state;
            s_carry[1] = T;
        }
        __syncthreads();
        if (lane == 0 && mine) {  // the starts of this block's chunks
            uint32_t state = s_block_r[warp], T = s_block_t[warp];
            for (uint32_t k = 0; k < mine; ++k) {
                const size_t c = cb + k0 + first + k;
                sc.start_range[c] = (uint8_t)state;
                sc.start_shift[c] = T;
                const uint32_t at = (first + k) * 128 + state - 128;
                T += s_s[at];
                state = s_r[at];
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t T = s_carry[1];
        const uint32_t m = T >= 24 ? (T - 24) / 8 + 1 : 0;
        const uint32_t bytes = m + 4, words = (bytes + 3) / 4;
        // (a partition that does not fit its slot cannot be written; the reference has the same limit)
        const bool fits = bytes <= (uint32_t)partition_step && words + 2 <= sc.words_per_partition;
        sc.params[4 * p + 0] = 24 + 8 * m;
        sc.params[4 * p + 1] = fits ? bytes : 0;
        sc.params[4 * p + 2] = fits ? words : 0;
        s_words = fits ? words + 2 : 0;
    }
    __syncthreads();
    unsigned long long *w = sc.words + (size_t)p * sc.words_per_partition;
    for (uint32_t i = tid; i < s_words; i += blockDim.x) w[i] = 0ull;
}

// C: one thread per chunk
__global__ void __launch_bounds__(128) k_boolcode_terms(const uint16_t *__restrict__ tokens, const uint32_t *__restrict__ part_info,
                                                        const uint32_t *__restrict__ coeff_probs, int P, BoolcodeScratch sc) {
    __shared__ uint8_t s_tab[1056 + 256];
    for (int i = threadIdx.x; i < 1056 + 256; i += blockDim.x) s_tab[i] = i < 1056 ? (uint8_t)coeff_probs[i] : (uint8_t)(i - 1056);
    __syncthreads();
    const int p = blockIdx.y;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = part_info[P + p];
    if (k * BC_CHUNK >= n || sc.params[4 * p + 2] == 0) return;
    const size_t c = (size_t)bc_chunk_base(part_info, P, p) + k;
    const uint16_t *t = tokens + part_info[p] + (size_t)k * BC_CHUNK;
    const uint32_t cnt = min((uint32_t)BC_CHUNK, n - k * BC_CHUNK);
    const uint32_t Tend = sc.params[4 * p + 0];
    unsigned long long *words = sc.words + (size_t)p * sc.words_per_partition;
    uint32_t R = sc.start_range[c], T = sc.start_shift[c];
    int cur = -1;
    unsigned long long acc = 0;
    auto flush = [&]() {
        if (cur >= 0 && acc) {
            atomicAdd(words + cur, acc & 0xffffffffull);
            if (acc >> 32) atomicAdd(words + cur + 1, acc >> 32);
        }
    };
    auto one = [&](uint32_t d) {
        const uint32_t E = Tend - T;  // exponent of this decision's split
        uint32_t split;
        R = bc_step(R, d & 255u, (d >> 8) != 0, T, split);
        if (d >> 8) {
            const int w = (int)(E >> 5);
            if (w != cur) {
                flush();
                cur = w;
                acc = 0;
            }
            acc += (unsigned long long)split << (E & 31);
        }
    };
    uint32_t j = 0;
    for (; j + 8 <= cnt; j += 8) {  // eight loads and table look-ups in flight ahead of the serial part
        uint32_t d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = t[j + i];
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = s_tab[d[i] & 0x7ff] | ((d[i] >> 15) << 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) one(d[i]);
    }
    for (; j < cnt; ++j) {
        const uint32_t e = t[j];
        one(s_tab[e & 0x7ff] | ((e >> 15) << 8));
    }
    flush();
}

// D: one CTA per partition: words -> 32 bits each, carries, bytes
constexpr int BC_D_THREADS = 1024;
__global__ void __launch_bounds__(BC_D_THREADS) k_boolcode_emit(BoolcodeScratch sc, uint8_t *output, int32_t *partition_sizes,
                                                                int partition_step) {
    __shared__ uint32_t s_g[32], s_p[32];
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bytes = sc.params[4 * p + 1], W = sc.params[4 * p + 2] + 2;  // (two spare words: always zero at the end)
    if (tid == 0) partition_sizes[p] = (int32_t)bytes;
    if (bytes == 0) return;
    const unsigned long long *A = sc.words + (size_t)p * sc.words_per_partition;
    uint8_t *out = output + (size_t)partition_step * p;
    // a thread owns `per` consecutive words; blocks of per * BC_D_THREADS words are chained through carry_in
    const uint32_t per = (W + BC_D_THREADS - 1) / BC_D_THREADS;
    const uint32_t w0 = min(W, tid * per), w1 = min(W, w0 + per);
    // local pass with carry-in 0: generate / propagate of the thread's words
    uint32_t g = 0, pr = 1, c = 0;
    for (uint32_t w = w0; w < w1; ++w) {
        const unsigned long long v = (A[w] & 0xffffffffull) + (w ? A[w - 1] >> 32 : 0ull) + c;
        c = (uint32_t)(v >> 32);
        pr &= (uint32_t)v == 0xffffffffu;
    }
    g = c;
    if (w0 == w1) pr = 1;  // no words: passes a carry through (there is none to pass)
    // inclusive scan of (g, p) over the threads: carry out of everything up to and including this thread
    uint32_t G = g, Pp = pr;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t g2 = __shfl_up_sync(0xffffffffu, G, o), p2 = __shfl_up_sync(0xffffffffu, Pp, o);
        if (lane >= o) {
            G = G | (Pp & g2);
            Pp = Pp & p2;
        }
    }
    if (lane == 31) {
        s_g[warp] = G;
        s_p[warp] = Pp;
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t wg = s_g[lane], wp = s_p[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t g2 = __shfl_up_sync(0xffffffffu, wg, o), p2 = __shfl_up_sync(0xffffffffu, wp, o);
            if (lane >= o) {
                wg = wg | (wp & g2);
                wp = wp & p2;
            }
        }
        s_g[lane] = wg;
    }
    __syncthreads();
    // carry into this thread = carry out of the previous thread (inclusive value of lane - 1, combined with the warps before)
    uint32_t cin_warp = warp ? s_g[warp - 1] : 0;
    uint32_t Gprev = __shfl_up_sync(0xffffffffu, G, 1), Pprev = __shfl_up_sync(0xffffffffu, Pp, 1);
    uint32_t cin = lane ? (Gprev | (Pprev & cin_warp)) : cin_warp;
    // final pass: the words with the true carry-in, bytes most significant first
    c = cin;
    for (uint32_t w = w0; w < w1; ++w) {
        const unsigned long long v = (A[w] & 0xffffffffull) + (w ? A[w - 1] >> 32 : 0ull) + c;
        c = (uint32_t)(v >> 32);
        const uint32_t word = (uint32_t)v;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t idx = 4 * w + b;  // byte index counted from the least significant end
            if (idx < bytes) out[bytes - 1 - idx] = (uint8_t)(word >> (8 * b));
        }
    }
}

}  // namespace vp8

using namespace vp8;

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

static size_t bc_align(size_t x) { return (x + 255) & ~(size_t)255; }
static void bc_carve(BoolcodeScratch &sc, void *scratch, uint32_t max_decisions, int P, int partition_step, size_t *total) {
    const size_t chunks = (size_t)max_decisions / BC_CHUNK + P + 1;
    const uint32_t wpp = (uint32_t)(partition_step / 4 + 4);
    char *base = (char *)scratch;
    size_t off = 0;
    sc.map_range = (uint8_t *)(base + off); off += bc_align(chunks * 128);
    sc.map_shift = (uint16_t *)(base + off); off += bc_align(chunks * 128 * 2);
    sc.start_range = (uint8_t *)(base + off); off += bc_align(chunks);
    sc.start_shift = (uint32_t *)(base + off); off += bc_align(chunks * 4);
    sc.params = (uint32_t *)(base + off); off += bc_align(8 * 4 * 4);
    sc.words = (unsigned long long *)(base + off); off += bc_align((size_t)P * wpp * 8);
    sc.words_per_partition = wpp;
    *total = off;
}

extern "C" size_t vp8b200_entropy_boolcode_scratch_bytes(uint32_t max_decisions, int num_partitions, int partition_step) {
    BoolcodeScratch sc;
    size_t total = 0;
    bc_carve(sc, nullptr, max_decisions, num_partitions, partition_step, &total);
    return total;
}

extern "C" int vp8b200_entropy_boolcode(void *stream, const uint16_t *tokens, const uint32_t *part_info,
                                        const uint32_t *coeff_probs, uint8_t *output, int32_t *partition_sizes,
                                        int num_partitions, int partition_step, uint32_t max_decisions, void *scratch) {
    const int P = num_partitions;
    if (P < 1 || P > 8 || partition_step < 4 || !scratch) return -(int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    BoolcodeScratch sc;
    size_t total = 0;
    bc_carve(sc, scratch, max_decisions, P, partition_step, &total);
    const unsigned chunks = max_decisions / BC_CHUNK + 1;  // per partition, at most
    k_boolcode_maps<<<dim3((chunks + BC_A_WARPS - 1) / BC_A_WARPS, P), BC_A_WARPS * 32, 0, st>>>(tokens, part_info, coeff_probs, P, sc);
    // (96 KB of dynamic shared memory: above the default limit; per device, so it is set on every call)
    if (cudaFuncSetAttribute(k_boolcode_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BC_B_SMEM) != cudaSuccess)
        return -(int)cudaGetLastError();
    k_boolcode_chain<<<P, BC_B_WARPS * 32, BC_B_SMEM, st>>>(part_info, P, sc, partition_step);
    k_boolcode_terms<<<dim3((chunks + 127) / 128, P), 128, 0, st>>>(tokens, part_info, coeff_probs, P, sc);
    k_boolcode_emit<<<P, BC_D_THREADS, 0, st>>>(sc, output, partition_sizes, partition_step);
    VP8_LAUNCH_CHECK();
}
