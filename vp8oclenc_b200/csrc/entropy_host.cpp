// See entropy_host.h.  VP8 coefficient token coding (RFC 6386 section 13) written as one block
// walker with pluggable sinks: a statistics sink (count_probs) and a boolean-coder sink
// (encode_coefficients).
//
// Inter frames are sparse (at 1080p, q=24: ~55 k non-zero coefficients in 204 k blocks), so the
// cost is per BLOCK, not per coefficient.  Each block is therefore summarised first with two SSE2
// compares (a 16-bit "which positions are non-zero" mask: x86-64 baseline, no -m flags), the walk
// visits only the positions up to the last non-zero one, and the reference's habit of counting an
// end-of-block decision at EVERY position after the end of the block is kept as a small histogram
// of "where the tail starts" that is expanded once per partition.
#include "entropy_host.h"

#include <emmintrin.h>

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace vp8host {
namespace {

enum Token { T_ZERO, T_ONE, T_TWO, T_THREE, T_FOUR, T_CAT1, T_CAT2, T_CAT3, T_CAT4, T_CAT5, T_CAT6, T_EOB };

// Decisions from the root of the RFC 6386 coefficient token tree to each token:
// (probability slot = tree node / 2, branch taken).
struct Path {
    int n;
    unsigned char slot[7], bit[7];
};
const Path kPath[12] = {
    /* ZERO  */ {2, {0, 1}, {1, 0}},
    /* ONE   */ {3, {0, 1, 2}, {1, 1, 0}},
    /* TWO   */ {5, {0, 1, 2, 3, 4}, {1, 1, 1, 0, 0}},
    /* THREE */ {6, {0, 1, 2, 3, 4, 5}, {1, 1, 1, 0, 1, 0}},
    /* FOUR  */ {6, {0, 1, 2, 3, 4, 5}, {1, 1, 1, 0, 1, 1}},
    /* CAT1  */ {6, {0, 1, 2, 3, 6, 7}, {1, 1, 1, 1, 0, 0}},
    /* CAT2  */ {6, {0, 1, 2, 3, 6, 7}, {1, 1, 1, 1, 0, 1}},
    /* CAT3  */ {7, {0, 1, 2, 3, 6, 8, 9}, {1, 1, 1, 1, 1, 0, 0}},
    /* CAT4  */ {7, {0, 1, 2, 3, 6, 8, 9}, {1, 1, 1, 1, 1, 0, 1}},
    /* CAT5  */ {7, {0, 1, 2, 3, 6, 8, 10}, {1, 1, 1, 1, 1, 1, 0}},
    /* CAT6  */ {7, {0, 1, 2, 3, 6, 8, 10}, {1, 1, 1, 1, 1, 1, 1}},
    /* EOB   */ {1, {0}, {0}},
};

// extra-bit probabilities of the six value categories (RFC 6386 13.2), base value and width
const unsigned char kCat1[] = {159}, kCat2[] = {165, 145}, kCat3[] = {173, 148, 140}, kCat4[] = {176, 155, 140, 135},
                    kCat5[] = {180, 157, 141, 134, 130},
                    kCat6[] = {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129};
const unsigned char *const kCatProb[6] = {kCat1, kCat2, kCat3, kCat4, kCat5, kCat6};
const int kCatBase[6] = {5, 7, 11, 19, 35, 67};
const int kCatBits[6] = {1, 2, 3, 4, 5, 11};

const int kBand[16] = {0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7};

inline Token classify(int mag) {
    if (mag <= 4) return (Token)mag;
    if (mag <= 6) return T_CAT1;
    if (mag <= 10) return T_CAT2;
    if (mag <= 18) return T_CAT3;
    if (mag <= 34) return T_CAT4;
    if (mag <= 66) return T_CAT5;
    return T_CAT6;
}

inline int ctx_index(int type, int band, int ctx, int slot) { return (((type << 3) + band) * 3 + ctx) * 11 + slot; }

inline const int16_t *block(const int16_t *MB, int mb, int b) { return MB + (size_t)mb * 400 + b * 16; }

// bit i set <=> coefficient i of the block is non-zero
inline unsigned nonzero_mask(const int16_t *c) {
    const __m128i z = _mm_setzero_si128();
    const __m128i a = _mm_cmpeq_epi16(_mm_loadu_si128(reinterpret_cast<const __m128i *>(c)), z);
    const __m128i b = _mm_cmpeq_epi16(_mm_loadu_si128(reinterpret_cast<const __m128i *>(c + 8)), z);
    return ~(unsigned)_mm_movemask_epi8(_mm_packs_epi16(a, b)) & 0xffffu;
}

// Walks one block.  type: 0 = Y after Y2 (starts at coefficient 1), 1 = Y2, 2 = chroma,
// 3 = Y without Y2.  ctx = number of non-empty neighbour blocks (above, left).  mask = the
// block's non-zero mask.  Returns the position of the end-of-block token (16: none, the block is
// full).  The reference's count_probs_in_block (src/CPU_kernels.cl:503-538) does NOT stop there:
// it visits an end-of-block decision with context 2 at every later position; the caller accounts
// for those from the returned position.
template <class Sink>
inline int walk_block(const int16_t *coef, int type, int ctx, unsigned mask, Sink &sink) {
    const int first = (type == 0) ? 1 : 0;
    if (first) mask &= ~1u;
    const int last = mask ? 31 - __builtin_clz(mask) : -1;  // positions > last are end-of-block tokens
    bool prev_zero = false;
    int i = first;
    for (; i <= last; ++i) {
        const int band = kBand[i];
        const int v = coef[i];
        if (v == 0) {
            // after a ZERO token an end-of-block cannot follow, so the first decision is implicit
            if (!prev_zero) sink.decision(ctx_index(type, band, ctx, 0), 1);
            sink.decision(ctx_index(type, band, ctx, 1), 0);
            prev_zero = true;
            ctx = 0;
            continue;
        }
        // the token tree of RFC 6386 13.2 written out (same decisions as kPath, no per-token loop)
        const int mag = v < 0 ? -v : v;
        const int base = ctx_index(type, band, ctx, 0);
        if (!prev_zero) sink.decision(base, 1);  // not end-of-block
        sink.decision(base + 1, 1);              // not ZERO
        if (mag == 1) {
            sink.decision(base + 2, 0);
            ctx = 1;
        } else {
            sink.decision(base + 2, 1);
            if (mag <= 4) {
                sink.decision(base + 3, 0);
                if (mag == 2) {
                    sink.decision(base + 4, 0);
                } else {
                    sink.decision(base + 4, 1);
                    sink.decision(base + 5, mag == 4);
                }
            } else {
                sink.decision(base + 3, 1);
                int c;  // value category
                if (mag <= 10) {
                    sink.decision(base + 6, 0);
                    sink.decision(base + 7, mag > 6);
                    c = mag > 6;
                } else {
                    sink.decision(base + 6, 1);
                    if (mag <= 34) {
                        sink.decision(base + 8, 0);
                        sink.decision(base + 9, mag > 18);
                        c = 2 + (mag > 18);
                    } else {
                        sink.decision(base + 8, 1);
                        sink.decision(base + 10, mag > 66);
                        c = 4 + (mag > 66);
                    }
                }
                const int extra = mag - kCatBase[c];
                for (int b = 0; b < kCatBits[c]; ++b) sink.literal(kCatProb[c][b], (extra >> (kCatBits[c] - 1 - b)) & 1);
            }
            ctx = 2;
        }
        sink.literal(128, v < 0);  // sign
        prev_zero = false;
    }
    if (i < 16) sink.decision(ctx_index(type, kBand[i], ctx, 0), 0);  // end of block (never right after a ZERO)
    return i;
}

struct StatSink {
    uint32_t *num, *den;  // this partition's [4][8][3][11] tables
    inline void decision(int idx, int bit) {
        num[idx] += 1 - bit;  // zeros are counted
        ++den[idx];
    }
    inline void literal(int, int) {}
};

// RFC 6386 section 7.3 boolean entropy encoder
struct BoolSink {
    uint8_t *out;
    const uint32_t *probs;
    uint32_t range = 255, bottom = 0, count = 0;
    int bit_count = 24;
    static void carry(uint8_t *q) {
        while (*--q == 255) *q = 0;
        ++*q;
    }
    inline void put(int prob, int bit) {
        const uint32_t split = 1 + (((range - 1) * (uint32_t)prob) >> 8);
        const uint32_t r = bit ? range - split : split;
        bottom += bit ? split : 0u;
        // The RFC renormalises bit by bit: before each of the s shifts a set top bit of `bottom` is a
        // carry into the bytes already written, and every time bit_count reaches 0 a byte leaves.
        // Same thing in at most two strides (s <= 7, so at most one byte leaves); s == 0 falls
        // through the first stride as a no-op, so there is no "needs renormalisation" branch.
        int s = __builtin_clz(r) - 24;
        range = r << s;
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
        const uint64_t wide = (uint64_t)bottom << s;
        for (uint32_t t = (uint32_t)(wide >> 32); t; t &= t - 1) carry(out);
        bottom = (uint32_t)wide;
        bit_count -= s;
    }
    inline void decision(int idx, int bit) { put((uint8_t)probs[idx], bit); }
    inline void literal(int prob, int bit) { put(prob, bit); }
    void finish() {
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

// number of non-empty neighbour blocks (above + left) of every block of one macroblock.
// own[b] = non-zero masks of this macroblock's blocks.
void neighbour_contexts(const int16_t *MB, const int32_t *parts, int mb, int mb_row, int mb_col, int mb_width,
                        const unsigned *own, uint8_t *ctx /* [25] */) {
    const int row_start = mb_row * mb_width;
    if (parts[mb] == 0) {
        // Y2: the neighbours are the nearest macroblocks above / to the left that have a Y2 block
        int n = 0;
        if (mb_row > 0) {
            int p = mb - mb_width;
            while (p >= 0 && parts[p] != 0) p -= mb_width;
            if (p >= 0) n += nonzero_mask(block(MB, p, 24)) != 0;
        }
        if (mb_col > 0) {
            int p = mb - 1;
            while (p >= row_start && parts[p] != 0) --p;
            if (p >= row_start) n += nonzero_mask(block(MB, p, 24)) != 0;
        }
        ctx[24] = (uint8_t)n;
    }
    // "non-empty" ignores the DC of a Y block whose macroblock has a Y2 block (the DC lives there)
    const unsigned own_y = parts[mb] == 0 ? 0xfffeu : 0xffffu;
    unsigned above_y[4] = {0, 0, 0, 0}, left_y[4] = {0, 0, 0, 0}, above_c[4] = {0, 0, 0, 0}, left_c[4] = {0, 0, 0, 0};
    if (mb_row > 0) {
        const int pm = mb - mb_width;
        const unsigned keep = parts[pm] == 0 ? 0xfffeu : 0xffffu;
        for (int k = 0; k < 4; ++k) above_y[k] = nonzero_mask(block(MB, pm, 12 + k)) & keep;
        above_c[0] = nonzero_mask(block(MB, pm, 18));
        above_c[1] = nonzero_mask(block(MB, pm, 19));
        above_c[2] = nonzero_mask(block(MB, pm, 22));
        above_c[3] = nonzero_mask(block(MB, pm, 23));
    }
    if (mb_col > 0) {
        const int pm = mb - 1;
        const unsigned keep = parts[pm] == 0 ? 0xfffeu : 0xffffu;
        for (int k = 0; k < 4; ++k) left_y[k] = nonzero_mask(block(MB, pm, 4 * k + 3)) & keep;
        left_c[0] = nonzero_mask(block(MB, pm, 17));
        left_c[1] = nonzero_mask(block(MB, pm, 19));
        left_c[2] = nonzero_mask(block(MB, pm, 21));
        left_c[3] = nonzero_mask(block(MB, pm, 23));
    }
    for (int b = 0; b < 16; ++b) {
        const unsigned up = (b >> 2) > 0 ? (own[b - 4] & own_y) : above_y[b & 3];
        const unsigned lf = (b & 3) > 0 ? (own[b - 1] & own_y) : left_y[b >> 2];
        ctx[b] = (uint8_t)((up != 0) + (lf != 0));
    }
    for (int pl = 0; pl < 2; ++pl) {
        const int base = 16 + 4 * pl;
        for (int k = 0; k < 4; ++k) {
            const unsigned up = (k >> 1) > 0 ? own[base + k - 2] : above_c[2 * pl + (k & 1)];
            const unsigned lf = (k & 1) > 0 ? own[base + k - 1] : left_c[2 * pl + (k >> 1)];
            ctx[base + k] = (uint8_t)((up != 0) + (lf != 0));
        }
    }
}

// ---- persistent helper threads, one job at a time (the shim is driven by a single host thread) ----
class Pool {
  public:
    void run(int n, const std::function<void(int)> &f) {
        if (n <= 1) {
            f(0);
            return;
        }
        {
            std::unique_lock<std::mutex> lk(m_);
            while ((int)threads_.size() < n - 1) {
                const int id = (int)threads_.size() + 1;
                threads_.emplace_back([this, id] { worker(id); });
                threads_.back().detach();  // they sleep on the condition variable until the process ends
            }
            job_ = &f;
            n_ = n;
            pending_ = n - 1;
            ++generation_;
        }
        wake_.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

  private:
    void worker(int id) {
        unsigned long seen = 0;
        for (;;) {
            const std::function<void(int)> *job;
            {
                std::unique_lock<std::mutex> lk(m_);
                wake_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (id >= n_) continue;
                job = job_;
            }
            (*job)(id);
            std::unique_lock<std::mutex> lk(m_);
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::mutex m_;
    std::condition_variable wake_, done_;
    std::vector<std::thread> threads_;
    const std::function<void(int)> *job_ = nullptr;
    int n_ = 0, pending_ = 0;
    unsigned long generation_ = 0;
};

// VP8B200_ENTROPY_THREADS=<k> caps the threads a frame's partitions are spread over (default: one per partition).
// With more encoder instances than cores on the machine fewer, longer jobs cost fewer wake-ups.
template <class F>
void for_each_partition(int n, F f) {
    static Pool *pool = new Pool;  // never destroyed: its threads outlive static destruction
    static const int cap = [] {
        const char *e = getenv("VP8B200_ENTROPY_THREADS");
        const int k = e ? atoi(e) : 0;
        return k > 0 ? k : 1 << 30;
    }();
    const int T = n < cap ? n : cap;
    if (T == n) {
        pool->run(n, std::function<void(int)>(f));
    } else {
        pool->run(T, std::function<void(int)>([&](int t) {
            for (int p = t; p < n; p += T) f(p);
        }));
    }
}

}  // namespace

void count_probs(const int16_t *MB, const int32_t *nz, const int32_t *parts, uint32_t *coeff_probs,
                 uint32_t *coeff_probs_denom, uint8_t *third_context, int mb_height, int mb_width, int P) {
    for_each_partition(P, [=](int p) {
        StatSink s{coeff_probs + (size_t)p * 1056, coeff_probs_denom + (size_t)p * 1056};
        for (int i = 0; i < 1056; ++i) {
            s.num[i] = 0;
            s.den[i] = 1;
        }
        // tail[type][i]: blocks of that type whose end-of-block token sat at position i-1, i.e. whose
        // context-2 end-of-block run starts at position i
        uint32_t tail[4][17];
        memset(tail, 0, sizeof(tail));
        for (int row = p; row < mb_height; row += P)
            for (int col = 0; col < mb_width; ++col) {
                const int mb = row * mb_width + col;
                if (nz[mb] == 0) continue;  // skipped macroblock: nothing is coded, contexts stay as they were
                uint8_t *ctx = third_context + (size_t)mb * 25;
                unsigned own[25];
                for (int b = 0; b < 25; ++b) own[b] = nonzero_mask(block(MB, mb, b));
                neighbour_contexts(MB, parts, mb, row, col, mb_width, own, ctx);
                int type = 3;
                if (parts[mb] == 0) {
                    const int e = walk_block(block(MB, mb, 24), 1, ctx[24], own[24], s);
                    ++tail[1][e < 16 ? e + 1 : 16];
                    type = 0;
                }
                for (int b = 0; b < 16; ++b) {
                    const int e = walk_block(block(MB, mb, b), type, ctx[b], own[b], s);
                    ++tail[type][e < 16 ? e + 1 : 16];
                }
                for (int b = 16; b < 24; ++b) {
                    const int e = walk_block(block(MB, mb, b), 2, ctx[b], own[b], s);
                    ++tail[2][e < 16 ? e + 1 : 16];
                }
            }
        // positions start..15 of each such block: one end-of-block decision with context 2 each
        for (int type = 0; type < 4; ++type) {
            uint32_t running = 0;
            for (int i = 1; i < 16; ++i) {
                running += tail[type][i];
                const int idx = ctx_index(type, kBand[i], 2, 0);
                s.num[idx] += running;
                s.den[idx] += running;
            }
        }
    });
}

void num_div_denom(uint32_t *coeff_probs, const uint32_t *coeff_probs_denom, int P) {
    for (int i = 0; i < 1056; ++i) {
        uint32_t num = 0, den = 0;
        for (int p = 0; p < P; ++p) {
            num += coeff_probs[(size_t)p * 1056 + i];
            den += coeff_probs_denom[(size_t)p * 1056 + i];
        }
        num = (num << 8) / den;
        coeff_probs[i] = num > 255 ? 255 : (num == 0 ? 1 : num);
    }
}

void encode_coefficients(const int16_t *MB, const int32_t *nz, const int32_t *parts, uint8_t *output,
                         int32_t *partition_sizes, const uint8_t *third_context, const uint32_t *coeff_probs,
                         int mb_height, int mb_width, int P, int partition_step) {
    for_each_partition(P, [=](int p) {
        BoolSink s;
        s.out = output + (size_t)partition_step * p;
        s.probs = coeff_probs;
        for (int row = p; row < mb_height; row += P)
            for (int col = 0; col < mb_width; ++col) {
                const int mb = row * mb_width + col;
                if (nz[mb] == 0) continue;
                const uint8_t *ctx = third_context + (size_t)mb * 25;
                int type = 3;
                if (parts[mb] == 0) {
                    walk_block(block(MB, mb, 24), 1, ctx[24], nonzero_mask(block(MB, mb, 24)), s);
                    type = 0;
                }
                for (int b = 0; b < 16; ++b) walk_block(block(MB, mb, b), type, ctx[b], nonzero_mask(block(MB, mb, b)), s);
                for (int b = 16; b < 24; ++b) walk_block(block(MB, mb, b), 2, ctx[b], nonzero_mask(block(MB, mb, b)), s);
            }
        s.finish();
        partition_sizes[p] = (int32_t)s.count;
    });
}

// The partitions from the decision streams the GPU prepared (entropy_kernels.cu): entry = bit 15 value,
// bits 0-10 probability slot (< 1056) or 1056 + fixed probability.  part_info = [P] stream bases, [P] counts.
void encode_token_streams(const uint16_t *tokens, const uint32_t *part_info, const uint32_t *coeff_probs, uint8_t *output,
                          int32_t *partition_sizes, int P, int partition_step) {
    // one table for both kinds of entry
    std::vector<uint8_t> table(1056 + 256);
    for (int i = 0; i < 1056; ++i) table[i] = (uint8_t)coeff_probs[i];
    for (int i = 0; i < 256; ++i) table[1056 + i] = (uint8_t)i;
    const uint8_t *tab = table.data();
    for_each_partition(P, [=](int p) {
        BoolSink s;
        s.out = output + (size_t)partition_step * p;
        s.probs = coeff_probs;
        const uint16_t *t = tokens + part_info[p];
        const uint32_t n = part_info[P + p];
        for (uint32_t i = 0; i < n; ++i) s.put(tab[t[i] & 0x7ff], t[i] >> 15);
        s.finish();
        partition_sizes[p] = (int32_t)s.count;
    });
}

}  // namespace vp8host

// C entry points (used by the host-logic tests; the shim calls the C++ functions directly)
extern "C" {
void vp8b200_host_count_probs(const int16_t *MB, const int32_t *nz, const int32_t *parts, uint32_t *coeff_probs,
                              uint32_t *coeff_probs_denom, uint8_t *third_context, int mb_height, int mb_width, int P) {
    vp8host::count_probs(MB, nz, parts, coeff_probs, coeff_probs_denom, third_context, mb_height, mb_width, P);
}
void vp8b200_host_encode_token_streams(const uint16_t *tokens, const uint32_t *part_info, const uint32_t *coeff_probs,
                                       uint8_t *output, int32_t *partition_sizes, int P, int partition_step) {
    vp8host::encode_token_streams(tokens, part_info, coeff_probs, output, partition_sizes, P, partition_step);
}
void vp8b200_host_num_div_denom(uint32_t *coeff_probs, const uint32_t *coeff_probs_denom, int P) {
    vp8host::num_div_denom(coeff_probs, coeff_probs_denom, P);
}
void vp8b200_host_encode_coefficients(const int16_t *MB, const int32_t *nz, const int32_t *parts, uint8_t *output,
                                      int32_t *partition_sizes, const uint8_t *third_context,
                                      const uint32_t *coeff_probs, int mb_height, int mb_width, int P,
                                      int partition_step) {
    vp8host::encode_coefficients(MB, nz, parts, output, partition_sizes, third_context, coeff_probs, mb_height, mb_width,
                                 P, partition_step);
}
}
