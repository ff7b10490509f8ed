// See entropy_host.h.  VP8 coefficient token coding (RFC 6386 section 13) written as one block
// walker with pluggable sinks: a statistics sink (count_probs) and a boolean-coder sink
// (encode_coefficients).
#include "entropy_host.h"

#include <cstring>
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

// Walks one block.  type: 0 = Y after Y2 (starts at coefficient 1), 1 = Y2, 2 = chroma,
// 3 = Y without Y2.  ctx = number of non-empty neighbour blocks (above, left).
// When keep_counting_after_eob is set the walk does what the reference's count_probs_in_block
// does (src/CPU_kernels.cl:503-538): it does NOT stop at the end-of-block token but goes on
// to visit an end-of-block decision for every remaining position, with context 2.
template <class Sink>
inline void walk_block(const int16_t *coef, int type, int ctx, bool keep_counting_after_eob, Sink &sink) {
    int last = 15;
    while (last >= 0 && coef[last] == 0) --last;  // positions > last are end-of-block tokens
    bool prev_zero = false;
    for (int i = (type == 0) ? 1 : 0; i < 16; ++i) {
        const int band = kBand[i];
        const int v = coef[i];
        const int mag = v < 0 ? -v : v;
        const Token t = (i > last) ? T_EOB : classify(mag);
        const Path &p = kPath[t];
        // after a ZERO token an end-of-block cannot follow, so the first decision is implicit
        for (int k = prev_zero ? 1 : 0; k < p.n; ++k) sink.decision(ctx_index(type, band, ctx, p.slot[k]), p.bit[k]);
        if (t == T_EOB) {
            if (!keep_counting_after_eob) return;
            ctx = 2;
            prev_zero = false;
            continue;
        }
        if (t >= T_CAT1) {
            const int c = t - T_CAT1, extra = mag - kCatBase[c];
            for (int b = 0; b < kCatBits[c]; ++b) sink.literal(kCatProb[c][b], (extra >> (kCatBits[c] - 1 - b)) & 1);
        }
        if (t == T_ZERO) {
            prev_zero = true;
            ctx = 0;
        } else {
            sink.literal(128, v < 0);  // sign
            prev_zero = false;
            ctx = (t == T_ONE) ? 1 : 2;
        }
    }
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
        if (bit) {
            bottom += split;
            range -= split;
        } else {
            range = split;
        }
        while (range < 128) {
            range <<= 1;
            if (bottom & (1u << 31)) carry(out);
            bottom <<= 1;
            if (!--bit_count) {
                *out++ = (uint8_t)(bottom >> 24);
                ++count;
                bottom &= (1u << 24) - 1;
                bit_count = 8;
            }
        }
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

inline const int16_t *block(const int16_t *MB, int mb, int b) { return MB + (size_t)mb * 400 + b * 16; }

inline int nonzero_from(const int16_t *c, int first) {
    for (int i = first; i < 16; ++i)
        if (c[i]) return 1;
    return 0;
}

// number of non-empty neighbour blocks (above + left) of every block of one macroblock
void neighbour_contexts(const int16_t *MB, const int32_t *parts, int mb, int mb_row, int mb_col, int mb_width,
                        uint8_t *ctx /* [25] */) {
    if (parts[mb] == 0) {
        // Y2: the neighbours are the nearest macroblocks above / to the left that have a Y2 block
        int n = 0;
        if (mb_row > 0) {
            int p = mb - mb_width;
            while (p >= 0 && parts[p] != 0) p -= mb_width;
            if (p >= 0) n += nonzero_from(block(MB, p, 24), 0);
        }
        if (mb_col > 0) {
            int p = mb - 1;
            while (p >= mb_row * mb_width && parts[p] != 0) --p;
            if (p >= mb_row * mb_width) n += nonzero_from(block(MB, p, 24), 0);
        }
        ctx[24] = (uint8_t)n;
    }
    for (int b = 0; b < 16; ++b) {
        int n = 0, pm = -1, pb = 0;
        if ((b >> 2) > 0) { pm = mb; pb = b - 4; } else if (mb_row > 0) { pm = mb - mb_width; pb = b + 12; }
        if (pm >= 0) n += nonzero_from(block(MB, pm, pb), parts[pm] == 0 ? 1 : 0);  // DC lives in Y2 there
        pm = -1;
        if ((b & 3) > 0) { pm = mb; pb = b - 1; } else if (mb_col > 0) { pm = mb - 1; pb = b + 3; }
        if (pm >= 0) n += nonzero_from(block(MB, pm, pb), parts[pm] == 0 ? 1 : 0);
        ctx[b] = (uint8_t)n;
    }
    for (int base = 16; base <= 20; base += 4)
        for (int k = 0; k < 4; ++k) {
            const int b = base + k;
            int n = 0, pm = -1, pb = 0;
            if ((k >> 1) > 0) { pm = mb; pb = b - 2; } else if (mb_row > 0) { pm = mb - mb_width; pb = b + 2; }
            if (pm >= 0) n += nonzero_from(block(MB, pm, pb), 0);
            pm = -1;
            if ((k & 1) > 0) { pm = mb; pb = b - 1; } else if (mb_col > 0) { pm = mb - 1; pb = b + 1; }
            if (pm >= 0) n += nonzero_from(block(MB, pm, pb), 0);
            ctx[b] = (uint8_t)n;
        }
}

template <class Sink>
inline void walk_macroblock(const int16_t *MB, const int32_t *parts, int mb, const uint8_t *ctx, bool counting, Sink &sink) {
    int type = 3;
    if (parts[mb] == 0) {
        walk_block(block(MB, mb, 24), 1, ctx[24], counting, sink);
        type = 0;
    }
    for (int b = 0; b < 16; ++b) walk_block(block(MB, mb, b), type, ctx[b], counting, sink);
    for (int b = 16; b < 24; ++b) walk_block(block(MB, mb, b), 2, ctx[b], counting, sink);
}

template <class F>
void for_each_partition(int n, F f) {
    if (n <= 1) {
        f(0);
        return;
    }
    std::vector<std::thread> th;
    th.reserve(n - 1);
    for (int p = 1; p < n; ++p) th.emplace_back(f, p);
    f(0);
    for (auto &t : th) t.join();
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
        for (int row = p; row < mb_height; row += P)
            for (int col = 0; col < mb_width; ++col) {
                const int mb = row * mb_width + col;
                if (nz[mb] == 0) continue;  // skipped macroblock: nothing is coded, contexts stay as they were
                uint8_t *ctx = third_context + (size_t)mb * 25;
                neighbour_contexts(MB, parts, mb, row, col, mb_width, ctx);
                walk_macroblock(MB, parts, mb, ctx, true, s);
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
                walk_macroblock(MB, parts, mb, third_context + (size_t)mb * 25, false, s);
            }
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
