// roc_small.cuh -- ROC decode and encode of SHORT units (graph rows, K <= 64 ids), one unit per THREAD.
//
// The lane-group decoder of roc_group.cuh is built for chains of up to 65 536 steps: a bucket structure in HBM, count
// levels in shared memory, a stream ring, one rendezvous per step -- about 560 warp instructions per step for the 8
// units of a warp. A graph row has at most 64 ids (altid_impl.h:53-67): "insert, then rank" (fenwick_tree.h:42-94) is a
// count over the ids decoded so far, which sit in shared memory, and the stream is ~30 words. One thread runs the
// reference's loop (codec.cpp:140-152) as it stands; a warp decodes 32 rows at ~10 warp instructions per row step.
//
// IDC_HD like idc_core.cuh: tests/hostsim runs these functions against the oracle on the CPU.
#pragma once

#include "idc_core.cuh"

namespace idc {

constexpr uint32_t kSmallUnit = 64;  // longest unit this decoder takes (shared memory: 4 bytes per id and thread)

struct SmallDec {
    uint64_t head;
    const uint32_t* words;  // the unit's stack in the blob, bottom first
    uint32_t sp;            // words of it not consumed yet
    uint32_t nxt;           // words[sp - 1], loaded when its predecessor was consumed (off the chain)
    uint32_t ov;            // the one word the decoder may hold above the blob's stack (see dec_push_uniform)
    uint32_t has_ov;
    uint32_t draws;         // words taken from the mt19937(1234) fallback (codec.h:32-40)
    uint32_t status;
};

IDC_HD uint32_t small_ld(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

IDC_HD void small_dec_init(SmallDec& s, uint64_t head, const uint32_t* words, uint32_t nwords) {
    s.head = head;
    s.words = words;
    s.sp = nwords;
    s.nxt = nwords ? small_ld(words + nwords - 1u) : 0u;
    s.ov = 0;
    s.has_ov = 0;
    s.draws = 0;
    s.status = 0;
}

// ANSState::stack_slice (codec.h:32-40): the top of the stack -- the overlay word, else the blob's next word, else the
// next output of std::mt19937(1234)
IDC_HD uint32_t small_refill(SmallDec& s, const uint32_t* mt) {
    if (s.has_ov) {
        s.has_ov = 0;
        return s.ov;
    }
    if (s.sp) {
        const uint32_t w = s.nxt;
        s.sp--;
        if (s.sp) s.nxt = small_ld(s.words + s.sp - 1u);
        return w;
    }
    const uint32_t d = s.draws++;
    if (d >= (uint32_t)kMtWords) {
        s.status |= kStMtDraws;
        return 0u;
    }
    return small_ld(mt + d);
}

// vrans_pop (codec.cpp:78-90), p in 0..16; p == 0 still runs the refill test
IDC_HD uint32_t small_pop_bits(SmallDec& s, uint32_t p, const uint32_t* mt) {
    uint64_t h = s.head;
    const uint32_t sym = (uint32_t)h & ((1u << p) - 1u);
    h >>= p;
    if (h < kRansL) h = (h << 32) | (uint64_t)small_refill(s, mt);
    s.head = h;
    return sym;
}

// codec_pop (codec.cpp:107-121) for precision <= 32: slices at lower = 48, 32 (precision 0: refill test only), 16, 0
IDC_HD uint32_t small_pop_id(SmallDec& s, int precision, const uint32_t* mt) {
    const uint32_t p0 = precision < 16 ? (uint32_t)precision : 16u;
    const uint32_t p1 = (uint32_t)precision - p0;
    (void)small_pop_bits(s, 0u, mt);
    (void)small_pop_bits(s, 0u, mt);
    const uint32_t hi = small_pop_bits(s, p1, mt);
    const uint32_t lo = small_pop_bits(s, p0, mt);
    return (hi << 16) | lo;
}

// push_with_finer_precision (codec.cpp:44-63); q31 = 2^31 / nmax
IDC_HD void small_push_uniform(SmallDec& s, uint32_t sym, uint32_t nmax, uint32_t q31, const uint32_t* mt) {
    uint64_t h = s.head;
    const uint32_t hi = (uint32_t)(h >> 32);
    if (hi >= q31) {  // h >= (2^31 / nmax) << 32: the low word goes on the stack
        if (s.has_ov) s.status |= kStOverlay;  // a second word above the blob's stack: not a stream of this codec
        s.ov = (uint32_t)h;
        s.has_ov = 1;
        h = (uint64_t)hi;
    }
    h = h * nmax + sym;
    if (h < kRansL) h = (uint64_t)small_refill(s, mt) | (h << 32);
    s.head = h;
}

// One unit, start to end (decompress, codec.cpp:140-152). seen(j) is a reference to the j-th id decoded (the caller's
// storage: a shared-memory column on the device); the output order is the reference's: data[n - 1 - i] = i-th decoded.
template <class Seen>
IDC_HD void small_dec_unit(SmallDec& s, uint32_t n, int precision, Seen&& seen, const uint32_t* mt) {
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t id = small_pop_id(s, precision, mt);
        uint32_t rank = 0;  // insert_then_forward_lookup(...).start: the ids below it (fenwick_tree.h:42-94)
        for (uint32_t j = 0; j < i; j++) rank += seen(j) < id ? 1u : 0u;
        seen(i) = id;
        small_push_uniform(s, rank, i + 1u, (uint32_t)(kRansL / (i + 1u)), mt);
    }
}

// ---- encode (compress, codec.cpp:123-138): the ids still in the set are a 64-bit mask over the ascending row --
// "select the k-th remaining id, remove it" (fenwick_tree.h:96-140) is a select on the mask.

// position of the r-th (0-based) set bit of a 64-bit mask with more than r ones
IDC_HD uint32_t small_select64(uint64_t m, uint32_t r) {
    const uint32_t lo = (uint32_t)m, c = (uint32_t)popc32(lo);
    return r < c ? select32(lo, r) : 32u + select32((uint32_t)(m >> 32), r - c);
}

// One encoder step: nmax ids are left (the caller walks nmax from n down to 1). ids(pos) is the pos-th id of the
// ascending row; returns pos (the caller records the sample order with it).
template <class Ids>
IDC_HD uint32_t small_enc_step(EncState& st, uint64_t& mask, uint32_t nmax, int precision, uint64_t rcp, uint32_t q31,
                               Ids&& ids, const uint32_t* mt) {
    const uint32_t k = enc_pop_uniform(st, nmax, rcp, q31, mt);
    const uint32_t pos = small_select64(mask, k);
    mask &= ~(1ull << pos);
    enc_push_id32(st, ids(pos), precision);
    return pos;
}

// ---- one unit per WARP: the coder runs in every lane (same inputs, same arithmetic), the order statistic is cooperative.
// GR is the warp as a lane group -- Grp<32> on the device (roc_group.cuh), 32 host threads in tests/hostsim: `sub` is the
// lane, ballot / shfl / sync are the warp collectives.

IDC_HD uint64_t small_ld64(const uint64_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

template <class GR>
IDC_HD uint32_t warp_sum(const GR& g, uint32_t v) {
#if defined(__CUDA_ARCH__)
    (void)g;
    return __reduce_add_sync(0xffffffffu, v);
#else
    for (uint32_t m = 16; m; m >>= 1) v += g.shfl_xor(v, m);
    return v;
#endif
}

// Decode of a unit of up to 2 048 ids (decompress, codec.cpp:140-152). seen[0 .. n): the warp's shared memory, the i-th
// decoded id at seen[i]; a rank is a strided count over it plus one warp reduction. q31[d] = 2^31 / d.
template <class GR>
IDC_HD void warp_dec_unit(const GR& g, SmallDec& st, uint32_t n, int prec, uint32_t* seen, const uint32_t* q31, const uint32_t* mt) {
    const uint32_t lane = g.sub;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t q = small_ld(q31 + i + 1u);  // requested before the step's chain starts
        const uint32_t id = small_pop_id(st, prec, mt);
        uint32_t cnt = 0;
        for (uint32_t j = lane; j < i; j += 32u) cnt += seen[j] < id ? 1u : 0u;
        const uint32_t rank = warp_sum(g, cnt);
        if (lane == 0) seen[i] = id;
        g.sync();
        small_push_uniform(st, rank, i + 1u, q, mt);
    }
}

// Decode of a graph row of up to 64 ids: lane j keeps the j-th and (j + 32)-th decoded id (s0, s1), a rank is two ballots.
template <class GR>
IDC_HD void warp_dec_row(const GR& g, SmallDec& st, uint32_t n, int prec, uint32_t& s0, uint32_t& s1, const uint32_t* mt) {
    const uint32_t lane = g.sub;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t id = small_pop_id(st, prec, mt);
        const uint32_t below0 = g.ballot(lane < i && s0 < id);
        const uint32_t below1 = g.ballot(lane + 32u < i && s1 < id);
        if (lane == (i & 31u)) {
            if (i < 32u) s0 = id; else s1 = id;
        }
        small_push_uniform(st, (uint32_t)(popc32(below0) + popc32(below1)), i + 1u, (uint32_t)(kRansL / (i + 1u)), mt);
    }
}

// Encode of a unit of up to 2 048 ascending ids sid[0 .. n) (compress, codec.cpp:123-138). Lane j keeps the presence
// masks of words j and j + 32 of the unit (m0, m1) and the number of ids still present in FRONT of those words (e0, e1) in
// registers. The front counts are non-decreasing over the words, so the words with front count <= k are a prefix and the
// last of them holds the k-th remaining id: two ballots, a shuffle of its count and mask, select32. A removal decrements
// the front counts of the words behind it. order(step, pos) records the sample order (pos = position in the ascending unit).
template <class GR, class Order>
IDC_HD void warp_enc_unit(const GR& g, EncState& st, uint32_t n, int prec, const uint32_t* sid, const uint64_t* rcp64,
                          const uint32_t* q31tab, Order&& order, const uint32_t* mt) {
    const uint32_t lane = g.sub;
    const uint32_t W = (n + 31u) >> 5;  // mask words in use (<= 64)
    auto word_mask = [&](uint32_t w) { return w * 32u + 32u <= n ? 0xffffffffu : (w * 32u < n ? (1u << (n - w * 32u)) - 1u : 0u); };
    uint32_t m0 = word_mask(lane), m1 = word_mask(lane + 32u);
    // a lane without a word never qualifies
    uint32_t e0 = lane < W ? lane * 32u : 0xffffffffu, e1 = lane + 32u < W ? (lane + 32u) * 32u : 0xffffffffu;
    uint64_t rcp = small_ld64(rcp64 + n);
    uint32_t q31 = small_ld(q31tab + n);
    for (uint32_t t = n; t >= 1u; --t) {
        const uint64_t rcp_n = small_ld64(rcp64 + (t - 1u));  // the next step's table entries, ahead of this step's chain
        const uint32_t q31_n = small_ld(q31tab + (t - 1u));
        const uint32_t k = enc_pop_uniform(st, t, rcp, q31, mt);
        const uint32_t word = (uint32_t)(popc32(g.ballot(e0 <= k)) + popc32(g.ballot(e1 <= k))) - 1u;
        const bool hi = word >= 32u;
        const uint32_t ew = g.shfl(hi ? e1 : e0, word & 31u);
        const uint32_t mw = g.shfl(hi ? m1 : m0, word & 31u);
        const uint32_t bit = select32(mw, k - ew);
        const uint32_t pos = word * 32u + bit;
        if (lane == (word & 31u)) {
            if (hi) m1 &= ~(1u << bit); else m0 &= ~(1u << bit);
        }
        e0 -= (lane > word && lane < W) ? 1u : 0u;
        e1 -= (lane + 32u > word && lane + 32u < W) ? 1u : 0u;
        enc_push_id32(st, sid[pos], prec);
        order(n - t, pos);
        g.sync();  // the stream words lane 0 stored are ordered before a later step's refill (enc_refill)
        rcp = rcp_n;
        q31 = q31_n;
    }
}

}  // namespace idc
