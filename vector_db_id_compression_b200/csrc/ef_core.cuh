// ef_core.cuh -- Elias-Fano word-level building blocks (host + device).
//
// Layout restated from the reference's elias_fano.hpp (modified ot/succinct):
//   l          = msb(universe / m) if m && universe / m else 0     (:28)
//   low bits   : element i's low l bits at bit offset i*l           (:40-42)
//   high bits  : (m + 1) + (universe >> l) + 1 bits, bit (v>>l)+i set (:29,:43)
// both as LSB-first 64-bit words (succinct bit_vector convention). `universe`
// is the value the reference passes as n: the list's max id
// (custom_invlists_impl.cpp:262-263, altid_impl.cpp:75-77).
//
// Encoding is formulated as a GATHER: every output word is produced by exactly
// one thread from the (ascending) ids it covers -- no atomics, no pre-zeroing,
// deterministic, and every store is a full coalesced 8-byte word.
#pragma once

#include "idc_core.cuh"
#include "idc_core.cuh"

namespace idc {

constexpr uint32_t kEfSampleLog = 8;  // a select sample every 256 ones (darray1 stand-in)
constexpr uint32_t kEfSample = 1u << kEfSampleLog;

IDC_HD uint32_t msb64(uint64_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
    return 63u - (uint32_t)__clzll((long long)x);
#else
    return 63u - (uint32_t)__builtin_clzll(x);
#endif
}

IDC_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

IDC_HD uint32_t ctz64(uint64_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffsll((long long)x) - 1u;
#else
    return (uint32_t)__builtin_ctzll(x);
#endif
}

struct EfShape {
    uint32_t l;
    uint64_t low_bits, high_bits, low_words, high_words, samples;
};

IDC_HD EfShape ef_shape(uint64_t universe, uint64_t m) {
    EfShape s;
    s.l = (m && universe / m) ? msb64(universe / m) : 0u;
    s.low_bits = m * s.l;
    s.high_bits = m ? (m + 1) + (universe >> s.l) + 1 : 0;  // empty lists own no bits (null elias_fano* in the reference)
    s.low_words = (s.low_bits + 63) / 64;
    s.high_words = (s.high_bits + 63) / 64;
    s.samples = (m + kEfSample - 1) >> kEfSampleLog;
    return s;
}

// position of element i's one in the high bit vector
template <typename IdT>
IDC_HD uint64_t ef_high_pos(const IdT* ids, uint64_t i, uint32_t l) {
    return (load_id(ids + i) >> l) + i;
}

// low-bits word w of a list: bits [64w, 64w+64) of the concatenated l-bit fields
template <typename IdT>
IDC_HD uint64_t ef_low_word(const IdT* ids, uint64_t m, uint32_t l, uint64_t w) {
    uint64_t bit0 = w * 64;
    uint64_t e = bit0 / l;
    uint64_t mask = (1ull << l) - 1ull;  // l < 64
    uint64_t out = 0;
    // first field may start before bit0
    int64_t shift = (int64_t)(e * l) - (int64_t)bit0;  // in (-l, 0]
    while (e < m && shift < 64) {
        uint64_t f = load_id(ids + e) & mask;
        out |= shift >= 0 ? (f << shift) : (f >> (-shift));
        shift += l;
        e++;
    }
    return out;
}

// high-bits word w of a list
template <typename IdT>
IDC_HD uint64_t ef_high_word(const IdT* ids, uint64_t m, uint32_t l, uint64_t universe, uint64_t w) {
    uint64_t p0 = w * 64;
    // first i with (ids[i] >> l) + i >= p0 ; hp is strictly increasing in i and i <= hp(i) <= i + (universe >> l)
    uint64_t span = universe >> l;
    uint64_t lo = p0 > span ? p0 - span : 0;
    uint64_t hi = p0 < m ? p0 : m;
    if (lo > hi) lo = hi;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (ef_high_pos(ids, mid, l) < p0)
            lo = mid + 1;
        else
            hi = mid;
    }
    uint64_t out = 0;
    for (uint64_t i = lo; i < m; i++) {
        uint64_t hp = ef_high_pos(ids, i, l);
        if (hp >= p0 + 64) break;
        out |= 1ull << (hp - p0);
    }
    return out;
}

// l-bit field i of the low bits (elias_fano.hpp:143 get_bits)
IDC_HD uint64_t ef_get_low(const uint64_t* low, uint64_t i, uint32_t l) {
    if (l == 0) return 0;
    uint64_t pos = i * l, sh = pos & 63;
#if defined(__CUDA_ARCH__)
    uint64_t w = __ldg(reinterpret_cast<const unsigned long long*>(low) + (pos >> 6)) >> sh;
    if (sh + l > 64) w |= __ldg(reinterpret_cast<const unsigned long long*>(low) + (pos >> 6) + 1) << (64 - sh);
#else
    uint64_t w = low[pos >> 6] >> sh;
    if (sh + l > 64) w |= low[(pos >> 6) + 1] << (64 - sh);
#endif
    return w & ((1ull << l) - 1ull);
}

// position of the r-th (0-based) set bit of a 64-bit word with more than r ones
IDC_HD uint32_t select64(uint64_t x, uint32_t r) {
    uint32_t pos = 0;
    uint32_t c = (uint32_t)popc64(x & 0xffffffffull);
    if (r >= c) { r -= c; pos += 32; x >>= 32; }
    c = (uint32_t)popc64(x & 0xffffull);
    if (r >= c) { r -= c; pos += 16; x >>= 16; }
    c = (uint32_t)popc64(x & 0xffull);
    if (r >= c) { r -= c; pos += 8; x >>= 8; }
    c = (uint32_t)popc64(x & 0xfull);
    if (r >= c) { r -= c; pos += 4; x >>= 4; }
    c = (uint32_t)popc64(x & 0x3ull);
    if (r >= c) { r -= c; pos += 2; x >>= 2; }
    if (r >= (uint32_t)(x & 1ull)) pos += 1;
    return pos;
}

// elias_fano::select(k) (elias_fano.hpp:141-145) using the select samples:
// sample j holds the high-bit position of one number j*256.
IDC_HD uint64_t ef_select(const uint64_t* low, const uint64_t* high, const uint32_t* samples, uint32_t l, uint64_t k) {
    uint64_t pos = samples ? samples[k >> kEfSampleLog] : 0;
    uint32_t r = samples ? (uint32_t)(k & (kEfSample - 1)) : (uint32_t)k;
    uint64_t w = pos >> 6;
#if defined(__CUDA_ARCH__)
    uint64_t bits = __ldg(reinterpret_cast<const unsigned long long*>(high) + w);
#else
    uint64_t bits = high[w];
#endif
    bits &= ~0ull << (pos & 63);
    for (;;) {
        uint32_t c = (uint32_t)popc64(bits);
        if (r < c) break;
        r -= c;
        w++;
#if defined(__CUDA_ARCH__)
        bits = __ldg(reinterpret_cast<const unsigned long long*>(high) + w);
#else
        bits = high[w];
#endif
    }
    uint64_t h = w * 64 + select64(bits, r);
    return ((h - k) << l) | ef_get_low(low, k, l);
}

}  // namespace idc
