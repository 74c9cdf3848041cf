// wt_core.cuh -- wavelet structure over the sequence "list number of id i" (host + device).
//
// Replaces sdsl::wt_int<> as used by CompressedIDInvertedListsWaveletTree
// (custom_invlist_cpp/custom_invlists_impl.cpp:346-397): the reference stores S[id] = list_no for
// id in [0, ntotal) (:354-362) and answers get_single_id(list_no, offset) = wt.select(offset + 1, list_no)
// (:377-379). SDSL is a third-party dependency that is absent from the reference tree, so the word layout of
// its wt_int cannot be pinned; what the reference's tests pin is the value of every select
// (test_compressed_ivfs.py:37-41,128-132), and that is what this structure reproduces.
//
// Layout (chosen for the GPU, "wavelet matrix"): `levels` = bit_length(nlist - 1) bit vectors of n bits each.
// Level 0 looks at the most significant bit of the symbols in id order; the order of level v + 1 is the STABLE
// partition of level v's order by its bit (zeros first). A stable partition of a whole level is one streaming
// pass with a prefix sum -- no per-node bookkeeping as in a pointer-based or level-wise wavelet tree -- and the
// per-level directories are flat arrays:
//   bits   levels x words      LSB-first 64-bit words, 8 words (512 bits) per rank block
//   rank   levels x (nblk+1)   ones before block j; entry nblk = ones of the level
//   sel1   levels x samp       block that holds one  number m * 2048
//   sel0   levels x samp       block that holds zero number m * 2048
//   start  nlist               position of the list's first id below the last level (lists in bit-reversed order)
// select(k, c): p = start[c] + k, then for v = levels-1 .. 0: p = position of the p-th zero of level v if bit v
// of c is 0, of the (p - zeros_v)-th one otherwise. The answer is p at level 0 = the id.
#pragma once

#include "ef_core.cuh"

namespace idc {

constexpr uint32_t kWtBlockLog = 9;  // 512 bits per rank block
constexpr uint32_t kWtBlockBits = 1u << kWtBlockLog;
constexpr uint32_t kWtBlockWords = kWtBlockBits / 64;
constexpr uint32_t kWtSampleLog = 11;  // a select sample every 2048 ones / zeros (> block: <= 1 sample per block)
constexpr uint32_t kWtSample = 1u << kWtSampleLog;
constexpr uint32_t kWtHole = 0xffffffffu;  // "no list owns this id" while the sequence is being filled

struct WtShape {
    uint32_t levels;
    uint64_t n, nblk, words, rank_stride, samp_stride;
};

IDC_HD WtShape wt_shape(uint64_t nlist, uint64_t n) {
    WtShape s;
    s.levels = 1;
    while (s.levels < 32 && (1ull << s.levels) < nlist) s.levels++;
    s.n = n;
    s.nblk = (n + kWtBlockBits - 1) >> kWtBlockLog;
    s.words = s.nblk * kWtBlockWords;
    s.rank_stride = s.nblk + 1;
    s.samp_stride = (n >> kWtSampleLog) + 2;
    return s;
}

// bit-reversal of the low `levels` bits: lists appear below the last level in this order
IDC_HD uint32_t wt_bitrev(uint32_t c, uint32_t levels) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < levels; i++) r |= ((c >> i) & 1u) << (levels - 1 - i);
    return r;
}

// Directory entry of rank block j of a level, from the ones before block j (o0) and before block j + 1 (o1):
// a block holds at most 512 ones / zeros, fewer than the sample distance, hence at most one sample of each kind.
struct WtDirEntry {
    bool has1, has0;
    uint64_t m1, m0;  // sample numbers: the block holds one number m1 * 2048 / zero number m0 * 2048
};

IDC_HD WtDirEntry wt_dir_entry(uint64_t j, uint64_t nblk, uint64_t n, uint64_t o0, uint64_t o1) {
    WtDirEntry d;
    d.m1 = (o0 + kWtSample - 1) >> kWtSampleLog;
    d.has1 = (d.m1 << kWtSampleLog) < o1;
    uint64_t z0 = (j << kWtBlockLog) - o0;
    uint64_t z1 = (j + 1 == nblk ? n : (j + 1) << kWtBlockLog) - o1;
    d.m0 = (z0 + kWtSample - 1) >> kWtSampleLog;
    d.has0 = (d.m0 << kWtSampleLog) < z1;
    return d;
}

// destination of element i of a level in the next level's order (stable partition, zeros first)
IDC_HD uint64_t wt_partition_dest(uint64_t i, uint32_t bit, uint64_t zeros_of_level, uint64_t ones_before) {
    return bit ? zeros_of_level + ones_before : i - ones_before;
}

IDC_HD uint32_t wt_ld32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

IDC_HD uint64_t wt_ld64(const uint64_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const unsigned long long*>(p));
#else
    return *p;
#endif
}

struct WtView {
    const uint64_t* bits;
    const uint32_t* rank;
    const uint32_t* sel1;
    const uint32_t* sel0;
    const uint32_t* start;
    WtShape sh;
    // wt_type = 1: the bit vectors are block-compressed (bits == nullptr), see "RRR(63) blocks" below
    const uint64_t* cls = nullptr;       // levels x nblk: eight 6-bit classes + the block's 8 tail bits
    const uint32_t* ptr = nullptr;       // levels x (nblk + 1): bit offset of the block's first offset field in its level's stream
    const uint64_t* off = nullptr;       // the offset streams of all levels
    const uint64_t* off_base = nullptr;  // levels + 1: first word of each level's stream
    const uint64_t* binom = nullptr;     // C(n, k), 64 x 64, followed by the 64 field widths
};

// ---- RRR(63) blocks (wt_type = 1; sdsl::rrr_vector<63> in the reference, custom_invlists_impl.h:105) -------------
// A 63-bit block is stored as its class k (number of ones, 6 bits) and its offset among the C(63, k) blocks of that
// class (ceil(log2 C(63, k)) bits, nothing for k = 0 and k = 63): the combinatorial number system -- with the ones at
// positions c_1 < c_2 < ... < c_k the offset is sum C(c_i, i). A 512-bit rank block is eight such blocks plus its last
// 8 bits verbatim, so the rank directory and the select samples of the plain layout stay what they are, and one
// 64-bit word holds the eight classes and the tail. SDSL is absent from the reference tree: the framing is this
// repository's, what is reproduced is the value of every select.
constexpr uint32_t kRrrBits = 63;
constexpr uint32_t kRrrPerBlock = 8;
constexpr uint64_t kRrrMask = (1ull << kRrrBits) - 1ull;

IDC_HD uint64_t rrr_binom(const uint64_t* tab, uint32_t n, uint32_t k) {
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const unsigned long long*>(tab) + n * 64u + k);
#else
    return tab[n * 64u + k];
#endif
}
IDC_HD uint32_t rrr_width(const uint64_t* tab, uint32_t k) { return (uint32_t)rrr_binom(tab, 64u, k); }  // row 64 = the widths

// offset of a 63-bit block among the blocks with as many ones
IDC_HD uint64_t rrr_offset_of(const uint64_t* tab, uint64_t x) {
    uint64_t off = 0;
    uint32_t i = 0;
    while (x) {
        const uint32_t c = ctz64(x);
        x &= x - 1;
        off += rrr_binom(tab, c, ++i);
    }
    return off;
}

// the block of class k with that offset
IDC_HD uint64_t rrr_block_of(const uint64_t* tab, uint32_t k, uint64_t off) {
    uint64_t x = 0;
    int32_t c = (int32_t)kRrrBits - 1;
    for (uint32_t i = k; i >= 1; i--) {
        while (rrr_binom(tab, (uint32_t)c, i) > off) c--;  // the largest position whose C(c, i) fits
        x |= 1ull << c;
        off -= rrr_binom(tab, (uint32_t)c, i);
        c--;
    }
    return x;
}

// the 63 bits at bit 63 j of a 512-bit block held in eight words
IDC_HD uint64_t rrr_piece(const uint64_t (&w)[8], uint32_t j) {
    const uint32_t p = kRrrBits * j, a = p >> 6, s = p & 63u;
    uint64_t x = w[a] >> s;
    if (s > 1u) x |= w[a + 1] << (64u - s);  // s + 63 > 64; a + 1 <= 7 because 63 j + 63 <= 504
    return x & kRrrMask;
}

// W bits (<= 60) at bit p of a stream
IDC_HD uint64_t rrr_read(const uint64_t* stream, uint64_t p, uint32_t W) {
    if (W == 0) return 0;
    const uint64_t a = p >> 6;
    const uint32_t s = (uint32_t)(p & 63u);
    uint64_t x = wt_ld64(stream + a) >> s;
    if (s + W > 64u) x |= wt_ld64(stream + a + 1) << (64u - s);
    return x & ((1ull << W) - 1ull);
}

// position of the k-th (0-based) bit of value b in level lev; k < number of such bits. ~0 = corrupt directory.
IDC_HD uint64_t wt_level_select(const WtView& v, uint32_t lev, uint32_t b, uint64_t k) {
    const uint32_t* rank = v.rank + (uint64_t)lev * v.sh.rank_stride;
    const uint32_t* samp = (b ? v.sel1 : v.sel0) + (uint64_t)lev * v.sh.samp_stride;
    const uint64_t ones = wt_ld32(rank + v.sh.nblk);
    const uint64_t cnt = b ? ones : v.sh.n - ones;
    const uint64_t m = k >> kWtSampleLog;
    uint64_t lo = wt_ld32(samp + m);
    uint64_t hi = ((m + 1) << kWtSampleLog) < cnt ? (uint64_t)wt_ld32(samp + m + 1) : v.sh.nblk - 1;
    // the largest block j in [lo, hi] with (b-bits before block j) <= k holds the k-th one
    while (lo < hi) {
        uint64_t mid = (lo + hi + 1) >> 1;
        uint64_t r = wt_ld32(rank + mid);
        uint64_t before = b ? r : (mid << kWtBlockLog) - r;
        if (before <= k)
            lo = mid;
        else
            hi = mid - 1;
    }
    uint64_t r0 = wt_ld32(rank + lo);
    uint32_t r = (uint32_t)(k - (b ? r0 : (lo << kWtBlockLog) - r0));
    if (v.cls) {
        // block-compressed level: walk the block's eight classes, decode the one 63-bit block that holds the target
        const uint64_t cw = wt_ld64(v.cls + (uint64_t)lev * v.sh.nblk + lo);
        uint64_t p = wt_ld32(v.ptr + (uint64_t)lev * (v.sh.nblk + 1) + lo);
        const uint64_t* stream = v.off + wt_ld64(v.off_base + lev);
        for (uint32_t j = 0; j < kRrrPerBlock; j++) {
            const uint32_t kc = (uint32_t)(cw >> (6u * j)) & 63u, W = rrr_width(v.binom, kc);
            const uint32_t c = b ? kc : kRrrBits - kc;
            if (r < c) {
                uint64_t x = rrr_block_of(v.binom, kc, rrr_read(stream, p, W));
                if (!b) x = ~x & kRrrMask;
                return (lo << kWtBlockLog) + kRrrBits * j + select64(x, r);
            }
            r -= c;
            p += W;
        }
        uint64_t x = (cw >> 48) & 0xffull;
        if (!b) x = ~x & 0xffull;
        if (r < (uint32_t)popc64(x)) return (lo << kWtBlockLog) + kRrrBits * kRrrPerBlock + select64(x, r);
        return ~0ull;
    }
    const uint64_t* w = v.bits + (uint64_t)lev * v.sh.words + lo * kWtBlockWords;
    for (uint32_t i = 0; i < kWtBlockWords; i++) {
        uint64_t x = wt_ld64(w + i);
        if (!b) x = ~x;
        uint32_t c = (uint32_t)popc64(x);
        if (r < c) return (lo << kWtBlockLog) + i * 64 + select64(x, r);
        r -= c;
    }
    return ~0ull;
}

// sdsl wt_int::select(k + 1, c) as called at custom_invlists_impl.cpp:377-379: the id of element k of list c
IDC_HD uint64_t wt_select(const WtView& v, uint32_t c, uint64_t k) {
    uint64_t p = (uint64_t)wt_ld32(v.start + c) + k;
    for (int lev = (int)v.sh.levels - 1; lev >= 0; lev--) {
        uint32_t b = (c >> (v.sh.levels - 1 - (uint32_t)lev)) & 1u;
        if (b) p -= v.sh.n - wt_ld32(v.rank + (uint64_t)lev * v.sh.rank_stride + v.sh.nblk);
        p = wt_level_select(v, (uint32_t)lev, b, p);
        if (p == ~0ull) return p;
    }
    return p;
}

// S[i] (sdsl wt_int::operator[]): walk down with rank. Used by the tests to check the structure both ways.
IDC_HD uint32_t wt_access(const WtView& v, uint64_t i) {
    uint32_t c = 0;
    uint64_t p = i;
    for (uint32_t lev = 0; lev < v.sh.levels; lev++) {
        const uint32_t* rank = v.rank + (uint64_t)lev * v.sh.rank_stride;
        const uint64_t* w = v.bits + (uint64_t)lev * v.sh.words;
        uint64_t blk = p >> kWtBlockLog;
        uint64_t ones = wt_ld32(rank + blk);
        uint32_t in = (uint32_t)(p & (kWtBlockBits - 1));
        for (uint32_t j = 0; j < (in >> 6); j++) ones += (uint32_t)popc64(wt_ld64(w + blk * kWtBlockWords + j));
        uint64_t x = wt_ld64(w + blk * kWtBlockWords + (in >> 6));
        ones += (uint32_t)popc64(x & ((1ull << (in & 63)) - 1ull));
        uint32_t b = (uint32_t)(x >> (in & 63)) & 1u;
        c = (c << 1) | b;
        uint64_t z = v.sh.n - wt_ld32(rank + v.sh.nblk);
        p = b ? z + ones : p - ones;
    }
    return c;
}

}  // namespace idc
