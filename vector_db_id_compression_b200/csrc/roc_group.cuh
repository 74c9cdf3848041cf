// roc_group.cuh -- one ROC unit per GROUP of G lanes (G = 2, 4 or 8): the per-step encode / decode bodies.
//
// Why groups (profiles/README.md, r1_lat_bench_b200.txt): with one unit per lane every step makes each lane
// fetch its own 128-byte line; 32 different lines per warp instruction cost 990 ns instead of 465 ns, the
// memory system accepts only ~9.6 G such lane-granular line reads per second, and the 560 instructions of a
// step are issued by one warp for a handful of resident warps per SM. A group of G lanes
//   * fetches the unit's line as ONE coalesced request (lane j loads bytes [128 j / G, 128 (j+1) / G)),
//   * searches the 16-entry count levels with one compare per lane and a ballot,
//   * runs the rANS arithmetic replicated in every lane (same inputs, same result; only lane 0 stores).
// The stream is still the reference's single serial rANS head per unit (codec.cpp:131-137,144-151): the lanes of
// a group cooperate on the order-statistic structure, never on the coder state, so the output is bit-exact.
//
// The functions are templates over a group type GR providing sub / ballot / shfl / shfl_xor / sync. On the
// device that is Grp<G> below (sub-warp __ballot_sync / __shfl_sync with the group's lane mask); tests/hostsim
// supplies an emulation whose lanes are host threads, so this very code is checked against the oracle on the CPU.
//
//   encode step  = custom_invlists_impl.cpp:178-192 / codec.cpp:131-137
//   decode step  = codec.cpp:144-151 / altid_impl.cpp:157-163
#pragma once

#include "idc_core.cuh"

namespace idc {

#if defined(__CUDACC__)
// Collectives are issued with the FULL warp mask by all 32 lanes in lock step (the groups of a warp always walk
// the same instruction stream; a group without work runs the step with its memory operations switched off).
// Sub-warp member masks would be legal, but a per-group mask makes the hardware split the warp into
// independently scheduled groups that never reconverge: measured 8x the issue slots per step.
template <int G>
struct Grp {
    uint32_t sub, base;
    __device__ __forceinline__ Grp() {
        uint32_t lane = threadIdx.x & 31u;
        sub = lane & (uint32_t)(G - 1);
        base = lane - sub;
    }
    __device__ __forceinline__ uint32_t ballot(bool p) const {
        return (__ballot_sync(0xffffffffu, p) >> base) & (G == 32 ? 0xffffffffu : (1u << G) - 1u);
    }
    __device__ __forceinline__ uint32_t shfl(uint32_t v, uint32_t src) const {
        return __shfl_sync(0xffffffffu, v, base + (src & (uint32_t)(G - 1)));
    }
    __device__ __forceinline__ uint32_t shfl_xor(uint32_t v, uint32_t m) const { return __shfl_xor_sync(0xffffffffu, v, m); }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ bool warp_any(bool p) const { return __any_sync(0xffffffffu, p); }  // all groups of the warp
    __device__ __forceinline__ void host_sync() const { __syncwarp(); }  // off the critical path where it is used
};
#endif

// A unit's shared-memory region. Logical word w of group q lives in 16-byte chunks interleaved over the NG groups
// of a warp: chunk (w >> 2) of group q is chunk ((w >> 2) * NG + q) of the warp's region. Lanes of different
// groups then never share a bank, and a lane's 8- / 16-byte vector load stays inside one chunk.
struct SmView {
    uint32_t* base;
    uint32_t ng, q;
    IDC_HD uint32_t* at(uint32_t w) const { return base + (((((w >> 2) * ng) + q) << 2) + (w & 3u)); }  // 32-bit index math: a warp's region is < 64 KB
};

template <int N>
IDC_HD void sm_load(const uint32_t* p, uint32_t (&v)[N]) {
#if defined(__CUDA_ARCH__)
    if (N == 4) {
        uint4 x = *reinterpret_cast<const uint4*>(p);
        v[0] = x.x, v[1 % N] = x.y, v[2 % N] = x.z, v[3 % N] = x.w;
    } else if (N == 2) {
        uint2 x = *reinterpret_cast<const uint2*>(p);
        v[0] = x.x, v[1 % N] = x.y;
    } else {
        v[0] = *p;
    }
#else
    for (int j = 0; j < N; j++) v[j] = p[j];
#endif
}
template <int N>
IDC_HD void sm_store(uint32_t* p, const uint32_t (&v)[N]) {
#if defined(__CUDA_ARCH__)
    if (N == 4)
        *reinterpret_cast<uint4*>(p) = make_uint4(v[0], v[1 % N], v[2 % N], v[3 % N]);
    else if (N == 2)
        *reinterpret_cast<uint2*>(p) = make_uint2(v[0], v[1 % N]);
    else
        *p = v[0];
#else
    for (int j = 0; j < N; j++) p[j] = v[j];
#endif
}

template <class GR>
IDC_HD uint32_t group_sum(const GR& g, uint32_t v, int G) {
    for (int m = G >> 1; m; m >>= 1) v += g.shfl_xor(v, (uint32_t)m);
    return v;
}

// this lane's slice of a 128-byte line: words [sub * W, (sub + 1) * W), W = 32 / G = 4, 8 or 16
template <int W>
IDC_HD void ld_line_slice(const uint32_t* line, uint32_t sub, uint32_t (&v)[W]) {
    if (W == 4) {
        uint4x s = ld_ws16(line + sub * 4u);
        v[0] = s.x, v[1 % W] = s.y, v[2 % W] = s.z, v[3 % W] = s.w;
    } else {
#pragma unroll
        for (int h = 0; h < W / 8; h++) {
            Sector s = ld_sector(reinterpret_cast<const uint16_t*>(line), sub * (uint32_t)(W / 8) + (uint32_t)h);
#pragma unroll
            for (int j = 0; j < 8; j++) v[(8 * h + j) % W] = s.w[j];
        }
    }
}

// v[idx % W] as a select tree (dynamic indexing of a register array would turn into divergent branches)
template <int W>
IDC_HD uint32_t pick_word(const uint32_t (&v)[W], uint32_t idx) {
    uint32_t a[W];
#pragma unroll
    for (int j = 0; j < W; j++) a[j] = v[j];
#pragma unroll
    for (int span = W / 2, bit = 0; span >= 1; span >>= 1, bit++) {
        const bool b = (idx >> bit) & 1u;
#pragma unroll
        for (int j = 0; j < span; j++) a[j] = b ? a[2 * j + 1] : a[2 * j];
    }
    return a[0];
}

// later memory operations are not scheduled before this point (keeps a long-latency load the first thing issued)
IDC_HD void issue_fence() {
#if defined(__CUDA_ARCH__)
    asm volatile("" ::: "memory");
#endif
}

// fire-and-forget add on a shared-memory word owned by the calling lane
IDC_HD void sm_add(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);  // result unused -> ATOMS/RED without a return value: no load latency on the issue path
#else
    *p += v;
#endif
}

// ============================================================================================ encoder
// Select-k-th-and-remove over the unit's id-sorted array (FenwickTree::reverse_lookup_then_remove,
// fenwick_tree.h:96-140):
//   global     128-byte records: word 0 presence mask, words 1..31 thirty-one consecutive ids
//   registers  A  sixteen 32-bit exclusive cumulative counts over superblocks of 192 records, 16/G per lane
//   shared     B  per superblock sixteen 16-bit exclusive cumulative counts over blocks of 12   (8 words each)
//              C  per block two words = twelve 5-bit counts of ids still present per record
// A and B are searched by the group (each lane owns 16/G consecutive entries), C by a carry-free multiply.
// Cumulative entries make a search one compare per entry; the decrements after a removal are independent updates
// of a lane's own entries, issued while the record line is in flight.

IDC_HD uint32_t genc_sm_words(uint32_t n) {
    EncTreeLayout L = enc_tree_layout(n);
    return (8u * L.supers + L.l0_words + 3u) & ~3u;
}

// initial value of logical shared-memory word w for a full unit of n ids
IDC_HD uint32_t genc_sm_init_word(uint32_t n, const EncTreeLayout& L, uint32_t w) {
    const uint32_t l0 = 8u * L.supers;
    if (w < l0) {
        uint32_t sb = w >> 3, j = w & 7u;
        uint32_t base = enc_full_before(n, sb * kRecPerSuper);
        uint32_t lo = enc_full_before(n, sb * kRecPerSuper + (2u * j) * kRecPerPair) - base;
        uint32_t hi = enc_full_before(n, sb * kRecPerSuper + (2u * j + 1u) * kRecPerPair) - base;
        return lo | (hi << 16);
    }
    uint32_t word = 0;
    for (uint32_t q = 0; q < kRecPerWord; q++) {
        uint32_t r = (w - l0) * kRecPerWord + q;
        word |= (enc_full_before(n, r + 1u) - enc_full_before(n, r)) << (5u * q);
    }
    return word;
}

// Levels A and B can be searched by the group (each lane 16/G entries, a ballot and two shuffles per level) or by
// every lane on its own (all 16 entries in registers / loaded by every lane: more instructions, no collectives).
// Measured (profiles/README.md): the 2-lane classes gain 8 % from the local search (equal-length control: encode
// 66.7 -> 61.3 ms), the 4-lane classes lose 5 % (Zipf: 103.9 -> 108.7 ms) -- so it follows the lane count.
template <int G>
struct EncSearch {
    static constexpr bool kLocal = G == 2;            // levels A and B by every lane on its own
#if defined(IDC_ENC_LOCAL_A)
    static constexpr bool kLocalA = true;             // experiment: level A alone (registers only) local for every width
#else
    static constexpr bool kLocalA = kLocal;
#endif
    static constexpr int kEa = kLocalA ? 16 : 16 / G;  // level-A entries held by a lane
};

template <int G>
struct GEncTree {
    uint32_t* rec;        // global: records of 32 words
    SmView sm;
    uint32_t sm_l0;       // logical word offset of level C
    // level A: all 16 entries (local search) or this lane's entries [sub * 16/G, (sub + 1) * 16/G)
    uint32_t ea[EncSearch<G>::kEa];
};

template <int G, class GR>
IDC_HD void genc_tree_init(const GR& g, GEncTree<G>& t, uint32_t n) {
    EncTreeLayout L = enc_tree_layout(n);
    t.sm_l0 = 8u * L.supers;
    const uint32_t total = t.sm_l0 + L.l0_words;
    for (uint32_t w = g.sub; w < total; w += (uint32_t)G) *t.sm.at(w) = genc_sm_init_word(n, L, w);
    if (EncSearch<G>::kLocalA) {
#pragma unroll
        for (int j = 0; j < 16; j++) t.ea[j % EncSearch<G>::kEa] = enc_full_before(n, (uint32_t)j * kRecPerSuper);
    } else {
#pragma unroll
        for (int j = 0; j < 16 / G; j++) t.ea[j] = enc_full_before(n, (g.sub * (16 / G) + (uint32_t)j) * kRecPerSuper);
    }
}

// 16 non-decreasing exclusive cumulative entries, all held by the calling lane (entry 0 is 0): index of the last
// entry <= k and its value
IDC_HD uint32_t local_cum_find(const uint32_t (&e)[16], uint32_t k, uint32_t& base) {
    uint32_t c = 0;
#pragma unroll
    for (int j = 1; j < 16; j++) c += e[j] <= k ? 1u : 0u;
    base = pick_word<16>(e, c);
    return c;
}

// Six 5-bit counts c0..c5 of one C word -> inclusive prefix sums in 10-bit fields:
//   E = [c0, c0+c1+c2, c0+..+c4]  (records 0, 2, 4)      O = [c0+c1, c0+..+c3, c0+..+c5]  (records 1, 3, 5)
IDC_HD void c6_prefix(uint32_t w, uint32_t& E, uint32_t& O) {
    const uint32_t M = 0x01F07C1Fu, K = 0x00100401u;
    uint32_t pa = (w & M) * K, pb = ((w >> 5) & M) * K;  // fields 0..2 hold running sums of the even / odd counts
    E = (pa + (pb << 10)) & 0x3FFFFFFFu;
    O = (pa + pb) & 0x3FFFFFFFu;
}
// number of the word's records whose inclusive prefix is <= k (0..6); needs k <= 511
IDC_HD uint32_t c6_count_le(uint32_t E, uint32_t O, uint32_t k) {
    const uint32_t add = (511u - k) * 0x00100401u, top = 0x20080200u;  // field + 511 - k has bit 9 set <=> field > k
    return 6u - (uint32_t)popc32((E + add) & top) - (uint32_t)popc32((O + add) & top);
}
// exclusive prefix of record r (0..5)
IDC_HD uint32_t c6_excl(uint32_t E, uint32_t O, uint32_t r) {
    uint32_t src = (r & 1u) ? E : O;
    uint32_t f = (r & 1u) ? (r - 1u) >> 1 : (r - 2u) >> 1;
    return r ? (src >> (10u * f)) & 0x3ffu : 0u;
}

// Search 16 non-decreasing exclusive cumulative entries spread over the group (lane j holds entries
// [j*EL, (j+1)*EL) in e[]): returns the index of the last entry <= k and that entry's value in `base`.
template <int EL, class GR>
IDC_HD uint32_t group_cum_find(const GR& g, const uint32_t (&e)[EL], uint32_t k, uint32_t& base) {
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < EL; j++) c += e[j] <= k ? 1u : 0u;
    uint32_t cand = e[0];
#pragma unroll
    for (int j = 1; j < EL; j++) cand = (c > (uint32_t)j) ? e[j] : cand;  // e[c - 1]
    // entries are non-decreasing and entry 0 is 0: the lanes with c >= 1 form a prefix that includes lane 0
    const uint32_t bl = (uint32_t)popc32(g.ballot(c >= 1u)) - 1u;
    const uint32_t cb = g.shfl(c, bl);
    base = g.shfl(cand, bl);
    return bl * (uint32_t)EL + cb - 1u;
}

// Select the k-th remaining id (0-based), remove it; returns its position in the sorted unit and the id.
// `act` false: the group has no work this step; it takes part in the collectives with dummy values and touches
// no memory.
template <int G, class GR>
IDC_HD uint32_t genc_select_remove(const GR& g, GEncTree<G>& t, uint32_t k, uint32_t& id_out, bool act) {
    constexpr bool kEncLocalSearch = EncSearch<G>::kLocal;
    constexpr int EA = 16 / G;   // level-A entries per lane (group search)
    constexpr int WB = 8 / G;    // level-B words per lane (two entries each)
    constexpr int WL = 32 / G;   // record words per lane
    uint32_t base, sb, p;
    uint32_t wb[kEncLocalSearch ? 8 : WB];
    uint32_t* pb;
    constexpr bool kLocalA = EncSearch<G>::kLocalA;
    if (kLocalA) {
        // ---- level A (registers, every lane all 16 entries)
        uint32_t ea16[16];
#pragma unroll
        for (int j = 0; j < 16; j++) ea16[j] = t.ea[j % EncSearch<G>::kEa];
        sb = local_cum_find(ea16, k, base);
        k -= base;
    } else {
        // ---- level A (registers, 16/G entries per lane)
        uint32_t eal[EA];
#pragma unroll
        for (int j = 0; j < EA; j++) eal[j] = t.ea[j % EncSearch<G>::kEa];
        sb = group_cum_find<EA>(g, eal, k, base);
        k -= base;
    }
    if (kEncLocalSearch) {
        // ---- level B (every lane loads the superblock's 8 words)
#pragma unroll
        for (int j = 0; j < 8; j++) wb[j % (kEncLocalSearch ? 8 : WB)] = 0u;
        pb = t.sm.at(8u * sb);
        if (act) {
            uint32_t h0[4], h1[4];
            sm_load<4>(pb, h0);
            sm_load<4>(t.sm.at(8u * sb + 4u), h1);
#pragma unroll
            for (int j = 0; j < 4; j++) wb[j % (kEncLocalSearch ? 8 : WB)] = h0[j], wb[(4 + j) % (kEncLocalSearch ? 8 : WB)] = h1[j];
        }
        uint32_t eb16[16];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            eb16[2 * j] = wb[j % (kEncLocalSearch ? 8 : WB)] & 0xffffu;
            eb16[2 * j + 1] = wb[j % (kEncLocalSearch ? 8 : WB)] >> 16;
        }
        p = local_cum_find(eb16, k, base);
        k -= base;
    } else {
        // ---- level B
        uint32_t eb[2 * WB];
#pragma unroll
        for (int j = 0; j < WB; j++) wb[j] = 0u;
        pb = t.sm.at(8u * sb + g.sub * WB);
        uint32_t wl[WB];
        if (act) {
            sm_load<WB>(pb, wl);
#pragma unroll
            for (int j = 0; j < WB; j++) wb[j] = wl[j];
        }
#pragma unroll
        for (int j = 0; j < WB; j++) eb[2 * j] = wb[j] & 0xffffu, eb[2 * j + 1] = wb[j] >> 16;
        p = group_cum_find<2 * WB>(g, eb, k, base);
        k -= base;
    }
    // ---- level C: twelve 5-bit counts, replicated in every lane
    const uint32_t l0 = t.sm_l0 + (sb * 16u + p) * 2u;
    uint32_t wc[2] = {0u, 0u};
    uint32_t* pc = t.sm.at(l0);
    if (act) sm_load<2>(pc, wc);
    uint32_t Ea, Oa, Eb, Ob;
    c6_prefix(wc[0], Ea, Oa);
    c6_prefix(wc[1], Eb, Ob);
    const uint32_t tot_a = Oa >> 20;
    const uint32_t ra = c6_count_le(Ea, Oa, k);
    uint32_t rb = c6_count_le(Eb, Ob, k - tot_a);  // only meaningful when k >= tot_a
    rb = rb < 5u ? rb : 5u;
    const bool second = ra >= 6u;
    const uint32_t r = second ? 6u + rb : ra;
    k -= second ? tot_a + c6_excl(Eb, Ob, rb) : c6_excl(Ea, Oa, ra);
    // ---- the record line: one coalesced 128-byte request per group
    const uint32_t rg = (sb * 16u + p) * kRecPerPair + r;
    uint32_t* rec = t.rec + (size_t)rg * 32u;
    uint32_t lw[WL];
#pragma unroll
    for (int j = 0; j < WL; j++) lw[j] = 0u;
    if (act) ld_line_slice<WL>(rec, g.sub, lw);
    issue_fence();
    // ---- the count updates ride in the shadow of the line fetch; every lane updates its own entries only
    g.host_sync();  // all lanes have read the C words before lane 0 rewrites one (racecheck: warp-level WAR)
    if (act) {
        if (kLocalA) {
#pragma unroll
            for (int j = 1; j < 16; j++) t.ea[j % EncSearch<G>::kEa] -= ((uint32_t)j > sb) ? 1u : 0u;
        } else {
#pragma unroll
            for (int j = 0; j < EA; j++) t.ea[j % EncSearch<G>::kEa] -= (g.sub * EA + (uint32_t)j > sb) ? 1u : 0u;
        }
        if (kEncLocalSearch) {
            if (g.sub == 0) {  // one lane rewrites the superblock's 8 words
                uint32_t h0[4], h1[4];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t i0 = (uint32_t)j * 2u;
                    const uint32_t w = wb[j % (kEncLocalSearch ? 8 : WB)] - ((i0 > p ? 1u : 0u) | (i0 + 1u > p ? 0x10000u : 0u));
                    if (j < 4) h0[j % 4] = w; else h1[j % 4] = w;
                }
                sm_store<4>(pb, h0);
                sm_store<4>(t.sm.at(8u * sb + 4u), h1);
            }
        } else {
            uint32_t wl[WB];
#pragma unroll
            for (int j = 0; j < WB; j++) {
                uint32_t i0 = (g.sub * WB + (uint32_t)j) * 2u;
                wl[j] = wb[j] - ((i0 > p ? 1u : 0u) | (i0 + 1u > p ? 0x10000u : 0u));
            }
            sm_store<WB>(pb, wl);
        }
        if (g.sub == 0) {
            const uint32_t wsel = second ? wc[1] : wc[0];
            *(second ? pc + 1 : pc) = wsel - (1u << (5u * (second ? rb : ra)));
        }
    }
    // ---- consume the line
    const uint32_t mask = g.shfl(lw[0], 0u);
    const uint32_t pos = select32(mask, k);
    const uint32_t idx = pos + 1u;  // word of the record that holds the id
    const uint32_t mine = pick_word<WL>(lw, idx);
    id_out = g.shfl(mine, idx / (uint32_t)WL);
    if (act && g.sub == 0) st_ws32(rec, mask & ~(1u << pos));
    return rg * kRecIds + pos;
}

template <int G, typename IdT>
struct GEncUnit {
    EncState st;
    GEncTree<G> tree;
    const uint32_t* sort_idx;  // original position of each sorted id inside its list (null: input was sorted)
    uint32_t* order;           // sample order out (null: not wanted)
    uint32_t pos_base;         // position of the unit's first id inside its list
    uint32_t n;
    int prec;
};

// nmax = ids still in the set (the caller walks it from n down to 1)
template <int G, class GR, typename IdT>
IDC_HD void genc_step(const GR& g, GEncUnit<G, IdT>& U, uint32_t nmax, uint64_t rcp, uint32_t q31, const uint32_t* mt,
                      bool act) {
    uint32_t k = 0;
    if (act) k = enc_pop_uniform(U.st, nmax, rcp, q31, mt);
    uint32_t id32;
    uint32_t pos = genc_select_remove<G>(g, U.tree, k, id32, act);
    if (act) {
        enc_push_id32(U.st, id32, U.prec);
        if (U.order && g.sub == 0) {
            uint32_t o = U.sort_idx ? U.sort_idx[pos] : U.pos_base + pos;
            U.order[U.n - nmax] = o;
        }
    }
    g.sync();  // this step's stores (stream words, mask, counts) are ordered before the next step's loads
}

// ============================================================================================ decoder
// Insert-and-rank (FenwickTree::insert_then_forward_lookup, fenwick_tree.h:42-94): ids decoded so far that are
// smaller than v. Same data as the lane design: value-indexed buckets of 64 ids (two 128-byte lines) in global
// memory, every count in shared memory (level 0: u8 per bucket, level 1: u16 per 16 buckets, level 2: u16 per
// 256 buckets). The group fetches a bucket line as one coalesced request, each lane compares its slice, the
// count levels are summed a slice per lane, and one shuffle reduction yields the rank.

struct GDecTree {
    uint32_t* rec;  // nb buckets of 64 ids
    uint32_t* ovf;  // pairs (bucket, id)
    SmView sm;
    uint32_t sm_l0;
    uint32_t nb, ovf_cap, ovf_n;
    uint32_t lo, hi;
    uint64_t scale;
    uint32_t degenerate;
};

IDC_HD GDecTree gdec_tree_at(uint8_t* ws, SmView sm, uint32_t n, uint32_t lo, uint32_t hi) {
    DecTreeLayout L = dec_tree_layout(n);
    GDecTree t;
    t.rec = reinterpret_cast<uint32_t*>(ws);
    t.ovf = reinterpret_cast<uint32_t*>(ws + 256ull * L.nb);
    t.sm = sm;
    t.sm_l0 = 4u + 8u * L.l1_sectors;
    t.nb = L.nb;
    t.ovf_cap = L.ovf_cap;
    t.ovf_n = 0;
    if (hi < lo) hi = lo;
    t.lo = lo;
    t.hi = hi;
    uint64_t range = (uint64_t)hi - lo + 1ull;
    t.scale = ((uint64_t)L.nb << 32) / range;
    t.degenerate = 0;
    return t;
}

IDC_HD uint32_t gdec_bucket(const GDecTree& t, uint32_t v) {
    uint32_t c = v < t.lo ? t.lo : (v > t.hi ? t.hi : v);
    uint64_t b = ((uint64_t)(c - t.lo) * t.scale) >> 32;
    return b >= t.nb ? t.nb - 1u : (uint32_t)b;
}

// what gdec_rank learned about the bucket, for gdec_insert
struct GDecHit {
    uint32_t* p0;   // shared-memory word holding the bucket's count
    uint32_t w0;    // its value
    uint32_t b, cnt;
    bool normal;    // bucket path taken (unit has work and is not degenerate)
};

template <class GR>
IDC_HD void gdec_insert(const GR& g, GDecTree& t, uint32_t v, const GDecHit& hit);

template <typename OutT>
struct GDecUnit {
    DecState st;
    GDecTree tree;
    OutT* out;  // the unit's n output slots
    uint32_t n;
    int prec;
    // The insert of the previous step, not applied yet: the bucket store, the three counters and the output store of
    // step i do not feed step i + 1's bucket REQUEST (only its count word, checked below), so they are issued in the
    // shadow of that request instead of in front of it -- a warp is one instruction stream, and whatever stands
    // between "rank known" and "next line requested" is on the unit's serial chain whether it depends on it or not.
    GDecHit pend_hit;
    uint32_t pend_id, pend_pos, pend_has;
};

template <typename OutT>
IDC_HD void gdec_unit_start(GDecUnit<OutT>& U) {
    U.pend_has = 0, U.pend_id = 0, U.pend_pos = 0;
    U.pend_hit.normal = false, U.pend_hit.b = 0, U.pend_hit.cnt = 0, U.pend_hit.w0 = 0, U.pend_hit.p0 = nullptr;
}

// apply the pending insert (lane 0 stores); the caller synchronises the group before anything reads it back
template <class GR, typename OutT>
IDC_HD void gdec_flush(const GR& g, GDecUnit<OutT>& U) {
    if (!U.pend_has) return;
    gdec_insert(g, U.tree, U.pend_id, U.pend_hit);
    if (g.sub == 0) U.out[U.pend_pos] = (OutT)U.pend_id;
    if (U.tree.degenerate) U.st.status |= kStDegenerate;
    U.pend_has = 0;
}

// rank of v among the ids decoded so far, the previous step's insert applied on the way. The collectives (the
// warp-wide vote, the rendezvous after the flush, the final sum) are reached by every lane of the warp; `act`
// false: no memory is touched. `shadow()` is work of the caller that nothing here depends on (table look-ahead);
// it runs while the bucket request is in flight.
template <int G, class GR, typename OutT, class Shadow>
IDC_HD uint32_t gdec_rank(const GR& g, GDecUnit<OutT>& U, uint32_t v, uint32_t decoded, bool act, GDecHit& hit,
                          const uint32_t* mt, Shadow&& shadow) {
    constexpr int SL = 32 / G;  // bucket slots per lane and line
    GDecTree& t = U.tree;
    uint32_t part = 0;
    hit.b = 0, hit.cnt = 0, hit.w0 = 0, hit.p0 = nullptr;
    bool normal = act && !t.degenerate;
    uint32_t b = 0;
    uint32_t* p0 = nullptr;
    if (normal) {
        b = gdec_bucket(t, v);
        p0 = t.sm.at(t.sm_l0 + (b >> 2));
    }
    // The pending insert must be in place BEFORE the request when this step reads what it writes ahead of the
    // request -- the bucket's count word (4 buckets per word), hence also the bucket itself -- or when it may change
    // the unit's mode (an overflowing bucket can fill the spill list: brute-force ranks from then on), or when this
    // step does not take the bucket path at all. Rare (a few steps in a thousand); decided for the whole warp.
    const bool hazard = U.pend_has && (!normal || !U.pend_hit.normal || U.pend_hit.cnt >= kBkSlots || (b >> 2) == (U.pend_hit.b >> 2));
    const bool early = g.warp_any(hazard);
    if (early) {
        dec_ring_advance(U.st, mt);
        gdec_flush(g, U);
        g.sync();
        normal = act && !t.degenerate;
    }
    hit.normal = normal;
    uint32_t s0[SL], s1[SL];
    bool need0 = false, need1 = false;
    uint32_t cnt = 0;
    if (normal) {
        // this bucket's count (same word in every lane) decides which slices of the bucket are fetched: DRAM
        // traffic and latency follow the number of 32-byte sectors asked for (fetching all four sectors of the
        // first line unconditionally, before the count is known, was measured 10 % slower)
        const uint32_t w0 = *p0;
        cnt = (w0 >> (8u * (b & 3u))) & 0xffu;
        hit.b = b, hit.cnt = cnt, hit.w0 = w0, hit.p0 = p0;
        const uint32_t sv = cnt < kBkSlots ? cnt : kBkSlots;
        const uint32_t* bk = t.rec + (size_t)b * kBkSlots;
        need0 = sv > g.sub * SL, need1 = sv > 32u + g.sub * SL;
        if (need0) ld_line_slice<SL>(bk, g.sub, s0);
        if (need1) ld_line_slice<SL>(bk + 32u, g.sub, s1);
    }
    issue_fence();  // the bucket fetch is in flight before anything below is issued
    if (!early) {
        dec_ring_advance(U.st, mt);
        gdec_flush(g, U);
        shadow();
        g.sync();  // the insert (and the stream words that have landed) are visible to every lane of the group
    } else {
        shadow();
    }
    if (act && !normal) {
        const OutT* out_prev = U.out + (U.n - decoded);
        for (uint32_t i = g.sub; i < decoded; i += (uint32_t)G) part += ((uint32_t)out_prev[i] < v) ? 1u : 0u;
    } else if (normal) {
        const uint32_t grp = b >> 4, sct = grp >> 4;
        // ---- counts below the bucket, one slice per lane
#pragma unroll
        for (uint32_t j0 = 0; j0 < 4u; j0 += (uint32_t)G) {
            const uint32_t j = j0 + g.sub;  // the four words of levels 0 and 2 are dealt round the group
            if (j < 4u) {
                // level 0: the 16 bucket bytes of this group of buckets, word j
                uint32_t nib = (((1u << (b & 15u)) - 1u) >> (4u * j)) & 15u;
                uint32_t sel = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
                part = dp4a_u(*t.sm.at(t.sm_l0 + 4u * grp + j), sel, part);
                // level 2: eight u16 entries, word j
                uint32_t two = (((1u << sct) - 1u) >> (2u * j)) & 3u;
                part = dp2a(*t.sm.at(j), (two & 1u) | ((two & 2u) << 7), part);
            }
        }
        {
            // level 1: sixteen u16 entries of sector sct, 8 / G words per lane
            constexpr int W1 = 8 / G;
            uint32_t w1[W1];
            sm_load<W1>(t.sm.at(4u + 8u * sct + g.sub * W1), w1);
            const uint32_t incl = (1u << (grp & 15u)) - 1u;
#pragma unroll
            for (int j = 0; j < W1; j++) {
                uint32_t two = (incl >> (2u * (g.sub * W1 + (uint32_t)j))) & 3u;
                part = dp2a(w1[j], (two & 1u) | ((two & 2u) << 7), part);
            }
        }
        // ---- exact tie-break inside the bucket: unused slots hold 0xffffffff (pre-filled), never < v
        // (straight-line compares over pre-set registers instead of a branch per lane: measured slower, 101 -> 104 ms)
        if (need0) {
#pragma unroll
            for (int j = 0; j < SL; j++) part += s0[j] < v ? 1u : 0u;
        }
        if (need1) {
#pragma unroll
            for (int j = 0; j < SL; j++) part += s1[j] < v ? 1u : 0u;
        }
        if (cnt > kBkSlots) {
            for (uint32_t e = g.sub; e < t.ovf_n; e += (uint32_t)G) {
                uint64_t pr = ld_ws64(reinterpret_cast<const uint64_t*>(t.ovf) + e);
                part += ((uint32_t)pr == b && (uint32_t)(pr >> 32) < v) ? 1u : 0u;
            }
        }
    }
    return group_sum(g, part, G);
}

// record v in its bucket and bump the three count levels
template <class GR>
IDC_HD void gdec_insert(const GR& g, GDecTree& t, uint32_t v, const GDecHit& hit) {
    if (!hit.normal) return;
    const uint32_t b = hit.b, cnt = hit.cnt;
    if (cnt < kBkSlots) {
        if (g.sub == 0) st_ws32(t.rec + (size_t)b * kBkSlots + cnt, v);
    } else if (cnt < 254u && t.ovf_n < t.ovf_cap) {
        if (g.sub == 0) st_ws64(reinterpret_cast<uint64_t*>(t.ovf) + t.ovf_n, (uint64_t)b | ((uint64_t)v << 32));
        t.ovf_n++;
    } else {
        t.degenerate = 1;  // from now on ranks come from the output array
        return;
    }
    if (g.sub == 0) {
        const uint32_t grp = b >> 4, sct = grp >> 4;
        *hit.p0 = hit.w0 + (1u << (8u * (b & 3u)));
        sm_add(t.sm.at(4u + (grp >> 1)), 1u << (16u * (grp & 1u)));
        sm_add(t.sm.at(sct >> 1), 1u << (16u * (sct & 1u)));
    }
}

struct NoShadow {
    IDC_HD void operator()() const {}
};

// i = 0-based step; q31 = 2^31 / (i + 1). The step's insert and output store stay pending until the next step (or
// gdec_finish) applies them.
template <int G, class GR, typename OutT, class Shadow = NoShadow>
IDC_HD void gdec_step(const GR& g, GDecUnit<OutT>& U, uint32_t i, uint32_t q31, const uint32_t* mt, bool act,
                      Shadow&& shadow = Shadow()) {
    uint32_t id = 0;
    // the fetching lane's copies of three and more steps ago have landed; the rendezvous in gdec_rank() makes them
    // visible to the group. (Not in the shadow of the bucket request: the wait is a DEPBAR on the scoreboard the
    // compiler also gives the bucket loads, and with a threshold this low it would wait for them -- measured.)
    dec_ring_sync();
    if (act) id = dec_pop_id32(U.st, U.prec, mt);
    GDecHit hit;
    const uint32_t rank = gdec_rank<G>(g, U, id, i, act, hit, mt, shadow);
    if (act) {
        dec_push_uniform(U.st, rank, i + 1u, q31, mt);
        U.pend_hit = hit;
        U.pend_id = id;
        U.pend_pos = U.n - 1u - i;
        U.pend_has = 1u;
    }
}

// after the last step: the last insert's output store (every lane of the warp calls this)
template <class GR, typename OutT>
IDC_HD void gdec_finish(const GR& g, GDecUnit<OutT>& U) {
    gdec_flush(g, U);
    g.sync();
}

}  // namespace idc
