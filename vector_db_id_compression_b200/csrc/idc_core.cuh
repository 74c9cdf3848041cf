// idc_core.cuh -- building blocks of the ROC codec: the exact rANS arithmetic of the reference, and the layouts
// of the order-statistic workspaces. The per-step bodies that use them are in roc_group.cuh.
//
// Everything here is written once as IDC_HD (host + device) code: the CUDA
// kernels in roc_kernels.cu run one rANS stream per group of lanes (interleaved
// across lists: several independent heads per warp, one per unit), and
// tests/hostsim compiles the very same functions with g++ to check them
// against the oracle before any GPU time is spent. The host build is test
// infrastructure only -- the shipped library contains no CPU code path.
//
// Reference semantics restated here (paths relative to the reference tree):
//   ANSState                         custom_invlist_cpp/codec.h:13-45
//   pop/push_with_finer_precision    custom_invlist_cpp/codec.cpp:21-63
//   vrans_push / vrans_pop           custom_invlist_cpp/codec.cpp:65-90
//   codec_push / codec_pop           custom_invlist_cpp/codec.cpp:92-121
//   order statistics                 fenwick_tree_cpp/src/fenwick_tree.h:42-140
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define IDC_HD __host__ __device__ __forceinline__
#else
#define IDC_HD inline
#endif

namespace idc {

constexpr uint64_t kRansL = 1ull << 31;  // codec.cpp:19
constexpr int kMtWords = 8;              // words of std::mt19937(1234) kept on the device
constexpr uint32_t kMaxUnit = 65536;     // reference round-trip domain

// status bits reported per unit
constexpr uint32_t kStOverlay = 1u;      // decoder stack rose 2 words above its low-water mark
constexpr uint32_t kStMtDraws = 2u;      // more than kMtWords draws from the mt19937 fallback
constexpr uint32_t kStUnsorted = 4u;     // ids not ascending although IDC_F_SORTED was given
constexpr uint32_t kStWide = 8u;         // id does not fit 32 bits
constexpr uint32_t kStScratch = 16u;     // encoder ran out of scratch words (cannot happen within the bound)
constexpr uint32_t kStDegenerate = 32u;  // decoder fell back to brute-force ranks (informational)
constexpr uint32_t kStRange = 64u;       // a row number on the device was out of range

struct uint4x {  // 16 bytes; uint4 on device
    uint32_t x, y, z, w;
};

// ------------------------------------------------------------- primitives --

IDC_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

IDC_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// sum of the two 16-bit halves of a, each multiplied by byte 0 / byte 1 of sel
IDC_HD uint32_t dp2a(uint32_t a, uint32_t sel, uint32_t acc) {
#if defined(__CUDA_ARCH__)
    return __dp2a_lo(a, sel, acc);
#else
    return acc + (a & 0xffffu) * (sel & 0xffu) + (a >> 16) * ((sel >> 8) & 0xffu);
#endif
}

// Workspace accesses. Device: L2-only loads (.cg) so that fire-and-forget
// reductions (RED at L2) and later loads of the same word stay coherent
// without relying on L1 invalidation; host: plain memory.
IDC_HD uint4x ld_ws16(const void* p) {
#if defined(__CUDA_ARCH__)
    uint4 v = __ldcg(reinterpret_cast<const uint4*>(p));
    return uint4x{v.x, v.y, v.z, v.w};
#else
    return *reinterpret_cast<const uint4x*>(p);
#endif
}
IDC_HD uint32_t ld_ws32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
IDC_HD uint64_t ld_ws64(const uint64_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(reinterpret_cast<const unsigned long long*>(p));
#else
    return *p;
#endif
}
IDC_HD void st_ws32(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    __stcg(p, v);
#else
    *p = v;
#endif
}
IDC_HD void st_ws64(uint64_t* p, uint64_t v) {
#if defined(__CUDA_ARCH__)
    __stcg(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
#else
    *p = v;
#endif
}
template <typename T>
IDC_HD T load_id_raw(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

template <typename IdT>
IDC_HD uint64_t load_id(const IdT* p) {
    IdT v = load_id_raw(p);
    if (sizeof(IdT) == 4) return (uint64_t)(uint32_t)v;
    return (uint64_t)v;
}

// ------------------------------------------------ sector-of-16-counts math --
// A "sector" is 32 bytes = 16 unsigned 16-bit counts = two 16-byte loads.

struct Sector {
    uint32_t w[8];
};

// one 32-byte sector from the global workspace: a single 256-bit L2-only load (LDG.E.256) per lane
IDC_HD Sector ld_sector(const uint16_t* base, uint32_t sector_idx) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(base) + (size_t)sector_idx * 32;
    Sector s;
#if defined(__CUDA_ARCH__)
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(s.w[0]), "=r"(s.w[1]), "=r"(s.w[2]), "=r"(s.w[3]), "=r"(s.w[4]), "=r"(s.w[5]), "=r"(s.w[6]),
                   "=r"(s.w[7])
                 : "l"(p));
#else
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
    for (int j = 0; j < 8; j++) s.w[j] = q[j];
#endif
    return s;
}

// predicated memory operations: straight-line code in every lane. A data-dependent `if` around a store or a
// load makes the groups of a warp take different paths; each divergence costs a branch resolve and a reconvergence
// barrier on the step's critical path (measured: 40 % of the rANS push).
IDC_HD void st_ws32_if(uint32_t* p, uint32_t v, bool cond) {
#if defined(__CUDA_ARCH__)
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.global.cg.u32 [%0], %1; }" ::"l"(p), "r"(v), "r"((uint32_t)cond)
                 : "memory");
#else
    if (cond) *p = v;
#endif
}
// ------------------------------------------------------ encoder rANS state --
// The encoder's stack is append-only in the codec's valid domain; the rare
// pop-from-stack cases are still implemented exactly (read back the word just
// written, or draw from the mt19937(1234) table when the stack is empty).

struct EncState {
    uint64_t head;
    uint32_t* words;   // this unit's scratch slot
    uint32_t sp;       // words on the stack
    uint32_t cap;
    uint32_t draws;
    uint32_t status;
    uint32_t wr;       // this lane stores the words (lane 0 of the group that owns the unit)
};

// push w if `cond`
IDC_HD void enc_spill_if(EncState& st, uint32_t w, bool cond) {
    const bool ok = st.sp < st.cap;
    st_ws32_if(st.words + st.sp, w, cond & ok & (st.wr != 0u));
    st.status |= (cond & !ok) ? kStScratch : 0u;
    st.sp += cond ? 1u : 0u;
}

IDC_HD uint32_t enc_refill(EncState& st, const uint32_t* mt) {
    if (st.sp) {
        st.sp--;
        return st.sp < st.cap ? ld_ws32(st.words + st.sp) : 0u;
    }
    uint32_t d = st.draws++;
    if (d >= (uint32_t)kMtWords) {
        st.status |= kStMtDraws;
        return 0;
    }
    return mt[d];
}

// codec.cpp:21-42. rcp = floor((2^64-1)/nmax), q31 = 2^31/nmax.
// The spill test `head >= nmax*((2^31/nmax) << 32)` only looks at the upper
// word because the threshold's lower word is zero.
IDC_HD uint32_t enc_pop_uniform(EncState& st, uint32_t nmax, uint64_t rcp, uint32_t q31, const uint32_t* mt) {
    uint64_t h = st.head;
    const uint32_t hi = (uint32_t)(h >> 32), low = (uint32_t)h;
    const bool spill = hi >= nmax * q31;  // nmax*q31 <= 2^31: no overflow
    h = spill ? (uint64_t)hi : h;
    uint64_t q = mulhi64(h, rcp);
    uint64_t r = h - q * nmax;
    const bool fix = r >= nmax;
    q += fix ? 1u : 0u;
    r -= fix ? nmax : 0u;
    const bool refill = h < kRansL;
    // spill && refill: the word just spilled is popped straight back, the stack is unchanged.
    // refill without spill needs head < 2^31 on entry: never inside the codec's valid domain (kept exact, cold).
    enc_spill_if(st, low, spill & !refill);
    uint32_t w = low;
    if (refill & !spill) w = enc_refill(st, mt);
    st.head = refill ? ((uint64_t)w | (q << 32)) : q;
    return (uint32_t)r;
}

// codec.cpp:65-76, start added unmasked; p in 0..16
IDC_HD void enc_push_bits(EncState& st, uint32_t start, uint32_t p) {
    uint64_t h = st.head;
    const uint32_t hi = (uint32_t)(h >> 32);
    const bool spill = hi >= (0x80000000u >> p);
    enc_spill_if(st, (uint32_t)h, spill);
    h = spill ? (uint64_t)hi : h;
    st.head = (h << p) + start;
}

// codec.cpp:92-105 for ids < 2^32 and precision <= 32 (the device domain). The reference pushes four 16-bit
// slices, low slice first; the slices at lower = 32 and 48 then have precision 0 and add 0, but still run the
// spill test `head >= 2^63` (codec.cpp:69) -- false after a push of an id < 2^precision, possible only for the
// mirrored power-of-two-max_id bug, so it is one cold check here (a second one could not fire: after a spill the
// head is below 2^32).
IDC_HD void enc_push_id32(EncState& st, uint32_t id, int precision) {
    const uint32_t p0 = precision < 16 ? (uint32_t)precision : 16u;
    const uint32_t p1 = (uint32_t)precision - p0;
    enc_push_bits(st, id & 0xffffu, p0);
    enc_push_bits(st, id >> 16, p1);
    if (st.head >> 63) enc_push_bits(st, 0u, 0u);
}

// ------------------------------------------------------ decoder rANS state --
// Reads the blob's words top-down; pushes go to a one-word overlay register
// (the decoder never holds more than one word above its low-water mark inside
// the codec's valid domain -- violations are flagged, not assumed away).

// The blob's words are consumed top-down through a ring of kDecRing words in SHARED memory that is kept full by
// asynchronous copies (cp.async: global -> shared without a destination register). Why not global loads into
// look-ahead registers: a warp's scoreboard is per register, not per lane -- with 8 units per warp some group
// re-loads its look-ahead register at almost every pop, and every other group's next read of "its" copy of that
// register then waits for that load (measured: 570 of the decoder's 3 100 cycles per step).
//
// The next THREE words sit in registers (pk0..pk2), loaded from the ring once per step by dec_ring_advance(), which
// the step runs in the shadow of its bucket request together with the refill of the slots that were freed. A
// renormalisation on the serial chain -- `if (h < 2^31) h = (h << 32) | pop()` -- is then a compare, a select among
// the three registers and a shift: no memory operation, no pointer update, no branch. Between two calls of
// dec_ring_advance() a stream of this codec takes at most three words (one by the push -- which leaves the head
// >= 2^31 -- and one by each of the two 16-bit pops of the next id); a fourth is flagged like a second overlay word.
constexpr uint32_t kDecRing = 16;      // words of look-ahead in shared memory
constexpr uint32_t kDecRingWait = 2;   // copy groups (one per step) allowed in flight at a step's rendezvous
constexpr uint32_t kDecPeek = 3;       // words of look-ahead in registers

struct DecRing {
    uint32_t* base;       // this unit's ring in shared memory: word i at base[(i >> 2) * chunk_stride + (i & 3)]
    uint32_t chunk_stride;
    IDC_HD uint32_t* at(uint32_t i) const { return base + ((i >> 2) * chunk_stride + (i & 3u)); }  // 32-bit index math: a ring is a few KB
};

IDC_HD void ring_fetch_if(uint32_t* dst_smem, const uint32_t* src, bool cond) {
#if defined(__CUDA_ARCH__)
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q cp.async.ca.shared.global [%0], [%1], 4; }" ::"r"(d), "l"(src),
                 "r"((uint32_t)cond)
                 : "memory");
#else
    if (cond) *dst_smem = *src;
#endif
}
IDC_HD void ring_commit() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
IDC_HD void ring_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

struct DecState {
    uint64_t head;
    const uint32_t* words;  // the unit's words (read directly only by the host emulation, see dec_ring_word)
    const uint32_t* fp;  // next word to FETCH into the ring is fp[-1]
    uint32_t fleft;      // words of the blob's stack not fetched yet
    uint32_t sp;         // words of the blob's stack not consumed yet, as of the last dec_ring_advance()
    uint32_t rpos;       // ring slot of pk0
    uint32_t pk0, pk1, pk2;  // the next three words pop() returns: the blob's, then the mt19937(1234) fallback of codec.h:32-40
    uint32_t used;       // how many of them have been taken since the last dec_ring_advance()
    DecRing ring;
    uint32_t fetcher;    // this lane issues the copies (lane 0 of the group that owns the unit)
    uint32_t ov;         // overlay: the one word the decoder may hold above the blob's stack
    uint32_t has_ov;
    uint32_t draws;      // fallback words consumed, as of the last dec_ring_advance()
    uint32_t status;
};

// every lane of the group calls this; afterwards the caller synchronises the group (the fetching lane has waited
// for the copies, the other lanes see them after the rendezvous) and calls dec_ring_prime()
IDC_HD void dec_state_init(DecState& st, uint64_t head, const uint32_t* words, uint32_t nwords, DecRing ring, bool fetcher) {
    st.head = head;
    st.words = words;
    st.sp = nwords;
    st.rpos = 0;
    st.ring = ring;
    st.fetcher = fetcher ? 1u : 0u;
    const uint32_t first = nwords < kDecRing ? nwords : kDecRing;
    for (uint32_t i = 0; i < first; i++) ring_fetch_if(ring.at(i), words + nwords - 1u - i, fetcher);
    ring_commit();
    ring_wait<0>();
    st.fp = words + nwords - first;
    st.fleft = nwords - first;
    st.ov = 0;
    st.has_ov = 0;
    st.draws = 0;
    st.status = 0;
    st.pk0 = st.pk1 = st.pk2 = 0;  // set by dec_ring_prime() once the group has met
    st.used = 0;
}

// The k-th word below the consumed part of the blob's stack: ring slot rpos + k. Every lane of the group reads it long
// after the copy that filled it has landed and a rendezvous has made it visible; the slots dec_ring_advance() reads
// (rpos .. rpos + 2 after the advance) are never the ones the fetching lane refills in the same or in the next call
// before the step's rendezvous, so the protocol holds for lanes that are a whole step phase apart -- the host
// emulation (free-running threads, copies that land at once) runs it as is.
IDC_HD uint32_t dec_ring_word(const DecState& st, uint32_t k) { return *st.ring.at((st.rpos + k) & (kDecRing - 1u)); }

// Once per step, before the step's rendezvous: all but the kDecRingWait most recent copy groups (one per step) have
// landed. A slot is refilled when its word is consumed and read again 14 words -- at three words per step at most,
// five steps -- later: its copy is waited for at the start of the third step after it and visible after that
// step's rendezvous.
IDC_HD void dec_ring_sync() { ring_wait<kDecRingWait>(); }

// d-th output of std::mt19937(1234). Device: a select chain over immediates, NOT a load from the table -- ptxas puts
// every global load of a kernel on one scoreboard, so a register that any LDG may write (even on a path never taken)
// is guarded by a wait for ALL global loads in flight, the step's bucket request included (measured: 1 050 cycles
// per step in the shadow of that request). tests/test_cabi_cpu.py checks the constants against std::mt19937.
IDC_HD uint32_t mt_word(uint32_t d, const uint32_t* mt) {
#if defined(__CUDA_ARCH__)
    (void)mt;
    constexpr uint32_t w[kMtWords] = {822569775u, 2137449171u, 2671936806u, 3512589365u,
                                      1880026316u, 2629000564u, 3373089432u, 3312965625u};
    uint32_t r = 0;
#pragma unroll
    for (uint32_t k = 0; k < (uint32_t)kMtWords; k++) r = d == k ? w[k] : r;
    return r;
#else
    return d < (uint32_t)kMtWords ? mt[d] : 0u;
#endif
}

// Account for the `used` words taken since the last call: free their ring slots and refill them with the words
// kDecRing further down (fetching lane), count fallback draws, reload the look-ahead registers. Off the serial chain:
// nothing here depends on the step's rank.
IDC_HD void dec_ring_advance(DecState& st, const uint32_t* mt) {
    const uint32_t u = st.used;
    const uint32_t fromblob = u < st.sp ? u : st.sp;
#pragma unroll
    for (uint32_t k = 0; k < kDecPeek; k++) {
        const bool more = (k < fromblob) & (st.fleft != 0u);
        ring_fetch_if(st.ring.at((st.rpos + k) & (kDecRing - 1u)), st.fp - 1, more & (st.fetcher != 0u));
        st.fp -= more ? 1 : 0;
        st.fleft -= more ? 1u : 0u;
    }
    ring_commit();
    st.sp -= fromblob;
    st.rpos = (st.rpos + fromblob) & (kDecRing - 1u);
    st.draws += u - fromblob;
    if (st.draws > (uint32_t)kMtWords) st.status |= kStMtDraws;
    uint32_t pk[kDecPeek];
#pragma unroll
    for (uint32_t k = 0; k < kDecPeek; k++) {
        if (k < st.sp) {
            pk[k] = dec_ring_word(st, k);
        } else {  // below the blob's bottom: the fallback words (cold: the last steps of a stream at most)
            const uint32_t d = st.draws + (k - st.sp);
            pk[k] = mt_word(d, mt);
        }
    }
    st.pk0 = pk[0], st.pk1 = pk[1], st.pk2 = pk[2];
    st.used = 0;
}

// after dec_state_init and a rendezvous of the group
IDC_HD void dec_ring_prime(DecState& st, const uint32_t* mt) { dec_ring_advance(st, mt); }

// `if (h < 2^31) h = (h << 32) | pop()` of codec.cpp:83-87 / :56-60 as straight-line register code. pop() = the
// overlay if there is one, else the next look-ahead word.
IDC_HD uint64_t dec_renorm(DecState& st, uint64_t h, const uint32_t* mt) {
    const bool rf = h < kRansL;
    const bool o = st.has_ov != 0u;
    const bool take = rf & !o;
    (void)mt;
    st.status |= (take & (st.used >= kDecPeek)) ? kStOverlay : 0u;  // a fourth word inside one step: not a stream of this codec
    const uint32_t pk = st.used == 0u ? st.pk0 : (st.used == 1u ? st.pk1 : st.pk2);
    const uint32_t w = o ? st.ov : pk;
    st.used += take ? 1u : 0u;
    st.has_ov = rf ? 0u : st.has_ov;
    return rf ? ((h << 32) | (uint64_t)w) : h;
}

// codec.cpp:78-90; p in 0..16
IDC_HD uint32_t dec_pop_bits(DecState& st, uint32_t p, const uint32_t* mt) {
    const uint64_t h = st.head;
    const uint32_t sym = (uint32_t)h & ((1u << p) - 1u);
    st.head = dec_renorm(st, h >> p, mt);
    return sym;
}

// codec.cpp:107-121 for precision <= 32: the slices at lower = 48 and 32 have precision 0 -- they pop nothing but
// still renormalise (codec.cpp:83-87). The encoder leaves every head >= 2^31 (its initial state is 2^31, a
// vrans_push never ends below it), and what the decoder pops from is the head a push_with_finer_precision left --
// the encoder's head before the matching pop -- so inside a stream of this codec those two renormalisations do
// nothing. dec_pop_start() runs them once, for whatever head a blob was imported with; the per-id pop flags a head
// below 2^31 (not a stream of this codec) instead of branching on it.
IDC_HD void dec_pop_start(DecState& st, const uint32_t* mt) {
    if (st.head < kRansL) st.head = dec_renorm(st, st.head, mt);
    if (st.head < kRansL) st.head = dec_renorm(st, st.head, mt);
}

IDC_HD uint32_t dec_pop_id32(DecState& st, int precision, const uint32_t* mt) {
    st.status |= st.head < kRansL ? kStOverlay : 0u;
    const uint32_t p0 = precision < 16 ? (uint32_t)precision : 16u;
    const uint32_t p1 = (uint32_t)precision - p0;
    const uint32_t hi = dec_pop_bits(st, p1, mt);
    const uint32_t lo = dec_pop_bits(st, p0, mt);
    return (hi << 16) | lo;
}

// codec.cpp:44-63; q31 = 2^31/nmax
IDC_HD void dec_push_uniform(DecState& st, uint32_t sym, uint32_t nmax, uint32_t q31, const uint32_t* mt) {
    uint64_t h = st.head;
    const uint32_t hi = (uint32_t)(h >> 32);
    const bool spill = hi >= q31;
    // the spilled word goes to the overlay; a second word above the blob's stack is flagged, not assumed away
    st.status |= (spill & (st.has_ov != 0u)) ? kStOverlay : 0u;
    st.ov = spill ? (uint32_t)h : st.ov;
    st.has_ov = spill ? 1u : st.has_ov;
    h = spill ? (uint64_t)hi : h;
    h = h * nmax + sym;
    st.head = dec_renorm(st, h, mt);
}

// ------------------------------------------------- encoder: select-remove --
// Order statistics over the unit's id-sorted array, replacing
// FenwickTree::reverse_lookup_then_remove (fenwick_tree.h:96-140).
//
// A random HBM access costs a whole 128-byte line whatever is asked for, so a step may afford ONE: the unit's
// ids are re-laid as 128-byte RECORDS -- word 0 a presence mask, words 1..31 thirty-one consecutive ids of the
// sorted unit -- and one line load yields both "which ids of this stretch are still there" and the id itself.
// Everything that decides WHICH record to touch lives in shared memory, private to the lane:
//   L0  one word per 6 records: six 5-bit counts of ids still present
//   L1  per 12 records (two L0 words) an exclusive cumulative 16-bit count, 16 entries (a "sector") per
//       superblock of 192 records
//   L2  sixteen 32-bit exclusive cumulative counts over the superblocks (n <= 65536 -> <= 12 of them)
// Cumulative entries make select a batch of independent compares (no serial prefix scan on the step's
// critical path); the matching decrements after a removal are independent word updates off that path.

constexpr uint32_t kRecIds = 31;          // ids per record
constexpr uint32_t kRecPerWord = 6;       // 5-bit counts per L0 word
constexpr uint32_t kRecPerPair = 12;      // records per L1 entry
constexpr uint32_t kRecPerSuper = 192;    // records per superblock (16 L1 entries)

struct EncTreeLayout {
    uint32_t records;   // ceil(n / 31)
    uint32_t l0_words;  // ceil(records / 6), rounded up to even
    uint32_t supers;    // ceil(records / 192)
};

IDC_HD EncTreeLayout enc_tree_layout(uint32_t n) {
    EncTreeLayout L;
    L.records = (n + kRecIds - 1u) / kRecIds;
    if (L.records == 0) L.records = 1;
    L.l0_words = ((L.records + kRecPerWord - 1u) / kRecPerWord + 1u) & ~1u;
    L.supers = (L.records + kRecPerSuper - 1u) / kRecPerSuper;
    return L;
}
// global workspace: the records
IDC_HD uint64_t enc_tree_bytes(uint32_t n) { return 128ull * enc_tree_layout(n).records; }
// shared-memory words per lane: [L2: 16][L1: 8 per superblock][L0]
IDC_HD uint32_t enc_tree_sm_words(uint32_t n) {
    EncTreeLayout L = enc_tree_layout(n);
    return 16u + 8u * L.supers + L.l0_words;
}

// ids present in records [0, r) of a full unit of n ids
IDC_HD uint32_t enc_full_before(uint32_t n, uint32_t r) {
    uint64_t c = (uint64_t)r * kRecIds;
    return (uint32_t)(c < n ? c : n);
}

// record word 0 (mask) and words 1..31 (ids) for record r of a unit of n ascending ids
template <typename IdT>
IDC_HD uint32_t enc_record_word(const IdT* src, uint32_t n, uint32_t r, uint32_t w) {
    uint32_t first = r * kRecIds;
    uint32_t have = first < n ? (n - first < kRecIds ? n - first : kRecIds) : 0u;
    if (w == 0) return have ? (0xffffffffu >> (32u - have)) : 0u;
    return (w - 1u) < have ? (uint32_t)load_id_raw(src + first + (w - 1u)) : 0xffffffffu;
}

// position of the r-th (0-based) set bit of a 32-bit mask with more than r ones
IDC_HD uint32_t select32(uint32_t m, uint32_t r) {
    uint32_t pos = 0, c;
    c = (uint32_t)popc32(m & 0xffffu);
    if (r >= c) { r -= c; pos += 16; m >>= 16; }
    c = (uint32_t)popc32(m & 0xffu);
    if (r >= c) { r -= c; pos += 8; m >>= 8; }
    c = (uint32_t)popc32(m & 0xfu);
    if (r >= c) { r -= c; pos += 4; m >>= 4; }
    c = (uint32_t)popc32(m & 0x3u);
    if (r >= c) { r -= c; pos += 2; m >>= 2; }
    if (r >= (m & 1u)) pos += 1;
    return pos;
}

// ---------------------------------------------------- decoder: insert-rank --
// Replaces FenwickTree::insert_then_forward_lookup (fenwick_tree.h:42-94):
// number of already decoded ids strictly smaller than v, then insert v.
//
// Measured on B200: a random access to HBM moves a whole 128-byte line (3.7 sectors per isolated 32-byte
// load; 43 G such accesses/s = 5.5 TB/s), and L2 turns over in ~25 us under that traffic, so per step the
// only affordable DRAM state is ONE spot. Hence:
//   * ids are not known in advance, so the structure is indexed by VALUE: a monotone map id -> bucket;
//   * a bucket is two 128-byte lines = 64 id slots, filled in arrival order (unsorted): the tie-break
//     inside a bucket is a brute-force compare of <= 64 ids that arrive with the same DRAM access;
//   * every count (per bucket u8, per 16 buckets u16, per 256 buckets u16) lives in SHARED MEMORY, private
//     to the lane: prefix counts cost no DRAM traffic and the bucket lines never need a header or memset.
// Buckets that overflow (skewed ids) spill (bucket, id) pairs to a small per-unit list; if that fills up too
// (adversarial input) the unit switches to brute-force counting over its already written output. Exact in
// every case; the [lo, hi] range hint only affects speed.

constexpr uint32_t kBkSlots = 64;    // ids per bucket
constexpr uint32_t kBkTarget = 40;   // mean load: P(Poisson(40) > 64) ~ 2e-4

struct DecTreeLayout {
    uint32_t nb;          // buckets
    uint32_t ngroups;     // groups of 16 buckets = level-1 entries
    uint32_t l1_sectors;  // sectors of 16 level-1 entries = level-2 entries (<= 8 for n <= 65536)
    uint32_t ovf_cap;     // (bucket, id) pairs
};

IDC_HD DecTreeLayout dec_tree_layout(uint32_t n) {
    DecTreeLayout L;
    L.nb = (n + kBkTarget - 1u) / kBkTarget;
    if (L.nb == 0) L.nb = 1;
    L.ngroups = (L.nb + 15u) / 16u;
    L.l1_sectors = (L.ngroups + 15u) / 16u;
    L.ovf_cap = n / 16u + 32u;
    return L;
}

// global workspace per unit: bucket lines + overflow pairs, 128-byte aligned
IDC_HD uint64_t dec_tree_bytes(uint32_t n) {
    DecTreeLayout L = dec_tree_layout(n);
    uint64_t b = 256ull * L.nb + 8ull * L.ovf_cap;
    return (b + 127ull) & ~127ull;
}
// shared-memory words per lane: [level 2: 4 words][level 1: 8 per sector][level 0: 4 per group]
IDC_HD uint32_t dec_tree_sm_words(uint32_t n) {
    DecTreeLayout L = dec_tree_layout(n);
    return 4u + 8u * L.l1_sectors + 4u * L.ngroups;
}

IDC_HD uint32_t dp4a_u(uint32_t a, uint32_t sel, uint32_t acc) {
#if defined(__CUDA_ARCH__)
    return __dp4a(a, sel, acc);
#else
    for (int k = 0; k < 4; k++) acc += ((a >> (8 * k)) & 0xffu) * ((sel >> (8 * k)) & 0xffu);
    return acc;
#endif
}

}  // namespace idc
