// TEST INFRASTRUCTURE ONLY. Compiles the lane-level codec functions of
// vector_db_id_compression_b200/csrc (idc_core.cuh, roc_group.cuh, ef_core.cuh)
// with g++ so their logic can be checked against the oracle in the CPU test
// suite (no GPU in the build container). Not part of libidcodec.so; nothing in
// the product loads it.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../vector_db_id_compression_b200/csrc/roc_group.cuh"
#include "../../vector_db_id_compression_b200/csrc/roc_small.cuh"
#include "grp_emu.h"
#include "../../vector_db_id_compression_b200/csrc/ef_core.cuh"
#include "../../vector_db_id_compression_b200/csrc/wt_core.cuh"

using namespace idc;

static void tables(uint32_t* mt) {
    std::mt19937 g(1234);
    for (int i = 0; i < kMtWords; i++) mt[i] = (uint32_t)g();
}

// ---- the group design of roc_group.cuh: G host threads play the G lanes of the group that owns the unit
template <int G>
static int64_t group_encode(uint32_t n, const uint64_t* ids, int prec, uint64_t* head_out, uint32_t* words_out,
                            uint32_t cap, uint32_t* order_out, uint32_t* status_out) {
    uint32_t mt[kMtWords];
    tables(mt);
    std::vector<uint8_t> ws(enc_tree_bytes(n) + 256, 0);
    std::vector<uint32_t> sm(genc_sm_words(n) + 8, 0);
    uint32_t* rec = reinterpret_cast<uint32_t*>(ws.data());
    EncTreeLayout lay = enc_tree_layout(n);
    for (uint32_t r = 0; r < lay.records; r++)
        for (uint32_t w = 0; w < 32; w++)
            rec[r * 32 + w] = enc_record_word(reinterpret_cast<const int64_t*>(ids), n, r, w);
    int64_t result = 0;
    run_group(G, [&](const HostGrp& g) {
        GEncUnit<G, int64_t> U;
        U.tree.rec = rec;
        U.tree.sm = SmView{sm.data(), 1, 0};
        genc_tree_init<G>(g, U.tree, n);
        g.sync();
        U.st = EncState{kRansL, words_out, 0, cap, 0, 0, g.sub == 0 ? 1u : 0u};
        U.sort_idx = nullptr;
        U.order = order_out;
        U.pos_base = 0;
        U.n = n;
        U.prec = prec;
        for (uint32_t t = n; t >= 1; --t)
            genc_step<G>(g, U, t, ~0ull / t, (uint32_t)((1ull << 31) / t), mt, true);
        if (g.sub == 0) {
            *head_out = U.st.head;
            *status_out = U.st.status;
            result = U.st.status & kStScratch ? -1 : (int64_t)U.st.sp;
        }
    });
    return result;
}

template <int G>
static void group_decode(uint64_t head, const uint32_t* words, uint32_t nwords, uint32_t n, int prec, uint32_t lo,
                         uint32_t hi, int64_t* out, uint32_t* status_out, uint32_t force_degenerate) {
    uint32_t mt[kMtWords];
    tables(mt);
    std::vector<uint8_t> ws(dec_tree_bytes(n) + 64, 0xff);
    std::vector<uint32_t> sm(dec_tree_sm_words(n) + 8, 0);
    std::vector<uint32_t> ring(kDecRing, 0);
    run_group(G, [&](const HostGrp& g) {
        GDecUnit<int64_t> U;
        U.tree = gdec_tree_at(ws.data(), SmView{sm.data(), 1, 0}, n, lo, hi);
        if (force_degenerate) U.tree.ovf_cap = force_degenerate - 1;
        dec_state_init(U.st, head, words, nwords, DecRing{ring.data(), 4u}, g.sub == 0);
        g.sync();
        dec_ring_prime(U.st, mt);
        dec_pop_start(U.st, mt);
        U.out = out;
        U.n = n;
        U.prec = prec;
        gdec_unit_start(U);
        for (uint32_t i = 0; i < n; i++)
            gdec_step<G>(g, U, i, (uint32_t)((1ull << 31) / (i + 1)), mt, true);
        gdec_finish(g, U);
        if (g.sub == 0) *status_out = U.st.status;
    });
}

extern "C" {

// roc_small.cuh: one unit per thread (graph rows); out[n - 1 - i] = i-th decoded id like the kernel's write-out
void sim_small_decode(uint64_t head, const uint32_t* words, uint32_t nwords, uint32_t n, int prec, int64_t* out,
                      uint32_t* status_out) {
    uint32_t mt[kMtWords];
    tables(mt);
    std::vector<uint32_t> seen(n + 1, 0);
    SmallDec s;
    small_dec_init(s, head, words, nwords);
    small_dec_unit(s, n, prec, [&](uint32_t j) -> uint32_t& { return seen[j]; }, mt);
    for (uint32_t i = 0; i < n; i++) out[n - 1 - i] = (int64_t)seen[i];
    *status_out = s.status;
}

// one unit per warp (warp_dec_unit / warp_dec_row / warp_enc_unit): 32 host threads play the lanes
void sim_warp_decode(uint64_t head, const uint32_t* words, uint32_t nwords, uint32_t n, int prec, int row, int64_t* out,
                     uint32_t* status_out) {
    uint32_t mt[kMtWords];
    tables(mt);
    std::vector<uint32_t> seen(n + 64, 0), q31(n + 2, 0);
    for (uint32_t d = 1; d < n + 2; d++) q31[d] = (uint32_t)((1ull << 31) / d);
    run_group(32, [&](const HostGrp& g) {
        SmallDec st;
        small_dec_init(st, head, words, nwords);
        if (row) {  // n <= 64: the decoded ids stay in two "registers" per lane
            uint32_t s0 = 0, s1 = 0;
            warp_dec_row(g, st, n, prec, s0, s1, mt);
            if (g.sub < n) seen[g.sub] = s0;
            if (g.sub + 32u < n) seen[g.sub + 32u] = s1;
        } else {
            warp_dec_unit(g, st, n, prec, seen.data(), q31.data(), mt);
        }
        g.sync();
        if (g.sub == 0) *status_out = st.status;
    });
    for (uint32_t i = 0; i < n; i++) out[n - 1 - i] = (int64_t)seen[i];
}

int64_t sim_warp_encode(uint32_t n, const uint64_t* ids, int prec, uint64_t* head_out, uint32_t* words_out, uint32_t cap,
                        uint32_t* order_out, uint32_t* status_out) {
    uint32_t mt[kMtWords];
    tables(mt);
    std::vector<uint32_t> sid(n + 1), q31(n + 2, 0);
    std::vector<uint64_t> rcp(n + 2, 0);
    for (uint32_t i = 0; i < n; i++) sid[i] = (uint32_t)ids[i];
    for (uint32_t d = 1; d < n + 2; d++) q31[d] = (uint32_t)((1ull << 31) / d), rcp[d] = ~0ull / d;
    int64_t result = 0;
    run_group(32, [&](const HostGrp& g) {
        EncState st{kRansL, words_out, 0, cap, 0, 0, g.sub == 0 ? 1u : 0u};
        warp_enc_unit(g, st, n, prec, sid.data(), rcp.data(), q31.data(),
                      [&](uint32_t step, uint32_t pos) {
                          if (g.sub == 0) order_out[step] = pos;
                      },
                      mt);
        if (g.sub == 0) {
            *head_out = st.head;
            *status_out = st.status;
            result = st.status & kStScratch ? -1 : (int64_t)st.sp;
        }
    });
    return result;
}

// encoder twin: ids ascending, n <= 64; returns the word count (-1: scratch overflow)
int64_t sim_small_encode(uint32_t n, const uint64_t* ids, int prec, uint64_t* head_out, uint32_t* words_out, uint32_t cap,
                         uint32_t* order_out, uint32_t* status_out) {
    uint32_t mt[kMtWords];
    tables(mt);
    EncState st{kRansL, words_out, 0, cap, 0, 0, 1u};
    uint64_t mask = n >= 64 ? ~0ull : ((1ull << n) - 1ull);
    for (uint32_t t = n; t >= 1; --t) {
        const uint32_t pos = small_enc_step(st, mask, t, prec, ~0ull / t, (uint32_t)((1ull << 31) / t),
                                            [&](uint32_t p) { return (uint32_t)ids[p]; }, mt);
        order_out[n - t] = pos;
    }
    *head_out = st.head;
    *status_out = st.status;
    return st.status & kStScratch ? -1 : (int64_t)st.sp;
}

int64_t sim_group_encode(int G, uint32_t n, const uint64_t* ids, int prec, uint64_t* head_out, uint32_t* words_out,
                         uint32_t cap, uint32_t* order_out, uint32_t* status_out) {
    return G == 8   ? group_encode<8>(n, ids, prec, head_out, words_out, cap, order_out, status_out)
           : G == 2 ? group_encode<2>(n, ids, prec, head_out, words_out, cap, order_out, status_out)
                    : group_encode<4>(n, ids, prec, head_out, words_out, cap, order_out, status_out);
}
void sim_group_decode(int G, uint64_t head, const uint32_t* words, uint32_t nwords, uint32_t n, int prec, uint32_t lo,
                      uint32_t hi, int64_t* out, uint32_t* status_out, uint32_t force_degenerate) {
    if (G == 8)
        group_decode<8>(head, words, nwords, n, prec, lo, hi, out, status_out, force_degenerate);
    else if (G == 2)
        group_decode<2>(head, words, nwords, n, prec, lo, hi, out, status_out, force_degenerate);
    else
        group_decode<4>(head, words, nwords, n, prec, lo, hi, out, status_out, force_degenerate);
}

uint64_t sim_dec_tree_bytes(uint32_t n) { return dec_tree_bytes(n); }
uint64_t sim_enc_tree_bytes(uint32_t n) { return enc_tree_bytes(n); }


// Elias-Fano: every output word through the gather functions of ef_core.cuh
void sim_ef_shape(uint64_t universe, uint64_t m, uint64_t* out6) {
    EfShape s = ef_shape(universe, m);
    out6[0] = s.l, out6[1] = s.low_bits, out6[2] = s.high_bits, out6[3] = s.low_words, out6[4] = s.high_words,
    out6[5] = s.samples;
}
void sim_ef_encode(const int64_t* ids, uint64_t m, uint64_t universe, uint64_t* low, uint64_t* high, uint32_t* samples) {
    EfShape s = ef_shape(universe, m);
    for (uint64_t w = 0; w < s.low_words; w++) low[w] = ef_low_word(ids, m, s.l, w);
    for (uint64_t w = 0; w < s.high_words; w++) high[w] = ef_high_word(ids, m, s.l, universe, w);
    for (uint64_t i = 0; i < m; i += kEfSample) samples[i >> kEfSampleLog] = (uint32_t)ef_high_pos(ids, i, s.l);
}
uint64_t sim_ef_select(const uint64_t* low, const uint64_t* high, const uint32_t* samples, uint32_t l, uint64_t k) {
    return ef_select(low, high, samples, l, k);
}


// Wavelet matrix: the build passes of wt_kernels.cu lane by lane (a ballot = a loop over the 32 lanes), with the
// directory and destination formulas of wt_core.cuh; queries through wt_select / wt_access themselves.
uint32_t sim_wt_levels(uint64_t nlist, uint64_t n) { return wt_shape(nlist, n).levels; }
void sim_wt_build(uint64_t nlist, uint64_t n, const uint32_t* S, uint64_t* bits, uint32_t* rank, uint32_t* sel1,
                  uint32_t* sel0, uint32_t* start, const uint64_t* list_size) {
    WtShape sh = wt_shape(nlist, n);
    std::vector<uint32_t> seq(S, S + n), next(n);
    std::vector<uint32_t> ones(sh.nblk + 1);
    for (uint32_t lev = 0; lev < sh.levels; lev++) {
        uint32_t shift = sh.levels - 1 - lev;
        uint32_t* B32 = reinterpret_cast<uint32_t*>(bits + (uint64_t)lev * sh.words);
        uint32_t* R = rank + (uint64_t)lev * sh.rank_stride;
        for (uint64_t blk = 0; blk < sh.nblk; blk++) {  // k_wt_level_bits
            uint32_t cnt = 0;
            for (int t = 0; t < 16; t++) {
                uint32_t m = 0;
                for (uint32_t lane = 0; lane < 32; lane++) {
                    uint64_t i = (blk << kWtBlockLog) + (uint64_t)t * 32 + lane;
                    uint32_t v = i < n ? seq[i] : 0u;
                    m |= ((v >> shift) & 1u) << lane;
                }
                B32[blk * 16 + t] = m;
                cnt += (uint32_t)__builtin_popcount(m);
            }
            ones[blk] = cnt;
        }
        ones[sh.nblk] = 0;
        uint32_t acc = 0;  // the scan kernels: exclusive prefix, entry nblk = total
        for (uint64_t j = 0; j <= sh.nblk; j++) {
            uint32_t c = ones[j];
            ones[j] = acc;
            acc += c;
        }
        for (uint64_t j = 0; j <= sh.nblk; j++) {  // k_wt_directory
            R[j] = ones[j];
            if (j == sh.nblk) break;
            WtDirEntry d = wt_dir_entry(j, sh.nblk, n, ones[j], ones[j + 1]);
            if (d.has1) sel1[(uint64_t)lev * sh.samp_stride + d.m1] = (uint32_t)j;
            if (d.has0) sel0[(uint64_t)lev * sh.samp_stride + d.m0] = (uint32_t)j;
        }
        if (lev + 1 == sh.levels) break;
        uint64_t z = n - R[sh.nblk];
        for (uint64_t blk = 0; blk < sh.nblk; blk++) {  // k_wt_level_scatter
            uint64_t r1 = R[blk];
            for (int t = 0; t < 16; t++) {
                uint32_t m = 0;
                for (uint32_t lane = 0; lane < 32; lane++) {
                    uint64_t i = (blk << kWtBlockLog) + (uint64_t)t * 32 + lane;
                    if (i < n && ((seq[i] >> shift) & 1u)) m |= 1u << lane;
                }
                for (uint32_t lane = 0; lane < 32; lane++) {
                    uint64_t i = (blk << kWtBlockLog) + (uint64_t)t * 32 + lane;
                    if (i >= n) continue;
                    uint32_t b = (seq[i] >> shift) & 1u;
                    uint64_t before = r1 + (uint32_t)__builtin_popcount(m & ((1u << lane) - 1u));
                    next[wt_partition_dest(i, b, z, before)] = seq[i];
                }
                r1 += (uint32_t)__builtin_popcount(m);
            }
        }
        seq.swap(next);
    }
    // list starts as idc_wt_encode computes them: prefix of the sizes in bit-reversed order
    std::vector<uint64_t> key(nlist);
    for (uint64_t l = 0; l < nlist; l++) key[l] = ((uint64_t)wt_bitrev((uint32_t)l, sh.levels) << 32) | l;
    std::sort(key.begin(), key.end());
    uint64_t a = 0;
    for (uint64_t i = 0; i < nlist; i++) {
        uint32_t l = (uint32_t)key[i];
        start[l] = (uint32_t)a;
        a += list_size[l];
    }
}
static WtView wt_view_of(uint64_t nlist, uint64_t n, const uint64_t* bits, const uint32_t* rank, const uint32_t* sel1,
                         const uint32_t* sel0, const uint32_t* start) {
    return WtView{bits, rank, sel1, sel0, start, wt_shape(nlist, n)};
}
uint64_t sim_wt_select(uint64_t nlist, uint64_t n, const uint64_t* bits, const uint32_t* rank, const uint32_t* sel1,
                       const uint32_t* sel0, const uint32_t* start, uint32_t c, uint64_t k) {
    return wt_select(wt_view_of(nlist, n, bits, rank, sel1, sel0, start), c, k);
}
uint32_t sim_wt_access(uint64_t nlist, uint64_t n, const uint64_t* bits, const uint32_t* rank, const uint32_t* sel1,
                       const uint32_t* sel0, const uint32_t* start, uint64_t i) {
    return wt_access(wt_view_of(nlist, n, bits, rank, sel1, sel0, start), i);
}
// k_wt_distribute + k_wt_apply CTA by CTA (8 tiles of 512 elements, per-CTA bucket counts, one cursor bump per
// (CTA, bucket), bucket b owning the slots [b << bucket_log, ...) of the pair array). Returns the status bits
// (1 hole / duplicate, 2 range, 4 unsorted); S is pre-filled with 0xffffffff like the device sequence.
uint32_t sim_wt_fill_bucketed(uint64_t nlist, uint64_t n, const int64_t* ids, const uint64_t* list_off, uint32_t bucket_log,
                              uint32_t* S) {
    const uint64_t nbuckets = ((n - 1) >> bucket_log) + 1;
    std::vector<uint64_t> cursor(nbuckets, 0), pairs(n, ~0ull);
    uint32_t st = 0;
    for (uint64_t i = 0; i < n; i++) S[i] = 0xffffffffu;
    const uint64_t per_cta = 8 * 512;
    for (uint64_t cta = 0; cta * per_cta < n; cta++) {
        std::vector<uint32_t> cnt(nbuckets, 0);
        std::vector<uint64_t> base(nbuckets, 0);
        uint64_t e0 = cta * per_cta, e1 = std::min(n, e0 + per_cta);
        std::vector<uint32_t> rank(e1 - e0), own(e1 - e0);
        std::vector<uint64_t> id(e1 - e0);
        for (uint64_t e = e0; e < e1; e++) {
            uint64_t l = std::upper_bound(list_off, list_off + nlist + 1, e) - list_off - 1;  // largest l with off[l] <= e
            own[e - e0] = (uint32_t)l;
            uint64_t v = (uint64_t)ids[e];
            bool bad = v >= n;
            if (bad)
                st |= 2u;
            else if (e > list_off[l] && (uint64_t)ids[e - 1] >= v)
                st |= 4u;
            id[e - e0] = bad ? ~0ull : v;
            if (!bad) rank[e - e0] = cnt[v >> bucket_log]++;
        }
        for (uint64_t b = 0; b < nbuckets; b++)
            if (cnt[b]) {
                base[b] = cursor[b];
                cursor[b] += cnt[b];
            }
        for (uint64_t e = e0; e < e1; e++) {
            uint64_t v = id[e - e0];
            if (v == ~0ull) continue;
            uint64_t b = v >> bucket_log, slot = base[b] + rank[e - e0];
            uint64_t cap = std::min<uint64_t>(1ull << bucket_log, n - (b << bucket_log));
            if (slot < cap)
                pairs[(b << bucket_log) + slot] = (v << 32) | own[e - e0];
            else
                st |= 1u;
        }
    }
    for (uint64_t i = 0; i < n; i++) {  // k_wt_apply
        uint64_t b = i >> bucket_log;
        if (i - (b << bucket_log) >= cursor[b]) continue;
        S[pairs[i] >> 32] = (uint32_t)pairs[i];
    }
    for (uint64_t i = 0; i < n; i++)
        if (S[i] == 0xffffffffu) st |= 1u;  // what level 0 of the build reports
    return st;
}
// k_wt_replay + k_wt_emit lane by lane: the partitions replayed on the ids, bits read back from the structure
void sim_wt_replay_all(uint64_t nlist, uint64_t n, const uint64_t* bits, const uint32_t* rank, const uint32_t* start,
                       const uint64_t* list_off, int64_t* out) {
    WtShape sh = wt_shape(nlist, n);
    std::vector<uint32_t> in(n), nxt(n);
    for (uint64_t i = 0; i < n; i++) in[i] = (uint32_t)i;
    for (uint32_t lev = 0; lev < sh.levels; lev++) {
        const uint32_t* B32 = reinterpret_cast<const uint32_t*>(bits + (uint64_t)lev * sh.words);
        const uint32_t* R = rank + (uint64_t)lev * sh.rank_stride;
        uint64_t z = n - R[sh.nblk];
        for (uint64_t blk = 0; blk < sh.nblk; blk++) {
            uint64_t r1 = R[blk];
            for (int t = 0; t < 16; t++) {
                uint32_t m = B32[blk * 16 + t];
                for (uint32_t lane = 0; lane < 32; lane++) {
                    uint64_t i = (blk << kWtBlockLog) + (uint64_t)t * 32 + lane;
                    uint64_t before = r1 + (uint32_t)__builtin_popcount(m & ((1u << lane) - 1u));
                    if (i < n) nxt[wt_partition_dest(i, (m >> lane) & 1u, z, before)] = in[i];
                }
                r1 += (uint32_t)__builtin_popcount(m);
            }
        }
        in.swap(nxt);
    }
    for (uint64_t l = 0; l < nlist; l++)
        for (uint64_t k = 0; k < list_off[l + 1] - list_off[l]; k++) out[list_off[l] + k] = in[start[l] + k];
}
// k_wt_replay2 + k_wt_emit lane by lane: two levels per pass through the per-tile windows of level v + 1 (32 words per
// side, exclusive popcount prefix on top of the window's directory entry), 32-bit positions -- the kernel's index
// arithmetic restated with plain loops; an odd last level takes the single-level pass
void sim_wt_replay2_all(uint64_t nlist, uint64_t n64, const uint64_t* bits, const uint32_t* rank, const uint32_t* start,
                        const uint64_t* list_off, int64_t* out) {
    WtShape sh = wt_shape(nlist, n64);
    const uint32_t n = (uint32_t)n64, nblk = (uint32_t)sh.nblk;
    std::vector<uint32_t> in(n), nxt(n);
    for (uint32_t i = 0; i < n; i++) in[i] = i;
    for (uint32_t lev = 0; lev < sh.levels;) {
        const uint32_t* B0 = reinterpret_cast<const uint32_t*>(bits + (uint64_t)lev * sh.words);
        const uint32_t* R0 = rank + (uint64_t)lev * sh.rank_stride;
        const uint32_t z0 = n - R0[nblk];
        if (lev + 1 < sh.levels) {
            const uint32_t* B1 = reinterpret_cast<const uint32_t*>(bits + (uint64_t)(lev + 1) * sh.words);
            const uint32_t* R1 = rank + (uint64_t)(lev + 1) * sh.rank_stride;
            const uint32_t z1 = n - R1[nblk], nwords = nblk * 16;
            for (uint32_t blk = 0; blk < nblk; blk++) {
                const uint32_t base = blk << kWtBlockLog, r1 = R0[blk];
                const uint32_t zs = base - r1, os = z0 + r1, zb = zs >> kWtBlockLog, ob = os >> kWtBlockLog;
                uint32_t win[128];
                uint32_t pz = R1[zb <= nblk ? zb : nblk], po = R1[ob <= nblk ? ob : nblk];
                for (uint32_t lane = 0; lane < 32; lane++) {
                    const uint32_t zi = zb * 16 + lane, oi = ob * 16 + lane;
                    win[lane] = zi < nwords ? B1[zi] : 0u;
                    win[32 + lane] = oi < nwords ? B1[oi] : 0u;
                    win[64 + lane] = pz;
                    win[96 + lane] = po;
                    pz += (uint32_t)__builtin_popcount(win[lane]);
                    po += (uint32_t)__builtin_popcount(win[32 + lane]);
                }
                const uint32_t zwin = zb << kWtBlockLog, owin = ob << kWtBlockLog;
                uint32_t ones = r1;
                for (int t = 0; t < 16; t++) {
                    const uint32_t m = B0[blk * 16 + t];
                    for (uint32_t lane = 0; lane < 32; lane++) {
                        const uint32_t i = base + (uint32_t)t * 32 + lane;
                        const uint32_t b0 = (m >> lane) & 1u;
                        const uint32_t before = ones + (uint32_t)__builtin_popcount(m & ((1u << lane) - 1u));
                        const uint32_t p = b0 ? z0 + before : i - before;
                        const uint32_t rel = p - (b0 ? owin : zwin);
                        const uint32_t src = ((rel >> 5) & 31u) + (b0 << 5);
                        const uint32_t x = win[src], y = win[64 + src], shf = rel & 31u;
                        const uint32_t b1 = (x >> shf) & 1u;
                        const uint32_t before1 = y + (uint32_t)__builtin_popcount(x & ((1u << shf) - 1u));
                        if (i < n) nxt[b1 ? z1 + before1 : p - before1] = in[i];
                    }
                    ones += (uint32_t)__builtin_popcount(m);
                }
            }
            lev += 2;
        } else {
            for (uint32_t blk = 0; blk < nblk; blk++) {
                uint32_t r1 = R0[blk];
                for (int t = 0; t < 16; t++) {
                    const uint32_t m = B0[blk * 16 + t];
                    for (uint32_t lane = 0; lane < 32; lane++) {
                        const uint32_t i = (blk << kWtBlockLog) + (uint32_t)t * 32 + lane;
                        const uint32_t before = r1 + (uint32_t)__builtin_popcount(m & ((1u << lane) - 1u));
                        if (i < n) nxt[((m >> lane) & 1u) ? z0 + before : i - before] = in[i];
                    }
                    r1 += (uint32_t)__builtin_popcount(m);
                }
            }
            lev += 1;
        }
        in.swap(nxt);
    }
    for (uint64_t l = 0; l < nlist; l++)
        for (uint64_t k = 0; k < list_off[l + 1] - list_off[l]; k++) out[list_off[l] + k] = in[start[l] + k];
}
// every id of every list in one call (list_off = CSR of the list sizes)
void sim_wt_decode_all(uint64_t nlist, uint64_t n, const uint64_t* bits, const uint32_t* rank, const uint32_t* sel1,
                       const uint32_t* sel0, const uint32_t* start, const uint64_t* list_off, int64_t* out) {
    WtView v = wt_view_of(nlist, n, bits, rank, sel1, sel0, start);
    for (uint64_t l = 0; l < nlist; l++)
        for (uint64_t k = 0; k < list_off[l + 1] - list_off[l]; k++) out[list_off[l] + k] = (int64_t)wt_select(v, (uint32_t)l, k);
}

}  // extern "C"
