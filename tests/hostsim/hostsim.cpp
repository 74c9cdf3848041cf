// TEST INFRASTRUCTURE ONLY. Compiles the lane-level codec functions of
// vector_db_id_compression_b200/csrc (idc_core.cuh, roc_group.cuh, ef_core.cuh)
// with g++ so their logic can be checked against the oracle in the CPU test
// suite (no GPU in the build container). Not part of libidcodec.so; nothing in
// the product loads it.
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../vector_db_id_compression_b200/csrc/roc_group.cuh"
#include "grp_emu.h"
#include "../../vector_db_id_compression_b200/csrc/ef_core.cuh"

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
        dec_ring_prime(U.st);
        U.out = out;
        U.n = n;
        U.prec = prec;
        for (uint32_t i = 0; i < n; i++)
            gdec_step<G>(g, U, i, (uint32_t)((1ull << 31) / (i + 1)), mt, true);
        if (g.sub == 0) *status_out = U.st.status;
    });
}

extern "C" {

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

}  // extern "C"
