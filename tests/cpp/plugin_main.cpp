// TEST PROGRAM: the C++ host side (csrc/plugin/idc_faiss_plugin.h, the reference's plugin classes on top of the C ABI)
// run the way the reference's own tests use the classes -- test_compressed_ivfs.py:26-90 (per list: decoded ids ==
// stored ids as a set, id j pairs with code j, get_single_id), test_altid.py:19-44 (per node: neighbours as a set).
// Then the free functions the SWIG modules bind, through the %inline blocks of the shipped .swig files
// (swig_inline_gen.h is extracted from csrc/plugin/custom_invlists.swig and altid.swig by the test harness):
// search_IVF_defer_id_decoding as in test_compressed_ivfs.py:95-156 (D, I equal to index.search for every flavour,
// decode_1by1, returned codes with list numbers), NSG search with replaced graphs and search_NSG_and_trace as in
// test_altid.py:28-62, BitstringReader_get_bits against BitstringReader::read.
// Faiss is absent from this image: faiss_shim.h declares the base classes, ArrayIL below stands in for
// faiss::ArrayInvertedLists. Needs a GPU (libidcodec.so has no CPU path). Exit code 0 = all checks passed.
#define IDC_FAISS_SHIM
#include "faiss_shim.h"

#include <algorithm>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

#include "idc_faiss_plugin.h"
#include "swig_inline_gen.h"

struct ArrayIL : faiss::InvertedLists {
    std::vector<std::vector<faiss::idx_t>> ids;
    std::vector<std::vector<uint8_t>> codes;
    ArrayIL(size_t nlist, size_t code_size) : InvertedLists(nlist, code_size), ids(nlist), codes(nlist) {}
    size_t list_size(size_t l) const override { return ids[l].size(); }
    const uint8_t* get_codes(size_t l) const override { return codes[l].data(); }
    const faiss::idx_t* get_ids(size_t l) const override { return ids[l].data(); }
    void add(size_t l, faiss::idx_t id) {
        ids[l].push_back(id);
        for (size_t b = 0; b < code_size; b++) codes[l].push_back(code_byte(id, b));
    }
    static uint8_t code_byte(faiss::idx_t id, size_t b) { return (uint8_t)(((uint64_t)id * 2654435761ull + 977 * b) >> 7); }
};

// IVFx,Flat content: list l holds (id, vector) pairs, the code of a vector is the vector
struct VecIL : faiss::InvertedLists {
    std::vector<std::vector<faiss::idx_t>> ids;
    std::vector<std::vector<uint8_t>> codes;
    VecIL(size_t nlist, int d) : InvertedLists(nlist, sizeof(float) * d), ids(nlist), codes(nlist) {}
    size_t list_size(size_t l) const override { return ids[l].size(); }
    const uint8_t* get_codes(size_t l) const override { return codes[l].data(); }
    const faiss::idx_t* get_ids(size_t l) const override { return ids[l].data(); }
    void add(size_t l, faiss::idx_t id, const float* v) {
        ids[l].push_back(id);
        const uint8_t* b = reinterpret_cast<const uint8_t*>(v);
        codes[l].insert(codes[l].end(), b, b + code_size);
    }
};

static int failures = 0;
#define CHECK(cond, ...)                      \
    do {                                      \
        if (!(cond)) {                        \
            failures++;                       \
            std::fprintf(stderr, "FAIL: ");   \
            std::fprintf(stderr, __VA_ARGS__); \
            std::fprintf(stderr, "\n");       \
        }                                     \
    } while (0)

template <class Inv>
static void check_invlists(const char* name, const Inv& inv, const ArrayIL& il, bool same_order, bool single_id) {
    CHECK(inv.nlist == il.nlist && inv.code_size == il.code_size, "%s: shape", name);
    for (size_t l = 0; l < il.nlist; l++) {
        size_t n = inv.list_size(l);
        CHECK(n == il.list_size(l), "%s: list %zu size %zu != %zu", name, l, n, il.list_size(l));
        const faiss::idx_t* got = inv.get_ids(l);
        if (n == 0) {
            continue;
        }
        CHECK(got != nullptr, "%s: list %zu get_ids returned null", name, l);
        if (!got) continue;
        std::vector<faiss::idx_t> a(got, got + n), b(il.ids[l]);
        if (same_order) CHECK(a == b, "%s: list %zu ids differ in order", name, l);
        std::sort(a.begin(), a.end());
        std::sort(b.begin(), b.end());
        CHECK(a == b, "%s: list %zu id set differs", name, l);
        const uint8_t* codes = inv.get_codes(l);
        bool paired = true;
        for (size_t j = 0; j < n && paired; j++)
            for (size_t c = 0; c < il.code_size; c++) paired &= codes[j * il.code_size + c] == ArrayIL::code_byte(got[j], c);
        CHECK(paired, "%s: list %zu code <-> id pairing broken", name, l);
        if (single_id)
            for (size_t j : {size_t(0), n / 2, n - 1})
                CHECK(inv.get_single_id(l, j) == got[j], "%s: get_single_id(%zu, %zu)", name, l, j);
        inv.release_ids(l, got);
    }
    CHECK(inv.compressed_ids_size_in_bytes > 0 && inv.compressed_ids_size_in_bytes < 8 * 1000, "%s: size %zu", name,
          (size_t)inv.compressed_ids_size_in_bytes);
    std::printf("%-44s compressed ids %6zu B, codes %zu B\n", name, (size_t)inv.compressed_ids_size_in_bytes,
                (size_t)inv.codes_size_in_bytes);
}

template <class G>
static void check_graph(const char* name, const G& g, const std::vector<int32_t>& rows, int N, int K, bool returns_k) {
    std::vector<int32_t> buf(K);
    for (int i = 0; i < N; i++) {
        std::vector<int32_t> want;
        for (int j = 0; j < K && rows[(size_t)i * K + j] >= 0; j++) want.push_back(rows[(size_t)i * K + j]);
        std::fill(buf.begin(), buf.end(), -7);
        size_t r = g.get_neighbors(i, buf.data());
        CHECK(r == (returns_k ? (size_t)K : want.size()), "%s: node %d returned %zu", name, i, r);
        std::vector<int32_t> got(buf.begin(), buf.begin() + want.size());
        std::sort(got.begin(), got.end());
        std::sort(want.begin(), want.end());
        CHECK(got == want, "%s: node %d neighbours differ", name, i);
    }
    std::printf("%-44s compressed ids %6zu B\n", name, (size_t)g.compressed_ids_size_in_bytes);
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "";
    const bool all = mode != "--core";       // --core: skip the fixed-width baselines
    const bool host_only = mode == "--host"; // --host: only what needs no device (shim, generic paths, free functions)
    try {
        std::mt19937 rng(4);
        if (!host_only) {
        ArrayIL il(8, 4);  // IVF8 over 1000 vectors, ids 0..999 in add order (ascending per list), list 5 left empty
        for (faiss::idx_t id = 0; id < 1000; id++) {
            size_t l = rng() % 8;
            il.add(l == 5 ? 6 : l, id);
        }
        {
            CompressedIDInvertedListsFenwickTree ft(il);
            check_invlists("CompressedIDInvertedListsFenwickTree", ft, il, false, false);
            CHECK(ft.get_ids(5) == nullptr, "FenwickTree: empty list must return null (custom_invlists_impl.cpp:212-214)");
            // deferred decoding: labels (list << 32 | offset) -> ids, in place (custom_invlists_impl.cpp:464-525)
            const faiss::idx_t* l3 = ft.get_ids(3);
            std::vector<faiss::idx_t> labels = {(3ll << 32) | 5, -1, (3ll << 32) | 0};
            ft.translate_labels(labels.data(), labels.size());
            CHECK(labels[0] == l3[5] && labels[1] == -1 && labels[2] == l3[0], "FenwickTree: translate_labels");
            ft.release_ids(3, l3);
            faiss::idx_t want_lists[2] = {1, 7};
            ft.prefetch_lists(want_lists, 2);
            check_invlists("  ... after prefetch_lists({1, 7})", ft, il, false, false);
        }
        {
            CompressedIDInvertedListsEliasFano ef(il);
            check_invlists("CompressedIDInvertedListsEliasFano", ef, il, true, true);
            CHECK(ef.get_ids(5) == nullptr, "EliasFano: empty list must return null (custom_invlists_impl.cpp:294-296)");
        }
        {
            CompressedIDInvertedListsWaveletTree wt(il, 0);
            check_invlists("CompressedIDInvertedListsWaveletTree", wt, il, true, true);
            CompressedIDInvertedListsWaveletTree rrr(il, 1);  // rrr_vector<63> flavour (custom_invlists_impl.cpp:371-372)
            check_invlists("CompressedIDInvertedListsWaveletTree(wt_type = 1)", rrr, il, true, true);
            bool threw = false;
            try {
                CompressedIDInvertedListsWaveletTree bad(il, 2);
            } catch (const std::exception&) {
                threw = true;
            }
            CHECK(threw, "WaveletTree: wt_type must be 0 or 1 (custom_invlists_impl.cpp:349)");
        }
        if (all) {
            CompressedIDInvertedListsPackedBits pb(il);
            check_invlists("CompressedIDInvertedListsPackedBits", pb, il, true, true);
            CHECK(pb.bits == 10, "PackedBits: bits %d for ntotal 1000", pb.bits);  // (1 << 10) >= 1001
        }
        const int N = 200, K = 16;
        std::vector<int32_t> rows((size_t)N * K, -1);
        for (int i = 0; i < N; i++) {  // distinct neighbours != i, degree 3..16, -1 padded (altid_impl.cpp:110-117)
            int deg = 3 + (int)(rng() % (K - 2));
            std::vector<int32_t> cand;
            while ((int)cand.size() < deg) {
                int32_t v = (int32_t)(rng() % N);
                if (v != i && std::find(cand.begin(), cand.end(), v) == cand.end()) cand.push_back(v);
            }
            std::copy(cand.begin(), cand.end(), rows.begin() + (size_t)i * K);
        }
        if (all) {
            std::vector<int32_t> data(rows);
            faiss::nsg::Graph<int32_t> g(data.data(), N, K);
            CompactBitNSGGraph cg(g);
            CHECK(cg.bits == 8 && cg.stride == 16, "CompactBit: bits %d stride %zu", cg.bits, cg.stride);
            struct Sized { const CompactBitNSGGraph& g; size_t compressed_ids_size_in_bytes;
                           size_t get_neighbors(int i, int32_t* nb) const { return g.get_neighbors(i, nb); } };
            check_graph("CompactBitNSGGraph", Sized{cg, cg.compressed_data.size()}, rows, N, K, false);
        }
        {
            std::vector<int32_t> data(rows);
            faiss::nsg::Graph<int32_t> g(data.data(), N, K);
            EliasFanoNSGGraph eg(g);
            CHECK(eg.overhead_in_bytes == 2 * (size_t)(N * 8 / 8.0), "EliasFanoNSGGraph: overhead %zu", eg.overhead_in_bytes);
            check_graph("EliasFanoNSGGraph", eg, rows, N, K, false);
        }
        {
            std::vector<int32_t> data(rows);
            faiss::nsg::Graph<int32_t> g(data.data(), N, K);
            ROCNSGGraph rg(g);
            check_graph("ROCNSGGraph", rg, rows, N, K, true);
            for (int i = 0; i < N; i++) {
                int deg = 0;
                while (deg < K && rows[(size_t)i * K + deg] >= 0) deg++;
                CHECK((int)rg.num_outgoing_edges[i] == deg, "ROCNSGGraph: num_outgoing_edges[%d]", i);
            }
        }

        }
        // ------------------------------------------------------------------ BitstringReader_get_bits (.cpp:35-58)
        {
            std::vector<uint8_t> str(64);
            for (auto& b : str) b = (uint8_t)rng();
            bool ok = true;
            for (int nbit = 1; nbit <= 64 && ok; nbit++)
                for (size_t i = 0; i + nbit <= str.size() * 8 && ok; i += 1 + rng() % 7) {
                    faiss::BitstringReader a(str.data(), str.size()), b(str.data(), str.size());
                    a.i = i;
                    ok = a.read(nbit) == BitstringReader_get_bits(b, i, nbit) && b.i == 0;
                }
            CHECK(ok, "BitstringReader_get_bits differs from BitstringReader::read");
            std::printf("%-44s ok\n", "BitstringReader_get_bits");
        }
        // ------------------------------------------------------------------ deferred id decoding (.cpp:407-526)
        {
            const int d = 8, nb = 3000, nl = 16, nq = 25, k = 10;
            std::normal_distribution<float> gauss;
            std::vector<float> xb((size_t)nb * d), xq((size_t)nq * d);
            for (auto& v : xb) v = gauss(rng);
            for (auto& v : xq) v = gauss(rng);
            faiss::IndexFlatL2 quant(d);
            quant.add(nl, xb.data());  // the first nl vectors are the centroids
            VecIL il(nl, d);
            {
                std::vector<float> dis(nb);
                std::vector<faiss::idx_t> assign(nb);
                quant.search(nb, xb.data(), 1, dis.data(), assign.data());
                for (int i = 0; i < nb; i++) il.add((size_t)assign[i], i, xb.data() + (size_t)i * d);
            }
            faiss::IndexIVF index(&quant, d, nl);
            index.ntotal = nb;
            index.nprobe = 4;
            index.replace_invlists(&il, false);
            std::vector<float> D0((size_t)nq * k), D((size_t)nq * k);
            std::vector<faiss::idx_t> I0((size_t)nq * k), I((size_t)nq * k);
            index.search(nq, xq.data(), k, D0.data(), I0.data());
            bool threw = false;
            try {
                search_IVF_defer_id_decoding(index, nq, xq.data(), k, D.data(), I.data());
            } catch (const faiss::FaissException&) {
                threw = true;
            }
            CHECK(threw, "search_IVF_defer_id_decoding must insist on parallel_mode == 3 (:420-422)");
            index.parallel_mode = 3;
            auto run = [&](const char* name, faiss::InvertedLists* inv, bool has_single_id) {
                index.replace_invlists(inv, false);
                std::fill(I.begin(), I.end(), -5);
                index.search(nq, xq.data(), k, D.data(), I.data());  // ids through get_ids / release_ids
                CHECK(D == D0 && I == I0, "%s: index.search differs from the uncompressed index", name);
                for (int one = 0; one <= (has_single_id ? 1 : 0); one++) {
                    std::fill(I.begin(), I.end(), -5);
                    search_IVF_defer_id_decoding_untyped(index, nq, xq.data(), k, D.data(), I.data(), one != 0, nullptr, false);
                    CHECK(D == D0 && I == I0, "%s: search_defer_id_decoding(decode_1by1=%d) differs", name, one);
                }
                const size_t cs1 = index.code_size + index.coarse_code_size();
                std::vector<uint8_t> codes((size_t)nq * k * cs1, 0x55);
                search_IVF_defer_id_decoding_untyped(index, nq, xq.data(), k, D.data(), I.data(), false, codes.data(), true);
                bool ok = D == D0 && I == I0 && index.coarse_code_size() == 1;
                for (size_t r = 0; r < (size_t)nq * k && ok; r++) {
                    const uint8_t* c = codes.data() + r * cs1;
                    if (I0[r] < 0) {
                        for (size_t b = 0; b < cs1; b++) ok &= c[b] == 0xff;
                        continue;
                    }
                    // byte 0 = the list of the result, the rest = its stored code = the vector itself
                    faiss::idx_t a;
                    float dis;
                    quant.search(1, xb.data() + (size_t)I0[r] * d, 1, &dis, &a);
                    ok &= c[0] == (uint8_t)a && std::memcmp(c + 1, xb.data() + (size_t)I0[r] * d, index.code_size) == 0;
                }
                CHECK(ok, "%s: returned codes (include_listno) wrong", name);
                std::printf("%-44s search + deferred decoding equal to the uncompressed index\n", name);
            };
            run("  IVF16,Flat uncompressed (generic path)", &il, true);
            if (!host_only) {
            {
                CompressedIDInvertedListsPackedBits inv(il);
                run("  IVF16,Flat PackedBits", &inv, true);
            }
            {
                CompressedIDInvertedListsFenwickTree inv(il);
                run("  IVF16,Flat FenwickTree (ROC)", &inv, false);
                inv.materialize_ans_states();
                size_t sz = 0;
                for (size_t l = 0; l < inv.nlist; l++) sz += inv.list_size(l) ? inv.ans_states[l].size() : 0;
                CHECK(sz == inv.compressed_ids_size_in_bytes, "FenwickTree: sum of ans_states[l].size() %zu != %zu", sz,
                      inv.compressed_ids_size_in_bytes);
                inv.set_cache_budget_ids(500);  // smaller than the probed lists together: bounded, still correct
                std::vector<faiss::idx_t> all(nl);
                for (int l = 0; l < nl; l++) all[l] = l;
                inv.prefetch_lists(all.data(), all.size());
                run("  ... with a 500-id decoded-list cache", &inv, false);
                inv.drop_cache();
            }
            {
                CompressedIDInvertedListsEliasFano inv(il);
                run("  IVF16,Flat EliasFano", &inv, true);
                inv.materialize_ef_bitstreams();
                size_t bits = 0;
                for (auto& e : inv.ef_bitstreams) bits += e.low_bits_size + e.high_bits_size;
                CHECK(bits / 8 == inv.compressed_ids_size_in_bytes, "EliasFano: ef_bitstreams bits %zu vs %zu bytes", bits,
                      inv.compressed_ids_size_in_bytes);
            }
            {
                CompressedIDInvertedListsWaveletTree inv(il, 0);
                run("  IVF16,Flat WaveletTree", &inv, true);
            }
            }
            index.replace_invlists(&il, false);
        }
        // ------------------------------------------------------------------ NSG search over replaced graphs
        {
            const int d = 8, NN = 400, KK = 16, nq = 20, k = 5;
            std::normal_distribution<float> gauss;
            std::vector<float> xb((size_t)NN * d), xq((size_t)nq * d);
            for (auto& v : xb) v = gauss(rng);
            for (auto& v : xq) v = gauss(rng);
            faiss::IndexFlatL2 storage(d);
            storage.add(NN, xb.data());
            // rows: the 6..12 nearest neighbours plus a few random long links, -1 padded
            std::vector<int32_t> g0((size_t)NN * KK, -1);
            {
                std::vector<float> dis((size_t)NN * 13);
                std::vector<faiss::idx_t> nn((size_t)NN * 13);
                storage.search(NN, xb.data(), 13, dis.data(), nn.data());
                for (int i = 0; i < NN; i++) {
                    std::vector<int32_t> row;
                    int near = 6 + (int)(rng() % 7);
                    for (int j = 1; j <= near; j++) row.push_back((int32_t)nn[(size_t)i * 13 + j]);
                    int far = (int)(rng() % 4);
                    while (far-- > 0) {
                        int32_t v = (int32_t)(rng() % NN);
                        if (v != i && std::find(row.begin(), row.end(), v) == row.end()) row.push_back(v);
                    }
                    std::copy(row.begin(), row.end(), g0.begin() + (size_t)i * KK);
                }
            }
            faiss::IndexNSG index(&storage);
            index.ntotal = NN;
            index.nsg.ntotal = NN;
            index.nsg.search_L = 24;
            std::vector<int32_t> data0(g0);
            NSG_replace_final_graph(index.nsg, new faiss::nsg::Graph<int32_t>(data0.data(), NN, KK));
            std::vector<float> D0((size_t)nq * k), D((size_t)nq * k);
            std::vector<faiss::idx_t> I0((size_t)nq * k), I((size_t)nq * k);
            index.search(nq, xq.data(), k, D0.data(), I0.data());
            auto run = [&](const char* name) {
                index.search(nq, xq.data(), k, D.data(), I.data());
                CHECK(D == D0 && I == I0, "%s: NSG search differs from the uncompressed graph", name);
                std::vector<faiss::idx_t> visited;
                std::fill(I.begin(), I.end(), -5);
                search_NSG_and_trace_untyped(index, nq, xq.data(), k, I.data(), D.data(), &visited);
                bool ok = D == D0 && I == I0 && !visited.empty();
                std::vector<faiss::idx_t> vs(visited);
                std::sort(vs.begin(), vs.end());
                for (faiss::idx_t id : I0) ok &= id < 0 || std::binary_search(vs.begin(), vs.end(), id);
                CHECK(ok, "%s: search_NSG_and_trace (results equal, ids within the trace)", name);
                std::printf("%-44s NSG search + trace equal (%zu distance computations)\n", name, visited.size());
            };
            run("  NSG uncompressed");
            if (!host_only) {
            {
                std::vector<int32_t> data(g0);
                faiss::nsg::Graph<int32_t> g(data.data(), NN, KK);
                NSG_replace_final_graph(index.nsg, new CompactBitNSGGraph(g));
                run("  NSG CompactBitNSGGraph");
            }
            for (int cached = 0; cached <= 1; cached++) {
                std::vector<int32_t> data(g0);
                faiss::nsg::Graph<int32_t> g(data.data(), NN, KK);
                auto* eg = new EliasFanoNSGGraph(g);
                if (cached) eg->set_cache_rows(64);
                NSG_replace_final_graph(index.nsg, eg);
                run(cached ? "  NSG EliasFanoNSGGraph, 64-row cache" : "  NSG EliasFanoNSGGraph");
            }
            for (int cached = 0; cached <= 1; cached++) {
                std::vector<int32_t> data(g0);
                faiss::nsg::Graph<int32_t> g(data.data(), NN, KK);
                auto* rg = new ROCNSGGraph(g);
                if (cached) rg->set_cache_rows(64);
                NSG_replace_final_graph(index.nsg, rg);
                run(cached ? "  NSG ROCNSGGraph, 64-row cache" : "  NSG ROCNSGGraph");
                if (cached) {
                    rg->materialize_ans_states();
                    size_t sz = 0;
                    for (int i = 0; i < NN; i++) sz += rg->num_outgoing_edges[i] ? rg->ans_states[i].size() : 0;
                    CHECK(sz == rg->compressed_ids_size_in_bytes, "ROCNSGGraph: sum of ans_states sizes");
                }
            }
            }
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "FAIL: exception: %s\n", e.what());
        return 2;
    }
    std::printf(failures ? "plugin_main: %d check(s) FAILED\n" : "plugin_main: all checks passed\n", failures);
    return failures ? 1 : 0;
}
