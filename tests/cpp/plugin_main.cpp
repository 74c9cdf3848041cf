// TEST PROGRAM: the C++ host side (csrc/plugin/idc_faiss_plugin.h, the reference's plugin classes on top of the C ABI)
// run the way the reference's own tests use the classes -- test_compressed_ivfs.py:26-90 (per list: decoded ids ==
// stored ids as a set, id j pairs with code j, get_single_id), test_altid.py:19-44 (per node: neighbours as a set).
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
    const bool all = argc > 1 && std::string(argv[1]) == "--all";  // also the classes not yet run on a GPU
    try {
        std::mt19937 rng(4);
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
            bool threw = false;
            try {
                CompressedIDInvertedListsWaveletTree rrr(il, 1);
            } catch (const std::exception&) {
                threw = true;
            }
            CHECK(threw, "WaveletTree: wt_type 1 must be rejected (not implemented)");
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
    } catch (const std::exception& e) {
        std::fprintf(stderr, "FAIL: exception: %s\n", e.what());
        return 2;
    }
    std::printf(failures ? "plugin_main: %d check(s) FAILED\n" : "plugin_main: all checks passed\n", failures);
    return failures ? 1 : 0;
}
