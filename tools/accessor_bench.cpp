// accessor_bench -- the per-call path the reference's drivers time, through the C++ adapter
// (csrc/plugin/idc_faiss_plugin.h over the C ABI; Faiss replaced by tests/faiss_shim.h):
//   (i)   single get_neighbors calls on EliasFanoNSGGraph / ROCNSGGraph (altid_impl.cpp:92-101,153-165): random rows
//         without a cache, a graph walk (next row = a neighbour of this one, what NSG search does) with the row
//         cache, and the same rows through one bulk get_neighbors_batch call;
//   (ii)  get_ids per list (custom_invlists_impl.cpp:210-223,292-311) vs prefetch_lists + get_ids;
//   (iii) the id translation of search_IVF_defer_id_decoding for nq * k = 10^4 * 100 labels (:464-525).
// Prints one JSON line. bench.py runs it and puts the CPU reference's per-row / per-list decode time beside it.
#define IDC_FAISS_SHIM
#include "faiss_shim.h"

#include <chrono>
#include <cstdio>
#include <random>
#include <vector>

#include "idc_faiss_plugin.h"

struct ArrayIL : faiss::InvertedLists {
    std::vector<std::vector<faiss::idx_t>> ids;
    std::vector<std::vector<uint8_t>> codes;
    ArrayIL(size_t nlist, size_t code_size) : InvertedLists(nlist, code_size), ids(nlist), codes(nlist) {}
    size_t list_size(size_t l) const override { return ids[l].size(); }
    const uint8_t* get_codes(size_t l) const override { return codes[l].data(); }
    const faiss::idx_t* get_ids(size_t l) const override { return ids[l].data(); }
};

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <class G>
static void graph_legs(const char* name, const std::vector<int32_t>& rows, int N, int K, std::mt19937& rng) {
    std::vector<int32_t> data(rows);
    faiss::nsg::Graph<int32_t> g0(data.data(), N, K);
    double t0 = now();
    G g(g0);
    double t_ctor = now() - t0;
    std::vector<int32_t> buf(K);
    const int calls = 10000;
    // random rows, no cache: one launch + copies per call
    g.get_neighbors(0, buf.data());
    t0 = now();
    long sink = 0;
    for (int c = 0; c < calls; c++) sink += g.get_neighbors((int)(rng() % N), buf.data()) + buf[0];
    double t_rand = now() - t0;
    // graph walk with the row cache: a miss decodes the row and all its neighbours' rows in one call
    g.set_cache_rows(8192);
    int cur = 0;
    t0 = now();
    for (int c = 0; c < calls; c++) {
        g.get_neighbors(cur, buf.data());
        int deg = 0;
        while (deg < K && buf[deg] >= 0) deg++;
        cur = deg ? buf[rng() % deg] : (int)(rng() % N);
    }
    double t_walk = now() - t0;
    g.drop_cache();
    // the same amount of rows through ONE bulk call
    const int nb = 100000;
    std::vector<int32_t> sel(nb), out((size_t)nb * K);
    std::vector<uint32_t> cnt(nb);
    for (auto& s : sel) s = (int32_t)(rng() % N);
    g.get_neighbors_batch(sel.data(), 1000, out.data(), cnt.data());
    t0 = now();
    g.get_neighbors_batch(sel.data(), nb, out.data(), cnt.data());
    double t_batch = now() - t0;
    std::printf("\"%s\": {\"rows\": %d, \"K\": %d, \"ctor_ms\": %.2f, \"get_neighbors_us_per_call\": %.2f, "
                "\"walk_cached_us_per_call\": %.2f, \"batch_rows\": %d, \"batch_us_per_row\": %.4f, \"sink\": %ld}",
                name, N, K, 1e3 * t_ctor, 1e6 * t_rand / calls, 1e6 * t_walk / calls, nb, 1e6 * t_batch / nb, sink % 7);
}

int main() {
    try {
        std::mt19937 rng(9);
        std::printf("{");
        // ---------------- (i) graphs
        const int N = 200000, K = 64;
        std::vector<int32_t> rows((size_t)N * K, -1);
        for (int i = 0; i < N; i++) {
            int deg = (rng() % 10) ? K : 16 + (int)(rng() % (K - 16));
            int32_t* r = rows.data() + (size_t)i * K;
            for (int j = 0; j < deg;) {
                int32_t v = (int32_t)(rng() % N);
                bool dup = v == i;
                for (int t = 0; t < j && !dup; t++) dup = r[t] == v;
                if (!dup) r[j++] = v;
            }
        }
        graph_legs<EliasFanoNSGGraph>("EliasFanoNSGGraph", rows, N, K, rng);
        std::printf(", ");
        graph_legs<ROCNSGGraph>("ROCNSGGraph", rows, N, K, rng);
        // ---------------- (ii) inverted lists: IVF1024 over 10^6 ids
        const size_t nlist = 1024, nb = 1000000, cs = 8;
        ArrayIL il(nlist, cs);
        for (size_t id = 0; id < nb; id++) {
            size_t l = rng() % nlist;
            il.ids[l].push_back((faiss::idx_t)id);
            for (size_t b = 0; b < cs; b++) il.codes[l].push_back((uint8_t)(id >> (8 * (b & 3))));
        }
        const int nl = 256;
        std::vector<faiss::idx_t> lists(nl);
        for (auto& l : lists) l = rng() % nlist;
        {
            double t0 = now();
            CompressedIDInvertedListsFenwickTree ft(il);
            double t_ctor = now() - t0;
            ft.set_cache_budget_ids(0);
            ft.release_ids(0, ft.get_ids(0));
            t0 = now();
            for (int i = 0; i < nl; i++) ft.release_ids(lists[i], ft.get_ids(lists[i]));
            double t_single = now() - t0;
            ft.set_cache_budget_ids(size_t(1) << 24);
            t0 = now();
            ft.prefetch_lists(lists.data(), nl);
            for (int i = 0; i < nl; i++) ft.release_ids(lists[i], ft.get_ids(lists[i]));
            double t_pref = now() - t0;
            // (iii) deferred translation: nq * k = 10^4 * 100 (list, offset) pairs, 2 % empty slots
            const size_t nlab = 1000000;
            std::vector<faiss::idx_t> labels(nlab), want(nlab);
            for (size_t i = 0; i < nlab; i++) {
                size_t l = rng() % nlist, o = rng() % il.ids[l].size();
                labels[i] = (rng() % 50) ? (faiss::idx_t)faiss::lo_build(l, o) : -1;
            }
            ft.drop_cache();
            std::vector<faiss::idx_t> lab2(labels);
            ft.translate_labels(lab2.data(), 1000);
            lab2 = labels;
            t0 = now();
            ft.translate_labels(lab2.data(), nlab);
            double t_tr = now() - t0;
            std::printf(", \"FenwickTree_IVF1024\": {\"ids\": %zu, \"ctor_ms\": %.2f, \"get_ids_us_per_list\": %.2f, "
                        "\"prefetch_then_get_ids_us_per_list\": %.2f, \"translate_labels\": %zu, \"translate_ms\": %.2f, "
                        "\"translate_labels_per_s\": %.3e}",
                        nb, 1e3 * t_ctor, 1e6 * t_single / nl, 1e6 * t_pref / nl, nlab, 1e3 * t_tr, nlab / t_tr);
            CompressedIDInvertedListsEliasFano ef(il);
            ef.release_ids(0, ef.get_ids(0));
            t0 = now();
            for (int i = 0; i < nl; i++) ef.release_ids(lists[i], ef.get_ids(lists[i]));
            double t_ef = now() - t0;
            std::vector<faiss::idx_t> lab3(labels);
            t0 = now();
            idc_plugin::translate_pairs(&ef, lab3.data(), nlab, true);
            double t_efsel = now() - t0;
            bool same = lab3.size() == lab2.size();
            // ROC stores a list in its own order, EF in id order: compare as sets per hit list is overkill here; both
            // must map empty slots to -1 and everything else to an id of the right list
            for (size_t i = 0; i < nlab && same; i++) same = (labels[i] < 0) == (lab2[i] < 0) && (labels[i] < 0) == (lab3[i] < 0);
            std::printf(", \"EliasFano_IVF1024\": {\"get_ids_us_per_list\": %.2f, \"select_labels\": %zu, \"select_ms\": %.2f, "
                        "\"select_labels_per_s\": %.3e, \"consistent\": %s}",
                        1e6 * t_ef / nl, nlab, 1e3 * t_efsel, nlab / t_efsel, same ? "true" : "false");
        }
        std::printf("}\n");
    } catch (const std::exception& e) {
        std::fprintf(stderr, "accessor_bench: %s\n", e.what());
        return 2;
    }
    return 0;
}
