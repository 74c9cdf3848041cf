// idc_faiss_plugin.h -- header-only C++ adapter: the reference's Faiss plugin classes on top of the C ABI.
//
// Drop this header next to custom_invlists_impl.h / altid_impl.h in a tree that has Faiss, include it from
// the .swig files instead of the reference's *_impl.h, and link libidcodec.so. Class names, base classes,
// constructor signatures and public data members are the reference's
// (custom_invlist_cpp/custom_invlists_impl.h:22-98, alt-graph-index/altid_impl.h:29-67); the per-list loops
// of the constructors and accessors are replaced by ONE bulk call each into the sm_100a codec.
//
// Without Faiss (this image) the header is compiled against tests/faiss_shim.h, which declares only the
// members used here (tests/test_cabi_cpu.py::test_plugin_header_compiles).
#pragma once

#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "idcodec.h"

#ifndef IDC_FAISS_SHIM
#include <faiss/IndexNSG.h>
#include <faiss/impl/FaissException.h>
#include <faiss/invlists/InvertedLists.h>
#endif

namespace idc_plugin {

inline void check(int rc) {
    if (rc != IDC_OK) throw std::runtime_error(std::string("idcodec: ") + idc_last_error());
}

// one context per process and device, created on first use
inline idc_ctx* context(int device = 0) {
    static std::mutex mu;
    static std::unordered_map<int, idc_ctx*> ctxs;
    std::lock_guard<std::mutex> g(mu);
    auto it = ctxs.find(device);
    if (it != ctxs.end()) return it->second;
    idc_ctx* c = nullptr;
    check(idc_ctx_create(device, &c));
    ctxs[device] = c;
    return c;
}

// CSR copy of a faiss::InvertedLists (ScopedIds per list, custom_invlists_impl.cpp:156-160)
struct Csr {
    std::vector<uint64_t> offsets;
    std::vector<faiss::idx_t> ids;
    bool ascending = true;
    explicit Csr(const faiss::InvertedLists& il) {
        offsets.resize(il.nlist + 1, 0);
        for (size_t l = 0; l < il.nlist; l++) offsets[l + 1] = offsets[l] + il.list_size(l);
        ids.resize(offsets.back());
        for (size_t l = 0; l < il.nlist; l++) {
            size_t ls = il.list_size(l);
            if (!ls) continue;
            faiss::InvertedLists::ScopedIds sids(&il, l);
            std::memcpy(ids.data() + offsets[l], sids.get(), ls * sizeof(faiss::idx_t));
            for (size_t i = 1; i < ls && ascending; i++) ascending = sids[i - 1] <= sids[i];
        }
    }
};

}  // namespace idc_plugin

/// custom_invlists_impl.h:22-33
struct InvertedListsArrayCodes : faiss::ReadOnlyInvertedLists {
    using idx_t = faiss::idx_t;
    std::vector<std::vector<uint8_t>> codes_all;
    explicit InvertedListsArrayCodes(const faiss::InvertedLists& il) : ReadOnlyInvertedLists(il.nlist, il.code_size) {}
    size_t list_size(size_t list_no) const override { return codes_all[list_no].size() / code_size; }
    const uint8_t* get_codes(size_t list_no) const override { return codes_all[list_no].data(); }
    void release_ids(size_t, const idx_t* ids) const override { delete[] ids; }

   protected:
    // copy the codes of list l permuted by order[0..ls) (order == nullptr: identity)
    void take_codes(const faiss::InvertedLists& il, size_t l, const uint32_t* order) {
        size_t ls = il.list_size(l);
        codes_all[l].resize(ls * code_size);
        if (!ls) return;
        faiss::InvertedLists::ScopedCodes codes(&il, l);
        for (size_t t = 0; t < ls; t++)
            std::memcpy(codes_all[l].data() + t * code_size, codes.get() + (order ? order[t] : t) * code_size, code_size);
    }
};

/// ROC-compressed ids. custom_invlists_impl.h:56-70, .cpp:133-223
struct CompressedIDInvertedListsFenwickTree : InvertedListsArrayCodes {
    idc_roc_blob* blob = nullptr;
    size_t compressed_ids_size_in_bytes = 0;
    size_t codes_size_in_bytes = 0;
    std::vector<uint64_t> id_symbol_precision;
    size_t overhead_in_bytes = 0;

    explicit CompressedIDInvertedListsFenwickTree(const faiss::InvertedLists& il) : InvertedListsArrayCodes(il) {
        idc_plugin::Csr csr(il);
        uint32_t flags = IDC_F_WANT_ORDER | (csr.ascending ? IDC_F_SORTED : 0u);
        idc_plugin::check(idc_roc_encode(idc_plugin::context(), nlist, csr.offsets.data(), csr.ids.data(), 8, IDC_MEM_HOST,
                                         flags, IDC_MAX_UNIT_DEFAULT, &blob));
        std::vector<uint32_t> order(csr.ids.size() + 1);
        idc_plugin::check(idc_roc_blob_order(blob, order.data(), IDC_MEM_HOST));
        codes_all.resize(nlist);
        for (size_t l = 0; l < nlist; l++) {  // codes in sample order, custom_invlists_impl.cpp:189-193
            take_codes(il, l, order.data() + csr.offsets[l]);
            codes_size_in_bytes += codes_all[l].size();
        }
        idc_roc_info info;
        idc_plugin::check(idc_roc_blob_info(blob, &info));
        compressed_ids_size_in_bytes = info.ans_bytes;  // sum of ANSState::size(), :199-202
        std::vector<uint64_t> unit_off(nlist + 1);
        std::vector<uint8_t> prec(info.nunits + 1);
        idc_plugin::check(idc_roc_blob_export(blob, nullptr, unit_off.data(), nullptr, prec.data(), nullptr, nullptr, nullptr));
        id_symbol_precision.resize(nlist);
        for (size_t l = 0; l < nlist; l++) id_symbol_precision[l] = prec[unit_off[l]];
    }
    ~CompressedIDInvertedListsFenwickTree() override { idc_roc_blob_free(blob); }

    /// bulk decode of the lists a search is about to touch; get_ids then serves from the cache
    void prefetch_lists(const idx_t* list_nos, size_t n) const {
        std::vector<uint64_t> ln(list_nos, list_nos + n), off(n + 1);
        size_t total = 0;
        for (size_t i = 0; i < n; i++) total += list_size(ln[i]);
        std::vector<idx_t> ids(total + 1);
        idc_plugin::check(idc_roc_decode(idc_plugin::context(), blob, ln.data(), n, ids.data(), 8, IDC_MEM_HOST, off.data()));
        std::lock_guard<std::mutex> g(mu_);
        for (size_t i = 0; i < n; i++) cache_[ln[i]].assign(ids.begin() + off[i], ids.begin() + off[i + 1]);
    }

    /// the id-translation step of search_IVF_defer_id_decoding (:464-525), in place: labels hold
    /// (list_no << 32 | offset) from search_preassigned(store_pairs = true), -1 for empty slots
    void translate_labels(idx_t* labels, size_t n) const {
        static_assert(sizeof(idx_t) == 8, "faiss::idx_t is int64");
        idc_plugin::check(idc_roc_translate(idc_plugin::context(), blob, reinterpret_cast<const int64_t*>(labels), IDC_MEM_HOST, n,
                                            reinterpret_cast<int64_t*>(labels), IDC_MEM_HOST));
    }

    const idx_t* get_ids(size_t list_no) const override {  // :210-219
        size_t ls = list_size(list_no);
        if (ls == 0) return nullptr;
        idx_t* the_ids = new idx_t[ls];
        {
            std::lock_guard<std::mutex> g(mu_);
            auto it = cache_.find(list_no);
            if (it != cache_.end()) {
                std::memcpy(the_ids, it->second.data(), ls * sizeof(idx_t));
                return the_ids;
            }
        }
        uint64_t ln = list_no;
        idc_plugin::check(idc_roc_decode(idc_plugin::context(), blob, &ln, 1, the_ids, 8, IDC_MEM_HOST, nullptr));
        return the_ids;
    }

   private:
    mutable std::mutex mu_;
    mutable std::unordered_map<size_t, std::vector<idx_t>> cache_;
};

/// Elias-Fano ids. custom_invlists_impl.h:72-98, .cpp:229-339
struct CompressedIDInvertedListsEliasFano : InvertedListsArrayCodes {
    idc_ef_blob* blob = nullptr;
    size_t overhead_in_bytes = 0;
    size_t compressed_ids_size_in_bytes = 0;
    size_t codes_size_in_bytes = 0;

    explicit CompressedIDInvertedListsEliasFano(const faiss::InvertedLists& il) : InvertedListsArrayCodes(il) {
        idc_plugin::Csr csr(il);
        codes_all.resize(nlist);
        std::vector<uint32_t> perm;
        for (size_t l = 0; l < nlist; l++) {  // canonicalize_order_inplace, :324-339
            size_t ls = il.list_size(l);
            perm.resize(ls);
            for (size_t i = 0; i < ls; i++) perm[i] = (uint32_t)i;
            faiss::idx_t* ids = csr.ids.data() + csr.offsets[l];
            if (!csr.ascending) {
                std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return ids[a] < ids[b]; });
                std::vector<faiss::idx_t> tmp(ls);
                for (size_t i = 0; i < ls; i++) tmp[i] = ids[perm[i]];
                std::memcpy(ids, tmp.data(), ls * sizeof(faiss::idx_t));
            }
            take_codes(il, l, csr.ascending ? nullptr : perm.data());
            codes_size_in_bytes += codes_all[l].size();
        }
        idc_plugin::check(idc_ef_encode(idc_plugin::context(), nlist, csr.offsets.data(), csr.ids.data(), 8, IDC_MEM_HOST,
                                        IDC_F_SORTED, &blob));
        idc_ef_info info;
        idc_plugin::check(idc_ef_blob_info(blob, &info));
        compressed_ids_size_in_bytes = info.bits_total / 8;  // :277,282
    }
    ~CompressedIDInvertedListsEliasFano() override { idc_ef_blob_free(blob); }

    const idx_t* get_ids(size_t list_no) const override {  // :292-311
        size_t ls = list_size(list_no);
        if (ls == 0) return nullptr;
        idx_t* the_ids = new idx_t[ls];
        uint64_t ln = list_no;
        idc_plugin::check(idc_ef_decode(idc_plugin::context(), blob, &ln, 1, the_ids, 8, IDC_MEM_HOST, nullptr));
        return the_ids;
    }
    idx_t get_single_id(size_t list_no, size_t offset) const override {  // :314-318
        uint64_t ln = list_no, of = offset;
        int64_t id = -1;
        idc_plugin::check(idc_ef_select(idc_plugin::context(), blob, &ln, &of, 1, IDC_MEM_HOST, &id, IDC_MEM_HOST));
        return id;
    }
    /// many (list, offset) pairs at once: the decode_1by1 branch of search_IVF_defer_id_decoding, :465-475
    void get_single_ids(const uint64_t* list_nos, const uint64_t* offsets, size_t n, int64_t* out) const {
        idc_plugin::check(idc_ef_select(idc_plugin::context(), blob, list_nos, offsets, n, IDC_MEM_HOST, out, IDC_MEM_HOST));
    }
};

/// Fixed-width ids (the baseline of every result table). custom_invlists_impl.h:37-53, .cpp:62-118.
/// NOTE: compiled and linked (tests/test_cabi_cpu.py), exercised by `plugin_main --all`; not yet run on a GPU.
struct CompressedIDInvertedListsPackedBits : InvertedListsArrayCodes {
    int bits = 0;
    std::vector<std::vector<uint8_t>> ids_all;
    size_t compressed_ids_size_in_bytes = 0;
    size_t codes_size_in_bytes = 0;
    size_t overhead_in_bytes = 0;

    explicit CompressedIDInvertedListsPackedBits(const faiss::InvertedLists& il) : InvertedListsArrayCodes(il) {
        idc_plugin::Csr csr(il);
        size_t ntotal = csr.ids.size();  // il.compute_ntotal(), :66
        while ((1ull << bits) < ntotal + 1) bits++;  // :67
        codes_all.resize(nlist);
        ids_all.resize(nlist);
        for (size_t l = 0; l < nlist; l++) {
            size_t ls = il.list_size(l);
            const faiss::idx_t* ids = csr.ids.data() + csr.offsets[l];
            for (size_t i = 0; i < ls; i++)
                if (ids[i] < 0 || (size_t)ids[i] >= ntotal)
                    throw std::runtime_error("Error: 'ids_in[i] >= 0 && ids_in[i] < ntotal' failed");  // :87
            ids_all[l].resize((ls * bits + 7) / 8);  // one BitstringWriter per list, :82-84
            if (ls)
                idc_plugin::check(idc_bits_pack(idc_plugin::context(), ls, ids, 8, IDC_MEM_HOST, bits, ids_all[l].data(),
                                                ids_all[l].size(), IDC_MEM_HOST));
            compressed_ids_size_in_bytes += ids_all[l].size();
            take_codes(il, l, nullptr);
            codes_size_in_bytes += codes_all[l].size();
        }
    }

    const idx_t* get_ids(size_t list_no) const override {  // :96-106
        size_t ls = list_size(list_no);
        idx_t* the_ids = new idx_t[ls];
        if (ls)
            idc_plugin::check(idc_bits_unpack(idc_plugin::context(), ls, ids_all[list_no].data(), ids_all[list_no].size(),
                                              IDC_MEM_HOST, bits, the_ids, 8, IDC_MEM_HOST));
        return the_ids;
    }
    idx_t get_single_id(size_t list_no, size_t offset) const override {  // BitstringReader_get_bits, :35-58,109-114
        const std::vector<uint8_t>& code = ids_all[list_no];
        uint64_t v = 0;
        for (size_t b = 0, pos = offset * bits; b < (size_t)bits; b++, pos++)
            v |= (uint64_t)((code[pos >> 3] >> (pos & 7)) & 1u) << b;
        return (idx_t)v;
    }
};

/// Wavelet-tree ids. custom_invlists_impl.h:100-124, .cpp:346-397: one structure over S[id] = list_no,
/// get_single_id(list_no, offset) = wt.select(offset + 1, list_no). wt_type 1 (rrr_vector<63>) throws: not implemented.
struct CompressedIDInvertedListsWaveletTree : InvertedListsArrayCodes {
    idc_wt_blob* blob = nullptr;
    int wt_type = 0;
    size_t overhead_in_bytes = 0;
    size_t compressed_ids_size_in_bytes = 0;
    size_t codes_size_in_bytes = 0;

    explicit CompressedIDInvertedListsWaveletTree(const faiss::InvertedLists& il, int wt_type = 0)
            : InvertedListsArrayCodes(il), wt_type(wt_type) {
        idc_plugin::Csr csr(il);  // ids ascending per list and < ntotal (asserts :358-359): verified on the device
        codes_all.resize(nlist);
        for (size_t l = 0; l < nlist; l++) {
            take_codes(il, l, nullptr);  // :363-364
            codes_size_in_bytes += codes_all[l].size();
        }
        idc_plugin::check(idc_wt_encode(idc_plugin::context(), nlist, csr.offsets.data(), csr.ids.data(), 8, IDC_MEM_HOST,
                                        wt_type, &blob));
        idc_wt_info info;
        idc_plugin::check(idc_wt_blob_info(blob, &info));
        compressed_ids_size_in_bytes = info.bits_bytes + info.aux_bytes;  // :368,371 size_in_bytes(wt), this layout's
    }
    ~CompressedIDInvertedListsWaveletTree() override { idc_wt_blob_free(blob); }

    idx_t get_single_id(size_t list_no, size_t offset) const override {  // :377-379
        uint64_t ln = list_no, of = offset;
        int64_t id = -1;
        idc_plugin::check(idc_wt_select(idc_plugin::context(), blob, &ln, &of, 1, IDC_MEM_HOST, &id, IDC_MEM_HOST));
        return id;
    }
    const idx_t* get_ids(size_t list_no) const override {  // :381-392, one bulk call instead of ls selects
        size_t ls = list_size(list_no);
        idx_t* the_ids = new idx_t[ls];
        uint64_t ln = list_no;
        idc_plugin::check(idc_wt_decode(idc_plugin::context(), blob, &ln, 1, the_ids, 8, IDC_MEM_HOST, nullptr));
        return the_ids;
    }
    void get_single_ids(const uint64_t* list_nos, const uint64_t* offsets, size_t n, int64_t* out) const {
        idc_plugin::check(idc_wt_select(idc_plugin::context(), blob, list_nos, offsets, n, IDC_MEM_HOST, out, IDC_MEM_HOST));
    }
};

/// Fixed-width edges, N marks the end of a row. altid_impl.h:29-39, .cpp:20-51.
/// NOTE: compiled and linked, exercised by `plugin_main --all`; not yet run on a GPU.
struct CompactBitNSGGraph : faiss::nsg::Graph<int32_t> {
    int bits = 0;
    size_t stride = 0;
    std::vector<uint8_t> compressed_data;
    explicit CompactBitNSGGraph(const faiss::nsg::Graph<int32_t>& graph) : faiss::nsg::Graph<int32_t>(graph.data, graph.N, graph.K) {
        while ((1 << bits) < N + 1) bits++;  // :22-23
        stride = ((size_t)K * bits + 7) / 8;
        compressed_data.assign((size_t)N * stride, 0);
        std::vector<int32_t> vals((size_t)N * K, 0);  // the row up to and including its end marker; zeros behind it
        for (size_t i = 0; i < (size_t)N; i++)
            for (size_t j = 0; j < (size_t)K; j++) {
                int32_t v = graph.data[i * K + j];
                vals[i * K + j] = v == -1 ? N : v;
                if (v == -1) break;
            }
        if (stride * 8 == (size_t)K * bits) {  // rows are byte-aligned: the whole graph is one packed string
            idc_plugin::check(idc_bits_pack(idc_plugin::context(), (uint64_t)N * K, vals.data(), 4, IDC_MEM_HOST, bits,
                                            compressed_data.data(), compressed_data.size(), IDC_MEM_HOST));
        } else {  // one BitstringWriter per row (:27), each padded to `stride` bytes
            for (size_t i = 0; i < (size_t)N; i++)
                idc_plugin::check(idc_bits_pack(idc_plugin::context(), K, vals.data() + i * K, 4, IDC_MEM_HOST, bits,
                                                compressed_data.data() + i * stride, stride, IDC_MEM_HOST));
        }
        data = nullptr;  // :38
    }
    size_t get_neighbors(int i, int32_t* neighbors) const override {  // :41-51
        std::vector<int32_t> row(K);
        idc_plugin::check(idc_bits_unpack(idc_plugin::context(), K, compressed_data.data() + (size_t)i * stride, stride,
                                          IDC_MEM_HOST, bits, row.data(), 4, IDC_MEM_HOST));
        for (int j = 0; j < K; j++) {
            if (row[j] == N) return j;
            neighbors[j] = row[j];
        }
        return K;
    }
};

/// altid_impl.h:42-50, .cpp:53-101
struct EliasFanoNSGGraph : faiss::nsg::Graph<int32_t> {
    idc_ef_blob* blob = nullptr;
    size_t compressed_ids_size_in_bytes = 0;
    size_t overhead_in_bytes = 0;
    explicit EliasFanoNSGGraph(const faiss::nsg::Graph<int32_t>& graph) : faiss::nsg::Graph<int32_t>(graph.data, graph.N, graph.K) {
        idc_plugin::check(idc_ef_encode_rows(idc_plugin::context(), N, K, graph.data, IDC_MEM_HOST, 0, &blob));
        idc_ef_info info;
        idc_plugin::check(idc_ef_blob_info(blob, &info));
        compressed_ids_size_in_bytes = info.bits_total / 8;
        overhead_in_bytes = 2 * (size_t)(N * std::ceil(std::log2((double)N)) / 8.0);  // :56-57 list sizes + max ids
        data = nullptr;  // :89
    }
    ~EliasFanoNSGGraph() override { idc_ef_blob_free(blob); }
    size_t get_neighbors(int i, int32_t* neighbors) const override {  // :92-101
        std::vector<int32_t> row(K);
        uint32_t cnt = 0;
        int32_t r = i;
        idc_plugin::check(idc_ef_decode_rows(idc_plugin::context(), blob, &r, IDC_MEM_HOST, 1, row.data(), &cnt, IDC_MEM_HOST));
        std::memcpy(neighbors, row.data(), cnt * sizeof(int32_t));
        return cnt;
    }
};

/// altid_impl.h:53-67, .cpp:103-165
struct ROCNSGGraph : faiss::nsg::Graph<int32_t> {
    idc_roc_blob* blob = nullptr;
    std::vector<uint64_t> id_symbol_precision;
    size_t compressed_ids_size_in_bytes = 0;
    std::vector<uint32_t> num_outgoing_edges;
    size_t overhead_in_bytes = 0;
    explicit ROCNSGGraph(const faiss::nsg::Graph<int32_t>& graph) : faiss::nsg::Graph<int32_t>(graph.data, graph.N, graph.K) {
        idc_plugin::check(idc_roc_encode_rows(idc_plugin::context(), N, K, graph.data, IDC_MEM_HOST, 0, &blob));
        idc_roc_info info;
        idc_plugin::check(idc_roc_blob_info(blob, &info));
        compressed_ids_size_in_bytes = info.ans_bytes;  // :148
        overhead_in_bytes = (size_t)(N * std::ceil(std::log2((double)N)) / 8.0);  // :106
        num_outgoing_edges.resize(N);
        std::vector<uint8_t> prec(N);
        idc_plugin::check(idc_roc_blob_export(blob, nullptr, nullptr, num_outgoing_edges.data(), prec.data(), nullptr, nullptr, nullptr));
        id_symbol_precision.assign(prec.begin(), prec.end());
        data = nullptr;  // :150
    }
    ~ROCNSGGraph() override { idc_roc_blob_free(blob); }
    size_t get_neighbors(int node, int32_t* neighbors) const override {  // :153-165
        std::vector<int32_t> row(K);
        uint32_t cnt = 0;
        int32_t r = node;
        idc_plugin::check(idc_roc_decode_rows(idc_plugin::context(), blob, &r, IDC_MEM_HOST, 1, row.data(), &cnt, IDC_MEM_HOST));
        std::memcpy(neighbors, row.data(), cnt * sizeof(int32_t));
        return K;  // sic: the reference returns K, not the row length (:164)
    }
};
