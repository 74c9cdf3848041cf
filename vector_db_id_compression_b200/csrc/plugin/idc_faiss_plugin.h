// idc_faiss_plugin.h -- header-only C++ adapter: the reference's Faiss plugin classes on top of the C ABI.
//
// Drop this header next to custom_invlists_impl.h / altid_impl.h in a tree that has Faiss, include it from
// the .swig files instead of the reference's *_impl.h, and link libidcodec.so. Class names, base classes,
// constructor signatures and public data members are the reference's
// (custom_invlist_cpp/custom_invlists_impl.h:22-98, alt-graph-index/altid_impl.h:29-67); the per-list loops
// of the constructors and accessors are replaced by ONE bulk call each into the sm_100a codec.
//
// The free functions the SWIG modules bind are here too, with the reference's signatures:
//   BitstringReader_get_bits        custom_invlists_impl.h:19,  .cpp:35-58
//   search_IVF_defer_id_decoding    custom_invlists_impl.h:130-139, .cpp:407-526
//   search_NSG_and_trace            altid_impl.h:18-25, .cpp:170-231
// so custom_invlists.swig / altid.swig (shipped next to this header) compile against it unchanged in shape.
//
// Without Faiss (this image) the header is compiled against tests/faiss_shim.h, which declares only the
// members used here (tests/test_cabi_cpu.py::test_plugin_header_compiles) and is enough to RUN the whole
// surface in tests/cpp/plugin_main.cpp.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstring>
#include <list>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "idcodec.h"

#ifndef IDC_FAISS_SHIM
#include <faiss/IndexIVF.h>
#include <faiss/IndexNSG.h>
#include <faiss/impl/AuxIndexStructures.h>
#include <faiss/impl/DistanceComputer.h>
#include <faiss/impl/FaissAssert.h>
#include <faiss/impl/FaissException.h>
#include <faiss/invlists/DirectMap.h>
#include <faiss/invlists/InvertedLists.h>
#include <faiss/utils/hamming.h>
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

// Host copies of the compressed structures, for callers that want to look at them the way the reference's public
// members allow (ans_states: custom_invlists_impl.h:59, altid_impl.h:58; ef_bitstreams: custom_invlists_impl.h:76,
// altid_impl.h:43). The device blob is the storage; these are filled on demand by materialize_*().
struct AnsStateView {  // ANSState, codec.h:13-45
    uint64_t head = 1ull << 31;
    std::vector<uint32_t> stack;
    size_t size() const { return sizeof(head) + stack.size() * sizeof(uint32_t); }  // codec.h:42-44
};
struct EliasFanoView {  // succinct::elias_fano as modified by the reference (elias_fano.hpp:264-276)
    uint64_t num_elements = 0;
    uint64_t universe = 0;  // the builder's n = max id
    uint8_t l = 0;          // m_l
    std::vector<uint64_t> m_low_bits, m_high_bits;  // LSB-first 64-bit words
    uint64_t low_bits_size = 0, high_bits_size = 0; // m_low_bits.size(), m_high_bits.size() in bits
};

inline std::vector<AnsStateView> ans_states_of(const idc_roc_blob* blob) {
    idc_roc_info info;
    check(idc_roc_blob_info(blob, &info));
    std::vector<uint64_t> heads(info.nunits + 1), woff(info.nunits + 1), uoff(info.nlist + 1);
    std::vector<uint32_t> words(info.total_words + 1), un(info.nunits + 1);
    check(idc_roc_blob_export(blob, nullptr, uoff.data(), un.data(), nullptr, heads.data(), woff.data(), words.data()));
    std::vector<AnsStateView> out(info.nlist);
    for (uint64_t l = 0; l < info.nlist; l++) {
        const uint64_t u = uoff[l];  // lists of <= 65536 ids are exactly one unit = the reference's ans_states[l]
        if (uoff[l + 1] != u + 1) throw std::runtime_error("ans_states: list is split into several units (> 65536 ids)");
        if (un[u] == 0) continue;  // an empty list keeps the default state (custom_invlists_impl.cpp:152-154)
        out[l].head = heads[u];
        out[l].stack.assign(words.begin() + woff[u], words.begin() + woff[u + 1]);
    }
    return out;
}

inline std::vector<EliasFanoView> ef_bitstreams_of(const idc_ef_blob* blob) {
    idc_ef_info info;
    check(idc_ef_blob_info(blob, &info));
    std::vector<uint64_t> loff(info.nlist + 1), uni(info.nlist + 1), lo(info.nlist + 1), ho(info.nlist + 1);
    std::vector<uint64_t> low(info.low_words + 1), high(info.high_words + 1);
    std::vector<uint8_t> l(info.nlist + 1);
    check(idc_ef_blob_export(blob, loff.data(), l.data(), uni.data(), lo.data(), ho.data(), low.data(), high.data()));
    std::vector<EliasFanoView> out(info.nlist);
    for (uint64_t i = 0; i < info.nlist; i++) {
        EliasFanoView& v = out[i];
        v.num_elements = loff[i + 1] - loff[i];
        if (!v.num_elements) continue;
        v.universe = uni[i];
        v.l = l[i];
        v.m_low_bits.assign(low.begin() + lo[i], low.begin() + lo[i + 1]);
        v.m_high_bits.assign(high.begin() + ho[i], high.begin() + ho[i + 1]);
        v.low_bits_size = v.num_elements * v.l;                                      // elias_fano.hpp:40-42
        v.high_bits_size = (v.num_elements + 1) + (v.universe >> v.l) + 1;           // elias_fano.hpp:29
    }
    return out;
}

// Bounded cache of decoded lists / rows keyed by number (least recently inserted goes first). The Faiss virtuals hand
// out one list / one row per call; a kernel launch per call cannot win, so bulk decodes (prefetch_*) park their
// results here and the per-call accessors serve from it. `budget` = entries of payload kept at most.
template <class T>
struct DecodedCache {
    size_t budget = 0, held = 0;
    std::unordered_map<size_t, std::vector<T>> map;
    std::list<size_t> order;
    mutable std::mutex mu;
    bool lookup(size_t key, T* out, size_t* n) const {
        std::lock_guard<std::mutex> g(mu);
        auto it = map.find(key);
        if (it == map.end()) return false;
        std::memcpy(out, it->second.data(), it->second.size() * sizeof(T));
        if (n) *n = it->second.size();
        return true;
    }
    bool contains(size_t key) const {
        std::lock_guard<std::mutex> g(mu);
        return map.count(key) != 0;
    }
    void insert(size_t key, const T* data, size_t n) {
        std::lock_guard<std::mutex> g(mu);
        if (budget == 0 || map.count(key)) return;
        while (held + n > budget && !order.empty()) {
            auto it = map.find(order.front());
            held -= it->second.size();
            map.erase(it);
            order.pop_front();
        }
        if (held + n > budget) return;
        map[key].assign(data, data + n);
        order.push_back(key);
        held += n;
    }
    void clear() {
        std::lock_guard<std::mutex> g(mu);
        map.clear();
        order.clear();
        held = 0;
    }
};

}  // namespace idc_plugin

/// nbit bits starting at bit i of a faiss::BitstringWriter string (LSB first), without moving the reader.
/// custom_invlists_impl.h:19, .cpp:35-58 (the reference masks with an `int` shift, undefined for nbit >= 31;
/// here every width up to 64 is exact)
inline uint64_t BitstringReader_get_bits(const faiss::BitstringReader& bs, size_t i, int nbit) {
    if (nbit <= 0) return 0;
    if (bs.code_size * 8 < (size_t)nbit + i) throw std::out_of_range("BitstringReader_get_bits: read past the end of the string");
    const size_t byte = i >> 3;
    const int sh = (int)(i & 7), need = sh + nbit;  // bits wanted counting from the start of `byte`
    uint64_t v = 0;
    for (int b = 0; b < 8 && 8 * b < need; b++) v |= (uint64_t)bs.code[byte + b] << (8 * b);
    v >>= sh;
    if (need > 64) v |= (uint64_t)bs.code[byte + 8] << (64 - sh);
    return nbit >= 64 ? v : v & ((1ull << nbit) - 1ull);
}

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
    /// host copy of the per-list coder states (custom_invlists_impl.h:59); empty until materialize_ans_states()
    std::vector<idc_plugin::AnsStateView> ans_states;
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
        size_t total_codes = 0, nonempty = 0;
        for (size_t l = 0; l < nlist; l++) {  // codes in sample order, custom_invlists_impl.cpp:189-193
            take_codes(il, l, order.data() + csr.offsets[l]);
            total_codes += codes_all[l].size();
            nonempty += codes_all[l].empty() ? 0 : 1;
        }
        // sic: the reference adds the size of ALL code arrays once per non-empty list (:198-205)
        codes_size_in_bytes = nonempty * total_codes;
        idc_roc_info info;
        idc_plugin::check(idc_roc_blob_info(blob, &info));
        compressed_ids_size_in_bytes = info.ans_bytes;  // sum of ANSState::size(), :199-202
        std::vector<uint64_t> unit_off(nlist + 1);
        std::vector<uint8_t> prec(info.nunits + 1);
        idc_plugin::check(idc_roc_blob_export(blob, nullptr, unit_off.data(), nullptr, prec.data(), nullptr, nullptr, nullptr));
        id_symbol_precision.resize(nlist);
        for (size_t l = 0; l < nlist; l++) id_symbol_precision[l] = prec[unit_off[l]];
        cache_.budget = kDefaultCacheIds;
    }
    ~CompressedIDInvertedListsFenwickTree() override { idc_roc_blob_free(blob); }

    void materialize_ans_states() { ans_states = idc_plugin::ans_states_of(blob); }

    /// bulk decode of the lists a search is about to touch; get_ids then serves from the cache. The cache holds at
    /// most cache_budget_ids() decoded ids (oldest lists are dropped first) so that it never grows into the
    /// uncompressed index; drop_cache() empties it (e.g. after a search batch).
    void prefetch_lists(const idx_t* list_nos, size_t n) const {
        std::vector<uint64_t> ln, off;
        size_t total = 0;
        for (size_t i = 0; i < n; i++) {
            if (cache_.contains((size_t)list_nos[i]) || list_size(list_nos[i]) == 0) continue;
            ln.push_back((uint64_t)list_nos[i]);
            total += list_size(list_nos[i]);
        }
        std::sort(ln.begin(), ln.end());
        ln.erase(std::unique(ln.begin(), ln.end()), ln.end());
        if (ln.empty()) return;
        off.resize(ln.size() + 1);
        std::vector<idx_t> ids(total + 1);
        idc_plugin::check(idc_roc_decode(idc_plugin::context(), blob, ln.data(), ln.size(), ids.data(), 8, IDC_MEM_HOST, off.data()));
        for (size_t i = 0; i < ln.size(); i++) cache_.insert(ln[i], ids.data() + off[i], off[i + 1] - off[i]);
    }
    void drop_cache() const { cache_.clear(); }
    void set_cache_budget_ids(size_t ids) const {
        cache_.clear();
        cache_.budget = ids;
    }
    size_t cache_budget_ids() const { return cache_.budget; }

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
        if (cache_.lookup(list_no, the_ids, nullptr)) return the_ids;
        uint64_t ln = list_no;
        idc_plugin::check(idc_roc_decode(idc_plugin::context(), blob, &ln, 1, the_ids, 8, IDC_MEM_HOST, nullptr));
        return the_ids;
    }

   private:
    static constexpr size_t kDefaultCacheIds = size_t(1) << 24;  // 128 MB of decoded ids at most
    mutable idc_plugin::DecodedCache<idx_t> cache_;
};

/// Elias-Fano ids. custom_invlists_impl.h:72-98, .cpp:229-339
struct CompressedIDInvertedListsEliasFano : InvertedListsArrayCodes {
    idc_ef_blob* blob = nullptr;
    size_t overhead_in_bytes = 0;
    /// host copy of the per-list bit vectors (custom_invlists_impl.h:76); empty until materialize_ef_bitstreams()
    std::vector<idc_plugin::EliasFanoView> ef_bitstreams;
    size_t compressed_ids_size_in_bytes = 0;
    size_t codes_size_in_bytes = 0;
    void materialize_ef_bitstreams() { ef_bitstreams = idc_plugin::ef_bitstreams_of(blob); }

    explicit CompressedIDInvertedListsEliasFano(const faiss::InvertedLists& il) : InvertedListsArrayCodes(il) {
        idc_plugin::Csr csr(il);
        codes_all.resize(nlist);
        std::vector<uint32_t> perm;
        size_t total_codes = 0, nonempty = 0;
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
            total_codes += codes_all[l].size();
            nonempty += ls ? 1 : 0;
        }
        codes_size_in_bytes = nonempty * total_codes;  // sic, :274-281: all code arrays once per non-empty list
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
        faiss::BitstringReader bs(ids_all[list_no].data(), ids_all[list_no].size());
        return (idx_t)BitstringReader_get_bits(bs, offset * bits, bits);
    }
};

/// Wavelet-tree ids. custom_invlists_impl.h:100-124, .cpp:346-397: one structure over S[id] = list_no,
/// get_single_id(list_no, offset) = wt.select(offset + 1, list_no). wt_type 1 (rrr_vector<63>): the levels as RRR(63) blocks.
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

namespace idc_plugin {
// Row cache policy of the compressed graphs. NSG search asks for ONE row per visited node, on the search thread
// (altid_impl.cpp:92-101,153-165); a launch + two copies per call costs more than decoding 64 ids on the CPU. So a
// miss decodes the row AND the rows of all its neighbours in one bulk call (the nodes a best-first search expands
// next are neighbours of the node it expands now) and parks them in a bounded cache. cache_rows = 0 switches it off.
template <class DecodeRows>
inline size_t cached_row(DecodedCache<int32_t>& cache, int K, int N, int i, int32_t* neighbors, const DecodeRows& decode_rows) {
    size_t cnt = 0;
    if (cache.budget && cache.lookup((size_t)i, neighbors, &cnt)) return cnt;
    std::vector<int32_t> row(K);
    uint32_t c0 = 0;
    int32_t r = i;
    decode_rows(&r, 1, row.data(), &c0);
    std::memcpy(neighbors, row.data(), c0 * sizeof(int32_t));
    if (cache.budget) {
        cache.insert((size_t)i, row.data(), c0);
        std::vector<int32_t> want;
        for (uint32_t j = 0; j < c0; j++)
            if (row[j] >= 0 && row[j] < N && !cache.contains((size_t)row[j])) want.push_back(row[j]);
        if (!want.empty()) {
            std::vector<int32_t> rows(want.size() * (size_t)K);
            std::vector<uint32_t> cnts(want.size());
            decode_rows(want.data(), want.size(), rows.data(), cnts.data());
            for (size_t t = 0; t < want.size(); t++) cache.insert((size_t)want[t], rows.data() + t * (size_t)K, cnts[t]);
        }
    }
    return c0;
}
}  // namespace idc_plugin

/// altid_impl.h:42-50, .cpp:53-101
struct EliasFanoNSGGraph : faiss::nsg::Graph<int32_t> {
    idc_ef_blob* blob = nullptr;
    /// host copy of the per-row bit vectors (altid_impl.h:43); empty until materialize_ef_bitstreams()
    std::vector<idc_plugin::EliasFanoView> ef_bitstreams;
    size_t compressed_ids_size_in_bytes = 0;
    size_t overhead_in_bytes = 0;
    explicit EliasFanoNSGGraph(const faiss::nsg::Graph<int32_t>& graph) : faiss::nsg::Graph<int32_t>(graph.data, graph.N, graph.K) {
        idc_plugin::check(idc_ef_encode_rows(idc_plugin::context(), N, K, graph.data, IDC_MEM_HOST, 0, &blob));
        idc_ef_info info;
        idc_plugin::check(idc_ef_blob_info(blob, &info));
        compressed_ids_size_in_bytes = info.bits_total / 8;
        // :56-57 list sizes + max ids; two `size_t += double` statements, each truncating
        overhead_in_bytes += (size_t)(N * std::ceil(std::log2((double)N)) / 8.0);
        overhead_in_bytes += (size_t)(N * std::ceil(std::log2((double)N)) / 8.0);
        data = nullptr;  // :89
    }
    ~EliasFanoNSGGraph() override { idc_ef_blob_free(blob); }
    void materialize_ef_bitstreams() { ef_bitstreams = idc_plugin::ef_bitstreams_of(blob); }
    /// bulk decode of many rows: out is n x K (-1 padded), counts the true lengths
    void get_neighbors_batch(const int32_t* rows, size_t n, int32_t* out, uint32_t* counts) const {
        idc_plugin::check(idc_ef_decode_rows(idc_plugin::context(), blob, rows, IDC_MEM_HOST, n, out, counts, IDC_MEM_HOST));
    }
    void set_cache_rows(size_t rows) const {
        cache_.clear();
        cache_.budget = rows * (size_t)K;
    }
    void drop_cache() const { cache_.clear(); }
    size_t get_neighbors(int i, int32_t* neighbors) const override {  // :92-101
        return idc_plugin::cached_row(cache_, K, N, i, neighbors,
                                      [&](const int32_t* r, size_t n, int32_t* o, uint32_t* c) { get_neighbors_batch(r, n, o, c); });
    }

   private:
    mutable idc_plugin::DecodedCache<int32_t> cache_;
};

/// altid_impl.h:53-67, .cpp:103-165
struct ROCNSGGraph : faiss::nsg::Graph<int32_t> {
    idc_roc_blob* blob = nullptr;
    /// host copy of the per-row coder states (altid_impl.h:58); empty until materialize_ans_states()
    std::vector<idc_plugin::AnsStateView> ans_states;
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
    void materialize_ans_states() { ans_states = idc_plugin::ans_states_of(blob); }
    void get_neighbors_batch(const int32_t* rows, size_t n, int32_t* out, uint32_t* counts) const {
        idc_plugin::check(idc_roc_decode_rows(idc_plugin::context(), blob, rows, IDC_MEM_HOST, n, out, counts, IDC_MEM_HOST));
    }
    void set_cache_rows(size_t rows) const {
        cache_.clear();
        cache_.budget = rows * (size_t)K;
    }
    void drop_cache() const { cache_.clear(); }
    size_t get_neighbors(int node, int32_t* neighbors) const override {  // :153-165
        size_t cnt = idc_plugin::cached_row(cache_, K, N, node, neighbors,
                                            [&](const int32_t* r, size_t n, int32_t* o, uint32_t* c) { get_neighbors_batch(r, n, o, c); });
        // sic: the reference returns K, not the row length (:164), and leaves neighbors[n..K) untouched -- whatever
        // the previous call left there. The slots are set to -1 here so that a caller trusting the returned K stops
        // at the row's end instead of walking stale ids.
        for (size_t j = cnt; j < (size_t)K; j++) neighbors[j] = -1;
        return K;
    }

   private:
    mutable idc_plugin::DecodedCache<int32_t> cache_;
};

// =============================================================================================================
// Free functions bound by the SWIG modules
// =============================================================================================================

namespace idc_plugin {

// labels[i] = (list_no << 32 | offset) or negative  ->  the id stored there. Our classes translate in bulk on the
// GPU; any other faiss::InvertedLists goes through its virtuals, hits grouped by list so that every hit list is
// fetched once (what the reference does for every class, custom_invlists_impl.cpp:477-525).
inline void translate_pairs(const faiss::InvertedLists* il, faiss::idx_t* labels, size_t n, bool decode_1by1) {
    using faiss::idx_t;
    if (auto* ft = dynamic_cast<const CompressedIDInvertedListsFenwickTree*>(il)) {
        ft->translate_labels(labels, n);  // hit lists decoded once in one launch + device gather (decode_1by1 or not)
        return;
    }
    auto* ef = dynamic_cast<const CompressedIDInvertedListsEliasFano*>(il);
    auto* wt = dynamic_cast<const CompressedIDInvertedListsWaveletTree*>(il);
    if (ef || wt) {  // random access is one select per hit: no list is decoded at all
        std::vector<uint64_t> ln, of;
        std::vector<size_t> where;
        for (size_t i = 0; i < n; i++)
            if (labels[i] >= 0) {
                ln.push_back(faiss::lo_listno(labels[i]));
                of.push_back(faiss::lo_offset(labels[i]));
                where.push_back(i);
            }
        if (where.empty()) return;
        std::vector<int64_t> ids(where.size());
        if (ef) ef->get_single_ids(ln.data(), of.data(), where.size(), ids.data());
        else wt->get_single_ids(ln.data(), of.data(), where.size(), ids.data());
        for (size_t t = 0; t < where.size(); t++) labels[where[t]] = ids[t];
        return;
    }
    if (decode_1by1) {  // :465-475
#pragma omp parallel for
        for (int64_t i = 0; i < (int64_t)n; i++)
            if (labels[i] >= 0) labels[i] = il->get_single_id(faiss::lo_listno(labels[i]), faiss::lo_offset(labels[i]));
        return;
    }
    std::vector<std::pair<uint64_t, size_t>> hits;  // (list_no, result slot), then grouped by list
    for (size_t i = 0; i < n; i++)
        if (labels[i] >= 0) hits.push_back({faiss::lo_listno(labels[i]), i});
    std::sort(hits.begin(), hits.end());
    std::vector<size_t> run_start;
    for (size_t t = 0; t < hits.size(); t++)
        if (t == 0 || hits[t].first != hits[t - 1].first) run_start.push_back(t);
    run_start.push_back(hits.size());
    const int64_t nruns = (int64_t)run_start.size() - 1;
#pragma omp parallel for
    for (int64_t r = 0; r < nruns; r++) {
        const size_t list_no = hits[run_start[r]].first;
        faiss::InvertedLists::ScopedIds sids(il, list_no);
        const idx_t* ids = sids.get();
        for (size_t t = run_start[r]; t < run_start[r + 1]; t++) {
            idx_t& l = labels[hits[t].second];
            l = ids[faiss::lo_offset(l)];
        }
    }
}

}  // namespace idc_plugin

/// Search an IVF index without decoding ids during the scan: collect (list, offset) pairs, translate them to ids
/// when the search is over. custom_invlists_impl.h:126-139, .cpp:407-526. `codes` (optional) receives the stored
/// code of every result, prefixed by the encoded list number when include_listno is set; empty result slots are
/// filled with 0xff.
inline void search_IVF_defer_id_decoding(const faiss::IndexIVF& index, faiss::idx_t n, const float* x, int k, float* distances,
                                         faiss::idx_t* labels, bool decode_1by1 = false, uint8_t* codes = nullptr,
                                         bool include_listno = false) {
    using faiss::idx_t;
    std::unique_ptr<float[]> Dq(new float[n * index.nprobe]);
    std::unique_ptr<idx_t[]> Iq(new idx_t[n * index.nprobe]);
    FAISS_THROW_IF_NOT_MSG(index.parallel_mode == 3, "set the parallel mode to 3 otherwise search will be single-threaded");  // :420-422
    index.quantizer->search(n, x, index.nprobe, Dq.get(), Iq.get());                           // :424
    index.search_preassigned(n, x, k, Iq.get(), Dq.get(), distances, labels, true);           // :427-428, store_pairs
    const faiss::InvertedLists* invlists = index.invlists;
    const size_t nres = (size_t)n * (size_t)k;
    if (codes) {  // :433-461
        const size_t code_size = index.code_size;
        const size_t code_size_1 = code_size + (include_listno ? index.coarse_code_size() : 0);
#pragma omp parallel for if (nres > 1000)
        for (int64_t ij = 0; ij < (int64_t)nres; ij++) {
            uint8_t* dst = codes + (size_t)ij * code_size_1;
            const idx_t key = labels[ij];
            if (key < 0) {
                std::memset(dst, 0xff, code_size_1);
                continue;
            }
            const uint64_t list_no = faiss::lo_listno(key), offset = faiss::lo_offset(key);
            if (include_listno) {
                index.encode_listno(list_no, dst);
                dst += code_size_1 - code_size;
            }
            std::memcpy(dst, invlists->get_single_code(list_no, offset), code_size);
        }
    }
    idc_plugin::translate_pairs(invlists, labels, nres, decode_1by1);  // :464-525
}

namespace idc_plugin {
/// records every node whose distance is computed (altid_impl.cpp:172-204)
struct TracingDistanceComputer : faiss::DistanceComputer {
    std::vector<faiss::idx_t> visited;
    std::unique_ptr<faiss::DistanceComputer> basedis;
    explicit TracingDistanceComputer(faiss::DistanceComputer* basedis) : basedis(basedis) {}
    void set_query(const float* x) override { basedis->set_query(x); }
    float operator()(faiss::idx_t i) override {
        visited.push_back(i);
        return (*basedis)(i);
    }
    float symmetric_dis(faiss::idx_t i, faiss::idx_t j) override {
        visited.push_back(i);
        visited.push_back(j);
        return basedis->symmetric_dis(i, j);
    }
};
}  // namespace idc_plugin

/// NSG search that also returns the ids of all nodes a distance was computed for. altid_impl.h:18-25, .cpp:207-231
inline void search_NSG_and_trace(const faiss::IndexNSG& index, faiss::idx_t n, const float* x, int k, faiss::idx_t* labels,
                                 float* distances, std::vector<faiss::idx_t>& visited_nodes) {
    faiss::VisitedTable vt(index.ntotal);
    idc_plugin::TracingDistanceComputer dis(faiss::nsg::storage_distance_computer(index.storage));
    for (faiss::idx_t i = 0; i < n; i++) {
        dis.set_query(x + i * index.d);
        index.nsg.search(dis, k, labels + i * k, distances + i * k, vt);
        vt.advance();
    }
    std::swap(visited_nodes, dis.visited);
}
