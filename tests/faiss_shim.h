// TEST ONLY: the handful of Faiss declarations idc_faiss_plugin.h uses, so the adapter can be compiled in an
// image without Faiss. Shapes follow faiss/invlists/InvertedLists.h and faiss/impl/NSG.h [third-party].
#pragma once
#include <cstddef>
#include <cstdint>
namespace faiss {
using idx_t = int64_t;
struct InvertedLists {
    size_t nlist, code_size;
    InvertedLists(size_t nlist, size_t code_size) : nlist(nlist), code_size(code_size) {}
    virtual size_t list_size(size_t list_no) const = 0;
    virtual const uint8_t* get_codes(size_t list_no) const = 0;
    virtual const idx_t* get_ids(size_t list_no) const = 0;
    virtual void release_codes(size_t, const uint8_t*) const {}
    virtual void release_ids(size_t, const idx_t*) const {}
    virtual idx_t get_single_id(size_t list_no, size_t offset) const { return get_ids(list_no)[offset]; }
    virtual ~InvertedLists() {}
    struct ScopedIds {
        const InvertedLists* il; const idx_t* ids; size_t list_no;
        ScopedIds(const InvertedLists* il, size_t l) : il(il), ids(il->get_ids(l)), list_no(l) {}
        const idx_t* get() { return ids; }
        idx_t operator[](size_t i) const { return ids[i]; }
        ~ScopedIds() { il->release_ids(list_no, ids); }
    };
    struct ScopedCodes {
        const InvertedLists* il; const uint8_t* codes; size_t list_no;
        ScopedCodes(const InvertedLists* il, size_t l) : il(il), codes(il->get_codes(l)), list_no(l) {}
        const uint8_t* get() { return codes; }
        ~ScopedCodes() { il->release_codes(list_no, codes); }
    };
};
struct ReadOnlyInvertedLists : InvertedLists {
    ReadOnlyInvertedLists(size_t nlist, size_t code_size) : InvertedLists(nlist, code_size) {}
};
namespace nsg {
template <class node_t>
struct Graph {
    node_t* data; int K; int N; bool own_fields;
    Graph(node_t* data, int N, int K) : data(data), K(K), N(N), own_fields(false) {}
    virtual size_t get_neighbors(int i, node_t* neighbors) const { (void)i; (void)neighbors; return 0; }
    virtual ~Graph() {}
};
}  // namespace nsg
}  // namespace faiss
