// TEST ONLY: the Faiss declarations idc_faiss_plugin.h and the shipped .swig %inline blocks use, so that the adapter
// can be compiled AND RUN in an image without Faiss. Shapes follow faiss/Index.h, faiss/IndexIVF.h,
// faiss/invlists/InvertedLists.h, faiss/impl/NSG.h, faiss/IndexNSG.h, faiss/utils/hamming.h,
// faiss/impl/DistanceComputer.h, faiss/impl/AuxIndexStructures.h [third-party, absent here]. The bodies are the
// simplest thing with the same contract: an exhaustive flat quantizer, an IVF whose codes are the raw float vectors,
// a best-first graph search that reads rows only through Graph::get_neighbors. Nothing here is product code.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace faiss {
using idx_t = int64_t;

struct FaissException : std::exception {
    std::string msg;
    explicit FaissException(const std::string& m) : msg(m) {}
    const char* what() const noexcept override { return msg.c_str(); }
};
#define FAISS_THROW_IF_NOT_MSG(X, MSG)                                   \
    do {                                                                 \
        if (!(X)) throw faiss::FaissException(std::string("Error: '") + #X + "' failed: " + MSG); \
    } while (0)
#define FAISS_THROW_IF_NOT(X)                                            \
    do {                                                                 \
        if (!(X)) throw faiss::FaissException(std::string("Error: '") + #X + "' failed"); \
    } while (0)

// faiss/invlists/DirectMap.h
inline uint64_t lo_build(uint64_t list_id, uint64_t offset) { return list_id << 32 | offset; }
inline uint64_t lo_listno(uint64_t lo) { return lo >> 32; }
inline uint64_t lo_offset(uint64_t lo) { return lo & 0xffffffff; }

// faiss/utils/hamming.h
struct BitstringReader {
    const uint8_t* code;
    size_t code_size;
    size_t i;
    BitstringReader(const uint8_t* code, size_t code_size) : code(code), code_size(code_size), i(0) {}
    uint64_t read(int nbit) {
        uint64_t v = 0;
        for (int b = 0; b < nbit; b++, i++) v |= (uint64_t)((code[i >> 3] >> (i & 7)) & 1u) << b;
        return v;
    }
};

struct InvertedLists {
    size_t nlist, code_size;
    InvertedLists(size_t nlist, size_t code_size) : nlist(nlist), code_size(code_size) {}
    virtual size_t list_size(size_t list_no) const = 0;
    virtual const uint8_t* get_codes(size_t list_no) const = 0;
    virtual const idx_t* get_ids(size_t list_no) const = 0;
    virtual void release_codes(size_t, const uint8_t*) const {}
    virtual void release_ids(size_t, const idx_t*) const {}
    virtual idx_t get_single_id(size_t list_no, size_t offset) const {
        const idx_t* ids = get_ids(list_no);
        idx_t r = ids[offset];
        release_ids(list_no, ids);
        return r;
    }
    virtual const uint8_t* get_single_code(size_t list_no, size_t offset) const { return get_codes(list_no) + offset * code_size; }
    size_t compute_ntotal() const {
        size_t t = 0;
        for (size_t l = 0; l < nlist; l++) t += list_size(l);
        return t;
    }
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

// faiss/impl/DistanceComputer.h
struct DistanceComputer {
    virtual void set_query(const float* x) = 0;
    virtual float operator()(idx_t i) = 0;
    virtual float symmetric_dis(idx_t i, idx_t j) = 0;
    virtual ~DistanceComputer() {}
};

// faiss/impl/AuxIndexStructures.h
struct VisitedTable {
    std::vector<uint8_t> visited;
    uint8_t visno;
    explicit VisitedTable(int size) : visited(size), visno(1) {}
    void set(int no) { visited[no] = visno; }
    bool get(int no) const { return visited[no] == visno; }
    void advance() {
        visno++;
        if (visno == 250) {
            std::fill(visited.begin(), visited.end(), 0);
            visno = 1;
        }
    }
};

struct Index {
    int d;
    idx_t ntotal = 0;
    explicit Index(int d) : d(d) {}
    virtual void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const = 0;
    virtual DistanceComputer* get_distance_computer() const { throw FaissException("get_distance_computer not implemented"); }
    virtual ~Index() {}
};

inline float shim_l2(const float* a, const float* b, int d) {
    float s = 0;
    for (int j = 0; j < d; j++) s += (a[j] - b[j]) * (a[j] - b[j]);
    return s;
}

// k smallest (distance, label) pairs of a scan, ties by scan order; empty slots: +inf / -1 like Faiss heaps
struct ShimTopK {
    size_t k;
    std::vector<std::pair<float, idx_t>> h;
    explicit ShimTopK(size_t k) : k(k) {}
    void add(float dis, idx_t label) {
        auto it = std::upper_bound(h.begin(), h.end(), dis, [](float v, const std::pair<float, idx_t>& p) { return v < p.first; });
        if ((size_t)(it - h.begin()) >= k) return;
        h.insert(it, {dis, label});
        if (h.size() > k) h.pop_back();
    }
    void write(float* D, idx_t* I) const {
        for (size_t j = 0; j < k; j++) {
            D[j] = j < h.size() ? h[j].first : std::numeric_limits<float>::infinity();
            I[j] = j < h.size() ? h[j].second : -1;
        }
    }
};

struct IndexFlatL2 : Index {
    std::vector<float> xb;
    explicit IndexFlatL2(int d) : Index(d) {}
    void add(idx_t n, const float* x) {
        xb.insert(xb.end(), x, x + n * d);
        ntotal += n;
    }
    void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const override {
        for (idx_t q = 0; q < n; q++) {
            ShimTopK top(k);
            for (idx_t i = 0; i < ntotal; i++) top.add(shim_l2(x + q * d, xb.data() + i * d, d), i);
            top.write(distances + q * k, labels + q * k);
        }
    }
    struct DC : DistanceComputer {
        const IndexFlatL2& ix;
        const float* q = nullptr;
        explicit DC(const IndexFlatL2& ix) : ix(ix) {}
        void set_query(const float* x) override { q = x; }
        float operator()(idx_t i) override { return shim_l2(q, ix.xb.data() + i * ix.d, ix.d); }
        float symmetric_dis(idx_t i, idx_t j) override { return shim_l2(ix.xb.data() + i * ix.d, ix.xb.data() + j * ix.d, ix.d); }
    };
    DistanceComputer* get_distance_computer() const override { return new DC(*this); }
};

struct IVFSearchParameters;
struct IndexIVFStats;

// IVFx,Flat: the code of a vector is the vector (code_size = 4 d)
struct IndexIVF : Index {
    size_t nlist;
    size_t nprobe = 1;
    int parallel_mode = 0;
    Index* quantizer;
    InvertedLists* invlists = nullptr;
    bool own_invlists = false;
    size_t code_size;
    IndexIVF(Index* quantizer, int d, size_t nlist) : Index(d), nlist(nlist), quantizer(quantizer), code_size(sizeof(float) * d) {}
    ~IndexIVF() override {
        if (own_invlists) delete invlists;
    }
    void replace_invlists(InvertedLists* il, bool own = false) {
        if (own_invlists) delete invlists;
        invlists = il;
        own_invlists = own;
    }
    size_t coarse_code_size() const {
        size_t nl = nlist - 1, nbyte = 0;
        while (nl > 0) nbyte++, nl >>= 8;
        return nbyte;
    }
    void encode_listno(idx_t list_no, uint8_t* code) const {
        size_t nl = nlist - 1;
        while (nl > 0) *code++ = list_no & 0xff, list_no >>= 8, nl >>= 8;
    }
    void search_preassigned(idx_t n, const float* x, idx_t k, const idx_t* assign, const float* /*centroid_dis*/,
                            float* distances, idx_t* labels, bool store_pairs, const IVFSearchParameters* = nullptr,
                            IndexIVFStats* = nullptr) const {
        for (idx_t q = 0; q < n; q++) {
            ShimTopK top(k);
            for (size_t p = 0; p < nprobe; p++) {
                idx_t key = assign[q * nprobe + p];
                if (key < 0) continue;
                size_t ls = invlists->list_size(key);
                if (ls == 0) continue;
                InvertedLists::ScopedCodes scodes(invlists, key);
                std::unique_ptr<InvertedLists::ScopedIds> sids;  // ids are only touched when pairs are not stored
                const idx_t* ids = nullptr;
                if (!store_pairs) {
                    sids.reset(new InvertedLists::ScopedIds(invlists, key));
                    ids = sids->get();
                }
                const float* codes = reinterpret_cast<const float*>(scodes.get());
                for (size_t j = 0; j < ls; j++)
                    top.add(shim_l2(x + q * d, codes + j * d, d), store_pairs ? (idx_t)lo_build(key, j) : ids[j]);
            }
            top.write(distances + q * k, labels + q * k);
        }
    }
    void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const override {
        std::vector<float> Dq(n * nprobe);
        std::vector<idx_t> Iq(n * nprobe);
        quantizer->search(n, x, nprobe, Dq.data(), Iq.data());
        search_preassigned(n, x, k, Iq.data(), Dq.data(), distances, labels, false);
    }
};

namespace nsg {
template <class node_t>
struct Graph {
    node_t* data; int K; int N; bool own_fields;
    Graph(node_t* data, int N, int K) : data(data), K(K), N(N), own_fields(false) {}
    virtual size_t get_neighbors(int i, node_t* neighbors) const {
        size_t n = 0;
        for (; n < (size_t)K && data[(size_t)i * K + n] >= 0; n++) neighbors[n] = data[(size_t)i * K + n];
        return n;
    }
    virtual ~Graph() {}
};
inline DistanceComputer* storage_distance_computer(const Index* storage) { return storage->get_distance_computer(); }
}  // namespace nsg

struct NSG {
    int ntotal = 0;
    int search_L = 16;
    int enterpoint = 0;
    std::shared_ptr<nsg::Graph<int32_t>> final_graph;
    // best-first search with a pool of L candidates; rows are read through Graph::get_neighbors only
    void search(DistanceComputer& dis, int k, idx_t* I, float* D, VisitedTable& vt) const {
        const int L = std::max(search_L, k);
        struct Cand { float d; int id; bool expanded; };
        std::vector<Cand> pool;
        std::vector<int32_t> nb(final_graph->K);
        pool.push_back({dis(enterpoint), enterpoint, false});
        vt.set(enterpoint);
        for (;;) {
            size_t c = 0;
            while (c < pool.size() && pool[c].expanded) c++;
            if (c == pool.size()) break;
            pool[c].expanded = true;
            const int node = pool[c].id;
            size_t nn = final_graph->get_neighbors(node, nb.data());
            for (size_t m = 0; m < nn; m++) {
                int id = nb[m];
                if (id < 0 || id >= ntotal) break;
                if (vt.get(id)) continue;
                vt.set(id);
                float dd = dis(id);
                if ((int)pool.size() == L && dd >= pool.back().d) continue;
                auto it = std::upper_bound(pool.begin(), pool.end(), dd, [](float v, const Cand& p) { return v < p.d; });
                pool.insert(it, Cand{dd, id, false});
                if ((int)pool.size() > L) pool.pop_back();
            }
        }
        for (int j = 0; j < k; j++) {
            I[j] = j < (int)pool.size() ? pool[j].id : -1;
            D[j] = j < (int)pool.size() ? pool[j].d : std::numeric_limits<float>::infinity();
        }
    }
};

struct IndexNSG : Index {
    NSG nsg;
    Index* storage;
    explicit IndexNSG(Index* storage) : Index(storage->d), storage(storage) {}
    void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const override {
        VisitedTable vt(ntotal);
        std::unique_ptr<DistanceComputer> dis(nsg::storage_distance_computer(storage));
        for (idx_t i = 0; i < n; i++) {
            dis->set_query(x + i * d);
            nsg.search(*dis, (int)k, labels + i * k, distances + i * k, vt);
            vt.advance();
        }
    }
};
}  // namespace faiss
