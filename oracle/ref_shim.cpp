// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Thin extern "C" shim around the UNMODIFIED reference ROC codec, compiled in
// place from /root/reference (custom_invlist_cpp/codec.{h,cpp} +
// fenwick_tree_cpp/src/fenwick_tree.h). No reference source is copied into
// this repository: the Makefile in this directory passes the reference files
// to g++ where they lie and writes the result to oracle/_ref/libref_roc.so.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the resulting library.
//
// What is wrapped (reference file:line):
//   compress / decompress                      codec.cpp:123-152
//   pop/push_with_finer_precision              codec.cpp:21-63
//   codec_push / codec_pop                     codec.cpp:92-121
//   FenwickTree<T>::insert_then_forward_lookup fenwick_tree.h:42-94
//   FenwickTree<T>::reverse_lookup_then_remove fenwick_tree.h:96-140
//   plugin list loop (shuffle, insert, sample) custom_invlists_impl.cpp:147-194
//   plugin get_ids (state copy + decompress)   custom_invlists_impl.cpp:210-219

#include "codec.h"
#include "../fenwick_tree_cpp/src/fenwick_tree.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <random>
#include <tuple>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

// Upper bound on stream words for a set of n ids at the given precision.
uint64_t ref_roc_max_words(uint64_t n, int precision) {
    return (n * (uint64_t)(precision > 0 ? precision : 1) + 31) / 32 + 4;
}

// compress() on a fresh ANSState. Returns number of stack words written to
// words_out (bottom -> top), or -1 if words_cap is too small.
int64_t ref_roc_compress(
        uint64_t n,
        const uint64_t* data,
        int precision,
        uint64_t* head_out,
        uint32_t* words_out,
        uint64_t words_cap) {
    ANSState st;
    compress(n, data, st, precision);
    *head_out = st.head;
    if (st.stack.size() > words_cap)
        return -1;
    if (!st.stack.empty())
        memcpy(words_out, st.stack.data(), st.stack.size() * sizeof(uint32_t));
    return (int64_t)st.stack.size();
}

// Plugin-style encode of one list (custom_invlists_impl.cpp:156-192): random
// order insertion (seeded here so the run is reproducible; the stream does not
// depend on the seed), then the sample loop. order_out[i] (optional) receives
// the position in the input array of the id sampled at step i, i.e. the
// permutation applied to the codes (custom_invlists_impl.cpp:189-190).
int64_t ref_roc_encode_list(
        uint64_t n,
        const uint64_t* ids,
        int precision,
        uint64_t seed,
        uint64_t* head_out,
        uint32_t* words_out,
        uint64_t words_cap,
        uint32_t* order_out) {
    using Sym = std::tuple<uint64_t, uint32_t>;
    ANSState st;
    FenwickTree<Sym> ftree;
    std::vector<uint32_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0u);
    std::mt19937 g((uint32_t)seed);
    std::shuffle(idx.begin(), idx.end(), g);
    for (uint32_t i : idx)
        ftree.insert_then_forward_lookup(std::make_tuple(ids[i], i));
    for (uint64_t i = 0; i < n; i++) {
        uint32_t nmax = (uint32_t)(n - i);
        size_t index = pop_with_finer_precision(st, nmax);
        auto range = ftree.reverse_lookup_then_remove((int)index);
        codec_push(st, std::get<0>(range.ftree->symbol), precision);
        if (order_out)
            order_out[i] = std::get<1>(range.ftree->symbol);
    }
    *head_out = st.head;
    if (st.stack.size() > words_cap)
        return -1;
    if (!st.stack.empty())
        memcpy(words_out, st.stack.data(), st.stack.size() * sizeof(uint32_t));
    return (int64_t)st.stack.size();
}

// decompress() on a copy of (head, words). Writes n ids to out (reference
// order: out[n-i-1] is the i-th decoded symbol). Optionally returns the final
// head / stack size for inspection.
void ref_roc_decompress(
        uint64_t head,
        const uint32_t* words,
        uint64_t nwords,
        uint64_t n,
        int precision,
        uint64_t* out,
        uint64_t* final_head,
        uint64_t* final_nwords) {
    ANSState st;
    st.head = head;
    st.stack.assign(words, words + nwords);
    decompress(st, n, out, precision);
    if (final_head)
        *final_head = st.head;
    if (final_nwords)
        *final_nwords = st.stack.size();
}

// The reference precision rule, custom_invlists_impl.cpp:163-164 /
// altid_impl.cpp:124-125 (int truncation + double log2 + ceil).
uint64_t ref_precision_rule(uint64_t max_id_u64) {
    int max_id = (int)max_id_u64;
    if (max_id <= 0)
        return 0; // log2(0) = -inf: UB in the reference; 0 by convention here
    return (uint64_t)std::ceil(std::log2(max_id));
}

// ---- step-wise access, used to pin the device functions one at a time ----

struct RefState {
    ANSState st;
};

void* ref_state_new(uint64_t head, const uint32_t* words, uint64_t nwords) {
    RefState* s = new RefState();
    s->st.head = head;
    if (nwords)
        s->st.stack.assign(words, words + nwords);
    return s;
}
void ref_state_free(void* p) { delete (RefState*)p; }
uint64_t ref_state_head(void* p) { return ((RefState*)p)->st.head; }
uint64_t ref_state_nwords(void* p) { return ((RefState*)p)->st.stack.size(); }
void ref_state_words(void* p, uint32_t* out) {
    auto& v = ((RefState*)p)->st.stack;
    if (!v.empty())
        memcpy(out, v.data(), v.size() * 4);
}
uint64_t ref_pop_uniform(void* p, uint64_t nmax) {
    return pop_with_finer_precision(((RefState*)p)->st, nmax);
}
void ref_push_uniform(void* p, uint64_t sym, uint64_t nmax) {
    push_with_finer_precision(((RefState*)p)->st, sym, nmax);
}
void ref_codec_push(void* p, uint64_t sym, int precision) {
    codec_push(((RefState*)p)->st, sym, precision);
}
uint64_t ref_codec_pop(void* p, int precision) {
    return codec_pop(((RefState*)p)->st, precision);
}

// ---- order-statistic tree, scripted access (test_fenwick_tree.cpp) ----

void* ref_ftree_new() { return new FenwickTree<int64_t>(); }
void ref_ftree_free(void* t) { delete (FenwickTree<int64_t>*)t; }
void ref_ftree_insert(void* t, int64_t sym, int64_t* out3) {
    auto r = ((FenwickTree<int64_t>*)t)->insert_then_forward_lookup(sym);
    out3[0] = r.ftree->symbol;
    out3[1] = r.start;
    out3[2] = r.freq;
}
void ref_ftree_remove(void* t, int index, int64_t* out3) {
    auto r = ((FenwickTree<int64_t>*)t)->reverse_lookup_then_remove(index);
    out3[0] = r.ftree->symbol;
    out3[1] = r.start;
    out3[2] = r.freq;
}

// ---- bulk CPU baseline: the plugin's OpenMP loop over lists ----
//
// offsets[nlist+1] CSR over ids; precision[nlist]; word_offsets[nlist+1] gives
// each list's slot in words_out (capacity = word_offsets[l+1]-word_offsets[l]).
// nwords_out[l] receives the words actually used. Returns 0, or -1 on overflow.
int ref_roc_encode_lists(
        uint64_t nlist,
        const uint64_t* offsets,
        const uint64_t* ids,
        const uint8_t* precision,
        const uint64_t* word_offsets,
        uint64_t* heads_out,
        uint32_t* words_out,
        uint64_t* nwords_out,
        int nthreads) {
    int bad = 0;
#ifdef _OPENMP
    if (nthreads > 0)
        omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t l = 0; l < (int64_t)nlist; l++) {
        uint64_t n = offsets[l + 1] - offsets[l];
        heads_out[l] = (uint64_t)1 << 31;
        nwords_out[l] = 0;
        if (n == 0)
            continue;
        int64_t w = ref_roc_encode_list(
                n,
                ids + offsets[l],
                precision[l],
                0x9e3779b9u ^ (uint64_t)l,
                &heads_out[l],
                words_out + word_offsets[l],
                word_offsets[l + 1] - word_offsets[l],
                nullptr);
        if (w < 0) {
#pragma omp atomic write
            bad = 1;
        } else {
            nwords_out[l] = (uint64_t)w;
        }
    }
    return bad ? -1 : 0;
}

void ref_roc_decode_lists(
        uint64_t nlist,
        const uint64_t* offsets,
        const uint8_t* precision,
        const uint64_t* word_offsets,
        const uint64_t* nwords,
        const uint64_t* heads,
        const uint32_t* words,
        uint64_t* ids_out,
        int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0)
        omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t l = 0; l < (int64_t)nlist; l++) {
        uint64_t n = offsets[l + 1] - offsets[l];
        if (n == 0)
            continue;
        ref_roc_decompress(
                heads[l],
                words + word_offsets[l],
                nwords[l],
                n,
                precision[l],
                ids_out + offsets[l],
                nullptr,
                nullptr);
    }
}

int ref_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

} // extern "C"
