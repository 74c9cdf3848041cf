// roc_lane.cuh -- one ROC unit per lane: the per-step encode / decode bodies.
// Shared by the CUDA kernels (roc_kernels.cu) and by tests/hostsim.
//
//   encode step  = custom_invlists_impl.cpp:178-192 / codec.cpp:131-137
//   decode step  = codec.cpp:144-151 / altid_impl.cpp:157-163
#pragma once

#include "idc_core.cuh"

namespace idc {

template <typename IdT>
struct EncLane {
    EncState st;
    EncTree tree;
    const uint32_t* sort_idx;  // original position of each sorted id inside its list (null: input was sorted)
    uint32_t* order;           // sample order out (null: not wanted)
    uint32_t pos_base;         // position of the unit's first id inside its list
    uint32_t n;
    int prec;
};

template <typename IdT>
IDC_HD uint64_t load_id(const IdT* p) {
#if defined(__CUDA_ARCH__)
    IdT v = __ldg(p);
#else
    IdT v = *p;
#endif
    if (sizeof(IdT) == 4)
        return (uint64_t)(uint32_t)v;
    return (uint64_t)v;
}

// nmax = ids still in the set (the caller walks it from n down to 1)
template <typename IdT>
IDC_HD void enc_lane_step(EncLane<IdT>& L, uint32_t nmax, uint64_t rcp, uint32_t q31, const uint32_t* mt) {
    uint32_t k = enc_pop_uniform(L.st, nmax, rcp, q31, mt);
    uint32_t id32;
    uint32_t pos = enc_tree_select_remove(L.tree, k, id32);
    enc_push_id(L.st, (uint64_t)id32, L.prec);
    if (L.order) {
        uint32_t o = L.sort_idx ? L.sort_idx[pos] : L.pos_base + pos;
        L.order[L.n - nmax] = o;
    }
}

template <typename OutT>
struct DecLane {
    DecState st;
    DecTree tree;
    OutT* out;  // the unit's n output slots
    uint32_t n;
    int prec;
};

// i = 0-based step; q31 = 2^31 / (i + 1)
template <typename OutT>
IDC_HD void dec_lane_step(DecLane<OutT>& L, uint32_t i, uint32_t q31, const uint32_t* mt) {
    uint64_t id = dec_pop_id(L.st, L.prec, mt);
    uint32_t rank = dec_tree_insert_rank(L.tree, (uint32_t)id, L.out + (L.n - i), i);
    dec_push_uniform(L.st, rank, i + 1u, q31, mt);
    L.out[L.n - 1u - i] = (OutT)id;
    if (L.tree.degenerate)
        L.st.status |= kStDegenerate;
}

}  // namespace idc
