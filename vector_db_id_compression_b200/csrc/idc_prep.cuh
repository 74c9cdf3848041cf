// idc_prep.cuh -- preparation kernels shared by the ROC and Elias-Fano paths:
// per-unit metadata (min / max id, precision rule, input checks), NSG row
// lengths, and the per-unit sort used when the caller's ids are not ascending.
#pragma once

#include <algorithm>
#include <vector>

#include "idc_core.cuh"
#include "idc_host.h"
#include "idc_scan.cuh"

namespace {

using namespace idc;

constexpr int kThreads = 128;  // 4 warps per CTA

template <typename T>
int dev_alloc(idc_ctx* c, T** p, size_t count, uint64_t* acct = nullptr) {
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    IDC_TRY(c->pool_alloc(reinterpret_cast<void**>(p), bytes));
    if (acct) *acct += bytes;
    return IDC_OK;
}

template <typename T>
int upload(idc_ctx* c, T* dst, const std::vector<T>& src) {
    if (src.empty()) return IDC_OK;
    IDC_CUDA(cudaMemcpyAsync(dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    return IDC_OK;
}

// ------------------------------------------------------------------ kernels

struct MetaArgs {
    const void* ids;
    const uint64_t* unit_src;
    const uint32_t* unit_n;
    uint32_t nunits;
    uint32_t expect_sorted;
    uint32_t precision_safe;
    uint8_t* unit_prec;
    uint32_t* unit_lo;
    uint32_t* unit_hi;
    uint32_t* status;  // single word, OR of per-unit flags
    // tiles of kMetaTile ids: a unit (an Elias-Fano "unit" is a whole list, possibly 10^8 ids) is scanned by as
    // many warps as it has tiles
    const uint32_t* tile_unit;
    const uint32_t* tile_idx;
    uint32_t ntiles;      // tiles of this launch: [tile_base, tile_base + ntiles)
    uint32_t tile_base;
    uint32_t unit_base;   // k_unit_meta_finish: units [unit_base, unit_base + unit_count)
    uint32_t unit_count;
};

constexpr uint32_t kMetaTile = 4096;

__device__ __forceinline__ uint32_t bit_length64(uint64_t x) { return x ? 64u - (uint32_t)__clzll((long long)x) : 0u; }

// One warp per tile: min / max id, ascending check (including the pair that straddles the tile's end) and 32-bit
// width check, merged into the unit's slots with atomics (unit_lo starts at 0xffffffff, unit_hi at 0).
template <typename IdT>
__global__ void __launch_bounds__(kThreads) k_unit_meta(MetaArgs a) {
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.ntiles) return;
    warp += a.tile_base;
    // null tables: every unit has at most one tile (tile t = unit t, empty units included)
    const uint32_t u = a.tile_unit ? a.tile_unit[warp] : warp;
    const IdT* src = reinterpret_cast<const IdT*>(a.ids) + a.unit_src[u];
    const uint32_t n = a.unit_n[u];
    const uint32_t i0 = (a.tile_idx ? a.tile_idx[warp] : 0u) * kMetaTile, i1 = i0 + kMetaTile < n ? i0 + kMetaTile : n;
    uint64_t mx = 0, mn = ~0ull;
    uint32_t bad = 0;
    for (uint32_t i = i0 + lane; i < i1; i += 32) {
        uint64_t v = load_id(src + i);
        if (sizeof(IdT) == 8 && (v >> 32)) bad |= kStWide;
        if (a.expect_sorted && i + 1 < n && load_id(src + i + 1) < v) bad |= kStUnsorted;
        mx = v > mx ? v : mx;
        mn = v < mn ? v : mn;
    }
    for (int o = 16; o; o >>= 1) {
        uint64_t omx = __shfl_xor_sync(0xffffffffu, mx, o), omn = __shfl_xor_sync(0xffffffffu, mn, o);
        mx = omx > mx ? omx : mx;
        mn = omn < mn ? omn : mn;
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if (lane == 0) {
        atomicMin(a.unit_lo + u, (uint32_t)(mn > 0xffffffffull ? 0xffffffffull : mn));
        atomicMax(a.unit_hi + u, (uint32_t)(mx > 0xffffffffull ? 0xffffffffull : mx));
        if (bad) atomicOr(a.status, bad);
    }
}

// One thread per unit: the reference precision rule (uint64_t)ceil(log2((int)max_id))
// (custom_invlists_impl.cpp:163-164, altid_impl.cpp:124-125) restated in integers:
// ceil(log2(m)) = bit_length(m - 1) for m >= 1.
__global__ void __launch_bounds__(kThreads) k_unit_meta_finish(MetaArgs a) {
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= a.unit_count) return;
    u += a.unit_base;
    const uint32_t n = a.unit_n[u];
    const uint64_t mx = a.unit_hi[u];
    uint32_t p;
    if (n == 0)
        p = 0;
    else if (a.precision_safe)
        p = bit_length64(mx);
    else
        p = mx ? bit_length64(mx - 1) : 0;  // max_id == 0 is undefined in the reference; 0 here
    a.unit_prec[u] = (uint8_t)p;
    if (n == 0) a.unit_lo[u] = 0u, a.unit_hi[u] = 0u;
}

// NSG rows: number of entries before the first -1 (altid_impl.cpp:110-117)
__global__ void __launch_bounds__(kThreads) k_row_counts(const int32_t* data, uint64_t nrows, uint32_t K, uint32_t* counts) {
    uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    const int32_t* row = data + warp * K;
    uint32_t cnt = K;
    for (uint32_t base = 0; base < K; base += 32) {
        uint32_t j = base + lane;
        bool stop = j < K && __ldg(row + j) == -1;
        uint32_t m = __ballot_sync(0xffffffffu, stop);
        if (m) {
            cnt = base + (uint32_t)__ffs((int)m) - 1u;
            break;
        }
    }
    if (lane == 0) counts[warp] = cnt;
}

// Per-unit bitonic sort of (id << 32 | position) keys. One CTA per unit; the
// keys live in shared memory when the padded unit fits (<= 4096), otherwise in
// a global scratch slot. Only used when the caller did not pass IDC_F_SORTED.
struct SortArgs {
    const void* ids;
    const uint64_t* unit_src;
    const uint32_t* unit_n;
    const uint32_t* unit_posbase;
    uint32_t nunits;
    uint32_t* sorted_ids;   // same element layout as ids
    uint32_t* sort_idx;
    uint64_t* big_scratch;  // 65536 keys per CTA slot (gridDim.x slots)
};

constexpr uint32_t kSortSmem = 4096;
constexpr uint32_t kSortWarp = 64;  // units up to this size are sorted by one warp in registers (k_sort_small)

// Units of <= 64 ids (NSG rows: K <= 64): one warp per unit, two keys per lane, bitonic network over shuffles --
// no shared memory, no block barriers (the CTA-wide sort below pays 21 __syncthreads per 64-id row).
template <typename IdT>
__global__ void __launch_bounds__(kThreads) k_sort_small(SortArgs a) {
    const uint32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (u >= a.nunits) return;
    const uint32_t n = a.unit_n[u];
    if (n == 0 || n > kSortWarp) return;
    const uint64_t src_off = a.unit_src[u];
    const IdT* src = reinterpret_cast<const IdT*>(a.ids) + src_off;
    const uint32_t base = a.unit_posbase[u];
    uint64_t key[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const uint32_t i = (uint32_t)r * 32u + lane;
        key[r] = i < n ? ((load_id(src + i) << 32) | (uint64_t)(base + i)) : ~0ull;
    }
#pragma unroll
    for (uint32_t k = 2; k <= 64; k <<= 1) {
#pragma unroll
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            if (j == 32) {  // partner = the lane's other key; k == 64: ascending
                const uint64_t lo = key[0] < key[1] ? key[0] : key[1], hi = key[0] < key[1] ? key[1] : key[0];
                key[0] = lo, key[1] = hi;
            } else {
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const uint32_t i = (uint32_t)r * 32u + lane;
                    const uint64_t other = __shfl_xor_sync(0xffffffffu, key[r], j);
                    const bool up = (i & k) == 0, lower = (lane & j) == 0;
                    const uint64_t mn = key[r] < other ? key[r] : other, mx = key[r] < other ? other : key[r];
                    key[r] = (lower == up) ? mn : mx;
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const uint32_t i = (uint32_t)r * 32u + lane;
        if (i < n) {
            a.sorted_ids[src_off + i] = (uint32_t)(key[r] >> 32);
            a.sort_idx[src_off + i] = (uint32_t)key[r];
        }
    }
}

template <typename IdT>
__global__ void __launch_bounds__(256) k_sort_units(SortArgs a) {
    __shared__ uint64_t skeys[kSortSmem];
    for (uint32_t u = blockIdx.x; u < a.nunits; u += gridDim.x) {
        uint32_t n = a.unit_n[u];
        if (n <= kSortWarp) continue;  // empty, or sorted by k_sort_small
        uint64_t src_off = a.unit_src[u];
        const IdT* src = reinterpret_cast<const IdT*>(a.ids) + src_off;
        uint32_t npad = 1;
        while (npad < n) npad <<= 1;
        uint64_t* keys = npad <= kSortSmem ? skeys : a.big_scratch + (size_t)blockIdx.x * kMaxUnit;
        uint32_t base = a.unit_posbase[u];
        for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x)
            keys[i] = i < n ? ((load_id(src + i) << 32) | (uint64_t)(base + i)) : ~0ull;
        __syncthreads();
        for (uint32_t k = 2; k <= npad; k <<= 1) {
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x) {
                    uint32_t p = i ^ j;
                    if (p > i) {
                        uint64_t x = keys[i], y = keys[p];
                        bool up = (i & k) == 0;
                        if ((x > y) == up) {
                            keys[i] = y;
                            keys[p] = x;
                        }
                    }
                }
                __syncthreads();
            }
        }
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            uint64_t kv = keys[i];
            a.sorted_ids[src_off + i] = (uint32_t)(kv >> 32);
            a.sort_idx[src_off + i] = (uint32_t)kv;
        }
        __syncthreads();
    }
}


inline uint32_t grid_for(uint64_t threads) { return (uint32_t)((threads + kThreads - 1) / kThreads); }

// both sort kernels: units of <= 64 ids by one warp each, longer ones by one CTA each (only launched when there
// are any: `any_big`)
inline int launch_sorts(idc_ctx* c, const SortArgs& s, int id_bytes, uint32_t sort_grid, bool any_big) {
    LaunchScope ls(c, "k_sort_units");
    if (s.nunits) {
        if (id_bytes == 8)
            k_sort_small<int64_t><<<grid_for((uint64_t)s.nunits * 32), kThreads, 0, c->stream>>>(s);
        else
            k_sort_small<uint32_t><<<grid_for((uint64_t)s.nunits * 32), kThreads, 0, c->stream>>>(s);
    }
    if (any_big) {
        if (id_bytes == 8)
            k_sort_units<int64_t><<<sort_grid, 256, 0, c->stream>>>(s);
        else
            k_sort_units<uint32_t><<<sort_grid, 256, 0, c->stream>>>(s);
    }
    return IDC_OK;
}

// Per-unit metadata for the units described by (m.unit_src, m.unit_n).
struct MetaPlan {
    std::vector<uint32_t> tile_unit, tile_idx;
    std::vector<uint64_t> tile_first;  // per unit (nunits + 1): index of its first tile
    bool identity = false;             // every unit <= kMetaTile ids: tile t = unit t, no tables
};

inline void plan_unit_meta(const std::vector<uint32_t>& unit_n, MetaPlan& p) {
    p.tile_unit.clear();
    p.tile_idx.clear();
    p.identity = true;
    for (uint64_t u = 0; u < unit_n.size() && p.identity; u++) p.identity = unit_n[u] <= kMetaTile;
    if (p.identity) {  // a million graph rows: no per-tile tables to build or upload
        p.tile_first.resize(unit_n.size() + 1);
        for (uint64_t u = 0; u <= unit_n.size(); u++) p.tile_first[u] = u;
        return;
    }
    p.tile_first.assign(unit_n.size() + 1, 0);
    p.tile_unit.reserve(unit_n.size() + unit_n.size() / 4);
    p.tile_idx.reserve(unit_n.size() + unit_n.size() / 4);
    for (uint64_t u = 0; u < unit_n.size(); u++) {
        p.tile_first[u] = p.tile_unit.size();
        for (uint32_t t = 0; (uint64_t)t * kMetaTile < unit_n[u]; t++) {
            p.tile_unit.push_back((uint32_t)u);
            p.tile_idx.push_back(t);
        }
    }
    p.tile_first[unit_n.size()] = p.tile_unit.size();
}

// the kernels for units [u0, u1); m.tile_unit / m.tile_idx are the uploaded tables of the whole plan, unit_lo /
// unit_hi hold 0xffffffff / 0 for these units
template <int kDummy = 0>
int launch_unit_meta(idc_ctx* c, MetaArgs m, int id_bytes, uint64_t u0, uint64_t u1, const MetaPlan& p) {
    if (u1 <= u0) return IDC_OK;
    m.tile_base = (uint32_t)p.tile_first[u0];
    m.ntiles = (uint32_t)(p.tile_first[u1] - p.tile_first[u0]);
    m.unit_base = (uint32_t)u0;
    m.unit_count = (uint32_t)(u1 - u0);
    LaunchScope ls(c, "k_unit_meta");
    if (m.ntiles) {
        if (id_bytes == 8)
            k_unit_meta<int64_t><<<grid_for((uint64_t)m.ntiles * 32), kThreads, 0, c->stream>>>(m);
        else
            k_unit_meta<uint32_t><<<grid_for((uint64_t)m.ntiles * 32), kThreads, 0, c->stream>>>(m);
    }
    k_unit_meta_finish<<<grid_for(m.unit_count), kThreads, 0, c->stream>>>(m);
    return check_last_launch("k_unit_meta");
}

// everything at once; the tile tables go to c->scratch (free at this point of an Elias-Fano encode call)
inline int run_unit_meta(idc_ctx* c, MetaArgs m, const std::vector<uint32_t>& unit_n_host, int id_bytes) {
    const uint64_t nu = unit_n_host.size();
    if (nu == 0) return IDC_OK;
    MetaPlan p;
    plan_unit_meta(unit_n_host, p);
    const uint64_t nt = p.tile_unit.size();
    IDC_REQUIRE(nt < (1ull << 32), IDC_ERR_ARG, "too many metadata tiles");
    uint32_t *d_tu = nullptr, *d_ti = nullptr;
    if (!p.identity) {
        IDC_TRY(c->scratch.reserve(nt * 8 + 256));
        d_tu = c->scratch.as<uint32_t>();
        d_ti = d_tu + nt;
        IDC_TRY(upload(c, d_tu, p.tile_unit));
        IDC_TRY(upload(c, d_ti, p.tile_idx));
    }
    IDC_CUDA(cudaMemsetAsync(m.unit_lo, 0xff, nu * 4, c->stream));
    IDC_CUDA(cudaMemsetAsync(m.unit_hi, 0, nu * 4, c->stream));
    m.tile_unit = d_tu;
    m.tile_idx = d_ti;
    IDC_TRY(launch_unit_meta(c, m, id_bytes, 0, nu, p));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    return IDC_OK;
}

// the same for units that are all known to hold <= kMetaTile ids (graph rows): tile t = unit t, no host tables, no
// synchronisation -- the caller reads the status word together with its other results
inline int run_unit_meta_small(idc_ctx* c, MetaArgs m, uint64_t nu, int id_bytes) {
    if (nu == 0) return IDC_OK;
    MetaPlan p;
    p.identity = true;
    p.tile_first.resize(2);
    IDC_CUDA(cudaMemsetAsync(m.unit_lo, 0xff, nu * 4, c->stream));
    IDC_CUDA(cudaMemsetAsync(m.unit_hi, 0, nu * 4, c->stream));
    m.tile_unit = nullptr;
    m.tile_idx = nullptr;
    m.tile_base = 0;
    m.ntiles = (uint32_t)nu;
    m.unit_base = 0;
    m.unit_count = (uint32_t)nu;
    LaunchScope ls(c, "k_unit_meta");
    if (id_bytes == 8)
        k_unit_meta<int64_t><<<grid_for(nu * 32), kThreads, 0, c->stream>>>(m);
    else
        k_unit_meta<uint32_t><<<grid_for(nu * 32), kThreads, 0, c->stream>>>(m);
    k_unit_meta_finish<<<grid_for(nu), kThreads, 0, c->stream>>>(m);
    return check_last_launch("k_unit_meta");
}

int status_to_error(uint32_t st, const char* what) {
    if (st & kStWide) {
        set_error("%s: an id does not fit 32 bits (the sm_100a path codes ids < 2^32)", what);
        return IDC_ERR_DOMAIN;
    }
    if (st & kStUnsorted) {
        set_error("%s: IDC_F_SORTED was given but a list is not ascending", what);
        return IDC_ERR_DOMAIN;
    }
    if (st & kStRange) {
        set_error("%s: row number out of range", what);
        return IDC_ERR_ARG;
    }
    if (st & (kStOverlay | kStMtDraws | kStScratch)) {
        set_error("%s: stream invariant violated (status 0x%x)", what, st);
        return IDC_ERR_STREAM;
    }
    return IDC_OK;
}


}  // namespace
