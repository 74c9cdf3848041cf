// wt_kernels.cu -- the wavelet-tree flavour of the IVF plugin surface on sm_100a
// (CompressedIDInvertedListsWaveletTree, custom_invlist_cpp/custom_invlists_impl.cpp:346-397).
//
// Build = `levels` streaming passes over the sequence "list number of id i" (4 bytes per id, ping-pong in the
// context workspace). Per level:
//   k_wt_level_bits     one warp per 512-id rank block: 16 coalesced 128-byte reads, __ballot_sync packs the
//                       level's bit of 32 symbols per instruction, lanes 0..15 store the block's 64 bytes, the
//                       block's popcount goes to the directory scratch
//   k_wt_scan_tiles / k_wt_scan_parts   exclusive prefix sum of the block popcounts (warp shuffles)
//   k_wt_directory      final rank directory + the select samples for ones and zeros
//   k_wt_level_scatter  the stable partition: destination = zeros-before (bit 0) or zeros + ones-before (bit 1),
//                       ones-before = directory entry + running ballot popcount
// Queries:
//   k_wt_select         one thread per (list, offset): wt_select of wt_core.cuh
//   k_wt_decode         a few whole lists: one thread per output id
//   k_wt_replay / k_wt_emit   most of the index: the partitions replayed on the ids themselves (streaming passes)
// All HBM-bound integer work; no tensor cores.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "idc_file.h"
#include "idc_host.h"
#include "idc_scan.cuh"
#include "wt_core.cuh"

using namespace idc;

namespace {

constexpr int kThreads = 128;
constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kWtStHole = 1u;      // an id of [0, ntotal) belongs to no list (or two lists share one)
constexpr uint32_t kWtStRange = 2u;     // id >= ntotal (custom_invlists_impl.cpp:359 assert)
constexpr uint32_t kWtStUnsorted = 4u;  // list not strictly ascending (custom_invlists_impl.cpp:358 assert)
constexpr uint32_t kWtStCorrupt = 8u;   // a select walked off the directory
constexpr uint32_t kScanThreads = 256;
constexpr uint32_t kScanPerThread = 4;
constexpr uint32_t kScanTile = kScanThreads * kScanPerThread;

inline uint32_t grid_for(uint64_t threads, uint32_t per_cta = kThreads) {
    return (uint32_t)((threads + per_cta - 1) / per_cta);
}

template <typename T>
int dev_alloc(idc_ctx* c, T** p, size_t count, uint64_t* acct) {
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    IDC_TRY(c->pool_alloc(reinterpret_cast<void**>(p), bytes));
    if (acct) *acct += bytes;
    return IDC_OK;
}

// largest s in [0, cnt) with off[s] <= e (off has cnt + 1 ascending entries, off[cnt] > e); skips empty lists
__device__ __forceinline__ uint64_t find_owner(const uint64_t* __restrict__ off, uint64_t cnt, uint64_t e) {
    uint64_t lo = 0, hi = cnt - 1;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo + 1) / 2;
        if (__ldg(off + mid) <= e)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

// Owners of the 512 consecutive elements [base, base + 512) of a warp's tile, 32 per round (element of lane j in
// round t: base + 32 t + j, clamped to last): ONE binary search per tile; after that a lane only checks that it is
// still inside the list of the element before its round (lists are long compared to a round), else it searches.
__device__ __forceinline__ void tile_owners(const uint64_t* __restrict__ off, uint64_t cnt, uint64_t base, uint64_t last,
                                            uint32_t lane, uint32_t (&own)[16]) {
    uint64_t s = 0;
    if (lane == 0) s = find_owner(off, cnt, base);
    s = __shfl_sync(kFull, s, 0);
#pragma unroll
    for (int t = 0; t < 16; t++) {
        uint64_t e = base + (uint64_t)t * 32 + lane;
        if (e > last) e = last;
        uint64_t o = e < __ldg(off + s + 1) ? s : find_owner(off, cnt, e);
        own[t] = (uint32_t)o;
        s = __shfl_sync(kFull, o, 31);
    }
}

// ---- S[id] = list_no (custom_invlists_impl.cpp:354-362), with the reference's asserts as status bits.
// One warp per tile of 512 elements of the CSR id array.
template <typename IdT>
__global__ void __launch_bounds__(kThreads) k_wt_fill(const IdT* __restrict__ ids, const uint64_t* __restrict__ list_off,
                                                      uint32_t nlist, uint64_t n, uint32_t* __restrict__ seq,
                                                      uint32_t* status) {
    const uint64_t tile = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t base = tile << 9;
    if (base >= n) return;
    uint64_t id[16];
#pragma unroll
    for (int t = 0; t < 16; t++) {
        uint64_t e = base + (uint64_t)t * 32 + lane;
        id[t] = e < n ? load_id(ids + e) : 0;  // a negative int64 id becomes >= 2^63: out of range
    }
    uint32_t own[16];
    tile_owners(list_off, nlist, base, n - 1, lane, own);
    uint64_t before = base ? load_id(ids + base - 1) : 0;  // the id in front of the round's first element
    uint32_t st = 0;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        uint64_t e = base + (uint64_t)t * 32 + lane;
        uint64_t prev = __shfl_up_sync(kFull, id[t], 1);
        if (lane == 0) prev = before;
        before = __shfl_sync(kFull, id[t], 31);
        if (e < n) {
            if (id[t] >= n)
                st |= kWtStRange;
            else
                seq[id[t]] = own[t];
            if (e > __ldg(list_off + own[t]) && prev >= id[t]) st |= kWtStUnsorted;
        }
    }
    if (st) atomicOr(status, st);
}

// ---- the same through id-range buckets, for sequences that do not fit L2: a scattered 4-byte store costs a whole
// DRAM sector (measured 22 G stores/s), so the (id, list) pairs are first distributed into buckets of 2^bucket_log
// consecutive ids -- the lists partition [0, n), hence bucket b receives exactly its 2^bucket_log pairs and owns
// the slots [b << bucket_log, ...) of the pair array, no histogram pass -- and then applied bucket after bucket,
// every store of a bucket landing in one L2-resident window of the sequence.
constexpr int kDistThreads = 256;      // 8 warps = 8 tiles of 512 elements per CTA
constexpr uint32_t kMaxBuckets = 2048;

template <typename IdT>
__global__ void __launch_bounds__(kDistThreads) k_wt_distribute(const IdT* __restrict__ ids, const uint64_t* __restrict__ list_off,
                                                                uint32_t nlist, uint64_t n, uint32_t bucket_log,
                                                                uint32_t nbuckets, unsigned long long* __restrict__ cursor,
                                                                uint64_t* __restrict__ pairs, uint32_t* status) {
    __shared__ uint32_t s_cnt[kMaxBuckets];
    __shared__ unsigned long long s_base[kMaxBuckets];
    for (uint32_t b = threadIdx.x; b < nbuckets; b += kDistThreads) s_cnt[b] = 0;
    __syncthreads();
    const uint64_t tile = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t base = tile << 9;
    const bool live = base < n;  // warp-uniform; dead warps still reach the barriers
    uint64_t id[16];
    uint32_t own[16], rank[16];
    uint32_t st = 0;
    if (live) {
#pragma unroll
        for (int t = 0; t < 16; t++) {
            uint64_t e = base + (uint64_t)t * 32 + lane;
            id[t] = e < n ? load_id(ids + e) : ~0ull;
        }
        tile_owners(list_off, nlist, base, n - 1, lane, own);
        uint64_t before = base ? load_id(ids + base - 1) : 0;
#pragma unroll
        for (int t = 0; t < 16; t++) {
            uint64_t e = base + (uint64_t)t * 32 + lane;
            uint64_t prev = __shfl_up_sync(kFull, id[t], 1);
            if (lane == 0) prev = before;
            before = __shfl_sync(kFull, id[t], 31);
            rank[t] = 0;
            if (e < n) {
                const bool bad = id[t] >= n;
                if (bad)
                    st |= kWtStRange;
                else if (e > __ldg(list_off + own[t]) && prev >= id[t])
                    st |= kWtStUnsorted;
                if (bad)
                    id[t] = ~0ull;  // dropped
                else
                    rank[t] = atomicAdd(&s_cnt[id[t] >> bucket_log], 1u);
            }
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbuckets; b += kDistThreads)
        s_base[b] = s_cnt[b] ? atomicAdd(cursor + b, (unsigned long long)s_cnt[b]) : 0ull;
    __syncthreads();
    if (live) {
#pragma unroll
        for (int t = 0; t < 16; t++) {
            if (id[t] == ~0ull) continue;
            uint64_t b = id[t] >> bucket_log;
            uint64_t slot = s_base[b] + rank[t];                     // offset inside the bucket
            uint64_t cap = min((uint64_t)1 << bucket_log, n - (b << bucket_log));
            if (slot < cap)
                pairs[(b << bucket_log) + slot] = (id[t] << 32) | own[t];
            else
                st |= kWtStHole;  // more ids than the bucket has room for: some id appears twice
        }
    }
    if (st) atomicOr(status, st);
}

__global__ void __launch_bounds__(kThreads) k_wt_apply(const uint64_t* __restrict__ pairs, const unsigned long long* __restrict__ cursor,
                                                       uint64_t n, uint32_t bucket_log, uint32_t* __restrict__ seq) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t b = i >> bucket_log;
    if (i - (b << bucket_log) >= cursor[b]) return;  // slot never filled (invalid input; reported as a hole)
    const uint64_t p = __ldg(pairs + i);
    seq[p >> 32] = (uint32_t)p;
}

// ---- one level: bits of every rank block + its popcount (32-bit positions, see k_wt_level_scatter)
template <typename SymT>
__global__ void __launch_bounds__(kThreads) k_wt_level_bits(const SymT* __restrict__ seq, uint64_t n64, uint64_t nblk64,
                                                            uint32_t shift, uint32_t check_holes,
                                                            uint64_t* __restrict__ bits, uint32_t* __restrict__ ones,
                                                            uint32_t* status) {
    const uint32_t n = (uint32_t)n64, nblk = (uint32_t)nblk64;
    const uint32_t blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // warp-uniform
    const uint32_t lane = threadIdx.x & 31u;
    if (blk >= nblk) return;
    const uint32_t base = blk << kWtBlockLog;
    uint32_t v[16];
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const uint32_t i = base + (uint32_t)t * 32 + lane;
        v[t] = i < n ? (uint32_t)__ldg(seq + i) : 0u;
    }
    uint32_t mine = 0, cnt = 0;
    bool hole = false;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const uint32_t i = base + (uint32_t)t * 32 + lane;
        hole |= sizeof(SymT) == 4 && check_holes && i < n && v[t] == kWtHole;
        const uint32_t m = __ballot_sync(kFull, (v[t] >> shift) & 1u);
        if (lane == (uint32_t)t) mine = m;
        cnt += (uint32_t)__popc(m);
    }
    // 16 32-bit words = the block's 8 little-endian 64-bit words
    if (lane < 16) reinterpret_cast<uint32_t*>(bits)[(size_t)blk * 16 + lane] = mine;
    if (lane == 0) ones[blk] = cnt;
    if (hole) atomicOr(status, kWtStHole);
}

__device__ __forceinline__ uint32_t warp_inclusive(uint32_t x, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(kFull, x, d);
        if (lane >= (uint32_t)d) x += t;
    }
    return x;
}

// data[0..E) -> tile-local exclusive prefix, part[tile] = tile total (tiles of 1024 entries)
__global__ void __launch_bounds__(kScanThreads) k_wt_scan_tiles(uint32_t* __restrict__ data, uint64_t E,
                                                                uint32_t* __restrict__ part) {
    __shared__ uint32_t wsum[kScanThreads / 32];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanPerThread;
    uint32_t v[kScanPerThread], s = 0;
#pragma unroll
    for (uint32_t k = 0; k < kScanPerThread; k++) {
        v[k] = base + k < E ? data[base + k] : 0u;
        s += v[k];
    }
    uint32_t inc = warp_inclusive(s, lane);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t x = lane < kScanThreads / 32 ? wsum[lane] : 0u;
        uint32_t xi = warp_inclusive(x, lane);
        if (lane < kScanThreads / 32) wsum[lane] = xi - x;
    }
    __syncthreads();
    uint32_t off = wsum[w] + inc - s;
#pragma unroll
    for (uint32_t k = 0; k < kScanPerThread; k++) {
        if (base + k < E) data[base + k] = off;
        off += v[k];
    }
    if (threadIdx.x == kScanThreads - 1) part[blockIdx.x] = off;
}

// exclusive prefix sum of the tile totals, one CTA
__global__ void __launch_bounds__(1024) k_wt_scan_parts(uint32_t* part, uint32_t P) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t total_s;
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < P; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t x = i < P ? part[i] : 0u;
        uint32_t inc = warp_inclusive(x, lane);
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            uint32_t y = wsum[lane];
            uint32_t yi = warp_inclusive(y, lane);
            wsum[lane] = yi - y;
            if (lane == 31) total_s = yi;
        }
        __syncthreads();
        if (i < P) part[i] = carry + wsum[w] + inc - x;
        carry += total_s;
        __syncthreads();
    }
}

// final directory of a level: rank[j] = ones before block j (j = nblk: all ones), and the select samples
__global__ void __launch_bounds__(kThreads) k_wt_directory(const uint32_t* __restrict__ local, const uint32_t* __restrict__ part,
                                                           uint64_t nblk, uint64_t n, uint32_t* __restrict__ rank,
                                                           uint32_t* __restrict__ sel1, uint32_t* __restrict__ sel0) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j > nblk) return;
    uint32_t o0 = local[j] + part[j / kScanTile];
    rank[j] = o0;
    if (j == nblk) return;
    uint32_t o1 = local[j + 1] + part[(j + 1) / kScanTile];
    WtDirEntry d = wt_dir_entry(j, nblk, n, o0, o1);
    if (d.has1) sel1[d.m1] = (uint32_t)j;
    if (d.has0) sel0[d.m0] = (uint32_t)j;
}

// stable partition of a level by its bit: zeros keep their order in [0, z), ones theirs in [z, n). All positions in
// 32 bits (ntotal < 2^32 - 4096 is a precondition of the structure): the 64-bit version of this pass was issue-bound.
template <typename InT, typename OutT>
__global__ void __launch_bounds__(kThreads) k_wt_level_scatter(const InT* __restrict__ seq, uint64_t n64, uint64_t nblk64,
                                                               uint32_t shift, const uint32_t* __restrict__ rank,
                                                               OutT* __restrict__ next) {
    const uint32_t n = (uint32_t)n64, nblk = (uint32_t)nblk64;
    const uint32_t blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (blk >= nblk) return;
    const uint32_t base = blk << kWtBlockLog;
    const uint32_t z = n - __ldg(rank + nblk);
    uint32_t r1 = __ldg(rank + blk);
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t v[16];
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const uint32_t i = base + (uint32_t)t * 32 + lane;
        v[t] = i < n ? (uint32_t)__ldg(seq + i) : 0u;
    }
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const uint32_t i = base + (uint32_t)t * 32 + lane;
        const bool valid = i < n;
        const uint32_t b = (v[t] >> shift) & 1u;
        const uint32_t m = __ballot_sync(kFull, valid && b);
        const uint32_t ones_before = r1 + (uint32_t)__popc(m & lt);
        if (valid) next[b ? z + ones_before : i - ones_before] = (OutT)v[t];
        r1 += (uint32_t)__popc(m);
    }
}

struct WtSelArgs {
    WtView v;
    const uint64_t* list_off;
    uint64_t nlist;
    const uint64_t* q_list;
    const uint64_t* q_off;
    int64_t* out;
    uint64_t nq;
};

__global__ void __launch_bounds__(kThreads) k_wt_select(WtSelArgs a) {
    uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.nq) return;
    uint64_t L = a.q_list[q], k = a.q_off[q];
    int64_t r = -1;
    if (L < a.nlist && k < a.list_off[L + 1] - a.list_off[L]) r = (int64_t)wt_select(a.v, (uint32_t)L, k);
    a.out[q] = r;
}

struct WtDecArgs {
    WtView v;
    const uint64_t* sel;      // selected list numbers (NULL = all lists in order)
    const uint64_t* out_off;  // nsel + 1 offsets of the output
    uint64_t nsel, total;
    void* out;
    uint32_t* status;
};

template <typename OutT>
__global__ void __launch_bounds__(kThreads) k_wt_decode(WtDecArgs a) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.total) return;
    const uint64_t lo = find_owner(a.out_off, a.nsel, e);
    uint64_t L = a.sel ? __ldg(a.sel + lo) : lo;
    uint64_t id = wt_select(a.v, (uint32_t)L, e - __ldg(a.out_off + lo));
    if (id == ~0ull) atomicOr(a.status, kWtStCorrupt);
    reinterpret_cast<OutT*>(a.out)[e] = (OutT)id;
}

// ---- get_ids of (nearly) everything as `levels` streaming passes: the build's stable partition replayed on a
// payload (the id itself), the bit of every element read back from the level's bit vector. After the last level
// payload[start[c] + k] is id number k of list c.
template <bool kIota>
__global__ void __launch_bounds__(kThreads) k_wt_replay(const uint32_t* __restrict__ in, uint64_t n, uint64_t nblk,
                                                        const uint64_t* __restrict__ bits, const uint32_t* __restrict__ rank,
                                                        uint32_t* __restrict__ out) {
    const uint64_t blk = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (blk >= nblk) return;
    const uint64_t base = blk << kWtBlockLog;
    const uint64_t z = n - __ldg(rank + nblk);
    uint64_t r1 = __ldg(rank + blk);
    const uint32_t word = lane < 16 ? __ldg(reinterpret_cast<const uint32_t*>(bits) + blk * 16 + lane) : 0u;
    uint32_t v[16];
#pragma unroll
    for (int t = 0; t < 16; t++) {
        uint64_t i = base + (uint64_t)t * 32 + lane;
        v[t] = kIota ? (uint32_t)i : (i < n ? __ldg(in + i) : 0u);
    }
#pragma unroll
    for (int t = 0; t < 16; t++) {
        uint64_t i = base + (uint64_t)t * 32 + lane;
        uint32_t m = __shfl_sync(kFull, word, t);  // bits past n are zero: they were built from zero symbols
        uint32_t b = (m >> lane) & 1u;
        uint64_t ones_before = r1 + (uint32_t)__popc(m & ((1u << lane) - 1u));
        if (i < n) out[wt_partition_dest(i, b, z, ones_before)] = v[t];
        r1 += (uint32_t)__popc(m);
    }
}

// Two levels per pass. The elements of a tile with bit 0 at level v sit next to each other at level v + 1 (stable
// partition: positions [base - r1, base - r1 + zeros of the tile)), those with bit 1 likewise ([z + r1, ...)): each
// run is at most 512 positions long, hence inside a window of two rank blocks of level v + 1 -- 32 words, one per
// lane, with a warp prefix sum of their popcounts on top of the window's directory entry. An element then finds its
// bit and the ones before it at level v + 1 with two shuffles per side, and moves straight to its place at level
// v + 2: the payload crosses HBM once for two levels.
template <bool kIota>
__global__ void __launch_bounds__(kThreads, 12) k_wt_replay2(const uint32_t* __restrict__ in, uint32_t n, uint32_t nblk,
                                                         const uint64_t* __restrict__ bits0, const uint32_t* __restrict__ rank0,
                                                         const uint64_t* __restrict__ bits1, const uint32_t* __restrict__ rank1,
                                                         uint32_t* __restrict__ out) {
    // per warp: [0, 32) window words of the zero side, [32, 64) of the one side, [64, 128) their exclusive prefixes
    __shared__ uint32_t s_win[kThreads / 32][128];
    const uint32_t blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // all positions fit 32 bits (ntotal < 2^32 - 4096)
    const uint32_t lane = threadIdx.x & 31u;
    if (blk >= nblk) return;
    uint32_t* win = s_win[threadIdx.x >> 5];
    const uint32_t base = blk << kWtBlockLog;
    const uint32_t z0 = n - __ldg(rank0 + nblk), z1 = n - __ldg(rank1 + nblk);
    const uint32_t r1 = __ldg(rank0 + blk);
    const uint32_t word = lane < 16 ? __ldg(reinterpret_cast<const uint32_t*>(bits0) + (size_t)blk * 16 + lane) : 0u;
    uint32_t v[16];
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const uint32_t i = base + (uint32_t)t * 32 + lane;
        v[t] = kIota ? i : (i < n ? __ldg(in + i) : 0u);
    }
    // the two windows of level v + 1
    const uint32_t zs = base - r1, os = z0 + r1;  // where the tile's zeros / ones start at level v + 1
    const uint32_t zb = zs >> kWtBlockLog, ob = os >> kWtBlockLog;
    const uint32_t nwords = nblk * 16;
    const uint32_t* w1 = reinterpret_cast<const uint32_t*>(bits1);
    const uint32_t zi = zb * 16 + lane, oi = ob * 16 + lane;
    const uint32_t wz = zi < nwords ? __ldg(w1 + zi) : 0u, wo = oi < nwords ? __ldg(w1 + oi) : 0u;
    const uint32_t cz = (uint32_t)__popc(wz), co = (uint32_t)__popc(wo);
    uint32_t pz = cz, po = co;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(kFull, pz, d), b = __shfl_up_sync(kFull, po, d);
        if (lane >= (uint32_t)d) pz += a, po += b;
    }
    // exclusive prefix = ones of level v + 1 before the lane's window word
    win[lane] = wz;
    win[32 + lane] = wo;
    win[64 + lane] = pz - cz + __ldg(rank1 + (zb <= nblk ? zb : nblk));
    win[96 + lane] = po - co + __ldg(rank1 + (ob <= nblk ? ob : nblk));
    __syncwarp();
    const uint32_t zwin = zb << kWtBlockLog, owin = ob << kWtBlockLog;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t ones = r1;  // ones of level v before the current round
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const uint32_t i = base + (uint32_t)t * 32 + lane;
        const uint32_t m = __shfl_sync(kFull, word, t);  // bits past n are zero: they were built from zero symbols
        const uint32_t b0 = (m >> lane) & 1u;
        const uint32_t ones_before = ones + (uint32_t)__popc(m & lt);
        const uint32_t p = b0 ? z0 + ones_before : i - ones_before;  // position at level v + 1
        const uint32_t rel = p - (b0 ? owin : zwin);                 // < 1024 for every valid element
        const uint32_t src = ((rel >> 5) & 31u) + (b0 << 5);
        const uint32_t x = win[src], y = win[64 + src];
        const uint32_t sh = rel & 31u;
        const uint32_t b1 = (x >> sh) & 1u;
        const uint32_t ones_before1 = y + (uint32_t)__popc(x & ((1u << sh) - 1u));
        if (i < n) out[b1 ? z1 + ones_before1 : p - ones_before1] = v[t];
        ones += (uint32_t)__popc(m);
    }
}

struct WtEmitArgs {
    const uint32_t* payload;
    const uint32_t* start;
    const uint64_t* sel;      // selected list numbers (NULL = all lists in order)
    const uint64_t* out_off;  // nsel + 1 offsets of the output
    uint64_t nsel, total;
    void* out;
};

// one warp per tile of 512 output ids
template <typename OutT>
__global__ void __launch_bounds__(kThreads) k_wt_emit(WtEmitArgs a) {
    const uint64_t tile = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t base = tile << 9;
    if (base >= a.total) return;
    uint32_t own[16];
    tile_owners(a.out_off, a.nsel, base, a.total - 1, lane, own);
    // lists are long compared to a tile: the two table entries of an element's list are re-read only when the list changes
    uint32_t cur = 0xffffffffu;
    int64_t delta = 0;  // payload position minus output position inside the current list
    uint32_t v[16];
#pragma unroll
    for (int t = 0; t < 16; t++) {
        uint64_t e = base + (uint64_t)t * 32 + lane;
        v[t] = 0;
        if (e < a.total) {
            if (own[t] != cur) {
                cur = own[t];
                uint64_t L = a.sel ? __ldg(a.sel + cur) : cur;
                delta = (int64_t)__ldg(a.start + L) - (int64_t)__ldg(a.out_off + cur);
            }
            v[t] = __ldg(a.payload + (uint64_t)((int64_t)e + delta));
        }
    }
#pragma unroll
    for (int t = 0; t < 16; t++) {
        uint64_t e = base + (uint64_t)t * 32 + lane;
        if (e < a.total) reinterpret_cast<OutT*>(a.out)[e] = (OutT)v[t];
    }
}

}  // namespace

struct idc_wt_blob {
    idc_ctx* ctx = nullptr;
    idc::CtxRef ref;  // declared right after ctx: destroyed last, after the arrays went back to the pool
    uint64_t nlist = 0, total_ids = 0;
    int wt_type = 0;
    WtShape sh{};
    std::vector<uint64_t> list_offsets;
    uint64_t* d_list_off = nullptr;
    uint64_t* d_bits = nullptr;
    uint32_t* d_rank = nullptr;
    uint32_t* d_sel1 = nullptr;
    uint32_t* d_sel0 = nullptr;
    uint32_t* d_start = nullptr;
    // wt_type = 1: d_bits is null, the levels are stored as RRR(63) blocks (wt_core.cuh)
    uint64_t* d_cls = nullptr;       // levels x nblk
    uint32_t* d_ptr = nullptr;       // levels x (nblk + 1)
    uint64_t* d_off = nullptr;       // offset streams, level after level
    uint64_t* d_off_base = nullptr;  // levels + 1
    std::vector<uint64_t> off_base;  // host copy; off_base[levels] = words of d_off in use
    uint64_t device_bytes = 0;
    WtView view() const {
        WtView v{d_bits, d_rank, d_sel1, d_sel0, d_start, sh};
        v.cls = d_cls;
        v.ptr = d_ptr;
        v.off = d_off;
        v.off_base = d_off_base;
        v.binom = ctx ? ctx->d_binom : nullptr;
        return v;
    }
    ~idc_wt_blob() {
        if (!ctx) return;
        ctx->pool_release(d_cls);
        ctx->pool_release(d_ptr);
        ctx->pool_release(d_off);
        ctx->pool_release(d_off_base);
        ctx->pool_release(d_list_off);
        ctx->pool_release(d_bits);
        ctx->pool_release(d_rank);
        ctx->pool_release(d_sel1);
        ctx->pool_release(d_sel0);
        ctx->pool_release(d_start);
    }
};

namespace {

// ---- wt_type = 1: RRR(63) block compression of the finished levels (layout: wt_core.cuh) -----------------------
// One thread per 512-bit rank block. k_rrr_sizes: the eight classes + the tail byte (one word) and the number of
// offset bits of the block; prefix sums over a level give every block its place in the level's stream; k_rrr_encode
// writes the offsets there (atomicOr: neighbouring blocks share words); k_rrr_expand is the inverse.
struct RrrArgs {
    uint64_t* bits;         // levels x words (source of the encoder, destination of the expander)
    uint64_t* cls;          // levels x nblk
    uint32_t* ptr;          // levels x (nblk + 1)
    uint64_t* off;
    const uint64_t* off_base;
    uint64_t* sizes;        // levels x nblk   (encoder scratch)
    const uint64_t* ptr64;  // levels x (nblk + 1)   (encoder scratch: the prefix sums)
    const uint64_t* binom;
    uint64_t nblk, words;
    uint32_t levels;
};

__global__ void __launch_bounds__(kThreads) k_rrr_sizes(RrrArgs a) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.nblk * a.levels) return;
    const uint64_t lev = g / a.nblk, blk = g - lev * a.nblk;
    uint64_t w[8];
    const uint64_t* src = a.bits + lev * a.words + blk * kWtBlockWords;
#pragma unroll
    for (int j = 0; j < 8; j++) w[j] = src[j];
    uint64_t cw = (w[7] >> 56) << 48, total = 0;
#pragma unroll
    for (uint32_t j = 0; j < kRrrPerBlock; j++) {
        const uint32_t k = (uint32_t)popc64(rrr_piece(w, j));
        cw |= (uint64_t)k << (6u * j);
        total += rrr_width(a.binom, k);
    }
    a.cls[g] = cw;
    a.sizes[g] = total;
}

__global__ void __launch_bounds__(kThreads) k_rrr_encode(RrrArgs a) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.nblk * a.levels) return;
    const uint64_t lev = g / a.nblk, blk = g - lev * a.nblk;
    uint64_t w[8];
    const uint64_t* src = a.bits + lev * a.words + blk * kWtBlockWords;
#pragma unroll
    for (int j = 0; j < 8; j++) w[j] = src[j];
    uint64_t p = a.ptr64[lev * (a.nblk + 1) + blk];
    a.ptr[lev * (a.nblk + 1) + blk] = (uint32_t)p;
    if (blk + 1 == a.nblk) a.ptr[lev * (a.nblk + 1) + a.nblk] = (uint32_t)a.ptr64[lev * (a.nblk + 1) + a.nblk];
    unsigned long long* stream = reinterpret_cast<unsigned long long*>(a.off + a.off_base[lev]);
#pragma unroll 1
    for (uint32_t j = 0; j < kRrrPerBlock; j++) {
        const uint64_t x = rrr_piece(w, j);
        const uint32_t W = rrr_width(a.binom, (uint32_t)popc64(x));
        if (W) {
            const uint64_t v = rrr_offset_of(a.binom, x);
            const uint32_t s = (uint32_t)(p & 63u);
            atomicOr(stream + (p >> 6), (unsigned long long)(v << s));
            if (s + W > 64u) atomicOr(stream + (p >> 6) + 1, (unsigned long long)(v >> (64u - s)));
            p += W;
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_rrr_expand(RrrArgs a) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.nblk * a.levels) return;
    const uint64_t lev = g / a.nblk, blk = g - lev * a.nblk;
    const uint64_t cw = a.cls[g];
    uint64_t p = a.ptr[lev * (a.nblk + 1) + blk];
    const uint64_t* stream = a.off + a.off_base[lev];
    uint64_t w[8];
#pragma unroll
    for (int j = 0; j < 8; j++) w[j] = 0;
#pragma unroll 1
    for (uint32_t j = 0; j < kRrrPerBlock; j++) {
        const uint32_t k = (uint32_t)(cw >> (6u * j)) & 63u, W = rrr_width(a.binom, k);
        const uint64_t x = rrr_block_of(a.binom, k, rrr_read(stream, p, W));
        p += W;
        const uint32_t q = kRrrBits * j, i = q >> 6, s = q & 63u;
        // (dynamic word index: the eight words live in local memory here; this kernel runs once per decode call)
        w[i] |= x << s;
        if (s > 1u) w[i + 1] |= x >> (64u - s);
    }
    w[7] |= ((cw >> 48) & 0xffull) << 56;
    uint64_t* dst = a.bits + lev * a.words + blk * kWtBlockWords;
#pragma unroll
    for (int j = 0; j < 8; j++) dst[j] = w[j];
}

// C(n, k) for n, k < 64 and, as row 64, the offset widths ceil(log2 C(63, k))
int wt_rrr_tables(idc_ctx* c) {
    if (c->d_binom) return IDC_OK;
    std::vector<uint64_t> t(65 * 64, 0);
    for (uint32_t n = 0; n < 64; n++) {
        t[n * 64] = 1;
        for (uint32_t k = 1; k <= n; k++) t[n * 64 + k] = t[(n - 1) * 64 + k - 1] + (k <= n - 1 ? t[(n - 1) * 64 + k] : 0);
    }
    for (uint32_t k = 0; k < 64; k++) {
        const uint64_t m = t[63 * 64 + k] - 1;  // offsets 0 .. C(63, k) - 1
        t[64 * 64 + k] = m ? 64 - (uint64_t)__builtin_clzll(m) : 0;
    }
    IDC_CUDA(cudaMalloc(&c->d_binom, t.size() * 8));
    IDC_CUDA(cudaMemcpy(c->d_binom, t.data(), t.size() * 8, cudaMemcpyHostToDevice));
    return IDC_OK;
}

// plain levels (b->d_bits) -> RRR(63) blocks; the plain array goes back to the pool
int wt_compress(idc_ctx* c, idc_wt_blob* b) {
    const WtShape sh = b->sh;
    const uint64_t nb = sh.nblk, L = sh.levels;
    IDC_TRY(wt_rrr_tables(c));
    IDC_TRY(dev_alloc(c, &b->d_cls, L * nb, &b->device_bytes));
    IDC_TRY(dev_alloc(c, &b->d_ptr, L * (nb + 1), &b->device_bytes));
    IDC_TRY(dev_alloc(c, &b->d_off_base, L + 1, &b->device_bytes));
    const size_t scan_bytes = scan_scratch_bytes(nb, 6);
    IDC_TRY(c->ws.reserve(L * nb * 8 + L * (nb + 1) * 8 + scan_bytes + 1024));
    uint64_t* d_sizes = c->ws.as<uint64_t>();
    uint64_t* d_ptr64 = d_sizes + L * nb;
    uint64_t* d_scan = d_ptr64 + L * (nb + 1);
    RrrArgs a{b->d_bits, b->d_cls, b->d_ptr, nullptr, b->d_off_base, d_sizes, d_ptr64, c->d_binom, nb, sh.words, (uint32_t)L};
    {
        LaunchScope ls(c, "k_rrr_sizes");
        k_rrr_sizes<<<grid_for(L * nb), kThreads, 0, c->stream>>>(a);
    }
    IDC_TRY(check_last_launch("k_rrr_sizes"));
    for (uint64_t l0 = 0; l0 < L; l0 += 6) {
        const int m = (int)std::min<uint64_t>(6, L - l0);
        const uint64_t* sin[6];
        uint64_t* sout[6];
        for (int j = 0; j < m; j++) sin[j] = d_sizes + (l0 + j) * nb, sout[j] = d_ptr64 + (l0 + j) * (nb + 1);
        IDC_TRY(device_scan(c, m, sin, sout, nb, d_scan));
    }
    std::vector<uint64_t> bits_of(L);
    for (uint64_t l = 0; l < L; l++)
        IDC_CUDA(cudaMemcpyAsync(&bits_of[l], d_ptr64 + l * (nb + 1) + nb, 8, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    b->off_base.assign(L + 1, 0);
    for (uint64_t l = 0; l < L; l++) {
        IDC_REQUIRE(bits_of[l] < (1ull << 32), IDC_ERR_DOMAIN, "wavelet level %llu: offset stream of 2^32 bits or more", (unsigned long long)l);
        b->off_base[l + 1] = b->off_base[l] + (bits_of[l] + 63) / 64 + 1;  // + 1: a field read may touch the next word
    }
    IDC_TRY(dev_alloc(c, &b->d_off, b->off_base[L], &b->device_bytes));
    IDC_CUDA(cudaMemsetAsync(b->d_off, 0, std::max<uint64_t>(b->off_base[L], 1) * 8, c->stream));
    IDC_CUDA(cudaMemcpyAsync(b->d_off_base, b->off_base.data(), (L + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    a.off = b->d_off;
    {
        LaunchScope ls(c, "k_rrr_encode");
        k_rrr_encode<<<grid_for(L * nb), kThreads, 0, c->stream>>>(a);
    }
    IDC_TRY(check_last_launch("k_rrr_encode"));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    b->device_bytes -= L * sh.words * 8;
    c->pool_release(b->d_bits);
    b->d_bits = nullptr;
    return IDC_OK;
}

// RRR(63) blocks -> plain levels at dst (levels x words)
int wt_expand(idc_ctx* c, const idc_wt_blob* b, uint64_t* dst) {
    const WtShape& sh = b->sh;
    RrrArgs a{dst, b->d_cls, b->d_ptr, b->d_off, b->d_off_base, nullptr, nullptr, c->d_binom, sh.nblk, sh.words, sh.levels};
    {
        LaunchScope ls(c, "k_rrr_expand");
        k_rrr_expand<<<grid_for((uint64_t)sh.levels * sh.nblk), kThreads, 0, c->stream>>>(a);
    }
    return check_last_launch("k_rrr_expand");
}

int wt_status_to_error(uint32_t st, const char* what) {
    if (st & kWtStRange) {
        set_error("%s: an id is negative or >= ntotal (the reference asserts ids < ntotal)", what);
        return IDC_ERR_DOMAIN;
    }
    if (st & kWtStUnsorted) {
        set_error("%s: a list is not strictly ascending (the reference asserts ordered ids)", what);
        return IDC_ERR_DOMAIN;
    }
    if (st & kWtStHole) {
        set_error("%s: the lists do not partition [0, ntotal): an id is missing or appears twice", what);
        return IDC_ERR_DOMAIN;
    }
    if (st & kWtStCorrupt) {
        set_error("%s: select left the directory (corrupt blob)", what);
        return IDC_ERR_STREAM;
    }
    return IDC_OK;
}

template <typename IdT>
int wt_build(idc_ctx* c, idc_wt_blob* b, const IdT* ids_dev) {
    const WtShape sh = b->sh;
    const uint64_t n = sh.n;
    cudaStream_t s = c->stream;
    // workspaces: two copies of the sequence (ping-pong), block popcounts + tile totals
    const uint64_t seq_elems = sh.nblk << kWtBlockLog;
    IDC_TRY(c->ws.reserve(2 * seq_elems * sizeof(uint32_t)));
    uint32_t* seq = c->ws.as<uint32_t>();
    uint32_t* next = seq + seq_elems;
    const uint64_t E = sh.nblk + 1;
    const uint32_t P = (uint32_t)((E + kScanTile - 1) / kScanTile);
    const uint64_t Epad = (E + 31) & ~uint64_t(31);
    // sequences beyond L2 are filled through id-range buckets (IDC_WT_FILL=direct|bucket, IDC_WT_BUCKET_LOG override)
    bool bucketed = n >= (1ull << 25);
    uint32_t bucket_log = 23;  // 2^23 ids = a 32 MB window of the sequence
    if (const char* e = getenv("IDC_WT_FILL")) bucketed = strcmp(e, "bucket") == 0 ? true : strcmp(e, "direct") == 0 ? false : bucketed;
    if (const char* e = getenv("IDC_WT_BUCKET_LOG")) bucket_log = (uint32_t)std::min(31, std::max(4, atoi(e)));
    while (((n - 1) >> bucket_log) + 1 > kMaxBuckets) bucket_log++;
    const uint32_t nbuckets = (uint32_t)(((n - 1) >> bucket_log) + 1);
    const uint64_t nb_pad = (nbuckets + 31) & ~uint64_t(31);
    const uint64_t dir_words = (Epad + P + 64 + 63) & ~uint64_t(63);  // 32-bit words; keeps the 64-bit arrays aligned
    IDC_TRY(c->scratch.reserve(dir_words * sizeof(uint32_t) + (bucketed ? (nb_pad + n) * 8 : 0)));
    uint32_t* local = c->scratch.as<uint32_t>();
    uint32_t* part = local + Epad;
    unsigned long long* cursor = reinterpret_cast<unsigned long long*>(local + dir_words);
    uint64_t* pairs = reinterpret_cast<uint64_t*>(cursor + nb_pad);
    IDC_TRY(c->status.reserve(128));
    uint32_t* d_status = c->status.as<uint32_t>();
    IDC_CUDA(cudaMemsetAsync(d_status, 0, 4, s));
    IDC_CUDA(cudaMemsetAsync(seq, 0xff, n * sizeof(uint32_t), s));
    if (bucketed) {
        IDC_CUDA(cudaMemsetAsync(cursor, 0, nbuckets * 8, s));
        {
            LaunchScope ls(c, "k_wt_distribute");
            k_wt_distribute<IdT><<<grid_for(sh.nblk * 32, kDistThreads), kDistThreads, 0, s>>>(
                    ids_dev, b->d_list_off, (uint32_t)b->nlist, n, bucket_log, nbuckets, cursor, pairs, d_status);
        }
        {
            LaunchScope ls(c, "k_wt_apply");
            k_wt_apply<<<grid_for(n), kThreads, 0, s>>>(pairs, cursor, n, bucket_log, seq);
        }
    } else {
        LaunchScope ls(c, "k_wt_fill");
        k_wt_fill<IdT><<<grid_for(sh.nblk * 32), kThreads, 0, s>>>(ids_dev, b->d_list_off, (uint32_t)b->nlist, n, seq, d_status);
    }
    IDC_TRY(check_last_launch("k_wt_fill"));
    const uint32_t warp_grid = grid_for(sh.nblk * 32);
    // Level 0 reads the 32-bit sequence (holes are 0xffffffff there); its partition already drops the bit it has
    // consumed, so with <= 17 levels every later pass moves 16-bit symbols: half the traffic.
    const bool narrow = sh.levels <= 17;
    bool cur16 = false;
    void* cur = seq;
    void* nxt = next;
    for (uint32_t lev = 0; lev < sh.levels; lev++) {
        const uint32_t shift = sh.levels - 1 - lev;
        uint64_t* bits = b->d_bits + (uint64_t)lev * sh.words;
        uint32_t* rank = b->d_rank + (uint64_t)lev * sh.rank_stride;
        IDC_CUDA(cudaMemsetAsync(local + sh.nblk, 0, 4, s));  // entry nblk of the scan input: becomes the level's ones
        {
            LaunchScope ls(c, "k_wt_level_bits");
            if (cur16)
                k_wt_level_bits<uint16_t><<<warp_grid, kThreads, 0, s>>>((const uint16_t*)cur, n, sh.nblk, shift, 0u, bits, local, d_status);
            else
                k_wt_level_bits<uint32_t><<<warp_grid, kThreads, 0, s>>>((const uint32_t*)cur, n, sh.nblk, shift, lev == 0 ? 1u : 0u, bits, local, d_status);
        }
        {
            LaunchScope ls(c, "k_wt_scan");
            k_wt_scan_tiles<<<P, kScanThreads, 0, s>>>(local, E, part);
        }
        {
            LaunchScope ls(c, "k_wt_scan");
            k_wt_scan_parts<<<1, 1024, 0, s>>>(part, P);
        }
        {
            LaunchScope ls(c, "k_wt_directory");
            k_wt_directory<<<grid_for(E), kThreads, 0, s>>>(local, part, sh.nblk, n, rank,
                                                            b->d_sel1 + (uint64_t)lev * sh.samp_stride,
                                                            b->d_sel0 + (uint64_t)lev * sh.samp_stride);
        }
        if (lev + 1 < sh.levels) {
            LaunchScope ls(c, "k_wt_level_scatter");
            if (cur16)
                k_wt_level_scatter<uint16_t, uint16_t><<<warp_grid, kThreads, 0, s>>>((const uint16_t*)cur, n, sh.nblk, shift, rank, (uint16_t*)nxt);
            else if (narrow)
                k_wt_level_scatter<uint32_t, uint16_t><<<warp_grid, kThreads, 0, s>>>((const uint32_t*)cur, n, sh.nblk, shift, rank, (uint16_t*)nxt);
            else
                k_wt_level_scatter<uint32_t, uint32_t><<<warp_grid, kThreads, 0, s>>>((const uint32_t*)cur, n, sh.nblk, shift, rank, (uint32_t*)nxt);
            cur16 = narrow;
            std::swap(cur, nxt);
        }
        IDC_TRY(check_last_launch("wavelet level"));
    }
    uint32_t st = 0;
    IDC_CUDA(cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, s));
    IDC_CUDA(cudaStreamSynchronize(s));
    return wt_status_to_error(st, "wt_encode");
}

}  // namespace

extern "C" {

int idc_wt_encode(idc_ctx* c, uint64_t nlist, const uint64_t* offsets, const void* ids, int id_bytes, int ids_mem,
                  int wt_type, idc_wt_blob** out) {
    IDC_REQUIRE(c && offsets && out, IDC_ERR_ARG, "idc_wt_encode: null argument");
    IDC_REQUIRE(id_bytes == 8 || id_bytes == 4, IDC_ERR_ARG, "id_bytes must be 4 or 8");
    IDC_REQUIRE(wt_type == 0 || wt_type == 1, IDC_ERR_ARG, "wt_type must be 0 or 1 (custom_invlists_impl.cpp:349)");
    IDC_REQUIRE(nlist <= (1ull << 31), IDC_ERR_ARG, "too many lists");
    *out = nullptr;
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    std::unique_ptr<idc_wt_blob> b(new idc_wt_blob());
    b->ctx = c;
    b->ref.bind(c);
    b->nlist = nlist;
    b->wt_type = wt_type;
    b->list_offsets.resize(nlist + 1);
    for (uint64_t l = 0; l <= nlist; l++) {
        IDC_REQUIRE(l == 0 || offsets[l] >= offsets[l - 1], IDC_ERR_ARG, "offsets must be non-decreasing");
        b->list_offsets[l] = offsets[l] - offsets[0];
    }
    const uint64_t n = b->list_offsets[nlist];
    IDC_REQUIRE(n < (1ull << 32) - 4096, IDC_ERR_ARG, "ntotal must be below 2^32 - 4096 (32-bit positions)");
    IDC_REQUIRE(ids != nullptr || n == 0, IDC_ERR_ARG, "ids is NULL");
    b->total_ids = n;
    b->sh = wt_shape(nlist, n);
    if (n == 0) {  // nothing to index: every list is empty
        b->sh.levels = 0;
        *out = b.release();
        return IDC_OK;
    }
    const WtShape sh = b->sh;
    IDC_TRY(dev_alloc(c, &b->d_list_off, nlist + 1, &b->device_bytes));
    IDC_TRY(dev_alloc(c, &b->d_bits, sh.levels * sh.words, &b->device_bytes));
    IDC_TRY(dev_alloc(c, &b->d_rank, sh.levels * sh.rank_stride, &b->device_bytes));
    IDC_TRY(dev_alloc(c, &b->d_sel1, sh.levels * sh.samp_stride, &b->device_bytes));
    IDC_TRY(dev_alloc(c, &b->d_sel0, sh.levels * sh.samp_stride, &b->device_bytes));
    IDC_TRY(dev_alloc(c, &b->d_start, nlist, &b->device_bytes));
    IDC_CUDA(cudaMemcpyAsync(b->d_list_off, b->list_offsets.data(), (nlist + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    IDC_CUDA(cudaMemsetAsync(b->d_sel1, 0, sh.levels * sh.samp_stride * 4, c->stream));
    IDC_CUDA(cudaMemsetAsync(b->d_sel0, 0, sh.levels * sh.samp_stride * 4, c->stream));
    // lists sit below the last level in bit-reversed order of their numbers
    std::vector<uint32_t> start(nlist);
    {
        std::vector<uint64_t> key(nlist);
        for (uint64_t l = 0; l < nlist; l++) key[l] = ((uint64_t)wt_bitrev((uint32_t)l, sh.levels) << 32) | l;
        std::sort(key.begin(), key.end());
        uint64_t acc = 0;
        for (uint64_t i = 0; i < nlist; i++) {
            uint32_t l = (uint32_t)key[i];
            start[l] = (uint32_t)acc;
            acc += b->list_offsets[l + 1] - b->list_offsets[l];
        }
    }
    IDC_CUDA(cudaMemcpyAsync(b->d_start, start.data(), nlist * 4, cudaMemcpyHostToDevice, c->stream));
    const uint8_t* src = static_cast<const uint8_t*>(ids) + offsets[0] * (uint64_t)id_bytes;
    const void* ids_dev = src;
    if (ids_mem == IDC_MEM_HOST) {
        IDC_TRY(c->stage.reserve(n * id_bytes));
        IDC_CUDA(cudaMemcpyAsync(c->stage.p, src, n * id_bytes, cudaMemcpyHostToDevice, c->stream));
        ids_dev = c->stage.p;
    }
    if (id_bytes == 8)
        IDC_TRY(wt_build<int64_t>(c, b.get(), static_cast<const int64_t*>(ids_dev)));
    else
        IDC_TRY(wt_build<uint32_t>(c, b.get(), static_cast<const uint32_t*>(ids_dev)));
    if (wt_type == 1) IDC_TRY(wt_compress(c, b.get()));  // rrr_vector<63> flavour: the finished levels, block-compressed
    *out = b.release();
    return IDC_OK;
}

int idc_wt_blob_info(const idc_wt_blob* b, idc_wt_info* info) {
    IDC_REQUIRE(b && info, IDC_ERR_ARG, "null argument");
    info->nlist = b->nlist;
    info->total_ids = b->total_ids;
    info->bits_bytes = (uint64_t)b->sh.levels * b->sh.words * 8;
    if (b->wt_type == 1 && b->total_ids)  // classes + tails, block pointers, offset streams
        info->bits_bytes = (uint64_t)b->sh.levels * b->sh.nblk * 8 + (uint64_t)b->sh.levels * (b->sh.nblk + 1) * 4 + b->off_base.back() * 8;
    info->aux_bytes = b->total_ids ? (uint64_t)b->sh.levels * (b->sh.rank_stride + 2 * b->sh.samp_stride) * 4 + b->nlist * 4 : 0;
    info->device_bytes = b->device_bytes;
    info->levels = b->sh.levels;
    info->wt_type = (uint32_t)b->wt_type;
    return IDC_OK;
}

int idc_wt_blob_export(const idc_wt_blob* b, uint64_t* list_offsets, uint64_t* bits, uint32_t* rank, uint32_t* sel1,
                       uint32_t* sel0, uint32_t* start) {
    IDC_REQUIRE(b, IDC_ERR_ARG, "null blob");
    IDC_CUDA(cudaSetDevice(b->ctx->device));
    cudaStream_t s = b->ctx->stream;
    if (list_offsets) memcpy(list_offsets, b->list_offsets.data(), (b->nlist + 1) * 8);
    if (b->total_ids) {
        const WtShape& sh = b->sh;
        if (bits) {
            const uint64_t* src = b->d_bits;
            if (b->wt_type == 1) {  // the plain form of the levels (what idc_wt_blob_import takes), expanded on the way out
                std::lock_guard<std::mutex> lock(b->ctx->mu);
                IDC_TRY(b->ctx->scratch.reserve(sh.levels * sh.words * 8));
                IDC_TRY(wt_expand(b->ctx, b, b->ctx->scratch.as<uint64_t>()));
                src = b->ctx->scratch.as<uint64_t>();
            }
            IDC_CUDA(cudaMemcpyAsync(bits, src, sh.levels * sh.words * 8, cudaMemcpyDeviceToHost, s));
        }
        if (rank) IDC_CUDA(cudaMemcpyAsync(rank, b->d_rank, sh.levels * sh.rank_stride * 4, cudaMemcpyDeviceToHost, s));
        if (sel1) IDC_CUDA(cudaMemcpyAsync(sel1, b->d_sel1, sh.levels * sh.samp_stride * 4, cudaMemcpyDeviceToHost, s));
        if (sel0) IDC_CUDA(cudaMemcpyAsync(sel0, b->d_sel0, sh.levels * sh.samp_stride * 4, cudaMemcpyDeviceToHost, s));
        if (start && b->nlist) IDC_CUDA(cudaMemcpyAsync(start, b->d_start, b->nlist * 4, cudaMemcpyDeviceToHost, s));
    }
    IDC_CUDA(cudaStreamSynchronize(s));
    return IDC_OK;
}

int idc_wt_blob_free(idc_wt_blob* b) {
    if (b) {
        std::lock_guard<std::mutex> lock(b->ctx->mu);
        cudaSetDevice(b->ctx->device);
        delete b;
    }
    return IDC_OK;
}

int idc_wt_select(idc_ctx* c, const idc_wt_blob* b, const uint64_t* list_nos, const uint64_t* offsets_in_list,
                  uint64_t nq, int query_mem, int64_t* ids_out, int out_mem) {
    IDC_REQUIRE(c && b && (nq == 0 || (list_nos && offsets_in_list && ids_out)), IDC_ERR_ARG,
                "idc_wt_select: null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    if (nq == 0) return IDC_OK;
    if (b->total_ids == 0) {  // every query is out of range
        if (out_mem == IDC_MEM_HOST)
            for (uint64_t q = 0; q < nq; q++) ids_out[q] = -1;
        else
            IDC_CUDA(cudaMemsetAsync(ids_out, 0xff, nq * 8, c->stream));
        IDC_CUDA(cudaStreamSynchronize(c->stream));
        return IDC_OK;
    }
    const uint64_t *d_ql = list_nos, *d_qo = offsets_in_list;
    // a few queries from the host (get_single_id, custom_invlists_impl.cpp:380-384): through the context's mailbox --
    // [list numbers | offsets | ids], read and written by the kernel in place: one launch, one synchronisation
    void *mh = nullptr, *md = nullptr;
    if (query_mem == IDC_MEM_HOST && out_mem == IDC_MEM_HOST && nq <= 2048) IDC_TRY(c->mailbox_get(nq * 24, &mh, &md));
    int64_t* out_dev = ids_out;
    if (mh) {
        uint64_t *h = static_cast<uint64_t*>(mh), *d = static_cast<uint64_t*>(md);
        std::memcpy(h, list_nos, nq * 8);
        std::memcpy(h + nq, offsets_in_list, nq * 8);
        d_ql = d;
        d_qo = d + nq;
        out_dev = reinterpret_cast<int64_t*>(d + 2 * nq);
    } else {
        size_t need = (query_mem == IDC_MEM_HOST ? nq * 16 : 0) + (out_mem == IDC_MEM_HOST ? nq * 8 : 0);
        IDC_TRY(c->stage.reserve(need + 256));
        uint8_t* sp = c->stage.as<uint8_t>();
        if (query_mem == IDC_MEM_HOST) {
            IDC_CUDA(cudaMemcpyAsync(sp, list_nos, nq * 8, cudaMemcpyHostToDevice, c->stream));
            IDC_CUDA(cudaMemcpyAsync(sp + nq * 8, offsets_in_list, nq * 8, cudaMemcpyHostToDevice, c->stream));
            d_ql = reinterpret_cast<uint64_t*>(sp);
            d_qo = reinterpret_cast<uint64_t*>(sp + nq * 8);
            sp += nq * 16;
        }
        if (out_mem == IDC_MEM_HOST) out_dev = reinterpret_cast<int64_t*>(sp);
    }
    WtSelArgs a{b->view(), b->d_list_off, b->nlist, d_ql, d_qo, out_dev, nq};
    {
        LaunchScope ls(c, "k_wt_select");
        k_wt_select<<<grid_for(nq), kThreads, 0, c->stream>>>(a);
    }
    IDC_TRY(check_last_launch("k_wt_select"));
    if (out_mem == IDC_MEM_HOST && !mh)
        IDC_CUDA(cudaMemcpyAsync(ids_out, out_dev, nq * 8, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    if (mh) std::memcpy(ids_out, static_cast<uint64_t*>(mh) + 2 * nq, nq * 8);
    return IDC_OK;
}

int idc_wt_decode(idc_ctx* c, const idc_wt_blob* b, const uint64_t* list_nos, uint64_t nsel, void* ids_out, int id_bytes,
                  int out_mem, uint64_t* out_offsets) {
    IDC_REQUIRE(c && b, IDC_ERR_ARG, "idc_wt_decode: null argument");
    IDC_REQUIRE(id_bytes == 8 || id_bytes == 4, IDC_ERR_ARG, "id_bytes must be 4 or 8");
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    uint64_t total = 0;
    const uint64_t* d_sel = nullptr;
    const uint64_t* d_off = nullptr;
    if (list_nos == nullptr) {
        nsel = b->nlist;
        total = b->total_ids;
        d_off = b->d_list_off;
        if (out_offsets) memcpy(out_offsets, b->list_offsets.data(), (b->nlist + 1) * 8);
    } else {
        std::vector<uint64_t> tab(2 * nsel + 1);  // selected lists, then their output offsets
        for (uint64_t i = 0; i < nsel; i++) {
            uint64_t L = list_nos[i];
            IDC_REQUIRE(L < b->nlist, IDC_ERR_ARG, "list_no out of range");
            tab[i] = L;
            tab[nsel + i] = total;
            total += b->list_offsets[L + 1] - b->list_offsets[L];
        }
        tab[2 * nsel] = total;
        if (out_offsets) memcpy(out_offsets, tab.data() + nsel, (nsel + 1) * 8);
        if (total) {
            IDC_TRY(c->meta.reserve(tab.size() * 8));
            IDC_CUDA(cudaMemcpyAsync(c->meta.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, c->stream));
            IDC_CUDA(cudaStreamSynchronize(c->stream));  // tab goes out of scope
            d_sel = c->meta.as<uint64_t>();
            d_off = d_sel + nsel;
        }
    }
    if (total == 0) return IDC_OK;
    IDC_REQUIRE(ids_out != nullptr, IDC_ERR_ARG, "ids_out is NULL");
    void* out_dev = ids_out;
    if (out_mem == IDC_MEM_HOST) {
        IDC_TRY(c->stage.reserve(total * id_bytes));
        out_dev = c->stage.p;
    }
    IDC_TRY(c->status.reserve(128));
    uint32_t* d_status = c->status.as<uint32_t>();
    IDC_CUDA(cudaMemsetAsync(d_status, 0, 4, c->stream));
    // A large share of the index: replay the partitions on the ids themselves (streaming passes, cost independent
    // of how many lists are asked for). Few lists: one select walk per id. IDC_WT_DECODE=select|replay forces a path.
    bool replay = total >= b->total_ids / 8;
    if (const char* e = getenv("IDC_WT_DECODE")) replay = strcmp(e, "replay") == 0 ? true : strcmp(e, "select") == 0 ? false : replay;
    if (replay) {
        const WtShape& sh = b->sh;
        const uint64_t seq_elems = sh.nblk << kWtBlockLog;
        IDC_TRY(c->ws.reserve(2 * seq_elems * sizeof(uint32_t)));
        uint32_t* bufA = c->ws.as<uint32_t>();
        uint32_t* bufB = bufA + seq_elems;
        const uint32_t* in = nullptr;
        uint32_t* out = bufA;
        const uint32_t warp_grid = grid_for(sh.nblk * 32);
        const uint64_t* all_bits = b->d_bits;
        if (b->wt_type == 1) {  // block-compressed levels: expand them once, the passes stream over plain bits
            IDC_TRY(c->scratch.reserve((uint64_t)sh.levels * sh.words * 8));
            IDC_TRY(wt_expand(c, b, c->scratch.as<uint64_t>()));
            all_bits = c->scratch.as<uint64_t>();
        }
        const bool two = getenv("IDC_WT_REPLAY1") == nullptr;  // (experiments: one level per pass)
        for (uint32_t lev = 0; lev < sh.levels;) {
            LaunchScope ls(c, "k_wt_replay");
            const uint64_t* bits = all_bits + (uint64_t)lev * sh.words;
            const uint32_t* rank = b->d_rank + (uint64_t)lev * sh.rank_stride;
            if (two && lev + 1 < sh.levels) {  // two levels per pass
                const uint64_t* bits1 = bits + sh.words;
                const uint32_t* rank1 = rank + sh.rank_stride;
                if (lev == 0)
                    k_wt_replay2<true><<<warp_grid, kThreads, 0, c->stream>>>(nullptr, (uint32_t)sh.n, (uint32_t)sh.nblk, bits, rank, bits1, rank1, out);
                else
                    k_wt_replay2<false><<<warp_grid, kThreads, 0, c->stream>>>(in, (uint32_t)sh.n, (uint32_t)sh.nblk, bits, rank, bits1, rank1, out);
                lev += 2;
            } else {
                if (lev == 0)
                    k_wt_replay<true><<<warp_grid, kThreads, 0, c->stream>>>(nullptr, sh.n, sh.nblk, bits, rank, out);
                else
                    k_wt_replay<false><<<warp_grid, kThreads, 0, c->stream>>>(in, sh.n, sh.nblk, bits, rank, out);
                lev += 1;
            }
            in = out;
            out = out == bufA ? bufB : bufA;
        }
        IDC_TRY(check_last_launch("k_wt_replay"));
        WtEmitArgs a{in, b->d_start, d_sel, d_off, nsel, total, out_dev};
        LaunchScope ls(c, "k_wt_emit");
        if (id_bytes == 8)
            k_wt_emit<int64_t><<<grid_for(((total + 511) >> 9) * 32), kThreads, 0, c->stream>>>(a);
        else
            k_wt_emit<int32_t><<<grid_for(((total + 511) >> 9) * 32), kThreads, 0, c->stream>>>(a);
    } else {
        WtDecArgs a{b->view(), d_sel, d_off, nsel, total, out_dev, d_status};
        LaunchScope ls(c, "k_wt_decode");
        if (id_bytes == 8)
            k_wt_decode<int64_t><<<grid_for(total), kThreads, 0, c->stream>>>(a);
        else
            k_wt_decode<int32_t><<<grid_for(total), kThreads, 0, c->stream>>>(a);
    }
    IDC_TRY(check_last_launch("k_wt_decode"));
    uint32_t st = 0;
    IDC_CUDA(cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, c->stream));
    if (out_mem == IDC_MEM_HOST)
        IDC_CUDA(cudaMemcpyAsync(ids_out, out_dev, total * id_bytes, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    return wt_status_to_error(st, "wt_decode");
}

/* Build a blob from the exported arrays (HOST or DEVICE per `mem`; shapes as idc_wt_blob_export documents them). */
int idc_wt_blob_import(idc_ctx* c, uint64_t nlist, const uint64_t* list_offsets, int wt_type, const uint64_t* bits, const uint32_t* rank,
                       const uint32_t* sel1, const uint32_t* sel0, const uint32_t* start, int mem, idc_wt_blob** out) {
    IDC_REQUIRE(c && out && list_offsets, IDC_ERR_ARG, "idc_wt_blob_import: null argument");
    IDC_REQUIRE(mem == IDC_MEM_HOST || mem == IDC_MEM_DEVICE, IDC_ERR_ARG, "mem must be IDC_MEM_HOST or IDC_MEM_DEVICE");
    IDC_REQUIRE(wt_type == 0 || wt_type == 1, IDC_ERR_ARG, "wt_type must be 0 or 1");
    IDC_REQUIRE(nlist <= (1ull << 31), IDC_ERR_ARG, "too many lists");
    *out = nullptr;
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    std::unique_ptr<idc_wt_blob> b(new idc_wt_blob());
    b->ctx = c;
    b->ref.bind(c);
    b->nlist = nlist;
    b->wt_type = wt_type;
    b->list_offsets.resize(nlist + 1);
    for (uint64_t l = 0; l <= nlist; l++) {
        IDC_REQUIRE(l == 0 || list_offsets[l] >= list_offsets[l - 1], IDC_ERR_ARG, "offsets must be non-decreasing");
        b->list_offsets[l] = list_offsets[l] - list_offsets[0];
    }
    const uint64_t n = b->list_offsets[nlist];
    IDC_REQUIRE(n < (1ull << 32) - 4096, IDC_ERR_ARG, "ntotal must be below 2^32 - 4096 (32-bit positions)");
    b->total_ids = n;
    b->sh = wt_shape(nlist, n);
    if (n == 0) {
        b->sh.levels = 0;
        *out = b.release();
        return IDC_OK;
    }
    IDC_REQUIRE(bits && rank && sel1 && sel0 && start, IDC_ERR_ARG, "idc_wt_blob_import: null array");
    const WtShape sh = b->sh;
    IDC_TRY(dev_alloc(c, &b->d_list_off, nlist + 1, &b->device_bytes));
    IDC_TRY(dev_alloc(c, &b->d_bits, sh.levels * sh.words, &b->device_bytes));
    IDC_TRY(dev_alloc(c, &b->d_rank, sh.levels * sh.rank_stride, &b->device_bytes));
    IDC_TRY(dev_alloc(c, &b->d_sel1, sh.levels * sh.samp_stride, &b->device_bytes));
    IDC_TRY(dev_alloc(c, &b->d_sel0, sh.levels * sh.samp_stride, &b->device_bytes));
    IDC_TRY(dev_alloc(c, &b->d_start, nlist, &b->device_bytes));
    const cudaMemcpyKind kind = mem == IDC_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    cudaStream_t s = c->stream;
    IDC_CUDA(cudaMemcpyAsync(b->d_list_off, b->list_offsets.data(), (nlist + 1) * 8, cudaMemcpyHostToDevice, s));
    IDC_CUDA(cudaMemcpyAsync(b->d_bits, bits, sh.levels * sh.words * 8, kind, s));
    IDC_CUDA(cudaMemcpyAsync(b->d_rank, rank, sh.levels * sh.rank_stride * 4, kind, s));
    IDC_CUDA(cudaMemcpyAsync(b->d_sel1, sel1, sh.levels * sh.samp_stride * 4, kind, s));
    IDC_CUDA(cudaMemcpyAsync(b->d_sel0, sel0, sh.levels * sh.samp_stride * 4, kind, s));
    IDC_CUDA(cudaMemcpyAsync(b->d_start, start, nlist * 4, kind, s));
    IDC_CUDA(cudaStreamSynchronize(s));
    if (wt_type == 1) IDC_TRY(wt_compress(c, b.get()));
    *out = b.release();
    return IDC_OK;
}

/* wt_type = 1 only: the compressed arrays as they lie in HBM (HOST copies; any pointer may be NULL) -- cls[levels * nblk]
 * (eight 6-bit classes + the block's 8 tail bits per 512-bit block), ptr[levels * (nblk + 1)] (bit offset of the
 * block's first offset field in its level's stream), off_base[levels + 1] (first 64-bit word of each level's stream in
 * off; off_base[levels] = words in use), off[off_base[levels]]. */
int idc_wt_blob_export_rrr(const idc_wt_blob* b, uint64_t* cls, uint32_t* ptr, uint64_t* off_base, uint64_t* off) {
    IDC_REQUIRE(b, IDC_ERR_ARG, "null blob");
    IDC_REQUIRE(b->wt_type == 1, IDC_ERR_ARG, "not a wt_type = 1 blob");
    IDC_CUDA(cudaSetDevice(b->ctx->device));
    cudaStream_t s = b->ctx->stream;
    if (b->total_ids) {
        const WtShape& sh = b->sh;
        if (cls) IDC_CUDA(cudaMemcpyAsync(cls, b->d_cls, sh.levels * sh.nblk * 8, cudaMemcpyDeviceToHost, s));
        if (ptr) IDC_CUDA(cudaMemcpyAsync(ptr, b->d_ptr, sh.levels * (sh.nblk + 1) * 4, cudaMemcpyDeviceToHost, s));
        if (off_base) memcpy(off_base, b->off_base.data(), (sh.levels + 1) * 8);
        if (off && b->off_base.back()) IDC_CUDA(cudaMemcpyAsync(off, b->d_off, b->off_base.back() * 8, cudaMemcpyDeviceToHost, s));
    }
    IDC_CUDA(cudaStreamSynchronize(s));
    return IDC_OK;
}

// ---- flat file form (idc_file.h): header words, the list CSR and the five arrays of idc_wt_blob_export
int idc_wt_blob_save(const idc_wt_blob* b, const char* path) {
    IDC_REQUIRE(b && path, IDC_ERR_ARG, "idc_wt_blob_save: null argument");
    const WtShape& sh = b->sh;
    const bool any = b->total_ids != 0;
    std::vector<uint64_t> bits(any ? sh.levels * sh.words : 0);
    std::vector<uint32_t> rank(any ? sh.levels * sh.rank_stride : 0), sel1(any ? sh.levels * sh.samp_stride : 0),
        sel0(any ? sh.levels * sh.samp_stride : 0), start(any ? b->nlist : 0);
    IDC_TRY(idc_wt_blob_export(b, nullptr, bits.data(), rank.data(), sel1.data(), sel0.data(), start.data()));
    std::vector<uint64_t> hdr{b->nlist, (uint64_t)b->wt_type, b->total_ids};
    FileWriter w;
    IDC_TRY(w.open(path, kFileWt, 7));
    w.vec(hdr);
    w.vec(b->list_offsets);
    w.vec(bits);
    w.vec(rank);
    w.vec(sel1);
    w.vec(sel0);
    w.vec(start);
    return w.close(path);
}

int idc_wt_blob_load(idc_ctx* c, const char* path, idc_wt_blob** out) {
    IDC_REQUIRE(c && path && out, IDC_ERR_ARG, "idc_wt_blob_load: null argument");
    *out = nullptr;
    FileReader r;
    IDC_TRY(r.open(path, kFileWt));
    std::vector<uint64_t> hdr, offs, bits;
    std::vector<uint32_t> rank, sel1, sel0, start;
    IDC_TRY(r.vec(hdr));
    IDC_TRY(r.vec(offs));
    IDC_TRY(r.vec(bits));
    IDC_TRY(r.vec(rank));
    IDC_TRY(r.vec(sel1));
    IDC_TRY(r.vec(sel0));
    IDC_TRY(r.vec(start));
    IDC_REQUIRE(hdr.size() == 3 && offs.size() == hdr[0] + 1, IDC_ERR_ARG, "%s: section sizes do not match the header", path);
    const uint64_t n = offs[hdr[0]] - offs[0];
    IDC_REQUIRE(n == hdr[2], IDC_ERR_ARG, "%s: list offsets and the stored id count disagree", path);
    if (n) {
        const WtShape sh = wt_shape(hdr[0], n);
        IDC_REQUIRE(bits.size() == sh.levels * sh.words && rank.size() == sh.levels * sh.rank_stride && sel1.size() == sh.levels * sh.samp_stride &&
                        sel0.size() == sh.levels * sh.samp_stride && start.size() == hdr[0],
                    IDC_ERR_ARG, "%s: array sizes do not follow from the index shape", path);
    }
    return idc_wt_blob_import(c, hdr[0], offs.data(), (int)hdr[1], bits.data(), rank.data(), sel1.data(), sel0.data(), start.data(),
                              IDC_MEM_HOST, out);
}

}  // extern "C"
