// ef_kernels.cu -- Elias-Fano encode / bulk decode / random-access select and
// fixed-width bit packing on sm_100a. These are the HBM-bound kernels of the
// codec: every id is read once and every packed word is written once, with
// coalesced 8-byte stores; there is no contraction, hence no tensor cores.
//
//   k_ef_encode   one warp per tile of 1024 ids: read once, transposed, lane-local packing of both bit vectors
//   k_ef_decode   one warp per tile of 1024 ids: popcount + warp prefix scan over
//                 the upper-bit words, upper parts staged in shared memory, then a
//                 coalesced pass that merges the lower bits and stores the ids
//   k_ef_select   one thread per (list, offset) query via the select samples
//   k_bits_pack / k_bits_unpack   fixed-width ids (packed-bits baselines)
#include <algorithm>
#include <cstring>
#include <memory>
#include <numeric>

#include "ef_core.cuh"
#include "idc_file.h"
#include "idc_host.h"
#include "idc_prep.cuh"

using namespace idc;

namespace {

constexpr uint32_t kEncTileIds = 1024;    // ids per warp in k_ef_encode (a multiple of 64: tiles own whole lower-bits words)
constexpr uint32_t kEncWinWords = 256;    // 32-bit words of the shared-memory window over the upper bits (8192 bits)
constexpr uint32_t kDecChunkWords = 16;   // 64-bit high words per warp in k_ef_decode (1024 bits, <= 1024 ids)
constexpr uint32_t kDecTile = 1024;       // max ids of one chunk
constexpr int kDecThreads = 256;          // 8 warps per CTA in k_ef_decode

// Chunk descriptor, written by the encoder, 32 bytes: everything k_ef_decode needs for one chunk.
struct EfChunk {
    uint64_t a;         // high32:40 | count:11 | l:5 | nwords32:6
                        //   high32 = offset (32-bit words) of the chunk's first upper-bits word
    uint64_t low32;     // offset (32-bit words) of the lower-bits word that holds id number (r0 & ~31)
    uint64_t out_base;  // element offset of the list's first id in a decode-everything output
    uint64_t b;         // r0:32 | zeros:32   r0 = ids of the list before this chunk,
                        //                    zeros = chunk's first bit position - r0
};

}  // namespace

struct idc_ef_blob {
    idc_ctx* ctx = nullptr;
    idc::CtxRef ref;  // declared right after ctx: destroyed last, after the arrays went back to the pool
    uint64_t nlist = 0, total_ids = 0, low_words = 0, high_words = 0, bits_total = 0, nsamples = 0;
    uint32_t row_stride = 0;
    uint32_t max_l = 0;
    // Host mirrors of the per-list tables. The tables are computed on the device (shapes + prefix sums); the mirrors
    // are fetched on first use (export, save, decoding a subset of lists): see ef_host_tables.
    mutable bool host_tables = false;
    mutable std::vector<uint64_t> list_offsets;  // nlist+1, ids per list (CSR)
    mutable std::vector<uint8_t> l;
    mutable std::vector<uint64_t> universe, low_off, high_off, samp_off, dir_off;  // nlist(+1)
    uint32_t* d_universe = nullptr;  // max id of each list
    uint64_t* d_list_off = nullptr;
    uint8_t* d_l = nullptr;
    uint64_t* d_low_off = nullptr;
    uint64_t* d_high_off = nullptr;
    uint64_t* d_samp_off = nullptr;
    uint64_t* d_low = nullptr;
    uint64_t* d_high = nullptr;
    uint32_t* d_samples = nullptr;
    uint64_t* d_dir_off = nullptr;
    EfChunk* d_dir = nullptr;        // one descriptor per chunk of 16 high words (1024 bits)
    uint64_t ndir = 0;
    uint64_t device_bytes = 0;
    // cached decode-everything tile table
    bool plan_ready = false;
    uint32_t* d_tile_list = nullptr;
    uint32_t* d_tile_idx = nullptr;
    uint64_t* d_tile_out = nullptr;
    uint64_t ntiles = 0;
    ~idc_ef_blob() {
        if (ctx) ctx->pool_release(d_list_off);
        if (ctx) ctx->pool_release(d_universe);
        if (ctx) ctx->pool_release(d_l);
        if (ctx) ctx->pool_release(d_low_off);
        if (ctx) ctx->pool_release(d_high_off);
        if (ctx) ctx->pool_release(d_samp_off);
        if (ctx) ctx->pool_release(d_low);
        if (ctx) ctx->pool_release(d_high);
        if (ctx) ctx->pool_release(d_samples);
        if (ctx) ctx->pool_release(d_dir_off);
        if (ctx) ctx->pool_release(d_dir);
        if (ctx) ctx->pool_release(d_tile_list);
        if (ctx) ctx->pool_release(d_tile_idx);
        if (ctx) ctx->pool_release(d_tile_out);
    }
};

namespace {

// Tile descriptor of the encoder, 64 bytes, written by k_ef_tile_desc: everything a warp needs to know about a tile of
// 1024 consecutive ids and about the tile's list, so that the encode kernel does no dependent global loads at all.
struct alignas(16) EfTile {
    uint64_t src;       // element offset of the tile's first id
    uint64_t low32;     // offset (32-bit words, from the start of the blob's lower-bits array) of the tile's first word
    uint64_t high_off;  // offset (64-bit words) of the list's upper-bits vector
    uint64_t out_base;  // element offset of the list's first id in a decode-everything output (chunk descriptors)
    uint64_t dir_off;   // first chunk descriptor of the list
    uint64_t samp;      // slot of the tile's first select sample
    uint32_t m;         // ids of the list
    uint32_t hw;        // 64-bit words of the list's upper-bits vector
    uint32_t idx_cnt;   // tile number inside the list << 10 | (ids in the tile - 1)
    uint32_t l;
};
static_assert(sizeof(EfTile) == 64, "EfTile is copied as four 16-byte pieces");

struct EfEncArgs {
    const void* ids;
    uint64_t* low;
    uint64_t* high;
    uint32_t* samples;
    EfChunk* dir;
    const EfTile* tiles;        // one descriptor per tile of 1024 ids (k_ef_tile_desc)
    uint32_t ntiles;
    uint32_t check_input;       // ascending input taken on trust so far: verify order and width while encoding
    uint32_t* status;
};

// Ascending input: a list's universe (max id) is its last id, no pass over the ids needed for it. One thread per
// list; the order / width of the ids is verified by k_ef_encode, which reads every id anyway.
template <typename IdT>
__global__ void __launch_bounds__(kThreads) k_list_ends(const void* ids, const uint64_t* list_src, const uint32_t* list_n,
                                                        uint32_t nlist, uint32_t* lo, uint32_t* hi, uint32_t* status) {
    uint32_t L = blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= nlist) return;
    const uint32_t n = list_n[L];
    const IdT* src = reinterpret_cast<const IdT*>(ids) + list_src[L];
    const uint64_t first = n ? load_id(src) : 0, last = n ? load_id(src + n - 1) : 0;
    if (sizeof(IdT) == 8 && ((first | last) >> 32)) atomicOr(status, kStWide);
    lo[L] = (uint32_t)first;
    hi[L] = (uint32_t)last;
}

// Chunk descriptor C of the tile's list (a chunk = kDecChunkWords 64-bit words = 1024 bits of the upper-bits vector);
// `before` = ids of the list in front of the chunk's first bit.
__device__ __forceinline__ void ef_emit_chunk(const EfEncArgs& a, const EfTile& t, uint64_t C, uint64_t before) {
    const uint64_t W = C * kDecChunkWords, hw = t.hw;
    if (W >= hw) return;
    const bool last = W + kDecChunkWords >= hw;
    const uint64_t cntc = last ? t.m - before : 0x7ffull;  // 0x7ff: k_ef_finish_chunks takes it from the next descriptor
    const uint64_t rest = hw - W;
    const uint64_t nw32 = 2 * (rest < kDecChunkWords ? rest : kDecChunkWords);
    const uint64_t low_base = t.low32 - (uint64_t)(t.idx_cnt >> 10) * (kEncTileIds / 32) * t.l;  // the list's first lower-bits word
    EfChunk d;
    d.a = (2 * (t.high_off + W)) | (cntc << 40) | ((uint64_t)t.l << 51) | (nw32 << 56);
    d.low32 = low_base + (before >> 5) * t.l;
    d.out_base = t.out_base;
    d.b = before | ((W * 64 - before) << 32);
    a.dir[t.dir_off + C] = d;
}

// ---- bulk asynchronous copies (the TMA unit's 1-D form, cp.async.bulk: ONE instruction moves a whole tile from
// global to shared memory, completion is signalled on an mbarrier; SASS: UBLKCP + SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// src and dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "EF_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra EF_DONE;\n"
        "bra EF_WAIT;\n"
        "EF_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// per-tile descriptors from the per-list tables
struct EfTileDescArgs {
    const uint64_t* list_src;
    const uint64_t* list_off;
    const uint8_t* l;
    const uint64_t* low_off;
    const uint64_t* high_off;
    const uint64_t* samp_off;
    const uint64_t* dir_off;
    const uint64_t* tile_base;
    uint32_t nlist;
    EfTile* tiles;
};

// one thread per tile; its list = the last one whose first tile is not past it (binary search over tile_base)
__global__ void __launch_bounds__(kThreads) k_ef_tile_desc(EfTileDescArgs a, uint32_t ntiles) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ntiles) return;
    uint32_t lo = 0, hi = a.nlist;  // tile_base[lo] <= g < tile_base[hi]
    while (hi - lo > 1u) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(reinterpret_cast<const unsigned long long*>(a.tile_base) + mid) <= g)
            lo = mid;
        else
            hi = mid;
    }
    const uint32_t L = lo, t = g - (uint32_t)a.tile_base[L], l = a.l[L];
    const uint64_t o0 = a.list_off[L], m = a.list_off[L + 1] - o0;
    const uint64_t h0 = a.high_off[L], hw = a.high_off[L + 1] - h0;
    const uint64_t i0 = (uint64_t)t * kEncTileIds;
    const uint32_t cnt = (uint32_t)(m - i0 < kEncTileIds ? m - i0 : kEncTileIds);
    EfTile d;
    d.src = a.list_src[L] + i0;
    d.low32 = 2 * a.low_off[L] + (uint64_t)t * (kEncTileIds / 32) * l;
    d.high_off = h0;
    d.out_base = o0;
    d.dir_off = a.dir_off[L];
    d.samp = a.samp_off[L] + (i0 >> kEfSampleLog);
    d.m = (uint32_t)m;
    d.hw = (uint32_t)hw;  // < 2^26: the vector has < 2^32 bits
    d.idx_cnt = (t << 10) | (cnt - 1u);
    d.l = l;
    a.tiles[g] = d;
}

constexpr int kEncWarps = 4;          // warps per CTA of k_ef_encode
constexpr uint32_t kEncSubIds = 128;  // ids per bulk copy: a tile arrives as 8 copies of 4 lane runs each
constexpr int kEncMaxFastL = 22;      // widest lower-bits field a full tile can have (1024 ids below 2^32)

// shared memory of one warp. `raw` is the bulk-copy target: sub-block c (ids 128 c .. 128 c + 127 of the tile, as they
// lie in global memory from the 16-byte boundary below the tile's first id) sits at 16 + c * kSubStride; the 16 bytes in
// front of sub-block 0 receive the id before the tile (tiles after a list's first one fetch 16 bytes more). The stride
// is the sub-block's size + 16 bytes, so the four 32-id runs of a sub-block go to the four quarter-warps and the eight
// lanes of a quarter-warp (one run from every sub-block) read different banks. `desc` is a ring of tile descriptors,
// filled two tiles ahead by cp.async: the persistent loop carries no prefetched state in registers.
template <typename IdT>
struct EfEncSmem {
    static constexpr uint32_t kSubStride = kEncSubIds * sizeof(IdT) + 16;
    alignas(16) uint8_t raw[16 + 8 * kSubStride];
    alignas(16) EfTile desc[3];
    uint32_t ts[kEncTileIds + 32];  // slow path: the transposed 32-bit tile; fast path: staging of the lower-bits words
    uint32_t win[kEncWinWords];
    alignas(8) uint64_t bar;
};

// Lower bits of one lane's run of 32 consecutive ids with a COMPILE-TIME field width: the 32 fields are exactly L
// 32-bit words; every shift and word index is a constant (mask + shift-add per id, nothing data dependent). The words
// are staged in shared memory (row stride L | 1: conflict-free) and leave as coalesced 128-byte stores.
template <int L>
__device__ __forceinline__ void ef_low_tile(const uint32_t (&v)[32], uint32_t* stage, uint32_t* low32, uint32_t run, uint32_t lane,
                                            uint32_t nlw) {  // nlw = 32-bit words the tile owns (a partial tile: fewer than 32 L)
    if constexpr (L > 0) {
        constexpr uint32_t kMask = (1u << L) - 1u, kRow = (uint32_t)L | 1u;
        uint32_t w[L];
#pragma unroll
        for (int k = 0; k < L; k++) w[k] = 0u;
#pragma unroll
        for (int r = 0; r < 32; r++) {
            const uint32_t f = v[r] & kMask;
            const int k = (r * L) >> 5, s = (r * L) & 31;
            w[k] += f << s;  // disjoint bits: + is | (one shift-add)
            if (s + L > 32) w[k + 1] = f >> (32 - s);
        }
        uint32_t* row = stage + run * kRow;
#pragma unroll
        for (int k = 0; k < L; k++) row[k] = w[k];
        __syncwarp();
        // (the lane number goes through an opaque move: otherwise the compiler hoists the 253 staging indices of all
        // field widths out of the persistent tile loop and keeps them in local memory)
        asm volatile("" : "+r"(lane));
#pragma unroll
        for (int k = 0; k < L; k++) {
            const uint32_t q = 32u * (uint32_t)k + lane;
            if (q < nlw) {
                if constexpr ((L & 1) != 0) {
                    low32[q] = stage[q];
                } else {
                    const uint32_t rr = q / (uint32_t)L;
                    low32[q] = stage[q + rr];
                }
            }
        }
    }
}

// SLOW path of k_ef_encode (partial tiles, upper bits wider than one window, input that is about to be refused): the
// tile has been transposed into `ts` (element e at e + e / 32, ids past the end of the list as 0), lane j owns the run
// of ids 32 j .. 32 j + 31. Out of line: it is rare and must not weigh on the fast path's register allocation.
__device__ __noinline__ void ef_tile_slow(const EfEncArgs& a, uint32_t* ts, uint32_t* win, const EfTile* tp, uint64_t id_prev_tile,
                                           uint32_t bad, uint32_t lane) {
    const EfTile t = *tp;
    const uint64_t m = t.m, hw = t.hw, i0 = (uint64_t)(t.idx_cnt >> 10) * kEncTileIds;
    const uint32_t l = t.l, i0w = (uint32_t)i0, cnt = (t.idx_cnt & 1023u) + 1u;
    const bool last_tile = i0 + cnt == m;
    uint32_t* low32 = reinterpret_cast<uint32_t*>(a.low) + t.low32;
    uint32_t* high32 = reinterpret_cast<uint32_t*>(a.high + t.high_off);
    const int64_t hp_prev_tile = i0 ? (int64_t)((id_prev_tile >> l) + i0 - 1) : -1;  // the one before this tile's first one
    uint32_t v[32];  // lane-major: v[r] = id of element 32 * lane + r
#pragma unroll
    for (int r = 0; r < 32; r++) v[r] = ts[33u * lane + (uint32_t)r];
    if (a.check_input) {
        // ascending? inside the lane, across lanes, across the tile's start (ids past the list's end were loaded as 0)
#pragma unroll
        for (int r = 0; r + 1 < 32; r++)
            if (32u * lane + (uint32_t)r + 1u < cnt && v[r + 1] < v[r]) bad |= kStUnsorted;
        const uint32_t next_first = __shfl_down_sync(0xffffffffu, v[0], 1);
        if (lane < 31u && 32u * (lane + 1u) < cnt && next_first < v[31]) bad |= kStUnsorted;
        if (lane == 0 && i0 && (uint64_t)v[0] < id_prev_tile) bad |= kStUnsorted;
        // The list's shapes were derived from its LAST id. If this tile is ascending its largest position is its
        // last one; should that lie outside the list's bit vector (only possible when the list as a whole is not
        // ascending), or the tile itself be out of order, nothing of it is written: the call fails anyway.
        const uint64_t last_pos = (uint64_t)(ts[(cnt - 1u) + ((cnt - 1u) >> 5)] >> l) + i0 + cnt - 1u;
        if (last_pos >= hw * 64) bad |= kStUnsorted;
        if (bad) atomicOr(a.status, bad);
        if (__any_sync(0xffffffffu, (bad & kStUnsorted) != 0)) return;
    } else if (bad) {
        atomicOr(a.status, bad);
    }
    const uint32_t hpF = (ts[0] >> l) + i0w;
    const uint32_t hpL = (ts[(cnt - 1u) + ((cnt - 1u) >> 5)] >> l) + i0w + cnt - 1u;
    // ---- chunk descriptors: id e announces the chunks C with prev < 1024 C <= hp(e), prev = the one before it: it
    // is the first id at or past their first bit, so `ids before the chunk` = e. A straight-line pass marks the
    // announcing ids (a few per tile); the descriptors are written in a rolled loop that re-reads those ids.
    {
        const uint32_t my_last = (v[31] >> l) + i0w + 32u * lane + 31u;
        int64_t prev = (int64_t)__shfl_up_sync(0xffffffffu, my_last, 1);
        if (lane == 0) prev = hp_prev_tile;
        uint32_t c_lo = prev < 0 ? 0u : (uint32_t)(prev >> 10) + 1u;  // first chunk not announced yet
        uint32_t bm = 0;
#pragma unroll
        for (int r = 0; r < 32; r++) {
            const uint32_t e = 32u * lane + (uint32_t)r;
            const uint32_t c_hi = ((v[r] >> l) + i0w + e) >> 10;
            if (e < cnt) {
                bm |= c_hi >= c_lo ? 1u << r : 0u;
                c_lo = c_hi + 1u;
            }
        }
        while (bm) {
            const uint32_t r = (uint32_t)__ffs((int)bm) - 1u;
            bm &= bm - 1u;
            const uint32_t e = 32u * lane + r;
            const uint32_t c_hi = ((ts[33u * lane + r] >> l) + i0w + e) >> 10;
            const int64_t pp = r ? (int64_t)((ts[33u * lane + r - 1u] >> l) + i0w + e - 1u) : prev;
            for (uint32_t C = pp < 0 ? 0u : (uint32_t)(pp >> 10) + 1u; C <= c_hi; C++) ef_emit_chunk(a, t, C, i0 + e);
        }
        if (last_tile) {
            const uint64_t nchunks = (hw + kDecChunkWords - 1) / kDecChunkWords;
            for (uint64_t C = (uint64_t)(hpL >> 10) + 1 + lane; C < nchunks; C += 32) ef_emit_chunk(a, t, C, m);
        }
    }
    __syncwarp();
    // ---- lower bits
    if (l) {
        const uint32_t nlw = ((cnt * l + 63u) / 64u) * 2u;  // 32-bit words, a whole number of 64-bit words
        const uint32_t fmask = (1u << l) - 1u;             // l <= 31 for ids < 2^32
        uint64_t acc = 0;
        uint32_t fill = 0, wq = lane * l;
#pragma unroll
        for (int r = 0; r < 32; r++) {
            acc |= (uint64_t)(v[r] & fmask) << fill;  // ids past the end of the list were loaded as 0
            fill += l;
            if (fill >= 32u) {
                ts[wq++] = (uint32_t)acc;
                acc >>= 32;
                fill -= 32u;
            }
        }
        __syncwarp();
        for (uint32_t q = lane; q < nlw; q += 32) low32[q] = ts[q];
    }
    // ---- upper bits
    const uint32_t gF = hpF >> 5, gL = hpL >> 5;  // the tile's first / last 32-bit word of the high vector
    if (lane % 8u == 0u && 32u * lane < cnt) a.samples[t.samp + (lane >> 3)] = (v[0] >> l) + i0w + 32u * lane;
    uint32_t wb = hpF & ~31u;  // window base (a bit position)
    for (;;) {
        for (uint32_t q = lane; q < kEncWinWords; q += 32) win[q] = 0u;
        __syncwarp();
        const bool more = hpL - wb >= 32u * kEncWinWords;  // warp-uniform: some ones lie past this window
        uint32_t next = 0xffffffffu;
#pragma unroll
        for (int r = 0; r < 32; r++) {
            const uint32_t e = 32u * lane + (uint32_t)r;
            const uint32_t hp = (v[r] >> l) + i0w + e;
            const uint32_t rel = hp - wb;  // wraps (huge) for positions below the window: handled in an earlier pass
            if (e < cnt && rel < 32u * kEncWinWords) atomicOr(win + (rel >> 5), 1u << (hp & 31u));
            if (more && e < cnt && hp >= wb && rel >= 32u * kEncWinWords && hp < next) next = hp;
        }
        __syncwarp();
        for (uint32_t q = lane; q < kEncWinWords; q += 32) {
            const uint32_t g = (wb >> 5) + q;
            if (g < gF || g > gL) continue;
            const uint32_t val = win[q];
            if (g == gF || g == gL) {
                if (val) atomicOr(high32 + g, val);
            } else {
                high32[g] = val;
            }
        }
        if (!more) break;
        next = __reduce_min_sync(0xffffffffu, next);
        wb = next & ~31u;
        __syncwarp();
    }
}

// PERSISTENT warps, one tile of 1024 consecutive ids of a list at a time. A tile is fetched by bulk asynchronous
// copies (cp.async.bulk = the TMA unit's 1-D form, completion on the warp's mbarrier) issued a whole tile ahead: while
// a warp packs tile k, the copy engine lands tile k + 1 in its raw buffer -- the global-memory latency is off the
// warp's critical path and no register tile of in-flight loads is needed. Every id is read exactly once. The tile
// descriptors (64 bytes, everything about the tile and its list) arrive two tiles ahead in a shared-memory ring.
// Lane j owns a run of 32 CONSECUTIVE ids; everything after the fetch is lane-local:
//   lower bits: a run's 32 fields are exactly l consecutive 32-bit words (1024 l bits per tile = a whole number of
//               64-bit words, so tiles own their lower-bits words), staged in shared memory, stored coalesced.
//   upper bits: a run's ones are strictly increasing and ~96 bits apart from the next run's, so setting them in an
//               8192-bit shared-memory window is an atomicOr without conflicts. Words strictly between the tile's
//               first and last one belong to the tile alone (plain stores); its first and last word may be shared
//               with the neighbouring tiles (atomicOr on the pre-zeroed array).
//   chunk descriptors for the decoder: the first id at or past a chunk's first bit announces it (ids before the
//               chunk = its number); the list's last tile adds the trailing ones.
// FAST path (full tiles whose upper bits fit one window and whose runs cross at most one chunk boundary each -- all
// but the last tile of a list): the run is read straight from the raw buffer into registers (lane 8 i + c owns run
// 4 c + i, see EfEncSmem), the lower bits are packed with the field width as a template parameter, the upper bits
// need one window pass without range checks, and a run that crosses a chunk boundary finds the announcing id by a
// binary search over its part of the raw buffer.
// SLOW path (everything else): ef_tile_slow, through a transposed 32-bit tile.
template <typename IdT>
__global__ void __launch_bounds__(kEncWarps * 32, 4) k_ef_encode(const __grid_constant__ EfEncArgs a) {
    extern __shared__ __align__(128) uint8_t ef_enc_smem[];
    using Smem = EfEncSmem<IdT>;
    Smem* sm = reinterpret_cast<Smem*>(ef_enc_smem) + (threadIdx.x >> 5);
    uint32_t lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * kEncWarps;
    uint32_t tile = blockIdx.x * kEncWarps + (threadIdx.x >> 5);
    uint32_t* ts = sm->ts;
    uint32_t* win = sm->win;
    if (tile >= a.ntiles) return;
    if (lane == 0) mbar_init(&sm->bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    // descriptor of tile t -> ring slot (asynchronous: lanes 0..3 move 16 bytes each)
    auto fetch_desc = [&](uint32_t t, uint32_t slot) {
        if (lane < 4u && t < a.ntiles)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(reinterpret_cast<uint8_t*>(sm->desc + slot) + 16u * lane)),
                         "l"(reinterpret_cast<const uint8_t*>(a.tiles + t) + 16u * lane)
                         : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // fetch of a tile's ids: sub-block c starts at the 16-byte boundary at or below its first id, sub-block 0 of a
    // tile that is not its list's first one 16 bytes earlier (the id before the tile)
    auto fetch = [&](const EfTile& d) {
        const uintptr_t p = reinterpret_cast<uintptr_t>(reinterpret_cast<const IdT*>(a.ids) + d.src);
        const uint32_t skew = (uint32_t)(p & 15u), cnt = (d.idx_cnt & 1023u) + 1u;
        const uint32_t first = lane * kEncSubIds;
        const uint32_t nsub = first < cnt ? (cnt - first < kEncSubIds ? cnt - first : kEncSubIds) : 0u;  // lane c: ids of sub-block c
        const uint32_t lead = (lane == 0u && (d.idx_cnt >> 10)) ? 16u : 0u;
        const uint32_t bytes = nsub ? lead + ((skew + nsub * (uint32_t)sizeof(IdT) + 15u) & ~15u) : 0u;
        const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
        if (lane == 0) mbar_expect_tx(&sm->bar, total);
        __syncwarp();
        if (bytes)
            bulk_g2s(sm->raw + 16u - lead + lane * Smem::kSubStride,
                     reinterpret_cast<const void*>(p - skew - lead + (size_t)first * sizeof(IdT)), bytes, &sm->bar);
    };
    // prologue: descriptors of the first two tiles, ids of the first
    fetch_desc(tile, 0);
    fetch_desc(tile + nwarps, 1);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    fetch(sm->desc[0]);
    const uint32_t run = ((lane & 7u) << 2) | (lane >> 3);  // fast path: the run of this lane
    const uint32_t lane_next = (((run + 1u) & 3u) << 3) | (((run + 1u) & 31u) >> 2), lane_prev = (((run - 1u) & 3u) << 3) | (((run - 1u) & 31u) >> 2);
    uint32_t parity = 0, slot = 0;
    for (; tile < a.ntiles; tile += nwarps) {
        const EfTile* tp = sm->desc + slot;
        const uint32_t slot1 = slot == 2u ? 0u : slot + 1u, slot2 = slot1 == 2u ? 0u : slot1 + 1u;
        const uint32_t l = tp->l, idx_cnt = tp->idx_cnt;
        const uint64_t i0 = (uint64_t)(idx_cnt >> 10) * kEncTileIds;
        const uint32_t cnt = (idx_cnt & 1023u) + 1u;
        const uint32_t i0w = (uint32_t)i0;  // positions in the high bit vector fit 32 bits (ef_build rejects longer vectors)
        const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(reinterpret_cast<const IdT*>(a.ids) + tp->src) & 15u);
        const uint8_t* raw0 = sm->raw + 16u + skew;  // the tile's first id
        // ---- the tile has landed
        mbar_wait(&sm->bar, parity);
        parity ^= 1u;
        const uint64_t id_prev_tile = i0 ? (sizeof(IdT) == 8 ? (uint64_t) * (reinterpret_cast<const IdT*>(raw0) - 1)
                                                             : (uint64_t)(uint32_t) * (reinterpret_cast<const IdT*>(raw0) - 1))
                                         : 0ull;
        uint32_t v[32];
        uint32_t hpF = 0, hpL = 0, hpL_pad = 0, hp_first = 0, ann_c = 0, ann_before = 0xffffffffu;
        const uint32_t hb = i0w + 32u * run;  // fast path: hp(r) = (v[r] >> l) + hb + r
        bool fast = l <= (uint32_t)kEncMaxFastL;
        if (fast) {
            uint32_t wide = 0;
            // the tile's last id; a PARTIAL tile (always the last one of its list) is filled up with ids just past it
            // whose lower bits are zero: they sort behind, pack to zero bits, and their ones land behind the tile's
            // last one, where the write-out stops
            uint64_t id_last;
            {
                const uint32_t e = cnt - 1u;
                const IdT* lp = reinterpret_cast<const IdT*>(raw0 + (e >> 7) * Smem::kSubStride) + (e & (kEncSubIds - 1u));
                id_last = sizeof(IdT) == 8 ? (uint64_t)*lp : (uint64_t)(uint32_t)*lp;
            }
            bool pad_overflow = false;
            if (cnt < kEncTileIds) {
                const uint64_t pad = ((id_last >> l) + 1ull) << l;
                pad_overflow = (pad >> 32) != 0ull;  // (only a list that ends at the very top of the 32-bit range)
                __syncwarp();  // every lane has read id_last
                for (uint32_t e = cnt + lane; e < kEncTileIds; e += 32u) {
                    IdT* pp = reinterpret_cast<IdT*>(const_cast<uint8_t*>(raw0) + (e >> 7) * Smem::kSubStride) + (e & (kEncSubIds - 1u));
                    *pp = (IdT)pad;
                }
                __syncwarp();
            }
            const IdT* rp = reinterpret_cast<const IdT*>(raw0 + (lane & 7u) * Smem::kSubStride) + (lane >> 3) * 32u;
            if (sizeof(IdT) == 8) {
                // 16-byte loads, two ids each: the eight lanes of a quarter-warp read different bank groups (the shared-
                // memory pipe is this kernel's busiest unit; 8-byte loads cost twice the wavefronts). A run starts 0 or
                // 8 bytes past a 16-byte boundary: in the second case 17 aligned loads cover it.
                const uint4* q4 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(rp) - skew);
                if (skew == 0u) {
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        const uint4 x = q4[k];
                        v[2 * k] = x.x;
                        v[2 * k + 1] = x.z;
                        wide |= x.y | x.w;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k <= 16; k++) {
                        const uint4 x = q4[k];
                        if (k > 0) {
                            v[2 * k - 1] = x.x;
                            wide |= x.y;
                        }
                        if (k < 16) {
                            v[2 * k] = x.z;
                            wide |= x.w;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < 32; r++) v[r] = (uint32_t)rp[r];
            }
            hp_first = (v[0] >> l) + hb;
            const uint32_t hp_last = (v[31] >> l) + hb + 31u;
            hpF = __shfl_sync(0xffffffffu, hp_first, 0);
            hpL_pad = __shfl_sync(0xffffffffu, hp_last, 31);                 // last position the window pass touches
            hpL = ((uint32_t)id_last >> l) + i0w + cnt - 1u;                 // the tile's last one
            const uint32_t next_first = __shfl_sync(0xffffffffu, v[0], lane_next);
            const uint32_t prev_last = __shfl_sync(0xffffffffu, hp_last, lane_prev);
            bool no = wide != 0u || pad_overflow;  // anything the slow path has to deal with (or refuse)
            if (a.check_input) {
                // ascending? inside the run, across runs, across the tile's start
#pragma unroll
                for (int r = 0; r + 1 < 32; r++) no |= v[r + 1] < v[r];
                no |= run < 31u && next_first < v[31];
                no |= run == 0u && i0 && (uint64_t)v[0] < id_prev_tile;
                // The list's shapes were derived from its LAST id: an ascending tile's largest position is its last
                // one; should that lie outside the list's bit vector the list as a whole is not ascending.
                no |= (uint64_t)hpL >= (uint64_t)tp->hw * 64;
            }
            // chunks (1024 bits of the high vector) this run announces: prev < 1024 C <= hp(last id of the run)
            const int64_t prev = run ? (int64_t)prev_last : (i0 ? (int64_t)((id_prev_tile >> l) + i0 - 1) : -1);
            const uint32_t c_lo = prev < 0 ? 0u : (uint32_t)(prev >> 10) + 1u, c_hi = hp_last >> 10;
            no |= c_hi > c_lo;                                  // two boundaries inside one run
            no |= hpL_pad - (hpF & ~31u) >= 32u * kEncWinWords; // upper bits wider than the window
            fast = !__any_sync(0xffffffffu, no);
            if (fast && c_hi == c_lo) {
                // ids of the run before the boundary: hp(r) < 1024 C  <=>  (id(r) >> l) + r < T; the keys ascend and
                // the last one is not below T: binary search over the run in the raw buffer
                const uint32_t T = (c_lo << 10) - hb;
                uint32_t before = 0;
#pragma unroll
                for (uint32_t step = 16; step; step >>= 1) {
                    const uint32_t r = before + step - 1u;
                    if (((uint32_t)rp[r] >> l) + r < T) before += step;
                }
                ann_c = c_lo;
                if (32u * run + before < cnt) ann_before = 32u * run + before;  // (a boundary behind the tile's last id: the trailing loop's)
            }
        }
        uint32_t bad = 0;
        if (!fast) {
            // ---- raw rows -> padded 32-bit tile (element e at e + e / 32)
#pragma unroll 4
            for (int r = 0; r < 32; r++) {
                const uint32_t e = (uint32_t)r * 32u + lane;
                const IdT* ep = reinterpret_cast<const IdT*>(raw0 + (e >> 7) * Smem::kSubStride) + (e & (kEncSubIds - 1u));
                uint64_t id = 0;
                if (e < cnt) id = sizeof(IdT) == 8 ? (uint64_t)*ep : (uint64_t)(uint32_t)*ep;
                if (sizeof(IdT) == 8 && (id >> 32)) bad |= kStWide;
                ts[33u * (uint32_t)r + lane] = (uint32_t)id;
            }
        }
        // every lane is done with the raw buffer: the next tile (its descriptor has arrived) may land in it, and the
        // descriptor after that sets out
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // (a partial tile's fill-up stores before the copy engine's writes)
        __syncwarp();
        if (tile + nwarps < a.ntiles) fetch(sm->desc[slot1]);
        fetch_desc(tile + 2u * nwarps, slot2);
        if (fast) {
            uint32_t* low32 = reinterpret_cast<uint32_t*>(a.low) + tp->low32;
            uint32_t* high32 = reinterpret_cast<uint32_t*>(a.high + tp->high_off);
            // ---- upper bits: one window, based at the word of the tile's first one
            // (nwin words are written out: up to the tile's last one; the window pass of a filled-up partial tile
            // touches nwin_pad words)
            const uint32_t wb = hpF & ~31u, nwin = ((hpL - wb) >> 5) + 1u, nwin_pad = ((hpL_pad - wb) >> 5) + 1u;
            for (uint32_t q = lane; q < nwin_pad; q += 32) win[q] = 0u;
            if ((run & 7u) == 0u && 32u * run < cnt) a.samples[tp->samp + (run >> 3)] = hp_first;
            __syncwarp();
            {
                const uint32_t hbw = hb - wb;
#pragma unroll
                for (int r = 0; r < 32; r++) {
                    const uint32_t rel = (v[r] >> l) + hbw + (uint32_t)r;
                    atomicOr(win + (rel >> 5), 1u << (rel & 31u));
                }
            }
            // ---- lower bits (the staging area is the slow path's tile)
            const uint32_t nlw = ((cnt * l + 63u) / 64u) * 2u;  // 32-bit words, a whole number of 64-bit words
            switch (l) {
#define IDC_EF_LOW_CASE(LL)                       \
    case LL:                                      \
        ef_low_tile<LL>(v, ts, low32, run, lane, nlw); \
        break;
                IDC_EF_LOW_CASE(1) IDC_EF_LOW_CASE(2) IDC_EF_LOW_CASE(3) IDC_EF_LOW_CASE(4) IDC_EF_LOW_CASE(5) IDC_EF_LOW_CASE(6)
                IDC_EF_LOW_CASE(7) IDC_EF_LOW_CASE(8) IDC_EF_LOW_CASE(9) IDC_EF_LOW_CASE(10) IDC_EF_LOW_CASE(11) IDC_EF_LOW_CASE(12)
                IDC_EF_LOW_CASE(13) IDC_EF_LOW_CASE(14) IDC_EF_LOW_CASE(15) IDC_EF_LOW_CASE(16) IDC_EF_LOW_CASE(17) IDC_EF_LOW_CASE(18)
                IDC_EF_LOW_CASE(19) IDC_EF_LOW_CASE(20) IDC_EF_LOW_CASE(21) IDC_EF_LOW_CASE(22)
#undef IDC_EF_LOW_CASE
                default:
                    break;
            }
            __syncwarp();
            for (uint32_t q = lane; q < nwin; q += 32) {
                uint32_t val = win[q];
                if (q == nwin - 1u) val &= 0xffffffffu >> (31u - (hpL & 31u));  // nothing behind the tile's last one
                if (q == 0u || q == nwin - 1u) {
                    if (val) atomicOr(high32 + (wb >> 5) + q, val);
                } else {
                    high32[(wb >> 5) + q] = val;
                }
            }
            // ---- chunk descriptors
            if (ann_before != 0xffffffffu) ef_emit_chunk(a, *tp, ann_c, i0 + ann_before);
            if (i0 + cnt == tp->m) {
                const uint64_t nchunks = ((uint64_t)tp->hw + kDecChunkWords - 1) / kDecChunkWords;
                for (uint64_t C = (uint64_t)(hpL >> 10) + 1 + lane; C < nchunks; C += 32) ef_emit_chunk(a, *tp, C, tp->m);
            }
        } else {
            ef_tile_slow(a, ts, win, tp, id_prev_tile, bad, lane);
        }
        __syncwarp();  // ts / win / the descriptor slot are rewritten by the following tiles
        slot = slot1;
    }
}

__global__ void __launch_bounds__(kThreads) k_ef_finish_chunks(EfChunk* dir, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t a = dir[i].a;
    if (((a >> 40) & 0x7ffull) == 0x7ffull) {
        uint64_t cnt = (dir[i + 1].b & 0xffffffffull) - (dir[i].b & 0xffffffffull);
        dir[i].a = (a & ~(0x7ffull << 40)) | (cnt << 40);
    }
}

struct EfDecArgs {
    const EfChunk* dir;
    const uint64_t* low;
    const uint64_t* high;
    const uint32_t* sel_desc;   // subset decode: descriptor index per tile (null: tile i = descriptor i)
    const uint64_t* sel_out;    // subset decode: element offset in out of the LIST's first id
    const int32_t* row_nos;     // row mode: descriptor = row_nos[slot] (null: slot), out = slot * row_stride
    void* out;
    uint32_t* counts;           // row mode: ids in the row
    uint64_t ntiles;
    uint32_t row_stride;        // != 0: row mode
    uint32_t low_stage_words;   // shared-memory words per warp for staging lower-bits words
    uint64_t nrows;             // row mode: rows of the blob (row numbers on the device are checked against it)
    uint32_t* status;           // row mode: kStRange is OR-ed in when a row number is out of range
};

constexpr uint32_t kDecStageCap = 384;  // staged lower-bits words per warp, at most
constexpr uint32_t kDecHighStage = 36;  // staged upper-bits words per warp: 32 + the 16-byte alignment slack on both sides

// Output pass of k_ef_decode with a COMPILE-TIME field width: groups of 32 consecutive id numbers; the L-bit lower
// fields of a group are exactly L consecutive 32-bit words of the staged stream, lane j's field starts at bit j L of
// them (two LDS + one funnel shift, all offsets immediates); upper part from shared memory, one coalesced store.
template <int L, typename OutT>
__device__ __forceinline__ void ef_dec_out(const uint32_t* s_low, const uint16_t* hi_part, OutT* out, uint32_t lane, uint32_t lead,
                                           uint32_t count, uint32_t ngroups, uint32_t zeros) {
    constexpr uint32_t kMask = L ? (0xffffffffu >> (32 - (L ? L : 1))) : 0u;
    const uint32_t fs = (lane * (uint32_t)L) & 31u;
    const uint32_t* sp = s_low + ((lane * (uint32_t)L) >> 5);
    uint32_t j = lane - lead;  // id number inside the chunk (wraps for the lanes in front of the chunk's first id)
    const uint16_t* hp = hi_part + (int32_t)j;  // (hi_part has 32 entries of slack in front and reads past the chunk's
    OutT* o = out + (int32_t)j;                 //  ids stay inside the CTA's shared memory: only the store is predicated)
    auto one = [&](uint32_t u) {
        uint32_t f = 0;
        if constexpr (L > 0) f = __funnelshift_r(sp[u * L], sp[u * L + 1], fs) & kMask;
        const uint32_t id = (((uint32_t)hp[32u * u] + zeros) << L) | f;  // ids < 2^32 on the device path
        if (j + 32u * u < count) o[32u * u] = (OutT)id;
    };
    // four groups per round, the last round's surplus groups are predicated off like the lanes past the chunk's end
    for (uint32_t g = 0; g < ngroups; g += 4u) {
        one(0);
        one(1);
        one(2);
        one(3);
        j += 128u;
        hp += 128;
        o += 128;
        sp += 4 * L;
    }
}

// ---- decoding one chunk of 16 64-bit words (= 32 32-bit words, one per lane) of the upper-bits vector
struct EfDecChunk {  // the 32-byte descriptor, unpacked
    uint64_t high32, low32o, out_base;
    uint32_t count, l, nw32, r0, zeros;
};

__device__ __forceinline__ EfDecChunk ef_dec_unpack(uint4 d0, uint4 d1) {
    EfDecChunk d;
    const uint64_t da = (uint64_t)d0.x | ((uint64_t)d0.y << 32);
    d.high32 = da & ((1ull << 40) - 1);
    d.count = (uint32_t)(da >> 40) & 0x7ffu;
    d.l = (uint32_t)(da >> 51) & 31u;
    d.nw32 = (uint32_t)(da >> 56) & 63u;
    d.low32o = (uint64_t)d0.z | ((uint64_t)d0.w << 32);
    d.out_base = (uint64_t)d1.x | ((uint64_t)d1.y << 32);
    d.r0 = d1.z;
    d.zeros = d1.w;
    return d;
}

// Staging of a chunk (one lane): its upper-bits words and the lower-bits words of its ids arrive in shared memory by two
// bulk asynchronous copies (cp.async.bulk, the TMA unit's 1-D form), completion on the mbarrier. Both sources are only
// 4- / 8-byte aligned: the copies start at the 16-byte boundary below them.
__device__ __forceinline__ void ef_dec_stage(const EfDecArgs& a, const EfDecChunk& d, uint32_t* s_low, uint32_t* s_high, uint64_t* bar) {
    const uint32_t lead = d.r0 & 31u, nlow = ((lead + d.count + 31u) >> 5) * d.l;
    const uint8_t* lsrc = reinterpret_cast<const uint8_t*>(reinterpret_cast<const uint32_t*>(a.low) + d.low32o);
    const uint8_t* hsrc = reinterpret_cast<const uint8_t*>(reinterpret_cast<const uint32_t*>(a.high) + d.high32);
    const uint32_t lskew = (uint32_t)(reinterpret_cast<uintptr_t>(lsrc) & 15u), hskew = (uint32_t)(reinterpret_cast<uintptr_t>(hsrc) & 15u);
    const uint32_t hbytes = d.count ? (hskew + 4u * d.nw32 + 15u) & ~15u : 0u;
    const uint32_t lbytes = (d.count && nlow && nlow <= a.low_stage_words) ? (lskew + 4u * nlow + 15u) & ~15u : 0u;
    mbar_expect_tx(bar, hbytes + lbytes);  // (an empty chunk: the arrival alone completes the phase)
    if (hbytes) bulk_g2s(s_high, hsrc - hskew, hbytes, bar);
    if (lbytes) bulk_g2s(s_low, lsrc - lskew, lbytes, bar);
}

// The staged chunk -> ids:
//   popcount + warp scan over the upper-bits words -> where each word's ids land;
//   every lane peels the set bits of its word (highest first: FLO, clear, store; four per round) into shared memory
//   as upper parts (position - id number);
//   a coalesced pass over the chunk's ids with the field width as a template parameter (ef_dec_out); ids leave as
//   full 256-byte warp stores.
template <typename OutT>
__device__ __forceinline__ void ef_dec_chunk(const EfDecArgs& a, const EfDecChunk& d, const uint32_t* s_low, const uint32_t* s_high,
                                             uint16_t* hi_part, OutT* out, uint32_t lane) {
    const uint32_t count = d.count, l = d.l, zeros = d.zeros;
    const uint32_t lead = d.r0 & 31u;
    const uint32_t ngroups = (lead + count + 31u) >> 5;   // groups of 32 consecutive id numbers
    const uint32_t nlow = ngroups * l;                     // 32-bit lower-bits words covering them
    const uint32_t* lsrc = reinterpret_cast<const uint32_t*>(a.low) + d.low32o;
    const uint32_t lskew = (uint32_t)(reinterpret_cast<uintptr_t>(lsrc) & 15u);
    const uint32_t hskew = (uint32_t)(reinterpret_cast<uintptr_t>(reinterpret_cast<const uint32_t*>(a.high) + d.high32) & 15u);
    uint32_t w = lane < d.nw32 ? s_high[(hskew >> 2) + lane] : 0u;
    uint32_t sc = (uint32_t)__popc(w);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, sc, o);
        if ((int)lane >= o) sc += v;
    }
    // The id with chunk-relative number j whose one sits at chunk bit P has upper part
    //   (chunk_bit0 + P) - (r0 + j) = zeros + (P - j);  shared memory gets P - j, `zeros` is added on the way out.
    {
        uint16_t* p = hi_part + sc;   // one past this word's last id
        uint32_t v = 32u * lane - sc; // (bit base) - (id number): grows by one for every step back
        while (w) {
            // four ones per round, straight-line: a lane that runs out of ones keeps w = 0 and skips the stores
#pragma unroll
            for (int u = 1; u <= 4; u++) {
                const uint32_t b = 31u - (uint32_t)__clz((int)w);
                const bool on = w != 0u;
                w &= ~(1u << (b & 31u));
                if (on) p[-u] = (uint16_t)(v + (uint32_t)u + b);
            }
            p -= 4;
            v += 4u;
        }
    }
    __syncwarp();
    if (nlow <= a.low_stage_words) {
        const uint32_t* sl = s_low + (lskew >> 2);
        switch (l) {
#define IDC_EF_DEC_CASE(LL)                                                        \
    case LL:                                                                       \
        ef_dec_out<LL, OutT>(sl, hi_part, out, lane, lead, count, ngroups, zeros); \
        break;
            IDC_EF_DEC_CASE(0) IDC_EF_DEC_CASE(1) IDC_EF_DEC_CASE(2) IDC_EF_DEC_CASE(3) IDC_EF_DEC_CASE(4) IDC_EF_DEC_CASE(5)
            IDC_EF_DEC_CASE(6) IDC_EF_DEC_CASE(7) IDC_EF_DEC_CASE(8) IDC_EF_DEC_CASE(9) IDC_EF_DEC_CASE(10) IDC_EF_DEC_CASE(11)
            IDC_EF_DEC_CASE(12) IDC_EF_DEC_CASE(13) IDC_EF_DEC_CASE(14) IDC_EF_DEC_CASE(15) IDC_EF_DEC_CASE(16) IDC_EF_DEC_CASE(17)
            IDC_EF_DEC_CASE(18) IDC_EF_DEC_CASE(19) IDC_EF_DEC_CASE(20) IDC_EF_DEC_CASE(21) IDC_EF_DEC_CASE(22) IDC_EF_DEC_CASE(23)
            IDC_EF_DEC_CASE(24) IDC_EF_DEC_CASE(25) IDC_EF_DEC_CASE(26) IDC_EF_DEC_CASE(27) IDC_EF_DEC_CASE(28) IDC_EF_DEC_CASE(29)
            IDC_EF_DEC_CASE(30) IDC_EF_DEC_CASE(31)
#undef IDC_EF_DEC_CASE
        }
    } else {
        // more lower-bits words than the staging area holds (wide fields): straight from global memory
        const uint32_t fw = (lane * l) >> 5, fs = (lane * l) & 31u, fmask = l ? (0xffffffffu >> (32u - l)) : 0u;
        const int32_t end = (int32_t)count;
        int32_t j = (int32_t)lane - (int32_t)lead;
        const uint32_t* lp = lsrc + lane;
        for (uint32_t g = 0; g < ngroups; g++, j += 32, lp += l) {
            uint32_t lwv = lane < l ? __ldg(lp) : 0u;
            uint32_t x0 = __shfl_sync(0xffffffffu, lwv, fw);
            uint32_t x1 = __shfl_sync(0xffffffffu, lwv, (fw + 1) & 31);
            uint32_t f = __funnelshift_r(x0, x1, fs) & fmask;
            if (j >= 0 && j < end) {
                uint64_t id = ((uint64_t)((uint32_t)hi_part[j] + zeros) << l) | f;
                out[j] = (OutT)id;
            }
        }
    }
}

// shared memory of one warp of the decoders: 1024 upper parts as u16 (position - id number inside the chunk <= 1023;
// 32 entries of slack in front), then per stage: the lower-bits words (+ 4: alignment slack) and the upper-bits words,
// then one mbarrier per stage
__host__ __device__ inline uint32_t ef_dec_stage_bytes(uint32_t low_stage_words) { return 4u * (low_stage_words + 4u) + 4u * kDecHighStage; }
__host__ __device__ inline uint32_t ef_dec_warp_bytes(uint32_t low_stage_words, uint32_t stages) {
    return 64u + 2u * kDecTile + stages * ef_dec_stage_bytes(low_stage_words) + 16u;
}

// One warp per chunk; subsets of lists (sel_desc) and graph rows (row mode), one chunk staged per warp:
//   1. one 32-byte descriptor (written by the encoder) tells the warp everything: where the chunk's words
//      are, the id number it starts at, how many ids it holds, l, where they go;
//   2. ef_dec_stage, 3. ef_dec_chunk.
template <typename OutT>
__global__ void __launch_bounds__(kDecThreads, 6) k_ef_decode(EfDecArgs a) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (kDecThreads / 32) + wib;
    if (warp >= a.ntiles) return;
    uint8_t* s_warp = s_dyn + (size_t)wib * ef_dec_warp_bytes(a.low_stage_words, 1);
    uint16_t* hi_part = reinterpret_cast<uint16_t*>(s_warp + 64u);
    uint32_t* s_low = reinterpret_cast<uint32_t*>(s_warp + 64u + 2u * kDecTile);
    uint32_t* s_high = s_low + a.low_stage_words + 4u;
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_high + kDecHighStage);
    uint64_t di = warp;
    if (a.row_stride) {
        if (a.row_nos) {
            // row numbers may live on the device (e.g. -1 padded neighbour arrays): checked here, like k_roc_decode
            const int64_t r = (int64_t)a.row_nos[warp];
            if (r < 0 || r >= (int64_t)a.nrows) {
                OutT* o = reinterpret_cast<OutT*>(a.out) + warp * a.row_stride;
                for (uint32_t t = lane; t < a.row_stride; t += 32) o[t] = (OutT)-1;
                if (lane == 0) {
                    if (a.counts) a.counts[warp] = 0u;
                    atomicOr(a.status, kStRange);
                }
                return;
            }
            di = (uint64_t)r;
        }
    } else if (a.sel_desc) {
        di = a.sel_desc[warp];
    }
    if (lane == 0) mbar_init(bar, 1);
    const uint4* dp = reinterpret_cast<const uint4*>(a.dir + di);
    const EfDecChunk d = ef_dec_unpack(__ldg(dp), __ldg(dp + 1));
    OutT* out;
    if (a.row_stride)
        out = reinterpret_cast<OutT*>(a.out) + warp * a.row_stride;
    else
        out = reinterpret_cast<OutT*>(a.out) + (a.sel_out ? a.sel_out[warp] : d.out_base) + d.r0;
    if (d.count) {
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
        if (lane == 0) ef_dec_stage(a, d, s_low, s_high, bar);
        mbar_wait(bar, 0);
        ef_dec_chunk<OutT>(a, d, s_low, s_high, hi_part, out, lane);
    }
    if (a.row_stride) {
        for (uint32_t t = d.count + lane; t < a.row_stride; t += 32) out[t] = (OutT)-1;
        if (a.counts && lane == 0) a.counts[warp] = d.count;
    }
}

struct EfSelArgs {
    const uint64_t* list_off;
    const uint8_t* l;
    const uint64_t* low_off;
    const uint64_t* high_off;
    const uint64_t* samp_off;
    const uint64_t* low;
    const uint64_t* high;
    const uint32_t* samples;
    const uint64_t* q_list;
    const uint64_t* q_off;
    int64_t* out;
    uint64_t nq;
    uint64_t nlist;
};

__global__ void __launch_bounds__(kThreads) k_ef_select(EfSelArgs a) {
    uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.nq) return;
    uint64_t L = a.q_list[q], k = a.q_off[q];
    int64_t r = -1;
    if (L < a.nlist && k < a.list_off[L + 1] - a.list_off[L])
        r = (int64_t)ef_select(a.low + a.low_off[L], a.high + a.high_off[L], a.samples + a.samp_off[L], a.l[L], k);
    a.out[q] = r;
}

// ---- fixed-width packing (BitstringWriter layout: value k at bits [k*bits, (k+1)*bits), LSB first)

template <typename T>
__global__ void __launch_bounds__(kThreads) k_bits_pack(const T* vals, uint64_t n, int bits, uint8_t* out, uint64_t out_bytes) {
    // one thread per output 32-bit word (tail bytes handled by the last thread)
    uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t nwords = (out_bytes + 3) / 4;
    if (w >= nwords) return;
    uint64_t bit0 = w * 32;
    uint64_t e = bit0 / (uint64_t)bits;
    int64_t shift = (int64_t)(e * bits) - (int64_t)bit0;
    uint64_t mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
    uint32_t acc = 0;
    while (e < n && shift < 32) {
        uint64_t f = (uint64_t)vals[e] & mask;
        acc |= shift >= 0 ? (uint32_t)(f << shift) : (uint32_t)(f >> (-shift));
        shift += bits;
        e++;
    }
    uint64_t byte0 = w * 4;
    if (byte0 + 4 <= out_bytes && ((uintptr_t)out & 3) == 0) {
        reinterpret_cast<uint32_t*>(out)[w] = acc;
    } else {
        for (int k = 0; k < 4 && byte0 + k < out_bytes; k++) out[byte0 + k] = (uint8_t)(acc >> (8 * k));
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) k_bits_unpack(const uint8_t* code, uint64_t code_bytes, uint64_t n, int bits, T* out) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint64_t pos = k * (uint64_t)bits, byte = pos >> 3;
    uint32_t sh = (uint32_t)(pos & 7);
    // up to 9 bytes cover bits+7 <= 71 bits
    uint64_t lo = 0;
    uint32_t need = (sh + (uint32_t)bits + 7) / 8;
    for (uint32_t j = 0; j < need && j < 8 && byte + j < code_bytes; j++) lo |= (uint64_t)code[byte + j] << (8 * j);
    uint64_t v = lo >> sh;
    if (need > 8 && byte + 8 < code_bytes) v |= (uint64_t)code[byte + 8] << (64 - sh);
    if (bits < 64) v &= (1ull << bits) - 1ull;
    out[k] = (T)v;
}

// ---- import: the derived arrays (chunk directory, select samples) from the bit vectors themselves.
// One CTA per list walks the list's chunks 128 at a time: popcount of the chunk's words, block-wide exclusive scan ->
// ids in front of every chunk; each thread then writes its chunk's descriptor and the samples (every 256-th one) that
// fall into it.
struct EfAuxArgs {
    const uint64_t* list_off;
    const uint8_t* l;
    const uint64_t* low_off;
    const uint64_t* high_off;
    const uint64_t* samp_off;
    const uint64_t* dir_off;
    const uint64_t* high;
    uint32_t* samples;
    EfChunk* dir;
    uint32_t nlist;
    uint32_t* status;  // kStRange: a list's upper-bits vector does not hold as many ones as the list has ids
};

__global__ void __launch_bounds__(kThreads) k_ef_rebuild_aux(EfAuxArgs a) {
    __shared__ uint32_t warp_tot[kThreads / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    for (uint32_t L = blockIdx.x; L < a.nlist; L += gridDim.x) {
        const uint64_t o0 = a.list_off[L], m = a.list_off[L + 1] - o0;
        if (m == 0) continue;  // the list's one descriptor stays zero: count 0
        const uint64_t h0 = a.high_off[L], hw = a.high_off[L + 1] - h0;
        const uint64_t nchunks = (hw + kDecChunkWords - 1) / kDecChunkWords;
        const uint32_t l = a.l[L];
        const uint64_t* hv = a.high + h0;
        uint64_t running = 0;
        for (uint64_t base = 0; base < nchunks; base += kThreads) {
            const uint64_t C = base + tid, W = C * kDecChunkWords;
            const uint32_t nw = C < nchunks ? (uint32_t)(hw - W < kDecChunkWords ? hw - W : kDecChunkWords) : 0u;
            uint32_t pop = 0;
            for (uint32_t w = 0; w < nw; w++) pop += (uint32_t)popc64(hv[W + w]);
            uint32_t sc = pop;  // inclusive scan over the CTA
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, sc, o);
                if ((int)lane >= o) sc += v;
            }
            __syncthreads();  // warp_tot of the previous round has been read
            if (lane == 31u) warp_tot[wid] = sc;
            __syncthreads();
            uint32_t wbase = 0, total = 0;
#pragma unroll
            for (int j = 0; j < kThreads / 32; j++) {
                const uint32_t t = warp_tot[j];
                wbase += (uint32_t)j < wid ? t : 0u;
                total += t;
            }
            const uint64_t before = running + wbase + sc - pop;
            if (nw) {
                EfChunk d;
                d.a = (2 * (h0 + W)) | ((uint64_t)pop << 40) | ((uint64_t)l << 51) | ((uint64_t)(2u * nw) << 56);
                d.low32 = 2 * a.low_off[L] + (before >> 5) * l;
                d.out_base = o0;
                d.b = before | ((W * 64 - before) << 32);
                a.dir[a.dir_off[L] + C] = d;
                if (pop) {
                    for (uint64_t s = (before + kEfSample - 1) >> kEfSampleLog; (s << kEfSampleLog) < before + pop; s++) {
                        uint32_t r = (uint32_t)((s << kEfSampleLog) - before);  // rank inside the chunk
                        for (uint32_t w = 0; w < nw; w++) {
                            const uint64_t word = hv[W + w];
                            const uint32_t cw = (uint32_t)popc64(word);
                            if (r < cw) {
                                (a.samples + a.samp_off[L])[s] = (uint32_t)((W + w) * 64 + select64(word, r));
                                break;
                            }
                            r -= cw;
                        }
                    }
                }
            }
            running += total;
        }
        if (tid == 0 && running != m) atomicOr(a.status, kStRange);
        __syncthreads();
    }
}

// ---- planning on the device: per-list shapes from (length, max id), then prefix sums (idc_prep.cuh: device_scan)
struct EfShapeArgs {
    const uint32_t* n;        // ids per list
    const uint32_t* hi;       // max id per list
    uint32_t nlist;
    uint8_t* l;
    uint64_t* sizes;          // [6][nlist]: lower-bits words, upper-bits words, samples, chunk descriptors, encoder tiles, ids
    uint64_t* totals;         // [0] bits_total, [1] max l, [2] a list whose upper-bits vector has >= 2^32 bits (its number + 1)
};

__global__ void __launch_bounds__(kThreads) k_ef_shapes(EfShapeArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
    uint64_t bits = 0;
    uint32_t l = 0;
    if (i < a.nlist) {
        const uint32_t n = a.n[i];
        const EfShape s = ef_shape(a.hi[i], n);
        l = s.l;
        a.l[i] = (uint8_t)s.l;
        a.sizes[0ull * a.nlist + i] = s.low_words;
        a.sizes[1ull * a.nlist + i] = s.high_words;
        a.sizes[2ull * a.nlist + i] = s.samples;
        const uint64_t nd = (s.high_words + kDecChunkWords - 1) / kDecChunkWords;
        a.sizes[3ull * a.nlist + i] = nd ? nd : 1ull;  // an empty list still owns one (zero) descriptor
        a.sizes[4ull * a.nlist + i] = ((uint64_t)n + kEncTileIds - 1) / kEncTileIds;
        a.sizes[5ull * a.nlist + i] = n;
        bits = s.low_bits + s.high_bits;
        if (s.high_bits >= (1ull << 32)) atomicMax(reinterpret_cast<unsigned long long*>(a.totals + 2), (unsigned long long)i + 1ull);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        bits += __shfl_xor_sync(0xffffffffu, bits, o);
        const uint32_t ol = __shfl_xor_sync(0xffffffffu, l, o);
        l = ol > l ? ol : l;
    }
    if (lane == 0) {
        if (bits) atomicAdd(reinterpret_cast<unsigned long long*>(a.totals), (unsigned long long)bits);
        if (l) atomicMax(reinterpret_cast<unsigned long long*>(a.totals + 1), (unsigned long long)l);
    }
}

// graph rows: row r starts at element r * K
__global__ void __launch_bounds__(kThreads) k_ef_row_src(uint64_t* src, uint64_t nrows, uint32_t K) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nrows) src[r] = r * K;
}

// ------------------------------------------------------------------ host

// Host mirrors of the per-list tables, fetched from the device on first use.
int ef_host_tables(const idc_ef_blob* b) {
    if (b->host_tables) return IDC_OK;
    idc_ctx* c = b->ctx;
    const uint64_t nl = b->nlist;
    cudaStream_t s = c->stream;
    std::vector<uint32_t> uni32(nl);
    b->list_offsets.resize(nl + 1);
    b->l.resize(nl);
    b->low_off.resize(nl + 1);
    b->high_off.resize(nl + 1);
    b->samp_off.resize(nl + 1);
    b->dir_off.resize(nl + 1);
    IDC_CUDA(cudaMemcpyAsync(b->list_offsets.data(), b->d_list_off, (nl + 1) * 8, cudaMemcpyDeviceToHost, s));
    IDC_CUDA(cudaMemcpyAsync(b->low_off.data(), b->d_low_off, (nl + 1) * 8, cudaMemcpyDeviceToHost, s));
    IDC_CUDA(cudaMemcpyAsync(b->high_off.data(), b->d_high_off, (nl + 1) * 8, cudaMemcpyDeviceToHost, s));
    IDC_CUDA(cudaMemcpyAsync(b->samp_off.data(), b->d_samp_off, (nl + 1) * 8, cudaMemcpyDeviceToHost, s));
    IDC_CUDA(cudaMemcpyAsync(b->dir_off.data(), b->d_dir_off, (nl + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (nl) IDC_CUDA(cudaMemcpyAsync(b->l.data(), b->d_l, nl, cudaMemcpyDeviceToHost, s));
    if (nl) IDC_CUDA(cudaMemcpyAsync(uni32.data(), b->d_universe, nl * 4, cudaMemcpyDeviceToHost, s));
    IDC_CUDA(cudaStreamSynchronize(s));
    b->universe.assign(uni32.begin(), uni32.end());
    b->host_tables = true;
    return IDC_OK;
}

// What the caller of ef_build has put on the device (c->meta is carved by ef_build itself: the caller fills the two
// arrays through the callbacks below once they exist).
struct EfBuildIn {
    const void* ids_dev;
    int id_bytes;
    uint32_t flags;
    uint64_t id_elems;
    const std::vector<uint64_t>* list_src;  // CSR mode: element offset of every list (host); null: graph rows
    const std::vector<uint32_t>* n_host;    // CSR mode: ids per list (host)
    uint32_t K;                             // graph rows: row r = elements [r K, r K + K), ids up to the first -1
    const uint32_t* d_cnt;                  // graph rows: ids per row (device)
};

// Everything per list is computed on the device: universe (max id), field width and vector sizes (k_ef_shapes), the
// offset tables (prefix sums); the host learns eight totals through one small copy. One million graph rows used to
// cost 20 ms of host loops and table uploads per call.
int ef_build(idc_ctx* c, idc_ef_blob* b, const EfBuildIn& in) {
    const uint64_t nl = b->nlist;
    const int id_bytes = in.id_bytes;
    const void* ids_dev = in.ids_dev;
    const bool rows = in.list_src == nullptr;
    HostTrace tr("ef_build");
    uint64_t acct = 0;
    IDC_TRY(dev_alloc(c, &b->d_list_off, nl + 1, &acct));
    IDC_TRY(dev_alloc(c, &b->d_universe, nl, &acct));
    IDC_TRY(dev_alloc(c, &b->d_l, nl, &acct));
    IDC_TRY(dev_alloc(c, &b->d_low_off, nl + 1, &acct));
    IDC_TRY(dev_alloc(c, &b->d_high_off, nl + 1, &acct));
    IDC_TRY(dev_alloc(c, &b->d_samp_off, nl + 1, &acct));
    IDC_TRY(dev_alloc(c, &b->d_dir_off, nl + 1, &acct));
    // per-call tables
    const size_t scan_bytes = scan_scratch_bytes(nl, 6);
    IDC_TRY(c->meta.reserve(nl * (8 + 4 + 4 + 1 + 6 * 8 + 8) + 8 + scan_bytes + 64 + 1024));
    uint8_t* mp = c->meta.as<uint8_t>();
    auto carve = [&](size_t bytes) {
        uint8_t* r = mp;
        mp += (bytes + 15) & ~size_t(15);
        return r;
    };
    uint64_t* d_src = (uint64_t*)carve(nl * 8);
    uint64_t* d_sizes = (uint64_t*)carve(nl * 6 * 8);
    uint64_t* d_tile_base = (uint64_t*)carve((nl + 1) * 8);
    uint64_t* d_scan = (uint64_t*)carve(scan_bytes);
    uint64_t* d_totals = (uint64_t*)carve(64);
    uint32_t* d_n = (uint32_t*)carve(nl * 4);
    uint32_t* d_lo = (uint32_t*)carve(nl * 4);
    uint8_t* d_prec = (uint8_t*)carve(nl);
    uint32_t* d_hi = b->d_universe;
    IDC_TRY(c->status.reserve(128));
    uint32_t* d_status = c->status.as<uint32_t>();
    IDC_CUDA(cudaMemsetAsync(d_status, 0, 4, c->stream));
    IDC_CUDA(cudaMemsetAsync(d_totals, 0, 64, c->stream));
    if (rows) {
        if (nl) {
            IDC_CUDA(cudaMemcpyAsync(d_n, in.d_cnt, nl * 4, cudaMemcpyDeviceToDevice, c->stream));
            LaunchScope ls(c, "k_ef_row_src");
            k_ef_row_src<<<grid_for(nl), kThreads, 0, c->stream>>>(d_src, nl, in.K);
        }
    } else {
        IDC_TRY(upload(c, d_src, *in.list_src));
        IDC_TRY(upload(c, d_n, *in.n_host));
    }
    const bool sorted_in = (in.flags & IDC_F_SORTED) != 0;
    if (sorted_in) {
        // the universe of an ascending list is its last id; order and width are verified by the encode kernel
        LaunchScope ls(c, "k_unit_meta");
        if (nl) {
            if (id_bytes == 8)
                k_list_ends<int64_t><<<grid_for(nl), kThreads, 0, c->stream>>>(ids_dev, d_src, d_n, (uint32_t)nl, d_lo, d_hi, d_status);
            else
                k_list_ends<uint32_t><<<grid_for(nl), kThreads, 0, c->stream>>>(ids_dev, d_src, d_n, (uint32_t)nl, d_lo, d_hi, d_status);
        }
    } else {
        MetaArgs m{ids_dev, d_src, d_n, (uint32_t)nl, 0u, 0u, d_prec, d_lo, d_hi, d_status, nullptr, nullptr, 0u};
        if (rows)
            IDC_TRY(run_unit_meta_small(c, m, nl, id_bytes));  // a row holds at most K <= 340 ids
        else
            IDC_TRY(run_unit_meta(c, m, *in.n_host, id_bytes));
    }
    IDC_TRY(check_last_launch("k_unit_meta"));
    // shapes + offset tables
    if (nl) {
        EfShapeArgs sa{d_n, d_hi, (uint32_t)nl, b->d_l, d_sizes, d_totals};
        {
            LaunchScope ls(c, "k_ef_shapes");
            k_ef_shapes<<<grid_for(nl), kThreads, 0, c->stream>>>(sa);
        }
        IDC_TRY(check_last_launch("k_ef_shapes"));
    }
    {
        const uint64_t* sin[6] = {d_sizes, d_sizes + nl, d_sizes + 2 * nl, d_sizes + 3 * nl, d_sizes + 4 * nl, d_sizes + 5 * nl};
        uint64_t* sout[6] = {b->d_low_off, b->d_high_off, b->d_samp_off, b->d_dir_off, d_tile_base, b->d_list_off};
        IDC_TRY(device_scan(c, 6, sin, sout, nl, d_scan));
    }
    uint64_t tot[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // low words, high words, samples, descriptors, tiles, ids, then d_totals[0..2)
    uint64_t tot2[3] = {0, 0, 0};
    uint32_t st = 0;
    {
        uint64_t* outs[6] = {b->d_low_off, b->d_high_off, b->d_samp_off, b->d_dir_off, d_tile_base, b->d_list_off};
        for (int j = 0; j < 6; j++) IDC_CUDA(cudaMemcpyAsync(&tot[j], outs[j] + nl, 8, cudaMemcpyDeviceToHost, c->stream));
        IDC_CUDA(cudaMemcpyAsync(tot2, d_totals, 24, cudaMemcpyDeviceToHost, c->stream));
        IDC_CUDA(cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, c->stream));
        IDC_CUDA(cudaStreamSynchronize(c->stream));
    }
    IDC_TRY(status_to_error(st, "ef_encode"));
    IDC_REQUIRE(tot2[2] == 0, IDC_ERR_DOMAIN, "list %llu: its upper-bits vector has 2^32 bits or more; the device path handles up to 2^32 - 1",
                (unsigned long long)(tot2[2] - 1));
    tr.mark("metadata, shapes, offset tables");
    b->low_words = tot[0];
    b->high_words = tot[1];
    b->nsamples = tot[2];
    b->ndir = tot[3];
    const uint64_t ntiles = tot[4];
    b->total_ids = tot[5];
    b->bits_total = tot2[0];
    b->max_l = (uint32_t)tot2[1];
    IDC_REQUIRE(ntiles < (1ull << 32), IDC_ERR_ARG, "too many encoder tiles");
    IDC_TRY(dev_alloc(c, &b->d_dir, b->ndir, &acct));
    IDC_TRY(dev_alloc(c, &b->d_low, b->low_words + 32, &acct));  // +256 B: the decoder's last group may read past the end
    IDC_TRY(dev_alloc(c, &b->d_high, b->high_words + 2, &acct));  // + 16 B: the decoder stages whole 16-byte pieces
    IDC_TRY(dev_alloc(c, &b->d_samples, b->nsamples, &acct));
    IDC_CUDA(cudaMemsetAsync(b->d_dir, 0, std::max<uint64_t>(b->ndir, 1) * sizeof(EfChunk), c->stream));  // empty lists: count 0
    tr.mark("alloc");
    // sort when needed (ids < 2^32 was checked by the metadata kernel)
    const void* enc_ids = ids_dev;
    int enc_id_bytes = id_bytes;
    size_t tile_bytes = (ntiles * sizeof(EfTile) + 255) & ~size_t(255);
    size_t ws_need = tile_bytes;
    size_t sorted_off = 0, sortidx_off = 0, big_off = 0, posbase_off = 0;
    uint32_t sort_grid = 0;
    bool any_big = false;
    if (!sorted_in) {
        sort_grid = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(nl, 1), (uint64_t)c->sm_count * 8);
        bool need_big = false;
        if (rows) {
            need_big = in.K > kSortSmem;
            any_big = in.K > kSortWarp;
        } else {
            for (uint64_t i = 0; i < nl; i++) {
                const uint32_t n = (*in.n_host)[i];
                IDC_REQUIRE(n <= kMaxUnit, IDC_ERR_DOMAIN,
                            "unsorted list %llu has %u ids: the device sort handles lists of <= 65536 ids; "
                            "pass ascending ids with IDC_F_SORTED", (unsigned long long)i, n);
                need_big |= n > kSortSmem;
                any_big |= n > kSortWarp;
            }
        }
        sorted_off = (ws_need + 255) & ~size_t(255);
        sortidx_off = sorted_off + ((in.id_elems * 4 + 255) & ~size_t(255));
        posbase_off = sortidx_off + ((in.id_elems * 4 + 255) & ~size_t(255));
        big_off = posbase_off + ((nl * 4 + 255) & ~size_t(255));
        ws_need = big_off + (need_big ? (size_t)sort_grid * kMaxUnit * 8 : 0);
    }
    IDC_TRY(c->ws.reserve(ws_need + 256));
    EfTile* d_tiles = reinterpret_cast<EfTile*>(c->ws.as<uint8_t>());
    if (!sorted_in && nl) {
        uint32_t* d_sorted = (uint32_t*)(c->ws.as<uint8_t>() + sorted_off);
        uint32_t* d_sort_idx = (uint32_t*)(c->ws.as<uint8_t>() + sortidx_off);
        uint32_t* d_posbase = (uint32_t*)(c->ws.as<uint8_t>() + posbase_off);
        IDC_CUDA(cudaMemsetAsync(d_posbase, 0, nl * 4, c->stream));
        SortArgs s{ids_dev, d_src, d_n, d_posbase, (uint32_t)nl, d_sorted, d_sort_idx,
                   (uint64_t*)(c->ws.as<uint8_t>() + big_off)};
        IDC_TRY(launch_sorts(c, s, id_bytes, sort_grid, any_big));
        enc_ids = d_sorted;
        enc_id_bytes = 4;
    }
    IDC_TRY(check_last_launch("k_sort_units"));
    // tiles OR their first / last upper-bits word into the array: it starts out zero
    IDC_CUDA(cudaMemsetAsync(b->d_high, 0, std::max<uint64_t>(b->high_words, 1) * 8, c->stream));
    if (ntiles) {
        {
            LaunchScope ls(c, "k_ef_tile_desc");
            EfTileDescArgs t{d_src, b->d_list_off, b->d_l, b->d_low_off, b->d_high_off, b->d_samp_off, b->d_dir_off, d_tile_base, (uint32_t)nl, d_tiles};
            k_ef_tile_desc<<<grid_for(ntiles), kThreads, 0, c->stream>>>(t, (uint32_t)ntiles);
        }
        IDC_TRY(check_last_launch("k_ef_tile_desc"));
        EfEncArgs e{enc_ids, b->d_low, b->d_high, b->d_samples, b->d_dir, d_tiles, (uint32_t)ntiles, sorted_in ? 1u : 0u, d_status};
        // persistent warps: as many CTAs as stay resident (shared memory: one raw tile + the 32-bit tile + the
        // window per warp), each warp walks the tile list with the stride of the whole grid
        const size_t smem = kEncWarps * (enc_id_bytes == 8 ? sizeof(EfEncSmem<int64_t>) : sizeof(EfEncSmem<uint32_t>));
        int per_sm = 0;
        if (enc_id_bytes == 8) {
            IDC_CUDA(cudaFuncSetAttribute(k_ef_encode<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            IDC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ef_encode<int64_t>, kEncWarps * 32, smem));
        } else {
            IDC_CUDA(cudaFuncSetAttribute(k_ef_encode<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            IDC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ef_encode<uint32_t>, kEncWarps * 32, smem));
        }
        if (const char* ev = getenv("IDC_EF_ENC_CTAS_PER_SM")) per_sm = std::min(per_sm, std::max(1, atoi(ev)));  // experiments
        const uint64_t want = (ntiles + kEncWarps - 1) / kEncWarps;
        const uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)std::max(per_sm, 1) * c->sm_count);
        LaunchScope ls(c, "k_ef_encode");
        if (enc_id_bytes == 8)
            k_ef_encode<int64_t><<<grid, kEncWarps * 32, smem, c->stream>>>(e);
        else
            k_ef_encode<uint32_t><<<grid, kEncWarps * 32, smem, c->stream>>>(e);
    }
    IDC_TRY(check_last_launch("k_ef_encode"));
    if (b->ndir) {
        LaunchScope ls(c, "k_ef_finish_chunks");
        k_ef_finish_chunks<<<grid_for(b->ndir), kThreads, 0, c->stream>>>(b->d_dir, b->ndir);
    }
    IDC_TRY(check_last_launch("k_ef_finish_chunks"));
    {
        uint32_t st2 = 0;
        IDC_CUDA(cudaMemcpyAsync(&st2, d_status, 4, cudaMemcpyDeviceToHost, c->stream));
        IDC_CUDA(cudaStreamSynchronize(c->stream));
        IDC_TRY(status_to_error(st2, "ef_encode"));
    }
    tr.mark("sort + encode kernels");
    b->device_bytes = acct;
    return IDC_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the function on a device, shared by every context of the
// process: raised when a launch needs more, never lowered (two driver calls less per single-row call).
int ef_dec_smem_limit(int device, size_t smem) {
    static std::mutex mu;
    static size_t have[64] = {};
    std::lock_guard<std::mutex> lock(mu);
    size_t& h = have[(unsigned)device & 63u];
    if (smem <= h) return IDC_OK;
    IDC_CUDA(cudaFuncSetAttribute(k_ef_decode<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IDC_CUDA(cudaFuncSetAttribute(k_ef_decode<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    h = smem;
    return IDC_OK;
}

// rows_checked: the row numbers were range-checked by the caller (host rows), the kernel cannot raise kStRange and the
// status word is neither cleared nor fetched. Returns with the stream synchronised.
int ef_run_decode(idc_ctx* c, const idc_ef_blob* b, const uint32_t* d_sel_desc, const uint64_t* d_sel_out,
                  const int32_t* d_row_nos, uint64_t ntiles, void* out_dev, int id_bytes, uint32_t* counts_dev,
                  uint32_t row_stride, bool rows_checked = false) {
    if (ntiles == 0) return IDC_OK;
    // stage up to 33 groups of max_l words per warp, capped so that 6 CTAs of 8 warps still fit an SM (chunks that need
    // more -- short lists with wide fields -- read their lower bits straight from global memory)
    uint32_t stage_words = std::min<uint32_t>(33u * b->max_l + 1u, kDecStageCap);
    if (const char* ev = getenv("IDC_EF_DEC_STAGE")) stage_words = std::min<uint32_t>(33u * b->max_l + 1u, (uint32_t)atoi(ev));  // experiments
    stage_words = (stage_words + 3u) & ~3u;
    IDC_TRY(c->status.reserve(128));
    uint32_t* d_status = c->status.as<uint32_t>();
    const bool want_status = row_stride && !rows_checked;
    if (want_status) IDC_CUDA(cudaMemsetAsync(d_status, 0, 4, c->stream));
    EfDecArgs a{b->d_dir, b->d_low, b->d_high, d_sel_desc, d_sel_out, d_row_nos, out_dev, counts_dev, ntiles, row_stride,
                stage_words, b->nlist, d_status};
    const uint32_t wpb = kDecThreads / 32;
    const uint32_t grid = (uint32_t)((ntiles + wpb - 1) / wpb);
    const size_t smem = (size_t)wpb * ef_dec_warp_bytes(stage_words, 1) + 512u;  // + 512: the output pass reads (and ignores) up to 3 groups past a chunk's words
    IDC_TRY(ef_dec_smem_limit(c->device, smem));
    {
        LaunchScope ls(c, "k_ef_decode");
        if (id_bytes == 8)
            k_ef_decode<int64_t><<<grid, kDecThreads, smem, c->stream>>>(a);
        else
            k_ef_decode<int32_t><<<grid, kDecThreads, smem, c->stream>>>(a);
    }
    IDC_TRY(check_last_launch("k_ef_decode"));
    uint32_t st = 0;
    if (want_status) IDC_CUDA(cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    IDC_REQUIRE((st & kStRange) == 0, IDC_ERR_ARG, "ef_decode_rows: a row number is out of range (such rows were set to -1)");
    return IDC_OK;
}

}  // namespace

extern "C" {

int idc_ef_encode(idc_ctx* c, uint64_t nlist, const uint64_t* offsets, const void* ids, int id_bytes, int ids_mem,
                  uint32_t flags, idc_ef_blob** out) {
    IDC_REQUIRE(c && offsets && out, IDC_ERR_ARG, "idc_ef_encode: null argument");
    IDC_REQUIRE(id_bytes == 8 || id_bytes == 4, IDC_ERR_ARG, "id_bytes must be 4 or 8");
    IDC_REQUIRE(nlist < (1ull << 32), IDC_ERR_ARG, "too many lists");
    *out = nullptr;
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    std::unique_ptr<idc_ef_blob> b(new idc_ef_blob());
    b->ctx = c;
    b->ref.bind(c);
    b->nlist = nlist;
    b->list_offsets.resize(nlist + 1);
    std::vector<uint64_t> src(nlist);
    std::vector<uint32_t> n32(nlist);
    for (uint64_t l = 0; l <= nlist; l++) {
        IDC_REQUIRE(l == 0 || offsets[l] >= offsets[l - 1], IDC_ERR_ARG, "offsets must be non-decreasing");
        b->list_offsets[l] = offsets[l] - offsets[0];
        if (l < nlist) {
            src[l] = offsets[l];
            IDC_REQUIRE(offsets[l + 1] < offsets[l] || offsets[l + 1] - offsets[l] < (1ull << 32), IDC_ERR_ARG, "list %llu too long",
                        (unsigned long long)l);
            n32[l] = (uint32_t)(offsets[l + 1] - offsets[l]);
        }
    }
    b->total_ids = b->list_offsets[nlist];
    uint64_t elems = offsets[nlist];
    IDC_REQUIRE(ids != nullptr || elems == 0, IDC_ERR_ARG, "ids is NULL");
    const void* ids_dev = ids;
    if (ids_mem == IDC_MEM_HOST && elems) {
        IDC_TRY(c->stage.reserve(elems * id_bytes + 64));  // + 64: the encoder's bulk copies move whole 16-byte pieces
        IDC_CUDA(cudaMemcpyAsync(c->stage.p, ids, elems * id_bytes, cudaMemcpyHostToDevice, c->stream));
        ids_dev = c->stage.p;
    }
    EfBuildIn in{ids_dev, id_bytes, flags, elems, &src, &n32, 0u, nullptr};
    IDC_TRY(ef_build(c, b.get(), in));
    *out = b.release();
    return IDC_OK;
}

int idc_ef_encode_rows(idc_ctx* c, uint64_t nrows, uint32_t K, const int32_t* data, int data_mem, uint32_t flags,
                       idc_ef_blob** out) {
    IDC_REQUIRE(c && out && (data || nrows == 0), IDC_ERR_ARG, "idc_ef_encode_rows: null argument");
    // what idc_ef_decode_rows can read back: a row's upper-bits vector (<= 3 K + 2 bits) must fit one decoder chunk
    IDC_REQUIRE(K >= 1 && K <= kDecTile && 3ull * K + 2 <= 64ull * kDecChunkWords, IDC_ERR_ARG,
                "K = %u out of range: rows of up to %u entries are supported", K, (unsigned)((64u * kDecChunkWords - 2u) / 3u));
    IDC_REQUIRE(nrows < (1ull << 32), IDC_ERR_ARG, "too many rows");
    *out = nullptr;
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    std::unique_ptr<idc_ef_blob> b(new idc_ef_blob());
    b->ctx = c;
    b->ref.bind(c);
    b->nlist = nrows;
    b->row_stride = K;
    const int32_t* d_data = data;
    uint64_t elems = nrows * K;
    if (data_mem == IDC_MEM_HOST && elems) {
        IDC_TRY(c->stage.reserve(elems * 4 + 64));  // + 64: bulk copies move whole 16-byte pieces
        IDC_CUDA(cudaMemcpyAsync(c->stage.p, data, elems * 4, cudaMemcpyHostToDevice, c->stream));
        d_data = c->stage.as<int32_t>();
    }
    // row lengths on the device; they stay there (shapes, offsets and the sort are planned on the device)
    IDC_TRY(c->scratch.reserve(nrows * 4 + 256));
    uint32_t* d_cnt = c->scratch.as<uint32_t>();
    if (nrows) {
        {
            LaunchScope ls(c, "k_row_counts");
            k_row_counts<<<grid_for(nrows * 32), kThreads, 0, c->stream>>>(d_data, nrows, K, d_cnt);
        }
        IDC_TRY(check_last_launch("k_row_counts"));
    }
    EfBuildIn in{d_data, 4, flags & ~IDC_F_SORTED, elems, nullptr, nullptr, K, d_cnt};
    IDC_TRY(ef_build(c, b.get(), in));
    *out = b.release();
    return IDC_OK;
}

int idc_ef_blob_info(const idc_ef_blob* b, idc_ef_info* info) {
    IDC_REQUIRE(b && info, IDC_ERR_ARG, "null argument");
    info->nlist = b->nlist;
    info->total_ids = b->total_ids;
    info->low_words = b->low_words;
    info->high_words = b->high_words;
    info->bits_total = b->bits_total;
    info->device_bytes = b->device_bytes;
    info->row_stride = b->row_stride;
    return IDC_OK;
}

int idc_ef_blob_export(const idc_ef_blob* b, uint64_t* list_offsets, uint8_t* l, uint64_t* universe,
                       uint64_t* low_offsets, uint64_t* high_offsets, uint64_t* low, uint64_t* high) {
    IDC_REQUIRE(b, IDC_ERR_ARG, "null blob");
    IDC_CUDA(cudaSetDevice(b->ctx->device));
    cudaStream_t s = b->ctx->stream;
    if (list_offsets || l || universe || low_offsets || high_offsets) {
        std::lock_guard<std::mutex> lock(b->ctx->mu);
        IDC_TRY(ef_host_tables(b));
    }
    if (list_offsets) memcpy(list_offsets, b->list_offsets.data(), (b->nlist + 1) * 8);
    if (l && b->nlist) memcpy(l, b->l.data(), b->nlist);
    if (universe && b->nlist) memcpy(universe, b->universe.data(), b->nlist * 8);
    if (low_offsets) memcpy(low_offsets, b->low_off.data(), (b->nlist + 1) * 8);
    if (high_offsets) memcpy(high_offsets, b->high_off.data(), (b->nlist + 1) * 8);
    if (low && b->low_words) IDC_CUDA(cudaMemcpyAsync(low, b->d_low, b->low_words * 8, cudaMemcpyDeviceToHost, s));
    if (high && b->high_words) IDC_CUDA(cudaMemcpyAsync(high, b->d_high, b->high_words * 8, cudaMemcpyDeviceToHost, s));
    IDC_CUDA(cudaStreamSynchronize(s));
    return IDC_OK;
}

int idc_ef_blob_free(idc_ef_blob* b) {
    if (b) {
        std::lock_guard<std::mutex> lock(b->ctx->mu);
        cudaSetDevice(b->ctx->device);
        delete b;
    }
    return IDC_OK;
}

int idc_ef_decode(idc_ctx* c, const idc_ef_blob* b, const uint64_t* list_nos, uint64_t nsel, void* ids_out,
                  int id_bytes, int out_mem, uint64_t* out_offsets) {
    IDC_REQUIRE(c && b, IDC_ERR_ARG, "idc_ef_decode: null argument");
    IDC_REQUIRE(id_bytes == 8 || id_bytes == 4, IDC_ERR_ARG, "id_bytes must be 4 or 8");
    IDC_REQUIRE(b->row_stride == 0, IDC_ERR_ARG, "row blob: use idc_ef_decode_rows");
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    uint64_t total_out = 0, ntiles = 0;
    uint32_t* t_desc = nullptr;
    uint64_t* t_out = nullptr;
    if (list_nos != nullptr || out_offsets != nullptr) IDC_TRY(ef_host_tables(b));
    {  // argument errors surface before anything is allocated
        uint64_t want = 0;
        if (list_nos == nullptr) want = b->total_ids;
        else
            for (uint64_t i = 0; i < nsel; i++) {
                IDC_REQUIRE(list_nos[i] < b->nlist, IDC_ERR_ARG, "list_no out of range");
                want += b->list_offsets[list_nos[i] + 1] - b->list_offsets[list_nos[i]];
            }
        IDC_REQUIRE(want == 0 || ids_out != nullptr, IDC_ERR_ARG, "ids_out is NULL");
    }
    if (list_nos == nullptr) {
        // decode everything: tile i is chunk descriptor i, no per-call planning at all
        ntiles = b->ndir;
        total_out = b->total_ids;
        if (out_offsets) memcpy(out_offsets, b->list_offsets.data(), (b->nlist + 1) * 8);
    } else {
        std::vector<uint32_t> td;
        std::vector<uint64_t> to;
        uint64_t pos = 0;
        for (uint64_t i = 0; i < nsel; i++) {
            uint64_t L = list_nos[i];
            IDC_REQUIRE(L < b->nlist, IDC_ERR_ARG, "list_no out of range");
            if (out_offsets) out_offsets[i] = pos;
            for (uint64_t d = b->dir_off[L]; d < b->dir_off[L + 1]; d++) {
                td.push_back((uint32_t)d);
                to.push_back(pos);
            }
            pos += b->list_offsets[L + 1] - b->list_offsets[L];
        }
        if (out_offsets) out_offsets[nsel] = pos;
        total_out = pos;
        ntiles = td.size();
        IDC_REQUIRE(b->ndir < (1ull << 32), IDC_ERR_ARG, "too many chunks for subset decode");
        IDC_TRY(dev_alloc(c, &t_desc, ntiles));
        IDC_TRY(dev_alloc(c, &t_out, ntiles));
        IDC_TRY(upload(c, t_desc, td));
        IDC_TRY(upload(c, t_out, to));
    }
    int rc = IDC_OK;
    if (total_out) {
        IDC_REQUIRE(ids_out != nullptr, IDC_ERR_ARG, "ids_out is NULL");
        void* out_dev = ids_out;
        if (out_mem == IDC_MEM_HOST) {
            rc = c->stage.reserve(total_out * id_bytes);
            out_dev = c->stage.p;
        }
        if (rc == IDC_OK) rc = ef_run_decode(c, b, t_desc, t_out, nullptr, ntiles, out_dev, id_bytes, nullptr, 0);
        if (rc == IDC_OK && out_mem == IDC_MEM_HOST) {
            cudaError_t e = cudaMemcpyAsync(ids_out, out_dev, total_out * id_bytes, cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) {
                set_error("D2H copy failed: %s", cudaGetErrorString(e));
                rc = IDC_ERR_CUDA;
            }
        }
    } else {
        cudaStreamSynchronize(c->stream);
    }
    c->pool_release(t_desc);
    c->pool_release(t_out);
    return rc;
}

int idc_ef_decode_rows(idc_ctx* c, const idc_ef_blob* b, const int32_t* row_nos, int rows_mem, uint64_t nsel,
                       int32_t* out, uint32_t* counts, int out_mem) {
    IDC_REQUIRE(c && b, IDC_ERR_ARG, "idc_ef_decode_rows: null argument");
    IDC_REQUIRE(b->row_stride != 0, IDC_ERR_ARG, "not a row blob");
    IDC_REQUIRE(b->row_stride <= kDecTile, IDC_ERR_ARG, "row stride > %u not supported", kDecTile);
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    const uint32_t K = b->row_stride;
    if (row_nos == nullptr) nsel = b->nlist;
    if (nsel == 0) return IDC_OK;
    IDC_REQUIRE(out != nullptr, IDC_ERR_ARG, "out is NULL");
    IDC_REQUIRE(3ull * K + 2 <= 64ull * kDecChunkWords, IDC_ERR_ARG, "row stride %u too large: a row must fit one 2048-bit chunk", K);
    // rows are addressed directly by the kernel (list = row number, chunk 0): no tile tables
    const int32_t* d_rows = row_nos;
    const bool host_rows = row_nos && rows_mem == IDC_MEM_HOST;
    if (host_rows)
        for (uint64_t i = 0; i < nsel; i++)
            IDC_REQUIRE(row_nos[i] >= 0 && (uint64_t)row_nos[i] < b->nlist, IDC_ERR_ARG, "row %d out of range", row_nos[i]);
    if (host_rows && out_mem == IDC_MEM_HOST && nsel <= 4096) {
        // a few rows, everything on the host (what NSG search does per visited node, altid_impl.cpp:92-101): through the
        // context's mailbox -- [row numbers | counts | rows], read and written by the kernel in place
        const size_t o_cnt = (nsel * 4 + 15) & ~size_t(15), o_out = 2 * o_cnt;
        void *mh = nullptr, *md = nullptr;
        IDC_TRY(c->mailbox_get(o_out + nsel * K * 4, &mh, &md));
        if (mh) {
            std::memcpy(mh, row_nos, nsel * 4);
            uint8_t *h8 = static_cast<uint8_t*>(mh), *d8 = static_cast<uint8_t*>(md);
            IDC_TRY(ef_run_decode(c, b, nullptr, nullptr, reinterpret_cast<const int32_t*>(d8), nsel, d8 + o_out, 4,
                                  reinterpret_cast<uint32_t*>(d8 + o_cnt), K, true));  // synchronises the stream
            std::memcpy(out, h8 + o_out, nsel * K * 4);
            if (counts) std::memcpy(counts, h8 + o_cnt, nsel * 4);
            return IDC_OK;
        }
    }
    if (host_rows) {
        IDC_TRY(c->meta.reserve(nsel * 4 + 256));
        IDC_CUDA(cudaMemcpyAsync(c->meta.p, row_nos, nsel * 4, cudaMemcpyHostToDevice, c->stream));
        d_rows = c->meta.as<int32_t>();
    }
    int32_t* out_dev = out;
    uint32_t* cnt_dev = counts;
    if (out_mem == IDC_MEM_HOST) {
        IDC_TRY(c->stage.reserve(nsel * K * 4 + nsel * 4 + 256));
        out_dev = c->stage.as<int32_t>();
        cnt_dev = reinterpret_cast<uint32_t*>(c->stage.as<uint8_t>() + nsel * K * 4);
    }
    IDC_TRY(ef_run_decode(c, b, nullptr, nullptr, d_rows, nsel, out_dev, 4, cnt_dev, K, host_rows));
    if (out_mem == IDC_MEM_HOST) {
        IDC_CUDA(cudaMemcpyAsync(out, out_dev, nsel * K * 4, cudaMemcpyDeviceToHost, c->stream));
        if (counts) IDC_CUDA(cudaMemcpyAsync(counts, cnt_dev, nsel * 4, cudaMemcpyDeviceToHost, c->stream));
        IDC_CUDA(cudaStreamSynchronize(c->stream));
    }
    return IDC_OK;
}

int idc_ef_select(idc_ctx* c, const idc_ef_blob* b, const uint64_t* list_nos, const uint64_t* offsets_in_list,
                  uint64_t nq, int query_mem, int64_t* ids_out, int out_mem) {
    IDC_REQUIRE(c && b && (nq == 0 || (list_nos && offsets_in_list && ids_out)), IDC_ERR_ARG,
                "idc_ef_select: null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    if (nq == 0) return IDC_OK;
    if (query_mem == IDC_MEM_HOST && out_mem == IDC_MEM_HOST && nq <= 2048) {
        // a few queries from the host (get_single_id, custom_invlists_impl.cpp:314-318): through the context's mailbox
        // -- [list numbers | offsets | ids], read and written by the kernel in place: one launch, one synchronisation
        void *mh = nullptr, *md = nullptr;
        IDC_TRY(c->mailbox_get(nq * 24, &mh, &md));
        if (mh) {
            uint64_t *h = static_cast<uint64_t*>(mh), *d = static_cast<uint64_t*>(md);
            std::memcpy(h, list_nos, nq * 8);
            std::memcpy(h + nq, offsets_in_list, nq * 8);
            EfSelArgs a{b->d_list_off, b->d_l, b->d_low_off, b->d_high_off, b->d_samp_off, b->d_low, b->d_high, b->d_samples,
                        d, d + nq, reinterpret_cast<int64_t*>(d + 2 * nq), nq, b->nlist};
            {
                LaunchScope ls(c, "k_ef_select");
                k_ef_select<<<grid_for(nq), kThreads, 0, c->stream>>>(a);
            }
            IDC_TRY(check_last_launch("k_ef_select"));
            IDC_CUDA(cudaStreamSynchronize(c->stream));
            std::memcpy(ids_out, h + 2 * nq, nq * 8);
            return IDC_OK;
        }
    }
    const uint64_t *d_ql = list_nos, *d_qo = offsets_in_list;
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
    int64_t* out_dev = out_mem == IDC_MEM_HOST ? reinterpret_cast<int64_t*>(sp) : ids_out;
    EfSelArgs a{b->d_list_off, b->d_l, b->d_low_off, b->d_high_off, b->d_samp_off, b->d_low, b->d_high, b->d_samples,
                d_ql, d_qo, out_dev, nq, b->nlist};
    {
        LaunchScope ls(c, "k_ef_select");
        k_ef_select<<<grid_for(nq), kThreads, 0, c->stream>>>(a);
    }
    IDC_TRY(check_last_launch("k_ef_select"));
    if (out_mem == IDC_MEM_HOST)
        IDC_CUDA(cudaMemcpyAsync(ids_out, out_dev, nq * 8, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    return IDC_OK;
}

int idc_bits_pack(idc_ctx* c, uint64_t n, const void* vals, int val_bytes, int vals_mem, int bits, uint8_t* out,
                  uint64_t out_bytes, int out_mem) {
    IDC_REQUIRE(c && (n == 0 || vals) && (out_bytes == 0 || out), IDC_ERR_ARG, "idc_bits_pack: null argument");
    IDC_REQUIRE(val_bytes == 8 || val_bytes == 4, IDC_ERR_ARG, "val_bytes must be 4 or 8");
    IDC_REQUIRE(bits >= 1 && bits <= 8 * val_bytes, IDC_ERR_ARG, "bits out of range");
    IDC_REQUIRE(out_bytes * 8 >= n * (uint64_t)bits, IDC_ERR_ARG, "output too small");
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    if (out_bytes == 0) return IDC_OK;
    size_t in_b = vals_mem == IDC_MEM_HOST ? ((n * val_bytes + 255) & ~size_t(255)) : 0;
    size_t out_b = out_mem == IDC_MEM_HOST ? out_bytes + 8 : 0;
    IDC_TRY(c->stage.reserve(in_b + out_b + 256));
    const void* d_vals = vals;
    uint8_t* d_out = out;
    if (vals_mem == IDC_MEM_HOST) {
        IDC_CUDA(cudaMemcpyAsync(c->stage.p, vals, n * val_bytes, cudaMemcpyHostToDevice, c->stream));
        d_vals = c->stage.p;
    }
    if (out_mem == IDC_MEM_HOST) d_out = c->stage.as<uint8_t>() + in_b;
    uint64_t nwords = (out_bytes + 3) / 4;
    {
        LaunchScope ls(c, "k_bits_pack");
        if (val_bytes == 8)
            k_bits_pack<uint64_t><<<grid_for(nwords), kThreads, 0, c->stream>>>((const uint64_t*)d_vals, n, bits, d_out, out_bytes);
        else
            k_bits_pack<uint32_t><<<grid_for(nwords), kThreads, 0, c->stream>>>((const uint32_t*)d_vals, n, bits, d_out, out_bytes);
    }
    IDC_TRY(check_last_launch("k_bits_pack"));
    if (out_mem == IDC_MEM_HOST) IDC_CUDA(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    return IDC_OK;
}

int idc_bits_unpack(idc_ctx* c, uint64_t n, const uint8_t* code, uint64_t code_bytes, int code_mem, int bits, void* out,
                    int val_bytes, int out_mem) {
    IDC_REQUIRE(c && (n == 0 || (code && out)), IDC_ERR_ARG, "idc_bits_unpack: null argument");
    IDC_REQUIRE(val_bytes == 8 || val_bytes == 4, IDC_ERR_ARG, "val_bytes must be 4 or 8");
    IDC_REQUIRE(bits >= 1 && bits <= 8 * val_bytes, IDC_ERR_ARG, "bits out of range");
    IDC_REQUIRE(code_bytes * 8 >= n * (uint64_t)bits, IDC_ERR_ARG, "code too small");
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    if (n == 0) return IDC_OK;
    size_t in_b = code_mem == IDC_MEM_HOST ? ((code_bytes + 255) & ~size_t(255)) : 0;
    size_t out_b = out_mem == IDC_MEM_HOST ? n * val_bytes : 0;
    IDC_TRY(c->stage.reserve(in_b + out_b + 256));
    const uint8_t* d_code = code;
    void* d_out = out;
    if (code_mem == IDC_MEM_HOST) {
        IDC_CUDA(cudaMemcpyAsync(c->stage.p, code, code_bytes, cudaMemcpyHostToDevice, c->stream));
        d_code = c->stage.as<uint8_t>();
    }
    if (out_mem == IDC_MEM_HOST) d_out = c->stage.as<uint8_t>() + in_b;
    {
        LaunchScope ls(c, "k_bits_unpack");
        if (val_bytes == 8)
            k_bits_unpack<uint64_t><<<grid_for(n), kThreads, 0, c->stream>>>(d_code, code_bytes, n, bits, (uint64_t*)d_out);
        else
            k_bits_unpack<uint32_t><<<grid_for(n), kThreads, 0, c->stream>>>(d_code, code_bytes, n, bits, (uint32_t*)d_out);
    }
    IDC_TRY(check_last_launch("k_bits_unpack"));
    if (out_mem == IDC_MEM_HOST) IDC_CUDA(cudaMemcpyAsync(out, d_out, n * val_bytes, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    return IDC_OK;
}

/* Build a blob from the exported form: the list CSR, every list's universe (max id) and the two bit vectors (HOST or
 * DEVICE per `mem`); field widths and word offsets follow from (universe, list length) exactly as at encode time,
 * the chunk directory and the select samples are rebuilt on the device. row_stride != 0: graph rows. */
int idc_ef_blob_import(idc_ctx* c, uint64_t nlist, const uint64_t* list_offsets, const uint64_t* universe, uint32_t row_stride,
                       const uint64_t* low, const uint64_t* high, int mem, idc_ef_blob** out) {
    IDC_REQUIRE(c && out && list_offsets && (universe || nlist == 0), IDC_ERR_ARG, "idc_ef_blob_import: null argument");
    IDC_REQUIRE(mem == IDC_MEM_HOST || mem == IDC_MEM_DEVICE, IDC_ERR_ARG, "mem must be IDC_MEM_HOST or IDC_MEM_DEVICE");
    IDC_REQUIRE(nlist < (1ull << 32), IDC_ERR_ARG, "too many lists");
    *out = nullptr;
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    std::unique_ptr<idc_ef_blob> b(new idc_ef_blob());
    b->ctx = c;
    b->ref.bind(c);
    b->nlist = nlist;
    b->row_stride = row_stride;
    const uint64_t nl = nlist;
    b->list_offsets.resize(nl + 1);
    for (uint64_t i = 0; i <= nl; i++) {
        IDC_REQUIRE(i == 0 || list_offsets[i] >= list_offsets[i - 1], IDC_ERR_ARG, "offsets must be non-decreasing");
        b->list_offsets[i] = list_offsets[i] - list_offsets[0];
    }
    b->total_ids = b->list_offsets[nl];
    b->l.resize(nl);
    b->universe.assign(universe, universe + nl);
    b->low_off.assign(nl + 1, 0);
    b->high_off.assign(nl + 1, 0);
    b->samp_off.assign(nl + 1, 0);
    b->dir_off.assign(nl + 1, 0);
    uint64_t bits_total = 0;
    for (uint64_t i = 0; i < nl; i++) {
        const uint64_t n = b->list_offsets[i + 1] - b->list_offsets[i];
        IDC_REQUIRE(n < (1ull << 32) && universe[i] < (1ull << 32), IDC_ERR_DOMAIN, "list %llu: length or universe beyond 32 bits",
                    (unsigned long long)i);
        IDC_REQUIRE(!row_stride || n <= row_stride, IDC_ERR_ARG, "row %llu holds more than row_stride entries", (unsigned long long)i);
        EfShape sh = ef_shape(universe[i], n);
        IDC_REQUIRE(sh.high_bits < (1ull << 32), IDC_ERR_DOMAIN, "list %llu: upper-bits vector of %llu bits; the device path handles up to 2^32 - 1",
                    (unsigned long long)i, (unsigned long long)sh.high_bits);
        b->l[i] = (uint8_t)sh.l;
        b->max_l = std::max<uint32_t>(b->max_l, sh.l);
        b->low_off[i + 1] = b->low_off[i] + sh.low_words;
        b->high_off[i + 1] = b->high_off[i] + sh.high_words;
        b->samp_off[i + 1] = b->samp_off[i] + sh.samples;
        b->dir_off[i + 1] = b->dir_off[i] + std::max<uint64_t>(1, (sh.high_words + kDecChunkWords - 1) / kDecChunkWords);
        bits_total += sh.low_bits + sh.high_bits;
    }
    if (row_stride)
        IDC_REQUIRE(row_stride <= kDecTile && 3ull * row_stride + 2 <= 64ull * kDecChunkWords, IDC_ERR_ARG, "row_stride %u out of range", row_stride);
    b->low_words = b->low_off[nl];
    b->high_words = b->high_off[nl];
    b->nsamples = b->samp_off[nl];
    b->ndir = b->dir_off[nl];
    b->bits_total = bits_total;
    IDC_REQUIRE((low || b->low_words == 0) && (high || b->high_words == 0), IDC_ERR_ARG, "idc_ef_blob_import: null bit vector");
    uint64_t acct = 0;
    IDC_TRY(dev_alloc(c, &b->d_list_off, nl + 1, &acct));
    IDC_TRY(dev_alloc(c, &b->d_universe, nl, &acct));
    IDC_TRY(dev_alloc(c, &b->d_l, nl, &acct));
    IDC_TRY(dev_alloc(c, &b->d_low_off, nl + 1, &acct));
    IDC_TRY(dev_alloc(c, &b->d_high_off, nl + 1, &acct));
    IDC_TRY(dev_alloc(c, &b->d_samp_off, nl + 1, &acct));
    IDC_TRY(dev_alloc(c, &b->d_dir_off, nl + 1, &acct));
    IDC_TRY(dev_alloc(c, &b->d_dir, b->ndir, &acct));
    IDC_TRY(dev_alloc(c, &b->d_low, b->low_words + 32, &acct));
    IDC_TRY(dev_alloc(c, &b->d_high, b->high_words + 2, &acct));
    IDC_TRY(dev_alloc(c, &b->d_samples, b->nsamples, &acct));
    b->host_tables = true;  // the mirrors were computed right here
    {
        std::vector<uint32_t> uni32(b->universe.begin(), b->universe.end());
        IDC_TRY(upload(c, b->d_universe, uni32));
        IDC_CUDA(cudaStreamSynchronize(c->stream));  // uni32 goes out of scope
    }
    IDC_TRY(upload(c, b->d_list_off, b->list_offsets));
    IDC_TRY(upload(c, b->d_l, b->l));
    IDC_TRY(upload(c, b->d_low_off, b->low_off));
    IDC_TRY(upload(c, b->d_high_off, b->high_off));
    IDC_TRY(upload(c, b->d_samp_off, b->samp_off));
    IDC_TRY(upload(c, b->d_dir_off, b->dir_off));
    IDC_CUDA(cudaMemsetAsync(b->d_dir, 0, std::max<uint64_t>(b->ndir, 1) * sizeof(EfChunk), c->stream));
    const cudaMemcpyKind kind = mem == IDC_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (b->low_words) IDC_CUDA(cudaMemcpyAsync(b->d_low, low, b->low_words * 8, kind, c->stream));
    if (b->high_words) IDC_CUDA(cudaMemcpyAsync(b->d_high, high, b->high_words * 8, kind, c->stream));
    IDC_TRY(c->status.reserve(128));
    uint32_t* d_status = c->status.as<uint32_t>();
    IDC_CUDA(cudaMemsetAsync(d_status, 0, 4, c->stream));
    if (nl) {
        EfAuxArgs a{b->d_list_off, b->d_l, b->d_low_off, b->d_high_off, b->d_samp_off, b->d_dir_off, b->d_high, b->d_samples, b->d_dir,
                    (uint32_t)nl, d_status};
        LaunchScope ls(c, "k_ef_rebuild_aux");
        k_ef_rebuild_aux<<<(uint32_t)std::min<uint64_t>(nl, (uint64_t)c->sm_count * 16), kThreads, 0, c->stream>>>(a);
    }
    IDC_TRY(check_last_launch("k_ef_rebuild_aux"));
    uint32_t st = 0;
    IDC_CUDA(cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    IDC_REQUIRE(st == 0, IDC_ERR_ARG, "idc_ef_blob_import: an upper-bits vector does not hold as many ones as its list has ids");
    b->device_bytes = acct;
    *out = b.release();
    return IDC_OK;
}

// ---- flat file form (idc_file.h): header words, the list CSR, the universes and the two bit vectors
int idc_ef_blob_save(const idc_ef_blob* b, const char* path) {
    IDC_REQUIRE(b && path, IDC_ERR_ARG, "idc_ef_blob_save: null argument");
    std::vector<uint64_t> low(b->low_words), high(b->high_words), offs(b->nlist + 1);
    IDC_TRY(idc_ef_blob_export(b, offs.data(), nullptr, nullptr, nullptr, nullptr, low.data(), high.data()));  // (fetches the host tables)
    std::vector<uint64_t> hdr{b->nlist, b->low_words, b->high_words, b->row_stride};
    FileWriter w;
    IDC_TRY(w.open(path, kFileEf, 5));
    w.vec(hdr);
    w.vec(b->list_offsets);
    w.vec(b->universe);
    w.vec(low);
    w.vec(high);
    return w.close(path);
}

int idc_ef_blob_load(idc_ctx* c, const char* path, idc_ef_blob** out) {
    IDC_REQUIRE(c && path && out, IDC_ERR_ARG, "idc_ef_blob_load: null argument");
    *out = nullptr;
    FileReader r;
    IDC_TRY(r.open(path, kFileEf));
    std::vector<uint64_t> hdr, offs, uni, low, high;
    IDC_TRY(r.vec(hdr));
    IDC_TRY(r.vec(offs));
    IDC_TRY(r.vec(uni));
    IDC_TRY(r.vec(low));
    IDC_TRY(r.vec(high));
    IDC_REQUIRE(hdr.size() == 4 && offs.size() == hdr[0] + 1 && uni.size() == hdr[0] && low.size() == hdr[1] && high.size() == hdr[2],
                IDC_ERR_ARG, "%s: section sizes do not match the header", path);
    // the sizes the import derives from (universe, length) must be the stored ones: checked before any array is read
    uint64_t lw = 0, hwords = 0;
    for (uint64_t i = 0; i < hdr[0]; i++) {
        IDC_REQUIRE(offs[i + 1] >= offs[i], IDC_ERR_ARG, "%s: list offsets decrease", path);
        const EfShape sh = ef_shape(uni[i], offs[i + 1] - offs[i]);
        lw += sh.low_words;
        hwords += sh.high_words;
    }
    IDC_REQUIRE(lw == hdr[1] && hwords == hdr[2], IDC_ERR_ARG, "%s: bit-vector sizes do not follow from the list shapes", path);
    return idc_ef_blob_import(c, hdr[0], offs.data(), uni.data(), (uint32_t)hdr[3], low.data(), high.data(), IDC_MEM_HOST, out);
}

}  // extern "C"
