// roc_kernels.cu -- ROC (bits-back rANS over sets) encode / decode on sm_100a.
//
// Execution model: ONE UNIT PER GROUP OF G LANES (G = 2, 4 or 8; roc_group.cuh). The reference
// stream is a single serial rANS head per list (codec.cpp:131-137,144-151), so the only
// bit-exact way to interleave coders inside a warp is across lists: a warp runs 32/G units,
// every group owns one unit's head (replicated in its lanes), stream pointer and
// order-statistic workspace, and the group's lanes share the work on that workspace (one
// coalesced 128-byte request per step, ballot searches of the count levels). Units are
// sorted by length so the groups of a warp run in lock step, and the encoder walks them
// END-ALIGNED so that `nmax` (ids left in the set) is the same in every group -- the
// reciprocal used by the uniform pop is then one broadcast load.
//
// Kernels (all integer, HBM/L2-latency bound; no tensor cores):
//   k_unit_meta     per 4096-id tile min/max id, sortedness / width checks; k_unit_meta_finish: precision rule
//   k_row_counts    NSG rows: length = entries before the first -1
//   k_sort_small / k_sort_units   per unit bitonic sort by a warp (<= 64 ids) / a CTA (only for unsorted input)
//   k_enc_records   the unit's ids re-laid as 128-byte records (presence mask + 31 ids)
//   k_roc_encode    the coder (one launch per size class, classes spread over 3 streams)
//   k_roc_compact   gather the per-unit scratch streams into the packed blob
//   k_roc_decode    the decoder (also the row mode: slot-derived tables for NSG rows)
//   k_translate_gather   ids of (list_no << 32 | offset) labels from the decoded hit lists
// Host buffers (IDC_MEM_HOST) are uploaded / downloaded in chunks on a copy stream, overlapped with the kernels.
#include <dlfcn.h>
#include <cuda.h>  // types of the driver entry point cuStreamWaitValue32 only (looked up at run time, not linked)

#include <algorithm>
#include <cstring>
#include <numeric>

#include <functional>
#include <memory>

#include "idc_file.h"
#include "idc_host.h"
#include "idc_prep.cuh"
#include "roc_group.cuh"
#include "roc_small.cuh"

using namespace idc;

// --------------------------------------------------------------------------
// blob
// --------------------------------------------------------------------------
struct idc_roc_blob {
    idc_ctx* ctx = nullptr;
    idc::CtxRef ref;  // declared right after ctx: destroyed last, after the arrays went back to the pool
    uint64_t nlist = 0, nunits = 0, total_ids = 0, total_words = 0, ans_bytes = 0;
    uint32_t max_unit = IDC_MAX_UNIT_DEFAULT, row_stride = 0;
    uint32_t max_n = 0;  // largest unit
    // host metadata (row blobs built by idc_roc_encode_rows are planned on the device: their host tables are fetched
    // on first use, see roc_host_tables)
    mutable bool host_tables = true;
    mutable std::vector<uint64_t> list_offsets;  // nlist+1 (CSR of ids)
    mutable std::vector<uint64_t> unit_offsets;  // nlist+1
    mutable std::vector<uint32_t> unit_n;        // nunits
    mutable std::vector<uint64_t> unit_src;      // nunits: element offset of the unit in the id array
    // device arrays
    uint32_t* d_unit_n = nullptr;
    uint8_t* d_unit_prec = nullptr;
    uint64_t* d_unit_head = nullptr;
    uint64_t* d_word_off = nullptr;  // nunits+1
    uint32_t* d_words = nullptr;
    uint32_t* d_unit_lo = nullptr;   // id range hints for the decoder's bucket map
    uint32_t* d_unit_hi = nullptr;
    uint32_t* d_order = nullptr;     // total_ids (rows: nlist*K), optional
    uint64_t device_bytes = 0;
    // cached decode-everything plan
    bool plan_ready = false;
    uint32_t* d_plan_unit = nullptr;
    uint64_t* d_plan_out = nullptr;
    uint64_t* d_plan_ws = nullptr;
    uint64_t plan_ws_bytes = 0;
    std::vector<uint32_t> plan_ns;   // unit lengths in launch order

    ~idc_roc_blob() {
        if (ctx) ctx->pool_release(d_unit_n);
        if (ctx) ctx->pool_release(d_unit_prec);
        if (ctx) ctx->pool_release(d_unit_head);
        if (ctx) ctx->pool_release(d_word_off);
        if (ctx) ctx->pool_release(d_words);
        if (ctx) ctx->pool_release(d_unit_lo);
        if (ctx) ctx->pool_release(d_unit_hi);
        if (ctx) ctx->pool_release(d_order);
        if (ctx) ctx->pool_release(d_plan_unit);
        if (ctx) ctx->pool_release(d_plan_out);
        if (ctx) ctx->pool_release(d_plan_ws);
    }
};

namespace {

struct EncArgs {
    const void* ids;             // ascending inside each unit
    const uint32_t* sort_idx;    // null when the caller's order was already ascending
    const uint64_t* unit_src;
    const uint32_t* unit_n;
    const uint32_t* unit_posbase;
    const uint8_t* unit_prec;
    const uint32_t* perm;        // launch slot -> unit, by descending n
    const uint64_t* ws_off;      // per unit byte offset into ws
    const uint64_t* scratch_off; // per unit word offset into scratch
    uint8_t* ws;
    uint32_t* scratch;
    uint64_t* unit_head;
    uint32_t* unit_nwords;
    uint32_t* order;             // null: not recorded
    uint32_t* status;
    const uint32_t* mt;
    const uint64_t* rcp64;
    const uint32_t* q31;
    uint32_t nunits;
    uint32_t sm_words;           // shared-memory words per unit (the count levels), a multiple of 4
    uint32_t slot_base;          // this launch covers launch slots [slot_base, slot_end)
    uint32_t slot_end;
    uint32_t rec_base;           // k_enc_records: units [rec_base, rec_base + rec_count)
    uint32_t rec_count;
};

// Re-lay every unit's ascending ids as 128-byte records (mask word + 31 ids, idc_core.cuh): one warp per unit,
// one record per warp iteration (lane w writes word w: coalesced 128-byte stores, coalesced id reads).
template <typename IdT>
__global__ void __launch_bounds__(kThreads) k_enc_records(EncArgs a) {
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.rec_count) return;
    warp += a.rec_base;
    uint32_t n = a.unit_n[warp];
    if (n == 0) return;
    const IdT* src = reinterpret_cast<const IdT*>(a.ids) + a.unit_src[warp];
    uint32_t* rec = reinterpret_cast<uint32_t*>(a.ws + a.ws_off[warp]);
    uint32_t records = enc_tree_layout(n).records;
    for (uint32_t r = 0; r < records; r++) rec[(size_t)r * 32u + lane] = enc_record_word(src, n, r, lane);
}

template <int G, typename IdT>
__global__ void __launch_bounds__(kThreads) k_roc_encode(EncArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr uint32_t NG = 32 / G;  // units per warp
    const Grp<G> g;
    const uint32_t lane_id = threadIdx.x & 31, q = lane_id / G, wic = threadIdx.x >> 5;
    const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + wic;
    const uint32_t slot = a.slot_base + gwarp * NG + q;
    const bool valid = slot < a.slot_end;
    const uint32_t u = valid ? a.perm[slot] : 0u;
    const uint32_t n = valid ? a.unit_n[u] : 0u;
    GEncUnit<G, IdT> U;
#pragma unroll
    for (int j = 0; j < 16 / G; j++) U.tree.ea[j] = 0u;
    U.n = n;
    U.prec = valid ? (int)a.unit_prec[u] : 0;
    const uint64_t src_off = valid ? a.unit_src[u] : 0ull;
    U.sort_idx = a.sort_idx ? a.sort_idx + src_off : nullptr;
    U.order = a.order ? a.order + src_off : nullptr;
    U.pos_base = valid ? a.unit_posbase[u] : 0u;
    U.tree.rec = reinterpret_cast<uint32_t*>(a.ws + (valid ? a.ws_off[u] : 0ull));
    U.tree.sm = SmView{smem + (size_t)wic * (a.sm_words * NG), NG, q};
    if (n) genc_tree_init<G>(g, U.tree, n);
    U.st.head = kRansL;
    U.st.words = a.scratch + (valid ? a.scratch_off[u] : 0ull);
    U.st.sp = 0;
    U.st.cap = n + 4u;
    U.st.draws = 0;
    U.st.status = 0;
    U.st.wr = g.sub == 0 ? 1u : 0u;
    __syncwarp();
    // end-aligned lock step: at warp step t every active group has nmax == t
    const uint32_t tmax = __reduce_max_sync(0xffffffffu, n);
    // Reciprocal tables: every warp of the launch walks the same t at about the same time, so a per-step table
    // load turns one L2 slice into a hot spot (measured: 67 % of all stall samples). Instead each warp fetches 32
    // consecutive entries with one coalesced load per 32 steps (lane j holds entry tb - j), one block ahead, and
    // broadcasts the step's entry with shuffles.
    auto tab_rcp = [&](uint32_t tb) { return tb >= lane_id ? __ldg(a.rcp64 + (tb - lane_id)) : 0ull; };
    auto tab_q31 = [&](uint32_t tb) { return tb >= lane_id ? __ldg(a.q31 + (tb - lane_id)) : 0u; };
    uint32_t tb = tmax;                       // block covers t = tb, tb-1, ..., tb-31
    uint64_t rcp_blk = tab_rcp(tb), rcp_nxt = tb >= 32u ? tab_rcp(tb - 32u) : 0ull;
    uint32_t q31_blk = tab_q31(tb), q31_nxt = tb >= 32u ? tab_q31(tb - 32u) : 0u;
    // the step's table entries are broadcast one step ahead: the shuffles for step t - 1 are issued at the top of
    // step t, so their latency is off the dependent chain
    uint64_t rcp_next = __shfl_sync(0xffffffffu, rcp_blk, 0);
    uint32_t q31_next = __shfl_sync(0xffffffffu, q31_blk, 0);
    for (uint32_t t = tmax; t >= 1u; --t) {
        const uint64_t rcp = rcp_next;
        const uint32_t q31 = q31_next;
        if (t > 1u) {
            const uint32_t tn = t - 1u;
            if (tb - tn == 32u) {
                tb -= 32u;
                rcp_blk = rcp_nxt;
                q31_blk = q31_nxt;
                rcp_nxt = tb >= 32u ? tab_rcp(tb - 32u) : 0ull;
                q31_nxt = tb >= 32u ? tab_q31(tb - 32u) : 0u;
            }
            rcp_next = __shfl_sync(0xffffffffu, rcp_blk, tb - tn);
            q31_next = __shfl_sync(0xffffffffu, q31_blk, tb - tn);
        }
        genc_step<G>(g, U, t, rcp, q31, a.mt, t <= n);
    }
    if (valid && g.sub == 0) {
        a.unit_head[u] = U.st.head;
        a.unit_nwords[u] = U.st.sp;
        if (U.st.status) atomicOr(a.status, U.st.status);
    }
}

// one warp per unit: scratch slot -> packed stream
__global__ void __launch_bounds__(kThreads) k_roc_compact(const uint32_t* scratch, const uint64_t* scratch_off,
                                                          const uint64_t* word_off, uint32_t* words, uint32_t nunits) {
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nunits) return;
    uint64_t d0 = word_off[warp], d1 = word_off[warp + 1];
    const uint32_t* src = scratch + scratch_off[warp];
    uint32_t cnt = (uint32_t)(d1 - d0);
    for (uint32_t i = lane; i < cnt; i += 32) words[d0 + i] = __ldcs(src + i);
}

// nwords[u] = word_off[u + 1] - word_off[u]
__global__ void __launch_bounds__(kThreads) k_unit_nwords(const uint64_t* word_off, uint32_t* nwords, uint64_t nunits) {
    uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < nunits) nwords[u] = (uint32_t)(word_off[u + 1] - word_off[u]);
}

// ids_out[i] = labels[i] < 0 ? labels[i] : decoded[src[i]]
__global__ void __launch_bounds__(kThreads) k_translate_gather(const int64_t* labels, const uint64_t* src,
                                                               const int64_t* decoded, int64_t* out, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t lab = labels[i];
    out[i] = lab < 0 ? lab : decoded[src[i]];
}

struct DecArgs {
    const uint32_t* sel_unit;   // launch slot -> blob unit, by descending n
    const uint64_t* sel_out;    // launch slot -> element offset in out
    const uint64_t* sel_ws;     // launch slot -> byte offset in ws
    const uint32_t* unit_n;
    const uint8_t* unit_prec;
    const uint64_t* unit_head;
    const uint64_t* word_off;
    const uint32_t* words;
    const uint32_t* unit_lo;
    const uint32_t* unit_hi;
    uint8_t* ws;
    void* out;
    uint32_t* counts;           // rows: true neighbour count per slot (may be null)
    uint32_t* status;
    const uint32_t* mt;
    const uint32_t* q31;
    uint32_t nsel;
    uint32_t row_stride;        // rows: pad the slot's output to this many entries with -1
    uint32_t sm_words;          // shared-memory words per unit (all count levels), a multiple of 4
    uint32_t slot_base;         // this launch covers launch slots [slot_base, slot_end)
    uint32_t slot_end;
    // row mode (sel_out == nullptr): slot s decodes row rows[s] (s itself when rows == nullptr) into out + s *
    // row_stride with the workspace slot s * slot_ws -- no per-slot tables, the row numbers may live on the device
    const int32_t* rows;
    uint64_t slot_ws;
    uint32_t nrows;
    uint32_t row_base;          // rows == nullptr: slot s decodes row row_base + s
    // progress milestones (the longest size class of a decode into host memory): every warp adds 1 to progress[m]
    // once all its units have finished (m + 1) * ms_seg steps -- their outputs out[n - (m + 1) * ms_seg, n) are then
    // final and visible device-wide, and the copy stream, which waits for the counter to reach the warp count, can
    // send them to the host while the chains go on
    uint32_t* progress;
    uint32_t ms_seg;
    uint32_t ms_count;
};

template <int G, typename OutT>
__global__ void __launch_bounds__(kThreads) k_roc_decode(DecArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr uint32_t NG = 32 / G;  // units per warp
    const Grp<G> g;
    const uint32_t lane_id = threadIdx.x & 31, q = lane_id / G, wic = threadIdx.x >> 5;
    const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + wic;
    const uint32_t slot = a.slot_base + gwarp * NG + q;
    bool valid = slot < a.slot_end;
    const bool row_mode = a.sel_out == nullptr;
    uint32_t u = 0;
    if (valid) {
        if (!row_mode) {
            u = a.sel_unit[slot];
        } else {
            const int64_t r = a.rows ? (int64_t)a.rows[slot] : (int64_t)a.row_base + slot;
            if (r < 0 || r >= (int64_t)a.nrows) {
                if (g.sub == 0) atomicOr(a.status, kStRange);
                valid = false;
            } else {
                u = (uint32_t)r;
            }
        }
    }
    const uint32_t n = valid ? a.unit_n[u] : 0u;
    GDecUnit<OutT> U;
    U.n = n;
    U.prec = valid ? (int)a.unit_prec[u] : 0;
    U.out = reinterpret_cast<OutT*>(a.out) + (row_mode ? (uint64_t)slot * a.row_stride : (valid ? a.sel_out[slot] : 0ull));
    const uint64_t w0 = valid ? a.word_off[u] : 0ull, w1 = valid ? a.word_off[u + 1] : 0ull;
    // the stream ring sits behind the count levels in the unit's shared-memory region
    const SmView sm{smem + (size_t)wic * (a.sm_words * NG), NG, q};
    dec_state_init(U.st, valid ? a.unit_head[u] : kRansL, a.words + w0, valid ? (uint32_t)(w1 - w0) : 0u,
                   DecRing{sm.at(a.sm_words - kDecRing), NG * 4u}, g.sub == 0);
    if (valid)
        for (uint32_t w = g.sub; w < a.sm_words - kDecRing; w += (uint32_t)G) *sm.at(w) = 0u;
    U.tree = gdec_tree_at(a.ws + (row_mode ? (uint64_t)slot * a.slot_ws : (valid ? a.sel_ws[slot] : 0ull)), sm, n ? n : 1u,
                          valid ? a.unit_lo[u] : 0u, valid ? a.unit_hi[u] : 0u);
    __syncwarp();
    dec_ring_prime(U.st, a.mt);
    dec_pop_start(U.st, a.mt);
    const uint32_t tmax = __reduce_max_sync(0xffffffffu, n);
    // 2^31 / (i + 1) from the table, 32 entries per coalesced load and one block ahead (see k_roc_encode)
    auto tab_q31 = [&](uint32_t ib) { uint32_t e = ib + lane_id + 1u; return __ldg(a.q31 + (e <= kMaxUnit ? e : kMaxUnit)); };
    uint32_t ib = 0;                          // block covers i = ib .. ib+31
    uint32_t q31_blk = tab_q31(0), q31_nxt = tab_q31(32);
    uint32_t q31_next = __shfl_sync(0xffffffffu, q31_blk, 0);  // one step ahead, see k_roc_encode
    gdec_unit_start(U);
    // milestone m is reported once step (m + 1) * ms_seg has RUN: it applied the pending output store of the step before
    uint32_t ms_next = a.progress ? a.ms_seg : 0xffffffffu, ms_done = 0;
    for (uint32_t i = 0; i < tmax; ++i) {
        const uint32_t q31 = q31_next;
        // The table look-ahead for step i + 1: the block refresh (a global load every 32 steps) stays in front of the
        // pops -- behind the bucket request it would wait for it (one scoreboard for all global loads) -- the
        // shuffle runs in the shadow of the request.
        {
            const uint32_t in = i + 1u;
            if (in - ib == 32u) {
                ib += 32u;
                q31_blk = q31_nxt;
                q31_nxt = tab_q31(ib + 32u);
            }
        }
        auto look_ahead = [&]() { q31_next = __shfl_sync(0xffffffffu, q31_blk, i + 1u - ib); };
        gdec_step<G>(g, U, i, q31, a.mt, i < n, look_ahead);
        if (i == ms_next) {  // warp-uniform
            if (ms_done < a.ms_count) {
                __threadfence();
                __syncwarp();
                if (lane_id == 0) atomicAdd(a.progress + ms_done, 1u);
            }
            ms_done++;
            ms_next += a.ms_seg;
        }
    }
    gdec_finish(g, U);
    if (a.progress) {  // a warp of shorter units reports the milestones it never reached: nothing of it is awaited
        __threadfence();
        __syncwarp();
        if (lane_id == 0)
            for (uint32_t m = ms_done; m < a.ms_count; m++) atomicAdd(a.progress + m, 1u);
    }
    if (valid) {
        if (a.row_stride)
            for (uint32_t t = n + g.sub; t < a.row_stride; t += (uint32_t)G) U.out[t] = (OutT)-1;
        if (g.sub == 0) {
            if (a.counts) a.counts[slot] = n;
            uint32_t st = U.st.status & ~kStDegenerate;
            if (st) atomicOr(a.status, st);
        }
    }
}

// ---- short units (graph rows, K <= kSmallUnit): one unit per THREAD (roc_small.cuh). The ids decoded so far are a
// column of shared memory (seen[j][thread], pitch + 1: the scan of a step and the transposed write-out are both free of
// bank conflicts); a warp writes its 32 rows out as full coalesced stores, in the reference's order (codec.cpp:150).
struct SmallDecArgs {
    const uint32_t* unit_n;
    const uint8_t* unit_prec;
    const uint64_t* unit_head;
    const uint64_t* word_off;
    const uint32_t* words;
    const int32_t* rows;   // row numbers (device-addressable; null: slot s decodes row row_base + s)
    uint32_t nrows;
    uint32_t row_base;
    uint32_t nsel;
    uint32_t row_stride;
    void* out;
    uint32_t* counts;
    uint32_t* status;
    const uint32_t* mt;
};

constexpr uint32_t kSmallThreads = 128;

template <typename OutT>
__global__ void __launch_bounds__(kSmallThreads) k_roc_decode_small(SmallDecArgs a) {
    extern __shared__ __align__(16) uint32_t s_seen[];
    constexpr uint32_t kPitch = kSmallThreads + 1u;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wbase = tid & ~31u;
    const uint32_t slot = blockIdx.x * kSmallThreads + tid;
    uint32_t n = 0;
    if (slot < a.nsel) {
        const int64_t r = a.rows ? (int64_t)a.rows[slot] : (int64_t)a.row_base + slot;
        if (r < 0 || r >= (int64_t)a.nrows) {
            atomicOr(a.status, kStRange);  // the row is written as empty (-1 everywhere, count 0)
        } else {
            const uint32_t u = (uint32_t)r;
            n = a.unit_n[u];
            if (n > a.row_stride) {  // not a row of this blob's shape
                atomicOr(a.status, kStRange);
                n = 0;
            }
            if (n) {
                const uint64_t w0 = a.word_off[u], w1 = a.word_off[u + 1];
                SmallDec st;
                small_dec_init(st, a.unit_head[u], a.words + w0, (uint32_t)(w1 - w0));
                small_dec_unit(st, n, (int)a.unit_prec[u], [&](uint32_t j) -> uint32_t& { return s_seen[j * kPitch + tid]; }, a.mt);
                if (st.status) atomicOr(a.status, st.status);
            }
        }
        if (a.counts) a.counts[slot] = n;
    }
    __syncwarp();
    OutT* out = reinterpret_cast<OutT*>(a.out);
    for (uint32_t r = 0; r < 32u; r++) {
        const uint32_t slot_r = blockIdx.x * kSmallThreads + wbase + r;
        if (slot_r >= a.nsel) break;  // warp-uniform
        const uint32_t n_r = __shfl_sync(0xffffffffu, n, r);
        OutT* o = out + (uint64_t)slot_r * a.row_stride;
        for (uint32_t t = lane; t < a.row_stride; t += 32u)
            o[t] = t < n_r ? (OutT)s_seen[(n_r - 1u - t) * kPitch + wbase + r] : (OutT)-1;
    }
}

// rows of at most kSmallUnit ids: the thread-per-row coders (no workspace, one launch). IDC_ROC_ROWS_GROUP=1 keeps the
// lane-group kernels for them (experiments, tests of both paths).
bool small_rows(uint32_t K) {
    return K <= kSmallUnit && getenv("IDC_ROC_ROWS_GROUP") == nullptr;
}

// The encoder twin (roc_small.cuh): one row per thread. The warp first lays its 32 ascending rows into shared-memory
// columns (coalesced reads); a step is pop-uniform, a select on the row's 64-bit presence mask, the push of that id.
// Stream words go to the row's scratch slot like k_roc_encode's (compacted afterwards by the same kernels).
__global__ void __launch_bounds__(kSmallThreads) k_roc_encode_small(EncArgs a) {
    extern __shared__ __align__(16) uint32_t s_ids[];
    constexpr uint32_t kPitch = kSmallThreads + 1u;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wbase = tid & ~31u;
    const uint32_t slot = a.slot_base + blockIdx.x * kSmallThreads + tid;
    const uint32_t* ids = reinterpret_cast<const uint32_t*>(a.ids);
    for (uint32_t r = 0; r < 32u; r++) {
        const uint32_t slot_r = a.slot_base + blockIdx.x * kSmallThreads + wbase + r;
        if (slot_r >= a.slot_end) break;  // warp-uniform
        const uint32_t u_r = a.perm[slot_r];
        const uint32_t n_r = min(a.unit_n[u_r], kSmallUnit);
        const uint32_t* src = ids + a.unit_src[u_r];
        for (uint32_t t = lane; t < n_r; t += 32u) s_ids[t * kPitch + wbase + r] = src[t];
    }
    __syncwarp();
    if (slot >= a.slot_end) return;
    const uint32_t u = a.perm[slot];
    const uint32_t n = a.unit_n[u];
    EncState st{kRansL, a.scratch + a.scratch_off[u], 0u, n + 4u, 0u, 0u, 1u};
    if (n > kSmallUnit) {  // not a unit for this kernel (the host never sends one)
        atomicOr(a.status, kStScratch);
    } else if (n) {
        const int prec = (int)a.unit_prec[u];
        const uint64_t src_off = a.unit_src[u];
        const uint32_t* sidx = a.sort_idx ? a.sort_idx + src_off : nullptr;
        uint32_t* order = a.order ? a.order + src_off : nullptr;
        const uint32_t pos_base = a.unit_posbase[u];
        uint64_t mask = n >= 64u ? ~0ull : ((1ull << n) - 1ull);
        uint64_t rcp = __ldg(a.rcp64 + n);
        uint32_t q31 = __ldg(a.q31 + n);
        for (uint32_t t = n; t >= 1u; --t) {
            // the next step's table entries are requested before this step's chain starts
            const uint64_t rcp_n = __ldg(a.rcp64 + (t - 1u));
            const uint32_t q31_n = __ldg(a.q31 + (t - 1u));
            const uint32_t pos = small_enc_step(st, mask, t, prec, rcp, q31, [&](uint32_t p) { return s_ids[p * kPitch + tid]; }, a.mt);
            if (order) order[n - t] = sidx ? sidx[pos] : pos_base + pos;
            rcp = rcp_n;
            q31 = q31_n;
        }
    }
    a.unit_head[u] = st.head;
    a.unit_nwords[u] = st.sp;
    if (st.status) atomicOr(a.status, st.status);
}

// The latency flavour of the same decoder for a handful of rows (one get_neighbors call, a row and its neighbours'
// rows): one WARP per row. Every lane runs the coder (same inputs, same arithmetic, nothing to exchange); lane j keeps
// the j-th and (j + 32)-th decoded id in two registers, a rank is two ballots. ~10x the issue slots of the
// thread-per-row kernel per row, a third of its latency: used when the rows would not fill the chip anyway.
template <typename OutT>
__global__ void __launch_bounds__(kSmallThreads) k_roc_decode_small_warp(SmallDecArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t slot = blockIdx.x * (kSmallThreads / 32u) + (threadIdx.x >> 5);
    if (slot >= a.nsel) return;  // warp-uniform
    uint32_t n = 0, s0 = 0, s1 = 0;
    const int64_t r = a.rows ? (int64_t)a.rows[slot] : (int64_t)a.row_base + slot;
    if (r < 0 || r >= (int64_t)a.nrows) {
        if (lane == 0) atomicOr(a.status, kStRange);
    } else {
        const uint32_t u = (uint32_t)r;
        n = a.unit_n[u];
        if (n > a.row_stride || n > kSmallUnit) {
            if (lane == 0) atomicOr(a.status, kStRange);
            n = 0;
        }
        if (n) {
            const uint64_t w0 = a.word_off[u], w1 = a.word_off[u + 1];
            const int prec = (int)a.unit_prec[u];
            SmallDec st;
            small_dec_init(st, a.unit_head[u], a.words + w0, (uint32_t)(w1 - w0));
            warp_dec_row(Grp<32>(), st, n, prec, s0, s1, a.mt);
            if (st.status && lane == 0) atomicOr(a.status, st.status);
        }
    }
    if (a.counts && lane == 0) a.counts[slot] = n;
    OutT* o = reinterpret_cast<OutT*>(a.out) + (uint64_t)slot * a.row_stride;
    if (lane < n) o[n - 1u - lane] = (OutT)s0;  // data[n - 1 - i] = i-th decoded id (codec.cpp:150)
    if (lane + 32u < n) o[n - 33u - lane] = (OutT)s1;
    for (uint32_t t = n + lane; t < a.row_stride; t += 32u) o[t] = (OutT)-1;
}

// ---- calls of at most kWarpCallUnits units whose LONGEST unit has at most kWarpUnit ids (IVF indexes with few, short
// lists: IVF256 over 10^5 ids, IVF1024 over 10^6): one unit per WARP, everything on chip. Such a call cannot fill the
// chip with lane-group chains, so the step's latency is what counts: every lane runs the coder (roc_small.cuh: same inputs,
// same arithmetic), the order statistic is cooperative --
//   decode: the ids decoded so far sit in shared memory, a rank is a strided count + one warp reduction;
//   encode: the ascending ids sit in shared memory, lane j keeps the presence mask of words j and j + 32 and the
//           number of ids still present in front of them in registers; select-k-th is two ballots, a shuffle, a
//           select on one mask word -- no memory on the chain but the id itself.
// ~450 cycles per step instead of ~2 300. Calls with longer units keep the lane-group kernels for all their classes.
constexpr uint32_t kWarpUnit = 2048;
constexpr uint32_t kWarpKThreads = 256;

constexpr uint64_t kWarpCallUnits = 4096;  // a warp per unit repeats the coder in 32 lanes: beyond a few units per
                                           // scheduler of the chip the lane-group kernels' shared instruction stream wins
                                           // (IVF65536 over 10^7 ids: 65 536 units of ~150 ids are issue-bound either way)

bool warp_units(uint32_t max_n, uint64_t units) {
    return max_n <= kWarpUnit && units <= kWarpCallUnits && getenv("IDC_ROC_NO_WARP_UNITS") == nullptr;
}

template <typename OutT>
__global__ void __launch_bounds__(kWarpKThreads) k_roc_decode_warp(DecArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t lane = threadIdx.x & 31u, wic = threadIdx.x >> 5;
    const uint32_t slot = a.slot_base + blockIdx.x * (kWarpKThreads / 32u) + wic;
    if (slot >= a.slot_end) return;  // warp-uniform
    const bool row_mode = a.sel_out == nullptr;
    uint32_t u = 0, n = 0;
    bool valid = true;
    if (!row_mode) {
        u = a.sel_unit[slot];
    } else {
        const int64_t r = a.rows ? (int64_t)a.rows[slot] : (int64_t)a.row_base + slot;
        if (r < 0 || r >= (int64_t)a.nrows) {
            if (lane == 0) atomicOr(a.status, kStRange);
            valid = false;
        } else {
            u = (uint32_t)r;
        }
    }
    if (valid) n = a.unit_n[u];
    if (n > a.sm_words) {  // longer than the class was sized for (the host never sends one)
        if (lane == 0) atomicOr(a.status, kStScratch);
        n = 0;
    }
    OutT* out = reinterpret_cast<OutT*>(a.out) + (row_mode ? (uint64_t)slot * a.row_stride : a.sel_out[slot]);
    uint32_t* seen = smem + (size_t)wic * a.sm_words;
    if (n) {
        const uint64_t w0 = a.word_off[u], w1 = a.word_off[u + 1];
        const int prec = (int)a.unit_prec[u];
        SmallDec st;
        small_dec_init(st, a.unit_head[u], a.words + w0, (uint32_t)(w1 - w0));
        warp_dec_unit(Grp<32>(), st, n, prec, seen, a.q31, a.mt);
        if (st.status && lane == 0) atomicOr(a.status, st.status);
        for (uint32_t t = lane; t < n; t += 32u) out[t] = (OutT)seen[n - 1u - t];  // codec.cpp:150
    }
    for (uint32_t t = n + lane; t < a.row_stride; t += 32u) out[t] = (OutT)-1;
    if (a.counts && lane == 0) a.counts[slot] = n;
}

template <typename IdT>
__global__ void __launch_bounds__(kWarpKThreads) k_roc_encode_warp(EncArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t lane = threadIdx.x & 31u, wic = threadIdx.x >> 5;
    const uint32_t slot = a.slot_base + blockIdx.x * (kWarpKThreads / 32u) + wic;
    if (slot >= a.slot_end) return;  // warp-uniform
    const uint32_t u = a.perm[slot];
    uint32_t n = a.unit_n[u];
    EncState st{kRansL, a.scratch + a.scratch_off[u], 0u, n + 4u, 0u, 0u, lane == 0 ? 1u : 0u};
    if (n > a.sm_words) {  // longer than the class was sized for (the host never sends one)
        st.status |= kStScratch;
        n = 0;
    }
    if (n) {
        uint32_t* sid = smem + (size_t)wic * a.sm_words;
        const uint64_t src_off = a.unit_src[u];
        const IdT* src = reinterpret_cast<const IdT*>(a.ids) + src_off;
        for (uint32_t t = lane; t < n; t += 32u) sid[t] = (uint32_t)load_id(src + t);
        __syncwarp();
        const int prec = (int)a.unit_prec[u];
        const uint32_t* sidx = a.sort_idx ? a.sort_idx + src_off : nullptr;
        uint32_t* order = a.order ? a.order + src_off : nullptr;
        const uint32_t pos_base = a.unit_posbase[u];
        warp_enc_unit(Grp<32>(), st, n, prec, sid, a.rcp64, a.q31,
                      [&](uint32_t step, uint32_t pos) {
                          if (order && lane == 0) order[step] = sidx ? sidx[pos] : pos_base + pos;
                      },
                      a.mt);
    }
    if (lane == 0) {
        a.unit_head[u] = st.head;
        a.unit_nwords[u] = st.sp;
        if (st.status) atomicOr(a.status, st.status);
    }
}

// --------------------------------------------------------------- host side

// launch order: units by descending n (counting sort; n <= 65536)
void length_sorted_order(const uint32_t* n, uint64_t count, std::vector<uint32_t>& perm) {
    std::vector<uint64_t> hist(kMaxUnit + 2, 0);
    for (uint64_t i = 0; i < count; i++) hist[kMaxUnit - n[i] + 1]++;
    for (size_t i = 1; i < hist.size(); i++) hist[i] += hist[i - 1];
    perm.resize(count);
    for (uint64_t i = 0; i < count; i++) perm[hist[kMaxUnit - n[i]]++] = (uint32_t)i;
}

// Size classes. Units are launched in descending length; a class is a contiguous slot range whose shared
// memory footprint is set by its first (longest) unit. The longest class bounds the kernel's duration (its units
// are serial chains of up to 65 536 steps). A class's duration is about (its longest unit) x (step latency), so the
// classes are spread over a few (3) streams by longest-processing-time-first on that estimate and run one
// after the other inside a stream: the longest class keeps a stream to itself, the short ones queue up in its
// shadow. Measured: all classes at once slow the long chains by 25 % (they share the issue slots); everything but
// the longest class on ONE stream makes that stream the critical path.
struct SizeClass {
    uint32_t slot_base, slot_end, max_n;
};

constexpr int kMaxClassStreams = 8;

inline int class_stream_count() {
    int n = 3;
    if (const char* e = getenv("IDC_CLASS_STREAMS")) n = atoi(e);  // experiments
    return n < 1 ? 1 : (n > kMaxClassStreams ? kMaxClassStreams : n);
}

// stream of each class: greedy LPT on max_n (classes come in descending max_n)
inline std::vector<int> class_streams(const std::vector<SizeClass>& cls) {
    std::vector<int> out(cls.size(), 0);
    const int kClassStreams = class_stream_count();
    uint64_t load[kMaxClassStreams] = {0};
    for (size_t k = 0; k < cls.size(); k++) {
        int best = 0;
        for (int j = 1; j < kClassStreams; j++)
            if (load[j] < load[best]) best = j;
        out[k] = best;
        load[best] += cls[k].max_n ? cls[k].max_n : 1u;
    }
    return out;
}

template <typename NofSlot>
std::vector<SizeClass> size_classes(uint64_t nslots, NofSlot n_of_slot) {
    static const uint32_t bounds[] = {32768, 16384, 8192, 4096, 2048, 1024, 256, 0};
    std::vector<SizeClass> cls;
    uint64_t s = 0;
    for (uint32_t lo : bounds) {
        if (s >= nslots) break;
        if (n_of_slot(s) <= lo && lo != 0) continue;
        // first slot whose n <= lo (slots are sorted by descending n)
        uint64_t a = s, b = nslots;
        while (a < b) {
            uint64_t m = (a + b) / 2;
            if (n_of_slot(m) > lo) a = m + 1; else b = m;
        }
        if (lo == 0) a = nslots;
        if (a > s) cls.push_back(SizeClass{(uint32_t)s, (uint32_t)a, n_of_slot(s)});
        s = a;
    }
    return cls;
}

// Lanes per unit, per size class.
//   4 for the classes with long units: their serial chains bound the kernel, and 4 lanes give the shortest step
//     (one coalesced line request, 4 count entries per lane);
//   2 for the classes of short units (n <= kG2MaxN): those are many and cheap, what counts there is the issue
//     cost per id -- 16 units per warp need ~30 instead of ~49 warp instructions per id.
// IDC_ROC_G=2|4|8 forces one width for all classes, IDC_ROC_G8_MIN_N=<n> selects 8 lanes for classes whose longest
// unit exceeds n, IDC_ROC_G2_MAX_N=<n> moves the 2-lane threshold (experiments; profiles/README.md has the results).
constexpr uint32_t kG2MaxN = 16384;

inline int group_lanes_for(uint32_t max_n) {
    if (const char* e = getenv("IDC_ROC_G")) {
        int v = atoi(e);
        if (v == 2 || v == 4 || v == 8) return v;
    }
    if (const char* e = getenv("IDC_ROC_G8_MIN_N")) {
        if (max_n > (uint32_t)atoi(e)) return 8;
    }
    uint32_t g2 = kG2MaxN;
    if (const char* e = getenv("IDC_ROC_G2_MAX_N")) g2 = (uint32_t)atoi(e);
    return max_n <= g2 ? 2 : 4;
}

// warps per CTA such that several CTAs fit an SM's 227 KB of shared memory
inline uint32_t warps_for(uint32_t sm_words, uint32_t units_per_warp) {
    size_t per_warp = (size_t)sm_words * 4 * units_per_warp;
    if (per_warp * 4 <= 56 * 1024) return 4;
    if (per_warp * 2 <= 56 * 1024) return 2;
    return 1;
}

template <typename K>
int set_max_smem(K kernel, size_t bytes) {
    IDC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    // All size classes ask for the maximum shared-memory carve-out: an SM keeps one carve-out while a kernel runs
    // on it, and with the driver's per-launch choice a class that needs more shared memory than the running class
    // left over could not become co-resident (measured: decode 146 ms -> 117 ms with the classes overlapping).
    IDC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    return IDC_OK;
}

// Shared encode driver. ids_dev: device pointer to the caller's ids (CSR or
// row-strided); blob has list_offsets / unit tables filled in already.
// ids_host != nullptr: the ids are still on the host and ids_dev is the device staging buffer they go to. With
// ascending input the upload is cut into chunks of whole units on the copy stream; metadata and record kernels
// follow chunk by chunk, and a size class starts as soon as the chunk holding its last unit has arrived (Zipf-length
// lists in CSR order: the longest class, which bounds the kernel, is complete after 72 % of the upload).
int roc_encode_units(idc_ctx* c, idc_roc_blob* b, const void* ids_dev, const void* ids_host, int id_bytes,
                     uint32_t flags, const std::vector<uint32_t>& unit_posbase, uint64_t id_elems) {
    const uint64_t nu = b->nunits;
    uint64_t acct = 0;
    HostTrace tr("roc_encode");
    IDC_TRY(dev_alloc(c, &b->d_unit_n, nu, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_prec, nu, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_head, nu, &acct));
    IDC_TRY(dev_alloc(c, &b->d_word_off, nu + 1, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_lo, nu, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_hi, nu, &acct));
    if (flags & IDC_F_WANT_ORDER) IDC_TRY(dev_alloc(c, &b->d_order, id_elems, &acct));
    IDC_TRY(upload(c, b->d_unit_n, b->unit_n));

    // per-call tables: unit_src, posbase, perm, ws_off, scratch_off, nwords
    std::vector<uint32_t> perm;
    length_sorted_order(b->unit_n.data(), nu, perm);
    std::vector<uint64_t> ws_off(nu), scratch_off(nu);
    uint64_t ws_bytes = 0, scratch_words = 0;
    for (uint64_t u = 0; u < nu; u++) {
        ws_off[u] = ws_bytes;
        ws_bytes += b->unit_n[u] ? enc_tree_bytes(b->unit_n[u]) : 0;
        scratch_off[u] = scratch_words;
        scratch_words += b->unit_n[u] ? (uint64_t)b->unit_n[u] + 4u : 0;
    }
    tr.mark("alloc + order + offsets");
    MetaPlan mplan;
    plan_unit_meta(b->unit_n, mplan);
    const uint64_t ntile = mplan.tile_unit.size();
    IDC_REQUIRE(ntile < (1ull << 32), IDC_ERR_ARG, "too many metadata tiles");
    size_t meta_bytes = nu * (8 + 4 + 4 + 8 + 8 + 4) + ntile * 8 + 256;
    IDC_TRY(c->meta.reserve(meta_bytes + 256));
    uint8_t* mp = c->meta.as<uint8_t>();
    auto carve = [&](size_t bytes) {
        uint8_t* r = mp;
        mp += (bytes + 15) & ~size_t(15);
        return r;
    };
    uint64_t* d_unit_src = (uint64_t*)carve(nu * 8);
    uint64_t* d_ws_off = (uint64_t*)carve(nu * 8);
    uint64_t* d_scratch_off = (uint64_t*)carve(nu * 8);
    uint32_t* d_posbase = (uint32_t*)carve(nu * 4);
    uint32_t* d_perm = (uint32_t*)carve(nu * 4);
    uint32_t* d_nwords = (uint32_t*)carve(nu * 4);
    uint32_t* d_tile_unit = (uint32_t*)carve(ntile * 4);
    uint32_t* d_tile_idx = (uint32_t*)carve(ntile * 4);
    IDC_TRY(c->status.reserve(128));
    uint32_t* d_status = c->status.as<uint32_t>();
    IDC_CUDA(cudaMemsetAsync(d_status, 0, 4, c->stream));
    IDC_TRY(upload(c, d_unit_src, b->unit_src));
    IDC_TRY(upload(c, d_ws_off, ws_off));
    IDC_TRY(upload(c, d_scratch_off, scratch_off));
    IDC_TRY(upload(c, d_posbase, unit_posbase));
    IDC_TRY(upload(c, d_perm, perm));
    IDC_TRY(upload(c, d_tile_unit, mplan.tile_unit));
    IDC_TRY(upload(c, d_tile_idx, mplan.tile_idx));
    if (nu) {
        IDC_CUDA(cudaMemsetAsync(b->d_unit_lo, 0xff, nu * 4, c->stream));
        IDC_CUDA(cudaMemsetAsync(b->d_unit_hi, 0, nu * 4, c->stream));
    }

    tr.mark("meta plan + uploads");
    const bool sorted_in = (flags & IDC_F_SORTED) != 0;
    // chunks of whole units: [cu[j], cu[j+1])
    const bool pipelined = ids_host != nullptr && sorted_in && nu >= 64 && id_elems >= (1ull << 22);
    std::vector<uint64_t> cu{0};
    if (pipelined) {
        const uint64_t first = b->unit_src[0], span = id_elems - first, target = (span + 15) / 16;
        for (uint64_t u = 1; u < nu; u++)
            if (b->unit_src[u] - b->unit_src[cu.back()] >= target) cu.push_back(u);
    }
    cu.push_back(nu);
    const size_t nchunk = cu.size() - 1;
    auto chunk_of_unit = [&](uint64_t u) { return (size_t)(std::upper_bound(cu.begin(), cu.end(), u) - cu.begin()) - 1; };
    cudaStream_t copy_s = nullptr;
    if (pipelined) IDC_TRY(c->copy_stream_get(&copy_s));
    if (ids_host && !pipelined && id_elems)
        IDC_CUDA(cudaMemcpyAsync(const_cast<void*>(ids_dev), ids_host, id_elems * id_bytes, cudaMemcpyHostToDevice, c->stream));

    // workspaces (the sort, when needed, works on a 32-bit copy of the ids)
    const void* enc_ids = ids_dev;
    int enc_id_bytes = id_bytes;
    uint32_t* d_sort_idx = nullptr;
    size_t ws_need = ws_bytes;
    size_t sorted_off = 0, sortidx_off = 0, big_off = 0;
    uint32_t sort_grid = 0;
    if (!sorted_in) {
        sort_grid = (uint32_t)std::min<uint64_t>(nu, (uint64_t)c->sm_count * 8);
        bool need_big = false;
        for (uint64_t u = 0; u < nu && !need_big; u++) need_big = b->unit_n[u] > kSortSmem;
        sorted_off = (ws_need + 255) & ~size_t(255);
        sortidx_off = sorted_off + ((id_elems * 4 + 255) & ~size_t(255));
        big_off = sortidx_off + ((id_elems * 4 + 255) & ~size_t(255));
        ws_need = big_off + (need_big ? (size_t)sort_grid * kMaxUnit * 8 : 0);
    }
    IDC_TRY(c->ws.reserve(ws_need + 256));
    IDC_TRY(c->scratch.reserve(scratch_words * 4 + 256));

    EncArgs e{};
    e.ids = enc_ids;
    e.unit_src = d_unit_src;
    e.unit_n = b->d_unit_n;
    e.unit_posbase = d_posbase;
    e.unit_prec = b->d_unit_prec;
    e.perm = d_perm;
    e.ws_off = d_ws_off;
    e.scratch_off = d_scratch_off;
    e.ws = c->ws.as<uint8_t>();
    e.scratch = c->scratch.as<uint32_t>();
    e.unit_head = b->d_unit_head;
    e.unit_nwords = d_nwords;
    e.order = b->d_order;
    e.status = d_status;
    e.mt = c->d_mt;
    e.rcp64 = c->d_rcp64;
    e.q31 = c->d_q31;
    e.nunits = (uint32_t)nu;
    uint32_t max_n = 0;
    for (uint64_t u = 0; u < nu; u++) max_n = std::max(max_n, b->unit_n[u]);
    b->max_n = max_n;

    // the class streams wait for what has been queued so far (tables, memsets), not for the chunks
    auto cls = size_classes(nu, [&](uint64_t slot) { return b->unit_n[perm[slot]]; });
    const std::vector<int> cstream = class_streams(cls);
    // Host input arriving in chunks: the longest class, whose chains bound the call, is launched in up to four pieces
    // that start as their own units arrive (below); the extra pieces get streams of their own.
    constexpr int kEncPieces = 4;
    const int nstreams = class_stream_count() + (pipelined ? kEncPieces - 1 : 0);
    IDC_TRY(c->fork(nstreams));

    // 1. per chunk: upload, unit metadata, [sort], records
    MetaArgs m{ids_dev, d_unit_src, b->d_unit_n, (uint32_t)nu, sorted_in ? 1u : 0u,
               (flags & IDC_F_PRECISION_SAFE) ? 1u : 0u, b->d_unit_prec, b->d_unit_lo, b->d_unit_hi, d_status,
               mplan.identity ? nullptr : d_tile_unit, mplan.identity ? nullptr : d_tile_idx, 0u, 0u, 0u, 0u};
    std::vector<cudaEvent_t> ev_ready(nchunk);
    for (size_t j = 0; j < nchunk; j++) {
        const uint64_t u0 = cu[j], u1 = cu[j + 1];
        if (pipelined) {
            const uint64_t e0 = j == 0 ? 0 : b->unit_src[u0], e1 = u1 < nu ? b->unit_src[u1] : id_elems;
            cudaEvent_t ev;
            IDC_TRY(c->sync_event(&ev));
            IDC_CUDA(cudaMemcpyAsync((uint8_t*)const_cast<void*>(ids_dev) + e0 * id_bytes, (const uint8_t*)ids_host + e0 * id_bytes,
                                     (e1 - e0) * id_bytes, cudaMemcpyHostToDevice, copy_s));
            IDC_CUDA(cudaEventRecord(ev, copy_s));
            IDC_CUDA(cudaStreamWaitEvent(c->stream, ev, 0));
        }
        IDC_TRY(launch_unit_meta(c, m, id_bytes, u0, u1, mplan));
        if (!pipelined) {
            // input errors surface before the heavy work (the pipelined path checks at the end: the kernels are
            // memory-safe on unsorted or too wide ids, their output is then discarded)
            uint32_t st = 0;
            IDC_CUDA(cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, c->stream));
            IDC_CUDA(cudaStreamSynchronize(c->stream));
            IDC_TRY(status_to_error(st, "roc_encode"));
        }
        if (!sorted_in) {
            uint32_t* d_sorted = (uint32_t*)(c->ws.as<uint8_t>() + sorted_off);
            d_sort_idx = (uint32_t*)(c->ws.as<uint8_t>() + sortidx_off);
            SortArgs s{ids_dev, d_unit_src, b->d_unit_n, d_posbase, (uint32_t)nu, d_sorted, d_sort_idx,
                       (uint64_t*)(c->ws.as<uint8_t>() + big_off)};
            bool any_big = false;
            for (uint64_t u = 0; u < nu && !any_big; u++) any_big = b->unit_n[u] > kSortWarp;
            IDC_TRY(launch_sorts(c, s, id_bytes, sort_grid, any_big));
            enc_ids = d_sorted;
            enc_id_bytes = 4;
            e.ids = enc_ids;
            IDC_TRY(check_last_launch("k_sort_units"));
        }
        e.sort_idx = d_sort_idx;
        {
            EncArgs er = e;
            er.rec_base = (uint32_t)u0;
            er.rec_count = (uint32_t)(u1 - u0);
            LaunchScope ls(c, "k_enc_records");
            if (er.rec_count) {
                if (enc_id_bytes == 8)
                    k_enc_records<int64_t><<<grid_for((uint64_t)er.rec_count * 32), kThreads, 0, c->stream>>>(er);
                else
                    k_enc_records<uint32_t><<<grid_for((uint64_t)er.rec_count * 32), kThreads, 0, c->stream>>>(er);
            }
        }
        IDC_TRY(check_last_launch("k_enc_records"));
        IDC_TRY(c->sync_event(&ev_ready[j]));
        IDC_CUDA(cudaEventRecord(ev_ready[j], c->stream));
    }

    tr.mark("chunks: meta, sort, records");
    // (IDC_TRACE_HOST: device-side time line of the size classes, relative to the first thing queued behind the tables)
    std::vector<cudaEvent_t> tl_ev;
    size_t tl_n = 0;
    if (tr.on) {
        tl_ev.resize(1 + 2 * (cls.size() + kEncPieces));
        for (auto& ev : tl_ev) cudaEventCreate(&ev);
        cudaEventRecord(tl_ev[0], c->aux[cstream.empty() ? 0 : cstream[0]]);  // (the class streams were forked before the chunks were queued)
    }
    // 3. encode: a class starts when the chunk with its last unit is ready. The longest class does not wait for ITS
    // last unit as a whole: its slots (descending length) are cut where the latest chunk needed so far changes, the
    // last three such values get a launch of their own -- with Zipf-length lists in CSR order the full-length units
    // (66 % of the upload) start 10 ms before the last unit of the class (72 %) has arrived, and their 65 536-step
    // chains are what the call waits for.
    struct Piece {
        uint32_t slot_base, slot_end, cls;
        int stream;
        size_t chunk;
    };
    std::vector<Piece> pieces;
    for (size_t k = 0; k < cls.size(); k++) {
        std::vector<size_t> need(cls[k].slot_end - cls[k].slot_base);  // running maximum of the chunk a slot's unit is in
        size_t run = 0;
        for (uint32_t sl = cls[k].slot_base; sl < cls[k].slot_end; sl++) {
            run = std::max(run, nchunk ? chunk_of_unit(perm[sl]) : (size_t)0);
            need[sl - cls[k].slot_base] = run;
        }
        std::vector<uint32_t> cuts{cls[k].slot_base};
        if (pipelined && k == 0 && nchunk > 1) {
            std::vector<uint32_t> steps;  // slots where the running maximum grows
            for (uint32_t i = 1; i < need.size(); i++)
                if (need[i] != need[i - 1]) steps.push_back(cls[k].slot_base + i);
            const size_t keep = std::min<size_t>(steps.size(), (size_t)kEncPieces - 1);
            cuts.insert(cuts.end(), steps.end() - keep, steps.end());
        }
        cuts.push_back(cls[k].slot_end);
        for (size_t i = 0; i + 1 < cuts.size(); i++)
            pieces.push_back(Piece{cuts[i], cuts[i + 1], (uint32_t)k, i == 0 ? cstream[k] : class_stream_count() + (int)i - 1,
                                   need[cuts[i + 1] - 1 - cls[k].slot_base]});
    }
    {
        LaunchScope ls(c, "k_roc_encode");
        for (const Piece& pc : pieces) {
            const size_t k = pc.cls;
            EncArgs ek = e;
            ek.slot_base = pc.slot_base;
            ek.slot_end = pc.slot_end;
            ek.sm_words = genc_sm_words(cls[k].max_n ? cls[k].max_n : 1u);
            const int G = group_lanes_for(cls[k].max_n);
            const uint32_t upw = 32u / (uint32_t)G;  // units per warp
            const uint32_t warps = warps_for(ek.sm_words, upw), threads = warps * 32;
            const uint32_t slots = ek.slot_end - ek.slot_base, nwarps = (slots + upw - 1) / upw;
            const uint32_t grid = (nwarps + warps - 1) / warps;
            const size_t smem = (size_t)ek.sm_words * 4 * upw * warps;
            cudaStream_t ps = c->aux[pc.stream];
            IDC_CUDA(cudaStreamWaitEvent(ps, ev_ready[nchunk ? pc.chunk : 0], 0));
            if (tr.on) cudaEventRecord(tl_ev[1 + 2 * tl_n], ps);
            if (warp_units(max_n, nu)) {
                // every unit of the call short: one unit per warp, on chip (k_roc_encode_warp)
                ek.sm_words = ((cls[k].max_n ? cls[k].max_n : 1u) + 31u) & ~31u;
                const uint32_t wpb = kWarpKThreads / 32u, wgrid = (slots + wpb - 1) / wpb;
                const size_t wsmem = (size_t)ek.sm_words * 4 * wpb;
                if (enc_id_bytes == 8) {
                    IDC_TRY(set_max_smem(k_roc_encode_warp<int64_t>, wsmem));
                    k_roc_encode_warp<int64_t><<<wgrid, kWarpKThreads, wsmem, ps>>>(ek);
                } else {
                    IDC_TRY(set_max_smem(k_roc_encode_warp<uint32_t>, wsmem));
                    k_roc_encode_warp<uint32_t><<<wgrid, kWarpKThreads, wsmem, ps>>>(ek);
                }
                if (tr.on) cudaEventRecord(tl_ev[2 + 2 * tl_n], ps), tl_n++;
                c->launches++;
                continue;
            }
#define IDC_LAUNCH_ENC(GG, TT)                                                       \
    do {                                                                             \
        IDC_TRY(set_max_smem(k_roc_encode<GG, TT>, smem));                            \
        k_roc_encode<GG, TT><<<grid, threads, smem, ps>>>(ek);                        \
    } while (0)
            if (G == 8) {
                if (enc_id_bytes == 8) IDC_LAUNCH_ENC(8, int64_t); else IDC_LAUNCH_ENC(8, uint32_t);
            } else if (G == 2) {
                if (enc_id_bytes == 8) IDC_LAUNCH_ENC(2, int64_t); else IDC_LAUNCH_ENC(2, uint32_t);
            } else {
                if (enc_id_bytes == 8) IDC_LAUNCH_ENC(4, int64_t); else IDC_LAUNCH_ENC(4, uint32_t);
            }
#undef IDC_LAUNCH_ENC
            if (tr.on) cudaEventRecord(tl_ev[2 + 2 * tl_n], ps), tl_n++;
            c->launches++;
        }
        c->launches--;  // LaunchScope counted one already
        IDC_TRY(c->join(nstreams));
    }
    IDC_TRY(check_last_launch("k_roc_encode"));

    tr.mark("class launches");
    // 4. sizes -> packed offsets -> compaction
    std::vector<uint32_t> nwords(nu);
    uint32_t st = 0;
    IDC_CUDA(cudaMemcpyAsync(nwords.data(), d_nwords, nu * 4, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    IDC_TRY(status_to_error(st, "roc_encode"));
    tr.mark("wait for the kernels");
    if (tr.on) {
        for (size_t k = 0; k < tl_n; k++) {
            float a = 0, z = 0;
            cudaEventElapsedTime(&a, tl_ev[0], tl_ev[1 + 2 * k]);
            cudaEventElapsedTime(&z, tl_ev[0], tl_ev[2 + 2 * k]);
            fprintf(stderr, "[idc host]   launch %zu (class %u, slots %u..%u, longest unit %u, stream %d, chunk %zu): starts %.1f ms, ends %.1f ms\n", k,
                    pieces[k].cls, pieces[k].slot_base, pieces[k].slot_end, b->unit_n[perm[pieces[k].slot_base]], pieces[k].stream,
                    pieces[k].chunk, a, z);
        }
        for (auto& ev : tl_ev) cudaEventDestroy(ev);
    }
    std::vector<uint64_t> word_off(nu + 1);
    word_off[0] = 0;
    uint64_t ans_bytes = 0;
    for (uint64_t u = 0; u < nu; u++) {
        if (b->unit_n[u] == 0) nwords[u] = 0;
        word_off[u + 1] = word_off[u] + nwords[u];
        if (b->unit_n[u]) ans_bytes += 8 + 4ull * nwords[u];  // ANSState::size(), codec.h:42-44
    }
    b->total_words = word_off[nu];
    b->ans_bytes = ans_bytes;
    IDC_TRY(dev_alloc(c, &b->d_words, b->total_words, &acct));
    IDC_TRY(upload(c, b->d_word_off, word_off));
    {
        LaunchScope ls(c, "k_roc_compact");
        k_roc_compact<<<grid_for(nu * 32), kThreads, 0, c->stream>>>(c->scratch.as<uint32_t>(), d_scratch_off,
                                                                       b->d_word_off, b->d_words, (uint32_t)nu);
    }
    IDC_TRY(check_last_launch("k_roc_compact"));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    tr.mark("word offsets + compaction");
    b->device_bytes = acct;
    return IDC_OK;
}

// list -> unit tables for CSR input
// ---- graph rows, planned on the device -------------------------------------------------------------------
// A million rows of <= K ids each: no host loop may touch them. Row r is unit r and launch slot r, it starts at
// element r K, owns a fixed-size workspace and scratch slot (what a full row needs), and the packed word offsets come
// from a prefix sum over the emitted word counts.
struct RowTabArgs {
    uint64_t* unit_src;
    uint64_t* ws_off;
    uint64_t* scratch_off;
    uint32_t* posbase;
    uint32_t* perm;
    uint64_t nrows;
    uint64_t K, slot_ws, slot_scratch;
};

__global__ void __launch_bounds__(kThreads) k_roc_row_tables(RowTabArgs a) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.nrows) return;
    a.unit_src[r] = r * a.K;
    a.ws_off[r] = r * a.slot_ws;
    a.scratch_off[r] = r * a.slot_scratch;
    a.posbase[r] = 0u;
    a.perm[r] = (uint32_t)r;
}

// per row: stream words (0 for an empty row), "row is not empty", ids -- the inputs of the three prefix sums
__global__ void __launch_bounds__(kThreads) k_roc_row_sizes(const uint32_t* unit_n, const uint32_t* nwords, uint64_t nrows, uint64_t* sizes) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const uint32_t n = unit_n[r];
    sizes[r] = n ? nwords[r] : 0u;
    sizes[nrows + r] = n ? 1u : 0u;
    sizes[2 * nrows + r] = n;
}

// host tables of a row blob that was planned on the device
int roc_host_tables(const idc_roc_blob* b) {
    if (b->host_tables) return IDC_OK;
    idc_ctx* c = b->ctx;
    const uint64_t nr = b->nlist;
    b->unit_n.resize(nr);
    if (nr) IDC_CUDA(cudaMemcpyAsync(b->unit_n.data(), b->d_unit_n, nr * 4, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    b->list_offsets.resize(nr + 1);
    b->unit_offsets.resize(nr + 1);
    b->unit_src.resize(nr);
    uint64_t total = 0;
    for (uint64_t r = 0; r < nr; r++) {
        b->list_offsets[r] = total;
        b->unit_offsets[r] = r;
        b->unit_src[r] = r * b->row_stride;
        total += b->unit_n[r];
    }
    b->list_offsets[nr] = total;
    b->unit_offsets[nr] = nr;
    b->host_tables = true;
    return IDC_OK;
}

int roc_encode_rows_device(idc_ctx* c, idc_roc_blob* b, const int32_t* d_data, uint32_t K, uint32_t flags) {
    const uint64_t nr = b->nlist, elems = nr * K;
    HostTrace tr("roc_encode_rows");
    uint64_t acct = 0;
    IDC_TRY(dev_alloc(c, &b->d_unit_n, nr, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_prec, nr, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_head, nr, &acct));
    IDC_TRY(dev_alloc(c, &b->d_word_off, nr + 1, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_lo, nr, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_hi, nr, &acct));
    if (flags & IDC_F_WANT_ORDER) IDC_TRY(dev_alloc(c, &b->d_order, elems, &acct));
    b->max_n = K;
    b->host_tables = false;
    const uint64_t slot_ws = enc_tree_bytes(K), slot_scratch = (uint64_t)K + 4u;
    const size_t scan_bytes = scan_scratch_bytes(nr, 3);
    IDC_TRY(c->meta.reserve(nr * (8 + 8 + 8 + 4 + 4 + 4 + 3 * 8 + 2 * 8) + scan_bytes + 1024));
    uint8_t* mp = c->meta.as<uint8_t>();
    auto carve = [&](size_t bytes) {
        uint8_t* r = mp;
        mp += (bytes + 15) & ~size_t(15);
        return r;
    };
    uint64_t* d_unit_src = (uint64_t*)carve(nr * 8);
    uint64_t* d_ws_off = (uint64_t*)carve(nr * 8);
    uint64_t* d_scratch_off = (uint64_t*)carve(nr * 8);
    uint64_t* d_sizes = (uint64_t*)carve(nr * 3 * 8);
    uint64_t* d_sums = (uint64_t*)carve((nr + 1) * 2 * 8);  // prefix sums nobody keeps: "not empty", ids
    uint64_t* d_scan = (uint64_t*)carve(scan_bytes);
    uint32_t* d_posbase = (uint32_t*)carve(nr * 4);
    uint32_t* d_perm = (uint32_t*)carve(nr * 4);
    uint32_t* d_nwords = (uint32_t*)carve(nr * 4);
    IDC_TRY(c->status.reserve(128));
    uint32_t* d_status = c->status.as<uint32_t>();
    IDC_CUDA(cudaMemsetAsync(d_status, 0, 4, c->stream));
    // workspaces: records, then the 32-bit sorted copy of the rows and the sort permutation
    const bool need_big = K > kSortSmem, any_big = K > kSortWarp;
    const uint32_t sort_grid = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(nr, 1), (uint64_t)c->sm_count * 8);
    const size_t sorted_off = (nr * slot_ws + 255) & ~size_t(255);
    const size_t sortidx_off = sorted_off + ((elems * 4 + 255) & ~size_t(255));
    const size_t big_off = sortidx_off + ((elems * 4 + 255) & ~size_t(255));
    IDC_TRY(c->ws.reserve(big_off + (need_big ? (size_t)sort_grid * kMaxUnit * 8 : 0) + 256));
    IDC_TRY(c->scratch.reserve(nr * slot_scratch * 4 + 256));
    tr.mark("alloc");
    if (nr) {
        {
            LaunchScope ls(c, "k_row_counts");
            k_row_counts<<<grid_for(nr * 32), kThreads, 0, c->stream>>>(d_data, nr, K, b->d_unit_n);
        }
        IDC_TRY(check_last_launch("k_row_counts"));
        {
            LaunchScope ls(c, "k_roc_row_tables");
            RowTabArgs t{d_unit_src, d_ws_off, d_scratch_off, d_posbase, d_perm, nr, K, slot_ws, slot_scratch};
            k_roc_row_tables<<<grid_for(nr), kThreads, 0, c->stream>>>(t);
        }
        IDC_TRY(check_last_launch("k_roc_row_tables"));
        MetaArgs m{d_data, d_unit_src, b->d_unit_n, (uint32_t)nr, 0u, (flags & IDC_F_PRECISION_SAFE) ? 1u : 0u, b->d_unit_prec,
                   b->d_unit_lo, b->d_unit_hi, d_status, nullptr, nullptr, 0u, 0u, 0u, 0u};
        IDC_TRY(run_unit_meta_small(c, m, nr, 4));
        uint32_t* d_sorted = (uint32_t*)(c->ws.as<uint8_t>() + sorted_off);
        uint32_t* d_sort_idx = (uint32_t*)(c->ws.as<uint8_t>() + sortidx_off);
        SortArgs s{d_data, d_unit_src, b->d_unit_n, d_posbase, (uint32_t)nr, d_sorted, d_sort_idx, (uint64_t*)(c->ws.as<uint8_t>() + big_off)};
        IDC_TRY(launch_sorts(c, s, 4, sort_grid, any_big));
        IDC_TRY(check_last_launch("k_sort_units"));
        EncArgs e{};
        e.ids = d_sorted;
        e.sort_idx = d_sort_idx;
        e.unit_src = d_unit_src;
        e.unit_n = b->d_unit_n;
        e.unit_posbase = d_posbase;
        e.unit_prec = b->d_unit_prec;
        e.perm = d_perm;
        e.ws_off = d_ws_off;
        e.scratch_off = d_scratch_off;
        e.ws = c->ws.as<uint8_t>();
        e.scratch = c->scratch.as<uint32_t>();
        e.unit_head = b->d_unit_head;
        e.unit_nwords = d_nwords;
        e.order = b->d_order;
        e.status = d_status;
        e.mt = c->d_mt;
        e.rcp64 = c->d_rcp64;
        e.q31 = c->d_q31;
        e.nunits = (uint32_t)nr;
        e.rec_base = 0;
        e.rec_count = (uint32_t)nr;
        if (small_rows(K)) {
            // rows of at most 64 ids: one row per thread, no record workspace (roc_small.cuh)
            e.slot_base = 0;
            e.slot_end = (uint32_t)nr;
            LaunchScope ls(c, "k_roc_encode_small");
            k_roc_encode_small<<<(uint32_t)((nr + kSmallThreads - 1) / kSmallThreads), kSmallThreads,
                                 (size_t)K * (kSmallThreads + 1u) * 4u, c->stream>>>(e);
        } else {
            {
                LaunchScope ls(c, "k_enc_records");
                k_enc_records<uint32_t><<<grid_for(nr * 32), kThreads, 0, c->stream>>>(e);
            }
            IDC_TRY(check_last_launch("k_enc_records"));
            // one size class: every slot is sized for a full row
            e.slot_base = 0;
            e.slot_end = (uint32_t)nr;
            e.sm_words = genc_sm_words(K);
            const int G = group_lanes_for(K);
            const uint32_t upw = 32u / (uint32_t)G;
            const uint32_t warps = warps_for(e.sm_words, upw), threads = warps * 32;
            const uint32_t nwarps = (uint32_t)((nr + upw - 1) / upw), grid = (nwarps + warps - 1) / warps;
            const size_t smem = (size_t)e.sm_words * 4 * upw * warps;
            LaunchScope ls(c, "k_roc_encode");
            if (G == 8) {
                IDC_TRY(set_max_smem(k_roc_encode<8, uint32_t>, smem));
                k_roc_encode<8, uint32_t><<<grid, threads, smem, c->stream>>>(e);
            } else if (G == 2) {
                IDC_TRY(set_max_smem(k_roc_encode<2, uint32_t>, smem));
                k_roc_encode<2, uint32_t><<<grid, threads, smem, c->stream>>>(e);
            } else {
                IDC_TRY(set_max_smem(k_roc_encode<4, uint32_t>, smem));
                k_roc_encode<4, uint32_t><<<grid, threads, smem, c->stream>>>(e);
            }
        }
        IDC_TRY(check_last_launch("k_roc_encode"));
        {
            LaunchScope ls(c, "k_roc_row_sizes");
            k_roc_row_sizes<<<grid_for(nr), kThreads, 0, c->stream>>>(b->d_unit_n, d_nwords, nr, d_sizes);
        }
        IDC_TRY(check_last_launch("k_roc_row_sizes"));
    }
    {
        const uint64_t* sin[3] = {d_sizes, d_sizes + nr, d_sizes + 2 * nr};
        uint64_t* sout[3] = {b->d_word_off, d_sums, d_sums + nr + 1};
        IDC_TRY(device_scan(c, 3, sin, sout, nr, d_scan));
    }
    uint64_t tot[3] = {0, 0, 0};
    uint32_t st = 0;
    IDC_CUDA(cudaMemcpyAsync(&tot[0], b->d_word_off + nr, 8, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaMemcpyAsync(&tot[1], d_sums + nr, 8, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaMemcpyAsync(&tot[2], d_sums + 2 * nr + 1, 8, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    IDC_TRY(status_to_error(st, "roc_encode_rows"));
    tr.mark("kernels");
    b->total_words = tot[0];
    b->ans_bytes = 8 * tot[1] + 4 * tot[0];  // sum over non-empty rows of ANSState::size(), codec.h:42-44
    b->total_ids = tot[2];
    IDC_TRY(dev_alloc(c, &b->d_words, b->total_words, &acct));
    if (nr) {
        LaunchScope ls(c, "k_roc_compact");
        k_roc_compact<<<grid_for(nr * 32), kThreads, 0, c->stream>>>(c->scratch.as<uint32_t>(), d_scratch_off, b->d_word_off, b->d_words,
                                                                       (uint32_t)nr);
    }
    IDC_TRY(check_last_launch("k_roc_compact"));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    tr.mark("compaction");
    b->device_bytes = acct;
    return IDC_OK;
}

int plan_units_csr(idc_roc_blob* b, uint64_t nlist, const uint64_t* offsets, uint32_t max_unit,
                   std::vector<uint32_t>& posbase) {
    b->nlist = nlist;
    b->max_unit = max_unit;
    b->list_offsets.assign(offsets, offsets + nlist + 1);
    b->unit_offsets.resize(nlist + 1);
    b->unit_n.clear();
    b->unit_src.clear();
    posbase.clear();
    for (uint64_t l = 0; l < nlist; l++) {
        IDC_REQUIRE(offsets[l + 1] >= offsets[l], IDC_ERR_ARG, "offsets must be non-decreasing (list %llu)",
                    (unsigned long long)l);
        b->unit_offsets[l] = b->unit_n.size();
        uint64_t n = offsets[l + 1] - offsets[l];
        if (n == 0) {  // an empty list still owns one (empty) unit so that list <-> unit stays total
            b->unit_n.push_back(0);
            b->unit_src.push_back(offsets[l]);
            posbase.push_back(0);
            continue;
        }
        for (uint64_t s = 0; s < n; s += max_unit) {
            b->unit_n.push_back((uint32_t)std::min<uint64_t>(max_unit, n - s));
            b->unit_src.push_back(offsets[l] + s);
            posbase.push_back((uint32_t)s);
        }
    }
    b->unit_offsets[nlist] = b->unit_n.size();
    b->nunits = b->unit_n.size();
    b->total_ids = offsets[nlist] - offsets[0];
    IDC_REQUIRE(b->nunits < (1ull << 32), IDC_ERR_ARG, "too many units");
    return IDC_OK;
}

int build_decode_plan(idc_ctx* c, const idc_roc_blob* b, const std::vector<uint32_t>& units,
                      const std::vector<uint64_t>& out_off, uint32_t** d_unit, uint64_t** d_out, uint64_t** d_ws,
                      uint64_t* ws_bytes, std::vector<uint32_t>* ns_sorted) {
    const uint64_t m = units.size();
    std::vector<uint32_t> ns(m), perm;
    for (uint64_t i = 0; i < m; i++) ns[i] = b->unit_n[units[i]];
    length_sorted_order(ns.data(), m, perm);
    std::vector<uint32_t> su(m);
    std::vector<uint64_t> so(m), sw(m);
    uint64_t wsb = 0;
    for (uint64_t i = 0; i < m; i++) {
        su[i] = units[perm[i]];
        so[i] = out_off[perm[i]];
        sw[i] = wsb;
        wsb += ns[perm[i]] ? dec_tree_bytes(ns[perm[i]]) : 0;
    }
    IDC_TRY(dev_alloc(c, d_unit, m));
    IDC_TRY(dev_alloc(c, d_out, m));
    IDC_TRY(dev_alloc(c, d_ws, m));
    IDC_TRY(upload(c, *d_unit, su));
    IDC_TRY(upload(c, *d_out, so));
    IDC_TRY(upload(c, *d_ws, sw));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    *ws_bytes = wsb;
    ns_sorted->resize(m);
    for (uint64_t i = 0; i < m; i++) (*ns_sorted)[i] = ns[perm[i]];
    return IDC_OK;
}

// For a caller that overlaps the device->host copy of the output with the kernels: one event per size class,
// recorded when that class is done, and the class's shortest unit (classes are ranges of unit lengths).
struct DecodeOverlap {
    std::vector<cudaEvent_t> class_done;
    std::vector<uint32_t> class_min_n;
    std::vector<uint64_t> class_est;  // estimated completion in steps: the classes before it on its stream + its own longest unit
    // milestones of the longest class (see DecArgs::progress); ms_count == 0: none
    bool want_ms = false;
    uint32_t ms_seg = 0, ms_count = 0, ms_expect = 0;
    uint32_t* d_progress = nullptr;
    cudaEvent_t ms_armed = nullptr;  // the counters are zero once this event has passed
};

// cuStreamWaitValue32 through the runtime's driver entry point lookup (the library does not link libcuda)
typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamWaitValue32Fn stream_wait_value32() {
    static StreamWaitValue32Fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<StreamWaitValue32Fn>(p);
    }();
    return fn;
}

// cudaMemcpyBatchAsync (CUDA 12.8): many copies in one call -- looked up at run time, an older runtime simply goes
// without the milestones
typedef cudaError_t (*MemcpyBatchFn)(void**, void**, size_t*, size_t, cudaMemcpyAttributes*, size_t*, size_t, size_t*, cudaStream_t);
MemcpyBatchFn memcpy_batch() {
    static MemcpyBatchFn fn = reinterpret_cast<MemcpyBatchFn>(dlsym(RTLD_DEFAULT, "cudaMemcpyBatchAsync"));
    return fn;
}

inline uint32_t ms_parts() {
    if (const char* e = getenv("IDC_MS_PARTS")) {  // experiments
        int v = atoi(e);
        if (v >= 2 && v <= 16) return (uint32_t)v;
    }
    return 8;  // the longest class's output leaves in this many pieces per unit
}
constexpr uint32_t kMsMinUnit = 16384;    // ... when its units are at least this long
constexpr uint32_t kMsWordOffset = 8;     // the counters live behind the status word (words 8 .. 23 of the status buffer)

int finish_decode(idc_ctx* c) {
    uint32_t st = 0;
    IDC_CUDA(cudaMemcpyAsync(&st, c->status.as<uint32_t>(), 4, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    return status_to_error(st, "roc_decode");
}

// ov == nullptr: launches, waits and checks. ov != nullptr: launches only; the caller runs finish_decode().
int run_decode(idc_ctx* c, const idc_roc_blob* b, const uint32_t* d_unit, const uint64_t* d_out, const uint64_t* d_ws,
               uint64_t ws_bytes, uint64_t nsel, void* out_dev, int id_bytes, uint32_t* counts_dev, uint32_t row_stride,
               uint32_t max_n, const std::function<uint32_t(uint64_t)>& n_of_slot, DecodeOverlap* ov = nullptr,
               const int32_t* rows_dev = nullptr, uint64_t slot_ws = 0, uint64_t row_base = 0) {
    if (nsel == 0) {
        if (ov) {
            IDC_TRY(c->status.reserve(128));
            IDC_CUDA(cudaMemsetAsync(c->status.p, 0, 4, c->stream));
        }
        return IDC_OK;
    }
    auto cls = size_classes(nsel, n_of_slot);
    // every unit of the call short: one unit per warp, on chip (k_roc_decode_warp) -- no bucket workspace
    const bool warp_mode = warp_units(cls.empty() ? 0u : cls[0].max_n, nsel);
    if (!warp_mode) IDC_TRY(c->ws.reserve(ws_bytes + 256));
    IDC_TRY(c->status.reserve(128));
    uint32_t* d_status = c->status.as<uint32_t>();
    IDC_CUDA(cudaMemsetAsync(d_status, 0, 128, c->stream));
    if (ov && ov->want_ms) {
        IDC_TRY(c->sync_event(&ov->ms_armed));
        IDC_CUDA(cudaEventRecord(ov->ms_armed, c->stream));
    }
    if (!warp_mode) {
        LaunchScope ls(c, "memset_ws");  // empty bucket slots must read as 0xffffffff (dec_tree_insert_rank)
        IDC_CUDA(cudaMemsetAsync(c->ws.p, 0xff, ws_bytes, c->stream));
    }
    DecArgs a{};
    a.sel_unit = d_unit;
    a.sel_out = d_out;
    a.sel_ws = d_ws;
    a.unit_n = b->d_unit_n;
    a.unit_prec = b->d_unit_prec;
    a.unit_head = b->d_unit_head;
    a.word_off = b->d_word_off;
    a.words = b->d_words;
    a.unit_lo = b->d_unit_lo;
    a.unit_hi = b->d_unit_hi;
    a.ws = c->ws.as<uint8_t>();
    a.out = out_dev;
    a.counts = counts_dev;
    a.status = d_status;
    a.mt = c->d_mt;
    a.q31 = c->d_q31;
    a.nsel = (uint32_t)nsel;
    a.row_stride = row_stride;
    a.rows = rows_dev;
    a.slot_ws = slot_ws;
    a.nrows = (uint32_t)b->nlist;
    a.row_base = (uint32_t)row_base;
    {
        (void)max_n;
        LaunchScope ls(c, "k_roc_decode");  // (the logical kernel: k_roc_decode_warp in warp mode)
        const std::vector<int> cstream = class_streams(cls);
        IDC_TRY(c->fork(class_stream_count()));
        uint64_t stream_load[kMaxClassStreams] = {0};
        // Launch order. Kernel alone: descending length (the LPT order the streams were balanced for). With the
        // output leaving for the host as it is produced (ov): the longest class first, then the SHORTEST classes --
        // inside a stream the order does not change when the stream is done, but the short lists' ids are final
        // within the first milliseconds and keep the copy engine busy until the long chains deliver.
        std::vector<size_t> korder(cls.size());
        std::iota(korder.begin(), korder.end(), (size_t)0);
        if (ov && cls.size() > 2) std::reverse(korder.begin() + 1, korder.end());
        if (ov) {
            ov->class_done.assign(cls.size(), nullptr);
            ov->class_min_n.assign(cls.size(), 0u);
            ov->class_est.assign(cls.size(), 0ull);
        }
        for (size_t k : korder) {
            DecArgs ak = a;
            ak.slot_base = cls[k].slot_base;
            ak.slot_end = cls[k].slot_end;
            ak.sm_words = dec_tree_sm_words(cls[k].max_n ? cls[k].max_n : 1u) + kDecRing;
            const int G = group_lanes_for(cls[k].max_n);
            const uint32_t upw = 32u / (uint32_t)G;
            const uint32_t warps = warps_for(ak.sm_words, upw), threads = warps * 32;
            const uint32_t slots = ak.slot_end - ak.slot_base, nwarps = (slots + upw - 1) / upw;
            const uint32_t grid = (nwarps + warps - 1) / warps;
            const size_t smem = (size_t)ak.sm_words * 4 * upw * warps;
            const bool ms = ov && ov->want_ms && k == 0 && cls[k].max_n >= kMsMinUnit;
            if (warp_mode) {
                ak.sm_words = ((cls[k].max_n ? cls[k].max_n : 1u) + 31u) & ~31u;
                const uint32_t wpb = kWarpKThreads / 32u, wslots = ak.slot_end - ak.slot_base;
                const uint32_t wgrid = (wslots + wpb - 1) / wpb;
                const size_t wsmem = (size_t)ak.sm_words * 4 * wpb;
                if (id_bytes == 8) {
                    IDC_TRY(set_max_smem(k_roc_decode_warp<int64_t>, wsmem));
                    k_roc_decode_warp<int64_t><<<wgrid, kWarpKThreads, wsmem, c->aux[cstream[k]]>>>(ak);
                } else {
                    IDC_TRY(set_max_smem(k_roc_decode_warp<int32_t>, wsmem));
                    k_roc_decode_warp<int32_t><<<wgrid, kWarpKThreads, wsmem, c->aux[cstream[k]]>>>(ak);
                }
            } else {
            if (ms) {
                ov->ms_seg = (cls[k].max_n + ms_parts() - 1) / ms_parts();
                ov->ms_count = ms_parts() - 1;
                ov->ms_expect = grid * warps;
                ov->d_progress = d_status + kMsWordOffset;
                ak.progress = ov->d_progress;
                ak.ms_seg = ov->ms_seg;
                ak.ms_count = ov->ms_count;
            }
#define IDC_LAUNCH_DEC(GG, TT)                                                       \
    do {                                                                             \
        IDC_TRY(set_max_smem(k_roc_decode<GG, TT>, smem));                            \
        k_roc_decode<GG, TT><<<grid, threads, smem, c->aux[cstream[k]]>>>(ak);                 \
    } while (0)
            if (G == 8) {
                if (id_bytes == 8) IDC_LAUNCH_DEC(8, int64_t); else IDC_LAUNCH_DEC(8, int32_t);
            } else if (G == 2) {
                if (id_bytes == 8) IDC_LAUNCH_DEC(2, int64_t); else IDC_LAUNCH_DEC(2, int32_t);
            } else {
                if (id_bytes == 8) IDC_LAUNCH_DEC(4, int64_t); else IDC_LAUNCH_DEC(4, int32_t);
            }
#undef IDC_LAUNCH_DEC
            }
            c->launches++;
            stream_load[cstream[k]] += cls[k].max_n ? cls[k].max_n : 1u;
            if (ov) {
                cudaEvent_t ev;
                IDC_TRY(c->sync_event(&ev));
                IDC_CUDA(cudaEventRecord(ev, c->aux[cstream[k]]));
                ov->class_done[k] = ev;
                ov->class_min_n[k] = n_of_slot(cls[k].slot_end - 1);
                ov->class_est[k] = stream_load[cstream[k]];
                // insurance: whatever happened inside the kernel, a wait on its counters ends with it
                if (ms) IDC_CUDA(cudaMemsetAsync(ov->d_progress, 0x40, 4 * ov->ms_count, c->aux[cstream[k]]));
            }
        }
        c->launches--;
        IDC_TRY(c->join(class_stream_count()));
    }
    IDC_TRY(check_last_launch("k_roc_decode"));
    if (ov) return IDC_OK;
    return finish_decode(c);
}

int run_decode_small(idc_ctx* c, const idc_roc_blob* b, const int32_t* rows_dev, uint64_t row_base, uint64_t nsel, int32_t* out_dev,
                     uint32_t* counts_dev, uint32_t K) {
    if (nsel == 0) return IDC_OK;
    IDC_TRY(c->status.reserve(128));
    uint32_t* d_status = c->status.as<uint32_t>();
    IDC_CUDA(cudaMemsetAsync(d_status, 0, 4, c->stream));
    SmallDecArgs a{b->d_unit_n, b->d_unit_prec, b->d_unit_head, b->d_word_off, b->d_words, rows_dev, (uint32_t)b->nlist,
                   (uint32_t)row_base, (uint32_t)nsel, K, out_dev, counts_dev, d_status, c->d_mt};
    const size_t smem = (size_t)K * (kSmallThreads + 1u) * 4u;  // <= 33 KB
    // up to a warp per scheduler of the chip the rows' chains run side by side whichever way: take the shorter chain
    static const uint64_t warp_rows = getenv("IDC_ROC_ROWS_WARP") ? (uint64_t)atoll(getenv("IDC_ROC_ROWS_WARP")) : 1024u;
    if (nsel <= warp_rows) {
        LaunchScope ls(c, "k_roc_decode_small_warp");
        const uint32_t wpb = kSmallThreads / 32u;
        k_roc_decode_small_warp<int32_t><<<(uint32_t)((nsel + wpb - 1) / wpb), kSmallThreads, 0, c->stream>>>(a);
    } else {
        LaunchScope ls(c, "k_roc_decode_small");
        k_roc_decode_small<int32_t><<<(uint32_t)((nsel + kSmallThreads - 1) / kSmallThreads), kSmallThreads, smem, c->stream>>>(a);
    }
    IDC_TRY(check_last_launch("k_roc_decode_small"));
    return finish_decode(c);
}

}  // namespace

// --------------------------------------------------------------------------
// C ABI
// --------------------------------------------------------------------------
extern "C" {

int idc_roc_encode(idc_ctx* c, uint64_t nlist, const uint64_t* offsets, const void* ids, int id_bytes, int ids_mem,
                   uint32_t flags, uint32_t max_unit, idc_roc_blob** out) {
    IDC_REQUIRE(c && offsets && out, IDC_ERR_ARG, "idc_roc_encode: null argument");
    IDC_REQUIRE(id_bytes == 8 || id_bytes == 4, IDC_ERR_ARG, "id_bytes must be 4 or 8");
    if (max_unit == 0) max_unit = IDC_MAX_UNIT_DEFAULT;
    IDC_REQUIRE(max_unit <= kMaxUnit, IDC_ERR_ARG,
                "max_unit %u > 65536: the reference codec does not round-trip larger sets", max_unit);
    *out = nullptr;
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    std::unique_ptr<idc_roc_blob> b(new idc_roc_blob());
    b->ctx = c;
    b->ref.bind(c);
    std::vector<uint32_t> posbase;
    IDC_TRY(plan_units_csr(b.get(), nlist, offsets, max_unit, posbase));
    uint64_t first = offsets[0], elems = offsets[nlist];
    IDC_REQUIRE(ids != nullptr || elems == 0, IDC_ERR_ARG, "ids is NULL");
    const void* ids_dev = ids;
    const void* ids_host = nullptr;
    if (ids_mem == IDC_MEM_HOST && elems) {
        IDC_TRY(c->stage.reserve(elems * id_bytes));
        ids_dev = c->stage.p;
        ids_host = ids;  // uploaded inside, overlapped with the kernels
    }
    (void)first;
    IDC_TRY(roc_encode_units(c, b.get(), ids_dev, ids_host, id_bytes, flags, posbase, elems));
    *out = b.release();
    return IDC_OK;
}

int idc_roc_encode_rows(idc_ctx* c, uint64_t nrows, uint32_t K, const int32_t* data, int data_mem, uint32_t flags,
                        idc_roc_blob** out) {
    IDC_REQUIRE(c && out && (data || nrows == 0), IDC_ERR_ARG, "idc_roc_encode_rows: null argument");
    IDC_REQUIRE(K >= 1 && K <= kMaxUnit, IDC_ERR_ARG, "K out of range");
    IDC_REQUIRE(nrows < (1ull << 32), IDC_ERR_ARG, "too many rows");
    *out = nullptr;
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    std::unique_ptr<idc_roc_blob> b(new idc_roc_blob());
    b->ctx = c;
    b->ref.bind(c);
    HostTrace tr("roc_encode_rows");
    b->row_stride = K;
    b->nlist = nrows;
    b->nunits = nrows;
    b->max_unit = kMaxUnit;
    const int32_t* d_data = data;
    uint64_t elems = nrows * K;
    if (data_mem == IDC_MEM_HOST && elems) {
        IDC_TRY(c->stage.reserve(elems * 4));
        IDC_CUDA(cudaMemcpyAsync(c->stage.p, data, elems * 4, cudaMemcpyHostToDevice, c->stream));
        d_data = c->stage.as<int32_t>();
    }
    // lengths, tables, sort, coder and word offsets: all on the device (roc_encode_rows_device)
    IDC_TRY(roc_encode_rows_device(c, b.get(), d_data, K, flags & ~IDC_F_SORTED));
    *out = b.release();
    return IDC_OK;
}

int idc_roc_blob_info(const idc_roc_blob* b, idc_roc_info* info) {
    IDC_REQUIRE(b && info, IDC_ERR_ARG, "null argument");
    info->nlist = b->nlist;
    info->nunits = b->nunits;
    info->total_ids = b->total_ids;
    info->total_words = b->total_words;
    info->ans_bytes = b->ans_bytes;
    info->device_bytes = b->device_bytes;
    info->max_unit = b->max_unit;
    info->row_stride = b->row_stride;
    return IDC_OK;
}

int idc_roc_blob_export(const idc_roc_blob* b, uint64_t* list_offsets, uint64_t* unit_offsets, uint32_t* unit_n,
                        uint8_t* unit_precision, uint64_t* unit_heads, uint64_t* word_offsets, uint32_t* words) {
    IDC_REQUIRE(b, IDC_ERR_ARG, "null blob");
    IDC_CUDA(cudaSetDevice(b->ctx->device));
    cudaStream_t s = b->ctx->stream;
    if (!b->host_tables) {
        std::lock_guard<std::mutex> lock(b->ctx->mu);
        IDC_TRY(roc_host_tables(b));
    }
    if (list_offsets) memcpy(list_offsets, b->list_offsets.data(), (b->nlist + 1) * 8);
    if (unit_offsets) memcpy(unit_offsets, b->unit_offsets.data(), (b->nlist + 1) * 8);
    if (unit_n && b->nunits) memcpy(unit_n, b->unit_n.data(), b->nunits * 4);
    if (unit_precision && b->nunits)
        IDC_CUDA(cudaMemcpyAsync(unit_precision, b->d_unit_prec, b->nunits, cudaMemcpyDeviceToHost, s));
    if (unit_heads && b->nunits)
        IDC_CUDA(cudaMemcpyAsync(unit_heads, b->d_unit_head, b->nunits * 8, cudaMemcpyDeviceToHost, s));
    if (word_offsets)
        IDC_CUDA(cudaMemcpyAsync(word_offsets, b->d_word_off, (b->nunits + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (words && b->total_words)
        IDC_CUDA(cudaMemcpyAsync(words, b->d_words, b->total_words * 4, cudaMemcpyDeviceToHost, s));
    IDC_CUDA(cudaStreamSynchronize(s));
    return IDC_OK;
}

int idc_roc_blob_import(idc_ctx* c, uint64_t nlist, const uint32_t* unit_n, const uint8_t* unit_precision,
                        const uint64_t* unit_heads, const uint64_t* word_offsets, const uint32_t* words,
                        idc_roc_blob** out) {
    IDC_REQUIRE(c && out && (nlist == 0 || (unit_n && unit_precision && unit_heads)) && word_offsets, IDC_ERR_ARG,
                "idc_roc_blob_import: null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    *out = nullptr;
    std::unique_ptr<idc_roc_blob> b(new idc_roc_blob());
    b->ctx = c;
    b->ref.bind(c);
    b->nlist = b->nunits = nlist;
    b->unit_n.assign(unit_n, unit_n + nlist);
    b->list_offsets.resize(nlist + 1);
    b->unit_offsets.resize(nlist + 1);
    b->unit_src.resize(nlist);
    uint64_t total = 0, ans = 0;
    std::vector<uint32_t> lo(nlist, 0), hi(nlist);
    for (uint64_t l = 0; l < nlist; l++) {
        IDC_REQUIRE(unit_n[l] <= kMaxUnit, IDC_ERR_DOMAIN, "unit %llu has %u ids (> 65536)", (unsigned long long)l,
                    unit_n[l]);
        IDC_REQUIRE(unit_precision[l] <= 32, IDC_ERR_DOMAIN, "unit %llu: precision %u > 32 not supported on device",
                    (unsigned long long)l, unit_precision[l]);
        IDC_REQUIRE(word_offsets[l + 1] >= word_offsets[l], IDC_ERR_ARG, "word_offsets must be non-decreasing");
        b->list_offsets[l] = total;
        b->unit_offsets[l] = l;
        b->unit_src[l] = total;
        total += unit_n[l];
        if (unit_n[l]) ans += 8 + 4 * (word_offsets[l + 1] - word_offsets[l]);
        hi[l] = unit_precision[l] >= 32 ? 0xffffffffu : ((1u << unit_precision[l]) - 1u);
    }
    b->list_offsets[nlist] = total;
    b->unit_offsets[nlist] = nlist;
    b->total_ids = total;
    for (uint64_t l = 0; l < nlist; l++) b->max_n = std::max(b->max_n, unit_n[l]);
    b->total_words = word_offsets[nlist] - word_offsets[0];
    b->ans_bytes = ans;
    uint64_t acct = 0;
    IDC_TRY(dev_alloc(c, &b->d_unit_n, nlist, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_prec, nlist, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_head, nlist, &acct));
    IDC_TRY(dev_alloc(c, &b->d_word_off, nlist + 1, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_lo, nlist, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_hi, nlist, &acct));
    IDC_TRY(dev_alloc(c, &b->d_words, b->total_words, &acct));
    std::vector<uint64_t> woff(nlist + 1);
    for (uint64_t l = 0; l <= nlist; l++) woff[l] = word_offsets[l] - word_offsets[0];
    cudaStream_t s = c->stream;
    if (nlist) {
        IDC_CUDA(cudaMemcpyAsync(b->d_unit_n, unit_n, nlist * 4, cudaMemcpyHostToDevice, s));
        IDC_CUDA(cudaMemcpyAsync(b->d_unit_prec, unit_precision, nlist, cudaMemcpyHostToDevice, s));
        IDC_CUDA(cudaMemcpyAsync(b->d_unit_head, unit_heads, nlist * 8, cudaMemcpyHostToDevice, s));
        IDC_CUDA(cudaMemcpyAsync(b->d_unit_lo, lo.data(), nlist * 4, cudaMemcpyHostToDevice, s));
        IDC_CUDA(cudaMemcpyAsync(b->d_unit_hi, hi.data(), nlist * 4, cudaMemcpyHostToDevice, s));
    }
    IDC_CUDA(cudaMemcpyAsync(b->d_word_off, woff.data(), (nlist + 1) * 8, cudaMemcpyHostToDevice, s));
    if (b->total_words)
        IDC_CUDA(cudaMemcpyAsync(b->d_words, words + word_offsets[0], b->total_words * 4, cudaMemcpyHostToDevice, s));
    IDC_CUDA(cudaStreamSynchronize(s));
    b->device_bytes = acct;
    *out = b.release();
    return IDC_OK;
}

int idc_roc_blob_export_payload(const idc_roc_blob* b, int mem, uint8_t* unit_precision, uint64_t* unit_heads,
                                uint32_t* unit_nwords, uint32_t* unit_lo, uint32_t* unit_hi, uint32_t* words) {
    IDC_REQUIRE(b, IDC_ERR_ARG, "null blob");
    IDC_REQUIRE(mem == IDC_MEM_HOST || mem == IDC_MEM_DEVICE, IDC_ERR_ARG, "mem must be IDC_MEM_HOST or IDC_MEM_DEVICE");
    idc_ctx* c = b->ctx;
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const cudaMemcpyKind kind = mem == IDC_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    const uint64_t nu = b->nunits;
    if (unit_precision && nu) IDC_CUDA(cudaMemcpyAsync(unit_precision, b->d_unit_prec, nu, kind, s));
    if (unit_heads && nu) IDC_CUDA(cudaMemcpyAsync(unit_heads, b->d_unit_head, nu * 8, kind, s));
    if (unit_lo && nu) IDC_CUDA(cudaMemcpyAsync(unit_lo, b->d_unit_lo, nu * 4, kind, s));
    if (unit_hi && nu) IDC_CUDA(cudaMemcpyAsync(unit_hi, b->d_unit_hi, nu * 4, kind, s));
    if (unit_nwords && nu) {
        uint32_t* d_nw = unit_nwords;
        if (mem == IDC_MEM_HOST) {
            IDC_TRY(c->meta.reserve(nu * 4 + 256));
            d_nw = c->meta.as<uint32_t>();
        }
        k_unit_nwords<<<grid_for(nu), kThreads, 0, s>>>(b->d_word_off, d_nw, nu);
        c->launches++;
        IDC_TRY(check_last_launch("k_unit_nwords"));
        if (mem == IDC_MEM_HOST) IDC_CUDA(cudaMemcpyAsync(unit_nwords, d_nw, nu * 4, cudaMemcpyDeviceToHost, s));
    }
    if (words && b->total_words) IDC_CUDA(cudaMemcpyAsync(words, b->d_words, b->total_words * 4, kind, s));
    IDC_CUDA(cudaStreamSynchronize(s));
    return IDC_OK;
}

int idc_roc_blob_assemble(idc_ctx* c, uint64_t nlist, const uint64_t* list_offsets, uint32_t max_unit, int mem,
                          const uint8_t* unit_precision, const uint64_t* unit_heads, const uint32_t* unit_nwords,
                          const uint32_t* unit_lo, const uint32_t* unit_hi, const uint32_t* words, uint64_t total_words,
                          idc_roc_blob** out) {
    IDC_REQUIRE(c && out && list_offsets, IDC_ERR_ARG, "idc_roc_blob_assemble: null argument");
    IDC_REQUIRE(mem == IDC_MEM_HOST || mem == IDC_MEM_DEVICE, IDC_ERR_ARG, "mem must be IDC_MEM_HOST or IDC_MEM_DEVICE");
    if (max_unit == 0) max_unit = IDC_MAX_UNIT_DEFAULT;
    IDC_REQUIRE(max_unit <= kMaxUnit, IDC_ERR_ARG, "max_unit %u > 65536", max_unit);
    *out = nullptr;
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    std::unique_ptr<idc_roc_blob> b(new idc_roc_blob());
    b->ctx = c;
    b->ref.bind(c);
    std::vector<uint32_t> posbase;
    IDC_TRY(plan_units_csr(b.get(), nlist, list_offsets, max_unit, posbase));  // the very split idc_roc_encode makes
    const uint64_t nu = b->nunits;
    IDC_REQUIRE(nu == 0 || (unit_precision && unit_heads && unit_nwords), IDC_ERR_ARG, "idc_roc_blob_assemble: null unit array");
    IDC_REQUIRE(total_words == 0 || words, IDC_ERR_ARG, "idc_roc_blob_assemble: words is NULL");
    for (uint64_t u = 0; u < nu; u++) b->max_n = std::max(b->max_n, b->unit_n[u]);
    uint64_t acct = 0;
    IDC_TRY(dev_alloc(c, &b->d_unit_n, nu, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_prec, nu, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_head, nu, &acct));
    IDC_TRY(dev_alloc(c, &b->d_word_off, nu + 1, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_lo, nu, &acct));
    IDC_TRY(dev_alloc(c, &b->d_unit_hi, nu, &acct));
    IDC_TRY(dev_alloc(c, &b->d_words, total_words, &acct));
    cudaStream_t s = c->stream;
    const cudaMemcpyKind kind = mem == IDC_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    IDC_TRY(upload(c, b->d_unit_n, b->unit_n));
    std::vector<uint32_t> nw(nu);
    std::vector<uint8_t> prec(nu);
    if (nu) {
        IDC_CUDA(cudaMemcpyAsync(b->d_unit_prec, unit_precision, nu, kind, s));
        IDC_CUDA(cudaMemcpyAsync(b->d_unit_head, unit_heads, nu * 8, kind, s));
        // word counts and precisions are metadata: a host copy feeds the offsets and the checks below
        IDC_CUDA(cudaMemcpyAsync(nw.data(), unit_nwords, nu * 4, mem == IDC_MEM_HOST ? cudaMemcpyHostToHost : cudaMemcpyDeviceToHost, s));
        IDC_CUDA(cudaMemcpyAsync(prec.data(), unit_precision, nu, mem == IDC_MEM_HOST ? cudaMemcpyHostToHost : cudaMemcpyDeviceToHost, s));
    }
    if (total_words) IDC_CUDA(cudaMemcpyAsync(b->d_words, words, total_words * 4, kind, s));
    IDC_CUDA(cudaStreamSynchronize(s));
    std::vector<uint64_t> woff(nu + 1, 0);
    uint64_t ans = 0;
    for (uint64_t u = 0; u < nu; u++) {
        IDC_REQUIRE(prec[u] <= 32, IDC_ERR_DOMAIN, "unit %llu: precision %u > 32 not supported on device", (unsigned long long)u, prec[u]);
        if (b->unit_n[u] == 0) nw[u] = 0;
        woff[u + 1] = woff[u] + nw[u];
        if (b->unit_n[u]) ans += 8 + 4ull * nw[u];
    }
    IDC_REQUIRE(woff[nu] == total_words, IDC_ERR_ARG, "idc_roc_blob_assemble: unit word counts sum to %llu, total_words is %llu",
                (unsigned long long)woff[nu], (unsigned long long)total_words);
    b->total_words = total_words;
    b->ans_bytes = ans;
    IDC_TRY(upload(c, b->d_word_off, woff));
    if (nu) {
        if (unit_lo && unit_hi) {
            IDC_CUDA(cudaMemcpyAsync(b->d_unit_lo, unit_lo, nu * 4, kind, s));
            IDC_CUDA(cudaMemcpyAsync(b->d_unit_hi, unit_hi, nu * 4, kind, s));
        } else {  // no hints: every id below 2^precision (only the decoder's speed depends on them)
            std::vector<uint32_t> lo(nu, 0), hi(nu);
            for (uint64_t u = 0; u < nu; u++) hi[u] = prec[u] >= 32 ? 0xffffffffu : ((1u << prec[u]) - 1u);
            IDC_TRY(upload(c, b->d_unit_lo, lo));
            IDC_TRY(upload(c, b->d_unit_hi, hi));
            IDC_CUDA(cudaStreamSynchronize(s));
        }
    }
    IDC_CUDA(cudaStreamSynchronize(s));
    b->device_bytes = acct;
    *out = b.release();
    return IDC_OK;
}

int idc_roc_blob_order(const idc_roc_blob* b, uint32_t* order, int order_mem) {
    IDC_REQUIRE(b && order, IDC_ERR_ARG, "null argument");
    IDC_REQUIRE(b->d_order != nullptr, IDC_ERR_ARG, "blob was encoded without IDC_F_WANT_ORDER");
    IDC_CUDA(cudaSetDevice(b->ctx->device));
    // d_order is indexed like the caller's id array (absolute element offsets); the export is rebased to the
    // first list, like every decode output: order[(offsets[l] - offsets[0]) + t]
    const uint64_t first = b->row_stride ? 0 : b->list_offsets[0];
    uint64_t elems = b->row_stride ? b->nlist * b->row_stride : b->total_ids;
    IDC_CUDA(cudaMemcpyAsync(order, b->d_order + first, elems * 4,
                             order_mem == IDC_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice,
                             b->ctx->stream));
    IDC_CUDA(cudaStreamSynchronize(b->ctx->stream));
    return IDC_OK;
}

int idc_roc_blob_free(idc_roc_blob* b) {
    if (b) {
        std::lock_guard<std::mutex> lock(b->ctx->mu);
        cudaSetDevice(b->ctx->device);
        delete b;
    }
    return IDC_OK;
}

int idc_roc_decode(idc_ctx* c, const idc_roc_blob* b, const uint64_t* list_nos, uint64_t nsel, void* ids_out,
                   int id_bytes, int out_mem, uint64_t* out_offsets) {
    IDC_REQUIRE(c && b, IDC_ERR_ARG, "idc_roc_decode: null argument");
    IDC_REQUIRE(id_bytes == 8 || id_bytes == 4, IDC_ERR_ARG, "id_bytes must be 4 or 8");
    IDC_REQUIRE(b->row_stride == 0, IDC_ERR_ARG, "row blob: use idc_roc_decode_rows");
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    idc_roc_blob* mb = const_cast<idc_roc_blob*>(b);
    const uint32_t* d_unit;
    const uint64_t *d_out, *d_ws;
    uint64_t ws_bytes, nunits_sel, total_out;
    uint32_t *t_unit = nullptr;
    uint64_t *t_out = nullptr, *t_ws = nullptr;
    std::vector<uint32_t> t_ns;
    {  // argument errors surface before any plan buffer is allocated
        uint64_t want = 0;
        if (list_nos == nullptr) want = b->total_ids;
        else
            for (uint64_t i = 0; i < nsel; i++) {
                IDC_REQUIRE(list_nos[i] < b->nlist, IDC_ERR_ARG, "list_no %llu out of range", (unsigned long long)list_nos[i]);
                want += b->list_offsets[list_nos[i] + 1] - b->list_offsets[list_nos[i]];
            }
        IDC_REQUIRE(want == 0 || ids_out != nullptr, IDC_ERR_ARG, "ids_out is NULL");
    }
    if (list_nos == nullptr) {
        if (!mb->plan_ready) {
            std::vector<uint32_t> units(b->nunits);
            std::iota(units.begin(), units.end(), 0u);
            std::vector<uint64_t> out_off(b->nunits);
            for (uint64_t u = 0; u < b->nunits; u++) out_off[u] = b->unit_src[u] - b->list_offsets[0];
            IDC_TRY(build_decode_plan(c, b, units, out_off, &mb->d_plan_unit, &mb->d_plan_out, &mb->d_plan_ws,
                                      &mb->plan_ws_bytes, &mb->plan_ns));
            mb->plan_ready = true;
        }
        d_unit = b->d_plan_unit;
        d_out = b->d_plan_out;
        d_ws = b->d_plan_ws;
        ws_bytes = b->plan_ws_bytes;
        nunits_sel = b->nunits;
        total_out = b->total_ids;
        if (out_offsets)
            for (uint64_t l = 0; l <= b->nlist; l++) out_offsets[l] = b->list_offsets[l] - b->list_offsets[0];
    } else {
        std::vector<uint32_t> units;
        std::vector<uint64_t> out_off;
        uint64_t pos = 0;
        for (uint64_t i = 0; i < nsel; i++) {
            uint64_t l = list_nos[i];
            IDC_REQUIRE(l < b->nlist, IDC_ERR_ARG, "list_no %llu out of range", (unsigned long long)l);
            if (out_offsets) out_offsets[i] = pos;
            for (uint64_t u = b->unit_offsets[l]; u < b->unit_offsets[l + 1]; u++) {
                units.push_back((uint32_t)u);
                out_off.push_back(pos);
                pos += b->unit_n[u];
            }
        }
        if (out_offsets) out_offsets[nsel] = pos;
        total_out = pos;
        nunits_sel = units.size();
        IDC_TRY(build_decode_plan(c, b, units, out_off, &t_unit, &t_out, &t_ws, &ws_bytes, &t_ns));
        d_unit = t_unit;
        d_out = t_out;
        d_ws = t_ws;
    }
    int rc = IDC_OK;
    if (total_out) {
        IDC_REQUIRE(ids_out != nullptr, IDC_ERR_ARG, "ids_out is NULL");
        void* out_dev = ids_out;
        if (out_mem == IDC_MEM_HOST) {
            rc = c->stage.reserve(total_out * id_bytes);
            out_dev = c->stage.p;
        }
        const std::vector<uint32_t>& ns = list_nos == nullptr ? b->plan_ns : t_ns;
        const bool overlap = out_mem == IDC_MEM_HOST && list_nos == nullptr && total_out >= (1ull << 22) && rc == IDC_OK;
        if (overlap) {
            // Decode everything into host memory: the output is copied in 16 chunks of whole units on the copy
            // stream, each as soon as the size classes of its units are done (Zipf-length lists in CSR order: the
            // short lists at the end of the array are finished, and on their way, long before the longest class).
            DecodeOverlap ov;
            HostTrace tr("roc_decode(host)");
            // milestone copies need the caller's buffer pinned / registered (copies into pageable memory are staged
            // by the driver and hold the host)
            void* host_dev = nullptr;
            {
                cudaPointerAttributes pa0{}, pa1{};
                if (cudaPointerGetAttributes(&pa0, ids_out) == cudaSuccess &&
                    cudaPointerGetAttributes(&pa1, (const uint8_t*)ids_out + total_out * id_bytes - 1) == cudaSuccess &&
                    pa0.type == cudaMemoryTypeHost && pa1.type == cudaMemoryTypeHost && pa0.devicePointer && pa1.devicePointer &&
                    (const uint8_t*)pa1.devicePointer - (const uint8_t*)pa0.devicePointer == (ptrdiff_t)(total_out * id_bytes - 1))
                    host_dev = pa0.devicePointer;
                (void)cudaGetLastError();
            }
            ov.want_ms = host_dev != nullptr && stream_wait_value32() != nullptr && memcpy_batch() != nullptr && getenv("IDC_NO_MILESTONES") == nullptr;
            rc = run_decode(c, b, d_unit, d_out, d_ws, ws_bytes, nunits_sel, out_dev, id_bytes, nullptr, 0,
                            ns.empty() ? 0u : ns[0], [&](uint64_t slot) { return ns[slot]; }, &ov);
            cudaStream_t copy_s = nullptr, copy_ms = nullptr;
            if (rc == IDC_OK) rc = c->copy_stream_get(&copy_s);
            if (rc == IDC_OK) rc = c->copy_stream_get(&copy_ms, 1);
            if (rc == IDC_OK) {
                auto class_of = [&](uint32_t n) {
                    size_t k = 0;
                    while (k + 1 < ov.class_min_n.size() && n < ov.class_min_n[k]) k++;
                    return k;
                };
                const uint64_t first = b->list_offsets[0], target = (total_out + 15) / 16;
                // A copy job is a range of whole units that waits for the size classes of its units (mask). The
                // units of the longest class, whose chains bound the kernel, do not wait for its end: the ids their
                // last ms_seg steps produced leave as ONE batch of copies (cudaMemcpyBatchAsync, a piece per unit)
                // per milestone counter of k_roc_decode -- the bulk of the output is on its way while the chains
                // are still running. Measured and dropped: a cudaMemcpy2DAsync per run of full-length units (10 456
                // calls at 17 us of host time each -- more than the kernels take), and a kernel that writes the
                // pieces into the pinned buffer itself (its PCIe stores back up into the SMs' memory pipes and slow
                // the chains next to them: decode 109 -> 165 ms).
                struct Job {
                    uint64_t e0, e1;   // elements [e0, e1) of the output; e1 == e0: milestone `part` of the longest class
                    uint32_t mask;
                    uint32_t part;
                    uint64_t est;      // estimated step at which it can go (the copy stream is in order)
                };
                std::vector<Job> jobs;
                const bool ms = ov.ms_count != 0;
                const uint32_t parts = ov.ms_count + 1u;
                auto est_of_mask = [&](uint32_t m) {
                    uint64_t e = 0;
                    for (size_t k = 0; k < ov.class_est.size(); k++)
                        if (m >> k & 1u) e = std::max(e, ov.class_est[k]);
                    return e;
                };
                auto by_milestone = [&](uint64_t u) { return ms && b->unit_n[u] && class_of(b->unit_n[u]) == 0; };
                for (uint64_t u0 = 0; u0 < b->nunits;) {
                    uint64_t u1 = u0;
                    if (by_milestone(u0)) {
                        u0++;
                        continue;
                    }
                    Job j{b->unit_src[u0] - first, b->unit_src[u0] - first, 0u, 0u, 0};
                    while (u1 < b->nunits && j.e1 - j.e0 < target && !by_milestone(u1)) {
                        if (b->unit_n[u1]) {
                            // a range of some size is not held back by (nor holds back) units that are ready at another time
                            const uint32_t k = (uint32_t)class_of(b->unit_n[u1]);
                            if (j.mask && !(j.mask >> k & 1u) && j.e1 - j.e0 >= (1ull << 20) && ov.class_est[k] != est_of_mask(j.mask)) break;
                            j.mask |= 1u << k;
                        }
                        j.e1 = b->unit_src[u1] - first + b->unit_n[u1];
                        u1++;
                    }
                    j.est = est_of_mask(j.mask);
                    if (j.e1 > j.e0) jobs.push_back(j);
                    u0 = u1;
                }
                std::vector<uint64_t> ms_units;
                std::vector<void*> bd, bs;
                std::vector<size_t> bn;
                if (ms)
                    for (uint64_t u = 0; u < b->nunits; u++)
                        if (by_milestone(u)) ms_units.push_back(u);
                if (ms)
                    for (uint32_t q = 0; q < parts; q++)
                        jobs.push_back(Job{0, 0, 1u, q, q + 1 < parts ? (uint64_t)(q + 1) * ov.ms_seg : ov.class_est[0]});
                // Two in-order copy streams: the ranges (queued first -- a batch call holds the host until the device
                // has taken its copies) and the milestone batches; each in the order its jobs become ready.
                std::stable_sort(jobs.begin(), jobs.end(), [](const Job& x, const Job& y) {
                    const bool mx = x.e1 == x.e0, my = y.e1 == y.e0;
                    return mx != my ? my : x.est < y.est;
                });
                tr.mark("launches + copy plan");
                cudaError_t e = cudaSuccess;
                if (ms) e = cudaStreamWaitEvent(copy_ms, ov.ms_armed, 0);
                uint32_t waited_mask = 0;  // classes the (in-order) copy stream is already behind
                const auto t_jobs = std::chrono::steady_clock::now();
                size_t job_no = 0;
                for (const Job& j : jobs) {
                    if (e != cudaSuccess) break;
                    if (tr.on && (j.e1 == j.e0 || job_no % 128 == 0 || (j.e1 - j.e0) * id_bytes > (32u << 20)))
                        fprintf(stderr, "[idc host]   job %zu (%s, est %llu, %.1f MB) issued at %.2f ms\n", job_no, j.e1 == j.e0 ? "milestone" : "range",
                                (unsigned long long)j.est, (j.e1 - j.e0) * id_bytes / 1e6,
                                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_jobs).count());
                    job_no++;
                    const bool by_counter = j.e1 == j.e0 && j.part + 1 < parts;
                    if (by_counter) {
                        if (stream_wait_value32()((CUstream)copy_ms, (CUdeviceptr)(uintptr_t)(ov.d_progress + j.part), ov.ms_expect,
                                                  CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                            e = cudaErrorUnknown;
                    } else if (j.e1 == j.e0) {
                        e = cudaStreamWaitEvent(copy_ms, ov.class_done[0], 0);
                    } else {
                        for (size_t k = 0; k < ov.class_done.size() && e == cudaSuccess; k++)
                            if ((j.mask & ~waited_mask) >> k & 1u) e = cudaStreamWaitEvent(copy_s, ov.class_done[k], 0);
                        waited_mask |= j.mask;
                    }
                    if (e != cudaSuccess) break;
                    if (j.e1 == j.e0) {
                        // a unit of n ids writes out[n - 1 - i] at step i: steps [part * seg, (part + 1) * seg)
                        const uint64_t s0 = (uint64_t)j.part * ov.ms_seg, s1 = s0 + ov.ms_seg;
                        bd.clear(), bs.clear(), bn.clear();
                        for (uint64_t u : ms_units) {
                            const uint64_t n = b->unit_n[u], base = b->unit_src[u] - first;
                            const uint64_t hi = n - std::min(n, s0), lo = n - std::min(n, s1);
                            if (hi == lo) continue;
                            bd.push_back((uint8_t*)ids_out + (base + lo) * id_bytes);
                            bs.push_back((uint8_t*)out_dev + (base + lo) * id_bytes);
                            bn.push_back((hi - lo) * id_bytes);
                        }
                        if (!bd.empty()) {
                            cudaMemcpyAttributes at{};
                            at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
                            at.flags = cudaMemcpyFlagPreferOverlapWithCompute;
                            size_t at_idx = 0, fail = 0;
                            e = memcpy_batch()(bd.data(), bs.data(), bn.data(), bd.size(), &at, &at_idx, 1, &fail, copy_ms);
                        }
                    } else {
                        e = cudaMemcpyAsync((uint8_t*)ids_out + j.e0 * id_bytes, (const uint8_t*)out_dev + j.e0 * id_bytes,
                                            (j.e1 - j.e0) * id_bytes, cudaMemcpyDeviceToHost, copy_s);
                    }
                }
                if (e != cudaSuccess) {
                    set_error("D2H copy failed: %s", cudaGetErrorString(e));
                    rc = IDC_ERR_CUDA;
                }
            }
            tr.mark("copies queued");
            int rc2 = finish_decode(c);
            tr.mark("kernels done");
            if (copy_s) cudaStreamSynchronize(copy_s);
            if (copy_ms) cudaStreamSynchronize(copy_ms);
            tr.mark("copies done");
            if (rc == IDC_OK) rc = rc2;
        } else {
            if (rc == IDC_OK)
                rc = run_decode(c, b, d_unit, d_out, d_ws, ws_bytes, nunits_sel, out_dev, id_bytes, nullptr, 0,
                                ns.empty() ? 0u : ns[0], [&](uint64_t slot) { return ns[slot]; });
            if (rc == IDC_OK && out_mem == IDC_MEM_HOST) {
                cudaError_t e = cudaMemcpyAsync(ids_out, out_dev, total_out * id_bytes, cudaMemcpyDeviceToHost, c->stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
                if (e != cudaSuccess) {
                    set_error("D2H copy failed: %s", cudaGetErrorString(e));
                    rc = IDC_ERR_CUDA;
                }
            }
        }
    }
    c->pool_release(t_unit);
    c->pool_release(t_out);
    c->pool_release(t_ws);
    return rc;
}

int idc_roc_translate(idc_ctx* c, const idc_roc_blob* b, const int64_t* labels, int labels_mem, uint64_t n,
                      int64_t* ids_out, int out_mem) {
    IDC_REQUIRE(c && b, IDC_ERR_ARG, "idc_roc_translate: null argument");
    IDC_REQUIRE(b->row_stride == 0, IDC_ERR_ARG, "row blob: labels address inverted lists");
    if (n == 0) return IDC_OK;
    IDC_REQUIRE(labels && ids_out, IDC_ERR_ARG, "idc_roc_translate: null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    // labels on the host: grouping the hits by list is metadata work (custom_invlists_impl.cpp:477-502)
    std::vector<int64_t> lab_h;
    const int64_t* lab = labels;
    if (labels_mem == IDC_MEM_DEVICE) {
        lab_h.resize(n);
        IDC_CUDA(cudaMemcpyAsync(lab_h.data(), labels, n * 8, cudaMemcpyDeviceToHost, c->stream));
        IDC_CUDA(cudaStreamSynchronize(c->stream));
        lab = lab_h.data();
    }
    // distinct hit lists, ascending, and where each of them starts in the decoded buffer. A flag per list when the batch
    // is large against the index (one pass, no sort: a million labels cost 5 ms instead of 90), a sort of the hit list
    // numbers when it is small.
    const bool dense = b->nlist <= 4 * n + 1024;
    std::vector<uint64_t> hit;
    std::vector<uint8_t> flag;
    if (dense) flag.assign(b->nlist, 0);
    else hit.reserve(n);
    for (uint64_t i = 0; i < n; i++) {
        if (lab[i] < 0) continue;
        const uint64_t l = (uint64_t)lab[i] >> 32, o = (uint64_t)lab[i] & 0xffffffffull;
        IDC_REQUIRE(l < b->nlist, IDC_ERR_ARG, "label %llu: list_no %llu out of range", (unsigned long long)i,
                    (unsigned long long)l);
        IDC_REQUIRE(o < b->list_offsets[l + 1] - b->list_offsets[l], IDC_ERR_ARG,
                    "label %llu: offset %llu past the end of list %llu", (unsigned long long)i, (unsigned long long)o,
                    (unsigned long long)l);
        if (dense) flag[l] = 1;
        else hit.push_back(l);
    }
    if (dense) {
        for (uint64_t l = 0; l < b->nlist; l++)
            if (flag[l]) hit.push_back(l);
    } else {
        std::sort(hit.begin(), hit.end());
        hit.erase(std::unique(hit.begin(), hit.end()), hit.end());
    }
    // decode plan for the hit lists; list_pos[h] = element offset of hit list h in the decoded buffer
    std::vector<uint32_t> units;
    std::vector<uint64_t> out_off, list_pos(hit.size());
    std::vector<uint64_t> pos_of;  // dense: the same, indexed by list number
    if (dense) pos_of.assign(b->nlist, 0);
    uint64_t pos = 0;
    for (size_t h = 0; h < hit.size(); h++) {
        list_pos[h] = pos;
        if (dense) pos_of[hit[h]] = pos;
        for (uint64_t u = b->unit_offsets[hit[h]]; u < b->unit_offsets[hit[h] + 1]; u++) {
            units.push_back((uint32_t)u);
            out_off.push_back(pos);
            pos += b->unit_n[u];
        }
    }
    std::vector<uint64_t> src(n, 0);
    for (uint64_t i = 0; i < n; i++) {
        if (lab[i] < 0) continue;
        const uint64_t l = (uint64_t)lab[i] >> 32, o = (uint64_t)lab[i] & 0xffffffffull;
        if (dense) {
            src[i] = pos_of[l] + o;
        } else {
            const size_t h = std::lower_bound(hit.begin(), hit.end(), l) - hit.begin();
            src[i] = list_pos[h] + o;
        }
    }
    // device buffers: decoded ids | src | labels (when they came from the host) | out (when it goes to the host)
    const size_t off_src = (pos * 8 + 255) & ~size_t(255), off_lab = off_src + ((n * 8 + 255) & ~size_t(255));
    const size_t off_out = off_lab + ((n * 8 + 255) & ~size_t(255));
    IDC_TRY(c->stage.reserve(off_out + n * 8 + 256));
    uint8_t* st = c->stage.as<uint8_t>();
    int64_t* d_dec = reinterpret_cast<int64_t*>(st);
    uint64_t* d_src = reinterpret_cast<uint64_t*>(st + off_src);
    const int64_t* d_lab = labels;
    if (labels_mem == IDC_MEM_HOST) {
        IDC_CUDA(cudaMemcpyAsync(st + off_lab, labels, n * 8, cudaMemcpyHostToDevice, c->stream));
        d_lab = reinterpret_cast<const int64_t*>(st + off_lab);
    }
    int64_t* d_out = out_mem == IDC_MEM_HOST ? reinterpret_cast<int64_t*>(st + off_out) : ids_out;
    IDC_TRY(upload(c, d_src, src));
    int rc = IDC_OK;
    uint32_t* t_unit = nullptr;
    uint64_t *t_out = nullptr, *t_ws = nullptr;
    if (!units.empty()) {
        uint64_t ws_bytes = 0;
        std::vector<uint32_t> t_ns;
        rc = build_decode_plan(c, b, units, out_off, &t_unit, &t_out, &t_ws, &ws_bytes, &t_ns);
        if (rc == IDC_OK)
            rc = run_decode(c, b, t_unit, t_out, t_ws, ws_bytes, units.size(), d_dec, 8, nullptr, 0, t_ns.empty() ? 0u : t_ns[0],
                            [&](uint64_t slot) { return t_ns[slot]; });
    }
    if (rc == IDC_OK) {
        {
            LaunchScope ls(c, "k_translate_gather");
            k_translate_gather<<<grid_for(n), kThreads, 0, c->stream>>>(d_lab, d_src, d_dec, d_out, n);
        }
        rc = check_last_launch("k_translate_gather");
    }
    if (rc == IDC_OK) {
        cudaError_t e = cudaSuccess;
        if (out_mem == IDC_MEM_HOST) e = cudaMemcpyAsync(ids_out, d_out, n * 8, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            set_error("idc_roc_translate: %s", cudaGetErrorString(e));
            rc = IDC_ERR_CUDA;
        }
    }
    c->pool_release(t_unit);
    c->pool_release(t_out);
    c->pool_release(t_ws);
    return rc;
}

int idc_roc_decode_rows(idc_ctx* c, const idc_roc_blob* b, const int32_t* row_nos, int rows_mem, uint64_t nsel,
                        int32_t* out, uint32_t* counts, int out_mem) {
    IDC_REQUIRE(c && b, IDC_ERR_ARG, "idc_roc_decode_rows: null argument");
    IDC_REQUIRE(b->row_stride != 0, IDC_ERR_ARG, "not a row blob");
    std::lock_guard<std::mutex> lock(c->mu);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    const uint32_t K = b->row_stride;
    if (row_nos == nullptr) nsel = b->nlist;
    if (nsel == 0) return IDC_OK;
    IDC_REQUIRE(out != nullptr, IDC_ERR_ARG, "out is NULL");
    // rows are short and near-uniform in length: no length sort, one fixed-size workspace slot per row, and the
    // kernel derives everything from the slot number (row mode of k_roc_decode): no host-side tables, the row
    // numbers stay where they are when they live on the device
    const uint64_t slot_ws = dec_tree_bytes(K);
    const uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(nsel, (4ull << 30) / slot_ws));
    if (row_nos && rows_mem == IDC_MEM_HOST && out_mem == IDC_MEM_HOST && nsel <= 4096 && nsel <= chunk) {
        // a few rows, everything on the host (what NSG search does per visited node, altid_impl.cpp:153-165): through the
        // context's mailbox -- [row numbers | counts | rows], read and written by the kernel in place: no copy calls
        const size_t o_cnt = (nsel * 4 + 15) & ~size_t(15), o_out = 2 * o_cnt;
        void *mh = nullptr, *md = nullptr;
        IDC_TRY(c->mailbox_get(o_out + nsel * K * 4, &mh, &md));
        if (mh) {
            std::memcpy(mh, row_nos, nsel * 4);
            uint8_t *h8 = static_cast<uint8_t*>(mh), *d8 = static_cast<uint8_t*>(md);
            // (run_decode synchronises the stream and turns an out-of-range row into IDC_ERR_ARG)
            if (small_rows(K))
                IDC_TRY(run_decode_small(c, b, reinterpret_cast<const int32_t*>(d8), 0, nsel, reinterpret_cast<int32_t*>(d8 + o_out),
                                         reinterpret_cast<uint32_t*>(d8 + o_cnt), K));
            else
                IDC_TRY(run_decode(c, b, nullptr, nullptr, nullptr, nsel * slot_ws, nsel, d8 + o_out, 4, reinterpret_cast<uint32_t*>(d8 + o_cnt),
                                   K, K, [&](uint64_t) { return K; }, nullptr, reinterpret_cast<const int32_t*>(d8), slot_ws, 0));
            std::memcpy(out, h8 + o_out, nsel * K * 4);
            if (counts) std::memcpy(counts, h8 + o_cnt, nsel * 4);
            return IDC_OK;
        }
    }
    int32_t* out_dev = out;
    uint32_t* cnt_dev = counts;
    if (out_mem == IDC_MEM_HOST) {
        IDC_TRY(c->stage.reserve(chunk * K * 4 + chunk * 4 + 256));
        out_dev = c->stage.as<int32_t>();
        cnt_dev = reinterpret_cast<uint32_t*>(c->stage.as<uint8_t>() + chunk * K * 4);
    }
    const int32_t* rows_dev = row_nos;
    if (row_nos && rows_mem == IDC_MEM_HOST) {
        IDC_TRY(c->meta.reserve(nsel * 4 + 256));
        IDC_CUDA(cudaMemcpyAsync(c->meta.p, row_nos, nsel * 4, cudaMemcpyHostToDevice, c->stream));
        rows_dev = c->meta.as<int32_t>();
    }
    for (uint64_t s = 0; s < nsel; s += chunk) {
        const uint64_t m = std::min(chunk, nsel - s);
        int32_t* od = out_mem == IDC_MEM_HOST ? out_dev : out_dev + s * K;
        uint32_t* cd = cnt_dev ? (out_mem == IDC_MEM_HOST ? cnt_dev : cnt_dev + s) : nullptr;
        if (small_rows(K))
            IDC_TRY(run_decode_small(c, b, rows_dev ? rows_dev + s : nullptr, s, m, od, cd, K));
        else
            IDC_TRY(run_decode(c, b, nullptr, nullptr, nullptr, m * slot_ws, m, od, 4, cd, K, K, [&](uint64_t) { return K; },
                               nullptr, rows_dev ? rows_dev + s : nullptr, slot_ws, s));
        if (out_mem == IDC_MEM_HOST) {
            IDC_CUDA(cudaMemcpyAsync(out + s * K, out_dev, m * K * 4, cudaMemcpyDeviceToHost, c->stream));
            if (counts) IDC_CUDA(cudaMemcpyAsync(counts + s, cnt_dev, m * 4, cudaMemcpyDeviceToHost, c->stream));
            IDC_CUDA(cudaStreamSynchronize(c->stream));
        }
    }
    return IDC_OK;
}

// ---- flat file form (idc_file.h): header words, the list CSR and the wire payload of idc_roc_blob_export_payload
int idc_roc_blob_save(const idc_roc_blob* b, const char* path) {
    IDC_REQUIRE(b && path, IDC_ERR_ARG, "idc_roc_blob_save: null argument");
    if (!b->host_tables) {
        std::lock_guard<std::mutex> lock(b->ctx->mu);
        IDC_TRY(roc_host_tables(b));
    }
    const uint64_t nu = b->nunits;
    std::vector<uint8_t> prec(nu);
    std::vector<uint64_t> heads(nu);
    std::vector<uint32_t> nw(nu), lo(nu), hi(nu), words(b->total_words);
    IDC_TRY(idc_roc_blob_export_payload(b, IDC_MEM_HOST, prec.data(), heads.data(), nw.data(), lo.data(), hi.data(), words.data()));
    std::vector<uint64_t> hdr{b->nlist, nu, b->total_words, b->max_unit, b->row_stride};
    FileWriter w;
    IDC_TRY(w.open(path, kFileRoc, 8));
    w.vec(hdr);
    w.vec(b->list_offsets);
    w.vec(prec);
    w.vec(heads);
    w.vec(nw);
    w.vec(lo);
    w.vec(hi);
    w.vec(words);
    return w.close(path);
}

int idc_roc_blob_load(idc_ctx* c, const char* path, idc_roc_blob** out) {
    IDC_REQUIRE(c && path && out, IDC_ERR_ARG, "idc_roc_blob_load: null argument");
    *out = nullptr;
    FileReader r;
    IDC_TRY(r.open(path, kFileRoc));
    std::vector<uint64_t> hdr, offs, heads;
    std::vector<uint8_t> prec;
    std::vector<uint32_t> nw, lo, hi, words;
    IDC_TRY(r.vec(hdr));
    IDC_TRY(r.vec(offs));
    IDC_TRY(r.vec(prec));
    IDC_TRY(r.vec(heads));
    IDC_TRY(r.vec(nw));
    IDC_TRY(r.vec(lo));
    IDC_TRY(r.vec(hi));
    IDC_TRY(r.vec(words));
    IDC_REQUIRE(hdr.size() == 5 && offs.size() == hdr[0] + 1 && prec.size() == hdr[1] && heads.size() == hdr[1] && nw.size() == hdr[1] &&
                    lo.size() == hdr[1] && hi.size() == hdr[1] && words.size() == hdr[2],
                IDC_ERR_ARG, "%s: section sizes do not match the header", path);
    // the units the list lengths imply (idc_roc_blob_assemble reads that many entries of the per-unit arrays)
    const uint64_t mu = hdr[3] ? hdr[3] : IDC_MAX_UNIT_DEFAULT;
    uint64_t implied = 0;
    for (uint64_t l = 0; l < hdr[0]; l++) {
        IDC_REQUIRE(offs[l + 1] >= offs[l], IDC_ERR_ARG, "%s: list offsets decrease", path);
        const uint64_t n = offs[l + 1] - offs[l];
        implied += n ? (n + mu - 1) / mu : 1;
    }
    IDC_REQUIRE(implied == hdr[1], IDC_ERR_ARG, "%s: %llu units stored, the list lengths imply %llu", path, (unsigned long long)hdr[1],
                (unsigned long long)implied);
    idc_roc_blob* b = nullptr;
    IDC_TRY(idc_roc_blob_assemble(c, hdr[0], offs.data(), (uint32_t)hdr[3], IDC_MEM_HOST, prec.data(), heads.data(), nw.data(), lo.data(),
                                  hi.data(), words.data(), words.size(), &b));
    b->row_stride = (uint32_t)hdr[4];  // graph rows: one unit per row (empty rows own an empty unit), decoded by idc_roc_decode_rows
    *out = b;
    return IDC_OK;
}

}  // extern "C"
