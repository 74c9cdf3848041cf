// ef_kernels.cu -- Elias-Fano encode / bulk decode / random-access select and
// fixed-width bit packing on sm_100a. These are the HBM-bound kernels of the
// codec: every id is read once and every packed word is written once, with
// coalesced 8-byte stores; there is no contraction, hence no tensor cores.
//
//   k_ef_encode   one thread per OUTPUT word (gather formulation, ef_core.cuh)
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
#include "idc_host.h"
#include "idc_prep.cuh"

using namespace idc;

struct idc_ef_blob {
    idc_ctx* ctx = nullptr;
    uint64_t nlist = 0, total_ids = 0, low_words = 0, high_words = 0, bits_total = 0, nsamples = 0;
    uint32_t row_stride = 0;
    std::vector<uint64_t> list_offsets;  // nlist+1, ids per list (CSR)
    std::vector<uint8_t> l;
    std::vector<uint64_t> universe, low_off, high_off, samp_off;  // nlist(+1)
    uint64_t* d_list_off = nullptr;
    uint8_t* d_l = nullptr;
    uint64_t* d_low_off = nullptr;
    uint64_t* d_high_off = nullptr;
    uint64_t* d_samp_off = nullptr;
    uint64_t* d_low = nullptr;
    uint64_t* d_high = nullptr;
    uint32_t* d_samples = nullptr;
    uint64_t device_bytes = 0;
    // cached decode-everything tile table
    bool plan_ready = false;
    uint32_t* d_tile_list = nullptr;
    uint32_t* d_tile_idx = nullptr;
    uint64_t* d_tile_out = nullptr;
    uint64_t ntiles = 0;
    ~idc_ef_blob() {
        cudaFree(d_list_off);
        cudaFree(d_l);
        cudaFree(d_low_off);
        cudaFree(d_high_off);
        cudaFree(d_samp_off);
        cudaFree(d_low);
        cudaFree(d_high);
        cudaFree(d_samples);
        cudaFree(d_tile_list);
        cudaFree(d_tile_idx);
        cudaFree(d_tile_out);
    }
};

namespace {

constexpr uint32_t kEncTileWords = 1024;  // output words per warp in k_ef_encode
constexpr uint32_t kDecTile = 1024;       // ids per warp in k_ef_decode (4 select samples)

struct EfEncArgs {
    const void* ids;
    const uint64_t* list_src;   // element offset of each list in ids
    const uint64_t* list_off;   // CSR (n = list_off[l+1] - list_off[l])
    const uint8_t* l;
    const uint32_t* list_hi;    // universe = max id
    const uint64_t* low_off;
    const uint64_t* high_off;
    const uint64_t* samp_off;
    uint64_t* low;
    uint64_t* high;
    uint32_t* samples;
    const uint32_t* tile_list;
    const uint32_t* tile_idx;
    uint32_t ntiles;
};

template <typename IdT>
__global__ void __launch_bounds__(kThreads) k_ef_encode(EfEncArgs a) {
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.ntiles) return;
    uint32_t L = a.tile_list[warp];
    uint64_t m = a.list_off[L + 1] - a.list_off[L];
    if (m == 0) return;
    const IdT* ids = reinterpret_cast<const IdT*>(a.ids) + a.list_src[L];
    uint32_t l = a.l[L];
    uint64_t universe = a.list_hi[L];
    uint64_t lw = a.low_off[L + 1] - a.low_off[L], hw = a.high_off[L + 1] - a.high_off[L];
    uint64_t w0 = (uint64_t)a.tile_idx[warp] * kEncTileWords;
    uint64_t w1 = w0 + kEncTileWords < lw + hw ? w0 + kEncTileWords : lw + hw;
    uint64_t* low = a.low + a.low_off[L];
    uint64_t* high = a.high + a.high_off[L];
    uint32_t* samples = a.samples + a.samp_off[L];
    for (uint64_t w = w0 + lane; w < w1; w += 32) {
        if (w < lw) {
            low[w] = ef_low_word(ids, m, l, w);
        } else {
            uint64_t hwi = w - lw, p0 = hwi * 64;
            // same search as ef_high_word, plus the select samples
            uint64_t span = universe >> l;
            uint64_t lo = p0 > span ? p0 - span : 0, hi = p0 < m ? p0 : m;
            if (lo > hi) lo = hi;
            while (lo < hi) {
                uint64_t mid = (lo + hi) >> 1;
                if (ef_high_pos(ids, mid, l) < p0)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            uint64_t out = 0;
            for (uint64_t i = lo; i < m; i++) {
                uint64_t hp = ef_high_pos(ids, i, l);
                if (hp >= p0 + 64) break;
                out |= 1ull << (hp - p0);
                if ((i & (kEfSample - 1)) == 0) samples[i >> kEfSampleLog] = (uint32_t)hp;
            }
            high[hwi] = out;
        }
    }
}

struct EfDecArgs {
    const uint64_t* list_off;
    const uint8_t* l;
    const uint64_t* low_off;
    const uint64_t* high_off;
    const uint64_t* samp_off;
    const uint64_t* low;
    const uint64_t* high;
    const uint32_t* samples;
    const uint32_t* tile_list;
    const uint32_t* tile_idx;
    const uint64_t* tile_out;   // element offset in out of the tile's first id
    void* out;
    uint32_t* counts;           // rows: ids in the row (indexed by tile = slot)
    uint32_t ntiles;
    uint32_t row_stride;
};

template <typename OutT>
__global__ void __launch_bounds__(kThreads) k_ef_decode(EfDecArgs a) {
    __shared__ uint32_t s_hi[kThreads / 32][kDecTile];
    uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t warp = blockIdx.x * (kThreads / 32) + wib;
    if (warp >= a.ntiles) return;
    uint32_t L = a.tile_list[warp];
    uint64_t m = a.list_off[L + 1] - a.list_off[L];
    uint64_t e0 = (uint64_t)a.tile_idx[warp] * kDecTile;
    uint32_t tn = (uint32_t)(m - e0 < kDecTile ? m - e0 : kDecTile);
    if (m <= e0) tn = 0;
    OutT* out = reinterpret_cast<OutT*>(a.out) + a.tile_out[warp];
    uint32_t l = a.l[L];
    const uint64_t* low = a.low + a.low_off[L];
    const uint64_t* high = a.high + a.high_off[L];
    uint64_t hw = a.high_off[L + 1] - a.high_off[L];
    uint32_t* hi_part = s_hi[wib];
    if (tn) {
        uint64_t pos = a.samples[a.samp_off[L] + (e0 >> kEfSampleLog)];
        uint64_t w = pos >> 6;
        uint32_t cnt = 0;
        bool first = true;
        while (cnt < tn) {
            uint64_t wi = w + lane;
            uint64_t bits = wi < hw ? __ldg(reinterpret_cast<const unsigned long long*>(high) + wi) : 0ull;
            if (first && lane == 0) bits &= ~0ull << (pos & 63);
            uint32_t c = (uint32_t)__popcll(bits);
            uint32_t incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += v;
            }
            uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t idx = cnt + incl - c;
            while (bits && idx < tn) {
                uint32_t b = (uint32_t)__ffsll((long long)bits) - 1u;
                bits &= bits - 1;
                hi_part[idx] = (uint32_t)(wi * 64 + b - (e0 + idx));
                idx++;
            }
            cnt += total;
            w += 32;
            first = false;
            if (total == 0 && w >= hw) break;  // corrupt blob guard
        }
        __syncwarp();
        for (uint32_t t = lane; t < tn; t += 32) {
            uint64_t id = ((uint64_t)hi_part[t] << l) | ef_get_low(low, e0 + t, l);
            out[t] = (OutT)id;
        }
    }
    if (a.row_stride) {
        for (uint32_t t = tn + lane; t < a.row_stride; t += 32) out[t] = (OutT)-1;
        if (a.counts && lane == 0) a.counts[warp] = tn;
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

// ------------------------------------------------------------------ host

int ef_build(idc_ctx* c, idc_ef_blob* b, const void* ids_dev, int id_bytes, uint32_t flags,
             const std::vector<uint64_t>& list_src, uint64_t id_elems) {
    const uint64_t nl = b->nlist;
    // per-list metadata via the shared unit kernel (a list is one "unit" here)
    std::vector<uint32_t> n32(nl);
    for (uint64_t i = 0; i < nl; i++) {
        uint64_t n = b->list_offsets[i + 1] - b->list_offsets[i];
        IDC_REQUIRE(n < (1ull << 32), IDC_ERR_ARG, "list %llu too long", (unsigned long long)i);
        n32[i] = (uint32_t)n;
    }
    IDC_TRY(c->meta.reserve(nl * (8 + 4 + 1 + 4 + 4) + 1024));
    uint8_t* mp = c->meta.as<uint8_t>();
    auto carve = [&](size_t bytes) {
        uint8_t* r = mp;
        mp += (bytes + 15) & ~size_t(15);
        return r;
    };
    uint64_t* d_src = (uint64_t*)carve(nl * 8);
    uint32_t* d_n = (uint32_t*)carve(nl * 4);
    uint32_t* d_lo = (uint32_t*)carve(nl * 4);
    uint32_t* d_hi = (uint32_t*)carve(nl * 4);
    uint8_t* d_prec = (uint8_t*)carve(nl);
    IDC_TRY(c->status.reserve(64));
    uint32_t* d_status = c->status.as<uint32_t>();
    IDC_CUDA(cudaMemsetAsync(d_status, 0, 4, c->stream));
    IDC_TRY(upload(c, d_src, list_src));
    IDC_TRY(upload(c, d_n, n32));
    const bool sorted_in = (flags & IDC_F_SORTED) != 0;
    {
        MetaArgs m{ids_dev, d_src, d_n, (uint32_t)nl, sorted_in ? 1u : 0u, 0u, d_prec, d_lo, d_hi, d_status};
        LaunchScope ls(c, "k_unit_meta");
        if (id_bytes == 8)
            k_unit_meta<int64_t><<<grid_for(nl * 32), kThreads, 0, c->stream>>>(m);
        else
            k_unit_meta<uint32_t><<<grid_for(nl * 32), kThreads, 0, c->stream>>>(m);
    }
    IDC_TRY(check_last_launch("k_unit_meta"));
    std::vector<uint32_t> hi(nl);
    uint32_t st = 0;
    if (nl) IDC_CUDA(cudaMemcpyAsync(hi.data(), d_hi, nl * 4, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, c->stream));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    IDC_TRY(status_to_error(st, "ef_encode"));

    // shapes
    b->l.resize(nl);
    b->universe.resize(nl);
    b->low_off.assign(nl + 1, 0);
    b->high_off.assign(nl + 1, 0);
    b->samp_off.assign(nl + 1, 0);
    std::vector<uint32_t> tile_list, tile_idx;
    uint64_t bits_total = 0;
    for (uint64_t i = 0; i < nl; i++) {
        EfShape s = ef_shape(hi[i], n32[i]);
        b->l[i] = (uint8_t)s.l;
        b->universe[i] = hi[i];
        b->low_off[i + 1] = b->low_off[i] + s.low_words;
        b->high_off[i + 1] = b->high_off[i] + s.high_words;
        b->samp_off[i + 1] = b->samp_off[i] + s.samples;
        bits_total += s.low_bits + s.high_bits;
        uint64_t words = s.low_words + s.high_words;
        for (uint64_t t = 0; t * kEncTileWords < words; t++) {
            tile_list.push_back((uint32_t)i);
            tile_idx.push_back((uint32_t)t);
        }
    }
    b->low_words = b->low_off[nl];
    b->high_words = b->high_off[nl];
    b->nsamples = b->samp_off[nl];
    b->bits_total = bits_total;
    uint64_t acct = 0;
    IDC_TRY(dev_alloc(&b->d_list_off, nl + 1, &acct));
    IDC_TRY(dev_alloc(&b->d_l, nl, &acct));
    IDC_TRY(dev_alloc(&b->d_low_off, nl + 1, &acct));
    IDC_TRY(dev_alloc(&b->d_high_off, nl + 1, &acct));
    IDC_TRY(dev_alloc(&b->d_samp_off, nl + 1, &acct));
    IDC_TRY(dev_alloc(&b->d_low, b->low_words, &acct));
    IDC_TRY(dev_alloc(&b->d_high, b->high_words, &acct));
    IDC_TRY(dev_alloc(&b->d_samples, b->nsamples, &acct));
    IDC_TRY(upload(c, b->d_list_off, b->list_offsets));
    IDC_TRY(upload(c, b->d_l, b->l));
    IDC_TRY(upload(c, b->d_low_off, b->low_off));
    IDC_TRY(upload(c, b->d_high_off, b->high_off));
    IDC_TRY(upload(c, b->d_samp_off, b->samp_off));

    // sort when needed (ids < 2^32 was checked by the metadata kernel)
    const void* enc_ids = ids_dev;
    int enc_id_bytes = id_bytes;
    const uint64_t ntiles = tile_list.size();
    size_t tile_bytes = ((ntiles * 4 + 255) & ~size_t(255)) * 2;
    size_t ws_need = tile_bytes;
    size_t sorted_off = 0, sortidx_off = 0, big_off = 0, posbase_off = 0;
    uint32_t sort_grid = 0;
    if (!sorted_in) {
        sort_grid = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(nl, 1), (uint64_t)c->sm_count * 8);
        bool need_big = false;
        for (uint64_t i = 0; i < nl; i++) {
            IDC_REQUIRE(n32[i] <= kMaxUnit, IDC_ERR_DOMAIN,
                        "unsorted list %llu has %u ids: the device sort handles lists of <= 65536 ids; "
                        "pass ascending ids with IDC_F_SORTED", (unsigned long long)i, n32[i]);
            need_big |= n32[i] > kSortSmem;
        }
        sorted_off = (ws_need + 255) & ~size_t(255);
        sortidx_off = sorted_off + ((id_elems * 4 + 255) & ~size_t(255));
        posbase_off = sortidx_off + ((id_elems * 4 + 255) & ~size_t(255));
        big_off = posbase_off + ((nl * 4 + 255) & ~size_t(255));
        ws_need = big_off + (need_big ? (size_t)sort_grid * kMaxUnit * 8 : 0);
    }
    IDC_TRY(c->ws.reserve(ws_need + 256));
    uint32_t* d_tile_list = c->ws.as<uint32_t>();
    uint32_t* d_tile_idx = reinterpret_cast<uint32_t*>(c->ws.as<uint8_t>() + tile_bytes / 2);
    IDC_TRY(upload(c, d_tile_list, tile_list));
    IDC_TRY(upload(c, d_tile_idx, tile_idx));
    if (!sorted_in && nl) {
        uint32_t* d_sorted = (uint32_t*)(c->ws.as<uint8_t>() + sorted_off);
        uint32_t* d_sort_idx = (uint32_t*)(c->ws.as<uint8_t>() + sortidx_off);
        uint32_t* d_posbase = (uint32_t*)(c->ws.as<uint8_t>() + posbase_off);
        IDC_CUDA(cudaMemsetAsync(d_posbase, 0, nl * 4, c->stream));
        SortArgs s{ids_dev, d_src, d_n, d_posbase, (uint32_t)nl, d_sorted, d_sort_idx,
                   (uint64_t*)(c->ws.as<uint8_t>() + big_off)};
        LaunchScope ls(c, "k_sort_units");
        if (id_bytes == 8)
            k_sort_units<int64_t><<<sort_grid, 256, 0, c->stream>>>(s);
        else
            k_sort_units<uint32_t><<<sort_grid, 256, 0, c->stream>>>(s);
        enc_ids = d_sorted;
        enc_id_bytes = 4;
    }
    IDC_TRY(check_last_launch("k_sort_units"));
    if (ntiles) {
        EfEncArgs e{enc_ids, d_src, b->d_list_off, b->d_l, d_hi, b->d_low_off, b->d_high_off, b->d_samp_off,
                    b->d_low, b->d_high, b->d_samples, d_tile_list, d_tile_idx, (uint32_t)ntiles};
        LaunchScope ls(c, "k_ef_encode");
        if (enc_id_bytes == 8)
            k_ef_encode<int64_t><<<grid_for(ntiles * 32), kThreads, 0, c->stream>>>(e);
        else
            k_ef_encode<uint32_t><<<grid_for(ntiles * 32), kThreads, 0, c->stream>>>(e);
    }
    IDC_TRY(check_last_launch("k_ef_encode"));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
    b->device_bytes = acct;
    return IDC_OK;
}

int ef_run_decode(idc_ctx* c, const idc_ef_blob* b, const uint32_t* d_tile_list, const uint32_t* d_tile_idx,
                  const uint64_t* d_tile_out, uint64_t ntiles, void* out_dev, int id_bytes, uint32_t* counts_dev,
                  uint32_t row_stride) {
    if (ntiles == 0) return IDC_OK;
    EfDecArgs a{b->d_list_off, b->d_l, b->d_low_off, b->d_high_off, b->d_samp_off, b->d_low, b->d_high, b->d_samples,
                d_tile_list, d_tile_idx, d_tile_out, out_dev, counts_dev, (uint32_t)ntiles, row_stride};
    {
        LaunchScope ls(c, "k_ef_decode");
        if (id_bytes == 8)
            k_ef_decode<int64_t><<<grid_for(ntiles * 32), kThreads, 0, c->stream>>>(a);
        else
            k_ef_decode<int32_t><<<grid_for(ntiles * 32), kThreads, 0, c->stream>>>(a);
    }
    IDC_TRY(check_last_launch("k_ef_decode"));
    IDC_CUDA(cudaStreamSynchronize(c->stream));
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
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    std::unique_ptr<idc_ef_blob> b(new idc_ef_blob());
    b->ctx = c;
    b->nlist = nlist;
    b->list_offsets.resize(nlist + 1);
    std::vector<uint64_t> src(nlist);
    for (uint64_t l = 0; l <= nlist; l++) {
        IDC_REQUIRE(l == 0 || offsets[l] >= offsets[l - 1], IDC_ERR_ARG, "offsets must be non-decreasing");
        b->list_offsets[l] = offsets[l] - offsets[0];
        if (l < nlist) src[l] = offsets[l];
    }
    b->total_ids = b->list_offsets[nlist];
    uint64_t elems = offsets[nlist];
    IDC_REQUIRE(ids != nullptr || elems == 0, IDC_ERR_ARG, "ids is NULL");
    const void* ids_dev = ids;
    if (ids_mem == IDC_MEM_HOST && elems) {
        IDC_TRY(c->stage.reserve(elems * id_bytes));
        IDC_CUDA(cudaMemcpyAsync(c->stage.p, ids, elems * id_bytes, cudaMemcpyHostToDevice, c->stream));
        ids_dev = c->stage.p;
    }
    IDC_TRY(ef_build(c, b.get(), ids_dev, id_bytes, flags, src, elems));
    *out = b.release();
    return IDC_OK;
}

int idc_ef_encode_rows(idc_ctx* c, uint64_t nrows, uint32_t K, const int32_t* data, int data_mem, uint32_t flags,
                       idc_ef_blob** out) {
    IDC_REQUIRE(c && out && (data || nrows == 0), IDC_ERR_ARG, "idc_ef_encode_rows: null argument");
    IDC_REQUIRE(K >= 1 && K <= kMaxUnit, IDC_ERR_ARG, "K out of range");
    IDC_REQUIRE(nrows < (1ull << 32), IDC_ERR_ARG, "too many rows");
    *out = nullptr;
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    std::unique_ptr<idc_ef_blob> b(new idc_ef_blob());
    b->ctx = c;
    b->nlist = nrows;
    b->row_stride = K;
    const int32_t* d_data = data;
    uint64_t elems = nrows * K;
    if (data_mem == IDC_MEM_HOST && elems) {
        IDC_TRY(c->stage.reserve(elems * 4));
        IDC_CUDA(cudaMemcpyAsync(c->stage.p, data, elems * 4, cudaMemcpyHostToDevice, c->stream));
        d_data = c->stage.as<int32_t>();
    }
    std::vector<uint32_t> cnt(nrows);
    if (nrows) {
        IDC_TRY(c->meta.reserve(nrows * 4 + 256));
        uint32_t* d_cnt = c->meta.as<uint32_t>();
        {
            LaunchScope ls(c, "k_row_counts");
            k_row_counts<<<grid_for(nrows * 32), kThreads, 0, c->stream>>>(d_data, nrows, K, d_cnt);
        }
        IDC_TRY(check_last_launch("k_row_counts"));
        IDC_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, nrows * 4, cudaMemcpyDeviceToHost, c->stream));
        IDC_CUDA(cudaStreamSynchronize(c->stream));
    }
    b->list_offsets.resize(nrows + 1);
    std::vector<uint64_t> src(nrows);
    uint64_t total = 0;
    for (uint64_t r = 0; r < nrows; r++) {
        b->list_offsets[r] = total;
        src[r] = r * K;
        total += cnt[r];
    }
    b->list_offsets[nrows] = total;
    b->total_ids = total;
    IDC_TRY(ef_build(c, b.get(), d_data, 4, flags & ~IDC_F_SORTED, src, elems));
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
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    idc_ef_blob* mb = const_cast<idc_ef_blob*>(b);
    std::vector<uint32_t> tl, ti;
    std::vector<uint64_t> to;
    uint64_t total_out = 0;
    auto add_list = [&](uint64_t L, uint64_t pos) {
        uint64_t m = b->list_offsets[L + 1] - b->list_offsets[L];
        for (uint64_t t = 0; t * kDecTile < m; t++) {
            tl.push_back((uint32_t)L);
            ti.push_back((uint32_t)t);
            to.push_back(pos + t * kDecTile);
        }
        return m;
    };
    const uint32_t *d_tl, *d_ti;
    const uint64_t* d_to;
    uint64_t ntiles;
    uint32_t *t_tl = nullptr, *t_ti = nullptr;
    uint64_t* t_to = nullptr;
    if (list_nos == nullptr) {
        if (!mb->plan_ready) {
            for (uint64_t L = 0; L < b->nlist; L++) add_list(L, b->list_offsets[L]);
            mb->ntiles = tl.size();
            IDC_TRY(dev_alloc(&mb->d_tile_list, tl.size()));
            IDC_TRY(dev_alloc(&mb->d_tile_idx, tl.size()));
            IDC_TRY(dev_alloc(&mb->d_tile_out, tl.size()));
            IDC_TRY(upload(c, mb->d_tile_list, tl));
            IDC_TRY(upload(c, mb->d_tile_idx, ti));
            IDC_TRY(upload(c, mb->d_tile_out, to));
            IDC_CUDA(cudaStreamSynchronize(c->stream));
            mb->plan_ready = true;
        }
        d_tl = b->d_tile_list;
        d_ti = b->d_tile_idx;
        d_to = b->d_tile_out;
        ntiles = b->ntiles;
        total_out = b->total_ids;
        if (out_offsets) memcpy(out_offsets, b->list_offsets.data(), (b->nlist + 1) * 8);
    } else {
        uint64_t pos = 0;
        for (uint64_t i = 0; i < nsel; i++) {
            IDC_REQUIRE(list_nos[i] < b->nlist, IDC_ERR_ARG, "list_no out of range");
            if (out_offsets) out_offsets[i] = pos;
            pos += add_list(list_nos[i], pos);
        }
        if (out_offsets) out_offsets[nsel] = pos;
        total_out = pos;
        ntiles = tl.size();
        IDC_TRY(dev_alloc(&t_tl, ntiles));
        IDC_TRY(dev_alloc(&t_ti, ntiles));
        IDC_TRY(dev_alloc(&t_to, ntiles));
        IDC_TRY(upload(c, t_tl, tl));
        IDC_TRY(upload(c, t_ti, ti));
        IDC_TRY(upload(c, t_to, to));
        d_tl = t_tl;
        d_ti = t_ti;
        d_to = t_to;
    }
    int rc = IDC_OK;
    if (total_out) {
        IDC_REQUIRE(ids_out != nullptr, IDC_ERR_ARG, "ids_out is NULL");
        void* out_dev = ids_out;
        if (out_mem == IDC_MEM_HOST) {
            rc = c->stage.reserve(total_out * id_bytes);
            out_dev = c->stage.p;
        }
        if (rc == IDC_OK) rc = ef_run_decode(c, b, d_tl, d_ti, d_to, ntiles, out_dev, id_bytes, nullptr, 0);
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
    cudaFree(t_tl);
    cudaFree(t_ti);
    cudaFree(t_to);
    return rc;
}

int idc_ef_decode_rows(idc_ctx* c, const idc_ef_blob* b, const int32_t* row_nos, int rows_mem, uint64_t nsel,
                       int32_t* out, uint32_t* counts, int out_mem) {
    IDC_REQUIRE(c && b, IDC_ERR_ARG, "idc_ef_decode_rows: null argument");
    IDC_REQUIRE(b->row_stride != 0, IDC_ERR_ARG, "not a row blob");
    IDC_REQUIRE(b->row_stride <= kDecTile, IDC_ERR_ARG, "row stride > %u not supported", kDecTile);
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    const uint32_t K = b->row_stride;
    if (row_nos == nullptr) nsel = b->nlist;
    if (nsel == 0) return IDC_OK;
    IDC_REQUIRE(out != nullptr, IDC_ERR_ARG, "out is NULL");
    std::vector<int32_t> rows_h;
    if (row_nos && rows_mem == IDC_MEM_DEVICE) {
        rows_h.resize(nsel);
        IDC_CUDA(cudaMemcpyAsync(rows_h.data(), row_nos, nsel * 4, cudaMemcpyDeviceToHost, c->stream));
        IDC_CUDA(cudaStreamSynchronize(c->stream));
        row_nos = rows_h.data();
    }
    std::vector<uint32_t> tl(nsel), ti(nsel, 0);
    std::vector<uint64_t> to(nsel);
    for (uint64_t i = 0; i < nsel; i++) {
        int64_t r = row_nos ? row_nos[i] : (int64_t)i;
        IDC_REQUIRE(r >= 0 && (uint64_t)r < b->nlist, IDC_ERR_ARG, "row %lld out of range", (long long)r);
        tl[i] = (uint32_t)r;
        to[i] = i * K;
    }
    IDC_TRY(c->meta.reserve(nsel * 16 + 1024));
    uint32_t* d_tl = c->meta.as<uint32_t>();
    uint32_t* d_ti = reinterpret_cast<uint32_t*>(c->meta.as<uint8_t>() + ((nsel * 4 + 255) & ~255ull));
    uint64_t* d_to = reinterpret_cast<uint64_t*>(c->meta.as<uint8_t>() + 2 * ((nsel * 4 + 255) & ~255ull));
    IDC_TRY(upload(c, d_tl, tl));
    IDC_TRY(upload(c, d_ti, ti));
    IDC_TRY(upload(c, d_to, to));
    int32_t* out_dev = out;
    uint32_t* cnt_dev = counts;
    if (out_mem == IDC_MEM_HOST) {
        IDC_TRY(c->stage.reserve(nsel * K * 4 + nsel * 4 + 256));
        out_dev = c->stage.as<int32_t>();
        cnt_dev = reinterpret_cast<uint32_t*>(c->stage.as<uint8_t>() + nsel * K * 4);
    }
    IDC_TRY(ef_run_decode(c, b, d_tl, d_ti, d_to, nsel, out_dev, 4, cnt_dev, K));
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
    IDC_CUDA(cudaSetDevice(c->device));
    c->begin_call();
    if (nq == 0) return IDC_OK;
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

}  // extern "C"
