// idc_scan.cuh -- exclusive prefix sums on the device, shared by the planning code of all three codecs.
#pragma once

#include "idc_host.h"

namespace {

using namespace idc;

// ---- exclusive prefix sums on the device (planning tables: per-list / per-unit offsets). Up to kScanMaxArrays arrays
// of n 64-bit values are scanned side by side (blockIdx.y picks the array): out[i] = sum of in[j], j < i, for
// i = 0 .. n (n + 1 outputs: the last one is the total). Three small kernels: tiles of 2048 values, the tile totals
// (one CTA), the tile bases added back.
constexpr int kScanMaxArrays = 6;
constexpr uint32_t kScanTileN = 2048;  // 256 threads x 8 values
struct ScanArgs {
    const uint64_t* in[kScanMaxArrays];
    uint64_t* out[kScanMaxArrays];  // n + 1 entries each
    uint64_t* tile_sum;             // [arrays][ntiles + 1] scratch
    uint64_t n, ntiles;
};

__device__ __forceinline__ uint64_t warp_incl_scan64(uint64_t v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t u = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)lane >= o) v += u;
    }
    return v;
}

// CTA-wide exclusive scan of one value per thread (256 threads); returns the CTA total through `total`
__device__ __forceinline__ uint64_t cta_excl_scan64(uint64_t v, uint64_t* sm_warp /* 8 */, uint64_t& total) {
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint64_t inc = warp_incl_scan64(v, lane);
    __syncthreads();
    if (lane == 31u) sm_warp[wid] = inc;
    __syncthreads();
    uint64_t base = 0, t = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint64_t w = sm_warp[j];
        base += (uint32_t)j < wid ? w : 0ull;
        t += w;
    }
    total = t;
    return base + inc - v;
}

__global__ void __launch_bounds__(256) k_scan_tiles(ScanArgs a) {
    __shared__ uint64_t sm_warp[8];
    const uint64_t* in = a.in[blockIdx.y];
    uint64_t* out = a.out[blockIdx.y];
    const uint64_t i0 = (uint64_t)blockIdx.x * kScanTileN + (uint64_t)threadIdx.x * 8u;
    uint64_t v[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        v[j] = i0 + j < a.n ? in[i0 + j] : 0ull;
        sum += v[j];
    }
    uint64_t total;
    uint64_t run = cta_excl_scan64(sum, sm_warp, total);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (i0 + j < a.n) out[i0 + j] = run;
        run += v[j];
    }
    if (threadIdx.x == 0) a.tile_sum[(uint64_t)blockIdx.y * (a.ntiles + 1) + blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) k_scan_sums(ScanArgs a) {
    __shared__ uint64_t sm_warp[8];
    uint64_t* ts = a.tile_sum + (uint64_t)blockIdx.y * (a.ntiles + 1);
    uint64_t carry = 0;
    for (uint64_t base = 0; base < a.ntiles; base += 256) {
        const uint64_t i = base + threadIdx.x;
        const uint64_t v = i < a.ntiles ? ts[i] : 0ull;
        uint64_t total;
        const uint64_t ex = cta_excl_scan64(v, sm_warp, total);
        if (i < a.ntiles) ts[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) {
        ts[a.ntiles] = carry;
        a.out[blockIdx.y][a.n] = carry;  // the total
    }
}

__global__ void __launch_bounds__(256) k_scan_add(ScanArgs a) {
    uint64_t* out = a.out[blockIdx.y];
    const uint64_t base = a.tile_sum[(uint64_t)blockIdx.y * (a.ntiles + 1) + blockIdx.x];
    const uint64_t i0 = (uint64_t)blockIdx.x * kScanTileN + (uint64_t)threadIdx.x * 8u;
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (i0 + j < a.n) out[i0 + j] += base;
}

// scratch bytes device_scan needs for `arrays` arrays of n values
inline size_t scan_scratch_bytes(uint64_t n, int arrays) { return (size_t)arrays * ((n + kScanTileN - 1) / kScanTileN + 1) * 8 + 64; }

inline int device_scan(idc_ctx* c, int arrays, const uint64_t* const* in, uint64_t* const* out, uint64_t n, uint64_t* scratch) {
    IDC_REQUIRE(arrays >= 1 && arrays <= kScanMaxArrays, IDC_ERR_ARG, "device_scan: %d arrays", arrays);
    ScanArgs a{};
    for (int j = 0; j < arrays; j++) a.in[j] = in[j], a.out[j] = out[j];
    a.tile_sum = scratch;
    a.n = n;
    a.ntiles = (n + kScanTileN - 1) / kScanTileN;
    LaunchScope ls(c, "k_scan");
    if (a.ntiles) k_scan_tiles<<<dim3((uint32_t)a.ntiles, (uint32_t)arrays), 256, 0, c->stream>>>(a);
    k_scan_sums<<<dim3(1, (uint32_t)arrays), 256, 0, c->stream>>>(a);
    if (a.ntiles > 1) k_scan_add<<<dim3((uint32_t)a.ntiles, (uint32_t)arrays), 256, 0, c->stream>>>(a);
    return check_last_launch("k_scan");
}

}  // namespace
