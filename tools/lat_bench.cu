// Micro-benchmark: latency of ONE dependent random 128-byte line read per step, in the shapes the ROC kernels
// could use (lane-per-chain vs. lane groups sharing a line), as a function of footprint and lanes per warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_bench lat_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ void ld256(const void* p, uint32_t* r) {
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ uint4 ld128(const void* p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld32(const void* p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// MODE 0: lane-per-chain, full line as 4 x 32 B      (active lanes = `lanes`)
// MODE 1: lane-per-chain, one 32-byte sector
// MODE 2: groups of G lanes share one chain; each lane reads 128/G bytes of the line; value exchanged by shuffle
template <int MODE, int G>
__global__ void chase(const uint8_t* data, uint64_t nlines, int steps, int lanes, uint64_t* out, int with_red) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    uint32_t acc = tid * 2654435761u;
    if (MODE == 2) acc = (tid / G) * 2654435761u;
    bool active = MODE == 2 ? true : (int)lane < lanes;
    if (active) {
        for (int s = 0; s < steps; s++) {
            uint64_t line = ((uint64_t)hash32(acc + s * 0x9e3779b9u) * nlines) >> 32;
            const uint8_t* p = data + line * 128;
            if (MODE == 0) {
                uint32_t r[32];
                ld256(p, r); ld256(p + 32, r + 8); ld256(p + 64, r + 16); ld256(p + 96, r + 24);
                acc += r[0] + r[9] + r[18] + r[31];
                if (with_red) atomicAnd((unsigned int*)p, ~(1u << (acc & 31)) | 0xffffffffu);
            } else if (MODE == 1) {
                uint32_t r[8];
                ld256(p, r);
                acc += r[0] + r[7];
            } else {
                uint32_t sub = lane % G, v;
                if (G == 32) v = ld32(p + sub * 4);
                else if (G == 8) { uint4 q = ld128(p + sub * 16); v = q.x + q.w; }
                else { uint32_t r[8]; ld256(p + sub * 32, r); v = r[0] + r[7]; }   // G == 4
                // every lane of the group needs the same next address: take the value of the lane picked by acc
                uint32_t src = (lane & ~(G - 1)) + (acc % G);
                acc += __shfl_sync(0xffffffffu, v, src) + 1u;
            }
        }
    }
    out[tid] = acc;
}

int main(int argc, char** argv) {
    int warps_per_sm = argc > 1 ? atoi(argv[1]) : 2;
    const int steps = 4000;
    uint64_t* o;
    cudaMalloc(&o, 148 * 64 * 32 * 8);
    for (double gb : {0.09, 1.0, 3.0, 24.0}) {
        uint64_t bytes = (uint64_t)(gb * (1ull << 30));
        uint64_t nlines = bytes / 128;
        uint8_t* d;
        if (cudaMalloc(&d, bytes) != cudaSuccess) { printf("alloc %.2f GB failed\n", gb); continue; }
        cudaMemset(d, 0, bytes);
        int blocks = 148 * warps_per_sm;  // one warp per CTA
        auto run = [&](const char* name, auto launch) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            float ms = 0;
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
            }
            printf("footprint %5.2f GB  warps/SM %d  %-44s %7.1f ns/step (%5.0f cyc @1.965GHz)\n", gb, warps_per_sm, name,
                   ms * 1e6 / steps, ms * 1e6 / steps * 1.965);
            cudaEventDestroy(a); cudaEventDestroy(b);
        };
        run("lane-per-chain 32 lanes, line 4x32B", [&] { chase<0, 1><<<blocks, 32>>>(d, nlines, steps, 32, o, 0); });
        run("lane-per-chain 32 lanes, line 4x32B + RED", [&] { chase<0, 1><<<blocks, 32>>>(d, nlines, steps, 32, o, 1); });
        run("lane-per-chain 16 lanes, line 4x32B", [&] { chase<0, 1><<<blocks, 32>>>(d, nlines, steps, 16, o, 0); });
        run("lane-per-chain  8 lanes, line 4x32B", [&] { chase<0, 1><<<blocks, 32>>>(d, nlines, steps, 8, o, 0); });
        run("lane-per-chain  4 lanes, line 4x32B", [&] { chase<0, 1><<<blocks, 32>>>(d, nlines, steps, 4, o, 0); });
        run("lane-per-chain  1 lane,  line 4x32B", [&] { chase<0, 1><<<blocks, 32>>>(d, nlines, steps, 1, o, 0); });
        run("lane-per-chain 32 lanes, one 32B sector", [&] { chase<1, 1><<<blocks, 32>>>(d, nlines, steps, 32, o, 0); });
        run("lane-per-chain  1 lane,  one 32B sector", [&] { chase<1, 1><<<blocks, 32>>>(d, nlines, steps, 1, o, 0); });
        run("group of 4 lanes per chain (8 lines/warp)", [&] { chase<2, 4><<<blocks, 32>>>(d, nlines, steps, 32, o, 0); });
        run("group of 8 lanes per chain (4 lines/warp)", [&] { chase<2, 8><<<blocks, 32>>>(d, nlines, steps, 32, o, 0); });
        run("group of 32 lanes per chain (1 line/warp)", [&] { chase<2, 32><<<blocks, 32>>>(d, nlines, steps, 32, o, 0); });
        cudaFree(d);
    }
    // the same with many warps per SM (throughput regime) for the group shapes
    {
        uint64_t bytes = 3ull << 30, nlines = bytes / 128;
        uint8_t* d;
        cudaMalloc(&d, bytes); cudaMemset(d, 0, bytes);
        for (int wps : {8, 16, 32}) {
            int blocks = 148 * wps;
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            float ms;
            cudaEventRecord(a); chase<2, 8><<<blocks, 32>>>(d, nlines, steps, 32, o, 0); cudaEventRecord(b);
            cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
            printf("3 GB, group of 8, warps/SM %2d: %7.1f ns/step\n", wps, ms * 1e6 / steps);
            cudaEventRecord(a); chase<0, 1><<<blocks, 32>>>(d, nlines, steps, 32, o, 0); cudaEventRecord(b);
            cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
            printf("3 GB, lane-per-chain 32, warps/SM %2d: %7.1f ns/step\n", wps, ms * 1e6 / steps);
        }
        cudaFree(d);
    }
    return 0;
}
