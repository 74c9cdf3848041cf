// Micro-benchmark: random-sector access rate of B200 HBM/L2 for the access shapes the ROC kernels use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
template <int MODE>
__device__ __forceinline__ uint64_t load8(const uint64_t* p) {
    uint64_t v;
    if (MODE == 0) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 1) asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
// each thread runs a dependent chain of `steps` random loads over `n64` uint64 (latency-bound per thread)
template <int MODE>
__global__ void chase(const uint64_t* data, uint64_t n64, int steps, uint64_t* out) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = tid;
    for (int s = 0; s < steps; s++) {
        uint64_t idx = ((uint64_t)hash32((uint32_t)acc + s * 0x9e3779b9u) * (n64 >> 5) >> 32) * 4;  // 32B-aligned sector
        acc += load8<MODE>(data + idx);
    }
    out[tid] = acc;
}
// same with a 32-byte (256-bit) load
__global__ void chase256(const uint64_t* data, uint64_t n64, int steps, uint64_t* out) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = tid;
    for (int s = 0; s < steps; s++) {
        uint64_t idx = ((uint64_t)hash32((uint32_t)acc + s * 0x9e3779b9u) * (n64 >> 5) >> 32) * 4;
        uint32_t r[8];
        asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(data + idx));
        acc += r[0] + r[7];
    }
    out[tid] = acc;
}
int main(int argc, char** argv) {
    size_t gb = argc > 1 ? atoi(argv[1]) : 8;
    int gran = argc > 2 ? atoi(argv[2]) : 0;
    if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); printf("set gran %d -> %s\n", gran, cudaGetErrorString(e)); }
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit = %zu\n", g);
    uint64_t n64 = gb * (1ull << 30) / 8;
    uint64_t *d, *o;
    cudaMalloc(&d, n64 * 8); cudaMemset(d, 0, n64 * 8);
    const int steps = 2000;
    for (int warps_per_sm : {4, 14, 32, 64}) {
        int threads = 148 * warps_per_sm * 32;
        cudaMalloc(&o, threads * 8);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int mode = 0; mode < 6; mode++) {
            float ms = 0;
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(a);
                int blocks = threads / 128;
                switch (mode) {
                    case 0: chase<0><<<blocks, 128>>>(d, n64, steps, o); break;
                    case 1: chase<1><<<blocks, 128>>>(d, n64, steps, o); break;
                    case 2: chase<2><<<blocks, 128>>>(d, n64, steps, o); break;
                    case 3: chase<3><<<blocks, 128>>>(d, n64, steps, o); break;
                    case 4: chase<4><<<blocks, 128>>>(d, n64, steps, o); break;
                    case 5: chase256<<<blocks, 128>>>(d, n64, steps, o); break;
                }
                cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
            }
            double acc = (double)threads * steps;
            const char* nm[] = {"cg.u64", "nc.u64", "cv.u64", "nc.noalloc", "ca.u64", "cg.v8.u32(32B)"};
            printf("warps/SM %2d  %-16s  %8.3f ms  %7.2f G loads/s  latency/step %6.0f ns\n", warps_per_sm, nm[mode], ms,
                   acc / ms / 1e6, ms * 1e6 / steps);
        }
        cudaFree(o);
    }
    return 0;
}
