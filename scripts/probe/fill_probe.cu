// Probe: how fast can 1.5 GB be zero-filled?  uint4 grid-stride stores (what fill_zero_kernel does), 256-bit stores,
// streaming stores, cudaMemsetAsync, and different grid sizes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probe/fill_probe scripts/probe/fill_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__global__ void __launch_bounds__(256) fill16(uint4* dst, int64_t n16) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = z;
}
__global__ void __launch_bounds__(256) fill16cs(uint4* dst, int64_t n16) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) __stcs(dst + i, z);
}
__global__ void __launch_bounds__(256) fill32(unsigned char* dst, int64_t n32) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += stride)
        asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(dst + i * 32), "r"(0) : "memory");
}
// each CTA owns a contiguous chunk (better DRAM page locality than a grid-wide stride)
__global__ void __launch_bounds__(256) fill16blk(uint4* dst, int64_t n16) {
    const int64_t per = (n16 + gridDim.x - 1) / gridDim.x;
    const int64_t b = per * blockIdx.x, e = b + per < n16 ? b + per : n16;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int64_t i = b + threadIdx.x; i < e; i += 256) dst[i] = z;
}

template <typename F>
static void timeit(const char* name, F f, double gb) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
    }
    printf("%-34s %.4f ms  %.0f GB/s  %s\n", name, best, gb / (best * 1e-3), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int64_t bytes = 24ll * 496 * 65536 * 2;
    void* p;
    cudaMalloc(&p, bytes);
    const double gb = bytes / 1e9;
    for (int mult : {4, 8, 16, 32}) {
        char nm[64];
        snprintf(nm, 64, "uint4 stride, grid 148x%d", mult);
        timeit(nm, [&] { fill16<<<148 * mult, 256>>>((uint4*)p, bytes / 16); }, gb);
    }
    timeit("uint4 stride st.cs, grid 148x16", [&] { fill16cs<<<148 * 16, 256>>>((uint4*)p, bytes / 16); }, gb);
    timeit("256-bit stores, grid 148x16", [&] { fill32<<<148 * 16, 256>>>((unsigned char*)p, bytes / 32); }, gb);
    timeit("256-bit stores, grid 148x8", [&] { fill32<<<148 * 8, 256>>>((unsigned char*)p, bytes / 32); }, gb);
    timeit("uint4 contiguous per CTA, 148x16", [&] { fill16blk<<<148 * 16, 256>>>((uint4*)p, bytes / 16); }, gb);
    timeit("uint4 contiguous per CTA, 148x4", [&] { fill16blk<<<148 * 4, 256>>>((uint4*)p, bytes / 16); }, gb);
    for (int mult : {32, 64, 128, 256}) {
        char nm[64];
        snprintf(nm, 64, "256-bit stores, grid 148x%d", mult);
        timeit(nm, [&] { fill32<<<148 * mult, 256>>>((unsigned char*)p, bytes / 32); }, gb);
    }
    timeit("256-bit, one store per thread", [&] { fill32<<<(unsigned)((bytes / 32 + 255) / 256), 256>>>((unsigned char*)p, bytes / 32); }, gb);
    timeit("256-bit, 4 stores per thread", [&] { fill32<<<(unsigned)((bytes / 32 + 1023) / 1024), 256>>>((unsigned char*)p, bytes / 32); }, gb);
    timeit("uint4, 8 stores per thread", [&] { fill16<<<(unsigned)((bytes / 16 + 2047) / 2048), 256>>>((uint4*)p, bytes / 16); }, gb);
    timeit("cudaMemsetAsync", [&] { cudaMemsetAsync(p, 0, bytes); }, gb);
    return 0;
}
