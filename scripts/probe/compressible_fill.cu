// Probe: does the B200 take compressible memory (cuMemCreate, CU_MEM_ALLOCATION_COMP_GENERIC), and what does a zero fill /
// a streaming read of zeros / of random data cost on it compared with ordinary device memory?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probe/compressible_fill scripts/probe/compressible_fill.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("%s -> %s\n", #x, s_); return 1; } } while (0)
#define RK(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(r_)); return 1; } } while (0)

__global__ void fill256(uint4* p, size_t n32) {            // n32 = number of 32-byte units
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n32) asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(reinterpret_cast<char*>(p) + i * 32), "r"(0) : "memory");
}
__global__ void fill_rand(uint4* p, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u;
        p[i] = make_uint4(h, h * 31u + 7u, h ^ 0x9e3779b9u, h * 17u);
    }
}
__global__ void read_sum(const uint4* __restrict__ p, size_t n, uint32_t* out) {
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = p[i];
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *out = acc;
}

static float time_ms(void (*launch)(void*, size_t), void* p, size_t bytes, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) launch(p, bytes);
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) launch(p, bytes);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}
static uint32_t* g_out;
static void l_fill(void* p, size_t bytes) { fill256<<<(unsigned)((bytes / 32 + 255) / 256), 256>>>((uint4*)p, bytes / 32); }
static void l_memset(void* p, size_t bytes) { cudaMemsetAsync(p, 0, bytes); }
static void l_rand(void* p, size_t bytes) { fill_rand<<<148 * 16, 256>>>((uint4*)p, bytes / 16); }
static void l_read(void* p, size_t bytes) { read_sum<<<148 * 16, 256>>>((const uint4*)p, bytes / 16, g_out); }

int main() {
    CK(cuInit(0));
    RK(cudaSetDevice(0));
    RK(cudaFree(0));
    CUdevice dev;
    CK(cuDeviceGet(&dev, 0));
    int comp = 0;
    CK(cuDeviceGetAttribute(&comp, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, dev));
    printf("GENERIC_COMPRESSION_SUPPORTED = %d\n", comp);
    RK(cudaMalloc(&g_out, 4));
    const size_t bytes = (size_t)1560 << 20;               // ~ the headline grad_rep (1.56 GB)

    void* plain = nullptr;
    RK(cudaMalloc(&plain, bytes));
    printf("plain      : fill256 %.4f ms  memset %.4f ms", time_ms(l_fill, plain, bytes, 20), time_ms(l_memset, plain, bytes, 20));
    l_fill(plain, bytes);
    printf("  read(zeros) %.4f ms", time_ms(l_read, plain, bytes, 20));
    l_rand(plain, bytes);
    printf("  read(random) %.4f ms\n", time_ms(l_read, plain, bytes, 20));

    if (!comp) return 0;
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
    size_t gran = 0;
    CK(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    const size_t sz = (bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h;
    CK(cuMemCreate(&h, sz, &prop, 0));
    CUmemAllocationProp got = {};
    CK(cuMemGetAllocationPropertiesFromHandle(&got, h));
    printf("granularity %zu, compressionType granted = %d\n", gran, (int)got.allocFlags.compressionType);
    CUdeviceptr va;
    CK(cuMemAddressReserve(&va, sz, 0, 0, 0));
    CK(cuMemMap(va, sz, 0, h, 0));
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CK(cuMemSetAccess(va, sz, &acc, 1));
    void* cp = (void*)va;
    printf("compressed : fill256 %.4f ms  memset %.4f ms", time_ms(l_fill, cp, bytes, 20), time_ms(l_memset, cp, bytes, 20));
    l_fill(cp, bytes);
    printf("  read(zeros) %.4f ms", time_ms(l_read, cp, bytes, 20));
    l_rand(cp, bytes);
    printf("  read(random) %.4f ms", time_ms(l_read, cp, bytes, 20));
    printf("  fill256 over random %.4f ms\n", time_ms(l_fill, cp, bytes, 20));
    // concurrent: zero fill of the compressible buffer while another stream streams the plain buffer (read-bound kernel)
    cudaStream_t s1, s2;
    cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaEvent_t a, b, c;
    cudaEventCreate(&a); cudaEventCreate(&b); cudaEventCreate(&c);
    for (int which = 0; which < 2; ++which) {
        void* tgt = which ? cp : plain;
        void* src = which ? plain : cp;          // read the other one (both hold data now; make the read source random)
        l_rand(src, bytes);
        cudaDeviceSynchronize();
        cudaEventRecord(a, s1);
        cudaStreamWaitEvent(s2, a, 0);
        for (int i = 0; i < 10; ++i) {
            read_sum<<<148 * 16, 256, 0, s1>>>((const uint4*)src, bytes / 16, g_out);
            fill256<<<(unsigned)((bytes / 32 + 255) / 256), 256, 0, s2>>>((uint4*)tgt, bytes / 32);
        }
        cudaEventRecord(c, s2);
        cudaStreamWaitEvent(s1, c, 0);
        cudaEventRecord(b, s1);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        printf("concurrent read(random, %s) + fill(%s): %.4f ms per pair\n", which ? "plain" : "compressed", which ? "compressed" : "plain", ms / 10);
    }
    return 0;
}
