// Probe: request rate of row gathers through the TMA unit, the way infonce_mma_kernel stages its negatives
// (1024 CTAs x 128 threads, 7 per SM, 32 chunks of 16 rows of 992 B, one stage per CTA, no math).
//   mode 0: one cp.async.bulk (UBLKCP) per row, issued by lanes 0-15 of warp 0            (what the kernel does)
//   mode 1: the same 16 copies issued by lanes 0-3 of all four warps
//   mode 2: cp.async.bulk.tensor.2d tile::gather4 -- 4 rows per request, two column halves (inner box <= 256 elements):
//           8 requests per chunk instead of 16
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_gather_rate scripts/probe/tma_gather_rate.cu -lcuda && /tmp/tma_gather_rate
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void bar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void gather4(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int r0, int r1, int r2, int r3) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(s32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(s32(bar)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

constexpr int ROWS = 30000, D = 496, ROWB = D * 2, KEYS = 16, CHUNKS = 32;

template <int MODE>
__global__ void __launch_bounds__(128, 7) probe(const __grid_constant__ CUtensorMap map, const unsigned short* __restrict__ bank, uint32_t* out) {
    extern __shared__ __align__(128) unsigned char stage[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { bar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    uint32_t acc = 0;
    for (int c = 0; c < CHUNKS; ++c) {
        if (tid == 0) bar_expect_tx(&bar, KEYS * ROWB);
        __syncthreads();
        if (MODE == 0) {
            if (warp == 0 && lane < KEYS) {
                const uint32_t r = hash(blockIdx.x * 7919u + c * 131u + lane) % ROWS;
                bulk_g2s(stage + lane * 1008, bank + (size_t)r * D, ROWB, &bar);
            }
        } else if (MODE == 1) {
            if (lane < 4) {
                const int k = warp * 4 + lane;
                const uint32_t r = hash(blockIdx.x * 7919u + c * 131u + k) % ROWS;
                bulk_g2s(stage + k * 1008, bank + (size_t)r * D, ROWB, &bar);
            }
        } else {
            if (warp == 0 && lane < 8) {
                const int g = lane >> 1, h = lane & 1;
                uint32_t r[4];
                for (int i = 0; i < 4; ++i) r[i] = hash(blockIdx.x * 7919u + c * 131u + g * 4 + i) % ROWS;
                gather4(stage + (g * 2 + h) * 1984, &map, &bar, h * 248, (int)r[0], (int)r[1], (int)r[2], (int)r[3]);
            }
        }
        bar_wait(&bar, (uint32_t)(c & 1));
        acc += reinterpret_cast<const uint32_t*>(stage)[tid * 7 % 4000];
        __syncthreads();
    }
    if (acc == 0x12345678u) out[0] = acc;
    if (MODE == 2 && blockIdx.x == 0 && tid == 0) {
        // correctness of the last chunk's first group: row r[i], columns 0..247 at stage + i*496 (half 0), 248.. at +1984 (half 1)
        const int c = CHUNKS - 1;
        uint32_t ok = 1;
        for (int i = 0; i < 4; ++i) {
            const uint32_t r = hash(blockIdx.x * 7919u + c * 131u + i) % ROWS;
            const unsigned short* h0 = reinterpret_cast<const unsigned short*>(stage + i * 496);
            const unsigned short* h1 = reinterpret_cast<const unsigned short*>(stage + 1984 + i * 496);
            for (int col = 0; col < 248; ++col) {
                ok &= h0[col] == bank[(size_t)r * D + col];
                ok &= h1[col] == bank[(size_t)r * D + 248 + col];
            }
        }
        out[1] = ok;
    }
}

int main() {
    unsigned short* bank;
    uint32_t* out;
    CHECK(cudaMalloc(&bank, (size_t)ROWS * ROWB));
    CHECK(cudaMalloc(&out, 16));
    CHECK(cudaMemset(out, 0, 16));
    unsigned short* h = (unsigned short*)malloc((size_t)ROWS * ROWB);
    for (size_t i = 0; i < (size_t)ROWS * D; ++i) h[i] = (unsigned short)((i * 2654435761u) >> 17);
    CHECK(cudaMemcpy(bank, h, (size_t)ROWS * ROWB, cudaMemcpyHostToDevice));
    CUtensorMap map;
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        const cuuint64_t gdim[2] = {D, ROWS};
        const cuuint64_t gstr[1] = {ROWB};
        const cuuint32_t box[2] = {248, 1};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = ((Enc)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, bank, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode (box 248 x 1): %d\n", (int)r);
    }
    const size_t smem = 16 * 1024 + 128;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](int mode) {
        for (int it = 0; it < 6; ++it) {
            if (it == 1) cudaEventRecord(e0);
            if (mode == 0) probe<0><<<1024, 128, smem>>>(map, bank, out);
            else if (mode == 1) probe<1><<<1024, 128, smem>>>(map, bank, out);
            else probe<2><<<1024, 128, smem>>>(map, bank, out);
        }
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        uint32_t ho[4];
        cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
        const double bytes = 1024.0 * CHUNKS * KEYS * ROWB;
        printf("mode %d: %s  %.1f us / launch, %.2f TB/s, gather4 content ok = %u\n", mode, cudaGetErrorString(e), ms / 5 * 1e3, bytes / (ms / 5 * 1e-3) / 1e12, ho[1]);
    };
    CHECK(cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CHECK(cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CHECK(cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    run(0); run(1); run(2);
    return 0;
}
