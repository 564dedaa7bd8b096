// Probe (not part of the library): how fast can 148 persistent CTAs stream a [B][D][S] bf16 tensor through TMA boxes of
// 128 rows x 64 pixels (SWIZZLE_128B) with nothing consuming the data?  Separates "the load pattern" from "the consumers"
// of proto_tc_kernel.   mode 0: a stage = all dim-boxes of one 64-px step (what proto_tc does)
//                       mode 1: a stage = NDB consecutive 64-px steps of one dim-box (long runs per feature row)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probe/tma_stream scripts/probe/tma_stream.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap tmap, int S, int D, int B, int NST,
                                                        int NDB, int mode, int box_bytes, unsigned long long* sink) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full[8];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int tiles = B * (S / 1024);                // 1024-px tiles like proto_tc
    const int steps_per_tile = mode == 0 ? 16 : 16 / NDB * NDB;   // mode 1: 16/NDB groups x NDB dim-boxes
    uint32_t issued = 0, waited = 0;
    unsigned long long acc = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int b = t / (S / 1024), s0 = (t % (S / 1024)) * 1024;
        for (int st = 0; st < steps_per_tile; ++st) {
            if (issued - waited == (uint32_t)NST) {          // ring full: retire the oldest stage
                const int ws = waited % NST;
                bar_wait(&full[ws], (waited / NST) & 1);
                acc += *(volatile unsigned int*)(base + (size_t)ws * NDB * box_bytes);
                ++waited;
            }
            const int s = issued % NST;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(NDB * box_bytes) : "memory");
            for (int k = 0; k < NDB; ++k) {
                int px, row;
                if (mode == 0) { px = s0 + st * 64; row = k * 128; }
                else { px = s0 + ((st / NDB) * NDB + k) * 64; row = (st % NDB) * 128; }
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(s32(base + ((size_t)s * NDB + k) * box_bytes)), "l"((uint64_t)&tmap), "r"(s32(&full[s])),
                               "r"(px), "r"(row), "r"(b) : "memory");
            }
            ++issued;
        }
    }
    while (waited < issued) { bar_wait(&full[waited % NST], (waited / NST) & 1); ++waited; }
    if (acc == 0x1234567) *sink = acc;
}

int main(int argc, char** argv) {
    const int S = 65536, D = argc > 1 ? atoi(argv[1]) : 496, B = 24;
    const int NDB = (D + 127) / 128;
    void* x;
    cudaMalloc(&x, (size_t)B * D * S * 2);
    cudaMemset(x, 0, (size_t)B * D * S * 2);
    unsigned long long* sink;
    cudaMalloc(&sink, 8);
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    Enc enc = (Enc)ptr;
    for (int promo = 0; promo < 3; ++promo)
        for (int mode = 0; mode < 2; ++mode)
            for (int nst = 2; nst <= 3; ++nst) {
                CUtensorMap map;
                const cuuint64_t gdim[3] = {(cuuint64_t)S, (cuuint64_t)D, (cuuint64_t)B};
                const cuuint64_t gstr[2] = {(cuuint64_t)S * 2, (cuuint64_t)S * D * 2};
                const cuuint32_t box[3] = {64, 128, 1};
                const cuuint32_t estr[3] = {1, 1, 1};
                const CUtensorMapL2promotion pr[3] = {CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B};
                CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, pr[promo], CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
                const int box_bytes = 128 * 128;
                const size_t smem = (size_t)nst * NDB * box_bytes + 1024;
                cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                float best = 1e9f;
                for (int it = 0; it < 5; ++it) {
                    cudaEventRecord(e0);
                    stream_kernel<<<148, 128, smem>>>(map, S, D, B, nst, NDB, mode, box_bytes, sink);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    if (it > 0 && ms < best) best = ms;
                }
                cudaError_t err = cudaGetLastError();
                const double gb = (double)B * D * S * 2 / 1e9;
                printf("D=%d promo=%d mode=%d stages=%d  %.3f ms  %.0f GB/s  %s\n", D, promo, mode, nst, best, gb / (best * 1e-3),
                       err == cudaSuccess ? "" : cudaGetErrorString(err));
            }
    return 0;
}
