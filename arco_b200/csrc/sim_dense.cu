// Config-5 dense similarity (section 8(b) "arco_infonce_fwd_dense").
//
// The reference pairs every query with its own N sampled bank rows (loss_helper_3d.py:466-486): Q*N row gathers of D
// elements, 0.5 flop/B.  When Q*N is much larger than the bank (M rows) the same logits are cheaper as a dense GEMM
//     S[q][m] = a_hat_q . k_m        ([Q, D] x [D, M], every bank row read once per 128 queries)
// followed by a SCALAR gather  logits[q][n] = S[q][row(idx[q][n])] / |k_row|.
//
// Precision: a_hat is split into three bf16 terms (hi + mid + lo carry all 24 mantissa bits), the ring rows are bf16
// already, bf16 x bf16 products are exact and tcgen05 accumulates in fp32 -- the three terms are accumulated into the SAME
// TMEM accumulator by running the K loop over [a_hi | a_mid | a_lo] against [k | k | k], so the result matches the FFMA
// gather form to fp32 summation order.
//
// Kernel: persistent, 1 CTA/SM, output tile 128 queries x 256 ring rows, K blocks of 64 (one 128-byte swizzle row).
//   warp 0  TMA producer: per K block one 256-row box of the ring + three 128-row boxes of the split anchors (80 KB stage, x2)
//   warp 1  tcgen05.mma issuer (M128 N256 K16, 12 per stage), accumulators double-buffered in TMEM (2 x 256 columns)
//   warp 2  TMEM allocator
//   warps 4-7 epilogue: tcgen05.ld -> scale by 1/|k_m| -> St[m][q] (ring-row major, so a warp stores 128 contiguous bytes)
// Forward only: this is the crossover study of config 5, not part of the training step.
#include <cuda.h>
#include <stdlib.h>

#include "arco_common.cuh"
#include "tc_common.cuh"

namespace arco {

constexpr int SD_BM = 128, SD_BN = 256, SD_BK = 64, SD_NST = 2;
constexpr int SD_MAX_SPLITS = 64;
constexpr int SD_A_BYTES = SD_BM * SD_BK * 2;               // 16 KB per anchor term
constexpr int SD_B_BYTES = SD_BN * SD_BK * 2;               // 32 KB
constexpr int SD_STAGE = SD_B_BYTES + 3 * SD_A_BYTES;       // 80 KB

// anchors [Q][D] f32 -> unit rows split into three bf16 terms, A3[term][q][d]
__global__ void __launch_bounds__(128) sim_prep_kernel(const float* __restrict__ anchors, unsigned short* __restrict__ a3, int Q,
                                                       int D) {
    const int q = blockIdx.x, tid = threadIdx.x;
    __shared__ float s_red[4];
    const float* row = anchors + (int64_t)q * D;
    float n2 = 0.f;
    for (int d = tid; d < D; d += 128) n2 += row[d] * row[d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = n2;
    __syncthreads();
    const float inv = 1.f / fmaxf(sqrtf(s_red[0] + s_red[1] + s_red[2] + s_red[3]), 1e-8f);
    for (int d = tid; d < D; d += 128) {
        const float x = row[d] * inv;
        const float hi = __bfloat162float(__float2bfloat16_rn(x));
        const float r1 = x - hi;
        const float mid = __bfloat162float(__float2bfloat16_rn(r1));
        const float lo = __bfloat162float(__float2bfloat16_rn(r1 - mid));
        a3[((int64_t)0 * Q + q) * D + d] = (unsigned short)(__float_as_uint(hi) >> 16);
        a3[((int64_t)1 * Q + q) * D + d] = (unsigned short)(__float_as_uint(mid) >> 16);
        a3[((int64_t)2 * Q + q) * D + d] = (unsigned short)(__float_as_uint(lo) >> 16);
    }
}

// 1 / max(|k|, eps) of every physical ring row of one class (bf16 rows); one warp per row
__global__ void __launch_bounds__(256) sim_row_norm_kernel(const unsigned short* __restrict__ rows, float* __restrict__ inv_nk,
                                                           int cap, int D) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= cap) return;
    const uint4* p = reinterpret_cast<const uint4*>(rows + (int64_t)r * D);
    float n2 = 0.f;
    for (int c = lane; c < D / 8; c += 32) {
        const uint4 u = p[c];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float a = __uint_as_float(w[i] << 16), b = __uint_as_float(w[i] & 0xffff0000u);
            n2 += a * a + b * b;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    if (lane == 0) inv_nk[r] = 1.f / fmaxf(sqrtf(n2), 1e-8f);
}

struct SimDenseParams {
    const float* inv_nk;       // forward: [cap]
    float* st;                 // forward: [cap][Q] cosines, ring-row major; backward: [Q][D] d loss / d a_hat
    int32_t Q, D, cap, n_mt, n_nt;
    int32_t kext;              // extent of the contraction: D (forward), padded ring rows (backward)
    int32_t grad;              // 0 forward epilogue (scale by 1/|k_m|, transposed store), 1 backward epilogue (plain [q][d] store)
    int32_t splits;            // split-K: work item = (output tile, K range); backward only (tiny Q x D output, K = ring rows)
};

__global__ void __launch_bounds__(256, 1) sim_dense_kernel(const __grid_constant__ CUtensorMap map_a,
                                                           const __grid_constant__ CUtensorMap map_b, SimDenseParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[SD_NST], empty_bar[SD_NST], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_tiles = p.n_mt * p.n_nt * p.splits;               // work items: (tile, K split)
    const int KB = (p.kext + SD_BK - 1) / SD_BK;
    constexpr uint32_t kIdescSD = umma_idesc_bf16(SD_BM, SD_BN);

    if (tid == 0) {
        for (int s = 0; s < SD_NST; ++s) { bar_init(&full_bar[s], 1); bar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { bar_init(&tfull_bar[a], 1); bar_init(&tempty_bar[a], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < n_tiles; item += gridDim.x) {
                const int tile = item / p.splits, sp = item % p.splits;
                const int nt = tile / p.n_mt, mt = tile % p.n_mt;         // the M tiles of one ring tile run side by side (L2)
                const int kb0 = (int)((int64_t)sp * KB / p.splits), kb1 = (int)((int64_t)(sp + 1) * KB / p.splits);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % SD_NST;
                    bar_wait(&empty_bar[s], ((it / SD_NST) & 1) ^ 1);
                    bar_expect_tx(&full_bar[s], (uint32_t)SD_STAGE);
                    unsigned char* dst = base + (size_t)s * SD_STAGE;
                    tma_load_2d(dst, &map_b, &full_bar[s], kb * SD_BK, nt * SD_BN);
                    for (int t = 0; t < 3; ++t)
                        tma_load_3d(dst + SD_B_BYTES + t * SD_A_BYTES, &map_a, &full_bar[s], kb * SD_BK, mt * SD_BM, t);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, ti = 0;
            for (int item = blockIdx.x; item < n_tiles; item += gridDim.x, ++ti) {
                const int sp = item % p.splits;
                const int kb0 = (int)((int64_t)sp * KB / p.splits), kb1 = (int)((int64_t)(sp + 1) * KB / p.splits);
                const uint32_t a = ti & 1;
                bar_wait(&tempty_bar[a], ((ti >> 1) & 1) ^ 1);            // the epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % SD_NST;
                    bar_wait(&full_bar[s], (it / SD_NST) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t b0 = s32(base + (size_t)s * SD_STAGE);
                    for (int t = 0; t < 3; ++t) {
                        const uint32_t a0 = b0 + SD_B_BYTES + t * SD_A_BYTES;
#pragma unroll
                        for (int kk = 0; kk < SD_BK / 16; ++kk)
                            umma_bf16(tmem + a * SD_BN, umma_desc(a0 + kk * 32), umma_desc(b0 + kk * 32), kIdescSD,
                                      (kb > kb0 || t > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tfull_bar[a]);
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;                                          // TMEM lanes 32*ew .. 32*ew+31
        uint32_t ti = 0;
        for (int item = blockIdx.x; item < n_tiles; item += gridDim.x, ++ti) {
            const int tile = item / p.splits, sp = item % p.splits;
            const int nt = tile / p.n_mt, mt = tile % p.n_mt;
            float* outp = p.st + (int64_t)sp * p.Q * p.D;                 // backward: partial sums of this K split
            const bool empty = (int)((int64_t)sp * KB / p.splits) == (int)((int64_t)(sp + 1) * KB / p.splits);   // no K block: accumulator never written
            const uint32_t a = ti & 1;
            bar_wait(&tfull_bar[a], (ti >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int q = mt * SD_BM + ew * 32 + lane;
            for (int c0 = 0; c0 < SD_BN; c0 += 16) {
                uint32_t v[16];
                const uint32_t taddr = tmem + ((uint32_t)(ew * 32) << 16) + a * SD_BN + c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int m0 = nt * SD_BN + c0;
                if (p.grad) {
                    // backward: column = feature d; a lane owns query q and stores 16 consecutive floats of its row
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (m0 + i < p.D && q < p.Q) outp[(int64_t)q * p.D + m0 + i] = empty ? 0.f : __uint_as_float(v[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int m = m0 + i;
                        if (m < p.cap && q < p.Q) p.st[(int64_t)m * p.Q + q] = __uint_as_float(v[i]) * __ldg(p.inv_nk + m);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid == 128) bar_arrive(&tempty_bar[a]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// logits[q][n] = St[ring row of idx[q][n]][q]
__global__ void __launch_bounds__(256) sim_gather_kernel(const float* __restrict__ st, const int32_t* __restrict__ idx,
                                                         const int32_t* __restrict__ head, const int32_t* __restrict__ len, int cls,
                                                         int cap, int Q, int N, float* __restrict__ logits) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (int64_t)Q * N) return;
    const int q = (int)(i / N);
    const int blen = len[cls], bhead = head[cls];
    int r = idx[i];
    r = min(max(r, 0), blen - 1);
    int phys = bhead + r;
    if (phys >= cap) phys -= cap;
    logits[i] = st[(int64_t)phys * Q + q];
}


// ---- backward of the dense similarity (the gradient the gather form gets for free) -------------------------------------------
//   d loss / d a_hat[q] = sum_n g[q][n] * k_hat[row(idx[q][n])]  =  (Wd . Ring)[q],   Wd[q][m] = sum_{n: row(idx[q][n]) = m} g[q][n] / |k_m|
// i.e. a second dense GEMM [Q, M] x [M, D] on the same tcgen05 kernel: A = Wd split into three bf16 terms (K-major: ring rows
// contiguous), B = the ring TRANSPOSED to [D][M] (K-major), K = the padded ring length.
__global__ void __launch_bounds__(256) sim_wscatter_kernel(const float* __restrict__ g, const int32_t* __restrict__ idx,
                                                           const int32_t* __restrict__ head, const int32_t* __restrict__ len, int cls, int cap,
                                                           const float* __restrict__ inv_nk, int Q, int N, int Mpad, float* __restrict__ wd) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (int64_t)Q * N) return;
    const int q = (int)(i / N);
    const int blen = len[cls], bhead = head[cls];
    int r = idx[i];
    r = min(max(r, 0), blen - 1);
    int phys = bhead + r;
    if (phys >= cap) phys -= cap;
    atomicAdd(wd + (int64_t)q * Mpad + phys, g[i] * inv_nk[phys]);   // duplicates of a row accumulate (float atomics)
}

__global__ void __launch_bounds__(256) sim_wsplit_kernel(const float* __restrict__ wd, unsigned short* __restrict__ w3, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float x = wd[i];
    const float hi = __bfloat162float(__float2bfloat16_rn(x));
    const float r1 = x - hi;
    const float mid = __bfloat162float(__float2bfloat16_rn(r1));
    const float lo = __bfloat162float(__float2bfloat16_rn(r1 - mid));
    w3[i] = (unsigned short)(__float_as_uint(hi) >> 16);
    w3[n + i] = (unsigned short)(__float_as_uint(mid) >> 16);
    w3[2 * n + i] = (unsigned short)(__float_as_uint(lo) >> 16);
}

// ring rows [cap][D] bf16 -> [D][Mpad] bf16 (columns >= cap zero), 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) sim_transpose_kernel(const unsigned short* __restrict__ rows, unsigned short* __restrict__ out, int cap,
                                                            int D, int Mpad) {
    __shared__ unsigned short tile[32][33];
    const int m0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int m = m0 + r, d = d0 + tx;
        tile[r][tx] = (m < cap && d < D) ? rows[(int64_t)m * D + d] : (unsigned short)0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int d = d0 + r, m = m0 + tx;
        if (d < D && m < Mpad) out[(int64_t)d * Mpad + m] = tile[tx][r];
    }
}

// out[i] = sum over the K splits of part[s][i], fixed order (deterministic)
__global__ void __launch_bounds__(256) sim_reduce_splits_kernel(const float* __restrict__ part, float* __restrict__ out, int64_t n, int splits) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += part[(int64_t)s * n + i];
    out[i] = acc;
}

}  // namespace arco

static int64_t sd_align(int64_t x) { return (x + 255) & ~(int64_t)255; }

extern "C" int64_t arco_similarity_dense_scratch(int32_t feat, int32_t queries, int32_t n_slots, const arco_bank* bank,
                                                 const int32_t* slot_class) {
    if (!bank || !slot_class || feat <= 0 || queries <= 0 || n_slots <= 0) return -1;
    int64_t total = sd_align((int64_t)n_slots * 3 * queries * feat * 2);
    for (int j = 0; j < n_slots; ++j) {
        const int c = slot_class[j];
        if (c < 0 || c >= ARCO_MAX_CLASSES) return -1;
        total += sd_align((int64_t)bank->cap[c] * 4) + sd_align((int64_t)bank->cap[c] * queries * 4);
    }
    return total;
}

extern "C" int arco_similarity_dense(int32_t feat, int32_t queries, int32_t negatives, int32_t n_slots,
                                     const int32_t* slot_class, const float* anchors, const arco_bank* bank,
                                     const int32_t* idx_neg, float* logits, void* scratch, void* stream) {
    using namespace arco;
    ARCO_REQUIRE(slot_class && anchors && bank && idx_neg && logits && scratch, "arco_similarity_dense: NULL argument");
    ARCO_REQUIRE(bank->row_dtype == ARCO_BF16, "arco_similarity_dense needs a bf16 ring (bf16 x bf16 products are exact)");
    ARCO_REQUIRE(feat % 8 == 0 && feat >= 8 && queries % SD_BM == 0 && queries > 0 && negatives > 0 && n_slots > 0,
                 "arco_similarity_dense: D must be a multiple of 8, Q a multiple of 128");
    EncodeTiledFn enc = encode_fn();
    ARCO_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cudaStream_t st = (cudaStream_t)stream;
    const int Q = queries, D = feat, N = negatives;
    char* sc = (char*)scratch;
    unsigned short* a3 = (unsigned short*)sc;
    int64_t off = sd_align((int64_t)n_slots * 3 * Q * D * 2);
    const size_t smem = (size_t)SD_NST * SD_STAGE + 1024;
    ARCO_CUDA_CHECK(cudaFuncSetAttribute(sim_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 148;
    {
        int devid = 0;
        cudaGetDevice(&devid);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, devid);
    }
    for (int j = 0; j < n_slots; ++j) {
        const int c = slot_class[j];
        ARCO_REQUIRE(c >= 0 && c < ARCO_MAX_CLASSES && bank->cap[c] > 0, "arco_similarity_dense: bad slot class");
        const int cap = bank->cap[c];
        float* inv_nk = (float*)(sc + off);
        off += sd_align((int64_t)cap * 4);
        float* stt = (float*)(sc + off);
        off += sd_align((int64_t)cap * Q * 4);
        unsigned short* a3j = a3 + (int64_t)j * 3 * Q * D;
        const unsigned short* rows = (const unsigned short*)bank->rows + bank->row_off[c] * D;
        sim_prep_kernel<<<Q, 128, 0, st>>>(anchors + (int64_t)j * Q * D, a3j, Q, D);
        sim_row_norm_kernel<<<(cap + 7) / 8, 256, 0, st>>>(rows, inv_nk, cap, D);
        CUtensorMap map_a, map_b;
        {
            const cuuint64_t gdim[3] = {(cuuint64_t)D, (cuuint64_t)Q, 3};
            const cuuint64_t gstr[2] = {(cuuint64_t)D * 2, (cuuint64_t)Q * D * 2};
            const cuuint32_t box[3] = {SD_BK, SD_BM, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a3j, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (anchors) failed with CUresult %d", (int)r); return ARCO_ERR_CUDA; }
        }
        {
            const cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)cap};
            const cuuint64_t gstr[1] = {(cuuint64_t)D * 2};
            const cuuint32_t box[2] = {SD_BK, SD_BN};
            const cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<unsigned short*>(rows), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (ring) failed with CUresult %d", (int)r); return ARCO_ERR_CUDA; }
        }
        SimDenseParams p;
        p.inv_nk = inv_nk; p.st = stt; p.Q = Q; p.D = D; p.cap = cap; p.kext = D; p.grad = 0; p.splits = 1;
        p.n_mt = Q / SD_BM; p.n_nt = (cap + SD_BN - 1) / SD_BN;
        const int tiles = p.n_mt * p.n_nt;
        sim_dense_kernel<<<tiles < sms ? tiles : sms, 256, smem, st>>>(map_a, map_b, p);
        const int64_t total = (int64_t)Q * N;
        sim_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(stt, idx_neg + (int64_t)j * Q * N, bank->head, bank->len, c, cap,
                                                                         Q, N, logits + (int64_t)j * Q * N);
    }
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

static int64_t sd_mpad(int cap) { return ((int64_t)cap + 63) / 64 * 64; }

extern "C" int64_t arco_similarity_dense_backward_scratch(int32_t feat, int32_t queries, int32_t n_slots, const arco_bank* bank,
                                                          const int32_t* slot_class) {
    if (!bank || !slot_class || feat <= 0 || queries <= 0 || n_slots <= 0) return -1;
    int64_t worst = 0;
    for (int j = 0; j < n_slots; ++j) {
        const int c = slot_class[j];
        if (c < 0 || c >= ARCO_MAX_CLASSES) return -1;
        const int64_t mp = sd_mpad(bank->cap[c]);
        const int64_t need = sd_align((int64_t)bank->cap[c] * 4) + sd_align((int64_t)queries * mp * 4) + sd_align(3ll * queries * mp * 2) +
                             sd_align((int64_t)feat * mp * 2) + sd_align((int64_t)arco::SD_MAX_SPLITS * queries * feat * 4);
        if (need > worst) worst = need;
    }
    return worst;                                            // the slots run one after the other in the same scratch
}

extern "C" int arco_similarity_dense_backward(int32_t feat, int32_t queries, int32_t negatives, int32_t n_slots,
                                              const int32_t* slot_class, const float* grad_logits, const arco_bank* bank,
                                              const int32_t* idx_neg, float* grad_anchor_hat, void* scratch, void* stream) {
    using namespace arco;
    ARCO_REQUIRE(slot_class && grad_logits && bank && idx_neg && grad_anchor_hat && scratch, "arco_similarity_dense_backward: NULL argument");
    ARCO_REQUIRE(bank->row_dtype == ARCO_BF16, "arco_similarity_dense_backward needs a bf16 ring");
    ARCO_REQUIRE(feat % 8 == 0 && feat >= 8 && queries % SD_BM == 0 && queries > 0 && negatives > 0 && n_slots > 0,
                 "arco_similarity_dense_backward: D must be a multiple of 8, Q a multiple of 128");
    EncodeTiledFn enc = encode_fn();
    ARCO_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cudaStream_t st = (cudaStream_t)stream;
    const int Q = queries, D = feat, N = negatives;
    const size_t smem = (size_t)SD_NST * SD_STAGE + 1024;
    ARCO_CUDA_CHECK(cudaFuncSetAttribute(sim_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = sm_count();
    for (int j = 0; j < n_slots; ++j) {
        const int c = slot_class[j];
        ARCO_REQUIRE(c >= 0 && c < ARCO_MAX_CLASSES && bank->cap[c] > 0, "arco_similarity_dense_backward: bad slot class");
        const int cap = bank->cap[c];
        const int64_t mp = sd_mpad(cap);
        char* sc = (char*)scratch;
        float* inv_nk = (float*)sc;                 sc += sd_align((int64_t)cap * 4);
        float* wd = (float*)sc;                     sc += sd_align((int64_t)Q * mp * 4);
        unsigned short* w3 = (unsigned short*)sc;   sc += sd_align(3ll * Q * mp * 2);
        unsigned short* ring_t = (unsigned short*)sc; sc += sd_align((int64_t)D * mp * 2);
        float* partials = (float*)sc;
        const unsigned short* rows = (const unsigned short*)bank->rows + bank->row_off[c] * D;
        sim_row_norm_kernel<<<(cap + 7) / 8, 256, 0, st>>>(rows, inv_nk, cap, D);
        ARCO_CUDA_CHECK(cudaMemsetAsync(wd, 0, (size_t)Q * mp * 4, st));
        const int64_t total = (int64_t)Q * N;
        sim_wscatter_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(grad_logits + (int64_t)j * Q * N, idx_neg + (int64_t)j * Q * N,
                                                                           bank->head, bank->len, c, cap, inv_nk, Q, N, (int)mp, wd);
        const int64_t nw = (int64_t)Q * mp;
        sim_wsplit_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(wd, w3, nw);
        sim_transpose_kernel<<<dim3((unsigned)(mp / 32), (unsigned)((D + 31) / 32)), 256, 0, st>>>(rows, ring_t, cap, D, (int)mp);
        CUtensorMap map_a, map_b;
        {
            const cuuint64_t gdim[3] = {(cuuint64_t)mp, (cuuint64_t)Q, 3};
            const cuuint64_t gstr[2] = {(cuuint64_t)mp * 2, (cuuint64_t)Q * mp * 2};
            const cuuint32_t box[3] = {SD_BK, SD_BM, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, w3, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (weights) failed with CUresult %d", (int)r); return ARCO_ERR_CUDA; }
        }
        {
            const cuuint64_t gdim[2] = {(cuuint64_t)mp, (cuuint64_t)D};
            const cuuint64_t gstr[1] = {(cuuint64_t)mp * 2};
            const cuuint32_t box[2] = {SD_BK, SD_BN};
            const cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ring_t, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (transposed ring) failed with CUresult %d", (int)r); return ARCO_ERR_CUDA; }
        }
        SimDenseParams p;
        p.inv_nk = inv_nk; p.Q = Q; p.D = D; p.cap = cap;
        p.n_mt = Q / SD_BM; p.n_nt = (D + SD_BN - 1) / SD_BN; p.kext = (int)mp; p.grad = 1;
        const int tiles = p.n_mt * p.n_nt;
        // the output is tiny (Q x D) and the contraction long (ring rows): split K so that every SM gets a work item
        int splits = sms / tiles;
        const int kb_total = (int)(mp / SD_BK);
        if (splits > SD_MAX_SPLITS) splits = SD_MAX_SPLITS;
        if (splits > kb_total) splits = kb_total;
        if (splits < 1) splits = 1;
        p.splits = splits;
        float* gout = grad_anchor_hat + (int64_t)j * Q * D;
        p.st = splits > 1 ? partials : gout;
        const int items = tiles * splits;
        sim_dense_kernel<<<items < sms ? items : sms, 256, smem, st>>>(map_a, map_b, p);
        if (splits > 1) {
            const int64_t n = (int64_t)Q * D;
            sim_reduce_splits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partials, gout, n, splits);
        }
    }
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}
