// Sub-systems 2 + 3a for bf16 representations on the 5th-generation tensor cores.
//
//   proto_sum[c][d] = sum_px onehot[c][px] * X[d][px]          (reference: torch.mean(rep_teacher[low_valid_c]),
//                                                                loss_helper_3d.py:380-384)
// is a GEMM with M = feature dims, N = classes, K = pixels, and X [B, D, S] is already K-major (pixels are
// contiguous).  So:  TMA (cp.async.bulk.tensor, SWIZZLE_128B) streams 128-row x 64-pixel bf16 boxes of X
// straight into shared memory, a 16-class x 64-pixel one-hot tile is built from the code bytes, and one thread
// issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) into TMEM accumulators that live for the CTA's whole
// life.  No CUDA-core instruction touches a feature element; products are exact (one-hot x bf16) and the
// accumulation is fp32, so results agree with the CUDA-core path to fp32 rounding.  Negative keys
// (loss_helper_3d.py:403-411) are copied out of the same shared boxes before the stage is released.
//
// Roles (256 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 one-hot builder +
// key enqueue; warps 0-3 run the TMEM -> partial-row epilogue.  3-stage mbarrier ring, 192 KB of loads in flight per SM.
#include <cuda.h>
#include <stdlib.h>

#include "arco_common.cuh"
#include "tc_common.cuh"
#include "proto_tail.cuh"

namespace arco {

constexpr int TC_KPX = 64;            // pixels per stage (one 128-byte swizzle row of bf16)
constexpr int TC_ROWS = 128;          // feature rows per TMA box / per MMA (M)
constexpr int TC_NCLS = 16;           // N of the MMA (classes, padded)
constexpr int TC_BOX_BYTES = TC_ROWS * TC_KPX * 2;      // 16 KB
constexpr int TC_B_BYTES = TC_NCLS * TC_KPX * 2;        // 2 KB one-hot tile
constexpr int TC_MAX_DB = 4;                            // D <= 512

struct ProtoTcParams {
    const uint8_t* codes;
    const uint32_t* tile_flagged;
    const uint32_t* off_key;
    const arco_plan* plan;
    void* bank_rows;
    int32_t bank_bf16;
    float* partials;
    double* proto_sums;          // [C][D+1] fp64, written by the in-kernel finalize (proto_tail.cuh)
    int64_t row_off[ARCO_MAX_CLASSES];
    int32_t cap[ARCO_MAX_CLASSES];
    int64_t S;
    int32_t B, C, D, tpi, NT, NDB;
};

// instruction descriptor: D fp32, A/B bf16, both K-major, N = 16, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((TC_NCLS >> 3) << 17) | ((TC_ROWS >> 4) << 24);

// byte offset of (feature row r, pixel kp) inside a SWIZZLE_128B box of 128 rows x 64 bf16
__device__ __forceinline__ uint32_t box_off(uint32_t r, uint32_t kp) {
    return (r >> 3) * 1024 + (r & 7) * 128 + (((kp >> 3) ^ (r & 7)) << 4) + ((kp & 7) << 1);
}

#ifdef ARCO_TC_TRACE
// debug build only: per-role clock64 stamps of CTA 0's first 256 steps (read back with arco_debug_tc_trace)
__device__ long long g_tc_trace[8][256];
__device__ unsigned long long g_tc_cta[4][160];        // globaltimer ns per CTA: entry, first full, last commit seen, exit
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TC_CTA(k) do { if (blockIdx.x < 160) g_tc_cta[k][blockIdx.x] = gtime(); } while (0)
#define TC_STAMP(role, it) do { if (blockIdx.x == 0 && (it) < 256) g_tc_trace[role][it] = clock64(); } while (0)
#else
#define TC_STAMP(role, it) do { } while (0)
#define TC_CTA(k) do { } while (0)
#endif

// A pipeline stage is one 64-pixel step with ALL feature blocks (NDB boxes of 128 rows x 128 B) plus its own one-hot
// tile; two builder warps (one thread per pixel) write the tile and the step's key list, two copier warps enqueue the
// key rows from the staged boxes, so the builders run up to NST stages ahead of the loads.  (Measured on B200: staging 128 rows x 256
// pixels per stage instead -- longer runs per row, fewer rows -- is 17 % slower, 0.434 vs 0.371 ms on the bf16
// D=496 shape; spreading every stage over all D rows keeps all HBM channels busy.)
__global__ void __launch_bounds__(256, 1) proto_tc_kernel(const __grid_constant__ CUtensorMap tmap, ProtoTcParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int NDB = p.NDB;
    const int stage_bytes = NDB * TC_BOX_BYTES + TC_B_BYTES;          // A boxes then the one-hot tile, 1024-aligned
    constexpr int NST = 3;
    __shared__ __align__(8) uint64_t full_bar[NST], bfull_bar[NST], kfull_bar[NST], empty_bar[NST], done_bar;
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_run[ARCO_MAX_CLASSES];
    __shared__ uint32_t s_keys[NST][TC_KPX];
    __shared__ uint32_t s_nkeys[NST];
    __shared__ __align__(16) uint8_t s_codes[ARCO_TILE];
    __shared__ int32_t s_skip[ARCO_MAX_CLASSES], s_base[ARCO_MAX_CLASSES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ngrid = gridDim.x;
    if (tid == 0) TC_CTA(0);
    const int64_t S = p.S;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { bar_init(&full_bar[s], 1); bar_init(&bfull_bar[s], 1); bar_init(&kfull_bar[s], 1); bar_init(&empty_bar[s], 2); }
        bar_init(&done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < ARCO_MAX_CLASSES) { s_skip[tid] = p.plan->bank_skip[tid]; s_base[tid] = p.plan->bank_write_base[tid]; }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&s_tmem)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    auto next_tile = [&](int t) {
        while (t < p.NT && p.tile_flagged[t] == 0) t += ngrid;
        return t;
    };
    // 64-pixel steps of tile t that hold a low-valid or key pixel (classify.cu writes one bit per 32-pixel group; bit 2*st of
    // the result stands for step st).  Steps without one are never loaded: on real images the entropy masks are spatially
    // coherent and whole runs of an unlabelled image drop out; pixels past the end of an image are never flagged.
    auto step_mask = [&](int t) {
        const uint32_t f = p.tile_flagged[t];
        return (f | (f >> 1)) & 0x55555555u;
    };

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = next_tile(blockIdx.x); t < p.NT; t = next_tile(t + ngrid)) {
                const int b = t / p.tpi;
                const int64_t s_tile = (int64_t)(t % p.tpi) * ARCO_TILE;
                for (uint32_t m = step_mask(t); m; m &= m - 1, ++it) {
                    const int st = (__ffs(m) - 1) >> 1;
                    const int s = it % NST;
                    bar_wait(&empty_bar[s], ((it / NST) & 1) ^ 1);
                    TC_STAMP(0, it);
                    bar_expect_tx(&full_bar[s], (uint32_t)NDB * TC_BOX_BYTES);
                    unsigned char* dst = base + (size_t)s * stage_bytes;
                    for (int db = 0; db < NDB; ++db)
                        tma_load_3d(dst + db * TC_BOX_BYTES, &tmap, &full_bar[s], (int)(s_tile + st * TC_KPX), db * TC_ROWS, b);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = next_tile(blockIdx.x); t < p.NT; t = next_tile(t + ngrid)) {
                for (uint32_t m = step_mask(t); m; m &= m - 1, ++it) {
                    const int s = it % NST;
                    const uint32_t ph = (it / NST) & 1;
                    bar_wait(&full_bar[s], ph);
                    TC_STAMP(1, it);
                    if (it == 0) TC_CTA(1);
                    bar_wait(&bfull_bar[s], ph);
                    TC_STAMP(2, it);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = s32(base + (size_t)s * stage_bytes);
                    const uint32_t b0 = a0 + NDB * TC_BOX_BYTES;
                    for (int db = 0; db < NDB; ++db) {
#pragma unroll
                        for (int kk = 0; kk < TC_KPX / 16; ++kk)
                            umma_bf16(tmem + db * TC_NCLS, umma_desc(a0 + db * TC_BOX_BYTES + kk * 32), umma_desc(b0 + kk * 32),
                                      kIdesc, (it > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);
                    TC_STAMP(3, it);
                }
            }
            umma_commit(&done_bar);
        }
    } else if (warp == 4 || warp == 5) {
        // ---- builders: one thread per pixel of the step; one-hot tile for the MMA, key list for the copiers ----
        const int at = tid - 128;
        uint32_t it = 0;
        int t = next_tile(blockIdx.x);
        uint4 pre = make_uint4(0u, 0u, 0u, 0u);                  // this thread's 16 codes of the tile, fetched one tile ahead
        auto fetch = [&](int tt) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (tt < p.NT) {
                const int64_t s0 = (int64_t)(tt % p.tpi) * ARCO_TILE + 16 * at;
                const uint8_t* src = p.codes + (int64_t)(tt / p.tpi) * S + s0;
                if (s0 + 16 <= S && ((S & 15) == 0)) v = *reinterpret_cast<const uint4*>(src);
                else {
                    uint32_t w[4] = {0u, 0u, 0u, 0u};
                    for (int i = 0; i < 16; ++i)
                        if (s0 + i < S) w[i >> 2] |= (uint32_t)src[i] << (8 * (i & 3));
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            return v;
        };
        pre = fetch(t);
        for (; t < p.NT;) {
            asm volatile("bar.sync 1, 64;" ::: "memory");        // every builder is done with the previous tile's codes
            reinterpret_cast<uint4*>(s_codes)[at] = pre;
            if (at < p.C) s_run[at] = p.off_key[(int64_t)at * (p.NT + 1) + t];
            const int t_next = next_tile(t + ngrid);
            pre = fetch(t_next);
            asm volatile("bar.sync 1, 64;" ::: "memory");
            for (uint32_t m = step_mask(t); m; m &= m - 1, ++it) {
                const int st = (__ffs(m) - 1) >> 1;
                const int s = it % NST;
                const uint32_t ph = (it / NST) & 1;
                unsigned char* btile = base + (size_t)s * stage_bytes + NDB * TC_BOX_BYTES;
                const uint32_t code = s_codes[st * TC_KPX + at];
                bar_wait(&empty_bar[s], ph ^ 1);                 // MMAs and key copies of the stage's previous use are done
                if (at == 0) TC_STAMP(4, it);
                reinterpret_cast<uint4*>(btile)[at] = make_uint4(0u, 0u, 0u, 0u);
                reinterpret_cast<uint4*>(btile)[at + 64] = make_uint4(0u, 0u, 0u, 0u);
                if (at == 0) s_nkeys[s] = 0;
                asm volatile("bar.sync 1, 64;" ::: "memory");
                if (code & CODE_LV) {
                    const uint32_t n = code & CODE_CLS_MASK;
                    *reinterpret_cast<unsigned short*>(btile + box_off(n, (uint32_t)at)) = 0x3F80;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 64;" ::: "memory");
                if (at == 0) bar_arrive(&bfull_bar[s]);
                // FIFO ordinal of every key pixel of the step (warp 4 owns pixels 0-31, warp 5 the next 32)
                const bool is_key = code & CODE_KEY;
                const uint32_t kcls = code & CODE_CLS_MASK;
                uint32_t ord = 0;
                const uint32_t peers = __match_any_sync(0xffffffffu, is_key ? kcls : 0xffffu);
                const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
                if (warp == 4) {
                    if (is_key) ord = s_run[kcls] + rank;
                    __syncwarp();
                    if (is_key && rank == 0) s_run[kcls] += __popc(peers);
                }
                asm volatile("bar.sync 1, 64;" ::: "memory");
                if (warp == 5) {
                    if (is_key) ord = s_run[kcls] + rank;
                    __syncwarp();
                    if (is_key && rank == 0) s_run[kcls] += __popc(peers);
                }
                if (is_key && ord >= (uint32_t)s_skip[kcls]) {                // not evicted within this very call
                    const uint32_t cap = (uint32_t)p.cap[kcls];
                    const uint32_t pos = ((uint32_t)s_base[kcls] + ord % cap) % cap;
                    const uint32_t slot = atomicAdd(&s_nkeys[s], 1u);
                    s_keys[s][slot] = (((uint32_t)p.row_off[kcls] + pos) << 6) | (uint32_t)at;   // ring row | pixel
                }
                asm volatile("bar.sync 1, 64;" ::: "memory");
                if (at == 0) { bar_arrive(&kfull_bar[s]); TC_STAMP(5, it); }
            }
            t = t_next;
        }
    } else if (warp == 2 || warp == 3 || warp >= 6) {
        // ---- copiers (4 warps; warps 2-3 go on to the epilogue): enqueue the step's key rows from the staged boxes.
        // Thread ct owns feature rows ct, ct+128, ...: inside a box that is a fixed (row, swizzle phase), so a key costs
        // one address, NDB shared loads at +box stride and NDB coalesced 2-byte stores per thread.
        const int ct = (warp < 4 ? (warp - 2) * 32 : 64 + (warp - 6) * 32) + lane;
        const uint32_t t_off = (uint32_t)(ct >> 3) * 1024u + (uint32_t)(ct & 7) * 128u;
        uint32_t it = 0;
        for (int t = next_tile(blockIdx.x); t < p.NT; t = next_tile(t + ngrid)) {
            for (uint32_t m = step_mask(t); m; m &= m - 1, ++it) {
                const int s = it % NST;
                const uint32_t ph = (it / NST) & 1;
                const uint32_t stage_a = s32(base + (size_t)s * stage_bytes) + t_off;
                bar_wait(&kfull_bar[s], ph);
                if (ct == 0) TC_STAMP(6, it);
                const uint32_t nkeys = s_nkeys[s];
                if (nkeys) {
                    bar_wait(&full_bar[s], ph);
                    for (uint32_t k = 0; k < nkeys; ++k) {
                        const uint32_t e = s_keys[s][k];
                        const uint32_t kp = e & 63u;
                        const int64_t row = (int64_t)(e >> 6) * p.D;
                        const uint32_t src = stage_a + ((((kp >> 3) ^ (uint32_t)(ct & 7))) << 4) + ((kp & 7u) << 1);
                        unsigned short v[TC_MAX_DB];
#pragma unroll
                        for (int i = 0; i < TC_MAX_DB; ++i)
                            if (ct + i * TC_ROWS < p.D)
                                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v[i]) : "r"(src + (uint32_t)i * TC_BOX_BYTES));
#pragma unroll
                        for (int i = 0; i < TC_MAX_DB; ++i) {
                            const int d = ct + i * TC_ROWS;
                            if (d < p.D) {
                                if (p.bank_bf16) reinterpret_cast<unsigned short*>(p.bank_rows)[row + d] = v[i];
                                else reinterpret_cast<float*>(p.bank_rows)[row + d] = bf16_bits_to_float(v[i]);
                            }
                        }
                    }
                }
                asm volatile("bar.sync 2, 128;" ::: "memory");
                if (ct == 0) { bar_arrive(&empty_bar[s]); TC_STAMP(7, it); }
            }
        }
    }

    __syncwarp();
    if (warp < 4) {
        uint32_t n_it = 0;
        for (int t = next_tile(blockIdx.x); t < p.NT; t = next_tile(t + ngrid)) n_it += __popc(step_mask(t));
        if (n_it > 0) {
            bar_wait(&done_bar, 0);
            if (tid == 0) TC_CTA(2);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int db = 0; db < NDB; ++db) {
            uint32_t v[TC_NCLS];
#pragma unroll
            for (int c = 0; c < TC_NCLS; ++c) v[c] = 0u;
            if (n_it > 0) {
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + db * TC_NCLS;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            const int d = db * TC_ROWS + warp * 32 + lane;
            if (d < p.D) {
#pragma unroll
                for (int c = 0; c < TC_NCLS; ++c)
                    if (c < p.C) p.partials[((int64_t)blockIdx.x * p.C + c) * p.D + d] = __uint_as_float(v[c]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
    if (tid == 0) TC_CTA(3);
    proto_finalize_tail(p.partials, (int)gridDim.x, p.C, p.D, const_cast<arco_plan*>(p.plan), p.proto_sums);
}

// ---------------------------------------------------------------------------------------------------
bool proto_tc_supported(const arco_dims& d) {
    return d.rep_dtype == ARCO_BF16 && d.classes <= TC_NCLS && d.feat <= TC_MAX_DB * TC_ROWS && d.space % 8 == 0 &&
           d.feat >= 64;
}

size_t proto_tc_smem(const arco_dims& d) {
    const int ndb = (d.feat + TC_ROWS - 1) / TC_ROWS;
    return (size_t)3 * (ndb * TC_BOX_BYTES + TC_B_BYTES) + 1024;
}

int launch_proto_tc(const arco_dims& d, const void* rep_teacher, const arco_bank* bank, const arco_ws_layout& L, char* ws,
                    int rows, double* proto_sums, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    ARCO_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    for (int c = 0; c < d.classes; ++c)
        ARCO_REQUIRE(bank->row_off[c] + bank->cap[c] < (1ll << 26), "memory bank too large for the packed key list (2^26 rows)");
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)d.space, (cuuint64_t)d.feat, (cuuint64_t)(d.n_lab + d.n_unlab)};
    const cuuint64_t gstr[2] = {(cuuint64_t)d.space * 2, (cuuint64_t)d.space * d.feat * 2};
    const cuuint32_t box[3] = {TC_KPX, TC_ROWS, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(rep_teacher), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return ARCO_ERR_CUDA;
    }
    ProtoTcParams p;
    p.codes = (const uint8_t*)(ws + L.codes);
    p.tile_flagged = (const uint32_t*)(ws + L.tile_flagged);
    p.off_key = (const uint32_t*)(ws + L.off_key);
    p.plan = (const arco_plan*)(ws + L.plan);
    p.bank_rows = bank->rows;
    p.bank_bf16 = bank->row_dtype == ARCO_BF16;
    p.partials = (float*)(ws + L.partials);
    p.proto_sums = proto_sums;
    for (int c = 0; c < ARCO_MAX_CLASSES; ++c) { p.row_off[c] = bank->row_off[c]; p.cap[c] = bank->cap[c] > 0 ? bank->cap[c] : 1; }
    p.S = d.space; p.B = d.n_lab + d.n_unlab; p.C = d.classes; p.D = d.feat;
    p.tpi = L.tiles_per_image; p.NT = L.n_tiles; p.NDB = (d.feat + TC_ROWS - 1) / TC_ROWS;
    const size_t smem = proto_tc_smem(d);
    ARCO_CUDA_CHECK(cudaFuncSetAttribute(proto_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    proto_tc_kernel<<<rows, 256, smem, st>>>(map, p);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

}  // namespace arco

#ifdef ARCO_TC_TRACE
extern "C" __attribute__((visibility("default"))) int arco_debug_tc_trace(long long* out) {
    cudaError_t e = cudaMemcpyFromSymbol(out, arco::g_tc_trace, sizeof(long long) * 8 * 256);
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out + 8 * 256, arco::g_tc_cta, sizeof(long long) * 4 * 160);
    return (int)e;
}
#endif
