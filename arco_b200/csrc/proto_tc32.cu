// Sub-systems 2 + 3a for fp32 representations on the 5th-generation tensor cores (TF32, exact two-term split).
//
//   proto_sum[c][d] = sum_px onehot[c][px] * X[d][px]          (loss_helper_3d.py:380-384)
// is the same GEMM as in proto_tc.cu (M = feature dims, N = classes, K = pixels, X [B, D, S] K-major), but tcgen05
// kind::tf32 reads only the top 19 bits of every fp32 operand.  The sum stays exact to fp32 accumulation by running
// TWO passes over each staged box:
//     pass hi : A = the box as TMA wrote it, read from SHARED memory (the hardware drops the low 13 mantissa bits:
//               hi = x & 0xFFFFE000)
//     pass lo : A = lo = x - hi (exact, <= 13 significant bits, so what TF32 drops of it is below 2^-22 |x|), which eight
//               converter warps compute from ONE conflict-free read of the box and store into TENSOR MEMORY
//               (tcgen05.st); the MMA takes its A operand from there (tcgen05.mma with A in TMEM)
// one-hot x fp32 products are exact, so hi + lo reproduces the fp32 sum up to accumulation order.
// (If the hardware rounded instead of truncating, hi + lo would be off by a TF32 ulp -- 5e-4 relative -- and the 2e-5
// prototype parity tests of tests/test_gpu_shapes.py would fail; they pass.)
//
// Why TMEM: at the HBM rate an SM must take in ~23 B/clk of boxes; round 1 rewrote every box in place for the lo pass
// (TMA write + hi read + converter read + converter write + lo read = 5x the box through shared memory = 115 of its
// 128 B/clk, ncu: 49 M bank conflicts, stage held ~1.9 us) and stalled at 0.51 of the HBM roofline.  With lo in TMEM the
// box crosses shared memory 3x (TMA write, hi read, converter read), it is never modified -- so keys are copied and the hi
// pass runs while the converters work -- and a stage is released as soon as both MMA passes have committed.
//
// Roles (384 threads): warp 0 TMA producer, warp 1 issues the hi passes, warp 2 the lo passes (own accumulators, summed
// in the epilogue), warp 3 builder (one thread per pixel of the 32-pixel step: one-hot tile, FIFO ordinals, key list),
// warps 4-11 converters (warp w owns TMEM lanes / feature rows 32*(w%4).. of boxes (w-4)/4 and +2) which also copy the
// step's keys out of the box; warps 4-7 run the TMEM -> partial-row epilogue.  Stage = all D rows of a 32-pixel step (NDB
// boxes of 128 rows x 128 B) + the one-hot tile; as many stages as fit (6 at D <= 256, 3 at D = 512).  TMEM: 512
// columns = hi + lo accumulators (2 x NDB x NCLS) + a ring of lo slots (NDB x 32 columns each: 6 at D = 256).
#include <cuda.h>
#include <stdlib.h>

#include "arco_common.cuh"
#include "tc_common.cuh"
#include "proto_tail.cuh"

namespace arco {

constexpr int T32_KPX = 32;            // pixels per stage (one 128-byte swizzle row of fp32)
constexpr int T32_ROWS = 128;          // feature rows per TMA box / per MMA (M)
constexpr int T32_BOX_BYTES = T32_ROWS * T32_KPX * 4;     // 16 KB
constexpr int T32_B_BYTES = 32 * T32_KPX * 4;             // 4 KB one-hot tile (up to N = 32 classes)
constexpr int T32_MAX_DB = 4;                             // D <= 512
constexpr int T32_MAX_ST = 6;
constexpr int T32_MAX_LO = 8;                             // lo operand slots in tensor memory

struct ProtoTc32Params {
    const uint8_t* codes;
    const uint32_t* tile_flagged;
    const uint32_t* off_key;
    const arco_plan* plan;
    float* bank_rows;
    float* partials;
    double* proto_sums;          // [C][D+1] fp64, written by the in-kernel finalize (proto_tail.cuh)
    int64_t row_off[ARCO_MAX_CLASSES];
    int32_t cap[ARCO_MAX_CLASSES];
    int64_t S;
    int32_t B, C, D, tpi, NT, NDB, NST, NCLS, NLO;       // NCLS = MMA N: 16 or 32; NLO = lo slots in TMEM
};

// same with the A operand in tensor memory (lane = row m, column = k; 8 columns per K = 8 step)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// byte offset of (row r, pixel kp) inside a SWIZZLE_128B box of rows x 32 fp32
__device__ __forceinline__ uint32_t box_off32(uint32_t r, uint32_t kp) {
    return (r >> 3) * 1024 + (r & 7) * 128 + (((kp >> 2) ^ (r & 7)) << 4) + ((kp & 3) << 2);
}

#ifdef ARCO_TC_TRACE
// debug build only: per-role clock64 stamps of CTA 0's first 512 steps (read back with arco_debug_tc32_trace)
__device__ long long g_tc32_trace[8][512];
#define T32_STAMP(role, it) do { if (blockIdx.x == 0 && (it) < 512) g_tc32_trace[role][it] = clock64(); } while (0)
#else
#define T32_STAMP(role, it) do { } while (0)
#endif

__global__ void __launch_bounds__(384, 1) proto_tc32_kernel(const __grid_constant__ CUtensorMap tmap, ProtoTc32Params p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int NDB = p.NDB, NST = p.NST, NLO = p.NLO;
    const int stage_bytes = NDB * T32_BOX_BYTES + T32_B_BYTES;
    __shared__ __align__(8) uint64_t full_bar[T32_MAX_ST], bfull_bar[T32_MAX_ST], kfull_bar[T32_MAX_ST], empty_bar[T32_MAX_ST],
        lofull_bar[T32_MAX_LO], loempty_bar[T32_MAX_LO], done_bar;
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_run[ARCO_MAX_CLASSES];
    __shared__ uint32_t s_keys[T32_MAX_ST][T32_KPX];
    __shared__ uint32_t s_nkeys[T32_MAX_ST];
    __shared__ __align__(16) uint8_t s_codes[ARCO_TILE];
    __shared__ int32_t s_skip[ARCO_MAX_CLASSES], s_base[ARCO_MAX_CLASSES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ngrid = gridDim.x;
    const int64_t S = p.S;
    const uint32_t idesc = umma_idesc_tf32(T32_ROWS, p.NCLS);
    // TMEM columns: [0, NDB*NCLS) hi accumulators, then the lo accumulators, then NLO slots of NDB*32 columns (lo operand)
    const uint32_t col_lo_acc = (uint32_t)(NDB * p.NCLS);
    const uint32_t col_slots = 2u * col_lo_acc;
    const uint32_t slot_cols = (uint32_t)(NDB * T32_KPX);

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            bar_init(&full_bar[s], 1); bar_init(&bfull_bar[s], 1); bar_init(&kfull_bar[s], 1);
            bar_init(&empty_bar[s], 2);                                  // hi commit + lo commit
        }
        for (int l = 0; l < NLO; ++l) { bar_init(&lofull_bar[l], 8); bar_init(&loempty_bar[l], 1); }
        bar_init(&done_bar, 2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < ARCO_MAX_CLASSES) { s_skip[tid] = p.plan->bank_skip[tid]; s_base[tid] = p.plan->bank_write_base[tid]; }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&s_tmem)), "r"(512));
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
    // 32-pixel steps of tile t that hold a low-valid or key pixel: exactly the bits classify.cu wrote.  Steps without one are
    // never loaded (coherent entropy masks drop whole runs of an unlabelled image); pixels past the image end are never flagged.
    auto step_mask = [&](int t) { return p.tile_flagged[t]; };
    auto total_steps = [&]() {
        uint32_t n = 0;
        for (int t = next_tile(blockIdx.x); t < p.NT; t = next_tile(t + ngrid)) n += __popc(step_mask(t));
        return n;
    };

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = next_tile(blockIdx.x); t < p.NT; t = next_tile(t + ngrid)) {
                const int b = t / p.tpi;
                const int64_t s_tile = (int64_t)(t % p.tpi) * ARCO_TILE;
                for (uint32_t m = step_mask(t); m; m &= m - 1, ++it) {
                    const int st = __ffs(m) - 1;
                    const int s = it % NST;
                    bar_wait(&empty_bar[s], ((it / NST) & 1) ^ 1);
                    T32_STAMP(0, it);
                    bar_expect_tx(&full_bar[s], (uint32_t)NDB * T32_BOX_BYTES);
                    unsigned char* dst = base + (size_t)s * stage_bytes;
                    for (int db = 0; db < NDB; ++db)
                        tma_load_3d(dst + db * T32_BOX_BYTES, &tmap, &full_bar[s], (int)(s_tile + st * T32_KPX), db * T32_ROWS, b);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // hi passes: A = the raw box in shared memory (the tensor core truncates to TF32), into accumulators [0, NDB*NCLS)
            const uint32_t total = total_steps();
            for (uint32_t k = 0; k < total; ++k) {
                const int s = k % NST;
                const uint32_t ph = (k / NST) & 1;
                bar_wait(&full_bar[s], ph);
                bar_wait(&bfull_bar[s], ph);
                T32_STAMP(1, k);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = s32(base + (size_t)s * stage_bytes);
                const uint32_t b0 = a0 + NDB * T32_BOX_BYTES;
                // feature block outer, K step inner (interleaving the accumulators measured 12 % slower)
                for (int db = 0; db < NDB; ++db) {
#pragma unroll
                    for (int kk = 0; kk < T32_KPX / 8; ++kk)
                        umma_tf32(tmem + db * p.NCLS, umma_desc(a0 + db * T32_BOX_BYTES + kk * 32), umma_desc(b0 + kk * 32), idesc,
                                  (k > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
                T32_STAMP(2, k);
            }
            umma_commit(&done_bar);
        }
    } else if (warp == 2) {
        if (lane == 0) {
            // lo passes: A = lo slot in tensor memory (written by the converters), B = the same one-hot tile
            const uint32_t total = total_steps();
            for (uint32_t k = 0; k < total; ++k) {
                const int s = k % NST, l = k % NLO;
                bar_wait(&lofull_bar[l], (k / NLO) & 1);
                bar_wait(&bfull_bar[s], (k / NST) & 1);
                T32_STAMP(3, k);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b0 = s32(base + (size_t)s * stage_bytes) + NDB * T32_BOX_BYTES;
                const uint32_t a_t = tmem + col_slots + (uint32_t)l * slot_cols;
                for (int db = 0; db < NDB; ++db) {
#pragma unroll
                    for (int kk = 0; kk < T32_KPX / 8; ++kk)
                        umma_tf32_ts(tmem + col_lo_acc + db * p.NCLS, a_t + db * T32_KPX + kk * 8, umma_desc(b0 + kk * 32), idesc,
                                     (k > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&loempty_bar[l]);
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&done_bar);
        }
    } else if (warp == 3) {
        // ---- builder: lane = pixel of the 32-pixel step ----
        uint32_t it = 0;
        int t = next_tile(blockIdx.x);
        auto fetch = [&](int tt, uint4& v0, uint4& v1) {                 // this lane's 32 codes of tile tt
            v0 = v1 = make_uint4(0u, 0u, 0u, 0u);
            if (tt < p.NT) {
                const int64_t s0 = (int64_t)(tt % p.tpi) * ARCO_TILE + 32 * lane;
                const uint8_t* src = p.codes + (int64_t)(tt / p.tpi) * S + s0;
                if (s0 + 32 <= S && ((S & 15) == 0)) {
                    v0 = reinterpret_cast<const uint4*>(src)[0];
                    v1 = reinterpret_cast<const uint4*>(src)[1];
                } else {
                    uint32_t w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                    for (int i = 0; i < 32; ++i)
                        if (s0 + i < S) w[i >> 2] |= (uint32_t)src[i] << (8 * (i & 3));
                    v0 = make_uint4(w[0], w[1], w[2], w[3]);
                    v1 = make_uint4(w[4], w[5], w[6], w[7]);
                }
            }
        };
        uint4 pre0, pre1;
        fetch(t, pre0, pre1);
        for (; t < p.NT;) {
            __syncwarp();
            reinterpret_cast<uint4*>(s_codes)[2 * lane] = pre0;
            reinterpret_cast<uint4*>(s_codes)[2 * lane + 1] = pre1;
            if (lane < p.C) s_run[lane] = p.off_key[(int64_t)lane * (p.NT + 1) + t];
            const int t_next = next_tile(t + ngrid);
            fetch(t_next, pre0, pre1);
            __syncwarp();
            for (uint32_t m = step_mask(t); m; m &= m - 1, ++it) {
                const int st = __ffs(m) - 1;
                const int s = it % NST;
                const uint32_t ph = (it / NST) & 1;
                unsigned char* btile = base + (size_t)s * stage_bytes + NDB * T32_BOX_BYTES;
                const uint32_t code = s_codes[st * T32_KPX + lane];
                bar_wait(&empty_bar[s], ph ^ 1);                         // both MMA passes of the stage's previous use are done
                if (lane == 0) T32_STAMP(4, it);
                for (int i = lane; i < p.NCLS * 8; i += 32) reinterpret_cast<uint4*>(btile)[i] = make_uint4(0u, 0u, 0u, 0u);
                if (lane == 0) s_nkeys[s] = 0;
                __syncwarp();
                if (code & CODE_LV) *reinterpret_cast<uint32_t*>(btile + box_off32(code & CODE_CLS_MASK, (uint32_t)lane)) = 0x3F800000u;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) bar_arrive(&bfull_bar[s]);
                const bool is_key = code & CODE_KEY;
                const uint32_t kcls = code & CODE_CLS_MASK;
                const uint32_t peers = __match_any_sync(0xffffffffu, is_key ? kcls : 0xffffu);
                const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
                uint32_t ord = 0;
                if (is_key) ord = s_run[kcls] + rank;
                __syncwarp();
                if (is_key && rank == 0) s_run[kcls] += __popc(peers);
                if (is_key && ord >= (uint32_t)s_skip[kcls]) {
                    const uint32_t cap = (uint32_t)p.cap[kcls];
                    const uint32_t pos = ((uint32_t)s_base[kcls] + ord % cap) % cap;
                    const uint32_t slot = atomicAdd(&s_nkeys[s], 1u);
                    s_keys[s][slot] = (((uint32_t)p.row_off[kcls] + pos) << 6) | (uint32_t)lane;
                }
                __syncwarp();
                if (lane == 0) { bar_arrive(&kfull_bar[s]); T32_STAMP(5, it); }
            }
            t = t_next;
        }
    } else {
        // ---- converters (warps 4-11): keys out of the box, then lo = x - (x & 0xFFFFE000) -> tensor memory ----
        const int q = warp & 3;                                          // TMEM lane quarter this warp may access
        const int half = (warp - 4) >> 2;                                // boxes half, half + 2
        const int cr = q * 32 + lane;                                    // feature row inside a box == TMEM lane
        const uint32_t t_off = (uint32_t)(cr >> 3) * 1024u + (uint32_t)(cr & 7) * 128u;
        const uint32_t sw = (uint32_t)(cr & 7);
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t it = 0;
        for (int t = next_tile(blockIdx.x); t < p.NT; t = next_tile(t + ngrid)) {
            for (uint32_t m = step_mask(t); m; m &= m - 1, ++it) {
                const int s = it % NST, l = it % NLO;
                const uint32_t ph = (it / NST) & 1;
                const uint32_t stage_a = s32(base + (size_t)s * stage_bytes) + t_off;
                bar_wait(&kfull_bar[s], ph);
                const uint32_t nkeys = s_nkeys[s];
                bar_wait(&full_bar[s], ph);
                if (warp == 4 && lane == 0) T32_STAMP(6, it);
                for (uint32_t k = 0; k < nkeys; ++k) {
                    const uint32_t e = s_keys[s][k];
                    const uint32_t kp = e & 63u;
                    const int64_t row = (int64_t)(e >> 6) * p.D;
                    const uint32_t src = stage_a + (((kp >> 2) ^ sw) << 4) + ((kp & 3u) << 2);
                    uint32_t v[T32_MAX_DB / 2];
#pragma unroll
                    for (int i = 0; i < T32_MAX_DB / 2; ++i) {                              // boxes half, half + 2
                        const int db = half + 2 * i;
                        if (db * T32_ROWS + cr < p.D)
                            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v[i]) : "r"(src + (uint32_t)db * T32_BOX_BYTES));
                    }
#pragma unroll
                    for (int i = 0; i < T32_MAX_DB / 2; ++i) {
                        const int d = (half + 2 * i) * T32_ROWS + cr;
                        if (d < p.D) p.bank_rows[row + d] = __uint_as_float(v[i]);
                    }
                }
                bar_wait(&loempty_bar[l], ((it / NLO) & 1) ^ 1);         // the lo pass that last read this slot has finished
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int i = 0; i < T32_MAX_DB / 2; ++i) {
                    const int db = half + 2 * i;
                    if (db < NDB) {                                      // warp-uniform
                        // this thread's feature row: 32 pixels = 8 swizzled 16-byte chunks (a quarter-warp covers all 32 banks)
                        uint32_t u[32];
                        const uint32_t row_a = stage_a + (uint32_t)db * T32_BOX_BYTES;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                         : "=r"(u[4 * j]), "=r"(u[4 * j + 1]), "=r"(u[4 * j + 2]), "=r"(u[4 * j + 3])
                                         : "r"(row_a + ((((uint32_t)j) ^ sw) << 4)));
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            u[e] = __float_as_uint(__fsub_rn(__uint_as_float(u[e]), __uint_as_float(u[e] & 0xFFFFE000u)));
                        const uint32_t taddr = t_lane + col_slots + (uint32_t)l * slot_cols + (uint32_t)db * T32_KPX;
                        asm volatile(
                            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                            "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                            ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]),
                              "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]),
                              "r"(u[16]), "r"(u[17]), "r"(u[18]), "r"(u[19]), "r"(u[20]), "r"(u[21]), "r"(u[22]), "r"(u[23]),
                              "r"(u[24]), "r"(u[25]), "r"(u[26]), "r"(u[27]), "r"(u[28]), "r"(u[29]), "r"(u[30]), "r"(u[31])
                            : "memory");
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) { bar_arrive(&lofull_bar[l]); if (warp == 4) T32_STAMP(7, it); }
            }
        }
    }

    __syncwarp();
    if (warp >= 4 && warp < 8) {
        const bool any = total_steps() > 0;                              // no step: the accumulators were never written
        bar_wait(&done_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        for (int db = 0; db < NDB; ++db) {
            for (int c0 = 0; c0 < p.NCLS; c0 += 16) {
                uint32_t v[16], w[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] = w[c] = 0u;
                if (any) {
                    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + db * p.NCLS + c0;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                        : "r"(taddr));
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                          "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                        : "r"(taddr + col_lo_acc));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 16; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) + __uint_as_float(w[c]));   // hi + lo
                }
                const int d = db * T32_ROWS + q * 32 + lane;
                if (d < p.D) {
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        if (c0 + c < p.C) p.partials[((int64_t)blockIdx.x * p.C + c0 + c) * p.D + d] = __uint_as_float(v[c]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    proto_finalize_tail(p.partials, (int)gridDim.x, p.C, p.D, const_cast<arco_plan*>(p.plan), p.proto_sums);
}

// ---------------------------------------------------------------------------------------------------
static int tc32_stages(const arco_dims& d) {
    const int ndb = (d.feat + T32_ROWS - 1) / T32_ROWS;
    const int stage = ndb * T32_BOX_BYTES + T32_B_BYTES;
    int n = (int)((227 * 1024 - 4096) / stage);
    return n > T32_MAX_ST ? T32_MAX_ST : n;
}

bool proto_tc32_supported(const arco_dims& d) {
    return d.rep_dtype == ARCO_F32 && d.classes <= 32 && d.feat <= T32_MAX_DB * T32_ROWS && d.feat >= 32 && d.space % 4 == 0 &&
           tc32_stages(d) >= 2;
}

size_t proto_tc32_smem(const arco_dims& d) {
    const int ndb = (d.feat + T32_ROWS - 1) / T32_ROWS;
    return (size_t)tc32_stages(d) * (ndb * T32_BOX_BYTES + T32_B_BYTES) + 1024;
}

int launch_proto_tc32(const arco_dims& d, const void* rep_teacher, const arco_bank* bank, const arco_ws_layout& L, char* ws,
                      int rows, double* proto_sums, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    ARCO_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    ARCO_REQUIRE(bank->row_dtype == ARCO_F32, "an fp32 representation head needs an fp32 ring");
    for (int c = 0; c < d.classes; ++c)
        ARCO_REQUIRE(bank->row_off[c] + bank->cap[c] < (1ll << 26), "memory bank too large for the packed key list (2^26 rows)");
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)d.space, (cuuint64_t)d.feat, (cuuint64_t)(d.n_lab + d.n_unlab)};
    const cuuint64_t gstr[2] = {(cuuint64_t)d.space * 4, (cuuint64_t)d.space * d.feat * 4};
    const cuuint32_t box[3] = {T32_KPX, T32_ROWS, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(rep_teacher), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return ARCO_ERR_CUDA;
    }
    ProtoTc32Params p;
    p.codes = (const uint8_t*)(ws + L.codes);
    p.tile_flagged = (const uint32_t*)(ws + L.tile_flagged);
    p.off_key = (const uint32_t*)(ws + L.off_key);
    p.plan = (const arco_plan*)(ws + L.plan);
    p.bank_rows = (float*)bank->rows;
    p.partials = (float*)(ws + L.partials);
    p.proto_sums = proto_sums;
    for (int c = 0; c < ARCO_MAX_CLASSES; ++c) { p.row_off[c] = bank->row_off[c]; p.cap[c] = bank->cap[c] > 0 ? bank->cap[c] : 1; }
    p.S = d.space; p.B = d.n_lab + d.n_unlab; p.C = d.classes; p.D = d.feat;
    p.tpi = L.tiles_per_image; p.NT = L.n_tiles; p.NDB = (d.feat + T32_ROWS - 1) / T32_ROWS;
    p.NST = tc32_stages(d);
    p.NCLS = d.classes <= 16 ? 16 : 32;
    {   // tensor memory: 512 columns = 2 * NDB * NCLS accumulator columns + NLO operand slots of NDB * 32 columns
        int nlo = (512 - 2 * p.NDB * p.NCLS) / (p.NDB * T32_KPX);
        p.NLO = nlo > T32_MAX_LO ? T32_MAX_LO : nlo;
        ARCO_REQUIRE(p.NLO >= 1, "proto_tc32: no tensor-memory room for the lo operand");
    }
    const size_t smem = proto_tc32_smem(d);
    ARCO_CUDA_CHECK(cudaFuncSetAttribute(proto_tc32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    proto_tc32_kernel<<<rows, 384, smem, st>>>(map, p);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

}  // namespace arco

#ifdef ARCO_TC_TRACE
extern "C" __attribute__((visibility("default"))) int arco_debug_tc32_trace(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, arco::g_tc32_trace, sizeof(long long) * 8 * 512);
}
#endif
