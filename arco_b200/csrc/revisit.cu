// the nearest-neighbour "revisiting" loss and its random-pool queue.
//
// Reference (train_arco_2d.py:126-136, :156-159, :400-402, :109-120):
//   rep_u, rep_u_teacher [bs, D, H, W] are flattened to [bs, L] (L = D*H*W = 32.5 M at D=496, 256x256) and L2-normalised
//   (two full read+write passes each), dist = 2 - 2 * rep_n @ pool^T against the K = 36 unit rows of random_pool
//   [K, L] fp32 (4.7 GB), once for the student (-> top-k smallest: the neighbours) and once for the teacher (-> the
//   distances that are averaged), then the normalised teacher rows overwrite pool[ptr : ptr+bs].
//
// Here: ONE streaming pass reads every element of rep_u, rep_u_teacher and random_pool exactly once and produces all
// 2*bs*K dot products plus the 2*bs squared norms (the normalised copies are never materialised: x_n . p = (x . p)/|x|);
// a one-CTA kernel turns them into distances, top-k, gather and the loss; the enqueue is one scale-and-copy pass.
//
// The pass is a [2bs x L] x [L x K] contraction with M = 24, N = 36 -- 864 FMAs per 60 loaded values, i.e. FMA- and
// HBM-time are about equal (0.8 vs 0.95 ms at the reference shape), so it runs on the CUDA cores in exact fp32 (a
// single-pass TF32 MMA would truncate both operands: 1e-3 relative on a dot, the loss needs 1e-5): a warp owns a
// 12-row x 9-column tile of the output, its lanes stride over L (conflict-free shared-memory reads, 10 FMAs per 8-byte read),
// chunks of L are staged in a 3-4 deep shared-memory ring by whole-row TMA bulk copies (mbarrier full / empty), CTAs are persistent and the 108 accumulators per lane
// meet in shuffles once, at the end.  Per-CTA partials are folded in fp64 in a fixed order (deterministic).
#include <cuda.h>
#include <stdlib.h>

#include "arco_common.cuh"
#include "tc_common.cuh"

namespace arco {

constexpr int RV_TM = 12, RV_TN = 9;          // output tile of a warp: rows (reps) x columns (pool rows)
constexpr int RV_TC = 256;                   // elements of L per staged chunk
constexpr int RV_MAX_ST = 4;                 // stages of the chunk ring
constexpr int RV_MAX_WARPS = 8;               // 256 threads x 255 registers: 108 + 12 accumulators and 42 operands per lane

struct RevisitParams {
    const void* rep_s;       // [bs][L]
    const void* rep_t;       // [bs][L]
    const float* pool;       // [K][L]
    float* partials;         // [grid][MP*NP + MP]
    int64_t L;
    int32_t bs, K, MT, NT_, MP, NP, NST;
    int32_t probe;           // tuning probe (ARCO_RV_PROBE): 1 = stage only (no math), 2 = math only (no copies)   // tiles and padded sizes: MP = MT*6 >= 2bs, NP = NT_*9 >= K
    int64_t nchunks;
};


template <typename T> __device__ __forceinline__ float rv_load(const T* p);
template <> __device__ __forceinline__ float rv_load<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float rv_load<__nv_bfloat16>(const __nv_bfloat16* p) {
    return bf16_bits_to_float(*reinterpret_cast<const unsigned short*>(p));
}

template <typename T> __device__ __forceinline__ void rv_load2(const T* p, float& x, float& y);
template <> __device__ __forceinline__ void rv_load2<float>(const float* p, float& x, float& y) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    x = v.x; y = v.y;
}
template <> __device__ __forceinline__ void rv_load2<__nv_bfloat16>(const __nv_bfloat16* p, float& x, float& y) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
    x = __uint_as_float(v << 16); y = __uint_as_float(v & 0xffff0000u);
}

// Staging history (B200, bs 12, K 36, L = 32.5 M; ARCO_RV_PROBE=1 / 2 time the staging / the math alone):
//   v1  every thread issues 16-byte cp.async copies with per-copy row / column arithmetic, two block-wide barriers per
//       256-element chunk, 16 warps x (6 x 9) tiles:                                    2.53 ms bf16 / 2.77 ms fp32
//   v2  one 1-D bulk copy per row, mbarrier ring, 8 warps x (12 x 9) tiles:             2.36 / 2.48  (a warp's lane 0 issued ~8
//       UBLKCP per chunk, serialised; 64-bit div/mod per chunk; a lone producer warp issuing all 60 was slower still: 2.9)
//   v3  incremental stage / phase bookkeeping, straight-line math for a whole chunk:    2.17 / 2.11
//   v4  a chunk = THREE 2-D TMA boxes (student rows, teacher rows, pool rows x 256 elements, no swizzle -> the row-major
//       layout the math loop reads; the tail of L arrives zero-filled), issued by one lane:   1.58 / 1.55 ms
//       = staging alone 0.95 / 1.17 ms (6.5 / 6.6 TB/s), math alone 1.46 / 1.41 ms (the FMA pipe at ~60 %).
// Stages form a ring of NST (3-4), handed back through an `empty` mbarrier every warp arrives on; no __syncthreads in the loop.
template <typename T>
__global__ void __launch_bounds__(256, 1) revisit_dots_kernel(const __grid_constant__ CUtensorMap map_s, const __grid_constant__ CUtensorMap map_t,
                                                              const __grid_constant__ CUtensorMap map_p, RevisitParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[RV_MAX_ST], empty_bar[RV_MAX_ST];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthr = blockDim.x, nwarps = nthr >> 5;
    const int MP = p.MP, NP = p.NP, NST = p.NST;
    const size_t rep_bytes = (size_t)MP * RV_TC * sizeof(T), pool_bytes = (size_t)NP * RV_TC * 4;
    const size_t buf_bytes = rep_bytes + pool_bytes;
    auto rep_buf = [&](int b) { return reinterpret_cast<T*>(smem_raw + (size_t)b * buf_bytes); };
    auto pool_buf = [&](int b) { return reinterpret_cast<float*>(smem_raw + (size_t)b * buf_bytes + rep_bytes); };

    // padding rows (>= 2bs, >= K) are zero and never loaded
    for (size_t i = tid; i < (size_t)NST * buf_bytes / 16; i += nthr) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { bar_init(&full_bar[s], 1); bar_init(&empty_bar[s], (uint32_t)nwarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zero fill (generic proxy) before any bulk copy lands
    __syncthreads();

    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int64_t n_my = first < p.nchunks ? (p.nchunks - first + stride - 1) / stride : 0;
    // The next chunk to issue goes to stage `is` (phase bit `iph`).  A chunk is THREE 2-D TMA boxes (student rows, teacher rows,
    // pool rows x 256 elements, no swizzle = the row-major stage layout the math loop reads), issued by lane 0 of warp 0; the
    // tail of L is zero-filled by the TMA unit.  (One bulk copy per row, the previous version, cost every warp ~8 serialised
    // UBLKCP issues per chunk.)  Stage and phase are carried incrementally -- a chunk is only four trips of the math loop, so
    // bookkeeping instructions count.
    int is = 0;
    uint32_t iph = 1;                                             // parity to wait for on `empty`: passes at once on first use
    int64_t ic = first;                                           // next chunk to issue
    const uint32_t tx_bytes = (uint32_t)(2 * p.bs) * RV_TC * (uint32_t)sizeof(T) + (uint32_t)p.K * RV_TC * 4u;
    auto issue = [&]() {
        if (warp == 0 && p.probe != 2) {
            bar_wait(&empty_bar[is], iph);                        // every warp has left the chunk that used this stage
            if (lane == 0) {
                const int c0 = (int)(ic * RV_TC);
                bar_expect_tx(&full_bar[is], tx_bytes);
                tma_load_2d(rep_buf(is), &map_s, &full_bar[is], c0, 0);
                tma_load_2d(rep_buf(is) + (size_t)p.bs * RV_TC, &map_t, &full_bar[is], c0, 0);
                tma_load_2d(pool_buf(is), &map_p, &full_bar[is], c0, 0);
            }
            __syncwarp();
        }
        ic += stride;
        if (++is == NST) { is = 0; iph ^= 1u; }
    };
    const int mt = warp / p.NT_, nt = warp % p.NT_;
    const bool has_tile = warp < p.MT * p.NT_;
    float acc[RV_TM][RV_TN], sq[RV_TM];
#pragma unroll
    for (int r = 0; r < RV_TM; ++r) {
        sq[r] = 0.f;
#pragma unroll
        for (int k = 0; k < RV_TN; ++k) acc[r][k] = 0.f;
    }

    const int n_chunks = (int)n_my;
    int n_issued = 0;
    for (; n_issued < NST - 1 && n_issued < n_chunks; ++n_issued) issue();
    int b = 0;
    uint32_t bph = 0;
    for (int i = 0; i < n_chunks; ++i) {
        if (n_issued < n_chunks) { issue(); ++n_issued; }
        if (p.probe != 2) bar_wait(&full_bar[b], bph);
        if (has_tile && p.probe != 1) {
            const T* ar = rep_buf(b) + (size_t)(mt * RV_TM) * RV_TC;
            const float* br = pool_buf(b) + (size_t)(nt * RV_TN) * RV_TC;
            // a lane takes PAIRS of consecutive elements: one 8-byte (fp32) / 4-byte (bf16) shared load feeds two FMAs per output
            auto body = [&](int c) {
                float a0[RV_TM], a1[RV_TM], b0[RV_TN], b1[RV_TN];
#pragma unroll
                for (int r = 0; r < RV_TM; ++r) rv_load2<T>(ar + (size_t)r * RV_TC + c, a0[r], a1[r]);
#pragma unroll
                for (int k = 0; k < RV_TN; ++k) {
                    const float2 v = *reinterpret_cast<const float2*>(br + (size_t)k * RV_TC + c);
                    b0[k] = v.x; b1[k] = v.y;
                }
#pragma unroll
                for (int r = 0; r < RV_TM; ++r) {
#pragma unroll
                    for (int k = 0; k < RV_TN; ++k) acc[r][k] = fmaf(a0[r], b0[k], acc[r][k]);
                }
#pragma unroll
                for (int r = 0; r < RV_TM; ++r) {
#pragma unroll
                    for (int k = 0; k < RV_TN; ++k) acc[r][k] = fmaf(a1[r], b1[k], acc[r][k]);
                }
                if (nt == 0) {
#pragma unroll
                    for (int r = 0; r < RV_TM; ++r) sq[r] = fmaf(a0[r], a0[r], sq[r]);
#pragma unroll
                    for (int r = 0; r < RV_TM; ++r) sq[r] = fmaf(a1[r], a1[r], sq[r]);
                }
            };
#pragma unroll
            for (int it = 0; it < RV_TC / 64; ++it) body(2 * lane + 64 * it);   // straight-line: the tail of L arrives as zeros
        }
        __syncwarp();
        if (lane == 0) bar_arrive(&empty_bar[b]);                 // this warp is done reading stage b
        if (++b == NST) { b = 0; bph ^= 1u; }
    }
    if (has_tile) {
        float* out = p.partials + (size_t)blockIdx.x * (MP * NP + MP);
#pragma unroll
        for (int r = 0; r < RV_TM; ++r) {
#pragma unroll
            for (int k = 0; k < RV_TN; ++k) {
                float v = acc[r][k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) out[(mt * RV_TM + r) * NP + nt * RV_TN + k] = v;
            }
            if (nt == 0) {
                float v = sq[r];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) out[MP * NP + mt * RV_TM + r] = v;
            }
        }
    }
}

// One CTA: fold the per-CTA partials in fp64 (fixed order), distances, top-k smallest student distances per row
// (train_arco_2d.py:133; ties: lower pool index first), gather of the teacher distances, mean (:134-135).
// Outputs: loss[1]; nn_index int32 [bs][topk]; stats float [2*bs*K dots (student rows first) | 2*bs squared norms].
__global__ void __launch_bounds__(256) revisit_finish_kernel(const float* __restrict__ partials, int rows, int bs, int K, int MP, int NP,
                                                            int topk, float* __restrict__ loss, int32_t* __restrict__ nn_index,
                                                            float* __restrict__ stats) {
    extern __shared__ double s_tot[];                         // [2bs*K + 2bs]
    const int tid = threadIdx.x;
    const int n_dot = 2 * bs * K, n_all = n_dot + 2 * bs;
    for (int i = tid; i < n_all; i += blockDim.x) {
        int src;
        if (i < n_dot) { const int m = i / K, k = i - m * K; src = m * NP + k; } else src = MP * NP + (i - n_dot);
        double s = 0.0;
        for (int r = 0; r < rows; ++r) s += (double)partials[(size_t)r * (MP * NP + MP) + src];
        s_tot[i] = s;
        stats[i] = (float)s;
    }
    __syncthreads();
    __shared__ float s_row[64];
    if (tid < bs) {
        // F.normalize: x / max(|x|, 1e-12)  (:128,:130)
        const float inv_s = 1.f / fmaxf(sqrtf((float)s_tot[n_dot + tid]), 1e-12f);
        const float inv_t = 1.f / fmaxf(sqrtf((float)s_tot[n_dot + bs + tid]), 1e-12f);
        unsigned long long taken = 0ull;
        float sum = 0.f;
        for (int j = 0; j < topk; ++j) {
            int best = -1;
            float bd = 0.f;
            for (int k = 0; k < K; ++k) {
                if (k < 64 && ((taken >> k) & 1ull)) continue;
                const float dt = 2.f - 2.f * ((float)s_tot[tid * K + k] * inv_s);          // dist_t, from the STUDENT rows (:131)
                if (best < 0 || dt < bd) { best = k; bd = dt; }
            }
            if (best >= 0) {
                if (best < 64) taken |= 1ull << best;
                nn_index[tid * topk + j] = best;
                sum += 2.f - 2.f * ((float)s_tot[(bs + tid) * K + best] * inv_t);          // dist_q, from the TEACHER rows (:132,:134)
            }
        }
        s_row[tid] = sum / (float)topk;
    }
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int i = 0; i < bs; ++i) s += s_row[i];
        loss[0] = s / (float)bs;
    }
}

// pool[(ptr + b) % K] = rep_t[b] / max(|rep_t[b]|, 1e-12)   (train_arco_2d.py:400-402 + :109-120); the pointer is host
// bookkeeping (it advances by bs per call whatever the data, :117), so it is a plain argument.
template <typename T>
__global__ void __launch_bounds__(256) revisit_enqueue_kernel(const T* __restrict__ rep_t, const float* __restrict__ stats, float* __restrict__ pool,
                                                             int64_t ptr, int bs, int K, int64_t L) {
    const int b = blockIdx.y;
    const float n2 = stats[2 * bs * K + bs + b];
    const float inv = 1.f / fmaxf(sqrtf(n2), 1e-12f);
    const int64_t row = (ptr + b) % K;
    const T* src = rep_t + (int64_t)b * L;
    float* dst = pool + row * L;
    constexpr int PER16 = 16 / (int)sizeof(T);
    const int64_t groups = L / PER16;
    // four independent 16-byte loads in flight per thread before the first store (a streaming read + 2x-wide write)
    constexpr int UN = 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g0 < groups; g0 += stride * UN) {
        uint4 raw[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int64_t g = g0 + u * stride;
            raw[u] = g < groups ? ldg_nc_u4(src + g * PER16) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int64_t g = g0 + u * stride;
            if (g >= groups) continue;
            const uint4 r = raw[u];
            if (PER16 == 4) {
                float4 o;
                o.x = __uint_as_float(r.x) * inv; o.y = __uint_as_float(r.y) * inv; o.z = __uint_as_float(r.z) * inv; o.w = __uint_as_float(r.w) * inv;
                __stcs(reinterpret_cast<float4*>(dst) + g, o);
            } else {
                float4 o0, o1;
                o0.x = __uint_as_float(r.x << 16) * inv; o0.y = __uint_as_float(r.x & 0xffff0000u) * inv;
                o0.z = __uint_as_float(r.y << 16) * inv; o0.w = __uint_as_float(r.y & 0xffff0000u) * inv;
                o1.x = __uint_as_float(r.z << 16) * inv; o1.y = __uint_as_float(r.z & 0xffff0000u) * inv;
                o1.z = __uint_as_float(r.w << 16) * inv; o1.w = __uint_as_float(r.w & 0xffff0000u) * inv;
                __stcs(reinterpret_cast<float4*>(dst) + 2 * g, o0);
                __stcs(reinterpret_cast<float4*>(dst) + 2 * g + 1, o1);
            }
        }
    }
}

static int revisit_geometry(int bs, int K, int* MT, int* NT_) {
    *MT = (2 * bs + RV_TM - 1) / RV_TM;
    *NT_ = (K + RV_TN - 1) / RV_TN;
    return (*MT) * (*NT_);
}

}  // namespace arco

extern "C" int64_t arco_revisit_scratch_bytes(int32_t bs, int32_t pool_rows) {
    int MT, NT_;
    arco::revisit_geometry(bs, pool_rows, &MT, &NT_);
    const int64_t per = (int64_t)(MT * arco::RV_TM) * (NT_ * arco::RV_TN) + MT * arco::RV_TM;
    return (int64_t)arco::sm_count() * per * 4 + 256;
}

extern "C" int arco_revisit_loss(const void* rep_u, const void* rep_u_teacher, const float* pool, int32_t bs, int32_t pool_rows,
                                 int64_t length, int32_t rep_dtype, int32_t topk, float* loss, int32_t* nn_index, float* stats,
                                 void* scratch, void* stream) {
    ARCO_REQUIRE(rep_u && rep_u_teacher && pool && loss && nn_index && stats && scratch, "arco_revisit_loss: NULL argument");
    ARCO_REQUIRE(bs >= 1 && bs <= 64 && pool_rows >= 1 && pool_rows <= 256 && topk >= 1 && topk <= pool_rows && length > 0, "arco_revisit_loss: bad sizes");
    const int esz = rep_dtype == ARCO_BF16 ? 2 : 4;
    ARCO_REQUIRE((length * esz) % 16 == 0 && (length * 4) % 16 == 0, "arco_revisit_loss: D*H*W must make 16-byte aligned rows");
    ARCO_REQUIRE((((uintptr_t)rep_u | (uintptr_t)rep_u_teacher | (uintptr_t)pool) & 15) == 0, "arco_revisit_loss: tensors must be 16-byte aligned");
    int MT, NT_;
    const int warps = arco::revisit_geometry(bs, pool_rows, &MT, &NT_);
    ARCO_REQUIRE(warps <= arco::RV_MAX_WARPS, "arco_revisit_loss: ceil(2*bs/12) * ceil(K/9) must be <= 8 (the trainers: bs 12, K 36)");
    arco::RevisitParams p;
    p.rep_s = rep_u; p.rep_t = rep_u_teacher; p.pool = pool; p.partials = (float*)scratch;
    p.L = length; p.bs = bs; p.K = pool_rows; p.MT = MT; p.NT_ = NT_; p.MP = MT * arco::RV_TM; p.NP = NT_ * arco::RV_TN;
    p.nchunks = (length + arco::RV_TC - 1) / arco::RV_TC;
    { const char* e = getenv("ARCO_RV_PROBE"); p.probe = e ? atoi(e) : 0; }
    int grid = arco::sm_count();
    if ((int64_t)grid > p.nchunks) grid = (int)p.nchunks;
    const int threads = warps * 32;
    const size_t stage = (size_t)p.MP * arco::RV_TC * esz + (size_t)p.NP * arco::RV_TC * 4;
    p.NST = (int)((size_t)(200 * 1024) / stage);
    if (p.NST > arco::RV_MAX_ST) p.NST = arco::RV_MAX_ST;
    ARCO_REQUIRE(p.NST >= 2, "arco_revisit_loss: a chunk of 2*bs + K rows does not fit two shared-memory stages");
    const size_t smem = (size_t)p.NST * stage;
    cudaStream_t st = (cudaStream_t)stream;
    ARCO_REQUIRE(length < (1ll << 31) - arco::RV_TC, "arco_revisit_loss: D*H*W must be below 2^31");
    CUtensorMap map_s, map_t, map_p;
    {
        arco::EncodeTiledFn enc = arco::encode_fn();
        ARCO_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
        auto make = [&](CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int e, int rows) -> int {
            const cuuint64_t gdim[2] = {(cuuint64_t)length, (cuuint64_t)rows};
            const cuuint64_t gstr[1] = {(cuuint64_t)length * (cuuint64_t)e};
            const cuuint32_t box[2] = {(cuuint32_t)arco::RV_TC, (cuuint32_t)rows};
            const cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(m, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { arco::set_error("cuTensorMapEncodeTiled (revisit) failed with CUresult %d", (int)r); return ARCO_ERR_CUDA; }
            return ARCO_OK;
        };
        const CUtensorMapDataType rdt = rep_dtype == ARCO_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
        int rc;
        if ((rc = make(&map_s, rep_u, rdt, esz, bs)) != ARCO_OK) return rc;
        if ((rc = make(&map_t, rep_u_teacher, rdt, esz, bs)) != ARCO_OK) return rc;
        if ((rc = make(&map_p, pool, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, pool_rows)) != ARCO_OK) return rc;
    }
    if (rep_dtype == ARCO_BF16) {
        ARCO_CUDA_CHECK(cudaFuncSetAttribute(arco::revisit_dots_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        arco::revisit_dots_kernel<__nv_bfloat16><<<grid, threads, smem, st>>>(map_s, map_t, map_p, p);
    } else {
        ARCO_CUDA_CHECK(cudaFuncSetAttribute(arco::revisit_dots_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        arco::revisit_dots_kernel<float><<<grid, threads, smem, st>>>(map_s, map_t, map_p, p);
    }
    ARCO_LAUNCH_CHECK();
    const int n_all = 2 * bs * pool_rows + 2 * bs;
    arco::revisit_finish_kernel<<<1, 256, (size_t)n_all * 8, st>>>(p.partials, grid, bs, pool_rows, p.MP, p.NP, topk, loss, nn_index, stats);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_revisit_enqueue(const void* rep_u_teacher, const float* stats, float* pool, int64_t pool_ptr, int32_t bs,
                                    int32_t pool_rows, int64_t length, int32_t rep_dtype, void* stream) {
    ARCO_REQUIRE(rep_u_teacher && stats && pool && pool_ptr >= 0, "arco_revisit_enqueue: bad argument");
    ARCO_REQUIRE(bs >= 1 && pool_rows >= bs && pool_rows % bs == 0, "arco_revisit_enqueue: K must be a multiple of the batch (train_arco_2d.py:113)");
    const int esz = rep_dtype == ARCO_BF16 ? 2 : 4;
    ARCO_REQUIRE((length * esz) % 16 == 0, "arco_revisit_enqueue: rows must be 16-byte multiples");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)(arco::sm_count() * 8 / bs + 1), (unsigned)bs);
    if (rep_dtype == ARCO_BF16)
        arco::revisit_enqueue_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)rep_u_teacher, stats, pool, pool_ptr, bs, pool_rows, length);
    else
        arco::revisit_enqueue_kernel<float><<<grid, 256, 0, st>>>((const float*)rep_u_teacher, stats, pool, pool_ptr, bs, pool_rows, length);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}
