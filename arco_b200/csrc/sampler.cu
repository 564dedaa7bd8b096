// Sub-system 3b: in-kernel restatement of ARCO's index samplers.
//
// Reference (host-side Python/NumPy/torch-CPU RNG): grid_monte_carlo_sample
// loss_helper_3d.py:120-184 ("smc"), grid_as_monte_carlo_sample :187-268 ("asmc"), their 1-D fallbacks
// monte_carlo_sample :83-117 / as_monte_carlo_sample :35-80, and plain torch.randint for any other
// `func` (:335-338).  The device sampler reproduces the DISTRIBUTION (identical strata, per-stratum
// counts, antithetic mirrors, drop of indices >= high, uniform shuffle of the survivors, iid uniform
// pads, truncation, float32 round trip of the drawn indices); the random streams themselves are
// Philox4x32-10 counters, so a seed reproduces a run bit-exactly on the GPU.  Parity tests inject the
// reference's own indices instead.
//
//   torch.randperm   -> keyed Feistel bijection on [0,n) (no sort, O(1) per element)
//   order-preserving drop of out-of-range draws -> scan across an 8-CTA thread-block cluster through
//                       distributed shared memory (one cluster per sampler call)
#include <cooperative_groups.h>

#include "arco_common.cuh"

namespace cg = cooperative_groups;

namespace arco {

constexpr int kClusterSize = 8;
constexpr int kThreads = 1024;
constexpr int kItems = 4;                   // draws per thread per round (scan kernel)
constexpr int kEmitThreads = 256;
constexpr int kEmitItems = 4;
constexpr int kMinGridEdge = 8;             // reference falls back below this (probe: high <= 56)
constexpr int kPatch = 16;

struct SampleParams {
    const arco_plan* plan;                  // fused mode: calls derived from the plan; else direct mode
    int32_t* idx_anchor;
    int32_t* idx_neg;
    int32_t* scratch;                       // per call: [0] = survivors of the bottom block row, [4..] their values
    int64_t scratch_stride;                 // int32 elements per call
    int32_t C, Q, N;
    // direct mode
    int64_t high, shape;
    int32_t* out;
    int32_t func;
    uint64_t seed, stream;
    int32_t only_if_replanned;              // fused mode: do nothing unless arco_replan_global changed the plan
};

enum { PURPOSE_DRAW = 0, PURPOSE_PAD = 1, PURPOSE_UNIFORM = 2, PURPOSE_PERM = 3 };

__device__ __forceinline__ uint32_t round_sqrt(uint64_t high) {
    uint64_t e = (uint64_t)sqrt((double)high);
    while (e * e > high) --e;
    while ((e + 1) * (e + 1) <= high) ++e;
    if (high - e * e > e) ++e;                        // round(math.sqrt(high)), :136
    return (uint32_t)e;
}

// Everything a thread needs to know about one sampler call (all fields are call-uniform).
struct Call {
    int64_t high, shape;
    int32_t* out;
    int32_t* scratch;
    uint64_t stream;
    uint32_t edge, step, per_block, half;
    bool active, anti, uniform, strata, drops;
};

__device__ __forceinline__ Call decode_call(const SampleParams& p, int call) {
    Call c;
    c.active = true;
    c.stream = p.stream;
    if (p.plan) {
        const int j = call >> 1;
        const bool neg = call & 1;
        c.active = p.plan->slot_active[j] != 0;                            // skipped positions draw nothing
        if (neg) {
            const int vc = p.plan->valid_class[j];
            c.high = c.active ? p.plan->bank_len[vc] : 0;                  // len(negative_feat), :471-473
            c.shape = (int64_t)p.Q * p.N;
            c.out = p.idx_neg + (int64_t)j * p.Q * p.N;
        } else {
            c.high = p.plan->n_anchor[j];                                  // len(seg_feat_low_entropy_list[i]), :444-446
            c.shape = p.Q;
            c.out = p.idx_anchor + (int64_t)j * p.Q;
        }
        c.stream = (p.stream + (uint64_t)p.plan->step_ctr) * 64ull + (uint64_t)call;   // fresh stream every device step
    } else {
        c.high = p.high; c.shape = p.shape; c.out = p.out;
    }
    c.scratch = p.scratch + (int64_t)call * p.scratch_stride;
    if (c.high <= 0 || c.shape <= 0) c.active = false;
    const bool structured = p.func == ARCO_FUNC_SMC || p.func == ARCO_FUNC_ASMC;
    c.anti = p.func == ARCO_FUNC_ASMC;
    c.edge = (structured && c.active) ? round_sqrt((uint64_t)c.high) : 0;
    c.strata = structured && c.edge < kMinGridEdge;
    // torch.randint(high, (shape,))  (:336 and :84-85 / :36-37)
    c.uniform = !structured || (c.strata && (c.high / kPatch > c.shape || c.high < kPatch));
    c.step = c.edge / 4;
    c.per_block = 0; c.half = 0; c.drops = false;
    if (c.active && !c.uniform && !c.strata) {
        c.per_block = (uint32_t)((uint64_t)c.shape * c.edge * c.edge / (uint64_t)c.high / 16ull);
        c.half = c.per_block / 2;
        if (c.anti) c.per_block = 2 * c.half;
        c.drops = (uint64_t)c.edge * c.edge > (uint64_t)c.high;           // only then can a draw be >= high (:165)
    }
    return c;
}

// value of canonical draw w (0 <= w < per_block) of grid block k; returns -1 if it is dropped (>= high)
__device__ __forceinline__ int32_t grid_draw(const Call& c, const Philox& rng, uint32_t k, uint32_t w) {
    const uint32_t bi = k >> 2, bj = k & 3;
    const uint32_t r0 = bi * c.step, c0 = bj * c.step;
    const uint32_t nr = (bi == 3) ? c.edge - r0 : c.step;                 // last block row/column absorbs the remainder (:146-153)
    const uint32_t nc = (bj == 3) ? c.edge - c0 : c.step;
    const bool mirror = c.anti && w >= c.half;
    const uint32_t m = mirror ? w - c.half : w;
    const uint32_t r = bounded(rng(k * c.per_block + m, PURPOSE_DRAW, (uint32_t)c.stream, (uint32_t)(c.stream >> 32)).x, nr * nc);
    int64_t v = (int64_t)(r0 + r / nc) * c.edge + (c0 + r % nc);
    if (mirror) v = ((int64_t)(2 * r0 + nr - 1) * c.edge + (2 * c0 + nc - 1)) - v;   // center - x, center = int(2*mean(block))
    v = (int64_t)(float)v;                                                // torch.Tensor(...) float32 round trip (:163,:245)
    return v < c.high ? (int32_t)v : -1;                                  // mask = cur_list < high (:165)
}

// Kernel 1 (only does work when edge^2 > high): order-preserving compaction of the bottom block row's
// draws -- the only ones that can fall outside [0, high) -- scanned across an 8-CTA cluster through DSMEM.
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kThreads)
sample_scan_kernel(SampleParams p) {
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    const int call = blockIdx.x / kClusterSize;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ uint32_t s_warp[32];
    if (p.only_if_replanned && p.plan->replanned == 0) return;            // uniform over the whole grid
    __shared__ uint32_t s_total;
    const Call c = decode_call(p, call);
    if (!c.active || c.uniform || c.strata || !c.drops) return;          // cluster-uniform
    const Philox rng(p.seed);
    const uint32_t gthread = rank * kThreads + tid;
    const uint32_t gstride = kClusterSize * kThreads;
    const uint32_t n_bottom = 4u * c.per_block;                           // blocks 12..15, canonical order
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_bottom; base += gstride * kItems) {
        int32_t vals[kItems];
        uint32_t cnt = 0;
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const uint32_t u = base + gthread * kItems + i;
            vals[i] = -1;
            if (u < n_bottom) {
                vals[i] = grid_draw(c, rng, 12u + u / c.per_block, u % c.per_block);
                cnt += vals[i] >= 0;
            }
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = s_warp[lane];
            uint32_t ws = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += y;
            }
            s_warp[lane] = ws - w;
            if (lane == 31) s_total = ws;
        }
        __syncthreads();
        uint32_t pos = s_warp[warp] + incl - cnt;
        cluster.sync();
        uint32_t before = 0, round_total = 0;
        for (uint32_t r = 0; r < kClusterSize; ++r) {
            const uint32_t tr = *cluster.map_shared_rank(&s_total, r);
            if (r < rank) before += tr;
            round_total += tr;
        }
        cluster.sync();                                                   // s_total / s_warp may be rewritten
        pos += carry + before;
#pragma unroll
        for (int i = 0; i < kItems; ++i)
            if (vals[i] >= 0) c.scratch[4 + pos++] = vals[i];
        carry += round_total;
    }
    if (gthread == 0) c.scratch[0] = (int32_t)carry;
}

// Kernel 2: every output position independently.  survivors S = (blocks 0..11 verbatim) ++ (compacted bottom
// row); out[t] = S[perm(t)] for t < min(|S|, shape), iid uniform pads after that (:170-180).
__global__ void __launch_bounds__(kEmitThreads) sample_emit_kernel(SampleParams p) {
    const int call = blockIdx.y;
    if (p.only_if_replanned && p.plan->replanned == 0) return;
    const Call c = decode_call(p, call);
    if (!c.active) return;
    const int64_t t0 = ((int64_t)blockIdx.x * kEmitThreads + threadIdx.x) * kEmitItems;
    if (t0 >= c.shape) return;
    const Philox rng(p.seed);
    const uint32_t st_lo = (uint32_t)c.stream, st_hi = (uint32_t)(c.stream >> 32);
    const uint32_t H = (uint32_t)c.high;
    int32_t res[kEmitItems];
    if (c.uniform) {
#pragma unroll
        for (int i = 0; i < kEmitItems; ++i)
            res[i] = (int32_t)bounded(rng((uint32_t)(t0 + i), PURPOSE_UNIFORM, st_lo, st_hi).x, H);
    } else {
        const uint4 ka = rng(0, PURPOSE_PERM, st_lo, st_hi), kb = rng(1, PURPOSE_PERM, st_lo, st_hi);
        if (c.strata) {
            // 1-D strata of 16 (:83-117 / :35-80): structured draws, uniform pads, shuffle of ALL `shape` entries
            const uint32_t strata = (uint32_t)(c.high / kPatch);
            uint32_t per = (uint32_t)(c.shape / strata);
            const uint32_t half = per / 2;
            if (c.anti) per = 2 * half;
            const uint64_t n_struct = (uint64_t)strata * per;
            const FeistelPerm perm((uint32_t)c.shape, ka, kb);
#pragma unroll
            for (int i = 0; i < kEmitItems; ++i) {
                const int64_t t = t0 + i;
                res[i] = 0;
                if (t >= c.shape) continue;
                const uint32_t u = perm((uint32_t)t);
                if (u < n_struct) {
                    const uint32_t k = u / per, w = u % per;
                    const bool mirror = c.anti && w >= half;
                    const uint32_t m = mirror ? w - half : w;
                    const uint32_t x = k * kPatch + bounded(rng(k * per + m, PURPOSE_DRAW, st_lo, st_hi).x, kPatch);
                    res[i] = mirror ? (int32_t)((2 * k + 1) * kPatch - 1 - x) : (int32_t)x;
                } else {
                    res[i] = (int32_t)bounded(rng(u, PURPOSE_PAD, st_lo, st_hi).x, H);
                }
            }
        } else {
            const uint32_t top = 12u * c.per_block;
            const uint32_t n_bottom = c.drops ? (uint32_t)c.scratch[0] : 4u * c.per_block;
            const uint32_t M = top + n_bottom;
            const FeistelPerm perm(M, ka, kb);
#pragma unroll
            for (int i = 0; i < kEmitItems; ++i) {
                const int64_t t = t0 + i;
                res[i] = 0;
                if (t >= c.shape) continue;
                if (t < M) {
                    const uint32_t rho = perm((uint32_t)t);
                    if (rho < top || !c.drops) res[i] = grid_draw(c, rng, rho / c.per_block, rho % c.per_block);
                    else res[i] = c.scratch[4 + (rho - top)];
                } else {
                    res[i] = (int32_t)bounded(rng((uint32_t)t, PURPOSE_PAD, st_lo, st_hi).x, H);   // (:176-177)
                }
            }
        }
    }
    if (t0 + kEmitItems <= c.shape && (((uintptr_t)(c.out + t0)) & 15) == 0) {
        *reinterpret_cast<int4*>(c.out + t0) = make_int4(res[0], res[1], res[2], res[3]);
    } else {
#pragma unroll
        for (int i = 0; i < kEmitItems; ++i)
            if (t0 + i < c.shape) c.out[t0 + i] = res[i];
    }
}

static int launch_sampler(const SampleParams& p, int calls, int64_t max_shape, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(calls * kClusterSize);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    ARCO_CUDA_CHECK(cudaLaunchKernelEx(&cfg, sample_scan_kernel, p));
    const int64_t per_cta = (int64_t)kEmitThreads * kEmitItems;
    dim3 grid((unsigned)((max_shape + per_cta - 1) / per_cta), (unsigned)calls);
    sample_emit_kernel<<<grid, kEmitThreads, 0, st>>>(p);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

}  // namespace arco

static int sample_impl(const arco_dims* dims, int32_t func, uint64_t seed, uint64_t step, int32_t* idx_anchor,
                       int32_t* idx_neg, void* workspace, void* stream, int only_if_replanned) {
    ARCO_REQUIRE(dims && idx_anchor && idx_neg && workspace, "arco_sample: NULL argument");
    const arco_dims& d = *dims;
    ARCO_REQUIRE((int64_t)d.queries * d.negatives < (int64_t)1 << 30, "Q*N too large");
    arco_ws_layout L;
    arco::compute_layout(d, &L);
    arco::SampleParams p = {};
    p.plan = (const arco_plan*)((char*)workspace + L.plan);
    p.idx_anchor = idx_anchor; p.idx_neg = idx_neg;
    p.scratch = (int32_t*)((char*)workspace + L.sample_scratch);
    const int64_t draws = (int64_t)d.queries * (d.negatives > 0 ? d.negatives : 1);
    p.scratch_stride = draws + draws / 4 + 4096;
    p.C = d.classes; p.Q = d.queries; p.N = d.negatives;
    p.func = func; p.seed = seed; p.stream = step;
    p.only_if_replanned = only_if_replanned;
    return arco::launch_sampler(p, 2 * d.classes, draws > d.queries ? draws : d.queries, (cudaStream_t)stream);
}

extern "C" int arco_sample(const arco_dims* dims, int32_t func, uint64_t seed, uint64_t step, int32_t* idx_anchor,
                           int32_t* idx_neg, void* workspace, void* stream) {
    return sample_impl(dims, func, seed, step, idx_anchor, idx_neg, workspace, stream, 0);
}

extern "C" int arco_sample_if_replanned(const arco_dims* dims, int32_t func, uint64_t seed, uint64_t step, int32_t* idx_anchor,
                                        int32_t* idx_neg, void* workspace, void* stream) {
    return sample_impl(dims, func, seed, step, idx_anchor, idx_neg, workspace, stream, 1);
}

extern "C" int arco_sample_one(int32_t func, int64_t high, int64_t shape, uint64_t seed, uint64_t stream_id,
                               int32_t* out, void* scratch, int64_t scratch_bytes, void* stream) {
    ARCO_REQUIRE(out && high > 0 && high < ((int64_t)1 << 31) && shape > 0 && shape < ((int64_t)1 << 30),
                 "arco_sample_one: bad argument");
    ARCO_REQUIRE(scratch && scratch_bytes >= (shape / 2 + 4096) * 4, "arco_sample_one: scratch too small (need (shape/2+4096)*4 bytes)");
    arco::SampleParams p = {};
    p.plan = nullptr;
    p.scratch = (int32_t*)scratch;
    p.scratch_stride = 0;
    p.high = high; p.shape = shape; p.out = out; p.func = func; p.seed = seed; p.stream = stream_id;
    return arco::launch_sampler(p, 1, shape, (cudaStream_t)stream);
}
