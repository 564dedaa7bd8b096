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
constexpr int kItems = 4;                   // draws per thread per round
constexpr int kMinGridEdge = 8;             // reference falls back below this (probe: high <= 56)
constexpr int kPatch = 16;

struct SampleParams {
    const arco_plan* plan;                  // fused mode: calls derived from the plan; else direct mode
    int32_t* idx_anchor;
    int32_t* idx_neg;
    int32_t C, Q, N;
    // direct mode
    int64_t high, shape;
    int32_t* out;
    int32_t func;
    uint64_t seed, stream;
};

enum { PURPOSE_DRAW = 0, PURPOSE_PAD = 1, PURPOSE_UNIFORM = 2, PURPOSE_PERM = 3 };

struct GridGeom {
    uint32_t edge, step, per_block, half;
    bool anti;
    __device__ void block_rect(uint32_t k, uint32_t& r0, uint32_t& c0, uint32_t& nr, uint32_t& nc) const {
        const uint32_t bi = k >> 2, bj = k & 3;
        r0 = bi * step; c0 = bj * step;
        nr = (bi == 3) ? edge - r0 : step;            // last block row/column absorbs the remainder (:146-153)
        nc = (bj == 3) ? edge - c0 : step;
    }
};

__device__ __forceinline__ uint32_t round_sqrt(uint64_t high) {
    uint64_t e = (uint64_t)sqrt((double)high);
    while (e * e > high) --e;
    while ((e + 1) * (e + 1) <= high) ++e;
    if (high - e * e > e) ++e;                        // round(math.sqrt(high)), :136
    return (uint32_t)e;
}

__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kThreads)
sample_kernel(SampleParams p) {
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    const int call = blockIdx.x / kClusterSize;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total;

    // ---- decode the call (cluster-uniform) ----
    int64_t high, shape;
    int32_t* out;
    uint64_t stream = p.stream;
    if (p.plan) {
        const int j = call >> 1;
        const bool neg = call & 1;
        if (!p.plan->slot_active[j]) return;                              // skipped positions draw nothing
        if (neg) {
            high = p.plan->bank_len[p.plan->valid_class[j]];              // len(negative_feat), :471-473
            shape = (int64_t)p.Q * p.N;
            out = p.idx_neg + (int64_t)j * p.Q * p.N;
        } else {
            high = p.plan->n_anchor[j];                                   // len(seg_feat_low_entropy_list[i]), :444-446
            shape = p.Q;
            out = p.idx_anchor + (int64_t)j * p.Q;
        }
        stream = p.stream * 64ull + (uint64_t)call;
    } else {
        high = p.high; shape = p.shape; out = p.out;
    }
    if (high <= 0 || shape <= 0) return;
    const Philox rng(p.seed);
    const uint32_t st_lo = (uint32_t)stream, st_hi = (uint32_t)(stream >> 32);
    const uint32_t gthread = rank * kThreads + tid;
    const uint32_t gstride = kClusterSize * kThreads;
    const uint32_t H = (uint32_t)high;

    const bool structured = p.func == ARCO_FUNC_SMC || p.func == ARCO_FUNC_ASMC;
    const bool anti = p.func == ARCO_FUNC_ASMC;
    const uint32_t edge = structured ? round_sqrt((uint64_t)high) : 0;

    if (!structured || (edge < kMinGridEdge && (high / kPatch > shape || high < kPatch))) {
        // torch.randint(high, (shape,))  (:336 and :84-85 / :36-37)
        for (int64_t i = gthread; i < shape; i += gstride)
            out[i] = (int32_t)bounded(rng((uint32_t)i, PURPOSE_UNIFORM, st_lo, st_hi).x, H);
        return;
    }
    const uint4 ka = rng(0, PURPOSE_PERM, st_lo, st_hi), kb = rng(1, PURPOSE_PERM, st_lo, st_hi);

    if (edge < kMinGridEdge) {
        // 1-D strata of 16 (:83-117 / :35-80): structured draws, uniform pads, shuffle of ALL `shape` entries
        const uint32_t strata = (uint32_t)(high / kPatch);
        uint32_t per = (uint32_t)(shape / strata);
        const uint32_t half = per / 2;
        if (anti) per = 2 * half;
        const uint64_t n_struct = (uint64_t)strata * per;
        const FeistelPerm perm((uint32_t)shape, ka, kb);
        for (int64_t i = gthread; i < shape; i += gstride) {
            const uint32_t u = perm((uint32_t)i);
            int32_t v;
            if (u < n_struct) {
                const uint32_t k = u / per, w = u % per;
                const uint32_t m = anti ? w % half : w;
                const uint32_t x = k * kPatch + bounded(rng(k * per + m, PURPOSE_DRAW, st_lo, st_hi).x, kPatch);
                v = (anti && w >= half) ? (int32_t)((2 * k + 1) * kPatch - 1 - x) : (int32_t)x;
            } else {
                v = (int32_t)bounded(rng(u, PURPOSE_PAD, st_lo, st_hi).x, H);
            }
            out[i] = v;
        }
        return;
    }

    // ---- 4x4 grid of blocks over an edge x edge image (:136-182 / :203-265) ----
    GridGeom g;
    g.edge = edge; g.step = edge / 4; g.anti = anti;
    g.per_block = (uint32_t)((uint64_t)shape * edge * edge / (uint64_t)high / 16ull);
    g.half = g.per_block / 2;
    if (anti) g.per_block = 2 * g.half;
    const uint64_t n_tot64 = 16ull * g.per_block;
    const uint32_t n_tot = (uint32_t)n_tot64;
    const FeistelPerm perm(n_tot, ka, kb);

    uint32_t carry = 0;                                                   // survivors emitted so far
    for (uint32_t base = 0; base < n_tot; base += gstride * kItems) {
        int32_t vals[kItems];
        uint32_t cnt = 0;
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const uint32_t t = base + gthread * kItems + i;
            vals[i] = -1;
            if (t < n_tot) {
                const uint32_t u = perm(t);                               // canonical draw behind shuffled slot t
                const uint32_t k = u / g.per_block, w = u % g.per_block;
                const uint32_t m = anti ? w % g.half : w;
                uint32_t r0, c0, nr, nc;
                g.block_rect(k, r0, c0, nr, nc);
                const uint32_t r = bounded(rng(k * g.per_block + m, PURPOSE_DRAW, st_lo, st_hi).x, nr * nc);
                int64_t v = (int64_t)(r0 + r / nc) * edge + (c0 + r % nc);
                if (anti && w >= g.half) {
                    const int64_t center = (int64_t)(2 * r0 + nr - 1) * edge + (2 * c0 + nc - 1);   // int(2*mean(block))
                    v = center - v;
                }
                v = (int64_t)(float)v;                                    // torch.Tensor(...) float32 round trip (:163,:245)
                if (v < high) { vals[i] = (int32_t)v; ++cnt; }            // mask = cur_list < high (:165)
            }
        }
        // block-level exclusive scan of cnt
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
        // cluster-level scan through distributed shared memory
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
            if (vals[i] >= 0) { if (pos < shape) out[pos] = vals[i]; ++pos; }   // truncate (:179-180)
        carry += round_total;
    }
    // iid uniform pads (:176-177)
    for (int64_t i = (int64_t)carry + gthread; i < shape; i += gstride)
        out[i] = (int32_t)bounded(rng((uint32_t)i, PURPOSE_PAD, st_lo, st_hi).x, H);
}

static int launch_sampler(const SampleParams& p, int calls, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(calls * kClusterSize);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    ARCO_CUDA_CHECK(cudaLaunchKernelEx(&cfg, sample_kernel, p));
    return ARCO_OK;
}

}  // namespace arco

extern "C" int arco_sample(const arco_dims* dims, int32_t func, uint64_t seed, uint64_t step, int32_t* idx_anchor,
                           int32_t* idx_neg, void* workspace, void* stream) {
    ARCO_REQUIRE(dims && idx_anchor && idx_neg && workspace, "arco_sample: NULL argument");
    const arco_dims& d = *dims;
    ARCO_REQUIRE((int64_t)d.queries * d.negatives < (int64_t)1 << 30, "Q*N too large");
    arco_ws_layout L;
    arco::compute_layout(d, &L);
    arco::SampleParams p = {};
    p.plan = (const arco_plan*)((char*)workspace + L.plan);
    p.idx_anchor = idx_anchor; p.idx_neg = idx_neg;
    p.C = d.classes; p.Q = d.queries; p.N = d.negatives;
    p.func = func; p.seed = seed; p.stream = step;
    return arco::launch_sampler(p, 2 * d.classes, (cudaStream_t)stream);
}

extern "C" int arco_sample_one(int32_t func, int64_t high, int64_t shape, uint64_t seed, uint64_t stream_id,
                               int32_t* out, void* scratch, int64_t scratch_bytes, void* stream) {
    (void)scratch; (void)scratch_bytes;
    ARCO_REQUIRE(out && high > 0 && high < ((int64_t)1 << 31) && shape > 0 && shape < ((int64_t)1 << 30),
                 "arco_sample_one: bad argument");
    arco::SampleParams p = {};
    p.plan = nullptr;
    p.high = high; p.shape = shape; p.out = out; p.func = func; p.seed = seed; p.stream = stream_id;
    return arco::launch_sampler(p, 1, (cudaStream_t)stream);
}
