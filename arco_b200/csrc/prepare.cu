// the mask / threshold preparation that feeds the loss, on the device.
//
// Replaces train_arco_2d.py:345-393 (train_arco_3d.py:315-353):
//   prob_*_teacher = softmax(pred_*_teacher, dim=1)                                   (:356-357)
//   entropy        = -sum(prob_u * log(prob_u + 1e-10), dim=1)                         (:353,358-359)
//   low_thresh     = np.percentile(entropy[label_u >= 0].cpu().numpy(), alpha_t)       (:360-362)  GPU -> CPU -> GPU
//   high_thresh    = np.percentile(..., 100 - alpha_t)                                 (:366-369)  GPU -> CPU -> GPU
//   low/high_mask  = entropy.le/ge(thresh) * (label_u >= 0), cat with (label_l >= 0)   (:363-392)
//
// np.percentile (numpy 2.x, float32 data, method "linear") is restated exactly and stays on the device:
//   q32 = float32(q) / float32(100);  v = float32(n - 1) * q32;  i = floor(v);  g = v - i      (all float32)
//   a = sorted[i], b = sorted[i + 1]  (both the last element when v >= n - 1)
//   thr = g >= 0.5 ? b - (b - a) * (1 - g) : a + (b - a) * g                                    (float32, no fma)
// The two order statistics per percentile come from a 3-level radix select (11 + 11 + 10 bits of the order-preserving
// integer image of the float) -- three histogram passes over the <= few million entropies, no sort; every pass begins by
// resolving the previous level from the global histogram (each CTA redundantly, deterministic), so the whole thing is
// four launches and never touches the host.
#include "arco_common.cuh"

namespace arco {

template <int MAXC>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ logits, int64_t B, int C, int64_t S,
                                                           float* __restrict__ prob, float* __restrict__ entropy) {
    const int64_t total = B * S;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t b = i / S, s = i - b * S;
        const float* x = logits + b * C * S + s;
        float v[MAXC];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            v[c] = c < C ? x[(int64_t)c * S] : -INFINITY;
            m = fmaxf(m, v[c]);
        }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            v[c] = c < C ? expf(v[c] - m) : 0.f;
            sum += v[c];
        }
        float ent = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            if (c < C) {
                const float p = v[c] / sum;
                if (prob) prob[b * C * S + (int64_t)c * S + s] = p;
                ent = __fadd_rn(ent, __fmul_rn(p, logf(p + 1e-10f)));
            }
        }
        if (entropy) entropy[i] = -ent;
    }
}

// order-preserving integer image of a float (ascending)
__device__ __forceinline__ uint32_t float_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}

constexpr int SEL_RANKS = 4;                 // (i, i+1) of the low and of the high percentile
constexpr int SEL_B1 = 2048, SEL_B2 = 2048, SEL_B3 = 1024;

struct SelectState {                          // device scratch, zeroed before every use
    uint32_t hist1[SEL_B1];
    uint32_t hist2[SEL_RANKS][SEL_B2];
    uint32_t hist3[SEL_RANKS][SEL_B3];
};

struct SelectResolved {
    uint32_t n;
    uint32_t rank[SEL_RANKS];                 // target ranks
    float g[2];                               // interpolation weights (low, high)
    uint32_t prefix[SEL_RANKS];               // resolved high bits so far
    uint32_t rem[SEL_RANKS];                  // rank inside the resolved bucket
};

// Resolve level `level` (1, 2 or 3) from the histograms; called by every thread of a CTA, result in shared memory.
// 1024 threads.
__device__ void select_resolve(const SelectState* st, int level, float q_low, float q_high, SelectResolved* out, uint32_t* s_scan) {
    const int tid = threadIdx.x;
    __shared__ uint32_t s_tot;
    // ---- level 1: totals, target ranks, bucket of each rank ----
    {
        // inclusive scan of hist1 (2048 bins, 2 per thread)
        const uint32_t a = st->hist1[2 * tid], b = st->hist1[2 * tid + 1];
        s_scan[tid] = a + b;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const uint32_t y = tid >= o ? s_scan[tid - o] : 0u;
            __syncthreads();
            s_scan[tid] += y;
            __syncthreads();
        }
        if (tid == 1023) s_tot = s_scan[1023];
        __syncthreads();
        const uint32_t n = s_tot;
        if (tid == 0) {
            out->n = n;
            const float qs[2] = {q_low, q_high};
            for (int k = 0; k < 2; ++k) {
                uint32_t i0 = 0, i1 = 0;
                float g = 0.f;
                if (n > 0) {
                    const float v = __fmul_rn((float)(n - 1), qs[k]);       // (n - 1) * q in float32 like numpy
                    if (v >= (float)(n - 1)) { i0 = i1 = n - 1; g = v - floorf(v); }
                    else if (v < 0.f) { i0 = i1 = 0; g = v - floorf(v); }
                    else { const float fl = floorf(v); i0 = (uint32_t)fl; i1 = i0 + 1; g = v - fl; }
                }
                out->rank[2 * k] = i0; out->rank[2 * k + 1] = i1; out->g[k] = g;
            }
        }
        __syncthreads();
        const uint32_t excl_pair = s_scan[tid] - (a + b);
#pragma unroll
        for (int r = 0; r < SEL_RANKS; ++r) {
            const uint32_t want = out->rank[r];
            if (n > 0) {
                if (want >= excl_pair && want < excl_pair + a) { out->prefix[r] = 2 * tid; out->rem[r] = want - excl_pair; }
                else if (want >= excl_pair + a && want < excl_pair + a + b) { out->prefix[r] = 2 * tid + 1; out->rem[r] = want - excl_pair - a; }
            } else if (tid == 0) { out->prefix[r] = 0; out->rem[r] = 0; }
        }
        __syncthreads();
    }
    if (level < 2) return;
    // ---- level 2: 2048 bins per rank ----
    for (int r = 0; r < SEL_RANKS; ++r) {
        const uint32_t a = st->hist2[r][2 * tid], b = st->hist2[r][2 * tid + 1];
        s_scan[tid] = a + b;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const uint32_t y = tid >= o ? s_scan[tid - o] : 0u;
            __syncthreads();
            s_scan[tid] += y;
            __syncthreads();
        }
        const uint32_t excl_pair = s_scan[tid] - (a + b);
        const uint32_t want = out->rem[r];
        __syncthreads();
        if (out->n > 0) {
            if (want >= excl_pair && want < excl_pair + a) { out->prefix[r] = (out->prefix[r] << 11) | (2 * tid); out->rem[r] = want - excl_pair; }
            else if (want >= excl_pair + a && want < excl_pair + a + b) { out->prefix[r] = (out->prefix[r] << 11) | (2 * tid + 1); out->rem[r] = want - excl_pair - a; }
        }
        __syncthreads();
    }
    if (level < 3) return;
    // ---- level 3: 1024 bins per rank -> the full 32-bit key ----
    for (int r = 0; r < SEL_RANKS; ++r) {
        const uint32_t a = st->hist3[r][tid];
        s_scan[tid] = a;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const uint32_t y = tid >= o ? s_scan[tid - o] : 0u;
            __syncthreads();
            s_scan[tid] += y;
            __syncthreads();
        }
        const uint32_t excl = s_scan[tid] - a;
        const uint32_t want = out->rem[r];
        __syncthreads();
        if (out->n > 0 && want >= excl && want < excl + a) { out->prefix[r] = (out->prefix[r] << 10) | tid; out->rem[r] = want - excl; }
        __syncthreads();
    }
}

// warp-aggregated shared-memory histogram increment: lanes that hit the same bin add once (entropies cluster in a
// handful of exponent buckets, so plain atomics would serialise)
__device__ __forceinline__ void hist_add(uint32_t* hist, uint32_t bin, bool on) {
    const uint32_t peers = __match_any_sync(0xffffffffu, on ? bin : 0xffffffffu);
    if (on && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
}

// level 1, 2, 3 histogram passes over the valid entropies (CTA-private histograms in shared memory, flushed once)
template <int LEVEL>
__global__ void __launch_bounds__(1024) select_hist_kernel(const float* __restrict__ entropy, const int64_t* __restrict__ label_u,
                                                           int64_t n_px, float q_low, float q_high, SelectState* st) {
    __shared__ SelectResolved res;
    __shared__ uint32_t s_scan[1024];
    extern __shared__ uint32_t s_hist[];                                   // LEVEL 1: [2048]; 2: [4][2048]; 3: [4][1024]
    constexpr int BINS = LEVEL == 3 ? SEL_B3 : SEL_B1;
    constexpr int NH = LEVEL == 1 ? 1 : SEL_RANKS;
    for (int i = threadIdx.x; i < NH * BINS; i += 1024) s_hist[i] = 0u;
    if (LEVEL > 1) select_resolve(st, LEVEL - 1, q_low, q_high, &res, s_scan);
    __syncthreads();
    // ranks that share a bucket share a histogram: count once, copy at the flush
    int owner[SEL_RANKS];
#pragma unroll
    for (int r = 0; r < SEL_RANKS; ++r) {
        owner[r] = r;
        if (LEVEL > 1)
            for (int o = r - 1; o >= 0; --o)
                if (res.prefix[o] == res.prefix[r]) owner[r] = o;
    }
    const int64_t stride = (int64_t)gridDim.x * 1024;
    const int64_t n_round = (n_px + 1023) / 1024 * 1024;                   // whole warps stay converged for match.any
    for (int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x; i < n_round; i += stride) {
        const bool valid = i < n_px && label_u[i] >= 0;
        const uint32_t k = valid ? float_key(entropy[i]) : 0u;
        if (LEVEL == 1) hist_add(s_hist, k >> 21, valid);
        else {
#pragma unroll
            for (int r = 0; r < SEL_RANKS; ++r) {
                if (owner[r] != r) continue;                                 // warp-uniform
                if (LEVEL == 2) hist_add(s_hist + r * BINS, (k >> 10) & 2047u, valid && (k >> 21) == res.prefix[r]);
                else hist_add(s_hist + r * BINS, k & 1023u, valid && (k >> 10) == res.prefix[r]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NH * BINS; i += 1024) {
        const int r = i / BINS, bin = i % BINS;
        const uint32_t c = s_hist[owner[LEVEL == 1 ? 0 : r] * BINS + bin];
        if (c) {
            if (LEVEL == 1) atomicAdd(&st->hist1[bin], c);
            else if (LEVEL == 2) atomicAdd(&st->hist2[r][bin], c);
            else atomicAdd(&st->hist3[r][bin], c);
        }
    }
}

// thresholds (numpy's float32 lerp) + the two masks for labelled and unlabelled images
__global__ void __launch_bounds__(1024) entropy_mask_kernel(const float* __restrict__ entropy, const int64_t* __restrict__ label_l,
                                                            const int64_t* __restrict__ label_u, int64_t n_l, int64_t n_u, float q_low,
                                                            float q_high, const SelectState* st, float* __restrict__ low_mask,
                                                            float* __restrict__ high_mask, float* __restrict__ thresholds) {
    __shared__ SelectResolved res;
    __shared__ uint32_t s_scan[1024];
    __shared__ float s_thr[2];
    select_resolve(st, 3, q_low, q_high, &res, s_scan);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 2; ++k) {
            const float a = key_float(res.prefix[2 * k]), b = key_float(res.prefix[2 * k + 1]);
            const float g = res.g[k];
            const float d = __fsub_rn(b, a);
            const float thr = g >= 0.5f ? __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.f, g))) : __fadd_rn(a, __fmul_rn(d, g));
            s_thr[k] = res.n > 0 ? thr : __int_as_float(0x7fc00000);          // empty selection: NaN like numpy
        }
        if (blockIdx.x == 0 && thresholds) { thresholds[0] = s_thr[0]; thresholds[1] = s_thr[1]; }
    }
    __syncthreads();
    if (low_mask == nullptr) return;                          // thresholds only (arco_entropy_thresholds)
    const float lo = s_thr[0], hi = s_thr[1];
    const int64_t total = n_l + n_u;
    for (int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 1024) {
        float lm, hm;
        if (i < n_l) {
            lm = hm = label_l[i] >= 0 ? 1.f : 0.f;
        } else {
            const int64_t u = i - n_l;
            const bool valid = label_u[u] >= 0;
            const float e = entropy[u];
            lm = (valid && e <= lo) ? 1.f : 0.f;
            hm = (valid && e >= hi) ? 1.f : 0.f;
        }
        low_mask[i] = lm;
        high_mask[i] = hm;
    }
}

}  // namespace arco

extern "C" int arco_softmax_rows(const float* logits, int64_t batch, int32_t classes, int64_t space, float* prob, float* entropy,
                                 void* stream) {
    ARCO_REQUIRE(logits && batch >= 0 && classes >= 1 && classes <= ARCO_MAX_CLASSES && space > 0 && (prob || entropy),
                 "arco_softmax_rows: bad argument");
    if (batch == 0) return ARCO_OK;
    const int64_t total = batch * space;
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (classes <= 4) arco::softmax_rows_kernel<4><<<grid, 256, 0, st>>>(logits, batch, classes, space, prob, entropy);
    else if (classes <= 8) arco::softmax_rows_kernel<8><<<grid, 256, 0, st>>>(logits, batch, classes, space, prob, entropy);
    else arco::softmax_rows_kernel<32><<<grid, 256, 0, st>>>(logits, batch, classes, space, prob, entropy);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int64_t arco_entropy_masks_scratch(void) { return (int64_t)sizeof(arco::SelectState); }

extern "C" int arco_entropy_masks(const float* entropy, const int64_t* label_l, const int64_t* label_u, int64_t n_lab_px,
                                  int64_t n_unlab_px, float q_low, float q_high, float* low_mask, float* high_mask,
                                  float* thresholds, void* scratch, void* stream) {
    ARCO_REQUIRE(low_mask && high_mask && scratch && n_lab_px >= 0 && n_unlab_px >= 0 && (n_unlab_px == 0 || (entropy && label_u)) &&
                     (n_lab_px == 0 || label_l), "arco_entropy_masks: bad argument");
    ARCO_REQUIRE(q_low >= 0.f && q_low <= 1.f && q_high >= 0.f && q_high <= 1.f, "quantiles must be in [0, 1]");
    ARCO_REQUIRE(n_unlab_px < (1ll << 32), "more than 2^32 unlabelled pixels");
    cudaStream_t st = (cudaStream_t)stream;
    arco::SelectState* s = (arco::SelectState*)scratch;
    ARCO_CUDA_CHECK(cudaMemsetAsync(s, 0, sizeof(arco::SelectState), st));
    const int grid = (int)((n_unlab_px + 4095) / 4096 < 148 ? (n_unlab_px + 4095) / 4096 : 148);
    if (n_unlab_px > 0) {
        arco::select_hist_kernel<1><<<grid, 1024, arco::SEL_B1 * 4, st>>>(entropy, label_u, n_unlab_px, q_low, q_high, s);
        arco::select_hist_kernel<2><<<grid, 1024, arco::SEL_RANKS * arco::SEL_B2 * 4, st>>>(entropy, label_u, n_unlab_px, q_low, q_high, s);
        arco::select_hist_kernel<3><<<grid, 1024, arco::SEL_RANKS * arco::SEL_B3 * 4, st>>>(entropy, label_u, n_unlab_px, q_low, q_high, s);
    }
    const int64_t total = n_lab_px + n_unlab_px;
    const int mgrid = (int)((total + 1023) / 1024 < 148 * 2 ? (total + 1023) / 1024 : 148 * 2);
    arco::entropy_mask_kernel<<<mgrid > 0 ? mgrid : 1, 1024, 0, st>>>(entropy, label_l, label_u, n_lab_px, n_unlab_px, q_low, q_high, s,
                                                                     low_mask, high_mask, thresholds);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_entropy_thresholds(const float* entropy, const int64_t* label_u, int64_t n_unlab_px, float q_low, float q_high,
                                       float* thresholds, void* scratch, void* stream) {
    ARCO_REQUIRE(entropy && label_u && thresholds && scratch && n_unlab_px > 0 && n_unlab_px < (1ll << 32), "arco_entropy_thresholds: bad argument");
    ARCO_REQUIRE(q_low >= 0.f && q_low <= 1.f && q_high >= 0.f && q_high <= 1.f, "quantiles must be in [0, 1]");
    cudaStream_t st = (cudaStream_t)stream;
    arco::SelectState* s = (arco::SelectState*)scratch;
    ARCO_CUDA_CHECK(cudaMemsetAsync(s, 0, sizeof(arco::SelectState), st));
    const int grid = (int)((n_unlab_px + 4095) / 4096 < 148 ? (n_unlab_px + 4095) / 4096 : 148);
    arco::select_hist_kernel<1><<<grid, 1024, arco::SEL_B1 * 4, st>>>(entropy, label_u, n_unlab_px, q_low, q_high, s);
    arco::select_hist_kernel<2><<<grid, 1024, arco::SEL_RANKS * arco::SEL_B2 * 4, st>>>(entropy, label_u, n_unlab_px, q_low, q_high, s);
    arco::select_hist_kernel<3><<<grid, 1024, arco::SEL_RANKS * arco::SEL_B3 * 4, st>>>(entropy, label_u, n_unlab_px, q_low, q_high, s);
    arco::entropy_mask_kernel<<<1, 1024, 0, st>>>(entropy, nullptr, label_u, 0, n_unlab_px, q_low, q_high, s, nullptr, nullptr, thresholds);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_prepare_contrast(const float* pred_u, const float* pred_l_teacher, const float* pred_u_teacher,
                                     const int64_t* label_l, const int64_t* label_u, int64_t n_lab, int64_t n_unlab, int32_t classes,
                                     int64_t space, float q_low, float q_high, float* prob_l_teacher, float* prob_u_teacher,
                                     float* entropy, float* low_mask, float* high_mask, float* thresholds, void* scratch,
                                     void* stream) {
    int rc;
    if (n_lab > 0 && (rc = arco_softmax_rows(pred_l_teacher, n_lab, classes, space, prob_l_teacher, nullptr, stream)) != ARCO_OK) return rc;
    if (n_unlab > 0) {
        if ((rc = arco_softmax_rows(pred_u_teacher, n_unlab, classes, space, prob_u_teacher, nullptr, stream)) != ARCO_OK) return rc;
        if ((rc = arco_softmax_rows(pred_u, n_unlab, classes, space, nullptr, entropy, stream)) != ARCO_OK) return rc;
    }
    return arco_entropy_masks(entropy, label_l, label_u, n_lab * space, n_unlab * space, q_low, q_high, low_mask, high_mask,
                              thresholds, scratch, stream);
}
