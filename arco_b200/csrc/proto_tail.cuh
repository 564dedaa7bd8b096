// In-kernel finalize of the prototype pass: the per-CTA partial rows [rows][C][D] (fp32) are folded into
// proto_sums[C][D+1] (fp64: feature sums, low-valid count) by the CTAs that finish LAST, instead of a second launch.
//
// Every CTA calls proto_finalize_tail() once, with all its threads, after its partial row is written.  The min(grid, 32)
// CTAs with the highest tickets wait until every CTA has arrived (a CTA takes its ticket after its own stores, so whoever
// it waits for is already running or done) and then fold blocks of 32 consecutive outputs each: lane <-> output (feature
// index fastest, so every load is a coalesced 128-byte line), the CTA's warps split the partial rows, fp64 partial sums per
// warp, combined in a fixed order -- deterministic run to run.
// The tickets live in the step's arco_plan (written fresh, as zero, by the plan derivation).
#pragma once
#include "arco_common.cuh"

namespace arco {

__device__ __forceinline__ uint32_t tail_ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void proto_finalize_tail(const float* __restrict__ partials, int rows, int C, int D,
                                                    arco_plan* plan, double* __restrict__ proto_sums) {
    if (proto_sums == nullptr) return;                           // A/B switch: the host launches proto_finalize_kernel instead
    if (proto_sums == nullptr) return;                           // A/B switch (ARCO_PROTO_TAIL=0): the host launches proto_finalize_kernel
    __shared__ int s_tail_role;
    __shared__ double s_tail_red[16][33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;   // nwarp <= 16
    const int grid = gridDim.x;
    const int n = C * (D + 1);
    const int nblk = (n + 31) >> 5;                              // blocks of 32 consecutive outputs (d fastest: coalesced)
    int K = grid < 32 ? grid : 32;
    if (K > nblk) K = nblk;
    __syncthreads();                                             // this CTA's partial row is complete
    if (tid == 0) {
        __threadfence();
        const uint32_t t = atomicAdd(&plan->proto_done, 1u);
        s_tail_role = (int)t >= grid - K ? (int)t - (grid - K) : -1;
    }
    __syncthreads();
    const int role = s_tail_role;
    if (role < 0) return;
    if (tid == 0) {
        while (tail_ld_acquire(&plan->proto_done) < (uint32_t)grid) __nanosleep(40);
    }
    __syncthreads();
    // lane <-> output, the CTA's warps split the rows (fixed partition), fp64 partial per warp, fixed-order combine
    for (int blk = role; blk < nblk; blk += K) {
        const int i = blk * 32 + lane;
        const bool valid = i < n;
        const int c = valid ? i / (D + 1) : 0, d = valid ? i % (D + 1) : 0;
        double s = 0.0;
        if (valid && d < D) {
            const float* src = partials + (int64_t)c * D + d;
            const int64_t rstride = (int64_t)C * D;
            // The fold is a chain of L2 round trips (a few hundred partial rows, nothing else left to run on the GPU): sixteen
            // independent loads in flight per lane -- with four, the LA shape's 592 rows took ~20 serial round trips (~12 us of
            // a 62 us kernel).
            int r = warp;
            for (; r + 15 * nwarp < rows; r += 16 * nwarp) {
                float a[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) a[u] = __ldcg(src + (int64_t)(r + u * nwarp) * rstride);
#pragma unroll
                for (int u = 0; u < 16; ++u) s += (double)a[u];
            }
            for (; r + 3 * nwarp < rows; r += 4 * nwarp) {
                const float a0 = __ldcg(src + (int64_t)r * rstride), a1 = __ldcg(src + (int64_t)(r + nwarp) * rstride);
                const float a2 = __ldcg(src + (int64_t)(r + 2 * nwarp) * rstride), a3 = __ldcg(src + (int64_t)(r + 3 * nwarp) * rstride);
                s += (double)a0; s += (double)a1; s += (double)a2; s += (double)a3;
            }
            for (; r < rows; r += nwarp) s += (double)__ldcg(src + (int64_t)r * rstride);
        }
        s_tail_red[warp][lane] = s;
        __syncthreads();
        if (warp == 0 && valid) {
            double t = 0.0;
            for (int w = 0; w < nwarp; ++w) t += s_tail_red[w][lane];
            proto_sums[i] = d == D ? (double)plan->lv_count[c] : t;
        }
        __syncthreads();
    }
    // re-arm the tickets (the entry point may be called again on the same plan, e.g. when a stage is timed on its own)
    if (tid == 0 && atomicAdd(&plan->proto_done2, 1u) == (uint32_t)K - 1) {
        plan->proto_done = 0;
        plan->proto_done2 = 0;
    }
}

// Stand-alone form of the same fold (ARCO_PROTO_TAIL=0: A/B measurements): one warp per output element.
__global__ void __launch_bounds__(128) proto_finalize_kernel(const float* __restrict__ partials, int rows, int C, int D,
                                                            const arco_plan* __restrict__ plan,
                                                            double* __restrict__ proto_sums);

}  // namespace arco
