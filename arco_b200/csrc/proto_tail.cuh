// In-kernel finalize of the prototype pass: the per-CTA partial rows [rows][C][D] (fp32) are folded into
// proto_sums[C][D+1] (fp64: feature sums, low-valid count) by the CTAs that finish LAST, instead of a second launch.
//
// Every CTA calls proto_finalize_tail() once, with all its threads, after its partial row is written.  The min(grid, 16)
// CTAs with the highest tickets wait until every CTA has arrived (a CTA takes its ticket after its own stores, so whoever
// it waits for is already running or done) and then fold a slice of the outputs each: one warp per output, lanes stride
// over the rows, fixed-order shuffle tree in fp64 -- the same deterministic reduction the separate kernel used.
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
    __shared__ int s_tail_role;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int grid = gridDim.x;
    const int K = grid < 16 ? grid : 16;
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
    const int n = C * (D + 1);
    for (int i = role * nwarp + warp; i < n; i += K * nwarp) {
        const int c = i / (D + 1), d = i % (D + 1);
        if (d == D) {
            if (lane == 0) proto_sums[i] = (double)plan->lv_count[c];
            continue;
        }
        double s = 0.0;
        for (int r = lane; r < rows; r += 32) s += (double)__ldcg(partials + ((int64_t)r * C + c) * D + d);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) proto_sums[i] = s;
    }
    // re-arm the tickets (the entry point may be called again on the same plan, e.g. when a stage is timed on its own)
    __syncthreads();
    if (tid == 0 && atomicAdd(&plan->proto_done2, 1u) == (uint32_t)K - 1) {
        plan->proto_done = 0;
        plan->proto_done2 = 0;
    }
}

}  // namespace arco
