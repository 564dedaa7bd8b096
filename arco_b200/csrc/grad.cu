// (a10) backward of the contrastive loss w.r.t. `rep`.
//
// The autograd contract forces a dense grad_rep [B,D,S]; the reference gets it from autograd's
// index backward (zeros + index_put_(accumulate=True) for the rows picked at loss_helper_3d.py:377,
// 455-457).  Here: one streaming zero-fill (P*D*e_g bytes written, the mandatory HBM term) and one
// scatter of grad_out * grad_anchor into the <= C*Q anchor pixels.  Duplicated anchors (sampling
// with replacement, trap 8) are summed in a fixed order by the first query that holds the pixel (no atomics).
#include "arco_common.cuh"

namespace arco {

// One 256-bit store per thread, one wave after another (no grid-stride loop): measured on B200 for 1.56 GB
// (scripts/probe/fill_probe.cu) 7430 GB/s, against 6424 for persistent 16-byte grid-stride stores and 7215 for cudaMemsetAsync.
__global__ void __launch_bounds__(256) fill_zero_kernel(unsigned char* __restrict__ dst, int64_t n32, int head_bytes, int tail_bytes) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n32) asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(dst + head_bytes + i * 32), "r"(0) : "memory");
    if (blockIdx.x == 0) {                                   // unaligned head / short tail, byte-wise
        if (threadIdx.x < head_bytes) dst[threadIdx.x] = 0;
        if (threadIdx.x < tail_bytes) dst[head_bytes + n32 * 32 + threadIdx.x] = 0;
    }
}

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// One CTA per (position j, query q).  Sampling is with replacement, so several queries of one position may hold the SAME
// pixel (trap 8: the reference's index_put_(accumulate=True) sums them).  Instead of atomics -- whose order, and with a
// bf16 grad_rep whose per-add rounding, changes from run to run -- the FIRST query of every pixel sums all its duplicates
// in ascending query order in fp32 and stores once: deterministic, one rounding, no read-modify-write.  (A pixel belongs to
// one class, so duplicates never cross positions; grad_rep is zero at every touched pixel on entry.)
template <typename T>
__global__ void __launch_bounds__(128) grad_scatter_kernel(const float* __restrict__ g_anchor,
                                                            const int32_t* __restrict__ anchor_pix,
                                                            const float* __restrict__ grad_out, T* __restrict__ grad_rep,
                                                            int D, int64_t S, int Q) {
    extern __shared__ int32_t s_dup[];                       // [Q] later queries of this position that hold the same pixel
    __shared__ int s_ndup;
    const int row = blockIdx.x;
    const int pix = anchor_pix[row];
    if (pix < 0) return;                                     // block-uniform
    const int j = row / Q, q = row - j * Q;
    const int tid = threadIdx.x;
    bool earlier = false, later = false;
    for (int k = tid; k < Q; k += 128) {
        const bool same = anchor_pix[j * Q + k] == pix;
        earlier |= same && k < q;
        later |= same && k > q;
    }
    if (__syncthreads_or(earlier)) return;                   // an earlier query owns this pixel
    const int any_later = __syncthreads_or(later);
    if (any_later) {                                         // rare: ordered list of the duplicates, built by one warp
        if (tid < 32) {
            int n = 0;
            for (int k0 = q + 1; k0 < Q; k0 += 32) {
                const int k = k0 + tid;
                const bool same = k < Q && anchor_pix[j * Q + k] == pix;
                const uint32_t bal = __ballot_sync(0xffffffffu, same);
                if (same) s_dup[n + __popc(bal & ((1u << tid) - 1u))] = k;
                n += __popc(bal);
            }
            if (tid == 0) s_ndup = n;
        }
        __syncthreads();
    }
    const int ndup = any_later ? s_ndup : 0;
    const float go = *grad_out;
    const int64_t b = pix / S, s = pix - b * S;
    for (int d = tid; d < D; d += 128) {
        float acc = go * g_anchor[(int64_t)row * D + d];
        for (int i = 0; i < ndup; ++i) acc += go * g_anchor[(int64_t)(j * Q + s_dup[i]) * D + d];
        grad_rep[(b * D + d) * S + s] = from_float<T>(acc);
    }
}

// Opt-in "sparse gradient" contract (arco_grad_scatter_sparse): grad_rep is a buffer the caller keeps across steps and that
// is zero everywhere except at the previous step's anchor pixels, so instead of rewriting P*D*e bytes of zeros it is enough
// to clear those <= C*Q pixel columns again before the new scatter.
template <typename T>
__global__ void __launch_bounds__(128) grad_unscatter_kernel(const int32_t* __restrict__ prev_pix, T* __restrict__ grad_rep,
                                                              int D, int64_t S) {
    const int pix = prev_pix[blockIdx.x];
    if (pix < 0) return;
    const int64_t b = pix / S, s = pix - b * S;
    for (int d = threadIdx.x; d < D; d += blockDim.x) grad_rep[(b * D + d) * S + s] = T(0.f);
}

__global__ void copy_pix_kernel(const int32_t* __restrict__ src, int32_t* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

}  // namespace arco

static int launch_fill(const arco_dims& d, void* grad_rep, cudaStream_t st) {
    const int64_t elems = ((int64_t)d.n_lab + d.n_unlab) * d.feat * d.space;
    const int64_t bytes = elems * (d.rep_dtype == ARCO_BF16 ? 2 : 4);
    int head = (int)((32 - ((uintptr_t)grad_rep & 31)) & 31);           // bytes up to the first 32-byte boundary
    if (head > bytes) head = (int)bytes;
    const int64_t n32 = (bytes - head) / 32;
    const int tail = (int)(bytes - head - n32 * 32);
    int64_t blocks = (n32 + 255) / 256;
    if (blocks < 1) blocks = 1;
    ARCO_REQUIRE(blocks < (1ll << 31), "grad_rep too large for one fill launch");
    arco::fill_zero_kernel<<<(unsigned)blocks, 256, 0, st>>>((unsigned char*)grad_rep, n32, head, tail);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

static int launch_scatter(const arco_dims& d, const float* grad_anchor, const int32_t* anchor_pix, const float* grad_out,
                          void* grad_rep, cudaStream_t st) {
    const int rows = d.classes * d.queries;
    ARCO_REQUIRE(d.queries <= 11264, "num_queries > 11264: the per-position pixel list no longer fits the scatter kernel's shared memory");
    if (d.rep_dtype == ARCO_BF16)
        arco::grad_scatter_kernel<__nv_bfloat16><<<rows, 128, (size_t)d.queries * 4, st>>>(grad_anchor, anchor_pix, grad_out,
                                                                                           (__nv_bfloat16*)grad_rep, d.feat, d.space, d.queries);
    else
        arco::grad_scatter_kernel<float><<<rows, 128, (size_t)d.queries * 4, st>>>(grad_anchor, anchor_pix, grad_out, (float*)grad_rep,
                                                                                   d.feat, d.space, d.queries);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_grad_zero(const arco_dims* dims, void* grad_rep, void* stream) {
    ARCO_REQUIRE(dims && grad_rep, "arco_grad_zero: NULL argument");
    return launch_fill(*dims, grad_rep, (cudaStream_t)stream);
}

extern "C" int arco_grad_scatter_add(const arco_dims* dims, const float* grad_anchor, const int32_t* anchor_pix,
                                     const float* grad_out, void* grad_rep, void* stream) {
    ARCO_REQUIRE(dims && grad_anchor && anchor_pix && grad_out && grad_rep, "arco_grad_scatter_add: NULL argument");
    return launch_scatter(*dims, grad_anchor, anchor_pix, grad_out, grad_rep, (cudaStream_t)stream);
}

extern "C" int arco_grad_scatter(const arco_dims* dims, const float* grad_anchor, const int32_t* anchor_pix,
                                 const float* grad_out, void* grad_rep, void* stream) {
    ARCO_REQUIRE(dims && grad_anchor && anchor_pix && grad_out && grad_rep, "arco_grad_scatter: NULL argument");
    int rc = launch_fill(*dims, grad_rep, (cudaStream_t)stream);
    if (rc != ARCO_OK) return rc;
    return launch_scatter(*dims, grad_anchor, anchor_pix, grad_out, grad_rep, (cudaStream_t)stream);
}

extern "C" int arco_grad_scatter_sparse(const arco_dims* dims, const float* grad_anchor, const int32_t* anchor_pix,
                                        const float* grad_out, void* grad_rep, int32_t* prev_pix, void* stream) {
    ARCO_REQUIRE(dims && grad_anchor && anchor_pix && grad_out && grad_rep && prev_pix, "arco_grad_scatter_sparse: NULL argument");
    const arco_dims& d = *dims;
    cudaStream_t st = (cudaStream_t)stream;
    const int rows = d.classes * d.queries;
    if (d.rep_dtype == ARCO_BF16)
        arco::grad_unscatter_kernel<__nv_bfloat16><<<rows, 128, 0, st>>>(prev_pix, (__nv_bfloat16*)grad_rep, d.feat, d.space);
    else
        arco::grad_unscatter_kernel<float><<<rows, 128, 0, st>>>(prev_pix, (float*)grad_rep, d.feat, d.space);
    ARCO_LAUNCH_CHECK();
    int rc = launch_scatter(d, grad_anchor, anchor_pix, grad_out, grad_rep, st);
    if (rc != ARCO_OK) return rc;
    arco::copy_pix_kernel<<<(rows + 255) / 256, 256, 0, st>>>(anchor_pix, prev_pix, rows);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}
