// the remaining per-pixel loss terms of the 2-D trainer's step, fused.
//
//   compute_unsupervised_loss        train_arco_2d.py:482-489   -> unsup_fwd_kernel / unsup_bwd_kernel
//   RandTPS grid (TPSGridGen)        tps/rand_tps.py:82-153, tps_stn_pytorch/tps_grid_gen.py:23-75 -> tps_grid_kernel
//   tps(x) = F.grid_sample(bilinear, align_corners=True)        tps/grid_sample.py:11-12 -> grid_sample_kernel
//   equivariance loss                train_arco_2d.py:404-423   -> eqv_fwd_kernel / scale_rows_kernel
//
// The reference runs each of them as a chain of ATen elementwise / reduction ops with full-size temporaries
// (softmax, log_softmax, KLDiv, two more grid_samples, masks, masked_select); every kernel here streams its inputs once
// and keeps the temporaries in registers.  All of them are HBM-bound elementwise / reduction work: one thread per pixel,
// channels strided by S (coalesced along the pixel axis), per-CTA partial sums folded in fp64 in a fixed order.
#include "arco_common.cuh"

namespace arco {

__device__ __forceinline__ float block_sum_256(float v, float* s_buf) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_buf[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? s_buf[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;                                                  // valid in thread 0
}

// ---------------------------------------------------------------------------------------------------------------------
// compute_unsupervised_loss(predict [B,C,S], target [B,S] (ignore = -1), logits [B,S], strong_threshold)
//   weighting[b] = #{logits[b] >= thr} / #{target[b] >= 0};  loss = CE(predict, target, ignore_index=-1, none);
//   result = mean over {loss > 0} of weighting[b] * loss                                    (:482-489)
// Forward: one pass; per (image, CTA) partials {sum of positive losses, #positive, #ge, #valid}.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unsup_fwd_kernel(const float* __restrict__ predict, const int64_t* __restrict__ target,
                                                       const float* __restrict__ logits, float thr, int C, int64_t S,
                                                       int ctas_per_image, float* __restrict__ partials) {
    __shared__ float s_buf[8];
    const int b = blockIdx.y;
    float sum = 0.f, npos = 0.f, nge = 0.f, nvalid = 0.f;
    for (int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x; s < S; s += (int64_t)ctas_per_image * 256) {
        const int64_t t = target[(int64_t)b * S + s];
        nge += logits[(int64_t)b * S + s] >= thr ? 1.f : 0.f;
        if (t >= 0) {
            nvalid += 1.f;
            const float* x = predict + (int64_t)b * C * S + s;
            float m = -INFINITY, se = 0.f, xt = 0.f;
            for (int c = 0; c < C; ++c) {                          // online log-sum-exp, one read per channel
                const float v = x[(int64_t)c * S];
                if (c == (int)t) xt = v;
                const float mn = fmaxf(m, v);
                se = se * __expf(m - mn) + __expf(v - mn);
                m = mn;
            }
            const float l = (m + __logf(se)) - xt;
            if (l > 0.f) { sum += l; npos += 1.f; }
        }
    }
    float* out = partials + ((size_t)b * ctas_per_image + blockIdx.x) * 4;
    float r;
    r = block_sum_256(sum, s_buf);    if (threadIdx.x == 0) out[0] = r;
    r = block_sum_256(npos, s_buf);   if (threadIdx.x == 0) out[1] = r;
    r = block_sum_256(nge, s_buf);    if (threadIdx.x == 0) out[2] = r;
    r = block_sum_256(nvalid, s_buf); if (threadIdx.x == 0) out[3] = r;
}

// stats: [B] weighting, then [1] 1/N, then [1] loss
__global__ void unsup_finish_kernel(const float* __restrict__ partials, int B, int ctas_per_image, float* __restrict__ stats) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double total = 0.0, n = 0.0;
    for (int b = 0; b < B; ++b) {
        double s = 0.0, np = 0.0, ge = 0.0, va = 0.0;
        for (int i = 0; i < ctas_per_image; ++i) {
            const float* p = partials + ((size_t)b * ctas_per_image + i) * 4;
            s += p[0]; np += p[1]; ge += p[2]; va += p[3];
        }
        const float w = (float)ge / (float)va;                     // int64 / float32 -> float32 (:486)
        stats[b] = w;
        total += (double)w * s;
        n += np;
    }
    stats[B] = (float)(1.0 / n);
    stats[B + 1] = (float)(total / n);
}

// d result / d predict[b,c,s] = grad_out * weighting[b] / N * (softmax_c - [c == target]) where loss > 0, else 0
__global__ void __launch_bounds__(256) unsup_bwd_kernel(const float* __restrict__ predict, const int64_t* __restrict__ target,
                                                       const float* __restrict__ stats, const float* __restrict__ grad_out, int B,
                                                       int C, int64_t S, float* __restrict__ grad) {
    const int b = blockIdx.y;
    const float scale = grad_out[0] * stats[b] * stats[B];
    for (int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x; s < S; s += (int64_t)gridDim.x * 256) {
        const int64_t t = target[(int64_t)b * S + s];
        const float* x = predict + (int64_t)b * C * S + s;
        float* g = grad + (int64_t)b * C * S + s;
        float m = -INFINITY, se = 0.f, xt = 0.f;
        if (t >= 0) {
            for (int c = 0; c < C; ++c) {
                const float v = x[(int64_t)c * S];
                if (c == (int)t) xt = v;
                const float mn = fmaxf(m, v);
                se = se * __expf(m - mn) + __expf(v - mn);
                m = mn;
            }
        }
        const float lse = m + __logf(se);
        const bool on = t >= 0 && (lse - xt) > 0.f;
        for (int c = 0; c < C; ++c) {
            float v = 0.f;
            if (on) v = scale * (__expf(x[(int64_t)c * S] - lse) - (c == (int)t ? 1.f : 0.f));
            g[(int64_t)c * S] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// TPS grid: grid[b,h,w,:] = [phi(|p - q_1|) ... phi(|p - q_N|), 1, x, y] @ mapping[b]   (tps_grid_gen.py:54-75)
//   p = (x, y) = (2w/(W-1) - 1, 2h/(H-1) - 1), phi(d2) = 0.5 * d2 * log(d2) with 0 log 0 := 0 (:9-21), q = the N control
//   points, mapping[b] = inverse_kernel @ [source_control_points[b]; 0] ([N+3, 2], host side: 28 x 28).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tps_grid_kernel(const float* __restrict__ mapping, const float* __restrict__ ctrl, int N, int B,
                                                      int H, int W, float* __restrict__ grid) {
    extern __shared__ float s_map[];                              // [N+3][2] of this image, then ctrl [N][2]
    const int b = blockIdx.y;
    float* s_ctrl = s_map + (N + 3) * 2;
    for (int i = threadIdx.x; i < (N + 3) * 2; i += 256) s_map[i] = mapping[(size_t)b * (N + 3) * 2 + i];
    for (int i = threadIdx.x; i < N * 2; i += 256) s_ctrl[i] = ctrl[i];
    __syncthreads();
    const int64_t HW = (int64_t)H * W;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < HW; i += (int64_t)gridDim.x * 256) {
        const int h = (int)(i / W), w = (int)(i - (int64_t)h * W);
        const float y = (float)h * 2.f / (float)(H - 1) - 1.f;     // Y * 2 / (H - 1) - 1 in float32 (:46-47)
        const float x = (float)w * 2.f / (float)(W - 1) - 1.f;
        float gx = 0.f, gy = 0.f;
        for (int k = 0; k < N; ++k) {
            const float dx = x - s_ctrl[2 * k], dy = y - s_ctrl[2 * k + 1];
            const float d2 = dx * dx + dy * dy;
            float r = 0.5f * d2 * logf(d2);
            if (r != r) r = 0.f;                                   // 0 * log(0) -> 0 (:19-20)
            gx += r * s_map[2 * k];
            gy += r * s_map[2 * k + 1];
        }
        gx += s_map[2 * N] + x * s_map[2 * (N + 1)] + y * s_map[2 * (N + 2)];
        gy += s_map[2 * N + 1] + x * s_map[2 * (N + 1) + 1] + y * s_map[2 * (N + 2) + 1];
        grid[((size_t)b * HW + i) * 2] = gx;
        grid[((size_t)b * HW + i) * 2 + 1] = gy;
    }
}

// F.grid_sample(bilinear, align_corners=True) coordinates and corner weights (zeros / border padding).
struct Bilinear {
    int x0, y0;                // top-left corner
    float w00, w01, w10, w11;  // weights of (y0,x0), (y0,x0+1), (y0+1,x0), (y0+1,x0+1)
    bool in00, in01, in10, in11;
};
__device__ __forceinline__ Bilinear bilinear_setup(float gx, float gy, int H, int W, bool border) {
    float ix = (gx + 1.f) * 0.5f * (float)(W - 1);                 // align_corners=True unnormalisation
    float iy = (gy + 1.f) * 0.5f * (float)(H - 1);
    if (border) {
        ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
        iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
    }
    const float fx = floorf(ix), fy = floorf(iy);
    Bilinear r;
    r.x0 = (int)fx; r.y0 = (int)fy;
    const float ax = ix - fx, ay = iy - fy;
    r.w00 = (1.f - ax) * (1.f - ay); r.w01 = ax * (1.f - ay); r.w10 = (1.f - ax) * ay; r.w11 = ax * ay;
    const bool xin0 = r.x0 >= 0 && r.x0 < W, xin1 = r.x0 + 1 >= 0 && r.x0 + 1 < W;
    const bool yin0 = r.y0 >= 0 && r.y0 < H, yin1 = r.y0 + 1 >= 0 && r.y0 + 1 < H;
    r.in00 = xin0 && yin0; r.in01 = xin1 && yin0; r.in10 = xin0 && yin1; r.in11 = xin1 && yin1;
    return r;
}
__device__ __forceinline__ float bilinear_fetch(const float* __restrict__ img, const Bilinear& q, int W) {
    const float* p = img + (int64_t)q.y0 * W + q.x0;
    float v = 0.f;
    if (q.in00) v += q.w00 * p[0];
    if (q.in01) v += q.w01 * p[1];
    if (q.in10) v += q.w10 * p[W];
    if (q.in11) v += q.w11 * p[W + 1];
    return v;
}

__global__ void __launch_bounds__(256) grid_sample_kernel(const float* __restrict__ in, const float* __restrict__ grid, int B, int C, int H,
                                                         int W, int border, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int64_t HW = (int64_t)H * W;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < HW; i += (int64_t)gridDim.x * 256) {
        const float2 g = *reinterpret_cast<const float2*>(grid + ((size_t)b * HW + i) * 2);
        const Bilinear q = bilinear_setup(g.x, g.y, H, W, border != 0);
        for (int c = 0; c < C; ++c) out[((int64_t)b * C + c) * HW + i] = bilinear_fetch(in + ((int64_t)b * C + c) * HW, q, W);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Equivariance loss (train_arco_2d.py:404-423), per output pixel of image b:
//   mask(src) = [labels != 0 and logits >= weak_threshold]  (:406-410),  mask_tps = tps(mask, 'zeros')       (:414)
//   t = softmax_c( tps(pred_d, 'zeros') ),  logp = log_softmax_c(pred_tps)                                   (:418-421)
//   kl = sum_c t_c (log t_c - logp_c);  loss = mean_b [ sum_px kl * mask_tps / (sum_px mask_tps + 1e-7) ]     (:420-423)
// Forward writes the UNSCALED gradient g = mask_tps * (softmax(pred_tps) - t) when asked to (backward only rescales it
// per image) and per-(image, CTA) partials {sum kl*mask, sum mask}.  MAXC channels live in registers.
// ---------------------------------------------------------------------------------------------------------------------
template <int MAXC>
__global__ void __launch_bounds__(256) eqv_fwd_kernel(const float* __restrict__ pred_tps, const float* __restrict__ pred_d,
                                                     const float* __restrict__ grid, const int64_t* __restrict__ labels,
                                                     const float* __restrict__ logits, float weak_thr, int B, int C, int H, int W,
                                                     int ctas_per_image, float* __restrict__ partials, float* __restrict__ g_out) {
    __shared__ float s_buf[8];
    const int b = blockIdx.y;
    const int64_t HW = (int64_t)H * W;
    float num = 0.f, den = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < HW; i += (int64_t)ctas_per_image * 256) {
        const float2 g = *reinterpret_cast<const float2*>(grid + ((size_t)b * HW + i) * 2);
        const Bilinear q = bilinear_setup(g.x, g.y, H, W, false);
        // mask_tps: bilinear sample of the 0/1 mask built on the fly at the four corners
        float mk = 0.f;
        {
            const int64_t* lb = labels + (int64_t)b * HW + (int64_t)q.y0 * W + q.x0;
            const float* lg = logits + (int64_t)b * HW + (int64_t)q.y0 * W + q.x0;
            if (q.in00 && lb[0] != 0 && !(lg[0] < weak_thr)) mk += q.w00;
            if (q.in01 && lb[1] != 0 && !(lg[1] < weak_thr)) mk += q.w01;
            if (q.in10 && lb[W] != 0 && !(lg[W] < weak_thr)) mk += q.w10;
            if (q.in11 && lb[W + 1] != 0 && !(lg[W + 1] < weak_thr)) mk += q.w11;
        }
        float t[MAXC], x[MAXC];
        float mt = -INFINITY, mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            if (c < C) {
                t[c] = bilinear_fetch(pred_d + ((int64_t)b * C + c) * HW, q, W);
                x[c] = pred_tps[((int64_t)b * C + c) * HW + i];
                mt = fmaxf(mt, t[c]);
                mx = fmaxf(mx, x[c]);
            }
        }
        float st = 0.f, sx = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            if (c < C) { st += __expf(t[c] - mt); sx += __expf(x[c] - mx); }
        }
        const float lst = mt + __logf(st), lsx = mx + __logf(sx);
        float kl = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            if (c < C) {
                const float logt = t[c] - lst, logp = x[c] - lsx;
                const float tc = __expf(logt);
                kl += tc > 0.f ? tc * (logt - logp) : 0.f;            // xlogy: 0 where the target is 0
                if (g_out) g_out[((int64_t)b * C + c) * HW + i] = mk * (__expf(logp) - tc);
            }
        }
        num += kl * mk;
        den += mk;
    }
    float* out = partials + ((size_t)b * ctas_per_image + blockIdx.x) * 2;
    float r;
    r = block_sum_256(num, s_buf); if (threadIdx.x == 0) out[0] = r;
    r = block_sum_256(den, s_buf); if (threadIdx.x == 0) out[1] = r;
}

// stats: [B] 1 / (B * (den_b + 1e-7)), then [1] loss
__global__ void eqv_finish_kernel(const float* __restrict__ partials, int B, int ctas_per_image, float* __restrict__ stats) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double total = 0.0;
    for (int b = 0; b < B; ++b) {
        double num = 0.0, den = 0.0;
        for (int i = 0; i < ctas_per_image; ++i) {
            num += partials[((size_t)b * ctas_per_image + i) * 2];
            den += partials[((size_t)b * ctas_per_image + i) * 2 + 1];
        }
        const float d = (float)den + 1e-7f;
        stats[b] = 1.f / ((float)B * d);
        total += (double)((float)num / d);
    }
    stats[B] = (float)(total / B);
}

// grad[b, :, :] = grad_out * scale[b] * g[b, :, :]   (in place allowed)
__global__ void __launch_bounds__(256) scale_rows_kernel(const float* __restrict__ g, const float* __restrict__ scale,
                                                        const float* __restrict__ grad_out, int64_t per_image, float* __restrict__ out) {
    const int b = blockIdx.y;
    const float s = grad_out[0] * scale[b];
    const float* src = g + (int64_t)b * per_image;
    float* dst = out + (int64_t)b * per_image;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < per_image; i += (int64_t)gridDim.x * 256) dst[i] = s * src[i];
}

static int ctas_per_image(int64_t S, int B) {
    int64_t want = (S + 1023) / 1024;
    const int64_t cap = (int64_t)sm_count() * 8 / (B > 0 ? B : 1) + 1;
    if (want > cap) want = cap;
    return (int)(want < 1 ? 1 : want);
}

}  // namespace arco

extern "C" int64_t arco_step_scratch_bytes(int32_t batch, int64_t space) {
    return (int64_t)batch * arco::ctas_per_image(space, batch) * 4 * 4 + 256;
}

extern "C" int arco_unsup_loss(const float* predict, const int64_t* target, const float* logits, float strong_threshold, int32_t batch,
                               int32_t classes, int64_t space, float* stats, void* scratch, void* stream) {
    ARCO_REQUIRE(predict && target && logits && stats && scratch && batch > 0 && classes > 0 && space > 0, "arco_unsup_loss: bad argument");
    const int cpi = arco::ctas_per_image(space, batch);
    cudaStream_t st = (cudaStream_t)stream;
    arco::unsup_fwd_kernel<<<dim3(cpi, batch), 256, 0, st>>>(predict, target, logits, strong_threshold, classes, space, cpi, (float*)scratch);
    ARCO_LAUNCH_CHECK();
    arco::unsup_finish_kernel<<<1, 32, 0, st>>>((const float*)scratch, batch, cpi, stats);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_unsup_loss_backward(const float* predict, const int64_t* target, const float* stats, const float* grad_out,
                                        int32_t batch, int32_t classes, int64_t space, float* grad_predict, void* stream) {
    ARCO_REQUIRE(predict && target && stats && grad_out && grad_predict, "arco_unsup_loss_backward: NULL argument");
    const int cpi = arco::ctas_per_image(space, batch);
    arco::unsup_bwd_kernel<<<dim3(cpi, batch), 256, 0, (cudaStream_t)stream>>>(predict, target, stats, grad_out, batch, classes, space, grad_predict);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_tps_grid(const float* mapping, const float* control_points, int32_t n_points, int32_t batch, int32_t height,
                             int32_t width, float* grid, void* stream) {
    ARCO_REQUIRE(mapping && control_points && grid && n_points > 0 && n_points <= 256 && batch > 0 && height > 1 && width > 1,
                 "arco_tps_grid: bad argument");
    const int cpi = arco::ctas_per_image((int64_t)height * width, batch);
    const size_t smem = (size_t)((n_points + 3) * 2 + n_points * 2) * 4;
    arco::tps_grid_kernel<<<dim3(cpi, batch), 256, smem, (cudaStream_t)stream>>>(mapping, control_points, n_points, batch, height, width, grid);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_grid_sample(const float* input, const float* grid, int32_t batch, int32_t channels, int32_t height, int32_t width,
                                int32_t border_padding, float* out, void* stream) {
    ARCO_REQUIRE(input && grid && out && batch > 0 && channels > 0 && height > 0 && width > 0, "arco_grid_sample: bad argument");
    const int cpi = arco::ctas_per_image((int64_t)height * width, batch);
    arco::grid_sample_kernel<<<dim3(cpi, batch), 256, 0, (cudaStream_t)stream>>>(input, grid, batch, channels, height, width, border_padding, out);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_eqv_loss(const float* pred_tps, const float* pred_detached, const float* grid, const int64_t* labels,
                             const float* logits, float weak_threshold, int32_t batch, int32_t classes, int32_t height, int32_t width,
                             float* stats, float* grad_unscaled, void* scratch, void* stream) {
    ARCO_REQUIRE(pred_tps && pred_detached && grid && labels && logits && stats && scratch, "arco_eqv_loss: NULL argument");
    ARCO_REQUIRE(batch > 0 && classes > 0 && classes <= 32 && height > 0 && width > 0, "arco_eqv_loss: bad sizes (classes <= 32)");
    const int cpi = arco::ctas_per_image((int64_t)height * width, batch);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 g(cpi, batch);
    if (classes <= 4)
        arco::eqv_fwd_kernel<4><<<g, 256, 0, st>>>(pred_tps, pred_detached, grid, labels, logits, weak_threshold, batch, classes, height, width, cpi, (float*)scratch, grad_unscaled);
    else if (classes <= 8)
        arco::eqv_fwd_kernel<8><<<g, 256, 0, st>>>(pred_tps, pred_detached, grid, labels, logits, weak_threshold, batch, classes, height, width, cpi, (float*)scratch, grad_unscaled);
    else
        arco::eqv_fwd_kernel<32><<<g, 256, 0, st>>>(pred_tps, pred_detached, grid, labels, logits, weak_threshold, batch, classes, height, width, cpi, (float*)scratch, grad_unscaled);
    ARCO_LAUNCH_CHECK();
    arco::eqv_finish_kernel<<<1, 32, 0, st>>>((const float*)scratch, batch, cpi, stats);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_scale_rows(const float* g, const float* scale, const float* grad_out, int32_t batch, int64_t per_image, float* out,
                               void* stream) {
    ARCO_REQUIRE(g && scale && grad_out && out && batch > 0 && per_image > 0, "arco_scale_rows: bad argument");
    const int cpi = arco::ctas_per_image(per_image, batch);
    arco::scale_rows_kernel<<<dim3(cpi, batch), 256, 0, (cudaStream_t)stream>>>(g, scale, grad_out, per_image, out);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}
