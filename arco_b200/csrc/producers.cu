// the representation producers, folded INTO the loss.
//
// Reference (train_arco_2d.py:317-333, model_2D.py:49-53, train_arco_2d.py:231-234):
//     rep_teacher = k_feature_extractor.fea4(x_t)                      one bias-free 1x1 convolution  W_k [D, D]
//     rep         = q_representation(q_feature_extractor.fea4(x_s))    three of them                   W_2 W_1 W_0
// on [24, 496, 256, 256] tensors (774 GFLOP per convolution), after which the contrastive loss reads of those tensors
//   * the per-class SUMS of rep_teacher over the low-valid pixels                (prototype, loss_helper_3d.py:380-384)
//   * the rows of rep_teacher at the K key pixels (~2 % of the pixels)            (enqueue,   :403-411)
//   * the rows of rep at the <= C*Q sampled anchor pixels (1024 of 1.57 M)        (:455-457)
// A 1x1 convolution without bias is linear and acts per pixel, so it commutes with every one of those selections:
//     sum_px (W x_px) = W (sum_px x_px),       (W x)[key pixel] = W (x[key pixel]),       rep[anchor] = W_2 W_1 W_0 x_s[anchor].
// The loss therefore runs its existing one-pass prototype/enqueue kernels over the convolutions' INPUT, and this file
// applies the weights afterwards, only where a value is consumed:
//   keys_transform_kernel   the ring rows enqueued by THIS step (located from the device plan) <- W_k . row, in place, as a
//                           tcgen05 GEMM: bf16 rings on kind::f16 (exact products, fp32 accumulation, one rounding to bf16 --
//                           what a bf16 autocast convolution does), fp32 rings on kind::tf32 with the three-term split
//                           a_hi w_hi + a_hi w_lo + a_lo w_hi (the dropped a_lo w_lo and the TF32 truncation of the lo parts
//                           are each <= 2^-22 relative)
//   proto_transform_kernel  the C x (D+1) fp64 class sums <- W_k . sums (fp64 accumulation), counts copied
//   (arco_anchor_gather / arco_infonce_rows in infonce.cu: anchors as rows; the student's three weights are applied to
//    those 1024 rows by the caller and autograd carries the gradient back through them)
// Neither rep nor rep_teacher is ever materialised; 4 x 774 GFLOP of forward convolutions and the student's dense backward
// become [K, D] x [D, D] and [C*Q, D] x [D, D] products.
#include <cuda.h>
#include <stdlib.h>

#include "arco_common.cuh"
#include "tc_common.cuh"

namespace arco {

constexpr int KT_BM = 128;                      // ring rows per tile (MMA M)
constexpr int KT_BN = 256;                      // output columns per MMA; two halves cover D <= 512 = all of tensor memory
constexpr int KT_A_BYTES = KT_BM * 128;         // one K block of the rows: 128 x 128 B (64 bf16 / 32 fp32 per row)
constexpr int KT_W_BYTES = KT_BN * 128;         // one K block of one half of the weight rows
constexpr int KT_MAX_SEG = 2 * ARCO_MAX_CLASSES;

struct KeysTransformParams {
    const arco_plan* plan;
    void* rows;
    int64_t row_off[ARCO_MAX_CLASSES];
    int32_t cap[ARCO_MAX_CLASSES];
    int32_t C, D;
};

// w_lo = w - (w & 0xFFFFE000): the part of an fp32 value the TF32 tensor core does not read
__global__ void __launch_bounds__(256) split_lo_kernel(const float* __restrict__ w, float* __restrict__ w_lo, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        const float x = w[i];
        w_lo[i] = __fsub_rn(x, __uint_as_float(__float_as_uint(x) & 0xFFFFE000u));
    }
}

// F32 = false: bf16 ring, bf16 weights, 2 stages of (rows 16 KB + weights 2 x 32 KB)
// F32 = true : fp32 ring, fp32 weights + their lo parts, 1 stage of (rows 16 KB + lo rows 16 KB + 4 x 32 KB)
template <bool F32>
__global__ void __launch_bounds__(256, 1) keys_transform_kernel(const __grid_constant__ CUtensorMap map_rows,
                                                                const __grid_constant__ CUtensorMap map_w,
                                                                const __grid_constant__ CUtensorMap map_wlo,
                                                                KeysTransformParams p) {
    constexpr int NST = F32 ? 1 : 2;
    constexpr int BK = F32 ? 32 : 64;                                     // elements per K block
    constexpr int KSTEP = F32 ? 8 : 16;                                   // elements per MMA
    constexpr int STAGE = F32 ? (2 * KT_A_BYTES + 4 * KT_W_BYTES) : (KT_A_BYTES + 2 * KT_W_BYTES);
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[NST], empty_bar[NST], cvt_bar[NST], tfull_bar, tempty_bar;
    __shared__ uint32_t s_tmem;
    __shared__ int32_t s_seg_row[KT_MAX_SEG], s_seg_len[KT_MAX_SEG], s_seg_tile0[KT_MAX_SEG + 1];
    __shared__ int32_t s_nseg;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = p.D;
    const int Dp = (D + 15) & ~15;
    const int n0 = Dp < KT_BN ? Dp : KT_BN, n1 = Dp - n0;                 // output columns of the two accumulators
    const int KB = (D + BK - 1) / BK;

    if (tid == 0) {
        // this step's written ring rows as <= 2 contiguous segments per class (plan_common.cuh: pos = (base + ord % cap) % cap)
        int ns = 0, nt = 0;
        for (int c = 0; c < p.C; ++c) {
            const int nk = (int)p.plan->n_key[c];
            const int cap = p.cap[c];
            if (nk <= 0 || cap <= 0) continue;
            const int r0 = (int)p.row_off[c];
            if (nk >= cap) {                                              // the whole ring was rewritten
                s_seg_row[ns] = r0; s_seg_len[ns] = cap; s_seg_tile0[ns] = nt; nt += (cap + KT_BM - 1) / KT_BM; ++ns;
            } else {
                const int wb = p.plan->bank_write_base[c];
                const int first = nk < cap - wb ? nk : cap - wb;
                s_seg_row[ns] = r0 + wb; s_seg_len[ns] = first; s_seg_tile0[ns] = nt; nt += (first + KT_BM - 1) / KT_BM; ++ns;
                if (nk > first) {
                    s_seg_row[ns] = r0; s_seg_len[ns] = nk - first; s_seg_tile0[ns] = nt; nt += (nk - first + KT_BM - 1) / KT_BM; ++ns;
                }
            }
        }
        s_seg_tile0[ns] = nt;
        s_nseg = ns;
        for (int s = 0; s < NST; ++s) { bar_init(&full_bar[s], 1); bar_init(&empty_bar[s], 1); bar_init(&cvt_bar[s], 128); }
        bar_init(&tfull_bar, 1); bar_init(&tempty_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const int n_tiles = s_seg_tile0[s_nseg];

    auto tile_rows = [&](int tile, int& row0, int& valid) {
        int sg = 0;
        while (tile >= s_seg_tile0[sg + 1]) ++sg;
        const int t = tile - s_seg_tile0[sg];
        row0 = s_seg_row[sg] + t * KT_BM;
        const int left = s_seg_len[sg] - t * KT_BM;
        valid = left < KT_BM ? left : KT_BM;
    };
    const uint32_t tx_bytes = (uint32_t)(KT_A_BYTES + (n1 > 0 ? 2 : 1) * KT_W_BYTES * (F32 ? 2 : 1));

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int row0, valid;
                tile_rows(tile, row0, valid);
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % NST;
                    bar_wait(&empty_bar[s], ((it / NST) & 1) ^ 1);
                    bar_expect_tx(&full_bar[s], tx_bytes);
                    unsigned char* dst = base + (size_t)s * STAGE;
                    tma_load_2d(dst, &map_rows, &full_bar[s], kb * BK, row0);
                    tma_load_2d(dst + KT_A_BYTES, &map_w, &full_bar[s], kb * BK, 0);
                    if (n1 > 0) tma_load_2d(dst + KT_A_BYTES + KT_W_BYTES, &map_w, &full_bar[s], kb * BK, KT_BN);
                    if (F32) {
                        tma_load_2d(dst + KT_A_BYTES + 2 * KT_W_BYTES, &map_wlo, &full_bar[s], kb * BK, 0);
                        if (n1 > 0) tma_load_2d(dst + KT_A_BYTES + 3 * KT_W_BYTES, &map_wlo, &full_bar[s], kb * BK, KT_BN);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc0 = F32 ? umma_idesc_tf32(KT_BM, n0) : umma_idesc_bf16(KT_BM, n0);
            const uint32_t idesc1 = F32 ? umma_idesc_tf32(KT_BM, n1 > 0 ? n1 : 16) : umma_idesc_bf16(KT_BM, n1 > 0 ? n1 : 16);
            uint32_t it = 0, ti = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
                bar_wait(&tempty_bar, (ti & 1) ^ 1);                      // the epilogue has drained the accumulators
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % NST;
                    bar_wait(&full_bar[s], (it / NST) & 1);
                    if (F32) bar_wait(&cvt_bar[s], (it / NST) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = s32(base + (size_t)s * STAGE);
                    const uint32_t w0 = a0 + KT_A_BYTES;
                    const uint32_t alo = a0 + KT_A_BYTES + 4 * KT_W_BYTES;   // F32 only
                    const int nterm = F32 ? 3 : 1;
                    for (int t = 0; t < nterm; ++t) {
                        // F32 terms: (a, w), (a, w_lo), (a_lo, w); the tensor core truncates a and w to their hi parts itself
                        const uint32_t aa = (t == 2) ? alo : a0;
                        const uint32_t ww = (t == 1) ? w0 + 2 * KT_W_BYTES : w0;
#pragma unroll
                        for (int kk = 0; kk < BK / KSTEP; ++kk) {
                            const uint32_t acc = (kb > 0 || t > 0 || kk > 0) ? 1u : 0u;
                            if (F32) {
                                umma_tf32(tmem, umma_desc(aa + kk * 32), umma_desc(ww + kk * 32), idesc0, acc);
                                if (n1 > 0) umma_tf32(tmem + KT_BN, umma_desc(aa + kk * 32), umma_desc(ww + KT_W_BYTES + kk * 32), idesc1, acc);
                            } else {
                                umma_bf16(tmem, umma_desc(aa + kk * 32), umma_desc(ww + kk * 32), idesc0, acc);
                                if (n1 > 0) umma_bf16(tmem + KT_BN, umma_desc(aa + kk * 32), umma_desc(ww + KT_W_BYTES + kk * 32), idesc1, acc);
                            }
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tfull_bar);
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;                                          // TMEM lanes 32*ew .. 32*ew+31 = tile rows
        const int et = tid - 128;
        uint32_t it = 0, ti = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
            int row0, valid;
            tile_rows(tile, row0, valid);
            if (F32) {
                // lo part of the staged rows, same swizzled position in a second buffer (pure element-wise map: thread t
                // handles the 16-byte chunks t, t+128, ... -> conflict-free)
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % NST;
                    bar_wait(&full_bar[s], (it / NST) & 1);
                    const uint4* src = reinterpret_cast<const uint4*>(base + (size_t)s * STAGE);
                    uint4* dst = reinterpret_cast<uint4*>(base + (size_t)s * STAGE + KT_A_BYTES + 4 * KT_W_BYTES);
#pragma unroll
                    for (int k = 0; k < KT_A_BYTES / 16 / 128; ++k) {
                        uint4 u = src[et + 128 * k];
                        uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            w[e] = __float_as_uint(__fsub_rn(__uint_as_float(w[e]), __uint_as_float(w[e] & 0xFFFFE000u)));
                        dst[et + 128 * k] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    bar_arrive(&cvt_bar[s]);
                }
            }
            bar_wait(&tfull_bar, ti & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int r = ew * 32 + lane;
            const bool ok = r < valid;
            unsigned char* out = reinterpret_cast<unsigned char*>(p.rows) + (int64_t)(row0 + r) * D * (F32 ? 4 : 2);
            for (int c0 = 0; c0 < Dp; c0 += 16) {
                uint32_t v[16];
                const uint32_t taddr = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ok && F32) {
                    float* o = reinterpret_cast<float*>(out) + c0;
                    if (c0 + 16 <= D) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            reinterpret_cast<uint4*>(o)[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    } else {
                        for (int i = 0; i < 16; ++i)
                            if (c0 + i < D) o[i] = __uint_as_float(v[i]);
                    }
                } else if (ok) {
                    unsigned short* o = reinterpret_cast<unsigned short*>(out) + c0;
                    uint32_t h[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                        h[i] = *reinterpret_cast<const uint32_t*>(&b2);
                    }
                    if (c0 + 16 <= D) {
                        reinterpret_cast<uint4*>(o)[0] = make_uint4(h[0], h[1], h[2], h[3]);
                        reinterpret_cast<uint4*>(o)[1] = make_uint4(h[4], h[5], h[6], h[7]);
                    } else {
                        for (int i = 0; i < 16; ++i)
                            if (c0 + i < D) o[i] = (unsigned short)((i & 1) ? (h[i >> 1] >> 16) : (h[i >> 1] & 0xffffu));
                    }
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid == 128) bar_arrive(&tempty_bar);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// out[c][o] = sum_i W[o][i] * in[c][i] (fp64), out[c][D] = in[c][D] (the count).  One warp per output element.
template <typename TW>
__global__ void __launch_bounds__(256) proto_transform_kernel(const TW* __restrict__ w, const double* __restrict__ in,
                                                              double* __restrict__ out, int C, int D) {
    const int o = blockIdx.x * 8 + (threadIdx.x >> 5), c = blockIdx.y, lane = threadIdx.x & 31;
    if (o > D) return;
    if (o == D) {
        if (lane == 0) out[(int64_t)c * (D + 1) + D] = in[(int64_t)c * (D + 1) + D];
        return;
    }
    const TW* wr = w + (int64_t)o * D;
    const double* x = in + (int64_t)c * (D + 1);
    double acc = 0.0;
    for (int i = lane; i < D; i += 32) acc = fma((double)(float)wr[i], x[i], acc);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) out[(int64_t)c * (D + 1) + o] = acc;
}

static int encode_2d(CUtensorMap* map, CUtensorMapDataType dt, int esz, void* ptr, uint64_t cols, uint64_t rows, uint32_t box_cols,
                     uint32_t box_rows) {
    EncodeTiledFn enc = encode_fn();
    ARCO_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstr[1] = {cols * (uint64_t)esz};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return ARCO_ERR_CUDA;
    }
    return ARCO_OK;
}

}  // namespace arco

extern "C" int64_t arco_keys_transform_scratch_bytes(int32_t feat, int32_t row_dtype) {
    return row_dtype == ARCO_F32 ? (int64_t)feat * feat * 4 : 0;
}

extern "C" int arco_keys_transform(const arco_dims* dims, const arco_bank* bank, const void* weight, void* scratch,
                                   void* workspace, void* stream) {
    using namespace arco;
    ARCO_REQUIRE(dims && bank && weight && workspace, "arco_keys_transform: NULL argument");
    const arco_dims& d = *dims;
    const bool f32 = bank->row_dtype == ARCO_F32;
    ARCO_REQUIRE(d.feat >= 16 && d.feat <= 512 && d.feat % (f32 ? 4 : 8) == 0,
                 "arco_keys_transform: D must be in [16, 512] and a multiple of 4 (fp32 ring) / 8 (bf16 ring)");
    ARCO_REQUIRE(!f32 || scratch != nullptr, "arco_keys_transform: an fp32 ring needs D*D*4 bytes of scratch");
    arco_ws_layout L;
    compute_layout(d, &L);
    int64_t total_rows = 0;
    for (int c = 0; c < d.classes; ++c) {
        const int64_t end = bank->row_off[c] + bank->cap[c];
        if (end > total_rows) total_rows = end;
    }
    ARCO_REQUIRE(total_rows > 0 && total_rows < (1ll << 31), "arco_keys_transform: empty or oversized ring");
    cudaStream_t st = (cudaStream_t)stream;
    KeysTransformParams p;
    p.plan = (const arco_plan*)((char*)workspace + L.plan);
    p.rows = bank->rows;
    for (int c = 0; c < ARCO_MAX_CLASSES; ++c) { p.row_off[c] = bank->row_off[c]; p.cap[c] = bank->cap[c]; }
    p.C = d.classes; p.D = d.feat;
    CUtensorMap map_rows, map_w, map_wlo;
    const int esz = f32 ? 4 : 2;
    const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const uint32_t bk = f32 ? 32 : 64;
    int rc = encode_2d(&map_rows, dt, esz, bank->rows, (uint64_t)d.feat, (uint64_t)total_rows, bk, KT_BM);
    if (rc != ARCO_OK) return rc;
    rc = encode_2d(&map_w, dt, esz, const_cast<void*>(weight), (uint64_t)d.feat, (uint64_t)d.feat, bk, KT_BN);
    if (rc != ARCO_OK) return rc;
    map_wlo = map_w;
    const int grid = sm_count();
    if (f32) {
        const int64_t n = (int64_t)d.feat * d.feat;
        split_lo_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float*)weight, (float*)scratch, n);
        rc = encode_2d(&map_wlo, dt, esz, scratch, (uint64_t)d.feat, (uint64_t)d.feat, bk, KT_BN);
        if (rc != ARCO_OK) return rc;
        const size_t smem = (size_t)(2 * KT_A_BYTES + 4 * KT_W_BYTES) + 1024;
        ARCO_CUDA_CHECK(cudaFuncSetAttribute(keys_transform_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        keys_transform_kernel<true><<<grid, 256, smem, st>>>(map_rows, map_w, map_wlo, p);
    } else {
        const size_t smem = (size_t)2 * (KT_A_BYTES + 2 * KT_W_BYTES) + 1024;
        ARCO_CUDA_CHECK(cudaFuncSetAttribute(keys_transform_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        keys_transform_kernel<false><<<grid, 256, smem, st>>>(map_rows, map_w, map_wlo, p);
    }
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_proto_transform(int32_t classes, int32_t feat, const void* weight, int32_t weight_dtype,
                                    const double* sums_in, double* sums_out, void* stream) {
    using namespace arco;
    ARCO_REQUIRE(weight && sums_in && sums_out && sums_in != sums_out, "arco_proto_transform: NULL or aliased argument");
    ARCO_REQUIRE(classes > 0 && classes <= ARCO_MAX_CLASSES && feat > 0, "arco_proto_transform: bad shape");
    const dim3 grid((unsigned)((feat + 1 + 7) / 8), (unsigned)classes);
    if (weight_dtype == ARCO_BF16)
        proto_transform_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)weight, sums_in, sums_out, classes, feat);
    else
        proto_transform_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)weight, sums_in, sums_out, classes, feat);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}
