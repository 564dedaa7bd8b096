// Sub-systems 2 + 3a: ONE pass over rep_teacher [B,D,S] that produces
//   * per-class feature sums of the low-valid pixels (prototype numerators,
//     reference: torch.mean(rep_teacher[low_valid_pixel_seg.bool()], dim=0), loss_helper_3d.py:380-384)
//   * the ordered enqueue of negative keys into the device ring buffer
//     (reference: keys = rep_teacher[negative_mask]; dequeue_and_enqueue(...), loss_helper_3d.py:403-411,
//      12-32).  Only keys that survive eviction are written ("tail-only").
//
// Design (HBM-bound: algorithmic bytes = P_lv*D*e_t read + K*D*(e_t+4) ; the kernel reads every flagged
// 32-byte sector once):
//   * channel-first data is read as 16-byte vectors along S (coalesced 512-B runs per row) and staged in
//     shared memory TRANSPOSED to pixel-major [pixel][chunk-of-4-dims] with a rotation swizzle, so the
//     16-B stores and the per-pixel 16-B reads are both bank-conflict free;
//   * lanes own 4 feature dims, warps own pixel ranges, accumulators are warp-private in shared memory
//     -> no atomics, summation order is fixed, results are bit-reproducible run to run;
//   * persistent CTAs keep one 64-dim (or 32-dim) chunk of D for their whole life, so the per-class
//     accumulators are flushed once, to a [rows][C][D] partial buffer that a tiny fp64 kernel folds
//     into proto_sums[C][D+1] (the buffer a multi-GPU caller all-reduces).
#include "arco_common.cuh"

namespace arco {

struct ProtoParams {
    const void* rep_t;
    const uint8_t* codes;
    const uint32_t* tile_flagged;
    const uint32_t* off_key;
    const arco_plan* plan;
    float* bank_rows;
    float* partials;
    int64_t row_off[ARCO_MAX_CLASSES];
    int32_t cap[ARCO_MAX_CLASSES];
    int64_t S;
    int32_t B, C, D, tpi, NT, NDC;
    int32_t vec_ok;        // 16-byte loads along S are legal
};

template <typename T> struct Elem;
template <> struct Elem<float> { static constexpr int PER16 = 4; };
template <> struct Elem<__nv_bfloat16> { static constexpr int PER16 = 8; };

__device__ __forceinline__ float load_scalar(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_scalar(const __nv_bfloat16* p) {
    return bf16_bits_to_float(*reinterpret_cast<const unsigned short*>(p));
}

// NCH: 16-byte chunks per staged pixel row (D-chunk = 4*NCH dims).  8 warps, 256 threads.
template <typename T, int NCH>
__global__ void __launch_bounds__(256) proto_enqueue_kernel(ProtoParams p) {
    constexpr int PXS = 32 / NCH;            // pixels processed per warp step
    constexpr int SP = 2048 / NCH;           // pixels per staged sub-tile (32 KB of fp32)
    constexpr int NSUB = ARCO_TILE / SP;
    constexpr int WR = SP / 8;               // pixel range per warp
    constexpr int PER16 = Elem<T>::PER16;
    constexpr int ROT = (PER16 == 4) ? 2 : 3;  // rotation granularity keeps the 16-B stores conflict free
    constexpr int BLOCKS = NCH * (SP / PER16) / 256;
    static_assert(BLOCKS >= 1, "tile too small");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* tile = reinterpret_cast<float4*>(smem_raw);                       // [SP][NCH]
    float4* acc = tile + SP * NCH;                                            // [8][C][NCH]
    uint32_t* sc_words = reinterpret_cast<uint32_t*>(acc + 8 * p.C * NCH);    // [SP/4]
    uint32_t* wrun = sc_words + SP / 4;                                       // [8][32]
    __shared__ int32_t s_skip[ARCO_MAX_CLASSES], s_base[ARCO_MAX_CLASSES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = lane / NCH, ci = lane % NCH;
    const int C = p.C, D = p.D;
    const int64_t S = p.S;
    const int dchunk = blockIdx.x % p.NDC, grp = blockIdx.x / p.NDC, ngrp = gridDim.x / p.NDC;
    const int d0 = dchunk * NCH * 4;
    const int nch_real = min(NCH, (D - d0) / 4);
    const uint8_t* sc = reinterpret_cast<const uint8_t*>(sc_words);
    const T* rep = reinterpret_cast<const T*>(p.rep_t);

    for (int i = tid; i < 8 * C * NCH; i += 256) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < ARCO_MAX_CLASSES) {
        s_skip[tid] = p.plan->bank_skip[tid];
        s_base[tid] = p.plan->bank_write_base[tid];
    }
    __syncthreads();

    for (int t = grp; t < p.NT; t += ngrp) {
        if (p.tile_flagged[t] == 0) continue;                      // CTA-uniform
        const int b = t / p.tpi;
        const int64_t s_tile = (int64_t)(t % p.tpi) * ARCO_TILE;
        if (lane < C) wrun[warp * 32 + lane] = p.off_key[(int64_t)lane * (p.NT + 1) + t];
        for (int sub = 0; sub < NSUB; ++sub) {
            const int64_t s_sub = s_tile + (int64_t)sub * SP;
            if (s_sub >= S) break;                                 // CTA-uniform
            const int64_t gpx = (int64_t)b * S + s_sub;
            // ---- stage the code bytes of this sub-tile; find out whether anything is flagged ----
            uint32_t cw = 0;
            if (tid < SP / 4) {
                const int64_t s4 = s_sub + 4 * tid;
                if (p.vec_ok) {
                    if (s4 < S) cw = *reinterpret_cast<const uint32_t*>(p.codes + gpx + 4 * tid);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (s4 + k < S) cw |= (uint32_t)p.codes[gpx + 4 * tid + k] << (8 * k);
                }
                sc_words[tid] = cw;
            }
            const int any = __syncthreads_or(cw & 0xA0A0A0A0u);     // low-valid or key anywhere?
            if (any) {
                // ---- global -> registers -> transposed, swizzled shared tile ----
                if (p.vec_ok) {
                    uint4 raw[BLOCKS][4];
#pragma unroll
                    for (int it = 0; it < BLOCKS; ++it) {
                        const int id = it * 256 + tid;
                        const int pg = id % (SP / PER16), bc = id / (SP / PER16);
                        const int64_t s = s_sub + (int64_t)pg * PER16;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            raw[it][j] = make_uint4(0u, 0u, 0u, 0u);
                            if (bc < nch_real && s < S)
                                raw[it][j] = ldg_nc_u4(rep + ((int64_t)b * D + d0 + 4 * bc + j) * S + s);
                        }
                    }
#pragma unroll
                    for (int it = 0; it < BLOCKS; ++it) {
                        const int id = it * 256 + tid;
                        const int pg = id % (SP / PER16), bc = id / (SP / PER16);
#pragma unroll
                        for (int k = 0; k < PER16; ++k) {
                            float4 v;
                            if (PER16 == 4) {
                                const uint32_t* r0 = &raw[it][0].x; const uint32_t* r1 = &raw[it][1].x;
                                const uint32_t* r2 = &raw[it][2].x; const uint32_t* r3 = &raw[it][3].x;
                                v = make_float4(__uint_as_float(r0[k]), __uint_as_float(r1[k]),
                                                __uint_as_float(r2[k]), __uint_as_float(r3[k]));
                            } else {
                                const uint32_t* r0 = &raw[it][0].x; const uint32_t* r1 = &raw[it][1].x;
                                const uint32_t* r2 = &raw[it][2].x; const uint32_t* r3 = &raw[it][3].x;
                                const int w = k >> 1, sh = (k & 1) * 16;
                                v = make_float4(bf16_bits_to_float((r0[w] >> sh) & 0xffffu),
                                                bf16_bits_to_float((r1[w] >> sh) & 0xffffu),
                                                bf16_bits_to_float((r2[w] >> sh) & 0xffffu),
                                                bf16_bits_to_float((r3[w] >> sh) & 0xffffu));
                            }
                            const int pxl = pg * PER16 + k;
                            tile[pxl * NCH + ((bc + (pxl >> ROT)) & (NCH - 1))] = v;
                        }
                    }
                } else {
                    // scalar path (S not a multiple of the vector width, or unaligned base)
                    float* tf = reinterpret_cast<float*>(tile);
                    for (int e = tid; e < SP * NCH * 4; e += 256) {
                        const int pxl = e % SP, r = e / SP;
                        const int64_t s = s_sub + pxl;
                        float v = 0.f;
                        if (r < nch_real * 4 && s < S) v = load_scalar(rep + ((int64_t)b * D + d0 + r) * S + s);
                        const int bc = r >> 2;
                        tf[(pxl * NCH + ((bc + (pxl >> ROT)) & (NCH - 1))) * 4 + (r & 3)] = v;
                    }
                }
            }
            __syncthreads();
            if (any) {
                const int px0 = warp * WR;                                  // this warp's pixel range
                const uint32_t mycode = lane < WR ? sc[px0 + lane] : 0u;
                uint32_t lvmask = __ballot_sync(0xffffffffu, mycode & CODE_LV);
                uint32_t keymask = __ballot_sync(0xffffffffu, mycode & CODE_KEY);
                // ---- prototype accumulation: PXS pixels per step, lanes own chunk ci ----
                while (lvmask) {
                    int pj = -1;
#pragma unroll
                    for (int j = 0; j < PXS; ++j) {
                        const int q = lvmask ? __ffs(lvmask) - 1 : -1;
                        if (lvmask) lvmask &= lvmask - 1;
                        if (j == slot) pj = q;
                    }
                    const uint32_t code = __shfl_sync(0xffffffffu, mycode, pj < 0 ? 0 : pj);
                    int cls = pj < 0 ? -1 : (int)(code & CODE_CLS_MASK);
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (pj >= 0 && ci < nch_real) {
                        const int pxl = px0 + pj;
                        v = tile[pxl * NCH + ((ci + (pxl >> ROT)) & (NCH - 1))];
                    }
                    bool leader = cls >= 0;
                    if (PXS > 1) {
                        // fold slots that hit the same class into the lowest such slot (fixed order)
#pragma unroll
                        for (int j = 0; j < PXS; ++j) {
                            const int src = j * NCH + ci;
                            const int oc = __shfl_sync(0xffffffffu, cls, src);
                            const float ox = __shfl_sync(0xffffffffu, v.x, src);
                            const float oy = __shfl_sync(0xffffffffu, v.y, src);
                            const float oz = __shfl_sync(0xffffffffu, v.z, src);
                            const float ow = __shfl_sync(0xffffffffu, v.w, src);
                            if (oc == cls && cls >= 0) {
                                if (j < slot) leader = false;
                                else if (j > slot) { v.x += ox; v.y += oy; v.z += oz; v.w += ow; }
                            }
                        }
                    }
                    if (leader && ci < nch_real) {
                        float4* a = acc + ((warp * C + cls) * NCH + ci);
                        float4 o = *a;
                        o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
                        *a = o;
                    }
                    __syncwarp();
                }
                // ---- negative keys: ordered ring-buffer enqueue (FIFO order == raster order) ----
                if (__any_sync(0xffffffffu, keymask != 0)) {
                    uint32_t km = keymask;
                    while (km) {
                        const int pk = __ffs(km) - 1;
                        km &= km - 1;
                        const int cls = (int)(__shfl_sync(0xffffffffu, mycode, pk) & CODE_CLS_MASK);
                        const int target = px0 + pk;
                        // keys of the same class earlier in this sub-tile
                        int mine = 0;
                        for (int wq = lane; wq < SP / 4; wq += 32) {
                            const uint32_t w4 = sc_words[wq];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t cd = (w4 >> (8 * k)) & 0xffu;
                                mine += ((cd & CODE_KEY) && (int)(cd & CODE_CLS_MASK) == cls && (4 * wq + k) < target);
                            }
                        }
                        const int before = __reduce_add_sync(0xffffffffu, mine);
                        const int64_t ord = (int64_t)wrun[warp * 32 + cls] + before;
                        if (ord >= s_skip[cls] && lane < nch_real) {
                            const int64_t pos = ((int64_t)s_base[cls] + ord) % p.cap[cls];
                            const float4 v = tile[target * NCH + ((lane + (target >> ROT)) & (NCH - 1))];
                            float4* dst = reinterpret_cast<float4*>(p.bank_rows + (p.row_off[cls] + pos) * D + d0) + lane;
                            *dst = v;
                        }
                    }
                }
            }
            // ---- advance the per-warp running key ordinals by this sub-tile's keys (all warps agree) ----
            if (any) {
                __syncwarp();
                for (int wq = lane; wq < SP / 4; wq += 32) {
                    const uint32_t w4 = sc_words[wq];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t cd = (w4 >> (8 * k)) & 0xffu;
                        const bool is_key = cd & CODE_KEY;
                        const uint32_t peers = __match_any_sync(__activemask(), is_key ? (cd & CODE_CLS_MASK) : 0xffffu);
                        if (is_key && (__ffs(peers) - 1) == lane) wrun[warp * 32 + (cd & CODE_CLS_MASK)] += __popc(peers);
                        __syncwarp(__activemask());
                    }
                }
                __syncwarp();
            }
            __syncthreads();
        }
    }
    __syncthreads();
    // ---- fold the 8 warp-private accumulators (fixed order) and publish this CTA's partial row ----
    for (int i = tid; i < C * NCH; i += 256) {
        const int c = i / NCH, k = i % NCH;
        if (k >= nch_real) continue;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const float4 a = acc[(w * C + c) * NCH + k];
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
        *reinterpret_cast<float4*>(p.partials + ((int64_t)grp * C + c) * D + d0 + 4 * k) = s;
    }
}

__global__ void proto_finalize_kernel(const float* __restrict__ partials, int rows, int C, int D,
                                      const arco_plan* __restrict__ plan, double* __restrict__ proto_sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * (D + 1)) return;
    const int c = i / (D + 1), d = i % (D + 1);
    if (d == D) { proto_sums[i] = (double)plan->lv_count[c]; return; }
    double s = 0.0;
    for (int r = 0; r < rows; ++r) s += (double)partials[((int64_t)r * C + c) * D + d];
    proto_sums[i] = s;
}

static int proto_nch(const arco_dims& d) { return d.feat <= 32 ? 8 : 16; }

static size_t proto_smem_bytes(const arco_dims& d, int nch) {
    const int sp = 2048 / nch;
    return (size_t)sp * nch * 16 + (size_t)8 * d.classes * nch * 16 + sp + 8 * 32 * 4;
}

// grid geometry shared by the workspace layout and the launch
static void proto_grid(const arco_dims& d, int* ndc, int* groups) {
    const int nch = proto_nch(d);
    *ndc = (d.feat + nch * 4 - 1) / (nch * 4);
    // occupancy is bounded by shared memory: 32 KB tile + C*NCH*128 B accumulators
    const size_t smem = proto_smem_bytes(d, nch);
    int occ = (int)((size_t)(227 * 1024) / (smem + 1024));
    if (occ > 6) occ = 6;
    if (occ < 1) occ = 1;
    int g = sm_count() * occ / *ndc;
    if (g < 1) g = 1;
    *groups = g;
}

int proto_partial_rows(const arco_dims& d) {
    int ndc, groups;
    proto_grid(d, &ndc, &groups);
    return groups;
}

template <typename T>
static int launch_proto(const arco_dims& d, const ProtoParams& p, int groups, cudaStream_t st) {
    const int nch = proto_nch(d);
    const size_t smem = proto_smem_bytes(d, nch);
    const int grid = groups * p.NDC;
    if (nch == 16) {
        ARCO_CUDA_CHECK(cudaFuncSetAttribute(proto_enqueue_kernel<T, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        proto_enqueue_kernel<T, 16><<<grid, 256, smem, st>>>(p);
    } else {
        ARCO_CUDA_CHECK(cudaFuncSetAttribute(proto_enqueue_kernel<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        proto_enqueue_kernel<T, 8><<<grid, 256, smem, st>>>(p);
    }
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

}  // namespace arco

extern "C" int arco_proto_enqueue(const arco_dims* dims, const void* rep_teacher, const arco_bank* bank,
                                  double* proto_sums, void* workspace, void* stream) {
    ARCO_REQUIRE(dims && rep_teacher && bank && proto_sums && workspace, "arco_proto_enqueue: NULL argument");
    const arco_dims& d = *dims;
    ARCO_REQUIRE(d.feat >= 4 && d.feat % 4 == 0, "feat (D) must be a positive multiple of 4");
    arco_ws_layout L;
    arco::compute_layout(d, &L);
    char* ws = (char*)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    arco::ProtoParams p;
    p.rep_t = rep_teacher;
    p.codes = (const uint8_t*)(ws + L.codes);
    p.tile_flagged = (const uint32_t*)(ws + L.tile_flagged);
    p.off_key = (const uint32_t*)(ws + L.off_key);
    p.plan = (const arco_plan*)(ws + L.plan);
    p.bank_rows = bank->rows;
    p.partials = (float*)(ws + L.partials);
    for (int c = 0; c < ARCO_MAX_CLASSES; ++c) { p.row_off[c] = bank->row_off[c]; p.cap[c] = bank->cap[c] > 0 ? bank->cap[c] : 1; }
    p.S = d.space; p.B = d.n_lab + d.n_unlab; p.C = d.classes; p.D = d.feat;
    p.tpi = L.tiles_per_image; p.NT = L.n_tiles;
    int ndc, groups;
    arco::proto_grid(d, &ndc, &groups);
    p.NDC = ndc;
    const int per16 = d.rep_dtype == ARCO_BF16 ? 8 : 4;
    p.vec_ok = (d.space % per16 == 0) && (((uintptr_t)rep_teacher & 15) == 0);
    int rc = d.rep_dtype == ARCO_BF16 ? arco::launch_proto<__nv_bfloat16>(d, p, groups, st)
                                      : arco::launch_proto<float>(d, p, groups, st);
    if (rc != ARCO_OK) return rc;
    const int n = d.classes * (d.feat + 1);
    arco::proto_finalize_kernel<<<(n + 127) / 128, 128, 0, st>>>(p.partials, groups, d.classes, d.feat, p.plan, proto_sums);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}
