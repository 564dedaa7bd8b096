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
#include <stdlib.h>

#include "arco_common.cuh"
#include "proto_tail.cuh"

namespace arco {

struct ProtoParams {
    const void* rep_t;
    const uint8_t* codes;
    const uint32_t* tile_flagged;
    const uint32_t* off_key;
    const arco_plan* plan;
    void* bank_rows;
    float* partials;
    int64_t row_off[ARCO_MAX_CLASSES];
    int32_t cap[ARCO_MAX_CLASSES];
    int64_t S;
    int32_t B, C, D, tpi, NT, NDC;
    int32_t vec_ok;        // 16-byte loads along S are legal
    int32_t bank_bf16;     // ring rows are bf16 (only with a bf16 rep_teacher: the narrowing is exact)
    double* proto_sums;    // [C][D+1] fp64, written by the in-kernel finalize (proto_tail.cuh)
    int32_t rows;          // partial rows (CTA groups)
};

// four floats that came from bf16 values -> their four bf16 bit patterns (exact: the low halves are zero)
__device__ __forceinline__ uint2 narrow4(const float4& v) {
    return make_uint2(__byte_perm(__float_as_uint(v.x), __float_as_uint(v.y), 0x7632u),
                      __byte_perm(__float_as_uint(v.z), __float_as_uint(v.w), 0x7632u));
}

template <typename T> struct Elem;
template <> struct Elem<float> { static constexpr int PER16 = 4; };
template <> struct Elem<__nv_bfloat16> { static constexpr int PER16 = 8; };

__device__ __forceinline__ float load_scalar(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_scalar(const __nv_bfloat16* p) {
    return bf16_bits_to_float(*reinterpret_cast<const unsigned short*>(p));
}

// NCH: 16-byte chunks per staged pixel row (D-chunk = 4*NCH dims).  8 warps, 256 threads.
// Each warp carries PXS = 32/NCH independent pixel STREAMS (a stream = NCH lanes, lane <-> 4 feature dims);
// stream s owns pixels [s*R, (s+1)*R) of every staged sub-tile and a private accumulator copy in shared
// memory, so there are no atomics and no cross-lane combines.  Consecutive pixels of one class are summed
// in registers and flushed on a class change (run merging).
template <typename T, int NCH>
__global__ void __launch_bounds__(256) proto_enqueue_kernel(ProtoParams p) {
    constexpr int PXS = 32 / NCH;            // streams per warp
    constexpr int NS = 8 * PXS;              // streams per CTA
    constexpr int SP = 2048 / NCH;           // pixels per staged sub-tile (32 KB of fp32)
    constexpr int NSUB = ARCO_TILE / SP;
    constexpr int R = SP / NS;               // pixels per stream per sub-tile (= 8)
    constexpr int PER16 = Elem<T>::PER16;
    constexpr int ROT = (PER16 == 4) ? 2 : 3;  // rotation granularity keeps the 16-B stores conflict free
    constexpr int BLOCKS = NCH * (SP / PER16) / 256;
    static_assert(BLOCKS >= 1 && R == 8, "tile geometry");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* tile = reinterpret_cast<float4*>(smem_raw);                       // [SP][NCH]
    float4* acc = tile + SP * NCH;                                            // [NS][C][NCH]
    uint32_t* sc_words = reinterpret_cast<uint32_t*>(acc + NS * p.C * NCH);   // [SP/4]
    uint32_t* wrun = sc_words + SP / 4;                                       // [8][32]
    __shared__ int32_t s_skip[ARCO_MAX_CLASSES], s_base[ARCO_MAX_CLASSES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = lane / NCH, ci = lane % NCH;
    const int stream = warp * PXS + slot;
    const int C = p.C, D = p.D;
    const int64_t S = p.S;
    const int dchunk = blockIdx.x % p.NDC, grp = blockIdx.x / p.NDC, ngrp = gridDim.x / p.NDC;
    const int d0 = dchunk * NCH * 4;
    const int nch_real = min(NCH, (D - d0) / 4);
    const bool lane_on = ci < nch_real;
    const T* rep = reinterpret_cast<const T*>(p.rep_t);
    float4* my_acc = acc + (size_t)stream * C * NCH + ci;

    for (int i = tid; i < NS * C * NCH; i += 256) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < ARCO_MAX_CLASSES) {
        s_skip[tid] = p.plan->bank_skip[tid];
        s_base[tid] = p.plan->bank_write_base[tid];
    }
    __syncthreads();

    int cur = -1;                                                            // class of the open run
    float4 racc = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int t = grp; t < p.NT; t += ngrp) {
        if (p.tile_flagged[t] == 0) continue;                                // CTA-uniform
        const int b = t / p.tpi;
        const int64_t s_tile = (int64_t)(t % p.tpi) * ARCO_TILE;
        if (lane < C) wrun[warp * 32 + lane] = p.off_key[(int64_t)lane * (p.NT + 1) + t];
        for (int sub = 0; sub < NSUB; ++sub) {
            const int64_t s_sub = s_tile + (int64_t)sub * SP;
            if (s_sub >= S) break;                                           // CTA-uniform
            const int64_t gpx = (int64_t)b * S + s_sub;
            // ---- stage the code bytes of this sub-tile; find out what is flagged ----
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
            const int any = __syncthreads_or(cw & 0xA0A0A0A0u);     // boolean: low-valid or key anywhere?
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
                        const uint32_t* r0 = &raw[it][0].x; const uint32_t* r1 = &raw[it][1].x;
                        const uint32_t* r2 = &raw[it][2].x; const uint32_t* r3 = &raw[it][3].x;
#pragma unroll
                        for (int k = 0; k < PER16; ++k) {
                            float4 v;
                            if (PER16 == 4) {
                                v = make_float4(__uint_as_float(r0[k]), __uint_as_float(r1[k]),
                                                __uint_as_float(r2[k]), __uint_as_float(r3[k]));
                            } else {
                                const int w = k >> 1;
                                if (k & 1)
                                    v = make_float4(__uint_as_float(r0[w] & 0xffff0000u), __uint_as_float(r1[w] & 0xffff0000u),
                                                    __uint_as_float(r2[w] & 0xffff0000u), __uint_as_float(r3[w] & 0xffff0000u));
                                else
                                    v = make_float4(__uint_as_float(r0[w] << 16), __uint_as_float(r1[w] << 16),
                                                    __uint_as_float(r2[w] << 16), __uint_as_float(r3[w] << 16));
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
                // what kind of work does the sub-tile hold (every warp derives the same answer)
                uint32_t agg = 0;
                for (int wq = lane; wq < SP / 4; wq += 32) agg |= sc_words[wq];
                const int flags = (__any_sync(0xffffffffu, agg & 0x20202020u) ? 1 : 0) |
                                  (__any_sync(0xffffffffu, agg & 0x80808080u) ? 2 : 0);
                // ---- this stream's 8 code bytes ----
                const int px0 = stream * R;
                const uint32_t w0 = sc_words[px0 / 4], w1 = sc_words[px0 / 4 + 1];
                if (flags & 1) {
                    // prototype accumulation: walk the low-valid pixels of the stream, merge same-class runs
                    uint32_t lv = ((w0 >> 5) & 1u) | ((w0 >> 12) & 2u) | ((w0 >> 19) & 4u) | ((w0 >> 26) & 8u) |
                                  ((w1 << 4 >> 5) & 16u) | ((w1 >> 8) & 32u) | ((w1 >> 15) & 64u) | ((w1 >> 22) & 128u);
                    while (lv) {
                        const int k = __ffs(lv) - 1;
                        lv &= lv - 1;
                        const uint32_t wsel = (k & 4) ? w1 : w0;
                        const int cls = (int)((wsel >> (8 * (k & 3))) & CODE_CLS_MASK);
                        const int pxl = px0 + k;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (lane_on) v = tile[pxl * NCH + ((ci + (pxl >> ROT)) & (NCH - 1))];
                        if (cls == cur) {
                            racc.x += v.x; racc.y += v.y; racc.z += v.z; racc.w += v.w;
                        } else {
                            if (cur >= 0 && lane_on) {
                                float4* a = my_acc + (size_t)cur * NCH;
                                float4 o = *a;
                                o.x += racc.x; o.y += racc.y; o.z += racc.z; o.w += racc.w;
                                *a = o;
                            }
                            cur = cls;
                            racc = v;
                        }
                    }
                }
                if (flags & 2) {
                    // ---- negative keys: ordered ring-buffer enqueue (FIFO order == raster order) ----
                    uint32_t kb = ((w0 >> 7) & 1u) | ((w0 >> 14) & 2u) | ((w0 >> 21) & 4u) | ((w0 >> 28) & 8u) |
                                  ((w1 << 4 >> 7) & 16u) | ((w1 >> 10) & 32u) | ((w1 >> 17) & 64u) | ((w1 >> 24) & 128u);
                    while (__any_sync(0xffffffffu, kb != 0)) {               // warp-uniform trip count
                        const bool have = kb != 0;
                        const int k = have ? __ffs(kb) - 1 : 0;
                        if (have) kb &= kb - 1;
                        const uint32_t wsel = (k & 4) ? w1 : w0;
                        const int cls = (int)((wsel >> (8 * (k & 3))) & CODE_CLS_MASK);
                        const int target = px0 + k;
                        // keys of the same class earlier in this sub-tile (counted by the stream's NCH lanes)
                        int mine = 0;
                        for (int wq = ci; wq < SP / 4; wq += NCH) {
                            const uint32_t w4 = sc_words[wq];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t cd = (w4 >> (8 * q)) & 0xffu;
                                mine += ((cd & CODE_KEY) && (int)(cd & CODE_CLS_MASK) == cls && (4 * wq + q) < target);
                            }
                        }
#pragma unroll
                        for (int o = NCH / 2; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
                        if (have && lane_on) {
                            const int64_t ord = (int64_t)wrun[warp * 32 + cls] + mine;
                            if (ord >= s_skip[cls]) {
                                const int64_t pos = ((int64_t)s_base[cls] + ord) % p.cap[cls];
                                const float4 v = tile[target * NCH + ((ci + (target >> ROT)) & (NCH - 1))];
                                if (p.bank_bf16)
                                    reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(p.bank_rows) +
                                                             (p.row_off[cls] + pos) * D + d0)[ci] = narrow4(v);
                                else
                                    reinterpret_cast<float4*>(reinterpret_cast<float*>(p.bank_rows) +
                                                              (p.row_off[cls] + pos) * D + d0)[ci] = v;
                            }
                        }
                    }
                    // advance the per-warp running key ordinals by this sub-tile's keys (all warps agree)
                    __syncwarp();
                    for (int wb = 0; wb < SP / 4; wb += 32) {                // warp-uniform trip count (match.any inside)
                        const int wq = wb + lane;
                        const uint32_t w4 = wq < SP / 4 ? sc_words[wq] : 0u;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t cd = (w4 >> (8 * q)) & 0xffu;
                            const bool is_key = cd & CODE_KEY;
                            const uint32_t peers = __match_any_sync(0xffffffffu, is_key ? (cd & CODE_CLS_MASK) : 0xffffu);
                            if (is_key && (__ffs(peers) - 1) == lane) wrun[warp * 32 + (cd & CODE_CLS_MASK)] += __popc(peers);
                            __syncwarp();
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
    // ---- close the open runs, fold the stream-private accumulators (fixed order), publish the partial row ----
    if (cur >= 0 && lane_on) {
        float4* a = my_acc + (size_t)cur * NCH;
        float4 o = *a;
        o.x += racc.x; o.y += racc.y; o.z += racc.z; o.w += racc.w;
        *a = o;
    }
    __syncthreads();
    for (int i = tid; i < C * NCH; i += 256) {
        const int c = i / NCH, k = i % NCH;
        if (k >= nch_real) continue;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int w = 0; w < NS; ++w) {
            const float4 a = acc[((size_t)w * C + c) * NCH + k];
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
        *reinterpret_cast<float4*>(p.partials + ((int64_t)grp * C + c) * D + d0 + 4 * k) = s;
    }
    proto_finalize_tail(p.partials, p.rows, C, D, const_cast<arco_plan*>(p.plan), p.proto_sums);
}

// ---------------------------------------------------------------------------------------------------
// Software-pipelined main kernel.  EPL = feature dims per lane:
//   EPL = 8 (C <= 8): fp32 data is staged as two float4 planes, bf16 data stays PACKED in shared memory
//                     (one uint4 = 8 dims) and is widened only when it is accumulated; D-chunk = 8*NCH;
//                     stream-private accumulators = 8 KB * C.
//   EPL = 4 (C  > 8): one float4 plane (bf16 widened at commit); D-chunk = 4*NCH; accumulators = 4 KB * C.
// Tile = 32 KB.  Stages: A (issue) global loads of sub-tile i+1 into registers, B (commit) registers ->
// transposed + rotation-swizzled shared tile, C (compute) per-stream accumulation and key enqueue of
// sub-tile i.  The loads of A stay in flight during C.
// ---------------------------------------------------------------------------------------------------
template <typename T> struct Wide;
template <> struct Wide<float> { static constexpr int PER16 = 4; };
template <> struct Wide<__nv_bfloat16> { static constexpr int PER16 = 8; };

template <typename T, int NCH, int EPL>
__global__ void __launch_bounds__(256, 2) proto_pipe_kernel(ProtoParams p) {
    constexpr int PER16 = Wide<T>::PER16;    // pixels per 16-byte global load
    constexpr bool PACKED = (EPL == 8 && PER16 == 8);            // bf16 kept packed in shared memory
    constexpr int LW = (EPL == 8 && PER16 == 4) ? 2 : 1;        // 16-byte words per (pixel, lane)
    constexpr int PXS = 32 / NCH;
    constexpr int NS = 8 * PXS;
    constexpr int SP = 2048 / (NCH * LW);    // pixels per staged sub-tile (32 KB)
    constexpr int NSUB = ARCO_TILE / SP;
    constexpr int R = SP / NS;               // pixels per stream per sub-tile (4 or 8)
    constexpr int RW = R / 4;                // code words per stream
    constexpr int ROT = (PER16 == 4) ? 2 : 3;   // swizzle rotation = pixel >> ROT: conflict-free 16-byte commits
    constexpr int RG = (R >> ROT) > 0 ? (R >> ROT) : 1;         // rotation groups inside a stream's R pixels
    constexpr int PLANE = SP * NCH;          // uint4 elements per plane
    constexpr int NBLK = NCH * (SP / PER16) / 256;              // (EPL rows x PER16 px) loader blocks per thread
    constexpr int AP = EPL / 4;              // accumulator planes (float4 each)
    static_assert(NBLK * 256 == NCH * (SP / PER16) && NBLK * EPL == 8 || (NBLK == 1 && EPL == 4), "loader geometry");
    static_assert(R == 4 || R == 8, "stream geometry");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* tile = reinterpret_cast<uint4*>(smem_raw);                         // [LW][SP][NCH]
    float4* acc = reinterpret_cast<float4*>(tile + LW * PLANE);               // [AP][NS][C][NCH]
    uint32_t* sc_tile = reinterpret_cast<uint32_t*>(acc + AP * NS * p.C * NCH);  // [2][256] code words of a tile, double buffered
    uint32_t* wrun = sc_tile + 2 * (ARCO_TILE / 4);                            // [8][32]
    __shared__ int32_t s_skip[ARCO_MAX_CLASSES], s_base[ARCO_MAX_CLASSES], s_cap[ARCO_MAX_CLASSES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = lane / NCH, ci = lane % NCH;
    const int stream = warp * PXS + slot;
    const int C = p.C, D = p.D;
    const int64_t S = p.S;
    const int dchunk = blockIdx.x % p.NDC, grp = blockIdx.x / p.NDC, ngrp = gridDim.x / p.NDC;
    const int d0 = dchunk * NCH * EPL;
    const int dreal = min(NCH * EPL, D - d0);                                 // multiple of 4
    const int rows_on = max(0, min(EPL, dreal - EPL * ci));                   // real dims in this lane (0, 4 or 8)
    const T* rep = reinterpret_cast<const T*>(p.rep_t);
    const int acc_plane = NS * C * NCH;
    float4* my_acc = acc + (size_t)stream * C * NCH + ci;

    for (int i = tid; i < AP * acc_plane; i += 256) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < ARCO_MAX_CLASSES) {
        s_skip[tid] = p.plan->bank_skip[tid];
        s_base[tid] = p.plan->bank_write_base[tid];
        s_cap[tid] = p.cap[tid];
    }

    int cur = -1;
    float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;

    auto flush = [&]() {
        if (cur >= 0 && rows_on > 0) {
            float4* a = my_acc + (size_t)cur * NCH;
            float4 o = a[0];
            o.x += ra.x; o.y += ra.y; o.z += ra.z; o.w += ra.w;
            a[0] = o;
            if (EPL == 8 && rows_on > 4) {
                o = a[acc_plane];
                o.x += rb.x; o.y += rb.y; o.z += rb.z; o.w += rb.w;
                a[acc_plane] = o;
            }
        }
    };
    // the EPL dims of a staged (pixel, lane) cell as float4 lo (dims 0-3) and hi (dims 4-7, EPL == 8 only)
    auto widen = [&](const uint4* cell, float4& lo, float4& hi) {
        const uint4 u = cell[0];
        if (PACKED) {                                  // 8 packed bf16: dims (0,1) (2,3) (4,5) (6,7)
            lo = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                             __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
            hi = make_float4(__uint_as_float(u.z << 16), __uint_as_float(u.z & 0xffff0000u),
                             __uint_as_float(u.w << 16), __uint_as_float(u.w & 0xffff0000u));
        } else {
            lo = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
            hi = make_float4(0.f, 0.f, 0.f, 0.f);
            if (LW == 2) {
                const uint4 w = cell[PLANE];
                hi = make_float4(__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w));
            }
        }
    };
    auto cell_of = [&](int pxl) { return tile + pxl * NCH + ((ci + (pxl >> ROT)) & (NCH - 1)); };

    int t_next = grp, sub_next = 0, par_next = 0;           // iterator of the issue stage
    auto skip_unflagged = [&]() {
        while (t_next < p.NT && p.tile_flagged[t_next] == 0) t_next += ngrp;
    };
    skip_unflagged();
    uint4 raw[NBLK * EPL];
    uint32_t cw_reg = 0;
    int t_ld = -1, sub_ld = 0, par_ld = 0;                   // what `raw` currently holds
    auto issue = [&]() {
        t_ld = t_next; sub_ld = sub_next; par_ld = par_next;
        if (t_next >= p.NT) { t_ld = -1; return; }
        const int b = t_next / p.tpi;
        const int64_t s_tile = (int64_t)(t_next % p.tpi) * ARCO_TILE;
#pragma unroll
        for (int blk = 0; blk < NBLK; ++blk) {
            // loader block: rows [EPL*bc, +EPL) x pixels [pg*PER16, +PER16) of the sub-tile
            const int id = blk * 256 + tid;
            const int pg = id % (SP / PER16), bc = id / (SP / PER16);
            const int ld_rows = max(0, min(EPL, dreal - EPL * bc));
            const int64_t s = s_tile + (int64_t)sub_next * SP + (int64_t)pg * PER16;
            const T* rowp = rep + ((int64_t)b * D + d0 + EPL * bc) * S + s;
            if (ld_rows == EPL && s < S) {
#pragma unroll
                for (int j = 0; j < EPL; ++j) { raw[blk * EPL + j] = ldg_nc_u4(rowp); rowp += S; }
            } else {
                const bool in = s < S;
#pragma unroll
                for (int j = 0; j < EPL; ++j) {
                    raw[blk * EPL + j] = make_uint4(0u, 0u, 0u, 0u);
                    if (j < ld_rows && in) raw[blk * EPL + j] = ldg_nc_u4(rowp + (int64_t)j * S);
                }
            }
        }
        if (sub_next == 0) {
            const int64_t s4 = s_tile + 4 * tid;
            cw_reg = s4 < S ? *reinterpret_cast<const uint32_t*>(p.codes + (int64_t)b * S + s4) : 0u;
        }
        ++sub_next;
        if (sub_next == NSUB || s_tile + (int64_t)sub_next * SP >= S) {
            sub_next = 0;
            par_next ^= 1;
            t_next += ngrp;
            skip_unflagged();
        }
    };
    issue();
    __syncthreads();                                         // accumulators zeroed, s_* visible
    while (t_ld >= 0) {
        const int t = t_ld, sub = sub_ld, par = par_ld;
        // ---- stage B: commit registers to the shared tile ----
#pragma unroll
        for (int blk = 0; blk < NBLK; ++blk) {
            const int id = blk * 256 + tid;
            const int pg = id % (SP / PER16), bc = id / (SP / PER16);
            const uint32_t* r0 = &raw[blk * EPL + 0].x; const uint32_t* r1 = &raw[blk * EPL + 1].x;
            const uint32_t* r2 = &raw[blk * EPL + 2].x; const uint32_t* r3 = &raw[blk * EPL + 3].x;
#pragma unroll
            for (int k = 0; k < PER16; ++k) {
                const int pxl = pg * PER16 + k;
                const int idx = pxl * NCH + ((bc + (pxl >> ROT)) & (NCH - 1));
                if (EPL == 8) {
                    const uint32_t* r4 = &raw[blk * EPL + (EPL == 8 ? 4 : 0)].x; const uint32_t* r5 = &raw[blk * EPL + (EPL == 8 ? 5 : 0)].x;
                    const uint32_t* r6 = &raw[blk * EPL + (EPL == 8 ? 6 : 0)].x; const uint32_t* r7 = &raw[blk * EPL + (EPL == 8 ? 7 : 0)].x;
                    if (PACKED) {
                        const int w = k >> 1;
                        const uint32_t sel = (k & 1) ? 0x7632u : 0x5410u;         // pick the k-th bf16 of two rows
                        tile[idx] = make_uint4(__byte_perm(r0[w], r1[w], sel), __byte_perm(r2[w], r3[w], sel),
                                               __byte_perm(r4[w], r5[w], sel), __byte_perm(r6[w], r7[w], sel));
                    } else {
                        tile[idx] = make_uint4(r0[k], r1[k], r2[k], r3[k]);
                        tile[PLANE + idx] = make_uint4(r4[k], r5[k], r6[k], r7[k]);
                    }
                } else if (PER16 == 4) {
                    tile[idx] = make_uint4(r0[k], r1[k], r2[k], r3[k]);
                } else {
                    const int w = k >> 1;                                          // widen bf16 -> fp32 bits
                    if (k & 1) tile[idx] = make_uint4(r0[w] & 0xffff0000u, r1[w] & 0xffff0000u, r2[w] & 0xffff0000u, r3[w] & 0xffff0000u);
                    else tile[idx] = make_uint4(r0[w] << 16, r1[w] << 16, r2[w] << 16, r3[w] << 16);
                }
            }
        }
        if (sub == 0) sc_tile[par * (ARCO_TILE / 4) + tid] = cw_reg;
        __syncthreads();
        // ---- stage A for the next sub-tile: loads stay in flight during the compute below ----
        issue();
        // ---- stage C ----
        const uint32_t* sc_words = sc_tile + par * (ARCO_TILE / 4) + sub * (SP / 4);
        if (sub == 0 && lane < C) wrun[warp * 32 + lane] = p.off_key[(int64_t)lane * (p.NT + 1) + t];
        __syncwarp();
        uint32_t agg = 0;
        for (int wq = lane; wq < SP / 4; wq += 32) agg |= sc_words[wq];
        const bool has_lv = __any_sync(0xffffffffu, agg & 0x20202020u);
        const bool has_key = __any_sync(0xffffffffu, agg & 0x80808080u);
        const int px0 = stream * R;
        const uint32_t w0 = sc_words[px0 / 4];
        const uint32_t w1 = RW == 2 ? sc_words[px0 / 4 + 1] : 0u;
        if (has_lv) {
            // R pixels per stream, unrolled: bit tests, class extraction and the tile offsets are compile-time
            // constants relative to RG per-stream base pointers (pixels of one rotation group share a rotation).
            const uint4* tbase[RG];
#pragma unroll
            for (int g = 0; g < RG; ++g) tbase[g] = cell_of(px0 + (g << ROT));
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const uint32_t wsel = (k & 4) ? w1 : w0;
                const uint32_t code = wsel >> (8 * (k & 3));
                if (code & CODE_LV) {
                    const int cls = (int)(code & CODE_CLS_MASK);
                    float4 lo, hi;
                    widen(tbase[RG == 1 ? 0 : (k >> ROT)] + (RG == 1 ? k : (k & ((1 << ROT) - 1))) * NCH, lo, hi);
                    if (cls != cur) {
                        flush();
                        cur = cls;
                        ra = lo; rb = hi;
                    } else {
                        ra.x += lo.x; ra.y += lo.y; ra.z += lo.z; ra.w += lo.w;
                        if (EPL == 8) { rb.x += hi.x; rb.y += hi.y; rb.z += hi.z; rb.w += hi.w; }
                    }
                }
            }
        }
        if (has_key) {
            uint32_t kb = ((w0 >> 7) & 1u) | ((w0 >> 14) & 2u) | ((w0 >> 21) & 4u) | ((w0 >> 28) & 8u) |
                          ((w1 >> 3) & 16u) | ((w1 >> 10) & 32u) | ((w1 >> 17) & 64u) | ((w1 >> 24) & 128u);
            while (__any_sync(0xffffffffu, kb != 0)) {                       // warp-uniform trip count
                const bool have = kb != 0;
                const int k = have ? __ffs(kb) - 1 : 0;
                if (have) kb &= kb - 1;
                const uint32_t wsel = (k & 4) ? w1 : w0;
                const int cls = (int)((wsel >> (8 * (k & 3))) & CODE_CLS_MASK);
                const int target = px0 + k;
                // keys of the same class earlier in this sub-tile (counted by the stream's NCH lanes)
                const uint32_t want = CODE_KEY | (uint32_t)cls;
                int mine = 0;
                for (int wq = ci; wq < SP / 4; wq += NCH) {
                    const uint32_t w4 = sc_words[wq];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        mine += (((w4 >> (8 * q)) & (CODE_KEY | CODE_CLS_MASK)) == want) && (4 * wq + q) < target;
                }
#pragma unroll
                for (int o = NCH / 2; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
                if (have && rows_on > 0) {
                    const uint32_t ord = wrun[warp * 32 + cls] + (uint32_t)mine;
                    if (ord >= (uint32_t)s_skip[cls]) {
                        const uint32_t cap = (uint32_t)s_cap[cls];
                        const uint32_t pos = ((uint32_t)s_base[cls] + ord % cap) % cap;
                        float4 lo, hi;
                        widen(cell_of(target), lo, hi);
                        const int64_t at = (p.row_off[cls] + pos) * D + d0 + EPL * ci;
                        if (p.bank_bf16) {
                            uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(p.bank_rows) + at);
                            dst[0] = narrow4(lo);
                            if (EPL == 8 && rows_on > 4) dst[1] = narrow4(hi);
                        } else {
                            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.bank_rows) + at);
                            dst[0] = lo;
                            if (EPL == 8 && rows_on > 4) dst[1] = hi;
                        }
                    }
                }
            }
            // advance the per-warp running key ordinals by this sub-tile's keys (all warps agree)
            __syncwarp();
            for (int wb = 0; wb < SP / 4; wb += 32) {                        // warp-uniform trip count (match.any inside)
                const int wq = wb + lane;
                const uint32_t w4 = wq < SP / 4 ? sc_words[wq] : 0u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t cd = (w4 >> (8 * q)) & 0xffu;
                    const bool is_key = cd & CODE_KEY;
                    const uint32_t peers = __match_any_sync(0xffffffffu, is_key ? (cd & CODE_CLS_MASK) : 0xffffu);
                    if (is_key && (__ffs(peers) - 1) == lane) wrun[warp * 32 + (cd & CODE_CLS_MASK)] += __popc(peers);
                    __syncwarp();
                }
            }
        }
        __syncthreads();                                     // tile and code buffer free for the next commit
    }
    flush();
    __syncthreads();
    for (int i = tid; i < C * NCH * AP; i += 256) {
        const int h = i / (C * NCH), c = (i / NCH) % C, k = i % NCH;
        if (EPL * k + 4 * h >= dreal) continue;
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int w = 0; w < NS; ++w) {
            const float4 a = acc[(size_t)h * acc_plane + ((size_t)w * C + c) * NCH + k];
            sum.x += a.x; sum.y += a.y; sum.z += a.z; sum.w += a.w;
        }
        *reinterpret_cast<float4*>(p.partials + ((int64_t)grp * C + c) * D + d0 + EPL * k + 4 * h) = sum;
    }
    proto_finalize_tail(p.partials, p.rows, C, D, const_cast<arco_plan*>(p.plan), p.proto_sums);
}

// ---------------------------------------------------------------------------------------------------
// Small-problem variant (C <= 3, D in {16, 32}; LA 3-D V-Net: C=2, D=16).  With C <= low_rank=3 no class
// can ever rank in [3,20), so no key exists (trap 3) and the pass is a pure masked reduction: thread <->
// 16-byte pixel group, all D rows of the group in flight at once, C*D register accumulators, no shared
// memory staging and no barrier in the streaming loop.
// ---------------------------------------------------------------------------------------------------
template <typename T, int CC, int DD>
__global__ void __launch_bounds__(256) proto_small_kernel(ProtoParams p) {
    constexpr int PER16 = Wide<T>::PER16;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t S = p.S;
    const int64_t gps = S / PER16;                        // pixel groups per image
    const int64_t groups = (int64_t)p.B * gps;
    const T* rep = reinterpret_cast<const T*>(p.rep_t);
    float acc[CC][DD];
#pragma unroll
    for (int c = 0; c < CC; ++c)
#pragma unroll
        for (int d = 0; d < DD; ++d) acc[c][d] = 0.f;
    // This kernel never writes ring rows.  That is exact for the reference's low_rank = 3 >= C (no class can rank in [3,20),
    // trap 3); a C-ABI caller that passed a smaller low_rank gets keys COUNTED by classify/scan_plan that nobody enqueues:
    // flag it (the Python layer raises) instead of silently exposing stale ring rows.
    if (blockIdx.x == 0 && tid < CC && p.plan->n_key[tid] != 0)
        atomicOr(const_cast<uint32_t*>(&p.plan->status), (uint32_t)ARCO_ST_KEYS_DROPPED);

    for (int64_t g = (int64_t)blockIdx.x * 256 + tid; g < groups; g += (int64_t)gridDim.x * 256) {
        const int64_t b = g / gps, s = (g - b * gps) * PER16;
        // code bytes and all D rows are requested together (no dependent round trip); with the iid / 20 % masks
        // of this workload every 32-byte sector holds a low-valid pixel anyway
        uint32_t cw[PER16 / 4];
#pragma unroll
        for (int q = 0; q < PER16 / 4; ++q) cw[q] = *reinterpret_cast<const uint32_t*>(p.codes + b * S + s + 4 * q);
        const T* rowp = rep + b * DD * S + s;
        uint4 raw[DD];
#pragma unroll
        for (int d = 0; d < DD; ++d) { raw[d] = ldg_nc_u4(rowp); rowp += S; }
#pragma unroll
        for (int k = 0; k < PER16; ++k) {
            const uint32_t code = cw[k >> 2] >> (8 * (k & 3));
            const bool lv = code & CODE_LV;
            const int cls = (int)(code & CODE_CLS_MASK);
#pragma unroll
            for (int d = 0; d < DD; ++d) {
                const uint32_t* r = &raw[d].x;
                float x;
                if (PER16 == 4) x = __uint_as_float(r[k]);
                else x = (k & 1) ? __uint_as_float(r[k >> 1] & 0xffff0000u) : __uint_as_float(r[k >> 1] << 16);
#pragma unroll
                for (int c = 0; c < CC; ++c)
                    if (lv && cls == c) acc[c][d] += x;
            }
        }
    }
    // deterministic block reduction: lanes, then warps in fixed order
    __shared__ float s_part[8][CC * DD];
#pragma unroll
    for (int c = 0; c < CC; ++c)
#pragma unroll
        for (int d = 0; d < DD; ++d) {
            float v = acc[c][d];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_part[warp][c * DD + d] = v;
        }
    __syncthreads();
    if (tid < CC * DD) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += s_part[w][tid];
        p.partials[(int64_t)blockIdx.x * CC * DD + tid] = v;
    }
    proto_finalize_tail(p.partials, p.rows, CC, DD, const_cast<arco_plan*>(p.plan), p.proto_sums);
}

__global__ void __launch_bounds__(128) proto_finalize_kernel(const float* __restrict__ partials, int rows, int C, int D,
                                                            const arco_plan* __restrict__ plan,
                                                            double* __restrict__ proto_sums) {
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= C * (D + 1)) return;
    const int c = i / (D + 1), d = i % (D + 1);
    if (d == D) {
        if (lane == 0) proto_sums[i] = (double)plan->lv_count[c];
        return;
    }
    double s = 0.0;
    for (int r = lane; r < rows; r += 32) s += (double)partials[((int64_t)r * C + c) * D + d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) proto_sums[i] = s;
}

// Kernel variant for a problem: the 8-dims-per-lane kernel needs 16-byte vector loads and its 8 KB * C
// stream accumulators to fit beside the tile; otherwise the 4-dims-per-lane kernel (4 KB * C) runs.
bool proto_tc_supported(const arco_dims& d);
int launch_proto_tc(const arco_dims& d, const void* rep_teacher, const arco_bank* bank, const arco_ws_layout& L, char* ws,
                    int rows, double* proto_sums, cudaStream_t st);
bool proto_tc32_supported(const arco_dims& d);
int launch_proto_tc32(const arco_dims& d, const void* rep_teacher, const arco_bank* bank, const arco_ws_layout& L, char* ws,
                      int rows, double* proto_sums, cudaStream_t st);

struct ProtoCfg {
    int kind;         // 0 scalar fallback (proto_enqueue_kernel), 1 pipelined (proto_pipe_kernel), 2 small (proto_small_kernel),
                      // 3 tensor cores (proto_tc_kernel, bf16), 4 tensor cores (proto_tc32_kernel, fp32 as TF32 hi + lo)
    int nch;          // lanes per stream
    int epl;          // feature dims per lane
    int dchunk;       // feature dims per CTA
    size_t smem;
};

static bool proto_vec_ok(const arco_dims& d) {
    const int per16 = d.rep_dtype == ARCO_BF16 ? 8 : 4;
    return d.space % per16 == 0;
}

static ProtoCfg proto_cfg(const arco_dims& d) {
    ProtoCfg c;
    const bool vec = proto_vec_ok(d);
    const char* tc_env = getenv("ARCO_PROTO_TC");
    if (proto_tc_supported(d) && !(tc_env && tc_env[0] == '0')) { c.kind = 3; c.nch = 0; c.epl = 0; c.dchunk = d.feat; c.smem = 0; return c; }
    if (vec && d.classes <= 3 && (d.feat == 16 || d.feat == 32)) { c.kind = 2; c.nch = 0; c.epl = 0; c.dchunk = d.feat; c.smem = 0; return c; }
    const char* tc32_env = getenv("ARCO_PROTO_TC32");
    // one 128-row box per step (D <= 128) leaves the tensor path latency-bound per 32-pixel step: measured 0.111 vs 0.050 ms
    // at D = 64, so the CUDA-core pipeline keeps those shapes (ARCO_PROTO_TC32=2 forces the tensor path, =0 disables it)
    if (proto_tc32_supported(d) && !(tc32_env && tc32_env[0] == '0') && (d.feat > 128 || (tc32_env && tc32_env[0] == '2'))) { c.kind = 4; c.nch = 0; c.epl = 0; c.dchunk = d.feat; c.smem = 0; return c; }
    if (vec) {
        c.kind = 1;
        c.epl = (d.classes <= 8 && d.feat >= 32) ? 8 : 4;
        c.nch = d.feat <= 8 * c.epl ? 8 : 16;
        c.dchunk = c.nch * c.epl;
        c.smem = (size_t)32768 + (size_t)(c.epl / 4) * 8 * (32 / c.nch) * d.classes * c.nch * 16 + 2 * ARCO_TILE + 8 * 32 * 4;
    } else {
        c.kind = 0;
        c.epl = 4;
        c.nch = d.feat <= 32 ? 8 : 16;
        c.dchunk = c.nch * 4;
        const int sp = 2048 / c.nch;
        c.smem = (size_t)32768 + (size_t)8 * (32 / c.nch) * d.classes * c.nch * 16 + sp + 8 * 32 * 4;
    }
    return c;
}

template <typename K>
static int kernel_occupancy(K kernel, size_t smem) {
    int occ = 0;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, smem) != cudaSuccess) {
        cudaGetLastError();
        occ = 0;
    }
    return occ;
}

template <typename T>
static int proto_occ(const ProtoCfg& c) {
    if (c.kind == 1) {
        if (c.epl == 8) return c.nch == 16 ? kernel_occupancy(proto_pipe_kernel<T, 16, 8>, c.smem) : kernel_occupancy(proto_pipe_kernel<T, 8, 8>, c.smem);
        return c.nch == 16 ? kernel_occupancy(proto_pipe_kernel<T, 16, 4>, c.smem) : kernel_occupancy(proto_pipe_kernel<T, 8, 4>, c.smem);
    }
    return c.nch == 16 ? kernel_occupancy(proto_enqueue_kernel<T, 16>, c.smem) : kernel_occupancy(proto_enqueue_kernel<T, 8>, c.smem);
}

// grid geometry shared by the workspace layout and the launch: one resident wave of persistent CTAs
static void proto_grid(const arco_dims& d, int* ndc, int* groups) {
    const ProtoCfg c = proto_cfg(d);
    if (c.kind == 2) { *ndc = 1; *groups = sm_count() * 4; return; }
    if (c.kind == 3 || c.kind == 4) { *ndc = 1; *groups = sm_count(); return; }
    *ndc = (d.feat + c.dchunk - 1) / c.dchunk;
    int occ = d.rep_dtype == ARCO_BF16 ? proto_occ<__nv_bfloat16>(c) : proto_occ<float>(c);
    // without a device (CPU-only build box) assume the shared-memory bound
    if (occ <= 0) occ = (int)((size_t)(227 * 1024) / (c.smem + 1024));
    if (occ > 8) occ = 8;
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
    const ProtoCfg c = proto_cfg(d);
    const int grid = groups * p.NDC;
    if (c.kind == 2) {
#define ARCO_SMALL(CC, DD) proto_small_kernel<T, CC, DD><<<grid, 256, 0, st>>>(p)
        if (d.feat == 16) { if (d.classes == 1) ARCO_SMALL(1, 16); else if (d.classes == 2) ARCO_SMALL(2, 16); else ARCO_SMALL(3, 16); }
        else { if (d.classes == 1) ARCO_SMALL(1, 32); else if (d.classes == 2) ARCO_SMALL(2, 32); else ARCO_SMALL(3, 32); }
#undef ARCO_SMALL
    } else if (c.kind == 1) {
        if (c.epl == 8) {
            if (c.nch == 16) proto_pipe_kernel<T, 16, 8><<<grid, 256, c.smem, st>>>(p);
            else proto_pipe_kernel<T, 8, 8><<<grid, 256, c.smem, st>>>(p);
        } else {
            if (c.nch == 16) proto_pipe_kernel<T, 16, 4><<<grid, 256, c.smem, st>>>(p);
            else proto_pipe_kernel<T, 8, 4><<<grid, 256, c.smem, st>>>(p);
        }
    } else {
        if (c.nch == 16) proto_enqueue_kernel<T, 16><<<grid, 256, c.smem, st>>>(p);
        else proto_enqueue_kernel<T, 8><<<grid, 256, c.smem, st>>>(p);
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
    p.bank_bf16 = bank->row_dtype == ARCO_BF16;
    ARCO_REQUIRE(!p.bank_bf16 || (d.rep_dtype == ARCO_BF16 && d.feat % 8 == 0),
                 "a bf16 bank needs a bf16 rep_teacher and D a multiple of 8");
    p.partials = (float*)(ws + L.partials);
    for (int c = 0; c < ARCO_MAX_CLASSES; ++c) { p.row_off[c] = bank->row_off[c]; p.cap[c] = bank->cap[c] > 0 ? bank->cap[c] : 1; }
    p.S = d.space; p.B = d.n_lab + d.n_unlab; p.C = d.classes; p.D = d.feat;
    p.tpi = L.tiles_per_image; p.NT = L.n_tiles;
    int ndc, groups;
    arco::proto_grid(d, &ndc, &groups);
    p.NDC = ndc;
    static const bool in_kernel_tail = [] { const char* e = getenv("ARCO_PROTO_TAIL"); return !(e && e[0] == '0'); }();
    double* tail_out = in_kernel_tail ? proto_sums : nullptr;
    p.proto_sums = tail_out; p.rows = groups;
    ARCO_REQUIRE(((uintptr_t)rep_teacher & 15) == 0, "rep_teacher must be 16-byte aligned");
    p.vec_ok = arco::proto_vec_ok(d);
    int rc;
    if (arco::proto_cfg(d).kind == 3) rc = arco::launch_proto_tc(d, rep_teacher, bank, L, ws, groups, tail_out, st);
    else if (arco::proto_cfg(d).kind == 4) rc = arco::launch_proto_tc32(d, rep_teacher, bank, L, ws, groups, tail_out, st);
    else rc = d.rep_dtype == ARCO_BF16 ? arco::launch_proto<__nv_bfloat16>(d, p, groups, st)
                                       : arco::launch_proto<float>(d, p, groups, st);
    if (rc != ARCO_OK) return rc;
    if (!in_kernel_tail) {
        const int n = d.classes * (d.feat + 1);
        arco::proto_finalize_kernel<<<(n + 3) / 4, 128, 0, st>>>(p.partials, groups, d.classes, d.feat, p.plan, proto_sums);
        ARCO_LAUNCH_CHECK();
    }
    return ARCO_OK;
}
