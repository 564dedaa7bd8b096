// Shared device/host helpers for libarco_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/arco_b200.h"

#define ARCO_SM_COUNT_FALLBACK 148

// per-pixel code byte
#define CODE_CLS_MASK 0x1Fu
#define CODE_LV 0x20u
#define CODE_ANCHOR 0x40u
#define CODE_KEY 0x80u

namespace arco {

void set_error(const char* fmt, ...);
int sm_count();

#define ARCO_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            arco::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return ARCO_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

#define ARCO_REQUIRE(cond, msg)                                     \
    do {                                                            \
        if (!(cond)) {                                              \
            arco::set_error("%s:%d %s", __FILE__, __LINE__, msg);   \
            return ARCO_ERR_INVALID;                                \
        }                                                           \
    } while (0)

#define ARCO_LAUNCH_CHECK() ARCO_CUDA_CHECK(cudaGetLastError())

__host__ __device__ inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// Rows of per-class prototype partial sums: one row per proto CTA group (see proto_enqueue.cu).
int proto_partial_rows(const arco_dims& d);

inline int compute_layout_uncached(const arco_dims& d, arco_ws_layout* L) {
    const int64_t B = (int64_t)d.n_lab + d.n_unlab;
    const int64_t P = B * d.space;
    const int64_t tpi = (d.space + ARCO_TILE - 1) / ARCO_TILE;
    const int64_t NT = B * tpi;
    const int64_t C = d.classes;
    int64_t off = 0;
    auto take = [&](int64_t bytes) { int64_t o = off; off = align_up(off + bytes, 256); return o; };
    L->plan = take(sizeof(arco_plan));
    L->codes = take(P + 16);
    L->tile_flagged = take(NT * 4);
    L->cnt_anchor = take(C * NT * 4);
    L->cnt_key = take(C * NT * 4);
    L->off_anchor = take(C * (NT + 1) * 4);
    L->off_key = take(C * (NT + 1) * 4);
    L->partial_rows = proto_partial_rows(d);
    L->partials = take((int64_t)L->partial_rows * C * d.feat * 4);
    L->loss_parts = take(C * (int64_t)d.queries * 4);
    // sampler staging: values + flags for the largest call (Q*N draws, <= 1.25x before the drop)
    const int64_t draws = (int64_t)d.queries * (d.negatives > 0 ? d.negatives : 1);
    L->sample_scratch = take(C * 2 * (draws + draws / 4 + 4096) * 4);
    L->total_bytes = off;
    L->n_tiles = (int32_t)NT;
    L->tiles_per_image = (int32_t)tpi;
    L->reserved = 0;
    return 0;
}

// Every stage entry point recomputes the workspace geometry from `dims`; the prototype grid inside it asks the
// runtime for kernel occupancy, so the result is memoised per thread (a training loop repeats one shape).
inline int compute_layout(const arco_dims& d, arco_ws_layout* L) {
    static thread_local arco_dims cached_dims;
    static thread_local arco_ws_layout cached_layout;
    static thread_local int cached_device = -1;
    static thread_local bool valid = false;
    int dev = -1;
    cudaGetDevice(&dev);
    if (valid && dev == cached_device && memcmp(&d, &cached_dims, sizeof(arco_dims)) == 0) {
        *L = cached_layout;
        return 0;
    }
    memset(L, 0, sizeof(*L));
    const int rc = compute_layout_uncached(d, L);
    if (rc == 0) { cached_dims = d; cached_layout = *L; cached_device = dev; valid = true; }
    return rc;
}

template <typename T>
__device__ __forceinline__ T* ws_ptr(void* ws, int64_t off) {
    return reinterpret_cast<T*>(reinterpret_cast<char*>(ws) + off);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al., SC'11) -- the same generator family torch and
// cuRAND use on device; implemented here so sampling needs no library state.
// ---------------------------------------------------------------------------------------------
struct Philox {
    uint32_t k0, k1;
    __host__ __device__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
        uint64_t p = (uint64_t)a * b;
        hi = (uint32_t)(p >> 32);
        lo = (uint32_t)p;
    }
    // counter = (c0, c1, c2, c3) -> 4 x 32 random bits
    __host__ __device__ inline uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        uint32_t ka = k0, kb = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t h0, l0, h1, l1;
            mulhilo(0xD2511F53u, c0, h0, l0);
            mulhilo(0xCD9E8D57u, c2, h1, l1);
            uint32_t n0 = h1 ^ c1 ^ ka, n1 = l1, n2 = h0 ^ c3 ^ kb, n3 = l0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            ka += 0x9E3779B9u;
            kb += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// uniform integer in [0, n) from 32 random bits (multiply-shift; bias < n / 2^32)
__host__ __device__ inline uint32_t bounded(uint32_t bits, uint32_t n) {
    return (uint32_t)(((uint64_t)bits * n) >> 32);
}

// ---------------------------------------------------------------------------------------------
// Keyed pseudo-random permutation of [0, n): alternating Feistel network on ceil(log2 n) bits with
// cycle walking.  Replaces torch.randperm (loss_helper_3d.py:170-173) without a sort.
// ---------------------------------------------------------------------------------------------
struct FeistelPerm {
    uint32_t n, lbits, rbits, lmask, rmask;
    uint32_t key[6];
    __host__ __device__ FeistelPerm() : n(0), lbits(0), rbits(0), lmask(0), rmask(0) {}
    __host__ __device__ FeistelPerm(uint32_t n_, uint4 keys_a, uint4 keys_b) : n(n_) {
        uint32_t bits = 2;
        while (bits < 32 && (1ull << bits) < (uint64_t)n_) ++bits;
        lbits = bits / 2;
        rbits = bits - lbits;
        lmask = (1u << lbits) - 1u;
        rmask = (1u << rbits) - 1u;
        key[0] = keys_a.x; key[1] = keys_a.y; key[2] = keys_a.z; key[3] = keys_a.w;
        key[4] = keys_b.x; key[5] = keys_b.y;
    }
    __host__ __device__ static inline uint32_t mix(uint32_t x, uint32_t k) {
        x ^= k;
        x *= 0x9E3779B1u;
        x ^= x >> 15;
        x *= 0x85EBCA77u;
        x ^= x >> 13;
        return x;
    }
    __host__ __device__ inline uint32_t encrypt(uint32_t x) const {
        uint32_t l = x >> rbits, r = x & rmask;   // l: lbits wide, r: rbits wide
#pragma unroll
        for (int i = 0; i < 6; i += 2) {
            l = (l ^ mix(r, key[i])) & lmask;
            r = (r ^ mix(l, key[i + 1])) & rmask;
        }
        return (l << rbits) | r;
    }
    __host__ __device__ inline uint32_t operator()(uint32_t i) const {
        if (n <= 1) return 0;
        uint32_t x = encrypt(i);
        while (x >= n) x = encrypt(x);
        return x;
    }
};

__device__ __forceinline__ float bf16_bits_to_float(uint32_t lo16) { return __uint_as_float(lo16 << 16); }

__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

}  // namespace arco
