// C-ABI plumbing: error strings, workspace geometry, and the inspection helpers used by the
// parity tests (ordered index lists, logical-order bank read-out).
#include <stdarg.h>
#include <string.h>

#include "arco_common.cuh"

namespace arco {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return ARCO_SM_COUNT_FALLBACK;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = ARCO_SM_COUNT_FALLBACK;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

// Materialise the raster-ordered index list of one class: what rep[mask] / nonzero would return in the
// reference (loss_helper_3d.py:376-377,403).  Not on the hot path -- the loss never builds these lists.
__global__ void __launch_bounds__(1024) export_list_kernel(const uint8_t* __restrict__ codes,
                                                            const uint32_t* __restrict__ off, int32_t* out,
                                                            int64_t out_cap, uint32_t* count, int64_t S, int tpi, int NT,
                                                            uint32_t want_mask, uint32_t want, bool scan_local) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_base;
    const int tile = blockIdx.x;
    const int b = tile / tpi;
    const int64_t s0 = (int64_t)(tile % tpi) * ARCO_TILE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t s = s0 + tid;
    const bool hit = s < S && (codes[(int64_t)b * S + s] & want_mask) == want;
    const uint32_t bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_warp[warp] = __popc(bal);
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
        if (lane == 31) {
            // low-valid lists have no precomputed offsets: tiles are serialised through an atomic ticket chain
            s_base = scan_local ? 0u : off[tile];
            if (!scan_local && tile == NT - 1 && count) *count = off[NT];
        }
    }
    __syncthreads();
    if (hit) {
        const int64_t pos = (int64_t)s_base + s_warp[warp] + __popc(bal & ((1u << lane) - 1u));
        if (pos < out_cap) out[pos] = (int32_t)((int64_t)b * S + s);
    }
}

__global__ void bank_read_kernel(const void* __restrict__ rows, int bf16, const int32_t* __restrict__ head,
                                 const int32_t* __restrict__ len, int cls, int cap, int64_t row_off, int D,
                                 float* __restrict__ out) {
    const int n = len[cls], h = head[cls];
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
        int phys = h + r;
        if (phys >= cap) phys -= cap;
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const int64_t at = (row_off + phys) * D + d;
            out[(int64_t)r * D + d] = bf16 ? bf16_bits_to_float(reinterpret_cast<const unsigned short*>(rows)[at])
                                           : reinterpret_cast<const float*>(rows)[at];
        }
    }
}

}  // namespace arco

extern "C" const char* arco_version(void) { return "arco_b200 0.1 (sm_100a)"; }
extern "C" const char* arco_last_error_string(void) { return arco::g_error; }

extern "C" int arco_workspace_layout(const arco_dims* dims, arco_ws_layout* out) {
    ARCO_REQUIRE(dims && out, "arco_workspace_layout: NULL argument");
    ARCO_REQUIRE(dims->classes >= 1 && dims->classes <= ARCO_MAX_CLASSES, "classes must be in [1, 32]");
    ARCO_REQUIRE(dims->feat >= 4 && dims->feat % 4 == 0, "feat (D) must be a positive multiple of 4");
    ARCO_REQUIRE(dims->space > 0 && dims->n_lab >= 0 && dims->n_unlab >= 0 && dims->n_lab + dims->n_unlab > 0,
                 "bad batch/space");
    ARCO_REQUIRE(dims->queries > 0 && dims->negatives >= 0, "bad queries/negatives");
    memset(out, 0, sizeof(*out));
    return arco::compute_layout(*dims, out);
}

extern "C" int arco_export_list(const arco_dims* dims, int32_t kind, int32_t cls, int32_t* out, int64_t out_cap,
                                uint32_t* count_dev, void* workspace, void* stream) {
    ARCO_REQUIRE(dims && out && workspace && cls >= 0 && cls < dims->classes && (kind == 0 || kind == 1),
                 "arco_export_list: bad argument (kind 0 = anchors, 1 = keys)");
    arco_ws_layout L;
    arco::compute_layout(*dims, &L);
    char* ws = (char*)workspace;
    const uint32_t flag = kind == 0 ? CODE_ANCHOR : CODE_KEY;
    const uint32_t* off = (const uint32_t*)(ws + (kind == 0 ? L.off_anchor : L.off_key)) + (int64_t)cls * (L.n_tiles + 1);
    arco::export_list_kernel<<<L.n_tiles, 1024, 0, (cudaStream_t)stream>>>(
        (const uint8_t*)(ws + L.codes), off, out, out_cap, count_dev, dims->space, L.tiles_per_image, L.n_tiles,
        flag | CODE_CLS_MASK, flag | (uint32_t)cls, false);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_bank_read(const arco_bank* bank, int32_t cls, int32_t feat, float* out, void* stream) {
    ARCO_REQUIRE(bank && out && cls >= 0 && cls < ARCO_MAX_CLASSES && feat > 0, "arco_bank_read: bad argument");
    arco::bank_read_kernel<<<256, 128, 0, (cudaStream_t)stream>>>(bank->rows, bank->row_dtype == ARCO_BF16, bank->head, bank->len, cls, bank->cap[cls],
                                                                bank->row_off[cls], feat, out);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}
