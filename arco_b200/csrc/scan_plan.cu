// Ordered-compaction offsets + the device-side "plan" of the step.
//
// The reference compacts with boolean indexing (rep[mask] -> nonzero + gather, loss_helper_3d.py:
// 376-377,382,403) and then makes host decisions from .item()/len() (:413-415, :417, :436-438).
// Here the per-tile counts from arco_classify_count are exclusive-scanned per class (raster order is
// preserved: tile order == pixel order), and the last CTA to finish derives everything the host used
// to decide: valid_classes, valid_seg, which LOOP-2 positions run, and the FIFO bookkeeping of
// dequeue_and_enqueue (loss_helper_3d.py:12-32) for a device-resident ring buffer.
#include "arco_common.cuh"
#include "plan_common.cuh"

namespace arco {

struct ScanParams {
    const uint32_t* cnt_anchor;
    const uint32_t* cnt_key;
    uint32_t* off_anchor;
    uint32_t* off_key;
    arco_plan* plan;
    PlanBank bank;
    const uint32_t* step_ctr;     // bank->counters + ARCO_CTR_STEP, or NULL
    int32_t C, NT, Q;
};

__global__ void __launch_bounds__(1024) scan_plan_kernel(ScanParams p) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    __shared__ bool s_last;
    const int row = blockIdx.x;                 // [0,C): anchor rows, [C,2C): key rows
    const int c = row % p.C;
    const bool is_key = row >= p.C;
    const uint32_t* cnt = (is_key ? p.cnt_key : p.cnt_anchor) + (int64_t)c * p.NT;
    uint32_t* off = (is_key ? p.off_key : p.off_anchor) + (int64_t)c * (p.NT + 1);
    const int tid = threadIdx.x;
    const uint32_t total = scan_row_block<1024>(cnt, off, p.NT, s_warp, &s_carry);
    if (tid == 0) {
        if (is_key) p.plan->n_key[c] = total; else p.plan->n_anchor[c] = total;
        __threadfence();
        const uint32_t ticket = atomicAdd(&p.plan->scan_done, 1u);
        s_last = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last || tid >= 32) return;
    __threadfence();
    // ---- plan: one warp, lane <-> class ----
    volatile arco_plan* vpl = p.plan;
    const bool on = tid < p.C;
    derive_plan_warp(p.plan, p.bank, p.C, p.Q, on ? vpl->lv_count[tid] : 0u, on ? vpl->n_anchor[tid] : 0u,
                     on ? vpl->n_key[tid] : 0u, vpl->status, p.step_ctr ? *(volatile const uint32_t*)p.step_ctr : 0u);
}

// Multi-GPU: after the per-class (feature sum, count) buffer has been all-reduced, the set of valid
// classes must come from the GLOBAL counts so that every rank runs the same LOOP-2 positions
// (SURVEY.md section 8(e)).  Anchors and banks stay rank-local.
__global__ void replan_global_kernel(arco_plan* pl, const double* __restrict__ proto_sums, int C, int D, int Q) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int nv = 0;
    uint32_t changed = 0;
    for (int k = 0; k < C; ++k)
        if (proto_sums[(int64_t)k * (D + 1) + D] > 0.0) { changed |= pl->valid_class[nv] != k; pl->valid_class[nv++] = k; }
    for (int k = nv; k < ARCO_MAX_CLASSES; ++k) { changed |= pl->valid_class[k] != -1; pl->valid_class[k] = -1; }
    pl->n_valid = nv;
    for (int pos = 0; pos < ARCO_MAX_CLASSES; ++pos) {
        int active = 0;
        if (nv > 1 && pos < nv) {
            const int bank_cls = pl->valid_class[pos];
            active = (pl->n_anchor[pos] > 0 && pl->bank_len[bank_cls] > 0) ? 1 : 0;
        }
        changed |= pl->slot_active[pos] != active;
        pl->slot_active[pos] = active;
    }
    pl->inv_scale = nv > 1 ? 1.0f / ((float)Q * (float)nv) : 0.f;
    pl->replanned = changed;                 // a speculative rank-local sampler run must be redone (arco_sample_if_replanned)
}

}  // namespace arco

extern "C" int arco_replan_global(const arco_dims* dims, const double* proto_sums, void* workspace, void* stream) {
    ARCO_REQUIRE(dims && proto_sums && workspace, "arco_replan_global: NULL argument");
    arco_ws_layout L;
    arco::compute_layout(*dims, &L);
    arco::replan_global_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((arco_plan*)((char*)workspace + L.plan), proto_sums,
                                                                 dims->classes, dims->feat, dims->queries);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_scan_plan(const arco_dims* dims, const arco_bank* bank, void* workspace, void* stream) {
    ARCO_REQUIRE(dims && bank && workspace, "arco_scan_plan: NULL argument");
    const arco_dims& d = *dims;
    arco_ws_layout L;
    arco::compute_layout(d, &L);
    char* ws = (char*)workspace;
    arco::ScanParams p;
    p.cnt_anchor = (const uint32_t*)(ws + L.cnt_anchor);
    p.cnt_key = (const uint32_t*)(ws + L.cnt_key);
    p.off_anchor = (uint32_t*)(ws + L.off_anchor);
    p.off_key = (uint32_t*)(ws + L.off_key);
    p.plan = (arco_plan*)(ws + L.plan);
    p.bank.head = bank->head;
    p.bank.len = bank->len;
    p.bank.ptr = bank->queue_ptr;
    for (int c = 0; c < ARCO_MAX_CLASSES; ++c) p.bank.cap[c] = c < d.classes ? bank->cap[c] : 0;
    for (int c = 0; c < d.classes; ++c) ARCO_REQUIRE(bank->cap[c] > 0, "queue_size must be positive");
    p.C = d.classes; p.NT = L.n_tiles; p.Q = d.queries;
    p.step_ctr = bank->counters ? bank->counters + ARCO_CTR_STEP : nullptr;
    arco::scan_plan_kernel<<<2 * d.classes, 1024, 0, (cudaStream_t)stream>>>(p);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}
