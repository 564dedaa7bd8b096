// The device-side "plan" of a step, derived by ONE WARP (lane k <-> class k / LOOP-2 position k), and the block-wide
// exclusive scan of one row of per-tile counts.  Shared by the stand-alone scan kernel (scan_plan.cu) and the in-kernel
// tail of arco_classify_plan (classify.cu).
//
// Round 1 derived the plan with a single thread walking ~100 dependent global loads/stores (13-17 us on its own for
// nothing but latency); here every per-class quantity is loaded by its own lane at once and the cross-class steps
// (valid-class list, which bank a position is contrasted against) are ballots and shuffles.
#pragma once
#include "arco_common.cuh"

namespace arco {

struct PlanBank {
    int32_t* head;
    int32_t* len;
    int64_t* ptr;
    int32_t cap[ARCO_MAX_CLASSES];
};

// lv / na / nk: this lane's class totals (low-valid pixels, anchor candidates, keys); zero for lanes >= C.
// Reference: dequeue_and_enqueue (loss_helper_3d.py:19-30) on a ring buffer; valid_classes (:413-415); LOOP-2 activity
// (:436-438, trap 1: anchors by POSITION, bank by CLASS ID); mean over Q and / valid_seg (:507-511).
__device__ __forceinline__ void derive_plan_warp(arco_plan* pl, const PlanBank& bank, int C, int Q, uint32_t lv, uint32_t na,
                                                 uint32_t nk, uint32_t status, uint32_t step_ctr) {
    const int k = threadIdx.x & 31;
    const bool on = k < C;
    int new_len = 0;
    {
        const int cap = on ? bank.cap[k] : 0;
        const int len = on ? bank.len[k] : 0, head = on ? bank.head[k] : 0;
        long long ptr = on ? bank.ptr[k] : 0;
        const long long merged = (long long)len + nk;
        const long long overflow = merged > cap ? merged - cap : 0;          // rows dropped from the front
        const int write_base = cap > 0 ? (int)(((long long)head + len) % cap) : 0;
        const int skip = (long long)nk > cap ? (int)((long long)nk - cap) : 0;   // this call's keys that never land
        new_len = (int)(merged > cap ? cap : merged);
        const int new_head = cap > 0 ? (int)(((long long)head + overflow) % cap) : 0;
        if (cap > 0) ptr = (merged >= cap) ? cap : (ptr + nk) % cap;         // reference pointer rule (:24-28)
        pl->lv_count[k] = lv;
        pl->n_anchor[k] = na;
        pl->n_key[k] = nk;
        pl->bank_write_base[k] = write_base;
        pl->bank_skip[k] = skip;
        pl->bank_len[k] = new_len;
        pl->bank_head[k] = new_head;
        pl->queue_ptr[k] = ptr;
        if (on) { bank.len[k] = new_len; bank.head[k] = new_head; bank.ptr[k] = ptr; }
    }
    const uint32_t vmask = __ballot_sync(0xffffffffu, on && lv > 0);         // (:413-415)
    const int nv = __popc(vmask);
    // as a POSITION: lane k is LOOP-2 position k
    const int bank_cls = k < nv ? (int)__fns(vmask, 0, k + 1) : -1;          // valid_classes[k]
    const int len_of_bank = __shfl_sync(0xffffffffu, new_len, bank_cls < 0 ? 0 : bank_cls);
    pl->valid_class[k] = bank_cls;
    pl->slot_active[k] = (nv > 1 && k < nv && na > 0 && len_of_bank > 0) ? 1 : 0;
    if (k == 0) {
        pl->n_valid = nv;
        pl->reserved0 = 0;
        pl->inv_scale = nv > 1 ? 1.0f / ((float)Q * (float)nv) : 0.f;
        pl->status = status;
        pl->scan_done = 0; pl->loss_done = 0; pl->replanned = 0; pl->proto_done = 0; pl->proto_done2 = 0; pl->step_ctr = step_ctr;
    }
}

// Exclusive scan of cnt[0..NT) into off[0..NT], off[NT] = total, by one CTA of NTHR threads (NTHR a multiple of 32, <= 1024).
// A thread owns ITEMS consecutive entries per round: all its loads are issued together (one L2 round trip), scanned in
// registers, one block-wide scan of the thread totals, one round of stores -- instead of a load / 4 barriers / store
// round per NTHR entries.  s_warp: 32 words of shared memory, s_carry: 1 word.  Returns the total (valid in every thread).
template <int NTHR>
__device__ __forceinline__ uint32_t scan_row_block(const uint32_t* cnt, uint32_t* off, int NT, uint32_t* s_warp, uint32_t* s_carry) {
    constexpr int NW = NTHR / 32;
    constexpr int ITEMS = 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) *s_carry = 0;
    __syncthreads();
    for (int base = 0; base < NT; base += NTHR * ITEMS) {
        const int i0 = base + tid * ITEMS;
        uint32_t v[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) v[k] = (i0 + k) < NT ? __ldcg(cnt + i0 + k) : 0u;   // written by other CTAs of this launch: L2
        uint32_t mine = 0;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) mine += v[k];
        uint32_t x = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = lane < NW ? s_warp[lane] : 0u;
            uint32_t ws = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += y;
            }
            if (lane < NW) s_warp[lane] = ws - w;                            // exclusive warp offsets
        }
        __syncthreads();
        uint32_t run = *s_carry + s_warp[warp] + x - mine;                   // exclusive prefix of this thread's first entry
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            if (i0 + k < NT) off[i0 + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (tid == NTHR - 1) *s_carry = run;
        __syncthreads();
    }
    const uint32_t total = *s_carry;
    if (tid == 0) off[NT] = total;
    return total;
}

}  // namespace arco
