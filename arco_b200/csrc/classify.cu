// Sub-system 1: fused label decode + confidence thresholds + teacher-rank test + per-tile counts.
//
// Replaces (reference, 2-D file; the 3-D twin is loss_helper.py +171 lines):
//   label one-hot           train_arco_2d.py:492-498          (index-label mode decodes it in-kernel)
//   low/high valid masks    loss_helper_3d.py:341-342,365-366
//   anchor / hard masks     loss_helper_3d.py:369-374
//   sort over classes       loss_helper_3d.py:352-358  -> rank_i = #{p_j > p_i} + #{j<i, p_j == p_i}
//   class_mask_l/u          loss_helper_3d.py:388-399
//   per-class totals        loss_helper_3d.py:413-415
//
// One CTA per 1024-pixel tile (a tile never straddles an image); 256 threads x 4 consecutive pixels
// so every global access is a 16-byte vector (int64 labels: two per pixel pair).  HBM-bound:
// algorithmic bytes = P * (8C + 4C + 8) read + P written.
#include <stdlib.h>

#include "arco_common.cuh"
#include "plan_common.cuh"

namespace arco {

struct ClassifyParams {
    const int64_t* label_l;
    const int64_t* label_u;
    const float* prob_l;
    const float* prob_u;
    const float* low_mask;
    const float* high_mask;
    // logits mode (arco_classify_plan_logits): prob_l / prob_u hold the TEACHER LOGITS, the masks are derived in the kernel from
    // the student entropy of the unlabelled images and the two percentile thresholds (train_arco_2d.py:356-392)
    const float* entropy_u;     // [n_unlab, S]
    const float* thresholds;    // device float[2]: low, high
    uint8_t* codes;
    uint32_t* tile_flagged;
    uint32_t* cnt_anchor;
    uint32_t* cnt_key;
    arco_plan* plan;
    int64_t S;
    int32_t n_lab, C, tpi, NT;
    int32_t label_kind, low_rank, high_rank;
    float delta_p, delta_n;
    // arco_classify_plan only: scan + plan run in this launch's tail (see classify_tail)
    uint32_t* ctr;             // bank->counters: [0,32) low-valid totals, [32] status, [33] tiles done, [34] rows scanned
    uint32_t* off_anchor;
    uint32_t* off_key;
    PlanBank bank;
    int32_t Q;
};

enum { CTR_STATUS = 32, CTR_TILES = 33, CTR_ROWS = 34 };

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Tail of arco_classify_plan, entered by every CTA after its tiles' counts are written.  The min(grid, 2C) CTAs that finish
// LAST wait until every tile is done (they took their ticket after their own stores, so everything they wait for is
// already running or finished -- at most 2C <= 64 CTAs ever spin), scan one row of tile counts each (row = anchor or key
// counts of one class), and the last scanner derives the plan with one warp and re-zeroes the persistent counters.
__device__ __forceinline__ void classify_tail(const ClassifyParams& p, int n_cta) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    __shared__ int s_role;
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    const int n_scan = min(n_cta, 2 * p.C);
    if (tid == 0) {
        __threadfence();
        const uint32_t t = atomicAdd(&p.ctr[CTR_TILES], 1u);
        s_role = (int)t >= n_cta - n_scan ? (int)t - (n_cta - n_scan) : -1;
    }
    __syncthreads();
    const int role = s_role;
    if (role < 0) return;
    if (tid == 0) {
        while (ld_acquire_gpu(&p.ctr[CTR_TILES]) < (uint32_t)n_cta) __nanosleep(40);
    }
    __syncthreads();
    for (int row = role; row < 2 * p.C; row += n_scan) {
        const int c = row % p.C;
        const bool is_key = row >= p.C;
        const uint32_t* cnt = (is_key ? p.cnt_key : p.cnt_anchor) + (int64_t)c * p.NT;
        uint32_t* off = (is_key ? p.off_key : p.off_anchor) + (int64_t)c * (p.NT + 1);
        const uint32_t total = scan_row_block<256>(cnt, off, p.NT, s_warp, &s_carry);
        if (tid == 0) { if (is_key) p.plan->n_key[c] = total; else p.plan->n_anchor[c] = total; }
        __syncthreads();
    }
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(&p.ctr[CTR_ROWS], 1u) == (uint32_t)n_scan - 1;
    }
    __syncthreads();
    if (!s_last || tid >= 32) return;
    __threadfence();
    volatile arco_plan* vpl = p.plan;
    volatile uint32_t* vc = p.ctr;
    const bool on = tid < p.C;
    const uint32_t lv = on ? vc[tid] : 0u;
    const uint32_t status = vc[CTR_STATUS];
    derive_plan_warp(p.plan, p.bank, p.C, p.Q, lv, on ? vpl->n_anchor[tid] : 0u, on ? vpl->n_key[tid] : 0u, status, vc[ARCO_CTR_STEP]);
    // leave the persistent counters zero for the next step
    __syncwarp();
    vc[tid] = 0u;
    if (tid == 0) { vc[CTR_STATUS] = 0u; vc[CTR_TILES] = 0u; vc[CTR_ROWS] = 0u; }
}


// tile_flagged[tile] is a bit mask: bit g is set when the 32-pixel group g of the tile holds a low-valid or key pixel, i.e.
// something the prototype / enqueue pass has to read.  The tensor-core kernels skip whole 32- / 64-pixel steps whose bits are
// clear (entropy masks are spatially coherent on real images); every consumer treats 0 as "skip the tile".
// With 4 consecutive pixels per thread a warp covers groups 4*warp .. 4*warp+3 (8 lanes each).
__device__ __forceinline__ uint32_t group_bits4(uint32_t ballot, int warp) {
    uint32_t bits = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if ((ballot >> (8 * k)) & 0xffu) bits |= 1u << (4 * warp + k);
    return bits;
}

template <int NV>
struct PixVec;
template <>
struct PixVec<4> {
    __device__ static void load_f(const float* p, float (&v)[4]) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ static void load_l(const int64_t* p, int64_t (&v)[4]) {
        longlong2 a = *reinterpret_cast<const longlong2*>(p);
        longlong2 b = *reinterpret_cast<const longlong2*>(p + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
};
template <>
struct PixVec<1> {
    __device__ static void load_f(const float* p, float (&v)[1]) { v[0] = *p; }
    __device__ static void load_l(const int64_t* p, int64_t (&v)[1]) { v[0] = *p; }
};

// NV = pixels per thread per pass (4: vector path, 1: scalar path for unaligned / odd S)
template <int NV, bool TAIL>
__global__ void __launch_bounds__(256, 4) classify_kernel(ClassifyParams p) {
    __shared__ uint32_t s_anchor[ARCO_MAX_CLASSES], s_key[ARCO_MAX_CLASSES], s_lv[ARCO_MAX_CLASSES];
    __shared__ uint32_t s_flagged, s_status;
    const int tid = threadIdx.x;
    // Persistent CTAs: tile, tile + grid, ...  Before a tile is processed, the NEXT tile's input lines are requested into
    // L2 (prefetch.global.L2, no registers held), so its three dependent load phases (labels -> own-class probability ->
    // rank pass) hit L2 instead of paying the HBM latency three times.
  for (int tile = blockIdx.x; tile < p.NT; tile += gridDim.x) {
    if (tid < ARCO_MAX_CLASSES) { s_anchor[tid] = 0; s_key[tid] = 0; s_lv[tid] = 0; }
    if (tid == 0) { s_flagged = 0; s_status = 0; }
    __syncthreads();
    {
        const int nt = tile + gridDim.x;
        if (nt < p.NT) {
            const int nb = nt / p.tpi;
            const int64_t ns0 = (int64_t)(nt % p.tpi) * ARCO_TILE;
            const int64_t npx = min((int64_t)ARCO_TILE, p.S - ns0);
            const bool nlab = nb < p.n_lab;
            const int nbx = nlab ? nb : nb - p.n_lab;
            const int C_ = p.C;
            const char* lab0 = reinterpret_cast<const char*>(nlab ? p.label_l : p.label_u);
            const char* prob0 = reinterpret_cast<const char*>(nlab ? p.prob_l : p.prob_u);
            const int lab_rows = p.label_kind == ARCO_LABEL_ONEHOT_I64 ? C_ : 1;
            const int lab_lines = (int)((npx * 8 + 127) >> 7), f_lines = (int)((npx * 4 + 127) >> 7);
            const int total = lab_rows * lab_lines + (C_ + 2) * f_lines;
            for (int i = tid; i < total; i += 256) {
                const char* a;
                if (i < lab_rows * lab_lines) {
                    const int r = i / lab_lines, l = i - r * lab_lines;
                    a = lab0 + (((int64_t)nbx * lab_rows + r) * p.S + ns0) * 8 + (int64_t)l * 128;
                } else {
                    const int k = i - lab_rows * lab_lines;
                    const int r = k / f_lines, l = k - r * f_lines;
                    if (r < C_) a = prob0 + (((int64_t)nbx * C_ + r) * p.S + ns0) * 4 + (int64_t)l * 128;
                    else a = reinterpret_cast<const char*>(r == C_ ? p.low_mask : p.high_mask) + ((int64_t)nb * p.S + ns0) * 4 + (int64_t)l * 128;
                }
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            }
        }
    }
    const int b = tile / p.tpi;
    const int64_t s0 = (int64_t)(tile % p.tpi) * ARCO_TILE;
    const bool labelled = b < p.n_lab;
    const int bx = labelled ? b : b - p.n_lab;
    const int C = p.C;
    const int64_t S = p.S;
    const int64_t* lab_base = labelled ? p.label_l : p.label_u;
    const float* prob_base = (labelled ? p.prob_l : p.prob_u) + (int64_t)bx * C * S;
    constexpr int PASSES = 4 / NV;

#pragma unroll 1
    for (int pass = 0; pass < PASSES; ++pass) {
        const int64_t s = (NV == 4) ? s0 + 4 * tid : s0 + tid + 256 * pass;
        const bool in_range = s < S;          // NV==4 requires S % 4 == 0, so the group is all-in or all-out
        int istar[NV];
        float pstar[NV], lm[NV], hm[NV];
        uint32_t status = 0;
#pragma unroll
        for (int v = 0; v < NV; ++v) { istar[v] = -1; pstar[v] = 0.f; lm[v] = 0.f; hm[v] = 0.f; }

        if (in_range) {
            if (p.label_kind == ARCO_LABEL_ONEHOT_I64) {
                const int64_t* lp = lab_base + (int64_t)bx * C * S + s;
#pragma unroll 4
                for (int c = 0; c < C; ++c) {
                    int64_t x[NV];
                    PixVec<NV>::load_l(lp + (int64_t)c * S, x);
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        if (x[v] != 0) {
                            if (istar[v] < 0) istar[v] = c; else status |= ARCO_ST_MULTI_HOT;
                        }
                    }
                }
            } else {
                int64_t x[NV];
                PixVec<NV>::load_l(lab_base + (int64_t)bx * S + s, x);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    int64_t l = x[v] < 0 ? 0 : x[v];             // relu: ignore label -1 -> class 0 (trap 4)
                    if (l >= C) { status |= ARCO_ST_LABEL_RANGE; istar[v] = -1; } else istar[v] = (int)l;
                }
            }
            PixVec<NV>::load_f(p.low_mask + (int64_t)b * S + s, lm);
            PixVec<NV>::load_f(p.high_mask + (int64_t)b * S + s, hm);
#pragma unroll
            for (int v = 0; v < NV; ++v)
                if (istar[v] >= 0) pstar[v] = __ldg(prob_base + (int64_t)istar[v] * S + s + v);
        }

        bool lv[NV], anchor[NV], hard[NV], key[NV];
        bool any_hard = false;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const bool has = istar[v] >= 0;
            lv[v] = has && (lm[v] != 0.f);
            anchor[v] = lv[v] && (pstar[v] > p.delta_p);
            // labelled images can never yield a key: the hard mask needs label_l[:,i]=1 while
            // class_mask_l is multiplied by label_l[:,i]==0 (trap 3, loss_helper_3d.py:372-374,397-399)
            hard[v] = has && !labelled && (hm[v] != 0.f) && (pstar[v] < p.delta_n);
            key[v] = false;
            any_hard |= hard[v];
        }
        // teacher rank of the labelled class, only where a hard pixel needs it (warp-uniform skip)
        if (__any_sync(0xffffffffu, any_hard)) {
            if (any_hard) {
                int rank[NV];
#pragma unroll
                for (int v = 0; v < NV; ++v) rank[v] = 0;
                const float* pp = prob_base + s;
#pragma unroll 4
                for (int c = 0; c < C; ++c) {
                    float q[NV];
                    PixVec<NV>::load_f(pp + (int64_t)c * S, q);
#pragma unroll
                    for (int v = 0; v < NV; ++v)
                        rank[v] += (q[v] > pstar[v]) || (q[v] == pstar[v] && c < istar[v]);
                }
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    key[v] = hard[v] && rank[v] >= p.low_rank && rank[v] < p.high_rank;
            }
        }

        uint32_t packed = 0;
        bool needed = false;                                 // this thread holds a pixel the prototype pass must read
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            uint32_t code = 0;
            if (istar[v] >= 0)
                code = (uint32_t)istar[v] | (lv[v] ? CODE_LV : 0u) | (anchor[v] ? CODE_ANCHOR : 0u) |
                       (key[v] ? CODE_KEY : 0u);
            packed |= code << (8 * v);
            // warp-aggregated counting: one shared atomic per distinct code value in the warp
            const uint32_t flags = code & (CODE_LV | CODE_ANCHOR | CODE_KEY);
            needed |= (flags & (CODE_LV | CODE_KEY)) != 0;
            const uint32_t peers = __match_any_sync(0xffffffffu, code);
            if (flags && (__ffs(peers) - 1) == (tid & 31)) {
                const uint32_t n = __popc(peers), c = code & CODE_CLS_MASK;
                if (flags & CODE_LV) atomicAdd(&s_lv[c], n);
                if (flags & CODE_ANCHOR) atomicAdd(&s_anchor[c], n);
                if (flags & CODE_KEY) atomicAdd(&s_key[c], n);
            }
        }
        {
            const uint32_t bal = __ballot_sync(0xffffffffu, needed);
            if ((tid & 31) == 0 && bal)
                atomicOr(&s_flagged, NV == 4 ? group_bits4(bal, tid >> 5) : 1u << ((tid >> 5) + 8 * pass));
        }
        if (in_range) {
            const int64_t gp = (int64_t)b * S + s;
            if (NV == 4) *reinterpret_cast<uint32_t*>(p.codes + gp) = packed;
            else p.codes[gp] = (uint8_t)packed;
        }
        if (status) atomicOr(&s_status, status);
    }
    __syncthreads();
    if (tid < C) {
        p.cnt_anchor[(int64_t)tid * p.NT + tile] = s_anchor[tid];
        p.cnt_key[(int64_t)tid * p.NT + tile] = s_key[tid];
        if (s_lv[tid]) atomicAdd(TAIL ? &p.ctr[tid] : &p.plan->lv_count[tid], s_lv[tid]);
    }
    if (tid == 0) {
        p.tile_flagged[tile] = s_flagged;
        if (s_status) atomicOr(TAIL ? &p.ctr[CTR_STATUS] : &p.plan->status, s_status);
    }
    __syncthreads();                                         // shared counters are reset for the next tile
  }
    if (TAIL) classify_tail(p, (int)gridDim.x);
}

// ---------------------------------------------------------------------------------------------------------------------
// C <= 8 (ACDC C=4, LA C=2, ...): every input of a 4-pixel group -- C probability vectors, the label vectors, both masks --
// is requested up front (no load depends on another); the own-class probability and the teacher rank are then selected
// from registers.  The generic kernel above needs three dependent round trips per tile (labels -> prob[label] -> rank
// pass), which kept it at 0.2-0.3 of the HBM roofline on the small shapes.  Per-class counts are packed 8 bits per class
// (<= 4 per thread, <= 128 per warp), reduced with redux.sync, and added to the tile counters by one lane per warp.
// ---------------------------------------------------------------------------------------------------------------------
template <int C, int KIND, bool TAIL, bool LOGITS = false>
__global__ void __launch_bounds__(256) classify_small_kernel(ClassifyParams p) {
    constexpr int NLAB = KIND == ARCO_LABEL_ONEHOT_I64 ? C : 1;
    __shared__ uint32_t s_anchor[8], s_key[8], s_lv[8];
    __shared__ uint32_t s_flagged, s_status;
    const int tid = threadIdx.x;
    if (tid < 8) { s_anchor[tid] = 0; s_key[tid] = 0; s_lv[tid] = 0; }
    if (tid == 0) { s_flagged = 0; s_status = 0; }
    __syncthreads();
    const int tile = blockIdx.x;
    const int b = tile / p.tpi;
    const int64_t s0 = (int64_t)(tile % p.tpi) * ARCO_TILE;
    const bool labelled = b < p.n_lab;
    const int bx = labelled ? b : b - p.n_lab;
    const int64_t S = p.S;
    const int64_t s = s0 + 4 * tid;
    const bool in_range = s < S;                             // S % 4 == 0: the group is all-in or all-out

    float4 pr[C];
    longlong2 la[NLAB][2];
    float4 lm4 = make_float4(0.f, 0.f, 0.f, 0.f), hm4 = lm4;
#pragma unroll
    for (int c = 0; c < C; ++c) pr[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NLAB; ++c) la[c][0] = la[c][1] = make_longlong2(0, 0);
    if (in_range) {
        const float* pp = (labelled ? p.prob_l : p.prob_u) + (int64_t)bx * C * S + s;
#pragma unroll
        for (int c = 0; c < C; ++c) pr[c] = __ldg(reinterpret_cast<const float4*>(pp + (int64_t)c * S));
        const int64_t* lp = (labelled ? p.label_l : p.label_u) + (int64_t)bx * NLAB * S + s;
#pragma unroll
        for (int c = 0; c < NLAB; ++c) {
            la[c][0] = __ldg(reinterpret_cast<const longlong2*>(lp + (int64_t)c * S));
            la[c][1] = __ldg(reinterpret_cast<const longlong2*>(lp + (int64_t)c * S + 2));
        }
        if (LOGITS) {
            if (!labelled) lm4 = __ldg(reinterpret_cast<const float4*>(p.entropy_u + (int64_t)bx * S + s));   // the student entropy
        } else {
            lm4 = __ldg(reinterpret_cast<const float4*>(p.low_mask + (int64_t)b * S + s));
            hm4 = __ldg(reinterpret_cast<const float4*>(p.high_mask + (int64_t)b * S + s));
        }
    }
    float lm[4] = {lm4.x, lm4.y, lm4.z, lm4.w}, hm[4] = {hm4.x, hm4.y, hm4.z, hm4.w};
    if (LOGITS) {
        // (1) teacher softmax in registers -- the operations of softmax_rows_kernel (prepare.cu) in the same order, so the
        //     probabilities are bit-identical to the materialised ones (train_arco_2d.py:356-357)
        // (2) the masks: labelled pixels count when their label is valid, unlabelled ones when in addition the student
        //     entropy is <= the low / >= the high percentile threshold (:363-392)
        const float lo = __ldg(p.thresholds), hi = __ldg(p.thresholds + 1);
        float x[C][4];
#pragma unroll
        for (int c = 0; c < C; ++c) { x[c][0] = pr[c].x; x[c][1] = pr[c].y; x[c][2] = pr[c].z; x[c][3] = pr[c].w; }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            float m = -INFINITY;
#pragma unroll
            for (int c = 0; c < C; ++c) m = fmaxf(m, x[c][v]);
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) { x[c][v] = expf(x[c][v] - m); sum += x[c][v]; }
#pragma unroll
            for (int c = 0; c < C; ++c) x[c][v] = x[c][v] / sum;
            const long long l = v == 0 ? la[0][0].x : v == 1 ? la[0][0].y : v == 2 ? la[0][1].x : la[0][1].y;
            const bool valid = in_range && l >= 0;
            const float e = lm[v];
            lm[v] = (valid && (labelled || e <= lo)) ? 1.f : 0.f;
            hm[v] = (valid && (labelled || e >= hi)) ? 1.f : 0.f;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) pr[c] = make_float4(x[c][0], x[c][1], x[c][2], x[c][3]);
    }
    uint32_t packed = 0, status = 0;
    unsigned long long n_lv = 0ull, n_an = 0ull, n_key = 0ull;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        int istar = -1;
        if (KIND == ARCO_LABEL_ONEHOT_I64) {
#pragma unroll
            for (int c = 0; c < NLAB; ++c) {
                const long long x = v == 0 ? la[c][0].x : v == 1 ? la[c][0].y : v == 2 ? la[c][1].x : la[c][1].y;
                if (x != 0) { if (istar < 0) istar = c; else status |= ARCO_ST_MULTI_HOT; }
            }
        } else if (in_range) {
            long long l = v == 0 ? la[0][0].x : v == 1 ? la[0][0].y : v == 2 ? la[0][1].x : la[0][1].y;
            l = l < 0 ? 0 : l;                               // relu: ignore label -1 -> class 0 (trap 4)
            if (l >= C) status |= ARCO_ST_LABEL_RANGE; else istar = (int)l;
        }
        float pstar = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float q = v == 0 ? pr[c].x : v == 1 ? pr[c].y : v == 2 ? pr[c].z : pr[c].w;
            pstar = (c == istar) ? q : pstar;
        }
        const bool has = istar >= 0;
        const bool lv = has && (lm[v] != 0.f);
        const bool anchor = lv && (pstar > p.delta_p);
        // labelled images can never yield a key (trap 3, loss_helper_3d.py:372-374,397-399)
        const bool hard = has && !labelled && (hm[v] != 0.f) && (pstar < p.delta_n);
        int rank = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float q = v == 0 ? pr[c].x : v == 1 ? pr[c].y : v == 2 ? pr[c].z : pr[c].w;
            rank += (q > pstar) || (q == pstar && c < istar);
        }
        const bool key = hard && rank >= p.low_rank && rank < p.high_rank;
        uint32_t code = 0;
        if (has) code = (uint32_t)istar | (lv ? CODE_LV : 0u) | (anchor ? CODE_ANCHOR : 0u) | (key ? CODE_KEY : 0u);
        packed |= code << (8 * v);
        const unsigned long long one = has ? 1ull << (8 * istar) : 0ull;
        n_lv += lv ? one : 0ull;
        n_an += anchor ? one : 0ull;
        n_key += key ? one : 0ull;
    }
    if (in_range) *reinterpret_cast<uint32_t*>(p.codes + (int64_t)b * S + s) = packed;
    {
        uint32_t w[6];
        w[0] = __reduce_add_sync(0xffffffffu, (uint32_t)n_lv);
        w[1] = __reduce_add_sync(0xffffffffu, (uint32_t)n_an);
        w[2] = __reduce_add_sync(0xffffffffu, (uint32_t)n_key);
        w[3] = C > 4 ? __reduce_add_sync(0xffffffffu, (uint32_t)(n_lv >> 32)) : 0u;
        w[4] = C > 4 ? __reduce_add_sync(0xffffffffu, (uint32_t)(n_an >> 32)) : 0u;
        w[5] = C > 4 ? __reduce_add_sync(0xffffffffu, (uint32_t)(n_key >> 32)) : 0u;
        const uint32_t st = __reduce_or_sync(0xffffffffu, status);
        const uint32_t bal = __ballot_sync(0xffffffffu, (n_lv | n_key) != 0ull);
        if ((tid & 31) == 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int sh = 8 * (c & 3), hi = c >> 2;
                const uint32_t a = (w[3 * hi + 0] >> sh) & 0xffu, bq = (w[3 * hi + 1] >> sh) & 0xffu, k = (w[3 * hi + 2] >> sh) & 0xffu;
                if (a) atomicAdd(&s_lv[c], a);
                if (bq) atomicAdd(&s_anchor[c], bq);
                if (k) atomicAdd(&s_key[c], k);
            }
            if (bal) atomicOr(&s_flagged, group_bits4(bal, tid >> 5));
            if (st) atomicOr(&s_status, st);
        }
    }
    __syncthreads();
    if (tid < C) {
        p.cnt_anchor[(int64_t)tid * p.NT + tile] = s_anchor[tid];
        p.cnt_key[(int64_t)tid * p.NT + tile] = s_key[tid];
        if (s_lv[tid]) atomicAdd(TAIL ? &p.ctr[tid] : &p.plan->lv_count[tid], s_lv[tid]);
    }
    if (tid == 0) {
        p.tile_flagged[tile] = s_flagged;
        if (s_status) atomicOr(TAIL ? &p.ctr[CTR_STATUS] : &p.plan->status, s_status);
    }
    if (TAIL) {
        __syncthreads();
        classify_tail(p, (int)gridDim.x);
    }
}

template <int KIND, bool TAIL, bool LOGITS = false>
static bool launch_classify_small(const ClassifyParams& p, int grid, cudaStream_t st) {
    switch (p.C) {
#define ARCO_CS(CC) case CC: classify_small_kernel<CC, KIND, TAIL, LOGITS><<<grid, 256, 0, st>>>(p); return true;
        ARCO_CS(2) ARCO_CS(3) ARCO_CS(4) ARCO_CS(5) ARCO_CS(6) ARCO_CS(7) ARCO_CS(8)
#undef ARCO_CS
        default: return false;
    }
}

// (a1) stand-alone drop-in for the trainers' label_onehot
__global__ void label_onehot_kernel(const int64_t* __restrict__ labels, float* __restrict__ out, int64_t batch,
                                    int classes, int64_t space) {
    const int64_t total = batch * classes * space;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i % space;
        const int64_t c = (i / space) % classes;
        const int64_t b = i / (space * classes);
        int64_t l = labels[b * space + s];
        l = l < 0 ? 0 : l;
        out[i] = (l == c) ? 1.f : 0.f;
    }
}

}  // namespace arco

extern "C" int arco_label_onehot(const int64_t* labels, float* out, int64_t batch, int32_t classes, int64_t space,
                                 void* stream) {
    ARCO_REQUIRE(labels && out && batch > 0 && classes > 0 && space > 0, "arco_label_onehot: bad argument");
    const int64_t total = batch * classes * space;
    int blocks = (int)((total + 255) / 256);
    const int cap = arco::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    arco::label_onehot_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(labels, out, batch, classes, space);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

static int classify_launch(const arco_dims* dims, const int64_t* label_l, const int64_t* label_u, const float* prob_l,
                           const float* prob_u, const float* low_mask, const float* high_mask, float delta_p, float delta_n,
                           int32_t low_rank, int32_t high_rank, const arco_bank* bank, void* workspace, void* stream,
                           const float* entropy_u = nullptr, const float* thresholds = nullptr) {
    ARCO_REQUIRE(dims && workspace, "arco_classify_count: NULL dims/workspace");
    const bool logits = thresholds != nullptr;
    if (logits) { low_mask = thresholds; high_mask = thresholds; }           // (only checked for NULL / alignment below)
    const arco_dims& d = *dims;
    ARCO_REQUIRE(d.classes >= 1 && d.classes <= ARCO_MAX_CLASSES, "classes must be in [1, 32]");
    ARCO_REQUIRE(d.n_lab >= 0 && d.n_unlab >= 0 && d.n_lab + d.n_unlab > 0 && d.space > 0, "bad batch/space");
    ARCO_REQUIRE((d.n_lab == 0 || (label_l && prob_l)) && (d.n_unlab == 0 || (label_u && prob_u)) && low_mask &&
                     high_mask, "NULL input tensor");
    ARCO_REQUIRE(((int64_t)d.n_lab + d.n_unlab) * d.space < (int64_t)0x7fffffff, "more than 2^31 pixels");
    arco_ws_layout L;
    arco::compute_layout(d, &L);
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    const bool tail = bank != nullptr;
    if (!tail) ARCO_CUDA_CHECK(cudaMemsetAsync(ws + L.plan, 0, sizeof(arco_plan), st));

    arco::ClassifyParams p;
    p.label_l = label_l; p.label_u = label_u; p.prob_l = prob_l; p.prob_u = prob_u;
    p.low_mask = low_mask; p.high_mask = high_mask; p.entropy_u = entropy_u; p.thresholds = thresholds;
    p.codes = (uint8_t*)(ws + L.codes);
    p.tile_flagged = (uint32_t*)(ws + L.tile_flagged);
    p.cnt_anchor = (uint32_t*)(ws + L.cnt_anchor);
    p.cnt_key = (uint32_t*)(ws + L.cnt_key);
    p.plan = (arco_plan*)(ws + L.plan);
    p.S = d.space; p.n_lab = d.n_lab; p.C = d.classes; p.tpi = L.tiles_per_image; p.NT = L.n_tiles;
    p.label_kind = d.label_kind; p.low_rank = low_rank; p.high_rank = high_rank;
    p.delta_p = delta_p; p.delta_n = delta_n;
    p.ctr = nullptr; p.off_anchor = (uint32_t*)(ws + L.off_anchor); p.off_key = (uint32_t*)(ws + L.off_key); p.Q = d.queries;
    p.bank.head = nullptr; p.bank.len = nullptr; p.bank.ptr = nullptr;
    for (int c = 0; c < ARCO_MAX_CLASSES; ++c) p.bank.cap[c] = 0;
    if (tail) {
        ARCO_REQUIRE(bank->counters != nullptr, "arco_classify_plan: bank->counters is NULL (needs ARCO_COUNTER_WORDS zeroed uint32)");
        p.ctr = bank->counters;
        p.bank.head = bank->head; p.bank.len = bank->len; p.bank.ptr = bank->queue_ptr;
        for (int c = 0; c < d.classes; ++c) {
            ARCO_REQUIRE(bank->cap[c] > 0, "queue_size must be positive");
            p.bank.cap[c] = bank->cap[c];
        }
    }

    auto aligned16 = [](const void* q) { return q == nullptr || ((uintptr_t)q & 15) == 0; };
    const bool vec = (d.space % 4 == 0) && aligned16(label_l) && aligned16(label_u) && aligned16(prob_l) &&
                     aligned16(prob_u) && aligned16(low_mask) && aligned16(high_mask);
    // ARCO_CLASSIFY_CTAS CTAs per SM; above ~28 that is one CTA per tile for every shape here (the default).  Measured on
    // B200: persistent CTAs (4 per SM) with an L2 prefetch of their next tile are no faster at C <= 4 and slower at C = 19
    // (0.285 vs 0.199 ms: 140 MB of prefetched lines thrash the 126 MB L2), so tiles are left to the block scheduler.
    static const int per_sm = [] { const char* e = getenv("ARCO_CLASSIFY_CTAS"); return e && atoi(e) > 0 ? atoi(e) : 32; }();
    int grid = arco::sm_count() * per_sm;
    if (grid > L.n_tiles) grid = L.n_tiles;
    static const bool small_ok = [] { const char* e = getenv("ARCO_CLASSIFY_SMALL"); return !(e && e[0] == '0'); }();
    if (logits) {
        ARCO_REQUIRE(tail && d.label_kind == ARCO_LABEL_INDEX_I64 && d.classes >= 2 && d.classes <= 8 && d.space % 4 == 0 &&
                         aligned16(label_l) && aligned16(label_u) && aligned16(prob_l) && aligned16(prob_u) && aligned16(entropy_u) &&
                         (d.n_unlab == 0 || entropy_u),
                     "arco_classify_plan_logits: needs integer label maps, 2 <= C <= 8, S % 4 == 0 and 16-byte aligned tensors");
        const bool done = arco::launch_classify_small<ARCO_LABEL_INDEX_I64, true, true>(p, L.n_tiles, st);
        ARCO_REQUIRE(done, "arco_classify_plan_logits: unsupported class count");
        ARCO_LAUNCH_CHECK();
        return ARCO_OK;
    }
    if (vec && small_ok && d.classes >= 2 && d.classes <= 8) {
        // one tile per CTA, every load independent (see classify_small_kernel)
        const int g = L.n_tiles;
        bool done;
        if (d.label_kind == ARCO_LABEL_ONEHOT_I64)
            done = tail ? arco::launch_classify_small<ARCO_LABEL_ONEHOT_I64, true>(p, g, st) : arco::launch_classify_small<ARCO_LABEL_ONEHOT_I64, false>(p, g, st);
        else
            done = tail ? arco::launch_classify_small<ARCO_LABEL_INDEX_I64, true>(p, g, st) : arco::launch_classify_small<ARCO_LABEL_INDEX_I64, false>(p, g, st);
        if (done) { ARCO_LAUNCH_CHECK(); return ARCO_OK; }
    }
    if (tail) {
        if (vec) arco::classify_kernel<4, true><<<grid, 256, 0, st>>>(p);
        else arco::classify_kernel<1, true><<<grid, 256, 0, st>>>(p);
    } else {
        if (vec) arco::classify_kernel<4, false><<<grid, 256, 0, st>>>(p);
        else arco::classify_kernel<1, false><<<grid, 256, 0, st>>>(p);
    }
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_classify_count(const arco_dims* dims, const int64_t* label_l, const int64_t* label_u,
                                   const float* prob_l, const float* prob_u, const float* low_mask,
                                   const float* high_mask, float delta_p, float delta_n, int32_t low_rank,
                                   int32_t high_rank, void* workspace, void* stream) {
    return classify_launch(dims, label_l, label_u, prob_l, prob_u, low_mask, high_mask, delta_p, delta_n, low_rank, high_rank,
                           nullptr, workspace, stream);
}

extern "C" int arco_classify_plan(const arco_dims* dims, const int64_t* label_l, const int64_t* label_u,
                                  const float* prob_l, const float* prob_u, const float* low_mask,
                                  const float* high_mask, float delta_p, float delta_n, int32_t low_rank,
                                  int32_t high_rank, const arco_bank* bank, void* workspace, void* stream) {
    ARCO_REQUIRE(bank != nullptr, "arco_classify_plan: NULL bank");
    return classify_launch(dims, label_l, label_u, prob_l, prob_u, low_mask, high_mask, delta_p, delta_n, low_rank, high_rank,
                           bank, workspace, stream);
}

extern "C" int arco_classify_plan_logits(const arco_dims* dims, const int64_t* label_l, const int64_t* label_u,
                                         const float* logits_l_teacher, const float* logits_u_teacher, const float* entropy_u,
                                         const float* thresholds, float delta_p, float delta_n, int32_t low_rank,
                                         int32_t high_rank, const arco_bank* bank, void* workspace, void* stream) {
    ARCO_REQUIRE(bank != nullptr && thresholds != nullptr, "arco_classify_plan_logits: NULL bank / thresholds");
    return classify_launch(dims, label_l, label_u, logits_l_teacher, logits_u_teacher, nullptr, nullptr, delta_p, delta_n, low_rank,
                           high_rank, bank, workspace, stream, entropy_u, thresholds);
}
