// Sub-systems 3c + 4: anchor rank-select/gather, memory-bank negative gather, cosine similarity,
// temperature-scaled InfoNCE and its gradient w.r.t. the anchor rows, in ONE kernel.
//
// Replaces loss_helper_3d.py:435-511 (and autograd's backward through it):
//   anchor_feat   = seg_feat_low_entropy_list[i][idx_a]            (:455-457)  -> rank-select on the
//                   per-tile anchor offsets + in-tile select on the code bytes (no compacted list exists)
//   negative_feat = memobank[valid_classes[i]][0][idx_n]           (:466-479)  -> ring-buffer rows fetched
//                   with cp.async.bulk (TMA 1-D bulk copies) into warp-private stages (one per warp; occupancy hides the wait)
//   all_feat / cosine_similarity / cross_entropy                   (:480-509)  -> online softmax; the
//                   [Q,1+N,D] tensor is never materialised
//   d loss / d anchor = (1/|a|) (Gw - (a_hat . Gw) a_hat),  Gw = sum_k w_k k_hat,  w_k = (p_k - [k=0])/temp
//
// One CTA (4 warps) per (LOOP-2 position, query); each warp owns a quarter of that query's negatives.
// FFMA form: 2 flop per 4 bytes gathered, i.e. bandwidth (L2 / HBM) bound -- see DESIGN.md for why the
// dense tensor-core reformulation is not used at the reference's Q, N.
#include "arco_common.cuh"

namespace arco {

struct InfoParams {
    const void* rep;
    const void* bank_rows;       // fp32 or bf16 rows (arco_bank.row_dtype)
    const double* proto_sums;
    const int32_t* idx_a;
    const int32_t* idx_n;
    const uint8_t* codes;
    const uint32_t* off_anchor;
    arco_plan* plan;
    float* loss;
    float* g_anchor;
    int32_t* anchor_pix;
    float* logits;
    float* loss_parts;
    // optional EMA prototypes (a11, loss_helper_3d.py:488-497)
    const float* momentum;       // [C][Q][D] or NULL
    const int32_t* momentum_on;  // device flag: momentum tensor has a non-zero entry (:489)
    float* proto_out;            // [C][Q][D] positive_feat written per (bank class, query) (:497), or NULL
    void* host_mirror;           // arco_bank.host_mirror (ring base) / host_queue_ptr
    int64_t* host_queue_ptr;
    uint32_t* step_ctr;          // bank->counters + ARCO_CTR_STEP or NULL
    float ema_decay, ema_keep;   // ema_keep = float32(1 - ema_decay) formed in DOUBLE by the caller, as the reference's Python scalar is (:491-495)
    int64_t row_off[ARCO_MAX_CLASSES];
    int32_t cap[ARCO_MAX_CLASSES];
    int64_t S;
    int32_t C, D, Q, N, tpi, NT;
    int32_t KC, RS16;        // keys per staged chunk, padded row stride in 16-byte chunks
    int32_t rep_dtype;
    float temp;
    // "rows mode" (arco_infonce_rows): the anchors were selected and produced upstream -- row j*Q+q of a dense
    // fp32 [C*Q][D] array and its pixel id -- instead of being rank-selected and gathered from a [B,D,S] tensor here
    const float* anchor_rows;
    const int32_t* anchor_pix_in;
    // Multi-GPU (arco_infonce_sharded): the exchange step of the path runs INSIDE this launch.  Block 0 is an extra CTA that
    // trades the C x (D+1) fp64 class sums with the peers over NVLink and re-derives the valid-class list while the query
    // CTAs (blocks 1..nq) run their negatives pass on the rank-local plan; they meet at plan->reserved0 before the merge.
    const unsigned long long* xchg_peers;   // NULL: no exchange in this launch
    double* xchg_out;                       // == proto_sums (written by block 0)
    unsigned long long xchg_seq;
    int64_t xchg_slot;                      // doubles per slot of the peer-mapped buffer
    int32_t xchg_rank, xchg_world;
    int32_t gate_replanned;                 // 1: this launch is the redo; it does nothing unless plan->replanned
    int32_t nq;                             // C * Q query CTAs
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

constexpr float kEps = 1e-8f;   // torch.cosine_similarity eps (ATen default), applied per norm

// 16-byte chunk of a bank row -> CHD floats (4 fp32, or 8 bf16 widened)
template <bool BF16BANK>
__device__ __forceinline__ void unpack_chunk(const uint4& u, float* v) {
    if (BF16BANK) {
        v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
        v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
        v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
        v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
    } else {
        v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
    }
}

struct AnchorInfo { int pix; float na, cos0; };

// Rank-select of one anchor: the idx-th anchor candidate (raster order) of LOOP-2 position j -> its pixel id in *s_pix.
// Called by ONE full warp.  No compacted list exists: 32-ary search over the per-tile exclusive counts, then an in-tile
// select on the code bytes (replaces seg_feat_low_entropy_list[i][idx], loss_helper_3d.py:455-457).
__device__ __forceinline__ void anchor_select_warp(const InfoParams& p, int j, int q, int* s_pix) {
    const int lane = threadIdx.x & 31;
    const arco_plan* pl = p.plan;
    const uint32_t n_anchor = pl->n_anchor[j];
    uint32_t idx = (uint32_t)p.idx_a[(int64_t)j * p.Q + q];
    if (idx >= n_anchor) {                                   // injected / foreign index out of range: flag it (the host raises), then stay in bounds
        if (lane == 0) atomicOr(&p.plan->status, (uint32_t)ARCO_ST_INDEX_RANGE);
        idx = n_anchor - 1;
    }
    const uint32_t* off = p.off_anchor + (int64_t)j * (p.NT + 1);
    // largest tile lo with off[lo] <= idx: 32 probes per round instead of a dependent load per halving
    int lo = 0, hi = p.NT;
    while (hi - lo > 1) {
        const int step = (hi - lo + 31) >> 5;
        const int pos = lo + lane * step;
        const bool le = pos < hi && off[pos] <= idx;
        const int k = __popc(__ballot_sync(0xffffffffu, le));         // >= 1: off[lo] <= idx holds throughout
        lo += (k - 1) * step;
        hi = min(hi, lo + step);
    }
    const uint32_t r = idx - off[lo];
    const int b = lo / p.tpi;
    const int64_t s0 = (int64_t)(lo % p.tpi) * ARCO_TILE;
    const int64_t n = min((int64_t)ARCO_TILE, p.S - s0);
    const uint8_t* cp = p.codes + (int64_t)b * p.S + s0;
    const uint32_t want = CODE_ANCHOR | (uint32_t)j;
    uint32_t mask = 0;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
        const int i = lane * 32 + k;
        const uint32_t cd = i < n ? cp[i] : 0u;
        mask |= (uint32_t)((cd & (CODE_ANCHOR | CODE_CLS_MASK)) == want) << k;
    }
    const uint32_t cnt = __popc(mask);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    const uint32_t excl = incl - cnt;
    if (r >= excl && r < incl) {
        uint32_t m = mask;
        for (uint32_t i = 0; i < r - excl; ++i) m &= m - 1;
        *s_pix = (int)((int64_t)b * p.S + s0 + lane * 32 + (__ffs(m) - 1));
    }
}

// Anchor rank-select and anchor row -> a_hat (unit vector in shared memory), |a|.  Called by all 128 threads of the CTA.
// The prototype half of the prologue (info_proto) runs AFTER the negatives pass: nothing in that pass needs the prototype,
// and on a batch shard the global class sums arrive while it runs.
template <bool ROWS>
__device__ __forceinline__ AnchorInfo info_anchor(const InfoParams& p, int j, int q, float* a_hat, float (*s_red)[2], int* s_pix) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = p.D;
    // ---- anchor rank-select: idx-th anchor candidate of class j (... anchors by POSITION j) ----
    if (!ROWS && warp == 0) anchor_select_warp(p, j, q, s_pix);
    __syncthreads();
    const int pix = ROWS ? p.anchor_pix_in[(int64_t)j * p.Q + q] : *s_pix;
    const int ab = (int)(pix / p.S);
    const int64_t as = pix - (int64_t)ab * p.S;
    // ---- anchor row (D strided loads, one 32-B sector each) ----
    float n2a = 0.f;
    for (int d = tid; d < D; d += 128) {
        float v;
        if (ROWS)
            v = p.anchor_rows[((int64_t)j * p.Q + q) * D + d];
        else if (p.rep_dtype == ARCO_BF16)
            v = bf16_bits_to_float(reinterpret_cast<const unsigned short*>(p.rep)[((int64_t)ab * D + d) * p.S + as]);
        else
            v = reinterpret_cast<const float*>(p.rep)[((int64_t)ab * D + d) * p.S + as];
        a_hat[d] = v;
        n2a += v * v;
    }
    n2a = warp_sum(n2a);
    if (lane == 0) s_red[warp][0] = n2a;
    __syncthreads();
    const float na = sqrtf(s_red[0][0] + s_red[1][0] + s_red[2][0] + s_red[3][0]);
    const float inv_na = 1.f / fmaxf(na, kEps);
    for (int d = tid; d < D; d += 128) a_hat[d] *= inv_na;
    __syncthreads();
    AnchorInfo out;
    out.pix = pix;
    out.na = na;
    out.cos0 = 0.f;
    return out;
}

// Prototype row (class mean, optionally the EMA blend) -> k0hat (unit vector) and cos(a, k0).  All 128 threads.
__device__ __forceinline__ float info_proto(const InfoParams& p, int j, int q, int bank_cls, const float* a_hat, float* k0hat) {
    __shared__ float s_k[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = p.D;
    const volatile double* ps = p.proto_sums;                 // on a batch shard block 0 of THIS launch wrote them
    const double cntj = ps[(int64_t)j * (D + 1) + D];
    float n2k = 0.f;
    for (int d = tid; d < D; d += 128) {
        float k = (float)(ps[(int64_t)j * (D + 1) + d] / cntj);   // class mean (:380-384)
        if (p.momentum) {
            // positive = (1-a)*proto + a*momentum_prototype[valid_classes[i]][q]  (:490-495); prototype[...] = positive (:497)
            const int64_t mo = ((int64_t)bank_cls * p.Q + q) * D + d;
            if (*p.momentum_on) k = p.ema_keep * k + p.ema_decay * p.momentum[mo];
            if (p.proto_out) p.proto_out[mo] = k;
        }
        k0hat[d] = k;
        n2k += k * k;
    }
    n2k = warp_sum(n2k);
    __syncthreads();                       // s_k may still be read by a previous use
    if (lane == 0) s_k[warp] = n2k;
    __syncthreads();
    const float nk0 = sqrtf(s_k[0] + s_k[1] + s_k[2] + s_k[3]);
    const float inv_nk0 = 1.f / fmaxf(nk0, kEps);
    float c0 = 0.f;
    for (int d = tid; d < D; d += 128) {
        const float k = k0hat[d] * inv_nk0;
        k0hat[d] = k;
        c0 += a_hat[d] * k;
    }
    c0 = warp_sum(c0);
    __syncthreads();                       // everyone is done reading s_k
    if (lane == 0) s_k[warp] = c0;
    __syncthreads();
    return s_k[0] + s_k[1] + s_k[2] + s_k[3];
}

// ---- the exchange step inside the launch (batch shards) -------------------------------------------------------------
__device__ __forceinline__ void x_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long x_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double x_ld_relaxed_sys_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long x_globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Block 0 of a sharded launch (128 threads).  Same protocol as proto_allreduce_p2p_kernel (allreduce.cu): raise this step's
// sequence number in every peer's flag row, wait for theirs, add the W peer slots in rank order (bit-identical sums on every
// rank) -- then re-derive valid_class / slot_active / inv_scale from the GLOBAL counts (replan_global_kernel, scan_plan.cu)
// and publish plan->reserved0 = 1 for the query CTAs of this launch.
struct XchgArgs {                                             // passed BY VALUE to the out-of-line exchange block
    const unsigned long long* xchg_peers;
    double* xchg_out;
    unsigned long long xchg_seq;
    int64_t xchg_slot;
    arco_plan* plan;
    int32_t xchg_rank, xchg_world, C, D, Q;
};
__device__ __noinline__ void exchange_block(const XchgArgs p) {
    const int tid = threadIdx.x;
    const int world = p.xchg_world, rank = p.xchg_rank;
    const int n = p.C * (p.D + 1);
    const int64_t flag_off = 2 * p.xchg_slot;
    arco_plan* pl = p.plan;
    // The step's sequence number: a launch parameter, or (replayed step, ARCO_XCHG_SEQ_FROM_DEVICE) the buffer's step word + 1.
    // Either way it is stored back, so direct and replayed steps can follow each other on the same buffer.
    unsigned long long seq = p.xchg_seq & ARCO_XCHG_SEQ_MASK;
    if (p.xchg_seq & ARCO_XCHG_STEP_WORD) {
        __shared__ unsigned long long s_seq;
        if (tid == 0) {
            unsigned long long* word = reinterpret_cast<unsigned long long*>(p.xchg_peers[rank]) + flag_off + 64;
            if (p.xchg_seq & ARCO_XCHG_SEQ_FROM_DEVICE) {
                const unsigned long long d = *word + 1ull;
                if ((d ^ seq) & 1ull) {                       // the prototype pass of this step wrote the other slot
                    atomicOr(&pl->status, (uint32_t)ARCO_ST_EXCHANGE_DESYNC);
                    __threadfence_system();
                    __trap();
                }
                seq = d;
            }
            *word = seq;
            s_seq = seq;
        }
        __syncthreads();
        seq = s_seq;
    }
    const int64_t slot = (int64_t)(seq & 1ull) * p.xchg_slot;
    __threadfence_system();                                   // the prototype kernel's sums (previous launch) before the flag
    if (tid < world && tid != rank)
        x_st_release_sys(reinterpret_cast<unsigned long long*>(p.xchg_peers[tid]) + flag_off + rank, seq);
    if (tid < world && tid != rank) {
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(p.xchg_peers[rank]) + flag_off + tid;
        const unsigned long long t0 = x_globaltimer_ns();     // a peer that never arrives must fail loudly: 10 s of wall time
        unsigned int polls = 0;
        while (x_ld_acquire_sys(mine) < seq) {
            if ((++polls & 1023u) == 0 && x_globaltimer_ns() - t0 > 10000000000ull) {
                atomicOr(&pl->status, (uint32_t)ARCO_ST_EXCHANGE_TIMEOUT);
                __threadfence_system();
                __trap();
            }
        }
    }
    __syncthreads();
    // 2 elements x up to 8 peers of independent NVLink loads in flight per thread, added in rank order afterwards.  Kept
    // small on purpose: this function is out of line, and a kernel's register allocation covers its callees -- a wider
    // unroll here raised EVERY InfoNCE kernel to 96 registers (5 instead of 7 CTAs per SM).
    for (int i0 = tid * 2; i0 < n; i0 += 128 * 2) {
        double acc0 = 0.0, acc1 = 0.0;
        for (int r0 = 0; r0 < world; r0 += 8) {
            double v[8][2];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const bool on = r0 + r < world;
                const double* src = reinterpret_cast<const double*>(p.xchg_peers[on ? r0 + r : rank]) + slot + i0;
                v[r][0] = on ? x_ld_relaxed_sys_f64(src) : 0.0;
                v[r][1] = (on && i0 + 1 < n) ? x_ld_relaxed_sys_f64(src + 1) : 0.0;
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) { acc0 += v[r][0]; acc1 += v[r][1]; }
        }
        p.xchg_out[i0] = acc0;
        if (i0 + 1 < n) p.xchg_out[i0 + 1] = acc1;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const volatile double* sums = p.xchg_out;
        const int C = p.C, D = p.D;
        int nv = 0;
        uint32_t changed = 0;
        for (int k = 0; k < C; ++k)
            if (sums[(int64_t)k * (D + 1) + D] > 0.0) { changed |= pl->valid_class[nv] != k; pl->valid_class[nv++] = k; }
        for (int k = nv; k < ARCO_MAX_CLASSES; ++k) { changed |= pl->valid_class[k] != -1; pl->valid_class[k] = -1; }
        pl->n_valid = nv;
        for (int pos = 0; pos < ARCO_MAX_CLASSES; ++pos) {
            int act = 0;
            if (nv > 1 && pos < nv) {
                const int bank_cls = pl->valid_class[pos];
                act = (pl->n_anchor[pos] > 0 && pl->bank_len[bank_cls] > 0) ? 1 : 0;
            }
            changed |= pl->slot_active[pos] != act;
            pl->slot_active[pos] = act;
        }
        pl->inv_scale = nv > 1 ? 1.0f / ((float)p.Q * (float)nv) : 0.f;
        pl->replanned = changed;
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&pl->reserved0), "r"(1u) : "memory");
    }
}

// Query CTAs of a sharded launch: wait (after the negatives pass) until block 0 has published the global sums and plan.
// Returns true when the plan changed under the speculation: this launch then emits nothing and the gated redo launch that
// follows it (after arco_sample_if_replanned) does the whole job on the global plan.  All 128 threads.
__device__ __forceinline__ bool exchange_wait(const InfoParams& p) {
    if (threadIdx.x == 0) {
        uint32_t v = 0;
        while (true) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&p.plan->reserved0) : "memory");
            if (v) break;
            __nanosleep(200);
        }
    }
    __syncthreads();
    return *reinterpret_cast<volatile uint32_t*>(&p.plan->replanned) != 0;
}

// d loss / d anchor and the per-query loss from the merged softmax statistics.  gsum(d) = sum_k e_k k_hat_k[d] over the
// negatives, S_all / S2_all include the positive key's e0 and e0*cos0, m_all is the offset the exponentials were taken at.
template <typename GSum>
__device__ __forceinline__ void info_epilogue(const InfoParams& p, int bid, int j, int q, const AnchorInfo& ai, float inv_scale,
                                              float e0, float m_all, float S_all, float S2_all, const float* a_hat,
                                              const float* k0hat, GSum gsum) {
    const int tid = threadIdx.x, D = p.D;
    const float inv_temp = 1.f / p.temp;
    const float z0 = ai.cos0 * inv_temp;
    const float inv_S = 1.f / S_all;
    const float sdot = (S2_all * inv_S - ai.cos0) * inv_temp;
    for (int d = tid; d < D; d += 128) {
        const float g = gsum(d) + e0 * k0hat[d];
        const float gw = (g * inv_S - k0hat[d]) * inv_temp;
        const float grad = (ai.na > kEps) ? (gw - sdot * a_hat[d]) / ai.na : gw / kEps;
        p.g_anchor[(int64_t)bid * D + d] = grad * inv_scale;
    }
    if (tid == 0) {
        p.loss_parts[bid] = (__logf(S_all) + m_all - z0) * inv_scale;
        p.anchor_pix[bid] = ai.pix;
        if (p.logits) p.logits[((int64_t)j * p.Q + q) * (1 + p.N)] = ai.cos0;
    }
}

// last CTA folds the per-query losses in a fixed order (deterministic); called by all 128 threads
__device__ __forceinline__ void info_fold_loss(const InfoParams& p) {
    __shared__ bool s_last;
    __shared__ float s_fold[128];
    const int tid = threadIdx.x;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&p.plan->loss_done, 1u) == (uint32_t)p.nq - 1u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const volatile float* lp = p.loss_parts;
    const int total = p.nq;
    float acc = 0.f;
    const int per = (total + 127) / 128;
    for (int i = tid * per; i < min(total, (tid + 1) * per); ++i) acc += lp[i];
    s_fold[tid] = acc;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int i = 0; i < 128; ++i) s += s_fold[i];
        p.loss[0] = s;
        p.plan->loss_done = 0;                               // re-arm: the entry point may run again on the same plan
    }
    // ---- end of the step: zero-copy host mirror of the summary, then advance the bank's device step counter ----
    const uint32_t seq = p.plan->step_ctr + 1u;
    if (p.host_mirror) {
        // (arco_bank.host_mirror) plan words -> slot seq % SLOTS, live queue pointers, system fence, sequence number
        char* slot = reinterpret_cast<char*>(p.host_mirror) + (size_t)(seq % ARCO_MIRROR_SLOTS) * ARCO_MIRROR_STRIDE;
        const volatile uint32_t* src = reinterpret_cast<const volatile uint32_t*>(p.plan);
        volatile uint32_t* dst = reinterpret_cast<volatile uint32_t*>(slot);
        for (int i = tid; i < (int)(sizeof(arco_plan) / 4); i += 128) dst[i] = src[i];
        if (p.host_queue_ptr && tid < p.C) reinterpret_cast<volatile long long*>(p.host_queue_ptr)[tid] = p.plan->queue_ptr[tid];
        __syncthreads();                                     // the release below is cumulative over the CTA's stores (one
        if (tid == 0) {                                      // system fence by one thread instead of one per thread)
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(slot + sizeof(arco_plan)), "l"((unsigned long long)seq) : "memory");
        }
    }
    if (p.step_ctr && tid == 0) *p.step_ctr = seq;
}

// Common entry of the three InfoNCE kernels.  Declares bid, j, q, pl, bank_cls0, active, redo.
//  * a gated launch (the redo after a changed plan) returns at once unless plan->replanned;
//  * block 0 of a sharded launch is the exchange block;
//  * `active` also requires a non-empty bank: on a batch shard the plan may be re-derived by block 0 WHILE the query CTAs read
//    it; valid classes are only ever added (global counts >= local counts), so valid_class[j] stays a class id, but its bank
//    may be empty -- such a launch is discarded (replanned != 0) and must merely stay in bounds.
#define INFO_ENTRY(p)                                                                                                  \
    constexpr bool ROWS = MODE == 1, SHARD = MODE == 2;           /* MODE: 0 plain, 1 anchors as rows, 2 batch shard */  \
    if (SHARD && (p).gate_replanned && *reinterpret_cast<volatile uint32_t*>(&(p).plan->replanned) == 0u) return;       \
    const int xc_ = (SHARD && (p).xchg_peers) ? 1 : 0;                                                                 \
    const bool SPEC = xc_ != 0;                                   /* speculative pass on the rank-local plan */         \
    if (SHARD && xc_ && blockIdx.x == 0) {                                                                             \
        XchgArgs xa_;                                                                                                   \
        xa_.xchg_peers = (p).xchg_peers; xa_.xchg_out = (p).xchg_out; xa_.xchg_seq = (p).xchg_seq;                      \
        xa_.xchg_slot = (p).xchg_slot; xa_.plan = (p).plan; xa_.xchg_rank = (p).xchg_rank;                              \
        xa_.xchg_world = (p).xchg_world; xa_.C = (p).C; xa_.D = (p).D; xa_.Q = (p).Q;                                   \
        exchange_block(xa_);                                                                                            \
        return;                                                                                                        \
    }                                                                                                                  \
    const int bid = (int)blockIdx.x - xc_;                                                                             \
    const int j = bid / (p).Q, q = bid % (p).Q;                   /* LOOP-2 position (loss_helper_3d.py:435), query */  \
    arco_plan* pl = (p).plan;                                                                                          \
    const int bank_cls0 = pl->valid_class[j];                                                                          \
    const bool active = pl->slot_active[j] != 0 && (!SHARD || (bank_cls0 >= 0 && pl->bank_len[bank_cls0] > 0));        \
    bool redo = false

// MAXIT: 16-byte chunks per lane in pass 2 (ceil(chunks per row / 32)); BF16BANK: the ring stores bf16 rows
template <int MAXIT, bool BF16BANK, int MODE>
__global__ void __launch_bounds__(128) infonce_kernel(InfoParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CHD = BF16BANK ? 8 : 4;                         // feature dims per 16-byte chunk of a bank row
    const int D = p.D, CPL = D / CHD;
    float* a_hat = reinterpret_cast<float*>(smem_raw);            // [D]
    float* k0hat = a_hat + D;                                     // [D]
    float* gbuf = k0hat + D;                                      // [4][D]
    uint4* stage = reinterpret_cast<uint4*>(gbuf + 4 * D);        // [4][KC*RS16]  one stage per warp
    __shared__ __align__(8) uint64_t bars[4];
    __shared__ float s_red[4][2];
    __shared__ float s_stats[4][3];
    __shared__ int s_pix;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    INFO_ENTRY(p);                                                // gated redo / exchange block / bid, j, q, pl, active, redo
    if (active) {
        const int bank_cls = bank_cls0;                           // trap 1: bank by CLASS ID ...
        const int blen = pl->bank_len[bank_cls];
        const int bhead = pl->bank_head[bank_cls];
        const int cap = p.cap[bank_cls];
        const uint32_t row_bytes = (uint32_t)D * (BF16BANK ? 2u : 4u);
        const unsigned char* bank = reinterpret_cast<const unsigned char*>(p.bank_rows) + p.row_off[bank_cls] * (int64_t)row_bytes;

        if (tid < 4) mbar_init(&bars[tid], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");

        AnchorInfo ai = info_anchor<ROWS>(p, j, q, a_hat, s_red, &s_pix);
        if (!SPEC) ai.cos0 = info_proto(p, j, q, bank_cls, a_hat, k0hat);   // single GPU: under the first gathers
        const float inv_temp = 1.f / p.temp;

        // ---- this warp's share of the negatives ----
        const int KC = p.KC, RS16 = p.RS16;
        const int npw = (p.N + 3) / 4;
        const int n_begin = min(p.N, warp * npw), n_end = min(p.N, n_begin + npw);
        const int cnt = n_end - n_begin;
        const int nchunks = (cnt + KC - 1) / KC;
        const int32_t* my_idx = p.idx_n + ((int64_t)j * p.Q + q) * p.N + n_begin;
        uint4* wstage = stage + (size_t)warp * KC * RS16;

        // Occupancy instead of double buffering: a warp owns ONE stage (<= 9 KB); with ~5 CTAs (20 warps) per
        // SM the other warps' gathers and math hide this warp's wait.
        auto issue = [&](int chunk) {
            const int nv = min(KC, cnt - chunk * KC);
            if (lane == 0) mbar_expect_tx(&bars[warp], (uint32_t)nv * row_bytes);
            __syncwarp();
            if (lane < nv) {
                int r = my_idx[chunk * KC + lane];
                if (r < 0 || r >= blen) { if (!SPEC) atomicOr(&p.plan->status, (uint32_t)ARCO_ST_INDEX_RANGE); r = min(max(r, 0), blen - 1); }
                int phys = bhead + r;
                if (phys >= cap) phys -= cap;
                bulk_g2s(wstage + (size_t)lane * RS16, bank + (int64_t)phys * row_bytes, row_bytes, &bars[warp]);
            }
        };

        const int kq = lane % KC, seg = lane / KC, DSEG = 32 / KC;
        int cplp = 1;
        while (cplp < CPL && cplp < 32) cplp <<= 1;              // lanes that tile one row in pass 2
        const int KP = 32 / cplp, kpar = lane / cplp, chl = lane % cplp;

        float m_run = -INFINITY, S_run = 0.f, S2_run = 0.f;
        float G[MAXIT][CHD];
#pragma unroll
        for (int it = 0; it < MAXIT; ++it)
#pragma unroll
            for (int e = 0; e < CHD; ++e) G[it][e] = 0.f;

        for (int chunk = 0; chunk < nchunks; ++chunk) {
            issue(chunk);
            mbar_wait(&bars[warp], (uint32_t)(chunk & 1));
            const int nv = min(KC, cnt - chunk * KC);
            const uint4* rows = wstage;
            // pass 1: lane (key kq, segment seg) -> dot and squared norm
            float dot = 0.f, n2 = 0.f;
            if (kq < nv) {
                const uint4* r16 = rows + (size_t)kq * RS16;
                float dacc[CHD], nacc[CHD];                                   // CHD independent FMA chains each
#pragma unroll
                for (int e = 0; e < CHD; ++e) { dacc[e] = 0.f; nacc[e] = 0.f; }
                for (int ch = seg; ch < CPL; ch += DSEG) {
                    float kv[CHD];
                    unpack_chunk<BF16BANK>(r16[ch], kv);
                    const float* av = a_hat + ch * CHD;
#pragma unroll
                    for (int e = 0; e < CHD; ++e) { dacc[e] += kv[e] * av[e]; nacc[e] += kv[e] * kv[e]; }
                }
#pragma unroll
                for (int e = 0; e < CHD; ++e) { dot += dacc[e]; n2 += nacc[e]; }
            }
            for (int o = KC; o < 32; o <<= 1) {
                dot += __shfl_xor_sync(0xffffffffu, dot, o);
                n2 += __shfl_xor_sync(0xffffffffu, n2, o);
            }
            const bool valid = kq < nv;
            const float inv_nk = 1.f / fmaxf(sqrtf(n2), kEps);
            const float cosv = dot * inv_nk;
            const float z = valid ? cosv * inv_temp : -INFINITY;
            if (p.logits && valid && seg == 0)
                p.logits[((int64_t)j * p.Q + q) * (1 + p.N) + 1 + n_begin + chunk * KC + kq] = cosv;
            const float m_new = fmaxf(m_run, warp_max(z));
            const float rescale = __expf(m_run - m_new);          // exp(-inf) = 0 on the first chunk
            const float e = valid ? __expf(z - m_new) : 0.f;
            const float es = (seg == 0) ? e : 0.f;
            S_run = S_run * rescale + warp_sum(es);
            S2_run = S2_run * rescale + warp_sum(es * cosv);
            m_run = m_new;
            const float coef = e * inv_nk;
            // pass 2: G += sum_k coef_k * key_k, lanes tile the row in 16-byte chunks
#pragma unroll
            for (int it = 0; it < MAXIT; ++it)
#pragma unroll
                for (int e2 = 0; e2 < CHD; ++e2) G[it][e2] *= rescale;
            for (int kk0 = 0; kk0 < nv; kk0 += KP) {
                const int kk = kk0 + kpar;
                const float ck = __shfl_sync(0xffffffffu, coef, kk < KC ? kk : 0);
                if (kk < nv) {
                    const uint4* r16 = rows + (size_t)kk * RS16;
#pragma unroll
                    for (int it = 0; it < MAXIT; ++it) {
                        const int ch = chl + 32 * it;
                        if (ch < CPL) {
                            float kv[CHD];
                            unpack_chunk<BF16BANK>(r16[ch], kv);
#pragma unroll
                            for (int e2 = 0; e2 < CHD; ++e2) G[it][e2] += ck * kv[e2];
                        }
                    }
                }
            }
            __syncwarp();
        }
        for (int o = cplp; o < 32; o <<= 1) {
#pragma unroll
            for (int it = 0; it < MAXIT; ++it)
#pragma unroll
                for (int e2 = 0; e2 < CHD; ++e2) G[it][e2] += __shfl_xor_sync(0xffffffffu, G[it][e2], o);
        }
        if (kpar == 0) {
#pragma unroll
            for (int it = 0; it < MAXIT; ++it) {
                const int ch = chl + 32 * it;
                if (ch < CPL) {
#pragma unroll
                    for (int e2 = 0; e2 < CHD; ++e2) gbuf[(size_t)warp * D + ch * CHD + e2] = G[it][e2];
                }
            }
        }
        if (lane == 0) { s_stats[warp][0] = m_run; s_stats[warp][1] = S_run; s_stats[warp][2] = S2_run; }
        __syncthreads();

        redo = SPEC ? exchange_wait(p) : false;                                  // batch shard: the global sums / plan are in place now
        if (!redo) {
            // ---- the positive key, then merge the 4 partial softmaxes with it, emit loss and d loss / d anchor ----
            const float inv_scale = pl->inv_scale;
            if (SPEC) ai.cos0 = info_proto(p, j, q, bank_cls, a_hat, k0hat);
            const float cos0 = ai.cos0;
            const float z0 = cos0 * inv_temp;
            float m_all = z0;
#pragma unroll
            for (int w = 0; w < 4; ++w) m_all = fmaxf(m_all, s_stats[w][0]);
            float f[4];
            const float e0 = __expf(z0 - m_all);
            float S_all = e0, S2_all = e0 * cos0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                f[w] = s_stats[w][1] > 0.f ? __expf(s_stats[w][0] - m_all) : 0.f;
                S_all += s_stats[w][1] * f[w];
                S2_all += s_stats[w][2] * f[w];
            }
            info_epilogue(p, bid, j, q, ai, inv_scale, e0, m_all, S_all, S2_all, a_hat, k0hat, [&](int d) {
                return gbuf[d] * f[0] + gbuf[D + d] * f[1] + gbuf[2 * D + d] * f[2] + gbuf[3 * D + d] * f[3];
            });
        }
    } else {
        redo = SPEC ? exchange_wait(p) : false;
        if (!redo && tid == 0) {
            p.loss_parts[bid] = 0.f;
            p.anchor_pix[bid] = -1;
        }
    }
    if (!redo) info_fold_loss(p);
}


// ======================================================================================================================
// Short rows (<= 128 bytes: LA D=16, D=32 fp32, D<=64 bf16): one LANE per key.  The staged kernel above issues one TMA bulk copy per row and
// then spends ~26 warp instructions per key on cross-lane reductions (ncu at D=64: issue slots 61 % busy, DRAM 10 %: it
// is instruction-bound, 0.40 of the HBM roofline; 0.07 at D=16).  Here a lane loads its OWN row with 16-byte loads (NCH of
// them in flight per lane, the row's two 128-byte lines stay in L1 between them), so the dot product, the norm and the
// gradient term  G += e_k/|k_k| * k_k  are plain per-lane FMA chains with no shuffle at all; the 32 lanes' G vectors meet
// once per warp at the end.  ~7 warp instructions per key.  cos <= 1, so every exponential is taken at the fixed offset
// 1/temp (no running maximum; needs temp >= 0.03 like the mma.sync kernel).
// ======================================================================================================================
template <int NCH, bool BF16BANK, int MODE>
__global__ void __launch_bounds__(128, 3) infonce_lane_kernel(InfoParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CHD = BF16BANK ? 8 : 4;                         // feature dims per 16-byte chunk
    constexpr int DD = NCH * CHD;                                 // == p.D
    float* a_hat = reinterpret_cast<float*>(smem_raw);            // [D]
    float* k0hat = a_hat + DD;                                    // [D]
    float* gbuf = k0hat + DD;                                     // [4][D]
    int32_t* s_idx = reinterpret_cast<int32_t*>(gbuf + 4 * DD);   // [N] this query's negatives as physical ring rows
    __shared__ float s_red[4][2];
    __shared__ float s_stats[4][2];
    __shared__ int s_pix;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    INFO_ENTRY(p);
    if (active) {
        const int bank_cls = bank_cls0;                           // trap 1: bank by CLASS ID
        const int blen = pl->bank_len[bank_cls], bhead = pl->bank_head[bank_cls], cap = p.cap[bank_cls];
        constexpr uint32_t row_bytes = NCH * 16;
        const unsigned char* bank = reinterpret_cast<const unsigned char*>(p.bank_rows) + p.row_off[bank_cls] * (int64_t)row_bytes;
        {
            const int32_t* src = p.idx_n + ((int64_t)j * p.Q + q) * p.N;
            bool bad = false;
            for (int n = tid; n < p.N; n += 128) {
                int r = src[n];
                if (r < 0 || r >= blen) { bad = true; r = min(max(r, 0), blen - 1); }
                int phys = bhead + r;
                if (phys >= cap) phys -= cap;
                s_idx[n] = phys;
            }
            if (bad && !SPEC) atomicOr(&p.plan->status, (uint32_t)ARCO_ST_INDEX_RANGE);   // (a speculative pass may see a plan in flux)
        }
        AnchorInfo ai = info_anchor<ROWS>(p, j, q, a_hat, s_red, &s_pix);   // (contains __syncthreads)
        if (!SPEC) ai.cos0 = info_proto(p, j, q, bank_cls, a_hat, k0hat);
        const float inv_temp = 1.f / p.temp;

        const int npw = (p.N + 3) / 4;
        const int n_begin = min(p.N, warp * npw), n_end = min(p.N, n_begin + npw);
        float S = 0.f, S2 = 0.f;
        float G[DD];
#pragma unroll
        for (int d = 0; d < DD; ++d) G[d] = 0.f;

        for (int base = n_begin; base < n_end; base += 32) {      // warp-uniform trip count
            const int n = base + lane;
            const bool val = n < n_end;
            uint4 raw[NCH];
            const uint4* row = reinterpret_cast<const uint4*>(bank + (int64_t)(val ? s_idx[n] : 0) * row_bytes);
#pragma unroll
            for (int c = 0; c < NCH; ++c) raw[c] = val ? __ldg(row + c) : make_uint4(0u, 0u, 0u, 0u);
            float kv[DD];
            float dacc[4] = {0.f, 0.f, 0.f, 0.f}, nacc[4] = {0.f, 0.f, 0.f, 0.f};      // independent FMA chains
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                unpack_chunk<BF16BANK>(raw[c], kv + c * CHD);
#pragma unroll
                for (int e4 = 0; e4 < CHD; e4 += 4) {
                    const float4 a4 = *reinterpret_cast<const float4*>(a_hat + c * CHD + e4);   // broadcast read
                    const float* kk = kv + c * CHD + e4;
                    dacc[0] += kk[0] * a4.x; dacc[1] += kk[1] * a4.y; dacc[2] += kk[2] * a4.z; dacc[3] += kk[3] * a4.w;
                    nacc[0] += kk[0] * kk[0]; nacc[1] += kk[1] * kk[1]; nacc[2] += kk[2] * kk[2]; nacc[3] += kk[3] * kk[3];
                }
            }
            const float dot = (dacc[0] + dacc[1]) + (dacc[2] + dacc[3]);
            const float n2 = (nacc[0] + nacc[1]) + (nacc[2] + nacc[3]);
            const float inv_nk = 1.f / fmaxf(sqrtf(n2), kEps);
            const float cosv = dot * inv_nk;
            const float e = val ? __expf((cosv - 1.f) * inv_temp) : 0.f;
            S += e;
            S2 += e * cosv;
            if (p.logits && val) p.logits[((int64_t)j * p.Q + q) * (1 + p.N) + 1 + n] = cosv;
            const float coef = e * inv_nk;
#pragma unroll
            for (int d = 0; d < DD; ++d) G[d] += coef * kv[d];
        }
        // the 32 lanes' partial G vectors meet once per warp
#pragma unroll
        for (int d = 0; d < DD; ++d) {
            float g = G[d];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
            if (lane == (d & 31)) gbuf[(size_t)warp * DD + d] = g;
        }
        S = warp_sum(S);
        S2 = warp_sum(S2);
        if (lane == 0) { s_stats[warp][0] = S; s_stats[warp][1] = S2; }
        __syncthreads();
        redo = SPEC ? exchange_wait(p) : false;
        if (!redo) {
            const float inv_scale = pl->inv_scale;
            if (SPEC) ai.cos0 = info_proto(p, j, q, bank_cls, a_hat, k0hat);
            const float z0 = ai.cos0 * inv_temp;
            const float m_all = inv_temp;                         // offset of every exponential: z <= 1/temp
            const float e0 = __expf(z0 - m_all);
            const float S_all = e0 + s_stats[0][0] + s_stats[1][0] + s_stats[2][0] + s_stats[3][0];
            const float S2_all = e0 * ai.cos0 + s_stats[0][1] + s_stats[1][1] + s_stats[2][1] + s_stats[3][1];
            info_epilogue(p, bid, j, q, ai, inv_scale, e0, m_all, S_all, S2_all, a_hat, k0hat,
                          [&](int d) { return gbuf[d] + gbuf[DD + d] + gbuf[2 * DD + d] + gbuf[3 * DD + d]; });
        }
    } else {
        redo = SPEC ? exchange_wait(p) : false;
        if (!redo && tid == 0) {
            p.loss_parts[bid] = 0.f;
            p.anchor_pix[bid] = -1;
        }
    }
    if (!redo) info_fold_loss(p);
}

// ======================================================================================================================
// bf16 ring: the same InfoNCE on the legacy tensor-core path (mma.sync m16n8k16, bf16 x bf16 -> fp32).
//
// The gathered rows cannot be a tcgen05 operand (a UMMA descriptor wants 8-row core matrices at a fixed stride, the
// gather delivers one row per bank position), ldmatrix takes one address per row.  Per query the work is a
// matrix-VECTOR product, so the vector operand is widened to use the 8/16-wide MMA dimension exactly:
//   pass 1  C[16 x 8 keys] += A[16 x 16 dims] . B[16 dims x 8 keys]     A rows 0-7 = the same 8 keys (diagonal of the
//           top block = |k|^2), rows 8-10 = a_hat split into three bf16 terms (hi + mid + lo carry all 24 mantissa bits;
//           the products are exact and the accumulation is fp32, so the dot equals the FFMA path's up to summation order)
//   pass 2  G[16 dims x 8] += K^T[16 dims x 16 keys] . W[16 keys x 8]  W columns 0-2 = the softmax weights e_k/|k| split
//           the same way; the three useful columns are added at the end.
// cos <= 1, so the exponentials are taken at the fixed offset 1/temp: no running maximum, no rescaling of G.
// The CTA works on 16-key chunks staged by TMA bulk copies in NSTG shared buffers; each warp owns a quarter of the
// feature dimension in BOTH passes (partial dots are added through shared memory), which keeps 32 accumulator registers
// per thread instead of D/4.
// ======================================================================================================================
constexpr int MMA_KEYS = 16;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// term `part` (0 hi, 1 mid, 2 lo) of the three-way bf16 split of x, as bf16 bits
__device__ __forceinline__ uint32_t bf16_term(float x, int part) {
    const float hi = __bfloat162float(__float2bfloat16_rn(x));
    if (part == 0) return __float_as_uint(hi) >> 16;
    const float r1 = x - hi;
    const float mid = __bfloat162float(__float2bfloat16_rn(r1));
    if (part == 1) return __float_as_uint(mid) >> 16;
    return __float_as_uint(__bfloat162float(__float2bfloat16_rn(r1 - mid))) >> 16;
}

template <int NSTG, int MODE>
__global__ void __launch_bounds__(128, NSTG == 1 ? 7 : 5) infonce_mma_kernel(InfoParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = p.D;
    const int KS = (D + 15) >> 4;                                 // 16-dim steps (pass 1 k-steps == pass 2 m-tiles)
    const int RS16 = p.RS16;                                      // odd, and RS16*8 >= KS*16: the pad stays zero
    float* a_hat = reinterpret_cast<float*>(smem_raw);            // [D]
    float* k0hat = a_hat + D;                                     // [D]
    uint2* asf = reinterpret_cast<uint2*>(k0hat + D);             // [KS][16]  A-fragment rows 8-10 (a_hat terms) + 4 zero slots
    float* s_part = reinterpret_cast<float*>(asf + KS * 16);      // [2][4][16][2]  per-warp partial (dot, |k|^2)
    uint4* stage = reinterpret_cast<uint4*>(s_part + 2 * 4 * MMA_KEYS * 2);   // [NSTG][16 * RS16]; G after the loop
    __shared__ __align__(8) uint64_t bars[NSTG];
    __shared__ float s_red[4][2];
    __shared__ int s_pix;
    __shared__ uint32_t s_done;                                   // warps finished with the current chunk's stage

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    INFO_ENTRY(p);
    if (active) {
        const int bank_cls = bank_cls0;
        const int blen = pl->bank_len[bank_cls];
        const int bhead = pl->bank_head[bank_cls];
        const int cap = p.cap[bank_cls];
        const uint32_t row_bytes = (uint32_t)D * 2u;
        const unsigned char* bank = reinterpret_cast<const unsigned char*>(p.bank_rows) + p.row_off[bank_cls] * (int64_t)row_bytes;
        const int stage_u4 = MMA_KEYS * RS16;

        if (tid < NSTG) mbar_init(&bars[tid], 1);
        if (tid == 0) s_done = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // zero the stages once: row pads and the unused rows of a short last chunk must read as finite values
        for (int i = tid; i < NSTG * stage_u4; i += 128) stage[i] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();

        const int nch = (p.N + MMA_KEYS - 1) / MMA_KEYS;
        const int32_t* my_idx = p.idx_n + ((int64_t)j * p.Q + q) * p.N;
        auto load_idx = [&](int chunk) {                          // ring row of key `lane` of the chunk (lanes 0-15)
            const int n = chunk * MMA_KEYS + lane;
            return (lane < MMA_KEYS && n < p.N) ? my_idx[n] : -1;
        };
        auto issue = [&](int chunk, int r) {                      // one whole warp
            const int s = chunk % NSTG;
            const int nv = min(MMA_KEYS, p.N - chunk * MMA_KEYS);
            if (lane == 0) mbar_expect_tx(&bars[s], (uint32_t)nv * row_bytes);
            __syncwarp();
            if (lane < nv) {
                if (r < 0 || r >= blen) { if (!SPEC) atomicOr(&p.plan->status, (uint32_t)ARCO_ST_INDEX_RANGE); r = min(max(r, 0), blen - 1); }
                int phys = bhead + r;
                if (phys >= cap) phys -= cap;
                bulk_g2s(stage + (size_t)s * stage_u4 + (size_t)lane * RS16, bank + (int64_t)phys * row_bytes, row_bytes, &bars[s]);
            }
        };
        // the first gathers do not depend on the anchor: they fly while the prologue chases its pointers
        if (warp == 0)
            for (int c = 0; c < NSTG && c < nch; ++c) issue(c, load_idx(c));
        int r_next = load_idx(NSTG);                              // every warp: any of them may issue the next chunk

        AnchorInfo ai = info_anchor<ROWS>(p, j, q, a_hat, s_red, &s_pix);
        if (!SPEC) ai.cos0 = info_proto(p, j, q, bank_cls, a_hat, k0hat);
        const float inv_temp = 1.f / p.temp;
        // A-fragment rows 8..10 of every k-step: lane (g < 3, t) holds terms g of a_hat[16ks + 2t, +1] and [.. + 8, + 9]
        // (slots 12..15 of a k-step are zero: the rows 11..15 of the A operand, read by the lanes with g >= 3)
        for (int i = tid; i < KS * 16; i += 128) {
            const int ks = i >> 4, gg = (i & 15) >> 2, tt = i & 3;
            const int d0 = 16 * ks + 2 * tt;
            auto av = [&](int d) { return d < D ? a_hat[d] : 0.f; };
            asf[i] = gg < 3 ? make_uint2(bf16_term(av(d0), gg) | (bf16_term(av(d0 + 1), gg) << 16),
                                         bf16_term(av(d0 + 8), gg) | (bf16_term(av(d0 + 9), gg) << 16))
                            : make_uint2(0u, 0u);
        }
        __syncthreads();

        const int ks_begin = (KS * warp) >> 2;                    // this warp's slice of D, both passes
        const int nks = ((KS * (warp + 1)) >> 2) - ks_begin;      // <= 8
        // ldmatrix lane address inside a stage: matrix i = lane >> 3 -> keys (i >> 1) * 8 + (lane & 7), dims + (i & 1) * 8
        const uint32_t lm_base = smem_u32(stage) +
                                 (uint32_t)(((lane >> 4) * 8 + (lane & 7)) * RS16 + ((lane >> 3) & 1) + 2 * ks_begin) * 16u;
        const uint2* af_ptr = asf + ks_begin * 16 + (g < 3 ? g * 4 + t : 12 + t);
        float acc[8][4];
#pragma unroll
        for (int m = 0; m < 8; ++m) { acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f; }
        float S_lane = 0.f, S2_lane = 0.f;
        const int key = lane & 15;

        for (int c = 0; c < nch; ++c) {
            const int s = c % NSTG;
            mbar_wait(&bars[s], (uint32_t)((c / NSTG) & 1));
            const uint32_t sbase = lm_base + (uint32_t)(s * stage_u4) * 16u;
            // ---- pass 1: partial dots / squared norms of the 16 keys over this warp's dims ----
            float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                if (m < nks) {
                    uint32_t r[4];
                    ldsm_x4(r, sbase + (uint32_t)m * 32u);
                    const uint2 af = af_ptr[m * 16];
                    mma_bf16_16816(c0, r[0], af.x, r[1], af.y, r[0], r[1]);
                    mma_bf16_16816(c1, r[2], af.x, r[3], af.y, r[2], r[3]);
                }
            }
            // rows 8..10 (c[2], c[3]) summed over g; the diagonal of rows 0..7 sits in lane 4n + n/2
#pragma unroll
            for (int o = 4; o <= 8; o <<= 1) {
                c0[2] += __shfl_xor_sync(0xffffffffu, c0[2], o); c0[3] += __shfl_xor_sync(0xffffffffu, c0[3], o);
                c1[2] += __shfl_xor_sync(0xffffffffu, c1[2], o); c1[3] += __shfl_xor_sync(0xffffffffu, c1[3], o);
            }
            float* part = s_part + ((c & 1) * 4 + warp) * (MMA_KEYS * 2);
            if (g == 0) {                                         // lane t: dots of keys 2t, 2t+1 (tile 0) and 8+2t, 9+2t (tile 1)
                part[4 * t] = c0[2]; part[4 * t + 2] = c0[3];
                part[16 + 4 * t] = c1[2]; part[16 + 4 * t + 2] = c1[3];
            }
            if (t == (g >> 1)) {                                  // |k|^2 of key g (tile 0) and key 8+g (tile 1)
                part[g * 2 + 1] = (g & 1) ? c0[1] : c0[0];
                part[(8 + g) * 2 + 1] = (g & 1) ? c1[1] : c1[0];
            }
            __syncthreads();
            // ---- softmax weights of the 16 keys (every warp computes all of them; lanes 16-31 mirror 0-15) ----
            const float2* pr = reinterpret_cast<const float2*>(s_part + (c & 1) * 4 * (MMA_KEYS * 2)) + key;
            const float2 p0 = pr[0], p1 = pr[MMA_KEYS], p2 = pr[2 * MMA_KEYS], p3 = pr[3 * MMA_KEYS];
            const float dot = (p0.x + p1.x) + (p2.x + p3.x);
            const float n2 = (p0.y + p1.y) + (p2.y + p3.y);
            const bool valid = c * MMA_KEYS + key < p.N;
            const float inv_nk = 1.f / fmaxf(sqrtf(n2), kEps);
            const float cosv = dot * inv_nk;
            const float e = valid ? __expf((cosv - 1.f) * inv_temp) : 0.f;
            const float coef = e * inv_nk;
            if (warp == 0 && lane < MMA_KEYS) {
                S_lane += e;
                S2_lane += e * cosv;
                if (p.logits && valid) p.logits[((int64_t)j * p.Q + q) * (1 + p.N) + 1 + c * MMA_KEYS + key] = cosv;
            }
            // B fragment of pass 2: column g (< 3) = term g of the weights of keys (2t, 2t+1) and (2t+8, 2t+9).
            // Every lane splits its own key's weight once; lanes 0-15 publish (hi | mid << 16), their mirrors 16-31 publish lo,
            // so one shuffle per key fetches the term a lane's column needs.
            uint32_t wpub;
            {
                const float hi = __bfloat162float(__float2bfloat16_rn(coef));
                const float r1 = coef - hi;
                const float mid = __bfloat162float(__float2bfloat16_rn(r1));
                const float lo = __bfloat162float(__float2bfloat16_rn(r1 - mid));
                wpub = lane < MMA_KEYS ? ((__float_as_uint(hi) >> 16) | (__float_as_uint(mid) & 0xffff0000u)) : (__float_as_uint(lo) >> 16);
            }
            const int src_hi = g == 2 ? MMA_KEYS : 0;
            const uint32_t sh = g == 1 ? 16u : 0u;
            const uint32_t u0 = (__shfl_sync(0xffffffffu, wpub, 2 * t + src_hi) >> sh) & 0xffffu;
            const uint32_t u1 = (__shfl_sync(0xffffffffu, wpub, 2 * t + 1 + src_hi) >> sh) & 0xffffu;
            const uint32_t u2 = (__shfl_sync(0xffffffffu, wpub, 2 * t + 8 + src_hi) >> sh) & 0xffffu;
            const uint32_t u3 = (__shfl_sync(0xffffffffu, wpub, 2 * t + 9 + src_hi) >> sh) & 0xffffu;
            const uint32_t b0 = g < 3 ? (u0 | (u1 << 16)) : 0u;
            const uint32_t b1 = g < 3 ? (u2 | (u3 << 16)) : 0u;
            // ---- pass 2: G[dims of this warp] += K^T . W ----
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                if (m < nks) {
                    uint32_t r[4];
                    ldsm_x4_trans(r, sbase + (uint32_t)m * 32u);
                    mma_bf16_16816(acc[m], r[0], r[1], r[2], r[3], b0, b1);
                }
            }
            // ---- the last warp to leave the stage refills it with chunk c + NSTG (nobody waits for anybody) ----
            __syncwarp();
            uint32_t prev = 0;
            if (lane == 0) { __threadfence_block(); prev = atomicAdd(&s_done, 1u); }
            prev = __shfl_sync(0xffffffffu, prev, 0);
            if (prev == 3u) {
                if (lane == 0) s_done = 0;
                if (c + NSTG < nch) issue(c + NSTG, r_next);
            }
            r_next = load_idx(c + NSTG + 1);
        }
        __syncthreads();                                          // every warp is done with the stages: reuse as G
        float* gbuf = reinterpret_cast<float*>(stage);
        // columns 0..2 of every accumulator tile -> G
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int mt = ks_begin + m;
            float lo = acc[m][0] + acc[m][1], hi = acc[m][2] + acc[m][3];
            lo += __shfl_xor_sync(0xffffffffu, lo, 1); lo += __shfl_xor_sync(0xffffffffu, lo, 2);
            hi += __shfl_xor_sync(0xffffffffu, hi, 1); hi += __shfl_xor_sync(0xffffffffu, hi, 2);
            if (m < nks && t == 0) { gbuf[16 * mt + g] = lo; gbuf[16 * mt + 8 + g] = hi; }
        }
        if (warp == 0) {
            const float S = warp_sum(S_lane), S2 = warp_sum(S2_lane);
            if (lane == 0) { s_red[0][0] = S; s_red[0][1] = S2; }
        }
        __syncthreads();
        redo = SPEC ? exchange_wait(p) : false;
        if (!redo) {
            const float inv_scale = pl->inv_scale;
            if (SPEC) ai.cos0 = info_proto(p, j, q, bank_cls, a_hat, k0hat);
            const float z0 = ai.cos0 * inv_temp;
            const float m_all = inv_temp;                         // offset of every exponential: z <= 1/temp
            const float e0 = __expf(z0 - m_all);
            info_epilogue(p, bid, j, q, ai, inv_scale, e0, m_all, e0 + s_red[0][0], e0 * ai.cos0 + s_red[0][1], a_hat, k0hat,
                          [&](int d) { return gbuf[d]; });
        }
    } else {
        redo = SPEC ? exchange_wait(p) : false;
        if (!redo && tid == 0) {
            p.loss_parts[bid] = 0.f;
            p.anchor_pix[bid] = -1;
        }
    }
    if (!redo) info_fold_loss(p);
}


// SURVEY 8(f) rank 2 (sampled producers): the step's C*Q anchors as ROWS of the tensor the student's 1x1 convolutions read.
// One CTA per (LOOP-2 position, query): warp 0 rank-selects the pixel, all threads gather its D channel values (one strided
// element per channel, the same access the InfoNCE prologue makes) into rows[j*Q+q][0..D) in fp32.  Inactive positions get
// pixel -1 and a zero row, so whatever is computed from the rows downstream stays finite and their gradient is dropped.
__global__ void __launch_bounds__(128) anchor_gather_kernel(InfoParams p, float* __restrict__ rows, int32_t* __restrict__ pix_out) {
    __shared__ int s_pix;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int bid = blockIdx.x;
    const int j = bid / p.Q, q = bid % p.Q;
    const bool active = p.plan->slot_active[j] != 0;
    if (!active) {
        for (int d = tid; d < p.D; d += 128) rows[(int64_t)bid * p.D + d] = 0.f;
        if (tid == 0) pix_out[bid] = -1;
        return;
    }
    if (warp == 0) anchor_select_warp(p, j, q, &s_pix);
    __syncthreads();
    const int pix = s_pix;
    const int ab = (int)(pix / p.S);
    const int64_t as = pix - (int64_t)ab * p.S;
    for (int d = tid; d < p.D; d += 128) {
        float v;
        if (p.rep_dtype == ARCO_BF16)
            v = bf16_bits_to_float(reinterpret_cast<const unsigned short*>(p.rep)[((int64_t)ab * p.D + d) * p.S + as]);
        else
            v = reinterpret_cast<const float*>(p.rep)[((int64_t)ab * p.D + d) * p.S + as];
        rows[(int64_t)bid * p.D + d] = v;
    }
    if (tid == 0) pix_out[bid] = pix;
}

}  // namespace arco

static int infonce_impl(const arco_dims* dims, const void* rep, const arco_bank* bank, const double* proto_sums,
                        const int32_t* idx_anchor, const int32_t* idx_neg, float temp, float* loss,
                        float* grad_anchor, int32_t* anchor_pix, float* logits, const float* momentum,
                        const int32_t* momentum_on, float ema_decay, float ema_keep, float* proto_out, void* workspace, void* stream,
                        const float* anchor_rows = nullptr, const int32_t* anchor_pix_in = nullptr,
                        const arco_exchange* xchg = nullptr, int gate_replanned = 0) {
    ARCO_REQUIRE(dims && (rep || anchor_rows) && bank && proto_sums && idx_anchor && idx_neg && loss && grad_anchor && anchor_pix &&
                     workspace, "arco_infonce: NULL argument");
    const arco_dims& d = *dims;
    ARCO_REQUIRE(d.feat % 4 == 0 && d.feat >= 4 && d.feat <= 512, "feat (D) must be a multiple of 4 in [4, 512]");
    ARCO_REQUIRE(d.queries > 0 && d.negatives >= 0 && temp > 0.f, "bad queries/negatives/temp");
    arco_ws_layout L;
    arco::compute_layout(d, &L);
    char* ws = (char*)workspace;
    arco::InfoParams p;
    p.rep = rep; p.bank_rows = bank->rows; p.proto_sums = proto_sums;
    p.idx_a = idx_anchor; p.idx_n = idx_neg;
    p.codes = (const uint8_t*)(ws + L.codes);
    p.off_anchor = (const uint32_t*)(ws + L.off_anchor);
    p.plan = (arco_plan*)(ws + L.plan);
    p.loss = loss; p.g_anchor = grad_anchor; p.anchor_pix = anchor_pix; p.logits = logits;
    p.anchor_rows = anchor_rows; p.anchor_pix_in = anchor_pix_in;
    p.xchg_peers = nullptr; p.xchg_out = nullptr; p.xchg_seq = 0; p.xchg_slot = 0; p.xchg_rank = 0; p.xchg_world = 1;
    p.gate_replanned = gate_replanned; p.nq = d.classes * d.queries;
    if (xchg) {
        ARCO_REQUIRE(xchg->peers && xchg->world >= 1 && xchg->world <= 64 && xchg->rank >= 0 && xchg->rank < xchg->world &&
                         ((xchg->seq & ARCO_XCHG_SEQ_MASK) > 0 || (xchg->seq & ARCO_XCHG_SEQ_FROM_DEVICE)) &&
                         (!(xchg->seq & ARCO_XCHG_SEQ_FROM_DEVICE) || (xchg->seq & ARCO_XCHG_STEP_WORD)) &&
                         xchg->slot_doubles >= (int64_t)d.classes * (d.feat + 1), "arco_infonce_sharded: bad exchange descriptor");
        p.xchg_peers = (const unsigned long long*)xchg->peers; p.xchg_out = const_cast<double*>(proto_sums);
        p.xchg_seq = xchg->seq; p.xchg_slot = xchg->slot_doubles; p.xchg_rank = xchg->rank; p.xchg_world = xchg->world;
    }
    p.loss_parts = (float*)(ws + L.loss_parts);
    p.momentum = momentum; p.momentum_on = momentum_on; p.proto_out = proto_out; p.ema_decay = ema_decay; p.ema_keep = ema_keep;
    p.host_mirror = bank->host_mirror; p.host_queue_ptr = bank->host_queue_ptr;
    p.step_ctr = bank->counters ? bank->counters + ARCO_CTR_STEP : nullptr;
    ARCO_REQUIRE(momentum == nullptr || momentum_on != nullptr, "momentum needs the device flag momentum_on");
    for (int c = 0; c < ARCO_MAX_CLASSES; ++c) { p.row_off[c] = bank->row_off[c]; p.cap[c] = bank->cap[c] > 0 ? bank->cap[c] : 1; }
    p.S = d.space; p.C = d.classes; p.D = d.feat; p.Q = d.queries; p.N = d.negatives;
    p.tpi = L.tiles_per_image; p.NT = L.n_tiles; p.rep_dtype = d.rep_dtype; p.temp = temp;
    // stage geometry: KC keys per chunk (<= 9 KB per stage), row stride padded for conflict-free 16-B reads
    const bool bf16bank = bank->row_dtype == ARCO_BF16;
    ARCO_REQUIRE(!bf16bank || d.feat % 8 == 0, "a bf16 bank needs D to be a multiple of 8");
    const int cpl = d.feat / (bf16bank ? 8 : 4);
    static const int64_t stage_budget = [] {
        const char* e = getenv("ARCO_INFONCE_STAGE");               // tuning knob: bytes of staged rows per warp
        return e && atoi(e) > 0 ? (int64_t)atoi(e) : (int64_t)9216;
    }();
    int kc = 32;
    while (kc > 4 && (int64_t)kc * (cpl + 2) * 16 > stage_budget) kc >>= 1;
    // One wave beats deeper stages: if halving the chunk lets every CTA of the grid be resident at once, do it
    // (measured at D = 64, Q = 256, C = 4: 0.045 vs 0.058 ms; the kernels use <= 96 registers -> at most 5 CTAs per SM).
    {
        const int grid_ctas = d.classes * d.queries;
        const int need = (grid_ctas + arco::sm_count() - 1) / arco::sm_count();
        auto per_sm = [&](int k) {
            const size_t sm = (size_t)6 * d.feat * 4 + (size_t)4 * k * (cpl + 2) * 16 + 1024;
            const int by_smem = (int)((size_t)(227 * 1024) / sm);
            return by_smem < 16 ? by_smem : 16;
        };
        while (kc > 8 && per_sm(kc) < need && per_sm(kc / 2) >= need) kc >>= 1;
    }
    int rs16 = cpl;
    if (kc >= 8) { while ((rs16 & 1) == 0) ++rs16; } else { while ((rs16 & 3) != 2) ++rs16; }
    p.KC = kc; p.RS16 = rs16;
    const size_t smem = (size_t)6 * d.feat * 4 + (size_t)4 * kc * rs16 * 16;
    const int grid = d.classes * d.queries + (xchg ? 1 : 0);         // + the exchange block (block 0)
    cudaStream_t st = (cudaStream_t)stream;
    // Three instantiations of every kernel: MODE 0 = the plain single-GPU op (nothing of the other two modes is compiled in:
    // they cost registers, and 72 registers = 7 CTAs per SM = ONE wave at C*Q = 1024 is what the small shapes live on),
    // 1 = anchors given as rows (arco_infonce_rows), 2 = batch shard (exchange block / gated redo, arco_infonce_sharded).
    const int mode = anchor_rows ? 1 : ((xchg || gate_replanned) ? 2 : 0);
    ARCO_REQUIRE(!(anchor_rows && (xchg || gate_replanned)), "arco_infonce: anchors-as-rows and the sharded exchange do not combine");
#define ARCO_LAUNCH_MODE(KERNEL, SMEM, ...)                                                                                \
    do {                                                                                                                   \
        if (mode == 0) {                                                                                                   \
            ARCO_CUDA_CHECK(cudaFuncSetAttribute(KERNEL<__VA_ARGS__, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM))); \
            KERNEL<__VA_ARGS__, 0><<<grid, 128, (SMEM), st>>>(p);                                                          \
        } else if (mode == 1) {                                                                                            \
            ARCO_CUDA_CHECK(cudaFuncSetAttribute(KERNEL<__VA_ARGS__, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM))); \
            KERNEL<__VA_ARGS__, 1><<<grid, 128, (SMEM), st>>>(p);                                                          \
        } else {                                                                                                           \
            ARCO_CUDA_CHECK(cudaFuncSetAttribute(KERNEL<__VA_ARGS__, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM))); \
            KERNEL<__VA_ARGS__, 2><<<grid, 128, (SMEM), st>>>(p);                                                          \
        }                                                                                                                  \
    } while (0)
#define ARCO_INFONCE(MI, BF) ARCO_LAUNCH_MODE(arco::infonce_kernel, smem, MI, BF)
    static const bool lane_ok = [] { const char* e = getenv("ARCO_INFONCE_LANE"); return !(e && e[0] == '0'); }();
    // Measured on B200 (cold bank, Q=256, N=512): 128-byte rows and shorter gain (LA D=16: 0.027 vs 0.036 ms staged), a 256-byte
    // row does not (ACDC D=64: 0.086 vs 0.051 ms -- 32 lanes x 2 lines per load instruction thrash L1), so longer rows stay staged.
    if (lane_ok && temp >= 0.03f && cpl <= 8 && (cpl == 2 || cpl == 4 || cpl == 8)) {
        // short rows: one lane per key (see infonce_lane_kernel)
        const size_t sm = (size_t)6 * d.feat * 4 + (size_t)(d.negatives > 0 ? d.negatives : 1) * 4;
#define ARCO_INFONCE_LANE(NC, BF) ARCO_LAUNCH_MODE(arco::infonce_lane_kernel, sm, NC, BF)
        if (bf16bank) { if (cpl == 2) ARCO_INFONCE_LANE(2, true); else if (cpl == 4) ARCO_INFONCE_LANE(4, true); else if (cpl == 8) ARCO_INFONCE_LANE(8, true); else goto staged; }
        else { if (cpl == 2) ARCO_INFONCE_LANE(2, false); else if (cpl == 4) ARCO_INFONCE_LANE(4, false); else ARCO_INFONCE_LANE(8, false); }
#undef ARCO_INFONCE_LANE
        ARCO_LAUNCH_CHECK();
        return ARCO_OK;
    }
staged:
    static const int mma_env = [] { const char* e = getenv("ARCO_INFONCE_MMA"); return e ? atoi(e) : 1; }();   // 0 = off, else stages (1 | 2)
    // Measured (profiles/r01_config5_sweep.md): the mma.sync kernel's per-chunk cost is flat in D, so it wins for long rows
    // (D = 496: 0.142 vs 0.190 ms) and loses to the FFMA kernel on the same bf16 ring at D <= 256 (0.114 vs 0.090 ms).
    if (bf16bank && mma_env > 0 && temp >= 0.03f && (d.feat >= 320 || mma_env >= 8)) {
        // tensor-core path (fixed softmax offset 1/temp needs exp(-2/temp) to stay a normal float)
        const int nstg = (mma_env & 7) >= 2 ? 2 : 1;
        const int ks = (d.feat + 15) / 16;
        int r16 = 2 * ks;                                               // >= 16*ks dims
        if ((r16 & 1) == 0) ++r16;                                      // odd stride: ldmatrix rows hit distinct banks
        p.RS16 = r16; p.KC = arco::MMA_KEYS;
        size_t stage_bytes = (size_t)nstg * arco::MMA_KEYS * r16 * 16;
        if (stage_bytes < (size_t)ks * 64) stage_bytes = (size_t)ks * 64;   // G[16*ks] lives there after the loop
        const size_t sm = (size_t)2 * d.feat * 4 + (size_t)ks * 16 * 8 + 2 * 4 * arco::MMA_KEYS * 2 * 4 + stage_bytes;
        if (nstg == 1) ARCO_LAUNCH_MODE(arco::infonce_mma_kernel, sm, 1);
        else ARCO_LAUNCH_MODE(arco::infonce_mma_kernel, sm, 2);
    } else if (bf16bank) {
        if (cpl <= 32) ARCO_INFONCE(1, true); else ARCO_INFONCE(2, true);
    } else {
        if (cpl <= 32) ARCO_INFONCE(1, false); else if (cpl <= 64) ARCO_INFONCE(2, false); else ARCO_INFONCE(4, false);
    }
#undef ARCO_INFONCE
#undef ARCO_LAUNCH_MODE
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_infonce(const arco_dims* dims, const void* rep, const arco_bank* bank, const double* proto_sums,
                            const int32_t* idx_anchor, const int32_t* idx_neg, float temp, float* loss,
                            float* grad_anchor, int32_t* anchor_pix, float* logits, void* workspace, void* stream) {
    return infonce_impl(dims, rep, bank, proto_sums, idx_anchor, idx_neg, temp, loss, grad_anchor, anchor_pix, logits,
                        nullptr, nullptr, 0.f, 1.f, nullptr, workspace, stream);
}

extern "C" int arco_infonce_ema(const arco_dims* dims, const void* rep, const arco_bank* bank, const double* proto_sums,
                                const int32_t* idx_anchor, const int32_t* idx_neg, float temp, float* loss,
                                float* grad_anchor, int32_t* anchor_pix, float* logits, const float* momentum,
                                const int32_t* momentum_on, float ema_decay, float ema_keep, float* proto_out,
                                void* workspace, void* stream) {
    ARCO_REQUIRE(momentum && momentum_on && proto_out, "arco_infonce_ema: NULL momentum argument");
    return infonce_impl(dims, rep, bank, proto_sums, idx_anchor, idx_neg, temp, loss, grad_anchor, anchor_pix, logits,
                        momentum, momentum_on, ema_decay, ema_keep, proto_out, workspace, stream);
}

// ---- sampled producers ------------------------------------------------------------------------------
extern "C" int arco_anchor_gather(const arco_dims* dims, const void* x, const int32_t* idx_anchor, float* rows,
                                  int32_t* anchor_pix, void* workspace, void* stream) {
    ARCO_REQUIRE(dims && x && idx_anchor && rows && anchor_pix && workspace, "arco_anchor_gather: NULL argument");
    const arco_dims& d = *dims;
    arco_ws_layout L;
    arco::compute_layout(d, &L);
    char* ws = (char*)workspace;
    arco::InfoParams p;
    memset(&p, 0, sizeof(p));
    p.rep = x; p.idx_a = idx_anchor;
    p.codes = (const uint8_t*)(ws + L.codes);
    p.off_anchor = (const uint32_t*)(ws + L.off_anchor);
    p.plan = (arco_plan*)(ws + L.plan);
    p.S = d.space; p.C = d.classes; p.D = d.feat; p.Q = d.queries; p.N = d.negatives;
    p.tpi = L.tiles_per_image; p.NT = L.n_tiles; p.rep_dtype = d.rep_dtype;
    arco::anchor_gather_kernel<<<d.classes * d.queries, 128, 0, (cudaStream_t)stream>>>(p, rows, anchor_pix);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_infonce_rows(const arco_dims* dims, const float* anchor_rows, const int32_t* anchor_pix_in, const arco_bank* bank,
                                 const double* proto_sums, const int32_t* idx_anchor, const int32_t* idx_neg, float temp,
                                 float* loss, float* grad_anchor, int32_t* anchor_pix, float* logits, void* workspace,
                                 void* stream) {
    ARCO_REQUIRE(anchor_rows && anchor_pix_in, "arco_infonce_rows: NULL anchor rows");
    return infonce_impl(dims, nullptr, bank, proto_sums, idx_anchor, idx_neg, temp, loss, grad_anchor, anchor_pix, logits,
                        nullptr, nullptr, 0.f, 1.f, nullptr, workspace, stream, anchor_rows, anchor_pix_in);
}

// ---- batch shards: InfoNCE with the exchange step of the path inside the launch -----------------------------
extern "C" int arco_infonce_sharded(const arco_dims* dims, const void* rep, const arco_bank* bank, const arco_exchange* exchange,
                                    int32_t gate_replanned, double* proto_sums, const int32_t* idx_anchor, const int32_t* idx_neg,
                                    float temp, float* loss, float* grad_anchor, int32_t* anchor_pix, float* logits,
                                    const float* momentum, const int32_t* momentum_on, float ema_decay, float ema_keep,
                                    float* proto_out, void* workspace, void* stream) {
    ARCO_REQUIRE((exchange != nullptr) != (gate_replanned != 0), "arco_infonce_sharded: pass the exchange descriptor OR gate_replanned");
    return infonce_impl(dims, rep, bank, proto_sums, idx_anchor, idx_neg, temp, loss, grad_anchor, anchor_pix, logits, momentum,
                        momentum_on, ema_decay, ema_keep, proto_out, workspace, stream, nullptr, nullptr, exchange, gate_replanned);
}
