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
    float ema_decay;
    int64_t row_off[ARCO_MAX_CLASSES];
    int32_t cap[ARCO_MAX_CLASSES];
    int64_t S;
    int32_t C, D, Q, N, tpi, NT;
    int32_t KC, RS16;        // keys per staged chunk, padded row stride in 16-byte chunks
    int32_t rep_dtype;
    float temp;
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

// MAXIT: 16-byte chunks per lane in pass 2 (ceil(chunks per row / 32)); BF16BANK: the ring stores bf16 rows
template <int MAXIT, bool BF16BANK>
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
    __shared__ bool s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bid = blockIdx.x;
    const int j = bid / p.Q;                                      // LOOP-2 position (loss_helper_3d.py:435)
    const int q = bid % p.Q;
    arco_plan* pl = p.plan;
    const bool active = pl->slot_active[j] != 0;

    if (active) {
        const int bank_cls = pl->valid_class[j];                  // trap 1: bank by CLASS ID ...
        const int blen = pl->bank_len[bank_cls];
        const int bhead = pl->bank_head[bank_cls];
        const int cap = p.cap[bank_cls];
        const uint32_t row_bytes = (uint32_t)D * (BF16BANK ? 2u : 4u);
        const unsigned char* bank = reinterpret_cast<const unsigned char*>(p.bank_rows) + p.row_off[bank_cls] * (int64_t)row_bytes;
        const float inv_scale = pl->inv_scale;

        if (tid < 4) mbar_init(&bars[tid], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");

        // ---- anchor rank-select: idx-th anchor candidate of class j (... anchors by POSITION j) ----
        if (warp == 0) {
            const uint32_t n_anchor = pl->n_anchor[j];
            uint32_t idx = (uint32_t)p.idx_a[(int64_t)j * p.Q + q];
            if (idx >= n_anchor) idx = n_anchor - 1;
            const uint32_t* off = p.off_anchor + (int64_t)j * (p.NT + 1);
            int lo = 0, hi = p.NT;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (off[mid] <= idx) lo = mid; else hi = mid;
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
                s_pix = (int)((int64_t)b * p.S + s0 + lane * 32 + (__ffs(m) - 1));
            }
        }
        __syncthreads();
        const int pix = s_pix;
        const int ab = (int)(pix / p.S);
        const int64_t as = pix - (int64_t)ab * p.S;

        // ---- anchor row (D strided loads, one 32-B sector each) and prototype row ----
        float n2a = 0.f, n2k = 0.f;
        const double cntj = p.proto_sums[(int64_t)j * (D + 1) + D];
        for (int d = tid; d < D; d += 128) {
            float v;
            if (p.rep_dtype == ARCO_BF16)
                v = bf16_bits_to_float(reinterpret_cast<const unsigned short*>(p.rep)[((int64_t)ab * D + d) * p.S + as]);
            else
                v = reinterpret_cast<const float*>(p.rep)[((int64_t)ab * D + d) * p.S + as];
            float k = (float)(p.proto_sums[(int64_t)j * (D + 1) + d] / cntj);   // class mean (:380-384)
            if (p.momentum) {
                // positive = (1-a)*proto + a*momentum_prototype[valid_classes[i]][q]  (:490-495); prototype[...] = positive (:497)
                const int64_t mo = ((int64_t)bank_cls * p.Q + q) * D + d;
                if (*p.momentum_on) k = (1.f - p.ema_decay) * k + p.ema_decay * p.momentum[mo];
                if (p.proto_out) p.proto_out[mo] = k;
            }
            a_hat[d] = v;
            k0hat[d] = k;
            n2a += v * v;
            n2k += k * k;
        }
        n2a = warp_sum(n2a);
        n2k = warp_sum(n2k);
        if (lane == 0) { s_red[warp][0] = n2a; s_red[warp][1] = n2k; }
        __syncthreads();
        const float na = sqrtf(s_red[0][0] + s_red[1][0] + s_red[2][0] + s_red[3][0]);
        const float nk0 = sqrtf(s_red[0][1] + s_red[1][1] + s_red[2][1] + s_red[3][1]);
        const float inv_na = 1.f / fmaxf(na, kEps), inv_nk0 = 1.f / fmaxf(nk0, kEps);
        float c0 = 0.f;
        for (int d = tid; d < D; d += 128) {
            const float a = a_hat[d] * inv_na, k = k0hat[d] * inv_nk0;
            a_hat[d] = a;
            k0hat[d] = k;
            c0 += a * k;
        }
        c0 = warp_sum(c0);
        __syncthreads();                       // everyone is done reading s_red
        if (lane == 0) s_red[warp][0] = c0;
        __syncthreads();
        const float cos0 = s_red[0][0] + s_red[1][0] + s_red[2][0] + s_red[3][0];
        const float inv_temp = 1.f / p.temp;
        const float z0 = cos0 * inv_temp;

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
                r = min(max(r, 0), blen - 1);
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

        // ---- merge the 4 partial softmaxes with the positive key, emit loss and d loss / d anchor ----
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
        const float inv_S = 1.f / S_all;
        const float sdot = (S2_all * inv_S - cos0) * inv_temp;
        for (int d = tid; d < D; d += 128) {
            const float g = gbuf[d] * f[0] + gbuf[D + d] * f[1] + gbuf[2 * D + d] * f[2] + gbuf[3 * D + d] * f[3] +
                            e0 * k0hat[d];
            const float gw = (g * inv_S - k0hat[d]) * inv_temp;
            const float grad = (na > kEps) ? (gw - sdot * a_hat[d]) / na : gw / kEps;
            p.g_anchor[(int64_t)bid * D + d] = grad * inv_scale;
        }
        if (tid == 0) {
            p.loss_parts[bid] = (__logf(S_all) + m_all - z0) * inv_scale;
            p.anchor_pix[bid] = pix;
            if (p.logits) p.logits[((int64_t)j * p.Q + q) * (1 + p.N)] = cos0;
        }
    } else if (tid == 0) {
        p.loss_parts[bid] = 0.f;
        p.anchor_pix[bid] = -1;
    }

    // ---- last CTA folds the per-query losses in a fixed order (deterministic) ----
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&pl->loss_done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const volatile float* lp = p.loss_parts;
    const int total = gridDim.x;
    float acc = 0.f;
    const int per = (total + 127) / 128;
    for (int i = tid * per; i < min(total, (tid + 1) * per); ++i) acc += lp[i];
    __shared__ float s_fold[128];
    s_fold[tid] = acc;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int i = 0; i < 128; ++i) s += s_fold[i];
        p.loss[0] = s;
    }
}

}  // namespace arco

static int infonce_impl(const arco_dims* dims, const void* rep, const arco_bank* bank, const double* proto_sums,
                        const int32_t* idx_anchor, const int32_t* idx_neg, float temp, float* loss,
                        float* grad_anchor, int32_t* anchor_pix, float* logits, const float* momentum,
                        const int32_t* momentum_on, float ema_decay, float* proto_out, void* workspace, void* stream) {
    ARCO_REQUIRE(dims && rep && bank && proto_sums && idx_anchor && idx_neg && loss && grad_anchor && anchor_pix &&
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
    p.loss_parts = (float*)(ws + L.loss_parts);
    p.momentum = momentum; p.momentum_on = momentum_on; p.proto_out = proto_out; p.ema_decay = ema_decay;
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
    int rs16 = cpl;
    if (kc >= 8) { while ((rs16 & 1) == 0) ++rs16; } else { while ((rs16 & 3) != 2) ++rs16; }
    p.KC = kc; p.RS16 = rs16;
    const size_t smem = (size_t)6 * d.feat * 4 + (size_t)4 * kc * rs16 * 16;
    const int grid = d.classes * d.queries;
    cudaStream_t st = (cudaStream_t)stream;
#define ARCO_INFONCE(MI, BF)                                                                                               \
    do {                                                                                                                   \
        ARCO_CUDA_CHECK(cudaFuncSetAttribute(arco::infonce_kernel<MI, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        arco::infonce_kernel<MI, BF><<<grid, 128, smem, st>>>(p);                                                           \
    } while (0)
    if (bf16bank) {
        if (cpl <= 32) ARCO_INFONCE(1, true); else ARCO_INFONCE(2, true);
    } else {
        if (cpl <= 32) ARCO_INFONCE(1, false); else if (cpl <= 64) ARCO_INFONCE(2, false); else ARCO_INFONCE(4, false);
    }
#undef ARCO_INFONCE
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}

extern "C" int arco_infonce(const arco_dims* dims, const void* rep, const arco_bank* bank, const double* proto_sums,
                            const int32_t* idx_anchor, const int32_t* idx_neg, float temp, float* loss,
                            float* grad_anchor, int32_t* anchor_pix, float* logits, void* workspace, void* stream) {
    return infonce_impl(dims, rep, bank, proto_sums, idx_anchor, idx_neg, temp, loss, grad_anchor, anchor_pix, logits,
                        nullptr, nullptr, 0.f, nullptr, workspace, stream);
}

extern "C" int arco_infonce_ema(const arco_dims* dims, const void* rep, const arco_bank* bank, const double* proto_sums,
                                const int32_t* idx_anchor, const int32_t* idx_neg, float temp, float* loss,
                                float* grad_anchor, int32_t* anchor_pix, float* logits, const float* momentum,
                                const int32_t* momentum_on, float ema_decay, float* proto_out, void* workspace,
                                void* stream) {
    ARCO_REQUIRE(momentum && momentum_on && proto_out, "arco_infonce_ema: NULL momentum argument");
    return infonce_impl(dims, rep, bank, proto_sums, idx_anchor, idx_neg, temp, loss, grad_anchor, anchor_pix, logits,
                        momentum, momentum_on, ema_decay, proto_out, workspace, stream);
}
