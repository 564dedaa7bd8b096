// Multi-GPU exchange step of the path: the C x (D+1) fp64 buffer of per-class feature sums and
// counts is summed over the ranks -- 16 KB at the headline shape, so the cost is pure latency.  Instead of an NCCL
// all-reduce (launch + protocol, ~25 us) every rank's prototype kernel writes its sums straight into a buffer that is
// mapped into all peers (torch symmetric memory: NVLink peer mappings), and one 1-CTA kernel per rank
//   1. raises this step's sequence number in every peer's flag row (st.release.sys over NVLink),
//   2. waits until every peer has raised it here (ld.acquire.sys),
//   3. loads all W peer buffers (P2P loads) and adds them in rank order 0..W-1 -- the same order on every rank, so all
//      ranks hold bit-identical global sums.
// Buffers are double-buffered by sequence parity and the sequence only grows: a rank can be at most one step ahead of its
// slowest peer (it needs that peer's flag for the step in between), so the slot written at step k+2 is never still read.
#include "arco_common.cuh"

namespace arco {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// peer_base[r]: this allocation as mapped for rank r (r == rank: the local one).  Layout in doubles:
// [slot 0: n][slot 1: n][flags: 64 u64, one per source rank][step word: u64, present when seq carries ARCO_XCHG_STEP_WORD]
__global__ void __launch_bounds__(1024) proto_allreduce_p2p_kernel(const unsigned long long* __restrict__ peer_base, int rank, int world,
                                                                   unsigned long long seq, int n, int64_t n_pad, double* __restrict__ out,
                                                                   uint32_t* status) {
    const int tid = threadIdx.x;
    const int64_t flag_off = 2 * n_pad;
    if ((seq & ARCO_XCHG_STEP_WORD) && tid == 0)              // keep the buffer's step word current (see arco_exchange.seq)
        *(reinterpret_cast<unsigned long long*>(peer_base[rank]) + flag_off + 64) = seq & ARCO_XCHG_SEQ_MASK;
    seq &= ARCO_XCHG_SEQ_MASK;
    const int64_t slot = (int64_t)(seq & 1ull) * n_pad;
    __threadfence_system();                                   // the prototype kernel's sums (previous launch) before the flag
    if (tid < world && tid != rank)
        st_release_sys(reinterpret_cast<unsigned long long*>(peer_base[tid]) + flag_off + rank, seq);
    if (tid < world && tid != rank) {
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peer_base[rank]) + flag_off + tid;
        // a peer that never arrives must fail loudly, not hang: bound the wait by WALL time (10 s on %globaltimer; a
        // poll over NVLink costs ~1 us, so a poll count would be minutes), raise the status bit, then trap
        const unsigned long long t0 = globaltimer_ns();
        unsigned int polls = 0;
        while (ld_acquire_sys(mine) < seq) {
            if ((++polls & 1023u) == 0 && globaltimer_ns() - t0 > 10000000000ull) {
                if (status) { atomicOr(status, (uint32_t)ARCO_ST_EXCHANGE_TIMEOUT); __threadfence_system(); }
                __trap();
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        double acc = 0.0;
        for (int r = 0; r < world; ++r)
            acc += ld_relaxed_sys_f64(reinterpret_cast<const double*>(peer_base[r]) + slot + i);
        out[i] = acc;
    }
}

}  // namespace arco

extern "C" int arco_proto_allreduce_p2p(const arco_dims* dims, const uint64_t* peer_base_dev, int32_t rank, int32_t world, uint64_t seq,
                                        int64_t slot_doubles, double* proto_sums_out, void* workspace, void* stream) {
    ARCO_REQUIRE(dims && peer_base_dev && proto_sums_out && workspace, "arco_proto_allreduce_p2p: NULL argument");
    ARCO_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world && (seq & ARCO_XCHG_SEQ_MASK) > 0 &&
                     !(seq & ARCO_XCHG_SEQ_FROM_DEVICE), "arco_proto_allreduce_p2p: bad rank / world / sequence");
    const int n = dims->classes * (dims->feat + 1);
    ARCO_REQUIRE(slot_doubles >= n, "arco_proto_allreduce_p2p: slot smaller than C*(D+1)");
    arco_ws_layout L;
    arco::compute_layout(*dims, &L);
    arco_plan* plan = (arco_plan*)((char*)workspace + L.plan);
    arco::proto_allreduce_p2p_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>((const unsigned long long*)peer_base_dev, rank, world, seq, n,
                                                                          slot_doubles, proto_sums_out, &plan->status);
    ARCO_LAUNCH_CHECK();
    return ARCO_OK;
}
