// One-call forward of the whole path (single GPU, or one batch shard with the peer-memory exchange step): classify+scan+plan (one launch)
// -> [sampler on a side stream] -> prototype + enqueue + fp64 finalize (one launch) -> InfoNCE, plus the optional early zero fill of grad_rep on a second side stream.
// Same kernels and order as driving the stage entry points one by one (arco_b200/contra.py does that when an
// all-reduce or injected indices sit between the stages); this entry exists to keep the host cost of a step at one
// FFI call, which matters for the launch-bound small shapes (ACDC 256x256 D=64, LA 112x112x80 D=16).
#include "arco_common.cuh"

namespace arco {

struct SideStreams {
    cudaStream_t s[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, join_sample = nullptr, join_fill = nullptr;
    bool ok = false;
};

static SideStreams* side_streams() {
    static thread_local SideStreams per_dev[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    SideStreams& ss = per_dev[dev];
    if (!ss.ok) {
        if (cudaStreamCreateWithFlags(&ss.s[0], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithFlags(&ss.s[1], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&ss.join_sample, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&ss.join_fill, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        ss.ok = true;
    }
    return &ss;
}

}  // namespace arco

extern "C" int arco_forward(const arco_dims* dims, const arco_step_io* io, const arco_bank* bank, void* workspace,
                            void* stream) {
    ARCO_REQUIRE(dims && io && bank && workspace, "arco_forward: NULL argument");
    ARCO_REQUIRE(io->rep && io->rep_teacher && io->proto_sums && io->idx_anchor && io->idx_neg && io->loss &&
                     io->grad_anchor && io->anchor_pix, "arco_forward: NULL tensor in arco_step_io");
    cudaStream_t main_st = (cudaStream_t)stream;
    arco::SideStreams* ss = arco::side_streams();
    ARCO_REQUIRE(ss != nullptr, "arco_forward: could not create the helper streams");
    int rc;
    if (io->grad_prefill) {
        // input-independent P*D*e bytes of zeros: start now, underneath the read-bound forward kernels
        ARCO_CUDA_CHECK(cudaEventRecord(ss->fork, main_st));
        ARCO_CUDA_CHECK(cudaStreamWaitEvent(ss->s[1], ss->fork, 0));
        if ((rc = arco_grad_zero(dims, io->grad_prefill, ss->s[1])) != ARCO_OK) return rc;
        ARCO_CUDA_CHECK(cudaEventRecord(ss->join_fill, ss->s[1]));
    }
    // classify + ordered-compaction scans + plan in ONE launch (no memset: self-cleaning counters in bank->counters)
    if ((rc = arco_classify_plan(dims, io->label_l, io->label_u, io->prob_l, io->prob_u, io->low_mask, io->high_mask,
                                 io->delta_p, io->delta_n, io->low_rank, io->high_rank, bank, workspace, main_st)) != ARCO_OK)
        return rc;
    // the sampler only needs the plan: run it underneath the prototype pass
    ARCO_CUDA_CHECK(cudaEventRecord(ss->fork, main_st));
    ARCO_CUDA_CHECK(cudaStreamWaitEvent(ss->s[0], ss->fork, 0));
    if ((rc = arco_sample(dims, io->func, io->seed, io->step, io->idx_anchor, io->idx_neg, workspace, ss->s[0])) != ARCO_OK)
        return rc;
    ARCO_CUDA_CHECK(cudaEventRecord(ss->join_sample, ss->s[0]));
    const bool sharded = io->exchange_peers != nullptr;
    ARCO_REQUIRE(!sharded || io->exchange_local, "arco_forward: exchange_local is NULL");
    if ((rc = arco_proto_enqueue(dims, io->rep_teacher, bank, sharded ? io->exchange_local : io->proto_sums, workspace, main_st)) != ARCO_OK)
        return rc;
    ARCO_CUDA_CHECK(cudaStreamWaitEvent(main_st, ss->join_sample, 0));
    if (sharded) {
        // The one exchange step (global class sums, then the valid-class list from the GLOBAL counts) runs INSIDE the InfoNCE
        // launch, underneath its negatives pass on the rank-local plan (arco_infonce_sharded).  The two gated launches after
        // it only do work when the global counts changed the plan (a rank lacks a class another rank has).
        arco_exchange x;
        x.peers = (const uint64_t*)io->exchange_peers; x.seq = io->exchange_seq; x.slot_doubles = io->exchange_slot;
        x.rank = io->exchange_rank; x.world = io->exchange_world;
        if ((rc = arco_infonce_sharded(dims, io->rep, bank, &x, 0, (double*)io->proto_sums, io->idx_anchor, io->idx_neg, io->temp,
                                       io->loss, io->grad_anchor, io->anchor_pix, io->logits, io->momentum, io->momentum_on,
                                       io->ema_decay, io->ema_keep, io->proto_out, workspace, main_st)) != ARCO_OK)
            return rc;
        if ((rc = arco_sample_if_replanned(dims, io->func, io->seed, io->step, io->idx_anchor, io->idx_neg, workspace, main_st)) != ARCO_OK)
            return rc;
        rc = arco_infonce_sharded(dims, io->rep, bank, nullptr, 1, (double*)io->proto_sums, io->idx_anchor, io->idx_neg, io->temp,
                                  io->loss, io->grad_anchor, io->anchor_pix, io->logits, io->momentum, io->momentum_on,
                                  io->ema_decay, io->ema_keep, io->proto_out, workspace, main_st);
    } else if (io->momentum)
        rc = arco_infonce_ema(dims, io->rep, bank, io->proto_sums, io->idx_anchor, io->idx_neg, io->temp, io->loss,
                              io->grad_anchor, io->anchor_pix, io->logits, io->momentum, io->momentum_on, io->ema_decay,
                              io->ema_keep, io->proto_out, workspace, main_st);
    else
        rc = arco_infonce(dims, io->rep, bank, io->proto_sums, io->idx_anchor, io->idx_neg, io->temp, io->loss,
                          io->grad_anchor, io->anchor_pix, io->logits, workspace, main_st);
    if (rc != ARCO_OK) return rc;
    if (io->grad_prefill) ARCO_CUDA_CHECK(cudaStreamWaitEvent(main_st, ss->join_fill, 0));   // buffer is main-stream memory again
    return ARCO_OK;
}
