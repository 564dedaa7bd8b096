// One-call forward of the whole path (single GPU, or one batch shard with the peer-memory exchange step): classify+scan+plan (one launch)
// -> [sampler on a side stream] -> prototype + enqueue + fp64 finalize (one launch) -> InfoNCE, plus the optional early zero fill of grad_rep on a second side stream.
// Same kernels and order as driving the stage entry points one by one (arco_b200/contra.py does that when an
// all-reduce or injected indices sit between the stages); this entry exists to keep the host cost of a step at one
// FFI call, which matters for the launch-bound small shapes (ACDC 256x256 D=64, LA 112x112x80 D=16).
#include <stdlib.h>
#include <string.h>

#include "arco_common.cuh"

namespace arco {

struct SideStreams {
    cudaStream_t s[2] = {nullptr, nullptr};
    cudaStream_t capture = nullptr;          // origin stream of the replay cache's captures (the legacy default stream cannot capture)
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
        if (cudaStreamCreateWithFlags(&ss.capture, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&ss.join_sample, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&ss.join_fill, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        ss.ok = true;
    }
    return &ss;
}

}  // namespace arco

// The launches of one forward, issued one by one (also the body that gets captured, below).
static int forward_launches(const arco_dims* dims, const arco_step_io* io, const arco_bank* bank, void* workspace,
                            cudaStream_t main_st) {
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

// ---------------------------------------------------------------------------------------------------------------------
// Replay cache.  A forward is 6-7 launches on three streams with two fork/join event pairs; on the small shapes (ACDC
// 256x256 D=64: 0.14 ms of kernels, LA 112x112x80 D=16: 0.17 ms) issuing them costs the host as much as the GPU needs to
// run them (scripts/host_overhead.py: 0.18 ms of host time per step), so the eager step was host-bound.  Nothing in the
// launch parameters changes from step to step except the tensor addresses (the Philox stream and the host-mirror slot come
// from the bank's DEVICE step counter), and a training loop's allocator hands back the same few blocks every iteration.  So:
// the second time the exact same (dims, arco_step_io, arco_bank, workspace) tuple is seen, the launch sequence above is
// captured into a CUDA graph, instantiated once, and from then on a step is ONE cudaGraphLaunch.
// Any difference in any byte of the tuple is a different key (a miss runs the launches directly); a stream that is already
// being captured (torch.cuda.graph around the whole trainer step) always runs the launches directly.  The multi-GPU step
// replays too when its exchange buffer carries a step word (ARCO_XCHG_STEP_WORD): a replayed exchange takes its sequence
// number from that word instead of a launch parameter.  ARCO_FWD_GRAPH=0 switches the cache off.
// ---------------------------------------------------------------------------------------------------------------------
namespace arco {

struct FwdKey {
    arco_dims dims;
    arco_step_io io;
    arco_bank bank;
    void* workspace;
};

static int g_replay_on = -1;                                  // -1: take ARCO_FWD_GRAPH on first use; 2 = every size

static bool replay_enabled() {
    if (g_replay_on < 0) { const char* e = getenv("ARCO_FWD_GRAPH"); g_replay_on = !e ? 1 : e[0] == '0' ? 0 : e[0] == '2' ? 2 : 1; }
    return g_replay_on != 0;
}

// Replay pays where the host is the limit, i.e. where the kernels of a step take less time than issuing them (~0.18 ms).
// Measured on B200 (ms per forward + backward, direct launches -> replay): ACDC D=64 0.168 -> 0.151, LA 0.229 -> 0.202; but
// ACDC D=496 bf16 (1.6 GB of representations) 0.663 -> 0.677 and Cityscapes (4.3 GB) 2.039 -> 2.059: a graph launch has a
// fixed start-up cost that an HBM-bound step does not hide.  So only steps whose representation tensor is <= 512 MiB replay.
static bool replay_worthwhile(const arco_dims& d) {
    if (g_replay_on == 2) return true;
    const int64_t rep_bytes = (int64_t)(d.n_lab + d.n_unlab) * d.feat * d.space * (d.rep_dtype == ARCO_BF16 ? 2 : 4);
    return rep_bytes <= ((int64_t)512 << 20);
}

struct FwdCache {
    static constexpr int N_EXEC = 8, N_SEEN = 8;
    int64_t n_replay = 0, n_capture = 0, n_direct = 0;
    FwdKey* key[N_EXEC] = {};
    cudaGraphExec_t exec[N_EXEC] = {};
    uint64_t used[N_EXEC] = {};
    FwdKey* seen[N_SEEN] = {};
    int n_seen = 0;
    uint64_t tick = 0;
};

static FwdCache& fwd_cache(int dev) {
    static thread_local FwdCache caches[64];
    return caches[dev];
}

static FwdKey* make_key(const arco_dims* dims, const arco_step_io* io, const arco_bank* bank, void* workspace) {
    FwdKey* k = (FwdKey*)calloc(1, sizeof(FwdKey));           // zeroed: struct padding compares equal
    if (!k) return nullptr;
    memcpy(&k->dims, dims, sizeof(arco_dims));
    memcpy(&k->io, io, sizeof(arco_step_io));
    // multi-GPU: the sequence number is taken from the buffer's step word when the step is replayed (exchange_local already
    // tells the slot parity apart), so it is not part of the key
    if (k->io.exchange_peers) k->io.exchange_seq = 0;
    memcpy(&k->bank, bank, sizeof(arco_bank));
    k->workspace = workspace;
    return k;
}

}  // namespace arco

extern "C" int arco_forward(const arco_dims* dims, const arco_step_io* io, const arco_bank* bank, void* workspace,
                            void* stream) {
    ARCO_REQUIRE(dims && io && bank && workspace, "arco_forward: NULL argument");
    ARCO_REQUIRE(io->rep && io->rep_teacher && io->proto_sums && io->idx_anchor && io->idx_neg && io->loss &&
                     io->grad_anchor && io->anchor_pix, "arco_forward: NULL tensor in arco_step_io");
    cudaStream_t main_st = (cudaStream_t)stream;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        cudaGetLastError();
        return forward_launches(dims, io, bank, workspace, main_st);
    }
    arco::FwdCache& fc = arco::fwd_cache(dev);
    const bool sharded_ok = io->exchange_peers == nullptr || (io->exchange_seq & ARCO_XCHG_STEP_WORD);
    if (cudaStreamIsCapturing(main_st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
        cudaGetLastError();                                          // the caller is capturing: its step is not counted
        return forward_launches(dims, io, bank, workspace, main_st);
    }
    if (!arco::replay_enabled() || !arco::replay_worthwhile(*dims) || !sharded_ok) {
        ++fc.n_direct;
        return forward_launches(dims, io, bank, workspace, main_st);
    }
    arco::FwdKey* key = arco::make_key(dims, io, bank, workspace);
    if (!key) return forward_launches(dims, io, bank, workspace, main_st);
    ++fc.tick;
    for (int i = 0; i < arco::FwdCache::N_EXEC; ++i)
        if (fc.key[i] && memcmp(fc.key[i], key, sizeof(arco::FwdKey)) == 0) {
            free(key);
            fc.used[i] = fc.tick;
            ++fc.n_replay;
            ARCO_CUDA_CHECK(cudaGraphLaunch(fc.exec[i], main_st));
            return ARCO_OK;
        }
    bool seen = false;
    for (int i = 0; i < arco::FwdCache::N_SEEN; ++i)
        if (fc.seen[i] && memcmp(fc.seen[i], key, sizeof(arco::FwdKey)) == 0) seen = true;
    if (!seen) {
        // first sighting: remember the tuple, run the launches directly
        const int slot = fc.n_seen++ % arco::FwdCache::N_SEEN;
        free(fc.seen[slot]);
        fc.seen[slot] = key;
        ++fc.n_direct;
        return forward_launches(dims, io, bank, workspace, main_st);
    }
    // second sighting: capture (on a library-owned origin stream: the caller's may be the legacy default stream, which cannot
    // capture; the graph does not remember where it was recorded), instantiate, replay on the caller's stream
    arco::SideStreams* ss = arco::side_streams();
    if (ss == nullptr || cudaStreamBeginCapture(ss->capture, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
        cudaGetLastError();
        free(key);
        return forward_launches(dims, io, bank, workspace, main_st);
    }
    arco_step_io io_cap = *io;
    if (io->exchange_peers)                                          // replayed steps read their sequence number on the device
        io_cap.exchange_seq = ARCO_XCHG_STEP_WORD | ARCO_XCHG_SEQ_FROM_DEVICE | (io->exchange_seq & 1ull);
    const int rc = forward_launches(dims, &io_cap, bank, workspace, ss->capture);
    cudaGraph_t graph = nullptr;
    const cudaError_t e_end = cudaStreamEndCapture(ss->capture, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc != ARCO_OK || e_end != cudaSuccess || graph == nullptr || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        free(key);
        if (rc != ARCO_OK) return rc;                                // argument error: already reported by the stage
        return forward_launches(dims, io, bank, workspace, main_st); // capture not possible here: nothing ran yet, run it now
    }
    cudaGraphDestroy(graph);
    int victim = 0;
    for (int i = 0; i < arco::FwdCache::N_EXEC; ++i) {
        if (!fc.key[i]) { victim = i; break; }
        if (fc.used[i] < fc.used[victim]) victim = i;
    }
    if (fc.key[victim]) { cudaGraphExecDestroy(fc.exec[victim]); free(fc.key[victim]); }
    fc.key[victim] = key; fc.exec[victim] = exec; fc.used[victim] = fc.tick;
    ++fc.n_capture;
    ARCO_CUDA_CHECK(cudaGraphLaunch(exec, main_st));
    return ARCO_OK;
}

extern "C" int arco_forward_replay(int32_t on) {
    const int prev = arco::replay_enabled() ? 1 : 0;
    if (on >= 0) arco::g_replay_on = on > 2 ? 1 : on;
    return prev;
}

extern "C" int arco_forward_replay_stats(int64_t* stats) {
    ARCO_REQUIRE(stats != nullptr, "arco_forward_replay_stats: NULL argument");
    int dev = -1;
    ARCO_CUDA_CHECK(cudaGetDevice(&dev));
    ARCO_REQUIRE(dev >= 0 && dev < 64, "arco_forward_replay_stats: device index out of range");
    const arco::FwdCache& fc = arco::fwd_cache(dev);
    stats[0] = fc.n_replay; stats[1] = fc.n_capture; stats[2] = fc.n_direct;
    return ARCO_OK;
}
