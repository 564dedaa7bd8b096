/*
 * arco_b200.h -- C ABI of libarco_b200.so: the B200 (sm_100a) implementation of ARCO's
 * stratified pixel/voxel contrastive loss.
 *
 * The reference has no FFI: its "plugin interface" is the Python function
 *   compute_contra_memobank_loss(rep, label_l, label_u, prob_l, prob_u, low_mask, high_mask,
 *                                memobank, queue_prtlis, queue_size, rep_teacher, ...)
 * (/root/reference/code/loss_helper_3d.py:271-290 for 2-D images, loss_helper.py:442-461 for 3-D
 * volumes), star-imported by train_arco_2d.py:24 / train_arco_3d.py:22 and called at
 * train_arco_2d.py:394-398 / train_arco_3d.py:356-360.  arco_b200/contra.py keeps that Python
 * signature and drives the entry points below through ctypes with raw device pointers.
 *
 * Conventions: every function returns 0 on success and a negative arco_status on failure, never
 * throws, never synchronises the device, and launches on the cudaStream_t passed as `stream`
 * (as void*).  All pointers are DEVICE pointers unless the name says `host_`.  Tensors are
 * contiguous; `rep`-like tensors are [B, D, S] (channel-first, S = flattened H*W or H*W*Z, the
 * layout the reference's trainers produce), probabilities [B_x, C, S] f32, masks [B, S] f32,
 * one-hot labels [B_x, C, S] int64 or plain integer label maps [B_x, S] int64.
 */
#ifndef ARCO_B200_H_
#define ARCO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ARCO_API __attribute__((visibility("default")))
#else
#define ARCO_API
#endif

#define ARCO_MAX_CLASSES 32
#define ARCO_TILE 1024            /* pixels per counting tile (never straddles an image) */

typedef enum arco_status {
    ARCO_OK = 0,
    ARCO_ERR_INVALID = -1,        /* bad argument (shape, alignment, NULL) */
    ARCO_ERR_CUDA = -2,           /* a CUDA runtime call failed; see arco_last_error_string() */
    ARCO_ERR_UNSUPPORTED = -3
} arco_status;

enum { ARCO_F32 = 0, ARCO_BF16 = 1 };
enum { ARCO_LABEL_ONEHOT_I64 = 0, ARCO_LABEL_INDEX_I64 = 1 };
/* sampler kinds: reference `func` argument, loss_helper_3d.py:327-338 */
enum { ARCO_FUNC_UNIFORM = 0, ARCO_FUNC_SMC = 1, ARCO_FUNC_ASMC = 2 };
/* device status bits (arco_plan.status) */
enum {
    ARCO_ST_MULTI_HOT = 1,        /* a one-hot label pixel has more than one non-zero class entry                 */
    ARCO_ST_LABEL_RANGE = 2,      /* an integer label map holds a class id >= classes                             */
    ARCO_ST_INDEX_RANGE = 4,      /* an injected / caller-provided sample index was outside its list (clamped)    */
    ARCO_ST_KEYS_DROPPED = 8,     /* keys were counted but the selected prototype kernel cannot enqueue them
                                     (register kernel for C <= 3: only legal with low_rank >= classes)            */
    ARCO_ST_EXCHANGE_DESYNC = 16, /* multi-GPU exchange: the device's step word and the slot the caller chose disagree
                                     (a replayed step ran without its prototype pass, or a launch was lost)           */
    ARCO_ST_EXCHANGE_TIMEOUT = (int)0x80000000u  /* multi-GPU exchange: a peer never raised its flag (10 s)       */
};

/* Problem geometry.  Mirrors the shapes at the reference call site. */
typedef struct arco_dims {
    int32_t n_lab;        /* labelled images, first in the batch (label_l.shape[0])       */
    int32_t n_unlab;      /* unlabelled images                                             */
    int32_t classes;      /* C  (label_l.shape[1]), <= ARCO_MAX_CLASSES                    */
    int32_t feat;         /* D  (rep.shape[1]), multiple of 4                              */
    int64_t space;        /* S  = prod(rep.shape[2:])                                      */
    int32_t queries;      /* Q  num_queries                                                */
    int32_t negatives;    /* N  num_negatives                                              */
    int32_t rep_dtype;    /* ARCO_F32 | ARCO_BF16 for rep, rep_teacher and grad_rep        */
    int32_t label_kind;   /* ARCO_LABEL_ONEHOT_I64 (reference call) | ARCO_LABEL_INDEX_I64 */
} arco_dims;

/* Byte offsets of the regions inside the caller-allocated workspace (256-B aligned base). */
typedef struct arco_ws_layout {
    int64_t total_bytes;
    int64_t plan;          /* arco_plan                                   */
    int64_t codes;         /* uint8  [B*S]   packed per-pixel class/flags */
    int64_t tile_flagged;  /* uint32 [NT]: bit g = 32-pixel group g of the tile holds a low-valid or key pixel */
    int64_t cnt_anchor;    /* uint32 [C][NT]                              */
    int64_t cnt_key;       /* uint32 [C][NT]                              */
    int64_t off_anchor;    /* uint32 [C][NT+1] exclusive scan             */
    int64_t off_key;       /* uint32 [C][NT+1]                            */
    int64_t partials;      /* float  [rows][C][D] prototype partial sums  */
    int64_t loss_parts;    /* float  [C*Q]                                */
    int64_t sample_scratch;/* int32  sampler staging                      */
    int32_t n_tiles;       /* NT = B * ceil(S / ARCO_TILE)                */
    int32_t tiles_per_image;
    int32_t partial_rows;
    int32_t reserved;
} arco_ws_layout;

/* Device-resident step summary written by arco_scan_plan (and read by every later stage).
 * The Python layer copies it to pinned host memory asynchronously to serve `new_keys`,
 * `memobank[c][0].shape[0]` and `queue_prtlis[c][0]` lazily (reference: loss_helper_3d.py:19-32,
 * :404-415). */
typedef struct arco_plan {
    uint32_t lv_count[ARCO_MAX_CLASSES];     /* low-valid pixels per class  (seg_num_list, :413-415) */
    uint32_t n_anchor[ARCO_MAX_CLASSES];     /* anchor candidates per class (len(seg_feat_low_entropy_list[c])) */
    uint32_t n_key[ARCO_MAX_CLASSES];        /* new_keys[c]                                           */
    int32_t  n_valid;                        /* valid_seg = len(valid_classes)                        */
    int32_t  valid_class[ARCO_MAX_CLASSES];  /* valid_classes[pos]                                    */
    int32_t  slot_active[ARCO_MAX_CLASSES];  /* LOOP-2 position pos runs (not skipped as "easy")      */
    int32_t  bank_write_base[ARCO_MAX_CLASSES]; /* ring position of this call's first key             */
    int32_t  bank_skip[ARCO_MAX_CLASSES];    /* leading keys dropped because n_key > capacity         */
    int32_t  bank_len[ARCO_MAX_CLASSES];     /* bank rows after this call                             */
    int32_t  bank_head[ARCO_MAX_CLASSES];    /* ring position of logical row 0 after this call        */
    int32_t  reserved0;                      /* padding (8-byte alignment of queue_ptr); 0 from the plan, set to 1 by the exchange block of arco_infonce_sharded */
    int64_t  queue_ptr[ARCO_MAX_CLASSES];    /* reference pointer bookkeeping (:24-30)                */
    float    inv_scale;                      /* 1 / (Q * valid_seg), 0 when valid_seg <= 1            */
    uint32_t status;                         /* ARCO_ST_* bits                                        */
    uint32_t scan_done;                      /* internal tickets                                      */
    uint32_t loss_done;
    uint32_t replanned;                      /* arco_replan_global changed valid_class / slot_active  */
    uint32_t proto_done;                     /* internal tickets of the prototype kernel's in-kernel finalize */
    uint32_t proto_done2;
    uint32_t step_ctr;                       /* value of the bank's device step counter when this step was planned   */
} arco_plan;

/* Device-resident ring-buffer memory bank (replaces the CPU list memobank[c] = [tensor[n,D]],
 * train_arco_2d.py:147-154).  Row r of class c lives at rows[(row_off[c] + (head[c]+r) % cap[c]) * D]. */
typedef struct arco_bank {
    void*    rows;                            /* [sum(cap), D] f32, or bf16 when row_dtype == ARCO_BF16 (bf16 rep only:
                                                 every enqueued key is then exactly representable)                   */
    int32_t* head;                            /* [C] device, updated by arco_scan_plan */
    int32_t* len;                             /* [C] device                            */
    int64_t* queue_ptr;                       /* [C] device                            */
    int32_t  cap[ARCO_MAX_CLASSES];           /* queue_size[c]                         */
    int64_t  row_off[ARCO_MAX_CLASSES];
    int32_t  row_dtype;                       /* ARCO_F32 | ARCO_BF16                  */
    int32_t  reserved;
    /* Optional zero-copy host mirror of the step summaries: NULL, or a DEVICE-ACCESSIBLE address of pinned host memory
       holding a ring of ARCO_MIRROR_SLOTS slots of ARCO_MIRROR_STRIDE bytes.  The last CTA of arco_infonce stores the final
       arco_plan into slot (seq % ARCO_MIRROR_SLOTS) and then seq as a uint64 at offset sizeof(arco_plan) of that slot
       (system-scope release), seq = 1, 2, 3, ... being the bank's DEVICE step counter (counters[ARCO_CTR_STEP], advanced by
       that same CTA).  The host follows new_keys, queue_prtlis and the status bits by polling memory: no memcpy, no event, no
       synchronisation -- and nothing in the launch parameters changes from step to step, so a whole step can be
       captured in a CUDA graph and replayed.  mirror_seq is unused (kept for layout stability). */
    void*    host_mirror;
    uint64_t mirror_seq;
    int64_t* host_queue_ptr;                  /* NULL, or device-accessible pinned int64[C]: the live queue_prtlis values */
    /* Device scratch of ARCO_COUNTER_WORDS uint32, ZERO when the bank is created and left zero by every step
       (self-cleaning tickets and accumulators of arco_classify_plan: no per-step memset).  Required by
       arco_classify_plan / arco_forward; the arco_classify_count + arco_scan_plan pair does not use it. */
    uint32_t* counters;
} arco_bank;
#define ARCO_COUNTER_WORDS 64
#define ARCO_CTR_STEP 40           /* counters[ARCO_CTR_STEP]: number of steps this bank has completed on the device */
#define ARCO_MIRROR_SLOTS 8
#define ARCO_MIRROR_STRIDE 1536

ARCO_API const char* arco_version(void);
ARCO_API const char* arco_last_error_string(void);

/* Workspace geometry for `dims` on the current device. */
ARCO_API int arco_workspace_layout(const arco_dims* dims, arco_ws_layout* out);

/* (a1) drop-in for the trainers' label_onehot (train_arco_2d.py:492-498): int64 labels [B,S] ->
 * float32 one-hot [B,C,S], ignore label -1 -> class 0. */
ARCO_API int arco_label_onehot(const int64_t* labels, float* out, int64_t batch, int32_t classes, int64_t space,
                      void* stream);

/* (a1-a3) fused one-hot decode, confidence thresholds and teacher-rank test -> per-pixel code byte
 * (bits 0-4 class, bit 5 low-valid, bit 6 anchor candidate, bit 7 negative key), per-tile per-class
 * anchor/key counts and per-class low-valid totals.  Replaces loss_helper_3d.py:341-342,352-374,
 * 388-401,413-415.  `label_*` are one-hot int64 [B_x,C,S] or index maps [B_x,S] per dims->label_kind. */
ARCO_API int arco_classify_count(const arco_dims* dims,
                        const int64_t* label_l, const int64_t* label_u,
                        const float* prob_l, const float* prob_u,
                        const float* low_mask, const float* high_mask,
                        float delta_p, float delta_n, int32_t low_rank, int32_t high_rank,
                        void* workspace, void* stream);

/* (a1-a4,a6) arco_classify_count + arco_scan_plan in ONE launch and without the memset of the plan: the CTAs that finish
 * last scan the per-tile counts (one class row each) and the very last one derives the plan with one warp; tickets and
 * the per-class low-valid accumulators live in bank->counters and are left zero for the next step. */
ARCO_API int arco_classify_plan(const arco_dims* dims,
                       const int64_t* label_l, const int64_t* label_u,
                       const float* prob_l, const float* prob_u,
                       const float* low_mask, const float* high_mask,
                       float delta_p, float delta_n, int32_t low_rank, int32_t high_rank,
                       const arco_bank* bank, void* workspace, void* stream);

/* (a4,a6) exclusive scans of the tile counts (ordered compaction offsets), valid-class list,
 * slot activity, ring-buffer bookkeeping (dequeue_and_enqueue, loss_helper_3d.py:12-32). */
ARCO_API int arco_scan_plan(const arco_dims* dims, const arco_bank* bank, void* workspace, void* stream);

/* Multi-GPU only: recompute valid_classes / slot activity from the all-reduced
 * proto_sums[C, D+1] counts so every rank runs the same LOOP-2 positions. */
ARCO_API int arco_replan_global(const arco_dims* dims, const double* proto_sums, void* workspace, void* stream);

/* (a5,a6) one pass over rep_teacher: per-class feature sums of low-valid pixels (prototype numerators,
 * :380-384) and ordered tail-only enqueue of the negative keys into the ring (:403-411).
 * proto_sums: float64 [C, D+1] = (sum over pixels, count) -- the buffer a multi-GPU caller all-reduces.  One launch: the
 * per-CTA partial rows are folded in fp64, in a fixed order, by the CTAs that finish last. */
ARCO_API int arco_proto_enqueue(const arco_dims* dims, const void* rep_teacher, const arco_bank* bank,
                       double* proto_sums, void* workspace, void* stream);

/* (a7) in-kernel Philox restatement of grid_monte_carlo_sample / grid_as_monte_carlo_sample and their
 * fallbacks (loss_helper_3d.py:35-268): idx_anchor int32 [C,Q], idx_neg int32 [C,Q*N], for every
 * active LOOP-2 position.  The Philox stream of a step is `step` + arco_plan.step_ctr (the bank's device step counter),
 * so replaying identical launch parameters (CUDA graph) still draws a fresh stream every step. */
ARCO_API int arco_sample(const arco_dims* dims, int32_t func, uint64_t seed, uint64_t step,
                int32_t* idx_anchor, int32_t* idx_neg, void* workspace, void* stream);

/* Stand-alone sampler: out[i] for i < shape drawn exactly like the reference sampler `func` called
 * with (high, shape).  Used by the drop-in sampler functions and the distribution tests. */
/* Multi-GPU: the sampler may run speculatively on the rank-local plan while the prototype pass and the all-reduce are in
   flight; after arco_replan_global this call redraws only if the global valid-class list changed the plan (device flag). */
/* Multi-GPU exchange step over NVLink peer memory instead of an NCCL all-reduce.  Every rank's
   arco_proto_enqueue writes its C*(D+1) fp64 sums into slot (seq & 1) of a buffer mapped into all peers (layout in doubles:
   [slot 0: slot_doubles][slot 1: slot_doubles][flags: one u64 per source rank, at least `world`]); this call signals the
   peers, waits for them and adds the W slots in rank order into proto_sums_out (bit-identical on every rank).
   peer_base_dev: device array of `world` addresses of that buffer as mapped for each rank; seq: 1, 2, 3, ... per call. */
ARCO_API int arco_proto_allreduce_p2p(const arco_dims* dims, const uint64_t* peer_base_dev, int32_t rank, int32_t world,
                                      uint64_t seq, int64_t slot_doubles, double* proto_sums_out, void* workspace,
                                      void* stream);
ARCO_API int arco_sample_if_replanned(const arco_dims* dims, int32_t func, uint64_t seed, uint64_t step, int32_t* idx_anchor,
                                      int32_t* idx_neg, void* workspace, void* stream);
ARCO_API int arco_sample_one(int32_t func, int64_t high, int64_t shape, uint64_t seed, uint64_t stream_id,
                    int32_t* out, void* scratch, int64_t scratch_bytes, void* stream);

/* (a8-a10) anchor rank-select + gather, memory-bank negative gather, cosine similarity, temperature
 * scaled InfoNCE and its gradient w.r.t. the anchor rows (loss_helper_3d.py:435-511).
 * Outputs: loss f32[1]; grad_anchor f32 [C,Q,D] (already scaled by 1/(Q*valid_seg));
 * anchor_pix int32 [C,Q] flat pixel id b*S+s or -1; logits f32 [C,Q,1+N] (optional, may be NULL). */
ARCO_API int arco_infonce(const arco_dims* dims, const void* rep, const arco_bank* bank, const double* proto_sums,
                 const int32_t* idx_anchor, const int32_t* idx_neg, float temp,
                 float* loss, float* grad_anchor, int32_t* anchor_pix, float* logits,
                 void* workspace, void* stream);

/* (a11) same with the optional EMA prototypes (momentum_prototype, loss_helper_3d.py:488-497): positive =
 * ema_keep*class_mean + ema_decay*momentum[bank_class][q] when *momentum_on != 0; the positive actually used is
 * written to proto_out[bank_class][q][:] (the reference's returned `prototype`).  momentum, proto_out: f32 [C,Q,D].
 * ema_keep = (float)(1.0 - (double)decay): the reference forms 1 - ema_decay as a Python double before it meets the
 * float32 tensor (:491-495); 1.f - (float)decay differs from that by 1.3e-5 relative at decay = 0.999. */
ARCO_API int arco_infonce_ema(const arco_dims* dims, const void* rep, const arco_bank* bank, const double* proto_sums,
                              const int32_t* idx_anchor, const int32_t* idx_neg, float temp,
                              float* loss, float* grad_anchor, int32_t* anchor_pix, float* logits,
                              const float* momentum, const int32_t* momentum_on, float ema_decay, float ema_keep,
                              float* proto_out, void* workspace, void* stream);

/* (a10) backward: grad_rep[B,D,S] = 0, then += grad_out * grad_anchor at the anchor pixels
 * (duplicates accumulate, trap 8).  grad_out: device f32 scalar. */
ARCO_API int arco_grad_scatter(const arco_dims* dims, const float* grad_anchor, const int32_t* anchor_pix,
                      const float* grad_out, void* grad_rep, void* stream);

/* The two halves of arco_grad_scatter, so that a caller can run the (input-independent) zero fill early on
 * a side stream, underneath the forward kernels, and only scatter in backward.  arco_grad_scatter_add expects the
 * zero-filled buffer: per pixel the duplicates of a position are summed in ascending query order in fp32 and STORED once
 * (deterministic; no atomics), so a non-zero value already at an anchor pixel would be overwritten, not added to. */
ARCO_API int arco_grad_zero(const arco_dims* dims, void* grad_rep, void* stream);
ARCO_API int arco_grad_scatter_add(const arco_dims* dims, const float* grad_anchor, const int32_t* anchor_pix,
                                   const float* grad_out, void* grad_rep, void* stream);

/* Opt-in sparse-gradient contract (not the reference's; arco_b200's `sparse_grad=True`): grad_rep is a buffer the CALLER
 * keeps alive across steps, all zero except at the pixels listed in prev_pix (int32 [C*Q], -1 = none; all -1 and an all-zero
 * buffer before the first call).  Clears those <= C*Q pixel columns, scatters this step's gradient (duplicates accumulate,
 * trap 8) and records this step's pixels in prev_pix: ~C*Q*D*(4+3e) bytes instead of the P*D*e-byte dense zero fill. */
ARCO_API int arco_grad_scatter_sparse(const arco_dims* dims, const float* grad_anchor, const int32_t* anchor_pix,
                                      const float* grad_out, void* grad_rep, int32_t* prev_pix, void* stream);

/* Every device pointer of one single-GPU forward step, for arco_forward. */
typedef struct arco_step_io {
    const void*    rep;            /* [B,D,S] student representation                                   */
    const void*    rep_teacher;    /* [B,D,S]                                                          */
    const int64_t* label_l;        /* per dims->label_kind                                             */
    const int64_t* label_u;
    const float*   prob_l;         /* [B_l,C,S]                                                        */
    const float*   prob_u;         /* [B_u,C,S]                                                        */
    const float*   low_mask;       /* [B,S]                                                            */
    const float*   high_mask;      /* [B,S]                                                            */
    double*        proto_sums;     /* out [C,D+1]                                                      */
    int32_t*       idx_anchor;     /* out [C,Q]                                                        */
    int32_t*       idx_neg;        /* out [C,Q*N]                                                      */
    float*         loss;           /* out [1]                                                          */
    float*         grad_anchor;    /* out [C,Q,D]                                                      */
    int32_t*       anchor_pix;     /* out [C,Q]                                                        */
    float*         logits;         /* out [C,Q,1+N] or NULL                                            */
    void*          grad_prefill;   /* NULL, or the grad_rep buffer to zero-fill on a side stream       */
    const float*   momentum;       /* NULL, or [C,Q,D] (a11)                                           */
    const int32_t* momentum_on;
    float*         proto_out;
    uint64_t       seed, step;     /* Philox seed / per-step stream id of the sampler                  */
    float          delta_p, delta_n, temp, ema_decay;
    int32_t        low_rank, high_rank, func;
    float          ema_keep;       /* (float)(1.0 - (double)ema_decay), see arco_infonce_ema                       */
    /* multi-GPU (batch shards); exchange_peers == NULL means single GPU.  See
       arco_proto_allreduce_p2p for the buffer layout. */
    const uint64_t* exchange_peers;   /* device array [exchange_world]: the exchange buffer as mapped for every rank  */
    double*        exchange_local;    /* this rank's slot (exchange_seq & 1) of it: the prototype kernel writes here   */
    uint64_t       exchange_seq;      /* 1, 2, 3, ... per call                                                         */
    int64_t        exchange_slot;     /* doubles per slot                                                              */
    int32_t        exchange_rank, exchange_world;
} arco_step_io;

/* The whole forward in one call (same kernels and order as the stage entry points; the sampler and the optional grad_rep
 * zero fill run on library-owned side streams and are joined with events, no host sync).  With exchange_peers set it is
 * the batch-sharded multi-GPU step: prototype sums -> arco_proto_allreduce_p2p -> arco_replan_global ->
 * arco_sample_if_replanned -> InfoNCE, still one call and no NCCL. */
ARCO_API int arco_forward(const arco_dims* dims, const arco_step_io* io, const arco_bank* bank, void* workspace,
                          void* stream);

/* Replay cache of arco_forward.  The second time a thread passes the byte-identical (dims, io, bank, workspace) tuple on a
 * device, arco_forward captures its launch sequence (6-7 kernels on three streams) into a CUDA graph
 * and from then on that tuple costs one cudaGraphLaunch; any other tuple, a stream that is already capturing and the
 * multi-GPU step run the launches directly.  Results are identical either way (same kernels, same parameters).
 * Only steps whose representation tensor is <= 512 MiB replay (the host-bound ones; an HBM-bound step does not hide a graph
 * launch's start-up cost), unless forced.  arco_forward_replay(1 / 0 / 2) switches the cache on / off / on for every size
 * for the process and returns whether it was on (-1 only queries; default 1, ARCO_FWD_GRAPH=0 / 2 in the environment).  arco_forward_replay_stats fills, for the calling
 * thread and current device, stats[0] = steps replayed, stats[1] = graphs captured, stats[2] = steps launched
 * directly (first sightings, cache off, too large, or an exchange buffer without a step word). */
ARCO_API int arco_forward_replay(int32_t on);
ARCO_API int arco_forward_replay_stats(int64_t* stats);

/* Parity / inspection helpers (not on the hot path). */
/* kind: 0 anchor candidates, 1 negative keys, 2 low-valid; writes the raster-ordered flat pixel ids of
 * class `cls` to out (capacity out_cap) and the count to *count_dev. */
ARCO_API int arco_export_list(const arco_dims* dims, int32_t kind, int32_t cls, int32_t* out, int64_t out_cap,
                     uint32_t* count_dev, void* workspace, void* stream);
/* Copy class `cls` of the ring to `out` in logical FIFO order ([cap, D], rows >= len untouched). */
ARCO_API int arco_bank_read(const arco_bank* bank, int32_t cls, int32_t feat, float* out, void* stream);


/* ---- config 5: dense tensor-core similarity, forward only ----------------------------------
   Replaces, for the crossover study, the paired gather of loss_helper_3d.py:466-486 (negative_feat = bank[idx]; cosine of
   every query against its own N rows) by  S = A_hat [Q, D] x Ring^T [D, cap]  on tcgen05 (anchor rows split into three bf16
   terms, ring rows bf16: exact products, fp32 accumulation) and a scalar gather  logits[q][n] = S[q][row(idx[q][n])] / |k_row|.
   anchors   f32 device [n_slots][Q][D]  raw (un-normalised) anchor rows
   slot_class int32 HOST [n_slots]       ring class each slot is contrasted against (valid_classes[i], trap 1)
   idx_neg   int32 device [n_slots][Q][N] logical ring rows (as arco_sample writes them)
   logits    f32 device [n_slots][Q][N]  cosines (not divided by temp)
   scratch   device, arco_similarity_dense_scratch() bytes.  Needs a bf16 ring, D % 8 == 0, Q % 128 == 0.          */
ARCO_API int64_t arco_similarity_dense_scratch(int32_t feat, int32_t queries, int32_t n_slots, const arco_bank* bank,
                                               const int32_t* slot_class);
ARCO_API int arco_similarity_dense(int32_t feat, int32_t queries, int32_t negatives, int32_t n_slots,
                                   const int32_t* slot_class, const float* anchors, const arco_bank* bank,
                                   const int32_t* idx_neg, float* logits, void* scratch, void* stream);
/* Backward of the dense form: grad_anchor_hat[slot][q][:] = sum_n grad_logits[slot][q][n] * k_hat[row(idx_neg[slot][q][n])]
   (the gradient with respect to the UNIT anchors; duplicates of a ring row accumulate) as a second tcgen05 GEMM
   [Q, M] x [M, D]: scattered weights split into three bf16 terms against the transposed ring. */
ARCO_API int64_t arco_similarity_dense_backward_scratch(int32_t feat, int32_t queries, int32_t n_slots, const arco_bank* bank,
                                                        const int32_t* slot_class);
ARCO_API int arco_similarity_dense_backward(int32_t feat, int32_t queries, int32_t negatives, int32_t n_slots,
                                            const int32_t* slot_class, const float* grad_logits, const arco_bank* bank,
                                            const int32_t* idx_neg, float* grad_anchor_hat, void* scratch, void* stream);

/* ---- mask / threshold preparation upstream of the loss -------------------------------
   Replaces train_arco_2d.py:345-393 (train_arco_3d.py:315-353): teacher softmax, student entropy, the two np.percentile
   thresholds (a GPU -> CPU -> GPU round trip each in the reference) and the low / high entropy masks.                  */
/* softmax over the class axis of logits [batch, classes, space] (f32); prob (same shape) and/or
   entropy[batch, space] = -sum_c p*log(p + 1e-10) may be NULL.  (:353-359)                                             */
ARCO_API int arco_softmax_rows(const float* logits, int64_t batch, int32_t classes, int64_t space, float* prob,
                               float* entropy, void* stream);
ARCO_API int64_t arco_entropy_masks_scratch(void);
/* entropy f32 [n_unlab_px] (student, unlabelled images), label_l / label_u int64 label maps (ignore = negative).
   q_low / q_high = float32(percent) / float32(100) exactly as numpy 2.x forms them for float32 data.
   low_mask / high_mask f32 [n_lab_px + n_unlab_px] (labelled part: label >= 0; unlabelled part: entropy <= / >= the
   numpy-"linear" percentile of the valid entropies, times label >= 0)  (:360-392); thresholds: optional device float[2]. */
ARCO_API int arco_entropy_masks(const float* entropy, const int64_t* label_l, const int64_t* label_u, int64_t n_lab_px,
                                int64_t n_unlab_px, float q_low, float q_high, float* low_mask, float* high_mask,
                                float* thresholds, void* scratch, void* stream);
/* The whole block in one call: teacher probabilities of both halves, student entropy of the unlabelled half, masks.
   logits f32 [n, classes, space]; labels int64 [n, space]; outputs as in the two functions above.                       */
ARCO_API int arco_prepare_contrast(const float* pred_u, const float* pred_l_teacher, const float* pred_u_teacher,
                                   const int64_t* label_l, const int64_t* label_u, int64_t n_lab, int64_t n_unlab,
                                   int32_t classes, int64_t space, float q_low, float q_high, float* prob_l_teacher,
                                   float* prob_u_teacher, float* entropy, float* low_mask, float* high_mask,
                                   float* thresholds, void* scratch, void* stream);

/* ---- nearest-neighbour "revisiting" loss + random-pool queue ----------------------------
   Replaces get_revisiting_loss (train_arco_2d.py:126-136) and the pool enqueue (:400-402 with _dequeue_and_enqueue :109-120).
   rep_u / rep_u_teacher: [bs, length] (the [bs, D, H, W] tensors flattened; f32 or bf16 per rep_dtype), pool: f32
   [pool_rows, length] unit rows.  ONE streaming pass yields every dot product and squared norm; the normalised copies the
   reference materialises are never written.
   loss f32[1]; nn_index int32 [bs, topk] (the top-k smallest STUDENT distances, ties: lower row first);
   stats f32 [2*bs*pool_rows dots (student rows, then teacher rows) | 2*bs squared norms] -- input of arco_revisit_enqueue.
   Limits: ceil(2*bs/12) * ceil(pool_rows/9) <= 8 (the trainers: bs 12, K 36), pool_rows <= 256, 16-byte aligned rows. */
ARCO_API int64_t arco_revisit_scratch_bytes(int32_t bs, int32_t pool_rows);
ARCO_API int arco_revisit_loss(const void* rep_u, const void* rep_u_teacher, const float* pool, int32_t bs, int32_t pool_rows,
                               int64_t length, int32_t rep_dtype, int32_t topk, float* loss, int32_t* nn_index, float* stats,
                               void* scratch, void* stream);
/* pool[(pool_ptr + b) % pool_rows] = rep_u_teacher[b] / max(|rep_u_teacher[b]|, 1e-12) with the norms of `stats`. */
ARCO_API int arco_revisit_enqueue(const void* rep_u_teacher, const float* stats, float* pool, int64_t pool_ptr, int32_t bs,
                                  int32_t pool_rows, int64_t length, int32_t rep_dtype, void* stream);

/* ---- the other per-pixel loss terms of the 2-D trainer's step ---------------------------
   All tensors f32 unless noted; [B, C, S] channel-first like the trainer's predictions; scratch: arco_step_scratch_bytes(). */
ARCO_API int64_t arco_step_scratch_bytes(int32_t batch, int64_t space);
/* compute_unsupervised_loss (train_arco_2d.py:482-489): target int64 [B,S] (ignore -1), logits [B,S] confidences.
   stats out: [B] weighting (= #{logits >= thr} / #{valid}), [1] 1/N (N = #{loss > 0}), [1] the loss. */
ARCO_API int arco_unsup_loss(const float* predict, const int64_t* target, const float* logits, float strong_threshold, int32_t batch,
                             int32_t classes, int64_t space, float* stats, void* scratch, void* stream);
ARCO_API int arco_unsup_loss_backward(const float* predict, const int64_t* target, const float* stats, const float* grad_out,
                                      int32_t batch, int32_t classes, int64_t space, float* grad_predict, void* stream);
/* TPSGridGen.forward (tps_stn_pytorch/tps_grid_gen.py:54-75) for RandTPS (tps/rand_tps.py:82-153): grid f32 [B,H,W,2] from
   mapping [B, n_points+3, 2] (= inverse_kernel @ [source control points; 0], 28 x 28, host side) and the n_points target
   control points [n_points, 2]. */
ARCO_API int arco_tps_grid(const float* mapping, const float* control_points, int32_t n_points, int32_t batch, int32_t height,
                           int32_t width, float* grid, void* stream);
/* tps(x) = F.grid_sample(x, grid, bilinear, align_corners=True, zeros | border padding) (tps/grid_sample.py:11-12); forward only. */
ARCO_API int arco_grid_sample(const float* input, const float* grid, int32_t batch, int32_t channels, int32_t height, int32_t width,
                              int32_t border_padding, float* out, void* stream);
/* Equivariance loss (train_arco_2d.py:404-423) in one pass: mask from labels int64 [B,H,W] / logits [B,H,W], its warp, the warp of
   pred_detached [B,C,H,W], both softmaxes, the masked KL.  stats out: [B] 1/(B*(sum mask_tps + 1e-7)), [1] the loss.
   grad_unscaled: NULL, or [B,C,H,W] receiving mask_tps * (softmax(pred_tps) - softmax(warped)); the gradient w.r.t. pred_tps is
   arco_scale_rows(grad_unscaled, stats, grad_out). */
ARCO_API int arco_eqv_loss(const float* pred_tps, const float* pred_detached, const float* grid, const int64_t* labels,
                           const float* logits, float weak_threshold, int32_t batch, int32_t classes, int32_t height, int32_t width,
                           float* stats, float* grad_unscaled, void* scratch, void* stream);
ARCO_API int arco_scale_rows(const float* g, const float* scale, const float* grad_out, int32_t batch, int64_t per_image, float* out,
                             void* stream);

/* ---- the representation producers, folded into the loss -------------------------------
   Replaces the bias-free 1x1 convolutions that PRODUCE the loss's two big operands (model_2D.py:49-53 `fea4` of the teacher's and
   the student's FeatureExtractor, train_arco_2d.py:231-234 `q_representation`, applied at train_arco_2d.py:317-329) by
   applying their weights only where the loss consumes a value: a 1x1 convolution is linear and per-pixel, so
   sum_px(W x) = W sum_px(x) and (W x)[pixel] = W (x[pixel]).  The caller runs arco_classify_plan and arco_proto_enqueue on
   the convolutions' INPUT tensors (same [B, D, S] layout, square weights [D_out = D][D_in = D] row-major as
   nn.Conv2d.weight[:, :, 0, 0]) and then: */
/* The ring rows enqueued by this step (found from the device plan in `workspace`) <- weight . row, in place, on tcgen05
   (bf16 ring: kind::f16; fp32 ring: kind::tf32, three-term split, needs arco_keys_transform_scratch_bytes() of scratch).
   `weight` must have the ring's row dtype. */
ARCO_API int64_t arco_keys_transform_scratch_bytes(int32_t feat, int32_t row_dtype);
ARCO_API int arco_keys_transform(const arco_dims* dims, const arco_bank* bank, const void* weight, void* scratch,
                                 void* workspace, void* stream);
/* sums_out[c][0..D) = weight . sums_in[c][0..D) in fp64, sums_out[c][D] = sums_in[c][D] (the prototype buffer of
   arco_proto_enqueue; multi-GPU: apply after the exchange).  weight_dtype ARCO_F32 | ARCO_BF16. */
ARCO_API int arco_proto_transform(int32_t classes, int32_t feat, const void* weight, int32_t weight_dtype,
                                  const double* sums_in, double* sums_out, void* stream);
/* The step's anchors as rows: rows[j*Q+q][0..D) = x[:, :, pixel of the idx_anchor[j][q]-th anchor candidate of position j]
   (fp32), anchor_pix[j*Q+q] = that pixel (-1 and a zero row for inactive positions).  x is [B, D, S] in dims->rep_dtype. */
ARCO_API int arco_anchor_gather(const arco_dims* dims, const void* x, const int32_t* idx_anchor, float* rows,
                                int32_t* anchor_pix, void* workspace, void* stream);
/* arco_infonce with the anchors given as dense fp32 rows [C*Q][D] (what the student's convolutions made of the gathered rows)
   and their pixels, instead of being gathered from a [B, D, S] tensor.  grad_anchor is d loss / d anchor_rows. */
ARCO_API int arco_infonce_rows(const arco_dims* dims, const float* anchor_rows, const int32_t* anchor_pix_in, const arco_bank* bank,
                               const double* proto_sums, const int32_t* idx_anchor, const int32_t* idx_neg, float temp,
                               float* loss, float* grad_anchor, int32_t* anchor_pix, float* logits, void* workspace,
                               void* stream);

/* ---- batch shards: the exchange step INSIDE the InfoNCE launch ---------------------------------------------------------------
   arco_proto_allreduce_p2p + arco_replan_global put the one exchange of the multi-GPU path between the
   prototype pass and InfoNCE: +30..43 us per step of launch gaps, NVLink round trips and rank skew.  The negatives pass of
   InfoNCE needs neither the prototype nor the global counts, so arco_infonce_sharded(exchange != NULL) runs it on the
   rank-local plan while an extra block 0 trades the class sums with the peers (same protocol and summation order as
   arco_proto_allreduce_p2p), re-derives the valid-class list and publishes both before the merge.  If the plan changed under
   the speculation the launch emits nothing; the caller always enqueues  arco_sample_if_replanned  and a second
   arco_infonce_sharded(exchange = NULL, gate_replanned = 1), both of which return at once unless plan->replanned. */
#define ARCO_XCHG_STEP_WORD       (1ull << 63)
#define ARCO_XCHG_SEQ_FROM_DEVICE (1ull << 62)
#define ARCO_XCHG_SEQ_MASK        ((1ull << 62) - 1)
typedef struct arco_exchange {
    const uint64_t* peers;        /* device array [world]: this rank's exchange buffer as mapped for every peer              */
    uint64_t        seq;          /* step sequence number (> 0, grows by one per step) in bits 0-61.
                                     Bit 63 (ARCO_XCHG_STEP_WORD): the buffer carries a u64 step word behind its 64 flags; every
                                     exchange stores its sequence number there.  Bit 62 (ARCO_XCHG_SEQ_FROM_DEVICE, needs bit 63):
                                     the sequence number is that word + 1 and bits 0-61 only give its expected parity -- this
                                     is what lets arco_forward replay a multi-GPU step as a CUDA graph (no launch parameter
                                     changes from step to step); a parity mismatch sets ARCO_ST_EXCHANGE_DESYNC and traps.    */
    int64_t         slot_doubles; /* doubles per slot; layout [slot 0][slot 1][flags: 64 u64, one per source rank][step word]   */
    int32_t         rank, world;
} arco_exchange;
ARCO_API int arco_infonce_sharded(const arco_dims* dims, const void* rep, const arco_bank* bank, const arco_exchange* exchange,
                                  int32_t gate_replanned, double* proto_sums, const int32_t* idx_anchor, const int32_t* idx_neg,
                                  float temp, float* loss, float* grad_anchor, int32_t* anchor_pix, float* logits,
                                  const float* momentum, const int32_t* momentum_on, float ema_decay, float ema_keep,
                                  float* proto_out, void* workspace, void* stream);

/* ---- logits-in loss (SURVEY.md section 8(f) rank 1 rationale): classify straight from the trainers' raw tensors ------------------
   arco_prepare_contrast materialises teacher probabilities [B,C,S] and two masks [B,1,S] only for arco_classify_plan to read them
   back.  Here the teacher softmax and both masks are formed in the classify kernel's registers (same operations in the same
   order: results are bit-identical to the two-call path): per pixel it reads the C teacher logits, the int64 label (ignore -1)
   and, on unlabelled images, the student entropy; nothing intermediate is written.  2 <= C <= 8, integer label maps
   (dims->label_kind = ARCO_LABEL_INDEX_I64), S % 4 == 0.
   arco_entropy_thresholds: the two np.percentile thresholds (train_arco_2d.py:360-369) of the valid unlabelled entropies ->
   thresholds[2] on the device (3-level radix select; scratch: arco_entropy_masks_scratch()). */
ARCO_API int arco_entropy_thresholds(const float* entropy, const int64_t* label_u, int64_t n_unlab_px, float q_low, float q_high,
                                     float* thresholds, void* scratch, void* stream);
ARCO_API int arco_classify_plan_logits(const arco_dims* dims, const int64_t* label_l, const int64_t* label_u,
                                       const float* logits_l_teacher, const float* logits_u_teacher, const float* entropy_u,
                                       const float* thresholds, float delta_p, float delta_n, int32_t low_rank,
                                       int32_t high_rank, const arco_bank* bank, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARCO_B200_H_ */
