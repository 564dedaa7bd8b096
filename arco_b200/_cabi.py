"""ctypes binding of ``libarco_b200.so`` (declared in ``include/arco_b200.h``).

The product path has no CPU or PyTorch fallback: if the shared library is missing, importing this
module raises with the build command, and every entry point raises :class:`ArcoError` on a
non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_CLASSES = 32
COUNTER_WORDS = 64
CTR_STEP = 40
MIRROR_SLOTS, MIRROR_STRIDE = 8, 1536
TILE = 1024
F32, BF16 = 0, 1
LABEL_ONEHOT_I64, LABEL_INDEX_I64 = 0, 1
FUNC_UNIFORM, FUNC_SMC, FUNC_ASMC = 0, 1, 2
ST_MULTI_HOT, ST_LABEL_RANGE, ST_INDEX_RANGE, ST_KEYS_DROPPED, ST_EXCHANGE_DESYNC, ST_EXCHANGE_TIMEOUT = 1, 2, 4, 8, 16, 0x80000000

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ARCO_B200_LIB") or os.path.join(_HERE, "lib", "libarco_b200.so")   # override: A/B builds


class ArcoError(RuntimeError):
    pass


class Dims(C.Structure):
    _fields_ = [
        ("n_lab", C.c_int32), ("n_unlab", C.c_int32), ("classes", C.c_int32), ("feat", C.c_int32),
        ("space", C.c_int64), ("queries", C.c_int32), ("negatives", C.c_int32),
        ("rep_dtype", C.c_int32), ("label_kind", C.c_int32),
    ]


class WsLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "total_bytes", "plan", "codes", "tile_flagged", "cnt_anchor", "cnt_key", "off_anchor", "off_key",
        "partials", "loss_parts", "sample_scratch")] + [
        ("n_tiles", C.c_int32), ("tiles_per_image", C.c_int32), ("partial_rows", C.c_int32), ("reserved", C.c_int32)]


_U32x = C.c_uint32 * MAX_CLASSES
_I32x = C.c_int32 * MAX_CLASSES
_I64x = C.c_int64 * MAX_CLASSES


class Plan(C.Structure):
    _fields_ = [
        ("lv_count", _U32x), ("n_anchor", _U32x), ("n_key", _U32x),
        ("n_valid", C.c_int32),
        ("valid_class", _I32x), ("slot_active", _I32x), ("bank_write_base", _I32x), ("bank_skip", _I32x),
        ("bank_len", _I32x), ("bank_head", _I32x), ("reserved0", C.c_int32),
        ("queue_ptr", _I64x),
        ("inv_scale", C.c_float), ("status", C.c_uint32), ("scan_done", C.c_uint32), ("loss_done", C.c_uint32),
        ("replanned", C.c_uint32), ("proto_done", C.c_uint32), ("proto_done2", C.c_uint32), ("step_ctr", C.c_uint32),
    ]


class Bank(C.Structure):
    _fields_ = [
        ("rows", C.c_void_p), ("head", C.c_void_p), ("len", C.c_void_p), ("queue_ptr", C.c_void_p),
        ("cap", _I32x), ("row_off", _I64x), ("row_dtype", C.c_int32), ("reserved", C.c_int32),
        ("host_mirror", C.c_void_p), ("mirror_seq", C.c_uint64), ("host_queue_ptr", C.c_void_p), ("counters", C.c_void_p),
    ]


class Exchange(C.Structure):
    _fields_ = [("peers", C.c_void_p), ("seq", C.c_uint64), ("slot_doubles", C.c_int64), ("rank", C.c_int32), ("world", C.c_int32)]


class StepIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "rep", "rep_teacher", "label_l", "label_u", "prob_l", "prob_u", "low_mask", "high_mask", "proto_sums",
        "idx_anchor", "idx_neg", "loss", "grad_anchor", "anchor_pix", "logits", "grad_prefill", "momentum",
        "momentum_on", "proto_out")] + [
        ("seed", C.c_uint64), ("step", C.c_uint64),
        ("delta_p", C.c_float), ("delta_n", C.c_float), ("temp", C.c_float), ("ema_decay", C.c_float),
        ("low_rank", C.c_int32), ("high_rank", C.c_int32), ("func", C.c_int32), ("ema_keep", C.c_float),
        ("exchange_peers", C.c_void_p), ("exchange_local", C.c_void_p), ("exchange_seq", C.c_uint64),
        ("exchange_slot", C.c_int64), ("exchange_rank", C.c_int32), ("exchange_world", C.c_int32)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. arco_b200 has no CPU/PyTorch fallback: build the CUDA library first with "
            f"`python -m arco_b200.build` (needs nvcc, targets sm_100a).")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
    dp, bp = C.POINTER(Dims), C.POINTER(Bank)
    sigs = {
        "arco_version": (C.c_char_p, []),
        "arco_last_error_string": (C.c_char_p, []),
        "arco_workspace_layout": (C.c_int, [dp, C.POINTER(WsLayout)]),
        "arco_label_onehot": (C.c_int, [vp, vp, i64, i32, i64, vp]),
        "arco_classify_count": (C.c_int, [dp, vp, vp, vp, vp, vp, vp, f32, f32, i32, i32, vp, vp]),
        "arco_classify_plan": (C.c_int, [dp, vp, vp, vp, vp, vp, vp, f32, f32, i32, i32, bp, vp, vp]),
        "arco_scan_plan": (C.c_int, [dp, bp, vp, vp]),
        "arco_replan_global": (C.c_int, [dp, vp, vp, vp]),
        "arco_proto_enqueue": (C.c_int, [dp, vp, bp, vp, vp, vp]),
        "arco_sample": (C.c_int, [dp, i32, u64, u64, vp, vp, vp, vp]),
        "arco_proto_allreduce_p2p": (C.c_int, [dp, vp, i32, i32, u64, i64, vp, vp, vp]),
        "arco_sample_if_replanned": (C.c_int, [dp, i32, u64, u64, vp, vp, vp, vp]),
        "arco_sample_one": (C.c_int, [i32, i64, i64, u64, u64, vp, vp, i64, vp]),
        "arco_infonce": (C.c_int, [dp, vp, bp, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp]),
        "arco_infonce_ema": (C.c_int, [dp, vp, bp, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp]),
        "arco_grad_scatter": (C.c_int, [dp, vp, vp, vp, vp, vp]),
        "arco_grad_zero": (C.c_int, [dp, vp, vp]),
        "arco_grad_scatter_add": (C.c_int, [dp, vp, vp, vp, vp, vp]),
        "arco_grad_scatter_sparse": (C.c_int, [dp, vp, vp, vp, vp, vp, vp]),
        "arco_forward": (C.c_int, [dp, C.POINTER(StepIO), bp, vp, vp]),
        "arco_forward_replay": (C.c_int, [i32]),
        "arco_forward_replay_stats": (C.c_int, [vp]),
        "arco_export_list": (C.c_int, [dp, i32, i32, vp, i64, vp, vp, vp]),
        "arco_bank_read": (C.c_int, [bp, i32, i32, vp, vp]),
        "arco_softmax_rows": (C.c_int, [vp, i64, i32, i64, vp, vp, vp]),
        "arco_entropy_masks_scratch": (C.c_int64, []),
        "arco_entropy_masks": (C.c_int, [vp, vp, vp, i64, i64, f32, f32, vp, vp, vp, vp, vp]),
        "arco_prepare_contrast": (C.c_int, [vp, vp, vp, vp, vp, i64, i64, i32, i64, f32, f32, vp, vp, vp, vp, vp, vp, vp, vp]),
        "arco_revisit_scratch_bytes": (C.c_int64, [i32, i32]),
        "arco_revisit_loss": (C.c_int, [vp, vp, vp, i32, i32, i64, i32, i32, vp, vp, vp, vp, vp]),
        "arco_revisit_enqueue": (C.c_int, [vp, vp, vp, i64, i32, i32, i64, i32, vp]),
        "arco_step_scratch_bytes": (C.c_int64, [i32, i64]),
        "arco_unsup_loss": (C.c_int, [vp, vp, vp, f32, i32, i32, i64, vp, vp, vp]),
        "arco_unsup_loss_backward": (C.c_int, [vp, vp, vp, vp, i32, i32, i64, vp, vp]),
        "arco_tps_grid": (C.c_int, [vp, vp, i32, i32, i32, i32, vp, vp]),
        "arco_grid_sample": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, vp, vp]),
        "arco_eqv_loss": (C.c_int, [vp, vp, vp, vp, vp, f32, i32, i32, i32, i32, vp, vp, vp, vp]),
        "arco_scale_rows": (C.c_int, [vp, vp, vp, i32, i64, vp, vp]),
        "arco_entropy_thresholds": (C.c_int, [vp, vp, i64, f32, f32, vp, vp, vp]),
        "arco_classify_plan_logits": (C.c_int, [dp, vp, vp, vp, vp, vp, vp, f32, f32, i32, i32, bp, vp, vp]),
        "arco_infonce_sharded": (C.c_int, [dp, vp, bp, C.POINTER(Exchange), i32, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp]),
        "arco_keys_transform_scratch_bytes": (C.c_int64, [i32, i32]),
        "arco_keys_transform": (C.c_int, [dp, bp, vp, vp, vp, vp]),
        "arco_proto_transform": (C.c_int, [i32, i32, vp, i32, vp, vp, vp]),
        "arco_anchor_gather": (C.c_int, [dp, vp, vp, vp, vp, vp, vp]),
        "arco_infonce_rows": (C.c_int, [dp, vp, vp, bp, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp]),
        "arco_similarity_dense_scratch": (C.c_int64, [i32, i32, i32, bp, vp]),
        "arco_similarity_dense": (C.c_int, [i32, i32, i32, i32, vp, vp, bp, vp, vp, vp, vp]),
        "arco_similarity_dense_backward_scratch": (C.c_int64, [i32, i32, i32, bp, vp]),
        "arco_similarity_dense_backward": (C.c_int, [i32, i32, i32, i32, vp, vp, bp, vp, vp, vp, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)      # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib, tuple(sigs)


lib, EXPORTS = _load()


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib.arco_last_error_string().decode(errors="replace")
        raise ArcoError(f"{what} failed with status {status}: {msg}")


def version() -> str:
    return lib.arco_version().decode()


def workspace_layout(dims: Dims) -> WsLayout:
    out = WsLayout()
    check(lib.arco_workspace_layout(C.byref(dims), C.byref(out)), "arco_workspace_layout")
    return out
