"""Device-resident ring-buffer memory bank that keeps the reference's list protocol.

The reference trainers own three Python lists (``/root/reference/code/train_arco_2d.py:147-154``,
``train_arco_3d.py:144-151``)::

    memobank[c]     = [cpu_tensor[n_c, D]]      # rows in FIFO order, newest last
    queue_ptrlis[c] = LongTensor(1)
    queue_size[c]   = int                        # 50000 for class 0, 30000 otherwise

and ``dequeue_and_enqueue`` (``loss_helper_3d.py:12-32``) copies every step's keys to the CPU,
concatenates, and keeps the newest ``queue_size`` rows; the loss then uploads the whole bank again
(``:466``).  Here the rows live in HBM as one ``[sum(cap), D]`` ring per class; (head, length,
pointer) are device scalars updated by ``arco_scan_plan``; nothing crosses PCIe in the step.

Row storage is fp32, or bf16 when the trainer runs the representation head in bf16 AND every adopted row is exactly
representable in bf16 (keys are teacher rows, so with a bf16 ``rep_teacher`` the narrow ring holds bit-identical
values at half the gather traffic).  Assigning a row that is not bf16-exact, or a later fp32 call, widens the ring
back to fp32 -- values never change.  ``ARCO_BANK_BF16=0`` forces fp32 storage.

The caller's lists are adopted lazily on the first call: ``memobank[c]`` is replaced by a
:class:`BankSlot` (a ``list`` subclass) whose element 0 still answers ``.shape[0]`` and row indexing in
logical FIFO order, and ``queue_ptrlis[c][0]`` is refreshed from the device bookkeeping whenever the
host mirror is refreshed: the last CTA of every step stores its ``arco_plan`` into a pinned, PCIe-mapped ring (zero-copy) and
:meth:`DeviceMemoryBank.poll` (non-blocking, called at the start of the next step) applies the plans that have landed, so the pointer is normally one step late and never stale by more than the steps still in flight;
:meth:`DeviceMemoryBank.settle` (blocking) is used on any inspection.
"""
from __future__ import annotations

import collections
import ctypes as C
import os
from typing import List, Optional, Sequence

import torch

from . import _cabi


class BankSlot(list):
    """Stands in for ``memobank[c]``.  ``slot[0]`` materialises the class's rows (FIFO order) on the
    bank's device; ``slot[0] = tensor`` replaces the class's content (as a trainer resetting its bank)."""

    def __init__(self, bank: "DeviceMemoryBank", cls: int):
        super().__init__([None])
        self.bank = bank
        self.cls = cls

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(1))]
        if i not in (0, -1):
            raise IndexError("memobank[c] holds exactly one tensor")
        return self.bank.rows_of(self.cls)

    def __setitem__(self, i, value):
        if i not in (0, -1):
            raise IndexError("memobank[c] holds exactly one tensor")
        self.bank.replace_rows(self.cls, value)

    def __iter__(self):
        yield self[0]

    def __repr__(self):
        return f"BankSlot(class={self.cls}, rows={self.bank.length(self.cls)}, cap={self.bank.caps[self.cls]})"


def _bf16_exact(t: torch.Tensor) -> bool:
    t = t.to(torch.float32)
    return bool((t.to(torch.bfloat16).to(torch.float32) == t).all())


class DeviceMemoryBank:
    def __init__(self, memobank: list, queue_ptrlis: list, queue_size: Sequence[int], feat: int, device,
                 prefer_bf16: bool = False):
        self.device = torch.device(device)
        self.feat = int(feat)
        self.classes = len(memobank)
        if self.classes > _cabi.MAX_CLASSES:
            raise ValueError(f"at most {_cabi.MAX_CLASSES} classes are supported, got {self.classes}")
        if len(queue_size) != self.classes or len(queue_ptrlis) != self.classes:
            raise ValueError("memobank, queue_prtlis and queue_size must have one entry per class")
        self.caps = [int(q) for q in queue_size]
        if min(self.caps) <= 0:
            raise ValueError("queue_size entries must be positive")
        self.row_off = [0]
        for cap in self.caps:
            self.row_off.append(self.row_off[-1] + cap)
        narrow = (prefer_bf16 and self.feat % 8 == 0 and os.environ.get("ARCO_BANK_BF16", "1") != "0"
                  and all(_bf16_exact(m[0]) for m in memobank))
        self.rows = torch.zeros((self.row_off[-1], self.feat), dtype=torch.bfloat16 if narrow else torch.float32,
                                device=self.device)
        head = torch.zeros(_cabi.MAX_CLASSES, dtype=torch.int32)
        length = torch.zeros(_cabi.MAX_CLASSES, dtype=torch.int32)
        ptr = torch.zeros(_cabi.MAX_CLASSES, dtype=torch.int64)
        for c in range(self.classes):
            init = memobank[c][0]
            if init.dim() != 2 or init.shape[1] != self.feat:
                raise ValueError(f"memobank[{c}][0] must be [n, {self.feat}], got {tuple(init.shape)}")
            init = init[-self.caps[c]:] if init.shape[0] > self.caps[c] else init
            n = init.shape[0]
            if n:
                self.rows[self.row_off[c]: self.row_off[c] + n] = init.to(self.device, self.rows.dtype)
            length[c] = n
            ptr[c] = int(queue_ptrlis[c].reshape(-1)[0])
        self.head = head.to(self.device)
        self.len = length.to(self.device)
        self.ptr = ptr.to(self.device)
        self._plan_view = None          # device view of the last step's arco_plan
        self._dirty = False
        self._edited = False            # rows replaced by the caller since the last step
        # Pinned, PCIe-mapped host mirror the device writes by itself (zero-copy; see begin_step / poll):
        #   [_MIRROR_SLOTS x _MIRROR_STRIDE] ring of finished steps' arco_plan + sequence number,
        #   then int64[MAX_CLASSES] "live" queue pointers -- the caller's queue_prtlis[c] are rebound to 1-element views of
        #   this array, so they follow the device with no host work at all.
        self._mirror = torch.zeros(self._MIRROR_SLOTS * self._MIRROR_STRIDE + 8 * _cabi.MAX_CLASSES, dtype=torch.uint8,
                                   pin_memory=True)
        self._mirror_np = self._mirror.numpy()
        n_plan = C.sizeof(_cabi.Plan)
        words = self._mirror_np[: self._MIRROR_SLOTS * self._MIRROR_STRIDE].reshape(self._MIRROR_SLOTS, self._MIRROR_STRIDE)
        self._seq_np = words[:, n_plan: n_plan + 8].view("<u8").reshape(-1)                       # [slots]
        self._status_np = words[:, _cabi.Plan.status.offset: _cabi.Plan.status.offset + 4].view("<u4").reshape(-1)
        self._ptr_alias = self._mirror[self._MIRROR_SLOTS * self._MIRROR_STRIDE:].view(torch.int64)   # [MAX_CLASSES]
        self._ptr_alias[: self.classes] = ptr[: self.classes]
        self._applied = 0               # last step number whose mirrored plan was looked at
        self._latest = None             # (slot, step) of the newest plan that landed
        self._host_len0 = [int(x) for x in length[: self.classes]]
        self._plan_cache = (None, None)
        self._settled_plan: Optional[_cabi.Plan] = None
        self._settled_step = -1
        self.step = 0
        self._bind_pointers(queue_ptrlis)
        self.c_struct = _cabi.Bank()
        self.c_struct.rows = self.rows.data_ptr()
        self.c_struct.row_dtype = _cabi.BF16 if narrow else _cabi.F32
        self.c_struct.head = self.head.data_ptr()
        self.c_struct.len = self.len.data_ptr()
        self.c_struct.queue_ptr = self.ptr.data_ptr()
        self.c_struct.host_queue_ptr = self._ptr_alias.data_ptr()
        self.c_struct.host_mirror = self._mirror.data_ptr()
        self._counters = torch.zeros(_cabi.COUNTER_WORDS, dtype=torch.int32, device=self.device)   # self-cleaning tickets
        self.c_struct.counters = self._counters.data_ptr()
        for c in range(self.classes):
            self.c_struct.cap[c] = self.caps[c]
            self.c_struct.row_off[c] = self.row_off[c]
        # adopt: the caller's list now fronts the device bank
        for c in range(self.classes):
            memobank[c] = BankSlot(self, c)

    # ------------------------------------------------------------------ adoption
    @staticmethod
    def adopt(memobank: list, queue_ptrlis: list, queue_size: Sequence[int], feat: int, device,
              rep_dtype: torch.dtype = torch.float32) -> "DeviceMemoryBank":
        first = memobank[0] if len(memobank) else None
        if isinstance(first, BankSlot):
            bank = first.bank
            if bank.device != torch.device(device) or bank.feat != int(feat) or bank.classes != len(memobank):
                raise ValueError("memobank was adopted for a different device / feature size / class count")
            if [int(q) for q in queue_size] != bank.caps:
                raise ValueError("queue_size changed after the memory bank was adopted")
            bank._bind_pointers(queue_ptrlis)
            if rep_dtype != torch.bfloat16:
                bank.widen()                     # fp32 keys are not bf16-exact in general
            return bank
        return DeviceMemoryBank(memobank, queue_ptrlis, queue_size, feat, device, rep_dtype == torch.bfloat16)

    @property
    def row_dtype(self) -> torch.dtype:
        return self.rows.dtype

    def widen(self) -> None:
        """Switch a bf16 ring to fp32 storage (exact); no-op for an fp32 ring."""
        if self.rows.dtype == torch.float32:
            return
        self.rows = self.rows.to(torch.float32)
        self.c_struct.rows = self.rows.data_ptr()
        self.c_struct.row_dtype = _cabi.F32

    # ------------------------------------------------------------------ host mirror
    _MIRROR_SLOTS = _cabi.MIRROR_SLOTS      # ring of pinned plan copies the device writes (ARCO_MIRROR_SLOTS)
    _MIRROR_STRIDE = _cabi.MIRROR_STRIDE    # bytes per slot: arco_plan + u64 sequence number at offset sizeof(arco_plan)

    def _bind_pointers(self, queue_ptrlis: list) -> None:
        """Rebind the caller's ``queue_prtlis[c]`` to 1-element views of the pinned live-pointer array (like
        ``memobank[c]`` is rebound to a :class:`BankSlot`): the device stores the reference's pointer values there at the
        end of every step, so ``queue_prtlis[c][0]`` follows the device without any host-side work."""
        views = self.__dict__.get("_alias_views")
        if views is None:
            views = self._alias_views = [self._ptr_alias[c: c + 1] for c in range(self.classes)]
        if queue_ptrlis is self.__dict__.get("_queue_ptrlis") and queue_ptrlis[0] is views[0]:
            return                                                # the steady state: one identity check per step
        self._queue_ptrlis = queue_ptrlis
        for c in range(self.classes):
            if queue_ptrlis[c] is not views[c]:
                queue_ptrlis[c] = views[c]

    def begin_step(self) -> None:
        """Nothing changes in the launch parameters from step to step (so a step can be captured in a CUDA graph and
        replayed): the LAST CTA of every step's InfoNCE kernel stores the final ``arco_plan`` into slot ``seq % 8`` of the
        pinned mirror ring through its PCIe-mapped address, then the live queue pointers, then ``seq`` -- the bank's
        DEVICE step counter -- which is how ``queue_prtlis``, the bank lengths and the label-error status reach the host
        without the step ever synchronising (see :meth:`poll`)."""

    def post_step(self, plan_view: torch.Tensor) -> None:
        """Remember the step's device-resident ``arco_plan`` (a view into its workspace); nothing is copied."""
        self._plan_view = plan_view
        self._dirty = True
        self._edited = False
        self._settled_plan = None
        self.step += 1

    def poll(self, block: bool = False) -> Optional[_cabi.Plan]:
        """Non-blocking look at the host mirror: for every step that has FINISHED on the device since the last call
        (device sequence numbers 1, 2, 3, ... -- replays of a captured CUDA graph count too), the status word is checked
        (invalid labels / indices raise ``ValueError`` here, i.e. at the start of the first call after the offending
        step has completed) and the newest landed plan is remembered for :attr:`host_len` / :attr:`last_plan`.  A handful
        of 8-byte reads; never waits for the device unless ``block``."""
        if block:
            return self.settle()
        slots = self._MIRROR_SLOTS
        newest = int(self._seq_np.max())
        if newest <= self._applied:
            return None
        first = max(self._applied + 1, newest - slots + 1)            # older ones were overwritten in the ring
        bad = 0
        for s_no in range(first, newest + 1):
            slot = s_no % slots
            if int(self._seq_np[slot]) != s_no:
                continue                                          # not landed yet / already reused
            bad |= int(self._status_np[slot])
            self._latest = (slot, s_no)
        self._applied = newest
        if bad:
            self.check_status(bad)
        return None

    def _landed_plan(self) -> Optional[_cabi.Plan]:
        """The newest mirrored plan, parsed on demand (None if none has landed or its slot was reused meanwhile)."""
        if self._latest is None:
            return None
        slot, s_no = self._latest
        if self._plan_cache[0] == s_no:
            return self._plan_cache[1]
        off = slot * self._MIRROR_STRIDE
        raw = self._mirror_np[off: off + C.sizeof(_cabi.Plan)].tobytes()
        if int(self._seq_np[slot]) != s_no:
            return None
        plan = _cabi.Plan.from_buffer_copy(raw)
        self._plan_cache = (s_no, plan)
        return plan

    @property
    def host_ptr(self) -> List[int]:
        """Reference pointer values (loss_helper_3d.py:24-30) as of the last step that finished on the device."""
        return [int(v) for v in self._ptr_alias[: self.classes].tolist()]

    @property
    def host_len(self) -> List[int]:
        """Bank lengths as of the newest step whose plan has landed on the host (no synchronisation)."""
        plan = self._landed_plan()
        if plan is None or self._edited:
            return list(self._host_len0)
        return [int(plan.bank_len[c]) for c in range(self.classes)]

    @property
    def last_plan(self) -> Optional[_cabi.Plan]:
        return self._settled_plan if self._settled_plan is not None else self._landed_plan()

    @staticmethod
    def check_status(status: int) -> None:
        if status & _cabi.ST_MULTI_HOT:
            raise ValueError("label_l/label_u are not one-hot: some pixel has more than one non-zero class entry")
        if status & _cabi.ST_LABEL_RANGE:
            raise ValueError("integer label map contains a class id >= num_classes")
        if status & _cabi.ST_INDEX_RANGE:
            raise ValueError("a sample index was outside its candidate list / memory bank (it was clamped on the device)")
        if status & _cabi.ST_KEYS_DROPPED:
            raise ValueError("negative keys were counted but not enqueued: the C<=3 register prototype kernel needs "
                             "low_rank >= num_classes")
        if status & _cabi.ST_EXCHANGE_DESYNC:
            raise RuntimeError("multi-GPU prototype exchange: the device's step word and the slot of this step disagree "
                               "(a replayed step ran out of order)")
        if status & _cabi.ST_EXCHANGE_TIMEOUT:
            raise RuntimeError("multi-GPU prototype exchange timed out waiting for a peer")

    def settle(self) -> Optional[_cabi.Plan]:
        """Read the device bookkeeping (host sync) and refresh the host mirrors, including the caller's
        ``queue_prtlis``.  Raises if the device flagged invalid labels.  Used by inspection only: the training step
        itself never calls it."""
        if not self._dirty:
            return self.last_plan
        self._dirty = False
        self._edited = False
        lens = self.len.cpu().tolist()
        ptrs = self.ptr.cpu().tolist()                            # (stream-ordered copy: every launched step has finished)
        self._applied = int(self._seq_np.max())
        self._latest = None
        self._host_len0 = [int(v) for v in lens[: self.classes]]
        self._ptr_alias[: self.classes] = torch.tensor(ptrs[: self.classes], dtype=torch.int64)
        if self._plan_view is not None:
            self._settled_plan = _cabi.Plan.from_buffer_copy(self._plan_view.cpu().numpy().tobytes())
            self._settled_step = self.step
            self.check_status(self._settled_plan.status)
        return self._settled_plan

    def length(self, cls: int) -> int:
        self.settle()
        return self.host_len[cls]

    # ------------------------------------------------------------------ list protocol
    def rows_of(self, cls: int) -> torch.Tensor:
        """Rows of class ``cls`` in logical FIFO order, ``[len, D]`` fp32 on the bank's device."""
        n = self.length(cls)
        out = torch.empty((self.caps[cls], self.feat), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream()
            _cabi.check(_cabi.lib.arco_bank_read(C.byref(self.c_struct), cls, self.feat, out.data_ptr(), st.cuda_stream),
                        "arco_bank_read")
        return out[:n]

    def replace_rows(self, cls: int, value: torch.Tensor) -> None:
        self.settle()
        if value.dim() != 2 or value.shape[1] != self.feat:
            raise ValueError(f"bank rows must be [n, {self.feat}]")
        value = value[-self.caps[cls]:] if value.shape[0] > self.caps[cls] else value
        n = value.shape[0]
        if self.rows.dtype == torch.bfloat16 and not _bf16_exact(value):
            self.widen()
        self.rows[self.row_off[cls]: self.row_off[cls] + n] = value.to(self.device, self.rows.dtype)
        self.head[cls] = 0
        self.len[cls] = n
        self._host_len0 = self.host_len
        self._host_len0[cls] = n
        self._dirty = True
        self._edited = True


def synchronize_bank(memobank: list) -> None:
    """Bring ``queue_prtlis`` and the bank lengths up to date with the device (host sync)."""
    if len(memobank) and isinstance(memobank[0], BankSlot):
        memobank[0].bank.settle()
