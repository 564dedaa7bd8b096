"""CPU: the C-ABI library loads, exports every symbol include/arco_b200.h declares, and the ctypes mirrors of
the header's structs have the C compiler's sizes.  No compute call is made (no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "arco_b200.h")


def _declared():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"ARCO_API\s+(?:const\s+char\*|int64_t|int)\s+(arco_\w+)\s*\(", src)))


def test_header_declares_the_path():
    names = _declared()
    for must in ("arco_classify_count", "arco_classify_plan", "arco_scan_plan", "arco_proto_enqueue", "arco_sample", "arco_infonce",
                 "arco_grad_scatter", "arco_workspace_layout", "arco_last_error_string", "arco_label_onehot"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from arco_b200 import _cabi
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} is declared in include/arco_b200.h but not exported"
    assert set(_declared()) == set(_cabi.EXPORTS), "ctypes prototypes and header disagree"
    lib.arco_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.arco_version()


def test_struct_sizes_match_the_c_compiler(tmp_path):
    from arco_b200 import _cabi
    prog = tmp_path / "sizes.c"
    prog.write_text('#include <stdio.h>\n#include "%s"\nint main(void){printf("%%zu %%zu %%zu %%zu %%zu\\n", sizeof(arco_dims), '
                    'sizeof(arco_ws_layout), sizeof(arco_plan), sizeof(arco_bank), sizeof(arco_step_io));return 0;}\n' % HEADER)
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", str(prog), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [ctypes.sizeof(_cabi.Dims), ctypes.sizeof(_cabi.WsLayout), ctypes.sizeof(_cabi.Plan),
                     ctypes.sizeof(_cabi.Bank), ctypes.sizeof(_cabi.StepIO)]


def test_layout_and_argument_errors_without_a_gpu():
    from arco_b200 import _cabi
    d = _cabi.Dims(4, 4, 4, 64, 65536, 256, 512, _cabi.F32, _cabi.LABEL_ONEHOT_I64)
    L = _cabi.workspace_layout(d)
    assert L.n_tiles == 8 * 64 and L.tiles_per_image == 64
    assert L.total_bytes > 8 * 65536 and L.codes % 256 == 0 and L.plan == 0
    bad = _cabi.Dims(4, 4, 40, 64, 65536, 256, 512, 0, 0)          # 40 classes > 32
    with pytest.raises(_cabi.ArcoError, match="classes"):
        _cabi.workspace_layout(bad)
    bad = _cabi.Dims(4, 4, 4, 30, 65536, 256, 512, 0, 0)           # D not a multiple of 4
    with pytest.raises(_cabi.ArcoError, match="multiple of 4"):
        _cabi.workspace_layout(bad)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "arco_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports the oracle"


def test_widened_ops_refuse_cpu_tensors():
    """The ops added for SURVEY 8(f) have no CPU / PyTorch fallback either: CPU tensors raise before anything runs."""
    import pytest
    import torch

    import arco_b200
    x = torch.zeros(2, 16, 8, 8)
    lab = torch.zeros(1, 8, 8, dtype=torch.int64)
    prob = torch.zeros(1, 4, 8, 8)
    mask = torch.zeros(2, 1, 8, 8)
    bank = [[torch.zeros(1, 16)] for _ in range(4)]
    ptr = [torch.zeros(1, dtype=torch.long) for _ in range(4)]
    w = torch.eye(16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        arco_b200.compute_contra_memobank_loss_from_features(x, x, [w, w, w], w, lab, lab, prob, prob, mask, mask, bank, ptr, [8] * 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        arco_b200.compute_contra_memobank_loss_from_logits(x, lab, lab, prob, prob, prob, 20.0, bank, ptr, [8] * 4, x)
    with pytest.raises((RuntimeError, ValueError)):
        arco_b200.get_revisiting_loss(torch.zeros(4, 16 * 64), x, x, topk=2)
