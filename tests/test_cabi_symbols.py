"""The C-ABI library loads and exports every symbol include/treat_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest

from treat_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "treat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(treat_b200_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 45
    L = ctypes.CDLL(cabi.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in treat_b200.h but not exported"
        assert n in cabi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(cabi.SIGNATURES) == names


def test_abi_version_and_strerror():
    L = cabi.lib()
    assert L.treat_b200_abi_version() == 1
    assert L.treat_b200_strerror(0) == b"ok"
    assert b"no CPU path" in L.treat_b200_strerror(cabi.ERR_NO_DEVICE)


def test_record_layout():
    assert cabi.RECORD_DTYPE.itemsize == 48
    # bytes 0..23 follow Alignment.MarshalBinary's field order (align.go:316-335)
    offs = {n: cabi.RECORD_DTYPE.fields[n][1] for n in cabi.RECORD_DTYPE.names}
    assert [offs[k] for k in ("edit_stop", "junc_start", "junc_end", "junc_len", "read_count", "has_mutation",
                              "alt_editing", "mismatches", "indel")] == [0, 4, 8, 12, 16, 20, 21, 22, 23]


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product must fail loudly, not fall back."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    ctx = ctypes.c_void_p()
    rc = cabi.lib().treat_b200_create(0, ctypes.byref(ctx))
    assert rc == cabi.ERR_NO_DEVICE and not ctx.value
    from treat_b200 import Aligner, TreatError
    with pytest.raises(TreatError):
        Aligner(0)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    for d, _, files in os.walk(os.path.join(ROOT, "treat_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(d, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("oracle/ ", ""), f"{f} mentions the oracle"


def test_header_is_plain_c(tmp_path):
    """cgo (and any other FFI generator) compiles include/treat_b200.h as C: it must be valid C99 on its own."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hdr.c"
    src.write_text('#include "treat_b200.h"\nint main(void) { return sizeof(treat_b200_record) == 48 ? 0 : 1; }\n')
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                           str(src), "-o", str(tmp_path / "hdr")])
    assert subprocess.call([str(tmp_path / "hdr")]) == 0
