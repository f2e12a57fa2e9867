"""CPU checks of the drop-in boundary: the shared library builds/loads without a GPU and exports
every symbol include/dana_b200.h declares; the ctypes table covers the same set; the product path
refuses CPU tensors loudly instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "dana_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dana_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import dana_b200  # noqa: F401
    from dana_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libdana_b200.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert lib.dana_abi_version() == 2
    assert lib.dana_error_string(-1).decode() == "invalid argument"


def test_workspace_queries_are_host_only():
    import dana_b200  # noqa: F401
    from dana_b200 import _lib
    lib = _lib.load()
    assert lib.dana_nms_workspace_bytes(6000) > 6000 * 94 * 8
    assert lib.dana_proposals_workspace_bytes(4, 38 * 63 * 12, 6000) > 0
    assert lib.dana_roi_align_workspace_bytes(4, 1024, 38, 63, 0) >= 4 * 4 * 1024 * 38 * 63


def test_conv_gemm_args_layout_matches_header():
    """Field order of the ctypes mirror == field order of the C struct."""
    import dana_b200  # noqa: F401
    from dana_b200._lib import ConvGemmArgs
    src = open(os.path.join(ROOT, "include", "dana_b200.h")).read()
    body = src[src.index("typedef struct dana_conv_gemm_args {"):src.index("} dana_conv_gemm_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1] if not decl.startswith("const") else decl.split(None, 2)[2]
        for n in names.split(","):
            fields.append(n.replace("*", "").strip())
    assert fields == [f[0] for f in ConvGemmArgs._fields_]


def test_no_cpu_fallback():
    import dana_b200  # noqa: F401
    from dana_b200 import _C, _lib, ops
    with pytest.raises(_lib.DanaError):
        ops.nms(torch.zeros(4, 4), torch.zeros(4), 0.5)
    with pytest.raises(RuntimeError):
        _C.nms(torch.zeros(4, 4), torch.zeros(4), 0.5)
    with pytest.raises(RuntimeError):
        _C.roi_align_forward(torch.zeros(1, 1, 4, 4), torch.zeros(1, 5), 1.0, 2, 2, 0)
    with pytest.raises(RuntimeError):
        _C.roi_pool_forward(torch.zeros(1, 1, 4, 4), torch.zeros(1, 5), 1.0, 2, 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dual-awareness-attention-for-few-shot-object-detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "dana_oracle" not in text and "import oracle" not in text and "oracle/" not in text, f


def test_header_is_plain_c():
    """include/dana_b200.h is the drop-in boundary: it must compile as C99 and as C++ without any CUDA / torch header."""
    import shutil
    import subprocess
    import tempfile
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "h.c")
        with open(src, "w") as f:
            f.write('#include "include/dana_b200.h"\nint main(void) { return dana_abi_version == 0; }\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-fsyntax-only", "-I", ROOT, src], cwd=ROOT)
        if shutil.which("g++") is not None:
            subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", ROOT, src], cwd=ROOT)
