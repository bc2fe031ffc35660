"""CPU: the C-ABI library loads and exports every symbol include/b200zk.h declares; the
product refuses to run without a GPU instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b200zk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200zk_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_header(zk):
    L = zk.lib()
    names = declared_symbols()
    assert len(names) > 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_no_cpu_fallback(zk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(zk.B200zkError):
        zk.Backend(0)


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "halo2-experiments_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "blockexec.cuh", (dirpath, f)
