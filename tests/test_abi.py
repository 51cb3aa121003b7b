"""The C-ABI library builds, loads and exports every symbol include/veomni_b200.h declares."""
import ctypes

from veomni_b200 import _lib, build


def test_library_builds_and_loads():
    path = build.build()
    assert path.exists()
    lib = _lib.load()
    assert lib.vb200_abi_version() == 1


def test_every_declared_symbol_is_exported_and_bound():
    build.build()
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    declared = _lib.declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in the header but not exported by the library"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in veomni_b200/_lib.py"
    for name in _lib.SIGNATURES:
        assert name in declared, f"{name} bound in _lib.py but not declared in include/veomni_b200.h"


def test_product_path_fails_loudly_without_cuda_tensors():
    import pytest
    import torch

    from veomni_b200 import functional as F

    with pytest.raises(_lib.VB200Error):
        F.rms_norm(torch.randn(4, 64, dtype=torch.bfloat16), torch.ones(64, dtype=torch.bfloat16), 1e-6)


def test_product_does_not_import_oracle():
    import pathlib
    import re

    pkg = pathlib.Path(_lib.__file__).parent
    for f in pkg.rglob("*.py"):
        text = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
