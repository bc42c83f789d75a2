"""CPU-side checks of the C ABI: the library builds for sm_100a, loads, and exports every symbol the public header
declares; the product package never imports the oracle; missing library / CPU tensors fail loudly."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "speech-to-image-translation-without-text_b200")


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__
    return __graft_entry__.build()


def test_library_exports_every_declared_symbol(lib_path):
    from sg2b200 import _lib
    sigs = _lib.parse_header()
    assert len(sigs) >= 30
    l = ctypes.CDLL(lib_path)
    missing = [n for n in list(sigs) + ["sg2_last_error"] if not hasattr(l, n)]
    assert not missing, missing
    assert l.sg2_version() >= 1


def test_library_contains_sm100a_tensor_core_and_tma_code(lib_path):
    """SASS evidence (B200_PROFILING.md): tcgen05.mma -> UTCHMMA, TMA -> UTMALDG, tcgen05.ld -> LDTM."""
    try:
        sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True, timeout=300).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass          # no legacy mma.sync path


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, "include", "sg2b200.h")).read()
    assert text.count("model.py:") >= 8 and text.count("trainer.py:") >= 3


def test_cpu_tensor_is_rejected_without_fallback(lib_path):
    import torch
    from sg2b200 import ops
    with pytest.raises(RuntimeError):
        ops.add_bf16(torch.zeros(8, dtype=torch.bfloat16), torch.zeros(8, dtype=torch.bfloat16))
