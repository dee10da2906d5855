"""CPU-side checks of the C-ABI boundary: the library loads without a GPU, exports every symbol
declared in include/sleap_b200.h, and refuses to create a handle without a device (no fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sleap_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|const char\*)\s+(sb_[a-z0-9_]+)\s*\(", src, flags=re.M)))


def test_header_symbols_exported():
    from sleap_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), set(names) ^ set(_lib.EXPORTED_SYMBOLS)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sleap_b200 import _lib
    with pytest.raises(_lib.SleapB200Error):
        _lib.Handle(0)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sleap_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
