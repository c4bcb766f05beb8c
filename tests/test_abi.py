"""The C-ABI shared library: loads, exports every symbol include/fmc.h declares, and refuses to
run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from fast_monte_carlo_b200 import native


def _declared():
    src = open(os.path.join(ROOT, "include", "fmc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fmc_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(native_lib):
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(native_lib, n), f"{n} declared in include/fmc.h but not exported"
    assert set(names) == set(native.EXPORTED_SYMBOLS)
    assert native_lib.fmc_abi_version() == 4


def test_no_cpu_fallback(native_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = native_lib.fmc_create(0, ctypes.byref(h))
    assert rc == -3                                   # FMC_ERR_NO_DEVICE
    assert b"no CPU fallback" in native_lib.fmc_last_error()
    with pytest.raises(native.FmcError):
        native.Context(0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "fast_monte_carlo_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "c_oracle" not in txt and "fmc_oracle" not in txt and "from oracle" not in txt and \
                    "import oracle" not in txt, fn
