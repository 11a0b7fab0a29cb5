"""The C-ABI libraries load on a CPU-only box and export every symbol include/*.h declares."""
import ctypes as C
import glob
import os
import re

import pytest

from conftest import ROOT
from dipgenie_b200 import _build


def declared_symbols(header):
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dgh?_[a-z0-9_]+)\s*\(", txt)))


@pytest.mark.parametrize("header", sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))))
def test_library_exports_every_declared_symbol(header):
    lib_path = _build.HOST_LIB if "host" in os.path.basename(header) else _build.CUDA_LIB
    if "host" in os.path.basename(header):
        _build.build_host()
    else:
        _build.build_cuda()
    lib = C.CDLL(lib_path)
    syms = declared_symbols(header)
    assert syms, header
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"{os.path.basename(lib_path)} lacks {missing}"


def test_no_gpu_means_loud_failure_not_fallback():
    """dg_create must fail (NULL) without a device; the Python wrapper turns that into an exception."""
    import torch
    from dipgenie_b200.cuda_api import Context, DipGenieCudaError
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(DipGenieCudaError):
        Context(0)
