"""The C-ABI shared library loads without a GPU / libcuda and exports every symbol that
include/fdx.h declares.  No compute entry point is called here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from flaxdiff_b200 import _lib
    if not os.path.exists(_lib.lib_path()):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "fdx.h")).read()
    names = sorted(set(re.findall(r"^(?:int|const char\*|unsigned long long)\s+(fdx_[a-z0-9_]+)\s*\(", hdr, re.M)))
    assert len(names) >= 30
    for n in names:
        assert getattr(lib, n) is not None, n


def test_version_and_error_string(lib):
    assert lib.fdx_version() >= 100
    lib.fdx_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.fdx_last_error(), bytes)


def test_no_libcuda_link_dependency():
    """libfdx must load on a box without the driver: TMA descriptors are encoded through
    cudaGetDriverEntryPoint at run time, cudart is linked statically."""
    import subprocess
    from flaxdiff_b200 import _lib
    out = subprocess.run(["ldd", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out and "libtorch" not in out


def test_ops_refuse_cpu_tensors(lib):
    import torch
    from flaxdiff_b200 import ops
    from flaxdiff_b200._lib import FdxError
    x = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16)
    with pytest.raises(FdxError):
        ops.groupnorm_stats(x, 8)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under flaxdiff_b200/ may reference it."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "flaxdiff_b200")):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)
