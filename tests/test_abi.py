"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol
include/blockcg_b200.h declares, refuses to run without a GPU (no CPU fallback), and the
C++ host mirror (blockcg_b200/host/*.hpp) compiles against it."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "blockcg_b200.h")
LIB = os.path.join(ROOT, "blockcg_b200", "libblockcg_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bcg_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "blockcg_b200", "csrc")], check=True)
    return ctypes.CDLL(LIB)


def test_header_symbols_are_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_lists_every_symbol():
    from blockcg_b200.capi import EXPORTS
    assert sorted(EXPORTS) == declared_symbols()


def test_supported_nrhs(lib):
    for n in (1, 2, 3, 4, 6, 8, 12, 16):
        assert lib.bcg_supports_nrhs(n) == 1
    assert lib.bcg_supports_nrhs(5) == 0
    assert lib.bcg_supports_nrhs(0) == 0


def test_no_cpu_fallback():
    """Without a CUDA device context creation must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import blockcg_b200
    with pytest.raises(blockcg_b200.BcgError) as ei:
        blockcg_b200.Context(16, 3)
    assert "no CPU fallback" in str(ei.value)
    import numpy as np
    D = blockcg_b200.dirac_op(16, 0.5, links=np.zeros((16, 3, 3), complex))
    X = np.zeros((16, 3, 3), complex)
    with pytest.raises(blockcg_b200.BcgError):
        blockcg_b200.BCGrQ(X, X.copy(), D, 1e-10)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under blockcg_b200/ may reference it."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "blockcg_b200")):
        if "build" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"\boracle\b|liboracle|libref_n|/root/reference", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_host_mirror_compiles(lib):
    """Reference-style C++ drivers (tests + benchmark) build against the drop-in headers."""
    host = os.path.join(ROOT, "blockcg_b200", "host")
    subprocess.run(["make", "-C", host, "-B", "all"], check=True, capture_output=True)
    assert os.path.exists(os.path.join(host, "test_solvers"))
    assert os.path.exists(os.path.join(host, "benchmark"))


def test_makefile_tracks_every_header():
    """Every header of the CUDA sources is a prerequisite of the objects: a header left out once kept a
    stale kernel in the library although `make` reported it up to date."""
    import glob
    import re
    csrc = os.path.join(ROOT, "blockcg_b200", "csrc")
    mk = open(os.path.join(csrc, "Makefile")).read()
    hdrs = re.search(r"^HDRS\s*=\s*(.*)$", mk, re.M).group(1).split()
    have = {os.path.basename(h) for h in hdrs}
    want = {os.path.basename(f) for f in glob.glob(os.path.join(csrc, "*.cuh")) + glob.glob(os.path.join(csrc, "*.h"))}
    assert want <= have, sorted(want - have)

