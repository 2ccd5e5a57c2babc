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



def test_update_schedules_apply_every_update_once_and_in_order():
    """The three schedules of the multishift update (bcg_shift_schedule = the function the kernels evaluate on
    the device): simulate whole solves with random retirements (highest system first, as
    block_solvers.hpp:161,179-181) and random stopping points, and check that every system receives the update
    of every iteration in which it was active exactly once, in iteration order, and that a deferred update is
    only ever applied together with the Q of its own iteration (kept as "previous Q")."""
    import random

    from blockcg_b200.capi import shift_schedule
    KQ, KQ_KEEP, KQPREV, KCUR, KPREV, KBOTH = range(6)
    rng = random.Random(7)
    for trial in range(300):
        S = rng.randint(1, 9)
        n_it = rng.randint(1, 40)
        # active counts per iteration: non-increasing, >= 1
        act, a = [], S
        for i in range(n_it):
            act.append(a)
            if a > 1 and rng.random() < 0.15:
                a -= rng.randint(1, min(2, a - 1))
        want = {s: [i + 1 for i in range(n_it) if s < act[i]] for s in range(S)}
        for sched in (0, 1, 2):
            got = {s: [] for s in range(S)}
            kept = None   # iteration whose Q is available as "previous Q"
            total = 0
            for i in range(1, n_it + 1):
                items, passes = shift_schedule(sched, i, i == n_it, act[i - 1], act[i - 2] if i >= 2 else 0)
                total += passes
                kinds = [k for k, _ in items]
                assert kinds[0] in (KQ, KQ_KEEP) and kinds.count(KQ) + kinds.count(KQ_KEEP) == 1
                have_prev = KQPREV in kinds
                nsys = 0
                for k, s in items:
                    if k in (KPREV, KBOTH):
                        assert have_prev and i >= 2
                        assert kept == i - 1 or sched == 2   # alternating: the odd iteration's Q was kept
                        got[s].append(i - 1)
                    if k in (KCUR, KBOTH):
                        got[s].append(i)
                    nsys += k in (KCUR, KPREV, KBOTH)
                if kinds[0] == KQ_KEEP:
                    kept = i
                assert passes == (3 if (have_prev or kinds[0] == KQ_KEEP) else 2) + 4 * nsys
            assert got == want, (sched, S, act, got, want)
            if sched == 0:
                plain = total
            else:
                assert total <= plain + 2 * n_it   # deferring never moves more than the plain loop (+ the kept Q)


def test_deep_staggered_schedule_applies_every_update_once_and_in_order():
    """Schedule 3 (bcg_stag_schedule = build_stag_items, the function shift_stag_kernel and the B-step evaluate):
    simulated solves with random retirements (highest system first, block_solvers.hpp:161,179-181) and stopping
    points, deferral depths 2, 3, 4.  Every system receives the update of every iteration in which it was active
    exactly once and in order; an update older than `depth - 1` iterations is never requested (its Q field would
    have been overwritten); system 0 and the Q item come first in every launch; the byte accounting matches."""
    import random

    from blockcg_b200.capi import stag_schedule
    rng = random.Random(11)
    for trial in range(300):
        S = rng.randint(1, 9)
        n_it = rng.randint(1, 40)
        act, a = [], S
        for i in range(n_it):
            act.append(a)
            if a > 1 and rng.random() < 0.15:
                a -= rng.randint(1, min(2, a - 1))
        want = {s: [i + 1 for i in range(n_it) if s < act[i]] for s in range(S)}
        plain = sum(2 + 4 * act[i] for i in range(n_it))
        for depth, overlap in ((2, 0), (3, 0), (4, 0), (2, 1), (3, 1)):
            R = depth + overlap   # ring of Q fields / operand sets (one more when the launch overlaps the next iterations)
            got = {s: [] for s in range(S)}
            ring = [0, 0, 0, 0]
            total = 0
            for i in range(1, n_it + 1):
                if overlap:   # the two launches of the overlapped variant together do what the whole launch does
                    crit, p1 = stag_schedule(depth, i, i == n_it, act[i - 1], ring, R, 1)
                    bulk, p2 = stag_schedule(depth, i, i == n_it, act[i - 1], ring, R, 2)
                    assert crit == [(-1, 0, 0), (0, 0, 1)] and p1 == 6
                    assert all(s >= 1 for s, _, _ in bulk)
                    items, passes = crit + bulk, p1 + p2
                else:
                    items, passes = stag_schedule(depth, i, i == n_it, act[i - 1], ring)
                assert items[0] == (-1, 0, 0) and items[1] == (0, 0, 1)
                earlier = set()
                for s, first_back, m in items[1:]:
                    assert 0 <= first_back <= depth - 1 and 1 <= m <= first_back + 1
                    if s >= 1 and i != n_it:
                        assert s % depth == i % depth
                    for u in range(m):
                        got[s].append(i - first_back + u)
                        if first_back - u > 0 or (overlap and s >= 1):   # overlapped: every Q comes from the ring
                            earlier.add(first_back - u)
                assert len({s for s, _, _ in items}) == len(items)
                assert passes == 2 + len(earlier) + 4 * (len(items) - 1)
                total += passes
                ring[i % R] = act[i - 1]
            assert got == want, (depth, overlap, S, act, got, want)
            assert total <= plain + depth * n_it


def test_every_diagnostic_switch_is_documented():
    """Every environment switch the library reads (getenv("BCG_...") in csrc/) has a row in INTEGRATION.md."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set()
    for f in glob.glob(os.path.join(root, "blockcg_b200", "csrc", "*.cu*")):
        names.update(re.findall(r'getenv\("(BCG_[A-Z0-9_]+)"\)', open(f).read()))
    assert names, "no switches found: the pattern is stale"
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    missing = sorted(n for n in names if n not in doc)
    assert not missing, missing

