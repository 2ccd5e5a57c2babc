"""CPU tests: the oracle restatement (oracle/oracle.cpp) against
 (a) the committed golden vectors generated from the unmodified reference
     (oracle/gen_golden.py), and
 (b) the reference itself (oracle/_ref) run live on fresh inputs, when present.
This is what pins the oracle (task rule 3); the GPU parity tests then compare
the CUDA path with the oracle / the same fixtures."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden
from oracle.pyoracle import RefShim

PRIMS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "prim_*.npz")))
SOLVES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "solve_*.npz")))


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("name", PRIMS)
def test_primitives_match_golden(oracle, name):
    g = golden(name)
    V, N, mass = int(g["V"]), int(g["N"]), float(g["mass"])
    U, B, M = g["U"], g["B"], g["M"]
    U2, B2 = oracle.make_inputs(V, N, int(g["seed"]))
    assert np.array_equal(U, U2) and np.array_equal(B, B2)  # RNG replay is bit-exact
    AB = oracle.op(U, B, mass)
    assert rel(AB, g["op"]) < 1e-14
    assert rel(oracle.hermitian_dot(B, g["op"]), g["gram_B_AB"]) < 1e-14
    assert rel(oracle.hermitian_dot(B, B), g["gram_BB"]) < 1e-14
    assert rel(oracle.add(B, g["op"], M), g["add"]) < 1e-14
    assert rel(oracle.add(B, g["op"], 0.375), g["add_scalar"]) < 1e-14
    assert rel(oracle.rescale_add(B, M, g["op"], 1.0), g["rescale_add"]) < 1e-14
    Q, R = oracle.thinQR(B)
    assert rel(R, g["thinqr_R"]) < 1e-13 and rel(Q, g["thinqr_Q"]) < 1e-12
    assert np.all(np.tril(R, -1) == 0)  # exactly zero below the diagonal (fields.hpp:142)
    assert rel(oracle.fullpivlu_inverse(M), g["lu_inv_M"]) < 1e-11
    assert rel(oracle.fullpivlu_inverse(g["gram_B_AB"]), g["lu_inv_G"]) < 1e-11
    Rl, info = oracle.llt_upper(g["gram_BB"])
    assert info == -1 and rel(Rl, g["llt_upper_BB"]) < 1e-13
    # tree-order Gram agrees with the sequential one to rounding
    assert rel(oracle.hermitian_dot(B, g["op"], chunk=4), g["gram_B_AB"]) < 1e-13


@pytest.mark.parametrize("name", SOLVES)
def test_solvers_match_golden(oracle, name):
    g = golden(name)
    U, B, mass, eps = g["U"], g["B"], float(g["mass"]), float(g["eps"])
    X, it, _ = oracle.BCG(U, B, mass, eps)
    assert abs(it - int(g["it_bcg"])) <= 1 and rel(X, g["X_bcg"]) < 1e-9
    X, it, _ = oracle.BCGrQ(U, B, mass, eps)
    assert abs(it - int(g["it_bcgrq"])) <= 1 and rel(X, g["X_bcgrq"]) < 1e-9
    Xs, it, _, _ = oracle.SBCGrQ(U, B, mass, g["shifts"], eps, float(g["eps_shifts"]))
    assert abs(it - int(g["it_sbcgrq"])) <= 1
    for s in range(len(g["shifts"])):
        assert rel(Xs[s], g["X_sbcgrq"][s]) < 1e-9
        # the reference's own acceptance rule (test/solvers.cpp:116)
        assert oracle.true_residual(U, B, Xs[s], mass, g["shifts"][s]).max() < 2 * eps
    assert np.array_equal(Xs[0], X)  # SBCGrQ shift 0 == BCGrQ (same arithmetic)


def test_benchmark_default_config(oracle):
    """./benchmark 1e3 1e-3 1e-10 (README.md:29): iteration count and residuals."""
    g = golden("bench_V1000_N12.npz")
    V, N, mass, eps = int(g["V"]), int(g["N"]), float(g["mass"]), float(g["eps"])
    U, B = oracle.make_inputs(V, N, 1)
    Xs, it, _, _ = oracle.SBCGrQ(U, B, mass, g["shifts"], eps, 1e-15)
    assert abs(it - int(g["it_sbcgrq"])) <= 1
    st = int(g["sample_stride"])
    for s in range(len(g["shifts"])):
        ref = g["X_sample"][s]
        assert np.abs(Xs[s][::st] - ref).max() / np.abs(ref).max() < 1e-9
    cn = np.sqrt((np.abs(Xs) ** 2).sum(axis=(1, 3)))
    assert np.abs(cn / g["X_colnorm"] - 1).max() < 1e-9
    # F7b: a tree-shaped Gram is more accurate and converges in fewer iterations
    _, it_tree, _, _ = oracle.SBCGrQ(U, B, mass, g["shifts"], eps, 1e-15, chunk=32)
    assert it_tree <= it


@pytest.mark.parametrize("N", [1, 3, 4, 12])
def test_oracle_vs_live_reference(oracle, N):
    if not RefShim.available(N):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    r = RefShim(N)
    V, mass = 96, 0.2
    U, B = r.make_inputs(V, 7)
    U2, B2 = oracle.make_inputs(V, N, 7)
    assert np.array_equal(U, U2) and np.array_equal(B, B2)
    assert rel(oracle.op(U, B, mass), r.op(U, B, mass)) < 1e-14
    sig = [0.0, 0.05, 0.5]
    Xr, itr, _ = r.SBCGrQ(U, B, mass, sig, 1e-10, 1e-15)
    Xo, ito, _, _ = oracle.SBCGrQ(U, B, mass, sig, 1e-10, 1e-15)
    assert abs(itr - ito) <= 1 and rel(Xo, Xr) < 1e-9
    Xr, itr, _ = r.BCG(U, B, mass, 1e-10)
    Xo, ito, _ = oracle.BCG(U, B, mass, 1e-10)
    assert abs(itr - ito) <= 1 and rel(Xo, Xr) < 1e-9


@pytest.mark.parametrize("V", [1, 2, 3, 5, 17, 101])
def test_oracle_vs_live_reference_ragged(oracle, V):
    """Edge volumes (V = 1, 2: every neighbour is a periodic image; odd and prime V) against the
    unmodified reference: operator, Gram, BCGrQ and the multishift solver."""
    for N in (1, 3, 4, 12):
        if not RefShim.available(N):
            pytest.skip("oracle/_ref not built (no /root/reference here)")
        r = RefShim(N)
        mass = 0.4
        U, B = r.make_inputs(V, 3 + V)
        assert rel(oracle.op(U, B, mass), r.op(U, B, mass)) < 1e-14
        assert rel(oracle.hermitian_dot(B, oracle.op(U, B, mass)), r.hermitian_dot(B, r.op(U, B, mass))) < 1e-13
        if 3 * V % N != 0 and 3 * V < 10 * N:
            # the block Krylov space runs out of dimensions in mid-iteration (3V not a multiple of N): P^dag A P
            # turns singular and the reference breaks down silently into NaNs (V=2, N=4: 395 iterations of them)
            continue
        sig = [0.0, 0.3]
        Xr, itr, _ = r.SBCGrQ(U, B, mass, sig, 1e-10, 1e-15)
        Xo, ito, _, _ = oracle.SBCGrQ(U, B, mass, sig, 1e-10, 1e-15)
        assert abs(itr - ito) <= 1 and rel(Xo, Xr) < 1e-9, (V, N, itr, ito)


def test_4d_extension_properties(oracle):
    """The 4-D extension of the operator has no reference counterpart (its oracle is 'parity
    unpinned'); what pins it are the properties of the construction: D anti-Hermitian for any
    links, m^2 - D^2 Hermitian positive definite with spectrum >= m^2, a 1 x 1 x 1 lattice in
    three directions degenerating to independent chains, and convergence of the solvers."""
    rng = np.random.default_rng(11)
    dims, N, mass = (4, 3, 2, 5), 2, 0.3
    V = int(np.prod(dims))
    U = rng.uniform(-1, 1, (V, 4, 3, 3)) + 1j * rng.uniform(-1, 1, (V, 4, 3, 3))
    a = rng.standard_normal((V, N, 3)) + 1j * rng.standard_normal((V, N, 3))
    b = rng.standard_normal((V, N, 3)) + 1j * rng.standard_normal((V, N, 3))
    oracle.set_lattice(dims)
    try:
        Da, Db = oracle.D(U, a), oracle.D(U, b)
        assert abs(np.vdot(a, Db) + np.vdot(Da, b)) < 1e-12 * abs(np.vdot(a, Db))
        Aa, Ab = oracle.op(U, a, mass), oracle.op(U, b, mass)
        assert abs(np.vdot(b, Aa) - np.vdot(Ab, a)) < 1e-12 * abs(np.vdot(b, Aa))
        assert np.vdot(a, Aa).real >= mass * mass * np.vdot(a, a).real
        shifts = [0.0, 0.1]
        Xs, it, _, _ = oracle.SBCGrQ(U, a, mass, shifts, 1e-10, 1e-15)
        for s, sig in enumerate(shifts):
            assert oracle.true_residual(U, a, Xs[s], mass, sig).max() < 2e-10
        # along one direction only, with the other three links zero, D is the reference's chain
        V1 = 7
        oracle.set_lattice((V1, 1, 1, 1))
        U1 = rng.uniform(-1, 1, (V1, 3, 3)) + 1j * rng.uniform(-1, 1, (V1, 3, 3))
        U4 = np.zeros((V1, 4, 3, 3), np.complex128)
        U4[:, 0] = U1
        x = rng.standard_normal((V1, N, 3)) + 1j * rng.standard_normal((V1, N, 3))
        D4x = oracle.D(U4, x)
        oracle.set_lattice(None)
        assert np.abs(D4x - oracle.D(U1, x)).max() < 1e-14
    finally:
        oracle.set_lattice(None)


def test_counter_uniform_generator():
    """Restatement of the device input generator: range, moments, independence of the split."""
    import oracle.pyoracle as ora
    a = ora.counter_uniform(7, 1, 0, 200000)
    assert a.min() >= -1.0 and a.max() < 1.0
    assert abs(a.mean()) < 5e-3 and abs(a.var() - 1.0 / 3.0) < 5e-3
    assert abs(np.corrcoef(a[:-1], a[1:])[0, 1]) < 1e-2
    # a slab is a slice of the global array; streams and seeds differ
    assert np.array_equal(ora.counter_uniform(7, 1, 1234, 100), a[1234:1334])
    assert not np.array_equal(ora.counter_uniform(7, 0, 0, 100), a[:100])
    assert not np.array_equal(ora.counter_uniform(8, 1, 0, 100), a[:100])
    assert np.all(a * 2.0 ** 52 == np.round(a * 2.0 ** 52))  # multiples of 2^-52: exact on any machine
    # SplitMix64 known answer: seed 0, first output of the published generator is 0xE220A8397B1DCDAF
    z = np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    assert int(z) == 0xE220A8397B1DCDAF
    assert ora.counter_uniform(0, 0, 0, 1)[0] == float(int(z) >> 11) * 2.0 ** -52 - 1.0



@pytest.mark.parametrize("name", ["scalar_V128.npz", "scalar_V200.npz"])
def test_oracle_cg_scg_vs_golden(oracle, name):
    """CG / SCG restatement (oracle.cpp: ora_CG / ora_SCG) against the unmodified reference's own CG / SCG
    (src/standard_solvers.cpp, linked into oracle/_ref/libref_n1.so; fixtures from gen_golden.py)."""
    g = golden(name)
    U, b, mass, eps = g["U"], g["B"], float(g["mass"]), float(g["eps"])
    x, it = oracle.CG(U, b, mass, eps)
    assert it == int(g["it_cg"])
    assert np.abs(x - g["X_cg"]).max() / np.abs(g["X_cg"]).max() < 1e-12
    xs, it = oracle.SCG(U, b, mass, list(g["shifts"]), eps, float(g["eps_shifts"]))
    assert it == int(g["it_scg"])
    for s in range(len(g["shifts"])):
        assert np.abs(xs[s] - g["X_scg"][s]).max() / np.abs(g["X_scg"][s]).max() < 1e-12
        assert oracle.true_residual(U, b, xs[s], mass, float(g["shifts"][s])).max() < 2 * eps
