"""Python mirror of the reference's host interface for the hot path.

Same names, argument meaning and return value as inc/block_solvers.hpp:
    BCG(X, B, D, eps=1e-15, max_iterations=1e6)            -> iterations   (:10-45)
    BCGrQ(X, B, D, eps=1e-15, max_iterations=1e6)          -> iterations   (:50-86)
    SBCGrQ(X, B, D, sigma, eps=1e-15, eps_shifts=1e-15, max_iterations=1e6) (:91-185)
and of inc/standard_solvers.hpp (CG, SCG for one right-hand side; src/standard_solvers.cpp:3-95).
`X` is overwritten (solvers start from 0), `B`, `D` are not modified, and the
return value is the number of operator applications.  Fields are numpy
complex128 arrays of shape (V, N, 3) in the reference's memory order; the
links of `dirac_op` are shape (V, 3, 3) with U[x, j, i] = U_x(i, j)
(column-major 3x3, inc/dirac_op.hpp:10-11).

Everything numeric happens in the CUDA library through the C-ABI; nothing
here computes on the CPU.
"""
import numpy as np

from .capi import Context

_ctx_cache = {}


def _context(V, N, S, device=0):
    key = (V, N, device)
    ctx = _ctx_cache.get(key)
    if ctx is None or ctx.S < S:
        if ctx is not None:
            ctx.close()
        ctx = Context(V, N, max_shifts=max(S, 1), device=device)
        _ctx_cache[key] = ctx
    return ctx


def release_contexts():
    """Free the cached device contexts (device memory of the last solves)."""
    for ctx in _ctx_cache.values():
        ctx.close()
    _ctx_cache.clear()


class block_fermion_field(np.ndarray):
    """(V, N, 3) complex128 array; `block_fermion_field(V, N)` as in inc/fields.hpp:25-38."""

    def __new__(cls, V, N_rhs=1):
        obj = np.zeros((V, N_rhs, 3), np.complex128).view(cls)
        return obj

    @property
    def V(self):
        return self.shape[0]


class dirac_op:
    """inc/dirac_op.hpp:8-44: public V, mass; links are public here."""

    def __init__(self, V, mass=0.1, links=None, rng=None):
        self.V = int(V)
        self.mass = float(mass)
        if links is None:
            # random 3x3 complex links, entries uniform in [-1,1]+i[-1,1] (dirac_op.hpp:27-32)
            rng = rng or np.random.default_rng()
            links = rng.uniform(-1, 1, (V, 3, 3)) + 1j * rng.uniform(-1, 1, (V, 3, 3))
        self.links = np.ascontiguousarray(links, dtype=np.complex128)
        assert self.links.shape == (self.V, 3, 3)

    def op(self, lhs, rhs, device=0):
        """lhs = (m^2 - D^2) rhs"""
        V, N, _ = rhs.shape
        ctx = _context(V, N, 1, device)
        ctx.set_links(self.links, self.mass)
        hi, ho = ctx.field(rhs), ctx.field()
        try:
            ctx.op(ho, hi)
            ctx.download(ho, lhs)
        finally:
            ctx.free(hi)
            ctx.free(ho)


def _check(X, B):
    if not (isinstance(B, np.ndarray) and B.dtype == np.complex128 and B.ndim == 3 and B.shape[2] == 3):
        raise TypeError("B must be a complex128 array of shape (V, N, 3)")
    if X.shape != B.shape or X.dtype != np.complex128 or not X.flags["C_CONTIGUOUS"]:
        raise TypeError("X must be a C-contiguous complex128 array shaped like B")


def BCG(X, B, D, eps=1e-15, max_iterations=int(1e6), device=0, info=None):
    _check(X, B)
    ctx = _context(B.shape[0], B.shape[1], 1, device)
    ctx.set_links(D.links, D.mass)
    r = ctx.solve_bcg(X, np.ascontiguousarray(B), eps, int(max_iterations))
    if info is not None:
        info.update(r.as_dict())
    return r.iterations


def BCGrQ(X, B, D, eps=1e-15, max_iterations=int(1e6), device=0, info=None):
    _check(X, B)
    ctx = _context(B.shape[0], B.shape[1], 1, device)
    ctx.set_links(D.links, D.mass)
    r = ctx.solve_bcgrq(X, np.ascontiguousarray(B), eps, int(max_iterations))
    if info is not None:
        info.update(r.as_dict())
    return r.iterations


def SBCGrQ(X, B, D, sigma, eps=1e-15, eps_shifts=1e-15, max_iterations=int(1e6), device=0, info=None):
    if len(X) != len(sigma):
        raise ValueError("number of shifts does not match number of solution vectors")
    for x in X:
        _check(x, B)
    ctx = _context(B.shape[0], B.shape[1], len(sigma), device)
    ctx.set_links(D.links, D.mass)
    r = ctx.solve_sbcgrq(list(X), np.ascontiguousarray(B), sigma, eps, eps_shifts, int(max_iterations))
    if info is not None:
        info.update(r.as_dict())
    return r.iterations


def CG(x, b, D, eps=1e-15, max_iterations=int(1e6), device=0, info=None):
    """src/standard_solvers.cpp:3-32 (inc/standard_solvers.hpp:10-13): fields of shape (V, 1, 3)."""
    _check(x, b)
    if b.shape[1] != 1:
        raise TypeError("CG takes one right-hand side: b must have shape (V, 1, 3)")
    ctx = _context(b.shape[0], 1, 1, device)
    ctx.set_links(D.links, D.mass)
    r = ctx.solve_cg(x, np.ascontiguousarray(b), eps, int(max_iterations))
    if info is not None:
        info.update(r.as_dict())
    return r.iterations


def SCG(x, b, D, sigma, eps=1e-15, eps_shifts=1e-15, max_iterations=int(1e6), device=0, info=None):
    """src/standard_solvers.cpp:34-95 (inc/standard_solvers.hpp:15-20)."""
    if len(x) != len(sigma):
        raise ValueError("number of shifts does not match number of solution vectors")
    for xs in x:
        _check(xs, b)
    if b.shape[1] != 1:
        raise TypeError("SCG takes one right-hand side: b must have shape (V, 1, 3)")
    ctx = _context(b.shape[0], 1, len(sigma), device)
    ctx.set_links(D.links, D.mass)
    r = ctx.solve_scg(list(x), np.ascontiguousarray(b), sigma, eps, eps_shifts, int(max_iterations))
    if info is not None:
        info.update(r.as_dict())
    return r.iterations
