"""TEST INFRASTRUCTURE ONLY: ctypes front-ends for the two CPU checkers.

* ``Oracle``  -> oracle/liboracle.so   (our restatement, oracle.cpp; runtime N)
* ``RefShim`` -> oracle/_ref/libref_n<N>.so (the unmodified reference compiled
  from /root/reference by oracle/Makefile; one library per compile-time N)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package never does.

numpy conventions (all complex128):
  field  F[x, r, c]   shape (V, N, 3)  == reference memory order [V][N][3]
  links  U[x, j, i]   shape (V, 3, 3)  == column-major 3x3 per site, so the
                                          mathematical U_x(i, j) is U[x, j, i]
  matrix M[i, j]      shape (N, N) logical; converted to column-major at the
                                          boundary.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)


def _p(a):
    assert a.dtype == np.complex128 or a.dtype == np.float64
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def mat_to_cm(M):
    """logical (N,N) -> column-major buffer"""
    return np.ascontiguousarray(np.asarray(M, dtype=np.complex128).T)


def mat_from_cm(buf, N):
    return np.ascontiguousarray(buf.reshape(N, N).T)


def build(force=False):
    """(Re)build liboracle.so and, when /root/reference exists, oracle/_ref."""
    if force or not os.path.exists(os.path.join(_HERE, "liboracle.so")) or os.path.exists("/root/reference/inc"):
        subprocess.run(["make", "-s", "-j8", "-C", _HERE, "all"], check=True)


class Oracle:
    def __init__(self):
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        L = self.lib
        L.ora_BCG.restype = C.c_int
        L.ora_BCGrQ.restype = C.c_int
        L.ora_SBCGrQ.restype = C.c_int
        L.ora_llt_upper.restype = C.c_int

    def set_lattice(self, dims=None):
        """dims = (L0, L1, L2, L3): every later call applies the 4-D extension of the operator
        (links (V, 4, 3, 3)); None: back to the reference's 1-D chain."""
        if dims is None:
            self.lib.ora_set_lattice(None)
        else:
            self.lib.ora_set_lattice((C.c_int * 4)(*[int(d) for d in dims]))

    def make_inputs(self, V, N, seed=1):
        U = np.empty((V, 3, 3), np.complex128)
        B = np.empty((V, N, 3), np.complex128)
        self.lib.ora_make_inputs(C.c_int(V), C.c_int(N), C.c_uint(seed), _p(U), _p(B))
        return U, B

    def op(self, U, x, mass, sigma=0.0):
        V, N, _ = x.shape
        out = np.empty_like(x)
        self.lib.ora_op(C.c_int(V), C.c_int(N), C.c_double(mass), _p(U), _p(x), _p(out), C.c_double(sigma))
        return out

    def D(self, U, x):
        V, N, _ = x.shape
        out = np.empty_like(x)
        self.lib.ora_D(C.c_int(V), C.c_int(N), _p(U), _p(x), _p(out))
        return out

    def hermitian_dot(self, a, b, chunk=0):
        V, N, _ = a.shape
        R = np.empty((N, N), np.complex128)
        self.lib.ora_hermitian_dot(C.c_int(V), C.c_int(N), _p(a), _p(b), _p(R), C.c_int(chunk))
        return mat_from_cm(R, N)

    def add(self, dst, src, M):
        V, N, _ = dst.shape
        out = dst.copy()
        if np.isscalar(M):
            self.lib.ora_add_scalar(C.c_int(V), C.c_int(N), _p(out), _p(src), C.c_double(M))
        else:
            self.lib.ora_add(C.c_int(V), C.c_int(N), _p(out), _p(src), _p(mat_to_cm(M)))
        return out

    def rescale_add(self, dst, L, src, r):
        V, N, _ = dst.shape
        out = dst.copy()
        self.lib.ora_rescale_add(C.c_int(V), C.c_int(N), _p(out), _p(mat_to_cm(L)), _p(src), C.c_double(r))
        return out

    def thinQR(self, q, chunk=0):
        V, N, _ = q.shape
        out = q.copy()
        R = np.empty((N, N), np.complex128)
        self.lib.ora_thinQR(C.c_int(V), C.c_int(N), _p(out), _p(R), C.c_int(chunk))
        return out, mat_from_cm(R, N)

    def llt_upper(self, A):
        N = A.shape[0]
        R = np.empty((N, N), np.complex128)
        info = self.lib.ora_llt_upper(C.c_int(N), _p(mat_to_cm(A)), _p(R))
        return mat_from_cm(R, N), info

    def fullpivlu_inverse(self, A):
        N = A.shape[0]
        X = np.empty((N, N), np.complex128)
        self.lib.ora_fullpivlu_inverse(C.c_int(N), _p(mat_to_cm(A)), _p(X))
        return mat_from_cm(X, N)

    def fullpivlu_solve(self, A, B):
        N = A.shape[0]
        X = np.empty((N, N), np.complex128)
        self.lib.ora_fullpivlu_solve(C.c_int(N), _p(mat_to_cm(A)), _p(mat_to_cm(B)), _p(X))
        return mat_from_cm(X, N)

    def _solve(self, fn, U, B, mass, eps, max_it, chunk):
        V, N, _ = B.shape
        X = np.empty_like(B)
        sec = C.c_double(0)
        it = fn(C.c_int(V), C.c_int(N), C.c_double(mass), _p(U), _p(B), _p(X), C.c_double(eps),
                C.c_int(max_it), C.c_int(chunk), C.byref(sec))
        return X, it, sec.value

    def BCG(self, U, B, mass, eps=1e-15, max_it=1000000, chunk=0):
        return self._solve(self.lib.ora_BCG, U, B, mass, eps, max_it, chunk)

    def BCGrQ(self, U, B, mass, eps=1e-15, max_it=1000000, chunk=0):
        return self._solve(self.lib.ora_BCGrQ, U, B, mass, eps, max_it, chunk)

    def SBCGrQ(self, U, B, mass, sigma, eps=1e-15, eps_shifts=1e-15, max_it=1000000, chunk=0):
        V, N, _ = B.shape
        sig = np.ascontiguousarray(sigma, dtype=np.float64)
        S = len(sig)
        X = np.empty((S, V, N, 3), np.complex128)
        sec = C.c_double(0)
        nun = C.c_int(0)
        it = self.lib.ora_SBCGrQ(C.c_int(V), C.c_int(N), C.c_double(mass), _p(U), _p(B), _p(X), _p(sig),
                                 C.c_int(S), C.c_double(eps), C.c_double(eps_shifts), C.c_int(max_it),
                                 C.c_int(chunk), C.byref(sec), C.byref(nun))
        return X, it, sec.value, nun.value

    def CG(self, U, b, mass, eps=1e-15, max_it=1000000):
        V = b.shape[0]
        assert b.shape == (V, 1, 3)
        x = np.empty_like(b)
        self.lib.ora_CG.restype = C.c_int
        it = self.lib.ora_CG(C.c_int(V), C.c_double(mass), _p(U), _p(b), _p(x), C.c_double(eps), C.c_int(max_it))
        return x, it

    def SCG(self, U, b, mass, sigma, eps=1e-15, eps_shifts=1e-15, max_it=1000000):
        V = b.shape[0]
        assert b.shape == (V, 1, 3)
        sig = np.ascontiguousarray(sigma, dtype=np.float64)
        x = np.empty((len(sig), V, 1, 3), np.complex128)
        self.lib.ora_SCG.restype = C.c_int
        it = self.lib.ora_SCG(C.c_int(V), C.c_double(mass), _p(U), _p(b), _p(x), _p(sig), C.c_int(len(sig)),
                              C.c_double(eps), C.c_double(eps_shifts), C.c_int(max_it))
        return x, it

    def true_residual(self, U, B, X, mass, sigma=0.0):
        V, N, _ = B.shape
        out = np.empty(N, np.float64)
        self.lib.ora_true_residual(C.c_int(V), C.c_int(N), C.c_double(mass), _p(U), _p(B), _p(X),
                                   C.c_double(sigma), _p(out))
        return out


class RefShim:
    """The unmodified reference, one shared library per compile-time N."""

    @staticmethod
    def available(N):
        return os.path.exists(os.path.join(_HERE, "_ref", "libref_n%d.so" % N))

    def __init__(self, N):
        self.N = N
        path = os.path.join(_HERE, "_ref", "libref_n%d.so" % N)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        for nm in ("ref_BCG", "ref_BCGrQ", "ref_SBCGrQ"):
            self._f(nm).restype = C.c_int
        for nm in ("ref_op", "ref_hermitian_dot"):
            self._f(nm).restype = C.c_double

    def _f(self, name):
        return getattr(self.lib, "%s_%d" % (name, self.N))

    def make_inputs(self, V, seed=1):
        U = np.empty((V, 3, 3), np.complex128)
        B = np.empty((V, self.N, 3), np.complex128)
        self._f("ref_make_inputs")(C.c_int(V), C.c_uint(seed), _p(U), _p(B))
        return U, B

    def op(self, U, x, mass, reps=1, want_time=False):
        V = x.shape[0]
        out = np.empty_like(x)
        t = self._f("ref_op")(C.c_int(V), C.c_double(mass), _p(U), _p(x), _p(out), C.c_int(reps))
        return (out, t) if want_time else out

    def hermitian_dot(self, a, b, reps=1, want_time=False):
        V = a.shape[0]
        R = np.empty((self.N, self.N), np.complex128)
        t = self._f("ref_hermitian_dot")(C.c_int(V), _p(a), _p(b), _p(R), C.c_int(reps))
        R = mat_from_cm(R, self.N)
        return (R, t) if want_time else R

    def add(self, dst, src, M):
        V = dst.shape[0]
        out = dst.copy()
        if np.isscalar(M):
            self._f("ref_add_scalar")(C.c_int(V), _p(out), _p(src), C.c_double(M))
        else:
            self._f("ref_add")(C.c_int(V), _p(out), _p(src), _p(mat_to_cm(M)))
        return out

    def rescale_add(self, dst, L, src, r):
        V = dst.shape[0]
        out = dst.copy()
        self._f("ref_rescale_add")(C.c_int(V), _p(out), _p(mat_to_cm(L)), _p(src), C.c_double(r))
        return out

    def thinQR(self, q):
        V = q.shape[0]
        out = q.copy()
        R = np.empty((self.N, self.N), np.complex128)
        self._f("ref_thinQR")(C.c_int(V), _p(out), _p(R))
        return out, mat_from_cm(R, self.N)

    def llt_upper(self, A):
        R = np.empty((self.N, self.N), np.complex128)
        self._f("ref_llt_upper")(_p(mat_to_cm(A)), _p(R))
        return mat_from_cm(R, self.N)

    def fullpivlu_inverse(self, A):
        X = np.empty((self.N, self.N), np.complex128)
        self._f("ref_fullpivlu_inverse")(_p(mat_to_cm(A)), _p(X))
        return mat_from_cm(X, self.N)

    def fullpivlu_solve(self, A, B):
        X = np.empty((self.N, self.N), np.complex128)
        self._f("ref_fullpivlu_solve")(_p(mat_to_cm(A)), _p(mat_to_cm(B)), _p(X))
        return mat_from_cm(X, self.N)

    def _solve(self, name, U, B, mass, eps, max_it):
        V = B.shape[0]
        X = np.empty_like(B)
        sec = C.c_double(0)
        it = self._f(name)(C.c_int(V), C.c_double(mass), _p(U), _p(B), _p(X), C.c_double(eps), C.c_int(max_it),
                           C.byref(sec))
        return X, it, sec.value

    def BCG(self, U, B, mass, eps=1e-15, max_it=1000000):
        return self._solve("ref_BCG", U, B, mass, eps, max_it)

    def BCGrQ(self, U, B, mass, eps=1e-15, max_it=1000000):
        return self._solve("ref_BCGrQ", U, B, mass, eps, max_it)

    def SBCGrQ(self, U, B, mass, sigma, eps=1e-15, eps_shifts=1e-15, max_it=1000000):
        V = B.shape[0]
        sig = np.ascontiguousarray(sigma, dtype=np.float64)
        S = len(sig)
        X = np.empty((S, V, self.N, 3), np.complex128)
        sec = C.c_double(0)
        it = self._f("ref_SBCGrQ")(C.c_int(V), C.c_double(mass), _p(U), _p(B), _p(X), _p(sig), C.c_int(S),
                                   C.c_double(eps), C.c_double(eps_shifts), C.c_int(max_it), C.byref(sec))
        return X, it, sec.value


    def CG(self, U, b, mass, eps=1e-15, max_it=1000000):
        assert self.N == 1
        V = b.shape[0]
        x = np.empty_like(b)
        sec = C.c_double(0)
        self.lib.ref_CG_1.restype = C.c_int
        it = self.lib.ref_CG_1(C.c_int(V), C.c_double(mass), _p(U), _p(b), _p(x), C.c_double(eps), C.c_int(max_it),
                               C.byref(sec))
        return x, it, sec.value

    def SCG(self, U, b, mass, sigma, eps=1e-15, eps_shifts=1e-15, max_it=1000000):
        assert self.N == 1
        V = b.shape[0]
        sig = np.ascontiguousarray(sigma, dtype=np.float64)
        x = np.empty((len(sig), V, 1, 3), np.complex128)
        sec = C.c_double(0)
        self.lib.ref_SCG_1.restype = C.c_int
        it = self.lib.ref_SCG_1(C.c_int(V), C.c_double(mass), _p(U), _p(b), _p(x), _p(sig), C.c_int(len(sig)),
                                C.c_double(eps), C.c_double(eps_shifts), C.c_int(max_it), C.byref(sec))
        return x, it, sec.value


def counter_uniform(seed, stream, first, n):
    """numpy restatement of the device input generator (blockcg_b200/csrc/field_kernels.cuh,
    counter_uniform): doubles number first .. first+n-1 of the global array, uniform in [-1, 1) --
    the distribution of the reference's setRandom (inc/dirac_op.hpp:24-32, benchmark.cpp:60-62)."""
    m = (1 << 64) - 1
    idx = np.arange(first + 1, first + n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & m) + np.uint64(0x9E3779B97F4A7C15) * idx + np.uint64((0xD1B54A32D192ED03 * stream) & m)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -52 - 1.0

