"""ctypes binding of include/blockcg_b200.h (libblockcg_b200.so).

This is plumbing only: every call goes straight into the CUDA library.  There
is no CPU fallback; loading fails loudly if the extension is missing, and
context creation fails loudly if there is no sm_100 device.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libblockcg_b200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

MAX_SHIFTS = 32
UNIQUE_ID_BYTES = 128
IPC_HANDLE_BYTES = 64

STATUS = {0: "BCG_OK", 1: "BCG_ERR_INVALID", 2: "BCG_ERR_CUDA", 3: "BCG_ERR_NOT_PD", 4: "BCG_ERR_NCCL",
          5: "BCG_ERR_NO_COMM", 6: "BCG_ERR_NAN"}

# every symbol include/blockcg_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "bcg_version", "bcg_supports_nrhs", "bcg_ctx_create", "bcg_ctx_destroy", "bcg_last_error",
    "bcg_ctx_create_4d", "bcg_set_links_4d", "bcg_comm_get_unique_id", "bcg_comm_init", "bcg_comm_ipc_handle", "bcg_comm_ipc_open", "bcg_comm_ipc_disable", "bcg_set_links", "bcg_field_alloc", "bcg_field_free",
    "bcg_field_upload", "bcg_field_download", "bcg_field_random", "bcg_set_links_random", "bcg_field_zero", "bcg_field_copy", "bcg_op", "bcg_gram",
    "bcg_add", "bcg_add_scalar", "bcg_rescale_add", "bcg_trsm", "bcg_thinqr", "bcg_true_residual",
    "bcg_solve_bcg_dev", "bcg_solve_bcgrq_dev", "bcg_solve_sbcgrq_dev", "bcg_solve_bcg", "bcg_solve_bcgrq",
    "bcg_solve_sbcgrq", "bcg_bench_kernel", "bcg_solve_cg_dev", "bcg_solve_scg_dev", "bcg_solve_cg", "bcg_solve_scg",
    "bcg_last_solve_stats", "bcg_small_inverse", "bcg_small_lu_solve", "bcg_set_loop_profile", "bcg_get_loop_profile", "bcg_shift_schedule", "bcg_stag_schedule",
]


class SolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("residual", C.c_double), ("n_unconverged", C.c_int),
                ("solve_ms", C.c_double), ("setup_ms", C.c_double), ("kernel_launches", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SolveStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("n_shifts", C.c_int), ("paired", C.c_int), ("depth", C.c_int),
                ("active_hist", C.c_uint32 * (MAX_SHIFTS + 1)), ("shift_update_field_passes", C.c_uint64),
                ("resid_shift", C.c_double * MAX_SHIFTS)]


class LoopProfile(C.Structure):
    _fields_ = [("iterations", C.c_int), ("first_iteration", C.c_int), ("active_systems", C.c_int),
                ("ms", C.c_double * 8)]


class BcgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (STATUS.get(code, code), msg))
        self.code = code


_lib = None


def load():
    """Load the CUDA extension (torch first, so its NCCL/CUDA libraries are the ones mapped)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not built: run `make -C blockcg_b200/csrc` (or __graft_entry__.build())" % LIB_PATH)
    try:  # plumbing only: makes the process use the torch-bundled libnccl.so.2 when torch is around
        import torch  # noqa: F401
    except Exception:  # pragma: no cover
        pass
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.bcg_version.restype = C.c_char_p
    lib.bcg_last_error.restype = C.c_char_p
    lib.bcg_last_error.argtypes = [C.c_void_p]
    lib.bcg_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.bcg_ctx_destroy.argtypes = [C.c_void_p]
    lib.bcg_comm_get_unique_id.argtypes = [C.c_void_p]
    lib.bcg_comm_init.argtypes = [C.c_void_p, C.c_void_p]
    lib.bcg_comm_ipc_handle.argtypes = [C.c_void_p, C.c_void_p]
    lib.bcg_comm_ipc_open.argtypes = [C.c_void_p, C.c_void_p]
    lib.bcg_comm_ipc_disable.argtypes = [C.c_void_p]
    lib.bcg_set_links.argtypes = [C.c_void_p, _dp, C.c_double]
    lib.bcg_set_links_4d.argtypes = [C.c_void_p, _dp, C.c_double]
    lib.bcg_ctx_create_4d.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int]
    lib.bcg_field_alloc.argtypes = [C.c_void_p, _ip]
    lib.bcg_field_free.argtypes = [C.c_void_p, C.c_int]
    lib.bcg_field_upload.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.bcg_field_download.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.bcg_field_zero.argtypes = [C.c_void_p, C.c_int]
    lib.bcg_field_random.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
    lib.bcg_set_links_random.argtypes = [C.c_void_p, C.c_uint64, C.c_double]
    lib.bcg_field_copy.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.bcg_op.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, _dp]
    lib.bcg_gram.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    lib.bcg_add.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    lib.bcg_add_scalar.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
    lib.bcg_rescale_add.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_double]
    lib.bcg_trsm.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.bcg_thinqr.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.bcg_true_residual.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, _dp]
    lib.bcg_solve_bcg_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(SolveInfo)]
    lib.bcg_solve_bcgrq_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(SolveInfo)]
    lib.bcg_solve_sbcgrq_dev.argtypes = [C.c_void_p, _ip, C.c_int, _dp, C.c_int, C.c_double, C.c_double, C.c_int,
                                         C.POINTER(SolveInfo)]
    lib.bcg_solve_bcg.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_int, C.POINTER(SolveInfo)]
    lib.bcg_solve_bcgrq.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_int, C.POINTER(SolveInfo)]
    lib.bcg_solve_sbcgrq.argtypes = [C.c_void_p, C.POINTER(_dp), _dp, _dp, C.c_int, C.c_double, C.c_double,
                                     C.c_int, C.POINTER(SolveInfo)]
    lib.bcg_solve_cg_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(SolveInfo)]
    lib.bcg_solve_scg_dev.argtypes = [C.c_void_p, _ip, C.c_int, _dp, C.c_int, C.c_double, C.c_double, C.c_int,
                                      C.POINTER(SolveInfo)]
    lib.bcg_solve_cg.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_int, C.POINTER(SolveInfo)]
    lib.bcg_solve_scg.argtypes = [C.c_void_p, C.POINTER(_dp), _dp, _dp, C.c_int, C.c_double, C.c_double,
                                  C.c_int, C.POINTER(SolveInfo)]
    lib.bcg_last_solve_stats.argtypes = [C.c_void_p, C.POINTER(SolveStats)]
    lib.bcg_set_loop_profile.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.bcg_get_loop_profile.argtypes = [C.c_void_p, C.POINTER(LoopProfile)]
    lib.bcg_shift_schedule.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _ip]
    lib.bcg_stag_schedule.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, _ip, _ip]
    lib.bcg_small_inverse.argtypes = [C.c_void_p, _dp, _dp, C.c_int, _ip]
    lib.bcg_small_lu_solve.argtypes = [C.c_void_p, _dp, _dp, _dp]
    lib.bcg_bench_kernel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _ip, C.c_int, _dp,
                                     C.POINTER(C.c_int64)]
    _lib = lib
    return lib


def _dptr(a):
    assert a.flags["C_CONTIGUOUS"] and a.dtype in (np.complex128, np.float64), (a.dtype, a.flags)
    return a.ctypes.data_as(_dp)


def mat_to_cm(M):
    return np.ascontiguousarray(np.asarray(M, dtype=np.complex128).T)


def mat_from_cm(buf, N):
    return np.ascontiguousarray(buf.reshape(N, N).T)


def shift_schedule(schedule, iteration, stop, n_active, n_active_prev):
    """Items of one launch of the multishift update: [(kind, system)], field passes (host-side, no device)."""
    lib = load()
    kinds, systems = (C.c_int * (MAX_SHIFTS + 2))(), (C.c_int * (MAX_SHIFTS + 2))()
    passes = C.c_int(0)
    n = lib.bcg_shift_schedule(schedule, iteration, 1 if stop else 0, n_active, n_active_prev, kinds, systems, C.byref(passes))
    return [(kinds[i], systems[i]) for i in range(n)], passes.value


def stag_schedule(depth, iteration, stop, n_active, n_active_ring, ring=None, part=0):
    """Items of one launch of the depth-`depth` staggered update (schedule 3): [(system, first_back, n_updates)],
    field passes (host-side, no device).  n_active_ring[j % ring] = active systems of the earlier iteration j;
    part = 0 whole launch, 1 / 2 the two launches of the overlapped variant (ring = depth + 1)."""
    lib = load()
    n_max = MAX_SHIFTS + 2
    systems, first, count = (C.c_int * n_max)(), (C.c_int * n_max)(), (C.c_int * n_max)()
    ringv = (C.c_int * 4)(*([int(v) for v in n_active_ring] + [0] * 4)[:4])
    passes = C.c_int(0)
    n = lib.bcg_stag_schedule(depth, ring or depth, part, iteration, 1 if stop else 0, n_active, ringv, systems, first,
                              count, C.byref(passes))
    assert n >= 0, "bad argument to bcg_stag_schedule"
    return [(systems[i], first[i], count[i]) for i in range(n)], passes.value


class Context:
    """One GPU, one slab of `v_local` sites, `n_rhs` right-hand sides."""

    def __init__(self, v_local, n_rhs, max_shifts=1, device=0, rank=0, nranks=1, dims=None):
        """dims = (L0, L1, L2, L3_local): a context of the 4-D extension of the operator (v_local is
        then ignored and set to the product); default: the reference's 1-D chain of v_local sites."""
        self.lib = load()
        self.dims = tuple(int(d) for d in dims) if dims is not None else None
        if self.dims is not None:
            v_local = int(np.prod(self.dims))
        self.V, self.N, self.S = int(v_local), int(n_rhs), int(max_shifts)
        self.rank, self.nranks = rank, nranks
        self._h = C.c_void_p()
        if self.dims is not None:
            rc = self.lib.bcg_ctx_create_4d(C.byref(self._h), (C.c_int64 * 4)(*self.dims), self.N, self.S, device,
                                            rank, nranks)
        else:
            rc = self.lib.bcg_ctx_create(C.byref(self._h), self.V, self.N, self.S, device, rank, nranks)
        if rc:
            msg = self.lib.bcg_last_error(self._h).decode() if self._h else "context creation failed"
            self.lib.bcg_ctx_destroy(self._h)
            self._h = None
            raise BcgError(rc, msg)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.bcg_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc:
            raise BcgError(rc, self.lib.bcg_last_error(self._h).decode())

    # ---- multi-GPU plumbing ----
    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(UNIQUE_ID_BYTES)
        rc = load().bcg_comm_get_unique_id(buf)
        if rc:
            raise BcgError(rc, "ncclGetUniqueId failed")
        return bytes(buf.raw)

    def ipc_handle(self):
        """CUDA-IPC handle of this rank's communication buffer (peer-memory exchange over NVLink)."""
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._ck(self.lib.bcg_comm_ipc_handle(self._h, buf))
        return buf.raw

    def ipc_open(self, handles):
        """handles: the ipc_handle() of every rank, concatenated in rank order."""
        buf = C.create_string_buffer(bytes(handles), len(handles))
        self._ck(self.lib.bcg_comm_ipc_open(self._h, buf))

    def ipc_disable(self):
        """Forget the peer mappings: the loop goes back to the NCCL exchange."""
        self._ck(self.lib.bcg_comm_ipc_disable(self._h))

    def comm_init(self, uid):
        buf = C.create_string_buffer(bytes(uid), UNIQUE_ID_BYTES)
        self._ck(self.lib.bcg_comm_init(self._h, buf))

    # ---- operator / fields ----
    def set_links(self, U, mass):
        U = np.ascontiguousarray(U, dtype=np.complex128)
        if self.dims is not None:
            assert U.shape == (self.V, 4, 3, 3), U.shape
            self._ck(self.lib.bcg_set_links_4d(self._h, _dptr(U), float(mass)))
            return
        assert U.shape == (self.V, 3, 3), U.shape
        self._ck(self.lib.bcg_set_links(self._h, _dptr(U), float(mass)))

    def set_links_random(self, seed, mass):
        """Links drawn on the device (uniform in [-1, 1), counter-based: independent of the rank count)."""
        self._ck(self.lib.bcg_set_links_random(self._h, int(seed), float(mass)))

    def field_random(self, h, seed):
        self._ck(self.lib.bcg_field_random(self._h, h, int(seed)))

    def field(self, data=None):
        h = C.c_int(-1)
        self._ck(self.lib.bcg_field_alloc(self._h, C.byref(h)))
        if data is not None:
            self.upload(h.value, data)
        return h.value

    def free(self, h):
        self._ck(self.lib.bcg_field_free(self._h, h))

    def upload(self, h, data):
        data = np.ascontiguousarray(data, dtype=np.complex128)
        assert data.shape == (self.V, self.N, 3), data.shape
        self._ck(self.lib.bcg_field_upload(self._h, h, _dptr(data)))

    def download(self, h, out=None):
        if out is None:
            out = np.empty((self.V, self.N, 3), np.complex128)
        self._ck(self.lib.bcg_field_download(self._h, h, _dptr(out)))
        return out

    def zero(self, h):
        self._ck(self.lib.bcg_field_zero(self._h, h))

    def copy(self, dst, src):
        self._ck(self.lib.bcg_field_copy(self._h, dst, src))

    # ---- primitives ----
    def op(self, out, inp, sigma=0.0, want_gram=False):
        g = np.empty((self.N, self.N), np.complex128) if want_gram else None
        self._ck(self.lib.bcg_op(self._h, out, inp, float(sigma), _dptr(g) if want_gram else None))
        return mat_from_cm(g, self.N) if want_gram else None

    def gram(self, a, b):
        g = np.empty((self.N, self.N), np.complex128)
        self._ck(self.lib.bcg_gram(self._h, a, b, _dptr(g)))
        return mat_from_cm(g, self.N)

    def add(self, dst, src, M):
        if np.isscalar(M):
            self._ck(self.lib.bcg_add_scalar(self._h, dst, src, float(M)))
        else:
            self._ck(self.lib.bcg_add(self._h, dst, src, _dptr(mat_to_cm(M))))

    def rescale_add(self, dst, L, src, r):
        self._ck(self.lib.bcg_rescale_add(self._h, dst, _dptr(mat_to_cm(L)), src, float(r)))

    def trsm(self, q, R):
        self._ck(self.lib.bcg_trsm(self._h, q, _dptr(mat_to_cm(R))))

    def thinqr(self, q):
        R = np.empty((self.N, self.N), np.complex128)
        self._ck(self.lib.bcg_thinqr(self._h, q, _dptr(R)))
        return mat_from_cm(R, self.N)

    def true_residual(self, x, b, sigma=0.0):
        out = np.empty(self.N, np.float64)
        self._ck(self.lib.bcg_true_residual(self._h, x, b, float(sigma), _dptr(out)))
        return out

    # ---- solvers on device handles ----
    def solve_bcg_dev(self, x, b, eps=1e-15, max_iterations=1000000):
        info = SolveInfo()
        self._ck(self.lib.bcg_solve_bcg_dev(self._h, x, b, eps, int(max_iterations), C.byref(info)))
        return info

    def solve_bcgrq_dev(self, x, b, eps=1e-15, max_iterations=1000000):
        info = SolveInfo()
        self._ck(self.lib.bcg_solve_bcgrq_dev(self._h, x, b, eps, int(max_iterations), C.byref(info)))
        return info

    def solve_sbcgrq_dev(self, xs, b, sigma, eps=1e-15, eps_shifts=1e-15, max_iterations=1000000):
        info = SolveInfo()
        xs_a = (C.c_int * len(xs))(*xs)
        sig = np.ascontiguousarray(sigma, dtype=np.float64)
        self._ck(self.lib.bcg_solve_sbcgrq_dev(self._h, xs_a, b, _dptr(sig), len(xs), eps, eps_shifts,
                                               int(max_iterations), C.byref(info)))
        return info

    # ---- solvers on host buffers (the drop-in entry points) ----
    def solve_bcg(self, X, B, eps=1e-15, max_iterations=1000000):
        info = SolveInfo()
        self._ck(self.lib.bcg_solve_bcg(self._h, _dptr(X), _dptr(B), eps, int(max_iterations), C.byref(info)))
        return info

    def solve_bcgrq(self, X, B, eps=1e-15, max_iterations=1000000):
        info = SolveInfo()
        self._ck(self.lib.bcg_solve_bcgrq(self._h, _dptr(X), _dptr(B), eps, int(max_iterations), C.byref(info)))
        return info

    def solve_sbcgrq(self, Xs, B, sigma, eps=1e-15, eps_shifts=1e-15, max_iterations=1000000):
        info = SolveInfo()
        ptrs = (_dp * len(Xs))(*[_dptr(x) for x in Xs])
        sig = np.ascontiguousarray(sigma, dtype=np.float64)
        self._ck(self.lib.bcg_solve_sbcgrq(self._h, ptrs, _dptr(B), _dptr(sig), len(Xs), eps, eps_shifts,
                                           int(max_iterations), C.byref(info)))
        return info

    # ---- CG / SCG (one right-hand side, src/standard_solvers.cpp) ----
    def solve_cg(self, X, B, eps=1e-15, max_iterations=1000000):
        info = SolveInfo()
        self._ck(self.lib.bcg_solve_cg(self._h, _dptr(X), _dptr(B), eps, int(max_iterations), C.byref(info)))
        return info

    def solve_scg(self, Xs, B, sigma, eps=1e-15, eps_shifts=1e-15, max_iterations=1000000):
        info = SolveInfo()
        ptrs = (_dp * len(Xs))(*[_dptr(x) for x in Xs])
        sig = np.ascontiguousarray(sigma, dtype=np.float64)
        self._ck(self.lib.bcg_solve_scg(self._h, ptrs, _dptr(B), _dptr(sig), len(Xs), eps, eps_shifts,
                                        int(max_iterations), C.byref(info)))
        return info

    def last_solve_stats(self):
        """Active-system histogram, bytes moved by the multishift update and per-system residual
        estimates of the last solve on this context."""
        st = SolveStats()
        self._ck(self.lib.bcg_last_solve_stats(self._h, C.byref(st)))
        return {"iterations": st.iterations, "n_shifts": st.n_shifts, "paired": bool(st.paired), "schedule": st.paired, "depth": st.depth,
                "active_hist": list(st.active_hist), "shift_update_field_passes": int(st.shift_update_field_passes),
                "resid_shift": list(st.resid_shift)[:max(st.n_shifts, 1)]}

    def set_loop_profile(self, n_iterations, after_iterations=0):
        """Time n_iterations of the next solve stage by stage (CUDA events inside the loop), starting once
        after_iterations have run."""
        self._ck(self.lib.bcg_set_loop_profile(self._h, int(n_iterations), int(after_iterations)))

    def loop_profile(self):
        lp = LoopProfile()
        self._ck(self.lib.bcg_get_loop_profile(self._h, C.byref(lp)))
        keys = ["dirac_gram", "step_a", "axpy_gram", "step_b", "shift_odd", "shift_even", "halo", "iteration"]
        return {"iterations": lp.iterations, "first_iteration": lp.first_iteration, "active_systems": lp.active_systems,
                "ms": dict(zip(keys, [float(v) for v in lp.ms]))}

    # ---- the device N x N routines in isolation (unit tests) ----
    def small_inverse(self, A, pivot=True):
        out = np.empty((self.N, self.N), np.complex128)
        info = C.c_int(0)
        self._ck(self.lib.bcg_small_inverse(self._h, _dptr(mat_to_cm(A)), _dptr(out), 1 if pivot else 0, C.byref(info)))
        return mat_from_cm(out, self.N), info.value

    def small_lu_solve(self, A, B):
        out = np.empty((self.N, self.N), np.complex128)
        self._ck(self.lib.bcg_small_lu_solve(self._h, _dptr(mat_to_cm(A)), _dptr(mat_to_cm(B)), _dptr(out)))
        return mat_from_cm(out, self.N)

    def bench_kernel(self, which, reps, handles, n_shifts=1):
        ms = C.c_double(0)
        nl = C.c_int64(0)
        h = (C.c_int * len(handles))(*handles)
        self._ck(self.lib.bcg_bench_kernel(self._h, which, reps, n_shifts, h, len(handles), C.byref(ms), C.byref(nl)))
        return ms.value, nl.value
