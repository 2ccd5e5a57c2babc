"""TEST INFRASTRUCTURE: generate tests/golden/*.npz from the UNMODIFIED reference.

Run here (the container that has /root/reference):
    make -C oracle && python oracle/gen_golden.py
Every array below is produced by oracle/_ref/libref_n<N>.so, i.e. by the
reference's own templates (inc/block_solvers.hpp, inc/dirac_op.hpp,
inc/fields.hpp) on inputs drawn with the reference's own RNG usage
(srand(1), links first then B: benchmark.cpp:36-40).  The fixtures travel to
the GPU box; /root/reference does not.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pyoracle import RefShim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TEST_SHIFTS = [0.0, 0.01, 0.10, 0.20, 0.9]  # test/solvers.cpp:17
BENCH_SHIFTS = [0, 0, 1e-10, 1e-8, 1e-6, 1e-5, 1e-4, 1e-2, 1e-1]  # benchmark.cpp:12-13


def primitives(V, N, mass, seed=1):
    r = RefShim(N)
    U, B = r.make_inputs(V, seed)
    rng = np.random.default_rng(1234 + V + N)
    M = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    AB = r.op(U, B, mass)
    G = r.hermitian_dot(B, AB)
    Q, R = r.thinQR(B)
    d = dict(V=V, N=N, mass=mass, seed=seed, U=U, B=B, M=M, op=AB, gram_B_AB=G, gram_BB=r.hermitian_dot(B, B),
             add=r.add(B, AB, M), add_scalar=r.add(B, AB, 0.375), rescale_add=r.rescale_add(B, M, AB, 1.0),
             thinqr_Q=Q, thinqr_R=R, lu_inv_M=r.fullpivlu_inverse(M), lu_inv_G=r.fullpivlu_inverse(G),
             llt_upper_BB=r.llt_upper(r.hermitian_dot(B, B)))
    return d


def solvers(V, N, mass, eps, shifts, seed=1, eps_shifts=1e-15):
    r = RefShim(N)
    U, B = r.make_inputs(V, seed)
    Xb, itb, _ = r.BCG(U, B, mass, eps)
    Xq, itq, _ = r.BCGrQ(U, B, mass, eps)
    Xs, its, _ = r.SBCGrQ(U, B, mass, shifts, eps, eps_shifts)
    return dict(V=V, N=N, mass=mass, seed=seed, eps=eps, eps_shifts=eps_shifts, shifts=np.array(shifts, float),
                U=U, B=B, X_bcg=Xb, it_bcg=itb, X_bcgrq=Xq, it_bcgrq=itq, X_sbcgrq=Xs, it_sbcgrq=its)


def main():
    os.makedirs(OUT, exist_ok=True)
    # primitive known-answer vectors incl. degenerate periodic wraps (V=1,2,3) and ragged V
    for V, N in [(1, 1), (2, 2), (3, 1), (5, 2), (7, 3), (16, 3), (33, 4), (16, 8), (20, 12)]:
        np.savez_compressed(os.path.join(OUT, "prim_V%d_N%d.npz" % (V, N)), **primitives(V, N, 0.5))
    # the reference's own test configuration (test/solvers.cpp:8-17)
    np.savez_compressed(os.path.join(OUT, "solve_V128_N3.npz"), **solvers(128, 3, 0.5, 1e-10, TEST_SHIFTS))
    np.savez_compressed(os.path.join(OUT, "solve_V64_N1.npz"), **solvers(64, 1, 0.5, 1e-10, TEST_SHIFTS))
    np.savez_compressed(os.path.join(OUT, "solve_V48_N4.npz"), **solvers(48, 4, 0.1, 1e-10, TEST_SHIFTS))
    np.savez_compressed(os.path.join(OUT, "solve_V40_N12.npz"), **solvers(40, 12, 0.05, 1e-10, BENCH_SHIFTS))
    # CG / SCG (src/standard_solvers.cpp) at the reference's test configuration (test/solvers.cpp:8-17,19-51:
    # V=128, mass 0.5, eps 1e-10, the five test shifts) and at a long-running one
    r1 = RefShim(1)
    for V, mass, shifts, tag in [(128, 0.5, TEST_SHIFTS, "V128"), (200, 0.01, BENCH_SHIFTS, "V200")]:
        U, b = r1.make_inputs(V, 1)
        xc, itc, _ = r1.CG(U, b, mass, 1e-10)
        xs, its_, _ = r1.SCG(U, b, mass, shifts, 1e-10, 1e-15)
        np.savez_compressed(os.path.join(OUT, "scalar_%s.npz" % tag), V=V, N=1, mass=mass, eps=1e-10, eps_shifts=1e-15,
                            shifts=np.array(shifts, float), U=U, B=b, X_cg=xc, it_cg=itc, X_scg=xs, it_scg=its_)
        print("%s: CG %d it, SCG %d it" % (tag, itc, its_))
    # README / benchmark default: ./benchmark 1e3 1e-3 1e-10 (README.md:29): keep it small --
    # iteration counts, per-shift true residuals and a strided sample of the solutions.
    V, N, mass, eps = 1000, 12, 1e-3, 1e-10
    r = RefShim(N)
    U, B = r.make_inputs(V, 1)
    Xs, its, secs = r.SBCGrQ(U, B, mass, BENCH_SHIFTS, eps, 1e-15)
    Xq, itq, secq = r.BCGrQ(U, B, mass, eps)
    assert np.array_equal(Xs[0], Xq), "SBCGrQ shift 0 == BCGrQ bit-for-bit (SURVEY 3.2)"
    from oracle.pyoracle import Oracle
    o = Oracle()
    res = np.array([o.true_residual(U, B, Xs[s], mass, BENCH_SHIFTS[s]) for s in range(len(BENCH_SHIFTS))])
    colnorm = np.sqrt((np.abs(Xs) ** 2).sum(axis=(1, 3)))  # [S][N]
    np.savez_compressed(os.path.join(OUT, "bench_V1000_N12.npz"), V=V, N=N, mass=mass, eps=eps, seed=1,
                        shifts=np.array(BENCH_SHIFTS, float), it_sbcgrq=its, it_bcgrq=itq,
                        ref_seconds_sbcgrq=secs, ref_seconds_bcgrq=secq, true_residual=res,
                        X_colnorm=colnorm, X_sample=Xs[:, ::25].copy(), sample_stride=25)
    print("V=1000 N=12: SBCGrQ %d it (%.2fs), BCGrQ %d it (%.2fs), worst true res/shift %s"
          % (its, secs, itq, secq, res.max(axis=1)))
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden fixtures: %d files, %.1f KB" % (len(os.listdir(OUT)), tot / 1024))


if __name__ == "__main__":
    main()
