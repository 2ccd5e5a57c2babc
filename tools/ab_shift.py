"""A/B of the multishift update kernels in one process (the library reads BCG_DMMA / BCG_PAIR per call):
    python tools/ab_shift.py [V] [N] [max_it]
Per variant: micro-benchmark of the paired update at 1 / 5 / 9 active systems, a capped solve (device time,
in-loop stage profile of iterations 200..263) and the solutions' relative difference to the first variant."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blockcg_b200  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 331776
N = int(sys.argv[2]) if len(sys.argv) > 2 else 12
max_it = int(sys.argv[3]) if len(sys.argv) > 3 else 600
variants = sys.argv[4:] or ["BCG_DMMA=0", "BCG_DMMA=1"]
shifts = [0, 0, 1e-10, 1e-8, 1e-6, 1e-5, 1e-4, 1e-2, 1e-1]  # benchmark.cpp:12-13
S = len(shifts)
rng = np.random.default_rng(1)
U = rng.uniform(-1, 1, (V, 3, 3)) + 1j * rng.uniform(-1, 1, (V, 3, 3))
B = rng.uniform(-1, 1, (V, N, 3)) + 1j * rng.uniform(-1, 1, (V, N, 3))
F = 48.0 * N * V
ref = None
for var in variants:
    for kv in var.split(","):
        k, v = kv.split("=")
        os.environ[k] = v
    out = {"variant": var, "V": V, "N": N}
    with blockcg_b200.Context(V, N, max_shifts=S) as ctx:
        ctx.set_links(U, 1e-3)
        hb = ctx.field(B)
        hs = [hb] + [ctx.field() for _ in range(2 * S)]
        for h in hs[1:]:
            ctx.copy(h, hb)
        for nm, which in (("dirac_gram", 0), ("dirac", 1), ("axpy_gram", 3), ("axpy", 5)):
            ms, _ = ctx.bench_kernel(which, 20, hs[:2], 1)
            out[nm + "_ms"] = ms
        for a in (1, 5, 9):
            ms, _ = ctx.bench_kernel(13, 8, hs[:1 + 2 * a], a)
            out["pair_ms_per_launch[%d]" % a] = ms / 2
            out["pair_effGBps[%d]" % a] = (2 + 4 * a) * F / (ms / 2) / 1e6
        ms, _ = ctx.bench_kernel(4, 8, hs[:1 + 2 * S], S)
        out["plain_ms[9]"] = ms
        for h in hs[1 + S:]:
            ctx.free(h)
        xs = hs[1:1 + S]
        ctx.set_loop_profile(64, 200)
        info = ctx.solve_sbcgrq_dev(xs, hb, shifts, 1e-10, 1e-15, max_it)
        out["profile"] = ctx.loop_profile()
        info = ctx.solve_sbcgrq_dev(xs, hb, shifts, 1e-10, 1e-15, max_it)
        out["solve_ms"] = info.solve_ms
        out["ms_per_iteration"] = info.solve_ms / max(info.iterations, 1)
        out["iterations"] = info.iterations
        out["residual"] = info.residual
        X = [ctx.download(x) for x in (xs[0], xs[4], xs[8])]
        if ref is None:
            ref = X
        else:
            out["x_rel_vs_first"] = [float(np.abs(a - b).max() / np.abs(b).max()) for a, b in zip(X, ref)]
    for kv in var.split(","):
        os.environ.pop(kv.split("=")[0], None)
    print(json.dumps(out), flush=True)
