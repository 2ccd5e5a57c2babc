"""BASELINE configs[1]: BCGrQ single-shift block solve, 16^4 sites, N = 4 / 8 / 12, mass 1e-3, tol 1e-10,
on one B200: time-to-solution, iterations, true residual (device verification path) and, as the CPU side,
the reference's seconds per iteration on a 3-iteration sample of the same inputs."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blockcg_b200
from oracle.pyoracle import RefShim

V, mass, eps = 16 ** 4, 1e-3, 1e-10
for N in (4, 8, 12):
    rng = np.random.default_rng(1)
    U = rng.uniform(-1, 1, (V, 3, 3)) + 1j * rng.uniform(-1, 1, (V, 3, 3))
    B = rng.uniform(-1, 1, (V, N, 3)) + 1j * rng.uniform(-1, 1, (V, N, 3))
    with blockcg_b200.Context(V, N) as ctx:
        ctx.set_links(U, mass)
        hb, hx = ctx.field(B), ctx.field()
        ctx.solve_bcgrq_dev(hx, hb, eps, 50)  # warm-up: graph build
        info = ctx.solve_bcgrq_dev(hx, hb, eps)
        res = float(ctx.true_residual(hx, hb, 0.0).max())
    out = {"config": "BCGrQ 16^4", "N": N, "iterations": info.iterations, "gpu_seconds": (info.setup_ms + info.solve_ms) / 1e3,
           "ms_per_iteration": info.solve_ms / info.iterations, "true_residual": res}
    if RefShim.available(N):
        _, it, sec = RefShim(N).BCGrQ(U, B, mass, eps, max_it=3)
        out["reference_cpu_s_per_iteration"] = sec / it
        out["reference_cpu_extrapolated_s"] = sec / it * info.iterations
    print(json.dumps(out), flush=True)
