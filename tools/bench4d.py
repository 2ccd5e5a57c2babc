"""4-D extension: stencil micro-benchmark and a capped SBCGrQ solve on an L^4 lattice.
    python tools/bench4d.py [L] [N] [max_it]
Algorithmic bytes of the 4-D apply: read P + write T + read the four links = (96 N + 576) V."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blockcg_b200

L = int(sys.argv[1]) if len(sys.argv) > 1 else 24
N = int(sys.argv[2]) if len(sys.argv) > 2 else 12
max_it = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dims = (L, L, L, L)
V = L ** 4
shifts = [0, 0, 1e-10, 1e-8, 1e-6, 1e-5, 1e-4, 1e-2, 1e-1]
rng = np.random.default_rng(1)
U = rng.uniform(-1, 1, (V, 4, 3, 3)) + 1j * rng.uniform(-1, 1, (V, 4, 3, 3))
B = rng.uniform(-1, 1, (V, N, 3)) + 1j * rng.uniform(-1, 1, (V, N, 3))
out = {"dims": dims, "N": N}
with blockcg_b200.Context(V, N, max_shifts=len(shifts), dims=dims) as ctx:
    ctx.set_links(U, 1e-3)
    hb, ha = ctx.field(B), ctx.field()
    nbytes = (96.0 * N + 576.0) * V
    for name, which in [("dirac4", 1), ("dirac4_gram", 0)]:
        ms, nl = ctx.bench_kernel(which, 10, [hb, ha], 1)
        out[name] = {"us": round(1e3 * ms, 1), "alg_GBps": round(nbytes / ms / 1e6), "launches_per_apply": nl // 10}
    xs = [ctx.field() for _ in shifts]
    info = ctx.solve_sbcgrq_dev(xs, hb, shifts, 1e-10, 1e-15, max_it)
    out["sbcgrq"] = {"iterations": info.iterations, "residual": info.residual, "ms_per_iteration": info.solve_ms / max(info.iterations, 1)}
print(json.dumps(out))
