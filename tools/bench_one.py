"""Run a few launches of selected kernels (for ncu captures): python tools/bench_one.py V N S which [which...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blockcg_b200

V, N, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(0)
U = rng.uniform(-1, 1, (V, 3, 3)) + 1j * rng.uniform(-1, 1, (V, 3, 3))
with blockcg_b200.Context(V, N, max_shifts=S) as ctx:
    ctx.set_links(U, 1e-3)
    data = rng.uniform(-1, 1, (V, N, 3)) + 1j * rng.uniform(-1, 1, (V, N, 3))
    hs = [ctx.field(data) for _ in range(2)] + [ctx.field() for _ in range(2 * S - 1)]
    for h in hs[2:]:
        ctx.copy(h, hs[0])
    for w in sys.argv[4:]:
        w = int(w)
        nh, ns = (1 + 2 * S, S) if w in (4, 7, 8, 13) else (2, 1)
        print(w, ctx.bench_kernel(w, 2, hs[:nh], ns))
