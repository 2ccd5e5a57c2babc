"""A few iterations of the headline solve (for ncu launch lists): python tools/solve_few.py [V] [N] [max_it]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blockcg_b200

V = int(sys.argv[1]) if len(sys.argv) > 1 else 331776
N = int(sys.argv[2]) if len(sys.argv) > 2 else 12
max_it = int(sys.argv[3]) if len(sys.argv) > 3 else 12
shifts = [0, 0, 1e-10, 1e-8, 1e-6, 1e-5, 1e-4, 1e-2, 1e-1]  # benchmark.cpp:12-13
rng = np.random.default_rng(1)
U = rng.uniform(-1, 1, (V, 3, 3)) + 1j * rng.uniform(-1, 1, (V, 3, 3))
B = rng.uniform(-1, 1, (V, N, 3)) + 1j * rng.uniform(-1, 1, (V, N, 3))
with blockcg_b200.Context(V, N, max_shifts=len(shifts)) as ctx:
    ctx.set_links(U, 1e-3)
    hb = ctx.field(B)
    xs = [ctx.field() for _ in shifts]
    info = ctx.solve_sbcgrq_dev(xs, hb, shifts, 1e-10, 1e-15, max_it)
    print("iterations", info.iterations, "residual", info.residual, "solve_ms", info.solve_ms, "launches", info.kernel_launches)
