"""Old-vs-new library check: python tools/ab_lib.py LIBPATH
Runs capped multishift solves (and a BCG / BCGrQ solve) at several (V, N) with the library given and prints,
per case, SHA-256 digests of the solutions plus the device time per iteration and the in-loop stage profile.
Two runs (two libraries) whose digests agree produced bit-identical solutions."""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blockcg_b200.capi as capi  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] != "-":
    capi.LIB_PATH = os.path.abspath(sys.argv[1])
import blockcg_b200  # noqa: E402

shifts = [0, 0, 1e-10, 1e-8, 1e-6, 1e-5, 1e-4, 1e-2, 1e-1]  # benchmark.cpp:12-13
S = len(shifts)
cases = [(1000, 12, 10 ** 6), (4096, 4, 300), (4096, 8, 300), (2048, 16, 200), (1500, 3, 300), (41472, 12, 600), (331776, 12, 600)]
if len(sys.argv) > 2:
    cases = [tuple(int(x) for x in c.split(",")) for c in sys.argv[2:]]
for V, N, max_it in cases:
    rng = np.random.default_rng(V + N)
    U = rng.uniform(-1, 1, (V, 3, 3)) + 1j * rng.uniform(-1, 1, (V, 3, 3))
    B = rng.uniform(-1, 1, (V, N, 3)) + 1j * rng.uniform(-1, 1, (V, N, 3))
    out = {"lib": os.path.basename(capi.LIB_PATH), "V": V, "N": N}
    with blockcg_b200.Context(V, N, max_shifts=S) as ctx:
        ctx.set_links(U, 1e-3 if V > 2000 else 0.05)
        hb = ctx.field(B)
        xs = [ctx.field() for _ in range(S)]
        if V >= 40000:
            ctx.set_loop_profile(64, 200)
        info = ctx.solve_sbcgrq_dev(xs, hb, shifts, 1e-10, 1e-15, max_it)
        if V >= 40000:
            out["profile_ms"] = {k: round(v, 5) for k, v in ctx.loop_profile()["ms"].items()}
            ctx.set_loop_profile(0, 0)
        info = ctx.solve_sbcgrq_dev(xs, hb, shifts, 1e-10, 1e-15, max_it)
        out["iterations"] = info.iterations
        out["residual"] = info.residual
        out["ms_per_iteration"] = info.solve_ms / max(info.iterations, 1)
        h = hashlib.sha256()
        for x in xs:
            h.update(np.ascontiguousarray(ctx.download(x)).tobytes())
        out["sbcgrq_sha256"] = h.hexdigest()[:16]
        if V <= 5000:
            hx = ctx.field()
            i2 = ctx.solve_bcg_dev(hx, hb, 1e-10, max_it) if hasattr(ctx, "solve_bcg_dev") else None
            if i2 is not None:
                out["bcg_iterations"] = i2.iterations
                out["bcg_sha256"] = hashlib.sha256(np.ascontiguousarray(ctx.download(hx)).tobytes()).hexdigest()[:16]
    print(json.dumps(out), flush=True)
