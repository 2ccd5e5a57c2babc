"""Kernel micro-benchmarks through the C-ABI (bcg_bench_kernel): mean device time per
launch with CUDA events on the library's stream, inputs larger than L2 where the
volume allows, algorithmic GB/s per SURVEY 8(d).

    python tools/microbench.py [--V 331776] [--N 12] [--S 9] [--reps 20]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blockcg_b200  # noqa: E402


def sweep(V, N, reps, seed=0):
    """BASELINE configs[4]: block Dirac apply / Gram / fused update at one (V, N); three fields only, inputs
    drawn on the device (bcg_set_links_random / bcg_field_random: a 64^4 x 32 field is 26 GB)."""
    F, Ub = 48.0 * N * V, 144.0 * V
    out = {"V": V, "N": N}
    with blockcg_b200.Context(V, N, max_shifts=1) as ctx:
        ctx.set_links_random(seed + 1, 1e-3)
        hs = [ctx.field(), ctx.field(), ctx.field()]
        ctx.field_random(hs[0], seed + 2)
        ctx.field_random(hs[1], seed + 3)
        ctx.field_random(hs[2], seed + 4)
        for name, which, nh, nbytes in [("dirac", 1, 2, 2 * F + Ub), ("dirac_gram", 0, 2, 2 * F + Ub),
                                        ("gram", 2, 2, 2 * F), ("axpy_gram", 3, 2, 3 * F),
                                        ("shift_update_S1", 4, 3, 6 * F)]:
            ms, _ = ctx.bench_kernel(which, reps, hs[:nh], 1)
            out[name] = {"us": round(1e3 * ms, 1), "GBps": round(nbytes / ms / 1e6)}
    return out


def run(V, N, S, reps, seed=0):
    rng = np.random.default_rng(seed)
    U = rng.uniform(-1, 1, (V, 3, 3)) + 1j * rng.uniform(-1, 1, (V, 3, 3))
    out = {"V": V, "N": N, "S": S, "reps": reps}
    F = 48.0 * N * V
    Ub = 144.0 * V
    with blockcg_b200.Context(V, N, max_shifts=S) as ctx:
        ctx.set_links(U, 1e-3)
        data = rng.uniform(-1, 1, (V, N, 3)) + 1j * rng.uniform(-1, 1, (V, N, 3))
        hs = [ctx.field(data) for _ in range(2)]
        hs += [ctx.field() for _ in range(2 * S - 1)]
        for h in hs[2:]:
            ctx.copy(h, hs[0])
        kernels = [
            ("dirac_gram", 0, hs[:2], 1, 2 * F + Ub),
            ("dirac", 1, hs[:2], 1, 2 * F + Ub),
            ("dirac_gram_v1", 9, hs[:2], 1, 2 * F + Ub),
            ("dirac_v1", 10, hs[:2], 1, 2 * F + Ub),
            ("gram", 2, hs[:2], 1, 2 * F),
            ("axpy_gram", 3, hs[:2], 1, 3 * F),
            ("axpy", 5, hs[:2], 1, 3 * F),
            ("axpy_gram_v1", 11, hs[:2], 1, 3 * F),
            ("axpy_v1", 12, hs[:2], 1, 3 * F),
            ("rescale_add", 6, hs[:2], 1, 3 * F),
            ("shift_update_S%d" % S, 4, hs[:1 + 2 * S], S, (2 + 4 * S) * F),
            ("shift_update_S1", 4, hs[:3], 1, 6 * F),
            ("shift_direct_S%d" % S, 7, hs[:1 + 2 * S], S, (2 + 4 * S) * F),
        ]
        if S > 1:  # paired update: one repetition = odd + even launch, "ms" is the pair
            kernels.append(("shift_pair_S%d" % S, 13, hs[:1 + 2 * S], S, 2 * (2 + 4 * S) * F))  # fixed per-unit bytes
        for name, which, handles, ns, bytes_alg in kernels:
            ms, nl = ctx.bench_kernel(which, reps, handles, ns)
            out[name] = {"ms": ms, "alg_GBps": bytes_alg / ms / 1e6, "launches": nl}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--V", type=int, nargs="+", default=[331776])
    ap.add_argument("--N", type=int, nargs="+", default=[12])
    ap.add_argument("--S", type=int, default=9)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--sweep", action="store_true", help="Dirac / Gram / update sweep over (V, N), 3 fields")
    a = ap.parse_args()
    if a.sweep:
        for V in a.V:
            for N in a.N:
                if 48.0 * N * V * 3.2 > 150e9:   # three fields must fit one GPU
                    continue
                try:
                    print(json.dumps(sweep(V, N, a.reps)), flush=True)
                except Exception as e:  # keep sweeping: one (V, N) that fails must not lose the rest
                    print(json.dumps({"V": V, "N": N, "error": str(e)[:200]}), flush=True)
        sys.exit(0)
    for V in a.V:
        for N in a.N:
            print(json.dumps(run(V, N, a.S, a.reps)), flush=True)
