"""Parity of a FULL multishift solve at a size the reference still finishes in minutes:
V = 8^4 = 4096 sites, N = 12, mass 1e-3, tol 1e-10, the nine benchmark shifts (benchmark.cpp:12-13).
Runs (a) the unmodified reference (oracle/_ref), (b) the CPU oracle with the GPU-like tree-shaped
Gram summation, (c) the GPU path, on the same inputs; prints iteration counts, per-shift relative
differences of the solutions against the reference, and true residuals."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blockcg_b200
from oracle.pyoracle import Oracle, RefShim

V = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N, mass, eps = 12, 1e-3, 1e-10
shifts = [0, 0, 1e-10, 1e-8, 1e-6, 1e-5, 1e-4, 1e-2, 1e-1]
o = Oracle()
U, B = o.make_inputs(V, N, 1)
out = {"V": V, "N": N, "mass": mass, "eps": eps, "n_shifts": len(shifts)}
D = blockcg_b200.dirac_op(V, mass, links=U)
Xg = [np.empty_like(B) for _ in shifts]
info = {}
t0 = time.time()
out["gpu_iterations"] = blockcg_b200.SBCGrQ(Xg, B, D, shifts, eps, 1e-15, info=info)
out["gpu_seconds"] = time.time() - t0
if os.environ.get("GPU_ONLY"):
    out["gpu_true_residual"] = [float(o.true_residual(U, B, Xg[s], mass, shifts[s]).max()) for s in range(len(shifts))]
    print(json.dumps(out))
    sys.exit(0)
t0 = time.time()
Xt, itt, _, _ = o.SBCGrQ(U, B, mass, shifts, eps, 1e-15, chunk=32)
out["tree_oracle_iterations"], out["tree_oracle_seconds"] = itt, time.time() - t0
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
out["gpu_vs_tree_oracle"] = [rel(Xg[s], Xt[s]) for s in range(len(shifts))]
if RefShim.available(N):
    t0 = time.time()
    Xr, itr, _ = RefShim(N).SBCGrQ(U, B, mass, shifts, eps, 1e-15)
    out["reference_iterations"], out["reference_seconds"] = itr, time.time() - t0
    out["gpu_vs_reference"] = [rel(Xg[s], Xr[s]) for s in range(len(shifts))]
    out["tree_oracle_vs_reference"] = [rel(Xt[s], Xr[s]) for s in range(len(shifts))]
    out["reference_true_residual"] = [float(o.true_residual(U, B, Xr[s], mass, shifts[s]).max()) for s in range(len(shifts))]
out["gpu_true_residual"] = [float(o.true_residual(U, B, Xg[s], mass, shifts[s]).max()) for s in range(len(shifts))]
print(json.dumps(out))
