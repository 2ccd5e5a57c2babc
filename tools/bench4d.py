"""4-D extension: stencil micro-benchmark and a capped SBCGrQ solve.
    python tools/bench4d.py L0 L1 L2 L3 [N] [max_it]                      (one GPU)
    torchrun --nproc-per-node G tools/bench4d.py L0 L1 L2 L3 [N] [max_it] (t-slabs: L3 split over G ranks)
Inputs are generated per rank (local slab only).  Algorithmic bytes of the 4-D apply: read P + write T
+ read the four links = (96 N + 576) V."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blockcg_b200

a = [int(x) for x in sys.argv[1:]]
dims_g = tuple(a[:4]) if len(a) >= 4 else (24, 24, 24, 24)
N = a[4] if len(a) > 4 else 12
max_it = a[5] if len(a) > 5 else 200
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from blockcg_b200 import distributed as D
assert dims_g[3] % world == 0
dims = dims_g[:3] + (dims_g[3] // world,)
V = int(np.prod(dims))
shifts = [0, 0, 1e-10, 1e-8, 1e-6, 1e-5, 1e-4, 1e-2, 1e-1]
rng = np.random.default_rng(1 + rank)
U = rng.uniform(-1, 1, (V, 4, 3, 3)) + 1j * rng.uniform(-1, 1, (V, 4, 3, 3))
B = rng.uniform(-1, 1, (V, N, 3)) + 1j * rng.uniform(-1, 1, (V, N, 3))
out = {"dims_global": dims_g, "n_gpus": world, "N": N}
ctx = blockcg_b200.Context(V, N, max_shifts=len(shifts), device=local, rank=rank, nranks=world, dims=dims)
if world > 1:
    ctx.comm_init(D.broadcast_unique_id(dist, torch.device("cuda", local)))
ctx.set_links(U, 1e-3)
del U
hb, ha = ctx.field(B), ctx.field()
del B
nbytes = (96.0 * N + 576.0) * V
for name, which in [("dirac4", 1), ("dirac4_gram", 0)]:
    ms, nl = ctx.bench_kernel(which, 10, [hb, ha], 1)
    out[name] = {"us": round(1e3 * ms, 1), "alg_GBps_per_gpu": round(nbytes / ms / 1e6), "launches_per_apply": nl // 10}
xs = [ctx.field() for _ in shifts]
info = ctx.solve_sbcgrq_dev(xs, hb, shifts, 1e-10, 1e-15, max_it)
out["sbcgrq"] = {"iterations": info.iterations, "residual": info.residual,
                 "ms_per_iteration": info.solve_ms / max(info.iterations, 1)}
if rank == 0:
    print(json.dumps(out), flush=True)
ctx.close()
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
