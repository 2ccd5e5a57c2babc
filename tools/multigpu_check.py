"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py
Slab-decomposed op / Gram / SBCGrQ (NCCL halo exchange + Gram all-reduce inside the CUDA
library) against the single-domain CPU oracle on the same inputs."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blockcg_b200  # noqa: E402
from blockcg_b200 import distributed as D  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    o = Oracle()
    V, N, mass = 4096, 12, 0.05
    shifts = [0.0, 1e-4, 1e-2, 1e-1]
    U, B = o.make_inputs(V, N, 1)
    p2p = os.environ.get("BCG_NO_P2P", "0") != "1"
    ctx, (b, e) = D.make_context(dist, V, N, len(shifts), local, U, mass, p2p=p2p)
    res = {"world": world, "V": V, "N": N, "p2p": p2p}
    hb, ha = ctx.field(np.ascontiguousarray(B[b:e])), ctx.field()
    G = ctx.op(ha, hb, sigma=0.125, want_gram=True)
    AB = o.op(U, B, mass, 0.125)
    res["op_rel"] = float(np.abs(ctx.download(ha) - AB[b:e]).max() / np.abs(AB).max())
    Gref = o.hermitian_dot(B, AB)
    res["gram_rel"] = float(np.abs(G - Gref).max() / np.abs(Gref).max())
    xs = [ctx.field() for _ in shifts]
    info = ctx.solve_sbcgrq_dev(xs, hb, shifts, 1e-10, 1e-15)
    Xo, ito, _, _ = o.SBCGrQ(U, B, mass, shifts, 1e-10, 1e-15, chunk=32)
    res["iterations"], res["oracle_iterations"] = info.iterations, ito
    res["x_rel"] = [float(np.abs(ctx.download(xs[s]) - Xo[s][b:e]).max() / np.abs(Xo[s]).max()) for s in range(len(shifts))]
    res["true_res"] = [float(ctx.true_residual(xs[s], hb, shifts[s]).max()) for s in range(len(shifts))]
    res["solve_ms"] = info.solve_ms
    # ---- 4-D extension: t-slabs (x3 slowest), halo = one x3-slice per side, two sweeps per apply ----
    dims = (6, 4, 6, 4 * world)
    V4 = int(np.prod(dims))
    rng = np.random.default_rng(3)
    U4 = rng.uniform(-1, 1, (V4, 4, 3, 3)) + 1j * rng.uniform(-1, 1, (V4, 4, 3, 3))
    B4 = rng.uniform(-1, 1, (V4, N, 3)) + 1j * rng.uniform(-1, 1, (V4, N, 3))
    sl = dims[0] * dims[1] * dims[2]
    b4, e4 = rank * 4 * sl, (rank + 1) * 4 * sl
    ctx4 = blockcg_b200.Context(0, N, max_shifts=2, device=local, rank=rank, nranks=world, dims=dims[:3] + (4,))
    ctx4.comm_init(D.broadcast_unique_id(dist, torch.device("cuda", local)))
    ctx4.set_links(np.ascontiguousarray(U4[b4:e4]), 0.3)
    o.set_lattice(dims)
    h4b, h4a = ctx4.field(np.ascontiguousarray(B4[b4:e4])), ctx4.field()
    G4 = ctx4.op(h4a, h4b, sigma=0.25, want_gram=True)
    AB4 = o.op(U4, B4, 0.3, 0.25)
    res["op4_rel"] = float(np.abs(ctx4.download(h4a) - AB4[b4:e4]).max() / np.abs(AB4).max())
    G4ref = o.hermitian_dot(B4, AB4)
    res["gram4_rel"] = float(np.abs(G4 - G4ref).max() / np.abs(G4ref).max())
    x4 = [ctx4.field(), ctx4.field()]
    info4 = ctx4.solve_sbcgrq_dev(x4, h4b, [0.0, 0.1], 1e-10, 1e-15)
    Xo4, ito4, _, _ = o.SBCGrQ(U4, B4, 0.3, [0.0, 0.1], 1e-10, 1e-15, chunk=32)
    o.set_lattice(None)
    res["iterations4"], res["oracle_iterations4"] = info4.iterations, ito4
    res["x4_rel"] = [float(np.abs(ctx4.download(x4[s]) - Xo4[s][b4:e4]).max() / np.abs(Xo4[s]).max()) for s in range(2)]
    ok4 = (res["op4_rel"] < 1e-13 and res["gram4_rel"] < 1e-12 and max(res["x4_rel"]) < 1e-9
           and abs(info4.iterations - ito4) <= 2)
    ctx4.close()
    ok = ok4 and (res["op_rel"] < 1e-13 and res["gram_rel"] < 1e-12 and max(res["x_rel"]) < 1e-9
          and abs(info.iterations - ito) <= max(2, ito // 100) and max(res["true_res"]) < 2e-10)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    res["ok_all_ranks"] = bool(t.item())
    if rank == 0:
        print(json.dumps(res), flush=True)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if res["ok_all_ranks"] else 1)


if __name__ == "__main__":
    main()
