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
    ok = (res["op_rel"] < 1e-13 and res["gram_rel"] < 1e-12 and max(res["x_rel"]) < 1e-9
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
