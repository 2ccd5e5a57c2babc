"""Multi-GPU plumbing: one process per GPU, contiguous site slabs (for V = L^3 x T this is the
t-slab split), torch.distributed only to agree on the NCCL id and to move test data.

The data path itself lives in the CUDA library: halo exchange (2 sites per side for the
reference's 1-D operator) and the N x N Gram all-reduce run on the library's own
communicator and stream (blockcg_b200/csrc/capi.cu: halo_refresh, gram_finalize).
"""
import numpy as np

from .capi import IPC_HANDLE_BYTES, UNIQUE_ID_BYTES, Context

HALO = 2  # op = m^2 - D^2 reaches x +- 2 (inc/dirac_op.hpp:14-21,36-43)


def slab_range(V, rank, nranks):
    """[begin, end) of the contiguous site range owned by `rank` (V must divide evenly)."""
    if V % nranks:
        raise ValueError("V=%d is not divisible by %d ranks" % (V, nranks))
    if V // nranks < HALO:
        raise ValueError("need at least %d sites per rank" % HALO)
    n = V // nranks
    return rank * n, (rank + 1) * n


def neighbours(rank, nranks):
    return (rank - 1) % nranks, (rank + 1) % nranks


def halo_sources(V, rank, nranks):
    """Global site indices whose values fill this rank's halo slots [-2,-1] and [Vl, Vl+1]."""
    b, e = slab_range(V, rank, nranks)
    return [(b - 2) % V, (b - 1) % V], [e % V, (e + 1) % V]


def broadcast_unique_id(dist, device=None):
    """Rank 0 creates the library's NCCL id, everybody receives it (any torch backend)."""
    import torch
    buf = torch.zeros(UNIQUE_ID_BYTES, dtype=torch.uint8)
    if dist.get_rank() == 0:
        buf = torch.frombuffer(bytearray(Context.unique_id()), dtype=torch.uint8).clone()
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


def exchange_ipc_handles(dist, ctx, device=None):
    """Map every rank's communication buffer into every other rank (CUDA IPC, one node): after
    this the iteration loop exchanges halos and Gram blocks with P2P stores over NVLink and
    contains no NCCL call."""
    import torch
    from .capi import BcgError
    ok = 1
    try:
        raw = ctx.ipc_handle()
    except BcgError:
        raw, ok = bytes(IPC_HANDLE_BYTES), 0
    mine = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    flag = torch.tensor([ok], dtype=torch.int32)
    if device is not None:
        mine, flag = mine.to(device), flag.to(device)
    allh = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(allh, mine)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 1:
        try:
            ctx.ipc_open(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
        except BcgError:
            ok = 0
    flag = torch.tensor([ok if int(flag.item()) == 1 else 0], dtype=torch.int32)
    if device is not None:
        flag = flag.to(device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # also the barrier: nobody pushes before everybody has mapped everybody
    if int(flag.item()) != 1:
        # some rank could not map its peers (no peer access / IPC forbidden): every rank stays on the
        # NCCL path (halo send/recv + all-reduce), which needs nothing beyond bcg_comm_init
        ctx.ipc_disable()
        return False
    return True


def make_context(dist, V, N, max_shifts, device, U_global, mass, p2p=True):
    """Slab context of this rank, communicator initialised, links (own slab) uploaded."""
    rank, world = dist.get_rank(), dist.get_world_size()
    b, e = slab_range(V, rank, world)
    ctx = Context(e - b, N, max_shifts=max_shifts, device=device, rank=rank, nranks=world)
    import torch
    uid = broadcast_unique_id(dist, torch.device("cuda", device) if dist.get_backend() == "nccl" else None)
    ctx.comm_init(uid)
    if p2p and world > 1 and dist.get_backend() == "nccl":
        exchange_ipc_handles(dist, ctx, torch.device("cuda", device))
    ctx.set_links(np.ascontiguousarray(U_global[b:e]), mass)
    return ctx, (b, e)
