"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: slab partition, halo bookkeeping
and unique-id broadcast of blockcg_b200/distributed.py, and the slab algorithm itself (2-site
halo exchange + Gram all-reduce) emulated with the oracle's primitives -- it must reproduce the
single-domain operator and Gram.  The CUDA library implements the same scheme with NCCL; that
path is exercised on GPUs by tools/multigpu_check.py."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, V, N, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from blockcg_b200 import distributed as D
    from oracle.pyoracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = Oracle()
        U, B = o.make_inputs(V, N, 5)
        mass = 0.4
        b, e = D.slab_range(V, rank, world)
        Vl = e - b
        left, right = D.neighbours(rank, world)

        def exchange(local, site_shape):
            """local: owned sites -> array with 2+2 halo sites, neighbour exchange over gloo."""
            ext = np.zeros((Vl + 4,) + site_shape, np.complex128)
            ext[2:-2] = local
            lo = torch.from_numpy(np.ascontiguousarray(local[:2]).view(np.float64).copy())
            hi = torch.from_numpy(np.ascontiguousarray(local[-2:]).view(np.float64).copy())
            rlo, rhi = torch.empty_like(lo), torch.empty_like(hi)
            reqs = [dist.isend(lo, left), dist.isend(hi, right), dist.irecv(rhi, right), dist.irecv(rlo, left)]
            for r in reqs:
                r.wait()
            ext[:2] = rlo.numpy().view(np.complex128).reshape((2,) + site_shape)
            ext[-2:] = rhi.numpy().view(np.complex128).reshape((2,) + site_shape)
            return ext

        # the halo slots must hold exactly the sites halo_sources() names
        lo_src, hi_src = D.halo_sources(V, rank, world)
        Bext = exchange(B[b:e], (N, 3))
        Uext = exchange(U[b:e], (3, 3))
        assert np.array_equal(Bext[:2], B[lo_src]) and np.array_equal(Bext[-2:], B[hi_src])
        assert np.array_equal(Uext[:2], U[lo_src]) and np.array_equal(Uext[-2:], U[hi_src])
        # local operator on the halo-extended slab: interior sites of a (Vl+4)-site periodic
        # oracle call are exact, because op reaches only x +- 2
        AB_loc = o.op(Uext, Bext, mass)[2:-2]
        AB_ref = o.op(U, B, mass)[b:e]
        assert np.abs(AB_loc - AB_ref).max() <= 1e-14 * np.abs(AB_ref).max()
        # Gram = all-reduce of slab Grams
        G = o.hermitian_dot(np.ascontiguousarray(B[b:e]), np.ascontiguousarray(AB_loc))
        t = torch.from_numpy(G.view(np.float64).copy())
        dist.all_reduce(t)
        G_all = t.numpy().view(np.complex128).reshape(N, N)
        G_ref = o.hermitian_dot(B, o.op(U, B, mass))
        assert np.abs(G_all - G_ref).max() <= 1e-13 * np.abs(G_ref).max()
        # the NCCL id travels unchanged through the broadcast helper (library call stubbed on CPU)
        from blockcg_b200 import capi
        capi.Context.unique_id = staticmethod(lambda: bytes(range(128)))
        assert D.broadcast_unique_id(dist) == bytes(range(128))
        out.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        out.put((rank, "FAIL: %r" % (ex,)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("V,N", [(64, 3), (10, 12)])
def test_slab_algorithm_world2(V, N, oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, V, N, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_partition_helpers():
    from blockcg_b200 import distributed as D
    assert D.slab_range(96, 0, 4) == (0, 24) and D.slab_range(96, 3, 4) == (72, 96)
    assert D.neighbours(0, 4) == (3, 1) and D.neighbours(3, 4) == (2, 0)
    assert D.halo_sources(96, 0, 4) == ([94, 95], [24, 25])
    assert D.halo_sources(96, 3, 4) == ([70, 71], [0, 1])
    assert D.halo_sources(8, 0, 1) == ([6, 7], [0, 1])
    with pytest.raises(ValueError):
        D.slab_range(10, 0, 3)
    with pytest.raises(ValueError):
        D.slab_range(4, 0, 4)
