"""GPU test (-m gpu) of the slab decomposition on real devices: runs tools/multigpu_check.py under
torchrun on 2 GPUs (skipped on a single-GPU box): block Dirac apply, Gram and a full multishift solve
on contiguous site slabs against the single-domain CPU oracle, for the reference's chain (halo sites and
Gram blocks exchanged by P2P stores over NVLink from inside the kernels, and again through NCCL) and for
the 4-D extension (x3-slabs, slice halos, exchange overlapped with the interior)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("no_p2p", ["0", "1"])
def test_two_slabs_match_single_domain_oracle(no_p2p):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, BCG_NO_P2P=no_p2p)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "2953" + no_p2p, os.path.join(ROOT, "tools", "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert res["ok_all_ranks"] and res["p2p"] == (no_p2p == "0")
    assert res["op_rel"] < 1e-13 and max(res["x_rel"]) < 1e-9 and abs(res["iterations"] - res["oracle_iterations"]) <= 2
    assert res["op4_rel"] < 1e-13 and max(res["x4_rel"]) < 1e-9
