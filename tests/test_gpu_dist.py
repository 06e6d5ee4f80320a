"""Multi-rank parity on the GPU box the driver gives the suite (ONE B200): two ranks share cuda:0, gloo
carries the set-up plumbing, and the data path is the product's own -- the peer-memory halo exchange and
all-reduce kernels (CUDA IPC between the two processes), the row-distributed multigrid cycle captured in a
CUDA graph, the replicated host logic.  Goldens: RijkeTube3D/Results/Active/active.log (config 1) and the
converged omega must be the single-GPU value (row e of the scope table)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, timeout=900):
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py"), "--same-device"] + extra
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stderr[-3000:]
    return json.loads(res.stdout.strip().split("\n")[-1])


def test_two_ranks_peer_transport_and_golden_fpi():
    out = _run(["--cases", "rijke3d"])
    assert len(out) == 2 and out[0]["transport"] == "peer"
    for r in out:
        assert all(v == 0 for v in r["unit"].values()), r["unit"]          # bitwise: halo values and rank-ordered sums
        assert r["rijke3d"]["max_abs_diff_vs_log"] < 2e-8, r["rijke3d"]     # the log prints 8 decimals
        assert r["rijke3d"]["distributed_levels"] >= 1
    # replicated host logic: both ranks hold the same omega, bit for bit
    assert out[0]["rijke3d"]["omega"] == out[1]["rijke3d"]["omega"]
