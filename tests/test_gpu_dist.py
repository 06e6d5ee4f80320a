"""Multi-GPU parity through the product's own data path -- peer-memory halo exchange and all-reduce
kernels, row-distributed multigrid cycle in a CUDA graph, replicated host logic -- on TWO real GPUs
(skipped on a one-GPU box: ranks must not share a device, their kernels wait on one another).  On the
driver's one-GPU box the multi-GPU path is covered by the gloo tests (tests/test_dist_cpu.py), by the
single-rank kernel test below, and by bench.py, which at N > 1 checks the converged omega against the
single-GPU value (tests/golden/bench_omega.json) to 1e-8.  Goldens: RijkeTube3D/Results/Active/active.log
(config 1), fullAnnulus/Results/Active/FPI/eigenvalues_dir.txt (config 3)."""
import ctypes as C
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (one rank per GPU)")
def test_two_gpus_peer_transport_and_golden_fpi():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads(res.stdout.strip().split("\n")[-1])
    assert len(out) == 2 and out[0]["transport"] == "peer"
    for r in out:
        assert all(v == 0 for v in r["unit"].values()), r["unit"]          # bitwise: halo values and rank-ordered sums
        assert r["rijke3d"]["max_abs_diff_vs_log"] < 2e-8, r["rijke3d"]     # the log prints 8 decimals
        assert r["annulus"]["rel_diff_vs_eigenvalues_dir"] < 1e-8
        assert r["rijke3d_p2"]["rel_diff_vs_oracle"] < 1e-8                 # degree 2 across ranks (dist.DofPartition)
        assert r["rijke3d"]["distributed_levels"] >= 1 and r["rijke3d"]["cycle_in_graph"]
    assert out[0]["annulus"]["omega"] == out[1]["annulus"]["omega"]          # replicated host logic: bit for bit


def test_peer_kernels_single_rank():
    """hx_peer_alloc / hx_peer_allreduce on a world of one (no other rank to wait for): the arena is a
    plain cudaMalloc with an IPC handle, the all-reduce is the identity and its sequence counter advances."""
    from helmholtz_x_b200 import _lib
    from helmholtz_x_b200.peer import AllreduceDesc, _Raw
    _lib.load()
    ptr = C.c_void_p()
    handle = (C.c_ubyte * 64)()
    nbytes = 1 << 20
    _lib.call("hx_peer_alloc", nbytes, C.byref(ptr), handle)
    assert any(handle)
    try:
        mem = torch.as_tensor(_Raw(ptr.value, nbytes), device="cuda")
        assert int(mem.sum()) == 0                                           # zero-filled
        d = AllreduceDesc()
        d.world, d.rank, d.slot_bytes = 1, 0, 65536
        d.slots[0] = ptr.value + 4096
        d.flags[0] = ptr.value
        d.my_flags = ptr.value
        d.seq = ptr.value + 1024
        d.block_counter = ptr.value + 2048
        d.err = ptr.value + 3072
        x = torch.randn(5000, dtype=torch.float64, device="cuda")
        y = torch.zeros_like(x)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            _lib.call("hx_peer_allreduce", C.byref(d), x.data_ptr(), y.data_ptr(), x.numel(), 0, st)
        torch.cuda.synchronize()
        assert torch.equal(x, y)
        assert int(mem[1024:1032].view(torch.int64)[0]) == 3                 # three all-reduces completed
        assert int(mem[3072:3076].view(torch.int32)[0]) == 0                 # no timeout
    finally:
        torch.cuda.synchronize()
        _lib.call("hx_peer_free", ptr)
