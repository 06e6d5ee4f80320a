#!/usr/bin/env python
"""Multi-GPU parity check (launch with torchrun): the reference's config-1 (Rijke3D EPS FPI)
and config-3 (full annulus PEP FPI, 16 pointwise flames) through the public API on N ranks,
compared with the golden logs."""
import contextlib
import io
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    import __graft_entry__ as ge
    ge.build()
    from tests import cases
    from tests.gpu_helpers import gpu_flame, gpu_operators
    from helmholtz_x_b200.eigensolvers import fixed_point_iteration
    from helmholtz_x_b200.eigenvectors import normalize_eigenvector
    G = cases.golden_values()
    out = {"world": world}
    for name, mk, gkey in (("rijke3d", cases.rijke3d, "rijke3d_active_fpi"), ("annulus", cases.annulus, "annulus_fpi_direct")):
        case = mk()
        quiet = io.StringIO()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(quiet):
            mats = gpu_operators(case)
            D = gpu_flame(case)
            D.assemble_submatrices()
            E = fixed_point_iteration(mats, D, case.target, nev=case.nev, i=0, tol=case.tol)
            omega, p = normalize_eigenvector(mats.mesh, E, 0, degree=1, matrices=mats, print_eigs=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gold = [cases.cplx(q) for q in G[gkey]["omegas"]]
        hist = E.omega_history[-len(gold):]
        err = max(abs(a - b) for a, b in zip(hist, gold))
        out[name] = {"seconds": round(dt, 3), "omega": [omega.real, omega.imag], "max_abs_diff_vs_log": err,
                     "n_own": mats.ops.n, "stats": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in mats.ops.stats.items()}}
        if name == "annulus":
            g1 = cases.cplx(G["annulus_fpi_eigenvalues_dir"]["direct_1"])
            out[name]["rel_diff_vs_eigenvalues_dir"] = abs(E.getEigenpair(0) - g1) / abs(g1)
    allout = [None] * world
    dist.all_gather_object(allout, out)
    if rank == 0:
        print(json.dumps(allout))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
