#!/usr/bin/env python
"""One preconditioned inner solve (GMRES + SA-AMG V-cycle) on the synthetic annulus inside a
cudaProfilerStart/Stop range, for `ncu --profile-from-start off` launch lists."""
import argparse
import contextlib
import io
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dofs", type=int, default=1_000_000)
    ap.add_argument("--nu", type=int, default=2)
    ap.add_argument("--agg", type=int, default=8)
    ap.add_argument("--omega", type=float, default=2.0 / 3.0)
    ap.add_argument("--coarse-max", type=int, default=900)
    ap.add_argument("--precision", default="single")
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from helmholtz_x_b200 import fem
    from helmholtz_x_b200.acoustic_matrices import AcousticMatrices
    from helmholtz_x_b200.operators import ShiftedSolver
    be = fem.default_backend()
    g = bench.workload(a.dofs, 1)
    mesh = fem.Mesh(g["x"], g["cells"], g["cell_tags"], g["facets"], g["facet_tags"])
    c = fem.Function(fem.DG0Space(mesh), g["c"], dtype=np.float64, name="soundspeed")
    with contextlib.redirect_stdout(io.StringIO()):
        mats = AcousticMatrices(mesh, fem.MeshTags(mesh.facet_tags), {11: {"Robin": -0.875 - 0.2j}}, c, degree=1)
    s = bench.TARGET
    mats.ops.amg_options = dict(nu=a.nu, agg_size=a.agg, omega=a.omega, coarse_max=a.coarse_max, precision=a.precision)
    t0 = time.perf_counter()
    solver = ShiftedSolver(mats.ops, {"A": 1.0, "B": s, "C": s ** 2})
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    n = mats.ops.n
    b = torch.randn(n, dtype=torch.float64, device=be.device).to(torch.complex128)
    x = be.zeros(n)
    solver.solve(b, x)
    torch.cuda.synchronize()
    it0 = mats.ops.stats["inner_iterations"]
    be.reset_launch_count()
    torch.cuda.profiler.start()
    t0 = time.perf_counter()
    solver.solve(b, x)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    torch.cuda.profiler.stop()
    its = mats.ops.stats["inner_iterations"] - it0
    P = solver.P.to_scipy() if n <= 300000 else None
    true_res = None
    if P is not None:
        xb, bb = x.cpu().numpy(), b.cpu().numpy()
        true_res = float(np.linalg.norm(P @ xb - bb) / np.linalg.norm(bb))
    print(json.dumps({"precision": a.precision, "true_residual": true_res, "nu": a.nu, "agg": a.agg, "omega": a.omega, "n": n, "setup_s": round(t_setup, 3), "solve_s": round(dt, 4), "iterations": its,
                      "ms_per_iteration": round(dt / its * 1e3, 4), "launches": be.launch_count(),
                      "amg_sizes": solver.mg.sizes, "operator_complexity": round(solver.mg.operator_complexity, 3),
                      "stats": mats.ops.stats}))


if __name__ == "__main__":
    main()
