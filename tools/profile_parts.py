#!/usr/bin/env python
"""Where one preconditioned GMRES iteration spends its device time (CUDA events, warm, L2-exceeding
working set): multigrid cycle (graph replay and kernel by kernel, per level), operator apply, CGS2
step as a function of the basis size.  Writes one JSON line."""
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


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e6
    return {"gpu_us": round(e0.elapsed_time(e1) / reps * 1e3, 1), "wall_us": round(wall, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dofs", type=int, default=1_000_000)
    a = ap.parse_args()
    from helmholtz_x_b200 import fem, krylov
    from helmholtz_x_b200.acoustic_matrices import AcousticMatrices
    from helmholtz_x_b200.operators import ShiftedSolver
    be = fem.default_backend()
    g = bench.workload(a.dofs, 1)
    mesh = fem.Mesh(g["x"], g["cells"], g["cell_tags"], g["facets"], g["facet_tags"])
    c = fem.Function(fem.DG0Space(mesh), g["c"], dtype=np.float64, name="soundspeed")
    with contextlib.redirect_stdout(io.StringIO()):
        mats = AcousticMatrices(mesh, fem.MeshTags(mesh.facet_tags), {11: {"Robin": -0.875 - 0.2j}}, c, degree=1)
    s = bench.TARGET
    solver = ShiftedSolver(mats.ops, {"A": 1.0, "B": s, "C": s ** 2})
    mg = solver.mg
    n = mats.ops.n
    out = {"n": n, "amg_sizes": mg.sizes, "level_nnz": [L.pattern.nnz for L in mg.levels]}
    b = torch.randn(n, dtype=torch.float64, device=be.device).to(torch.complex128)
    x = be.zeros(n)
    y = be.zeros(n)
    solver.solve(b, x)
    out["vcycle_graph"] = timed(lambda: mg.apply(b, y), 100)
    L0 = mg.levels[0]
    out["vcycle_eager"] = timed(lambda: mg._cycle(0, L0.v_w), 100)
    for i in range(1, len(mg.levels)):
        Li = mg.levels[i]
        out[f"cycle_from_level{i}"] = timed(lambda i=i, Li=Li: mg._cycle(i, Li.b_), 100)
    out["fine_jacobi_sweep"] = timed(lambda: be.jacobi_sweep(L0.Mop, L0.dinv_w, L0.v_w, L0.x, L0.t, mg.omega), 200)
    out["fine_residual"] = timed(lambda: be.spmv(L0.Mop, L0.x, L0.r, alpha=-1.0, beta=1.0, y0=L0.v_w), 200)
    L1 = mg.levels[1]
    # SELL kernel variants (unroll, min blocks) for the complex64 sweep on levels 0 and 1
    for lvl, L in ((0, L0), (1, L1)):
        if getattr(L.Mop, "is_sell", False):
            keep = L.Mop.variant
            for v in (1, 2, 3, 4, 5):
                L.Mop.variant = v
                out[f"jacobi_c64_level{lvl}_variant{v}"] = timed(
                    lambda L=L: be.jacobi_sweep(L.Mop, L.dinv_w, L.b_ if lvl else L0.v_w, L.x, L.t, mg.omega), 200)
            L.Mop.variant = keep
    # level 1 as CSR-vector (complex64) with different sub-warp widths
    from helmholtz_x_b200.backend import CsrMatrix

    class Csr(CsrMatrix):
        lanes = 8

    p1 = L1.pattern
    csr1 = Csr(p1.n_rows, p1.n_cols, p1.indptr, p1.indices, L1.M.values.to(torch.complex64))
    for lanes in (8, 16, 32):
        csr1.lanes = lanes
        out[f"jacobi_c64_level1_csr_lanes{lanes}"] = timed(
            lambda: be.jacobi_sweep(csr1, L1.dinv_w, L1.b_, L1.x, L1.t, mg.omega), 200)
    L2 = mg.levels[2]
    if len(mg.levels) > 3:
        p2 = L2.pattern
        csr2 = Csr(p2.n_rows, p2.n_cols, p2.indptr, p2.indices, L2.M.values.to(torch.complex64))
        out["level2_default_lanes"] = int(L2.Mop.lanes) if hasattr(L2.Mop, "lanes") else None
        for lanes in (8, 16, 32):
            csr2.lanes = lanes
            out[f"jacobi_c64_level2_csr_lanes{lanes}"] = timed(
                lambda: be.jacobi_sweep(csr2, L2.dinv_w, L2.b_, L2.x, L2.t, mg.omega), 200)
    Lc = mg.levels[-1]
    out["coarse_solve"] = timed(lambda: mg._cycle(len(mg.levels) - 1, Lc.b_), 200)
    out["coarse_gemv_only"] = timed(lambda: be.dense_gemv(mg.coarse_inv, Lc.b64, Lc.x64), 200)
    out["fine_restrict"] = timed(lambda: be.spmv(L0.R, L0.r, L1.b_), 200)
    out["fine_prolong"] = timed(lambda: be.spmv(L0.P, L1.x, L0.x, alpha=1.0, beta=1.0, y0=L0.x), 200)
    out["P_nnz"] = int(L0.P.nnz)
    out["operator_apply_c128"] = timed(lambda: be.spmv(solver.Pop, b, y), 200)
    basis = solver.basis
    basis.V.copy_(torch.randn(basis.V.shape, dtype=torch.float64, device=be.device).to(torch.complex128) * 1e-3)
    for j in (0, 7, 15, 23, 31, 47, 62):
        def step(j=j):
            basis.w.copy_(b)
            return basis.orthogonalize(j)
        out[f"cgs2_j{j}"] = timed(step, 30)
    out["w_copy"] = timed(lambda: basis.w.copy_(b), 100)
    it0 = mats.ops.stats["inner_iterations"]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    solver.solve(b, x)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    its = mats.ops.stats["inner_iterations"] - it0
    out["solve"] = {"iterations": its, "ms_per_iteration": round(dt / its * 1e3, 4)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
