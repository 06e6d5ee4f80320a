#!/usr/bin/env python
"""A handful of launches of each hot kernel on the bench workload's operator, for one `ncu --set full`
capture: the complex128 SELL-32 SpMV (the roofline entry), the complex64 SELL-32 Jacobi sweep of the
multigrid cycle (the kernel the step spends most of its time in), the Gram-Schmidt pair (multi_dot,
multi_axpy) at k = 12, and the restart rotation on the FP64 tensor cores."""
import argparse
import contextlib
import io
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dofs", type=int, default=bench.DEFAULT_DOFS)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from helmholtz_x_b200 import fem
    from helmholtz_x_b200.acoustic_matrices import AcousticMatrices
    from helmholtz_x_b200.operators import ShiftedSolver
    be = fem.default_backend()
    g = bench.workload(a.dofs, 1)
    mesh, c_dev = bench.upload(g)
    c = fem.Function.from_device(fem.DG0Space(mesh), c_dev, name="soundspeed")
    with contextlib.redirect_stdout(io.StringIO()):
        mats = AcousticMatrices(mesh, fem.MeshTags(mesh.facet_tags), {11: {"Robin": -0.875 - 0.2j}}, c, degree=1)
    s = bench.TARGET
    solver = ShiftedSolver(mats.ops, {"A": 1.0, "B": s, "C": s ** 2})
    mg = solver.mg
    n = mats.ops.n
    gen = torch.Generator(be.device).manual_seed(0)
    b = torch.randn(n, dtype=torch.float64, device=be.device, generator=gen).to(torch.complex128)
    y = be.zeros(n)
    L0 = mg.levels[0]
    b32 = b.to(torch.complex64)
    basis = solver.basis
    k = 12
    for j in range(k + 1):
        basis.V[j].copy_(torch.randn(n, dtype=torch.float64, device=be.device, generator=gen).to(torch.complex128))
    Q = torch.view_as_complex(torch.randn(10, 19, 2, dtype=torch.float64, device=be.device, generator=gen)).contiguous()
    Vout = be.zeros(10, n)
    h = be.zeros(k + 2)
    nr = torch.view_as_real(h)[k]
    torch.cuda.synchronize()
    for _ in range(a.reps):
        be.spmv(solver.Pop, b, y)                                            # sell_kernel<4,4,0,double2>
        be.jacobi_sweep(L0.Mop, L0.dinv_w, b32, L0.x, L0.t, 0.6)             # sell_kernel<4,6,2,float2>
        be.multi_dot(basis.V, k, b, h)
        be.multi_axpy(basis.V, k, h, y, nrm2=nr)
        be.basis_rotate(basis.V, 19, Q, 10, Vout, tensor_cores=True)
    torch.cuda.synchronize()
    Pop = solver.Pop
    print(json.dumps({"n": n, "nnz": int(Pop.nnz), "k": k, "sell_c128_bytes_model": 20.0 * Pop.nnz + 36.0 * n,
                      "sell_c64_jacobi_bytes_model": 12.0 * Pop.nnz + 32.0 * n,
                      "multi_dot_bytes_model": 16.0 * n * (k + 1), "multi_axpy_bytes_model": 16.0 * n * (k + 2),
                      "rotate_bytes_model": 16.0 * n * 29}))


if __name__ == "__main__":
    main()
