#!/usr/bin/env python
"""Standalone complex128 SpMV measurement on the synthetic annulus operator P(sigma):
CSR-vector (all lane widths) and SELL-32, CUDA events, GB/s with the 20*nnz+36*n model.
Used for the ncu captures under profiles/."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dofs", type=int, default=10_000_000)
    ap.add_argument("--degree", type=int, default=1)
    ap.add_argument("--launches", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from helmholtz_x_b200 import fem
    from helmholtz_x_b200.sell import SellMatrix
    be = fem.default_backend()
    g = bench.workload(a.dofs, a.degree)
    mesh = fem.Mesh(g["x"], g["cells"], g["cell_tags"], g["facets"], g["facet_tags"])
    V = fem.functionspace(mesh, ("Lagrange", a.degree))
    av, cv = fem.assemble_AC(V, g["c"])
    vals = be.empty(av.numel())
    be.combine_abc(av, None, cv, 1.0, 0.0, bench.TARGET ** 2, vals)
    M = V.matrix(vals)
    peak, src = bench.measured_peak()
    nbytes = 20.0 * M.nnz + 36.0 * M.n_rows
    x = torch.randn(M.n_cols, dtype=torch.float64, device=be.device).to(torch.complex128)
    y = be.zeros(M.n_rows)
    st = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {"n": M.n_rows, "nnz": M.nnz, "nnz_per_row": M.nnz / M.n_rows, "bytes": nbytes, "peak": peak, "peak_source": src}

    def timeit(fn):
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize()
        e0.record(st)
        for _ in range(a.launches):
            fn()
        e1.record(st)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.launches

    for lanes in (4, 8, 16, 32):
        ms = timeit(lambda: be.spmv(M, x, y, lanes=lanes))
        out[f"csr_lanes{lanes}"] = {"ms": round(ms, 4), "gbs": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / peak, 4)}
    S = SellMatrix.from_csr(be, M)
    for variant in range(6):
        ms = timeit(lambda: S.spmv(x, y, variant=variant))
        out[f"sell32_v{variant}"] = {"ms": round(ms, 4), "gbs": round(nbytes / ms / 1e6, 1),
                                     "frac": round(nbytes / ms / 1e6 / peak, 4), "padding": round(S.padding_ratio, 4)}
    # copy reference on the same box: y <- x (read 16n + write 16n)
    big = torch.empty(int(nbytes // 32), dtype=torch.complex128, device=be.device)
    big2 = torch.empty_like(big)
    ms = timeit(lambda: big2.copy_(big))
    out["torch_copy_same_bytes"] = {"ms": round(ms, 4), "gbs": round(big.numel() * 32 / ms / 1e6, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
