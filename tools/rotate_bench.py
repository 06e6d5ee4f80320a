#!/usr/bin/env python
"""Restart rotation Vout = Q^T V (n x m by m x k, complex128): FP64 CUDA-core kernel (basis_rotate_kernel)
against the FP64 tensor-core kernel (basis_rotate_dmma_kernel, DMMA), CUDA events, inputs far larger than L2.
Algorithmic bytes 16 n (m + k); flops 8 n m k.  One JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    import __graft_entry__ as ge
    ge.build()
    from helmholtz_x_b200 import fem
    be = fem.default_backend()
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
    out = {"hbm_peak_gbs": peak, "cases": []}
    g = torch.Generator(be.device).manual_seed(0)
    for n in (1_000_000, 10_000_000):
        for m, k in ((19, 10), (19, 2), (24, 12), (64, 32), (64, 8)):
            if n * (m + k) * 16 > 30e9:
                continue
            V = torch.randn(m, n, 2, dtype=torch.float64, device=be.device, generator=g)
            V = torch.view_as_complex(V)
            Q = torch.view_as_complex(torch.randn(k, m, 2, dtype=torch.float64, device=be.device, generator=g)).contiguous()
            Vout = be.zeros(k, n)
            rec = {"n": n, "m": m, "k": k, "bytes": 16.0 * n * (m + k), "flops": 8.0 * n * m * k}
            for name, tc in (("fma", False), ("dmma", True)):
                for _ in range(3):
                    be.basis_rotate(V, m, Q, k, Vout, tensor_cores=tc)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                reps = 20
                e0.record()
                for _ in range(reps):
                    be.basis_rotate(V, m, Q, k, Vout, tensor_cores=tc)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                rec[name] = {"ms": round(ms, 4), "gbs": round(rec["bytes"] / ms / 1e6, 1), "tflops": round(rec["flops"] / ms / 1e9, 2),
                             "frac_of_hbm_peak": round(rec["bytes"] / ms / 1e6 / peak, 3)}
            rec["dmma_speedup"] = round(rec["fma"]["ms"] / rec["dmma"]["ms"], 3)
            out["cases"].append(rec)
            del V, Vout
    print(json.dumps(out))


if __name__ == "__main__":
    main()
