#!/usr/bin/env python
"""Multi-rank parity check (launch with torchrun): peer-memory primitives against torch references, then
the reference's config-1 (Rijke3D EPS FPI) and config-3 (full annulus PEP FPI, 16 pointwise flames) through
the public API on N ranks, compared with the golden logs.

One rank per GPU, NCCL for the set-up plumbing.  (Ranks must NOT share a GPU: the peer-memory kernels
of different ranks wait on one another, and nothing guarantees that two processes' kernels run at the same
time on one device -- B200_PROFILING.md reports Xid 109 for exactly that.)"""
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
import torch.distributed as dist  # noqa: E402


def unit_checks(rank, world):
    """hx_peer_allreduce and hx_peer_halo_exchange against plain torch arithmetic."""
    from helmholtz_x_b200.peer import HaloExchanger, PeerGroup
    grp = PeerGroup.get()
    dev = grp.device
    res = {}
    # all-reduce: complex128 (Gram column), float32 (restricted residual of the cycle), many repetitions
    for dtype, n in ((torch.complex128, 33), (torch.float64, 1), (torch.complex64, 70001)):
        bad = 0
        for rep in range(20):
            gen = torch.Generator().manual_seed(100 * rep + 7)
            parts = [torch.randn(n, dtype=torch.float64, generator=gen).to(dtype) * (q + 1) for q in range(world)]
            want = parts[0].clone()
            for q in range(1, world):
                want = want + parts[q]                       # rank order, like the kernel
            t = parts[rank].to(dev).contiguous()
            grp.allreduce_(t)
            bad += int(not torch.equal(t.cpu(), want))
        res[f"allreduce_{str(dtype).split('.')[-1]}_{n}_mismatches"] = bad
    # halo exchange on a ring: rank r owns n_own entries, needs the first m of rank r+1 and the last m of r-1
    n_own, m = 1000 + 10 * rank, 37
    nxt, prv = (rank + 1) % world, (rank - 1) % world
    recv = np.zeros(world, np.int64)
    send = np.zeros(world, np.int64)
    if world == 2:
        recv[nxt] = 2 * m
        send[nxt] = 2 * m
        n_own_nb = 1000 + 10 * nxt
        send_idx = torch.cat([torch.arange(m), torch.arange(n_own - m, n_own)])       # what the other rank reads
        want_ghost = lambda v_nb: torch.cat([v_nb[:m], v_nb[n_own_nb - m:n_own_nb]])  # noqa: E731
    else:
        recv[nxt] = m
        recv[prv] = m
        send[prv] = m
        send[nxt] = m
        chunks = {prv: torch.arange(m), nxt: torch.arange(n_own - m, n_own)}           # prv reads my head, nxt my tail
        send_idx = torch.cat([chunks[q] for q in sorted(chunks)])
    ex = HaloExchanger(world, rank, n_own, send_idx.to(dev), send, recv)
    n_loc = n_own + int(recv.sum())
    _, (mx,) = grp._all_min_max([n_loc])
    arena = grp.lease(2 * (mx * 16 + 512))
    bad = 0
    for dtype in (torch.complex128, torch.complex64):
        x = arena.take(mx, dtype, n_loc)
        for rep in range(30):
            vals = [torch.randn(1000 + 10 * q, dtype=torch.float64, generator=torch.Generator().manual_seed(rep * 31 + q)).to(dtype)
                    for q in range(world)]
            x.zero_()
            x[:n_own] = vals[rank].to(dev)
            ex.exchange(x)
            got = x[n_own:].cpu()
            if world == 2:
                want = want_ghost(vals[nxt])
            else:
                pieces = {nxt: vals[nxt][:m], prv: vals[prv][1000 + 10 * prv - m:]}
                want = torch.cat([pieces[q] for q in sorted(pieces)])                  # ghosts grouped by owner rank
            bad += int(not torch.equal(got, want))
    res["halo_mismatches"] = bad
    grp.release(arena)
    grp.check()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="rijke3d,annulus,rijke3d_p2")
    ap.add_argument("--no-unit", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if torch.cuda.device_count() < world:
        raise SystemExit("dist_check.py needs one GPU per rank")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    dist.barrier()
    from tests import cases
    from tests.gpu_helpers import gpu_flame, gpu_operators
    from helmholtz_x_b200 import peer
    from helmholtz_x_b200.eigensolvers import fixed_point_iteration
    from helmholtz_x_b200.eigenvectors import normalize_eigenvector
    G = cases.golden_values()
    out = {"world": world, "transport": peer.transport()}
    if not args.no_unit and peer.transport() == "peer":
        out["unit"] = unit_checks(rank, world)
    def rijke3d_p2():
        c = cases.rijke3d()
        c["degree"] = 2
        return c
    table = {"rijke3d": (cases.rijke3d, "rijke3d_active_fpi"), "annulus": (cases.annulus, "annulus_fpi_direct"),
             "rijke3d_p2": (rijke3d_p2, None)}
    for name in [c for c in args.cases.split(",") if c]:
        mk, gkey = table[name]
        case = mk()
        if gkey is None:
            # degree 2 across ranks (dist.DofPartition): no reference golden exists for P2, the
            # CPU oracle's fixed-point iteration on the same mesh is the target
            from oracle import hx_oracle as ox
            quiet = io.StringIO()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(quiet):
                mats = gpu_operators(case)
                D = gpu_flame(case)
                D.assemble_submatrices()
                E = fixed_point_iteration(mats, D, case.target, nev=2, i=0, tol=1e-8)
                omega, p = normalize_eigenvector(mats.mesh, E, 0, degree=2, matrices=mats, print_eigs=False)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            Eo, _ = ox.fixed_point_iteration(cases.oracle_operators(case), cases.oracle_flame(case), case.target, nev=2, i=0, tol=1e-8)
            out[name] = {"seconds": round(dt, 3), "omega": [omega.real, omega.imag], "n_own": mats.ops.n, "n_global": mats.ops.n_global,
                         "rel_diff_vs_oracle": abs(omega - Eo.omega(0)) / abs(Eo.omega(0)),
                         "stats": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in mats.ops.stats.items()}}
            del mats, D, E
            continue
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
        hier = mats.ops.hierarchy()
        out[name] = {"seconds": round(dt, 3), "omega": [omega.real, omega.imag], "max_abs_diff_vs_log": err,
                     "n_own": mats.ops.n, "distributed_levels": getattr(hier, "n_dist", None),
                     "cycle_in_graph": bool(getattr(hier, "use_graph", False)),
                     "stats": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in mats.ops.stats.items()}}
        if name == "annulus":
            g1 = cases.cplx(G["annulus_fpi_eigenvalues_dir"]["direct_1"])
            out[name]["rel_diff_vs_eigenvalues_dir"] = abs(E.getEigenpair(0) - g1) / abs(g1)
        del mats, D, E
    if peer.transport() == "peer":
        peer.PeerGroup.get().check()
    allout = [None] * world
    dist.all_gather_object(allout, out)
    if rank == 0:
        print(json.dumps(allout))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
