"""N>1 host logic on CPU: world_size-2 gloo group, CPU test double as the per-rank
backend.  Covers the row partition, the halo exchange plan, all-reduced Gram columns,
block-Jacobi AMG, and the full eigen-solve / fixed-point iteration against the goldens."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _single_thread():
    """The matrices are tiny: BLAS / OpenMP teams of two ranks spinning against each other cost 10x
    more than the arithmetic (measured: 9 ms per 900 x 900 gemv)."""
    torch.set_num_threads(1)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except ImportError:
        pass


class _FakeVloc:
    """Local function space of a rank built from oracle matrices (the CUDA assembly is
    not available on CPU): pattern over the rank's local (owned+ghost) dofs."""

    def __init__(self, be, part, Aglob, x):
        import scipy.sparse as sp
        self.be, self.degree = be, 1
        sub = sp.csr_matrix(Aglob)[part.l2g][:, part.l2g].tocsr()
        sub.sort_indices()
        self._p = (torch.from_numpy(sub.indptr.astype(np.int32)), torch.from_numpy(sub.indices.astype(np.int32)))
        self.dof_coords = torch.from_numpy(np.ascontiguousarray(x[part.l2g]))

    def pattern(self):
        return self._p


def _sub_values(part, Mglob, pattern):
    import scipy.sparse as sp
    sub = sp.csr_matrix(Mglob)[part.l2g][:, part.l2g].tocsr()
    sub.sort_indices()
    # lay out on the (A-derived) local pattern
    n = sub.shape[0]
    P = sp.csr_matrix((np.arange(1, len(pattern[1]) + 1), pattern[1].numpy(), pattern[0].numpy()), shape=(n, n))
    out = np.zeros(len(pattern[1]), dtype=sub.dtype)
    coo = sub.tocoo()
    if coo.nnz == 0:
        return out
    pos = np.asarray(P[coo.row, coo.col]).ravel().astype(np.int64) - 1
    assert (pos >= 0).all()
    out[pos] = coo.data
    return out


def _synthetic_worker(rank, world, port, q, hierarchy=False):
    """Structured synthetic annulus, file-order ('input') partition = z-slabs: SpMV + PEP solve."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["HX_DIST_HIERARCHY"] = "1" if hierarchy else "0"
    os.environ["HX_DIST_MIN_ROWS"] = "50"           # small mesh: still distribute two levels
    _single_thread()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helmholtz_x_b200 import eigensolvers, synthetic
        from helmholtz_x_b200.dist import DistSpace, Partition
        from helmholtz_x_b200.operators import Mat, OperatorSet
        from oracle.host_backend import HostBackend
        from oracle import hx_oracle as ox
        g = synthetic.annulus_grid(5, 32, 14)
        m = ox.Mesh(g["x"], g["cells"].astype(np.int64), g["cell_tags"], g["facets"].astype(np.int64), g["facet_tags"])
        c = synthetic.annulus_sound_speed(g["x"], g["cells"])
        ops_o = ox.acoustic_matrices(m, {11: {"Robin": -0.875 - 0.2j}}, c, 1, c_is_dg0=True)
        part = Partition(m.x, m.cells, world, rank, "input", m.facets)
        assert len(part.neighbours()) == 1
        be = HostBackend()
        Vloc = _FakeVloc(be, part, ops_o.A, m.x)
        space = DistSpace(part, Vloc)
        vals = {k: torch.from_numpy(_sub_values(part, M, Vloc.pattern())) for k, M in
                (("A", ops_o.A.real), ("C", ops_o.C.real), ("B", ops_o.B))}
        ops = OperatorSet(space, space.own_values(vals["A"]), space.own_values(vals["C"]), space.own_values(vals["B"]))
        A, B, C = Mat(ops, {"A": 1.0}), Mat(ops, {"B": 1.0}), Mat(ops, {"C": 1.0})
        target = 3225.12 + 481.0j
        if hierarchy:
            ops.amg_options = {"coarse_max": 100}      # 2240 -> 140 -> 9 rows: two levels get distributed
        E = eigensolvers.pep_solver(A, B, C, target, nev=2)
        Eo = ox.pep_solve(ops_o.A, ops_o.B, ops_o.C, target, 2)
        assert abs(E.getEigenpair(0) - Eo.eigenvalues[0]) / abs(Eo.eigenvalues[0]) < 1e-8
        single = None
        if hierarchy:
            # the same eigen-solve on one rank: the row-distributed cycle is the single-rank cycle up to
            # summation order, so the inner iteration counts must agree
            assert ops.hierarchy() is not None and ops.amg() is ops.hierarchy().mg
            assert ops.hierarchy().n_dist == 2 and all(D.halo.n_ghost > 0 for D in ops.hierarchy().dl)
            from tests.host_helpers import HostSpace
            V1 = HostSpace(ops_o.space, HostBackend())
            ops1 = OperatorSet(V1, torch.from_numpy(ox._on_pattern(ops_o.space, ops_o.A).real.copy()),
                               torch.from_numpy(ox._on_pattern(ops_o.space, ops_o.C).real.copy()),
                               torch.from_numpy(ox._on_pattern(ops_o.space, ops_o.B)))
            ops1.amg_options = {"coarse_max": 100}
            E1 = eigensolvers.pep_solver(Mat(ops1, {"A": 1.0}), Mat(ops1, {"B": 1.0}), Mat(ops1, {"C": 1.0}), target, nev=2)
            assert abs(E1.getEigenpair(0) - E.getEigenpair(0)) / abs(E.getEigenpair(0)) < 1e-9
            single = ops1.stats["inner_iterations"]
            per_solve = ops.stats["inner_iterations"] / ops.stats["inner_solves"]
            per_solve1 = single / ops1.stats["inner_solves"]
            assert abs(per_solve - per_solve1) <= 0.05 * per_solve1 + 0.5, (ops.stats, ops1.stats)
            # the start vectors differ (each rank seeds its own slice), so one run may need a Krylov-Schur
            # restart the other does not: at most one basis (ncv = 19) of extra solves
            assert abs(ops.stats["inner_solves"] - ops1.stats["inner_solves"]) <= 20, (ops.stats, ops1.stats)
        q.put((rank, "ok", (ops.stats["inner_iterations"], single)))
    except Exception:      # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _worker(rank, world, port, ordering, q, hierarchy=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(rank)
    os.environ["HX_DIST_HIERARCHY"] = "1" if hierarchy else "0"
    _single_thread()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helmholtz_x_b200 import eigensolvers
        from helmholtz_x_b200.dist import DistSpace, Partition
        from helmholtz_x_b200.operators import LowRankMat, Mat, OperatorSet, build_lowrank
        from oracle.host_backend import HostBackend
        from oracle import hx_oracle as ox
        case = cases.rijke3d()
        m = case.mesh
        ops_o = cases.oracle_operators(case)
        part = Partition(m.x, m.cells.astype(np.int64), world, rank, ordering, m.facets)
        assert part.n_own > 0 and part.n_ghost > 0
        owned_total = torch.tensor([part.n_own])
        dist.all_reduce(owned_total)
        assert int(owned_total) == m.n_nodes
        be = HostBackend()
        Vloc = _FakeVloc(be, part, ops_o.A, m.x)
        space = DistSpace(part, Vloc)
        a = torch.from_numpy(_sub_values(part, ops_o.A.real, Vloc.pattern()))
        c = torch.from_numpy(_sub_values(part, ops_o.C.real, Vloc.pattern()))
        ops = OperatorSet(space, space.own_values(a), space.own_values(c), None)
        A, C = Mat(ops, {"A": 1.0}), Mat(ops, {"C": 1.0})
        # (i) distributed SpMV with halo exchange == global SpMV
        rng = np.random.default_rng(0)
        xg = rng.standard_normal(m.n_nodes) + 1j * rng.standard_normal(m.n_nodes)
        xl = torch.from_numpy(ops.to_local(xg))
        yl = torch.zeros(part.n_own, dtype=torch.complex128)
        A.apply(xl, yl)
        ref = (ops_o.A @ xg)[part.l2g[:part.n_own]]
        assert np.abs(yl.numpy() - ref).max() < 1e-12 * np.abs(ref).max()
        # (ii) all-reduced dot and gather
        out = torch.zeros(2, dtype=torch.complex128)
        ops.be.multi_dot(xl.view(1, -1), 1, xl, out)
        assert abs(out[0].item() - np.vdot(xg, xg)) < 1e-9 * abs(np.vdot(xg, xg))
        assert np.allclose(ops.to_global(xl), xg)
        # (iii) passive eigen-solve against the golden log
        G = cases.golden_values()
        hops = type("H", (), {})()
        c0 = ox.acoustic_matrices(m, case.bcs, case.c_passive, 1)
        a0 = torch.from_numpy(_sub_values(part, c0.A.real, Vloc.pattern()))
        ops0 = OperatorSet(space, space.own_values(a0), space.own_values(c), None)
        E = eigensolvers.eps_solver(Mat(ops0, {"A": 1.0}), Mat(ops0, {"C": 1.0}), case.target, nev=2)
        lam = np.array([E.getEigenvalue(i) for i in range(2)])
        gold = G["rijke3d_passive_eps"]["lambdas"][0]
        assert min(abs(lam - gold)) / gold < 1e-8
        # (iv) full fixed-point iteration with the flame vectors split over ranks
        fl = cases.oracle_flame(case)

        def owned_list(v):
            loc = v[part.l2g[:part.n_own]]
            idx = np.flatnonzero(loc).astype(np.int32)
            return [(idx, loc[idx])]
        L, R = owned_list(fl.left[:, 0]), owned_list(fl.right[:, 0])
        lr, lrT = build_lowrank(be, part.n_own, L, R), build_lowrank(be, part.n_own, R, L)

        class D:
            FTF = fl.FTF
            _D = LowRankMat(part.n_own, lr, lrT, 1.0, (L, R))
            matrix = None

            def assemble_matrix(self, omega, problem_type='direct'):
                self.matrix = self._D * self.FTF(omega)
        hops.A, hops.C, hops.B, hops.B_adj, hops.mesh = A, C, None, None, m
        Ef = eigensolvers.fixed_point_iteration(hops, D(), case.target, nev=2, i=0, tol=1e-8)
        gold = [cases.cplx(p) for p in G["rijke3d_active_fpi"]["omegas"]]
        assert len(Ef.omega_history) == len(gold)
        for u, v in zip(Ef.omega_history, gold):
            assert abs(u - v) < 2e-8 * abs(v) + 1e-8
        vr, _ = A.createVecs()
        Ef.getEigenvector(0, vr)
        assert vr.array.shape[0] == m.n_nodes and np.linalg.norm(vr.array) > 0
        q.put((rank, "ok", ops.stats["inner_iterations"]))
    except Exception as e:      # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _p2_worker(rank, world, port, q):
    """Degree 2 over two ranks: dist.DofPartition (vertex + edge dofs), permuted local space, SpMV, eigen-solve."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(rank)
    os.environ["HX_DIST_HIERARCHY"] = "1"
    _single_thread()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse as sp
        from helmholtz_x_b200 import eigensolvers, fem
        from helmholtz_x_b200.dist import DistSpace, DofPartition, Partition
        from helmholtz_x_b200.operators import Mat, OperatorSet
        from oracle.host_backend import HostBackend
        from oracle import hx_oracle as ox
        case = cases.rijke3d()
        case["degree"] = 2
        m = case.mesh
        ops_o = cases.oracle_operators(case, passive=True)
        be = HostBackend()
        gmesh = fem.Mesh(m.x, m.cells, m.cell_tags, m.facets, m.facet_tags, backend=be)
        part = Partition(gmesh.xd, gmesh.cellsd, world, rank, "morton", gmesh.facetsd)
        part.local_mesh = fem.Mesh(gmesh.xd[part.dev("l2g")], part.dev("local_cells"), gmesh.cell_tagsd[part.dev("cell_ids")],
                                   part.dev("local_facets"), gmesh.facet_tagsd[part.dev("facet_ids")], backend=be)
        Vglob, Vloc = fem.FunctionSpace(gmesh, 2), fem.FunctionSpace(part.local_mesh, 2)
        assert Vglob.n == ops_o.A.shape[0]
        dpart = DofPartition(part, Vglob, Vloc)
        # every global dof is owned exactly once
        owned = [None] * world
        dist.all_gather_object(owned, dpart.l2g[:dpart.n_own])
        allo = np.concatenate(owned)
        assert len(allo) == Vglob.n and len(np.unique(allo)) == Vglob.n
        assert np.all(np.diff(dpart.l2g[:dpart.n_own]) > 0)                       # ascending global ids
        assert dpart.n_ghost > 0 and (dpart.owner[dpart.l2g[dpart.n_own:]] != rank).all()
        gl_old = dpart.l2g[dpart.dof_inv_perm]                                    # global dof of every dof of the local space
        assert np.array_equal(dpart.g2l_old[gl_old], np.arange(len(gl_old)))       # Dirichlet dofs go through this map
        assert np.array_equal(dpart.g2l[dpart.l2g], np.arange(len(gl_old))) and (dpart.g2l_old >= 0).sum() == len(gl_old)
        assert np.array_equal(gl_old[:part.n_loc], part.l2g)                       # vertex dofs first, in sub-mesh order

        class FakeV:
            degree = 2

            def __init__(self, A):
                self.be = be
                sub = sp.csr_matrix(A)[gl_old][:, gl_old].tocsr()
                sub.sort_indices()
                self.sub = sub
                self._p = (torch.from_numpy(sub.indptr.astype(np.int32)), torch.from_numpy(sub.indices.astype(np.int32)))
                self.dof_coords = Vloc.dof_coords

            def pattern(self):
                return self._p
        pat_src = abs(ops_o.A) + abs(ops_o.C)
        V = FakeV(pat_src)

        def vals(Mg):
            sub = sp.csr_matrix(Mg)[gl_old][:, gl_old].tocsr()
            P = sp.csr_matrix((np.arange(1, V.sub.nnz + 1), V.sub.indices, V.sub.indptr), shape=V.sub.shape)
            out = np.zeros(V.sub.nnz)
            coo = sub.tocoo()
            out[np.asarray(P[coo.row, coo.col]).ravel().astype(np.int64) - 1] = coo.data.real
            return torch.from_numpy(out)
        space = DistSpace(dpart, V)
        ops = OperatorSet(space, space.own_values(vals(ops_o.A)), space.own_values(vals(ops_o.C)), None)
        A, C = Mat(ops, {"A": 1.0}), Mat(ops, {"C": 1.0})
        rng = np.random.default_rng(0)
        xg = rng.standard_normal(Vglob.n) + 1j * rng.standard_normal(Vglob.n)
        xl = torch.from_numpy(ops.to_local(xg))
        yl = torch.zeros(dpart.n_own, dtype=torch.complex128)
        A.apply(xl, yl)
        ref = (ops_o.A @ xg)[dpart.l2g[:dpart.n_own]]
        assert np.abs(yl.numpy() - ref).max() < 1e-12 * np.abs(ref).max()
        assert np.allclose(ops.to_global(xl), xg)
        E = eigensolvers.eps_solver(A, C, case.target, nev=2)
        Eo = ox.eps_solve(ops_o.A, -ops_o.C, case.target ** 2, 2)
        lam = np.array([E.getEigenvalue(i) for i in range(2)])
        for lo in Eo.eigenvalues[:2]:
            assert min(abs(lam - lo)) < 1e-8 * max(abs(lo), 1.0) + 1e-3, (lam, Eo.eigenvalues)
        assert ops.hierarchy() is not None
        q.put((rank, "ok", ops.stats["inner_iterations"]))
    except Exception:      # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_degree_two_partition():
    """P2 across ranks: edge dofs owned by the lower-ranked owner of their vertices (dist.DofPartition)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_p2_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in res:
        assert status == "ok", f"rank {rank}: {info}"


@pytest.mark.parametrize("ordering,hierarchy", [("morton", False), ("morton", True)])
def test_two_rank_partitioned_solve_matches_goldens(ordering, hierarchy):
    """'input' ordering is only meaningful for meshes whose node order is already local
    (the structured synthetic annulus); gmsh node order is not, so Morton is the default."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ordering, q, hierarchy)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in res:
        assert status == "ok", f"rank {rank}: {info}"
    print("inner iterations (distributed, single rank):", [info for _, _, info in res])
    print("inner GMRES iterations per rank:", [info for _, _, info in res])


@pytest.mark.parametrize("hierarchy", [False, True])
def test_two_rank_slab_partition_of_structured_annulus(hierarchy):
    """hierarchy=False: two-level Schwarz (default).  True: the row-distributed multigrid cycle
    (HX_DIST_HIERARCHY=1) -- same iteration count as one rank."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_synthetic_worker, args=(r, 2, port, q, hierarchy)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in res:
        assert status == "ok", f"rank {rank}: {info}"
    print("inner iterations (distributed, single rank):", [info for _, _, info in res])
