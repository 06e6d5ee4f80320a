"""Host-side logic of the product (Krylov-Schur, GMRES, AMG cycle order, Woodbury flame
term, fixed-point recurrences) on the CPU test double -- no GPU, no CUDA kernels."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla
import torch

from helmholtz_x_b200 import eigensolvers, krylov
from oracle import hx_oracle as ox
from oracle.host_backend import HostBackend
from tests import cases
from tests.host_helpers import HostFlame, HostOperators

G = cases.golden_values()


@pytest.fixture(scope="module")
def rijke():
    case = cases.rijke3d()
    return case, HostOperators(case)


def test_krylov_schur_matches_passive_golden(rijke):
    """numerical_examples/Longitudinal/NetworkCode/RijkeTube3D/Results/Passive/passive.log:30-33"""
    case, _ = rijke
    hops = HostOperators(case, passive=True)
    E = eigensolvers.eps_solver(hops.A, hops.C, case.target, nev=2)
    lam = np.sort_complex(np.array([E.getEigenvalue(i) for i in range(2)]))
    gold = np.array(G["rijke3d_passive_eps"]["lambdas"])
    assert abs(lam[1] - gold[0]) / gold[0] < 1e-9
    assert abs(lam[0]) < 1e-3      # the null (constant) mode


def test_krylov_schur_restart_path():
    rng = np.random.default_rng(1)
    n = 300
    d = np.concatenate([[50, 40, 30, 20], rng.uniform(0.1, 5, n - 4)]) * np.exp(1j * rng.uniform(0, 0.3, n))
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    Mx = (Q * d) @ Q.conj().T
    be = HostBackend()
    res = krylov.krylov_schur(be, lambda v, o: o.copy_(torch.from_numpy(Mx @ v.numpy())), n, nev=4, ncv=10, tol=1e-12)
    assert res.its > 1 and res.nconv >= 4
    assert np.allclose(np.sort_complex(res.theta[:4]), np.sort_complex(d[:4]), rtol=1e-10)
    for i in range(4):
        x = res.X[i].numpy()
        assert np.linalg.norm(Mx @ x - res.theta[i] * x) < 1e-8 * abs(res.theta[i])


def test_gmres_restarts_and_true_residual(rijke):
    _, hops = rijke
    be = hops.ops.be
    P = hops.ops.space.matrix(hops.ops.combine({"A": 1.0, "C": (400 * np.pi) ** 2}))
    Ps = P.to_scipy()
    ilu = spla.spilu(Ps.tocsc(), drop_tol=1e-2, fill_factor=2)
    b = torch.randn(P.n_rows, dtype=torch.float64, generator=torch.Generator().manual_seed(0)).to(torch.complex128)
    x = torch.zeros_like(b)
    its, rel = krylov.gmres(be, lambda v, o: be.spmv(P, v, o), b, x, rtol=1e-10, restart=30, maxiter=400,
                            precond=lambda v, o: o.copy_(torch.from_numpy(ilu.solve(v.numpy()))))
    assert 30 < its < 400, "restart path not exercised / stagnation"
    assert np.linalg.norm(Ps @ x.numpy() - b.numpy()) / np.linalg.norm(b.numpy()) < 2e-10


def test_amg_preconditioned_solve(rijke):
    _, hops = rijke
    mg = hops.ops.amg()
    assert mg.sizes[0] == 2426 and len(mg.sizes) >= 2
    from helmholtz_x_b200.operators import ShiftedSolver
    s = ShiftedSolver(hops.ops, {"A": 1.0, "C": (400 * np.pi) ** 2}, rtol=1e-11)
    b = torch.ones(hops.ops.n, dtype=torch.complex128)
    x = torch.zeros_like(b)
    s.solve(b, x)
    Ps = s.P.to_scipy()
    assert np.linalg.norm(Ps @ x.numpy() - b.numpy()) / np.linalg.norm(b.numpy()) < 1e-10
    assert hops.ops.stats["inner_iterations"] < 40


def test_woodbury_flame_term_matches_densified(rijke):
    case, hops = rijke
    D = HostFlame(case, hops)
    om = 1248.5 + 3.4j
    D.assemble_matrix(om)
    K = hops.A - D.matrix
    from helmholtz_x_b200.operators import ShiftedSolver
    terms = {"A": 1.0, "C": (400 * np.pi) ** 2}
    s = ShiftedSolver(hops.ops, terms, K.lowrank, rtol=1e-12)
    b = torch.randn(hops.ops.n, dtype=torch.float64, generator=torch.Generator().manual_seed(2)).to(torch.complex128)
    x = torch.zeros_like(b)
    s.solve(b, x)
    dense = (K.to_scipy() + (400 * np.pi) ** 2 * hops.C.to_scipy()).tocsc()      # densified reference operator
    ref = spla.spsolve(dense, b.numpy())
    assert np.linalg.norm(x.numpy() - ref) / np.linalg.norm(ref) < 1e-8
    # matrix-free apply == densified apply
    y = torch.zeros_like(b)
    K.apply(b, y)
    assert np.allclose(y.numpy(), K.to_scipy() @ b.numpy(), rtol=1e-12, atol=1e-9)


def test_fpi_eps_matches_golden_log(rijke, capsys):
    """config 1: .../RijkeTube3D/Results/Active/active.log:23-52 through the product's
    fixed_point_iteration (host logic) on the CPU double."""
    case, hops = rijke
    D = HostFlame(case, hops)
    E = eigensolvers.fixed_point_iteration(hops, D, case.target, nev=2, i=0, tol=1e-8)
    gold = [cases.cplx(p) for p in G["rijke3d_active_fpi"]["omegas"]]
    hist = E.omega_history
    assert len(hist) == len(gold)
    for a, b in zip(hist, gold):
        assert abs(a - b) < 2e-8 * abs(b) + 1e-8
    out = capsys.readouterr().out
    assert "+ Starting eigenvalue is found: +1249.66494052  +0.00000000j." in out
    assert "* iter =  4" in out


def test_fpi_pep_direct_and_adjoint_match_golden():
    """PRF Rijke3D (Robin, PEP): .../PRF/RijkeTube3D/Results/Active/active.log:21-50,57-86"""
    case = cases.prf_rijke3d()
    hops = HostOperators(case)
    D = HostFlame(case, hops)
    E = eigensolvers.fixed_point_iteration(hops, D, case.target, nev=2, i=0)
    gold = [cases.cplx(p) for p in G["prf_rijke3d_direct_fpi"]["omegas"]]
    for a, b in zip(E.omega_history[1:], gold):
        assert abs(a - b) < 2e-8
    E2 = eigensolvers.fixed_point_iteration(hops, D, case.target, nev=2, i=0, problem_type='adjoint')
    gold = [cases.cplx(p) for p in G["prf_rijke3d_adjoint_fpi"]["omegas"]]
    for a, b in zip(E2.omega_history[1:], gold):
        assert abs(a - b) < 2e-8


def test_two_sided_left_vectors(rijke):
    case, hops = rijke
    D = HostFlame(case, hops)
    om = 1247.4 + 6.8j
    D.assemble_matrix(om)
    L = hops.A + om ** 2 * hops.C - D.matrix
    E = eigensolvers.eps_solver(L, -hops.C, 0, 2, two_sided=True)
    Ls, Cs = L.to_scipy(), hops.C.to_scipy()
    lam = E.getEigenvalue(0)
    y = E.device_vector(0, "left").numpy()
    x = E.device_vector(0, "right").numpy()
    assert np.linalg.norm(Ls @ x - lam * (Cs @ x)) < 1e-7 * np.linalg.norm(Ls @ x)
    r = y.conj() @ Ls - lam * (y.conj() @ Cs)
    assert np.linalg.norm(r) < 1e-7 * np.linalg.norm(y.conj() @ Ls)


def test_native_spgemm_coarsening_path_matches_library_path(rijke, monkeypatch):
    """amg._coarsen_native (hx_spgemm_* kernels on the GPU) against the torch.sparse path, with the
    kernels replaced by SciPy stand-ins: checks the prolongator assembly, transpose and the shared
    Galerkin pattern logic on CPU."""
    import scipy.sparse as sp
    from helmholtz_x_b200 import amg, spgemm
    from helmholtz_x_b200.backend import CsrMatrix

    def to_sp(M, vals=None):
        v = M.values if vals is None else vals
        v = np.ones(M.indices.numel()) if v is None else v.numpy()
        return sp.csr_matrix((v, M.indices.numpy(), M.indptr.numpy()), shape=M.shape)

    def fake_symbolic(be, A, B):
        C = (abs(to_sp(A, torch.ones(A.indices.numel(), dtype=torch.float64)))
             @ abs(to_sp(B, torch.ones(B.indices.numel(), dtype=torch.float64)))).tocsr()
        C.sort_indices()
        return torch.from_numpy(C.indptr.astype(np.int32)), torch.from_numpy(C.indices.astype(np.int32))

    def fake_numeric(be, A, B, indptr, indices, out=None):
        C = (to_sp(A) @ to_sp(B)).tocsr()
        n = len(indptr) - 1
        Pm = sp.csr_matrix((np.arange(1, len(indices) + 1), indices.numpy(), indptr.numpy()), shape=(n, B.n_cols))
        coo = C.tocoo()
        vals = np.zeros(len(indices))
        vals[np.asarray(Pm[coo.row, coo.col]).ravel() - 1] = coo.data
        return torch.from_numpy(vals)

    monkeypatch.setattr(spgemm, "symbolic", fake_symbolic)
    monkeypatch.setattr(spgemm, "numeric", fake_numeric)
    _, hops = rijke
    be = hops.ops.be
    pat = hops.V.matrix
    A, C = pat(hops.ops.base["A"]), pat(hops.ops.base["C"])
    ref = amg.AMG(be, A, C, None, hops.V.dof_coords, agg_size=8)

    class NativeAMG(amg.AMG):          # force the native path on this small problem
        native_min_rows = property(lambda self: 0, lambda self, v: None)
    be.supports_spgemm = True
    nat = NativeAMG(be, A, C, None, hops.V.dof_coords, agg_size=8)
    be.supports_spgemm = False
    assert nat.sizes == ref.sizes and nat.native_levels == len(nat.sizes) - 1
    for Ln, Lr in zip(nat.levels[:-1], ref.levels[:-1]):
        assert abs(to_sp(Ln.P) - to_sp(Lr.P)).max() < 1e-12
        assert abs(to_sp(Ln.R) - to_sp(Lr.R)).max() < 1e-12
    a1n = to_sp(nat.levels[1].pattern, nat.levels[1].a)
    a1r = to_sp(ref.levels[1].pattern, ref.levels[1].a)
    assert abs(a1n - a1r).max() < 1e-9 * abs(a1r).max()
    c1n = to_sp(nat.levels[1].pattern, nat.levels[1].c)
    c1r = to_sp(ref.levels[1].pattern, ref.levels[1].c)
    assert abs(c1n - c1r).max() < 1e-9 * abs(c1r).max()
    # impedance matrix B (nonzero on the Robin boundary only: compacted Galerkin product) and the prolongator filter
    hp = HostOperators(cases.prf_rijke3d())
    bp = hp.ops.be
    A, C, B = (hp.V.matrix(hp.ops.base[k]) for k in ("A", "C", "B"))
    assert int(torch.count_nonzero(B.values)) * 8 < B.values.numel()
    ref = amg.AMG(bp, A, C, B, hp.V.dof_coords, agg_size=8, p_filter=0.1)
    bp.supports_spgemm = True
    nat = NativeAMG(bp, A, C, B, hp.V.dof_coords, agg_size=8, p_filter=0.1)
    bp.supports_spgemm = False
    assert nat.sizes == ref.sizes
    assert nat.levels[0].P.nnz == ref.levels[0].P.nnz < amg.AMG(bp, A, C, B, hp.V.dof_coords, agg_size=8, p_filter=0.0).levels[0].P.nnz
    for name in ("a", "c", "b"):
        mn = to_sp(nat.levels[1].pattern, getattr(nat.levels[1], name))
        mr = to_sp(ref.levels[1].pattern, getattr(ref.levels[1], name))
        assert abs(mn - mr).max() < 1e-9 * abs(mr).max(), name
    ones = torch.ones(ref.levels[1].n, dtype=torch.float64)
    rowsum = to_sp(ref.levels[0].P) @ ones.numpy()
    unf = to_sp(amg.AMG(bp, A, C, B, hp.V.dof_coords, agg_size=8, p_filter=0.0).levels[0].P) @ ones.numpy()
    assert np.allclose(rowsum, unf, rtol=1e-12, atol=1e-14)            # the filter keeps every row sum


def test_bloch_reduction_host_logic():
    """config 4 (Micca sector, N = 16): the product's Blochifier (relabelled CSR entries, no BN/NB
    products) against the oracle's NB*M*BN, then the reference's passive and active drivers
    (bloch/passive.py, bloch/active.py) through eps_solver / fixed_point_iteration on the CPU double."""
    from helmholtz_x_b200.bloch_operator import Blochifier
    from helmholtz_x_b200.fem import _Vec
    case = cases.bloch()
    hops = HostOperators(case)
    sp_ = hops.oracle.space
    numb = cases.bloch_numbering(sp_)
    for pairing in ("sorted", "geometric"):
        bl = Blochifier(case.mesh, case.bcs, case.N, hops, pairing=pairing, numbering=numb)
        md, sd = ox.bloch_pairs(sp_, case.master, case.slave, case.N, pairing=pairing, numbering=numb)
        assert np.array_equal(bl.dofs_master, md) and np.array_equal(bl.dofs_slave, sd)
        BN, NB = ox.bloch_maps(sp_.n, md, sd, case.N)
        for name, M in (("A", hops.oracle.A), ("B", hops.oracle.B), ("C", hops.oracle.C)):
            want = ox.blochify(M, BN, NB)
            got = getattr(bl, name).to_scipy()
            assert abs(got - want).max() <= 1e-13 * abs(want).max(), (pairing, name)
        assert bl.B_adj is None
        # remapper = BN
        x = np.random.default_rng(0).standard_normal(bl.n_red) + 0j
        xin, y = bl.remapper.createVecs()
        xin.setArray(x)
        bl.remapper.mult(xin, y)
        assert np.allclose(y.array, BN @ x, rtol=0, atol=1e-14)
    # passive: bloch/Results/Passive/passive.log:27-31 (sorted pairing = the reference's)
    bl = Blochifier(case.mesh, case.bcs, case.N, hops, pairing="sorted", numbering=numb)
    E = eigensolvers.eps_solver(bl.A, bl.C, case.passive_target, nev=case.passive_nev, print_results=False)
    for k, g in enumerate(G["bloch_passive"]["omegas"]):
        assert abs(np.sqrt(E.getEigenvalue(k)) - g) < 1e-6 + 1e-9 * g
    # active: bloch/Results/Active/active.log:38-75
    D = HostFlame(case, hops)
    D.blochify(bl)
    assert D._D_ij.n == bl.n_red
    E = eigensolvers.fixed_point_iteration(bl, D, case.target, nev=case.nev, i=0, tol=case.tol)
    gold = [cases.cplx(p) for p in G["bloch_active_fpi"]["omegas"]]
    hist = E.omega_history
    assert len(hist) == len(gold) + 1
    for a, b in zip(hist[1:], gold):
        assert abs(a - b) < 6e-4
    vr, vi = bl.A.createVecs()
    omega = E.getEigenpair(0, vr, vi)
    assert abs(omega - cases.cplx(G["bloch_active_fpi"]["final"])) < 2e-6
    with pytest.raises(NotImplementedError):
        eigensolvers.eps_solver(bl.A, bl.C, case.passive_target, nev=2, two_sided=True)


def test_bloch_reduction_p2_geometric_pairing():
    """P2 on the sector: master/slave edge dofs are paired as rotation images as well, and the reduced
    operators equal the oracle's NB*M*BN."""
    from helmholtz_x_b200.bloch_operator import Blochifier
    case = cases.bloch()
    case["degree"] = 2
    hops = HostOperators(case)
    sp_ = hops.oracle.space
    bl = Blochifier(case.mesh, case.bcs, case.N, hops)
    md, sd = ox.bloch_pairs(sp_, case.master, case.slave, case.N, pairing="geometric")
    assert len(md) > 311 and np.array_equal(bl.dofs_master, md) and np.array_equal(bl.dofs_slave, sd)
    X = sp_.dof_x
    a = 2 * np.pi / case.N
    Rm = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    d = min(np.abs(X[md] @ Rm.T - X[sd]).max(), np.abs(X[md] @ Rm - X[sd]).max())
    assert d < 1e-9
    BN, NB = ox.bloch_maps(sp_.n, md, sd, case.N)
    for name, M in (("A", hops.oracle.A), ("B", hops.oracle.B), ("C", hops.oracle.C)):
        want = ox.blochify(M, BN, NB)
        got = getattr(bl, name).to_scipy()
        assert abs(got - want).max() <= 1e-13 * abs(want).max(), name
    assert abs(want - want.conj().T).max() <= 1e-13 * abs(want).max()       # C_b Hermitian


def test_single_pass_gram_schmidt_and_verified_residual(rijke, monkeypatch):
    """Inner GMRES: one Gram-Schmidt pass per step (the default) against two -- same iteration counts,
    true residual below rtol; a shift next to an eigenvalue floors above rtol and must return at the
    floor instead of spinning to maxiter."""
    import helmholtz_x_b200.operators as O
    case, hops = rijke
    ops = hops.ops
    s2 = case.target ** 2
    rng = np.random.default_rng(11)
    b = torch.from_numpy(rng.standard_normal(ops.n) + 1j * rng.standard_normal(ops.n))
    counts = {}
    for passes in (2, 1):
        monkeypatch.setattr(O, "GMRES_ORTH_PASSES", passes)
        ops._shift_state = None
        solver = O.ShiftedSolver(ops, {"A": 1.0, "C": s2})
        x = torch.zeros_like(b)
        i0 = ops.stats["inner_iterations"]
        solver.solve(b, x)
        counts[passes] = ops.stats["inner_iterations"] - i0
        P = solver.P.to_scipy()
        assert np.linalg.norm(P @ x.numpy() - b.numpy()) / np.linalg.norm(b.numpy()) < 1.05e-11
    assert abs(counts[1] - counts[2]) <= 2, counts
    # the second solve of this eigen-solve floors at ~5e-11 (> rtol): it must return there after a short
    # refinement cycle (34 iterations), not run to maxiter = 512
    monkeypatch.setattr(O, "GMRES_ORTH_PASSES", 1)
    ops._shift_state = None
    i0, s0 = ops.stats["inner_iterations"], ops.stats["inner_solves"]
    eigensolvers.eps_solver(hops.A, hops.C, case.target, nev=2)
    n_solves = ops.stats["inner_solves"] - s0
    assert ops.stats["inner_iterations"] - i0 < 40 * n_solves, ops.stats


def test_flexible_gmres_with_complex64_cycle_on_the_cpu_double(monkeypatch):
    """The CUDA backend runs the multigrid cycle in complex64 inside flexible GMRES; the CPU double does
    the same when asked, and the complex128 residual is still reached in the same number of iterations."""
    from helmholtz_x_b200.operators import ShiftedSolver
    monkeypatch.setattr(HostBackend, "supports_mixed", True)
    case = cases.prf_rijke3d()
    hops = HostOperators(case)
    s = case.target
    solver = ShiftedSolver(hops.ops, {"A": 1.0, "B": s, "C": s ** 2}, rtol=1e-11)
    assert solver.mg.single and solver.st["zbasis"] is not None
    rng = np.random.default_rng(12)
    b = torch.from_numpy(rng.standard_normal(hops.ops.n) + 1j * rng.standard_normal(hops.ops.n))
    x = torch.zeros_like(b)
    solver.solve(b, x)
    P = solver.P.to_scipy()
    assert np.linalg.norm(P @ x.numpy() - b.numpy()) / np.linalg.norm(b.numpy()) < 1.05e-11
    assert hops.ops.stats["inner_iterations"] < 45


def test_w_cycle_option_reduces_iterations_and_keeps_the_solution(monkeypatch):
    """AMG(w_from=1): levels >= 1 visited twice.  Same solution, fewer GMRES iterations (the option is
    off by default until it is timed on the GPU)."""
    from helmholtz_x_b200.operators import ShiftedSolver
    monkeypatch.delenv("HX_AMG_WCYCLE", raising=False)
    case = cases.annulus()
    rng = np.random.default_rng(21)
    its, sols = {}, {}
    for w_from in (None, 1):
        hops = HostOperators(case)
        hops.ops.amg_options = {"w_from": w_from}
        s = case.target
        solver = ShiftedSolver(hops.ops, {"A": 1.0, "B": s, "C": s ** 2})
        assert len(solver.mg.levels) >= 3 and solver.mg.w_from == w_from
        if w_from is None:
            b = torch.from_numpy(rng.standard_normal(hops.ops.n) + 1j * rng.standard_normal(hops.ops.n))
        x = torch.zeros_like(b)
        solver.solve(b, x)
        its[w_from], sols[w_from] = hops.ops.stats["inner_iterations"], x.numpy().copy()
    assert its[1] < its[None], its
    assert np.linalg.norm(sols[1] - sols[None]) / np.linalg.norm(sols[None]) < 1e-9


def test_chebyshev_damping_option(rijke):
    """AMG(smoother="chebyshev"): per-level, per-shift estimate of rho(D^-1 P_l) and the Chebyshev roots
    as damping factors of the nu sweeps -- same sweeps, fewer iterations, same solution."""
    from helmholtz_x_b200.operators import ShiftedSolver
    case, _ = rijke
    rng = np.random.default_rng(22)
    out = {}
    for sm in ("jacobi", "chebyshev"):
        hops = HostOperators(case)
        hops.ops.amg_options = {"smoother": sm}
        solver = ShiftedSolver(hops.ops, {"A": 1.0, "C": case.target ** 2})
        if sm == "jacobi":
            b = torch.from_numpy(rng.standard_normal(hops.ops.n) + 1j * rng.standard_normal(hops.ops.n))
        x = torch.zeros_like(b)
        solver.solve(b, x)
        out[sm] = (hops.ops.stats["inner_iterations"], x.numpy().copy())
        if sm == "chebyshev":
            L0 = solver.mg.levels[0]
            assert 1.5 < L0.rho < 3.5 and len(L0.omegas) == 2 and L0.omegas[0] < 2.0 / 3.0 < L0.omegas[1]
            # the polynomial is a contraction on [rho/8, rho]
            t = np.linspace(L0.rho / 8, L0.rho, 50)
            assert np.abs((1 - L0.omegas[0] * t) * (1 - L0.omegas[1] * t)).max() < 0.45
    assert out["chebyshev"][0] < out["jacobi"][0], (out["chebyshev"][0], out["jacobi"][0])
    assert np.linalg.norm(out["chebyshev"][1] - out["jacobi"][1]) / np.linalg.norm(out["jacobi"][1]) < 1e-9


def test_relaxed_inner_tolerance_reproduces_golden_with_fewer_iterations(monkeypatch):
    """HX_INNER_RELAX: inner solves of late Arnoldi steps stop at INNER_RTOL / (Ritz residual); the PRF
    golden log (8 decimals) is reproduced with about a third fewer inner iterations."""
    case = cases.prf_rijke3d()
    gold = [cases.cplx(p) for p in G["prf_rijke3d_direct_fpi"]["omegas"]]
    its = {}
    for relax in (False, True):
        monkeypatch.setattr(eigensolvers, "INNER_RELAX", relax)
        hops = HostOperators(case)
        D = HostFlame(case, hops)
        E = eigensolvers.fixed_point_iteration(hops, D, case.target, nev=2, i=0)
        for a, b in zip(E.omega_history[1:], gold):
            assert abs(a - b) < 2e-8
        its[relax] = hops.ops.stats["inner_iterations"]
    assert its[True] < 0.8 * its[False], its


@pytest.mark.parametrize("mixed", [False, True])
def test_shift_next_to_an_eigenvalue_is_solved_to_backward_stability(monkeypatch, mixed):
    """shift-invert with sigma at relative distance 1e-6 / 1e-9 / 0 from an eigenvalue
    (helmholtz_x/eigensolvers.py:41-67 with an exact LU simply works there).  ||b - P x|| / ||b|| cannot
    reach 1e-11 next to a singular P; the inner solve stops on the normwise backward error instead and the
    eigenvalue still comes out to 1e-9."""
    monkeypatch.setattr(HostBackend, "supports_mixed", mixed)
    case = cases.rijke3d()
    hops = HostOperators(case, passive=True)
    gold = np.sqrt(G["rijke3d_passive_eps"]["lambdas"][0])
    E0 = eigensolvers.eps_solver(hops.A, hops.C, case.target, nev=2)
    om0 = min((np.sqrt(E0.getEigenvalue(i)) for i in range(2)), key=lambda z: abs(z - gold))
    assert abs(om0 - gold) < 1e-9 * gold
    for delta in (1e-6, 1e-9, 0.0):
        hops.ops._shift_state = None
        s0, i0 = hops.ops.stats["inner_solves"], hops.ops.stats["inner_iterations"]
        E = eigensolvers.eps_solver(hops.A, hops.C, om0.real * (1 + delta), nev=2)
        om = min((np.sqrt(E.getEigenvalue(i)) for i in range(2)), key=lambda z: abs(z - gold))
        assert abs(om - gold) < 1e-9 * gold, (delta, om, gold)
        n_solves = hops.ops.stats["inner_solves"] - s0
        assert hops.ops.stats["inner_iterations"] - i0 < 120 * n_solves, (delta, hops.ops.stats)
    assert hops.ops.stats.get("floor_accepts", 0) > 0            # the backward-error exit was actually taken


def test_gmres_reports_the_true_residual_on_every_exit(rijke):
    """ADVICE r1 (krylov.py:187): maxiter exits return a recomputed residual, never the recurrence estimate."""
    _, hops = rijke
    be = hops.ops.be
    P = hops.ops.space.matrix(hops.ops.combine({"A": 1.0, "C": (400 * np.pi) ** 2}))
    Ps = P.to_scipy()
    b = torch.randn(P.n_rows, dtype=torch.float64, generator=torch.Generator().manual_seed(3)).to(torch.complex128)
    x = torch.zeros_like(b)
    info = {}
    its, rel = krylov.gmres(be, lambda v, o: be.spmv(P, v, o), b, x, rtol=1e-12, restart=10, maxiter=25, info=info)
    assert its == 25 and info["status"] == "maxiter"
    true = np.linalg.norm(Ps @ x.numpy() - b.numpy()) / np.linalg.norm(b.numpy())
    assert abs(rel - true) <= 1e-12 * max(true, 1.0) + 1e-14
    # continuing from that x (x0=True) picks up where it stopped
    its2, rel2 = krylov.gmres(be, lambda v, o: be.spmv(P, v, o), b, x, rtol=1e-12, restart=10, maxiter=25, x0=True)
    assert rel2 < rel


def test_newton_to_a_tight_tolerance_on_the_singular_operator(monkeypatch):
    """helmholtz_x/eigensolvers.py:278-348: every Newton step shift-inverts L(omega_k) at sigma = 0, and
    L(omega_k) becomes singular as omega_k converges.  tol 1e-6 (the golden runs stop at 1e-2)."""
    monkeypatch.setattr(HostBackend, "supports_mixed", True)
    case = cases.rijke3d()
    hops = HostOperators(case)
    D = HostFlame(case, hops)
    gold = cases.cplx(G["rijke3d_active_fpi"]["omegas"][-1])
    omega, p = eigensolvers.newtonSolver(hops, D, gold * (1 + 2e-3), nev=2, i=0, tol=1e-6)
    om_o, _, _ = ox.newton_solver(hops.oracle, cases.oracle_flame(case), gold * (1 + 2e-3), nev=2, i=0, tol=1e-6)
    assert abs(omega - om_o) < 1e-8 * abs(om_o), (omega, om_o)
    assert abs(omega - gold) < 1e-5 * abs(gold)
