"""GPU parity tests: every product call goes through libhx_b200.so (the C-ABI) and is
compared with the CPU oracle on the same inputs, and with the reference's golden
logs where one exists.  Bars: bit-exact for integer/index work (CSR pattern, dof
maps, point ownership); relative 1e-12 for assembled FP64 values; relative 1e-8 for
eigenvalues (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import hx_oracle as ox
from tests import cases
from tests.gpu_helpers import gpu_flame, gpu_mesh, gpu_operators

pytestmark = pytest.mark.gpu

G = cases.golden_values()
VAL_RTOL = 1e-12     # assembled values, relative to the largest entry
EIG_RTOL = 1e-8      # eigenvalue parity (north star)


def be():
    from helmholtz_x_b200.fem import default_backend
    return default_backend()


def relmax(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def on_pattern(space, Mx):
    return ox._on_pattern(space, Mx)


CASES = {"rijke3d": cases.rijke3d, "prf": cases.prf_rijke3d, "rijkeffd": cases.rijkeffd, "annulus": cases.annulus}


def degree_case(name, degree):
    c = CASES[name]()
    c["degree"] = degree
    return c


# ---------------------------------------------------------------------------- K4
@pytest.mark.parametrize("name,degree", [("rijke3d", 1), ("rijke3d", 2), ("annulus", 1)])
def test_csr_pattern_and_dofmap_bit_exact(name, degree):
    from helmholtz_x_b200 import fem
    case = degree_case(name, degree)
    V = fem.functionspace(gpu_mesh(case), ("Lagrange", degree))
    sp_ = ox.function_space(case.mesh, degree)
    assert V.n == sp_.n
    assert np.array_equal(V.cell_dofs.cpu().numpy(), sp_.cell_dofs)
    assert np.array_equal(V.facet_dofs.cpu().numpy(), sp_.facet_dofs)
    indptr, indices = V.pattern()
    ip, ix = ox.csr_pattern(sp_)
    assert np.array_equal(indptr.cpu().numpy(), ip)
    assert np.array_equal(indices.cpu().numpy(), ix)


def test_pattern_is_run_to_run_deterministic():
    from helmholtz_x_b200 import fem
    case = cases.rijke3d()
    m = gpu_mesh(case)
    a = fem.FunctionSpace(m, 2).pattern()
    b = fem.FunctionSpace(m, 2).pattern()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


# ---------------------------------------------------------------------------- K1-K3
@pytest.mark.parametrize("name,degree", [("rijke3d", 1), ("prf", 1), ("rijkeffd", 1), ("annulus", 1), ("prf", 2),
                                         ("annulus", 2)])
def test_assembled_A_B_C_match_oracle(name, degree):
    case = degree_case(name, degree)
    mats = gpu_operators(case)
    ops = cases.oracle_operators(case)
    sp_ = ops.space
    a = mats.ops.base["A"].cpu().numpy()
    c = mats.ops.base["C"].cpu().numpy()
    assert relmax(a, on_pattern(sp_, ops.A).real) < VAL_RTOL
    assert relmax(c, on_pattern(sp_, ops.C).real) < VAL_RTOL
    if ops.B is None:
        assert mats.B is None
    else:
        b = mats.ops.base["B"].cpu().numpy()
        assert relmax(b, on_pattern(sp_, ops.B)) < VAL_RTOL
        ip, ix, v = mats.B_adj.getValuesCSR()
        assert relmax(v, on_pattern(sp_, ops.B_adj)) < VAL_RTOL


def test_assembly_is_bitwise_reproducible():
    from helmholtz_x_b200 import fem
    case = degree_case("prf", 2)
    V = fem.functionspace(gpu_mesh(case), ("Lagrange", 2))
    a1, c1 = fem.assemble_AC(V, case.c)
    a2, c2 = fem.assemble_AC(V, case.c)
    assert torch.equal(a1, a2) and torch.equal(c1, c2)


@pytest.mark.parametrize("name", ["rijke3d", "annulus"])
def test_device_colouring_is_valid_and_deterministic(name):
    """hx_color_cells (Jones-Plassmann rounds on the device): cells of one colour share no vertex, the
    result depends on the mesh alone (two fresh runs are identical), the colour count stays near the host
    greedy first-fit (hx_color_cells_h, the checker), facets likewise."""
    import ctypes as C
    from helmholtz_x_b200 import _lib, fem
    case = CASES[name]()
    m = case.mesh
    runs = []
    for _ in range(2):
        mesh = fem.Mesh(m.x, m.cells, m.cell_tags, m.facets, m.facet_tags)
        nc, ptr, order = mesh.cell_colors()
        runs.append((nc, ptr.copy(), order.cpu().numpy()))
    assert runs[0][0] == runs[1][0] and np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][2], runs[1][2])
    nc, ptr, order = runs[0]
    cells = np.asarray(m.cells)
    assert np.array_equal(np.sort(order), np.arange(len(cells)))
    for c in range(nc):
        members = order[ptr[c]:ptr[c + 1]]
        assert np.all(np.diff(members) > 0)                       # ascending cell index inside a colour
        v = cells[members].ravel()
        assert len(np.unique(v)) == len(v), f"colour {c}: two cells share a vertex"
    ent = np.ascontiguousarray(cells, dtype=np.int32)
    col_h = np.empty(len(ent), np.int32)
    nc_h = _lib.call("hx_color_cells_h", len(ent), 4, ent.ctypes.data_as(C.c_void_p), m.x.shape[0], col_h.ctypes.data_as(C.c_void_p))
    assert nc <= nc_h + 16, (nc, nc_h)
    tag = int(np.unique(m.facet_tags)[0])
    ncf, ptrf, orderf = mesh.facet_colors(tag)
    fac = np.asarray(m.facets)
    orderf = orderf.cpu().numpy()
    assert np.array_equal(np.sort(orderf), np.flatnonzero(m.facet_tags == tag))
    for c in range(ncf):
        v = fac[orderf[ptrf[c]:ptrf[c + 1]]].ravel()
        assert len(np.unique(v)) == len(v)


def test_dirichlet_rows_cols_and_unit_diagonal():
    case = cases.rijke3d()
    case["bcs"] = {1: {"Dirichlet"}, 2: {"Neumann"}, 3: {"Neumann"}}
    mats = gpu_operators(case)
    ops = cases.oracle_operators(case)
    assert relmax(mats.ops.base["A"].cpu().numpy(), on_pattern(ops.space, ops.A).real) < VAL_RTOL
    assert relmax(mats.ops.base["C"].cpu().numpy(), on_pattern(ops.space, ops.C).real) < VAL_RTOL


def test_choked_boundaries_match_oracle():
    """ChokedInlet/ChokedOutlet with T-dependent gamma (acoustic_matrices.py:75-97) on the Rijke mesh."""
    case = cases.rijkeffd()
    case["bcs"] = {1: {"Neumann"}, 2: {"ChokedOutlet": 0.05}, 3: {"ChokedInlet": 0.02}}
    mats = gpu_operators(case)
    ops = cases.oracle_operators(case)
    assert relmax(mats.ops.base["B"].cpu().numpy(), on_pattern(ops.space, ops.B)) < 1e-11


# ---------------------------------------------------------------------------- K5/K6
@pytest.mark.parametrize("name,degree", [("rijke3d", 1), ("rijkeffd", 1), ("rijke3d", 2)])
def test_distributed_flame_vectors(name, degree):
    case = degree_case(name, degree)
    D = gpu_flame(case)
    D.assemble_submatrices()
    fl = cases.oracle_flame(case)
    (li, lv), = D.submatrices.lists[0]
    (ri, rv), = D.submatrices.lists[1]
    ol, orr = fl.left[:, 0], fl.right[:, 0]
    # thresholding near tol may flip single entries; compare dense vectors
    dl = np.zeros_like(ol); dl[li] = lv
    dr = np.zeros_like(orr); dr[ri] = rv
    assert np.abs(dl - ol).max() < 1e-12 * np.abs(ol).max() + 2e-5 * (np.count_nonzero(dl) != np.count_nonzero(ol))
    assert np.abs(dr - orr).max() < 1e-12 * np.abs(orr).max() + 2e-5 * (np.count_nonzero(dr) != np.count_nonzero(orr))
    if degree == 1 and name == "rijke3d":
        assert (len(li), len(ri)) == (558, 569)        # SURVEY K5 (reference block 558 x 569)


@pytest.mark.parametrize("degree", [1, 2])
def test_pointwise_flame_vectors_and_point_ownership(degree):
    case = degree_case("annulus", degree)
    D = gpu_flame(case)
    D.assemble_submatrices()
    fl = cases.oracle_flame(case)
    lefts, rights = D.submatrices.lists
    assert len(lefts) == 16
    for f in range(16):
        dl = np.zeros(fl.left.shape[0]); dl[lefts[f][0]] = lefts[f][1]
        dr = np.zeros(fl.left.shape[0]); dr[rights[f][0]] = rights[f][1]
        assert np.abs(dl - fl.left[:, f]).max() < 1e-11 * np.abs(fl.left[:, f]).max()
        assert np.abs(dr - fl.right[:, f]).max() < 1e-11 * np.abs(fl.right[:, f]).max()
    if degree == 1:
        assert D.submatrices.dense_block_nnz() == fl.dense_block_nnz()


# ---------------------------------------------------------------------------- K7/K8
@pytest.mark.parametrize("lanes", [2, 4, 8, 16, 32])
def test_spmv_complex_csr_all_lane_widths(lanes):
    case = cases.annulus()
    mats = gpu_operators(case)
    s = case.target
    P = (mats.A + s * mats.B + s ** 2 * mats.C)
    csr = P.csr()
    rng = np.random.default_rng(0)
    x = rng.standard_normal(csr.n_cols) + 1j * rng.standard_normal(csr.n_cols)
    xd = be().asarray(x, dtype=torch.complex128)
    yd = be().zeros(csr.n_rows)
    be().spmv(csr, xd, yd, lanes=lanes)
    ref = ox.spmv(csr.indptr.cpu().numpy(), csr.indices.cpu().numpy(), csr.values.cpu().numpy(), x)
    assert relmax(yd.cpu().numpy(), ref) < 1e-13
    # alpha/beta form, real-valued matrix path
    y0 = be().asarray(rng.standard_normal(csr.n_rows) + 0j, dtype=torch.complex128)
    Cr = mats.ops.space.matrix(mats.ops.base["C"])
    be().spmv(Cr, xd, yd, alpha=2 - 1j, beta=0.5j, y0=y0, lanes=lanes)
    refc = (2 - 1j) * ox.spmv(Cr.indptr.cpu().numpy(), Cr.indices.cpu().numpy(), Cr.values.cpu().numpy(), x) + 0.5j * y0.cpu().numpy()
    assert relmax(yd.cpu().numpy(), refc) < 1e-13


def test_spmv_edge_cases_empty_rows_and_ragged():
    from helmholtz_x_b200.backend import CsrMatrix
    b = be()
    indptr = np.array([0, 0, 3, 3, 4, 4 + 70], np.int32)       # empty rows, a long ragged row
    rng = np.random.default_rng(3)
    indices = np.concatenate([[0, 2, 4], [1], np.sort(rng.choice(100, 70, replace=False))]).astype(np.int32)
    vals = rng.standard_normal(74) + 1j * rng.standard_normal(74)
    x = rng.standard_normal(100) + 1j * rng.standard_normal(100)
    M = CsrMatrix(5, 100, b.asarray(indptr), b.asarray(indices), b.asarray(vals, dtype=torch.complex128))
    for lanes in (2, 8, 32):
        y = b.zeros(5)
        b.spmv(M, b.asarray(x, dtype=torch.complex128), y, lanes=lanes)
        assert relmax(y.cpu().numpy(), ox.spmv(indptr, indices, vals, x)) < 1e-14
    # n = 0 is a no-op
    M0 = CsrMatrix(0, 100, b.asarray(np.zeros(1, np.int32)), b.asarray(np.zeros(0, np.int32)), b.zeros(0))
    b.spmv(M0, b.asarray(x, dtype=torch.complex128), b.zeros(1))


def test_spmv_sell_matches_csr():
    from helmholtz_x_b200.sell import SellMatrix
    case = cases.annulus()
    mats = gpu_operators(case)
    s = case.target
    csr = (mats.A + s * mats.B + s ** 2 * mats.C).csr()
    sell = SellMatrix.from_csr(be(), csr)
    rng = np.random.default_rng(1)
    x = be().asarray(rng.standard_normal(csr.n_cols) + 1j * rng.standard_normal(csr.n_cols), dtype=torch.complex128)
    y1, y2 = be().zeros(csr.n_rows), be().zeros(csr.n_rows)
    be().spmv(csr, x, y1)
    for variant in range(6):
        y2.zero_()
        sell.spmv(x, y2, variant=variant)
        assert relmax(y2.cpu().numpy(), y1.cpu().numpy()) < 1e-13
    assert sell.padding_ratio < 1.25
    # alpha/beta epilogue and the Jacobi sweep on the SELL matrix against the CSR kernels
    y0 = be().asarray(rng.standard_normal(csr.n_rows) + 0j, dtype=torch.complex128)
    be().spmv(csr, x, y1, alpha=-1.0, beta=1.0, y0=y0)
    be().spmv(sell, x, y2, alpha=-1.0, beta=1.0, y0=y0)
    assert relmax(y2.cpu().numpy(), y1.cpu().numpy()) < 1e-13
    dinv = be().zeros(csr.n_rows)
    be().diag_inv(csr, dinv)
    be().jacobi_sweep(csr, dinv, y0, x, y1, 0.6)
    be().jacobi_sweep(sell, dinv, y0, x, y2, 0.6)
    assert relmax(y2.cpu().numpy(), y1.cpu().numpy()) < 1e-13
    be().jacobi_sweep(sell, dinv, y0, None, y2, 0.6)
    assert relmax(y2.cpu().numpy(), (0.6 * dinv * y0).cpu().numpy()) < 1e-14


@pytest.mark.parametrize("shape", [(5000, 300), (300, 5000), (4097, 4097)])
def test_spmv_sell_real_rectangular_transfer_operators(shape):
    """hx_spmv_sell_sc: float32 matrix (the prolongation / restriction operators of the multigrid cycle,
    tall and wide) times complex64 vector, with the x += P xc epilogue, against SciPy.  Padding entries
    must stay inside the column range (a tall matrix has fewer columns than rows)."""
    import scipy.sparse as sp
    from helmholtz_x_b200.backend import CsrMatrix
    from helmholtz_x_b200.sell import SellMatrix, SellPattern
    b = be()
    m, n = shape
    rng = np.random.default_rng(m + n)
    M = sp.random(m, n, density=min(6.0 / n, 1.0), random_state=rng, format="csr", dtype=np.float32)
    M.sort_indices()
    csr = CsrMatrix(m, n, b.asarray(M.indptr, dtype=torch.int32), b.asarray(M.indices, dtype=torch.int32),
                    b.asarray(M.data, dtype=torch.float32))
    p = SellPattern(b, csr.indptr, csr.indices, m, n)
    assert int(p.cols.min()) >= 0 and int(p.cols.max()) < n
    S = SellMatrix(p, p.values_from_csr(csr.values))
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    y0 = (rng.standard_normal(m) + 1j * rng.standard_normal(m)).astype(np.complex64)
    xd, yd = b.asarray(x, dtype=torch.complex64), b.asarray(y0, dtype=torch.complex64)
    out = torch.zeros(m, dtype=torch.complex64, device=b.device)
    b.spmv(S, xd, out)
    ref = M.astype(np.float64) @ x.astype(np.complex128)
    assert relmax(out.cpu().numpy(), ref) < 2e-6
    b.spmv(S, xd, yd, alpha=1.0, beta=1.0, y0=yd)                 # in place: y += M x
    assert relmax(yd.cpu().numpy(), y0 + ref) < 2e-6
    out2 = torch.zeros(m, dtype=torch.complex64, device=b.device)
    b.spmv(csr, xd, out2)                                         # the CSR kernel it replaces
    assert relmax(out.cpu().numpy(), out2.cpu().numpy()) < 2e-6


def test_fused_operator_apply_with_flame_term():
    case = cases.rijke3d()
    mats = gpu_operators(case)
    D = gpu_flame(case)
    D.assemble_submatrices()
    om = 1248.5 + 3.4j
    D.assemble_matrix(om)
    Lm = mats.A + om ** 2 * mats.C - D.matrix
    ops = cases.oracle_operators(case)
    fl = cases.oracle_flame(case)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(ops.A.shape[0]) + 1j * rng.standard_normal(ops.A.shape[0])
    y = be().zeros(len(x))
    Lm.apply(be().asarray(x, dtype=torch.complex128), y)
    ref = ox.fused_apply(ops, fl, om, fl.FTF(om), x)
    assert relmax(y.cpu().numpy(), ref) < 1e-12


# ---------------------------------------------------------------------------- K10/K11
def test_basis_kernels_match_numpy_and_are_deterministic():
    b = be()
    rng = np.random.default_rng(7)
    n, m = 100003, 19
    Vh = rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
    wh = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    V = b.asarray(Vh, dtype=torch.complex128)
    w = b.asarray(wh, dtype=torch.complex128)
    out = b.zeros(m + 2)
    for k in (1, 7, 8, 9, 19):
        b.multi_dot(V, k, w, out)
        assert relmax(out[:k].cpu().numpy(), Vh[:k].conj() @ wh) < 1e-13
        o1 = out[:k].clone()
        b.multi_dot(V, k, w, out)
        assert torch.equal(o1, out[:k]), "reduction is not run-to-run deterministic"
        b.multi_dot(V, k, w, out, conj=False)
        assert relmax(out[:k].cpu().numpy(), Vh[:k] @ wh) < 1e-13
    h = rng.standard_normal(m) + 1j * rng.standard_normal(m)
    hd = b.asarray(h, dtype=torch.complex128)
    hacc = b.zeros(m + 2)
    nr = torch.view_as_real(hacc)[m + 1]
    w2 = w.clone()
    b.multi_axpy(V, m, hd, w2, hacc=hacc, nrm2=nr)
    ref = wh - h @ Vh
    assert relmax(w2.cpu().numpy(), ref) < 1e-13
    assert relmax(hacc[:m].cpu().numpy(), h) < 1e-15
    assert abs(float(nr[0]) - np.vdot(ref, ref).real) < 1e-12 * np.vdot(ref, ref).real
    o = b.zeros(n)
    b.scale_copy(w2, o, nrm2=nr)
    assert abs(np.linalg.norm(o.cpu().numpy()) - 1) < 1e-13
    b.axpby(2j, w, -1.0, o)
    Q = rng.standard_normal((11, m)) + 1j * rng.standard_normal((11, m))
    Vout = b.zeros(11, n)
    b.basis_rotate(V, m, b.asarray(Q, dtype=torch.complex128), 11, Vout)
    assert relmax(Vout.cpu().numpy(), Q @ Vh) < 1e-13


@pytest.mark.parametrize("n,m,kout", [(100003, 19, 11), (4099, 64, 33), (7, 3, 1), (50000, 20, 4)])
def test_basis_rotate_on_fp64_tensor_cores_matches_the_fma_kernel(n, m, kout):
    """hx_basis_rotate_dmma (mma.sync.m8n8k4.f64, the restart rotation V <- V Q of Krylov-Schur; SLEPc
    BVMultInPlace behind helmholtz_x/eigensolvers.py:62,113) against NumPy and the CUDA-core kernel."""
    b = be()
    rng = np.random.default_rng(n + m)
    Vh = rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
    Q = rng.standard_normal((kout, m)) + 1j * rng.standard_normal((kout, m))
    V = b.asarray(Vh, dtype=torch.complex128)
    Qd = b.asarray(Q, dtype=torch.complex128)
    ref = Q @ Vh
    out_t = torch.full((kout, n), float("nan"), dtype=torch.complex128, device=b.device)
    b.basis_rotate(V, m, Qd, kout, out_t, tensor_cores=True)
    assert relmax(out_t.cpu().numpy(), ref) < 1e-13
    out_f = b.zeros(kout, n)
    b.basis_rotate(V, m, Qd, kout, out_f, tensor_cores=False)
    assert relmax(out_t.cpu().numpy(), out_f.cpu().numpy()) < 1e-13
    again = b.zeros(kout, n)
    b.basis_rotate(V, m, Qd, kout, again, tensor_cores=True)
    assert torch.equal(again, out_t)


def test_dense_inverse_and_gemv():
    b = be()
    rng = np.random.default_rng(11)
    for n in (1, 5, 64, 300):
        A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        A[0, 0] = 0.0 if n > 1 else A[0, 0]          # force a pivot exchange
        Acm = b.asarray(np.ascontiguousarray(A.T), dtype=torch.complex128)
        info = b.dense_inverse(Acm)
        assert int(info[0]) == 0
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        y = b.zeros(n)
        b.dense_gemv(Acm, b.asarray(x, dtype=torch.complex128), y)
        assert relmax(y.cpu().numpy(), np.linalg.solve(A, x)) < 1e-9


def test_spgemm_kernels_match_scipy_and_are_deterministic():
    """hx_spgemm_symbolic / hx_spgemm_numeric (multigrid set-up) against SciPy: pattern bit-exact."""
    import scipy.sparse as sp
    from helmholtz_x_b200 import spgemm
    from helmholtz_x_b200.backend import CsrMatrix
    b = be()
    rng = np.random.default_rng(5)

    def dev(Ms):
        Ms = Ms.tocsr(); Ms.sort_indices()
        return CsrMatrix(Ms.shape[0], Ms.shape[1], b.asarray(Ms.indptr.astype(np.int32)), b.asarray(Ms.indices.astype(np.int32)),
                         b.asarray(Ms.data.astype(np.float64)))
    case = cases.annulus()
    ops = cases.oracle_operators(case)
    A = ops.A.real.tocsr()
    n = A.shape[0]
    agg = np.arange(n) // 7
    T = sp.csr_matrix((rng.standard_normal(n), (np.arange(n), agg)), shape=(n, agg.max() + 1))
    P = (A @ T).tocsr()                                       # a realistic, wider right factor
    for X, Y in ((A, T), (A, P), (P.T.tocsr(), (A @ P).tocsr())):
        C1 = spgemm.multiply(b, dev(X), dev(Y))
        C2 = spgemm.multiply(b, dev(X), dev(Y))
        assert torch.equal(C1.values, C2.values)
        ref = (abs(X) @ abs(Y)).tocsr(); ref.sort_indices()
        assert np.array_equal(C1.indptr.cpu().numpy(), ref.indptr)
        assert np.array_equal(C1.indices.cpu().numpy(), ref.indices)
        val = (X @ Y).tocsr()
        got = sp.csr_matrix((C1.values.cpu().numpy(), C1.indices.cpu().numpy(), C1.indptr.cpu().numpy()), shape=val.shape)
        assert abs(got - val).max() < 1e-12 * abs(val).max()
    Rt = spgemm.transpose(dev(P))
    assert abs(Rt.to_scipy() - P.T).max() == 0.0
    # empty rows and an overflowing row (> 2048 distinct columns) are handled
    E = sp.csr_matrix((3, n))
    assert spgemm.multiply(b, dev(E), dev(A)).nnz == 0
    wide = sp.csr_matrix((np.ones(3000), (np.zeros(3000, int), np.arange(3000))), shape=(1, n))
    D = sp.identity(n, format="csr")
    with pytest.raises(spgemm.Overflow):
        spgemm.multiply(b, dev(wide), dev(D))
    # a product with rows of 513..1500 distinct columns: the small size class overflows, the large one takes over
    m = 40
    rows = np.repeat(np.arange(m), 1500)
    cols = np.concatenate([rng.choice(n, 1500, replace=False) for _ in range(m)])
    W = sp.csr_matrix((rng.standard_normal(m * 1500), (rows, cols)), shape=(m, n))
    Cw = spgemm.multiply(b, dev(W), dev(D))
    ref = W.tocsr(); ref.sort_indices()
    assert np.array_equal(Cw.indices.cpu().numpy(), ref.indices)
    assert abs(sp.csr_matrix((Cw.values.cpu().numpy(), Cw.indices.cpu().numpy(), Cw.indptr.cpu().numpy()), shape=ref.shape) - ref).max() < 1e-14


def test_ilu0_level_scheduled_matches_host_ilu0():
    from helmholtz_x_b200.ilu import ILU0
    case = cases.rijke3d()
    mats = gpu_operators(case)
    csr = (mats.A + (400 * np.pi) ** 2 * mats.C).csr()
    ilu = ILU0(be(), csr)
    rng = np.random.default_rng(2)
    bvec = rng.standard_normal(csr.n_rows) + 1j * rng.standard_normal(csr.n_rows)
    x = be().zeros(csr.n_rows)
    ilu.solve(be().asarray(bvec, dtype=torch.complex128), x)
    # host ILU(0) on the same pattern
    ip, ix, v = csr.indptr.cpu().numpy(), csr.indices.cpu().numpy(), csr.values.cpu().numpy().copy()
    n = csr.n_rows
    diag = np.array([ip[i] + np.searchsorted(ix[ip[i]:ip[i + 1]], i) for i in range(n)])
    for i in range(n):
        for kk in range(ip[i], diag[i]):
            k = ix[kk]
            v[kk] /= v[diag[k]]
            cols_k = ix[diag[k] + 1:ip[k + 1]]
            pos = ip[i] + np.searchsorted(ix[ip[i]:ip[i + 1]], cols_k)
            ok = (pos < ip[i + 1]) & (ix[np.minimum(pos, ip[i + 1] - 1)] == cols_k)
            v[pos[ok]] -= v[kk] * v[diag[k] + 1:ip[k + 1]][ok]
    xs = bvec.copy()
    for i in range(n):
        xs[i] -= v[ip[i]:diag[i]] @ xs[ix[ip[i]:diag[i]]]
    for i in range(n - 1, -1, -1):
        xs[i] = (xs[i] - v[diag[i] + 1:ip[i + 1]] @ xs[ix[diag[i] + 1:ip[i + 1]]]) / v[diag[i]]
    assert relmax(x.cpu().numpy(), xs) < 1e-10


# ---------------------------------------------------------------------------- K9 + a7-a10
def test_amg_vcycle_matches_host_double_and_gmres_converges():
    from helmholtz_x_b200.operators import ShiftedSolver
    case = cases.annulus()
    mats = gpu_operators(case)
    s = case.target
    solver = ShiftedSolver(mats.ops, {"A": 1.0, "B": s, "C": s ** 2}, rtol=1e-11)
    rng = np.random.default_rng(4)
    n = mats.ops.n
    bvec = be().asarray(rng.standard_normal(n) + 1j * rng.standard_normal(n), dtype=torch.complex128)
    x = be().zeros(n)
    solver.solve(bvec, x)
    P = solver.P.to_scipy()
    res = np.linalg.norm(P @ x.cpu().numpy() - bvec.cpu().numpy()) / np.linalg.norm(bvec.cpu().numpy())
    assert res < 1e-10
    assert mats.ops.stats["inner_iterations"] <= 60, mats.ops.stats


def test_amg_fused_tail_matches_the_kernel_by_kernel_cycle(monkeypatch):
    """hx_amg_tail (last smoothed level + dense coarsest solve as one persistent kernel with grid barriers)
    against the same cycle launched kernel by kernel: complex64 round-off apart, V- and W-cycle, repeated
    applications (the barrier state persists across launches)."""
    from helmholtz_x_b200.amg import AMG
    case = cases.annulus()
    mats = gpu_operators(case)
    ops = mats.ops
    pat = ops.space.matrix
    s = case.target
    outs = {}
    for w_from in (None, 1):
        for tail in ("0", "1"):
            monkeypatch.setenv("HX_AMG_TAIL", tail)           # off by default (slower in the graph-replayed cycle)
            mg = AMG(be(), pat(ops.base["A"]), pat(ops.base["C"]), pat(ops.base["B"]), ops.space.dof_coords, w_from=w_from or "off")
            mg.set_shift(1.0, s, s ** 2)
            assert (mg._tail is not None) == (tail == "1") and len(mg.levels) >= 3
            rng = np.random.default_rng(5)
            v = be().asarray(rng.standard_normal(ops.n) + 1j * rng.standard_normal(ops.n), dtype=torch.complex128)
            y = be().zeros(ops.n)
            for _ in range(3):
                mg.apply(v, y)
            outs[(w_from, tail)] = y.cpu().numpy().copy()
        assert relmax(outs[(w_from, "1")], outs[(w_from, "0")]) < 5e-5, w_from


@pytest.mark.parametrize("precision,sell_min_rows,graph", [("single", 1000, True), ("double", 1000, True),
                                                           ("single", 10 ** 9, True), ("single", 1000, False)])
def test_amg_precision_modes_reach_full_accuracy(precision, sell_min_rows, graph, monkeypatch):
    """The complex64 V-cycle is only a preconditioner: the complex128 GMRES must still reach 1e-11 --
    with the levels in SELL-32 (what million-row levels use) or CSR, replayed from a CUDA graph or
    launched kernel by kernel."""
    from helmholtz_x_b200.operators import ShiftedSolver
    monkeypatch.setenv("HX_AMG_GRAPH", "1" if graph else "0")
    case = cases.annulus()
    mats = gpu_operators(case)
    mats.ops.amg_options = {"precision": precision, "sell_min_rows": sell_min_rows}
    s = case.target
    solver = ShiftedSolver(mats.ops, {"A": 1.0, "B": s, "C": s ** 2}, rtol=1e-11)
    assert solver.mg.single == (precision == "single")
    assert solver.mg.use_graph == graph
    assert getattr(solver.mg.levels[0].Mop, "is_sell", False) == (sell_min_rows == 1000)
    rng = np.random.default_rng(8)
    n = mats.ops.n
    bvec = be().asarray(rng.standard_normal(n) + 1j * rng.standard_normal(n), dtype=torch.complex128)
    x = be().zeros(n)
    solver.solve(bvec, x)
    P = solver.P.to_scipy()
    res = np.linalg.norm(P @ x.cpu().numpy() - bvec.cpu().numpy()) / np.linalg.norm(bvec.cpu().numpy())
    assert res < 5e-11, res
    assert mats.ops.stats["inner_iterations"] <= 70, mats.ops.stats


def test_passive_eps_matches_golden():
    """.../RijkeTube3D/Results/Passive/passive.log:30-33"""
    from helmholtz_x_b200.eigensolvers import eps_solver
    case = cases.rijke3d()
    mats = gpu_operators(case, passive=True)
    E = eps_solver(mats.A, mats.C, case.target, nev=2)
    lam = np.array([E.getEigenvalue(i) for i in range(2)])
    gold = G["rijke3d_passive_eps"]["lambdas"][0]
    assert min(abs(lam - gold)) / gold < EIG_RTOL


def _run_fpi(case, problem_type="direct", **kw):
    from helmholtz_x_b200.eigensolvers import fixed_point_iteration
    mats = gpu_operators(case)
    D = gpu_flame(case)
    D.assemble_submatrices(problem_type)
    target = case.target if problem_type == "direct" else np.conj(case.target)
    E = fixed_point_iteration(mats, D, target, nev=case.nev, i=0, tol=case.tol, problem_type=problem_type, **kw)
    return mats, D, E


def _check_history(hist, gold, rtol=EIG_RTOL, atol=0.0):
    assert len(hist) >= len(gold)
    for a, b in zip(hist[-len(gold):], gold):
        assert abs(a - b) <= rtol * abs(b) + atol, (a, b)


def test_fpi_rijke3d_config1_matches_golden_log_and_eigenvector():
    """config 1: .../RijkeTube3D/Results/Active/active.log:23-52 and Results/Active/p.h5"""
    from helmholtz_x_b200.eigenvectors import normalize_eigenvector
    from scipy.spatial import cKDTree
    case = cases.rijke3d()
    mats, D, E = _run_fpi(case)
    gold = [cases.cplx(p) for p in G["rijke3d_active_fpi"]["omegas"]]
    _check_history(E.omega_history, gold, atol=6e-9)       # log prints 8 decimals
    omega, p = normalize_eigenvector(mats.mesh, E, 0, degree=1, which='right', matrices=mats)
    gp = np.load(cases.GOLDEN_DIR + "/rijke3d_active_p.npz")
    _, idx = cKDTree(case.mesh.x).query(gp["geometry"])
    pm, pg = p.x.array[idx], gp["p"]
    sgn = 1 if abs(pm[0] - pg[0]) < abs(pm[0] + pg[0]) else -1
    assert np.abs(sgn * pm - pg).max() / np.abs(pg).max() < 1e-7


def test_fpi_flamedduct_choked_boundaries_match_golden():
    """.../NetworkCode/FlamedDuct/Results/Active/active.log:21-56: choked inlet / outlet, temperature
    parameter (variable gamma), half-Gaussian heat release, 33,855 dofs, PEP fixed-point iteration."""
    case = cases.flamedduct()
    _, _, E = _run_fpi(case)
    gold = [cases.cplx(p) for p in G["flamedduct_active_fpi"]["omegas"]]
    _check_history(E.omega_history, gold, atol=7.1e-9)      # log prints 8 decimals


def test_fpi_prf_pep_direct_and_adjoint_match_golden():
    case = cases.prf_rijke3d()
    _, _, E = _run_fpi(case)
    _check_history(E.omega_history, [cases.cplx(p) for p in G["prf_rijke3d_direct_fpi"]["omegas"]], atol=6e-9)
    _, _, E2 = _run_fpi(case, "adjoint")
    _check_history(E2.omega_history, [cases.cplx(p) for p in G["prf_rijke3d_adjoint_fpi"]["omegas"]], atol=6e-9)


def test_fpi_rijkeffd_config5_eigenpair_matches_golden():
    """.../RijkeFFD/Results/Original/eigenvalues.txt"""
    case = cases.rijkeffd()
    _, _, E = _run_fpi(case)
    g = cases.cplx(G["rijkeffd_eigenvalues"]["direct"])
    assert abs(E.getEigenpair(0) - g) / abs(g) < EIG_RTOL
    _, _, E2 = _run_fpi(case, "adjoint")
    g = cases.cplx(G["rijkeffd_eigenvalues"]["adjoint"])
    assert abs(E2.getEigenpair(0) - g) / abs(g) < EIG_RTOL


def test_normalize_adjoint_config5_self_check():
    """a13: eigenvectors.py:125-177; the reference prints '! Normalization Check:' = 1
    (RijkeFFD/Results/ShapeDerivatives/results.log:113)."""
    from helmholtz_x_b200.eigenvectors import normalize_adjoint, normalize_eigenvector
    from helmholtz_x_b200.petsc4py_utils import vector_matrix_vector
    case = cases.rijkeffd()
    mats, D, E = _run_fpi(case)
    omega_dir, p_dir = normalize_eigenvector(mats.mesh, E, 0, degree=1, which='right', matrices=mats, print_eigs=False)
    _, _, E2 = _run_fpi(case, "adjoint")
    omega_adj, p_adj = normalize_eigenvector(mats.mesh, E2, 0, degree=1, which='right', matrices=mats, print_eigs=False)
    D.assemble_submatrices('direct')
    p_adj_n = normalize_adjoint(omega_dir, p_dir, p_adj, mats, D)
    dL = mats.B + mats.C * (2 * omega_dir) - D.get_derivative(omega_dir)
    check = vector_matrix_vector(p_adj_n.x.petsc_vec, dL, p_dir.x.petsc_vec)
    assert abs(check - 1.0) < 1e-10
    # against the oracle's normalisation of its own eigenvectors (sign-fixed, so comparable)
    ops, fl = cases.oracle_operators(case), cases.oracle_flame(case)
    Eo, _ = ox.fixed_point_iteration(ops, fl, case.target, nev=2, i=0)
    Ea, _ = ox.fixed_point_iteration(ops, fl, case.target, nev=2, i=0, problem_type="adjoint")
    _, po = ox.normalize_eigenvector(ops, Eo, 0)
    _, pa = ox.normalize_eigenvector(ops, Ea, 0)
    pan, chk = ox.normalize_adjoint(ops, Eo.omega(0), po, pa, fl)
    assert abs(chk - 1.0) < 1e-10
    assert relmax(np.abs(p_adj_n.x.array), np.abs(pan)) < 1e-6


@pytest.mark.parametrize("degree", [1, 2])
def test_shape_derivative_boundary_integral_matches_oracle(degree):
    """(f-1) helmholtz_x/shape_derivatives.py:12-37: int (V.n) div(conj(p_adj) c^2 grad p) ds on the
    lateral wall of the RijkeFFD tube for a radial and an axial-bump displacement field."""
    from helmholtz_x_b200.shape_derivatives import boundary_shape_integral
    case = degree_case("rijkeffd", degree)
    m = case.mesh
    mesh = gpu_mesh(case)
    sp_ = ox.function_space(m, degree)
    rng = np.random.default_rng(12)
    p = rng.standard_normal(sp_.n) + 1j * rng.standard_normal(sp_.n)
    pa = rng.standard_normal(sp_.n) + 1j * rng.standard_normal(sp_.n)
    c = ox.sound_speed_variable_gamma(case.T)
    for V in (np.stack([m.x[:, 0], m.x[:, 1], 0 * m.x[:, 2]], 1),
              np.stack([m.x[:, 0], m.x[:, 1], 0 * m.x[:, 2]], 1) * np.exp(-((m.x[:, 2] - 0.5) / 0.1) ** 2)[:, None]):
        got = boundary_shape_integral(mesh, degree, 1, V, p, pa, c)
        ref = ox.shape_derivative(sp_, 1, V, p, pa, c)
        assert abs(got - ref) < 1e-11 * abs(ref), (got, ref)


def test_xdmf_reader_writer_round_trip_drives_a_solve(tmp_path):
    """(f-2) io_utils: XDMFReader on a meshio-layout XDMF/HDF5 pair, the reference's driver flow
    (XDMFReader -> c_step -> AcousticMatrices -> eps_solver -> normalize_eigenvector -> xdmf_writer),
    and the DOLFINx-layout result file read back."""
    from helmholtz_x_b200.acoustic_matrices import AcousticMatrices
    from helmholtz_x_b200.eigensolvers import eps_solver
    from helmholtz_x_b200.eigenvectors import normalize_eigenvector
    from helmholtz_x_b200.h5lite import H5File
    from helmholtz_x_b200.io_utils import XDMFReader, write_mesh_xdmf, xdmf_writer
    from helmholtz_x_b200.parameters_utils import c_step
    m = cases.mesh("rijke3d")
    stem = str(tmp_path / "mesh")
    write_mesh_xdmf(stem, m.x, m.cells, m.cell_tags, m.facets, m.facet_tags)
    geo = XDMFReader(stem)
    mesh, subdomains, facet_tags = geo.getAll()
    assert geo.getInfo() == 8530 and mesh.n_nodes == 2426
    assert np.array_equal(mesh.cells, m.cells) and np.array_equal(mesh.facet_tags, m.facet_tags)
    c_u = np.sqrt(1.4 * 1e5 / 1.22)
    c = c_step(mesh, np.array([[0.0, 0.0, 0.25]]), c_u, c_u)
    matrices = AcousticMatrices(mesh, facet_tags, {1: {'Neumann'}, 2: {'Neumann'}, 3: {'Neumann'}}, c, degree=1)
    E = eps_solver(matrices.A, matrices.C, 200 * 2 * np.pi, nev=2)
    omega, p = normalize_eigenvector(mesh, E, 0, degree=1, which='right')
    gold = np.sqrt(G["rijke3d_passive_eps"]["lambdas"][0])
    lam = [abs(np.sqrt(E.getEigenvalue(i)) - gold) / gold for i in range(2)]
    assert min(lam) < EIG_RTOL
    xdmf_writer(str(tmp_path / "p"), mesh, p)
    f = H5File(str(tmp_path / "p.h5"))
    assert np.array_equal(f["/Mesh/Grid/topology"], m.cells)
    back = f["/Function/real_f/0"][:, 0] + 1j * f["/Function/imag_f/0"][:, 0]
    assert np.array_equal(back, p.x.array)


def _bloch_setup(pairing):
    from helmholtz_x_b200.bloch_operator import Blochifier
    case = cases.bloch()
    mats = gpu_operators(case)
    space = ox.function_space(case.mesh, 1)
    numb = cases.bloch_numbering(space)
    bl = Blochifier(geometry=gpu_mesh(case), boundary_conditions=case.bcs, N=case.N, passive_matrices=mats,
                    pairing=pairing, numbering=numb)
    md, sd = ox.bloch_pairs(space, case.master, case.slave, case.N, pairing=pairing, numbering=numb)
    BN, NB = ox.bloch_maps(space.n, md, sd, case.N)
    return case, mats, bl, (md, sd, BN, NB)


@pytest.mark.parametrize("pairing", ["sorted", "geometric"])
def test_bloch_config4_reduced_operators_match_oracle(pairing):
    """bloch_operator.py:42-78,104-111: NB*M*BN for A, B, C; the pairing itself is index work (bit-exact)."""
    case, mats, bl, (md, sd, BN, NB) = _bloch_setup(pairing)
    assert np.array_equal(bl.dofs_master, md) and np.array_equal(bl.dofs_slave, sd)
    oo = cases.oracle_operators(case)
    for name, M in (("A", oo.A), ("B", oo.B), ("C", oo.C)):
        want = ox.blochify(M, BN, NB)
        got = getattr(bl, name).to_scipy()
        got.sort_indices()
        assert got.shape == want.shape
        assert abs(got - want).max() <= VAL_RTOL * abs(want).max(), name
    x = np.random.default_rng(1).standard_normal(bl.n_red) + 1j * np.random.default_rng(2).standard_normal(bl.n_red)
    xin, y = bl.remapper.createVecs()
    xin.setArray(x)
    bl.remapper.mult(xin, y)
    assert np.array_equal(y.array, BN @ x) or relmax(y.array, BN @ x) < 1e-15


def test_bloch_config4_passive_and_active_match_golden():
    """config 4, the reference's drivers bloch/passive.py and bloch/active.py through the product API:
    Results/Passive/passive.log:27-31 + p_1.h5, Results/Active/active.log:38-75 + p_1_dir.h5."""
    from helmholtz_x_b200.eigensolvers import eps_solver, fixed_point_iteration
    from helmholtz_x_b200.eigenvectors import normalize_eigenvector
    from scipy.spatial import cKDTree
    case, mats, bl, _ = _bloch_setup("sorted")

    def vs_golden(p, name):
        gp = np.load(cases.GOLDEN_DIR + "/" + name)
        d, idx = cKDTree(case.mesh.x).query(gp["geometry"])
        assert d.max() < 1e-9
        pm, pg = p.x.array[idx], gp["p"]
        sgn = 1 if abs(pm[0] - pg[0]) < abs(pm[0] + pg[0]) else -1
        return np.abs(sgn * pm - pg).max() / np.abs(pg).max()

    E = eps_solver(bl.A, bl.C, case.passive_target, nev=case.passive_nev, print_results=False)
    for k, g in enumerate(G["bloch_passive"]["omegas"]):
        omega, p = normalize_eigenvector(mats.mesh, E, k, BlochRemapper=bl.remapper)
        assert abs(omega - g) <= EIG_RTOL * g + 6e-7, (k, omega, g)          # log prints 6 decimals
        if k == 0:
            assert p.x.array.shape == (case.mesh.n_nodes,)
            assert vs_golden(p, "bloch_passive1_p.npz") < 1e-6

    D = gpu_flame(case, bloch_object=bl)
    D.assemble_submatrices('direct')
    D.blochify()
    assert D.submatrices.n == bl.n_red
    E = fixed_point_iteration(bl, D, case.target, nev=case.nev, i=0, tol=case.tol)
    gold = [cases.cplx(p) for p in G["bloch_active_fpi"]["omegas"]]
    assert len(E.omega_history) == len(gold) + 1
    _check_history(E.omega_history, gold, rtol=0, atol=6e-4)                 # log prints 3 decimals
    omega, p = normalize_eigenvector(mats.mesh, E, 0, degree=1, BlochRemapper=bl.remapper)
    gf = cases.cplx(G["bloch_active_fpi"]["final"])
    assert abs(omega - gf) <= EIG_RTOL * abs(gf) + 1e-6                      # printed with 6 decimals
    assert vs_golden(p, "bloch_active1_p.npz") < 1e-6


def test_fpi_annulus_config3_matches_golden():
    """.../fullAnnulus/Results/Active/FPI/active.log:43-92 and eigenvalues_dir.txt"""
    case = cases.annulus()
    _, _, E = _run_fpi(case)
    gold = [cases.cplx(p) for p in G["annulus_fpi_direct"]["omegas"]]
    _check_history(E.omega_history, gold, rtol=0, atol=6e-4)        # log prints 3 decimals
    g1 = cases.cplx(G["annulus_fpi_eigenvalues_dir"]["direct_1"])
    g2 = cases.cplx(G["annulus_fpi_eigenvalues_dir"]["direct_2"])
    assert abs(E.getEigenpair(0) - g1) / abs(g1) < EIG_RTOL
    assert abs(E.getEigenpair(1) - g2) / abs(g2) < EIG_RTOL


def test_newton_annulus_config3_matches_golden():
    """.../fullAnnulus/Results/Active/NewtonSolver/{active.log:41-149,eigenvalues.txt}, i=0"""
    from helmholtz_x_b200.eigensolvers import newtonSolver
    case = cases.annulus()
    mats = gpu_operators(case)
    D = gpu_flame(case)
    D.assemble_submatrices('direct')
    omega, p = newtonSolver(mats, D, case.newton_init, i=0, nev=case.newton_nev, tol=case.newton_tol)
    g = cases.cplx(G["annulus_newton_eigenvalues"]["direct_1"])
    assert abs(omega - g) / abs(g) < EIG_RTOL


def test_p2_fpi_matches_oracle():
    """P2 has no usable reference golden ('parity unpinned'); the oracle is the target."""
    case = degree_case("prf", 2)
    _, _, E = _run_fpi(case)
    ops = cases.oracle_operators(case)
    fl = cases.oracle_flame(case)
    Eo, hist = ox.fixed_point_iteration(ops, fl, case.target, nev=case.nev, i=0, tol=case.tol)
    assert abs(E.getEigenpair(0) - Eo.omega(0)) / abs(Eo.omega(0)) < EIG_RTOL


# ---------------------------------------------------------------------------- size-independent properties
def test_large_synthetic_annulus_properties():
    """Full-size properties where the oracle is too slow: linearity of the SpMV, A 1 = 0
    (constants are in the kernel of the stiffness form), sum(C) = domain volume."""
    from helmholtz_x_b200 import fem, synthetic
    g = synthetic.annulus_grid(24, 384, 96, device="cuda")
    mesh = fem.Mesh(g["x"], g["cells"], g["cell_tags"], g["facets"], g["facet_tags"])
    V = fem.functionspace(mesh, ("Lagrange", 1))
    c = synthetic.annulus_sound_speed(mesh.x, mesh.cells)
    a, cv = fem.assemble_AC(V, c)
    A, C = V.matrix(a), V.matrix(cv)
    b = be()
    one = b.zeros(V.n) + 1.0
    y = b.zeros(V.n)
    b.spmv(A, one, y)
    assert float(y.abs().max()) < 1e-6 * float(a.abs().max())
    b.spmv(C, one, y)
    vol = float(mesh.volumes().sum())
    assert abs(float(y.real.sum()) - vol) < 1e-12 * vol * V.n ** 0.5
    x1 = torch.randn(V.n, dtype=torch.float64, device=b.device).to(torch.complex128)
    x2 = torch.randn(V.n, dtype=torch.float64, device=b.device).to(torch.complex128) * 1j
    y1, y2, y3 = b.zeros(V.n), b.zeros(V.n), b.zeros(V.n)
    b.spmv(A, x1, y1); b.spmv(A, x2, y2); b.spmv(A, x1 + 2 * x2, y3)
    assert float((y3 - y1 - 2 * y2).abs().max()) < 1e-12 * float(y3.abs().max())


# ---------------------------------------------------------------------------- options not yet timed on the GPU
@pytest.mark.parametrize("options", [("jacobi",), ("norelax",), ("cgs2",), ("wcycle",), ("jacobi", "norelax", "vcycle")])
def test_solver_switches_reproduce_goldens(options, monkeypatch):
    """The defaults are Chebyshev damping + relaxed inner tolerance + one Gram-Schmidt pass (+ W-cycle on
    large meshes); every switch away from them (constant Jacobi damping, full inner tolerance, two passes,
    forced W- / V-cycle) must reproduce the Rijke3D (config 1), PRF and annulus (config 3) goldens as well."""
    from helmholtz_x_b200 import eigensolvers
    import helmholtz_x_b200.operators as O
    if "wcycle" in options:
        monkeypatch.setenv("HX_AMG_WCYCLE", "1")
    if "vcycle" in options:
        monkeypatch.setenv("HX_AMG_WCYCLE", "off")
    if "jacobi" in options:
        monkeypatch.setenv("HX_AMG_SMOOTHER", "jacobi")
    if "norelax" in options:
        monkeypatch.setattr(eigensolvers, "INNER_RELAX", False)
    if "cgs2" in options:
        monkeypatch.setattr(O, "GMRES_ORTH_PASSES", 2)
    _, _, E = _run_fpi(cases.rijke3d())
    _check_history(E.omega_history, [cases.cplx(p) for p in G["rijke3d_active_fpi"]["omegas"]], atol=6e-9)
    _, _, E = _run_fpi(cases.prf_rijke3d())
    _check_history(E.omega_history, [cases.cplx(p) for p in G["prf_rijke3d_direct_fpi"]["omegas"]], atol=6e-9)
    _, _, E = _run_fpi(cases.annulus())
    g1 = cases.cplx(G["annulus_fpi_eigenvalues_dir"]["direct_1"])
    assert abs(E.getEigenpair(0) - g1) / abs(g1) < EIG_RTOL
