"""CPU oracle for the helmholtz-x hot path  --  TEST INFRASTRUCTURE ONLY.

A NumPy/SciPy restatement of what the reference computes on the path
``AcousticMatrices -> FlameMatrix -> eps/pep shift-invert -> fixed-point / Newton``.
The reference's arithmetic lives in un-vendored third-party stacks (DOLFINx 0.9.0,
PETSc/SLEPc complex builds, MUMPS) that are not installed here, so this file
restates their *published algorithms* (P1/P2 Lagrange assembly, shift-invert
Krylov eigen-solve with an exact sparse LU) and is pinned against the reference's
committed golden logs / eigenvalue files / eigenvector .h5 files
(``tests/test_oracle_golden.py``; fixtures made by ``tests/golden/make_fixtures.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module.  The product package
(``helmholtz_x_b200``) never does: it fails loudly without its CUDA library.

Parity status: P1 paths are pinned by goldens to >= 8 digits.  P2 is
"parity unpinned" (the reference's only P2 golden was produced on a mesh that is
not in the repository) -- for P2 this oracle is the target.

Each function cites the reference file:line it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.special import roots_jacobi

# ---------------------------------------------------------------------------
# mesh
# ---------------------------------------------------------------------------


@dataclass
class Mesh:
    """Tetrahedral mesh in the *input file* node order (meshio/XDMF order)."""
    x: np.ndarray            # (n_nodes, 3) float64
    cells: np.ndarray        # (n_cells, 4) int64
    cell_tags: np.ndarray    # (n_cells,) int32
    facets: np.ndarray       # (n_facets, 3) int64  tagged boundary triangles
    facet_tags: np.ndarray   # (n_facets,) int32
    _cache: dict = field(default_factory=dict, repr=False)

    @property
    def n_nodes(self):
        return self.x.shape[0]

    @property
    def n_cells(self):
        return self.cells.shape[0]


def load_mesh_npz(path) -> Mesh:
    d = np.load(path)
    return Mesh(d["x"].astype(np.float64), d["cells"].astype(np.int64), d["cell_tags"].astype(np.int32),
                d["facets"].astype(np.int64), d["facet_tags"].astype(np.int32))


def geometry(mesh: Mesh):
    """Per-cell volume |K| and barycentric gradients G[c,a,:] = grad L_a."""
    if "geom" not in mesh._cache:
        X = mesh.x[mesh.cells]                       # (nc,4,3)
        J = X[:, 1:, :] - X[:, :1, :]                # rows = edge vectors (nc,3,3)
        det = np.linalg.det(J)
        Jinv = np.linalg.inv(J)                      # (nc,3,3): columns are grad L_1..3
        G = np.empty((mesh.n_cells, 4, 3))
        G[:, 1:, :] = np.transpose(Jinv, (0, 2, 1))
        G[:, 0, :] = -G[:, 1:, :].sum(axis=1)
        mesh._cache["geom"] = (np.abs(det) / 6.0, G)
    return mesh._cache["geom"]


# ---------------------------------------------------------------------------
# reference element: P1 / P2 Lagrange on the tetrahedron and triangle
# ---------------------------------------------------------------------------

TET_EDGES = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))
TRI_EDGES = ((0, 1), (0, 2), (1, 2))


def tabulate(degree, L, edges=TET_EDGES):
    """phi (nq,nd) and dphi/dL (nq,nd,nv) at barycentric points L (nq,nv)."""
    nq, nv = L.shape
    if degree == 1:
        phi = L.copy()
        d = np.broadcast_to(np.eye(nv), (nq, nv, nv)).copy()
        return phi, d
    nd = nv + len(edges)
    phi = np.zeros((nq, nd))
    d = np.zeros((nq, nd, nv))
    for a in range(nv):
        phi[:, a] = L[:, a] * (2 * L[:, a] - 1)
        d[:, a, a] = 4 * L[:, a] - 1
    for e, (a, b) in enumerate(edges):
        phi[:, nv + e] = 4 * L[:, a] * L[:, b]
        d[:, nv + e, a] = 4 * L[:, b]
        d[:, nv + e, b] = 4 * L[:, a]
    return phi, d


def tet_rule(n=5):
    """Collapsed Gauss-Jacobi rule, exact to degree 2n-1; weights sum to 1."""
    xa, wa = roots_jacobi(n, 2, 0)
    xb, wb = roots_jacobi(n, 1, 0)
    xc, wc = roots_jacobi(n, 0, 0)
    u = (xa + 1) / 2; v = (xb + 1) / 2; w = (xc + 1) / 2
    U, V, W = np.meshgrid(u, v, w, indexing="ij")
    WT = wa[:, None, None] * wb[None, :, None] * wc[None, None, :]
    x = U.ravel(); y = (V * (1 - U)).ravel(); z = (W * (1 - U) * (1 - V)).ravel()
    wt = WT.ravel()
    wt = wt / wt.sum()
    return np.stack([1 - x - y - z, x, y, z], axis=1), wt


def tri_rule(n=5):
    xa, wa = roots_jacobi(n, 1, 0)
    xb, wb = roots_jacobi(n, 0, 0)
    u = (xa + 1) / 2; v = (xb + 1) / 2
    U, V = np.meshgrid(u, v, indexing="ij")
    WT = wa[:, None] * wb[None, :]
    x = U.ravel(); y = (V * (1 - U)).ravel()
    wt = WT.ravel(); wt = wt / wt.sum()
    return np.stack([1 - x - y, x, y], axis=1), wt


#: 4-point degree-2 rule used by FFCx for the P1 right flame vector
TET_DEG2_A = 0.1381966011250105
TET_DEG2_B = 0.5854101966249685


def tet_rule_deg2():
    a, b = TET_DEG2_A, TET_DEG2_B
    L = np.full((4, 4), a)
    # Basix point order: (x,y,z) = (a,a,a),(b,a,a),(a,b,a),(a,a,b)
    L[0, 0] = b; L[1, 1] = b; L[2, 2] = b; L[3, 3] = b
    return L, np.full(4, 0.25)


def tet_rule_deg5_14():
    """14-point degree-5 rule (Walkington/Keast); the rule the CUDA path uses for the
    non-polynomial P2 right flame vector.  Weights sum to 1."""
    a1, w1 = 0.31088591926330060980, 0.11268792571801585080
    a2, w2 = 0.092735250310891226402, 0.073493043116361949544
    b3, w3 = 0.045503704125649649492, 0.042546020777081466438
    L, w = [], []
    for a, ww in ((a1, w1), (a2, w2)):
        for i in range(4):
            p = [a] * 4; p[i] = 1 - 3 * a
            L.append(p); w.append(ww)
    for i in range(4):
        for j in range(i + 1, 4):
            p = [0.5 - b3] * 4; p[i] = b3; p[j] = b3
            L.append(p); w.append(w3)
    return np.array(L), np.array(w)


def reference_tensors(degree):
    """Exact reference integrals (unit-volume scaling) used by the assembly.

    M[a,b]          = int phi_a phi_b
    S0[a,b,k,l]     = int dphi_a/dL_k dphi_b/dL_l
    S2[a,b,k,l,m,n] = int dphi_a/dL_k dphi_b/dL_l L_m L_n
    F1[a,m,n]       = int phi_a L_m L_n          F0[a] = int phi_a ;  F01[a,m] = int phi_a L_m
    """
    L, w = tet_rule(5)
    phi, d = tabulate(degree, L)
    M = np.einsum("q,qa,qb->ab", w, phi, phi)
    S0 = np.einsum("q,qak,qbl->abkl", w, d, d)
    S2 = np.einsum("q,qak,qbl,qm,qn->abklmn", w, d, d, L, L)
    F1 = np.einsum("q,qa,qm,qn->amn", w, phi, L, L)
    F01 = np.einsum("q,qa,qm->am", w, phi, L)
    F0 = np.einsum("q,qa->a", w, phi)
    return dict(M=M, S0=S0, S2=S2, F1=F1, F01=F01, F0=F0)


def reference_facet_tensors(degree):
    """T1[i,j,m] = int_F phi_i phi_j L_m ; T0[i,j] = int_F phi_i phi_j (unit area)."""
    L, w = tri_rule(5)
    phi, _ = tabulate(degree, L, TRI_EDGES)
    T1 = np.einsum("q,qi,qj,qm->ijm", w, phi, phi, L)
    T0 = np.einsum("q,qi,qj->ij", w, phi, phi)
    G1 = np.einsum("q,qm->m", w, L)
    return dict(T1=T1, T0=T0, G1=G1)


# ---------------------------------------------------------------------------
# dof maps and sparsity pattern
# ---------------------------------------------------------------------------


@dataclass
class Space:
    mesh: Mesh
    degree: int
    n: int
    cell_dofs: np.ndarray     # (nc, 4 | 10)
    facet_dofs: np.ndarray    # (nf, 3 | 6)
    edges: np.ndarray | None  # (n_edges,2) sorted vertex pairs (P2)
    dof_x: np.ndarray         # (n,3) dof coordinates
    _cache: dict = field(default_factory=dict, repr=False)


def _edge_keys(a, b, n):
    lo = np.minimum(a, b); hi = np.maximum(a, b)
    return lo * n + hi


def function_space(mesh: Mesh, degree: int) -> Space:
    """P1: dof = mesh node.  P2: vertex dofs first (node order) then one dof per
    edge, edges numbered by ascending (min vertex, max vertex).

    DOLFINx's own (graph-reordered) numbering is not reproducible without DOLFINx;
    eigenvalues are numbering-invariant (SURVEY hard part "bit-exact DOF maps")."""
    key = ("space", degree)
    if key in mesh._cache:
        return mesh._cache[key]
    nn = mesh.n_nodes
    if degree == 1:
        sp_ = Space(mesh, 1, nn, mesh.cells.copy(), mesh.facets.copy(), None, mesh.x)
    elif degree == 2:
        c = mesh.cells
        ek = np.stack([_edge_keys(c[:, a], c[:, b], nn) for a, b in TET_EDGES], axis=1)
        uniq, inv = np.unique(ek.ravel(), return_inverse=True)
        cell_dofs = np.concatenate([c, nn + inv.reshape(ek.shape)], axis=1)
        edges = np.stack([uniq // nn, uniq % nn], axis=1)
        f = mesh.facets
        if len(f):
            fk = np.stack([_edge_keys(f[:, a], f[:, b], nn) for a, b in TRI_EDGES], axis=1)
            pos = np.searchsorted(uniq, fk.ravel())
            assert np.all(uniq[pos] == fk.ravel())
            facet_dofs = np.concatenate([f, nn + pos.reshape(fk.shape)], axis=1)
        else:
            facet_dofs = np.zeros((0, 6), np.int64)
        dof_x = np.concatenate([mesh.x, 0.5 * (mesh.x[edges[:, 0]] + mesh.x[edges[:, 1]])])
        sp_ = Space(mesh, 2, nn + len(uniq), cell_dofs, facet_dofs, edges, dof_x)
    else:
        raise ValueError("degree must be 1 or 2")
    mesh._cache[key] = sp_
    return sp_


def csr_pattern(space: Space):
    """Sorted, duplicate-free CSR pattern from the cell dofmaps (what DOLFINx's
    SparsityPattern + MatCreateAIJ produce; acoustic_matrices.py:102)."""
    if "pattern" not in space._cache:
        cd = space.cell_dofs
        nd = cd.shape[1]
        rows = np.repeat(cd, nd, axis=1).ravel()
        cols = np.tile(cd, (1, nd)).ravel()
        key = np.unique(rows * space.n + cols)
        r = key // space.n; c = key % space.n
        indptr = np.zeros(space.n + 1, np.int64)
        np.add.at(indptr, r + 1, 1)
        indptr = np.cumsum(indptr)
        space._cache["pattern"] = (indptr.astype(np.int32), c.astype(np.int32))
    return space._cache["pattern"]


def _scatter(space, Ke, dofs=None, n=None):
    """COO -> CSR sum of element matrices Ke (ne, nd, nd)."""
    dofs = space.cell_dofs if dofs is None else dofs
    nd = dofs.shape[1]
    rows = np.repeat(dofs, nd, axis=1).ravel()
    cols = np.tile(dofs, (1, nd)).ravel()
    n = space.n
    Mx = sp.coo_matrix((Ke.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    Mx.sum_duplicates(); Mx.sort_indices()
    return Mx


def _on_pattern(space, Mx):
    """Return Mx's values laid out on the full cell pattern (zeros where absent)."""
    indptr, indices = csr_pattern(space)
    P = sp.csr_matrix((np.zeros(len(indices), Mx.dtype), indices, indptr), shape=Mx.shape)
    R = (P + Mx).tocsr(); R.sort_indices()
    # P+Mx may drop nothing (explicit zeros are kept by scipy's binop only if both absent) -> rebuild explicitly
    out = np.zeros(len(indices), Mx.dtype)
    Mc = Mx.tocoo()
    rowstart = indptr[Mc.row]
    # position by searchsorted inside each row
    key_p = np.repeat(np.arange(space.n, dtype=np.int64), np.diff(indptr)) * space.n + indices
    key_m = Mc.row.astype(np.int64) * space.n + Mc.col
    pos = np.searchsorted(key_p, key_m)
    assert np.all(key_p[pos] == key_m)
    out[pos] = Mc.data
    return out


# ---------------------------------------------------------------------------
# a1: acoustic operators A, B, C   (helmholtz_x/acoustic_matrices.py:12-125)
# ---------------------------------------------------------------------------


def _chunks(n, size=20000):
    for s in range(0, n, size):
        yield slice(s, min(n, s + size))


def assemble_A(space: Space, c, c_is_dg0=False):
    """A = -int c^2 grad(phi_k).grad(phi_j) dx  (acoustic_matrices.py:101-103).

    c is P1 nodal (c_step / sound_speed_variable_gamma, parameters_utils.py:80-153)
    or DG0 per cell (fullAnnulus/params.py:53-70)."""
    mesh = space.mesh
    vol, G = geometry(mesh)
    ref = reference_tensors(space.degree)
    nd = space.cell_dofs.shape[1]
    Ke = np.empty((mesh.n_cells, nd, nd))
    for s in _chunks(mesh.n_cells):
        GG = np.einsum("cki,cli->ckl", G[s], G[s])
        if c_is_dg0:
            Ke[s] = -(vol[s] * c[s] ** 2)[:, None, None] * np.einsum("abkl,ckl->cab", ref["S0"], GG)
        else:
            cc = c[mesh.cells[s]]
            Ke[s] = -vol[s][:, None, None] * np.einsum("abklmn,ckl,cm,cn->cab", ref["S2"], GG, cc, cc, optimize=True)
    return _scatter(space, Ke).astype(np.complex128)


def assemble_C(space: Space):
    """C = int phi_k phi_j dx  (acoustic_matrices.py:121-123)."""
    vol, _ = geometry(space.mesh)
    Ke = vol[:, None, None] * reference_tensors(space.degree)["M"][None]
    return _scatter(space, Ke).astype(np.complex128)


def facet_areas(mesh: Mesh):
    X = mesh.x[mesh.facets]
    return 0.5 * np.linalg.norm(np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), axis=1)


def facet_owner_cells(mesh: Mesh):
    """Index of the cell owning each tagged boundary facet."""
    if "fowner" not in mesh._cache:
        c = mesh.cells
        faces = np.concatenate([c[:, [1, 2, 3]], c[:, [0, 2, 3]], c[:, [0, 1, 3]], c[:, [0, 1, 2]]])
        owner = np.tile(np.arange(mesh.n_cells), 4)
        fs = np.sort(faces, axis=1)
        n = mesh.n_nodes
        key = (fs[:, 0] * n + fs[:, 1]) * n + fs[:, 2]
        order = np.argsort(key, kind="stable")
        ks = key[order]
        q = np.sort(mesh.facets, axis=1)
        qk = (q[:, 0] * n + q[:, 1]) * n + q[:, 2]
        pos = np.searchsorted(ks, qk)
        assert np.all(ks[pos] == qk), "facet without owning cell"
        mesh._cache["fowner"] = owner[order[pos]]
    return mesh._cache["fowner"]


def boundary_mean(mesh: Mesh, tag, nodal):
    """(int_tag f ds)/(int_tag ds) for a P1 field (acoustic_matrices.py:76-78,88-90)."""
    sel = mesh.facet_tags == tag
    ar = facet_areas(mesh)[sel]
    return float((ar * nodal[mesh.facets[sel]].mean(axis=1)).sum() / ar.sum())


def reflection_to_impedance(R):
    return (1 + R) / (1 - R)


def boundary_reflection(mesh, tag, bc, gamma_nodal):
    """Reflection coefficient of one boundary entry (acoustic_matrices.py:68-97)."""
    if "Robin" in bc:
        return bc["Robin"]
    if "ChokedInlet" in bc:
        g = boundary_mean(mesh, tag, gamma_nodal); M = bc["ChokedInlet"]
        t = g * M / (1 + (g - 1) * M ** 2)
        return (1 - t) / (1 + t)
    if "ChokedOutlet" in bc:
        g = boundary_mean(mesh, tag, gamma_nodal); M = bc["ChokedOutlet"]
        return (1 - 0.5 * (g - 1) * M) / (1 + 0.5 * (g - 1) * M)
    return None


def assemble_B(space: Space, c, boundary_conditions, c_is_dg0=False, gamma_nodal=None):
    """B = sum_tags int (i c / Z) phi_k phi_j ds   (acoustic_matrices.py:68-110).
    Returns None when no Robin/choked boundary exists (B falsy selects EPS)."""
    mesh = space.mesh
    if gamma_nodal is None:
        gamma_nodal = np.full(mesh.n_nodes, 1.4)
    ref = reference_facet_tensors(space.degree)
    area = facet_areas(mesh)
    nfd = space.facet_dofs.shape[1]
    Ke = np.zeros((len(mesh.facets), nfd, nfd), np.complex128)
    any_term = False
    for tag, bc in boundary_conditions.items():
        R = boundary_reflection(mesh, tag, bc, gamma_nodal) if isinstance(bc, (dict,)) else None
        if R is None:
            continue
        any_term = True
        Z = reflection_to_impedance(R)
        sel = np.flatnonzero(mesh.facet_tags == tag)
        if c_is_dg0:
            cf = c[facet_owner_cells(mesh)[sel]]
            Ke[sel] += (1j / Z) * (area[sel] * cf)[:, None, None] * ref["T0"][None]
        else:
            cv = c[mesh.facets[sel]]
            Ke[sel] += (1j / Z) * area[sel][:, None, None] * np.einsum("ijm,fm->fij", ref["T1"], cv)
    if not any_term:
        return None
    return _scatter(space, Ke, dofs=space.facet_dofs)


def dirichlet_dofs(space: Space, boundary_conditions):
    dofs = []
    for tag, bc in boundary_conditions.items():
        if "Dirichlet" in bc:
            dofs.append(space.facet_dofs[space.mesh.facet_tags == tag].ravel())
    return np.unique(np.concatenate(dofs)) if dofs else np.zeros(0, np.int64)


def apply_dirichlet(Mx, dofs):
    """Zero rows+cols, 1.0 on the diagonal (DOLFINx assemble_matrix(bcs=...);
    both A and C get diagonal 1, acoustic_matrices.py:102,122)."""
    if len(dofs) == 0:
        return Mx
    n = Mx.shape[0]
    keep = np.ones(n); keep[dofs] = 0
    Dk = sp.diags(keep)
    out = (Dk @ Mx @ Dk + sp.diags(1 - keep)).tocsr()
    out.sort_indices()
    return out


@dataclass
class Operators:
    """Stand-in for AcousticMatrices' .A .B .B_adj .C (acoustic_matrices.py:127-138)."""
    space: Space
    A: sp.csr_matrix
    B: sp.csr_matrix | None
    B_adj: sp.csr_matrix | None
    C: sp.csr_matrix
    C_nobc: sp.csr_matrix
    c: np.ndarray
    gamma: np.ndarray


def gamma_function(T):
    """parameters_utils.py:62-78."""
    cp = 973.60091 + 0.1333 * T
    return cp / (cp - 287.1)


def sound_speed_variable_gamma(T):
    """parameters_utils.py:80-93."""
    return np.sqrt(gamma_function(T) * 287.1 * T)


def acoustic_matrices(mesh, boundary_conditions, parameter, degree=1, parameter_is_temperature=False,
                      c_is_dg0=False) -> Operators:
    """AcousticMatrices.__init__ (acoustic_matrices.py:12-125)."""
    space = function_space(mesh, degree)
    if parameter_is_temperature:
        c = sound_speed_variable_gamma(parameter)
        gamma = gamma_function(parameter)
    else:
        c = parameter
        gamma = np.full(mesh.n_nodes, 1.4)
    bcd = dirichlet_dofs(space, boundary_conditions)
    A = apply_dirichlet(assemble_A(space, c, c_is_dg0), bcd)
    B = assemble_B(space, c, boundary_conditions, c_is_dg0, gamma)
    C0 = assemble_C(space)
    C = apply_dirichlet(C0, bcd)
    B_adj = B.conj().T.tocsr() if B is not None else None
    if B_adj is not None:
        B_adj.sort_indices()
    return Operators(space, A, B, B_adj, C, C0, c, gamma)


# ---------------------------------------------------------------------------
# parameter fields (helmholtz_x/parameters_utils.py) -- P1 nodal arrays
# ---------------------------------------------------------------------------


def p1_integral(mesh, f):
    vol, _ = geometry(mesh)
    return float((vol * f[mesh.cells].mean(axis=1)).sum())


def gaussian_function(mesh, x_ref, sigma):
    """gaussianFunction (parameters_utils.py:8-43) + normalize (dolfinx_utils.py:32-48)."""
    x_ref = np.asarray(x_ref, float).reshape(-1)[:3]
    r2 = ((mesh.x - x_ref) ** 2).sum(axis=1)
    f = np.exp(-r2 / (2 * sigma ** 2)) / (sigma ** 3 * (2 * np.pi) ** 1.5)
    return f / p1_integral(mesh, f)


def half_gaussian_function(mesh, x_ref, sigma):
    """halfGaussianFunction (parameters_utils.py:45-60)."""
    x_ref = np.asarray(x_ref, float).reshape(-1)[:3]
    h = gaussian_function(mesh, x_ref, sigma)
    h = np.where(mesh.x[:, 2] < x_ref[2], 0.0, h)
    return h / p1_integral(mesh, h)


def rho_step(mesh, x_f, a_f, rho_d, rho_u):
    """parameters_utils.py:103-121 (3-D: along z)."""
    zf = np.asarray(x_f, float).reshape(-1)[2]
    return rho_u + (rho_d - rho_u) / 2 * (1 + np.tanh((mesh.x[:, 2] - zf) / a_f))


def step_field(mesh, x_f, up, down):
    """c_step / temperature_step (parameters_utils.py:129-153,185-208)."""
    zf = np.asarray(x_f, float).reshape(-1)[2]
    return np.where(mesh.x[:, 2] < zf, up, down).astype(float)


def q_multiple(mesh, n_sector):
    """Q_multiple (parameters_utils.py:228-247): DG0, 1/V_f on cells tagged f."""
    vol, _ = geometry(mesh)
    q = np.zeros(mesh.n_cells)
    for f in range(n_sector):
        sel = mesh.cell_tags == f
        q[sel] = 1.0 / vol[sel].sum()
    return q


# ---------------------------------------------------------------------------
# a6: flame transfer functions (helmholtz_x/flame_transfer_function.py)
# ---------------------------------------------------------------------------


class NTau:
    """nTau (flame_transfer_function.py:5-14)."""

    def __init__(self, n, tau):
        self.n, self.tau = n, tau

    def __call__(self, omega):
        return self.n * np.exp(1j * omega * self.tau)

    def derivative(self, omega):
        return self.n * (1j * self.tau) * np.exp(1j * omega * self.tau)


class StateSpace:
    """stateSpace (flame_transfer_function.py:16-42)."""

    def __init__(self, S1, s2, s3, s4):
        self.A, self.b, self.c, self.d = (np.asarray(v) for v in (S1, s2, s3, s4))
        self.Id = np.eye(self.A.shape[0])

    def _eval(self, omega, k):
        omega = np.conj(omega)
        Mat = (-1j) ** k * math.factorial(k) * np.linalg.matrix_power(1j * omega * self.Id - self.A, -(k + 1))
        H = np.dot(np.dot(self.c, Mat), self.b)
        if k == 0:
            H = H + self.d
        return np.conj(H[0][0])

    def __call__(self, omega):
        return self._eval(omega, 0)

    def derivative(self, omega):
        return self._eval(omega, 1)


# ---------------------------------------------------------------------------
# a2-a5: flame operator kept as sparse left/right vectors
# ---------------------------------------------------------------------------


@dataclass
class Flame:
    """D_ij = sum_f left[:,f] right[:,f]^T  (flame_matrices.py:75-94,171-176), kept
    factored; D(omega)=FTF(omega) D_ij (flame_matrices.py:96-108)."""
    left: np.ndarray    # (n, r) real, thresholded
    right: np.ndarray   # (n, r)
    FTF: object

    def factors(self, problem_type="direct"):
        return (self.left, self.right) if problem_type == "direct" else (self.right, self.left)

    def ftf(self, omega, problem_type="direct"):
        if problem_type == "direct":
            return self.FTF(omega)
        return np.conj(self.FTF(np.conj(omega)))

    def dense_block_nnz(self):
        return int(sum(np.count_nonzero(self.left[:, f]) * np.count_nonzero(self.right[:, f])
                       for f in range(self.left.shape[1])))


def _threshold(v, tol):
    """flame_matrices.py:67-68 (values are real here)."""
    v = v.copy()
    v[np.abs(v) < tol] = 0.0
    return v


def _assemble_vector(space, be):
    out = np.zeros(space.n)
    np.add.at(out, space.cell_dofs.ravel(), be.ravel())
    return out


def distributed_flame(mesh, w, h, rho, T, q_0, u_b, FTF, degree=1, gamma=None, tol=1e-5) -> Flame:
    """DistributedFlameMatrix (flame_matrices.py:191-244).
    left_i  = int (gamma-1) q0/u_b h phi_i dx            (:199)
    right_j = int (e_z . grad phi_j) w / rho dx          (:200)
    The right integrand is not polynomial: FFCx estimates degree 2 (P1) / 3 (P2);
    P1 uses the 4-point degree-2 rule (pinned by goldens), P2 the 14-point degree-5
    rule (unpinned, sensitivity ~1e-10 rel. in omega)."""
    space = function_space(mesh, degree)
    vol, G = geometry(mesh)
    ref = reference_tensors(degree)
    cells = mesh.cells
    hc = h[cells]
    if gamma is None:
        gamma = gamma_function(T)
    if np.ndim(gamma) == 0:
        be = (gamma - 1) * q_0 / u_b * vol[:, None] * np.einsum("am,cm->ca", ref["F01"], hc)
    else:
        gm1 = gamma[cells] - 1.0
        be = q_0 / u_b * vol[:, None] * np.einsum("amn,cm,cn->ca", ref["F1"], gm1, hc)
    left = _assemble_vector(space, be)
    if degree == 1:
        Lq, wq = tet_rule_deg2()
    else:
        Lq, wq = tet_rule_deg5_14()
    _, d = tabulate(degree, Lq)
    wq_rho = (w[cells] @ Lq.T) / (rho[cells] @ Lq.T)                # (nc,nq)
    dz = np.einsum("qak,ck->cqa", d, G[:, :, 2])                     # d phi_a/dz at q
    be = vol[:, None] * np.einsum("q,cq,cqa->ca", wq, wq_rho, dz)
    right = _assemble_vector(space, be)
    return Flame(_threshold(left, tol)[:, None], _threshold(right, tol)[:, None], FTF)


def locate_point(mesh, p, tol=1e-10):
    """Lowest-index cell containing p (documented tie-break)."""
    vol, G = geometry(mesh)
    X0 = mesh.x[mesh.cells[:, 0]]
    lam = np.einsum("cai,ci->ca", G[:, 1:, :], p[None, :] - X0)
    L = np.concatenate([1 - lam.sum(axis=1, keepdims=True), lam], axis=1)
    ok = np.flatnonzero(L.min(axis=1) >= -tol)
    if len(ok) == 0:
        raise ValueError(f"point {p} not found in mesh")
    return int(ok[0]), L[ok[0]]


def pointwise_flame(mesh, x_r, h_dg0, rho_u, q_0, u_b, FTF, degree=1, gamma=1.4, tol=1e-10) -> Flame:
    """PointwiseFlameMatrix (flame_matrices.py:129-189): per flame f,
    left^f = (gamma-1) q0/u_b int_{tag f} h phi_j dx  (:141), right^f =
    d(phi_a)/dz (x_r^f) / rho_u on the owning cell's dofs (:144-156)."""
    space = function_space(mesh, degree)
    vol, G = geometry(mesh)
    ref = reference_tensors(degree)
    nf = len(x_r)
    left = np.zeros((space.n, nf)); right = np.zeros((space.n, nf))
    for f in range(nf):
        sel = np.flatnonzero(mesh.cell_tags == f)
        be = (gamma - 1) * q_0 / u_b * (vol[sel] * h_dg0[sel])[:, None] * ref["F0"][None, :]
        np.add.at(left[:, f], space.cell_dofs[sel].ravel(), be.ravel())
        cell, L = locate_point(mesh, np.asarray(x_r[f], float))
        _, d = tabulate(degree, L[None, :])
        dz = d[0] @ G[cell, :, 2]
        right[space.cell_dofs[cell], f] += dz / rho_u
        left[:, f] = _threshold(left[:, f], tol)
        right[:, f] = _threshold(right[:, f], tol)
    return Flame(left, right, FTF)


# ---------------------------------------------------------------------------
# a7/a8: linear eigen-solves (shift-invert, exact LU, flame term by Woodbury)
# ---------------------------------------------------------------------------


class ShiftedSolve:
    """x = (P - U W^T)^{-1} b with sparse LU of P and the rank-r Woodbury update
    (SURVEY App. A "matrix-free flame term").  ``hermitian=True`` solves with the
    conjugate-transposed operator (left eigenvectors)."""

    def __init__(self, P, U=None, W=None):
        self.lu = spla.splu(sp.csc_matrix(P))
        self.U, self.W = U, W
        if U is not None and U.shape[1] > 0:
            self.Z = np.column_stack([self.lu.solve(np.ascontiguousarray(U[:, k])) for k in range(U.shape[1])])
            self.S = np.eye(U.shape[1]) - W.T @ self.Z
            self.ZH = np.column_stack([self.lu.solve(np.ascontiguousarray(np.conj(W[:, k])), trans="H")
                                       for k in range(W.shape[1])])
            self.SH = np.eye(U.shape[1]) - U.conj().T @ self.ZH
        else:
            self.U = None

    def solve(self, b):
        y = self.lu.solve(b)
        if self.U is not None:
            y = y + self.Z @ np.linalg.solve(self.S, self.W.T @ y)
        return y

    def solve_H(self, b):
        y = self.lu.solve(b, trans="H")
        if self.U is not None:
            y = y + self.ZH @ np.linalg.solve(self.SH, self.U.conj().T @ y)
        return y


def _arnoldi_eigs(op, n, nev, ncv=None, v0=None):
    ncv = ncv or max(2 * nev + 1, 20)
    ncv = min(ncv, n - 1)
    if v0 is None:
        v0 = np.random.default_rng(0).standard_normal(n) + 0j
    mu, X = spla.eigs(spla.LinearOperator((n, n), matvec=op, dtype=np.complex128), k=nev, which="LM",
                      ncv=ncv, tol=0, v0=v0, maxiter=50 * n)
    order = np.argsort(-np.abs(mu), kind="stable")
    return mu[order], X[:, order]


@dataclass
class EigenResult:
    """What callers read off a SLEPc EPS/PEP handle (eigenvectors.py:20-33)."""
    kind: str                 # 'eps' | 'pep'
    eigenvalues: np.ndarray   # eps: lambda (=omega^2 in FPI); pep: omega
    vectors: np.ndarray       # (n, nev)
    left_vectors: np.ndarray | None = None

    def omega(self, i):
        return np.sqrt(self.eigenvalues[i]) if self.kind == "eps" else self.eigenvalues[i]


def eps_solve(K, M, sigma, nev, U=None, W=None, two_sided=False) -> EigenResult:
    """K x = lambda M x nearest sigma with K := K - U W^T applied by Woodbury.
    Mirrors eps_solver(A, C, target, nev) (eigensolvers.py:41-67), where the
    caller passes M = -C and sigma = target**2 (:45,:53)."""
    n = K.shape[0]
    S = ShiftedSolve((K - sigma * M).tocsc(), U, W)
    mu, X = _arnoldi_eigs(lambda v: S.solve(M @ v), n, nev)
    lam = sigma + 1.0 / mu
    Y = None
    if two_sided:
        MH = M.conj().T.tocsr()
        muL, YL = _arnoldi_eigs(lambda v: S.solve_H(MH @ v), n, nev)
        lamL = np.conj(sigma + 1.0 / muL)      # eigenvalues of the pencil seen from the left
        Y = np.empty_like(X)
        for i in range(nev):
            j = int(np.argmin(np.abs(lamL - lam[i])))
            Y[:, i] = YL[:, j]
    return EigenResult("eps", lam, X, Y)


def pep_solve(K, B, C, sigma, nev, U=None, W=None) -> EigenResult:
    """(K + omega B + omega^2 C) p = 0 nearest sigma, K := K - U W^T.
    Mirrors pep_solver (eigensolvers.py:69-120; SLEPc TOAR + sinvert) through the
    first companion linearisation, shift-inverted (SURVEY App. A "PEP")."""
    n = K.shape[0]
    S = ShiftedSolve((K + sigma * B + sigma ** 2 * C).tocsc(), U, W)
    BsC = (B + sigma * C).tocsr()

    def op(z):
        u, v = z[:n], z[n:]
        p = -S.solve(C @ v + BsC @ u)
        return np.concatenate([p, u + sigma * p])

    mu, Z = _arnoldi_eigs(op, 2 * n, nev)
    om = sigma + 1.0 / mu
    return EigenResult("pep", om, Z[:n, :])


def _flame_UW(flame, omega, problem_type):
    if flame is None:
        return None, None
    Lf, Rf = flame.factors(problem_type)
    return (flame.ftf(omega, problem_type) * Lf).astype(np.complex128), Rf.astype(np.complex128)


# ---------------------------------------------------------------------------
# a9/a10: nonlinear iterations (helmholtz_x/eigensolvers.py:122-348)
# ---------------------------------------------------------------------------


def fixed_point_iteration(ops: Operators, flame: Flame, target, nev=2, i=0, tol=1e-8, maxiter=50,
                          problem_type="direct", log=None):
    """fixed_point_iteration (eigensolvers.py:261-276) -> (EigenResult, omegas).
    EPS variant :122-195, PEP variant :197-259.  Returns the LAST linear solve
    (callers read omega off it) and the omega_k history."""
    if problem_type not in ("direct", "adjoint"):
        raise ValueError("The problem type should be specified as 'direct' or 'adjoint'.")
    A, C = ops.A, ops.C
    B = ops.B if problem_type == "direct" else ops.B_adj
    quadratic = ops.B is not None
    if quadratic:
        E = pep_solve(A, B, C, target, nev)
        omega = [E.eigenvalues[i]]
    else:
        E = eps_solve(A, -C, target ** 2, nev)
        omega = [np.sqrt(E.eigenvalues[i])]
    f = []; alpha = [0.5]
    hist = [omega[0]]
    domega = 2 * tol
    k = -1
    while abs(domega) > tol:
        k += 1
        if k >= maxiter:
            raise IndexError("fixed_point_iteration: maxiter exceeded (the reference overruns its arrays here)")
        U, W = _flame_UW(flame, omega[k], problem_type)
        if quadratic:
            E = pep_solve(A, B, C, target, nev, U, W)
            f.append(E.eigenvalues[i])
        else:
            E = eps_solve(A, -C, target ** 2, nev, U, W)
            f.append(np.sqrt(E.eigenvalues[i]))
        if k != 0:
            alpha.append(1 / (1 - ((f[k] - f[k - 1]) / (omega[k] - omega[k - 1]))))
        omega.append(alpha[k] * f[k] + (1 - alpha[k]) * omega[k])
        domega = omega[k + 1] - omega[k]
        hist.append(omega[k + 1])
        if log:
            log(k, omega[k + 1], abs(domega))
    return E, np.array(hist)


def fix_sign(v):
    """FixSign (petsc4py_utils.py:100-111)."""
    x0 = v[0]
    return v / (x0 / abs(x0))


def normalize_eigenvector(ops: Operators, E: EigenResult, i, which="right", absolute=False):
    """normalize_eigenvector (eigenvectors.py:11-64): FixSign, divide by
    sqrt(p^T M p) with the *unconjugated* mass form (:47)."""
    omega = E.omega(i)
    v = E.vectors[:, i] if which == "right" else E.left_vectors[:, i]
    v = fix_sign(v)
    meas = np.sqrt(v @ (ops.C_nobc @ v))
    out = v / meas
    if absolute:
        out = np.abs(out) / np.abs(out).max()
    return omega, out


def vector_matrix_vector(y, Mx, x):
    """petsc4py_utils.py:67-89 with petsc4py's Vec.dot convention: y.dot(Ax) =
    sum_i y_i conj((Ax)_i)  (SURVEY App. C.1)."""
    return np.vdot(Mx @ x, y)


def newton_solver(ops: Operators, flame: Flame, init, nev=2, i=0, tol=1e-3, maxiter=100, log=None):
    """newtonSolver (eigensolvers.py:278-348), bug-compatible with the
    conjugated derivative (App. C.1) and relaxation *= 0.8 (:337)."""
    A, B, C = ops.A, ops.B, ops.C
    omega = [complex(init)]
    domega = 2 * tol
    k = 0
    relaxation = 1.0
    p = None
    while abs(domega) > tol:
        om = omega[k]
        ftf = flame.FTF(om); dftf = flame.FTF.derivative(om)
        Lf, Rf = flame.factors("direct")
        L = A + om ** 2 * C if B is None else A + om * B + om ** 2 * C
        dL = 2 * om * C if B is None else B + 2 * om * C
        U = (ftf * Lf).astype(np.complex128); W = Rf.astype(np.complex128)
        E = eps_solve(L.tocsr(), C, 0.0, nev, U, W, two_sided=True)
        eig = E.eigenvalues[i]
        _, p = normalize_eigenvector(ops, E, i, "right")
        _, p_adj = normalize_eigenvector(ops, E, i, "left")
        dLp = dL @ p - dftf * (Lf @ (Rf.T @ p))
        num = np.vdot(dLp, p_adj)
        den = np.vdot(C @ p, p_adj)
        deig = num / den
        domega = -relaxation * eig / deig
        relaxation *= 0.8
        omega.append(om + domega)
        if log:
            log(k, omega[k + 1], abs(domega))
        k += 1
        if k >= maxiter - 1:
            break
    return omega[k], p, np.array(omega)


def normalize_adjoint(ops: Operators, omega_dir, p_dir, p_adj, flame: Flame | None = None):
    """normalize_adjoint (eigenvectors.py:125-177)."""
    dLp = 2 * omega_dir * (ops.C @ p_dir)
    if ops.B is not None:
        dLp = dLp + ops.B @ p_dir
    if flame is not None:
        Lf, Rf = flame.factors("direct")
        dLp = dLp - flame.FTF.derivative(omega_dir) * (Lf @ (Rf.T @ p_dir))
    meas = np.vdot(dLp, p_adj)
    out = p_adj / meas
    return out, np.vdot(dLp, out)


# ---------------------------------------------------------------------------
# K7/K8 reference kernels used by the GPU parity tests
# ---------------------------------------------------------------------------


def spmv(indptr, indices, values, x):
    """y = M x for complex128 CSR (PETSc MatMult; petsc4py_utils.py:86,96)."""
    Mx = sp.csr_matrix((values, indices, indptr), shape=(len(indptr) - 1, len(x)))
    return Mx @ x


def fused_apply(ops: Operators, flame, sigma, ftf, x, problem_type="direct"):
    """y = (A + sigma B + sigma^2 C) x - ftf * L (R^T x)   (eigensolvers.py:174-176,
    240, 309-315 without forming the sum)."""
    y = ops.A @ x + sigma ** 2 * (ops.C @ x)
    if ops.B is not None:
        B = ops.B if problem_type == "direct" else ops.B_adj
        y = y + sigma * (B @ x)
    if flame is not None:
        Lf, Rf = flame.factors(problem_type)
        y = y - ftf * (Lf @ (Rf.T @ x))
    return y


# ---------------------------------------------------------------------------
# (f-1) adjoint sensitivity tail: boundary shape-derivative integral
# ---------------------------------------------------------------------------


def facet_normals(mesh: Mesh):
    """Outward unit normals of the tagged boundary facets (UFL FacetNormal)."""
    X = mesh.x[mesh.facets]
    nrm = np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0])
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    owner = facet_owner_cells(mesh)
    centroid = mesh.x[mesh.cells[owner]].mean(axis=1)
    sign = np.sign(np.einsum("fi,fi->f", nrm, X.mean(axis=1) - centroid))
    return nrm * sign[:, None]


def shape_derivative(space: Space, tag, V_nodal, p_dir, p_adj_normalized, c_nodal):
    """int_{ds(tag)} (V.n) div( conj(p_adj) c^2 grad p ) ds  for one displacement field V
    (helmholtz_x/shape_derivatives.py:12-37: G_neu = div(p_adj_conj * c**2 * grad(p_dir)), the
    form inner(V_ffd, normal) * G_neu * ds(tag)).  V, c are P1 nodal; p, p_adj live in `space`.
    Evaluated in the owning cell of each facet with a degree-7 triangle rule."""
    mesh = space.mesh
    sel = np.flatnonzero(mesh.facet_tags == tag)
    owner = facet_owner_cells(mesh)[sel]
    vol, G = geometry(mesh)
    nrm = facet_normals(mesh)[sel]
    area = facet_areas(mesh)[sel]
    Lq, wq = tri_rule(4)
    fac = mesh.facets[sel]
    cells = mesh.cells[owner]
    # barycentric coordinates (in the owner cell) of the facet quadrature points
    loc = np.array([[int(np.flatnonzero(cells[f] == fac[f, k])[0]) for k in range(3)] for f in range(len(sel))])
    pa = np.conj(p_adj_normalized)
    total = 0.0 + 0.0j
    for q in range(len(wq)):
        L = np.zeros((len(sel), 4))
        for k in range(3):
            L[np.arange(len(sel)), loc[:, k]] = Lq[q, k]
        val = np.zeros(len(sel), complex)
        for f in range(len(sel)):
            Gc = G[owner[f]]                                   # (4,3) grad L_a
            phi, d = tabulate(space.degree, L[f][None, :])
            dofs = space.cell_dofs[owner[f]]
            gphi = d[0] @ Gc                                   # (nd,3) physical gradients
            p, gp = phi[0] @ p_dir[dofs], gphi.T @ p_dir[dofs]
            q_, gq = phi[0] @ pa[dofs], gphi.T @ pa[dofs]
            cc = c_nodal[cells[f]]
            cval, gc = L[f] @ cc, Gc.T @ cc
            lap = 0.0
            if space.degree == 2:
                # Laplacian of P2 basis: vertex a: 4 |G_a|^2 ; edge (a,b): 8 G_a.G_b
                GG = Gc @ Gc.T
                lapphi = np.array([4 * GG[a, a] for a in range(4)] + [8 * GG[a, b] for a, b in TET_EDGES])
                lap = lapphi @ p_dir[dofs]
            div = cval ** 2 * (gq @ gp) + q_ * 2 * cval * (gc @ gp) + q_ * cval ** 2 * lap
            Vq = L[f] @ V_nodal[cells[f]]
            val[f] = (Vq @ nrm[f]) * div
        total += wq[q] * (area * val).sum()
    return total


# ---------------------------------------------------------------------------
# (f-3) Bloch-periodic reduction   (helmholtz_x/bloch_operator.py)
# ---------------------------------------------------------------------------


def bloch_pairs(space: Space, master_tag, slave_tag, N, pairing="geometric", numbering=None, tol=1e-8):
    """(master dofs, slave dofs) paired.  pairing="geometric": a master dof is the image of its
    slave under the rotation by -2 pi / N about z.  pairing="sorted" restates the reference:
    k-th master with k-th slave in ascending dof index (bloch_operator.py:33-40); `numbering`
    (reference dof index of every dof here) makes that order the reference's DOLFINx order
    (only 49 % of those pairs are geometric images)."""
    mesh = space.mesh
    md = np.unique(space.facet_dofs[mesh.facet_tags == master_tag])
    sd = np.unique(space.facet_dofs[mesh.facet_tags == slave_tag])
    assert len(md) == len(sd)
    if pairing == "sorted":
        if numbering is not None:
            md = md[np.argsort(numbering[md])]
            sd = sd[np.argsort(numbering[sd])]
        return md, sd
    from scipy.spatial import cKDTree
    X = space.dof_x
    best = None
    for sgn in (1.0, -1.0):
        a = sgn * 2 * np.pi / N
        Rm = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        d, j = cKDTree(X[sd]).query(X[md] @ Rm.T)
        if best is None or d.max() < best[0]:
            best = (d.max(), j)
    assert best[0] < tol * max(np.abs(X).max(), 1.0) * 1e3, f"master/slave faces are not rotation images (max dist {best[0]})"
    return md, sd[best[1]]


def bloch_maps(n, masters, slaves, N, b=1.0):
    """BN (n x n_red) and NB = BN^H (bloch_operator.py:42-68): reduced dofs = all but the
    masters; full[master_k] = f_b * reduced[slave_k], f_b = exp(2 pi i b / N)."""
    f_b = np.exp(b * 1j * 2 * np.pi / N)
    keep = np.ones(n, bool)
    keep[masters] = False
    red = -np.ones(n, np.int64)
    red[keep] = np.arange(keep.sum())
    rows = np.concatenate([np.flatnonzero(keep), masters])
    cols = np.concatenate([red[keep], red[slaves]])
    vals = np.concatenate([np.ones(keep.sum(), complex), np.full(len(masters), f_b)])
    BN = sp.csr_matrix((vals, (rows, cols)), shape=(n, int(keep.sum())))
    return BN, BN.conj().T.tocsr()


def blochify(Mx, BN, NB):
    out = (NB @ Mx @ BN).tocsr()
    out.sort_indices()
    return out


def bloch_operators(ops: Operators, BN, NB) -> Operators:
    """Blochifier.A/.B/.C (bloch_operator.py:70-78); B_adj stays unset as in the reference (:25,:89-90)."""
    return Operators(ops.space, blochify(ops.A, BN, NB), blochify(ops.B, BN, NB) if ops.B is not None else None, None,
                     blochify(ops.C, BN, NB), ops.C_nobc, ops.c, ops.gamma)


def bloch_flame(flame: Flame, BN, NB) -> Flame:
    """FlameMatrix.blochify (flame_matrices.py:117-127): NB (l r^T) BN = (NB l)(BN^T r)^T."""
    return Flame(np.asarray(NB @ flame.left), np.asarray(BN.T @ flame.right), flame.FTF)
