"""Build product-side (GPU) objects from the shared case definitions in tests/cases.py."""
import numpy as np

from tests import cases


def gpu_mesh(case):
    from helmholtz_x_b200 import fem
    m = case.mesh
    if "gpu" not in m._cache:
        m._cache["gpu"] = fem.Mesh(m.x, m.cells, m.cell_tags, m.facets, m.facet_tags)
    return m._cache["gpu"]


def gpu_field(mesh, values, dg0=False, name="f"):
    from helmholtz_x_b200 import fem
    V = fem.DG0Space(mesh) if dg0 else fem.functionspace(mesh, ("CG", 1))
    return fem.Function(V, np.asarray(values, float), dtype=np.float64, name=name)


def gpu_operators(case, passive=False):
    from helmholtz_x_b200.acoustic_matrices import AcousticMatrices
    from helmholtz_x_b200.fem import MeshTags
    mesh = gpu_mesh(case)
    cvals = case["c_passive"] if passive else case["c"]
    name = "temperature" if case.parameter_is_temperature else "soundspeed"
    param = gpu_field(mesh, cvals, dg0=case.c_is_dg0, name=name)
    return AcousticMatrices(mesh, MeshTags(mesh.facet_tags), case.bcs, param, degree=case.degree)


def gpu_ftf(case):
    from helmholtz_x_b200.flame_transfer_function import nTau, stateSpace
    spec = case.ftf
    return nTau(spec[1], spec[2]) if spec[0] == "ntau" else stateSpace(*spec[1:])


def gpu_flame(case, mesh=None, bloch_object=None):
    from helmholtz_x_b200.flame_matrices import DistributedFlameMatrix, PointwiseFlameMatrix
    from helmholtz_x_b200.fem import MeshTags
    mesh = mesh or gpu_mesh(case)
    ftf = gpu_ftf(case)
    if case.flame == "distributed":
        w, h, rho, T = (gpu_field(mesh, case[k]) for k in ("w", "h", "rho", "T"))
        return DistributedFlameMatrix(mesh, w, h, rho, T, case.q_0, case.u_b, ftf, degree=case.degree, gamma=case.gamma)
    h = gpu_field(mesh, case.h, dg0=True)
    return PointwiseFlameMatrix(mesh, MeshTags(mesh.cell_tags), case.x_r, h, case.rho_u, case.q_0, case.u_b, ftf,
                                degree=case.degree, bloch_object=bloch_object, gamma=case.gamma)
