"""Build the committed fixtures under tests/golden/ from the reference tree.

Run once in the build container (``/root/reference`` is not present on the GPU
box, so nothing at test time may read it):

    python tests/golden/make_fixtures.py

Writes
  * ``<case>_mesh.npz``      mesh arrays read from the reference's XDMF/HDF5 pair
                             (data files, not source code),
  * ``golden_values.json``   eigenvalues / iteration histories transcribed from the
                             reference's committed logs and result files, each with
                             its file:line provenance,
  * ``<case>_p.npz``         golden eigenvectors from result .h5 files with their
                             DOLFINx geometry (used to map node orders).
"""
import ast
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from helmholtz_x_b200.h5lite import H5File  # noqa: E402

REF = "/root/reference/numerical_examples/"

MESHES = {
    "rijke3d": "Longitudinal/NetworkCode/RijkeTube3D/MeshDir/mesh",
    "prf_rijke3d": "Longitudinal/PRF/RijkeTube3D/MeshDir/mesh",
    "rijkeffd": "ShapeSensitivities/RijkeFFD/MeshDir/Original/mesh",
    "annulus": "AnnularCombustor/Micca/fullAnnulus/MeshDir/mesh",
    "bloch": "AnnularCombustor/Micca/bloch/MeshDir/mesh",
    "flamedduct": "Longitudinal/NetworkCode/FlamedDuct/MeshDir/mesh",
}


def read_mesh(stem):
    m = H5File(REF + stem + ".h5")
    t = H5File(REF + stem + "_tags.h5")
    x = m["/data0"]
    assert np.array_equal(x, t["/data0"])
    return dict(x=x, cells=m["/data1"].astype(np.int32), cell_tags=m["/data2"].astype(np.int32),
                facets=t["/data1"].astype(np.int32), facet_tags=t["/data2"].astype(np.int32))


def omegas_from_log(path, first, last):
    """'+ omega = +1248.52758621  +3.44467515j' / 'Starting eigenvalue' lines."""
    out = []
    with open(REF + path) as fh:
        lines = fh.read().splitlines()
    pat = re.compile(r"([+-]\d+\.\d+)\s+([+-]\d+\.\d+)j")
    for ln in lines[first - 1:last]:
        if "omega =" in ln or "Starting eigenvalue" in ln:
            m = pat.search(ln)
            out.append([float(m.group(1)), float(m.group(2))])
    return out


def dict_file(path):
    with open(REF + path) as fh:
        s = json.load(fh)
    s = re.sub(r"np\.complex128\(([^)]*)\)", r"(\1)", s)
    d = ast.literal_eval(s)
    return {k: [complex(v).real, complex(v).imag] for k, v in d.items()}


def main():
    for name, stem in MESHES.items():
        if name == "prf_rijke3d":
            a = read_mesh(stem); b = read_mesh(MESHES["rijke3d"])
            assert all(np.array_equal(a[k], b[k]) for k in a), "PRF mesh differs from Rijke3D mesh"
            continue
        np.savez_compressed(os.path.join(HERE, name + "_mesh.npz"), **read_mesh(stem))
        print("mesh", name)

    g = {}
    L = "Longitudinal/NetworkCode/RijkeTube3D/Results/"
    g["rijke3d_active_fpi"] = dict(
        source=L + "Active/active.log:23-52", omegas=omegas_from_log(L + "Active/active.log", 23, 52))
    g["rijke3d_passive_eps"] = dict(source=L + "Passive/passive.log:30-33",
                                    lambdas=[1133475.711956, 0.0, 4530504.974779, 10202766.022649])
    P = "Longitudinal/PRF/RijkeTube3D/Results/Active/active.log"
    g["prf_rijke3d_direct_fpi"] = dict(source=P + ":21-50", omegas=omegas_from_log(P, 21, 50))
    g["prf_rijke3d_adjoint_fpi"] = dict(source=P + ":57-86", omegas=omegas_from_log(P, 57, 86))
    F = "ShapeSensitivities/RijkeFFD/Results/Original/"
    g["rijkeffd_direct_fpi"] = dict(source=F + "results.log:21-57", omegas=omegas_from_log(F + "results.log", 21, 57))
    g["rijkeffd_adjoint_fpi"] = dict(source=F + "results.log:64-100", omegas=omegas_from_log(F + "results.log", 64, 100))
    g["rijkeffd_eigenvalues"] = dict(source=F + "eigenvalues.txt", **dict_file(F + "eigenvalues.txt"))
    A = "AnnularCombustor/Micca/fullAnnulus/Results/Active/"
    g["annulus_fpi_direct"] = dict(source=A + "FPI/active.log:43-92", omegas=omegas_from_log(A + "FPI/active.log", 43, 92))
    g["annulus_fpi_eigenvalues_dir"] = dict(source=A + "FPI/eigenvalues_dir.txt", **dict_file(A + "FPI/eigenvalues_dir.txt"))
    g["annulus_fpi_eigenvalues_adj"] = dict(source=A + "FPI/eigenvalues_adj.txt", **dict_file(A + "FPI/eigenvalues_adj.txt"))
    g["annulus_newton_eigenvalues"] = dict(source=A + "NewtonSolver/eigenvalues.txt", **dict_file(A + "NewtonSolver/eigenvalues.txt"))
    nl = []
    with open(REF + A + "NewtonSolver/active.log") as fh:
        for ln in fh.read().splitlines()[40:149]:
            m = re.search(r"iter =\s*(\d+),\s+omega = ([+-]\d+\.\d+)\s+([+-]\d+\.\d+)j", ln)
            if m:
                nl.append([float(m.group(2)), float(m.group(3))])
    g["annulus_newton_mode1"] = dict(source=A + "NewtonSolver/active.log:41-149", omegas=nl)
    D = "Longitudinal/NetworkCode/FlamedDuct/Results/Active/active.log"
    g["flamedduct_active_fpi"] = dict(source=D + ":28-56", omegas=omegas_from_log(D, 21, 56))
    B = "AnnularCombustor/Micca/bloch/Results/Passive/passive.log"
    g["bloch_passive"] = dict(source=B + ":27-31", omegas=[2931.177998, 4633.352640, 11107.674019])
    BA = "AnnularCombustor/Micca/bloch/Results/Active/active.log"
    g["bloch_active_fpi"] = dict(source=BA + ":38-75", omegas=omegas_from_log(BA, 38, 75),
                                 final=[3235.145363, 436.054594])
    # config 2: manufactured 2-D duct, analytic dispersion roots written by the reference's MATLAB script
    rows = [ln.split() for ln in open(REF + "manufacturedSolution/matlab_data/analytical.txt").read().splitlines() if ln.strip()]
    sel = [0, 60, 140, 260, 340, 399]
    zb = np.linspace(-10j, 10j, 400); za = np.linspace(-10, 10, 400)
    g["manufactured_analytic"] = dict(
        source="manufacturedSolution/matlab_data/analytical.txt (rows %s; columns Re f_b, Im f_b, Re f_a, Im f_a in Hz, 1 decimal)" % sel,
        Z_imag=[[0.0, float(zb[i].imag)] for i in sel], f_imagZ=[[float(rows[i][0]), float(rows[i][1])] for i in sel],
        Z_real=[[float(za[i]), 0.0] for i in sel], f_realZ=[[float(rows[i][2]), float(rows[i][3])] for i in sel])
    with open(os.path.join(HERE, "golden_values.json"), "w") as fh:
        json.dump(g, fh, indent=1)
    print("golden values", {k: len(v.get("omegas", [])) for k, v in g.items()})

    # golden eigenvectors (DOLFINx node order + geometry)
    for name, path in {"rijke3d_active": L + "Active/p.h5", "rijke3d_passive": L + "Passive/p.h5",
                       "rijkeffd_dir": F + "p_dir.h5", "rijkeffd_adj": F + "p_adj.h5",
                       "bloch_passive1": "AnnularCombustor/Micca/bloch/Results/Passive/p_1.h5",
                       "bloch_active1": "AnnularCombustor/Micca/bloch/Results/Active/p_1_dir.h5"}.items():
        f = H5File(REF + path)
        re_k = [k for k in f.keys() if "/real_" in k][0]
        im_k = [k for k in f.keys() if "/imag_" in k][0]
        np.savez_compressed(os.path.join(HERE, name + "_p.npz"), geometry=f["/Mesh/Grid/geometry"],
                            p=(f[re_k][:, 0] + 1j * f[im_k][:, 0]))
        print("vector", name)
    # annulus FTF state-space matrices
    from scipy.io import loadmat
    m = loadmat(REF + "AnnularCombustor/Micca/fullAnnulus/ftf.mat")
    np.savez(os.path.join(HERE, "annulus_ftf.npz"), A=m["A"], b=m["b"], c=m["c"], d=m["d"])
    mb = loadmat(REF + "AnnularCombustor/Micca/bloch/ftf.mat")
    assert all(np.array_equal(m[k], mb[k]) for k in "Abcd")


if __name__ == "__main__":
    main()
