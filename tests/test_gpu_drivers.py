"""The reference's own driver scripts, byte for byte (tests/golden/reference_drivers.json, packed by
make_driver_fixture.py with their sha256), executed on the GPU with PYTHONPATH=compat -- the import shim
that maps `helmholtz_x`, `dolfinx.fem`, `petsc4py.PETSc`, `mpi4py.MPI` onto helmholtz_x_b200 -- in a
scratch directory holding the mesh the driver expects (MeshDir/mesh.xdmf, written from the committed
mesh fixture) and, for the annulus, ftf.mat.  What they print / save is compared with the golden logs:

  numerical_examples/Longitudinal/NetworkCode/RijkeTube3D/active.py    vs Results/Active/active.log:23-52
  numerical_examples/AnnularCombustor/Micca/fullAnnulus/active_fpi.py  vs Results/Active/FPI/eigenvalues_{dir,adj}.txt
(BASELINE north_star: "the numerical_examples drivers run unchanged"; VERDICT r1 missing-5.)"""
import hashlib
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = cases.golden_values()


def _stage(tmp_path, case, mesh_name, results):
    from helmholtz_x_b200.io_utils import write_mesh_xdmf
    with open(os.path.join(cases.GOLDEN_DIR, "reference_drivers.json")) as fh:
        files = json.load(fh)[case]
    for name, rec in files.items():
        raw = "\n".join(rec["lines"]).encode()
        assert hashlib.sha256(raw).hexdigest() == rec["sha256"], name          # verbatim
        (tmp_path / name).write_bytes(raw)
    m = cases.mesh(mesh_name)
    os.makedirs(tmp_path / "MeshDir")
    write_mesh_xdmf(str(tmp_path / "MeshDir" / "mesh"), m.x, m.cells, m.cell_tags, m.facets, m.facet_tags)
    os.makedirs(tmp_path / results)
    return files


def _run(tmp_path, script):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "compat"), ROOT, env.get("PYTHONPATH", "")])
    res = subprocess.run([sys.executable, script], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    return res.stdout


def _omegas(stdout):
    pat = re.compile(r"([+-]\d+\.\d+)\s+([+-]\d+\.\d+)j")
    out = []
    for ln in stdout.splitlines():
        if "omega =" in ln or "Starting eigenvalue" in ln:
            mt = pat.search(ln)
            out.append(complex(float(mt.group(1)), float(mt.group(2))))
    return out


def test_reference_driver_rijke3d_active_runs_unchanged(tmp_path):
    _stage(tmp_path, "rijke3d", "rijke3d", os.path.join("Results", "Active"))
    out = _run(tmp_path, "active.py")
    got = _omegas(out)
    gold = [cases.cplx(p) for p in G["rijke3d_active_fpi"]["omegas"]]
    assert len(got) == len(gold), out[-1500:]
    for a, b in zip(got, gold):
        assert abs(a - b) < 2e-8, (a, b)                       # the log prints 8 decimals
    assert "Total Execution Time" in out
    for f in ("p", "p_abs", "p_phase"):
        assert os.path.exists(tmp_path / "Results" / "Active" / (f + ".xdmf"))
        assert os.path.exists(tmp_path / "Results" / "Active" / (f + ".h5"))


def test_reference_driver_annulus_active_fpi_runs_unchanged(tmp_path):
    from scipy.io import savemat
    from helmholtz_x_b200.io_utils import dict_loader
    _stage(tmp_path, "annulus", "annulus", os.path.join("Results", "Active", "FPI"))
    f = np.load(os.path.join(cases.GOLDEN_DIR, "annulus_ftf.npz"))
    savemat(str(tmp_path / "ftf.mat"), {"A": f["A"], "b": f["b"], "c": f["c"], "d": f["d"]})
    out = _run(tmp_path, "active_fpi.py")
    assert "Total Execution Time" in out
    for name, gkey in (("eigenvalues_dir", "annulus_fpi_eigenvalues_dir"), ("eigenvalues_adj", "annulus_fpi_eigenvalues_adj")):
        d = dict_loader(str(tmp_path / "Results" / "Active" / "FPI" / name))
        for k, g in G[gkey].items():
            if k == "source":
                continue
            g = cases.cplx(g)
            assert abs(complex(d[k]) - g) / abs(g) < 1e-8, (k, d[k], g)
    for fn in ("p_1_dir", "p_2_dir", "u_1_dir", "p_1_adj", "p_2_adj"):
        assert os.path.exists(tmp_path / "Results" / "Active" / "FPI" / (fn + ".h5"))
