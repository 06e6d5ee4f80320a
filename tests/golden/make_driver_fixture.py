"""Pack the reference's own driver scripts into tests/golden/reference_drivers.json (TEST DATA: the
files are kept byte for byte so that the `-m gpu` test can execute them UNMODIFIED against the compat
import shim; /root/reference does not exist on the GPU box).

    python tests/golden/make_driver_fixture.py

Sources (numerical_examples/...):
  Longitudinal/NetworkCode/RijkeTube3D/{active.py, params.py}         BASELINE config 1
  AnnularCombustor/Micca/fullAnnulus/{active_fpi.py, params.py}       BASELINE config 3
"""
import hashlib
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/numerical_examples/"
FILES = {
    "rijke3d": ["Longitudinal/NetworkCode/RijkeTube3D/active.py", "Longitudinal/NetworkCode/RijkeTube3D/params.py"],
    "annulus": ["AnnularCombustor/Micca/fullAnnulus/active_fpi.py", "AnnularCombustor/Micca/fullAnnulus/params.py"],
}


def main():
    out = {}
    for case, paths in FILES.items():
        out[case] = {}
        for p in paths:
            with open(REF + p, "rb") as fh:
                raw = fh.read()
            out[case][os.path.basename(p)] = {"source": "numerical_examples/" + p, "sha256": hashlib.sha256(raw).hexdigest(),
                                              "lines": raw.decode().split("\n")}
    with open(os.path.join(HERE, "reference_drivers.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print({k: list(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
