"""helmholtz_x/io_utils.py subset: XDMF(+HDF5) mesh reader, XDMF writer, dict I/O."""
import ast
import json
import os
import re

import numpy as np

from .fem import Mesh, MeshTags
from .h5lite import H5File, write_h5
from .solver_utils import info, rank0


def dict_writer(filename, dictionary, extension=".txt"):
    with open(filename + extension, 'w') as file:
        file.write(json.dumps(str(dictionary)))
    if rank0():
        print(filename + extension, " is saved.")


def dict_loader(filename, extension=".txt"):
    with open(filename + extension) as f:
        data = json.load(f)
    data = ast.literal_eval(re.sub(r"np\.complex128\(([^)]*)\)", r"(\1)", data))
    if rank0():
        print(filename + extension, " is loaded.")
    return data


def _h5_items(xdmf_path):
    with open(xdmf_path) as fh:
        txt = fh.read()
    return re.findall(r">\s*([^<>:\s]+\.h5):(/[^<\s]+)\s*<", txt)


class XDMFReader:
    """XDMFReader(name) reads name.xdmf/.h5 and name_tags.xdmf/.h5 (io_utils.py:161-217)."""

    def __init__(self, name):
        self.name = name
        base = os.path.dirname(name)
        items = _h5_items(name + ".xdmf")
        f = H5File(os.path.join(base, items[0][0]))
        x, cells = f[items[0][1]], f[items[1][1]]
        ctags = f[items[2][1]] if len(items) > 2 else np.zeros(len(cells), np.int32)
        info("\nXDMF Mesh - Cell data loaded.")
        titems = _h5_items(name + "_tags.xdmf")
        t = H5File(os.path.join(base, titems[0][0]))
        facets, ftags = t[titems[1][1]], t[titems[2][1]]
        info("XDMF Mesh - Facet data loaded.")
        self._mesh = Mesh(x, cells, ctags, facets, ftags)
        self._cell_tags = MeshTags(self._mesh.cell_tags)
        self._facet_tags = MeshTags(self._mesh.facet_tags)

    @property
    def mesh(self):
        return self._mesh

    @property
    def subdomains(self):
        return self._cell_tags

    @property
    def facet_tags(self):
        return self._facet_tags

    @property
    def dimension(self):
        return 3

    def getAll(self):
        return self.mesh, self.subdomains, self.facet_tags

    def getInfo(self):
        if rank0():
            print("Number of cells:  {:,}".format(self._mesh.n_cells))
            print("Number of cores: ", int(os.environ.get("WORLD_SIZE", "1")), "\n")
        return self._mesh.n_cells


def xdmf_writer(name, mesh, function):
    """Write mesh + P1 function as XDMF/HDF5 with DOLFINx's dataset layout
    (/Mesh/Grid/{geometry,topology}, /Function/{real_,imag_}<name>/0; io_utils.py:40-60)."""
    bs = getattr(function.function_space, "bs", 1)            # 3: blocked vector field (velocity_eigenvector)
    vals = np.asarray(function.x.array)[:mesh.n_nodes * bs]
    fname = function.name
    h5 = os.path.basename(name) + ".h5"
    write_h5(name + ".h5", {
        "/Mesh/Grid/geometry": mesh.x, "/Mesh/Grid/topology": mesh.cells.astype(np.int64),
        f"/Function/real_{fname}/0": np.real(vals).reshape(-1, bs).astype(np.float64),
        f"/Function/imag_{fname}/0": np.imag(vals).reshape(-1, bs).astype(np.float64)})
    n, nc = mesh.n_nodes, mesh.n_cells
    kind = "Scalar" if bs == 1 else "Vector"

    def attr(part):
        return (f'<Attribute Name="{part}_{fname}" AttributeType="{kind}" Center="Node"><DataItem Dimensions="{n} {bs}" '
                f'Format="HDF">{h5}:/Function/{part}_{fname}/0</DataItem></Attribute>')
    with open(name + ".xdmf", "w") as fh:
        fh.write(f'<Xdmf Version="3.0"><Domain><Grid Name="Grid" GridType="Uniform">'
                 f'<Topology TopologyType="Tetrahedron" NumberOfElements="{nc}" NodesPerElement="4">'
                 f'<DataItem Dimensions="{nc} 4" NumberType="Int" Format="HDF">{h5}:/Mesh/Grid/topology</DataItem></Topology>'
                 f'<Geometry GeometryType="XYZ"><DataItem Dimensions="{n} 3" Format="HDF">{h5}:/Mesh/Grid/geometry</DataItem>'
                 f'</Geometry>{attr("real")}{attr("imag")}</Grid></Domain></Xdmf>')


def write_mesh_xdmf(name, x, cells, cell_tags, facets, facet_tags):
    """Write the XDMF/HDF5 mesh pair XDMFReader expects (the layout meshio produces in the
    reference's write_xdmf_mesh, io_utils.py:98-139): name.{xdmf,h5} with points / tetrahedra /
    cell tags and name_tags.{xdmf,h5} with points / boundary triangles / facet tags."""
    base = os.path.basename(name)
    x = np.asarray(x, np.float64)
    for suffix, conn, tags, ttype, npe in (("", cells, cell_tags, "Tetrahedron", 4), ("_tags", facets, facet_tags, "Triangle", 3)):
        conn = np.asarray(conn, np.int64)
        tags = np.asarray(tags, np.int64)
        h5 = base + suffix + ".h5"
        write_h5(name + suffix + ".h5", {"/data0": x, "/data1": conn, "/data2": tags})
        with open(name + suffix + ".xdmf", "w") as fh:
            fh.write(f'<Xdmf Version="3.0"><Domain><Grid Name="Grid"><Geometry GeometryType="XYZ">'
                     f'<DataItem DataType="Float" Dimensions="{len(x)} 3" Format="HDF" Precision="8">{h5}:/data0</DataItem>'
                     f'</Geometry><Topology TopologyType="{ttype}" NumberOfElements="{len(conn)}" NodesPerElement="{npe}">'
                     f'<DataItem DataType="Int" Dimensions="{len(conn)} {npe}" Format="HDF" Precision="8">{h5}:/data1</DataItem>'
                     f'</Topology><Attribute Name="name_to_read" AttributeType="Scalar" Center="Cell">'
                     f'<DataItem DataType="Int" Dimensions="{len(conn)}" Format="HDF" Precision="8">{h5}:/data2</DataItem>'
                     f'</Attribute></Grid></Domain></Xdmf>')
