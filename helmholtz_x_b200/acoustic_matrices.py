"""Drop-in for helmholtz_x/acoustic_matrices.py: AcousticMatrices assembles A, B, C on
the device (kernels K1-K4) and exposes them as .A .B .B_adj .C."""
import numpy as np

from . import fem
from .operators import Mat, OperatorSet
from .parameters_utils import gamma_function, sound_speed_variable_gamma
from .phases import phase
from .solver_utils import info, rank0


class AcousticMatrices:

    def __init__(self, mesh, facet_tags, boundary_conditions, parameter, degree=1):
        self.mesh = mesh
        self.facet_tags = facet_tags
        self.boundary_conditions = boundary_conditions
        self.parameter = parameter
        self.degree = degree
        self.dimension = 3
        self.fdim = 2
        self.V = fem.functionspace(mesh, ("Lagrange", degree))
        self.bcs_Dirichlet = []
        self._A = self._B = self._B_adj = self._C = None

        if rank0():
            print("Degree of basis functions: ", self.degree, "\n")

        if getattr(self.parameter, "name", None) == "temperature":
            self.c = sound_speed_variable_gamma(self.mesh, parameter, degree=degree)
            self.T = self.parameter
            self.gamma = gamma_function(self.T)
            info("/\\ Temperature function is used for passive flame matrices.")
        else:
            self.c = parameter
            self.gamma = self.c.copy().fill(1.4)
            info("\\/ Speed of sound function is used for passive flame matrices.")

        terms = []          # (tag, i/Z) of the impedance boundaries (acoustic_matrices.py:68-97)
        bc_dofs = []
        for boundary in boundary_conditions:
            bc = boundary_conditions[boundary]
            if 'Neumann' in bc:
                info("- Neumann boundaries on boundary " + str(boundary))
            if 'Dirichlet' in bc:
                sel = np.flatnonzero(mesh.facet_tags == boundary)
                bc_dofs.append(self.V.facet_dofs.cpu().numpy()[sel].ravel())
                info("- Dirichlet boundary on boundary " + str(boundary))
            if 'Robin' in bc:
                R = bc['Robin']
                Z = (1 + R) / (1 - R)
                terms.append((boundary, 1j / Z))
                info("- Robin boundary on boundary " + str(boundary))
            if 'ChokedInlet' in bc:
                area, gint = fem.facet_integrals(mesh, boundary, np.real(self.gamma.x.array)[:mesh.n_nodes])
                gamma_inlet = gint / area
                Mach = bc['ChokedInlet']
                R = (1 - gamma_inlet * Mach / (1 + (gamma_inlet - 1) * Mach ** 2)) / \
                    (1 + gamma_inlet * Mach / (1 + (gamma_inlet - 1) * Mach ** 2))
                Z = (1 + R) / (1 - R)
                terms.append((boundary, 1j / Z))
                info("- Choked inlet boundary on boundary " + str(boundary))
            if 'ChokedOutlet' in bc:
                area, gint = fem.facet_integrals(mesh, boundary, np.real(self.gamma.x.array)[:mesh.n_nodes])
                gamma_outlet = gint / area
                Mach = bc['ChokedOutlet']
                R = (1 - 0.5 * (gamma_outlet - 1) * Mach) / (1 + 0.5 * (gamma_outlet - 1) * Mach)
                Z = (1 + R) / (1 - R)
                terms.append((boundary, 1j / Z))
                info("- Choked outlet boundary on boundary " + str(boundary))
        self.impedance_terms = terms

        info("- Passive matrices are assembling..")
        part = mesh.partition()
        if part is None:
            Vasm, c_asm = self.V, self.c
        else:
            # multi-GPU: this rank assembles the cells touching its rows on its sub-mesh
            from .dist import DistSpace
            dpart = mesh.dof_partition(degree)            # degree 2: vertex + edge dofs (dist.DofPartition)
            Vasm = fem.functionspace(part.local_mesh, ("Lagrange", degree))
            cvals = self.c.real_device()                  # restricted on the device, no host round trip
            if isinstance(self.c.function_space, fem.DG0Space):
                c_asm = fem.Function.from_device(fem.DG0Space(part.local_mesh), part.restrict_cell(cvals))
            else:
                c_asm = fem.Function.from_device(Vasm, part.restrict_nodal(cvals[:mesh.n_nodes]))
            g2l_loc = part.g2l if degree == 1 else dpart.g2l_old          # global dof -> dof of the local space
            bc_dofs = [g2l_loc[d][g2l_loc[d] >= 0] for d in bc_dofs]
        with phase("assembly_fields"):
            a_vals, c_vals = fem.assemble_AC(Vasm, c_asm)
        self.C_nobc_values = c_vals
        if bc_dofs:
            dofs = np.unique(np.concatenate(bc_dofs))
            self.bcs_Dirichlet = dofs
            self.C_nobc_values = c_vals.clone()
            fem.apply_dirichlet(Vasm, a_vals, dofs)
            fem.apply_dirichlet(Vasm, c_vals, dofs)
        info("- Matrix A is assembled.")
        b_vals = None
        if terms:
            with phase("assembly_B"):
                b_vals = fem.assemble_B(Vasm, c_asm, terms)
            info("- Matrix B is assembled.")
        if part is not None:
            self.V = DistSpace(dpart, Vasm)
            a_vals, c_vals = self.V.own_values(a_vals), self.V.own_values(c_vals)
            self.C_nobc_values = self.V.own_values(self.C_nobc_values)
            b_vals = self.V.own_values(b_vals) if b_vals is not None else None
        self.ops = OperatorSet(self.V, a_vals, c_vals, b_vals)
        self._A = Mat(self.ops, {"A": 1.0})
        self._C = Mat(self.ops, {"C": 1.0})
        if b_vals is not None:
            self._B = Mat(self.ops, {"B": 1.0})
            self._B_adj = Mat(self.ops, {"Bh": 1.0})
        info("- Matrix C is assembled.\n")

    @property
    def A(self):
        return self._A

    @property
    def B(self):
        return self._B

    @property
    def B_adj(self):
        return self._B_adj

    @property
    def C(self):
        return self._C
