"""Synthetic annular-combustor mesh "of the named shape" (SURVEY section 8d): a structured
cylinder shell r in [r0,r1], theta periodic, z in [z0,z1]; every hexahedron split into
6 Kuhn tetrahedra; 16 flame cell-tag blocks at theta_f = f*22.5 deg, z in [0, z_flame];
outlet facets (tag 11) on the z = z1 plane.  Geometry constants follow
numerical_examples/AnnularCombustor/Micca/fullAnnulus/params.py:7-18,43-48.
Built with torch ops on the chosen device; deterministic (no RNG)."""
import itertools
import math

import numpy as np
import torch


def annulus_grid(n_r, n_theta, n_z, device="cpu", r0=0.14, r1=0.21, z0=-0.09, z1=0.20, n_flames=16, z_flame=0.006,
                 flame_halfwidth=None):
    """Returns dict(x, cells, cell_tags, facets, facet_tags) of numpy arrays.
    Nodes: n_r x n_theta x n_z (theta periodic); node id = (iz*n_theta + it)*n_r + ir."""
    dev = torch.device(device)
    ir = torch.arange(n_r, device=dev)
    it = torch.arange(n_theta, device=dev)
    iz = torch.arange(n_z, device=dev)
    r = r0 + (r1 - r0) * ir.double() / (n_r - 1)
    th = 2 * math.pi * it.double() / n_theta
    z = z0 + (z1 - z0) * iz.double() / (n_z - 1)
    Z, T, R = torch.meshgrid(z, th, r, indexing="ij")
    x = torch.stack([R * torch.cos(T), R * torch.sin(T), Z], dim=-1).reshape(-1, 3)

    def nid(a, b, c):   # (ir, it, iz) -> node id, theta wraps
        return ((c * n_theta + (b % n_theta)) * n_r + a)

    cz, ct, cr = torch.meshgrid(torch.arange(n_z - 1, device=dev), it, torch.arange(n_r - 1, device=dev), indexing="ij")
    cr, ct, cz = cr.reshape(-1), ct.reshape(-1), cz.reshape(-1)
    tets = []
    for perm in itertools.permutations(range(3)):
        off = [0, 0, 0]
        verts = [nid(cr, ct, cz)]
        for ax in perm:
            off[ax] = 1
            verts.append(nid(cr + off[0], ct + off[1], cz + off[2]))
        tets.append(torch.stack(verts, dim=1))
    cells = torch.stack(tets, dim=1).reshape(-1, 4)          # 6 consecutive tets per hex
    # cell tags: flame f for hexes with z-centre in [0, z_flame] and theta within the flame block
    zc = z0 + (z1 - z0) * (cz.double() + 0.5) / (n_z - 1)
    tc = 2 * math.pi * (ct.double() + 0.5) / n_theta
    hw = flame_halfwidth if flame_halfwidth is not None else math.pi / n_flames / 3
    hw = max(hw, 1.01 * math.pi / n_theta)
    tag = torch.full_like(cr, n_flames)
    sector = 2 * math.pi / n_flames
    f = torch.round(tc / sector).long() % n_flames
    dth = torch.abs(torch.remainder(tc - f * sector + math.pi, 2 * math.pi) - math.pi)
    rc = r0 + (r1 - r0) * (cr.double() + 0.5) / (n_r - 1)
    rmid, rhalf = 0.5 * (r0 + r1), 0.25 * (r1 - r0)
    inflame = (zc >= 0.0) & (zc <= max(z_flame, (z1 - z0) / (n_z - 1))) & (dth <= hw) & (torch.abs(rc - rmid) <= rhalf)
    tag = torch.where(inflame, f, tag)
    cell_tags = tag.repeat_interleave(6)
    # outlet facets on z = z1: hexes of the top layer, two triangles each
    top = cz == (n_z - 2)
    a, b, c = cr[top], ct[top], cz[top]
    v001, v101, v011, v111 = nid(a, b, c + 1), nid(a + 1, b, c + 1), nid(a, b + 1, c + 1), nid(a + 1, b + 1, c + 1)
    facets = torch.cat([torch.stack([v001, v101, v111], 1), torch.stack([v001, v011, v111], 1)])
    facet_tags = torch.full((facets.shape[0],), 11, device=dev)
    return dict(x=x.cpu().numpy(), cells=cells.to(torch.int32).cpu().numpy(),
                cell_tags=cell_tags.to(torch.int32).cpu().numpy(), facets=facets.to(torch.int32).cpu().numpy(),
                facet_tags=facet_tags.to(torch.int32).cpu().numpy())


def annulus_sound_speed(x, cells):
    """DG0 c(z) three-zone profile from the cell midpoint (fullAnnulus/params.py:53-70)."""
    z = x[cells].mean(axis=1)[:, 2]
    gamma, r, l_cc, T_amb, T_a, T_b = 1.4, 287.0, 0.2, 300.0, 1521.0, 1200.0
    c = np.full(len(cells), math.sqrt(gamma * r * T_b))
    c[z < 0] = math.sqrt(gamma * r * T_amb)
    mid = (z > 0) & (z < l_cc)
    c[mid] = np.sqrt(gamma * r * ((T_b - T_a) * (z[mid] / l_cc) ** 2 + T_a))
    return c


def grid_for_dofs(n_dofs, degree=1):
    """(n_r, n_theta, n_z) giving roughly n_dofs dofs with near-cubic cells."""
    n_nodes = n_dofs if degree == 1 else n_dofs / 7.0     # P2 on Kuhn tets: ~7 dofs per node
    # aspect: radial 0.07 m, circumference ~1.1 m, height 0.29 m
    h = (0.07 * 1.1 * 0.29 / n_nodes) ** (1 / 3)
    n_r = max(3, int(round(0.07 / h)) + 1)
    n_theta = max(16, int(round(1.1 / h / 16)) * 16)
    n_z = max(4, int(round(n_nodes / (n_r * n_theta))))
    return n_r, n_theta, n_z


def slab_grid(nx, ny, L, h, thickness=None):
    """The reference's 2-D rectangle [0,L]x[0,h] (dolfinx_utils.RectangleSetup, used by
    numerical_examples/manufacturedSolution/manufacturedHelmholtz.py:12-15) as a one-cell-thick
    extruded slab of Kuhn tetrahedra: modes that do not vary in z are exactly those of the 2-D
    problem, so the P1-tetrahedron path covers BASELINE config 2.  Facet tags as in
    RectangleSetup: 1 left (x=0), 2 right (x=L), 3 bottom (y=0), 4 top (y=h)."""
    t = thickness if thickness is not None else min(L / nx, h / ny)
    xs, ys, zs = np.linspace(0, L, nx + 1), np.linspace(0, h, ny + 1), np.array([0.0, t])
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    x = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)

    def nid(i, j, k):
        return (k * (ny + 1) + j) * (nx + 1) + i
    I, J = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    I, J = I.ravel(), J.ravel()
    K = np.zeros_like(I)
    tets = []
    for perm in itertools.permutations(range(3)):
        off = [0, 0, 0]
        verts = [nid(I, J, K)]
        for ax in perm:
            off[ax] = 1
            verts.append(nid(I + off[0], J + off[1], K + off[2]))
        tets.append(np.stack(verts, axis=1))
    cells = np.stack(tets, axis=1).reshape(-1, 4)
    facets, tags = [], []

    def quad(a, b, c, d, tag):       # a-b-d-c around the quad, split along the Kuhn diagonal a-d
        facets.extend([np.stack([a, b, d], 1), np.stack([a, c, d], 1)])
        tags.extend([np.full(len(a), tag), np.full(len(a), tag)])
    jj = np.arange(ny)
    ii = np.arange(nx)
    z0, z1 = np.zeros_like(jj), np.ones_like(jj)
    quad(nid(0 * jj, jj, z0), nid(0 * jj, jj + 1, z0), nid(0 * jj, jj, z1), nid(0 * jj, jj + 1, z1), 1)
    quad(nid(0 * jj + nx, jj, z0), nid(0 * jj + nx, jj + 1, z0), nid(0 * jj + nx, jj, z1), nid(0 * jj + nx, jj + 1, z1), 2)
    z0, z1 = np.zeros_like(ii), np.ones_like(ii)
    quad(nid(ii, 0 * ii, z0), nid(ii + 1, 0 * ii, z0), nid(ii, 0 * ii, z1), nid(ii + 1, 0 * ii, z1), 3)
    quad(nid(ii, 0 * ii + ny, z0), nid(ii + 1, 0 * ii + ny, z0), nid(ii, 0 * ii + ny, z1), nid(ii + 1, 0 * ii + ny, z1), 4)
    return dict(x=x, cells=cells.astype(np.int32), cell_tags=np.zeros(len(cells), np.int32),
                facets=np.concatenate(facets).astype(np.int32), facet_tags=np.concatenate(tags).astype(np.int32))
