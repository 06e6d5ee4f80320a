"""Multi-GPU row partition: one process per GPU, rows (P1 dofs = mesh
nodes) split into contiguous chunks of a locality ordering, the way PETSc splits
ownership ranges across MPI ranks.  Each rank assembles the cells touching its rows on
its own sub-mesh, so assembly needs no exchange.  Per operator apply there is ONE
exchange step -- the packed halo of interface x entries (NCCL send/recv) -- and per
orthogonalisation pass one all-reduce of the Gram column.  The preconditioner is
block-Jacobi across GPUs: each rank runs its AMG cycle on its diagonal block.

The same code runs over gloo with the CPU test double (tests/test_dist_cpu.py).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from .backend import CsrMatrix
from .peer import HaloExchanger, PeerGroup, transport

c128 = torch.complex128


def _gloo():
    return dist.get_backend() == "gloo"


def all_gather_tensors(bufs, t):
    """dist.all_gather that also works for CUDA tensors over gloo (several ranks on one GPU in the tests)."""
    if t.is_cuda and _gloo():
        cb = [torch.zeros_like(b, device="cpu") for b in bufs]
        dist.all_gather(cb, t.cpu())
        for b, c in zip(bufs, cb):
            b.copy_(c)
    else:
        dist.all_gather(bufs, t)


def use_peer(t):
    """Peer-memory transport for this tensor's collectives?  (CUDA + more than one rank + not switched off.)"""
    return t.is_cuda and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and transport() == "peer"


def all_reduce_sum_(t):
    """In-place sum over ranks on the per-iteration path: one peer-memory kernel, or torch.distributed."""
    if use_peer(t):
        PeerGroup.get().allreduce_(t)
    else:
        dist.all_reduce(torch.view_as_real(t) if t.is_complex() else t)
    return t


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def morton_keys_np(x, bits=16):
    lo = x.min(axis=0)
    ext = max(float((x.max(axis=0) - lo).max()), 1e-300)
    q = np.minimum(((x - lo) / ext * (2 ** bits - 1)).astype(np.uint64), 2 ** bits - 1)
    key = np.zeros(len(x), np.uint64)
    for b in range(bits):
        for d in range(3):
            key |= ((q[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + d)
    return key


class Partition:
    """Node ownership + this rank's sub-mesh numbering + the halo exchange plan.  Integer work on the
    device the mesh lives on (torch; the CPU tests run the same code on CPU tensors); the host (numpy)
    views the drivers' host-side code needs -- l2g, g2l, cell_ids, ... -- are materialised on first use."""

    _HOST_VIEWS = ("owner", "pos", "cell_ids", "l2g", "g2l", "local_cells", "facet_ids", "local_facets")

    def __init__(self, x, cells, world, rank, ordering="input", facets=None):
        x = torch.as_tensor(x)
        dev = x.device
        cells = torch.as_tensor(cells).to(dev).long()
        n = int(x.shape[0])
        self.world, self.rank, self.n_global, self.device = world, rank, n, dev
        ar = torch.arange(n, device=dev)
        if ordering == "morton":
            from .amg import morton_order
            perm = morton_order(x.to(torch.float64))
        else:
            perm = ar
        owner = torch.empty(n, dtype=torch.int64, device=dev)
        owner[perm] = ar * world // n
        pos = torch.empty(n, dtype=torch.int64, device=dev)
        pos[perm] = ar                                   # rank of every node along the locality ordering
        own = torch.nonzero(owner == rank).reshape(-1)
        cell_mask = (owner[cells] == rank).any(dim=1)
        cell_ids = torch.nonzero(cell_mask).reshape(-1)
        lc = cells[cell_mask]
        nodes = torch.unique(lc)
        ghost = nodes[owner[nodes] != rank]
        ghost = ghost[torch.sort(owner[ghost] * n + ghost).indices]          # grouped by owner, ascending id
        self.n_own, self.n_ghost = int(own.numel()), int(ghost.numel())
        l2g = torch.cat([own, ghost])
        g2l = torch.full((n,), -1, dtype=torch.int64, device=dev)
        g2l[l2g] = torch.arange(l2g.numel(), device=dev)
        self._d = {"owner": owner, "pos": pos, "cell_ids": cell_ids, "l2g": l2g, "g2l": g2l,
                   "local_cells": g2l[lc].to(torch.int32).contiguous()}
        if facets is not None and len(facets):
            fac = torch.as_tensor(facets).to(dev).long()
            fmask = (owner[fac] == rank).any(dim=1)
            self._d["facet_ids"] = torch.nonzero(fmask).reshape(-1)
            self._d["local_facets"] = g2l[fac[fmask]].to(torch.int32).contiguous()
            assert int(self._d["local_facets"].min()) >= 0 if self._d["local_facets"].numel() else True
        else:
            self._d["facet_ids"] = torch.zeros(0, dtype=torch.int64, device=dev)
            self._d["local_facets"] = torch.zeros((0, 3), dtype=torch.int32, device=dev)
        self._h = {}
        self._halo_plan(owner, ghost, g2l)

    def _halo_plan(self, owner, ghost, g2l):
        """Who needs which of my rows: every rank learns every ghost list (global ids, grouped by owner) and
        picks the ids it owns, in the requester's order."""
        world, rank, dev = self.world, self.rank, owner.device
        self.ghost_owner_counts = torch.bincount(owner[ghost], minlength=world).cpu().numpy().astype(np.int64)
        if world > 1:
            counts = _all_gather_rows(torch.tensor([self.n_ghost, self.n_own], dtype=torch.int64, device=dev).view(1, 2), world).cpu().numpy()
            allg = _all_gather_rows(ghost, world)
        else:
            counts, allg = np.array([[self.n_ghost, self.n_own]]), ghost
        self.own_counts = counts[:, 1].astype(np.int64)
        off = np.concatenate([[0], np.cumsum(counts[:, 0])])
        send_idx, send_counts = [], np.zeros(world, np.int64)
        for q in range(world):
            if q == rank:
                continue
            gq = allg[int(off[q]):int(off[q + 1])]
            mine = gq[owner[gq] == rank]                              # in q's ghost order
            send_idx.append(g2l[mine])
            send_counts[q] = int(mine.numel())
        self.send_counts = send_counts
        self.send_idx = torch.cat(send_idx) if send_idx else torch.zeros(0, dtype=torch.int64, device=dev)
        self._dev = {}
        self._all_ids = None

    def dev(self, name):
        """Device tensor of one of the index arrays (owner, pos, cell_ids, l2g, g2l, local_cells, ...)."""
        return self._d[name]

    def __getattr__(self, name):
        if name in type(self)._HOST_VIEWS and name in self.__dict__.get("_d", {}):
            h = self.__dict__["_h"]
            if name not in h:
                h[name] = self.__dict__["_d"][name].cpu().numpy()
            return h[name]
        raise AttributeError(name)

    @property
    def send_idx_h(self):
        return self.send_idx.cpu().numpy()

    @property
    def n_loc(self):
        return self.n_own + self.n_ghost

    def neighbours(self):
        return [q for q in range(self.world) if q != self.rank and (self.send_counts[q] or self.ghost_owner_counts[q])]

    def exchanger(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = HaloExchanger(self.world, self.rank, self.n_own, self.send_idx.to(device), self.send_counts,
                                           self.ghost_owner_counts)
        return self._dev[key]

    def exchange(self, x_loc):
        """Fill the ghost tail of x_loc (length n_loc) from the owners: one peer-memory kernel, or one
        gather + one grouped send/recv (peer.HaloExchanger)."""
        if self.world == 1:
            return x_loc
        return self.exchanger(x_loc.device).exchange(x_loc)

    def all_reduce_max(self, t):
        """In-place maximum over ranks of a real device scalar / tensor."""
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    # ---- global <-> distributed host helpers -------------------------------------------
    def restrict_nodal(self, arr):
        """Values at this rank's local (owned + ghost) nodes; device tensor in -> device tensor out."""
        if torch.is_tensor(arr):
            return arr[self._d["l2g"].to(arr.device)].contiguous()
        return np.asarray(arr)[self.l2g]

    def restrict_cell(self, arr):
        if torch.is_tensor(arr):
            return arr[self._d["cell_ids"].to(arr.device)].contiguous()
        return np.asarray(arr)[self.cell_ids]

    def gather_global(self, x_own):
        """All ranks get the global vector (host numpy) from the owned device pieces."""
        if self.world == 1:
            out = np.zeros(self.n_global, complex)
            out[self.l2g[:self.n_own]] = x_own.cpu().numpy()
            return out
        if getattr(self, "_all_ids", None) is None:           # static per partition; tensor collective, no pickling
            self._all_ids = _all_gather_rows(self._d["l2g"][:self.n_own].contiguous(), self.world).cpu().numpy()
        vals = _all_gather_rows(x_own[:self.n_own].contiguous(), self.world).cpu().numpy()
        out = np.zeros(self.n_global, complex)
        out[self._all_ids] = vals
        return out


class DofPartition(Partition):
    """Row partition of a degree-2 space on top of the node partition: vertex dofs follow their node, an
    edge dof belongs to the lower-ranked owner of its two vertices (that rank owns a vertex of the edge, so
    its sub-mesh holds every cell around the edge and the row it assembles is complete).  The rank's local
    space numbers [vertices | edges]; `dof_perm` reorders it to [owned | ghosts grouped by owner], the layout
    every distributed vector and matrix uses.  Same interface as Partition (n_own, l2g, exchange, ...)."""

    _HOST_VIEWS = Partition._HOST_VIEWS + ("dof_perm", "dof_inv_perm", "g2l_old")

    def __init__(self, node_part: Partition, Vglob, Vloc):
        self.node_part = node_part
        self.world, self.rank, self.device = node_part.world, node_part.rank, node_part.device
        dev = self.device
        nn = node_part.n_global
        ge = Vglob.edges.to(dev)                                          # (n_edges, 2) global vertex ids, sorted by key
        gkeys = ge[:, 0] * nn + ge[:, 1]
        self.n_global = nn + int(ge.shape[0])
        own_node = node_part.dev("owner")
        owner = torch.cat([own_node, torch.minimum(own_node[ge[:, 0]], own_node[ge[:, 1]])])
        l2g_node = node_part.dev("l2g")
        le = Vloc.edges.to(dev)                                           # local vertex pairs of the local edges
        ga, gb = l2g_node[le[:, 0]], l2g_node[le[:, 1]]
        lkeys = torch.minimum(ga, gb) * nn + torch.maximum(ga, gb)
        gidx = torch.searchsorted(gkeys, lkeys)
        assert bool((gkeys[gidx.clamp_max(gkeys.numel() - 1)] == lkeys).all()), "a local edge is missing from the global mesh"
        gl_old = torch.cat([l2g_node, nn + gidx])                         # global dof of every local dof (local order)
        mine = owner[gl_old] == self.rank
        own_old = torch.nonzero(mine).reshape(-1)
        own_old = own_old[torch.sort(gl_old[own_old]).indices]            # owned dofs by ascending global id
        ghost_old = torch.nonzero(~mine).reshape(-1)
        ghost_old = ghost_old[torch.sort(owner[gl_old[ghost_old]] * self.n_global + gl_old[ghost_old]).indices]
        perm = torch.cat([own_old, ghost_old])                            # new local index -> old local index
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(perm.numel(), device=dev)
        self.n_own, self.n_ghost = int(own_old.numel()), int(ghost_old.numel())
        l2g = gl_old[perm]
        g2l = torch.full((self.n_global,), -1, dtype=torch.int64, device=dev)
        g2l[l2g] = torch.arange(l2g.numel(), device=dev)
        g2l_old = torch.full((self.n_global,), -1, dtype=torch.int64, device=dev)
        g2l_old[gl_old] = torch.arange(gl_old.numel(), device=dev)
        self._d = {"owner": owner, "l2g": l2g, "g2l": g2l, "g2l_old": g2l_old, "dof_perm": perm, "dof_inv_perm": inv,
                   "cell_ids": node_part.dev("cell_ids")}
        self._h = {}
        self._halo_plan(owner, l2g[self.n_own:], g2l)
        self.local_mesh = node_part.local_mesh

    def restrict_nodal(self, arr):
        return self.node_part.restrict_nodal(arr)

    def restrict_cell(self, arr):
        return self.node_part.restrict_cell(arr)


class DistMatrix:
    """Owned rows of a distributed matrix: local CSR/SELL operator (n_own x n_loc)."""
    is_dist = True

    def __init__(self, part: Partition, local, be_local, space=None):
        self.part, self.op = part, local
        self.n_rows = self.n_cols = part.n_own
        self.nnz = local.nnz
        self._xbuf = None
        self.be_local = be_local
        self.space = space

    @property
    def shape(self):
        return (self.n_rows, self.n_cols)

    def xbuf(self):
        """Input vector [owned | ghosts] of the local kernel; shared by all matrices of one space (it lives
        in the space's peer arena when the halo travels over peer memory)."""
        if self._xbuf is None:
            self._xbuf = self.space.xbuf() if self.space is not None else self.be_local.zeros(self.part.n_loc)
        return self._xbuf

    def to_scipy_local(self):
        return self.op.to_scipy()


class DistBackend:
    """Wraps the per-GPU backend: vectors hold owned entries; reductions are all-reduced,
    SpMV on a DistMatrix does the halo exchange first.  Everything else is local."""
    name = "dist"

    def __init__(self, local, part: Partition):
        self.local, self.part = local, part
        self.device = local.device
        self.supports_sell = getattr(local, "supports_sell", False)
        # the Hessenberg column can return through the pinned copy stream when no collective needs the host
        self.supports_pipelining = getattr(local, "supports_pipelining", False) and transport() == "peer"

    def __getattr__(self, name):
        return getattr(self.local, name)

    def spmv(self, M, x, y, alpha=1.0, beta=None, y0=None, lanes=None):
        if getattr(M, "is_dist", False):
            xl = M.xbuf()
            xl[:self.part.n_own].copy_(x[:self.part.n_own])
            self.part.exchange(xl)
            return self.local.spmv(M.op, xl, y, alpha=alpha, beta=beta, y0=y0)
        return self.local.spmv(M, x, y, alpha=alpha, beta=beta, y0=y0, lanes=lanes)

    def multi_dot(self, V, k, w, out, conj=True):
        self.local.multi_dot(V, k, w, out, conj)
        if self.part.world > 1:
            all_reduce_sum_(out[:k])
        return out

    def multi_axpy(self, V, k, h, w, hacc=None, nrm2=None):
        self.local.multi_axpy(V, k, h, w, hacc=hacc, nrm2=nrm2)
        if nrm2 is not None and self.part.world > 1:
            all_reduce_sum_(nrm2[:1])
        return w

    def lowrank_dots(self, lr, x, t):
        self.local.lowrank_dots(lr, x, t)
        if self.part.world > 1:
            all_reduce_sum_(t[:max(lr.r, 1)])
        return t


class DistSpace:
    """What OperatorSet / AMG need from a function space, for the owned rows of a rank:
    pattern of the owned rows (columns in local numbering [owned | ghosts]), the diagonal block,
    coordinates.  A degree-2 partition carries the permutation from the local space's [vertices | edges]
    numbering to that layout; for degree 1 the sub-mesh numbering already has it."""

    def __init__(self, part: Partition, Vloc):
        self.part, self.Vloc = part, Vloc
        self.local_be = Vloc.be
        self.be = DistBackend(Vloc.be, part)
        self.n = part.n_own
        self.n_global = part.n_global
        self.degree = Vloc.degree
        ip, ix = Vloc.pattern()
        dev = ip.device
        n_own = part.n_own
        if isinstance(part, DofPartition):
            perm, inv = part.dev("dof_perm").to(dev), part.dev("dof_inv_perm").to(dev)
            ipl = ip.long()
            old_rows = perm[:n_own]
            cnt = (ipl[1:] - ipl[:-1])[old_rows]
            ptr = torch.zeros(n_own + 1, dtype=torch.int64, device=dev)
            ptr[1:] = torch.cumsum(cnt, 0)
            total = int(ptr[-1])
            pos = torch.repeat_interleave(ipl[old_rows] - ptr[:-1], cnt, output_size=total) + torch.arange(total, device=dev)
            rows = torch.repeat_interleave(torch.arange(n_own, device=dev), cnt, output_size=total)
            cols = inv[ix[pos].long()]
            order = torch.sort(rows * int(perm.numel()) + cols).indices          # columns of a row ascending again
            self._value_perm = pos[order]
            self._pattern = (ptr.to(torch.int32).contiguous(), cols[order].to(torch.int32).contiguous())
            self.nnz_own = total
            self.dof_coords = Vloc.dof_coords[old_rows].contiguous()
        else:
            self._value_perm = None
            self.nnz_own = int(ip[n_own])
            self._pattern = (ip[:n_own + 1].contiguous(), ix[:self.nnz_own].contiguous())
            self.dof_coords = Vloc.dof_coords[:n_own].contiguous()
        mask = self._pattern[1] < n_own
        self.diag_sel = torch.nonzero(mask).reshape(-1)
        rows = torch.repeat_interleave(torch.arange(n_own, device=dev),
                                       (self._pattern[0][1:] - self._pattern[0][:-1]).long())
        counts = torch.bincount(rows[mask], minlength=n_own)
        dptr = torch.zeros(n_own + 1, dtype=torch.int64, device=dev)
        dptr[1:] = torch.cumsum(counts, 0)
        self._diag_pattern = (dptr.to(torch.int32).contiguous(), self._pattern[1][mask].contiguous())
        self._xbuf = None
        self._arena = None

    def xbuf(self):
        if self._xbuf is None:
            n_loc = self.part.n_loc
            if use_peer(self._pattern[0]):
                grp = PeerGroup.get()
                _, (mx,) = grp._all_min_max([n_loc])
                self._arena = grp.lease(mx * 16 + 4096)
                self._xbuf = self._arena.take(mx, c128, n_loc)
            else:
                self._xbuf = self.local_be.zeros(n_loc)
        return self._xbuf

    def __del__(self):
        if getattr(self, "_arena", None) is not None:
            self._arena.group.release(self._arena)

    def pattern(self):
        return self._pattern

    def own_values(self, full_values):
        """Values of the owned rows (this space's pattern order) from values on the local space's pattern."""
        if self._value_perm is not None:
            return full_values[self._value_perm].contiguous()
        return full_values[:self.nnz_own].contiguous()

    def own_vector(self, local_vector):
        """Owned entries (distributed layout) of a vector in the local space's numbering."""
        if isinstance(self.part, DofPartition):
            return local_vector[self.part.dev("dof_perm").to(local_vector.device)[:self.part.n_own]].contiguous()
        return local_vector[:self.part.n_own].contiguous()

    def matrix(self, values):
        local = CsrMatrix(self.part.n_own, self.part.n_loc, self._pattern[0], self._pattern[1], values)
        return DistMatrix(self.part, local, self.local_be, self)

    def matrix_sell(self, values):
        """The owned rows in SELL-32 (the fine-level SpMV format of the single-GPU path) when the block is
        large enough to pay; None otherwise."""
        be = self.local_be
        if not getattr(be, "supports_sell", False) or self.part.n_own < 250000:
            return None
        from .sell import SellMatrix, SellPattern
        if getattr(self, "_sellp", None) is None:
            self._sellp = SellPattern(be, self._pattern[0], self._pattern[1], self.part.n_own, self.part.n_loc)
            self._sell_vals = be.empty(max(self._sellp.total, 1))
        local = SellMatrix(self._sellp, self._sellp.values_from_csr(values, out=self._sell_vals))
        return DistMatrix(self.part, local, be, self)

    def diag_matrix(self, values):
        return CsrMatrix(self.part.n_own, self.part.n_own, self._diag_pattern[0], self._diag_pattern[1],
                         values[self.diag_sel].contiguous())


class CoarseCorrection:
    """Global coarse space for the multi-GPU preconditioner: ~nc aggregates = equal runs of
    the global locality (Morton) ordering, piecewise-constant prolongator P0.  The dense
    coarse operators P0^T X P0 (X = A, B, C) are summed over ranks once per mesh; per shift
    they are combined and inverted (replicated); per application: local restriction, one
    all-reduce of nc complex numbers, dense GEMV, local prolongation.  Block-Jacobi AMG
    alone misses the global low (indefinite) modes: measured on the 35 k-DoF annulus,
    GMRES iterations 255 -> 81 (2 ranks) and 515 -> 114 (8 ranks) with nc = 1000."""

    def __init__(self, space: "DistSpace", base, nc=1000):
        part, be = space.part, space.local_be
        self.part, self.be = part, be
        dev = space._pattern[0].device
        nc = int(min(nc, max(part.n_global // 8, 1)))
        self.nc = nc
        agg_loc = (part.dev("pos")[part.dev("l2g")] * nc // part.n_global).to(dev)
        n_own = part.n_own
        ip, ix = space._pattern
        rows = torch.repeat_interleave(torch.arange(n_own, device=dev), (ip[1:] - ip[:-1]).long())
        key = agg_loc[rows] * nc + agg_loc[ix.long()]
        self._key = key
        # P0 (n_own x nc) and R0 = P0^T (nc x n_own) as real CSR with unit entries
        one = torch.ones(n_own, dtype=torch.float64, device=dev)
        self.P0 = CsrMatrix(n_own, nc, torch.arange(n_own + 1, device=dev, dtype=torch.int32),
                            agg_loc[:n_own].to(torch.int32).contiguous(), one)
        order = torch.sort(agg_loc[:n_own], stable=True).indices
        counts = torch.bincount(agg_loc[:n_own], minlength=nc)
        rptr = torch.zeros(nc + 1, dtype=torch.int64, device=dev)
        rptr[1:] = torch.cumsum(counts, 0)
        self.R0 = CsrMatrix(nc, n_own, rptr.to(torch.int32).contiguous(), order.to(torch.int32).contiguous(), one)
        self.dense = {}
        for name in ("A", "B", "C", "Bh"):
            v = base.get(name)
            if v is None:
                continue
            self.dense[name] = self._galerkin(v.to(c128))
        self.t = be.zeros(nc)
        self.y = be.zeros(nc)
        self.inv = None

    def _galerkin(self, vals):
        """sum_{i in I, j in J} a_ij over owned rows, all-reduced; deterministic (sort + segmented sum)."""
        nc = self.nc
        sp_ = torch.sparse_coo_tensor(self._key.view(1, -1), vals, size=(nc * nc,)).coalesce()
        dense = torch.zeros(nc * nc, dtype=c128, device=vals.device)
        dense[sp_.indices()[0]] = sp_.values()
        if self.part.world > 1:
            dist.all_reduce(torch.view_as_real(dense))
        return dense.view(nc, nc)

    def set_shift(self, terms):
        M = torch.zeros(self.nc, self.nc, dtype=c128, device=self.t.device)
        for k, v in terms.items():
            if v != 0:
                M += complex(v) * self.dense[k]
        self.inv = M.t().contiguous()          # column-major view of M for the in-place inverse kernel
        self.info = self.be.dense_inverse(self.inv)
        if int(self.info[0]) != 0:
            raise RuntimeError("multi-GPU coarse correction: the coarse operator is singular")

    def apply(self, v, x):
        """x = P0 A0^-1 P0^T v  (v, x: owned entries)."""
        be = self.be
        be.spmv(self.R0, v, self.t)
        if self.part.world > 1:
            all_reduce_sum_(self.t)
        be.dense_gemv(self.inv, self.t, self.y)
        be.spmv(self.P0, self.y, x)
        return x


# ---------------------------------------------------------------------------------------------
# Row-distributed multigrid cycle (HX_DIST_HIERARCHY=1; default off until measured on GPUs).
#
# The two-level Schwarz preconditioner above loses the couplings across ranks on every level
# and needs 1.5-1.8x the iterations of the single-GPU hierarchy.  Here every rank builds the
# *same* global hierarchy once per mesh (replicated set-up: the assembled rows are all-gathered,
# 32 B per nonzero), the fine level -- where the bytes are -- is smoothed on the owned rows with
# one halo exchange per sweep, the restricted residual is all-reduced, and the levels below run
# replicated.  The cycle is then the single-GPU cycle up to summation order, so the iteration
# count does not depend on the number of ranks.
# ---------------------------------------------------------------------------------------------
def _all_gather_rows(t, world):
    """Concatenation over ranks of a 1-D/2-D tensor whose first dimension differs per rank."""
    if world == 1:
        return t
    cplx = t.is_complex()
    src = torch.view_as_real(t.contiguous()) if cplx else t.contiguous()
    n = torch.tensor([src.shape[0]], dtype=torch.int64, device=src.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    all_gather_tensors(counts, n)
    counts = [int(c) for c in counts]
    pad = torch.zeros((max(counts),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    pad[:src.shape[0]] = src
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    all_gather_tensors(bufs, pad)
    out = torch.cat([b[:c] for b, c in zip(bufs, counts)])
    return torch.view_as_complex(out) if cplx else out


def _csr_rows(M: CsrMatrix, rows):
    """Sub-matrix of the given rows (all columns)."""
    ip = M.indptr.long()
    cnt = (ip[1:] - ip[:-1])[rows]
    ptr = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=rows.device)
    ptr[1:] = torch.cumsum(cnt, 0)
    total = int(ptr[-1])
    idx = torch.repeat_interleave(ip[rows] - ptr[:-1], cnt, output_size=total) + torch.arange(total, device=rows.device)
    return CsrMatrix(int(rows.numel()), M.n_cols, ptr.to(torch.int32).contiguous(), M.indices[idx].contiguous(),
                     M.values[idx].contiguous())


def _csr_transpose(M: CsrMatrix):
    ip = M.indptr.long()
    rows = torch.repeat_interleave(torch.arange(M.n_rows, device=ip.device), ip[1:] - ip[:-1], output_size=M.nnz)
    cols = M.indices.long()
    order = torch.sort(cols * M.n_rows + rows).indices
    ptr = torch.zeros(M.n_cols + 1, dtype=torch.int64, device=ip.device)
    ptr[1:] = torch.cumsum(torch.bincount(cols, minlength=M.n_cols), 0)
    return CsrMatrix(M.n_cols, M.n_rows, ptr.to(torch.int32).contiguous(), rows[order].to(torch.int32).contiguous(),
                     M.values[order].contiguous())


class HaloPlan:
    """Exchange plan of one distributed level: `own` global ids (ascending) first, then the `ghost` ids
    grouped by owner.  Same exchange step as Partition.exchange (peer.HaloExchanger)."""

    def __init__(self, world, rank, own, ghost, owner):
        dev = own.device
        self.world, self.rank = world, rank
        self.n_own, self.n_ghost = int(own.numel()), int(ghost.numel())
        gown = owner[ghost]
        order = torch.sort(gown * (int(owner.numel()) + 1) + ghost).indices
        ghost = ghost[order]
        self.l2g = torch.cat([own, ghost])
        self.ghost_owner_counts = torch.bincount(owner[ghost], minlength=world).cpu().numpy()
        # every rank learns every ghost list and picks the ids it owns, in the requester's order
        counts = _all_gather_rows(torch.tensor([self.n_ghost], dtype=torch.int64, device=dev), world).cpu().numpy()
        allg = _all_gather_rows(ghost, world)
        off = np.concatenate([[0], np.cumsum(counts)])
        send_idx, send_counts = [], np.zeros(world, np.int64)
        for q in range(world):
            if q == rank:
                continue
            gq = allg[off[q]:off[q + 1]]
            mine = gq[owner[gq] == rank]
            send_idx.append(torch.searchsorted(own, mine))
            send_counts[q] = int(mine.numel())
        self.send_counts = send_counts
        self.send_idx = torch.cat(send_idx) if send_idx else torch.zeros(0, dtype=torch.int64, device=dev)
        self.ex = HaloExchanger(world, rank, self.n_own, self.send_idx, send_counts, self.ghost_owner_counts)

    @property
    def n_loc(self):
        return self.n_own + self.n_ghost

    def exchange(self, x_loc):
        if self.world == 1:
            return x_loc
        return self.ex.exchange(x_loc)


def _local_rows(M: CsrMatrix, own, g2l, n_loc):
    """Rows `own` of the global matrix M with the columns renumbered by g2l; also the positions of
    their values in M.values (to refresh the values at every shift)."""
    ip = M.indptr.long()
    cnt = (ip[1:] - ip[:-1])[own]
    ptr = torch.zeros(own.numel() + 1, dtype=torch.int64, device=own.device)
    ptr[1:] = torch.cumsum(cnt, 0)
    total = int(ptr[-1])
    pos = torch.repeat_interleave(ip[own] - ptr[:-1], cnt, output_size=total) + torch.arange(total, device=own.device)
    cols = g2l[M.indices[pos].long()]
    assert int(cols.min()) >= 0 if total else True, "a column of an owned row is neither owned nor in the halo"
    # keep the columns of each row sorted (the CSR kernels and the diagonal search expect it)
    rows = torch.repeat_interleave(torch.arange(own.numel(), device=own.device), cnt, output_size=total)
    order = torch.sort(rows * n_loc + cols).indices
    local = CsrMatrix(int(own.numel()), n_loc, ptr.to(torch.int32).contiguous(), cols[order].to(torch.int32).contiguous(),
                      M.values[pos[order]].contiguous() if M.values is not None else None)
    return local, pos[order]


class _DLevel:
    pass


SELL_MIN_ROWS = 250000      # owned rows of a distributed level from which its block is stored as SELL-32


class DistHierarchy:
    """Replicated set-up, row-distributed cycle.  Levels with at least `min_rows` rows per rank are
    distributed (smoothing, residual, restriction and prolongation on the owned rows, one halo exchange
    per operator application); the rest of the hierarchy runs replicated on the all-reduced residual."""

    def __init__(self, space: "DistSpace", base, min_rows=None, **amg_options):
        import os
        from .amg import AMG
        part, be = space.part, space.local_be
        self.part, self.be, self.space = part, be, space
        world, rank = part.world, part.rank
        if min_rows is None:
            min_rows = int(os.environ.get("HX_DIST_MIN_ROWS", "20000"))
        ip, ix = space._pattern
        dev = ip.device
        n_own, ng = part.n_own, part.n_global
        l2g = part.dev("l2g").to(dev)
        rows_loc = torch.repeat_interleave(torch.arange(n_own, device=dev), (ip[1:] - ip[:-1]).long())
        key = _all_gather_rows(l2g[rows_loc] * ng + l2g[ix.long()], world)
        order = torch.sort(key).indices
        key = key[order]
        grow = torch.div(key, ng, rounding_mode="floor")
        gptr = torch.zeros(ng + 1, dtype=torch.int64, device=dev)
        gptr[1:] = torch.cumsum(torch.bincount(grow, minlength=ng), 0)
        gptr, gidx = gptr.to(torch.int32).contiguous(), (key - grow * ng).to(torch.int32).contiguous()

        def glob(v):
            return CsrMatrix(ng, ng, gptr, gidx, _all_gather_rows(v, world)[order].contiguous())
        coords = torch.zeros(ng, 3, dtype=space.dof_coords.dtype, device=dev)
        coords[_all_gather_rows(l2g[:n_own], world)] = _all_gather_rows(space.dof_coords, world)
        B = glob(base["B"]) if base.get("B") is not None else None
        # the replicated hierarchy only feeds the owned rows of the distributed levels and runs the small levels:
        # no SELL copies of the global operators (the distributed blocks build their own)
        opts = dict(amg_options)
        opts.setdefault("sell_min_rows", 1 << 62)
        self.mg = mg = AMG(be, glob(base["A"]), glob(base["C"]), B, coords, **opts)
        mg.use_graph = False
        self.single, self.wdtype, self.nu = mg.single, mg.wdtype, mg.nu
        nlev = len(mg.levels)
        if nlev < 2:
            raise ValueError("the distributed cycle needs at least two levels")
        # ---- ownership per level: a coarse dof belongs to the lowest rank owning one of its members
        owners = [part.dev("owner").to(dev)]
        n_dist = 1
        for l in range(1, nlev - 1):
            if mg.levels[l].n < min_rows * world:
                break
            prev, agg = owners[-1], mg.levels[l - 1].agg
            own_l = torch.full((mg.levels[l].n,), world, dtype=torch.int64, device=dev)
            own_l.scatter_reduce_(0, agg, prev, reduce="amin", include_self=True)
            owners.append(own_l)
            n_dist = l + 1
        self.n_dist = n_dist
        wd = self.wdtype
        self.dl = []
        for l in range(n_dist):
            D = _DLevel()
            L = mg.levels[l]
            D.own = torch.nonzero(owners[l] == rank).reshape(-1)
            # halo of level l: columns of the owned rows of M_l and of the prolongator rows of the level above
            D.need = [_csr_rows(L.pattern, D.own).indices.long()]
            if l > 0:
                D.need.append(_csr_rows(mg.levels[l - 1].P, self.dl[l - 1].own).indices.long())
            self.dl.append(D)
        # restriction rows of the next distributed level read level-l residuals: extend level l's halo
        for l in range(n_dist - 1):
            Rn = _csr_rows(mg.levels[l].R, self.dl[l + 1].own)
            self.dl[l].need.append(Rn.indices.long())
        for l, D in enumerate(self.dl):
            L = mg.levels[l]
            allc = torch.unique(torch.cat(D.need))
            ghost = allc[owners[l][allc] != rank]
            D.halo = HaloPlan(world, rank, D.own, ghost, owners[l])
            g2l = torch.full((L.n,), -1, dtype=torch.int64, device=dev)
            g2l[D.halo.l2g] = torch.arange(D.halo.n_loc, device=dev)
            D.g2l = g2l
            D.M_pat, D.M_pos = _local_rows(L.pattern, D.own, g2l, D.halo.n_loc)
            n_o, n_l = D.halo.n_own, D.halo.n_loc
            D.b = be.zeros(n_o, dtype=wd)
            D.dinv = be.zeros(n_o)
            D.sellp = None
            if getattr(be, "supports_sell", False) and n_o >= SELL_MIN_ROWS:
                from .sell import SellPattern
                D.sellp = SellPattern(be, D.M_pat.indptr, D.M_pat.indices, n_o, n_l)
                D.sell_vals = be.empty(max(D.sellp.total, 1), dtype=wd)
            if mg.w_from is not None:
                D.xs = be.zeros(n_l, dtype=wd)
                D.bs = be.zeros(n_o, dtype=wd)
            del D.need
        # the vectors whose ghost tails the neighbours fill: in a peer arena when the halo travels over
        # peer memory (sizes = maximum over ranks, so every rank computes the same offsets)
        self._arena = None
        n_locs = [D.halo.n_loc for D in self.dl]
        if use_peer(ip):
            grp = PeerGroup.get()
            _, mx = grp._all_min_max(n_locs)
            item = torch.empty(0, dtype=wd).element_size()
            self._arena = grp.lease(sum(3 * (m_ * item + 256) for m_ in mx) + 4096)
            for D, m_, n_l in zip(self.dl, mx, n_locs):
                D.r, D.xa, D.xb = (self._arena.take(m_, wd, n_l) for _ in range(3))
        else:
            for D, n_l in zip(self.dl, n_locs):
                D.r, D.xa, D.xb = be.zeros(n_l, dtype=wd), be.zeros(n_l, dtype=wd), be.zeros(n_l, dtype=wd)
        self._graph = None
        self.use_graph = (self._arena is not None and getattr(be, "supports_graphs", False)
                          and os.environ.get("HX_AMG_GRAPH", "1") != "0")
        for l, D in enumerate(self.dl):
            L = mg.levels[l]
            if l + 1 < n_dist:
                Dn = self.dl[l + 1]
                D.P, _ = _local_rows(L.P, D.own, Dn.g2l, Dn.halo.n_loc)           # own_l x loc_{l+1}
                D.R, _ = _local_rows(L.R, Dn.own, D.g2l, D.halo.n_loc)            # own_{l+1} x loc_l
            else:
                D.P = _csr_rows(L.P, D.own)                                       # own_l x n_{l+1} (replicated below)
                D.R = _csr_transpose(D.P)                                         # partial sums, all-reduced
            D.P, D.R = self._transfer_op(D.P), self._transfer_op(D.R)
        assert torch.equal(self.dl[0].own, l2g[:n_own]), "level-0 ownership differs from the partition"

    def _transfer_op(self, M):
        """SELL-32 copy of a float32 transfer block with many rows (as amg.AMG._transfer_op on one GPU)."""
        if (not getattr(self.be, "supports_sell", False) or M.n_rows < SELL_MIN_ROWS or M.values is None
                or M.values.dtype != torch.float32 or os.environ.get("HX_AMG_SELL_TRANSFER", "1") == "0"):
            return M
        from .sell import SellMatrix, SellPattern
        p = SellPattern(self.be, M.indptr, M.indices, M.n_rows, M.n_cols)
        return SellMatrix(p, p.values_from_csr(M.values))

    def set_fine(self, values_own=None):
        """Refresh the owned rows of every distributed level from the replicated level operators
        (call after mg.set_shift)."""
        be = self.be
        for l, D in enumerate(self.dl):
            vals = self.mg.levels[l].M.values[D.M_pos].contiguous()
            M = D.M_pat.with_values(vals)
            be.diag_inv(M, D.dinv)
            D.dinv_w = D.dinv.to(self.wdtype) if self.single else D.dinv
            if D.sellp is not None:
                from .sell import SellMatrix
                D.M = SellMatrix(D.sellp, D.sellp.values_from_csr(vals, out=D.sell_vals), variant=2 if (self.single and l == 0) else 0)
            else:
                D.M = M.with_values(vals.to(self.wdtype)) if self.single else M
            D.omegas = self.mg.levels[l].omegas       # constant, or the level's Chebyshev roots of this shift
        self._graph = None                            # the captured cycle points at the previous shift's operators

    def _sweeps(self, D, first_zero):
        be = self.be
        k = 0
        if first_zero:
            be.jacobi_sweep(D.M, D.dinv_w, D.b, None, D.xa, D.omegas[0])
            k = 1
        for s in range(k, len(D.omegas)):
            D.halo.exchange(D.xa)
            be.jacobi_sweep(D.M, D.dinv_w, D.b, D.xa, D.xb, D.omegas[s])
            D.xa, D.xb = D.xb, D.xa

    def _cycle(self, l):
        """V-cycle from distributed level l for the right-hand side in dl[l].b; result in dl[l].xa[:n_own]."""
        be, mg, D = self.be, self.mg, self.dl[l]
        self._sweeps(D, first_zero=True)
        D.halo.exchange(D.xa)
        be.spmv(D.M, D.xa, D.r, alpha=-1.0, beta=1.0, y0=D.b)                     # r = b - M x (owned rows)
        if l + 1 < self.n_dist:
            Dn = self.dl[l + 1]
            D.halo.exchange(D.r)
            be.spmv(D.R, D.r, Dn.b)
            self._cycle(l + 1)
            if mg.w_from is not None and mg.w_from <= l + 1 <= mg.w_to:
                # W-cycle: second visit of the (distributed) coarse level on its residual
                Dn.xs.copy_(Dn.xa)
                Dn.bs.copy_(Dn.b)
                Dn.halo.exchange(Dn.xa)
                be.spmv(Dn.M, Dn.xa, Dn.b, alpha=-1.0, beta=1.0, y0=Dn.bs)
                self._cycle(l + 1)
                Dn.xa[:Dn.halo.n_own] += Dn.xs[:Dn.halo.n_own]
            Dn.halo.exchange(Dn.xa)
            be.spmv(D.P, Dn.xa, D.xa, alpha=1.0, beta=1.0, y0=D.xa)
        else:
            b1 = mg.levels[l + 1].b_
            be.spmv(D.R, D.r, b1)
            if self.part.world > 1:
                all_reduce_sum_(b1)
            x1 = mg._cycle(l + 1, b1)
            if mg.w_from is not None and mg.w_from <= l + 1 <= mg.w_to and l + 1 < len(mg.levels) - 1:
                Lc = mg.levels[l + 1]
                Lc.xs.copy_(x1)
                Lc.bs.copy_(b1)
                be.spmv(Lc.Mop, Lc.xs, b1, alpha=-1.0, beta=1.0, y0=Lc.bs)
                x1 = mg._cycle(l + 1, b1)
                x1.add_(Lc.xs)
            be.spmv(D.P, x1, D.xa, alpha=1.0, beta=1.0, y0=D.xa)
        self._sweeps(D, first_zero=False)

    def __del__(self):
        if getattr(self, "_arena", None) is not None:
            self._arena.group.release(self._arena)

    def _capture(self):
        """Record one distributed cycle -- kernels, halo exchanges and all-reduces, all plain launches over
        peer memory -- into a CUDA graph (input dl[0].b, output dl[0].xa)."""
        be = self.be
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self._cycle(0)                                       # eager pass: lazy buffers / plans exist before capture
            n0 = be.launch_count()
            g = torch.cuda.CUDAGraph()
            g.capture_begin(capture_error_mode="thread_local")
            try:
                self._cycle(0)
            finally:
                g.capture_end()
            self._graph_kernels = be.launch_count() - n0
            be.add_launches(-self._graph_kernels)                # recorded, not run
        cur.wait_stream(side)
        self._graph = g
        self._graph_x = self.dl[0].xa                            # the buffer roles after one cycle are fixed

    def apply(self, v, out):
        D = self.dl[0]
        D.b.copy_(v)
        if self.use_graph:
            if self._graph is None:
                self._capture()
            self._graph.replay()
            self.be.add_launches(self._graph_kernels)
            out.copy_(self._graph_x[:D.halo.n_own])
            return out
        self._cycle(0)
        out.copy_(D.xa[:D.halo.n_own])
        return out
