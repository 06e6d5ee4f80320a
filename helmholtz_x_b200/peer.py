"""NVLink peer-memory transport of the multi-GPU data path (one process per GPU, one node).

What PETSc does with MPI for the reference under ``mpirun`` -- the VecScatter inside every MatMult
on an MPIAIJ matrix (helmholtz_x/petsc4py_utils.py:86,96) and the MPI_Allreduce inside
VecDot / VecNorm / BVOrthogonalize (helmholtz_x/eigensolvers.py:62,113) -- is done here by two
kernels of libhx_b200 (csrc/hx_peer.cu) that store straight into the neighbours' HBM through
NVSwitch and synchronise with sequence flags: ``hx_peer_halo_exchange`` and ``hx_peer_allreduce``.
No NCCL call and no host round trip sits on the per-iteration path, so whole multigrid cycles with
their exchanges are captured in CUDA graphs.  ``torch.distributed`` (NCCL, or gloo when several
ranks share one GPU in the tests) is only the set-up plumbing: it carries the CUDA IPC handles and
the integer exchange plans.

Buffers that neighbours write into live in *arenas*: cudaMalloc'ed by ``hx_peer_alloc``, exported
with CUDA IPC and mapped by every rank.  Arenas are pooled and never freed during a run (an importer
must not outlive the exporter's cudaFree); the layout inside an arena uses the maximum size over ranks
for every buffer, so the same offset is valid on every rank.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

HX_PEER_MAX = 15
FLAG_KINDS = 3


class HaloDesc(C.Structure):
    """hx_peer_halo_desc (include/hx_b200.h)."""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("n_nb", C.c_int32), ("pad_", C.c_int32),
                ("nb_rank", C.c_int32 * (HX_PEER_MAX + 1)), ("send_ptr", C.c_int64 * (HX_PEER_MAX + 1)),
                ("send_idx", C.c_void_p), ("dst", C.c_void_p * HX_PEER_MAX), ("nb_flags", C.c_void_p * HX_PEER_MAX),
                ("my_flags", C.c_void_p), ("chan_seq", C.c_void_p), ("block_counter", C.c_void_p), ("err", C.c_void_p)]


class AllreduceDesc(C.Structure):
    """hx_peer_allreduce_desc (include/hx_b200.h)."""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("slot_bytes", C.c_int64),
                ("slots", C.c_void_p * (HX_PEER_MAX + 1)), ("flags", C.c_void_p * (HX_PEER_MAX + 1)),
                ("my_flags", C.c_void_p), ("seq", C.c_void_p), ("block_counter", C.c_void_p), ("err", C.c_void_p)]


class _Raw:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _align(n, a=256):
    return (int(n) + a - 1) // a * a


def transport():
    """'peer' (default on CUDA with more than one rank) or 'nccl' (torch.distributed send/recv + all_reduce,
    HX_DIST_TRANSPORT=nccl; also what the CPU test double uses over gloo)."""
    return os.environ.get("HX_DIST_TRANSPORT", "peer")


class Arena:
    """One IPC-shared allocation: `base` on this rank, `peers[q]` = the same allocation of rank q mapped here."""

    def __init__(self, group, base, peers, capacity):
        self.group, self.base, self.peers, self.capacity = group, base, peers, capacity
        self.free = False
        self.offset = 0
        self.bytes = torch.as_tensor(_Raw(base, capacity), device=group.device)

    def reset(self):
        self.offset = 0
        self.bytes.zero_()

    def take(self, max_numel, dtype, numel):
        """Sub-buffer sized for max_numel (the maximum over ranks, so offsets agree); returns a tensor of
        `numel` elements registered with the group."""
        item = torch.empty(0, dtype=dtype).element_size()
        off = self.offset
        self.offset = _align(off + max_numel * item)
        if self.offset > self.capacity:
            raise RuntimeError("peer arena overflow")
        t = self.bytes[off:off + max(numel, 1) * item].view(dtype)[:numel]
        self.group._registry[t.data_ptr()] = (self, off)
        return t


class PeerGroup:
    """Process-wide: control arena (flags, sequence counters, all-reduce slots) + the arena pool."""
    _instance = None

    @classmethod
    def get(cls):
        if cls._instance is None:
            cls._instance = PeerGroup()
        return cls._instance

    def __init__(self, slot_bytes=4 << 20):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerGroup needs an initialised torch.distributed process group")
        if not torch.cuda.is_available():
            raise _lib.HxLibraryError("the peer-memory transport needs CUDA devices")
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > HX_PEER_MAX + 1:
            raise RuntimeError(f"peer transport supports up to {HX_PEER_MAX + 1} ranks on one node")
        self.device = torch.device(f"cuda:{torch.cuda.current_device()}")
        self.slot_bytes = int(slot_bytes)
        self._registry = {}
        self._pool = []
        W = self.world
        # control arena layout
        self.o_flags = 0
        self.o_chan = _align(FLAG_KINDS * W * 8)
        self.o_arseq = self.o_chan + _align(W * 8)
        self.o_cnt = self.o_arseq + 256
        self.o_err = self.o_cnt + 256
        self.o_slots = self.o_err + 256
        total = self.o_slots + 2 * W * self.slot_bytes
        self.ctrl_base, self.ctrl_peers = self._alloc(total)
        self.ctrl = torch.as_tensor(_Raw(self.ctrl_base, total), device=self.device)
        self.err = self.ctrl[self.o_err:self.o_err + 4].view(torch.int32)
        d = AllreduceDesc()
        d.world, d.rank, d.slot_bytes = W, self.rank, self.slot_bytes
        for q in range(W):
            d.slots[q] = self.ctrl_peers[q] + self.o_slots
            d.flags[q] = self.ctrl_peers[q] + self.o_flags
        d.my_flags = self.ctrl_base + self.o_flags
        d.seq = self.ctrl_base + self.o_arseq
        d.block_counter = self.ctrl_base + self.o_cnt + 16
        d.err = self.ctrl_base + self.o_err
        self._ar = d
        dist.barrier()

    # -- allocation ---------------------------------------------------------------------------------
    def _alloc(self, nbytes):
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        _lib.call("hx_peer_alloc", int(nbytes), C.byref(ptr), handle)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle))
        peers = []
        for q in range(self.world):
            if q == self.rank:
                peers.append(ptr.value)
            else:
                p = C.c_void_p()
                hq = (C.c_ubyte * 64).from_buffer_copy(handles[q])
                _lib.call("hx_peer_open", hq, C.byref(p))
                peers.append(p.value)
        return ptr.value, peers

    def _all_min_max(self, values):
        """(min, max) over ranks of a small list of integers (host plumbing)."""
        v = torch.tensor(list(values), dtype=torch.int64)
        lo, hi = v.clone(), v.clone()
        if dist.get_backend() == "nccl":
            lo, hi = lo.to(self.device), hi.to(self.device)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return lo.cpu().tolist(), hi.cpu().tolist()

    def lease(self, nbytes):
        """An arena of at least `nbytes` (collective; every rank passes the same number).  Pooled:
        a free arena is reused when EVERY rank has it free, otherwise all ranks allocate a new one."""
        nbytes = _align(max(int(nbytes), 256), 1 << 20)
        cand = [i for i, a in enumerate(self._pool) if a.free and a.capacity >= nbytes]
        have = [1 if i in cand else 0 for i in range(len(self._pool))]
        lo, _ = self._all_min_max(have + [1])
        common = [i for i in range(len(self._pool)) if lo[i] == 1]
        torch.cuda.synchronize()
        if common:
            arena = self._pool[common[0]]
            arena.free = False
            arena.reset()
            torch.cuda.synchronize()
            dist.barrier()
            return arena
        base, peers = self._alloc(nbytes)
        arena = Arena(self, base, peers, nbytes)
        self._pool.append(arena)
        return arena

    def release(self, arena):
        for k in [k for k, v in self._registry.items() if v[0] is arena]:
            del self._registry[k]
        arena.free = True

    def lookup(self, t):
        """(arena, byte offset) of a tensor handed out by Arena.take (views at the same start address)."""
        hit = self._registry.get(t.data_ptr())
        if hit is None:
            raise RuntimeError("halo exchange on a vector that does not live in a peer arena")
        return hit

    # -- the two data-path operations ------------------------------------------------------------------
    def allreduce_(self, t):
        """In-place sum over ranks of a contiguous float32 / float64 / complex tensor (one kernel)."""
        assert t.is_contiguous()
        if t.is_complex():
            t = torch.view_as_real(t)
        is_f32 = t.dtype == torch.float32
        assert is_f32 or t.dtype == torch.float64
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.call("hx_peer_allreduce", C.byref(self._ar), t.data_ptr(), t.data_ptr(), int(t.numel()), int(is_f32), st)
        return t

    def halo_desc(self, nb_ranks, send_counts, send_idx32, dst_ptrs):
        d = HaloDesc()
        d.world, d.rank, d.n_nb = self.world, self.rank, len(nb_ranks)
        off = 0
        for i, q in enumerate(nb_ranks):
            d.nb_rank[i] = int(q)
            d.send_ptr[i] = off
            off += int(send_counts[i])
            d.dst[i] = int(dst_ptrs[i])
            d.nb_flags[i] = self.ctrl_peers[q] + self.o_flags
        d.send_ptr[len(nb_ranks)] = off
        d.send_idx = send_idx32.data_ptr()
        d.my_flags = self.ctrl_base + self.o_flags
        d.chan_seq = self.ctrl_base + self.o_chan
        d.block_counter = self.ctrl_base + self.o_cnt
        d.err = self.ctrl_base + self.o_err
        return d

    def halo_exchange(self, desc, x_loc):
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.call("hx_peer_halo_exchange", C.byref(desc), x_loc.data_ptr(), int(x_loc.element_size()), st)

    def check(self):
        """Raise if a device-side wait timed out (synchronises)."""
        code = int(self.err.cpu()[0])
        if code:
            raise RuntimeError(f"peer transport: a wait on another GPU timed out (code {code}: 1/2 = halo phase A/B, 3 = all-reduce)")


class HaloExchanger:
    """Exchange plan of one vector layout [owned | ghosts grouped by owner]: which local entries go to which
    rank (send_idx, send_counts) and how many arrive from each (recv_counts).  Two transports behind one
    call: the peer-memory kernel, or torch.distributed send/recv (NCCL / gloo)."""

    def __init__(self, world, rank, n_own, send_idx, send_counts, recv_counts):
        self.world, self.rank, self.n_own = world, rank, int(n_own)
        self.send_idx = send_idx                        # device int64, concatenated over destination ranks
        self.send_counts = np.asarray(send_counts, np.int64)
        self.recv_counts = np.asarray(recv_counts, np.int64)
        self.n_ghost = int(self.recv_counts.sum())
        self._plans = {}
        self._peer = None
        self.active = world > 1 and (self.n_ghost > 0 or int(self.send_counts.sum()) > 0)
        if world > 1 and send_idx.is_cuda and transport() == "peer":
            self._setup_peer()

    def _setup_peer(self):
        grp = PeerGroup.get()
        info = [None] * self.world
        dist.all_gather_object(info, (self.n_own, self.recv_counts.tolist()))
        self.nb = [q for q in range(self.world) if q != self.rank and (self.send_counts[q] or self.recv_counts[q])]
        # where my values land in neighbour q's vector: after its owned entries and the ghosts owned by lower ranks
        self.remote_start = []
        for q in self.nb:
            n_own_q, recv_q = info[q]
            self.remote_start.append(int(n_own_q) + int(sum(recv_q[r] for r in range(self.rank) if r != q)))
        self.send_idx32 = self.send_idx.to(torch.int32).contiguous()
        self._peer = grp

    def exchange(self, x_loc):
        """Fill the ghost tail of x_loc from the owners."""
        if not self.active and self._peer is None:
            return x_loc
        if self._peer is not None:
            if not self.nb:
                return x_loc
            key = x_loc.data_ptr()
            desc = self._plans.get(key)
            if desc is None:
                arena, off = self._peer.lookup(x_loc)
                item = x_loc.element_size()
                dst = [arena.peers[q] + off + self.remote_start[i] * item for i, q in enumerate(self.nb)]
                desc = self._peer.halo_desc(self.nb, [self.send_counts[q] for q in self.nb], self.send_idx32, dst)
                self._plans[key] = desc
            self._peer.halo_exchange(desc, x_loc)
            return x_loc
        key = (x_loc.data_ptr(), x_loc.numel())
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) > 16:
                self._plans.clear()
            sendbuf = torch.zeros(max(int(self.send_idx.numel()), 1), dtype=x_loc.dtype, device=x_loc.device)
            sreal = torch.view_as_real(sendbuf)
            ghost = torch.view_as_real(x_loc[self.n_own:])
            ops, soff, roff = [], 0, 0
            for q in range(self.world):
                if q == self.rank:
                    continue
                ns, nr = int(self.send_counts[q]), int(self.recv_counts[q])
                if ns:
                    ops.append(dist.P2POp(dist.isend, sreal[soff:soff + ns], q))
                if nr:
                    ops.append(dist.P2POp(dist.irecv, ghost[roff:roff + nr], q))
                soff += ns
                roff += nr
            plan = (sendbuf, ops, x_loc)          # keep x_loc alive: the plan is keyed by its address
            self._plans[key] = plan
        sendbuf, ops, _ = plan
        if self.send_idx.numel():
            torch.index_select(x_loc, 0, self.send_idx, out=sendbuf)
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return x_loc
