"""Minimal pure-Python HDF5 reader/writer (superblock v0 files only).

helmholtz-x reads its meshes through XDMF+HDF5 (reference
``helmholtz_x/io_utils.py:161-217``, ``XDMFReader``) and writes results the same
way (``io_utils.py:40-60``).  h5py is not a dependency of this package, so the
small subset of the HDF5 file format those files use is handled here:

* superblock version 0, 8-byte offsets/lengths,
* "old style" groups (symbol-table message 0x11, v1 B-tree + SNOD + local heap),
* version-1 object headers with continuation blocks,
* dataspace v1/v2, fixed-point / IEEE-float datatypes (little endian),
* data layout v3: contiguous, or chunked with an optional deflate filter
  (meshio writes ``/data0``, ``/data1``, ``/data2`` gzip-chunked; DOLFINx writes
  ``/Mesh/Grid/geometry`` … contiguous).

This is host-side file I/O, not part of the measured hot path.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(IOError):
    pass


class H5File:
    """Read-only view of an HDF5 (superblock v0) file: ``f["/a/b"] -> ndarray``."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != _SIG:
            raise H5Error(f"{path}: not an HDF5 file")
        if b[8] != 0:
            raise H5Error(f"{path}: superblock version {b[8]} unsupported (only v0)")
        self.O, self.L = b[13], b[14]
        if (self.O, self.L) != (8, 8):
            raise H5Error("only 8-byte offsets/lengths supported")
        # 8 sig + 8 versions/sizes + 2+2 K + 4 flags = 24 ; then 4 addresses
        self.base = struct.unpack_from("<Q", b, 24)[0]
        root_entry = 24 + 4 * 8
        self.root = self._read_symbol_entry(root_entry)
        self._datasets = {}
        self._walk("", self.root)

    # -- low level -------------------------------------------------------
    def _read_symbol_entry(self, off):
        name_off, hdr, cache = struct.unpack_from("<QQI", self.buf, off)
        scratch = self.buf[off + 24:off + 40]
        ent = {"name_off": name_off, "hdr": hdr, "cache": cache}
        if cache == 1:
            ent["btree"], ent["heap"] = struct.unpack("<QQ", scratch)
        return ent

    def _messages(self, hdr_addr):
        b = self.buf
        ver, _, nmsg, _refc, hsize = struct.unpack_from("<BBHII", b, hdr_addr)
        if ver != 1:
            raise H5Error(f"object header version {ver} unsupported")
        blocks = [(hdr_addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            off, size = blocks.pop(0)
            end = off + size
            while off + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, off)
                data = b[off + 8:off + 8 + msize]
                off += 8 + msize
                if mtype == 0x10:
                    coff, clen = struct.unpack("<QQ", data[:16])
                    blocks.append((coff, clen))
                out.append((mtype, data))
        return out

    def _heap_string(self, heap_addr, off):
        b = self.buf
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5Error("bad local heap")
        data_addr = struct.unpack_from("<Q", b, heap_addr + 24)[0]
        s = data_addr + off
        e = b.index(b"\0", s)
        return b[s:e].decode()

    def _group_entries(self, btree, heap):
        b = self.buf
        if b[btree:btree + 4] != b"TREE":
            raise H5Error("bad B-tree node")
        _ntype, level, nent = struct.unpack_from("<BBH", b, btree + 4)
        off = btree + 8 + 16
        children = []
        off += 8  # key 0
        for _ in range(nent):
            children.append(struct.unpack_from("<Q", b, off)[0])
            off += 16  # child + next key
        for ch in children:
            if level > 0:
                yield from self._group_entries(ch, heap)
            else:
                if b[ch:ch + 4] != b"SNOD":
                    raise H5Error("bad symbol node")
                nsym = struct.unpack_from("<H", b, ch + 6)[0]
                for k in range(nsym):
                    ent = self._read_symbol_entry(ch + 8 + 40 * k)
                    yield self._heap_string(heap, ent["name_off"]), ent

    def _walk(self, prefix, ent):
        msgs = self._messages(ent["hdr"])
        st = [d for t, d in msgs if t == 0x11]
        if st:
            btree, heap = struct.unpack("<QQ", st[0][:16])
            for name, child in self._group_entries(btree, heap):
                self._walk(prefix + "/" + name, child)
        elif any(t == 0x08 for t, _ in msgs):
            self._datasets[prefix] = msgs

    # -- datasets ----------------------------------------------------------
    def keys(self):
        return sorted(self._datasets)

    def __contains__(self, name):
        return name in self._datasets

    @staticmethod
    def _dtype(d):
        cls = d[0] & 0x0F
        bits0 = d[1]
        size = struct.unpack_from("<I", d, 4)[0]
        if bits0 & 1:
            raise H5Error("big-endian data unsupported")
        if cls == 0:
            return np.dtype(("<i" if bits0 & 0x08 else "<u") + str(size))
        if cls == 1:
            return np.dtype("<f" + str(size))
        raise H5Error(f"datatype class {cls} unsupported")

    @staticmethod
    def _shape(d):
        ver, rank, flags = d[0], d[1], d[2]
        off = 8 if ver == 1 else 4
        return tuple(struct.unpack_from("<" + "Q" * rank, d, off)) if rank else ()

    def _chunks(self, addr, ndim):
        b = self.buf
        if b[addr:addr + 4] != b"TREE":
            raise H5Error("bad chunk B-tree")
        _ntype, level, nent = struct.unpack_from("<BBH", b, addr + 4)
        off = addr + 24
        keysz = 8 + 8 * ndim
        for _ in range(nent):
            csize, fmask = struct.unpack_from("<II", b, off)
            offs = struct.unpack_from("<" + "Q" * ndim, b, off + 8)
            child = struct.unpack_from("<Q", b, off + keysz)[0]
            off += keysz + 8
            if level > 0:
                yield from self._chunks(child, ndim)
            else:
                yield csize, fmask, offs, child

    def __getitem__(self, name):
        if not name.startswith("/"):
            name = "/" + name
        msgs = self._datasets[name]
        m = {t: d for t, d in msgs}
        shape = self._shape(m[0x01])
        dt = self._dtype(m[0x03])
        lay = m[0x08]
        if lay[0] != 3:
            raise H5Error(f"data layout version {lay[0]} unsupported")
        cls = lay[1]
        n = int(np.prod(shape)) if shape else 1
        if cls == 1:
            addr, size = struct.unpack_from("<QQ", lay, 2)
            if addr == _UNDEF:
                return np.zeros(shape, dt)
            return np.frombuffer(self.buf, dt, n, addr).reshape(shape).copy()
        if cls == 0:
            size = struct.unpack_from("<H", lay, 2)[0]
            return np.frombuffer(lay[4:4 + size], dt, n).reshape(shape).copy()
        if cls == 2:
            ndim = lay[2]
            addr = struct.unpack_from("<Q", lay, 3)[0]
            cdims = struct.unpack_from("<" + "I" * ndim, lay, 11)
            cshape = cdims[:-1]
            deflate = False
            if 0x0B in m:
                f = m[0x0B]
                fver, nf = f[0], f[1]
                off = 8 if fver == 1 else 2
                for _ in range(nf):
                    fid, nlen, _fl, ncd = struct.unpack_from("<HHHH", f, off)
                    off += 8 + ((nlen + 7) // 8 * 8 if fver == 1 else nlen)
                    off += 4 * ncd + (4 if (fver == 1 and ncd % 2) else 0)
                    if fid == 1:
                        deflate = True
                    elif fid == 2:
                        raise H5Error("shuffle filter unsupported")
                    else:
                        raise H5Error(f"filter {fid} unsupported")
            out = np.zeros(shape, dt)
            if addr == _UNDEF:
                return out
            for csize, fmask, offs, caddr in self._chunks(addr, ndim):
                raw = self.buf[caddr:caddr + csize]
                if deflate and not (fmask & 1):
                    raw = zlib.decompress(raw)
                chunk = np.frombuffer(raw, dt, int(np.prod(cshape))).reshape(cshape)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, shape))
                csl = tuple(slice(0, s.stop - s.start) for s in sl)
                out[sl] = chunk[csl]
            return out
        raise H5Error(f"layout class {cls} unsupported")


# ---------------------------------------------------------------------------
# writer: contiguous datasets inside nested old-style groups
# ---------------------------------------------------------------------------

def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


class _Writer:
    def __init__(self):
        self.buf = bytearray()

    def alloc(self, data):
        self.buf += b"\0" * (-len(self.buf) % 8)
        off = len(self.buf)
        self.buf += data
        return off


def _msg(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data


def _obj_header(msgs):
    body = b"".join(msgs)
    return struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\0" * 4 + body


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        size = dt.itemsize
        if size == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits = bytes([0x20, 0x3F, 0x00])
        else:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits = bytes([0x20, 0x1F, 0x00])
        return bytes([0x11]) + bits + struct.pack("<I", size) + props
    if dt.kind in "iu":
        bits = bytes([0x08 if dt.kind == "i" else 0x00, 0, 0])
        return bytes([0x10]) + bits + struct.pack("<I", dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    raise H5Error(f"cannot write dtype {dt}")


def write_h5(path, datasets):
    """Write ``{"/grp/name": ndarray}`` as contiguous little-endian datasets."""
    w = _Writer()
    w.buf += b"\0" * 96  # superblock v0 with root symbol-table entry

    tree = {}
    for name, arr in datasets.items():
        parts = [p for p in name.split("/") if p]
        node = tree
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = np.ascontiguousarray(arr)

    def emit_dataset(arr):
        dt = arr.dtype.newbyteorder("<") if arr.dtype.byteorder == ">" else arr.dtype
        raw = arr.astype(dt, copy=False).tobytes()
        addr = w.alloc(raw) if raw else _UNDEF
        rank = arr.ndim
        space = struct.pack("<BBBB", 1, rank, 0, 0) + b"\0" * 4 + struct.pack("<" + "Q" * rank, *arr.shape)
        layout = struct.pack("<BB", 3, 1) + struct.pack("<QQ", addr, len(raw))
        hdr = _obj_header([_msg(0x01, space), _msg(0x03, _dtype_msg(dt), 1), _msg(0x08, layout)])
        return w.alloc(hdr)

    def emit_group(node):
        names = sorted(node)
        if len(names) > 32:
            raise H5Error("more than 32 entries in one group unsupported by this writer")
        children = {}
        for nm in names:
            v = node[nm]
            children[nm] = emit_group(v) if isinstance(v, dict) else (emit_dataset(v), None, None)
        heap_data = bytearray(b"\0" * 8)
        name_off = {}
        for nm in names:
            name_off[nm] = len(heap_data)
            heap_data += _pad8(nm.encode() + b"\0")
        heap_data += b"\0" * 16
        free_off = len(heap_data) - 16
        struct.pack_into("<QQ", heap_data, free_off, 1, 16)
        data_addr = w.alloc(bytes(heap_data))
        heap = w.alloc(b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(heap_data), free_off, data_addr))
        snod = bytearray(b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(names)))
        for nm in names:
            hdr, bt, hp = children[nm]
            if bt is None:
                snod += struct.pack("<QQII", name_off[nm], hdr, 0, 0) + b"\0" * 16
            else:
                snod += struct.pack("<QQII", name_off[nm], hdr, 1, 0) + struct.pack("<QQ", bt, hp)
        snod += b"\0" * (8 + 40 * 32 - len(snod))
        snod_addr = w.alloc(bytes(snod))
        last = name_off[names[-1]] if names else 0
        bt = b"TREE" + struct.pack("<BBH", 0, 0, 1) + struct.pack("<QQ", _UNDEF, _UNDEF)
        bt += struct.pack("<QQQ", 0, snod_addr, last)
        bt += b"\0" * (24 + 8 + 32 * 16 - len(bt))
        bt_addr = w.alloc(bt)
        hdr = w.alloc(_obj_header([_msg(0x11, struct.pack("<QQ", bt_addr, heap))]))
        return hdr, bt_addr, heap

    root_hdr, root_bt, root_heap = emit_group(tree)
    w.buf += b"\0" * (-len(w.buf) % 8)
    sb = _SIG + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, _UNDEF, len(w.buf), _UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", root_bt, root_heap)
    w.buf[:len(sb)] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(w.buf))
