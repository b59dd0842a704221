"""Minimal pure-Python reader for the HDF5 files samurai's `save()` writes.

TEST INFRASTRUCTURE ONLY (oracle/): used to read the reference's golden files
(`/root/reference/tests/reference/finite_volume/*.h5`) when generating the
fixtures under tests/golden/.  Handles exactly what those files use: superblock
version 0, old-style groups (symbol-table B-tree v1 + local heap), version-1
object headers, contiguous layout (version 3), fixed-point / IEEE float types.
No h5py is available in this image.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        assert b[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        assert b[8] == 0, "only superblock v0 supported"
        assert b[13] == 8 and b[14] == 8, "only 8-byte offsets/lengths supported"
        # 8 sig + 8 versions/sizes + 4 (leaf k, internal k) + 4 flags = 24
        # then base, free-space, eof, driver-info addresses (4 x 8)
        root_entry = 24 + 32
        self.root = self._read_symbol_entry(root_entry)

    # symbol table entry: link name offset(8) obj header addr(8) cache type(4) reserved(4) scratch(16)
    def _read_symbol_entry(self, off):
        name_off, hdr_addr, cache_type = struct.unpack_from("<QQI", self.buf, off)
        scratch = self.buf[off + 24 : off + 40]
        return dict(name_off=name_off, hdr=hdr_addr, cache=cache_type, scratch=scratch)

    def _messages(self, hdr_addr):
        b = self.buf
        version, _, nmsg, _refcnt, hdr_size = struct.unpack_from("<BBHII", b, hdr_addr)
        assert version == 1, "only object header v1 supported"
        msgs = []
        blocks = [(hdr_addr + 16, hdr_size)]
        while blocks and len(msgs) < nmsg:
            off, size = blocks.pop(0)
            end = off + size
            while off + 8 <= end and len(msgs) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, off)
                body = off + 8
                if mtype == 0x10:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", b, body)
                    blocks.append((caddr, clen))
                msgs.append((mtype, body, msize))
                off = body + msize
        return msgs

    def _group_entries(self, btree_addr, heap_addr):
        b = self.buf
        assert b[heap_addr : heap_addr + 4] == b"HEAP"
        heap_data = struct.unpack_from("<Q", b, heap_addr + 24)[0]
        out = {}

        def walk(addr):
            assert b[addr : addr + 4] == b"TREE"
            ntype, level, nused = struct.unpack_from("<BBH", b, addr + 4)
            assert ntype == 0
            p = addr + 24  # sig4 + 4 + left8 + right8
            # keys and children interleaved: key0 child0 key1 child1 ... keyN
            children = []
            for i in range(nused):
                p += 8  # key
                children.append(struct.unpack_from("<Q", b, p)[0])
                p += 8
            for c in children:
                if level > 0:
                    walk(c)
                else:
                    assert b[c : c + 4] == b"SNOD"
                    nsym = struct.unpack_from("<H", b, c + 6)[0]
                    for k in range(nsym):
                        e = self._read_symbol_entry(c + 8 + 40 * k)
                        s = heap_data + e["name_off"]
                        name = b[s : b.index(b"\x00", s)].decode()
                        out[name] = e

        walk(btree_addr)
        return out

    def _open_group(self, entry):
        if entry["cache"] == 1:
            btree, heap = struct.unpack_from("<QQ", entry["scratch"], 0)
            return self._group_entries(btree, heap)
        for mtype, body, _ in self._messages(entry["hdr"]):
            if mtype == 0x11:  # symbol table message
                btree, heap = struct.unpack_from("<QQ", self.buf, body)
                return self._group_entries(btree, heap)
        raise KeyError("not a group")

    def _lookup(self, path):
        entry = self.root
        for part in [p for p in path.split("/") if p]:
            entry = self._open_group(entry)[part]
        return entry

    def listdir(self, path="/"):
        return sorted(self._open_group(self._lookup(path)).keys())

    def read(self, path):
        b = self.buf
        entry = self._lookup(path)
        shape = dtype = addr = None
        for mtype, body, _msize in self._messages(entry["hdr"]):
            if mtype == 0x01:  # dataspace
                ver, rank, flags = struct.unpack_from("<BBB", b, body)
                p = body + (8 if ver == 1 else 4)
                shape = struct.unpack_from("<%dQ" % rank, b, p)
            elif mtype == 0x03:  # datatype
                cls_ver, bits0, _b1, _b2, size = struct.unpack_from("<BBBBI", b, body)
                cls = cls_ver & 0x0F
                endian = ">" if (bits0 & 1) else "<"
                if cls == 0:
                    signed = (bits0 >> 3) & 1
                    dtype = np.dtype("%s%s%d" % (endian, "i" if signed else "u", size))
                elif cls == 1:
                    dtype = np.dtype("%sf%d" % (endian, size))
                else:
                    raise NotImplementedError("datatype class %d" % cls)
            elif mtype == 0x08:  # layout
                ver = b[body]
                assert ver == 3, "only layout v3 supported"
                lclass = b[body + 1]
                assert lclass == 1, "only contiguous layout supported"
                addr, _size = struct.unpack_from("<QQ", b, body + 2)
        assert shape is not None and dtype is not None and addr is not None
        n = int(np.prod(shape)) if len(shape) else 1
        if addr == UNDEF:
            return np.zeros(shape, dtype)
        return np.frombuffer(b, dtype=dtype, count=n, offset=addr).reshape(shape).copy()


def read_samurai_mesh(path, dim=2):
    """Return (level, idx[N,dim], u[N]) per cell, in file (= for_each_cell) order.

    Cells are quads given by 4 point ids into /mesh/points; (level, i, j) is
    recovered from the quad corners: length = x1-x0 = 2^-level, i = x0/length.
    (reference io/hdf5.hpp: points + connectivity + fields/<name>)
    """
    h5 = H5File(path)
    pts = h5.read("/mesh/points")
    conn = h5.read("/mesh/connectivity")
    fields = {name: h5.read("/mesh/fields/" + name) for name in h5.listdir("/mesh/fields")}
    nper = 1 << dim
    conn = conn.reshape(-1, nper)
    corners = pts[conn.astype(np.int64)]  # [N, nper, 3]
    lo = corners.min(axis=1)[:, :dim]
    hi = corners.max(axis=1)[:, :dim]
    length = hi[:, 0] - lo[:, 0]
    level = np.rint(-np.log2(length)).astype(np.int64)
    idx = np.rint(lo / length[:, None]).astype(np.int64)
    return level, idx, fields
