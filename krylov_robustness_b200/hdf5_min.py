"""A minimal HDF5 reader for MATLAB v7.3 MAT-files (row f4 of SURVEY.md section 8: the three files of
datasets_paper/Misc that are not MAT-v5 - CollegeMsg, Drugs, as_735).  The image has no h5py / HDF5 library.

Covers what MATLAB's `save -v7.3` writes: a 512-byte user block, superblock version 0, old-style groups (B-tree v1 +
local heap + symbol nodes), version-1 object headers with continuation blocks, dataspace / datatype / layout (compact,
contiguous, chunked with a v1 chunk B-tree) / filter-pipeline (deflate, shuffle) / attribute messages, fixed-point and
IEEE floating-point datatypes.  Sparse matrices come back as scipy CSC (MATLAB stores `data`, `ir`, `jc` datasets in a
group whose attribute MATLAB_sparse holds the row count), dense numeric arrays as column-major-correct ndarrays,
structs as dicts.  Anything else raises Hdf5Error - this is a dataset loader, not a general HDF5 implementation."""
import struct
import zlib

import numpy as np
import scipy.sparse as sp

UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(Exception):
    pass


class _File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.b = fh.read()
        self.base = None
        for off in (0, 512, 1024, 2048):
            if self.b[off:off + 8] == b"\x89HDF\r\n\x1a\n":
                self.base = off
                break
        if self.base is None:
            raise Hdf5Error("no HDF5 signature")
        sb = self.base
        ver = self.b[sb + 8]
        if ver not in (0, 1):
            raise Hdf5Error("superblock version %d not supported" % ver)
        self.so, self.sl = self.b[sb + 13], self.b[sb + 14]
        if (self.so, self.sl) != (8, 8):
            raise Hdf5Error("only 8-byte offsets / lengths are supported")
        p = sb + 24 + (4 if ver == 1 else 0)
        # base address, free-space address, end of file, driver info, then the root group's symbol table entry
        # the base address is absolute (MATLAB: 512, the user block); every other address is relative to it
        self.base = self.u64(p)
        root = p + 32
        self.root_header = self.u64(root + 8)
        cache = self.u32(root + 16)
        self.root_btree = self.root_heap = None
        if cache == 1:
            self.root_btree, self.root_heap = self.u64(root + 24), self.u64(root + 32)

    # addresses in the file are relative to the base address
    def u8(self, p):
        return self.b[p]

    def u16(self, p):
        return struct.unpack_from("<H", self.b, p)[0]

    def u32(self, p):
        return struct.unpack_from("<I", self.b, p)[0]

    def u64(self, p):
        return struct.unpack_from("<Q", self.b, p)[0]

    def at(self, addr):
        return self.base + addr

    # ---------------------------------------------------------------- object headers (version 1)
    def messages(self, addr):
        p = self.at(addr)
        if self.b[p] != 1:
            raise Hdf5Error("object header version %d not supported" % self.b[p])
        nmsg = self.u16(p + 2)
        size = self.u32(p + 8)
        out = []
        blocks = [(p + 16, size)]
        while blocks and len(out) < nmsg:
            q, left = blocks.pop(0)
            end = q + left
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self.u16(q), self.u16(q + 2), self.b[q + 4]
                body = q + 8
                if mtype == 0x10:
                    blocks.append((self.at(self.u64(body)), self.u64(body + 8)))
                out.append((mtype, body, msize, flags))
                q = body + msize
        return out

    # ---------------------------------------------------------------- groups
    def heap_data(self, heap_addr):
        p = self.at(heap_addr)
        if self.b[p:p + 4] != b"HEAP":
            raise Hdf5Error("bad local heap")
        return self.at(self.u64(p + 24))

    def links(self, btree_addr, heap_addr):
        """name -> object header address for an old-style group"""
        data = self.heap_data(heap_addr)
        out = {}

        def walk(addr):
            p = self.at(addr)
            sig = self.b[p:p + 4]
            if sig == b"TREE":
                ntype, level, used = self.b[p + 4], self.b[p + 5], self.u16(p + 6)
                if ntype != 0:
                    raise Hdf5Error("group B-tree expected")
                q = p + 24
                for i in range(used):
                    child = self.u64(q + 8)          # key, child, key, child ...
                    walk(child)
                    q += 16
            elif sig == b"SNOD":
                n = self.u16(p + 6)
                q = p + 8
                for i in range(n):
                    name_off, hdr = self.u64(q), self.u64(q + 8)
                    e = self.b.index(b"\0", data + name_off)
                    out[self.b[data + name_off:e].decode()] = hdr
                    q += 40
            else:
                raise Hdf5Error("unexpected node %r in a group B-tree" % sig)
        walk(btree_addr)
        return out

    # ---------------------------------------------------------------- messages
    def datatype(self, p):
        cv = self.b[p]
        cls, ver = cv & 0x0F, cv >> 4
        bits0 = self.b[p + 1]
        size = self.u32(p + 4)
        if cls == 0:                                   # fixed point
            dt = np.dtype("%s%s%d" % (">" if bits0 & 1 else "<", "i" if bits0 & 8 else "u", size))
            return dt, 8 + 4
        if cls == 1:                                   # floating point
            dt = np.dtype("%sf%d" % (">" if bits0 & 1 else "<", size))
            return dt, 8 + 12
        if cls == 3:                                   # fixed-length string
            return np.dtype("S%d" % size), 8
        raise Hdf5Error("datatype class %d not supported" % cls)

    def dataspace(self, p):
        ver, rank, flags = self.b[p], self.b[p + 1], self.b[p + 2]
        q = p + (8 if ver == 1 else 4)
        dims = [self.u64(q + 8 * i) for i in range(rank)]
        return dims

    def attributes(self, msgs):
        out = {}
        for mtype, p, msize, flags in msgs:
            if mtype != 0x0C:
                continue
            ver = self.b[p]
            nsz, tsz, ssz = self.u16(p + 2), self.u16(p + 4), self.u16(p + 6)
            if ver == 1:
                pad = lambda x: (x + 7) & ~7
                q = p + 8
            elif ver in (2, 3):
                pad = lambda x: x
                q = p + 8 + (1 if ver == 3 else 0)
            else:
                continue
            name = self.b[q:q + nsz].split(b"\0")[0].decode()
            q += pad(nsz)
            try:
                dt, _ = self.datatype(q)
            except Hdf5Error:
                continue
            q += pad(tsz)
            dims = self.dataspace(q)
            q += pad(ssz)
            n = int(np.prod(dims)) if dims else 1
            val = np.frombuffer(self.b, dtype=dt, count=n, offset=q)
            out[name] = val[0] if n == 1 else val.copy()
        return out

    # ---------------------------------------------------------------- datasets
    def read_dataset(self, msgs):
        dt = dims = layout = None
        filters = []
        for mtype, p, msize, flags in msgs:
            if mtype == 0x01:
                dims = self.dataspace(p)
            elif mtype == 0x03:
                dt, _ = self.datatype(p)
            elif mtype == 0x08:
                layout = p
            elif mtype == 0x0B:
                ver, nf = self.b[p], self.b[p + 1]
                q = p + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid = self.u16(q)
                    if ver == 1 or fid >= 256:
                        nlen = self.u16(q + 2)
                        ncv = self.u16(q + 6)
                        q += 8 + ((nlen + 7) & ~7 if ver == 1 else nlen)
                    else:
                        ncv = self.u16(q + 4)
                        q += 6
                    cvals = [self.u32(q + 4 * i) for i in range(ncv)]
                    q += 4 * ncv
                    if ver == 1 and ncv % 2:
                        q += 4
                    filters.append((fid, cvals))
        if dt is None or dims is None or layout is None:
            raise Hdf5Error("not a dataset")
        n = int(np.prod(dims)) if dims else 1
        ver = self.b[layout]
        if ver != 3:
            raise Hdf5Error("layout message version %d not supported" % ver)
        cls = self.b[layout + 1]
        if cls == 0:                                   # compact
            size = self.u16(layout + 2)
            raw = self.b[layout + 4:layout + 4 + size]
            return np.frombuffer(raw, dtype=dt, count=n).reshape(dims)
        if cls == 1:                                   # contiguous
            addr = self.u64(layout + 2)
            if addr == UNDEF:
                return np.zeros(dims, dtype=dt)
            return np.frombuffer(self.b, dtype=dt, count=n, offset=self.at(addr)).reshape(dims).copy()
        if cls != 2:
            raise Hdf5Error("layout class %d not supported" % cls)
        nd = self.b[layout + 2]                        # dimensionality + 1
        btree = self.u64(layout + 3)
        cdims = [self.u32(layout + 11 + 4 * i) for i in range(nd)][:-1]
        out = np.zeros(dims, dtype=dt)
        if btree == UNDEF:
            return out

        def walk(addr):
            p = self.at(addr)
            if self.b[p:p + 4] != b"TREE" or self.b[p + 4] != 1:
                raise Hdf5Error("chunk B-tree expected")
            level, used = self.b[p + 5], self.u16(p + 6)
            q = p + 24
            ksz = 8 + 8 * nd
            for i in range(used):
                csize, mask = self.u32(q), self.u32(q + 4)
                offs = [self.u64(q + 8 + 8 * d) for d in range(nd - 1)]
                child = self.u64(q + ksz)
                if level > 0:
                    walk(child)
                else:
                    raw = self.b[self.at(child):self.at(child) + csize]
                    for k in range(len(filters) - 1, -1, -1):
                        if mask & (1 << k):
                            continue
                        fid, cv = filters[k]
                        if fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:
                            es = cv[0] if cv else dt.itemsize
                            a = np.frombuffer(raw, dtype=np.uint8)
                            m = a.size // es
                            raw = a[:m * es].reshape(es, m).T.tobytes() + a[m * es:].tobytes()
                        elif fid == 3:
                            raw = raw[:-4]             # fletcher32 checksum
                        else:
                            raise Hdf5Error("filter %d not supported" % fid)
                    chunk = np.frombuffer(raw, dtype=dt, count=int(np.prod(cdims))).reshape(cdims)
                    sl = tuple(slice(o, min(o + c, d)) for o, c, d in zip(offs, cdims, dims))
                    out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
                q += ksz + 8
        walk(btree)
        return out

    # ---------------------------------------------------------------- MATLAB objects
    def read_object(self, header_addr):
        msgs = self.messages(header_addr)
        attrs = self.attributes(msgs)
        stab = [(p) for mtype, p, msize, flags in msgs if mtype == 0x11]
        cls = attrs.get("MATLAB_class", b"")
        cls = cls.decode() if isinstance(cls, bytes) else str(cls)
        if stab:
            members = self.links(self.u64(stab[0]), self.u64(stab[0] + 8))
            if "MATLAB_sparse" in attrs:
                nrows = int(attrs["MATLAB_sparse"])
                jc = self.read_dataset(self.messages(members["jc"])).astype(np.int64).ravel()
                if "data" in members:
                    data = self.read_dataset(self.messages(members["data"])).ravel().astype(np.float64)
                    ir = self.read_dataset(self.messages(members["ir"])).astype(np.int64).ravel()
                else:
                    data, ir = np.zeros(0), np.zeros(0, dtype=np.int64)
                return sp.csc_matrix((data, ir, jc), shape=(nrows, jc.size - 1))
            return {k: self.read_object(v) for k, v in members.items() if not k.startswith("#")}
        arr = self.read_dataset(msgs)
        if "MATLAB_empty" in attrs and int(attrs["MATLAB_empty"]):
            return np.zeros(tuple(int(v) for v in arr.ravel()))
        if cls == "char":
            return "".join(chr(int(c)) for c in arr.ravel())
        if arr.dtype.kind in "fiu":
            a = arr.astype(np.float64) if cls in ("double", "single", "") or arr.dtype.kind == "f" else arr
            return a.T if a.ndim == 2 else a           # HDF5 stores the transposed (row-major) array
        return arr


def loadmat73(path):
    """{variable name: value} of a MATLAB v7.3 MAT-file (sparse -> scipy CSC, numeric -> ndarray, struct -> dict)."""
    f = _File(path)
    if f.root_btree is None:
        msgs = f.messages(f.root_header)
        st = [p for mtype, p, msize, flags in msgs if mtype == 0x11]
        if not st:
            raise Hdf5Error("root group without a symbol table")
        f.root_btree, f.root_heap = f.u64(st[0]), f.u64(st[0] + 8)
    out = {}
    for name, hdr in f.links(f.root_btree, f.root_heap).items():
        if name.startswith("#"):
            continue                                   # #refs#, #subsystem#
        out[name] = f.read_object(hdr)
    return out


def is_v73(path):
    with open(path, "rb") as fh:
        return fh.read(10) == b"MATLAB 7.3"
