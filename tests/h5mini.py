"""Independent minimal HDF5 reader (test infrastructure): parses the subset of the HDF5 file format that
hemocell_b200/host/hemo_h5.cpp writes -- and that h5py/libhdf5 wrote for the reference's output in the
same era (superblock v0, v1 object headers, group B-tree + SNOD + local heap, contiguous or chunked
(B-tree v1) + deflate datasets, v1 attribute messages).  No libhdf5 / h5py exists in this image, so this
reader, written from the format specification, is what checks the writer's files structurally."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


class H5Error(ValueError):
    pass


def _dtype(msg):
    cls_ver = msg[0]
    cls, ver = cls_ver & 15, cls_ver >> 4
    if ver != 1:
        raise H5Error("datatype version %d" % ver)
    bits0 = msg[1]
    size = struct.unpack_from("<I", msg, 4)[0]
    if bits0 & 1:
        raise H5Error("big endian")
    if cls == 0:
        off, prec = struct.unpack_from("<HH", msg, 8)
        if off != 0 or prec != 8*size:
            raise H5Error("odd integer layout")
        return np.dtype("<%s%d" % ("i" if bits0 & 8 else "u", size))
    if cls == 1:
        off, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", msg, 8)
        want = {4: (0, 32, 23, 8, 0, 23, 127, 31), 8: (0, 64, 52, 11, 0, 52, 1023, 63)}[size]
        if (off, prec, eloc, esize, mloc, msize, bias, msg[2]) != want or (bits0 >> 4) & 3 != 2:
            raise H5Error("not IEEE float")
        return np.dtype("<f%d" % size)
    raise H5Error("datatype class %d" % cls)


def _dataspace(msg):
    ver, rank, flags = msg[0], msg[1], msg[2]
    if ver != 1:
        raise H5Error("dataspace version %d" % ver)
    return tuple(struct.unpack_from("<%dQ" % rank, msg, 8))


class File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        if b[:8] != SIG:
            raise H5Error("bad signature")
        if b[8] != 0 or b[13] != 8 or b[14] != 8:
            raise H5Error("superblock version / offset sizes")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, 16)
        base, fs, eof, drv = struct.unpack_from("<4Q", b, 24)
        if base != 0 or eof != len(b):
            raise H5Error("base/eof address: eof %d, file %d" % (eof, len(b)))
        name_off, root, cache = struct.unpack_from("<QQI", b, 56)
        self.attrs = {}
        self.datasets = {}
        msgs = self._header(root)
        st = [m for t, m in msgs if t == 0x11]
        if len(st) != 1:
            raise H5Error("root group without symbol table message")
        btree, heap = struct.unpack_from("<QQ", st[0], 0)
        if cache == 1 and struct.unpack_from("<QQ", b, 80) != (btree, heap):
            raise H5Error("root scratch pad disagrees with the symbol table message")
        for t, m in msgs:
            if t == 0x0C:
                name, val = self._attribute(m)
                self.attrs[name] = val
        for name, addr in self._group(btree, heap):
            self.datasets[name] = self._dataset(addr)

    # ---- object header v1
    def _header(self, addr):
        b = self.b
        ver, _, nmsg, refc, size = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise H5Error("object header version %d" % ver)
        p, end, out = addr + 16, addr + 16 + size, []
        while p < end and len(out) < nmsg:
            t, sz, fl = struct.unpack_from("<HHB", b, p)
            if sz % 8:
                raise H5Error("unaligned message size")
            out.append((t, b[p + 8:p + 8 + sz]))
            p += 8 + sz
        if len(out) != nmsg or p != end:
            raise H5Error("object header message count / size mismatch")
        return out

    def _attribute(self, m):
        ver, _, nsz, tsz, ssz = struct.unpack_from("<BBHHH", m, 0)
        if ver != 1:
            raise H5Error("attribute version")
        pad = lambda n: (n + 7)//8*8
        p = 8
        name = m[p:p + nsz].rstrip(b"\0").decode(); p += pad(nsz)
        dt = _dtype(m[p:p + tsz]); p += pad(tsz)
        shape = _dataspace(m[p:p + ssz]); p += pad(ssz)
        n = int(np.prod(shape)) if shape else 1
        return name, np.frombuffer(m, dt, n, p).reshape(shape)

    # ---- root group
    def _group(self, btree, heap):
        b = self.b
        if b[heap:heap + 4] != b"HEAP":
            raise H5Error("local heap signature")
        dsize, free, daddr = struct.unpack_from("<QQQ", b, heap + 8)
        if free != 1 and free >= dsize:
            raise H5Error("bad heap free list")
        if b[btree:btree + 4] != b"TREE" or b[btree + 4] != 0:
            raise H5Error("group B-tree signature / type")
        level, used = struct.unpack_from("<BH", b, btree + 5)
        if level != 0:
            raise H5Error("multi-level group B-tree")
        if btree + 24 + (2*self.internal_k + 1)*8 + 2*self.internal_k*8 > len(b):
            raise H5Error("group B-tree node overruns the file")
        out = []
        for k in range(used):
            key0, child, key1 = struct.unpack_from("<QQQ", b, btree + 24 + 16*k)
            if b[child:child + 4] != b"SNOD":
                raise H5Error("symbol table node signature")
            if child + 8 + 2*self.leaf_k*40 > len(b):
                raise H5Error("symbol table node overruns the file")
            nsym = struct.unpack_from("<H", b, child + 6)[0]
            if nsym > 2*self.leaf_k:
                raise H5Error("too many symbols for leaf K")
            names = []
            for s in range(nsym):
                noff, oaddr, ctype = struct.unpack_from("<QQI", b, child + 8 + 40*s)
                end = b.index(b"\0", daddr + noff)
                names.append(b[daddr + noff:end].decode())
                out.append((names[-1], oaddr))
            if names != sorted(names):
                raise H5Error("symbol table entries not sorted")
            e1 = b.index(b"\0", daddr + key1)
            if b[daddr + key1:e1].decode() != names[-1]:
                raise H5Error("B-tree right key is not the largest name of the child")
        return out

    # ---- datasets
    def _dataset(self, addr):
        msgs = dict()
        for t, m in self._header(addr):
            msgs[t] = m
        shape = _dataspace(msgs[1])
        dt = _dtype(msgs[3])
        lay = msgs[8]
        if lay[0] != 3:
            raise H5Error("layout version")
        n = int(np.prod(shape)) if shape else 1
        info = {"dtype": dt, "shape": shape}
        if lay[1] == 1:
            a, size = struct.unpack_from("<QQ", lay, 2)
            if size != n*dt.itemsize:
                raise H5Error("contiguous size mismatch")
            info["layout"] = "contiguous"
            info["data"] = np.frombuffer(self.b, dt, n, a).reshape(shape) if n else np.zeros(shape, dt)
            return info
        if lay[1] != 2:
            raise H5Error("layout class %d" % lay[1])
        nd = lay[2]
        bt = struct.unpack_from("<Q", lay, 3)[0]
        cdims = struct.unpack_from("<%dI" % nd, lay, 11)
        if nd != len(shape) + 1 or cdims[-1] != dt.itemsize:
            raise H5Error("chunk dimensionality")
        chunk = cdims[:-1]
        deflate = None
        if 0x0B in msgs:
            f = msgs[0x0B]
            if f[0] != 1 or f[1] != 1:
                raise H5Error("filter pipeline")
            fid, nlen, fflags, ncd = struct.unpack_from("<HHHH", f, 8)
            if fid != 1 or nlen != 0:
                raise H5Error("not deflate")
            deflate = struct.unpack_from("<I", f, 16)[0]
        info.update(layout="chunked", chunk=chunk, deflate=deflate, nchunks=0)
        out = np.zeros(shape, dt)
        if bt != UNDEF:
            self._chunks(bt, len(shape), chunk, dt, deflate, out, info)
        info["data"] = out
        return info

    def _chunks(self, addr, rank, chunk, dt, deflate, out, info, lo=None):
        b = self.b
        if b[addr:addr + 4] != b"TREE" or b[addr + 4] != 1:
            raise H5Error("chunk B-tree signature / type")
        level, used = struct.unpack_from("<BH", b, addr + 5)
        ks = 8 + 8*(rank + 1)
        if addr + 24 + 65*ks + 64*8 > len(b):
            raise H5Error("chunk B-tree node overruns the file")
        if used > 64:
            raise H5Error("too many entries")
        p = addr + 24
        prev = None
        for k in range(used):
            nbytes, mask = struct.unpack_from("<II", b, p)
            offs = struct.unpack_from("<%dQ" % (rank + 1), b, p + 8)
            child = struct.unpack_from("<Q", b, p + ks)[0]
            nxt = struct.unpack_from("<%dQ" % (rank + 1), b, p + ks + 8 + 8)
            if offs[-1] != 0 or not (tuple(offs) < tuple(nxt)):
                raise H5Error("chunk keys not strictly ascending")
            if prev is not None and not prev < tuple(offs):
                raise H5Error("chunk keys out of order")
            prev = tuple(offs)
            if level > 0:
                self._chunks(child, rank, chunk, dt, deflate, out, info)
            else:
                raw = b[child:child + nbytes]
                if mask == 0 and deflate is not None:
                    raw = zlib.decompress(raw)
                blk = np.frombuffer(raw, dt).reshape(chunk)
                if any(o % c for o, c in zip(offs, chunk)):
                    raise H5Error("chunk offset not a multiple of the chunk size")
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, out.shape))
                out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
                info["nchunks"] += 1
            p += ks + 8

    def __getitem__(self, name):
        return self.datasets[name]["data"]
