"""Minimal HDF5 reader / writer -- the subset Keras weight files use (h5py is not available offline).

The reference loads its checkpoints with ``model.load_weights('....h5')`` (predict.py:21-49).  Such files are written by
h5py with the library's "earliest" format: superblock version 0, version-1 object headers, groups stored as a
symbol table (version-1 B-tree + local heap + symbol-table nodes), contiguous (or compact) little-endian datasets and
small attributes holding fixed-length strings.  ``File`` reads exactly that (plus superblock 1, continuation blocks,
attribute message versions 1-3, dataspace versions 1-2); anything else -- chunked / filtered datasets, new-style
groups, variable-length data -- raises ``Hdf5Error`` with the feature's name rather than guessing.
``write_file`` produces the same layout (one B-tree node and one symbol-table node per group, leaf K raised in the
superblock so that a group of any size fits), which the reader, and the HDF5 library, can parse; it is what
``Model.save_weights('x.h5')`` and the tests use."""
from __future__ import annotations

import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(OSError):
    """Raised for files this subset cannot read (an OSError, like h5py's, so that predict.load_model's handler sees it)."""


# ======================================================================================= reader
class _Obj:
    def __init__(self, f, addr):
        self.f, self.addr = f, addr
        self.msgs = f._object_messages(addr)
        self._attrs = None

    @property
    def attrs(self):
        if self._attrs is None:
            self._attrs = {}
            for typ, data in self.msgs:
                if typ == 0x000C:
                    name, val = self.f._attribute(data)
                    self._attrs[name] = val
        return self._attrs


class Group(_Obj):
    def __init__(self, f, addr):
        super().__init__(f, addr)
        self._links = None

    def _load(self):
        if self._links is None:
            self._links = {}
            for typ, data in self.msgs:
                if typ == 0x0011:  # symbol table
                    btree, heap = struct.unpack_from("<QQ", data, 0)
                    self._links.update(self.f._group_entries(btree, heap))
                elif typ in (0x0002, 0x0006):
                    raise Hdf5Error("new-style HDF5 groups (link messages, libver='latest') are not supported")
        return self._links

    def keys(self):
        return list(self._load())

    def __contains__(self, k):
        return k in self._load()

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            links = node._load()
            if part not in links:
                raise KeyError(path)
            node = node.f._open(links[part])
        return node


class Dataset(_Obj):
    def read(self):
        dt = shape = layout = None
        for typ, data in self.msgs:
            if typ == 0x0001:
                shape = self.f._dataspace(data)
            elif typ == 0x0003:
                dt = self.f._datatype(data)
            elif typ == 0x0008:
                layout = data
            elif typ == 0x000B:
                raise Hdf5Error("filtered (compressed) datasets are not supported")
        if dt is None or shape is None or layout is None:
            raise Hdf5Error("dataset without datatype / dataspace / layout message")
        n = int(np.prod(shape)) if shape else 1
        ver, cls = layout[0], layout[1]
        if ver != 3:
            raise Hdf5Error(f"data layout message version {ver} is not supported")
        if cls == 0:
            size = struct.unpack_from("<H", layout, 2)[0]
            raw = bytes(layout[4:4 + size])
        elif cls == 1:
            addr, size = struct.unpack_from("<QQ", layout, 2)
            raw = b"" if addr == UNDEF else self.f._read(self.f.base + addr, size)
        else:
            raise Hdf5Error("chunked datasets are not supported")
        if len(raw) < n * dt.itemsize:  # never written: fill value (zeros)
            raw = raw + b"\0" * (n * dt.itemsize - len(raw))
        return np.frombuffer(raw, dt, n).reshape(shape).copy()

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a if dtype is None else a.astype(dtype)


class File(Group):
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        off = 0
        while True:
            if self.buf[off:off + 8] == SIG:
                break
            off = 512 if off == 0 else off * 2
            if off >= len(self.buf):
                raise Hdf5Error(f"Unable to open file (file signature not found): {path}")
        ver = self.buf[off + 8]
        if ver in (0, 1):
            so, sl = self.buf[off + 13], self.buf[off + 14]
            if (so, sl) != (8, 8):
                raise Hdf5Error("only 8-byte offsets and lengths are supported")
            p = off + 24 + (4 if ver == 1 else 0)
            self.base = struct.unpack_from("<Q", self.buf, p)[0]
            root = struct.unpack_from("<Q", self.buf, p + 32 + 8)[0]  # root symbol-table entry: name offset, header address
        elif ver in (2, 3):
            if (self.buf[off + 9], self.buf[off + 10]) != (8, 8):
                raise Hdf5Error("only 8-byte offsets and lengths are supported")
            self.base = struct.unpack_from("<Q", self.buf, off + 12)[0]
            root = struct.unpack_from("<Q", self.buf, off + 36)[0]
        else:
            raise Hdf5Error(f"superblock version {ver} is not supported")
        self.f = self
        super().__init__(self, root)

    # ---- low level
    def _read(self, addr, n):
        if addr + n > len(self.buf):
            raise Hdf5Error("truncated file")
        return self.buf[addr:addr + n]

    def _open(self, addr):
        msgs = self._object_messages(addr)
        kinds = {t for t, _ in msgs}
        return Dataset(self, addr) if 0x0008 in kinds else Group(self, addr)

    def _object_messages(self, addr):
        a = self.base + addr
        if self.buf[a:a + 4] == b"OHDR":
            raise Hdf5Error("version-2 object headers (libver='latest') are not supported")
        ver, _, nmsg, _refs, hsize = struct.unpack_from("<BBHII", self.buf, a)
        if ver != 1:
            raise Hdf5Error(f"object header version {ver} is not supported")
        blocks = [(a + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                typ, size, _flags = struct.unpack_from("<HHB", self.buf, p)
                data = self.buf[p + 8:p + 8 + size]
                p += 8 + size
                if typ == 0x0010:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((self.base + caddr, clen))
                out.append((typ, data))
        return out

    def _heap_name(self, heap_addr, off):
        h = self.base + heap_addr
        if self.buf[h:h + 4] != b"HEAP":
            raise Hdf5Error("bad local heap signature")
        data_addr = struct.unpack_from("<Q", self.buf, h + 24)[0]
        s = self.base + data_addr + off
        e = self.buf.index(b"\0", s)
        return self.buf[s:e].decode("utf-8")

    def _group_entries(self, btree, heap):
        out = {}
        a = self.base + btree
        if self.buf[a:a + 4] == b"SNOD":
            nodes = [btree]
        else:
            if self.buf[a:a + 4] != b"TREE":
                raise Hdf5Error("bad group B-tree signature")
            _typ, level, used = struct.unpack_from("<BBH", self.buf, a + 4)
            children = [struct.unpack_from("<Q", self.buf, a + 24 + 8 + i * 16)[0] for i in range(used)]
            if level > 0:
                for c in children:
                    out.update(self._group_entries(c, heap))
                return out
            nodes = children
        for nd in nodes:
            s = self.base + nd
            if self.buf[s:s + 4] != b"SNOD":
                raise Hdf5Error("bad symbol table node signature")
            nsym = struct.unpack_from("<H", self.buf, s + 6)[0]
            for i in range(nsym):
                name_off, hdr = struct.unpack_from("<QQ", self.buf, s + 8 + i * 40)
                out[self._heap_name(heap, name_off)] = hdr
        return out

    # ---- messages
    @staticmethod
    def _dataspace(d):
        ver, rank, flags = d[0], d[1], d[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if d[3] == 2:
                return (0,)  # null dataspace
            p = 4
        else:
            raise Hdf5Error(f"dataspace version {ver} is not supported")
        return tuple(struct.unpack_from("<" + "Q" * rank, d, p)) if rank else ()

    @staticmethod
    def _datatype(d):
        cls, bits0 = d[0] & 0x0F, d[1]
        size = struct.unpack_from("<I", d, 4)[0]
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{size}")
        if cls == 1:
            if size not in (2, 4, 8):
                raise Hdf5Error(f"{size}-byte floats are not supported")
            return np.dtype(f"{order}f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        if cls == 9:
            raise Hdf5Error("variable-length data")
        raise Hdf5Error(f"datatype class {cls} is not supported")

    def _attribute(self, d):
        ver = d[0]
        nsz, dsz, ssz = struct.unpack_from("<HHH", d, 2)
        pad = (lambda n: (n + 7) & ~7) if ver == 1 else (lambda n: n)
        p = 8 if ver in (1, 2) else 9
        if ver not in (1, 2, 3):
            raise Hdf5Error(f"attribute message version {ver} is not supported")
        name = bytes(d[p:p + nsz]).split(b"\0")[0].decode("utf-8")
        p += pad(nsz)
        dt_raw = d[p:p + dsz]
        p += pad(dsz)
        shape = self._dataspace(d[p:p + ssz])
        p += pad(ssz)
        try:
            dt = self._datatype(dt_raw)
        except Hdf5Error:
            return name, None  # e.g. variable-length strings (keras_version / backend in newer h5py): not needed
        n = int(np.prod(shape)) if shape else 1
        val = np.frombuffer(bytes(d[p:p + n * dt.itemsize]), dt, n).reshape(shape)
        return name, (val if shape else val[()])


# ======================================================================================= writer
class _W:
    def __init__(self):
        self.b = bytearray()

    def align(self, n=8):
        self.b += b"\0" * (-len(self.b) % n)

    def tell(self):
        return len(self.b)


def _dt_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        props = {4: struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127), 8: struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)}[dt.itemsize]
        bits = bytes([0x20, {4: 31, 8: 63}[dt.itemsize], 0])  # little-endian, implied msb normalisation, sign position
        return bytes([0x11]) + bits + struct.pack("<I", dt.itemsize) + props
    if dt.kind in "iu":
        return bytes([0x10, 0x08 if dt.kind == "i" else 0, 0, 0]) + struct.pack("<I", dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        return bytes([0x13, 0, 0, 0]) + struct.pack("<I", dt.itemsize)  # null-terminated ASCII
    raise Hdf5Error(f"cannot write dtype {dt}")


def _ds_msg(shape):
    return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def _msg(typ, data):
    data = bytes(data) + b"\0" * (-len(data) % 8)
    return struct.pack("<HHBBBB", typ, len(data), 0, 0, 0, 0) + data


def _attr_msg(name, value):
    a = np.asarray(value)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf-8")
    nm = name.encode() + b"\0"
    dt, ds = _dt_msg(a.dtype), _ds_msg(a.shape)
    p8 = lambda b: b + b"\0" * (-len(b) % 8)  # noqa: E731
    return _msg(0x000C, struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + p8(nm) + p8(dt) + p8(ds) + a.tobytes())


def _header(w, msgs):
    w.align()
    addr = w.tell()
    body = b"".join(msgs)
    w.b += struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\0" * 4 + body
    return addr


def _write_dataset(w, arr):
    """contiguous dataset; returns the object header address"""
    a = np.ascontiguousarray(arr)
    w.align()
    daddr = w.tell()
    w.b += a.tobytes()
    layout = struct.pack("<BBQQ", 3, 1, daddr, a.nbytes)
    return _header(w, [_msg(0x0001, _ds_msg(a.shape)), _msg(0x0003, _dt_msg(a.dtype)), _msg(0x0008, layout)])


def write_file(path, tree):
    """tree: nested dict; ndarray leaves are datasets, keys starting with '@' are attributes of the enclosing group."""
    w = _W()
    w.b += b"\0" * 96  # superblock, filled in last

    def rec(node):
        if isinstance(node, dict):
            kids = {k: (v if k.startswith("@") else _Placed(rec(v))) for k, v in node.items()}
            return _write_group(w, kids)
        return _write_dataset(w, node)
    root = rec(tree)
    eof = w.tell()
    sb = SIG + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4096, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root[0], 1, 0) + struct.pack("<QQ", root[1], root[2])
    w.b[:len(sb)] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(w.b))


class _Placed:
    def __init__(self, res):
        self.addr = res[0] if isinstance(res, tuple) else res


def _write_group(w, kids):
    node = {}
    placed = {}
    for k, v in kids.items():
        if isinstance(v, _Placed):
            placed[k] = v.addr
        else:
            node[k] = v
    names = sorted(placed, key=lambda s: s.encode())
    heap = bytearray(b"\0" * 8)
    offs = {}
    for k in names:
        offs[k] = len(heap)
        heap += k.encode() + b"\0"
        heap += b"\0" * (-len(heap) % 8)
    w.align()
    heap_data = w.tell()
    w.b += heap
    w.align()
    heap_addr = w.tell()
    w.b += b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(heap), UNDEF, heap_data)
    w.align()
    snod = w.tell()
    w.b += b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for k in names:
        w.b += struct.pack("<QQII", offs[k], placed[k], 0, 0) + b"\0" * 16
    w.align()
    tree = w.tell()
    last = offs[names[-1]] if names else 0
    w.b += b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod, last)
    msgs = [_msg(0x0011, struct.pack("<QQ", tree, heap_addr))]
    for k, v in node.items():
        msgs.append(_attr_msg(k[1:], v))
    return _header(w, msgs), tree, heap_addr
