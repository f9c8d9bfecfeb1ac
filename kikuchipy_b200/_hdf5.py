"""A small HDF5 reader and writer, enough for h5ebsd files (no h5py in this image, and none needed).

Reads the "earliest" file layout that h5py / the HDF5 library write by default and that every h5ebsd
file in the reference's test data uses: superblock version 0 or 1, version-1 object headers (with
continuation blocks), groups as symbol tables (version-1 B-tree + local heap + symbol-table nodes),
datasets with compact, contiguous or chunked (version-1 chunk B-tree) layout, the deflate, shuffle and
Fletcher-32 filters, fixed-point, floating-point and fixed-length string types, and variable-length
strings through the global heap.  Anything newer (superblock 2/3, version-2 object headers, fractal
heaps) raises ``NotImplementedError`` naming the structure.

Writes the same subset (contiguous datasets only), which is what ``save_h5ebsd`` needs and what the
tests use to build files.  Format: "HDF5 File Format Specification Version 3.0", sections II-IV.
"""

from __future__ import annotations

import mmap
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(IOError):
    pass


# ---------------------------------------------------------------------------------------------------
# reading
# ---------------------------------------------------------------------------------------------------

class _Buf:
    """File bytes (memory mapped) with little-endian field readers."""

    def __init__(self, path):
        self._f = open(path, "rb")
        try:
            self.m = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:
            self._f.close()
            raise Hdf5Error(f"{path!r} is empty")

    def close(self):
        try:
            self.m.close()
        except BufferError:  # arrays that view the mapping are still alive: the mapping stays until they go
            pass
        self._f.close()

    def u(self, off, size):
        return int.from_bytes(self.m[off:off + size], "little")


class Dataset:
    def __init__(self, file, name, shape, dtype, layout, filters, strtype):
        self._file, self.name, self.shape, self.dtype = file, name, tuple(shape), dtype
        self._layout, self._filters, self._strtype = layout, filters, strtype

    @property
    def chunks(self):
        return tuple(self._layout["chunk"][:-1]) if self._layout["class"] == 2 else None

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    def read(self, copy=True):
        """The whole dataset as an array (``copy=False``: a read-only view of the file mapping where
        the layout allows it - contiguous, unfiltered)."""
        f = self._file
        lay = self._layout
        n = self.size
        if self._strtype == "vlen":
            raw = self._raw_bytes(16 * n)
            out = np.empty(n, dtype=object)
            for i in range(n):
                length, addr, index = struct.unpack_from("<IQI", raw, 16 * i)
                out[i] = f._global_heap_object(addr, index)[:length] if length else b""
            return out.reshape(self.shape)
        if lay["class"] == 1 and not self._filters:
            if lay["addr"] == UNDEF or n == 0:
                return np.zeros(self.shape, self.dtype)
            a = np.frombuffer(f._b.m, dtype=self.dtype, count=n, offset=lay["addr"]).reshape(self.shape)
            return a.copy() if copy else a
        raw = self._raw_bytes(n * self.dtype.itemsize)
        return np.frombuffer(raw, dtype=self.dtype, count=n).reshape(self.shape).copy()

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, key):
        a = self.read()
        if isinstance(key, tuple) and len(key) == 0:
            return a if a.shape else a[()]
        return a[key]

    # -- storage -----------------------------------------------------------------------------------
    def _raw_bytes(self, nbytes):
        f, lay = self._file, self._layout
        if lay["class"] == 0:
            return bytes(lay["data"][:nbytes]).ljust(nbytes, b"\0")
        if lay["class"] == 1:
            if lay["addr"] == UNDEF:
                return bytes(nbytes)
            return bytes(f._b.m[lay["addr"]:lay["addr"] + nbytes])
        # chunked: assemble in element units
        item = lay["chunk"][-1]
        cshape = lay["chunk"][:-1]
        rank = len(cshape)
        full = np.zeros(self.shape, dtype=np.dtype((np.void, item)))
        if lay["addr"] != UNDEF:
            for offsets, mask, blob in f._chunks(lay["addr"], rank):
                data = self._unfilter(blob, mask, int(np.prod(cshape)) * item)
                chunk = np.frombuffer(data, dtype=full.dtype, count=int(np.prod(cshape))).reshape(cshape)
                sel_f, sel_c = [], []
                for d in range(rank):
                    lo = offsets[d]
                    hi = min(lo + cshape[d], self.shape[d])
                    if hi <= lo:
                        break
                    sel_f.append(slice(lo, hi))
                    sel_c.append(slice(0, hi - lo))
                else:
                    full[tuple(sel_f)] = chunk[tuple(sel_c)]
        return full.tobytes()[:nbytes]

    def _unfilter(self, blob, mask, expect):
        data = blob
        for i, (fid, cvals) in reversed(list(enumerate(self._filters))):
            if mask & (1 << i):
                continue
            if fid == 1:
                data = zlib.decompress(data)
            elif fid == 2:
                size = cvals[0] if cvals else self.dtype.itemsize
                n = len(data) // size
                a = np.frombuffer(data, np.uint8, n * size).reshape(size, n).T
                data = a.tobytes() + data[n * size:]
            elif fid == 3:
                data = data[:-4]
            else:
                raise NotImplementedError(f"HDF5 filter {fid} (dataset {self.name!r})")
        if len(data) < expect:
            data = bytes(data).ljust(expect, b"\0")
        return data


class Group:
    def __init__(self, file, name, btree, heap):
        self._file, self.name, self._btree, self._heap = file, name, btree, heap
        self._entries = None

    def _load(self):
        if self._entries is None:
            self._entries = dict(self._file._group_entries(self._btree, self._heap))
        return self._entries

    def keys(self):
        return list(self._load().keys())

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __iter__(self):
        return iter(self.keys())

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def get(self, path, default=None):
        try:
            return self[path]
        except KeyError:
            return default

    def __getitem__(self, path):
        node = self
        parts = [p for p in path.split("/") if p]
        if path.startswith("/"):
            node = self._file.root
        for i, part in enumerate(parts):
            if not isinstance(node, Group):
                raise KeyError(path)
            entries = node._load()
            if part not in entries:
                raise KeyError(f"{path!r}: no object {part!r} in group {node.name!r}")
            child_name = (node.name.rstrip("/") + "/" + part)
            node = self._file._object(entries[part], child_name)
        return node


class File(Group):
    """``File(path)[...]`` - groups behave like dictionaries, datasets have ``shape``, ``dtype``,
    ``chunks``, ``read()`` and ``[()]``."""

    def __init__(self, path):
        self.filename = str(path)
        self._b = _Buf(path)
        try:
            self._open()
        except Exception:
            self._b.close()
            raise

    def _open(self):
        b = self._b
        base = None
        off = 0
        while off + 8 <= len(b.m):  # the superblock may sit at 0, 512, 1024, ...
            if b.m[off:off + 8] == SIGNATURE:
                base = off
                break
            off = 512 if off == 0 else off * 2
        if base is None:
            raise Hdf5Error(f"{self.filename!r} is not an HDF5 file")
        version = b.m[base + 8]
        if version not in (0, 1):
            raise NotImplementedError(f"HDF5 superblock version {version} (only the 'earliest' layout, 0 / 1, is read)")
        so, sl = b.m[base + 13], b.m[base + 14]
        if so != 8 or sl != 8:
            raise NotImplementedError(f"HDF5 files with {so}-byte offsets / {sl}-byte lengths")
        p = base + 24 + (4 if version == 1 else 0)
        self._base = b.u(p, 8)
        # root symbol table entry follows base, free-space, end-of-file and driver-info addresses
        ent = p + 32
        self._objects = {}
        root_header = b.u(ent + 8, 8)
        Group.__init__(self, self, "/", None, None)
        self.root = self._object(root_header, "/")
        if not isinstance(self.root, Group):
            raise Hdf5Error("the root object is not a group")
        self._btree, self._heap = self.root._btree, self.root._heap

    def close(self):
        self._b.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    # -- object headers ------------------------------------------------------------------------------
    def _messages(self, addr):
        b = self._b
        addr += self._base
        if b.m[addr:addr + 4] == b"OHDR":
            raise NotImplementedError("version-2 object headers (file written with libver='latest')")
        if b.m[addr] != 1:
            raise Hdf5Error(f"object header version {b.m[addr]} at {addr}")
        n_msgs = b.u(addr + 2, 2)
        size = b.u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < n_msgs:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < n_msgs:
                mtype, msize, flags = b.u(p, 2), b.u(p + 2, 2), b.m[p + 4]
                body = p + 8
                if mtype == 0x10:  # continuation
                    blocks.append((b.u(body, 8) + self._base, b.u(body + 8, 8)))
                out.append((mtype, body, msize, flags))
                p = body + msize
        return out

    def _object(self, header_addr, name):
        if header_addr in self._objects:
            obj = self._objects[header_addr]
            return obj
        b = self._b
        msgs = self._messages(header_addr)
        kinds = {m[0] for m in msgs}
        if 0x11 in kinds:
            body = next(m[1] for m in msgs if m[0] == 0x11)
            obj = Group(self, name, b.u(body, 8), b.u(body + 8, 8))
        elif 0x02 in kinds or 0x06 in kinds:
            raise NotImplementedError("groups stored as link messages / fractal heaps (libver='latest')")
        elif 0x08 in kinds:
            obj = self._dataset(msgs, name)
        else:
            raise Hdf5Error(f"object {name!r} is neither a group nor a dataset")
        self._objects[header_addr] = obj
        return obj

    def _dataset(self, msgs, name):
        b = self._b
        shape, dtype, layout, filters, strtype = (), None, None, [], None
        for mtype, p, size, flags in msgs:
            if flags & 2:
                raise NotImplementedError(f"shared header messages (dataset {name!r})")
            if mtype == 0x01:
                ver, rank, fl = b.m[p], b.m[p + 1], b.m[p + 2]
                q = p + (8 if ver == 1 else 4)
                if ver not in (1, 2):
                    raise Hdf5Error(f"dataspace message version {ver}")
                if ver == 2 and b.m[p + 3] == 2:
                    raise NotImplementedError("null dataspaces")
                shape = tuple(b.u(q + 8 * i, 8) for i in range(rank))
            elif mtype == 0x03:
                dtype, strtype = self._datatype(p, name)
            elif mtype == 0x08:
                ver = b.m[p]
                if ver != 3:
                    raise NotImplementedError(f"data layout message version {ver} (dataset {name!r})")
                cls = b.m[p + 1]
                if cls == 0:
                    n = b.u(p + 2, 2)
                    layout = {"class": 0, "data": bytes(b.m[p + 4:p + 4 + n])}
                elif cls == 1:
                    a = b.u(p + 2, 8)
                    layout = {"class": 1, "addr": a if a == UNDEF else a + self._base, "size": b.u(p + 10, 8)}
                elif cls == 2:
                    rank = b.m[p + 2]
                    a = b.u(p + 3, 8)
                    layout = {"class": 2, "addr": a if a == UNDEF else a + self._base,
                              "chunk": [b.u(p + 11 + 4 * i, 4) for i in range(rank)]}
                else:
                    raise NotImplementedError(f"data layout class {cls}")
            elif mtype == 0x0B:
                ver, nf = b.m[p], b.m[p + 1]
                q = p + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid = b.u(q, 2)
                    if ver == 1 or fid >= 256:
                        nlen, ncv = b.u(q + 2, 2), b.u(q + 6, 2)
                        q += 8 + ((nlen + 7) // 8 * 8 if ver == 1 else nlen)
                    else:  # version 2, library filter: no name fields
                        ncv = b.u(q + 4, 2)
                        q += 6
                    cvals = [b.u(q + 4 * i, 4) for i in range(ncv)]
                    q += 4 * ncv
                    if ver == 1 and ncv % 2:
                        q += 4
                    filters.append((fid, cvals))
        if dtype is None or layout is None:
            raise Hdf5Error(f"dataset {name!r} lacks a datatype or layout message")
        return Dataset(self, name, shape, dtype, layout, filters, strtype)

    def _datatype(self, p, name):
        b = self._b
        cls, ver = b.m[p] & 0x0F, b.m[p] >> 4
        bits0 = b.m[p + 1]
        size = b.u(p + 4, 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{size}"), None
        if cls == 1:
            if size not in (2, 4, 8):
                raise NotImplementedError(f"{size}-byte floating-point type (dataset {name!r})")
            return np.dtype(f"{order}f{size}"), None
        if cls == 3:
            return np.dtype(f"S{size}"), "fixed"
        if cls == 9 and (bits0 & 0x0F) == 1:
            return np.dtype("O"), "vlen"
        if cls == 8:  # enumeration over an integer base type; h5py stores bool as {FALSE: 0, TRUE: 1} over int8
            base, _ = self._datatype(p + 8, name)
            n_members = b.u(p + 1, 2)
            if base.itemsize == 1 and n_members == 2:
                q = p + 8 + 8 + 4  # base type message: 8 bytes + 4 bytes of fixed-point properties
                names = bytes(b.m[q:q + 16])
                if names.startswith(b"FALSE\0") and b"TRUE\0" in names:
                    return np.dtype(np.bool_), None
            return base, None
        raise NotImplementedError(f"HDF5 datatype class {cls} version {ver} (dataset {name!r})")

    # -- groups --------------------------------------------------------------------------------------
    def _heap_string(self, heap_addr, offset):
        b = self._b
        h = heap_addr + self._base
        if b.m[h:h + 4] != b"HEAP":
            raise Hdf5Error("local heap signature missing")
        data = b.u(h + 24, 8) + self._base
        end = b.m.find(b"\0", data + offset)
        return b.m[data + offset:end].decode("utf-8", "replace")

    def _group_entries(self, btree, heap):
        b = self._b
        out = []

        def walk(addr):
            a = addr + self._base
            sig = b.m[a:a + 4]
            if sig == b"TREE":
                if b.m[a + 4] != 0:
                    raise Hdf5Error("group B-tree of the wrong node type")
                n = b.u(a + 6, 2)
                p = a + 24
                for i in range(n):
                    walk(b.u(p + 8 + 16 * i, 8))  # key, child, key, child, ..., key
            elif sig == b"SNOD":
                n = b.u(a + 6, 2)
                for i in range(n):
                    e = a + 8 + 40 * i
                    out.append((self._heap_string(heap, b.u(e, 8)), b.u(e + 8, 8)))
            else:
                raise Hdf5Error(f"unexpected structure {bytes(sig)!r} in a group B-tree")

        if btree is not None and btree != UNDEF:
            walk(btree)
        return out

    def _chunks(self, addr, rank):
        """(offsets, filter mask, raw bytes) of every stored chunk."""
        b = self._b
        stack = [addr]
        while stack:
            a = stack.pop()
            if b.m[a:a + 4] != b"TREE" or b.m[a + 4] != 1:
                raise Hdf5Error("chunk B-tree node signature missing")
            level, n = b.m[a + 5], b.u(a + 6, 2)
            key = 8 + 8 * (rank + 1)
            p = a + 24
            for i in range(n):
                k = p + i * (key + 8)
                nbytes, mask = b.u(k, 4), b.u(k + 4, 4)
                offs = [b.u(k + 8 + 8 * d, 8) for d in range(rank)]
                child = b.u(k + key, 8) + self._base
                if level == 0:
                    yield offs, mask, b.m[child:child + nbytes]
                else:
                    stack.append(child)

    def _global_heap_object(self, addr, index):
        b = self._b
        a = addr + self._base
        if b.m[a:a + 4] != b"GCOL":
            raise Hdf5Error("global heap collection signature missing")
        end = a + b.u(a + 8, 8)
        p = a + 16
        while p + 16 <= end:
            idx, size = b.u(p, 2), b.u(p + 8, 8)
            if idx == 0:
                break
            if idx == index:
                return bytes(b.m[p + 16:p + 16 + size])
            p += 16 + (size + 7) // 8 * 8
        raise Hdf5Error(f"global heap object {index} not found")


# ---------------------------------------------------------------------------------------------------
# writing
# ---------------------------------------------------------------------------------------------------

def _pad8(x: bytes) -> bytes:
    return x + b"\0" * (-len(x) % 8)


def _dtype_message(dt: np.dtype) -> bytes:
    if dt.kind in "iu":
        bits = (1 if dt.byteorder == ">" else 0) | (8 if dt.kind == "i" else 0)
        return struct.pack("<BBBBI", 0x10, bits, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f":
        spec = {2: (15, 10, 5, 0, 10, 15), 4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[dt.itemsize]
        sign, eloc, esize, mloc, msize, bias = spec
        bits0 = (1 if dt.byteorder == ">" else 0) | 0x20  # implied leading mantissa bit
        return (struct.pack("<BBBBI", 0x11, bits0, sign, 0, dt.itemsize)
                + struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, eloc, esize, mloc, msize, bias))
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0, 0, 0, dt.itemsize)  # null-terminated ASCII
    raise TypeError(f"cannot store dtype {dt} in an HDF5 file")


class Writer:
    """Build a file from nested dictionaries: ``Writer().write(path, {"group": {"dataset": array}})``.
    Strings become fixed-length byte strings (length + 1, like the reference's ``_dict2hdf5group``),
    scalars one-element datasets of shape ``(1,)``."""

    LEAF_K = 64  # symbols per symbol-table node: 2 K

    def __init__(self):
        self._chunks = []
        self._pos = 0

    def _alloc(self, data: bytes, align=8) -> int:
        pad = -self._pos % align
        if pad:
            self._chunks.append(b"\0" * pad)
            self._pos += pad
        addr = self._pos
        self._chunks.append(data)
        self._pos += len(data)
        return addr

    def _header(self, messages) -> int:
        body = b"".join(struct.pack("<HHBBBB", t, len(_pad8(m)), 0, 0, 0, 0) + _pad8(m) for t, m in messages)
        head = struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\0" * 4
        return self._alloc(head + body)

    def _dataset(self, value) -> int:
        if isinstance(value, str):
            raw = value.encode("latin-1", "replace")
            arr = np.array([raw], dtype=f"S{len(raw) + 1}")
        elif isinstance(value, bytes):
            arr = np.array([value], dtype=f"S{len(value) + 1}")
        else:
            arr = np.asarray(value)
            if arr.dtype == bool:
                arr = arr.astype(np.uint8)
            if arr.dtype.kind == "U":
                arr = np.char.encode(arr, "latin-1")
            if arr.ndim == 0:
                arr = arr.reshape(1)
        arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        data_addr = self._alloc(arr.tobytes()) if arr.size else UNDEF
        space = struct.pack("<BBBB4x", 1, arr.ndim, 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape)
        layout = struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes)
        fill = struct.pack("<BBBB", 2, 2, 2, 0)  # version 2, allocate late, write at allocation, undefined
        return self._header([(0x01, space), (0x03, _dtype_message(arr.dtype)), (0x05, fill), (0x08, layout)])

    def _group(self, mapping) -> int:
        children = {}
        for key, val in mapping.items():
            children[str(key)] = self._group(val) if isinstance(val, dict) else self._dataset(val)
        names = sorted(children)
        # local heap: offset 0 holds the empty string the B-tree's first key points at
        heap_data, offsets = bytearray(b"\0" * 8), {}
        for n in names:
            offsets[n] = len(heap_data)
            heap_data += _pad8(n.encode("utf-8") + b"\0")
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)  # one free block: (next = 1 "none", size)
        data_addr = self._alloc(bytes(heap_data))
        heap_addr = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, data_addr))
        per = 2 * self.LEAF_K
        nodes = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        if len(nodes) > 32:
            raise ValueError("too many objects in one group for this writer")
        snods = []
        for chunk in nodes:
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
            for n in chunk:
                body += struct.pack("<QQII16x", offsets[n], children[n], 0, 0)
            body += b"\0" * (40 * (per - len(chunk)))
            snods.append(self._alloc(body))
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(nodes) if names else 0, UNDEF, UNDEF)
        tree += struct.pack("<Q", 0)
        for chunk, addr in zip(nodes, snods):
            if names:
                tree += struct.pack("<QQ", addr, offsets[chunk[-1]])
        tree += b"\0" * (8 + 16 * 32 - (len(tree) - 24))  # room for 2 K = 32 children
        tree_addr = self._alloc(tree)
        return self._header([(0x11, struct.pack("<QQ", tree_addr, heap_addr))])

    def write(self, path, tree: dict):
        self._chunks, self._pos = [], 0
        self._alloc(b"\0" * 96)  # superblock, filled in last
        root = self._group(tree)
        eof = self._pos
        sb = SIGNATURE + struct.pack("<BBBBBBBxHHI", 0, 0, 0, 0, 0, 8, 8, self.LEAF_K, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII16x", 0, root, 0, 0)
        assert len(sb) == 96, len(sb)
        self._chunks[0] = sb
        with open(path, "wb") as f:
            for c in self._chunks:
                f.write(c)


def write(path, tree: dict):
    Writer().write(path, tree)
