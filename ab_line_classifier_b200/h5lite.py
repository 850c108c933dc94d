"""Minimal pure-Python HDF5 reader / writer for Keras weight files (h5py is not installed in this image).

The reference saves its models with ``save_model(model.model, '...h5')`` (src/train.py:139-144), restores them with
``load_model(path, compile=False)`` (src/predict.py:50) and transplants weights by layer name from a saved ``.h5``
(src/models/model_utils.py:6-23).  Those files are HDF5 written by h5py with default settings; this module implements
just the subset of the HDF5 file format (HDF5 File Format Specification, version 0/1 structures) such files use:

  reader  superblock v0/v1 (and v2/v3), object headers v1 (and v2 without dense storage), groups as symbol tables
          (B-tree v1 + local heap + SNOD) or compact link messages, contiguous / compact / chunked (unfiltered) dataset
          layouts, little-endian integer / IEEE-float / fixed-length-string / variable-length-string datatypes,
          attributes (message versions 1-3) incl. arrays of fixed-length strings (Keras ``layer_names`` /
          ``weight_names``) and global-heap strings (``model_config``).
  writer  superblock v0, symbol-table groups, contiguous datasets, fixed-length-string and numeric attributes: the layout
          Keras' legacy H5 format uses, so a file written here can be read by ``tf.keras`` ``load_weights`` where
          TensorFlow exists -- and it provides the fixtures for the reader's tests.

Compressed / filtered datasets, dense (fractal-heap) groups and big-endian types are reported as NotImplementedError.
The reader has been validated against files produced by the writer below and against hand-assembled structures only:
no file written by the HDF5 C library is available in this image.
"""
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


# --------------------------------------------------------------------------------------------------------- reader
class _Datatype:
    def __init__(self, cls, size, np_dtype=None, base=None, vlen_string=False):
        self.cls, self.size, self.np_dtype, self.base, self.vlen_string = cls, size, np_dtype, base, vlen_string


def _parse_datatype(buf, off=0):
    cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", buf, off)
    cls, ver = cv & 0x0F, cv >> 4
    if ver not in (1, 2, 3):
        raise H5FormatError(f"datatype version {ver}")
    if cls == 0:                      # fixed-point
        if b0 & 1:
            raise NotImplementedError("big-endian integers")
        signed = bool(b0 & 8)
        return _Datatype(cls, size, np.dtype(("<i" if signed else "<u") + str(size)))
    if cls == 1:                      # floating point
        if b0 & 1:
            raise NotImplementedError("big-endian floats")
        if size not in (2, 4, 8):
            raise NotImplementedError(f"{size}-byte float")
        return _Datatype(cls, size, np.dtype("<f" + str(size)))
    if cls == 3:                      # fixed-length string
        return _Datatype(cls, size, np.dtype("S" + str(size)))
    if cls == 9:                      # variable length: type in b0 low nibble (0 sequence, 1 string)
        base = _parse_datatype(buf, off + 8)
        return _Datatype(cls, size, None, base=base, vlen_string=(b0 & 0x0F) == 1)
    raise NotImplementedError(f"HDF5 datatype class {cls}")


def _parse_dataspace(buf, off=0):
    ver, rank, flags = struct.unpack_from("<BBB", buf, off)
    if ver == 1:
        p = off + 8
    elif ver == 2:
        p = off + 4
        if buf[off + 3] == 2:         # null dataspace
            return None
    else:
        raise H5FormatError(f"dataspace version {ver}")
    return tuple(struct.unpack_from("<" + "Q" * rank, buf, p)) if rank else ()


class H5Object:
    """A group or dataset: ``attrs`` (dict), ``keys()`` / ``[name]`` for groups, ``read()`` / ``shape`` for datasets."""

    def __init__(self, f, addr, name="/"):
        self._f, self.addr, self.name = f, addr, name
        self.attrs = {}
        self._links = None            # name -> object header address (groups)
        self._stab = None
        self._dtype = self._shape = self._layout = None
        self._filtered = False
        self._parse_header()

    # ---- object header
    def _messages(self):
        buf, a = self._f.buf, self.addr
        if buf[a:a + 4] == b"OHDR":   # version 2 header
            flags = buf[a + 5]
            p = a + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            csize = (1, 2, 4, 8)[flags & 3]
            chunk0 = int.from_bytes(buf[p:p + csize], "little")
            p += csize
            blocks = [(p, p + chunk0)]
            track = bool(flags & 0x04)
            while blocks:
                p, end = blocks.pop(0)
                while p + 4 <= end:
                    mtype = buf[p]
                    msize = struct.unpack_from("<H", buf, p + 1)[0]
                    p += 4 + (2 if track else 0)
                    if mtype == 0x10:
                        o, ln = struct.unpack_from("<QQ", buf, p)
                        blocks.append((o + 4, o + ln - 4))       # skip 'OCHK', drop the checksum
                    else:
                        yield mtype, buf[p:p + msize]
                    p += msize
            return
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", buf, a)
        if ver != 1:
            raise H5FormatError(f"object header version {ver} at {a}")
        blocks = [(a + 16, a + 16 + hsize)]
        seen = 0
        while blocks and seen < nmsg:
            p, end = blocks.pop(0)
            while p + 8 <= end and seen < nmsg:
                mtype, msize = struct.unpack_from("<HH", buf, p)
                data = buf[p + 8:p + 8 + msize]
                p += 8 + msize
                seen += 1
                if mtype == 0x10:
                    o, ln = struct.unpack_from("<QQ", data, 0)
                    blocks.append((o, o + ln))
                else:
                    yield mtype, data

    def _parse_header(self):
        for mtype, d in self._messages():
            if mtype == 0x01:
                self._shape = _parse_dataspace(d)
            elif mtype == 0x03:
                self._dtype = _parse_datatype(d)
            elif mtype == 0x08:
                self._layout = bytes(d)
            elif mtype == 0x0B:
                self._filtered = True
            elif mtype == 0x0C:
                k, v = self._parse_attribute(d)
                self.attrs[k] = v
            elif mtype == 0x11:
                self._stab = struct.unpack_from("<QQ", d, 0)
            elif mtype == 0x06:       # link message (compact new-style group)
                k, a = self._parse_link(d)
                if self._links is None:
                    self._links = {}
                if a is not None:
                    self._links[k] = a
            elif mtype == 0x02 and len(d) >= 2:      # link info: dense storage?
                flags = d[1]
                p = 2 + (8 if flags & 1 else 0)
                fheap = struct.unpack_from("<Q", d, p)[0]
                if fheap != UNDEF:
                    raise NotImplementedError("dense (fractal heap) HDF5 groups")
                if self._links is None:
                    self._links = {}

    def _parse_link(self, d):
        ver, flags = d[0], d[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = d[p]; p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        lsz = (1, 2, 4, 8)[flags & 3]
        nlen = int.from_bytes(d[p:p + lsz], "little"); p += lsz
        name = bytes(d[p:p + nlen]).decode("utf8"); p += nlen
        if ltype != 0:
            return name, None         # soft / external links are not followed
        return name, struct.unpack_from("<Q", d, p)[0]

    def _parse_attribute(self, d):
        ver = d[0]
        nsz, tsz, ssz = struct.unpack_from("<HHH", d, 2)
        if ver == 1:
            p = 8
            pad = lambda n: (n + 7) // 8 * 8   # noqa: E731
        elif ver in (2, 3):
            p = 8 + (1 if ver == 3 else 0)
            pad = lambda n: n                  # noqa: E731
        else:
            raise H5FormatError(f"attribute message version {ver}")
        name = bytes(d[p:p + nsz]).split(b"\0", 1)[0].decode("utf8"); p += pad(nsz)
        dt = _parse_datatype(d, p); p += pad(tsz)
        shape = _parse_dataspace(d, p); p += pad(ssz)
        return name, self._f._decode(dt, shape, d[p:])

    # ---- group interface
    def _load_links(self):
        if self._links is not None and self._stab is None:
            return self._links
        links = dict(self._links or {})
        if self._stab is not None:
            btree, heap = self._stab
            heap_data = self._f._local_heap(heap)
            for off, addr in self._f._walk_group_btree(btree):
                nm = bytes(heap_data[off:heap_data.index(b"\0", off)]).decode("utf8")
                links[nm] = addr
        self._links = links
        self._stab = None
        return links

    @property
    def is_group(self):
        return self._layout is None

    def keys(self):
        return sorted(self._load_links()) if self.is_group else []

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        obj = self
        for part in [p for p in path.split("/") if p]:
            links = obj._load_links() if obj.is_group else {}
            if part not in links:
                raise KeyError(f"{path!r} not found under {self.name!r}")
            obj = H5Object(self._f, links[part], (obj.name.rstrip("/") + "/" + part))
        return obj

    def visit_datasets(self, prefix=""):
        """(path, dataset) for every dataset below this group, depth first in name order."""
        for k in self.keys():
            child = self[k]
            if child.is_group:
                yield from child.visit_datasets(prefix + k + "/")
            else:
                yield prefix + k, child

    # ---- dataset interface
    @property
    def shape(self):
        return self._shape

    def read(self):
        if self.is_group:
            raise TypeError(f"{self.name} is a group")
        if self._filtered:
            raise NotImplementedError(f"{self.name}: filtered (compressed) datasets are not supported")
        dt, shape, lay = self._dtype, self._shape, self._layout
        n = int(np.prod(shape)) if shape else 1
        nbytes = n * dt.size
        buf = self._f.buf
        ver = lay[0]
        if ver == 3:
            cls = lay[1]
            if cls == 0:
                size = struct.unpack_from("<H", lay, 2)[0]
                raw = lay[4:4 + size]
            elif cls == 1:
                addr, size = struct.unpack_from("<QQ", lay, 2)
                raw = b"\0" * nbytes if addr == UNDEF else buf[addr:addr + nbytes]
            elif cls == 2:
                rank = lay[2]
                btree = struct.unpack_from("<Q", lay, 3)[0]
                cdims = struct.unpack_from("<" + "I" * rank, lay, 11)
                raw = self._read_chunked(btree, cdims[:-1], dt, shape)
            else:
                raise NotImplementedError(f"layout class {cls}")
        elif ver in (1, 2):
            rank, cls = lay[1], lay[2]
            p = 8
            addr = None
            if cls != 0:
                addr = struct.unpack_from("<Q", lay, p)[0]; p += 8
            dims = struct.unpack_from("<" + "I" * rank, lay, p); p += 4 * rank
            if cls == 0:
                size = struct.unpack_from("<I", lay, p)[0]
                raw = lay[p + 4:p + 4 + size]
            elif cls == 1:
                raw = buf[addr:addr + nbytes]
            else:
                raw = self._read_chunked(addr, dims[:-1], dt, shape)
        else:
            raise NotImplementedError(f"data layout message version {ver}")
        return self._f._decode(dt, shape, raw)

    def _read_chunked(self, btree, cdims, dt, shape):
        out = np.zeros(shape, dtype=np.dtype("V" + str(dt.size)))
        rank = len(shape)
        for offs, addr, size in self._f._walk_chunk_btree(btree, rank):
            chunk = np.frombuffer(self._f.buf[addr:addr + size], dtype=out.dtype).reshape(cdims)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
            out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out.tobytes()


class H5File(H5Object):
    """``H5File(path)`` or ``H5File(data=bytes)``: the root group of an HDF5 file (read-only)."""

    def __init__(self, path=None, data=None):
        if data is None:
            with open(path, "rb") as fh:
                data = fh.read()
        self.buf = memoryview(data)
        base = 0
        while bytes(self.buf[base:base + 8]) != SIGNATURE:       # the superblock may sit at 0, 512, 1024, ...
            base = 512 if base == 0 else base * 2
            if base + 8 > len(self.buf):
                raise H5FormatError("not an HDF5 file (signature not found)")
        ver = self.buf[base + 8]
        if ver in (0, 1):
            if self.buf[base + 13] != 8 or self.buf[base + 14] != 8:
                raise NotImplementedError("only 8-byte offsets / lengths are supported")
            p = base + 24 + (4 if ver == 1 else 0)
            p += 32                                              # base, free-space, EOF, driver addresses
            root = struct.unpack_from("<Q", self.buf, p + 8)[0]  # symbol table entry: name offset, header address
            stab = None
            if struct.unpack_from("<I", self.buf, p + 16)[0] == 1:
                stab = struct.unpack_from("<QQ", self.buf, p + 24)
        elif ver in (2, 3):
            if self.buf[base + 9] != 8 or self.buf[base + 10] != 8:
                raise NotImplementedError("only 8-byte offsets / lengths are supported")
            root = struct.unpack_from("<Q", self.buf, base + 12 + 24)[0]
            stab = None
        else:
            raise H5FormatError(f"superblock version {ver}")
        super().__init__(self, root, "/")
        if self._stab is None and stab is not None and self._links is None:
            self._stab = stab

    # ---- shared structures
    def _local_heap(self, addr):
        if bytes(self.buf[addr:addr + 4]) != b"HEAP":
            raise H5FormatError("local heap signature")
        size, _, data_addr = struct.unpack_from("<QQQ", self.buf, addr + 8)
        return bytes(self.buf[data_addr:data_addr + size])

    def _walk_group_btree(self, addr):
        """(heap offset of the link name, object header address) of every entry below a group B-tree node."""
        if bytes(self.buf[addr:addr + 4]) == b"SNOD":
            n = struct.unpack_from("<H", self.buf, addr + 6)[0]
            for i in range(n):
                yield struct.unpack_from("<QQ", self.buf, addr + 8 + 40 * i)
            return
        if bytes(self.buf[addr:addr + 4]) != b"TREE":
            raise H5FormatError("B-tree node signature")
        ntype, level, used = struct.unpack_from("<BBH", self.buf, addr + 4)
        if ntype != 0:
            raise H5FormatError("expected a group B-tree")
        p = addr + 24
        for i in range(used):
            child = struct.unpack_from("<Q", self.buf, p + 8 + 16 * i)[0]
            yield from self._walk_group_btree(child)

    def _walk_chunk_btree(self, addr, rank):
        if addr == UNDEF:
            return
        if bytes(self.buf[addr:addr + 4]) != b"TREE":
            raise H5FormatError("chunk B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", self.buf, addr + 4)
        if ntype != 1:
            raise H5FormatError("expected a chunk B-tree")
        ksz = 8 + 8 * (rank + 1)
        p = addr + 24
        for i in range(used):
            size, fmask = struct.unpack_from("<II", self.buf, p)
            offs = struct.unpack_from("<" + "Q" * rank, self.buf, p + 8)
            child = struct.unpack_from("<Q", self.buf, p + ksz)[0]
            if level == 0:
                if fmask:
                    raise NotImplementedError("filtered chunks")
                yield offs, child, size
            else:
                yield from self._walk_chunk_btree(child, rank)
            p += ksz + 8

    def _global_heap_object(self, addr, index):
        if bytes(self.buf[addr:addr + 4]) != b"GCOL":
            raise H5FormatError("global heap signature")
        csize = struct.unpack_from("<Q", self.buf, addr + 8)[0]
        p, end = addr + 16, addr + csize
        while p + 16 <= end:
            idx, _, _, osize = struct.unpack_from("<HHIQ", self.buf, p)
            if idx == 0:
                break
            if idx == index:
                return bytes(self.buf[p + 16:p + 16 + osize])
            p += 16 + (osize + 7) // 8 * 8
        raise H5FormatError("global heap object not found")

    def _decode(self, dt, shape, raw):
        if shape is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        if dt.cls == 9:
            if not dt.vlen_string:
                raise NotImplementedError("variable-length sequences")
            out = []
            for i in range(n):
                ln, addr, idx = struct.unpack_from("<IQI", raw, 16 * i)
                out.append(self._global_heap_object(addr, idx)[:ln] if addr not in (0, UNDEF) else b"")
            out = [o.decode("utf8", "replace") for o in out]
            return out[0] if shape == () else np.array(out, dtype=object).reshape(shape)
        arr = np.frombuffer(bytes(raw[:n * dt.size]), dtype=dt.np_dtype, count=n)
        if shape == ():
            v = arr[0]
            return bytes(v).rstrip(b"\0") if dt.cls == 3 else v
        return arr.reshape(shape).copy()


# --------------------------------------------------------------------------------------------------------- writer
def _dt_bytes(dtype):
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        size = dtype.itemsize
        exp, man = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[size]
        bias = {2: 15, 4: 127, 8: 1023}[size]
        # class 1, version 1; bit field: little-endian, mantissa normalisation 2 (implied msb), sign at the top bit
        return struct.pack("<BBBBI", 0x11, 0x20, size * 8 - 1, 0, size) + struct.pack(
            "<HHBBBBI", 0, size * 8, man, exp, 0, man, bias)
    if dtype.kind in "iu":
        return struct.pack("<BBBBI", 0x10, 0x08 if dtype.kind == "i" else 0, 0, 0, dtype.itemsize) + struct.pack(
            "<HH", 0, dtype.itemsize * 8)
    if dtype.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dtype.itemsize)      # null-terminated ASCII
    raise NotImplementedError(f"cannot write dtype {dtype}")


def _ds_bytes(shape):
    if shape == ():
        return struct.pack("<BBB5x", 1, 0, 0)
    return struct.pack("<BBB5x", 1, len(shape), 0) + struct.pack("<" + "Q" * len(shape), *shape)


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _attr_msg(name, value):
    if isinstance(value, str):
        value = value.encode("utf8")
    if isinstance(value, bytes):
        arr = np.array(value, dtype="S" + str(max(1, len(value))))
    else:
        arr = np.asarray(value)
        if arr.dtype.kind == "U":
            arr = np.char.encode(arr, "utf8")
        if arr.dtype.kind == "O":
            arr = np.array([v if isinstance(v, bytes) else str(v).encode("utf8") for v in arr.ravel()]).reshape(arr.shape)
    nm = name.encode("utf8") + b"\0"
    dt, ds = _dt_bytes(arr.dtype), _ds_bytes(arr.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + arr.tobytes()
    return _msg(0x0C, body)


class H5Writer:
    """``w = H5Writer(); g = w.group(w.root, 'name'); w.dataset(g, 'x', array); w.attr(g, 'k', v); w.save(path)``."""

    LEAF_K = 4          # symbol-table nodes hold up to 2K entries (HDF5 default)

    def __init__(self):
        self.root = {"kind": "group", "children": {}, "attrs": []}

    def group(self, parent, name):
        g = parent["children"].get(name)
        if g is None:
            g = parent["children"][name] = {"kind": "group", "children": {}, "attrs": []}
        return g

    def dataset(self, parent, name, array):
        parent["children"][name] = {"kind": "dataset", "data": np.ascontiguousarray(array), "attrs": []}
        return parent["children"][name]

    def attr(self, node, name, value):
        node["attrs"].append((name, value))

    def create_path(self, path, array):
        """Dataset at a '/'-separated path below the root, intermediate groups created on the way."""
        parts = [p for p in path.split("/") if p]
        g = self.root
        for p in parts[:-1]:
            g = self.group(g, p)
        return self.dataset(g, parts[-1], array)

    # ---- serialisation
    def _alloc(self, data):
        data = _pad8(bytes(data))
        addr = len(self._out)
        self._out += data
        return addr

    def _object_header(self, messages):
        body = b"".join(messages)
        return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body

    def _emit(self, node):
        attrs = [_attr_msg(k, v) for k, v in node["attrs"]]
        if node["kind"] == "dataset":
            a = node["data"]
            if a.dtype.byteorder == ">":
                a = a.astype(a.dtype.newbyteorder("<"))
            daddr = self._alloc(a.tobytes()) if a.nbytes else UNDEF
            msgs = [_msg(0x01, _ds_bytes(a.shape)), _msg(0x03, _dt_bytes(a.dtype), 1),
                    _msg(0x08, struct.pack("<BBQQ", 3, 1, daddr, a.nbytes))] + attrs
            return self._alloc(self._object_header(msgs))
        names = sorted(node["children"], key=lambda s: s.encode("utf8"))
        child_addr = {n: self._emit(node["children"][n]) for n in names}
        heap = bytearray(b"\0" * 8)                  # offset 0 = empty string (the first B-tree key)
        name_off = {}
        for n in names:
            name_off[n] = len(heap)
            heap += _pad8(n.encode("utf8") + b"\0")
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 16)            # one free block: (next = 1 -> none, size 16)
        heap_data = self._alloc(heap)
        heap_addr = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data))
        per = 2 * self.LEAF_K
        chunks = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        snods, keys = [], [0]
        for ch in chunks:
            ent = b"".join(struct.pack("<QQII16x", name_off[n], child_addr[n], 0, 0) for n in ch)
            ent += b"\0" * (40 * (per - len(ch)))
            snods.append(self._alloc(b"SNOD" + struct.pack("<BBH", 1, 0, len(ch)) + ent))
            keys.append(name_off[ch[-1]] if ch else 0)
        if len(snods) > 32:
            raise NotImplementedError("more than 256 links in one group")
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
        for i, s in enumerate(snods):
            tree += struct.pack("<QQ", keys[i], s)
        tree += struct.pack("<Q", keys[-1])
        tree += b"\0" * (16 * (32 - len(snods)))      # room for 2K = 32 children (internal K = 16)
        btree_addr = self._alloc(tree)
        node["_stab"] = (btree_addr, heap_addr)
        return self._alloc(self._object_header([_msg(0x11, struct.pack("<QQ", btree_addr, heap_addr))] + attrs))

    def tobytes(self):
        self._out = bytearray(b"\0" * 96)            # superblock v0 with 8-byte offsets: 56 + 40 bytes
        root_addr = self._emit(self.root)
        btree, heap = self.root["_stab"]
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self._out), UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree, heap)
        self._out[:len(sb)] = sb
        return bytes(self._out)

    def save(self, path):
        with open(path, "wb") as fh:
            fh.write(self.tobytes())
