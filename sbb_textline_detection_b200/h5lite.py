"""Minimal read-only HDF5 parser (pure Python + numpy) -- just enough of the file format to read the
Keras ``.h5`` model files the reference loads with ``keras.models.load_model`` (main.py:58-60, 221;
SURVEY.md section 8(f) rank 2).  This image has neither h5py nor libhdf5.

Supported (what h5py 2.x / HDF5 1.8-1.10 write by default, plus the common variations):
  superblock v0/v1/v2/v3 (with a user block / non-zero base address)
  object headers v1 and v2, continuation blocks
  old-style groups (symbol table: B-tree v1 + local heap + SNOD) and compact new-style groups
    (link messages); dense groups (fractal heap) raise NotImplementedError
  datasets: contiguous, compact and chunked (B-tree v1) layouts; deflate + shuffle (+fletcher32) filters
  datatypes: integers, IEEE floats, fixed-length strings, variable-length strings (global heap)
  attributes v1/v2/v3 stored in the object header (attributes moved to dense storage are reported missing)

Follows the published "HDF5 File Format Specification Version 3.0" (The HDF Group); no code of any
HDF5 implementation was consulted.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
_SIG = b"\x89HDF\r\n\x1a\n"


class H5Error(Exception):
    pass


class _Type:
    """Decoded datatype message."""

    def __init__(self, cls, size, dtype=None, vlen_str=False, base=None, strpad=0):
        self.cls, self.size, self.dtype, self.vlen_str, self.base, self.strpad = cls, size, dtype, vlen_str, base, strpad


class File:
    def __init__(self, path_or_bytes):
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            self.buf = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, "rb") as f:
                self.buf = f.read()
        self._parse_superblock()
        self.root = Group(self, self._root_addr, "/")

    # ------------------------------------------------------------------ low level
    def _u(self, off, n):
        return int.from_bytes(self.buf[off:off + n], "little")

    def _addr(self, off):
        """Read a file offset field; returns the absolute buffer position (base address applied)."""
        v = self._u(off, self.O)
        if v == (1 << (8 * self.O)) - 1:
            return UNDEF
        return v + self.base

    def _len(self, off):
        return self._u(off, self.L)

    def _parse_superblock(self):
        pos = 0
        while True:  # the signature sits at 0, 512, 1024, 2048, ... (user block)
            if self.buf[pos:pos + 8] == _SIG:
                break
            pos = 512 if pos == 0 else pos * 2
            if pos + 8 > len(self.buf):
                raise H5Error("not an HDF5 file (signature not found)")
        ver = self.buf[pos + 8]
        self.base = 0
        if ver in (0, 1):
            self.O, self.L = self.buf[pos + 13], self.buf[pos + 14]
            p = pos + 24 + (4 if ver == 1 else 0)
            base = self._u(p, self.O)
            self.base = base if base != 0 else 0
            # a file with a user block may record base address 0 and absolute offsets relative to the
            # signature position; HDF5 itself records the user-block size as the base address
            if base == 0 and pos != 0:
                self.base = pos
            p += 4 * self.O  # base, free-space info, end of file, driver info
            # root group symbol table entry: link name offset, object header address, cache type, ...
            self._root_addr = self._addr(p + self.O)
        elif ver in (2, 3):
            self.O, self.L = self.buf[pos + 9], self.buf[pos + 10]
            p = pos + 12
            base = self._u(p, self.O)
            self.base = base if base != 0 else pos
            self._root_addr = self._addr(p + 3 * self.O)
        else:
            raise H5Error(f"unsupported superblock version {ver}")

    # ------------------------------------------------------------------ object headers
    def _messages(self, addr):
        """Yield (type, flags, data_offset, size) for every header message of the object at addr."""
        b = self.buf
        if b[addr:addr + 4] == b"OHDR":
            yield from self._messages_v2(addr)
            return
        if b[addr] != 1:
            raise H5Error(f"unsupported object header version {b[addr]} at {addr}")
        nmsg = self._u(addr + 2, 2)
        size = self._u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        seen = 0
        while blocks and seen < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and seen < nmsg:
                mtype, msize, mflags = self._u(p, 2), self._u(p + 2, 2), b[p + 4]
                d = p + 8
                seen += 1
                if mtype == 0x10:
                    blocks.append((self._addr(d), self._len(d + self.O)))
                else:
                    yield mtype, mflags, d, msize
                p = d + msize

    def _messages_v2(self, addr):
        b = self.buf
        flags = b[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        nsz = 1 << (flags & 3)
        size0 = self._u(p, nsz)
        p += nsz
        blocks = [(p, size0)]
        track = bool(flags & 0x04)
        while blocks:
            p, n = blocks.pop(0)
            end = p + n
            while p + 4 <= end:
                mtype, msize, mflags = b[p], self._u(p + 1, 2), b[p + 3]
                d = p + 4 + (2 if track else 0)
                if d + msize > end:
                    break
                if mtype == 0x10:
                    ca, cl = self._addr(d), self._len(d + self.O)
                    if b[ca:ca + 4] != b"OCHK":
                        raise H5Error("bad object header continuation block")
                    blocks.append((ca + 4, cl - 8))
                elif mtype != 0:
                    yield mtype, mflags, d, msize
                p = d + msize

    # ------------------------------------------------------------------ message decoders
    def _dataspace(self, d):
        b = self.buf
        ver, rank, flags = b[d], b[d + 1], b[d + 2]
        if ver == 1:
            p = d + 8
        elif ver == 2:
            if b[d + 3] == 2:  # null dataspace
                return None
            p = d + 4
        else:
            raise H5Error(f"dataspace version {ver}")
        return tuple(self._len(p + i * self.L) for i in range(rank))

    def _datatype(self, d):
        b = self.buf
        cls, ver = b[d] & 15, b[d] >> 4
        bits = b[d + 1] | (b[d + 2] << 8) | (b[d + 3] << 16)
        size = self._u(d + 4, 4)
        if cls == 0:
            order = ">" if bits & 1 else "<"
            kind = "i" if bits & 8 else "u"
            return _Type(cls, size, np.dtype(f"{order}{kind}{size}"))
        if cls == 1:
            order = ">" if bits & 1 else "<"
            if size not in (2, 4, 8):
                raise H5Error(f"float size {size}")
            return _Type(cls, size, np.dtype(f"{order}f{size}"))
        if cls == 3:
            return _Type(cls, size, np.dtype(f"S{size}"), strpad=bits & 15)
        if cls == 9:
            base = self._datatype(d + 8)
            return _Type(cls, size, None, vlen_str=(bits & 15) == 1, base=base)
        if cls == 7:  # object reference: keep the raw address bytes
            return _Type(cls, size, np.dtype(f"V{size}"))
        if cls == 8:  # enum: decode as its base integer type
            return self._datatype(d + 8)
        raise NotImplementedError(f"HDF5 datatype class {cls}")

    def _read_vlen(self, raw, n, t):
        out = []
        esz = 4 + self.O + 4
        for i in range(n):
            p = i * esz
            ln = int.from_bytes(raw[p:p + 4], "little")
            addr = int.from_bytes(raw[p + 4:p + 4 + self.O], "little")
            idx = int.from_bytes(raw[p + 4 + self.O:p + esz], "little")
            if ln == 0 or addr == 0:
                out.append(b"" if t.vlen_str else np.zeros(0, t.base.dtype))
                continue
            data = self._global_heap_object(addr + self.base, idx)
            if t.vlen_str:
                out.append(data[:ln])
            else:
                out.append(np.frombuffer(data[:ln * t.base.size], t.base.dtype).copy())
        return out

    def _global_heap_object(self, addr, idx):
        b = self.buf
        if b[addr:addr + 4] != b"GCOL":
            raise H5Error("bad global heap collection")
        size = self._len(addr + 8)
        p = addr + 8 + self.L
        end = addr + size
        while p + 8 + self.L <= end:
            oi = self._u(p, 2)
            osz = self._len(p + 8)
            if oi == 0:
                break
            if oi == idx:
                return b[p + 8 + self.L:p + 8 + self.L + osz]
            p += 8 + self.L + ((osz + 7) & ~7)
        raise H5Error(f"global heap object {idx} not found")

    def _decode(self, raw, t, shape):
        n = 1
        for s in shape or ():
            n *= s
        if t.cls == 9:
            vals = self._read_vlen(raw, n, t)
            if shape == ():
                return vals[0]
            arr = np.empty(n, object)
            for i, v in enumerate(vals):
                arr[i] = v
            return arr.reshape(shape)
        arr = np.frombuffer(raw, t.dtype, count=n).reshape(shape if shape is not None else ())
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        return arr.copy()

    def _attribute(self, d, size):
        b = self.buf
        ver = b[d]
        nsz, tsz, ssz = self._u(d + 2, 2), self._u(d + 4, 2), self._u(d + 6, 2)
        if ver == 1:
            pad = lambda v: (v + 7) & ~7
            p = d + 8
        elif ver == 2:
            pad = lambda v: v
            p = d + 8
        elif ver == 3:
            pad = lambda v: v
            p = d + 9
        else:
            raise H5Error(f"attribute version {ver}")
        name = b[p:p + nsz].split(b"\0")[0].decode("utf-8", "replace")
        p += pad(nsz)
        t = self._datatype(p)
        p += pad(tsz)
        shape = self._dataspace(p)
        p += pad(ssz)
        if shape is None:
            return name, None
        raw = b[p:d + size]
        return name, self._decode(raw, t, shape)


class _Object:
    def __init__(self, f: File, addr: int, name: str):
        self._f, self._addr, self.name = f, addr, name
        self._attrs = None
        self.has_dense_attrs = False

    @property
    def attrs(self) -> dict:
        if self._attrs is None:
            out = {}
            for mtype, _, d, size in self._f._messages(self._addr):
                if mtype == 0x0C:
                    k, v = self._f._attribute(d, size)
                    out[k] = v
                elif mtype == 0x15:
                    fheap = self._f._addr(d + 2 + (2 if self._f.buf[d + 1] & 1 else 0))
                    if fheap != UNDEF:
                        # dense attribute storage (fractal heap): not parsed.  Attributes kept there (HDF5 moves
                        # an attribute out of the header when it outgrows 64 KB) are reported as missing.
                        self.has_dense_attrs = True
            self._attrs = out
        return self._attrs


class Group(_Object):
    def __init__(self, f, addr, name):
        super().__init__(f, addr, name)
        self._links = None

    def _load(self):
        f = self._f
        links = {}
        for mtype, _, d, size in f._messages(self._addr):
            if mtype == 0x11:
                btree, heap = f._addr(d), f._addr(d + f.O)
                hb = f.buf
                if hb[heap:heap + 4] != b"HEAP":
                    raise H5Error("bad local heap")
                data_seg = f._addr(heap + 8 + 2 * f.L)
                self._walk_group_btree(btree, data_seg, links)
            elif mtype == 0x06:
                b = f.buf
                flags = b[d + 1]
                p = d + 2
                ltype = 0
                if flags & 0x08:
                    ltype = b[p]
                    p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                lsz = 1 << (flags & 3)
                ln = f._u(p, lsz)
                p += lsz
                nm = b[p:p + ln].decode("utf-8", "replace")
                p += ln
                if ltype == 0:
                    links[nm] = f._addr(p)
            elif mtype == 0x02:
                b = f.buf
                flags = b[d + 1]
                p = d + 2 + (8 if flags & 1 else 0)
                if f._addr(p) != UNDEF:
                    raise NotImplementedError("dense link storage (fractal heap) is not supported")
        self._links = links

    def _walk_group_btree(self, addr, data_seg, links):
        f, b = self._f, self._f.buf
        if b[addr:addr + 4] == b"SNOD":
            n = f._u(addr + 6, 2)
            p = addr + 8
            esz = 2 * f.O + 24
            for _ in range(n):
                noff = f._u(p, f.O)
                oaddr = f._addr(p + f.O)
                s = data_seg + noff
                e = b.index(b"\0", s)
                links[b[s:e].decode("utf-8", "replace")] = oaddr
                p += esz
            return
        if b[addr:addr + 4] != b"TREE":
            raise H5Error("bad group B-tree node")
        n = f._u(addr + 6, 2)
        p = addr + 8 + 2 * f.O
        for i in range(n):
            child = f._addr(p + f.L + i * (f.L + f.O))
            self._walk_group_btree(child, data_seg, links)

    def keys(self):
        if self._links is None:
            self._load()
        return list(self._links.keys())

    def __contains__(self, k):
        if self._links is None:
            self._load()
        return k.strip("/").split("/")[0] in self._links

    def __getitem__(self, path):
        if self._links is None:
            self._load()
        parts = [p for p in path.split("/") if p]
        obj = self
        for i, part in enumerate(parts):
            if not isinstance(obj, Group):
                raise KeyError(path)
            if obj._links is None:
                obj._load()
            if part not in obj._links:
                raise KeyError(f"{path} (no '{part}' in {obj.name})")
            obj = obj._child(part)
        return obj

    def _child(self, name):
        f = self._f
        addr = self._links[name]
        full = self.name.rstrip("/") + "/" + name
        for mtype, _, _, _ in f._messages(addr):
            if mtype == 0x08:
                return Dataset(f, addr, full)
        return Group(f, addr, full)

    def visit_datasets(self, prefix=""):
        for k in self.keys():
            c = self[k]
            if isinstance(c, Group):
                yield from c.visit_datasets(prefix + k + "/")
            else:
                yield prefix + k, c


class Dataset(_Object):
    def _meta(self):
        f = self._f
        shape = t = layout = None
        filters = []
        for mtype, _, d, size in f._messages(self._addr):
            if mtype == 0x01:
                shape = f._dataspace(d)
            elif mtype == 0x03:
                t = f._datatype(d)
            elif mtype == 0x08:
                layout = d
            elif mtype == 0x0B:
                filters = self._filters(d)
        return shape, t, layout, filters

    @property
    def shape(self):
        return self._meta()[0]

    def _filters(self, d):
        f, b = self._f, self._f.buf
        ver, n = b[d], b[d + 1]
        p = d + (8 if ver == 1 else 2)
        out = []
        for _ in range(n):
            fid = f._u(p, 2)
            p += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = f._u(p, 2)
                p += 2
            p += 2  # flags
            ncd = f._u(p, 2)
            p += 2
            p += ((nlen + 7) & ~7) if ver == 1 else nlen
            cd = [f._u(p + 4 * i, 4) for i in range(ncd)]
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def read(self):
        f, b = self._f, self._f.buf
        shape, t, d, filters = self._meta()
        if shape is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        ver = b[d]
        if ver == 3:
            cls = b[d + 1]
            if cls == 0:
                sz = f._u(d + 2, 2)
                return f._decode(b[d + 4:d + 4 + sz], t, shape)
            if cls == 1:
                addr, sz = f._addr(d + 2), f._len(d + 2 + f.O)
                if addr == UNDEF:
                    return np.zeros(shape, t.dtype) if t.dtype is not None else None
                return f._decode(b[addr:addr + sz], t, shape)
            if cls == 2:
                ndim = b[d + 2]
                btree = f._addr(d + 3)
                cdims = [f._u(d + 3 + f.O + 4 * i, 4) for i in range(ndim)]
                return self._read_chunked(btree, cdims[:-1], shape, t, filters)
            raise NotImplementedError(f"data layout class {cls}")
        if ver in (1, 2):
            ndim, cls = b[d + 1], b[d + 2]
            p = d + 8
            addr = UNDEF
            if cls != 0:
                addr = f._addr(p)
                p += f.O
            dims = [f._u(p + 4 * i, 4) for i in range(ndim)]
            p += 4 * ndim
            if cls == 1:
                return f._decode(b[addr:addr + n * t.size], t, shape)
            if cls == 2:
                return self._read_chunked(addr, dims[:-1] if len(dims) > len(shape) else dims, shape, t, filters)
            sz = f._u(p, 4)
            return f._decode(b[p + 4:p + 4 + sz], t, shape)
        raise NotImplementedError(f"data layout version {ver}")

    def _read_chunked(self, btree, cdims, shape, t, filters):
        if t.dtype is None:
            raise NotImplementedError("chunked variable-length data")
        f = self._f
        out = np.zeros(shape, t.dtype.newbyteorder("<") if t.dtype.byteorder == ">" else t.dtype)
        if btree == UNDEF:
            return out
        ndim = len(shape)
        csize = int(np.prod(cdims)) * t.size
        for offs, fmask, addr, nbytes in self._chunks(btree, ndim):
            raw = f.buf[addr:addr + nbytes]
            for k in range(len(filters) - 1, -1, -1):
                if fmask & (1 << k):
                    continue
                fid, cd = filters[k]
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else t.size
                    a = np.frombuffer(raw, np.uint8)
                    m = len(a) // es
                    raw = a[:m * es].reshape(es, m).T.tobytes() + a[m * es:].tobytes()
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise NotImplementedError(f"HDF5 filter {fid}")
            chunk = np.frombuffer(raw[:csize], t.dtype).reshape(cdims)
            sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
            sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
            out[sl_out] = chunk[sl_in]
        return out

    def _chunks(self, addr, ndim):
        f, b = self._f, self._f.buf
        if b[addr:addr + 4] != b"TREE" or b[addr + 4] != 1:
            raise H5Error("bad chunk B-tree node")
        level, n = b[addr + 5], f._u(addr + 6, 2)
        p = addr + 8 + 2 * f.O
        ksz = 8 + 8 * (ndim + 1)
        for i in range(n):
            k = p + i * (ksz + f.O)
            nbytes, fmask = f._u(k, 4), f._u(k + 4, 4)
            offs = [f._u(k + 8 + 8 * j, 8) for j in range(ndim)]
            child = f._addr(k + ksz)
            if level == 0:
                yield offs, fmask, child, nbytes
            else:
                yield from self._chunks(child, ndim)

    def __array__(self, dtype=None):
        a = self.read()
        return a if dtype is None else a.astype(dtype)


def _to_str(v):
    if isinstance(v, np.ndarray) and v.shape == ():
        v = v[()]
    if isinstance(v, bytes):
        return v.rstrip(b"\0").decode("utf-8", "replace")
    return str(v)


def attr_strings(v):
    """An h5py string-array attribute (``layer_names`` / ``weight_names``) -> list of str."""
    if v is None:
        return []
    a = np.atleast_1d(v)
    return [_to_str(x) for x in a.tolist()]
