"""Test-side minimal HDF5 WRITER (superblock v0, symbol-table groups, contiguous datasets, v1
attributes) producing the structures h5py 2.x / Keras 2.3 write, so that the product's h5lite reader
and keras_h5 importer can be exercised without h5py.  The reader is additionally checked against a
real HDF5-library-written file (scipy's testhdf5_7.4_GLNX86.mat) in tests/test_keras_h5.py."""
import struct

import numpy as np

O = L = 8
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K = 4  # SNOD holds up to 2*LEAF_K entries


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        if dt.itemsize == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits = (0x20, 31, 0)
        else:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits = (0x20, 63, 0)
        return bytes([0x11, bits[0], bits[1], bits[2]]) + struct.pack("<I", dt.itemsize) + props
    if dt.kind in "iu":
        return bytes([0x10, 0x08 if dt.kind == "i" else 0, 0, 0]) + struct.pack("<I", dt.itemsize) + \
            struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        return bytes([0x13, 0x00, 0, 0]) + struct.pack("<I", dt.itemsize)
    raise TypeError(dt)


def _space_msg(shape):
    return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _attr_msg(name, value):
    a = np.asarray(value)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf-8")
    nm = name.encode() + b"\0"
    dt, sp = _dtype_msg(a.dtype), _space_msg(a.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + a.tobytes()
    return 0x0C, body


class Writer:
    def __init__(self, user_block=0):
        self.user_block = user_block
        self.buf = bytearray(b"\0" * (user_block + 96))  # superblock v0 with 8-byte offsets is 96 bytes

    def _alloc(self, data):
        while len(self.buf) % 8:
            self.buf.append(0)
        pos = len(self.buf) - self.user_block  # addresses are relative to the base address
        self.buf += data
        return pos

    def _header(self, msgs):
        body = b""
        for mtype, data in msgs:
            data = _pad8(data)
            if len(data) >= 65536:
                raise ValueError("header message too large")
            body += struct.pack("<HHB3x", mtype, len(data), 0) + data
        hdr = struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body))
        return self._alloc(hdr + body)

    def dataset(self, arr, attrs=None):
        arr = np.ascontiguousarray(arr)
        addr = self._alloc(arr.tobytes()) if arr.size else UNDEF
        layout = struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)
        msgs = [(0x01, _space_msg(arr.shape)), (0x03, _dtype_msg(arr.dtype)), (0x08, layout)]
        msgs += [_attr_msg(k, v) for k, v in (attrs or {}).items()]
        return self._header(msgs)

    def chunked_dataset(self, arr, chunks, gzip=True, shuffle=True):
        """Chunked layout (B-tree v1, node type 1) with the shuffle + deflate pipeline h5py's
        ``compression='gzip', shuffle=True`` produces."""
        import itertools
        import zlib
        arr = np.ascontiguousarray(arr)
        nd, es = arr.ndim, arr.dtype.itemsize
        grid = [range(0, arr.shape[d], chunks[d]) for d in range(nd)]
        entries = []
        for offs in itertools.product(*grid):
            blk = np.zeros(chunks, arr.dtype)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunks, arr.shape))
            blk[tuple(slice(0, x.stop - x.start) for x in sl)] = arr[sl]
            raw = blk.tobytes()
            if shuffle:
                raw = np.frombuffer(raw, np.uint8).reshape(-1, es).T.tobytes()
            if gzip:
                raw = zlib.compress(raw, 4)
            entries.append((offs, len(raw), self._alloc(raw)))
        if len(entries) > 60:
            raise ValueError("too many chunks for a single B-tree node in this test writer")
        node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), UNDEF, UNDEF)
        for offs, nbytes, addr in entries:
            node += struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)
            node += struct.pack("<Q", addr)
        node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape) + struct.pack("<Q", 0)
        btree = self._alloc(node)
        layout = struct.pack("<BBB", 3, 2, nd + 1) + struct.pack("<Q", btree) + \
            b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", es)
        filt = b""
        nf = 0
        if shuffle:
            filt += struct.pack("<HHHH", 2, 0, 1, 1) + struct.pack("<I", es) + struct.pack("<I", 0)
            nf += 1
        if gzip:
            filt += struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<I", 4) + struct.pack("<I", 0)
            nf += 1
        msgs = [(0x01, _space_msg(arr.shape)), (0x03, _dtype_msg(arr.dtype)), (0x08, layout)]
        if nf:
            msgs.append((0x0B, struct.pack("<BB6x", 1, nf) + filt))
        return self._header(msgs)

    def group(self, children: dict, attrs=None):
        """children: name -> object header address (already written)."""
        names = sorted(children)
        heap_data = bytearray(b"\0" * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += _pad8(n.encode() + b"\0")
        data_addr = self._alloc(bytes(heap_data))
        heap = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, data_addr))
        snods = []
        for i in range(0, max(len(names), 1), 2 * LEAF_K):
            part = names[i:i + 2 * LEAF_K]
            ent = b"".join(struct.pack("<QQII16x", offs[n], children[n], 0, 0) for n in part)
            ent += b"\0" * (40 * (2 * LEAF_K - len(part)))
            snods.append((self._alloc(b"SNOD" + struct.pack("<BBH", 1, 0, len(part)) + ent), part))
        if len(snods) > 64:
            raise ValueError("too many entries for a single B-tree node in this test writer")
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", 0)
        for addr, part in snods:
            node += struct.pack("<QQ", addr, offs[part[-1]] if part else 0)
        btree = self._alloc(node)
        msgs = [(0x11, struct.pack("<QQ", btree, heap))] + [_attr_msg(k, v) for k, v in (attrs or {}).items()]
        return self._header(msgs), btree, heap

    def finish(self, root):
        root_addr, btree, heap = root
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, O, L, 0, LEAF_K, 32, 0)
        sb += struct.pack("<QQQQ", self.user_block, UNDEF, len(self.buf) - self.user_block, UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[self.user_block:self.user_block + 96] = sb
        return bytes(self.buf)


def write_keras_model(weights_by_layer: dict, layer_order, model_config: str | None = None, wrap=True,
                      user_block=0) -> bytes:
    """weights_by_layer: layer -> list of (weight name like 'kernel:0', array).  Layers without weights
    get an empty group, like Keras does for activations / pads."""
    w = Writer(user_block)
    layer_groups = {}
    for ln in layer_order:
        ws = weights_by_layer.get(ln, [])
        inner = {}
        for wn, arr in ws:
            inner[wn] = w.dataset(np.asarray(arr, np.float32))
        attrs = {"weight_names": np.array([f"{ln}/{wn}".encode() for wn, _ in ws] or [], dtype="S64")}
        if ws:
            sub, _, _ = w.group(inner)
            g, _, _ = w.group({ln: sub}, attrs)
        else:
            g, _, _ = w.group({}, {})
        layer_groups[ln] = g
    top_attrs = {"layer_names": np.array([n.encode() for n in layer_order], dtype="S64"),
                 "backend": np.bytes_(b"tensorflow"), "keras_version": np.bytes_(b"2.3.1")}
    mw = w.group(layer_groups, top_attrs)
    if not wrap:
        return w.finish(mw)
    root_attrs = {"keras_version": np.bytes_(b"2.3.1"), "backend": np.bytes_(b"tensorflow")}
    if model_config is not None:
        root_attrs["model_config"] = np.bytes_(model_config.encode())
    root = w.group({"model_weights": mw[0]}, root_attrs)
    return w.finish(root)
