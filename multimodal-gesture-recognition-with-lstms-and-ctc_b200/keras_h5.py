"""Minimal pure-Python HDF5 reader / writer for Keras 2.1.4 weight files (`*_weights_best.h5`).

The reference restores its towers with `model.load_weights('.../sp_ctc_lstm_weights_best.h5')`
(/root/reference/multimodal_fusion/multimodal.py:68-85) and writes them with `model.save_weights(...)`
(/root/reference/audio_network/data_generator.py:277-281).  h5py is not part of this image, so this module reads
the subset of the HDF5 file format that `h5py` (libver 'earliest', its default) produces for
`keras.engine.topology.save_weights_to_hdf5_group`:

  * superblock version 0 / 1, "old style" groups (symbol-table message -> v1 B-tree of SNOD nodes + local heap),
    version-1 object headers with continuation blocks;
  * datasets with contiguous (or compact, or un-filtered chunked) layout, little-endian IEEE float / integer types;
  * attributes (message 0x000C, versions 1-3) holding fixed-length string arrays (`layer_names`, `weight_names`,
    `backend`, `keras_version`) or numeric scalars/arrays.

Layout Keras 2.1.4 writes:  root attrs['layer_names'];  per layer a group with attrs['weight_names'];  per weight a
dataset named like 'bidirectional_1/forward_blstm_1/kernel:0' (the '/' creates nested groups).  A full `model.save()`
file keeps the same tree under the group 'model_weights'.

`write_keras_weights` emits the same structures (it is what `keras_io.save_weights(model, 'x.h5')` uses and what
produces the fixtures of tests/test_keras_h5.py).  *No real Keras file could be read in this sandbox* (no h5py, no
network): the reader follows the published HDF5 File Format Specification 2.0 and is cross-checked against this
writer and against hand-assembled byte sequences of the spec's structures, not against libhdf5 output.  Anything
outside the subset raises `H5FormatError` instead of guessing.
"""
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ reader
class H5Reader:
    def __init__(self, path_or_bytes):
        if isinstance(path_or_bytes, (bytes, bytearray)):
            self.buf = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, "rb") as f:
                self.buf = f.read()
        b = self.buf
        if b[:8] != SIGNATURE:
            raise H5FormatError("not an HDF5 file (signature at offset 0 missing; user blocks are not supported)")
        ver = b[8]
        if ver not in (0, 1):
            raise H5FormatError("superblock version %d: only the classic format (h5py libver='earliest') is supported" % ver)
        self.O, self.L = b[13], b[14]          # size of offsets / lengths
        if self.O != 8 or self.L != 8:
            raise H5FormatError("only 8-byte offsets and lengths are supported")
        pos = 24 if ver == 0 else 28
        self.base = self._u(pos, 8)
        pos += 4 * 8                             # base, free-space info, end of file, driver info
        # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
        self.root = self._u(pos + 8, 8)

    # -- primitives
    def _u(self, pos, n):
        return int.from_bytes(self.buf[pos:pos + n], "little")

    def _messages(self, addr):
        """(type, flags, data bytes) of every message of the version-1 object header at `addr`."""
        b = self.buf
        a = self.base + addr
        if b[a:a + 4] == b"OHDR":
            raise H5FormatError("version-2 object headers (libver='latest') are not supported")
        if b[a] != 1:
            raise H5FormatError("object header version %d at %d" % (b[a], addr))
        nmsg = self._u(a + 2, 2)
        size = self._u(a + 8, 4)
        blocks = [(a + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, remaining = blocks.pop(0)
            end = p + remaining
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self._u(p, 2), self._u(p + 2, 2), b[p + 4]
                data = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:                                   # continuation
                    blocks.append((self.base + int.from_bytes(data[:8], "little"), int.from_bytes(data[8:16], "little")))
                out.append((mtype, flags, data))
        return out

    # -- datatype / dataspace
    @staticmethod
    def _dtype(d):
        cls, ver = d[0] & 0x0F, d[0] >> 4
        bits0 = d[1]
        size = int.from_bytes(d[4:8], "little")
        if ver not in (1, 2, 3):
            raise H5FormatError("datatype message version %d" % ver)
        if cls == 0:      # fixed point
            if bits0 & 1:
                raise H5FormatError("big-endian integers are not supported")
            return np.dtype("<%s%d" % ("i" if bits0 & 0x08 else "u", size)), 8 + 4
        if cls == 1:      # floating point
            if bits0 & 1:
                raise H5FormatError("big-endian floats are not supported")
            if size not in (2, 4, 8):
                raise H5FormatError("float size %d" % size)
            return np.dtype("<f%d" % size), 8 + 12
        if cls == 3:      # fixed-length string (null-terminated / null-padded / space-padded)
            return np.dtype("S%d" % size), 8
        if cls == 9:
            raise H5FormatError("variable-length types are not supported (Keras 2.1.4 writes fixed-length byte strings)")
        raise H5FormatError("datatype class %d is not supported" % cls)

    def _dspace(self, d):
        ver, rank, flags = d[0], d[1], d[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            p = 4
            if d[3] == 2:          # null dataspace
                return None
        else:
            raise H5FormatError("dataspace message version %d" % ver)
        return tuple(int.from_bytes(d[p + 8 * i:p + 8 * i + 8], "little") for i in range(rank))

    # -- attributes
    def attrs(self, addr):
        out = {}
        for mtype, _, d in self._messages(addr):
            if mtype != 0x000C:
                continue
            ver = d[0]
            nsz, tsz, ssz = (int.from_bytes(d[2 + 2 * i:4 + 2 * i], "little") for i in range(3))
            if ver == 1:
                pad = lambda n: (n + 7) & ~7
                p = 8
            elif ver in (2, 3):
                if d[1] & 3:
                    raise H5FormatError("shared attribute datatypes / dataspaces are not supported")
                pad = lambda n: n
                p = 8 if ver == 2 else 9
            else:
                raise H5FormatError("attribute message version %d" % ver)
            name = d[p:p + nsz].split(b"\0", 1)[0].decode("utf8")
            p += pad(nsz)
            dt, _ = self._dtype(d[p:p + tsz])
            p += pad(tsz)
            shape = self._dspace(d[p:p + ssz])
            p += pad(ssz)
            if shape is None:
                out[name] = None
                continue
            n = int(np.prod(shape)) if shape else 1
            out[name] = np.frombuffer(d[p:p + n * dt.itemsize], dtype=dt).reshape(shape).copy()
        return out

    # -- groups
    def _heap_name(self, heap_addr, off):
        a = self.base + heap_addr
        if self.buf[a:a + 4] != b"HEAP":
            raise H5FormatError("local heap signature missing at %d" % heap_addr)
        data = self.base + self._u(a + 24, 8)
        end = self.buf.index(b"\0", data + off)
        return self.buf[data + off:end].decode("utf8")

    def _btree_snods(self, addr):
        a = self.base + addr
        if self.buf[a:a + 4] != b"TREE":
            raise H5FormatError("B-tree signature missing at %d" % addr)
        if self.buf[a + 4] != 0:
            raise H5FormatError("expected a group B-tree node")
        level, n = self.buf[a + 5], self._u(a + 6, 2)
        p = a + 24
        kids = []
        for i in range(n):
            p += 8                                   # key i
            kids.append(self._u(p, 8))
            p += 8
        if level == 0:
            return kids
        out = []
        for k in kids:
            out.extend(self._btree_snods(k))
        return out

    def children(self, addr):
        """name -> object header address of the members of the group at `addr` (in B-tree = name order)."""
        for mtype, _, d in self._messages(addr):
            if mtype == 0x0011:
                btree, heap = int.from_bytes(d[:8], "little"), int.from_bytes(d[8:16], "little")
                out = {}
                for sn in self._btree_snods(btree):
                    a = self.base + sn
                    if self.buf[a:a + 4] != b"SNOD":
                        raise H5FormatError("symbol table node signature missing at %d" % sn)
                    for i in range(self._u(a + 6, 2)):
                        e = a + 8 + 40 * i
                        out[self._heap_name(heap, self._u(e, 8))] = self._u(e + 8, 8)
                return out
            if mtype in (0x0002, 0x0006):
                raise H5FormatError("new-style groups (link messages) are not supported; write the file with libver='earliest'")
        raise H5FormatError("object at %d is not a group" % addr)

    def resolve(self, addr, path):
        for part in [p for p in path.split("/") if p]:
            kids = self.children(addr)
            if part not in kids:
                raise KeyError("%r not found (members: %s)" % (part, sorted(kids)))
            addr = kids[part]
        return addr

    # -- datasets
    def dataset(self, addr):
        dt = shape = layout = None
        for mtype, _, d in self._messages(addr):
            if mtype == 0x0001:
                shape = self._dspace(d)
            elif mtype == 0x0003:
                dt, _ = self._dtype(d)
            elif mtype == 0x0008:
                layout = d
            elif mtype == 0x000B:
                raise H5FormatError("filtered (compressed) datasets are not supported")
        if dt is None or shape is None or layout is None:
            raise H5FormatError("object at %d is not a dataset" % addr)
        n = int(np.prod(shape)) if shape else 1
        nbytes = n * dt.itemsize
        if layout[0] != 3:
            raise H5FormatError("data layout message version %d" % layout[0])
        cls = layout[1]
        if cls == 0:      # compact
            size = int.from_bytes(layout[2:4], "little")
            raw = layout[4:4 + size]
        elif cls == 1:    # contiguous
            a = int.from_bytes(layout[2:10], "little")
            raw = b"\0" * nbytes if a == UNDEF else self.buf[self.base + a:self.base + a + nbytes]
        elif cls == 2:    # chunked, no filters
            rank = layout[2]
            bt = int.from_bytes(layout[3:11], "little")
            cdims = [int.from_bytes(layout[11 + 4 * i:15 + 4 * i], "little") for i in range(rank)]
            out = np.zeros(shape, dtype=dt)
            if bt != UNDEF:
                self._read_chunks(bt, rank, cdims[:-1], out)
            return out
        else:
            raise H5FormatError("data layout class %d" % cls)
        if len(raw) < nbytes:
            raise H5FormatError("dataset storage is truncated")
        return np.frombuffer(raw[:nbytes], dtype=dt).reshape(shape).copy()

    def _read_chunks(self, addr, rank, cdims, out):
        a = self.base + addr
        if self.buf[a:a + 4] != b"TREE" or self.buf[a + 4] != 1:
            raise H5FormatError("chunk B-tree node expected at %d" % addr)
        level, n = self.buf[a + 5], self._u(a + 6, 2)
        p = a + 24
        ksz = 8 + 8 * rank
        for i in range(n):
            csize = self._u(p, 4)
            offs = [self._u(p + 8 + 8 * j, 8) for j in range(rank - 1)]
            child = self._u(p + ksz, 8)
            p += ksz + 8
            if level > 0:
                self._read_chunks(child, rank, cdims, out)
                continue
            chunk = np.frombuffer(self.buf[self.base + child:self.base + child + csize], dtype=out.dtype)
            chunk = chunk[:int(np.prod(cdims))].reshape(cdims)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, out.shape))
            out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]


def _as_names(a):
    return [x.decode("utf8") if isinstance(x, bytes) else str(x) for x in np.asarray(a).reshape(-1).tolist()]


def read_keras_weights(path):
    """[(layer name, [(weight name, array), ...]), ...] in `layer_names` / `weight_names` order -- what
    `keras.engine.topology.load_weights_from_hdf5_group` iterates (layers without weights are dropped there too)."""
    f = H5Reader(path)
    root = f.root
    top = f.children(root)
    if "layer_names" not in f.attrs(root) and "model_weights" in top:      # file written by model.save()
        root = top["model_weights"]
    a = f.attrs(root)
    if "layer_names" not in a:
        raise H5FormatError("no 'layer_names' attribute: not a Keras weight file")
    out = []
    for lname in _as_names(a["layer_names"]):
        g = f.resolve(root, lname)
        wn = f.attrs(g).get("weight_names")
        names = _as_names(wn) if wn is not None else []
        if not names:
            continue
        out.append((lname, [(n, f.dataset(f.resolve(g, n))) for n in names]))
    return out


# ------------------------------------------------------------------------------------------------ writer
class _Writer:
    """Classic-format HDF5 writer: superblock v0, symbol-table groups (one leaf B-tree node + one SNOD per group, names in
    heap order = sorted, as the B-tree requires), v1 object headers, contiguous datasets, v1 attribute messages."""

    def __init__(self):
        self.buf = bytearray(b"\0" * 96)        # superblock: 24 + 32 (addresses) + 40 (root symbol table entry)

    def _align(self):
        while len(self.buf) % 8:
            self.buf.append(0)

    def _alloc(self, data):
        self._align()
        a = len(self.buf)
        self.buf += data
        return a

    @staticmethod
    def _msg(mtype, data):
        data = bytes(data) + b"\0" * (-len(data) % 8)
        return struct.pack("<HHB3x", mtype, len(data), 0) + data

    @staticmethod
    def _dtype_msg(dt):
        dt = np.dtype(dt)
        if dt.kind == "f":
            size = dt.itemsize
            ebits, mbits = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[size]
            bias = (1 << (ebits - 1)) - 1
            # class 1 v1; bit field: little-endian, mantissa normalisation = implied msb (2 << 4), sign position in byte 2
            head = struct.pack("<BBBBI", 0x11, 0x20, size * 8 - 1, 0, size)
            return head + struct.pack("<HHBBBBI", 0, size * 8, mbits, ebits, 0, mbits, bias)
        if dt.kind in "iu":
            head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0, 0, 0, dt.itemsize)
            return head + struct.pack("<HH", 0, dt.itemsize * 8)
        if dt.kind == "S":
            return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)       # null-padded ASCII
        raise H5FormatError("cannot write dtype %r" % dt)

    @staticmethod
    def _dspace_msg(shape):
        return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)

    def _attr_msg(self, name, value):
        value = np.asarray(value)
        if value.dtype.kind == "U":
            value = np.char.encode(value, "utf8")
        nm = name.encode("utf8") + b"\0"
        dt, ds = self._dtype_msg(value.dtype), self._dspace_msg(value.shape)
        pad = lambda x: x + b"\0" * (-len(x) % 8)
        body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + pad(nm) + pad(dt) + pad(ds) + value.tobytes()
        return self._msg(0x000C, body)

    def _header(self, msgs):
        body = b"".join(msgs)
        return self._alloc(struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body)

    def dataset(self, arr, attrs=None):
        arr = np.ascontiguousarray(arr)
        data = self._alloc(arr.tobytes()) if arr.size else UNDEF
        msgs = [self._msg(0x0001, self._dspace_msg(arr.shape)), self._msg(0x0003, self._dtype_msg(arr.dtype)),
                self._msg(0x0008, struct.pack("<BBQQ", 3, 1, data, arr.nbytes))]
        msgs += [self._attr_msg(k, v) for k, v in (attrs or {}).items()]
        return self._header(msgs)

    def group(self, members, attrs=None):
        """members: name -> object header address."""
        names = sorted(members, key=lambda s: s.encode("utf8"))
        heap = bytearray(b"\0" * 8)                 # offset 0 = the empty string (B-tree key 0)
        offs = []
        for n in names:
            offs.append(len(heap))
            heap += n.encode("utf8") + b"\0"
            heap += b"\0" * (-len(heap) % 8)
        heap_data = self._alloc(bytes(heap))
        heap_addr = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), UNDEF, heap_data))
        snod = b"SNOD" + struct.pack("<BxH", 1, len(names))
        for o, n in zip(offs, names):
            snod += struct.pack("<QQII16x", o, members[n], 0, 0)
        snod += b"\0" * (40 * (32 - len(names)))    # 2 * group leaf node K (= 16) entries
        if len(names) > 32:
            raise H5FormatError("more than 32 members per group are not supported by this writer")
        snod_addr = self._alloc(snod)
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, offs[-1] if offs else 0)
        tree += b"\0" * (8 * (2 * 32 + 1) - 24)     # room for 2 * internal node K (= 16) children
        tree_addr = self._alloc(tree)
        msgs = [self._msg(0x0011, struct.pack("<QQ", tree_addr, heap_addr))] + [self._attr_msg(k, v) for k, v in (attrs or {}).items()]
        return self._header(msgs), tree_addr, heap_addr

    def finish(self, root_addr, tree_addr, heap_addr):
        self._align()
        sb = SIGNATURE + struct.pack("<BBBBBBBxHHI", 0, 0, 0, 0, 0, 8, 8, 16, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", tree_addr, heap_addr)
        self.buf[:len(sb)] = sb
        return bytes(self.buf)


def write_keras_weights(path, layers, backend="tensorflow", keras_version="2.1.4"):
    """layers: [(layer name, [(weight name, array), ...]), ...] in Keras order -> the tree `save_weights` writes."""
    w = _Writer()

    def build(tree, attrs=None):        # tree: dict name -> (dict | ndarray)
        members = {}
        for k, v in tree.items():
            members[k] = build(v)[0] if isinstance(v, dict) else w.dataset(v)
        return w.group(members, attrs)
    root_members = {}
    for lname, weights in layers:
        tree = {}
        for wname, arr in weights:
            node = tree
            parts = wname.split("/")
            for part in parts[:-1]:
                node = node.setdefault(part, {})
            node[parts[-1]] = np.asarray(arr)
        members = {}
        for k, v in tree.items():
            members[k] = build(v)[0] if isinstance(v, dict) else w.dataset(v)
        root_members[lname] = w.group(members, {"weight_names": np.array([n.encode("utf8") for n, _ in weights] or [b""])[:len(weights)]})[0]
    root, tree_addr, heap_addr = w.group(root_members, {
        "layer_names": np.array([n.encode("utf8") for n, _ in layers]),
        "backend": np.array(backend.encode("utf8")), "keras_version": np.array(keras_version.encode("utf8"))})
    data = w.finish(root, tree_addr, heap_addr)
    if path is not None:
        with open(path, "wb") as f:
            f.write(data)
    return data
