"""Dependency-free reader (and a matching writer) for the HDF5 subset Keras ``save_weights`` files use.

The reference stores weights with ``self.model.save_weights(weight_path)`` and loads them with
``self.model.load_weights(weight_path)`` (/root/reference/src/chess_zero/agent/model_chess.py:146,164): Keras 2.0.8 on
h5py 2.x, i.e. libhdf5 "earliest" format.  h5py / libhdf5 are not installable here and the trained file is not shipped
(.MISSING_LARGE_BLOBS), so this module reads the on-disk format directly, from the HDF5 File Format Specification v2:

  superblock v0/v1 -> root symbol-table entry -> version-1 object headers (+ continuation blocks)
  groups: symbol-table message -> v1 B-tree ("TREE") -> symbol nodes ("SNOD") -> names in the local heap ("HEAP");
          compact "link message" groups are accepted too
  datasets: dataspace v1/v2, datatype classes fixed-point / IEEE float, data layout v3 contiguous or compact
  attributes v1-v3 holding fixed-length or variable-length (global heap, "GCOL") strings

Layout of a Keras weight file: root attribute ``layer_names``; one group per layer with attribute ``weight_names``;
each weight a dataset at ``<layer>/<weight_name>`` (weight names contain '/', so datasets sit in nested groups).
``model.save`` files keep the same tree under ``/model_weights``.

Chunked / compressed datasets, version-2 object headers ("OHDR") and dense attribute storage are outside the subset
(Keras does not produce them for weights); they raise ``Hdf5Error`` naming what was met.

PARITY NOTE: pinned only by round trip against ``write_keras_weights`` below (tests/test_hdf5_min.py) -- there is no
libhdf5 in this environment to cross-check either direction.
"""
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(ValueError):
    pass


def _pad8(n):
    return (n + 7) & ~7


# =====================================================================================================
# reader
# =====================================================================================================
class H5File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.buf = f.read()
        b = self.buf
        if b[:8] != SIGNATURE:
            raise Hdf5Error("not an HDF5 file")
        ver = b[8]
        if ver not in (0, 1):
            raise Hdf5Error(f"superblock version {ver} not supported (only the libhdf5 'earliest' format, v0/v1)")
        if b[13] != 8 or b[14] != 8:
            raise Hdf5Error("only 8-byte offsets and lengths are supported")
        pos = 24 if ver == 0 else 28
        self.base = struct.unpack_from("<Q", b, pos)[0]
        root_entry = pos + 32
        self.root = struct.unpack_from("<Q", b, root_entry + 8)[0]

    # ---- object headers -----------------------------------------------------------------------------
    def messages(self, addr):
        """[(type, flags, data bytes)] of the version-1 object header at `addr`, continuation blocks included."""
        b = self.buf
        addr += self.base
        if b[addr:addr + 4] == b"OHDR":
            raise Hdf5Error("version-2 object header (libver='latest' file) not supported")
        if b[addr] != 1:
            raise Hdf5Error(f"object header version {b[addr]} not supported")
        n_msgs, = struct.unpack_from("<H", b, addr + 2)
        size, = struct.unpack_from("<I", b, addr + 8)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < n_msgs:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < n_msgs:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, pos)
                data = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:                       # continuation
                    off, length = struct.unpack_from("<QQ", data, 0)
                    blocks.append((off + self.base, length))
                out.append((mtype, mflags, data))
        return out

    # ---- groups -------------------------------------------------------------------------------------
    def children(self, addr):
        """{name: object header address} of the group at `addr`."""
        out = {}
        for mtype, _, data in self.messages(addr):
            if mtype == 0x0011:                           # symbol table: B-tree + local heap
                btree, heap = struct.unpack_from("<QQ", data, 0)
                heap_data = self._local_heap(heap)
                self._walk_btree(btree, heap_data, out)
            elif mtype == 0x0006:                         # link message (compact new-style group)
                name, target = self._link(data)
                if target is not None:
                    out[name] = target
            elif mtype == 0x0002 and len(data) >= 18:
                fractal = struct.unpack_from("<Q", data, len(data) - 16)[0]
                if fractal != UNDEF:
                    raise Hdf5Error("dense link storage (fractal heap) not supported")
        return out

    def _local_heap(self, addr):
        b = self.buf
        addr += self.base
        if b[addr:addr + 4] != b"HEAP":
            raise Hdf5Error("local heap signature missing")
        size, _, data_addr = struct.unpack_from("<QQQ", b, addr + 8)
        return b[data_addr + self.base:data_addr + self.base + size]

    @staticmethod
    def _cstr(heap, off):
        end = heap.index(b"\0", off)
        return heap[off:end].decode("utf-8")

    def _walk_btree(self, addr, heap, out):
        b = self.buf
        addr += self.base
        if b[addr:addr + 4] != b"TREE":
            raise Hdf5Error("B-tree signature missing")
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        if ntype != 0:
            raise Hdf5Error("expected a group B-tree")
        pos = addr + 24
        for i in range(used):
            child, = struct.unpack_from("<Q", b, pos + 8)          # key_i (8), child_i (8)
            pos += 16
            if level > 0:
                self._walk_btree(child, heap, out)
            else:
                self._symbol_node(child, heap, out)

    def _symbol_node(self, addr, heap, out):
        b = self.buf
        addr += self.base
        if b[addr:addr + 4] != b"SNOD":
            raise Hdf5Error("symbol node signature missing")
        n, = struct.unpack_from("<H", b, addr + 6)
        for i in range(n):
            name_off, obj = struct.unpack_from("<QQ", b, addr + 8 + 40 * i)
            out[self._cstr(heap, name_off)] = obj

    @staticmethod
    def _link(data):
        flags = data[1]
        pos = 2
        ltype = 0
        if flags & 0x08:
            ltype = data[pos]
            pos += 1
        if flags & 0x04:
            pos += 8
        if flags & 0x10:
            pos += 1
        lsz = 1 << (flags & 3)
        nlen = int.from_bytes(data[pos:pos + lsz], "little")
        pos += lsz
        name = data[pos:pos + nlen].decode("utf-8")
        pos += nlen
        return name, (struct.unpack_from("<Q", data, pos)[0] if ltype == 0 else None)

    def resolve(self, path, start=None):
        addr = self.root if start is None else start
        for part in [p for p in path.split("/") if p]:
            kids = self.children(addr)
            if part not in kids:
                raise KeyError(f"'{part}' not found (have {sorted(kids)[:8]}...)")
            addr = kids[part]
        return addr

    # ---- datatypes / dataspaces -------------------------------------------------------------------
    @staticmethod
    def _dataspace(data):
        ver, rank, flags = data[0], data[1], data[2]
        pos = 8 if ver == 1 else 4
        return tuple(struct.unpack_from("<Q", data, pos + 8 * i)[0] for i in range(rank))

    def _datatype(self, data):
        """-> ('num', numpy dtype) | ('str', length) | ('vstr',)"""
        cls, bits0 = data[0] & 0x0F, data[1]
        size, = struct.unpack_from("<I", data, 4)
        order = ">" if (bits0 & 1) else "<"
        if cls == 0:
            signed = bool(bits0 & 0x08)
            return ("num", np.dtype(f"{order}{'i' if signed else 'u'}{size}"))
        if cls == 1:
            return ("num", np.dtype(f"{order}f{size}"))
        if cls == 3:
            return ("str", size)
        if cls == 9:
            if (bits0 & 0x0F) != 1:
                raise Hdf5Error("variable-length sequences not supported")
            return ("vstr",)
        raise Hdf5Error(f"datatype class {cls} not supported")

    def _global_heap_object(self, coll_addr, index):
        b = self.buf
        pos = coll_addr + self.base
        if b[pos:pos + 4] != b"GCOL":
            raise Hdf5Error("global heap signature missing")
        size, = struct.unpack_from("<Q", b, pos + 8)
        end = pos + size
        pos += 16
        while pos + 16 <= end:
            idx, _, _, osz = struct.unpack_from("<HHIQ", b, pos)
            if idx == 0:
                break
            if idx == index:
                return b[pos + 16:pos + 16 + osz]
            pos += 16 + _pad8(osz)
        raise Hdf5Error("global heap object not found")

    def _decode(self, kind, shape, raw):
        count = int(np.prod(shape)) if shape else 1
        if kind[0] == "num":
            return np.frombuffer(raw, dtype=kind[1], count=count).reshape(shape)
        if kind[0] == "str":
            n = kind[1]
            vals = [raw[i * n:(i + 1) * n].split(b"\0")[0].rstrip(b" ").decode("utf-8") for i in range(count)]
        else:
            vals = []
            for i in range(count):
                length, coll, idx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append(self._global_heap_object(coll, idx)[:length].decode("utf-8"))
        return vals if shape else vals[0]

    # ---- attributes -------------------------------------------------------------------------------
    def attrs(self, addr):
        out = {}
        for mtype, _, data in self.messages(addr):
            if mtype == 0x0015:                           # attribute info: dense storage lives in a fractal heap
                off = 2 + (2 if data[1] & 1 else 0)
                if len(data) >= off + 8 and struct.unpack_from("<Q", data, off)[0] != UNDEF:
                    raise Hdf5Error("dense attribute storage not supported")
            if mtype != 0x000C:
                continue
            ver = data[0]
            nsz, tsz, ssz = struct.unpack_from("<HHH", data, 2)
            pos = 8 if ver < 3 else 9
            pad = _pad8 if ver == 1 else (lambda n: n)
            name = data[pos:pos + nsz].split(b"\0")[0].decode("utf-8")
            pos += pad(nsz)
            kind = self._datatype(data[pos:pos + tsz])
            pos += pad(tsz)
            shape = self._dataspace(data[pos:pos + ssz])
            pos += pad(ssz)
            out[name] = self._decode(kind, shape, data[pos:])
        return out

    # ---- datasets ---------------------------------------------------------------------------------
    def read_dataset(self, addr):
        kind = shape = None
        raw = None
        for mtype, _, data in self.messages(addr):
            if mtype == 0x0001:
                shape = self._dataspace(data)
            elif mtype == 0x0003:
                kind = self._datatype(data)
            elif mtype == 0x000B:
                raise Hdf5Error("filtered (compressed) datasets not supported")
            elif mtype == 0x0008:
                ver, cls = data[0], data[1]
                if ver != 3:
                    raise Hdf5Error(f"data layout version {ver} not supported")
                if cls == 1:
                    daddr, dsize = struct.unpack_from("<QQ", data, 2)
                    raw = b"" if daddr == UNDEF else self.buf[daddr + self.base:daddr + self.base + dsize]
                elif cls == 0:
                    dsize, = struct.unpack_from("<H", data, 2)
                    raw = data[4:4 + dsize]
                else:
                    raise Hdf5Error("chunked datasets not supported")
        if kind is None or shape is None or raw is None:
            raise Hdf5Error("object is not a readable dataset")
        if kind[0] != "num":
            raise Hdf5Error("only numeric datasets are supported")
        need = int(np.prod(shape)) * kind[1].itemsize if shape else kind[1].itemsize
        if len(raw) < need:
            raise Hdf5Error("dataset storage shorter than its extent")
        return self._decode(kind, shape, raw[:need])


def read_keras_weights(path, names_in_layer_order=None):
    """{keras weight name: ndarray} for a Keras ``save_weights`` (or ``model.save``) HDF5 file."""
    f = H5File(path)
    root = f.root
    if "layer_names" not in f.attrs(root):
        kids = f.children(root)
        if "model_weights" not in kids:
            raise Hdf5Error("no 'layer_names' attribute: not a Keras weight file")
        root = kids["model_weights"]
    out = {}
    for lname in f.attrs(root)["layer_names"]:
        grp = f.resolve(lname, root)
        for wname in f.attrs(grp).get("weight_names", []):
            out[wname] = np.array(f.read_dataset(f.resolve(wname, grp)))
    return out


# =====================================================================================================
# writer (the same subset; used by ChessModel.save for '.h5' paths and by the round-trip tests)
# =====================================================================================================
class _Writer:
    LEAF_K, INTERNAL_K = 4, 16

    def __init__(self):
        self.buf = bytearray(96)          # superblock placeholder

    def _alloc(self, data):
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    @staticmethod
    def _msg(mtype, data, flags=0):
        data = bytes(data) + b"\0" * (_pad8(len(data)) - len(data))
        return struct.pack("<HHB3x", mtype, len(data), flags) + data

    def _object_header(self, msgs):
        body = b"".join(msgs)
        return self._alloc(struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body)

    @staticmethod
    def _dataspace(shape):
        return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)

    @staticmethod
    def _dtype_f32():
        return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)

    @staticmethod
    def _dtype_str(n):
        return struct.pack("<BBBBI", 0x13, 0x00, 0x00, 0x00, n)        # null-terminated ASCII, fixed length

    def _attr_strings(self, name, values):
        enc = [v.encode("utf-8") for v in values]
        n = max([len(e) for e in enc] + [1])
        dt, ds = self._dtype_str(n), self._dataspace((len(enc),))
        nm = name.encode() + b"\0"
        data = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds))
        for part in (nm, dt, ds):
            data += part + b"\0" * (_pad8(len(part)) - len(part))
        data += b"".join(e.ljust(n, b"\0") for e in enc)
        return self._msg(0x000C, data)

    def dataset(self, arr):
        arr = np.ascontiguousarray(arr, dtype="<f4")
        daddr = self._alloc(arr.tobytes()) if arr.size else UNDEF
        layout = struct.pack("<BBQQ", 3, 1, daddr, arr.nbytes)
        fill = struct.pack("<BBBB", 2, 2, 2, 0)                       # version 2, late allocation, never written
        return self._object_header([self._msg(0x0001, self._dataspace(arr.shape)), self._msg(0x0003, self._dtype_f32()),
                                    self._msg(0x0005, fill), self._msg(0x0008, layout)])

    def group(self, entries, attrs=None):
        """entries: {name: object header address}; attrs: {name: [str]}.  Old-style group (symbol table)."""
        names = sorted(entries, key=lambda s: s.encode("utf-8"))
        heap = bytearray(b"\0" * 8)                                     # offset 0: the empty string
        offs = {}
        for nm in names:
            offs[nm] = len(heap)
            e = nm.encode("utf-8") + b"\0"
            heap += e + b"\0" * (_pad8(len(e)) - len(e))
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 16)                               # one free block: next = 1 (none), size 16
        heap_data = self._alloc(heap)
        heap_addr = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data))
        per = 2 * self.LEAF_K
        if len(names) > per * 2 * self.INTERNAL_K:
            raise Hdf5Error("too many entries for a one-level group B-tree")
        chunks = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        keys, kids = [0], []
        for ch in chunks:
            node = b"SNOD" + struct.pack("<BxH", 1, len(ch))
            for nm in ch:
                node += struct.pack("<QQII16x", offs[nm], entries[nm], 0, 0)
            node += b"\0" * (8 + 40 * per - len(node))
            kids.append(self._alloc(node))
            keys.append(offs[ch[-1]] if ch else 0)
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(kids), UNDEF, UNDEF)
        for i, k in enumerate(kids):
            tree += struct.pack("<QQ", keys[i], k)
        tree += struct.pack("<Q", keys[-1])
        tree += b"\0" * (24 + 16 * 2 * self.INTERNAL_K + 8 - len(tree))
        tree_addr = self._alloc(tree)
        msgs = [self._msg(0x0011, struct.pack("<QQ", tree_addr, heap_addr))]
        for k, v in (attrs or {}).items():
            msgs.append(self._attr_strings(k, v))
        return self._object_header(msgs), tree_addr, heap_addr

    def finish(self, root, tree, heap):
        sb = SIGNATURE + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_keras_weights(path, layers):
    """layers: ordered [(layer_name, [(weight_name, ndarray), ...])] -- the tree ``save_weights`` makes:
    root attr layer_names, one group per layer with attr weight_names, weights at <layer>/<weight_name>."""
    w = _Writer()
    root_entries = {}
    for lname, weights in layers:
        tree = {}                                                       # nested dict of the '/'-separated weight paths
        for wname, arr in weights:
            node = tree
            parts = wname.split("/")
            for part in parts[:-1]:
                node = node.setdefault(part, {})
            node[parts[-1]] = w.dataset(arr)

        def emit(node):
            return w.group({k: (emit(v) if isinstance(v, dict) else v) for k, v in node.items()})[0]

        entries = {k: (emit(v) if isinstance(v, dict) else v) for k, v in tree.items()}
        root_entries[lname] = w.group(entries, {"weight_names": [n for n, _ in weights]})[0]
    root, tree_addr, heap_addr = w.group(root_entries, {"layer_names": [n for n, _ in layers],
                                                          "backend": ["tensorflow"], "keras_version": ["2.0.8"]})
    with open(path, "wb") as f:
        f.write(w.finish(root, tree_addr, heap_addr))
