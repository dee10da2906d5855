"""Minimal pure-Python reader for the subset of HDF5 that SLEAP's files use.

The reference reads ``best_model.h5`` through Keras/h5py (``sleap/nn/inference.py:3203-3213``,
``sleap/nn/model.py``) and ``.slp`` label files through h5py (``sleap/io/format/hdf5.py:132-263``).
h5py is not available on the deployment image, so this module restates the *file format*
(HDF5 File Format Specification 1.x: superblock v0/v1, v1 object headers, symbol-table groups
with v1 B-trees + local heaps, contiguous / compact / chunked layouts with deflate + shuffle
filters, fixed/float/string/compound/enum/array/vlen datatypes, attributes, global heaps).
Read-only; everything is returned as NumPy arrays / Python objects.

    f = File(path)
    f["model_weights"].attrs["layer_names"]; f["model_weights/conv/conv/kernel:0"][()]
"""
import struct
import zlib
from typing import Dict, List, Optional

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(IOError):
    pass


class _Datatype:
    __slots__ = ("cls", "size", "np_dtype", "vlen_kind", "base", "consumed")


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error("not an HDF5 file")
        ver = buf[8]
        if ver not in (0, 1):
            raise H5Error(f"unsupported superblock version {ver} (only the classic v0/v1 layout h5py writes by default)")
        self.O, self.L = buf[13], buf[14]
        if self.O != 8 or self.L != 8:
            raise H5Error("only 8-byte offsets/lengths are supported")
        p = 24 + (4 if ver == 1 else 0)
        self.base = self.u64(p)
        p += 32                                   # base, free-space, EOF, driver-info addresses
        self.root_header = self.u64(p + 8)        # root symbol-table entry: name offset, object header address

    def u8(self, p):
        return self.b[p]

    def u16(self, p):
        return struct.unpack_from("<H", self.b, p)[0]

    def u32(self, p):
        return struct.unpack_from("<I", self.b, p)[0]

    def u64(self, p):
        return struct.unpack_from("<Q", self.b, p)[0]

    # ---- object headers -----------------------------------------------------------------
    def messages(self, addr):
        """All (type, flags, body_offset, body_size) header messages of the object at `addr` (v1 headers)."""
        b = self.b
        if b[addr:addr + 4] == b"OHDR":
            return self._messages_v2(addr)
        if b[addr] != 1:
            raise H5Error(f"unsupported object header version {b[addr]} at {addr}")
        nmsg = self.u16(addr + 2)
        size = self.u32(addr + 8)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(out) < nmsg:
                t, sz, fl = self.u16(p), self.u16(p + 2), b[p + 4]
                body = p + 8
                if t == 0x10:
                    blocks.append((self.u64(body), self.u64(body + 8)))
                out.append((t, fl, body, sz))
                p = body + sz
        return out

    def _messages_v2(self, addr):
        b = self.b
        flags = b[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        szb = 1 << (flags & 3)
        chunk0 = int.from_bytes(b[p:p + szb], "little")
        p += szb
        blocks = [(p, chunk0)]
        out = []
        track = bool(flags & 0x04)
        while blocks:
            p, n = blocks.pop(0)
            end = p + n
            while p + 4 + (2 if track else 0) <= end:
                t, sz, fl = b[p], self.u16(p + 1), b[p + 3]
                body = p + 4 + (2 if track else 0)
                if t == 0x10:
                    ca, cl = self.u64(body), self.u64(body + 8)
                    blocks.append((ca + 4, cl - 8))       # skip "OCHK", drop checksum
                out.append((t, fl, body, sz))
                p = body + sz
        return out

    # ---- groups -------------------------------------------------------------------------
    def heap_string(self, heap_addr, off):
        if self.b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5Error("bad local heap")
        data = self.u64(heap_addr + 24)
        e = self.b.index(b"\x00", data + off)
        return self.b[data + off:e].decode("utf-8")

    def group_links(self, addr) -> Dict[str, int]:
        links: Dict[str, int] = {}
        for t, fl, p, sz in self.messages(addr):
            if t == 0x11:
                self._walk_group_btree(self.u64(p), self.u64(p + 8), links)
            elif t == 0x06:                      # new-style link message
                ver, lf = self.b[p], self.b[p + 1]
                q = p + 2
                ltype = 0
                if lf & 0x08:
                    ltype = self.b[q]; q += 1
                if lf & 0x04:
                    q += 8
                if lf & 0x10:
                    q += 1
                nb = 1 << (lf & 3)
                ln = int.from_bytes(self.b[q:q + nb], "little"); q += nb
                name = self.b[q:q + ln].decode("utf-8"); q += ln
                if ltype == 0:
                    links[name] = self.u64(q)
        return links

    def _walk_group_btree(self, node, heap, links):
        b = self.b
        if node == UNDEF:
            return
        if b[node:node + 4] == b"SNOD":
            n = self.u16(node + 6)
            p = node + 8
            for _ in range(n):
                links[self.heap_string(heap, self.u64(p))] = self.u64(p + 8)
                p += 40
            return
        if b[node:node + 4] != b"TREE":
            raise H5Error("bad group B-tree node")
        n = self.u16(node + 6)
        p = node + 24 + 8                        # skip key 0
        for _ in range(n):
            self._walk_group_btree(self.u64(p), heap, links)
            p += 16                              # child pointer + next key

    # ---- datatypes ----------------------------------------------------------------------
    def datatype(self, p) -> _Datatype:
        b = self.b
        dt = _Datatype()
        cv = b[p]
        cls, ver = cv & 0x0F, cv >> 4
        bits = b[p + 1] | (b[p + 2] << 8) | (b[p + 3] << 16)
        size = self.u32(p + 4)
        dt.cls, dt.size, dt.vlen_kind, dt.base = cls, size, None, None
        q = p + 8
        if cls == 0:
            dt.np_dtype = np.dtype((">" if bits & 1 else "<") + ("i" if bits & 8 else "u") + str(size))
            q += 4
        elif cls == 1:
            dt.np_dtype = np.dtype((">" if bits & 1 else "<") + "f" + str(size))
            q += 12
        elif cls == 3:
            dt.np_dtype = np.dtype(f"S{size}")
        elif cls == 4:                           # bitfield
            dt.np_dtype = np.dtype(f"<u{size}")
            q += 4
        elif cls == 5:                           # opaque
            tag = 0
            dt.np_dtype = np.dtype(f"V{size}")
            q += (bits & 0xFF + 7) // 8 * 8 if tag else ((bits & 0xFF) + 7) // 8 * 8
        elif cls == 6:
            n = bits & 0xFFFF
            names, fmts, offs = [], [], []
            for _ in range(n):
                e = b.index(b"\x00", q)
                name = b[q:e].decode("utf-8")
                if ver < 3:
                    q += (e - q + 8) // 8 * 8
                else:
                    q = e + 1
                if ver == 1:
                    off = self.u32(q); q += 4 + 1 + 3 + 4 + 4 + 16
                elif ver == 2:
                    off = self.u32(q); q += 4
                else:
                    nb = 1
                    while (1 << (8 * nb)) <= size and nb < 8:
                        nb += 1
                    off = int.from_bytes(b[q:q + nb], "little"); q += nb
                m = self.datatype(q)
                q = m.consumed
                names.append(name); offs.append(off)
                fmts.append(m.np_dtype if m.vlen_kind is None else np.dtype("V16"))
            dt.np_dtype = np.dtype(dict(names=names, formats=fmts, offsets=offs, itemsize=size))
        elif cls == 7:                           # object reference
            dt.np_dtype = np.dtype("<u8")
        elif cls == 8:                           # enum: base type, names, values (h5py bool = enum of int8)
            base = self.datatype(q)
            q = base.consumed
            n = bits & 0xFFFF
            for _ in range(n):
                e = b.index(b"\x00", q)
                q = q + (e - q + 8) // 8 * 8 if ver < 3 else e + 1
            q += n * base.size
            dt.np_dtype = base.np_dtype
        elif cls == 9:
            base = self.datatype(q)
            q = base.consumed
            dt.vlen_kind = "str" if (bits & 0x0F) == 1 else "seq"
            dt.base = base
            dt.np_dtype = np.dtype("O")
        elif cls == 10:
            rank = b[q]
            q += 4 if ver < 3 else 1
            dims = [self.u32(q + 4 * i) for i in range(rank)]
            q += 4 * rank * (2 if ver < 3 else 1)
            base = self.datatype(q)
            q = base.consumed
            dt.np_dtype = np.dtype((base.np_dtype, tuple(dims)))
        else:
            raise H5Error(f"unsupported datatype class {cls}")
        dt.consumed = q
        return dt

    def dataspace(self, p):
        ver, rank, flags = self.b[p], self.b[p + 1], self.b[p + 2]
        if ver == 1:
            q = p + 8
        elif ver == 2:
            if self.b[p + 3] == 2:               # null dataspace
                return None
            q = p + 4
        else:
            raise H5Error(f"unsupported dataspace version {ver}")
        return tuple(self.u64(q + 8 * i) for i in range(rank))

    # ---- vlen / global heap -------------------------------------------------------------
    def global_heap_object(self, coll, index) -> bytes:
        b = self.b
        if b[coll:coll + 4] != b"GCOL":
            raise H5Error("bad global heap collection")
        size = self.u64(coll + 8)
        p, end = coll + 16, coll + size
        while p + 16 <= end:
            idx, n = self.u16(p), self.u64(p + 8)
            if idx == 0:
                break
            if idx == index:
                return b[p + 16:p + 16 + n]
            p += 16 + (n + 7) // 8 * 8
        raise H5Error(f"global heap object {index} not found")

    def decode(self, raw: bytes, dt: _Datatype, shape):
        n = int(np.prod(shape)) if shape else 1
        if dt.vlen_kind is not None:
            out = np.empty(n, dtype=object)
            for i in range(n):
                ln, coll, idx = struct.unpack_from("<IQI", raw, 16 * i)
                if coll == 0 or ln == 0:
                    out[i] = "" if dt.vlen_kind == "str" else np.zeros(0, dt.base.np_dtype)
                    continue
                data = self.global_heap_object(coll, idx)
                if dt.vlen_kind == "str":
                    out[i] = data[:ln].decode("utf-8")
                else:
                    out[i] = np.frombuffer(data, dt.base.np_dtype, ln).copy()
            return out.reshape(shape) if shape else out[0]
        arr = np.frombuffer(raw, dt.np_dtype, n).copy()
        return arr.reshape(shape) if shape else arr[0]

    # ---- datasets -----------------------------------------------------------------------
    def read_dataset(self, addr):
        dt = shape = layout = None
        filters = []
        for t, fl, p, sz in self.messages(addr):
            if t == 0x01:
                shape = self.dataspace(p)
            elif t == 0x03:
                dt = self.datatype(p)
            elif t == 0x08:
                layout = p
            elif t == 0x0B:
                filters = self._filters(p)
        if dt is None or layout is None:
            raise H5Error("object is not a dataset")
        if shape is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        b = self.b
        ver = b[layout]
        if ver == 3:
            lc = b[layout + 1]
            if lc == 0:
                sz = self.u16(layout + 2)
                raw = b[layout + 4:layout + 4 + sz]
            elif lc == 1:
                a, sz = self.u64(layout + 2), self.u64(layout + 10)
                raw = b"\x00" * (n * dt.size) if a == UNDEF else b[a:a + n * dt.size]
            elif lc == 2:
                nd = b[layout + 2]
                bt = self.u64(layout + 3)
                cdims = [self.u32(layout + 11 + 4 * i) for i in range(nd)]
                raw = self._read_chunked(bt, shape, cdims[:-1], dt.size, filters)
            else:
                raise H5Error(f"unsupported layout class {lc}")
        elif ver in (1, 2):
            nd, lc = b[layout + 1], b[layout + 2]
            q = layout + 8
            a = None
            if lc != 0:
                a = self.u64(q); q += 8
            dims = [self.u32(q + 4 * i) for i in range(nd)]
            q += 4 * nd
            if lc == 0:
                sz = self.u32(q)
                raw = b[q + 4:q + 4 + sz]
            elif lc == 1:
                raw = b"\x00" * (n * dt.size) if a == UNDEF else b[a:a + n * dt.size]
            else:
                raw = self._read_chunked(a, shape, dims[:-1], dt.size, filters)
        else:
            raise H5Error(f"unsupported data layout version {ver}")
        return self.decode(raw, dt, shape)

    def _filters(self, p):
        b = self.b
        ver, nf = b[p], b[p + 1]
        q = p + (8 if ver == 1 else 2)
        out = []
        for _ in range(nf):
            fid = self.u16(q); q += 2
            nl = 0
            if ver == 1 or fid >= 256:
                nl = self.u16(q); q += 2
            q += 2
            ncd = self.u16(q); q += 2
            q += (nl + 7) // 8 * 8 if ver == 1 else nl
            cd = [self.u32(q + 4 * i) for i in range(ncd)]
            q += 4 * ncd
            if ver == 1 and ncd % 2:
                q += 4
            out.append((fid, cd))
        return out

    def _read_chunked(self, btree, shape, cdims, esize, filters):
        out = np.zeros(tuple(shape), dtype=f"V{esize}")
        if btree == UNDEF or out.size == 0:
            return out.tobytes()
        nd = len(shape)
        chunks = []
        self._walk_chunk_btree(btree, nd, chunks)
        for caddr, csize, mask, offs in chunks:
            raw = self.b[caddr:caddr + csize]
            for i, (fid, cd) in reversed(list(enumerate(filters))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else esize
                    raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise H5Error(f"unsupported filter id {fid}")
            c = np.frombuffer(raw, dtype=f"V{esize}", count=int(np.prod(cdims))).reshape(cdims)
            sl_out, sl_in = [], []
            for d in range(nd):
                lo = offs[d]
                hi = min(lo + cdims[d], shape[d])
                sl_out.append(slice(lo, hi)); sl_in.append(slice(0, hi - lo))
            out[tuple(sl_out)] = c[tuple(sl_in)]
        return out.tobytes()

    def _walk_chunk_btree(self, node, nd, chunks):
        b = self.b
        if b[node:node + 4] != b"TREE" or b[node + 4] != 1:
            raise H5Error("bad chunk B-tree node")
        level, n = b[node + 5], self.u16(node + 6)
        ksz = 8 + 8 * (nd + 1)
        p = node + 24
        for _ in range(n):
            csize, mask = self.u32(p), self.u32(p + 4)
            offs = [self.u64(p + 8 + 8 * i) for i in range(nd)]
            child = self.u64(p + ksz)
            if level == 0:
                chunks.append((child, csize, mask, offs))
            else:
                self._walk_chunk_btree(child, nd, chunks)
            p += ksz + 8

    # ---- attributes ---------------------------------------------------------------------
    def attributes(self, addr):
        out = {}
        b = self.b
        for t, fl, p, sz in self.messages(addr):
            if t != 0x0C:
                continue
            ver = b[p]
            nsz, dsz, ssz = self.u16(p + 2), self.u16(p + 4), self.u16(p + 6)
            q = p + 8 + (1 if ver == 3 else 0)
            pad = (lambda x: (x + 7) // 8 * 8) if ver == 1 else (lambda x: x)
            name = b[q:q + nsz].split(b"\x00")[0].decode("utf-8"); q += pad(nsz)
            dt = self.datatype(q); q += pad(dsz)
            shape = self.dataspace(q); q += pad(ssz)
            if shape is None:
                out[name] = None
                continue
            n = int(np.prod(shape)) if shape else 1
            out[name] = self.decode(b[q:q + n * dt.size], dt, shape)
        return out


class _Node:
    def __init__(self, r: _Reader, addr: int, name: str):
        self._r, self._addr, self.name = r, addr, name
        self._attrs = None

    @property
    def attrs(self):
        if self._attrs is None:
            self._attrs = self._r.attributes(self._addr)
        return self._attrs


class Dataset(_Node):
    def __getitem__(self, key):
        v = self._r.read_dataset(self._addr)
        if key == () or key is Ellipsis:
            return v
        return v[key]

    def read(self):
        return self._r.read_dataset(self._addr)


class Group(_Node):
    def __init__(self, r, addr, name):
        super().__init__(r, addr, name)
        self._links = None

    def _l(self):
        if self._links is None:
            self._links = self._r.group_links(self._addr)
        return self._links

    def keys(self) -> List[str]:
        return list(self._l().keys())

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __iter__(self):
        return iter(self._l())

    def __getitem__(self, path: str):
        node = self
        for part in [s for s in path.split("/") if s]:
            if not isinstance(node, Group):
                raise KeyError(path)
            links = node._l()
            if part not in links:
                raise KeyError(path)
            addr = links[part]
            types = {t for t, *_ in node._r.messages(addr)}
            cls = Dataset if (0x08 in types) else Group
            node = cls(node._r, addr, (node.name.rstrip("/") + "/" + part))
        return node

    def visit_datasets(self, prefix=""):
        """Yields (path, Dataset) depth-first in link-name order."""
        for k in sorted(self._l()):
            n = self[k]
            if isinstance(n, Group):
                yield from n.visit_datasets(prefix + k + "/")
            else:
                yield prefix + k, n


class File(Group):
    def __init__(self, path: str):
        with open(path, "rb") as f:
            buf = f.read()
        r = _Reader(buf)
        super().__init__(r, r.root_header + r.base, "/")
        self.filename = path

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _as_str(x):
    if isinstance(x, bytes):
        return x.decode("utf-8")
    if isinstance(x, np.bytes_):
        return bytes(x).decode("utf-8")
    return str(x)


def read_keras_weights(path: str) -> Dict[str, Dict[str, np.ndarray]]:
    """``best_model.h5`` (Keras ``save_model``/``save_weights`` HDF5) -> {layer_name: {weight_short_name: array}}.

    Layout (Keras 2.x ``hdf5_format.save_weights_to_hdf5_group``): group ``model_weights`` (or the
    root for weights-only files) has attr ``layer_names``; each layer group has attr
    ``weight_names`` = paths like ``stack0_enc0_conv0/kernel:0`` relative to the layer group.
    Short names drop the ``:0`` suffix and the scope (``kernel``, ``bias``, ``gamma``, ``beta``,
    ``moving_mean``, ``moving_variance``).
    """
    f = File(path)
    g = f["model_weights"] if "model_weights" in f else f
    out: Dict[str, Dict[str, np.ndarray]] = {}
    for ln in g.attrs.get("layer_names", []):
        ln = _as_str(ln)
        lg = g[ln]
        names = [_as_str(w) for w in np.atleast_1d(lg.attrs.get("weight_names", []))]
        if not names:
            continue
        d = {}
        for wn in names:
            short = wn.split("/")[-1].split(":")[0]
            d[short] = np.asarray(lg[wn].read())
        out[ln] = d
    return out
