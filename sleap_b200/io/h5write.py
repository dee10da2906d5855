"""Minimal HDF5 *writer* for the ``.slp`` container (the counterpart of ``h5lite``'s reader).

Emits the same classic layout h5py/libhdf5 1.10 writes by default and that the reference's
``LabelsV1Adaptor.read`` opens (sleap/io/format/hdf5.py:132-263): superblock v0, version-1 object headers,
symbol-table groups (one v1 B-tree node + one SNOD per group, names in a local heap), contiguous datasets,
fixed-point / IEEE float / fixed-length string / compound (v1) datatypes, version-1 attributes.
Restrictions (enough for labels files): <= 8 links per group (one symbol node, leaf K = 4), 1-D or scalar
dataspaces, no filters, no variable-length types.  The datatype / dataspace encoders are checked byte for
byte against messages found in files written by h5py (tests/test_slp_writer.py).
"""
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


# ---- datatype / dataspace messages -----------------------------------------------------------------
def encode_datatype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.names:                                           # compound, version 1
        out = struct.pack("<BBBBI", 0x16, len(dt.names) & 0xFF, (len(dt.names) >> 8) & 0xFF, 0, dt.itemsize)
        for name in dt.names:
            sub, off = dt.fields[name][0], dt.fields[name][1]
            nb = name.encode("utf-8") + b"\x00"
            out += nb + b"\x00" * (-len(nb) % 8)
            out += struct.pack("<IB3xII4I", off, 0, 0, 0, 0, 0, 0, 0)
            out += encode_datatype(sub)
        return out
    if dt.kind == "b":                                     # numpy bool = HDF5 enum {FALSE = 0, TRUE = 1} over int8 (what h5py writes)
        return (struct.pack("<BBBBI", 0x18, 2, 0, 0, 1) + encode_datatype(np.dtype("i1")) + b"FALSE\x00\x00\x00" + b"TRUE\x00\x00\x00\x00" +
                b"\x00\x01")
    if dt.kind in "iu":
        return struct.pack("<BBBBIHH", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        if dt.itemsize == 8:
            return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x3F, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x1F, 0, 4, 0, 32, 23, 8, 0, 23, 127)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)      # null-padded ASCII
    raise TypeError(f"unsupported dtype {dt}")


def encode_dataspace(shape: Tuple[int, ...], unlimited: bool = False) -> bytes:
    """Version 1; ``unlimited`` adds the maximum-dimension array (all H5S_UNLIMITED), as h5py does for
    ``maxshape=(None,)`` (only meaningful with chunked layouts, kept for the byte-level tests)."""
    body = struct.pack("<BBB5x", 1, len(shape), 1 if unlimited else 0)
    body += b"".join(struct.pack("<Q", int(d)) for d in shape)
    if unlimited:
        body += struct.pack("<Q", UNDEF) * len(shape)
    return body


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages: List[bytes]) -> bytes:
    blob = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(blob)) + blob


def _attribute(name: str, value) -> bytes:
    if isinstance(value, (bytes, str)):
        raw = value.encode("utf-8") if isinstance(value, str) else value
        raw = raw if raw else b"\x00"
        dt, ds, data = struct.pack("<BBBBI", 0x13, 0x01, 0, 0, len(raw)), encode_dataspace(()), raw
    else:
        arr = np.asarray(value)
        dt, ds, data = encode_datatype(arr.dtype), encode_dataspace(arr.shape), arr.tobytes()
    nb = name.encode("utf-8") + b"\x00"
    body = struct.pack("<BBHHH", 1, 0, len(nb), len(dt), len(ds)) + _pad8(nb) + _pad8(dt) + _pad8(ds) + data
    return _message(0x000C, body)


class _Alloc:
    def __init__(self, start: int):
        self.buf = bytearray(start)

    def reserve(self, n: int) -> int:
        addr = len(self.buf)
        self.buf += b"\x00" * (n + (-n % 8))
        return addr

    def put(self, addr: int, data: bytes):
        self.buf[addr:addr + len(data)] = data

    def add(self, data: bytes) -> int:
        addr = self.reserve(len(data))
        self.put(addr, data)
        return addr


class Group:
    def __init__(self):
        self.children: Dict[str, object] = {}
        self.attrs: Dict[str, object] = {}

    def create_group(self, name: str) -> "Group":
        g = Group()
        self.children[name] = g
        return g

    def create_dataset(self, name: str, data: np.ndarray):
        data = np.ascontiguousarray(data)
        if data.ndim > 1:
            raise ValueError("only 1-D and scalar datasets are written")
        self.children[name] = data
        return data


class File(Group):
    """``with File(path) as f: f.create_dataset(...); g = f.create_group(...); g.attrs[...] = ...`` — everything is
    laid out and written on close."""

    def __init__(self, path: str):
        super().__init__()
        self.path = path

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *a):
        if exc_type is None:
            self.close()
        return False

    # -- layout ----------------------------------------------------------------------------------------
    def _write_dataset(self, al: _Alloc, data: np.ndarray) -> int:
        raw = data.tobytes()
        addr = al.add(raw) if len(raw) else UNDEF
        msgs = [_message(0x0001, encode_dataspace(data.shape)),
                _message(0x0003, encode_datatype(data.dtype), flags=1),
                _message(0x0005, struct.pack("<BBBBI", 2, 2, 0, 1, 0), flags=1),       # fill value v2: late alloc, default fill
                _message(0x0008, struct.pack("<BBQQ", 3, 1, addr, len(raw)))]            # layout v3, contiguous
        return al.add(_object_header(msgs))

    def _write_group(self, al: _Alloc, g: Group) -> Tuple[int, int, int]:
        """-> (object header address, B-tree address, local heap address)."""
        if len(g.children) > 2 * LEAF_K:
            raise ValueError(f"at most {2 * LEAF_K} links per group are supported")
        names = sorted(g.children)                          # symbol nodes are ordered by name
        child_addr = {}
        for n in names:
            c = g.children[n]
            child_addr[n] = self._write_group(al, c) if isinstance(c, Group) else (self._write_dataset(al, c), None, None)
        # local heap: offset 0 holds the empty string (the B-tree's first key)
        heap_data = bytearray(8)
        name_off = {}
        for n in names:
            name_off[n] = len(heap_data)
            heap_data += _pad8(n.encode("utf-8") + b"\x00")
        heap_data_addr = al.add(bytes(heap_data))
        # free-list head 1 = H5HL_FREE_NULL: the data segment is exactly full, there is no free block
        heap_addr = al.add(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, heap_data_addr))
        snod_addr = UNDEF
        if names:
            snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(names)))
            for n in names:
                hdr, bt, hp = child_addr[n]
                if bt is None:
                    snod += struct.pack("<QQII16x", name_off[n], hdr, 0, 0)
                else:                                       # groups cache their B-tree / heap addresses in the scratch pad
                    snod += struct.pack("<QQIIQQ", name_off[n], hdr, 1, 0, bt, hp)
            snod += b"\x00" * (8 + 2 * LEAF_K * 40 - len(snod))
            snod_addr = al.add(bytes(snod))
        node = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF))
        if names:
            node += struct.pack("<QQQ", 0, snod_addr, name_off[names[-1]])   # key0 = "", child, key1 = largest name
        node += b"\x00" * (24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8 - len(node))
        bt_addr = al.add(bytes(node))
        msgs = [_message(0x0011, struct.pack("<QQ", bt_addr, heap_addr))]
        msgs += [_attribute(k, v) for k, v in g.attrs.items()]
        return al.add(_object_header(msgs)), bt_addr, heap_addr

    def close(self):
        al = _Alloc(96)                                     # superblock v0 (56 B) + root symbol-table entry (40 B)
        hdr, bt, hp = self._write_group(al, self)
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(al.buf), UNDEF)
        sb += struct.pack("<QQIIQQ", 0, hdr, 1, 0, bt, hp)
        assert len(sb) == 96
        al.put(0, sb)
        with open(self.path, "wb") as f:
            f.write(bytes(al.buf))
