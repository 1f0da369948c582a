"""Minimal read-only HDF5 reader for Keras checkpoints (``models_tracking/BaseTracker.py:74-80`` writes
``<saved_model_dir>/<name>-CHKPNT-<epoch>-<val_loss>.hdf5`` with ``ModelCheckpoint``; ``MultiObjDetTracker.py:291-293``
loads one with ``model.load_weights``).  h5py is not installable here, so this module parses the subset of the HDF5
file format that h5py's default settings (``libver='earliest'``) produce:

* superblock version 0 / 1, 8-byte or 4-byte offsets;
* old-style groups: symbol-table message -> v1 B-tree (``TREE``) -> symbol nodes (``SNOD``) -> names in a local heap;
* version-1 object headers with continuation blocks;
* datasets of fixed-point and IEEE floating-point types, little or big endian; contiguous, compact and chunked
  layouts (v1 chunk B-tree), deflate + shuffle filters.

Attributes, links, new-style (fractal-heap) groups, variable-length types and external storage are not read -- the
weights are found by walking the group tree and matching dataset names, which is all ``load_weights`` needs.

``read_hdf5(path)`` -> ``{"/group/sub/dataset": ndarray}``.  Pinned against h5py-written files that ship in the reference
tree (``py-faster-rcnn/caffe-fast-rcnn/src/caffe/test/test_data/*.h5``: contiguous float32 and gzip-chunked uint8 /
float32 datasets with contents known from ``generate_sample_data.py``) by ``tests/test_formats_cpu.py``.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, Optional

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = {4: 0xFFFFFFFF, 8: 0xFFFFFFFFFFFFFFFF}


class HDF5Error(ValueError):
    pass


class _File:
    def __init__(self, buf: bytes):
        self.b = buf
        base = buf.find(_SIG)
        if base != 0:
            raise HDF5Error("not an HDF5 file (or a user block precedes the superblock)")
        ver = buf[8]
        if ver not in (0, 1):
            raise HDF5Error(f"superblock version {ver} is not supported (file written with libver='latest'?)")
        self.O, self.L = buf[13], buf[14]                         # size of offsets / lengths
        if self.O not in (4, 8) or self.L not in (4, 8):
            raise HDF5Error("bad offset / length size")
        p = 24 + (4 if ver == 1 else 0)
        self.base = self.off(p)
        p += 4 * self.O                                           # base, free-space, end-of-file, driver-info addresses
        self.root = self.symbol_entry(p)

    # ---- primitives
    def off(self, p: int) -> int:
        return int.from_bytes(self.b[p:p + self.O], "little")

    def length(self, p: int) -> int:
        return int.from_bytes(self.b[p:p + self.L], "little")

    def u16(self, p: int) -> int:
        return struct.unpack_from("<H", self.b, p)[0]

    def u32(self, p: int) -> int:
        return struct.unpack_from("<I", self.b, p)[0]

    def symbol_entry(self, p: int) -> dict:
        """link-name offset, object-header address, cache type, scratch pad (B-tree + heap address when cached)"""
        e = {"name_off": self.off(p), "header": self.off(p + self.O), "cache": self.u32(p + 2 * self.O)}
        s = p + 2 * self.O + 8
        if e["cache"] == 1:
            e["btree"], e["heap"] = self.off(s), self.off(s + self.O)
        return e

    # ---- object headers
    def messages(self, addr: int):
        """(type, payload offset, payload size) of every message of a version-1 object header, continuations followed."""
        b = self.b
        if b[addr] != 1:
            raise HDF5Error(f"object header version {b[addr]} at {addr} is not supported (new-style file)")
        n_msg = self.u16(addr + 2)
        size = self.u32(addr + 8)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < n_msg:
            p, remaining = blocks.pop(0)
            end = p + remaining
            while p + 8 <= end and len(out) < n_msg:
                mtype, msize = self.u16(p), self.u16(p + 2)
                body = p + 8
                out.append((mtype, body, msize))
                if mtype == 0x0010:                               # continuation: offset, length
                    blocks.append((self.off(body) + self.base, self.length(body + self.O)))
                p = body + msize
        return out

    # ---- groups
    def heap_name(self, heap_addr: int, off: int) -> str:
        if self.b[heap_addr:heap_addr + 4] != b"HEAP":
            raise HDF5Error("bad local heap")
        data = self.off(heap_addr + 8 + 2 * self.L) + self.base
        p = data + off
        q = self.b.index(b"\0", p)
        return self.b[p:q].decode("utf-8", "replace")

    def group_entries(self, btree: int, heap: int):
        b = self.b
        if b[btree:btree + 4] != b"TREE" or b[btree + 4] != 0:
            raise HDF5Error("bad group B-tree node")
        level, used = b[btree + 5], self.u16(btree + 6)
        p = btree + 8 + 2 * self.O
        for i in range(used):
            child = self.off(p + self.L + i * (self.L + self.O)) + self.base      # key_i, child_i, key_i+1, ...
            if level > 0:
                yield from self.group_entries(child, heap)
                continue
            if b[child:child + 4] != b"SNOD":
                raise HDF5Error("bad symbol node")
            n = self.u16(child + 6)
            q = child + 8
            for _ in range(n):
                e = self.symbol_entry(q)
                yield self.heap_name(heap, e["name_off"]), e
                q += 2 * self.O + 24

    # ---- datasets
    def dataset(self, msgs) -> Optional[np.ndarray]:
        shape = dtype = layout = None
        filters = []
        for mtype, p, size in msgs:
            b = self.b
            if mtype == 0x0001:                                   # dataspace
                ver, rank = b[p], b[p + 1]
                q = p + (8 if ver == 1 else 4)
                shape = tuple(self.length(q + i * self.L) for i in range(rank))
            elif mtype == 0x0003:                                 # datatype
                cls, bits0, nbytes = b[p] & 15, b[p + 1], self.u32(p + 4)
                order = ">" if bits0 & 1 else "<"
                if cls == 0:
                    dtype = np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{nbytes}")
                elif cls == 1:
                    dtype = np.dtype(f"{order}f{nbytes}")
                else:
                    return None                                   # strings, compounds ...: not a weight tensor
            elif mtype == 0x0008:                                 # data layout
                ver = b[p]
                if ver == 3:
                    cls = b[p + 1]
                    if cls == 0:
                        n = self.u16(p + 2)
                        layout = ("compact", p + 4, n)
                    elif cls == 1:
                        layout = ("contiguous", self.off(p + 2), self.length(p + 2 + self.O))
                    elif cls == 2:
                        rank = b[p + 2]
                        addr = self.off(p + 3)
                        dims = [self.u32(p + 3 + self.O + 4 * i) for i in range(rank)]
                        layout = ("chunked", addr, dims)
                elif ver in (1, 2):
                    rank, cls = b[p + 1], b[p + 2]
                    q = p + 8
                    addr = None
                    if cls != 0:
                        addr = self.off(q)
                        q += self.O
                    dims = [self.u32(q + 4 * i) for i in range(rank)]
                    q += 4 * rank
                    if cls == 1:
                        layout = ("contiguous", addr, None)
                    elif cls == 2:
                        layout = ("chunked", addr, dims + [self.u32(q)])
                    else:
                        layout = ("compact", q + 4, self.u32(q))
                else:
                    raise HDF5Error(f"data layout version {ver} is not supported")
            elif mtype == 0x000B:                                 # filter pipeline
                ver, nf = b[p], b[p + 1]
                q = p + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid = self.u16(q)
                    if ver == 1 or fid >= 256:
                        name_len = self.u16(q + 2); q += 4
                    else:
                        name_len = 0; q += 2
                    ncd = self.u16(q + 2)
                    q += 4
                    q += (name_len + 7) // 8 * 8 if ver == 1 else name_len
                    q += 4 * ncd
                    if ver == 1 and ncd % 2:
                        q += 4
                    filters.append(fid)
        if shape is None or dtype is None or layout is None:
            return None
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if layout[0] == "compact":
            raw = self.b[layout[1]:layout[1] + layout[2]]
            return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).copy()
        if layout[0] == "contiguous":
            addr = layout[1]
            if addr == _UNDEF[self.O]:
                return np.zeros(shape, dtype=dtype)               # never written: fill value
            a = addr + self.base
            return np.frombuffer(self.b, dtype=dtype, count=n, offset=a).reshape(shape).copy()
        # chunked
        addr, dims = layout[1], layout[2]
        chunk_shape = tuple(dims[:-1])
        out = np.zeros(shape, dtype=dtype)
        if addr == _UNDEF[self.O]:
            return out
        for offsets, caddr, csize, mask in self.chunks(addr + self.base, len(shape)):
            raw = self.b[caddr:caddr + csize]
            for i, fid in reversed(list(enumerate(filters))):     # filters are undone in reverse order
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:                                    # shuffle: bytes of all elements grouped by significance
                    es = dtype.itemsize
                    raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                elif fid == 3:                                    # fletcher32 checksum trails the data
                    raw = raw[:-4]
                else:
                    raise HDF5Error(f"HDF5 filter {fid} is not supported")
            c = np.frombuffer(raw, dtype=dtype, count=int(np.prod(chunk_shape))).reshape(chunk_shape)
            sl = tuple(slice(o, min(o + cs, s)) for o, cs, s in zip(offsets, chunk_shape, shape))
            out[sl] = c[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out

    def chunks(self, node: int, rank: int):
        b = self.b
        if b[node:node + 4] != b"TREE" or b[node + 4] != 1:
            raise HDF5Error("bad chunk B-tree node")
        level, used = b[node + 5], self.u16(node + 6)
        key = 8 + 8 * (rank + 1)
        p = node + 8 + 2 * self.O
        for i in range(used):
            k = p + i * (key + self.O)
            csize, mask = self.u32(k), self.u32(k + 4)
            offsets = [int.from_bytes(b[k + 8 + 8 * j:k + 16 + 8 * j], "little") for j in range(rank)]
            child = self.off(k + key) + self.base
            if level > 0:
                yield from self.chunks(child, rank)
            else:
                yield offsets, child, csize, mask

    # ---- walk
    def walk(self, entry: dict, prefix: str, out: Dict[str, np.ndarray], depth: int = 0):
        if depth > 32:
            raise HDF5Error("group nesting too deep (cycle?)")
        msgs = self.messages(entry["header"] + self.base)
        btree = heap = None
        if entry.get("cache") == 1:
            btree, heap = entry["btree"], entry["heap"]
        for mtype, p, _ in msgs:
            if mtype == 0x0011:
                btree, heap = self.off(p), self.off(p + self.O)
        if btree is not None:
            for name, e in self.group_entries(btree + self.base, heap + self.base):
                self.walk(e, prefix + "/" + name, out, depth + 1)
            return
        a = self.dataset(msgs)
        if a is not None:
            out[prefix] = a


def read_hdf5(path: str) -> Dict[str, np.ndarray]:
    """Every numeric dataset of the file, keyed by its absolute path."""
    with open(path, "rb") as f:
        buf = f.read()
    h = _File(buf)
    out: Dict[str, np.ndarray] = {}
    h.walk(h.root, "", out)
    return out
