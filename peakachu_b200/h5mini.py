"""Minimal read-only HDF5 reader: just enough to open a cooler file.

The reference opens its contact map with ``cooler.Cooler(uri)``
(``score_chromosome.py:33-34``, ``score_genome.py:28-31``), i.e. through
h5py / libhdf5, neither of which exists in this image. A ``.cool`` is a plain
HDF5 file holding a few one-dimensional columns (``pixels/{bin1_id,bin2_id,
count}``, ``bins/{start,end,<weight>}``, ``chroms/{name,length}``,
``indexes/{chrom_offset,bin1_offset}``), chunked and gzip(+shuffle)
compressed. This module reads exactly that subset of the HDF5 file format, in
numpy + zlib:

* superblock versions 0-3 (with or without a user block),
* object headers version 1 and 2 (with continuation blocks),
* old-style groups (symbol table: v1 B-tree + local heap + SNOD nodes) and
  new-style *compact* groups (link messages); dense groups (fractal heap) raise,
* datasets with fixed-point, floating-point, fixed-length string and enum
  element types; layouts compact, contiguous and chunked (v1 B-tree chunk
  index; for the version-4 layout message: single-chunk, implicit and
  fixed-array indexes, paged or not); filters deflate, shuffle and fletcher32,
* scalar numeric / fixed-string attributes (enough for cooler's ``bin-size``).

Anything else raises ``H5Unsupported`` naming the feature, never a silent guess.
Partial reads (``Dataset.read(lo, hi)``) touch only the chunks that overlap the
range, so fetching one chromosome of a genome-wide file does not inflate the rest.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


_POOL = None


def _pool():
    """Threads that inflate chunks (shared by every file of the process)."""
    global _POOL
    if _POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        try:
            ncpu = len(os.sched_getaffinity(0))
        except AttributeError:
            ncpu = os.cpu_count() or 1
        _POOL = ThreadPoolExecutor(max_workers=max(1, min(16, ncpu)), thread_name_prefix="h5mini")
    return _POOL


class H5Error(Exception):
    pass


def lookup3(data: bytes, init: int = 0) -> int:
    """Bob Jenkins' lookup3 ``hashlittle``: the checksum HDF5 puts behind its version-2 metadata structures
    (``H5_checksum_metadata``). Known answer: ``lookup3(b"Four score and seven years ago") == 0x17770551``."""
    M = 0xFFFFFFFF

    def rot(x, k):
        return ((x << k) | (x >> (32 - k))) & M
    n = len(data)
    a = b = c = (0xDEADBEEF + n + init) & M
    i = 0
    while n - i > 12:
        a = (a + int.from_bytes(data[i:i + 4], "little")) & M
        b = (b + int.from_bytes(data[i + 4:i + 8], "little")) & M
        c = (c + int.from_bytes(data[i + 8:i + 12], "little")) & M
        a = (a - c) & M; a ^= rot(c, 4); c = (c + b) & M
        b = (b - a) & M; b ^= rot(a, 6); a = (a + c) & M
        c = (c - b) & M; c ^= rot(b, 8); b = (b + a) & M
        a = (a - c) & M; a ^= rot(c, 16); c = (c + b) & M
        b = (b - a) & M; b ^= rot(a, 19); a = (a + c) & M
        c = (c - b) & M; c ^= rot(b, 4); b = (b + a) & M
        i += 12
    tail = data[i:]
    if not tail:
        return c
    tail = tail + b"\0" * (12 - len(tail))
    a = (a + int.from_bytes(tail[0:4], "little")) & M
    b = (b + int.from_bytes(tail[4:8], "little")) & M
    c = (c + int.from_bytes(tail[8:12], "little")) & M
    c ^= b; c = (c - rot(b, 14)) & M
    a ^= c; a = (a - rot(c, 11)) & M
    b ^= a; b = (b - rot(a, 25)) & M
    c ^= b; c = (c - rot(b, 16)) & M
    a ^= c; a = (a - rot(c, 4)) & M
    b ^= a; b = (b - rot(a, 14)) & M
    c ^= b; c = (c - rot(b, 24)) & M
    return c


class H5Unsupported(H5Error):
    pass


class _Buf:
    """File bytes with HDF5's variable-width little-endian integers."""

    def __init__(self, data, base: int, so: int, sl: int):
        self.d, self.base, self.so, self.sl = data, base, so, sl

    def u(self, pos: int, size: int) -> int:
        return int.from_bytes(self.d[pos:pos + size], "little")

    def off(self, pos: int) -> int:
        v = self.u(pos, self.so)
        return UNDEF if v == (1 << (8 * self.so)) - 1 else v

    def length(self, pos: int) -> int:
        return self.u(pos, self.sl)


class File:
    def __init__(self, path: str):
        self.path = path
        self._fh = open(path, "rb")
        import mmap
        self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        d = self._mm
        pos = 0
        while True:                                  # the superblock sits at 0, 512, 1024, 2048, ...
            if pos + 8 > len(d):
                raise H5Error("%s: not an HDF5 file (no superblock signature)" % path)
            if d[pos:pos + 8] == SIGNATURE:
                break
            pos = 512 if pos == 0 else pos * 2
        ver = d[pos + 8]
        if ver in (0, 1):
            so, sl = d[pos + 13], d[pos + 14]
            p = pos + 24 + (4 if ver == 1 else 0)
            b = _Buf(d, 0, so, sl)
            base = b.off(p)
            # base, free-space, end-of-file, driver block; then the root symbol-table entry
            root_entry = p + 4 * so
            self._b = _Buf(d, base, so, sl)
            self._root_addr = self._b.off(root_entry + so)
        elif ver in (2, 3):
            so, sl = d[pos + 9], d[pos + 10]
            b = _Buf(d, 0, so, sl)
            base = b.off(pos + 12)
            self._b = _Buf(d, base, so, sl)
            self._root_addr = self._b.off(pos + 12 + 3 * so)
        else:
            raise H5Unsupported("superblock version %d" % ver)
        if self._b.base == UNDEF:
            self._b.base = pos
        self.root = Group(self, self._root_addr, "/")

    def close(self):
        try:
            self._mm.close()
        finally:
            self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __getitem__(self, path: str):
        return self.root[path]

    def __contains__(self, path: str):
        return path in self.root

    # ---- object headers ------------------------------------------------------
    def _messages(self, addr: int):
        """[(type, flags, data_pos, size)] of the object header at `addr` (relative)."""
        b = self._b
        d = b.d
        a = b.base + addr
        out = []
        if d[a:a + 4] == b"OHDR":
            if d[a + 4] != 2:
                raise H5Unsupported("object header version %d" % d[a + 4])
            fl = d[a + 5]
            p = a + 6
            if fl & 0x20:
                p += 16
            if fl & 0x10:
                p += 4
            w = 1 << (fl & 3)
            size0 = b.u(p, w)
            p += w
            blocks = [(p, p + size0)]                # checksum follows the chunk
            track = bool(fl & 0x04)
            while blocks:
                p, end = blocks.pop(0)
                while p + 4 <= end:
                    mtype = d[p]
                    msize = b.u(p + 1, 2)
                    mflags = d[p + 3]
                    p += 4 + (2 if track else 0)
                    if p + msize > end:
                        break
                    if mtype == 0x10:
                        co, cl = b.off(p), b.length(p + b.so)
                        ca = b.base + co
                        if d[ca:ca + 4] != b"OCHK":
                            raise H5Error("bad object header continuation at %d" % co)
                        blocks.append((ca + 4, ca + cl - 4))
                    elif mtype != 0:
                        out.append((mtype, mflags, p, msize))
                    p += msize
            return out
        if d[a] != 1:
            raise H5Error("no object header at address %d" % addr)
        nmsg = b.u(a + 2, 2)
        size0 = b.u(a + 8, 4)
        blocks = [(a + 16, a + 16 + size0)]
        while blocks and len(out) < nmsg + 64:
            p, end = blocks.pop(0)
            while p + 8 <= end:
                mtype = b.u(p, 2)
                msize = b.u(p + 2, 2)
                mflags = d[p + 4]
                p += 8
                if mtype == 0x10:
                    co, cl = b.off(p), b.length(p + b.so)
                    blocks.append((b.base + co, b.base + co + cl))
                elif mtype != 0:
                    out.append((mtype, mflags, p, msize))
                p += msize
        return out

    def _open(self, addr: int, name: str):
        msgs = self._messages(addr)
        types = {m[0] for m in msgs}
        if 0x08 in types:
            return Dataset(self, addr, name, msgs)
        if 0x11 in types or 0x02 in types or 0x06 in types:
            return Group(self, addr, name, msgs)
        if 0x03 in types:
            raise H5Unsupported("%s: committed datatype object" % name)
        return Group(self, addr, name, msgs)       # empty new-style group


def _shared(flags: int, what: str):
    if flags & 0x02:
        raise H5Unsupported("shared %s message" % what)


class _Attrs:
    """Attributes of an object: scalar / 1-D numeric and fixed-length strings."""

    def __init__(self, f: File, msgs, name: str):
        self._f, self._msgs, self._name = f, msgs, name
        self._cache = None

    def _load(self):
        if self._cache is not None:
            return self._cache
        b = self._f._b
        d = b.d
        out = {}
        for mtype, mflags, p, size in self._msgs:
            if mtype != 0x0C:
                continue
            try:
                ver = d[p]
                if ver == 1:
                    nlen, tlen, slen = b.u(p + 2, 2), b.u(p + 4, 2), b.u(p + 6, 2)
                    q = p + 8
                    pad = lambda n: (n + 7) & ~7
                elif ver in (2, 3):
                    if d[p + 1] & 3:
                        raise H5Unsupported("shared attribute datatype/dataspace")
                    nlen, tlen, slen = b.u(p + 2, 2), b.u(p + 4, 2), b.u(p + 6, 2)
                    q = p + 8 + (1 if ver == 3 else 0)
                    pad = lambda n: n
                else:
                    raise H5Unsupported("attribute message version %d" % ver)
                aname = bytes(d[q:q + nlen]).split(b"\0")[0].decode("utf-8", "replace")
                q += pad(nlen)
                dt = _parse_dtype(b, q)
                q += pad(tlen)
                shape = _parse_space(b, q)
                q += pad(slen)
                n = int(np.prod(shape)) if shape else 1
                val = np.frombuffer(bytes(d[q:q + n * dt.itemsize]), dtype=dt, count=n)
                if dt.kind == "S":
                    val = np.array([v.split(b"\0")[0] for v in val])
                out[aname] = val.reshape(shape) if shape else val[0]
            except H5Unsupported:
                continue                               # e.g. variable-length strings: not needed on this path
        self._cache = out
        return out

    def __getitem__(self, k):
        return self._load()[k]

    def __contains__(self, k):
        return k in self._load()

    def get(self, k, default=None):
        return self._load().get(k, default)

    def keys(self):
        return self._load().keys()


class Group:
    def __init__(self, f: File, addr: int, name: str, msgs=None):
        self._f, self._addr, self.name = f, addr, name
        self._msgs = msgs if msgs is not None else f._messages(addr)
        self._links = None
        self.attrs = _Attrs(f, self._msgs, name)

    def _load(self):
        if self._links is not None:
            return self._links
        b = self._f._b
        d = b.d
        links = {}
        for mtype, mflags, p, size in self._msgs:
            if mtype == 0x11:                        # symbol table: v1 B-tree + local heap
                _shared(mflags, "symbol table")
                btree, heap = b.off(p), b.off(p + b.so)
                ha = b.base + heap
                if d[ha:ha + 4] != b"HEAP":
                    raise H5Error("%s: bad local heap" % self.name)
                heap_data = b.base + b.off(ha + 8 + 2 * b.sl)
                for name_off, ohdr in self._walk_group_btree(btree):
                    s = heap_data + name_off
                    e = d.find(b"\0", s)
                    links[bytes(d[s:e]).decode("utf-8")] = ohdr
            elif mtype == 0x02:                      # link info: dense storage?
                fl = d[p + 1]
                q = p + 2 + (8 if fl & 1 else 0)
                if b.off(q) != UNDEF:
                    raise H5Unsupported("%s: densely stored links (fractal heap)" % self.name)
            elif mtype == 0x06:                      # link message (compact new-style group)
                _shared(mflags, "link")
                fl = d[p + 1]
                q = p + 2
                ltype = 0
                if fl & 0x08:
                    ltype = d[q]
                    q += 1
                if fl & 0x04:
                    q += 8
                if fl & 0x10:
                    q += 1
                w = 1 << (fl & 3)
                nlen = b.u(q, w)
                q += w
                lname = bytes(d[q:q + nlen]).decode("utf-8")
                q += nlen
                if ltype == 0:
                    links[lname] = b.off(q)
                # soft / external links are not followed
        self._links = links
        return links

    def _walk_group_btree(self, addr: int):
        b = self._f._b
        d = b.d
        a = b.base + addr
        if d[a:a + 4] != b"TREE" or d[a + 4] != 0:
            raise H5Error("%s: bad group B-tree node" % self.name)
        level, used = d[a + 5], b.u(a + 6, 2)
        p = a + 8 + 2 * b.so
        for i in range(used):
            child = b.off(p + b.sl + i * (b.sl + b.so))
            if level > 0:
                yield from self._walk_group_btree(child)
            else:
                s = b.base + child
                if d[s:s + 4] != b"SNOD":
                    raise H5Error("%s: bad symbol table node" % self.name)
                n = b.u(s + 6, 2)
                esz = 2 * b.so + 24
                for k in range(n):
                    e = s + 8 + k * esz
                    yield b.off(e), b.off(e + b.so)

    def keys(self):
        return list(self._load().keys())

    def __contains__(self, path: str):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __getitem__(self, path: str):
        node = self
        parts = [s for s in path.split("/") if s]
        if path.startswith("/"):
            node = self._f.root
        for i, part in enumerate(parts):
            if not isinstance(node, Group):
                raise KeyError(path)
            links = node._load()
            if part not in links:
                raise KeyError("%s: no object %r in group %s" % (self._f.path, part, node.name))
            node = self._f._open(links[part], node.name.rstrip("/") + "/" + part)
        return node


def _parse_dtype(b: _Buf, p: int) -> np.dtype:
    d = b.d
    cls, ver = d[p] & 0x0F, d[p] >> 4
    bits0 = d[p + 1]
    size = b.u(p + 4, 4)
    if cls == 0:                                     # fixed point
        order = ">" if bits0 & 1 else "<"
        return np.dtype("%s%s%d" % (order, "i" if bits0 & 0x08 else "u", size))
    if cls == 1:                                     # floating point (IEEE layouts only)
        if bits0 & 0x40:
            raise H5Unsupported("VAX floating point")
        if size not in (2, 4, 8):
            raise H5Unsupported("%d-byte floating point" % size)
        return np.dtype("%sf%d" % (">" if bits0 & 1 else "<", size))
    if cls == 3:                                     # fixed-length string
        return np.dtype("S%d" % size)
    if cls == 8:                                     # enum: values of the base integer type
        return _parse_dtype(b, p + 8)
    names = {2: "time", 4: "bitfield", 5: "opaque", 6: "compound", 7: "reference", 9: "variable-length", 10: "array"}
    raise H5Unsupported("datatype class %s (version %d)" % (names.get(cls, cls), ver))


def _parse_space(b: _Buf, p: int):
    d = b.d
    ver, rank = d[p], d[p + 1]
    if ver == 1:
        q = p + 8
    elif ver == 2:
        if d[p + 3] == 2:
            return (0,)                              # null dataspace
        q = p + 4
    else:
        raise H5Unsupported("dataspace version %d" % ver)
    return tuple(b.length(q + i * b.sl) for i in range(rank))


class Dataset:
    def __init__(self, f: File, addr: int, name: str, msgs):
        self._f, self.name = f, name
        self.attrs = _Attrs(f, msgs, name)
        b = f._b
        d = b.d
        self.filters = []                            # [(id, client_values)]
        self.shape = None
        self.dtype = None
        self._layout = None
        for mtype, mflags, p, size in msgs:
            if mtype == 0x01:
                _shared(mflags, "dataspace")
                self.shape = _parse_space(b, p)
            elif mtype == 0x03:
                _shared(mflags, "datatype")
                self.dtype = _parse_dtype(b, p)
            elif mtype == 0x0B:
                _shared(mflags, "filter pipeline")
                self.filters = self._parse_filters(p)
            elif mtype == 0x08:
                self._layout = self._parse_layout(p)
        if self.shape is None or self.dtype is None or self._layout is None:
            raise H5Error("%s: incomplete dataset header" % name)
        self._chunks = None
        self._firsts = None

    def __len__(self):
        return self.shape[0] if self.shape else 1

    # ---- header messages -------------------------------------------------
    def _parse_filters(self, p: int):
        b = self._f._b
        d = b.d
        ver, n = d[p], d[p + 1]
        q = p + (8 if ver == 1 else 2)
        out = []
        for _ in range(n):
            fid = b.u(q, 2)
            q += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = b.u(q, 2)
                q += 2
            q += 2                                   # flags
            nval = b.u(q, 2)
            q += 2
            q += ((nlen + 7) & ~7) if ver == 1 else nlen
            vals = [b.u(q + 4 * i, 4) for i in range(nval)]
            q += 4 * nval
            if ver == 1 and nval % 2:
                q += 4
            out.append((fid, vals))
        return out

    def _parse_layout(self, p: int):
        b = self._f._b
        d = b.d
        ver = d[p]
        if ver in (1, 2):
            ndim, cls = d[p + 1], d[p + 2]
            q = p + 8
            addr = None
            if cls != 0:
                addr = b.off(q)
                q += b.so
            dims = [b.u(q + 4 * i, 4) for i in range(ndim)]
            q += 4 * ndim
            if cls == 2:
                return ("btree1", addr, dims[:-1] if len(dims) > len(self.shape or ()) else dims)
            if cls == 1:
                return ("contiguous", addr, None)
            size = b.u(q, 4)
            return ("compact", q + 4, size)
        if ver == 3:
            cls = d[p + 1]
            if cls == 0:
                return ("compact", p + 4, b.u(p + 2, 2))
            if cls == 1:
                return ("contiguous", b.off(p + 2), b.length(p + 2 + b.so))
            if cls == 2:
                ndim = d[p + 2]
                addr = b.off(p + 3)
                dims = [b.u(p + 3 + b.so + 4 * i, 4) for i in range(ndim)]
                return ("btree1", addr, dims[:-1])
            raise H5Unsupported("%s: layout class %d" % (self.name, cls))
        if ver == 4:
            cls = d[p + 1]
            if cls == 0:
                return ("compact", p + 4, b.u(p + 2, 2))
            if cls == 1:
                return ("contiguous", b.off(p + 2), b.length(p + 2 + b.so))
            if cls != 2:
                raise H5Unsupported("%s: layout class %d (virtual)" % (self.name, cls))
            fl, ndim, enc = d[p + 2], d[p + 3], d[p + 4]
            q = p + 5
            dims = [b.u(q + enc * i, enc) for i in range(ndim)]
            q += enc * ndim
            itype = d[q]
            q += 1
            dims = dims[:-1]
            if itype == 1:                           # single chunk
                fsize, fmask = None, 0
                if fl & 0x02:
                    fsize, fmask = b.length(q), b.u(q + b.sl, 4)
                    q += b.sl + 4
                return ("single", b.off(q), dims, fsize, fmask)
            if itype == 2:                           # implicit: chunks stored back to back, no filters
                return ("implicit", b.off(q), dims)
            if itype == 3:                           # fixed array
                return ("farray", b.off(q + 1), dims)
            names = {4: "extensible array", 5: "version-2 B-tree"}
            raise H5Unsupported("%s: chunk index type %s" % (self.name, names.get(itype, itype)))
        raise H5Unsupported("%s: data layout message version %d" % (self.name, ver))

    # ---- chunk index -----------------------------------------------------
    def _walk_chunk_btree(self, addr: int, out):
        b = self._f._b
        d = b.d
        a = b.base + addr
        if d[a:a + 4] != b"TREE" or d[a + 4] != 1:
            raise H5Error("%s: bad chunk B-tree node" % self.name)
        level, used = d[a + 5], b.u(a + 6, 2)
        rank = len(self.shape)
        ksz = 8 + 8 * (rank + 1)
        p = a + 8 + 2 * b.so
        for i in range(used):
            k = p + i * (ksz + b.so)
            child = b.off(k + ksz)
            if level > 0:
                self._walk_chunk_btree(child, out)
            else:
                csize, fmask = b.u(k, 4), b.u(k + 4, 4)
                offs = tuple(b.u(k + 8 + 8 * j, 8) for j in range(rank))
                out.append((offs, child, csize, fmask))

    def _chunk_list(self):
        """[(element offsets, address, stored size, filter mask)] sorted by offsets."""
        if self._chunks is not None:
            return self._chunks
        b = self._f._b
        d = b.d
        kind = self._layout[0]
        out = []
        itemsize = self.dtype.itemsize
        cdims = self._layout[2]
        cbytes = int(np.prod(cdims)) * itemsize if cdims else 0
        grid = [(-(-s // c)) for s, c in zip(self.shape, cdims)] if cdims else []
        if kind == "btree1":
            if self._layout[1] != UNDEF:
                self._walk_chunk_btree(self._layout[1], out)
        elif kind == "single":
            _, addr, dims, fsize, fmask = self._layout
            if addr != UNDEF:
                out.append((tuple(0 for _ in dims), addr, fsize if fsize is not None else cbytes, fmask))
        elif kind == "implicit":
            addr = self._layout[1]
            if addr != UNDEF:
                for i, idx in enumerate(np.ndindex(*grid)):
                    out.append((tuple(i_ * c for i_, c in zip(idx, cdims)), addr + i * cbytes, cbytes, 0))
        elif kind == "farray":
            addr = self._layout[1]
            if addr != UNDEF:
                a = b.base + addr
                if d[a:a + 4] != b"FAHD":
                    raise H5Error("%s: bad fixed array header" % self.name)
                client, esize, pbits = d[a + 5], d[a + 6], d[a + 7]
                nent = b.length(a + 8)
                db = b.base + b.off(a + 8 + b.sl)
                if d[db:db + 4] != b"FADB":
                    raise H5Error("%s: bad fixed array data block" % self.name)
                def entry(e, idx):
                    ca = b.off(e)
                    if ca == UNDEF:
                        return
                    if client == 1:                  # filtered chunks: address, size, mask
                        sw = esize - b.so - 4
                        out.append((tuple(i_ * c for i_, c in zip(idx, cdims)), ca, b.u(e + b.so, sw), b.u(e + b.so + sw, 4)))
                    else:
                        out.append((tuple(i_ * c for i_, c in zip(idx, cdims)), ca, cbytes, 0))
                cells = list(np.ndindex(*grid))
                if len(cells) > nent:
                    raise H5Error("%s: fixed array holds %d entries for %d chunks" % (self.name, nent, len(cells)))
                if len(self.shape) > 1 and nent != len(cells):
                    # sized for the maximum dimensions: entries are numbered over that grid, not the current one
                    raise H5Unsupported("%s: fixed array of a multi-dimensional dataset below its maximum size" % self.name)
                per_page = 1 << pbits
                if nent <= per_page:                 # the elements sit in the data block itself
                    q = db + 6 + b.so
                    for i, idx in enumerate(cells):
                        entry(q + i * esize, idx)
                else:
                    # paged data block: a bitmap of initialised pages closes the prefix; the pages follow it back
                    # to back, 2^pbits elements + a checksum each (the last one holds the remainder). Every
                    # checksum on the way is verified (lookup3), so a file laid out differently from this
                    # reading of the format fails here instead of yielding chunk addresses from the wrong bytes.
                    npages = -(-nent // per_page)
                    nbitmap = (npages + 7) // 8
                    bm = db + 6 + b.so
                    prefix_end = bm + nbitmap
                    if b.u(prefix_end, 4) != lookup3(bytes(d[db:prefix_end])):
                        raise H5Error("%s: fixed array data block checksum mismatch" % self.name)
                    first_page = prefix_end + 4
                    page_bytes = per_page * esize + 4
                    for pg in range(npages):
                        if not (d[bm + pg // 8] >> (7 - pg % 8)) & 1:        # page never written: all chunks absent
                            continue
                        lo_e = pg * per_page
                        cnt = min(per_page, nent - lo_e)
                        q = first_page + pg * page_bytes
                        if q + cnt * esize + 4 > len(d):
                            raise H5Error("%s: fixed array page beyond the end of the file" % self.name)
                        if b.u(q + cnt * esize, 4) != lookup3(bytes(d[q:q + cnt * esize])):
                            raise H5Error("%s: fixed array page %d checksum mismatch" % (self.name, pg))
                        for i in range(lo_e, min(lo_e + cnt, len(cells))):
                            entry(q + (i - lo_e) * esize, cells[i])
        out.sort(key=lambda t: t[0])
        self._chunks = out
        return out

    def _decode_chunk(self, addr: int, csize: int, fmask: int) -> np.ndarray:
        b = self._f._b
        raw = bytes(b.d[b.base + addr: b.base + addr + csize])
        for i in range(len(self.filters) - 1, -1, -1):   # undo the pipeline, last filter first
            if fmask & (1 << i):
                continue
            fid, vals = self.filters[i]
            if fid == 1:
                # sized output buffer: one inflate call, all of it outside the GIL (the chunks of a large
                # read are inflated on a thread pool)
                raw = zlib.decompress(raw, 15, int(np.prod(self._layout[2])) * self.dtype.itemsize + 16)
            elif fid == 2:
                # byte planes back to elements: one strided store per plane (a transposed copy through
                # .T.tobytes() is ten times slower and dominated the read)
                es = vals[0] if vals else self.dtype.itemsize
                n = len(raw) // es
                planes = np.frombuffer(raw, dtype=np.uint8, count=n * es).reshape(es, n)
                body = np.empty((n, es), dtype=np.uint8)
                for j in range(es):
                    body[:, j] = planes[j]
                raw = body.reshape(-1).data if n * es == len(raw) else body.tobytes() + bytes(raw[n * es:])
            elif fid == 3:
                raw = raw[:-4]
            else:
                names = {4: "szip", 5: "nbit", 6: "scaleoffset", 32000: "lzf", 32001: "blosc"}
                raise H5Unsupported("%s: filter %s" % (self.name, names.get(fid, fid)))
        cdims = self._layout[2]
        return np.frombuffer(raw, dtype=self.dtype, count=int(np.prod(cdims))).reshape(cdims)

    def _decode_native(self, hits, lo: int, hi: int, out: np.ndarray) -> bool:
        """Large reads of 1-D columns (a chromosome's pixel columns) go to the library's threaded decoder
        (``pk_h5_decode_chunks``: zlib + un-shuffle from the mapped file into ``out``). Returns False when the
        read is small, the filter pipeline is not [shuffle,] deflate [, fletcher32], or the library is not built --
        the Python loop below then does the same work."""
        if len(self.shape) != 1 or len(hits) < 4 or sum(rec[2] for rec in hits) < (1 << 20) or not self.filters:
            return False
        if self.dtype.byteorder == ">" or self.dtype.kind not in "iuf" or any(rec[3] for rec in hits):
            return False
        ids = [fid for fid, _ in self.filters]
        if ids not in ([1], [2, 1], [1, 3], [2, 1, 3]):
            return False
        if 2 in ids:
            vals = dict(self.filters)[2]
            if vals and vals[0] != self.dtype.itemsize:
                return False
        try:
            from . import _lib
            L = _lib.lib()
        except Exception:
            return False
        import ctypes as C
        b = self._f._b
        view = np.frombuffer(b.d, dtype=np.uint8)              # the mapped file, no copy
        off = np.array([b.base + rec[1] for rec in hits], dtype=np.int64)
        size = np.array([rec[2] for rec in hits], dtype=np.int64)
        first = np.array([rec[0][0] for rec in hits], dtype=np.int64)
        if int((off + size).max()) > view.size:
            return False
        rc = L.pk_h5_decode_chunks(C.c_void_p(view.ctypes.data), len(hits), _lib.ptr(off, _lib.c_i64p), _lib.ptr(size, _lib.c_i64p),
                                   _lib.ptr(first, _lib.c_i64p), int(self._layout[2][0]), self.dtype.itemsize, 1, int(2 in ids),
                                   int(3 in ids), int(lo), int(hi), C.c_void_p(out.ctypes.data), 0)
        if rc != 0:
            raise H5Error("%s: %s" % (self.name, L.pk_last_error().decode("utf-8", "replace")))
        return True

    # ---- reads -----------------------------------------------------------
    def read(self, lo: int = 0, hi: int | None = None) -> np.ndarray:
        """Rows [lo, hi) along the first axis, as a native-endian array."""
        shape = self.shape if self.shape else (1,)
        n0 = shape[0]
        hi = n0 if hi is None else min(hi, n0)
        lo = max(0, min(lo, hi))
        b = self._f._b
        kind = self._layout[0]
        rest = shape[1:]
        rowbytes = int(np.prod(rest)) * self.dtype.itemsize if rest else self.dtype.itemsize
        out_shape = (hi - lo,) + tuple(rest)
        native = self.dtype.newbyteorder("=")
        if kind in ("contiguous", "compact"):
            if kind == "contiguous":
                if self._layout[1] == UNDEF:         # never written: fill value (zeros)
                    return np.zeros(out_shape, dtype=native)
                start = b.base + self._layout[1]
            else:
                start = self._layout[1]
            buf = bytes(b.d[start + lo * rowbytes: start + hi * rowbytes])
            return np.frombuffer(buf, dtype=self.dtype).reshape(out_shape).astype(native, copy=True)
        out = np.zeros(out_shape, dtype=native)
        cdims = self._layout[2]
        chunks = self._chunk_list()
        if len(shape) == 1:
            # sorted by first element: the chunks of a range by bisection (a genome-wide cooler column has tens of
            # thousands of chunks and is read once per chromosome)
            if self._firsts is None:
                self._firsts = np.array([rec[0][0] for rec in chunks], dtype=np.int64)
            i0 = int(np.searchsorted(self._firsts, lo - cdims[0], side="right"))
            i1 = int(np.searchsorted(self._firsts, hi, side="left"))
            hits = chunks[i0:i1]
        else:
            hits = [rec for rec in chunks if rec[0][0] + cdims[0] > lo and rec[0][0] < hi]
        if self._decode_native(hits, lo, hi, out):
            return out
        # zlib releases the GIL: inflate the chunks of a large read on a few threads (a chromosome of a
        # genome-wide cooler is hundreds of megabytes of pixel columns)
        if len(hits) >= 4 and sum(rec[2] for rec in hits) >= (1 << 20) and self.filters:
            decoded = _pool().map(lambda rec: self._decode_chunk(rec[1], rec[2], rec[3]), hits)
        else:
            decoded = (self._decode_chunk(rec[1], rec[2], rec[3]) for rec in hits)
        for (offs, addr, csize, fmask), chunk in zip(hits, decoded):
            c_lo, c_hi = offs[0], offs[0] + cdims[0]
            src = [slice(max(lo, c_lo) - c_lo, min(hi, c_hi) - c_lo)]
            dst = [slice(max(lo, c_lo) - lo, min(hi, c_hi) - lo)]
            for ax in range(1, len(shape)):
                e = min(offs[ax] + cdims[ax], shape[ax])
                src.append(slice(0, e - offs[ax]))
                dst.append(slice(offs[ax], e))
            out[tuple(dst)] = chunk[tuple(src)]
        return out

    def __getitem__(self, key):
        if key is Ellipsis or key == ():
            return self.read() if self.shape else self.read()[0]
        if isinstance(key, slice):
            lo, hi, step = key.indices(len(self))
            if step != 1:
                return self.read()[key]
            return self.read(lo, hi)
        if isinstance(key, (int, np.integer)):
            k = int(key) + (len(self) if key < 0 else 0)
            return self.read(k, k + 1)[0]
        return self.read()[key]
