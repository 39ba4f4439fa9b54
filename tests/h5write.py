"""Test-only HDF5 writer: emits a cooler-layout file the way h5py / libhdf5 1.10 lays one
out with its defaults (superblock version 0, version-1 object headers, symbol-table groups
with a local heap, chunked datasets indexed by version-1 B-trees, version-1 filter pipeline
messages with shuffle + deflate), written from the HDF5 file-format specification. It exists
because nothing in this image can write HDF5; the reader under test (`peakachu_b200.h5mini`)
is additionally checked against the one libhdf5-written file the image holds (a MATLAB 7.3
file in scipy's test data).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


class Writer:
    def __init__(self, userblock: int = 0):
        self.buf = bytearray(userblock + 96)           # user block + superblock placeholder
        self.base = userblock

    def alloc(self, data: bytes) -> int:
        """Append 8-byte aligned; returns the address relative to the base."""
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf) - self.base
        self.buf += data
        return addr

    # ---- datatypes -------------------------------------------------------
    @staticmethod
    def dtype_msg(dt: np.dtype, enum=None) -> bytes:
        dt = np.dtype(dt)
        if enum is not None:                           # {name: value} over an integer base type
            base = Writer.dtype_msg(dt)
            n = len(enum)
            head = struct.pack("<BBBBI", 0x18, n & 0xFF, (n >> 8) & 0xFF, 0, dt.itemsize)
            names = b"".join(_pad8(k.encode() + b"\0") for k in enum)
            vals = np.array(list(enum.values()), dtype=dt).tobytes()
            return head + base + names + vals
        big = 1 if dt.byteorder == ">" else 0
        if dt.kind in "iu":
            return struct.pack("<BBBBIHH", 0x10, big | (0x08 if dt.kind == "i" else 0), 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
        if dt.kind == "f":
            if dt.itemsize == 8:
                return struct.pack("<BBBBIHHBBBBI", 0x11, big | 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
            if dt.itemsize == 4:
                return struct.pack("<BBBBIHHBBBBI", 0x11, big | 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
        if dt.kind == "S":
            return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)     # null-padded ASCII
        raise NotImplementedError(dt)

    @staticmethod
    def space_msg(shape, unlimited=False) -> bytes:
        m = struct.pack("<BBBBI", 1, len(shape), 1 if unlimited else 0, 0, 0)
        m += b"".join(struct.pack("<Q", s) for s in shape)
        if unlimited:
            m += b"".join(struct.pack("<Q", UNDEF) for _ in shape)
        return m

    @staticmethod
    def attr_msg(name: str, value) -> bytes:
        v = np.asarray(value)
        nm = name.encode() + b"\0"
        dtm = Writer.dtype_msg(v.dtype)
        spm = Writer.space_msg(v.shape)
        return (struct.pack("<BBHHH", 1, 0, len(nm), len(dtm), len(spm)) + _pad8(nm) + _pad8(dtm) + _pad8(spm) + v.tobytes())

    def object_header(self, messages, split_after=None) -> int:
        """messages: [(type, bytes)]. With split_after = k the messages after the k-th go to a
        continuation block (as libhdf5 does when attributes are added later)."""
        def pack(msgs):
            out = b""
            for t, data in msgs:
                data = _pad8(data)
                out += struct.pack("<HHBBBB", t, len(data), 0, 0, 0, 0) + data
            return out
        nmsg = len(messages)
        if split_after is not None and split_after < len(messages):
            tail = pack(messages[split_after:])
            tail_addr = self.alloc(tail)
            head_msgs = list(messages[:split_after]) + [(0x10, struct.pack("<QQ", tail_addr, len(tail)))]
            nmsg += 1
        else:
            head_msgs = messages
        body = pack(head_msgs)
        return self.alloc(struct.pack("<BBHII", 1, 0, nmsg, 1, len(body)) + b"\0" * 4 + body)

    # ---- datasets ----------------------------------------------------------
    def dataset(self, arr, chunk=None, gzip=None, shuffle=False, fletcher=False, unlimited=False,
                enum=None, attrs=None, compact=False, skip_filter_on=None) -> int:
        arr = np.ascontiguousarray(arr)
        msgs = [(0x01, self.space_msg(arr.shape, unlimited)), (0x03, self.dtype_msg(arr.dtype, enum)),
                (0x05, struct.pack("<BBBB", 2, 2, 2, 0))]
        es = arr.dtype.itemsize
        if chunk is None:
            if compact:
                raw = arr.tobytes()
                msgs.append((0x08, struct.pack("<BBH", 3, 0, len(raw)) + raw))
            else:
                addr = self.alloc(arr.tobytes()) if arr.size else UNDEF
                msgs.append((0x08, struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)))
        else:
            if isinstance(chunk, int):
                chunk = (chunk,)
            assert arr.ndim == 1 and len(chunk) == 1, "the test writer chunks one-dimensional data"
            filters = []
            if shuffle:
                filters.append((2, b"shuffle\0", [es]))
            if gzip is not None:
                filters.append((1, b"deflate\0", [gzip]))
            if fletcher:
                filters.append((3, b"fletcher32\0", []))
            if filters:
                fm = struct.pack("<BBHI", 1, len(filters), 0, 0)
                for fid, name, vals in filters:
                    name = _pad8(name)
                    fm += struct.pack("<HHHH", fid, len(name), 1, len(vals)) + name
                    fm += b"".join(struct.pack("<I", v) for v in vals)
                    if len(vals) % 2:
                        fm += b"\0" * 4
                msgs.append((0x0B, fm))
            c = chunk[0]
            records = []
            for k, o in enumerate(range(0, arr.shape[0], c)):
                blk = np.zeros(c, dtype=arr.dtype)
                part = arr[o:o + c]
                blk[:part.size] = part
                raw = blk.tobytes()
                mask = 0
                for i, (fid, _, vals) in enumerate(filters):
                    if skip_filter_on is not None and k == skip_filter_on and fid == 1:
                        mask |= 1 << i                  # libhdf5 skips an optional filter that fails
                        continue
                    if fid == 2:
                        raw = np.frombuffer(raw, dtype=np.uint8).reshape(c, es).T.tobytes()
                    elif fid == 1:
                        raw = zlib.compress(raw, vals[0])
                    elif fid == 3:
                        raw = raw + struct.pack("<I", zlib.adler32(raw) & 0xFFFFFFFF)    # value is not checked by readers here
                records.append((o, self.alloc(raw), len(raw), mask))
            btree = self._chunk_btree(records, arr.shape[0], c, es) if records else UNDEF
            msgs.append((0x08, struct.pack("<BBBQII", 3, 2, 2, btree, c, es)))
        split = None
        if attrs:
            split = len(msgs)
            for k, v in attrs.items():
                msgs.append((0x0C, self.attr_msg(k, v)))
        return self.object_header(msgs, split_after=split)

    def _chunk_btree(self, records, n, c, es, K=32) -> int:
        ksz = 8 + 8 * 2

        def node(level, entries, end_key):
            # entries: [(key tuple (size, mask, offset), child address)]
            body = b"TREE" + struct.pack("<BBHQQ", 1, level, len(entries), UNDEF, UNDEF)
            for (size, mask, off), child in entries:
                body += struct.pack("<IIQQ", size, mask, off, 0) + struct.pack("<Q", child)
            body += struct.pack("<IIQQ", 0, 0, end_key, 0)
            full = 8 + 16 + (2 * K + 1) * ksz + 2 * K * 8
            return self.alloc(body + b"\0" * (full - len(body)))

        level = 0
        entries = [((size, mask, off), addr) for off, addr, size, mask in records]
        while True:
            groups = [entries[i:i + 2 * K] for i in range(0, len(entries), 2 * K)]
            parents = []
            for gi, g in enumerate(groups):
                end = groups[gi + 1][0][0][2] if gi + 1 < len(groups) else ((n + c - 1) // c) * c
                parents.append((g[0][0], node(level, g, end)))
            if len(parents) == 1:
                return parents[0][1]
            entries = parents
            level += 1

    # ---- groups --------------------------------------------------------------
    def group(self, children: dict, attrs=None, leaf_k: int = 4) -> int:
        names = sorted(children)
        heap = bytearray(8)                            # offset 0: the empty string
        offs = {}
        for nm in names:
            offs[nm] = len(heap)
            heap += _pad8(nm.encode() + b"\0")
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap), UNDEF, heap_data))
        snods = []
        for i in range(0, max(len(names), 1), 2 * leaf_k):
            part = names[i:i + 2 * leaf_k]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for nm in part:
                body += struct.pack("<QQII", offs[nm], children[nm], 0, 0) + b"\0" * 16
            body += b"\0" * (8 + 2 * leaf_k * 40 - len(body))
            snods.append((offs[part[-1]] if part else 0, self.alloc(body)))
        body = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", 0)
        for last, addr in snods:
            body += struct.pack("<QQ", addr, last)
        btree = self.alloc(body + b"\0" * 512)
        msgs = [(0x11, struct.pack("<QQ", btree, heap_addr))]
        split = None
        if attrs:
            split = 1
            for k, v in attrs.items():
                msgs.append((0x0C, self.attr_msg(k, v)))
        return self.object_header(msgs, split_after=split)

    def finish(self, root_addr: int, path: str):
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack("<QQQQ", self.base, UNDEF, len(self.buf) - self.base, UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 0, 0) + b"\0" * 16
        assert len(sb) == 96
        self.buf[self.base:self.base + 96] = sb
        with open(path, "wb") as fh:
            fh.write(self.buf)


def write_cool(path: str, chroms, binsize: int, weight_name: str = "weight", trans=None, group: str = "",
               chunk: int = 4096, userblock: int = 0, extra_bins: dict | None = None, latest: bool = False):
    """Write synth.SynthChrom-like objects (name, n, bin1, bin2, count, weights) as a cooler file:
    cooler's schema version 3 columns and dtypes (int64 bin ids, int32 counts and coordinates, float64
    weights, enum chromosome ids, fixed-length ASCII names), gzip level 6 + shuffle like cooler's
    writer. `trans` adds inter-chromosomal pixels [(genome-wide bin1, bin2, count)] that a reader of
    one chromosome must drop. `group` nests the cooler (mcool layout: 'resolutions/10000')."""
    chroms = list(chroms)
    nb = np.array([c.n for c in chroms], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
    b1 = np.concatenate([c.bin1.astype(np.int64) + off[i] for i, c in enumerate(chroms)])
    b2 = np.concatenate([c.bin2.astype(np.int64) + off[i] for i, c in enumerate(chroms)])
    cnt = np.concatenate([c.count for c in chroms]).astype(np.int32)
    if trans is not None and len(trans):
        t = np.asarray(trans, dtype=np.int64)
        b1 = np.concatenate([b1, t[:, 0]]); b2 = np.concatenate([b2, t[:, 1]]); cnt = np.concatenate([cnt, t[:, 2].astype(np.int32)])
    order = np.lexsort((b2, b1))
    b1, b2, cnt = b1[order], b2[order], cnt[order]
    nbins = int(off[-1])
    bin1_offset = np.searchsorted(b1, np.arange(nbins + 1), side="left").astype(np.int64)
    w = np.concatenate([c.weights for c in chroms]).astype(np.float64)
    start = np.concatenate([np.arange(n, dtype=np.int32) * binsize for n in nb]).astype(np.int32)
    lengths = (nb * binsize - binsize // 3).astype(np.int32)         # last bin of a chromosome is short
    end = np.concatenate([np.minimum(np.arange(1, n + 1, dtype=np.int64) * binsize, lengths[i]) for i, n in enumerate(nb)]).astype(np.int32)
    chrom_id = np.repeat(np.arange(len(chroms), dtype=np.int32), nb)
    names = np.array([c.name.encode() for c in chroms], dtype="S%d" % max(len(c.name) for c in chroms))

    W = (Writer2 if latest else Writer)(userblock=userblock)      # Writer2 is defined below (resolved at call time)
    z = dict(chunk=chunk, gzip=6, shuffle=True, unlimited=True)
    g_chroms = W.group({"name": W.dataset(names, chunk=max(1, len(chroms)), gzip=6, shuffle=True),
                        "length": W.dataset(lengths, chunk=max(1, len(chroms)), gzip=6, shuffle=True)})
    bins = {"chrom": W.dataset(chrom_id, enum={c.name: i for i, c in enumerate(chroms)}, **z),
            "start": W.dataset(start, **z), "end": W.dataset(end, **z),
            weight_name: W.dataset(w, attrs={"ignore_diags": np.int64(2), "converged": np.uint8(1)}, skip_filter_on=1, **z)}
    for k, v in (extra_bins or {}).items():
        # a column named like hic2cool's KR / VC carries cooler's `divisive_weights` attribute
        at = {"divisive_weights": np.uint8(1)} if k in ("KR", "VC", "VC_SQRT") else None
        bins[k] = W.dataset(v, attrs=at, **z)
    g_bins = W.group(bins)
    g_pixels = W.group({"bin1_id": W.dataset(b1, **z), "bin2_id": W.dataset(b2, **z),
                        "count": W.dataset(cnt, fletcher=True, **z)})
    g_idx = W.group({"chrom_offset": W.dataset(off), "bin1_offset": W.dataset(bin1_offset, **z)})
    top = W.group({"chroms": g_chroms, "bins": g_bins, "pixels": g_pixels, "indexes": g_idx},
                  attrs={"bin-size": np.int64(binsize), "nbins": np.int64(nbins), "nnz": np.int64(b1.size),
                         "format": np.bytes_(b"HDF5::Cooler"), "format-version": np.int64(3)})
    for part in reversed([s for s in group.split("/") if s]):
        top = W.group({part: top})
    W.finish(top, path)
    return dict(bin1=b1, bin2=b2, count=cnt, chrom_offset=off, bin1_offset=bin1_offset, weights=w)


class Writer2(Writer):
    """The same objects in HDF5's "latest" file format (what h5py writes with libver='latest'):
    superblock version 2, version-2 object headers (OHDR / OCHK), groups as link messages,
    version-2 dataspaces and filter pipelines, version-3 attributes, version-4 data layouts with
    single-chunk, implicit and fixed-array chunk indexes. Checksum fields are written as zero (the
    reader under test does not verify them)."""

    def __init__(self, userblock: int = 0):
        super().__init__(userblock)
        del self.buf[self.base + 48:]                  # the version-2 superblock is 48 bytes

    @staticmethod
    def space_msg(shape, unlimited=False) -> bytes:
        m = struct.pack("<BBBB", 2, len(shape), 1 if unlimited else 0, 1 if len(shape) else 0)
        m += b"".join(struct.pack("<Q", s) for s in shape)
        if unlimited:
            m += b"".join(struct.pack("<Q", UNDEF) for _ in shape)
        return m

    @staticmethod
    def attr_msg(name: str, value) -> bytes:
        v = np.asarray(value)
        nm = name.encode() + b"\0"
        dtm = Writer.dtype_msg(v.dtype)
        spm = Writer2.space_msg(v.shape)
        return struct.pack("<BBHHHB", 3, 0, len(nm), len(dtm), len(spm), 0) + nm + dtm + spm + v.tobytes()

    def object_header(self, messages, split_after=None) -> int:
        def pack(msgs):
            return b"".join(struct.pack("<BHB", t, len(data), 0) + data for t, data in msgs)
        if split_after is not None and split_after < len(messages):
            tail = b"OCHK" + pack(messages[split_after:]) + b"\0" * 4
            tail_addr = self.alloc(tail)
            head = pack(list(messages[:split_after]) + [(0x10, struct.pack("<QQ", tail_addr, len(tail)))])
        else:
            head = pack(messages)
        # flags 0x22: chunk-0 size in 4 bytes, times stored
        return self.alloc(b"OHDR" + struct.pack("<BB", 2, 0x22) + b"\0" * 16 + struct.pack("<I", len(head)) + head + b"\0" * 4)

    def dataset(self, arr, chunk=None, gzip=None, shuffle=False, fletcher=False, unlimited=False,
                enum=None, attrs=None, compact=False, skip_filter_on=None, page_bits=10, absent_pages=()) -> int:
        """``page_bits``: a fixed-array chunk index with more than 2**page_bits entries is written paged (libhdf5's
        default is 10); ``absent_pages``: pages left uninitialised (their chunks do not exist and read as zeros)."""
        arr = np.ascontiguousarray(arr)
        msgs = [(0x01, self.space_msg(arr.shape, unlimited)), (0x03, self.dtype_msg(arr.dtype, enum)),
                (0x05, struct.pack("<BB", 3, 0x09))]
        es = arr.dtype.itemsize
        if chunk is None:
            if compact:
                raw = arr.tobytes()
                msgs.append((0x08, struct.pack("<BBH", 4, 0, len(raw)) + raw))
            else:
                addr = self.alloc(arr.tobytes()) if arr.size else UNDEF
                msgs.append((0x08, struct.pack("<BBQQ", 4, 1, addr, arr.nbytes)))
        else:
            c = chunk if isinstance(chunk, int) else chunk[0]
            assert arr.ndim == 1
            filters = ([(2, [es])] if shuffle else []) + ([(1, [gzip])] if gzip is not None else []) + ([(3, [])] if fletcher else [])
            if filters:
                fm = struct.pack("<BB", 2, len(filters))
                for fid, vals in filters:
                    fm += struct.pack("<HHH", fid, 1, len(vals)) + b"".join(struct.pack("<I", v) for v in vals)
                msgs.append((0x0B, fm))
            records = []
            for k, o in enumerate(range(0, arr.shape[0], c)):
                blk = np.zeros(c, dtype=arr.dtype)
                part = arr[o:o + c]
                blk[:part.size] = part
                raw = blk.tobytes()
                mask = 0
                for i, (fid, vals) in enumerate(filters):
                    if skip_filter_on is not None and k == skip_filter_on and fid == 1:
                        mask |= 1 << i
                        continue
                    if fid == 2:
                        raw = np.frombuffer(raw, dtype=np.uint8).reshape(c, es).T.tobytes()
                    elif fid == 1:
                        raw = zlib.compress(raw, vals[0])
                    elif fid == 3:
                        raw = raw + b"\0" * 4
                records.append((self.alloc(raw), len(raw), mask))
            head = struct.pack("<BBBBB", 4, 2, 0x02 if (filters and len(records) == 1) else 0, 2, 4) + struct.pack("<II", c, es)
            if len(records) <= 1:                      # single-chunk index
                addr, size, mask = records[0] if records else (UNDEF, 0, 0)
                lay = head + struct.pack("<B", 1) + (struct.pack("<QI", size, mask) if filters and records else b"") + struct.pack("<Q", addr)
                if not records:
                    lay = struct.pack("<BBBBB", 4, 2, 0, 2, 4) + struct.pack("<II", c, es) + struct.pack("<BQ", 1, UNDEF)
            elif not filters:                          # implicit index: the chunks lie back to back
                first = self.alloc(b"".join(bytes(self.buf[self.base + a: self.base + a + n]) for a, n, _ in records))
                lay = head + struct.pack("<BQ", 2, first)
            else:                                      # fixed array of (address, size, mask) entries
                from peakachu_b200.h5mini import lookup3
                esize = 8 + 4 + 4
                fahd = self.alloc(b"\0" * (4 + 4 + 8 + 8 + 4))
                ents = [struct.pack("<QII", a, n, m) for a, n, m in records]
                per_page = 1 << page_bits
                if len(ents) <= per_page:
                    body = b"FADB" + struct.pack("<BBQ", 0, 1, fahd) + b"".join(ents) + b"\0" * 4
                else:
                    # paged: prefix = signature .. page bitmap (MSB first) + checksum, then the pages, each
                    # 2**page_bits entries (the last one the remainder) + checksum; an absent page keeps its room
                    npages = -(-len(ents) // per_page)
                    bitmap = bytearray((npages + 7) // 8)
                    pages = b""
                    for pg in range(npages):
                        part = b"".join(ents[pg * per_page:(pg + 1) * per_page])
                        if pg in absent_pages:
                            pages += b"\xAA" * (len(part) + 4)
                        else:
                            bitmap[pg // 8] |= 0x80 >> (pg % 8)
                            pages += part + struct.pack("<I", lookup3(part))
                            if pg < npages - 1:
                                assert len(part) == per_page * esize
                    prefix = b"FADB" + struct.pack("<BBQ", 0, 1, fahd) + bytes(bitmap)
                    body = prefix + struct.pack("<I", lookup3(prefix)) + pages
                fadb = self.alloc(body)
                self.buf[self.base + fahd: self.base + fahd + 28] = b"FAHD" + struct.pack("<BBBBQQ", 0, 1, esize, page_bits, len(records), fadb) + b"\0" * 4
                lay = head + struct.pack("<BBQ", 3, page_bits, fahd)
            msgs.append((0x08, lay))
        split = None
        if attrs:
            split = len(msgs)
            for k, v in attrs.items():
                msgs.append((0x0C, self.attr_msg(k, v)))
        return self.object_header(msgs, split_after=split)

    def group(self, children: dict, attrs=None, leaf_k: int = 4) -> int:
        msgs = [(0x02, struct.pack("<BBQQ", 0, 0, UNDEF, UNDEF)), (0x0A, struct.pack("<BB", 0, 0))]
        for i, nm in enumerate(sorted(children)):
            name = nm.encode()
            if i % 2:                                  # alternate the optional fields of a link message
                msgs.append((0x06, struct.pack("<BBBQBB", 1, 0x1C, 0, i, 0, len(name)) + name + struct.pack("<Q", children[nm])))
            else:
                msgs.append((0x06, struct.pack("<BBH", 1, 0x01, len(name)) + name + struct.pack("<Q", children[nm])))
        split = None
        if attrs:
            split = len(msgs)
            for k, v in attrs.items():
                msgs.append((0x0C, self.attr_msg(k, v)))
        return self.object_header(msgs, split_after=split)

    def finish(self, root_addr: int, path: str):
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBB", 2, 8, 8, 0)
        sb += struct.pack("<QQQQ", self.base, UNDEF, len(self.buf) - self.base, root_addr) + b"\0" * 4
        assert len(sb) == 48
        self.buf[self.base:self.base + 48] = sb
        with open(path, "wb") as fh:
            fh.write(self.buf)
