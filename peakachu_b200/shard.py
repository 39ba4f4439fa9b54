"""Multi-GPU sharding of ``score_genome``.

One process per GPU (torchrun: RANK / LOCAL_RANK / WORLD_SIZE). Work units are
chromosomes, plus band row tiles of chromosomes that are larger than an even
share; units are assigned greedily, largest first, to the least loaded rank.
Nothing is reduced on the device: every rank returns its records and rank 0
gathers them over the host backend (gloo) and formats the bedpe text. Row tiles of
one chromosome return their per-batch window counts so that the reference's "drop
a 100,000-batch with <= 1 window" rule (scoreUtils.py:104-108) is applied on the
gathered counts.
"""
from __future__ import annotations

import os

import numpy as np


def genome_cname(key):
    """Chromosome label of ``score_genome`` records (score_genome.py:48-51): the name itself when it
    starts with 'chr', else 'chr' + name. ``score_chromosome`` has a different rule
    (score_chromosome.py:37-38, ``'chr' + name.lstrip('chr')``), kept in score_chromosome.py."""
    return key if key.startswith("chr") else "chr" + key


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def band_pixels(n, lower, upper, w):
    lo, up = max(lower, w + 1), min(upper, n - 2 * w)
    if up < lo:
        return 0
    k = up - lo + 1
    return k * n - (lo + up) * k // 2


def plan(sizes: dict, world: int, lower: int, upper: int, w: int, split: bool = True):
    """Greedy largest-first assignment.

    sizes: chromosome -> n_bins. Returns a list (one per rank) of units
    ``(chrom, row_begin, row_end)``. A chromosome whose band is larger than
    total/world is cut into equal row tiles (at most ``world``)."""
    cost = {k: band_pixels(n, lower, upper, w) for k, n in sizes.items()}
    total = sum(cost.values())
    units = []
    for k, n in sizes.items():
        parts = 1
        if split and world > 1 and total > 0 and cost[k] > total / world:
            parts = min(world, int(np.ceil(cost[k] / (total / world))))
        edges = np.linspace(0, n, parts + 1).astype(int)
        for i in range(parts):
            units.append((cost[k] / parts, k, int(edges[i]), int(edges[i + 1])))
    units.sort(key=lambda u: (-u[0], u[1], u[2]))
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for c, k, a, b in units:
        r = int(np.argmin(load))
        load[r] += c
        out[r].append((k, a, b))
    return out


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def bind_to_device_node(device, sysfs="/sys"):
    """Run this process on the CPUs of the NUMA node its GPU is attached to, so that host buffers
    allocated (and pinned) from now on are local to the GPU's DMA engine. With one process per GPU on
    a two-socket box the ranks otherwise land on either socket and half of the uploads cross the
    socket interconnect. Returns the node, or None when the topology is not exposed (containers,
    single-node hosts) or PEAKACHU_B200_NUMA=0."""
    import ctypes as C

    from . import _lib
    if os.environ.get("PEAKACHU_B200_NUMA", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    buf = C.create_string_buffer(32)
    _lib.check(_lib.lib().pk_device_pci_bus_id(int(device), buf, 32))
    bus = buf.value.decode().lower()
    try:
        node = int(open(os.path.join(sysfs, "bus/pci/devices", bus, "numa_node")).read())
        if node < 0:
            return None
        cpus = _cpulist(open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)).read())
        cpus &= os.sched_getaffinity(0)          # stay inside a cgroup / taskset restriction
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError):
        return None


_GROUP = None


def _host_group():
    global _GROUP
    import torch.distributed as dist
    if _GROUP is None:
        if not dist.is_initialized():
            dist.init_process_group(backend="gloo")
            _GROUP = dist.group.WORLD
        elif dist.get_backend() == "gloo":
            _GROUP = dist.group.WORLD
        else:
            _GROUP = dist.new_group(backend="gloo")
    return _GROUP


# -- host-side gather ---------------------------------------------------------------
# Records are numpy columns. Pickling them through gloo's gather_object costs milliseconds per pass
# (serialise, TCP loopback, deserialise; 6 ms at two ranks), as much as a rank's share of an hg19-sized
# genome takes to score on eight GPUs, and even a metadata-only gather_object + barrier costs 2-3 ms at
# eight ranks. On one node (the deployment SURVEY.md 8(e) describes) nothing goes through a collective
# after the first call: every rank owns a small control segment and a data segment in POSIX shared
# memory. A rank copies its columns into its data segment and publishes the pass number in its control
# segment; rank 0 polls for it, copies the columns out and acknowledges in its own control segment,
# which the rank checks before it overwrites the data in the next pass.
#   control segment: [0:8] last published pass, [8:16] bytes of the data block, [16:80] name of the data
#   segment, [128 + 8 r] rank 0 only: last pass of rank r it has finished reading
_CTL_BYTES = 4096
_SHM = {"mode": None, "pass": 0, "ctl": None, "seg": None, "cap": 0, "ctl_peers": {}, "peers": {}, "peer_names": {},
        "fence": None, "held": None, "world": 1}


def _strip_arrays(obj, arrays):
    """Replace every ndarray in a nest of dicts / lists by a placeholder."""
    if isinstance(obj, np.ndarray):
        arrays.append(np.ascontiguousarray(obj))
        return ("__nd__", len(arrays) - 1)
    if isinstance(obj, dict):
        return {k: _strip_arrays(v, arrays) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_strip_arrays(v, arrays) for v in obj]
    return obj


def _restore_arrays(obj, arrays):
    if isinstance(obj, tuple) and len(obj) == 2 and obj[0] == "__nd__":
        return arrays[obj[1]]
    if isinstance(obj, dict):
        return {k: _restore_arrays(v, arrays) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_restore_arrays(v, arrays) for v in obj]
    return obj


def _release_shm():
    # A rank other than 0 must not unlink its data segment before rank 0 has read the last pass it
    # published: rank 0 attaches a peer's data segment lazily, at its first read, and usually finishes
    # last (the greedy plan gives it the largest unit). Wait for its acknowledgement first.
    release_gathered()                    # rank 0: views of the last pass are not needed any more
    root_name = _SHM.get("root_ctl")
    if _SHM["mode"] == "shm" and root_name is not None and _SHM["pass"] > 0 and _SHM.get("rank", 0) != 0:
        root = _SHM["ctl_peers"].get(root_name)
        if root is not None:
            try:
                _wait_for(root.buf, 128 + 8 * _SHM["rank"], _SHM["pass"],
                          "rank 0 to read pass %d before exit" % _SHM["pass"], timeout=120.0)
            except Exception:
                pass                      # rank 0 died or never gathered: nothing left to protect
    for d in (_SHM["peers"], _SHM["ctl_peers"]):
        for peer in d.values():
            try:
                peer.close()
            except Exception:
                pass
        d.clear()
    for key in ("seg", "ctl"):
        seg, _SHM[key] = _SHM[key], None
        if seg is not None:
            try:
                seg.close()
                seg.unlink()
            except Exception:
                pass


def _attach(name, cache):
    """Map another rank's segment (the owner unlinks it)."""
    from multiprocessing import resource_tracker, shared_memory
    seg = shared_memory.SharedMemory(name=name)
    try:        # Python < 3.13 registers attached segments too; the owner is responsible for this one
        resource_tracker.unregister(seg._name, "shared_memory")
    except Exception:
        pass
    cache[name] = seg
    return seg


def _fence():
    """Full memory barrier between writing a block and publishing its sequence number (and between
    seeing the number and reading the block): a mutex round trip."""
    lock = _SHM["fence"]
    lock.acquire()
    lock.release()


def _wait_for(buf, offset, value, what, timeout=600.0):
    import struct
    import time
    t0, spins = time.monotonic(), 0
    while struct.unpack_from("<Q", buf, offset)[0] < value:
        spins += 1
        time.sleep(0 if spins < 20000 else 1e-4)
        if (spins & 1023) == 0 and time.monotonic() - t0 > timeout:
            raise RuntimeError("host gather: timed out waiting for %s" % what)


def _setup_shm(rank, world, group):
    """First call: decide the mode, create the control segments, exchange their names once."""
    import atexit
    import socket
    import threading
    from multiprocessing import shared_memory

    import torch.distributed as dist
    want = os.environ.get("PEAKACHU_B200_GATHER", "shm")
    ctl = None
    if want == "shm":
        try:
            ctl = shared_memory.SharedMemory(create=True, size=_CTL_BYTES)
            ctl.buf[:_CTL_BYTES] = bytes(_CTL_BYTES)
        except OSError:
            ctl = None
    info = [None] * world
    dist.all_gather_object(info, (socket.gethostname(), ctl.name if ctl is not None else None), group=group)
    ok = want == "shm" and len({h for h, _ in info}) == 1 and all(n is not None for _, n in info)
    if not ok:
        if ctl is not None:
            ctl.close()
            ctl.unlink()
        _SHM["mode"] = "pickle"
        return
    _SHM.update(mode="shm", ctl=ctl, fence=threading.Lock(), rank=rank)
    atexit.register(_release_shm)
    if rank == 0:
        for r in range(1, world):
            _attach(info[r][1], _SHM["ctl_peers"])
        _SHM["ctl_names"] = [n for _, n in info]
    else:
        _attach(info[0][1], _SHM["ctl_peers"])
        _SHM["root_ctl"] = info[0][1]


def release_gathered():
    """Rank 0: done with the views a ``gather_to_rank0(..., copy=False)`` returned -- acknowledge the pass, so
    that the other ranks may overwrite their blocks. A no-op elsewhere, and when nothing is held."""
    import struct
    p = _SHM.get("held")
    if p is None or _SHM["ctl"] is None:
        _SHM["held"] = None
        return
    _fence()
    mine = _SHM["ctl"].buf
    for r in range(1, _SHM["world"]):
        struct.pack_into("<Q", mine, 128 + 8 * r, p)
    _SHM["held"] = None


def _gather_shm(obj, rank, world, copy=True):
    import pickle
    import struct
    from multiprocessing import shared_memory
    if rank == 0:
        release_gathered()                # views of the previous pass die here at the latest
    _SHM["world"] = world
    p = _SHM["pass"] = _SHM["pass"] + 1
    if rank != 0:
        arrays = []
        meta = _strip_arrays(obj, arrays)
        layout, off = [], 0
        for a in arrays:
            layout.append((a.dtype.str, a.shape, off))
            off += (a.nbytes + 63) & ~63
        head = pickle.dumps((meta, layout), protocol=pickle.HIGHEST_PROTOCOL)
        base = (16 + len(head) + 63) & ~63
        total = base + off
        # rank 0 must have finished with the previous pass before its block is overwritten
        root = _SHM["ctl_peers"][_SHM["root_ctl"]]
        _wait_for(root.buf, 128 + 8 * rank, p - 1, "rank 0 to read pass %d" % (p - 1))
        if _SHM["seg"] is None or _SHM["cap"] < total:
            if _SHM["seg"] is not None:
                _SHM["seg"].close()
                _SHM["seg"].unlink()
            _SHM["cap"] = max(1 << 20, int(total * 1.5))
            _SHM["seg"] = shared_memory.SharedMemory(create=True, size=_SHM["cap"])
        buf = _SHM["seg"].buf
        struct.pack_into("<QQ", buf, 0, len(head), base)
        buf[16:16 + len(head)] = head
        for a, (_, _, o) in zip(arrays, layout):
            if a.nbytes:
                np.frombuffer(buf, dtype=np.uint8, count=a.nbytes, offset=base + o)[:] = a.reshape(-1).view(np.uint8)
        ctl = _SHM["ctl"].buf
        name = _SHM["seg"].name.encode()
        struct.pack_into("<Q64s", ctl, 8, total, name)
        _fence()
        struct.pack_into("<Q", ctl, 0, p)                 # publish
        return None
    result = [obj]
    mine = _SHM["ctl"].buf
    for r in range(1, world):
        ctl = _SHM["ctl_peers"][_SHM["ctl_names"][r]].buf
        _wait_for(ctl, 0, p, "rank %d to publish pass %d" % (r, p))
        _fence()
        total, name = struct.unpack_from("<Q64s", ctl, 8)
        name = name.rstrip(b"\0").decode()
        seg = _SHM["peers"].get(r)
        if seg is None or _SHM["peer_names"].get(r) != name:      # first pass, or rank r grew its segment
            if seg is not None:
                try:
                    seg.close()
                except BufferError:               # views of a released pass are still referenced somewhere
                    _SHM.setdefault("stale", []).append(seg)
            seg = _attach(name, {})
            _SHM["peers"][r] = seg
            _SHM["peer_names"][r] = name
        hlen, base = struct.unpack_from("<QQ", seg.buf, 0)
        meta, layout = pickle.loads(bytes(seg.buf[16:16 + hlen]))
        got = [np.frombuffer(seg.buf, dtype=np.dtype(dt), count=int(np.prod(shape, dtype=np.int64)), offset=base + o)
               .reshape(shape) for dt, shape, o in layout]
        if copy:
            got = [a.copy() for a in got]
        result.append(_restore_arrays(meta, got))
        if copy:
            _fence()
            struct.pack_into("<Q", mine, 128 + 8 * r, p)      # acknowledge: rank r may overwrite its block
    if not copy:
        _SHM["held"] = p                  # acknowledged by release_gathered() (or the next gather, or at exit)
    return result


def gather_to_rank0(obj, rank, world, copy=True):
    """Host-side gather of the ranks' records to rank 0: numpy columns through shared memory when every
    rank runs on the same host (PEAKACHU_B200_GATHER=shm, the default), else pickled through gloo's
    gather_object. Returns the list of per-rank objects on rank 0, None elsewhere. With ``copy=False`` the
    arrays of the other ranks are views of their shared-memory blocks (no copy on rank 0: 8 MB per pass of the
    hg19-shaped genome, 1.1 ms of a 2.6 ms pass on 8 GPUs); they stay valid until ``release_gathered()``, which
    the caller owes the other ranks as soon as it has consumed them (the next gather releases them at the latest)."""
    if world == 1:
        return [obj]
    import torch.distributed as dist
    if _SHM["mode"] is None:
        _setup_shm(rank, world, _host_group())
    if _SHM["mode"] == "shm":
        return _gather_shm(obj, rank, world, copy=copy)
    out = [None] * world if rank == 0 else None
    dist.gather_object(obj, out, dst=0, group=_host_group())
    return out


class Engine:
    """The library's persistent scoring pipeline of one device (``pk_engine_*``): streams, chromosome
    handles and pinned staging live as long as the object; ``submit`` queues a unit without waiting for
    the device, ``collect`` returns the records of everything queued."""

    _cache: dict = {}

    def __init__(self, forest, lower, upper, device, depth=6):
        import ctypes as C

        from . import _lib
        self._L = _lib.lib()
        self.forest, self.device = forest, device
        self._h = C.c_void_p()
        _lib.check(self._L.pk_engine_create(device, forest.handle, forest.width, int(lower), int(upper), int(depth),
                                            C.byref(self._h)))
        self._keep = []

    @classmethod
    def of(cls, forest, lower, upper, device, depth=6):
        key = (id(forest), int(lower), int(upper), int(device), int(depth))
        hit = cls._cache.get(key)
        if hit is None or hit.forest is not forest or not hit._h:
            hit = cls._cache[key] = cls(forest, lower, upper, device, depth)
        return hit

    def submit(self, tag, n_bins, row_begin, row_end, encoding, a, b, c, size, weights, min_prob, poisson_weights=None):
        import ctypes as C

        from . import _lib
        arrays = [v for v in (a, b, c, weights, poisson_weights) if v is not None]
        self._keep.extend(arrays)                        # uploads are asynchronous when the source is pinned
        u = _lib.Unit(int(tag), int(n_bins), int(row_begin), int(row_end), int(encoding),
                      a.ctypes.data if a is not None else None, b.ctypes.data if b is not None else None,
                      c.ctypes.data if c is not None else None, int(size),
                      weights.ctypes.data if weights is not None else None,
                      poisson_weights.ctypes.data if poisson_weights is not None else None, float(min_prob))
        rc = self._L.pk_engine_submit(self._h, C.byref(u))
        if rc != 0:
            msg = self._L.pk_last_error().decode("utf-8", "replace")
            self.reset()
            raise _lib.PKError(rc, msg)

    def collect(self, copy=True):
        """List of result dicts in submission order. With ``copy=False`` the record columns are views of
        the engine's pinned block, valid until the next ``submit``."""
        import ctypes as C

        from . import _lib
        n = C.c_int64()
        _lib.check(self._L.pk_engine_collect(self._h, None, 0, C.byref(n)))
        res = (_lib.UnitResult * max(n.value, 1))()
        rc = self._L.pk_engine_collect(self._h, res, n.value, C.byref(n))
        self._keep = []
        if rc != 0:
            msg = self._L.pk_last_error().decode("utf-8", "replace")
            self.reset()
            raise _lib.PKError(rc, msg)

        def col(ptr, count, dtype):
            if count == 0:
                return np.zeros(0, dtype)
            buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
            v = np.frombuffer(buf, dtype=dtype, count=count)
            return v.copy() if copy else v
        out = []
        for r in res[:n.value]:
            m = r.n_records
            out.append(dict(tag=r.tag, row_begin=r.row_begin, whole=bool(r.whole), x=col(r.x, m, np.int32),
                            y=col(r.y, m, np.int32), p=col(r.prob, m, np.float64), v=col(r.value, m, np.float64),
                            batch=col(r.batch, m, np.int32),
                            batch_windows=col(r.batch_windows, r.n_batches, np.int32).astype(np.int64),
                            n_candidates=int(r.n_candidates), n_windows=int(r.n_windows)))
        return out

    def reset(self):
        self._keep = []
        self._L.pk_engine_reset(self._h)

    def close(self):
        if self._h:
            self._L.pk_engine_destroy(self._h)
            self._h = None


def map_weights(Lib, key, correct):
    """(weights that balance the pixel values, weights of the Poisson filter or None when they are the same)
    of one chromosome. cooler's ``matrix(balance=name)`` multiplies by the column -- or by its reciprocal
    when the column carries ``divisive_weights`` (hic2cool's KR / VC columns) -- while the reference hands
    the column's raw values to ``Chromosome`` (score_chromosome.py:42-44), which divides the expected value
    by them in the Poisson filter (scoreUtils.py:55-57)."""
    from . import _lib
    if not correct:
        return None, None
    w = _lib.as_c(Lib.weights(key, correct), np.float64)
    if getattr(Lib, "weights_divisive", None) is not None and Lib.weights_divisive(correct):
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.ascontiguousarray(1.0 / w), w
    return w, None


def _slim_far():
    """PEAKACHU_B200_SLIM_FAR=0 makes a cooler file's packed rows carry every pixel beyond the band, not only
    the ones the `valid` mask needs (see H5Cool.upper_pixels_rows)."""
    import os
    return os.environ.get("PEAKACHU_B200_SLIM_FAR", "1") != "0"


def _unit_columns(Lib, key, nd_need, encoding, scoring_weights=False):
    """The pixel columns of one chromosome in the most compact form the reader offers (or the one asked
    for): (PK_ENC_*, a, b, c, size). ``scoring_weights`` (the balancing weights, None for raw counts): the caller
    scores the chromosome, so a reader that packs rows on the fly may leave out the pixels beyond the band that the
    `valid` mask does not need."""
    from . import _lib
    n = Lib.nbins(key)
    if encoding in (None, "rows") and hasattr(Lib, "upper_pixels_rows"):
        if scoring_weights is not False and _slim_far() and getattr(Lib, "packs_rows_on_the_fly", False):
            blob = Lib.upper_pixels_rows(key, nd_need, scoring_weights=scoring_weights)
        else:
            blob = Lib.upper_pixels_rows(key, nd_need)
        if blob is not None:
            blob = _lib.as_c(blob, np.uint8)
            return _lib.PK_ENC_ROWS, blob, None, None, blob.size
    if encoding == "rows":
        raise ValueError("no packed rows covering %d distances for chromosome %s" % (nd_need, key))
    if encoding in (None, "csr16") and hasattr(Lib, "upper_pixels_csr16"):
        narrow = Lib.upper_pixels_csr16(key)
        if narrow is not None:
            rp, d16, c16 = narrow
            if np.asarray(d16).dtype != np.uint16 or np.asarray(c16).dtype != np.uint16:
                raise TypeError("upper_pixels_csr16 must return uint16 columns")
            rp, d16, c16 = _lib.as_c(rp, np.int64), _lib.as_c(d16, np.uint16), _lib.as_c(c16, np.uint16)
            return _lib.PK_ENC_CSR16, rp, d16, c16, d16.size
    if encoding == "csr16":
        raise ValueError("chromosome %s has pixels that uint16 columns cannot hold" % key)
    if encoding in (None, "csr32") and hasattr(Lib, "upper_pixels_csr"):
        rp, b2, cnt = Lib.upper_pixels_csr(key)
        rp, b2, cnt = _lib.as_c(rp, np.int64), _lib.as_c(b2, np.int32), _lib.as_c(cnt, np.int32)
        if rp.size != n + 1:
            raise ValueError("bin1_offset must have n_bins + 1 entries")
        return _lib.PK_ENC_CSR32, rp, b2, cnt, b2.size
    b1, b2, cnt = Lib.upper_pixels(key)
    b1, b2, cnt = (_lib.as_c(v, np.int32) for v in (b1, b2, cnt))
    return _lib.PK_ENC_COO, b1, b2, cnt, b1.size


def score_units(Lib, units, flat, *, correct, lower, upper, res, device, min_prob, depth=6, copy=True,
                encoding=None):
    """Score this rank's units ``(chrom, row_begin, row_end)``. Returns {chrom: [tile dict, ...]}.

    Everything runs in the library's engine (``pk_engine_*``): the units are queued back to back --
    while one chromosome's kernels run, the next one's pixel columns are already crossing the bus --
    and collected with one wait at the end. ``encoding`` forces a column format (``"rows"``,
    ``"csr16"``, ``"csr32"``, ``"coo"``); by default the most compact one the reader offers is used."""
    from . import _lib
    from .scoreUtils import DeviceForest
    _lib.require_device()
    forest = DeviceForest.of(flat, device)
    if _lib.expected_mode() == "host":
        return _score_units_host_fit(Lib, units, forest, correct, lower, upper, res, device, min_prob, encoding)
    eng = Engine.of(forest, lower, upper, device, depth)
    w = forest.width
    keys = []
    try:
        for key, a, b in units:
            n = Lib.nbins(key)
            nd_need = min(upper, n - 2 * w) + 2 * w + 1              # stored distances (scoreUtils.py:14,31)
            weights, pweights = map_weights(Lib, key, correct)
            if weights is not None and weights.size != n:
                raise ValueError("weight column of %s has %d entries for %d bins" % (key, weights.size, n))
            enc, ca, cb, cc, size = _unit_columns(Lib, key, nd_need, encoding, scoring_weights=weights)
            eng.submit(len(keys), n, a, b, enc, ca, cb, cc, size, weights, min_prob, pweights)
            keys.append(key)
        results = eng.collect(copy=copy)
    except BaseException:
        eng.reset()
        raise
    out = {}
    for key, r in zip(keys, results):
        out.setdefault(key, []).append(r)
    return out


def _score_units_host_fit(Lib, units, forest, correct, lower, upper, res, device, min_prob, encoding):
    """PEAKACHU_B200_EXPECTED=host: the engine fits the expected curve on the device, so in this mode the units
    go one at a time through ``scoreUtils.Chromosome`` (which then fits with the installed scikit-learn)."""
    from .scoreUtils import Chromosome
    out = {}
    for key, a, b in units:
        n = Lib.nbins(key)
        weights, pweights = map_weights(Lib, key, correct)
        X = Chromosome.from_map(Lib, key, weights, forest, lower=lower, upper=upper, cname=key, res=res,
                                width=forest.width, device=device, encoding=encoding, first_tile=(a, b))
        if pweights is not None:
            X.set_poisson_weights(pweights)
        x, y, p, v, batch, bw = X.score_records(min_prob, with_batches=True)
        whole = a == 0 and b == n
        out.setdefault(key, []).append(dict(tag=len(out), row_begin=a, whole=whole, x=x, y=y, p=p, v=v, batch=batch,
                                            batch_windows=bw, n_candidates=int(X.n_candidates), n_windows=int(X.n_windows)))
        X.close()
    return out


def merge_tiles(parts):
    """Combine the row tiles of one chromosome: apply the batch rule on summed window
    counts (a tile that covered the whole chromosome already applied it on the
    device), then order by (x, y)."""
    if len(parts) == 1 and parts[0]["whole"]:
        q = parts[0]
        return q["x"], q["y"], q["p"], q["v"]
    nb = max((q["batch_windows"].size for q in parts), default=0)
    tot = np.zeros(nb, np.int64)
    for q in parts:
        tot[:q["batch_windows"].size] += q["batch_windows"]
    x = np.concatenate([q["x"] for q in parts])
    y = np.concatenate([q["y"] for q in parts])
    p = np.concatenate([q["p"] for q in parts])
    v = np.concatenate([q["v"] for q in parts])
    bt = np.concatenate([q["batch"] for q in parts])
    ok = tot[bt] > 1 if bt.size else np.zeros(0, bool)
    x, y, p, v = x[ok], y[ok], p[ok], v[ok]
    order = np.lexsort((y, x))
    return x[order], y[order], p[order], v[order]


def assemble_text(queue, gathered, res, verbose=False):
    """bedpe text per chromosome from the gathered per-rank tile dicts."""
    from .scoreUtils import format_bedpe
    text = {}
    for key in queue:
        parts = []
        for g in gathered:
            parts.extend(g.get(key, []))
        parts.sort(key=lambda q: q["row_begin"])
        cname = genome_cname(key)
        if verbose:                                           # scoreUtils.py:97-98
            print("scoring matrix {}".format(cname))
            print("number of candidates {}".format(sum(q["n_candidates"] for q in parts)))
        x, y, p, v = merge_tiles(parts)
        nz = p != 0                                            # prob_csr.nonzero() drops exact zeros
        text[key] = format_bedpe(cname, res, x[nz], y[nz], p[nz], v[nz])
    return text


def score_chromosomes(Lib, queue, flat, *, correct, lower, upper, res, min_prob, device=None, verbose=False):
    """Score ``queue`` across all ranks; on rank 0 returns {chrom: bedpe text}."""
    rank, world = rank_world()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        bind_to_device_node(device)
    sizes = {k: Lib.nbins(k) for k in queue}
    assignment = plan(sizes, world, lower, upper, flat.width)
    # smallest unit first: the pass starts computing after a short upload, and the large uploads that
    # follow hide behind kernels
    order = sorted(assignment[rank], key=lambda u: band_pixels(sizes[u[0]], lower, upper, flat.width) * (u[2] - u[1]) / max(sizes[u[0]], 1))
    mine = score_units(Lib, order, flat, correct=correct, lower=lower, upper=upper, res=res,
                       device=device, min_prob=min_prob)
    gathered = gather_to_rank0(mine, rank, world, copy=False)
    if rank != 0:
        return {}
    text = assemble_text(queue, gathered, res, verbose=verbose)
    release_gathered()                    # the text holds everything: the other ranks may go on (or exit)
    return text
