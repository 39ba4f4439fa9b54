#!/usr/bin/env python
"""Benchmark of the loop-scoring hot path (BASELINE.json metric: candidate pixels
scored / second, window features + RF proba).

  python bench.py --gpus N --steps K --warmup W            (N>1: under torchrun)
  python bench.py --impl reference --steps K --warmup W    CPU arm (oracle port)

A step = one pass of the whole path over one chromosome of the workload
(BASELINE configs[1]: chr1-scale synthetic, 24,900 bins at 10 kb, w=5, l=6, u=300,
100-tree forest): pixel scatter -> band, per-diagonal sums, expected fit, Poisson
candidate scan, window features, forest, threshold emit. "Pixels" are band pixels
sum_{d=lower..upper}(n-d), the count BASELINE.json quotes (~7.3 M for this map).

`value`  : device-resident inputs (pixel columns + weights already in HBM), timed
           with CUDA events on the library's stream, L2 flushed between steps.
`e2e`    : the public API call (Chromosome.from_pixels + score_records) on HOST
           buffers: H2D of pixels/weights and D2H of the records inside the timing.
With N ranks every rank scores its own chromosome of the same shape (different
seed): weak scaling, no collective on the data path; time = max over ranks.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_bins, res, lower, upper, w, forest, depth, band)
    "c2": dict(n=24900, res=10000, lower=6, upper=300, w=5, forest="c2", depth=300.0, band=330,
               desc="score_chromosome chr1-scale synthetic (24,900 bins, 10 kb, w=5, l=6, u=300, 100-tree RF)"),
    "c4": dict(n=49850, res=5000, lower=6, upper=600, w=7, forest="c4", depth=300.0, band=640,
               desc="score_chromosome chr1-scale synthetic at 5 kb (49,850 bins, w=7 / 15x15 windows, l=6, u=600, 200-tree RF)"),
    "c1": dict(n=2000, res=10000, lower=6, upper=300, w=5, forest="c2", depth=300.0, band=330,
               desc="score_chromosome 2,000-bin synthetic 10 kb (w=5, l=6, u=300, 100-tree RF)"),
}
# score_genome workloads (the `genome` block of the JSON line): hg19-shaped genomes sharded over the ranks
# (chromosomes + band row tiles, greedy), records gathered on rank 0 -- strong scaling
GENOMES = {
    # BASELINE configs[2]
    "c3": dict(res=10000, lower=6, upper=300, w=5, forest="c2", depth=300.0, band=330,
               desc="score_genome hg19-shaped synthetic 10 kb (23 chromosomes, 303,641 bins, w=5, l=6, u=300, 100-tree RF)"),
    # BASELINE configs[3]
    "c4": dict(res=5000, lower=6, upper=600, w=7, forest="c4", depth=300.0, band=640,
               desc="score_genome hg19-shaped synthetic 5 kb (23 chromosomes, 607,271 bins, w=7 / 15x15 windows, l=6, u=600, 200-tree RF)"),
}
KERNELS_PER_STEP = 12  # band_csr, valid_bits, diag_sums, fit_expected, cand_mark, scan2, cand_write, score_fused, emit, row_offsets, record_place, record_pack


def band_pixels(n, lower, upper, w):
    lo, up = max(lower, w + 1), min(upper, n - 2 * w)
    k = up - lo + 1
    return k * n - (lo + up) * k // 2 if k > 0 else 0


def make_map(wl, seed):
    from peakachu_b200 import synth
    return synth.make_chromosome("chr1", wl["n"], seed=seed, depth=wl["depth"], band=wl["band"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([s.strip() for s in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def workload_config(wl, args):
    """The `config` object of the JSON line: identical in the GPU arm and the --impl reference arm."""
    px = band_pixels(wl["n"], wl["lower"], wl["upper"], wl["w"]) if "n" in wl else None
    return {"workload": wl["desc"],
            "pixels": "band pixels sum_{d=l..u}(n-d)" + (" = %d per chromosome" % px if px else ""),
            "per_rank": "one chromosome of the workload per rank per step (weak scaling); the `genome` block of the "
                        "GPU arm is score_genome on the hg19-shaped map sharded over the ranks (strong scaling)",
            "forest": "bench_data/%s.pkl" % wl["forest"],
            "l2": "value: several chromosomes in flight, each ~150 MB of working set, together > 126 MB L2; stage_ms / roofline: "
                  "one chromosome at a time with a 256 MiB L2 flush between steps; e2e: every step uploads its "
                  "inputs from pinned host memory (the host is the cold side)"}


# ---------------------------------------------------------------------------
# CPU arm: the oracle port on host cores (the only place bench.py runs oracle/)
# ---------------------------------------------------------------------------
def cpu_pass(wl, n_bins, seed, model):
    """One oracle pass (score_chromosome body) on an n_bins chromosome of the same
    synthetic distribution. Returns (seconds, band pixels, candidates, records)."""
    import tempfile
    from oracle import peakachu_oracle as po
    from peakachu_b200 import coolio, synth
    ch = synth.make_chromosome("chr1", n_bins, seed=seed, depth=wl["depth"], band=wl["band"])
    path = os.path.join(tempfile.mkdtemp(), "cpu.pkcool")
    coolio.PKCool.write(path, [ch], wl["res"], rows_nd=0)
    lib = coolio.Cooler(path)
    t0 = time.perf_counter()
    st = po.score_map(lib, model, ["chr1"], weight_name="weight", lower=wl["lower"], upper=wl["upper"],
                      res=wl["res"], min_prob=0.5, output=os.path.join(os.path.dirname(path), "o.bedpe"))
    dt = time.perf_counter() - t0
    return dt, band_pixels(n_bins, wl["lower"], wl["upper"], wl["w"]), st[0]["candidates"], st[0]["rows"]


def load_sklearn_model(name):
    import joblib
    return joblib.load(os.path.join(ROOT, "bench_data", name + ".pkl"))


_WORKER_MODEL = {}


def _cpu_worker(job):
    wl, n_bins, seed = job
    if wl["forest"] not in _WORKER_MODEL:
        _WORKER_MODEL[wl["forest"]] = load_sklearn_model(wl["forest"])
    return cpu_pass(wl, n_bins, seed, _WORKER_MODEL[wl["forest"]])


def run_reference_arm(args, wl):
    """The reference path on the host cores, SAME configuration as the GPU arm: every step scores
    full-size chromosomes of the workload (24,900 bins for c2). The reference is single-threaded
    (forest n_jobs=1, sequential chromosome loop), so "all host threads" means one process per
    chromosome, the only parallelism its design admits."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    workers = max(1, min(os.cpu_count() or 1, args.cpu_workers))
    n_sample = args.cpu_bins or wl["n"]
    with ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("fork")) as ex:
        for _ in range(max(args.warmup, 1)):
            list(ex.map(_cpu_worker, [(wl, 600, 5)] * workers))          # imports, model load
        tot_t, tot_px = 0.0, 0
        for s in range(args.steps):
            t0 = time.perf_counter()
            res = list(ex.map(_cpu_worker, [(wl, n_sample, 1234 + s * workers + i) for i in range(workers)]))
            tot_t += time.perf_counter() - t0
            tot_px += sum(r[1] for r in res)
    val = tot_px / tot_t
    sample = ("per step: %d chromosomes of %d bins from the workload's synthetic distribution (%d band px each; the "
              "GPU arm's chromosome is seed 1234), one process each on %d of %d host threads, numpy oracle port of "
              "score_chromosome (oracle/peakachu_oracle.py, pinned to the reference's own outputs at this size: "
              "tests/golden/c2.json)" % (workers, n_sample, band_pixels(n_sample, wl["lower"], wl["upper"], wl["w"]),
                                         workers, os.cpu_count() or 1))
    print(json.dumps({
        "impl": "reference", "metric": METRIC,
        "value": val, "unit": "pixels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(wl, args),
        "same_config": n_sample == wl.get("n"),
        "cpu_baseline": {"value": val, "unit": "pixels/s", "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------
# pinned host maps (what a reader hands to the scoring API, minus the file)
# ---------------------------------------------------------------------------
def _pinned(a):
    import torch
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint16:                       # torch has no pinned uint16 on every build: pin the bytes
        return _pinned(a.view(np.uint8)).view(np.uint16)
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True)
    v = t.numpy()
    v[...] = a
    _PIN_KEEP.append(t)
    return v


_PIN_KEEP = []


class PinnedMap:
    """Chromosomes in pinned host memory in every column format of the C ABI. `alias` maps the
    names the bench scores (chr1, chr2, ... of identical content) to the stored chromosome."""

    def __init__(self, nd_enc):
        self.nd_enc, self.ch, self.alias = nd_enc, {}, {}

    def add(self, key, ch, formats=("rows", "csr16", "csr32")):
        from peakachu_b200 import rowpack
        rp = np.searchsorted(ch.bin1, np.arange(ch.n + 1)).astype(np.int64)
        e = dict(n=ch.n, w=_pinned(ch.weights), rp=_pinned(rp), nnz=int(ch.bin1.size))
        if "csr32" in formats:
            e["b2"], e["cnt"] = _pinned(ch.bin2), _pinned(ch.count)
        delta = ch.bin2 - ch.bin1
        if "csr16" in formats and ch.count.size and int(delta.max()) <= 65535 and int(ch.count.max()) <= 65535:
            e["d16"], e["c16"] = _pinned(delta.astype(np.uint16)), _pinned(ch.count.astype(np.uint16))
        if "rows" in formats:
            e["rows"] = _pinned(rowpack.pack_rows(rp, ch.bin2, ch.count, ch.n, self.nd_enc))
        self.ch[key] = e

    def _e(self, key): return self.ch[self.alias.get(key, key)]
    def nbins(self, key): return self._e(key)["n"]
    def weights(self, key, name): return self._e(key)["w"]
    def upper_pixels_csr(self, key): e = self._e(key); return e["rp"], e["b2"], e["cnt"]
    def upper_pixels_csr16(self, key): e = self._e(key); return (e["rp"], e["d16"], e["c16"]) if "d16" in e else None
    def upper_pixels_rows(self, key, nd_min): e = self._e(key); return e.get("rows") if self.nd_enc >= nd_min else None

    def h2d_bytes(self, key, encoding):
        e = self._e(key)
        if encoding == "rows":
            return int(e["rows"].nbytes) + 8 * e["n"]
        return (4 if encoding == "csr16" else 8) * e["nnz"] + 8 * (e["n"] + 1) + 8 * e["n"]


COLUMNS = {"rows": "packed pixel rows (pk_chrom_upload_rows: presence bitmap + 1 byte per count, one blob per chromosome; "
                   "what a .pkcool container stores) + weights f64",
           "csr16": "bin1_offset int64 + (bin2 - bin1) uint16 + count uint16 + weights f64 (pk_chrom_upload_csr16)",
           "csr32": "cooler's own columns: bin1_offset int64 + bin2 int32 + count int32 + weights f64 (pk_chrom_upload_csr)"}


# ---------------------------------------------------------------------------
# score_genome (BASELINE configs[2], [3]): the whole multi-rank path, strong scaling
# ---------------------------------------------------------------------------
def run_genome(args, name, flat, rank, world, local, steps, encodings):
    import torch
    import torch.distributed as dist
    from peakachu_b200 import shard, synth
    wl = GENOMES[name]
    sizes = synth.hg19_bins(wl["res"])
    queue = list(sizes)
    plan = shard.plan(sizes, world, wl["lower"], wl["upper"], wl["w"])
    # smallest unit first, as shard.score_chromosomes orders them
    mine = sorted(plan[rank], key=lambda u: band_pixels(sizes[u[0]], wl["lower"], wl["upper"], wl["w"]) * (u[2] - u[1]) / sizes[u[0]])
    need = sorted({k for k, _, _ in mine}, key=queue.index)
    pm = PinnedMap(nd_enc=(wl["upper"] + 2 * wl["w"] + 1 + 31) // 32 * 32)
    t_gen = time.perf_counter()
    for k in need:
        pm.add(k, synth.make_chromosome(k, sizes[k], seed=5000 + queue.index(k), depth=wl["depth"], band=wl["band"]),
               formats=encodings)
    t_gen = time.perf_counter() - t_gen
    px = sum(band_pixels(n, wl["lower"], wl["upper"], wl["w"]) for n in sizes.values())
    my_px = sum(band_pixels(sizes[k], wl["lower"], wl["upper"], wl["w"]) * (b - a) / sizes[k] for k, a, b in mine)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    out = {}
    for enc in encodings:
        phase = np.zeros(3)          # score_units | gather | merge (this rank, all timed passes)

        def one_pass():
            t_a = time.perf_counter()
            res = shard.score_units(pm, mine, flat, correct="weight", lower=wl["lower"], upper=wl["upper"],
                                    res=wl["res"], device=local, min_prob=0.5, copy=False, encoding=enc)
            t_b = time.perf_counter()
            gathered = shard.gather_to_rank0(res, rank, world, copy=False)
            t_c = time.perf_counter()
            n_out = 0
            if rank == 0:
                merged = {k: shard.merge_tiles(sorted([q for g in gathered for q in g.get(k, [])],
                                                      key=lambda q: q["row_begin"])) for k in queue}
                n_out = sum(int(m[0].size) for m in merged.values())
                chk = sum(int(m[0][-1]) for m in merged.values() if m[0].size)      # touch the gathered columns
                shard.release_gathered()                                            # ... before handing the blocks back
                n_out += 0 * chk
            t_d = time.perf_counter()
            phase[:] += (t_b - t_a, t_c - t_b, t_d - t_c)
            return n_out

        for _ in range(3):
            one_pass()
        barrier()
        phase[:] = 0
        t0 = time.perf_counter()
        nrec = 0
        for _ in range(steps):
            nrec = one_pass()
        barrier()
        dt = time.perf_counter() - t0
        h2d = sum(pm.h2d_bytes(k, enc) for k, _, _ in mine)
        t = torch.tensor([dt] + (phase / steps).tolist() + [h2d], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt = float(tmax[0].item())
        out[enc] = {"ms_per_pass": 1e3 * dt / steps, "pixels_per_s": px * steps / dt,
                    "pixels_per_s_per_gpu": px * steps / dt / world, "records_per_pass": nrec,
                    "phase_ms_max_over_ranks": {"score_units": 1e3 * float(tmax[1]), "gather": 1e3 * float(tmax[2]),
                                                "merge_rank0": 1e3 * float(tmax[3])},
                    "h2d_bytes_per_pass_all_ranks": int(float(t[4]) + 0.5),
                    "h2d_bytes_per_pass_max_rank": int(float(tmax[4]) + 0.5),
                    "pcie_gbs_busiest_rank": float(tmax[4]) / (dt / steps) / 1e9, "columns": COLUMNS[enc]}
    return {"workload": wl["desc"], "band_pixels": px, "scaling": "strong", "steps": steps,
            "timed": "whole passes back to back: pinned host columns -> H2D -> kernels -> records D2H -> host gather "
                     "to rank 0 -> records of every chromosome in the reference's order (bedpe text not formatted)",
            "units_per_rank": [len(u) for u in plan],
            "share_of_busiest_rank": max(sum(band_pixels(sizes[k], wl["lower"], wl["upper"], wl["w"]) * (b - a) / sizes[k]
                                             for k, a, b in u) for u in plan) / px,
            "synth_s_this_rank": t_gen, "by_columns": out}


# ---------------------------------------------------------------------------
# file -> bedpe: the CLI's own call (score_chromosome.main) on a map file of the workload's chromosome
# ---------------------------------------------------------------------------
def run_file_e2e(wl, ch, local, reps=3):
    """Wall time of `peakachu_b200 score_chromosome -p FILE -C chr1 -m MODEL -O out.bedpe` for the workload's
    chromosome stored (a) as an HDF5 cooler file, chunked + gzip + shuffle like cooler writes them, read by the
    built-in h5mini reader, and (b) as the repo's .pkcool container (packed pixel rows). The model is the
    workload's .pkl (joblib); the bedpe is written to local disk. Best of `reps` runs after one warm-up."""
    import argparse
    import contextlib
    import io
    import tempfile

    from peakachu_b200 import coolio, score_chromosome
    from tests import h5write                      # the HDF5 writer of the test fixtures (bench input only)
    tmp = tempfile.mkdtemp(prefix="pk_file_e2e_")
    px = band_pixels(ch.n, wl["lower"], wl["upper"], wl["w"])
    paths = {"cool": os.path.join(tmp, "wl.cool"), "pkcool": os.path.join(tmp, "wl.pkcool")}
    t0 = time.perf_counter()
    h5write.write_cool(paths["cool"], [ch], wl["res"])
    t_write = time.perf_counter() - t0
    coolio.PKCool.write(paths["pkcool"], [ch], wl["res"])
    out = {}
    for kind, path in paths.items():
        ns = argparse.Namespace(path=path, model=os.path.join(ROOT, "bench_data", wl["forest"] + ".pkl"), output=os.path.join(tmp, kind + ".bedpe"),
                                resolution=wl["res"], lower=wl["lower"], upper=wl["upper"], minimum_prob=0.5,
                                clr_weight_name="weight", chrom=ch.name, device=local)
        best, read_best = None, None
        for r in range(reps + 1):
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                score_chromosome.main(ns)
            dt = time.perf_counter() - t0
            # the reader's share: open the file and pull the chromosome's columns the way from_map does
            t1 = time.perf_counter()
            Lib = coolio.open_map(path)
            n = Lib.nbins(ch.name)
            from peakachu_b200 import shard
            shard._unit_columns(Lib, ch.name, min(wl["upper"], n - 2 * wl["w"]) + 2 * wl["w"] + 1, None)
            shard.map_weights(Lib, ch.name, "weight")
            dr = time.perf_counter() - t1
            if r > 0:
                best = dt if best is None else min(best, dt)
                read_best = dr if read_best is None else min(read_best, dr)
        out[kind] = {"ms": 1e3 * best, "pixels_per_s": px / best, "read_ms": 1e3 * read_best,
                     "file_bytes": os.path.getsize(path), "records": sum(1 for _ in open(ns.output))}
    out["note"] = ("score_chromosome.main(args) from the file path to the bedpe on disk: model load (joblib .pkl -> node tables), "
                   "file read (read_ms: open + the chromosome's pixel columns + weights, measured separately), upload, kernels, "
                   "record fetch, bedpe text. The .cool is written by tests/h5write.py (%.1f s, not timed) with cooler's layout: "
                   "chunked, gzip-6 + shuffle." % t_write)
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return out


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
METRIC = "candidate pixels scored/sec (window features + RF proba)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c4"])
    ap.add_argument("--cpu-bins", type=int, default=0, help="chromosome size of the CPU arms (0: the workload's own, 24,900 for c2)")
    ap.add_argument("--cpu-workers", type=int, default=32, help="processes of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--genome", default="auto", help="score_genome blocks in the JSON line: auto (c3 at every N, c4 too at N = 8), "
                                                     "none, or a comma list of c3,c4")
    ap.add_argument("--genome-steps", type=int, default=0, help="timed passes of a genome block (0: min(steps, 10))")
    ap.add_argument("--reserve-sms", type=int, default=8, help="pk_set_tuning('reserve_sms'): SMs the fused kernel leaves to the short stages of the chromosomes queued behind (engine passes)")
    ap.add_argument("--inflight", type=int, default=4, help="chromosomes in flight (streams) of the device-resident `value` pass")
    ap.add_argument("--no-file-e2e", action="store_true", help="skip the file -> bedpe block (N = 1, about 40 s of host work)")
    ap.add_argument("--numa", type=int, default=1, help="N > 1: run each rank on the NUMA node of its GPU (0: leave the affinity alone)")
    ap.add_argument("--fused", type=int, default=-1, help="pk_set_tuning('fused'): -1 auto, 0 off, 1, 2")
    ap.add_argument("--prune", type=int, default=1, help="pk_set_tuning('prune'): retire pixels that cannot exceed min_prob")
    ap.add_argument("--child-features", type=int, default=-1, help="pk_set_tuning('child_features'): forest walk on the child-feature node encoding (-1 auto, 0 off, 1 on)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    from peakachu_b200 import _lib, shard
    from peakachu_b200.forest import FlatForest
    from peakachu_b200.scoreUtils import DeviceForest

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    L = _lib.lib()
    _lib.require_device()
    numa_node = None
    if world > 1 and args.numa:
        numa_node = shard.bind_to_device_node(local)       # before any pinned allocation
    print("rank %d: device %d, NUMA node %s" % (rank, local, numa_node), file=sys.stderr)
    _lib.check(L.pk_set_tuning(b"fused", args.fused))
    _lib.check(L.pk_set_tuning(b"prune", args.prune))
    _lib.check(L.pk_set_tuning(b"child_features", args.child_features))
    _lib.check(L.pk_set_tuning(b"reserve_sms", args.reserve_sms))
    args.warmup = max(args.warmup, 3)

    flat = FlatForest.load(os.path.join(ROOT, "bench_data", wl["forest"] + "_forest.npz"))
    forest = DeviceForest.of(flat, local)
    ch = make_map(wl, seed=1234 + rank)
    n, w = ch.n, wl["w"]
    px = band_pixels(n, wl["lower"], wl["upper"], w)
    nnz = ch.bin1.size

    # ---- device-resident inputs (torch tensors only as buffers): cooler's CSR columns ----
    stream = torch.cuda.Stream(device=local)
    rowptr = np.searchsorted(ch.bin1, np.arange(n + 1)).astype(np.int64)      # indexes/bin1_offset
    d_rp = torch.from_numpy(rowptr).cuda(); d_b2 = torch.from_numpy(ch.bin2).cuda()
    d_cnt = torch.from_numpy(ch.count).cuda(); d_w = torch.from_numpy(ch.weights).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    h = C.c_void_p()
    _lib.check(L.pk_chrom_create(local, n, w, wl["lower"], wl["upper"], 1, C.c_void_p(stream.cuda_stream), C.byref(h)))

    def device_step(hh):
        _lib.check(L.pk_chrom_upload_csr(hh, C.c_void_p(d_rp.data_ptr()), C.c_void_p(d_b2.data_ptr()),
                                         C.c_void_p(d_cnt.data_ptr()), nnz, C.c_void_p(d_w.data_ptr()),
                                         _lib.PK_MEM_DEVICE))
        _lib.check(L.pk_chrom_fit_expected(hh))
        _lib.check(L.pk_chrom_find_candidates(hh, 0, n, None))
        _lib.check(L.pk_chrom_score(hh, forest.handle, 0.5))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)       # nvidia-smi clocks / throttle reasons over all timed regions
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        device_step(h)
    barrier()
    # ---- pass 1: one chromosome at a time, L2 flushed between steps: per-kernel times ----
    k1 = min(args.steps, 10)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k1)]
    stage_acc = np.zeros(8)
    for a, b in ev:
        flush.fill_(1)                                  # evict L2 (256 MiB > 126 MB), untimed
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            a.record(stream)
            device_step(h)
            b.record(stream)
        stream.synchronize()
        ms = np.zeros(8, dtype=np.float32)
        _lib.check(L.pk_chrom_stage_ms(h, _lib.ptr(ms, _lib.c_f32p)))
        stage_acc += ms
    serial_ms = sum(a.elapsed_time(b) for a, b in ev) / k1
    nrec, ncand, nwin = C.c_int64(), C.c_int64(), C.c_int64()
    _lib.check(L.pk_chrom_result_count(h, C.byref(nrec), C.byref(ncand), C.byref(nwin)))

    # ---- pass 2 (the reported value): K chromosomes, NFLIGHT in flight, arranged as score_genome's engine runs
    # them (pk_engine_*): every handle has a high-priority stream for its upload and short stages, the fused
    # scoring kernels go back to back to one ordinary stream (pk_chrom_set_score_stream) and leave a few SMs to
    # the short stages of the chromosomes queued behind (tuning "reserve_sms"). The working sets in flight
    # (NFLIGHT x ~150 MB) exceed L2, so no flush is needed.
    NFLIGHT = args.inflight
    streams = []
    for _ in range(NFLIGHT):
        sp = C.c_void_p()
        _lib.check(L.pk_stream_create_priority(local, 1, C.byref(sp)))
        streams.append(torch.cuda.ExternalStream(sp.value, device=local))
    score_stream = torch.cuda.Stream(device=local)
    handles = []
    for st in streams:
        hh = C.c_void_p()
        _lib.check(L.pk_chrom_create(local, n, w, wl["lower"], wl["upper"], 1, C.c_void_p(st.cuda_stream), C.byref(hh)))
        _lib.check(L.pk_chrom_set_score_stream(hh, C.c_void_p(score_stream.cuda_stream)))
        handles.append(hh)
    for i in range(2 * NFLIGHT):
        device_step(handles[i % NFLIGHT])
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(NFLIGHT)]
    e0.record(torch.cuda.current_stream())
    for st in streams:
        st.wait_event(e0)
    for i in range(args.steps):
        device_step(handles[i % NFLIGHT])
    for st, e in zip(streams, e1):
        e.record(st)
    barrier()
    dev_ms = max(e0.elapsed_time(e) for e in e1)
    for hh in handles:
        nr2 = C.c_int64()
        _lib.check(L.pk_chrom_result_count(hh, C.byref(nr2), None, None))   # also checks the device flags
        assert nr2.value == nrec.value
        _lib.check(L.pk_chrom_destroy(hh))
    for st in streams:
        _lib.check(L.pk_stream_destroy(local, C.c_void_p(st.cuda_stream)))
    _lib.check(L.pk_chrom_destroy(h))
    del d_rp, d_b2, d_cnt, d_w

    # ---- end to end through the public API (shard.score_units, the engine of score_genome) with HOST buffers:
    # per chromosome an H2D of its columns from pinned memory, the kernels, a D2H of its records ----
    pm = PinnedMap(nd_enc=(wl["upper"] + 2 * w + 1 + 31) // 32 * 32)
    pm.add("chr1", ch)
    encs = [e for e in ("rows", "csr16", "csr32") if e != "csr16" or "d16" in pm.ch["chr1"]]

    def e2e_run(k, enc):
        units = []
        for i in range(k):
            pm.alias["c%d" % i] = "chr1"
            units.append(("c%d" % i, 0, n))
        return shard.score_units(pm, units, flat, correct="weight", lower=wl["lower"], upper=wl["upper"],
                                 res=wl["res"], device=local, min_prob=0.5, copy=False, encoding=enc)

    e2e_ms, rec_x = {}, {}
    for enc in encs:
        e2e_run(max(12, args.steps), enc)      # warm the engine's handles and its pinned staging
        barrier()
        t0 = time.perf_counter()
        res = e2e_run(args.steps, enc)
        torch.cuda.synchronize()
        e2e_ms[enc] = 1e3 * (time.perf_counter() - t0)
        rec_x[enc] = res["c0"][0]["x"].copy()
        assert all(np.array_equal(rec_x[enc], res["c%d" % i][0]["x"]) for i in range(args.steps))
    assert all(np.array_equal(rec_x[encs[0]], v) for v in rec_x.values()) and rec_x[encs[0]].size == nrec.value
    clocks = sampler.stop() if rank == 0 else None
    d2h = 28 * int(nrec.value) + 64 + 4 * (px // 100000 + 2)

    # max over ranks
    t = torch.tensor([dev_ms] + [e2e_ms[e] for e in encs], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t[0])
    e2e_ms = {e: float(v) for e, v in zip(encs, t[1:].tolist())}

    # ---- score_genome blocks (strong scaling of BASELINE configs[2] / [3]) ----
    want = args.genome
    if want == "auto":
        want = "c3,c4" if world >= 8 else "c3"
    genome = {}
    gsteps = args.genome_steps or min(args.steps, 10)
    for name in [g for g in want.split(",") if g and g != "none"]:
        gflat = FlatForest.load(os.path.join(ROOT, "bench_data", GENOMES[name]["forest"] + "_forest.npz"))
        genome[name] = run_genome(args, name, gflat, rank, world, local, gsteps, ("rows", "csr32"))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    steps = args.steps
    value = world * px * steps / (dev_ms * 1e-3)
    stage = dict(zip(("band_build", "diag_sums", "expected_fit", "candidate_scan", "features", "forest", "emit"),
                     (stage_acc[:7] / k1).tolist()))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    # algorithmic bytes of one chromosome: pixel columns once, weights,
    # expected curve, emitted records, forest tables once
    forest_bytes = 8 * flat.n_nodes
    bytes_alg = 12 * nnz + 8 * n + 8 * (wl["upper"] + 2 * w + 1) + 24 * int(nrec.value) + forest_bytes
    dom = max(stage, key=stage.get)
    dom_ms = stage[dom]
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(args.workload, {}).get(dom, {}).get("dram_bytes")
    except (OSError, ValueError):
        pass
    kernel_names = {"features": "k_score_fused (window features + forest, fused)", "forest": "k_forest",
                    "band_build": "k_band_csr", "diag_sums": "k_diag_sums", "expected_fit": "k_fit_expected",
                    "candidate_scan": "k_cand_mark+k_scan2+k_cand_write", "emit": "k_emit"}
    roofline = {"bound": "hbm", "kernel": kernel_names.get(dom, dom), "achieved": bytes_alg / (dom_ms * 1e-3) / 1e9, "peak": peak_gbs,
                "unit": "GB/s", "frac": bytes_alg / (dom_ms * 1e-3) / 1e9 / peak_gbs, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s",
                "bytes_alg_per_launch": bytes_alg, "kernel_ms": dom_ms,
                "note": "the dominant kernel is bound by shared-memory load throughput and round-trip latency of the forest "
                        "walk plus issue/FP64 in the feature phase, not HBM (DESIGN.md section 4); frac is algorithmic "
                        "bytes of the chromosome over its duration",
                "whole_step_frac": bytes_alg / (dev_ms / steps * 1e-3) / 1e9 / peak_gbs,
                "serial_step_ms": serial_ms}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        model = load_sklearn_model(wl["forest"])
        cpu_pass(wl, 600, 5, model)
        nb = args.cpu_bins or n
        dt, cpx, ccand, crows = cpu_pass(wl, nb, 1234, model)
        cpu = {"value": cpx / dt, "unit": "pixels/s", "cores": 1, "kind": "port",
               "sample": "the workload's own chromosome (%d bins, seed 1234: %d band px, %d candidates, %d records), "
                         "numpy oracle port of score_chromosome, 1 of %d host threads (the reference is "
                         "single-threaded), %.1f s" % (nb, cpx, ccand, crows, os.cpu_count(), dt)}
        if nb == n:
            assert ccand == ncand.value and crows == nrec.value, "CPU oracle and GPU disagree on the bench map"

    def e2e_block(enc):
        ms = e2e_ms[enc] / steps
        hb = pm.h2d_bytes("chr1", enc)
        return {"value": world * px / (ms * 1e-3), "unit": "pixels/s", "ms_per_step": ms, "h2d_bytes_per_step": hb,
                "d2h_bytes_per_step": d2h, "pcie_gbs_per_gpu": (hb + d2h) / (ms * 1e-3) / 1e9, "columns": COLUMNS[enc]}
    head = e2e_block(encs[0])
    head["api"] = ("peakachu_b200.shard.score_units (the engine of score_genome: pk_engine_submit / pk_engine_collect), "
                   "pinned host buffers, six chromosomes in flight")
    line = {
        "metric": METRIC, "value": value, "unit": "pixels/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(wl, args),
        "run": {"candidates_per_step": int(ncand.value), "windows_per_step": int(nwin.value),
                "records_per_step": int(nrec.value), "forest": "%d trees, %d nodes" % (flat.n_trees, flat.n_nodes),
                "timed": "%d chromosomes, %d in flight (device-resident CSR columns; short stages on high-priority streams, fused kernels back to back on one stream)" % (steps, args.inflight)},
        "candidates_per_s": world * int(ncand.value) * steps / (dev_ms * 1e-3),
        "stage_ms": stage, "roofline": roofline, "cpu_baseline": cpu, "e2e": head,
    }
    for enc in encs[1:]:
        line["e2e_%s_columns" % ("uint16" if enc == "csr16" else "int32")] = e2e_block(enc)
    line["genome"] = genome
    if world == 1 and not args.no_file_e2e and args.workload != "c4":
        line["e2e_file"] = run_file_e2e(wl, ch, local)
    line["gpu_launches"] = KERNELS_PER_STEP * steps
    line["clocks"] = clocks
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
