"""score_genome under torchrun with two or three ranks: chromosomes and band row tiles sharded
across the ranks, host-side gather, output identical to the reference's bedpe. Every case runs
with all ranks on device 0 (a one-GPU box exercises the whole N-rank path) and, when the box
has a GPU per rank, again with one rank per GPU."""
import os
import socket
import subprocess
import sys

import pytest

from tests.cases import Case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_torchrun(nproc, argv, tmp):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), "-m", "peakachu_b200"] + argv
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r


def _score_genome_args(case, cool, out, chroms):
    cfg = case.cfg
    return ["score_genome", "-p", cool, "-m", case.pkl, "-O", out, "-r", str(cfg["res"]),
            "-l", str(cfg["lower"]), "-u", str(cfg["upper"]), "--minimum-prob", str(cfg["min_prob"]),
            "--clr-weight-name", cfg["weight"], "-C"] + chroms


@pytest.mark.parametrize("name,chroms,world", [("genome", ["#", "X"], 2), ("c1", ["1"], 2), ("gnames", [], 3),
                                               ("batchrule", [], 2), ("batchrule", [], 3)])
def test_score_genome_ranks_sharing_one_device(name, chroms, world, tmp_path):
    """The N-rank path of score_genome (plan, chromosome shards, band row tiles of a chromosome larger
    than an even share, shared-memory gather, per-batch window sums over the tiles) with every rank on
    device 0 (--device 0), so that it runs on a one-GPU box too. `batchrule` is a single chromosome: its
    rows are tiled over the ranks and the 100,000-candidate batch rule is decided on the gathered counts."""
    import torch
    case = Case(name)
    cool = case.write_cool(tmp_path)
    out = os.path.join(str(tmp_path), "multi.bedpe")
    _run_torchrun(world, _score_genome_args(case, cool, out, chroms) + ["--device", "0"], tmp_path)
    assert open(out).read() == case.bedpe
    if torch.cuda.device_count() >= world:
        # the same run with one rank per GPU (LOCAL_RANK picks the device)
        out2 = os.path.join(str(tmp_path), "multi_spread.bedpe")
        _run_torchrun(world, _score_genome_args(case, cool, out2, chroms), tmp_path)
        assert open(out2).read() == case.bedpe


def test_two_devices_in_one_process(tmp_path):
    """Handles on different GPUs in one process (the C ABI takes a device per handle): per-device
    kernel attributes, allocator caches and streams; the second device gives the same bedpe. On a
    one-GPU box the same sequence runs on device 0 alone (handles re-created between runs)."""
    import argparse

    import torch
    devices = (0, 1, 0) if torch.cuda.device_count() >= 2 else (0, 0)
    from peakachu_b200 import score_chromosome
    case = Case("c1")
    cfg = case.cfg
    cool = case.write_cool(tmp_path)
    for dev in devices:
        out = os.path.join(str(tmp_path), "dev%d.bedpe" % dev)
        score_chromosome.main(argparse.Namespace(path=cool, model=case.pkl, output=out, resolution=cfg["res"],
                                                 lower=cfg["lower"], upper=cfg["upper"], minimum_prob=cfg["min_prob"],
                                                 clr_weight_name=cfg["weight"], chrom=case.chroms[0].name, device=dev))
        assert open(out).read() == case.bedpe
