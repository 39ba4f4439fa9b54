"""score_genome under torchrun with two ranks (one per GPU): chromosomes and band row
tiles sharded across the ranks, host-side gather, output identical to the reference's
bedpe. Skipped on a box with fewer than two GPUs."""
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


@pytest.mark.parametrize("name,chroms", [("genome", ["#", "X"]), ("c1", ["1"])])
def test_score_genome_two_ranks(name, chroms, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    case = Case(name)
    cfg = case.cfg
    cool = case.write_cool(tmp_path)
    out = os.path.join(str(tmp_path), "multi.bedpe")
    _run_torchrun(2, ["score_genome", "-p", cool, "-m", case.pkl, "-O", out, "-r", str(cfg["res"]),
                      "-l", str(cfg["lower"]), "-u", str(cfg["upper"]), "--minimum-prob", str(cfg["min_prob"]),
                      "--clr-weight-name", cfg["weight"], "-C"] + chroms, tmp_path)
    assert open(out).read() == case.bedpe


def test_two_devices_in_one_process(tmp_path):
    """Handles on different GPUs in one process (the C ABI takes a device per handle): per-device
    kernel attributes, allocator caches and streams; the second device gives the same bedpe."""
    import argparse

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from peakachu_b200 import score_chromosome
    case = Case("c1")
    cfg = case.cfg
    cool = case.write_cool(tmp_path)
    for dev in (0, 1, 0):
        out = os.path.join(str(tmp_path), "dev%d.bedpe" % dev)
        score_chromosome.main(argparse.Namespace(path=cool, model=case.pkl, output=out, resolution=cfg["res"],
                                                 lower=cfg["lower"], upper=cfg["upper"], minimum_prob=cfg["min_prob"],
                                                 clr_weight_name=cfg["weight"], chrom=case.chroms[0].name, device=dev))
        assert open(out).read() == case.bedpe
