"""The N > 1 host path of score_genome on CPU: two gloo ranks take their share of the
plan (chromosomes and band row tiles), "score" it (records cut from the reference's
golden bedpe), rank 0 gathers on the host and must reproduce the file byte for byte."""
import os
import socket

import numpy as np
import pytest

from peakachu_b200 import shard
from tests.cases import GOLDEN


def _records():
    rec = {}
    for line in open(os.path.join(GOLDEN, "genome.bedpe")):
        p = line.split("\t")
        key = p[0][3:]
        rec.setdefault(key, []).append((int(p[1]) // 10000, int(p[4]) // 10000, float(p[6]), float(p[7])))
    return {k: np.array(v) for k, v in rec.items()}


SIZES = {"1": 700, "2": 450, "X": 520}


def _fake_units(units, rec):
    out = {}
    for key, a, b in units:
        r = rec[key]
        sel = r[(r[:, 0] >= a) & (r[:, 0] < b)]
        out.setdefault(key, []).append(dict(
            row_begin=a, whole=(a == 0 and b == SIZES[key]), x=sel[:, 0].astype(np.int32), y=sel[:, 1].astype(np.int32),
            p=sel[:, 2].copy(), v=sel[:, 3].copy(), batch=np.zeros(len(sel), np.int32),
            batch_windows=np.array([max(len(sel), 2)], np.int64), n_candidates=len(sel)))
    return out


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rec = _records()
    queue = ["1", "2", "X"]
    asg = shard.plan(SIZES, world, 6, 80, 5)
    assert any(len({k for k, _, _ in u}) for u in asg)
    mine = _fake_units(asg[rank], rec)
    for i in range(4):                 # the shared-memory segments are reused from pass to pass, and regrown
        big = dict(mine, pad=[dict(row_begin=0, junk=np.arange((i % 2) * 400000 + rank, dtype=np.int64))]) if i < 3 else mine
        zero_copy = i == 1                     # rank 0 reads views of the peers' blocks and hands them back itself
        gathered = shard.gather_to_rank0(big, rank, world, copy=not zero_copy)
        if rank == 0 and i < 3:
            for r in range(world):
                assert np.array_equal(gathered[r]["pad"][0]["junk"], np.arange((i % 2) * 400000 + r))
            if zero_copy:
                assert shard._SHM["mode"] != "shm" or not gathered[1]["pad"][0]["junk"].flags["OWNDATA"]
                gathered = None
                shard.release_gathered()
    assert shard._SHM["mode"] == os.environ.get("PEAKACHU_B200_GATHER", "shm")
    if rank == 0:
        assert len(gathered) == world
        text = shard.assemble_text(queue, gathered, 10000)
        with open(os.path.join(tmp, "out.bedpe"), "w") as fh:
            for k in queue:
                fh.write(text[k])
    dist.barrier()
    dist.destroy_process_group()


def _worker_early_exit(rank, world, port, tmp):
    """Rank 1 publishes one pass and leaves at once (no barrier, as the torchrun CLI path does); rank 0
    turns up a second later. Rank 1's exit hook must keep its data segment alive until rank 0 has read it."""
    import time

    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard._setup_shm(rank, world, shard._host_group())          # the one collective of the gather
    assert shard._SHM["mode"] == "shm"
    payload = dict(c=[dict(row_begin=0, junk=np.arange(300000, dtype=np.int64) + rank)])
    if rank == 0:
        time.sleep(1.0)
    gathered = shard.gather_to_rank0(payload, rank, world)
    if rank == 0:
        assert np.array_equal(gathered[1]["c"][0]["junk"], np.arange(300000, dtype=np.int64) + 1)
        open(os.path.join(tmp, "ok"), "w").write("ok")
    # no barrier: a non-zero rank returns immediately and its atexit hook runs
    if rank != 0:
        shard._release_shm()


def test_rank_that_exits_right_after_the_gather_keeps_its_segment(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker_early_exit, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok"))


@pytest.mark.parametrize("how", ["shm", "pickle"])
def test_two_rank_gather_reproduces_bedpe(tmp_path, how, monkeypatch):
    """Both host gathers: numpy columns through shared memory (one node) and pickled objects."""
    import torch.multiprocessing as mp
    monkeypatch.setenv("PEAKACHU_B200_GATHER", how)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(os.path.join(str(tmp_path), "out.bedpe")).read() == open(os.path.join(GOLDEN, "genome.bedpe")).read()


def test_plan_splits_large_chromosome_into_row_tiles():
    asg = shard.plan({"1": 24926, "21": 4813}, 4, 6, 300, 5)
    tiles = sorted((a, b) for u in asg for k, a, b in u if k == "1")
    assert len(tiles) == 4 and tiles[0][0] == 0 and tiles[-1][1] == 24926
    assert all(t0[1] == t1[0] for t0, t1 in zip(tiles, tiles[1:]))


def test_bind_to_device_node_reads_sysfs(tmp_path, monkeypatch):
    """NUMA binding of a rank: PCI bus id of the device -> numa_node -> cpulist -> sched_setaffinity
    (sysfs faked; the affinity call recorded, not applied)."""
    from peakachu_b200 import _lib

    class Stub:
        @staticmethod
        def pk_device_pci_bus_id(device, buf, n):
            buf.value = b"0000:1B:00.0"
            return 0
    monkeypatch.setattr(_lib, "lib", lambda: Stub)
    dev = tmp_path / "bus/pci/devices/0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    node = tmp_path / "devices/system/node/node1"
    node.mkdir(parents=True)
    have = sorted(os.sched_getaffinity(0))
    (node / "cpulist").write_text("%d-%d,9999\n" % (have[0], have[-1]))
    calls = []
    monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: calls.append((pid, set(cpus))))
    assert shard.bind_to_device_node(0, sysfs=str(tmp_path)) == 1
    assert calls == [(0, set(have))]                      # intersected with the current affinity
    (dev / "numa_node").write_text("-1\n")                # topology not exposed
    assert shard.bind_to_device_node(0, sysfs=str(tmp_path)) is None
    monkeypatch.setenv("PEAKACHU_B200_NUMA", "0")
    (dev / "numa_node").write_text("1\n")
    assert shard.bind_to_device_node(0, sysfs=str(tmp_path)) is None
    assert len(calls) == 1


def test_plan_covers_every_row_once_and_balances():
    """Property test of the greedy plan: every band row of every chromosome is assigned
    exactly once, tiles of a chromosome are contiguous, and no rank carries more than the even share plus
    the largest unit (the bound of greedy largest-first)."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(min_value=1, max_value=30000), min_size=1, max_size=30),
           st.integers(min_value=1, max_value=8), st.sampled_from([(6, 300, 5), (6, 600, 7), (6, 60, 5)]))
    def check(sizes, world, lu):
        lower, upper, w = lu
        named = {"c%d" % i: n for i, n in enumerate(sizes)}
        asg = shard.plan(named, world, lower, upper, w)
        assert len(asg) == world
        seen = {}
        for units in asg:
            for k, a, b in units:
                assert 0 <= a <= b <= named[k]
                seen.setdefault(k, []).append((a, b))
        assert set(seen) == set(named)
        for k, tiles in seen.items():
            tiles.sort()
            assert tiles[0][0] == 0 and tiles[-1][1] == named[k]
            assert all(t0[1] == t1[0] for t0, t1 in zip(tiles, tiles[1:]))
        cost = {k: shard.band_pixels(n, lower, upper, w) for k, n in named.items()}
        unit_cost = [cost[k] / len(seen[k]) for units in asg for k, _, _ in units]
        loads = [sum(cost[k] / len(seen[k]) for k, _, _ in units) for units in asg]
        assert max(loads) <= sum(cost.values()) / world + max(unit_cost) + 1e-6

    check()
