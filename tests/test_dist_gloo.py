"""CPU, world_size 2, gloo: the N>1 plumbing of bench.py — per-rank robot record blocks, one
all-gather per step, every rank ends up with every stream's robots in rank order."""
import os
import socket
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rm_radar_b200 import dist as rdist

MAX_CARS = 20


def fake_recs(rank):
    """Deterministic per-rank robots shaped like the C ABI's rmr_robot_t."""
    recs = []
    for i in range(3 + rank):
        located = (i + rank) % 2 == 0
        recs.append(SimpleNamespace(label=(5 * rank + i) % 12, is_detected=int(i != 1), confidence=0.5 + 0.01 * i,
                                    is_located=int(located), location=(1.0 * rank + i, 2.0, 0.5 * i),
                                    rect=(10.0, 20.0, 30.0 + i, 40.0)))
    return recs


def worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        recs = fake_recs(rank)
        block = rdist.pack_records(recs, len(recs), MAX_CARS)
        assert block.shape == (MAX_CARS, rdist.RECORD_FLOATS) and block[len(recs):].abs().sum() == 0
        for _ in range(3):      # one collective per step, buffers reused
            gathered = rdist.all_gather_records(block)
        assert gathered.shape == (world, MAX_CARS, rdist.RECORD_FLOATS)
        per_rank = rdist.unpack_records(gathered)
        for r in range(world):
            want = fake_recs(r)
            assert len(per_rank[r]) == len(want)
            for got, w in zip(per_rank[r], want):
                assert got["label"] == (w.label if w.is_detected else -1)
                assert (got["location"] is not None) == bool(w.is_located)
                if w.is_located:
                    assert got["location"] == pytest.approx(w.location)
        # capacity: more robots than max_cars are truncated, never overflow the block
        many = fake_recs(0) * 10
        assert rdist.pack_records(many, len(many), MAX_CARS).shape[0] == MAX_CARS
    finally:
        dist.destroy_process_group()


def test_all_gather_of_robot_records_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(worker, args=(2, port), nprocs=2, join=True)


def test_pack_records_fast_path_matches_generic():
    """The vectorised view of the C ABI's rmr_robot_t array and the generic per-robot path agree."""
    import ctypes
    from rm_radar_b200 import _lib
    assert ctypes.sizeof(_lib.RobotRec) == rdist.ROBOT_DTYPE.itemsize
    recs = (_lib.RobotRec * MAX_CARS)()
    for i in range(6):
        recs[i].label = i
        recs[i].is_detected = int(i != 2)
        recs[i].confidence = 0.5 + 0.1 * i
        recs[i].is_located = i % 2
        recs[i].location = (ctypes.c_float * 3)(1.0 + i, 2.0, 3.0)
        recs[i].rect = (ctypes.c_float * 4)(1, 2, 3 + i, 4)
    fast = rdist.pack_records(recs, 6, MAX_CARS)
    slow = rdist.pack_records([recs[i] for i in range(6)], 6, MAX_CARS)
    assert torch.equal(fast, slow)


def tracker_worker(rank, world, port):
    """Two camera streams, each seeing its own moving robots; every rank tracks the union after the all-gather."""
    import rm_radar_b200 as rr
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tracker = rr.Tracker([0.2, 0.2, 0.2], 12, init_thresh=3)
        t = 0
        for frame in range(6):
            t += 40_000_000
            recs = [SimpleNamespace(label=3 * rank + i, is_detected=1, confidence=0.8, is_located=1,
                                    location=(10.0 * rank + i + 0.02 * frame, 1.0 + i, 0.0), rect=(0.0, 0.0, 10.0, 10.0))
                    for i in range(2 + rank)]
            block = rdist.pack_records(recs, len(recs), MAX_CARS)
            robots = rdist.robots_from_records(rdist.all_gather_records(block))
            assert len(robots) == 2 + 3                      # rank 0 publishes 2 robots, rank 1 publishes 3
            tracker.update(robots, t)
        tracks = tracker.tracks()
        assert len(tracks) == 5 and all(k["state"] == 1 for k in tracks)          # confirmed on every rank
        assert sorted(k["label"] for k in tracks) == [0, 1, 3, 4, 5]
        assert all(r.track_state == 1 for r in robots)
        # every rank holds the same field-level picture
        summary = torch.tensor([[k["id"], k["label"], *k["location"]] for k in tracks], dtype=torch.float32)
        both = [torch.empty_like(summary) for _ in range(world)]
        dist.all_gather(both, summary)
        assert torch.allclose(both[0], both[1])
    finally:
        dist.destroy_process_group()


def test_field_level_tracker_over_gathered_records_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(tracker_worker, args=(2, port), nprocs=2, join=True)
