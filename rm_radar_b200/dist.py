"""Multi-GPU plumbing: one process per GPU, one camera + LiDAR stream per rank.

The hot path shards by stream with no data-path collective; the only exchange is the fixed-size block
of world-frame robot records every rank publishes once per step (the reference has no distributed
code at all).  On the GPU box the exchange is the library's own (`rm_radar_b200.Comm` = `rmr_comm_*`, NCCL all-gather
issued from `csrc/comm.cu`); this module is the host-side restatement of it over `torch.distributed` — the transport of
the world-size-2 gloo tests on CPU, the checker of `rmr_comm_pack`, and the unpacking of a gathered block into robots.
Record layout (8 float32 per robot, `max_cars` rows per rank):
    [valid, label (-1 = undetected), confidence, is_located, x, y, z (metres, world), rect area]
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

RECORD_FLOATS = 8

# numpy view of include/rm_radar_b200.h `rmr_robot_t` (same layout as _lib.RobotRec)
ROBOT_DTYPE = np.dtype([("rect", np.float32, 4), ("has_rect", np.int32), ("is_detected", np.int32), ("label", np.int32),
                        ("confidence", np.float32), ("n_armors", np.int32), ("armors", np.float32, (16, 6)),
                        ("is_located", np.int32), ("location", np.float32, 3), ("cluster", np.int32),
                        ("cluster_points", np.int32)])


def pack_records(recs, n: int, max_cars: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """ctypes RobotRec array (or any objects with the same fields) -> [max_cars, 8] float32 CPU tensor."""
    if out is None:
        out = torch.zeros(max_cars, RECORD_FLOATS)
    else:
        out.zero_()
    n = min(n, max_cars)
    if isinstance(recs, ctypes.Array) and ctypes.sizeof(recs._type_) == ROBOT_DTYPE.itemsize:
        # the C ABI's record array: one vectorised pass, no per-robot Python work
        a = np.frombuffer(recs, dtype=ROBOT_DTYPE, count=n)
        o = out.numpy()
        o[:n, 0] = 1.0
        o[:n, 1] = np.where(a["is_detected"] != 0, a["label"], -1)
        o[:n, 2] = a["confidence"]
        o[:n, 3] = a["is_located"] != 0
        o[:n, 4:7] = a["location"] * (a["is_located"] != 0)[:, None]
        o[:n, 7] = a["rect"][:, 2] * a["rect"][:, 3]
        return out
    for i in range(n):
        r = recs[i]
        out[i, 0] = 1.0
        out[i, 1] = float(r.label) if r.is_detected else -1.0
        out[i, 2] = float(r.confidence)
        out[i, 3] = float(r.is_located)
        if r.is_located:
            out[i, 4] = r.location[0]; out[i, 5] = r.location[1]; out[i, 6] = r.location[2]
        out[i, 7] = float(r.rect[2]) * float(r.rect[3])
    return out


def all_gather_records(block: torch.Tensor, gathered: torch.Tensor | None = None, group=None) -> torch.Tensor:
    """One collective per step: every rank's [max_cars, 8] block -> [world, max_cars, 8] on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if gathered is None:
        gathered = torch.empty(world * block.shape[0], block.shape[1], dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(gathered, block.contiguous(), group=group)
    return gathered.view(world, block.shape[0], block.shape[1])


def unpack_records(gathered: torch.Tensor):
    """[world, max_cars, 8] -> list (per rank) of dicts for the valid robots, in record order."""
    out = []
    g = gathered.cpu()
    for rank in range(g.shape[0]):
        robots = []
        for row in g[rank]:
            if row[0] < 0.5:
                continue
            robots.append(dict(label=int(row[1]), confidence=float(row[2]),
                               location=tuple(float(v) for v in row[4:7]) if row[3] > 0.5 else None,
                               area=float(row[7])))
        out.append(robots)
    return out


def robots_from_records(gathered: torch.Tensor) -> list:
    """[world, max_cars, 8] -> one flat list of `Robot` (rank order, record order) for a field-level `Tracker`.

    The exchanged record carries the winning label and its confidence, not the armour list, so a detected robot gets one
    stand-in armour (label, confidence): `Robot::feature` (robot.cpp:102-122) of it is the one-hot vector the reference
    would build from a single armour.  This is the consumer SURVEY §8f names for the tracker: "the all-gathered positions".
    """
    from . import Detection, Robot
    robots = []
    for per_rank in unpack_records(gathered):
        for r in per_rank:
            robot = Robot()
            if r["label"] >= 0:
                robot.label, robot.confidence = r["label"], r["confidence"]
                robot.armors = [Detection(0.0, 0.0, 0.0, 0.0, float(r["label"]), r["confidence"])]
            robot.location = r["location"]
            robots.append(robot)
    return robots
