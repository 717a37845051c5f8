"""Detector-level sharding of a visit over the GPUs of one box.

The reference's only scaling mechanism is GalSim's ``output.nproc``: one OS
process per output file = per CCD (config/imsim-config.yaml:326, imsim/ccd.py:72-89,
tests/test_multiproc.py:64).  The B200 path keeps that unit: one process per GPU,
each simulating its own list of detectors; photons never cross detectors, so there
is NO collective on the photon path.  ``torch.distributed`` (NCCL on the GPU box,
gloo in the CPU tests) is used only for the end-of-visit gather of per-CCD metadata.
"""
from __future__ import annotations

from typing import Dict, List, Sequence


def lpt_partition(costs: Dict[str, float], n_ranks: int) -> List[List[str]]:
    """Longest-processing-time-first assignment of detectors to ranks by estimated
    photon count (bright-star CCDs dominate; 189 = 8*23+5 CCDs => <= 24 per GPU).
    Deterministic: ties broken by detector name."""
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    shards: List[List[str]] = [[] for _ in range(n_ranks)]
    load = [0.0] * n_ranks
    for det in sorted(costs, key=lambda d: (-costs[d], d)):
        r = min(range(n_ranks), key=lambda k: (load[k], k))
        shards[r].append(det)
        load[r] += costs[det]
    return shards


def my_detectors(costs: Dict[str, float], rank: int, world_size: int) -> List[str]:
    return lpt_partition(costs, world_size)[rank]


def gather_visit_metadata(local: Sequence[dict], group=None) -> List[dict]:
    """End-of-visit gather of the per-CCD records (photons simulated, flux deposited,
    boundary-photon counts, timings).  Every rank returns the full list ordered by
    detector name.  The e-images themselves are written by the owning rank and never move."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sorted(local, key=lambda r: r["det_name"])
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, list(local), group=group)
    merged = [rec for part in out for rec in part]
    names = [r["det_name"] for r in merged]
    if len(set(names)) != len(names):
        raise RuntimeError("a detector was simulated by more than one rank")
    return sorted(merged, key=lambda r: r["det_name"])
