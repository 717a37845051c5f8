"""Multi-GPU host logic on CPU: LPT detector sharding and the gloo metadata gather (world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

from imsim_b200.detector import lsstcam_science_detectors
from imsim_b200.sharding import gather_visit_metadata, lpt_partition


def test_lpt_partition_covers_every_detector_once():
    dets = lsstcam_science_detectors()
    rng = np.random.default_rng(0)
    costs = {d: float(c) for d, c in zip(dets, rng.lognormal(18, 0.7, len(dets)))}
    for n in (1, 2, 4, 8):
        shards = lpt_partition(costs, n)
        flat = [d for s in shards for d in s]
        assert sorted(flat) == sorted(dets)
        loads = [sum(costs[d] for d in s) for s in shards]
        assert max(loads) <= min(loads) + max(costs.values())  # LPT bound
        assert max(len(s) for s in shards) <= -(-len(dets) // n) + 6
    assert lpt_partition(costs, 8) == lpt_partition(dict(reversed(list(costs.items()))), 8)  # deterministic


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dets = lsstcam_science_detectors()[:10]
    costs = {d: 1.0 + i for i, d in enumerate(dets)}
    mine = lpt_partition(costs, world)[rank]
    local = [{"det_name": d, "photons": int(costs[d] * 1000), "rank": rank} for d in mine]
    allrec = gather_visit_metadata(local)
    dist.destroy_process_group()
    q.put((rank, [r["det_name"] for r in allrec], sum(r["photons"] for r in allrec)))


def test_gloo_metadata_gather_world_size_2():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dets = sorted(lsstcam_science_detectors()[:10])
    for rank, names, total in res:
        assert names == dets
        assert total == sum(int((1.0 + i) * 1000) for i in range(10))


def test_run_many_keeps_two_detectors_queued_and_yields_in_order():
    """Host logic of the software-pipelined visit loop (imsim_b200/visit.py, no GPU): detector k+2 is prepared
    while k and k+1 are queued, k is finished before k+2 is launched, results come back in job order."""
    from imsim_b200.visit import DetectorRunner

    class Fake(DetectorRunner):
        def __init__(self):
            self.log = []

        def prepare(self, name):
            self.log.append(("prepare", name))
            return name

        def launch(self, prep):
            self.log.append(("launch", prep))
            return prep

        def finish(self, h):
            self.log.append(("finish", h))
            return {"det_name": h}, None

    for n in (0, 1, 2, 3, 6):
        r = Fake()
        out = [rec["det_name"] for rec, _ in r.run_many(dict(name=k) for k in range(n))]
        assert out == list(range(n))
        launched, finished = set(), set()
        for op, k in r.log:
            if op == "launch":
                launched.add(k)
                assert len(launched - finished) <= 2  # never more than two detectors on the stream
            elif op == "finish":
                assert k in launched
                finished.add(k)
            else:  # prepare(k) runs while k-1 and k-2 (if any) are still queued: the overlap
                assert {j for j in (k - 1, k - 2) if j >= 0} <= launched - finished
        assert finished == set(range(n))


def _visit_line_worker(rank, world, port, q, fail_rank):
    import argparse

    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def fake_simulate(opts, rank, world, local, barrier=None):
        assert barrier is None and opts.visit_repeat == 2 and opts.visit_catalog and opts.visit_readout
        if rank == fail_rank:
            raise RuntimeError("boom on rank %d" % rank)
        n = 95 if rank == 0 else 94
        recs = [{"photons": 1000 + rank, "gpu_ms": 10.0 * (rank + 1)} for _ in range(n)]
        return [3.0 + rank, 2.0 + 0.5 * rank], recs

    bench.simulate_visit = fake_simulate
    args = argparse.Namespace(visit_ccds=189, visit_photons=1e8)
    out = bench.visit_for_line(args, rank, world, 0, torch.device("cpu"))
    dist.destroy_process_group()
    q.put((rank, out))


@pytest.mark.parametrize("fail_rank", [-1, 1])
def test_bench_visit_summary_reductions_world_size_2(fail_rank):
    """bench.py's `visit` sub-object on two ranks (gloo): wall and GPU time are maxima over the ranks, photons and
    CCDs sums; a rank whose simulation raises still joins the reductions and every rank reports the failure."""
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_visit_line_worker, args=(r, 2, port, q, fail_rank)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        out = res[rank]
        if fail_rank >= 0:
            assert "error" in out and ("boom" in out["error"] or out["error"] == "a rank failed")
        else:
            assert out["ccds"] == 189 and out["photons"] == 95 * 1000 + 94 * 1001 and out["n_gpus"] == 2
            assert out["wall_s_max_rank"] == 2.5 and out["first_visit_wall_s_max_rank"] == 4.0
            assert abs(out["gpu_s_max_rank"] - 94 * 0.020) < 1e-12 and abs(out["visits_per_hour"] - 1440.0) < 1e-9
            assert out["note"] == ""
