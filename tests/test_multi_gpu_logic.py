"""CPU (gloo, world_size 2): the host-side logic of the one-process-per-GPU prove — point-range shares, the
gather of per-rank partial sums and the host combine (za_point_sum) — with the oracle standing in for the
per-rank multiexps.  The GPU version of the same flow is tests/test_multi_gpu.py / bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import oracle as O
from tests import pyref as P


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_share_rule_partitions_every_count():
    import za_b200
    for count in (0, 1, 2, 7, 8, 1000, (1 << 20) - 1):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for rank in range(world):
                lo, hi = za_b200.share(count, rank, world)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == count


def test_weighted_share_rule():
    """Rank 0 also runs the H pipeline, so it may take a smaller part of the witness multiexps: the ranges still
    partition every count, equal weights reproduce the plain rule, and rank 0's share shrinks with its weight."""
    import za_b200
    for count in (0, 1, 5, 1000, (1 << 20) + 3):
        for world in (1, 2, 4, 8):
            for w in (1.0, 0.5, 0.106, 0.001):
                prev = 0
                for rank in range(world):
                    lo, hi = za_b200.share_weighted(count, rank, world, w)
                    assert lo == prev and hi >= lo
                    if w == 1.0:
                        assert (lo, hi) == za_b200.share(count, rank, world)
                    prev = hi
                assert prev == count
    lo, hi = za_b200.share_weighted(8000, 0, 8, 0.106)
    assert hi - lo == 8000 * 106 // 7106
    assert za_b200.share_weighted(8000, 0, 8, 1.0) == (0, 1000)


def _worker(rank, world, port, n, q):
    import za_b200
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bases = O.g1_multiples(n)
        scalars = O.random_frs(n, 99)
        lo, hi = za_b200.share(n, rank, world)
        rc, part = O.multiexp("g1", bases[lo:hi], scalars[lo:hi]) if hi > lo else (0, b"\0" * 64)
        assert rc == 0
        # XYZZ record of an affine point: X = x, Y = y, ZZ = ZZZ = 1 (infinity: all zero)
        one = (1).to_bytes(32, "little")
        rec = part + (one + one if part != b"\0" * 64 else b"\0" * 64)
        t = torch.from_numpy(np.frombuffer(rec, np.uint8).copy())
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        if rank == 0:
            total = za_b200.point_sum(1, [g.numpy().tobytes() for g in gathered])
            rc, full = O.multiexp("g1", bases, scalars)
            q.put((total == full, total.hex()[:16]))
    finally:
        dist.destroy_process_group()


def test_two_rank_partial_sums_combine_to_the_full_multiexp():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 301, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, _ = q.get(timeout=5)
    assert ok


def test_point_sum_handles_infinity_and_opposites():
    import za_b200
    one = (1).to_bytes(32, "little")
    g = O.g1_bytes(P.G1_GEN)
    neg = O.g1_bytes(P.ec_neg(P.Fq1Ops, P.G1_GEN))
    inf = b"\0" * 128
    assert za_b200.point_sum(1, [g + one + one, inf]) == g
    assert za_b200.point_sum(1, [g + one + one, neg + one + one]) == b"\0" * 64
    assert za_b200.point_sum(1, [g + one + one, g + one + one]) == O.g1_bytes(P.g1_mul(P.G1_GEN, 2))
