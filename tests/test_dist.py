"""N > 1 path on CPU: world_size-2 gloo processes merging LOD accumulators exactly as
LODscores::merge_results does (lod_score.h:98-105), and the replicate placement."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import ROOT
from oracle import orcapi


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, raws, counts, out):
    sys.path.insert(0, ROOT)
    from swiftlink_b200 import dist as sdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    merged, count = sdist.merge_lod(torch.from_numpy(raws[rank].copy()), counts[rank])
    ok, bad = sdist.merge_swap_stats([3 + rank, 5], [1, 2 * rank])
    out[rank] = (merged.numpy().copy(), count, ok.tolist(), bad.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_merge_lod_two_ranks_gloo():
    rng = np.random.default_rng(3)
    n = 495
    raws = [rng.uniform(-40.0, -20.0, n), rng.uniform(-40.0, -20.0, n)]
    big = -np.finfo(np.float64).max
    raws[0][5] = big                      # empty on one rank
    raws[0][7] = big; raws[1][7] = big    # empty on both
    counts = [17, 21]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), raws, counts, out), nprocs=2, join=True)
    want = np.array([orcapi.log_sum(a, b) for a, b in zip(raws[0], raws[1])])   # logarithms.cc:14-23
    for r in (0, 1):
        merged, count, ok, bad = out[r]
        assert count == 38
        assert ok == [7, 10] and bad == [2, 2]          # MC3 swap counters summed over the two ladders
        assert merged[7] == big and merged[5] == raws[1][5]
        assert np.abs(merged - want).max() <= 1e-12 * np.abs(want[want > big]).max()


def test_single_process_merge_is_identity_and_normalise():
    from swiftlink_b200 import dist as sdist
    raw = torch.tensor([-30.0, -31.5], dtype=torch.float64)
    m, c = sdist.merge_lod(raw, 4)
    assert c == 4 and torch.equal(m, raw)
    lod = sdist.normalise(raw, 4, -21.0)
    assert abs(float(lod[0]) - orcapi.lod_normalise(-30.0, 4, -21.0)) < 1e-15


def test_chain_placement():
    from swiftlink_b200 import dist as sdist
    assert sdist.chain_placement(8, 8) == [[i] for i in range(8)]
    assert sdist.chain_placement(8, 2) == [[0, 2, 4, 6], [1, 3, 5, 7]]
    assert sum(len(x) for x in sdist.chain_placement(10, 4)) == 10


def test_mc3_ladder():
    """mc3.cc:31-42: T_0 = 1, T_i = 1 / (1 + 0.001 * 2^i), or the user's temperatures"""
    from swiftlink_b200 import host as H
    assert H.mc3_temperature(0, 4) == 1.0
    for i in range(1, 8):
        assert H.mc3_temperature(i, 8) == 1.0 / (1.0 + 0.001 * 2.0 ** i)
    assert H.mc3_temperature(2, 3, [1.0, 0.9, 0.5]) == 0.5
