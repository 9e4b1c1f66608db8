"""CPU, world_size 2, gloo: the multi-GPU plumbing of bench.py (pair sharding, max-over-ranks timing, whole-job
throughput, the relocalization all-gather) -- same code path the GPU box runs over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from airslam_b200 import dist as D


def test_shard_range_partitions():
    for total in (0, 1, 7, 8, 64, 10000):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                b, e = D.shard_range(total, r, world)
                seen += list(range(b, e))
                for i in range(b, e):
                    assert D.shard_owner(i, total, world) == r
            assert seen == list(range(total))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = D.shard_range(9, rank, world)
    t = 1.0 + rank          # rank 1 is slower
    thr = D.aggregate_throughput(e - b, t)
    mx = D.max_over_ranks(t)
    feat = torch.full((e - b, 4, 259), float(rank))
    cnt = torch.full((e - b,), rank + 1, dtype=torch.int32)
    # all_gather needs equal shapes: pad to the largest shard like the relocalization path does
    q = 5
    fpad = torch.zeros(q, 4, 259); fpad[: e - b] = feat
    cpad = torch.zeros(q, dtype=torch.int32); cpad[: e - b] = cnt
    fa, ca = D.all_gather_features(fpad, cpad)
    out[rank] = (b, e, thr, mx, tuple(fa.shape), ca.tolist())
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    (b0, e0, thr0, mx0, shp0, c0), (b1, e1, thr1, mx1, shp1, c1) = out[0], out[1]
    assert (b0, e0, b1, e1) == (0, 5, 5, 9)
    assert mx0 == mx1 == 2.0
    assert abs(thr0 - 9 / 2.0) < 1e-12 and thr0 == thr1          # all units / slowest rank
    assert shp0 == (10, 4, 259) and c0 == c1 == [1] * 5 + [2] * 4 + [0]


def _reloc_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from airslam_b200 import reloc
    n_kf, q_total = 10, 5
    qb, qe = D.shard_range(q_total, rank, world)
    # query i has (10 + i) keypoints whose score row encodes the query id
    qf = []
    for i in range(qb, qe):
        f = np.zeros((259, 10 + i), dtype=np.float32)
        f[0] = i
        qf.append(f)
    cand = np.array([[(3 * q + c) % n_kf for c in range(3)] for q in range(q_total)])
    # mock matcher: "number of matches" = 100*query + keyframe id, computed only by the owner of the keyframe
    seen = {}

    def fake(jobs, feat_all, cnt_all):
        # the all-gathered query tensor carries every rank's queries in global order: row q has 10 + q valid columns, score row == q
        assert feat_all.shape[0] == q_total and cnt_all.tolist() == [10 + i for i in range(q_total)]
        for q, c, kf in jobs:
            assert float(feat_all[q, 0, 0]) == float(q) and float(feat_all[q, 10 + q - 1, 0]) == float(q)
            seen[(q, kf)] = True
        return [100 * q + kf for q, c, kf in jobs]
    best, cnt, table = reloc.relocalize(None, 0, qf, cand, n_kf, rank, world, device="cpu", match_fn=fake)
    out[rank] = (best.tolist(), cnt.tolist(), table.tolist(), sorted(seen))
    dist.destroy_process_group()


def test_relocalization_exchange_two_ranks():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_reloc_worker, args=(2, port, out), nprocs=2, join=True)
    b0, c0, t0, s0 = out[0]
    b1, c1, t1, s1 = out[1]
    assert (b0, c0, t0) == (b1, c1, t1)                           # every rank ends with the same decision
    exp = [[100 * q + (3 * q + c) % 10 for c in range(3)] for q in range(5)]
    assert t0 == exp
    # winner rule of src/map_user.cc:370-373: strictly more matches than every earlier candidate
    assert c0 == [max(r) for r in exp] and b0 == [[(3 * q + c) % 10 for c in range(3)][int(np.argmax(exp[q]))] for q in range(5)]
    assert not (set(map(tuple, s0)) & set(map(tuple, s1)))       # each (query, keyframe) job ran on exactly one rank
    assert all(kf < 5 for _, kf in s0) and all(kf >= 5 for _, kf in s1)
