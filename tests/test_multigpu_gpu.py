"""Tests that need TWO GPUs on the box (skipped on a 1-GPU lease; run with `gpurun --gpus 2`):
  * two contexts on two devices in ONE process (the cudaFuncSetAttribute opt-ins and the SM count are per device);
  * config 5 over NCCL, world size 2: keyframes sharded by id, query features all-gathered as device tensors and handed to
    airfe_reloc_match by device pointer; both ranks end with the single-process answer."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _two_gpus():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


def test_two_devices_in_one_process():
    if not _two_gpus():
        pytest.skip("needs 2 GPUs")
    import _parity as P
    from airslam_b200 import capi
    from oracle import synth
    l, r, _ = synth.stereo_pair(752, 480, 91)
    outs = []
    ctxs = [capi.Context(device=d, max_batch=1, enable_superpoint=0) for d in (0, 1)]     # both alive at once
    for c in ctxs:
        outs.append(c.stereo_batch(capi.NET_PLNET, capi.MATCHER_LIGHTGLUE, l[None], r[None], lines=True, junctions=True)[0])
    for c in ctxs:
        c.close()
    a, b = outs
    P.exact("device 0 == device 1 in one process (all outputs, bitwise)",
            np.array_equal(a["feat_l"], b["feat_l"]) and np.array_equal(a["lines_l"], b["lines_l"]) and np.array_equal(a["junc"], b["junc"]) and
            np.array_equal(a["matches"][0], b["matches"][0]) and np.array_equal(a["matches"][1], b["matches"][1]))
    assert len(a["matches"][0]) > 100


def _problem():
    from oracle import synth
    n_kf, n_q = 20, 6
    kfs = [synth.keypoint_set(400, 752, 480, 0xA1750005 + k) for k in range(n_kf)]
    rs = np.random.RandomState(11)
    src = rs.permutation(n_kf)[:n_q]
    queries = [synth.keypoint_set(380, 752, 480, 700 + q, perturb_of=kfs[src[q]])[0] for q in range(n_q)]
    cand = np.zeros((n_q, 3), dtype=np.int64)
    for q in range(n_q):
        others = [k for k in rs.permutation(n_kf) if k != src[q]][:2]
        c = [src[q]] + others
        rs.shuffle(c)
        cand[q] = c
    return n_kf, n_q, kfs, src, queries, cand


def _nccl_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    from airslam_b200 import capi, dist as D, reloc
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    n_kf, n_q, kfs, src, queries, cand = _problem()
    kb, ke = D.shard_range(n_kf, rank, world)
    qb, qe = D.shard_range(n_q, rank, world)
    ctx = capi.Context(device=rank, max_batch=4, enable_superpoint=0, enable_plnet=0)
    reloc.upload_keyframes(ctx, kfs[kb:ke])
    best, cnt, table = reloc.relocalize(ctx, capi.MATCHER_LIGHTGLUE, queries[qb:qe], cand, n_kf, rank, world)
    ctx.close()
    out[rank] = (best.tolist(), cnt.tolist(), table.tolist())
    dist.destroy_process_group()


def test_config5_relocalization_nccl_two_ranks():
    if not _two_gpus():
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import _parity as P
    from airslam_b200 import capi, reloc
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, port, out), nprocs=2, join=True)
    n_kf, n_q, kfs, src, queries, cand = _problem()
    ctx = capi.Context(max_batch=4, enable_superpoint=0, enable_plnet=0)
    reloc.upload_keyframes(ctx, kfs)
    best, cnt, table = reloc.relocalize(ctx, capi.MATCHER_LIGHTGLUE, queries, cand, n_kf)
    ctx.close()
    assert np.array_equal(best, src)
    P.exact("config 5 over NCCL (2 ranks) == single process (winner, counts, table)",
            out[0] == out[1] and out[0] == (best.tolist(), cnt.tolist(), table.tolist()))
