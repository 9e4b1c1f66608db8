"""Batched relocalization matching (BASELINE.json config 5; the reference does it sequentially on one GPU:
MapUser::Relocalization, src/map_user.cc:363-376 -- up to GoodCandidateNum = 3 MatchingPoints calls per query, each re-uploading
the keyframe's features).

This module is only the multi-rank CHOREOGRAPHY (which rank owns which keyframe, the two collectives); everything that computes runs
behind the C ABI:
  * keyframe features live on the device of the rank that owns them (block partition by keyframe id, airslam_b200.dist.shard_range),
    uploaded once with airfe_kf_put -- the device-resident keyframe cache of SURVEY.md 8f rank 3;
  * query feature sets are produced replica-parallel (Q / world per rank) and ALL-GATHERED as device tensors (the one collective on
    the path: NCCL over NVLink on the GPU box, gloo in the CPU tests); the gathered tensor is handed to airfe_reloc_match by device
    pointer, so query features never touch host memory;
  * every rank matches all (query, candidate) jobs whose candidate it owns in batched LightGlue launches that read both sides in place;
  * the [Q, C] table of match counts is summed over ranks (disjoint ownership: sum == gather) and airfe_reloc_pick applies the reference's
    winner rule."""
import numpy as np

from . import dist as D


def local_jobs(candidates, n_keyframes, rank, world):
    """candidates: int array [Q, C] of global keyframe ids per query.  Returns the (query, slot, keyframe) jobs this rank owns."""
    jobs = []
    for q in range(candidates.shape[0]):
        for c in range(candidates.shape[1]):
            kf = int(candidates[q, c])
            if kf >= 0 and D.shard_owner(kf, n_keyframes, world) == rank:
                jobs.append((q, c, kf))
    return jobs


def upload_keyframes(ctx, keyframe_feats_local, feat_cap=None):
    """Device-resident cache of this rank's keyframes (slot i = local keyframe i)."""
    feat_cap = feat_cap or max([f.shape[1] for f in keyframe_feats_local] + [1])
    ctx.kf_reserve(max(len(keyframe_feats_local), 1), feat_cap)
    for i, f in enumerate(keyframe_feats_local):
        ctx.kf_put(i, f)


def gather_queries(feat_local, cnt_local, world):
    """feat_local: torch tensor [Ql, cap, 259] (device of the backend), cnt_local int32 [Ql].  Returns (feat_all, cnt_all) with the same
    capacity on every rank.  world == 1: identity."""
    import torch
    if world <= 1:
        return feat_local, cnt_local
    dev = feat_local.device
    mq = int(D.max_over_ranks(feat_local.shape[0], device=dev))
    mc = int(D.max_over_ranks(feat_local.shape[1], device=dev))
    fpad = torch.zeros(mq, mc, 259, device=dev)
    fpad[:feat_local.shape[0], :feat_local.shape[1]] = feat_local
    cpad = torch.full((mq,), -1, dtype=torch.int32, device=dev)
    cpad[:cnt_local.shape[0]] = cnt_local
    fa, ca = D.all_gather_features(fpad, cpad)          # the collective
    keep = ca >= 0
    return fa[keep].contiguous(), ca[keep].contiguous()


def relocalize(ctx, matcher, query_feats_local, candidates, n_keyframes, rank=0, world=1, device=None, match_fn=None):
    """query_feats_local: list of [259, N] arrays detected on this rank (queries are block-partitioned over ranks like keyframes).
    The keyframe cache of `ctx` must hold this rank's shard (upload_keyframes).  Returns (best_keyframe [Q], best_count [Q],
    table [Q, C] of match counts) -- identical on every rank.  match_fn(jobs, feat_all, cnt_all) replaces the C-ABI call in CPU tests."""
    import torch
    import torch.distributed as dist
    cap = max([f.shape[1] for f in query_feats_local] + [1])
    ql = len(query_feats_local)
    feat = torch.zeros(ql, cap, 259)
    cnt = torch.zeros(ql, dtype=torch.int32)
    for i, f in enumerate(query_feats_local):
        feat[i, :f.shape[1]] = torch.from_numpy(np.ascontiguousarray(f.T))
        cnt[i] = f.shape[1]
    if device is None:
        device = "cuda" if (world > 1 and dist.get_backend() == "nccl") or (world == 1 and ctx is not None and match_fn is None and torch.cuda.is_available()) else "cpu"
    feat, cnt = feat.to(device), cnt.to(device)
    feat_all, cnt_all = gather_queries(feat, cnt, world)
    kf_begin, _ = D.shard_range(n_keyframes, rank, world)
    jobs = local_jobs(np.asarray(candidates), n_keyframes, rank, world)
    if match_fn is not None:
        counts = match_fn(jobs, feat_all, cnt_all)
    else:
        counts = ctx.reloc_match(matcher, feat_all.data_ptr(), cnt_all.cpu().numpy(), feat_all.shape[1], [q for q, _, _ in jobs],
                                 [kf - kf_begin for _, _, kf in jobs])
    table = torch.zeros(np.asarray(candidates).shape, dtype=torch.int32)
    for (q, c, _), n in zip(jobs, counts):
        table[q, c] = int(n)
    if world > 1:
        t = table.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)                              # disjoint ownership: sum == gather of (query, candidate, #matches)
        table = t.cpu()
    tbl = table.numpy().astype(np.int32)
    tbl[np.asarray(candidates) < 0] = -1
    if ctx is not None:
        from . import capi
        best, best_cnt = capi.reloc_pick(tbl)
    else:                                                                      # CPU tests without the library: same rule in numpy
        best = np.full(tbl.shape[0], -1, np.int32); best_cnt = np.zeros(tbl.shape[0], np.int32)
        for q in range(tbl.shape[0]):
            for k in range(tbl.shape[1]):
                if tbl[q, k] > best_cnt[q]:
                    best_cnt[q], best[q] = tbl[q, k], k
    cand = np.asarray(candidates)
    best_kf = np.where(best >= 0, cand[np.arange(len(best)), np.maximum(best, 0)], -1)
    return best_kf, best_cnt, tbl
