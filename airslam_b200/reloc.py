"""Batched relocalization matching (BASELINE.json config 5; the reference does it sequentially on one GPU:
MapUser::Relocalization, src/map_user.cc:363-376 -- up to GoodCandidateNum = 3 MatchingPoints calls per query).

Keyframe features are sharded over ranks by keyframe id (block partition, airslam_b200.dist.shard_range); query features are
detected replica-parallel and ALL-GATHERED once (the only collective on the path, NCCL over NVLink on the GPU box), then every rank
LightGlue-matches all queries against the candidates it owns, in batches through airfe_match_batch; a second tiny all-gather of
(query, candidate, #matches) lets every rank pick the winners.  No host code here computes anything but bookkeeping."""
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


def match_jobs(ctx, matcher, jobs, query_feats, keyframe_feats_local, kf_begin):
    """Runs the jobs in batches of ctx.cfg.max_batch pairs.  Returns int array [len(jobs)] of match counts and the match lists."""
    counts, matches = [], []
    B = ctx.cfg.max_batch
    for i in range(0, len(jobs), B):
        chunk = jobs[i:i + B]
        res = ctx.match_batch(matcher, [query_feats[q] for q, _, _ in chunk], [keyframe_feats_local[kf - kf_begin] for _, _, kf in chunk])
        for idx, sc in res:
            counts.append(len(idx))
            matches.append((idx, sc))
    return np.array(counts, dtype=np.int64), matches


def relocalize(ctx, matcher, query_feats_local, keyframe_feats_local, candidates, n_keyframes, rank=0, world=1, match_fn=None):
    """query_feats_local: list of [259, N] arrays detected on this rank (queries are block-partitioned over ranks like keyframes).
    Returns (best_candidate [Q], best_count [Q], table [Q, C] of match counts) -- identical on every rank."""
    import torch
    import torch.distributed as dist
    cap = max([f.shape[1] for f in query_feats_local] + [1])
    ql = len(query_feats_local)
    feat = torch.zeros(ql, cap, 259)
    cnt = torch.zeros(ql, dtype=torch.int32)
    for i, f in enumerate(query_feats_local):
        feat[i, :f.shape[1]] = torch.from_numpy(np.ascontiguousarray(f.T))
        cnt[i] = f.shape[1]
    if world > 1:
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        mx = int(D.max_over_ranks(cap, device=dev if dev == "cuda" else None))
        mq = int(D.max_over_ranks(ql, device=dev if dev == "cuda" else None))
        fpad = torch.zeros(mq, mx, 259); fpad[:ql, :cap] = feat
        cpad = torch.full((mq,), -1, dtype=torch.int32); cpad[:ql] = cnt
        fa, ca = D.all_gather_features(fpad.to(dev), cpad.to(dev))          # the collective: query features to every rank
        fa, ca = fa.cpu(), ca.cpu()
        keep = ca >= 0
        feat, cnt = fa[keep], ca[keep]
    queries = [feat[i, :int(cnt[i])].numpy().T.copy() for i in range(feat.shape[0])]
    kf_begin, _ = D.shard_range(n_keyframes, rank, world)
    jobs = local_jobs(candidates, n_keyframes, rank, world)
    fn = match_fn or (lambda jb: match_jobs(ctx, matcher, jb, queries, keyframe_feats_local, kf_begin)[0])
    counts = fn(jobs)
    table = torch.zeros(candidates.shape, dtype=torch.int64)
    for (q, c, _), n in zip(jobs, counts):
        table[q, c] = int(n)
    if world > 1:
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = table.to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)                              # disjoint ownership: sum == gather of (query, candidate, #matches)
        table = t.cpu()
    best = table.argmax(dim=1)
    return np.asarray(candidates)[np.arange(len(best)), best.numpy()], table.max(dim=1).values.numpy(), table.numpy()
