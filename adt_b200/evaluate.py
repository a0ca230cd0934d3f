"""Full-catalog evaluation on B200: encoder forward -> K7 scoring with fused per-user top-K -> HR/NDCG/MRR.

Replaces evaluate_loader_full / get_full_sort_score (/root/reference/sasrec/utils.py:710-740, :686-708) and the
sampled-candidate ranker evaluate_loader (:395-428).  The score row [U, I+1] is never materialised: the scoring
kernel keeps a sorted top-K list per user on chip (include/adt_b200.h: adt_score_topk).

Item-sharded mode (one 8xB200 box): every rank holds the whole (small) model, scores the contiguous catalog slice
[lo, hi) of its rank, and the per-shard lists are all-gathered ([world, U, K] pairs) and merged on every rank.
"""
import ctypes
import math
import numpy as np
import torch

from . import _lib as L
from .model import _as_ids


def shard_bounds(n_rows, world, rank):
    """contiguous, 4-row aligned catalog slice of `rank` (rows 0..n_rows-1 of item_emb, row 0 = padding item)."""
    per = (n_rows + world - 1) // world
    per = (per + 3) // 4 * 4
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def merge_topk(scores, ids, K):
    """merge per-shard lists: scores/ids [S, U, K] -> [U, K], order (score desc, id asc); -1 ids are padding.
    Pure torch; runs on CPU (gloo tests) and CUDA alike."""
    S, U, k = scores.shape
    sc = scores.permute(1, 0, 2).reshape(U, S * k)
    idx = ids.permute(1, 0, 2).reshape(U, S * k)
    sc = torch.where(idx < 0, torch.full_like(sc, float("-inf")), sc)
    # lexicographic (score desc, id asc): stable sort by id asc first, then stable by score desc
    o1 = torch.argsort(idx, dim=1, stable=True)
    sc1, id1 = torch.gather(sc, 1, o1), torch.gather(idx, 1, o1)
    o2 = torch.argsort(sc1, dim=1, descending=True, stable=True)
    return torch.gather(sc1, 1, o2)[:, :K], torch.gather(id1, 1, o2)[:, :K]


class CatalogScorer:
    def __init__(self, model, K=10, n_splits=None, process_group=None):
        self.model, self.K = model, int(K)
        self.lib = L.lib()
        self.pg = process_group
        self.world, self.rank = 1, 0
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
            self.rank = torch.distributed.get_rank(process_group)
        self.n_splits = n_splits
        self._buf = {}

    def _buffers(self, U, S, dev):
        key = (U, S)
        if key not in self._buf:
            K = self.K
            self._buf[key] = (torch.empty(S, U, K, device=dev), torch.empty(S, U, K, dtype=torch.int32, device=dev),
                              torch.empty(U, K, device=dev), torch.empty(U, K, dtype=torch.int32, device=dev))
        return self._buf[key]

    @torch.no_grad()
    def topk(self, log_seqs, seen_indptr=None, seen_idx=None):
        """-> (scores [U,K] fp32, ids [U,K] int32), best first, seen items excluded (utils.py:725)."""
        m = self.model
        dev = m.item_emb.weight.device
        seq = _as_ids(log_seqs, dev)
        feats = m.final_feats(seq)
        return self.topk_from_feats(feats, seen_indptr, seen_idx)

    @torch.no_grad()
    def topk_from_feats(self, feats, seen_indptr=None, seen_idx=None):
        m = self.model
        dev = feats.device
        U, H = feats.shape
        E = m.item_emb.weight
        lo, hi = shard_bounds(E.shape[0], self.world, self.rank)
        n_items = hi - lo
        tiles = (U + 63) // 64
        S = self.n_splits or max(1, min(256, (2 * 148 + tiles - 1) // tiles, (n_items + 255) // 256))
        ps, pi, os_, oi = self._buffers(U, S, dev)
        ip = _as_ids(seen_indptr, dev) if seen_indptr is not None else None
        ix = _as_ids(seen_idx, dev) if seen_idx is not None else None
        a = L.fill(L.adt_score_topk_args(), feats=feats, U=U, H=H, item_emb=E[lo:hi], n_items=n_items, item_offset=lo,
                   seen_indptr=ip, seen_idx=ix, K=self.K, n_splits=S, part_scores=ps, part_ids=pi, out_scores=os_, out_ids=oi)
        L.check(self.lib.adt_score_topk(ctypes.byref(a), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "adt_score_topk")
        if self.world == 1:
            return os_, oi
        gs = torch.empty(self.world, U, self.K, device=dev)
        gi = torch.empty(self.world, U, self.K, dtype=torch.int32, device=dev)
        torch.distributed.all_gather_into_tensor(gs, os_, group=self.pg)
        torch.distributed.all_gather_into_tensor(gi, oi, group=self.pg)
        return merge_topk(gs, gi, self.K)


def hit_ndcg_mrr(answers, topk_ids, ks=(5, 10)):
    """get_full_sort_score (utils.py:686-708) for one held-out answer per user.  Host side, exact integers.
    returns dict HIT@k, NDCG@k, MRR (MRR over the returned list, 0 when absent: utils.py:546-569)."""
    ids = np.asarray(topk_ids.cpu() if isinstance(topk_ids, torch.Tensor) else topk_ids)
    ans = np.asarray(answers).reshape(-1, 1)
    hit = ids == ans
    U = ids.shape[0]
    has = hit.any(axis=1)
    first = np.where(has, hit.argmax(axis=1), ids.shape[1])
    res = {}
    for k in ks:
        inside = first < k
        res[f"HIT@{k}"] = float(inside.sum()) / U
        res[f"NDCG@{k}"] = float((1.0 / np.log2(first[inside] + 2.0)).sum()) / U
    res["MRR"] = float((1.0 / (first[has] + 1.0)).sum()) / U
    return res


@torch.no_grad()
def rank_of_first_candidate(model, user_ids, log_seqs, item_idx):
    """evaluate_loader's `(-predict).argsort().argsort()[:, 0]` (utils.py:407-410) without the two sorts:
    rank = #{j : s_j > s_0} (ties broken like a stable argsort: earlier columns first)."""
    logits = model.predict(user_ids, log_seqs, item_idx)
    s0 = logits[:, :1]
    return (logits[:, 1:] > s0).sum(dim=1)


def sampled_metrics(rank, n_candidates, ks=(5, 10)):
    """HR/NDCG@k and the reference's AUC (with its C = 1 + n_candidates quirk, utils.py:424-427)."""
    rank = np.asarray(rank.cpu() if isinstance(rank, torch.Tensor) else rank).astype(np.int64)
    U = len(rank)
    ndcg = {k: float((1.0 / np.log2(rank[rank < k] + 2.0)).sum()) / U for k in ks}
    hr = {k: float((rank < k).sum()) / U for k in ks}
    C = 1 + n_candidates
    auc = float(np.mean((C - (rank + 1)) / (C - 1)))
    return (ndcg, hr), auc
