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
    def __init__(self, model, K=10, n_splits=None, process_group=None, use_tensor_cores=True, tc_min_items=32768):
        self.model, self.K = model, int(K)
        self.use_tc = use_tensor_cores
        self.tc_min_items = tc_min_items   # below this catalog size the exact fp32 kernel is already latency bound and cheaper
        self._tab = None           # (bf16 copy of this rank's catalog shard, max |e|^2 scalar)
        self.fallback_users = 0    # users re-run on the exact fp32 kernel because the bf16 bound was inconclusive
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

    def refresh_table(self):
        """(re)build the bf16 copy of the catalog shard -- call after the item table changed (e.g. once per eval pass)."""
        E = self.model.item_emb.weight
        lo, hi = shard_bounds(E.shape[0], self.world, self.rank)
        shard = E[lo:hi]
        tab = torch.empty(shard.shape, dtype=torch.bfloat16, device=E.device)
        mx = torch.zeros(1, dtype=torch.float32, device=E.device)
        st = ctypes.c_void_p(torch.cuda.current_stream(E.device).cuda_stream)
        L.check(self.lib.adt_to_bf16(L.ptr(shard), L.ptr(tab), ctypes.c_int64(shard.shape[0]), ctypes.c_int32(shard.shape[1]), L.ptr(mx), st),
                "adt_to_bf16")
        self._tab = (tab, mx)

    @torch.no_grad()
    def _tc_launch(self, feats, ip, ix, lo, hi):
        """tensor-core path, device side: tcgen05 candidate generation + exact fp32 re-score.  No host synchronisation
        (capturable); users whose bf16 bound was inconclusive are marked in `flags` for _tc_fixup."""
        dev = feats.device
        U, H = feats.shape
        E = self.model.item_emb.weight
        if self._tab is None:
            self.refresh_table()
        tab, mx = self._tab
        n_items = hi - lo
        K = self.K
        KC = min(64, max(K + 8, 2 * K))
        tiles = (U + 127) // 128
        S = max(1, min(2048 // KC, max(1, 148 // tiles), (n_items + 127) // 128))
        key = ("tc", U, S, KC)
        if key not in self._buf:
            self._buf[key] = (torch.empty(S, U, KC, device=dev), torch.empty(S, U, KC, dtype=torch.int32, device=dev),
                              torch.empty(S, U, device=dev), torch.empty(U, K, device=dev), torch.empty(U, K, dtype=torch.int32, device=dev),
                              torch.empty(U, dtype=torch.int32, device=dev), torch.empty(U, H, dtype=torch.bfloat16, device=dev))
        ps, pi, pt, os_, oi, flags, fb = self._buf[key]
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        L.check(self.lib.adt_to_bf16(L.ptr(feats), L.ptr(fb), ctypes.c_int64(U), ctypes.c_int32(H), None, st), "adt_to_bf16")
        a = L.fill(L.adt_score_topk_tc_args(), feats=feats, feats_bf16=fb, U=U, H=H, item_emb=E[lo:hi], item_emb_bf16=tab, n_items=n_items,
                   item_offset=lo, max_normsq=mx, seen_indptr=ip, seen_idx=ix, K=K, KC=KC, n_splits=S, part_scores=ps, part_ids=pi,
                   part_thr=pt, out_scores=os_, out_ids=oi, flags=flags)
        L.check(self.lib.adt_score_topk_tc(ctypes.byref(a), st), "adt_score_topk_tc")
        return os_, oi, flags

    @torch.no_grad()
    def _tc_fixup(self, feats, ip, ix, lo, hi, os_, oi, flags):
        """host side of the tensor-core path: re-run flagged users on the exact fp32 kernel (reads `flags`: synchronises)."""
        dev = feats.device
        bad = torch.nonzero(flags, as_tuple=False).flatten()
        if bad.numel():
            self.fallback_users += int(bad.numel())
            sub_ip = sub_ix = None
            if ip is not None:
                ipc, ixc = ip.cpu().numpy(), ix.cpu().numpy()
                rows = bad.cpu().numpy()
                lens = ipc[rows + 1] - ipc[rows]
                sub_ip = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
                sub_ix = np.concatenate([ixc[ipc[r]:ipc[r + 1]] for r in rows]).astype(np.int32) if lens.sum() else np.zeros(0, np.int32)
            es, ei = self._topk_exact(feats[bad].contiguous(), _as_ids(sub_ip, dev) if sub_ip is not None else None,
                                      _as_ids(sub_ix, dev) if sub_ix is not None else None, lo, hi)
            os_[bad] = es
            oi[bad] = ei
        return os_, oi

    def _uses_tc(self, H, lo, hi):
        return self.use_tc and H % 64 == 0 and (hi - lo) >= self.tc_min_items

    @torch.no_grad()
    def _merge(self, os_, oi):
        """item-sharded mode: all-gather the per-shard lists and merge them on every rank."""
        if self.world == 1:
            return os_, oi
        U = os_.shape[0]
        gs = torch.empty(self.world, U, self.K, device=os_.device)
        gi = torch.empty(self.world, U, self.K, dtype=torch.int32, device=os_.device)
        torch.distributed.all_gather_into_tensor(gs, os_.contiguous(), group=self.pg)
        torch.distributed.all_gather_into_tensor(gi, oi.contiguous(), group=self.pg)
        return merge_topk(gs, gi, self.K)

    @torch.no_grad()
    def topk_from_feats(self, feats, seen_indptr=None, seen_idx=None):
        dev = feats.device
        U, H = feats.shape
        lo, hi = shard_bounds(self.model.item_emb.weight.shape[0], self.world, self.rank)
        ip = _as_ids(seen_indptr, dev) if seen_indptr is not None else None
        ix = _as_ids(seen_idx, dev) if seen_idx is not None else None
        if self._uses_tc(H, lo, hi):
            os_, oi = self._tc_fixup(feats, ip, ix, lo, hi, *self._tc_launch(feats, ip, ix, lo, hi))
        else:
            os_, oi = self._topk_exact(feats, ip, ix, lo, hi)
        return self._merge(os_, oi)

    @torch.no_grad()
    def _topk_exact(self, feats, ip, ix, lo, hi):
        m = self.model
        dev = feats.device
        U, H = feats.shape
        E = m.item_emb.weight
        n_items = hi - lo
        tiles = (U + 63) // 64
        S = self.n_splits or max(1, min(256, (2 * 148 + tiles - 1) // tiles, (n_items + 255) // 256))
        ps, pi, os_, oi = self._buffers(U, S, dev)
        a = L.fill(L.adt_score_topk_args(), feats=feats, U=U, H=H, item_emb=E[lo:hi], n_items=n_items, item_offset=lo,
                   seen_indptr=ip, seen_idx=ix, K=self.K, n_splits=S, part_scores=ps, part_ids=pi, out_scores=os_, out_ids=oi)
        L.check(self.lib.adt_score_topk(ctypes.byref(a), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "adt_score_topk")
        return os_, oi


class GraphedScorer:
    """Fixed-shape evaluation batch (U users x L positions, seen-set CSR of at most `max_seen` ids) replayed as ONE CUDA
    graph: encoder forward + catalog scoring + fused top-K.  Per batch the host only copies the ids into the static
    buffers and replays; the exact-kernel fix-up of the tensor-core path and the item-shard merge stay outside the graph
    (one reads a device flag on the host, the other is a collective)."""

    def __init__(self, scorer, U, L_, max_seen):
        self.sc = scorer
        m = scorer.model
        dev = m.item_emb.weight.device
        self.U, self.L = int(U), int(L_)
        self.seq = torch.zeros(U, L_, dtype=torch.int32, device=dev)
        self.ip = torch.zeros(U + 1, dtype=torch.int32, device=dev)
        self.ix = torch.zeros(max(1, int(max_seen)), dtype=torch.int32, device=dev)
        self.lo, self.hi = shard_bounds(m.item_emb.weight.shape[0], scorer.world, scorer.rank)
        self.tc = scorer._uses_tc(m.hidden, self.lo, self.hi)
        if self.tc:
            scorer.refresh_table()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):          # warm-up outside capture (lazy allocations, cudaFuncSetAttribute)
            for _ in range(2):
                self._device_part()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._device_part()

    @torch.no_grad()
    def _device_part(self):
        self.feats = self.sc.model.final_feats(self.seq)
        if self.tc:
            return self.sc._tc_launch(self.feats, self.ip, self.ix, self.lo, self.hi)
        return self.sc._topk_exact(self.feats, self.ip, self.ix, self.lo, self.hi)

    @torch.no_grad()
    def topk(self, log_seqs, seen_indptr, seen_idx):
        """same contract as CatalogScorer.topk for a [U, L] batch; the returned tensors are overwritten by the next call."""
        def src(a):   # host numpy or torch tensor (any device) -> int32 tensor that copy_ can read
            return a.to(torch.int32) if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32))
        seq, ip, ix = src(log_seqs), src(seen_indptr), src(seen_idx)
        if tuple(seq.shape) != (self.U, self.L) or ip.numel() != self.U + 1 or ix.numel() > self.ix.numel():
            raise ValueError(f"GraphedScorer was built for [{self.U},{self.L}] batches with at most {self.ix.numel()} seen ids")
        self.seq.copy_(seq, non_blocking=True)
        self.ip.copy_(ip, non_blocking=True)
        self.ix[:ix.numel()].copy_(ix, non_blocking=True)
        self.graph.replay()
        if self.tc:
            os_, oi = self.sc._tc_fixup(self.feats, self.ip, self.ix, self.lo, self.hi, *self.out)
        else:
            os_, oi = self.out
        return self.sc._merge(os_, oi)


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
