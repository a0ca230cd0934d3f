"""Device-side batch assembly and sampling -- the GPU replacement of the reference's DataLoader workers:
WarpDataset.sample_data + random_neq (/root/reference/sasrec/utils.py:288-307, :73-77) for training batches and
EvalDataset.sample_data + PopularSampler.get_negative_samples (:162-191, :57-69) for the sampled-candidate evaluation.

The user histories (train / valid / test dicts of the reference's data_partition, utils.py:124-160) are uploaded ONCE as CSR; every
batch is then one kernel launch (adt_assemble_train_batch / adt_assemble_eval_batch) writing the int32 [B, L] id tensors the model
consumes -- no host work, no H2D copy per step.  Randomness is counter based: sample (user, position, draw) of epoch e is a pure
function of (seed, e, user, position, draw), independent of batch composition, batch size and rank; oracle/sampler_oracle.py
restates it on the host bit for bit.
"""
import ctypes
import numpy as np
import torch

from . import _lib as L


def _csr(rows, n_users, sort=False):
    """dict / list of per-user id lists (user ids 1..n_users) -> (indptr [n_users + 2], items) indexed by user id"""
    indptr = np.zeros(n_users + 2, np.int64)
    chunks = []
    for u in range(1, n_users + 1):
        r = np.asarray(rows.get(u, []) if isinstance(rows, dict) else rows[u], dtype=np.int32)
        if sort:
            r = np.unique(r)
        chunks.append(r)
        indptr[u + 1] = indptr[u] + len(r)
    items = np.concatenate(chunks).astype(np.int32) if chunks and indptr[-1] else np.zeros(1, np.int32)
    return indptr.astype(np.int32), items


def alias_table(p):
    """Vose's alias method for a probability vector p (float64) -> (prob float32 [n], alias int32 [n])"""
    p = np.asarray(p, dtype=np.float64)
    n = len(p)
    scaled = p * n / p.sum()
    prob, alias = np.ones(n, np.float64), np.arange(n, dtype=np.int32)
    small = [i for i in range(n) if scaled[i] < 1.0]
    large = [i for i in range(n) if scaled[i] >= 1.0]
    while small and large:
        s, l = small.pop(), large.pop()
        prob[s], alias[s] = scaled[s], l
        scaled[l] = scaled[l] - (1.0 - scaled[s])
        (small if scaled[l] < 1.0 else large).append(l)
    return prob.astype(np.float32), alias


class DeviceSampler:
    def __init__(self, user_train, user_valid, user_test, usernum, itemnum, maxlen, device="cuda", seed=23):
        self.usernum, self.itemnum, self.L, self.seed = int(usernum), int(itemnum), int(maxlen), int(seed)
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise L.AdtError("adt_b200.DeviceSampler assembles batches on a CUDA device (no CPU fallback)")
        self.lib = L.lib()
        ip, it = _csr(user_train, usernum)
        # seen sets of the candidate sampler (utils.py:60-63): train + valid (val mode), + test (test mode)
        seen_v = {u: list(user_train.get(u, [])) + list(user_valid.get(u, [])) for u in range(1, usernum + 1)}
        seen_t = {u: seen_v[u] + list(user_test.get(u, [])) for u in range(1, usernum + 1)}
        vip, vit = _csr(seen_v, usernum, sort=True)
        tip, tit = _csr(seen_t, usernum, sort=True)
        # popularity over train + valid + test, indexed 0..itemnum-1 (utils.py:30-39, quirk B8)
        pop = np.zeros(itemnum, np.float64)
        for d in (user_train, user_valid, user_test):
            for u in range(1, usernum + 1):
                for i in d.get(u, []):
                    if i < itemnum:
                        pop[i] += 1.0
        self.popular_p = pop / pop.sum()
        ap, ai = alias_table(self.popular_p)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        self.hist_indptr, self.hist_items = t(ip), t(it)
        self.seen = {"val": (t(vip), t(vit)), "test": (t(tip), t(tit))}
        self.alias_prob, self.alias_idx = t(ap), t(ai)
        self.valid_first = t(np.array([0] + [(user_valid.get(u) or [0])[0] for u in range(1, usernum + 1)], np.int32))
        self.test_first = t(np.array([0] + [(user_test.get(u) or [0])[0] for u in range(1, usernum + 1)], np.int32))
        self._host = dict(train=user_train, valid=user_valid, test=user_test)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def train_batch(self, users, epoch=0, out=None):
        """users: int array / tensor [B] of user ids -> (seq, dec, pos, neg) int32 device tensors [B, L]
        (`out` = four preallocated tensors, e.g. the static buffers of a captured training graph)"""
        u = users.to(self.dev, torch.int32) if isinstance(users, torch.Tensor) else torch.from_numpy(np.asarray(users, np.int32)).to(self.dev)
        B = u.numel()
        if out is None:
            out = [torch.empty(B, self.L, dtype=torch.int32, device=self.dev) for _ in range(4)]
        a = L.fill(L.adt_train_batch_args(), users=u, hist_indptr=self.hist_indptr, hist_items=self.hist_items,
                   hist_sorted=self._sorted_aligned(), seq=out[0], dec=out[1], pos=out[2], neg=out[3], B=B, L=self.L, itemnum=self.itemnum,
                   seed=self.seed, epoch=int(epoch))
        L.check(self.lib.adt_assemble_train_batch(ctypes.byref(a), self._stream()), "adt_assemble_train_batch")
        return tuple(out)

    def _sorted_aligned(self):
        """per-user history sorted ascending WITH duplicates kept, so that it shares hist_indptr with the ordered history"""
        if not hasattr(self, "_sorted_dup"):
            ip = self.hist_indptr.cpu().numpy()
            it = self.hist_items.cpu().numpy().copy()
            for u in range(1, self.usernum + 1):
                it[ip[u]:ip[u + 1]].sort()
            self._sorted_dup = torch.from_numpy(it).to(self.dev)
        return self._sorted_dup

    def eval_batch(self, users, mode="val", n_candidates=100, epoch=0):
        """-> (seq [U, L], item_idx [U, 1 + n_candidates]) : column 0 is the held-out item (valid / test), the rest popularity-sampled
        negatives outside the user's seen set (utils.py:162-191, :57-69)"""
        u = users.to(self.dev, torch.int32) if isinstance(users, torch.Tensor) else torch.from_numpy(np.asarray(users, np.int32)).to(self.dev)
        U = u.numel()
        seq = torch.empty(U, self.L, dtype=torch.int32, device=self.dev)
        idx = torch.empty(U, 1 + n_candidates, dtype=torch.int32, device=self.dev)
        sip, sit = self.seen[mode]
        ul = u.long()
        answers = (self.valid_first if mode == "val" else self.test_first)[ul].contiguous()
        last = self.valid_first[ul].contiguous() if mode == "test" else None
        a = L.fill(L.adt_eval_batch_args(), users=u, hist_indptr=self.hist_indptr, hist_items=self.hist_items, seen_indptr=sip, seen_sorted=sit,
                   last_item=last, answers=answers, alias_prob=self.alias_prob, alias_idx=self.alias_idx, seq=seq, item_idx=idx, U=U, L=self.L,
                   itemnum=self.itemnum, n_candidates=int(n_candidates), seed=self.seed, epoch=int(epoch))
        L.check(self.lib.adt_assemble_eval_batch(ctypes.byref(a), self._stream()), "adt_assemble_eval_batch")
        return seq, idx


class ClozeSampler:
    """Bert4Rec-ADT training instances on the device (BertTrainDataset._generate_data / sample_data / _mask_last,
    /root/reference/bert4rec/datasets/dataset.py:70-158).  The reference materialises dupe_factor masked copies of every window of
    every user as Python lists of tensors before training; here only the instance TABLE (user, window start, window length, copy
    index) is built once (integer arithmetic on the host, same window enumeration), and the masked token / decoder / label tensors of a
    batch are generated by one kernel launch from the resident histories, with fresh masks every epoch if wanted."""

    def __init__(self, user_train, usernum, itemnum, maxlen, mask_prob, dupe_factor=10, prop_sliding_window=0.1, device="cuda", seed=23):
        self.usernum, self.itemnum, self.L = int(usernum), int(itemnum), int(maxlen)
        self.mask_prob, self.seed, self.mask_token = float(mask_prob), int(seed), int(itemnum) + 1
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise L.AdtError("adt_b200.ClozeSampler generates batches on a CUDA device (no CPU fallback)")
        self.lib = L.lib()
        ip, it = _csr(user_train, usernum)
        self.table = self.instance_table(user_train, usernum, maxlen, dupe_factor, prop_sliding_window)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        self.hist_indptr, self.hist_items = t(ip), t(it)
        self.tab_dev = t(self.table)

    @staticmethod
    def instance_table(user_train, usernum, maxlen, dupe_factor, prop_sliding_window):
        """dataset.py:70-98 -> int32 [n_instances, 4] = (user, window start, window length, copy index | -1 for mask-last)"""
        rows = []
        for user in range(1, usernum + 1):
            seqs = user_train.get(user, []) if isinstance(user_train, dict) else user_train[user]
            n = len(seqs)
            if n < 1:
                continue
            if n <= maxlen:
                rows += [(user, 0, n, d) for d in range(dupe_factor)]
            else:
                step = int(prop_sliding_window * maxlen) if prop_sliding_window != -1 else maxlen
                beg = list(range(n - maxlen, 0, -step)) + [0]
                for i in beg[::-1]:
                    rows += [(user, i, min(maxlen, n - i), d) for d in range(dupe_factor)]
            ln = min(n, maxlen)
            rows.append((user, n - ln, ln, -1))
        return np.asarray(rows, np.int32).reshape(-1, 4)

    def __len__(self):
        return len(self.table)

    def batch(self, idx, epoch=0):
        """idx: int array / tensor of instance indices -> (tokens, dec_tokens, labels) int32 device tensors [B, L]"""
        ix = idx.to(self.dev, torch.long) if isinstance(idx, torch.Tensor) else torch.from_numpy(np.asarray(idx, np.int64)).to(self.dev)
        rows = self.tab_dev[ix].t().contiguous()
        B = ix.numel()
        out = [torch.empty(B, self.L, dtype=torch.int32, device=self.dev) for _ in range(3)]
        a = L.fill(L.adt_cloze_batch_args(), users=rows[0], win_start=rows[1], win_len=rows[2], dup=rows[3], hist_indptr=self.hist_indptr,
                   hist_items=self.hist_items, tokens=out[0], dec_tokens=out[1], labels=out[2], B=B, L=self.L, itemnum=self.itemnum,
                   mask_token=self.mask_token, mask_prob=self.mask_prob, seed=self.seed, epoch=int(epoch))
        L.check(self.lib.adt_cloze_batch(ctypes.byref(a), ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)), "adt_cloze_batch")
        return tuple(out)
