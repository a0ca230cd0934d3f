"""Synthetic workloads of the shapes BASELINE.json names.  Host-side numpy only.

Batches follow the reference's WarpDataset.sample_data layout (/root/reference/sasrec/utils.py:288-307):
right-aligned, left-padded with 0; dec = seq shifted right by one; pos = next item; neg = uniform item not in the
user's history."""
import numpy as np

CONFIGS = {
    # name: (items, maxlen, hidden, heads, layers, batch, dropout, dataset key for get_lambdas, weight_decay, mean_len)
    "C2": dict(items=12101, L=50, H=64, nh=2, nl=2, B=256, p=0.5, dataset="beauty", wd=1e-4, geo=1.0 / 9, lo=3, add=2),
    "C1": dict(items=3416, L=200, H=256, nh=2, nl=2, B=256, p=0.5, dataset="ml-1m", wd=1e-3, geo=1.0 / 160, lo=16, add=15),
    "beauty": dict(items=54542, L=50, H=256, nh=2, nl=2, B=256, p=0.5, dataset="beauty", wd=1e-4, geo=1.0 / 9, lo=3, add=2),
}


def zipf_items(rng, n, items):
    """Zipf(1.0) over 1..items by inverse-CDF on the harmonic weights."""
    w = 1.0 / np.arange(1, items + 1, dtype=np.float64)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    return (np.searchsorted(cdf, rng.random(n)) + 1).astype(np.int64)


def make_batch(rng, cfg, B=None):
    B = B or cfg["B"]
    L, I = cfg["L"], cfg["items"]
    seq = np.zeros((B, L), np.int32)
    dec = np.zeros((B, L), np.int32)
    pos = np.zeros((B, L), np.int32)
    neg = np.zeros((B, L), np.int32)
    lens = np.clip(rng.geometric(cfg["geo"], size=B) + cfg["add"], cfg["lo"], L + 1)
    for b in range(B):
        n = int(lens[b])
        items = zipf_items(rng, n, I)
        hist, nxt = items[:-1][-L:], items[1:][-L:]
        m = len(hist)
        seq[b, L - m:] = hist
        pos[b, L - m:] = nxt
        ng = rng.integers(1, I + 1, size=m)
        bad = np.isin(ng, items)
        while bad.any():   # random_neq (utils.py:73-77)
            ng[bad] = rng.integers(1, I + 1, size=int(bad.sum()))
            bad = np.isin(ng, items)
        neg[b, L - m:] = ng
        dec[b, L - m + 1:] = hist[:-1]
    return seq, dec, pos, neg


def make_eval_batch(rng, cfg, U):
    """sequences + held-out answer + seen-set CSR (sorted ids) for full-catalog evaluation."""
    L, I = cfg["L"], cfg["items"]
    seq = np.zeros((U, L), np.int32)
    answers = np.zeros((U,), np.int32)
    indptr = np.zeros(U + 1, np.int32)
    seen = []
    lens = np.clip(rng.geometric(cfg["geo"], size=U) + cfg["add"], cfg["lo"], L + 1)
    for u in range(U):
        n = int(lens[u])
        items = zipf_items(rng, n, I)
        hist = items[:-1][-L:]
        seq[u, L - len(hist):] = hist
        s = np.unique(items[:-1])
        ans = int(items[-1])
        if ans in s:  # keep the answer scoreable
            s = s[s != ans]
        answers[u] = ans
        seen.append(s.astype(np.int32))
        indptr[u + 1] = indptr[u] + len(s)
    return seq, answers, indptr, (np.concatenate(seen) if seen else np.zeros(0, np.int32))


def make_histories(rng, cfg, n_users):
    """per-user interaction histories of the configured shape, split like data_partition (/root/reference/sasrec/utils.py:124-160):
    the last item is the test item, the one before the validation item, the rest is training history -> (train, valid, test) dicts"""
    I = cfg["items"]
    lens = np.clip(rng.geometric(cfg["geo"], size=n_users) + cfg["add"] + 2, cfg["lo"] + 2, 4 * cfg["L"])
    train, valid, test = {}, {}, {}
    for u in range(1, n_users + 1):
        items = [int(x) for x in zipf_items(rng, int(lens[u - 1]), I)]
        train[u], valid[u], test[u] = items[:-2], [items[-2]], [items[-1]]
    return train, valid, test
