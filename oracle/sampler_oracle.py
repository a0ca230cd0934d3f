"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path.

Host restatement of the device-side batch assembly (adt_b200/csrc/sampler.cu) following the reference's samplers:
  train_sample()   WarpDataset.sample_data + random_neq      /root/reference/sasrec/utils.py:288-307, :73-77
  eval_sample()    EvalDataset.sample_data (:162-191) + PopularSampler.get_negative_samples (:57-69)
with the reference's np.random draws replaced by the same Philox4x32-10 counters the kernels use (oracle/philox.py), so the
comparison is bit exact.  The layout logic (right alignment, dec = seq shifted right, pos = next item, the validation item closing
the test sequence, popularity over ids 0..itemnum-1 / quirk B8, seen-set per mode) is the reference's.
"parity": pinned against the reference's own sample_data for everything that is not random (tests/test_sampler.py runs the
unmodified WarpDataset / EvalDataset from /root/reference when it is present and compares seq / dec / pos / answers / seen exclusion).
"""
import numpy as np

from . import philox


def _bounded(r, n):
    return int((int(r) * int(n)) >> 32)


def train_sample(hist, user, L, itemnum, seed, epoch):
    """-> seq, dec, pos, neg  (int32 [L]) for one user"""
    seq, dec, pos, neg = (np.zeros(L, np.int32) for _ in range(4))
    h = list(hist)
    n = len(h)
    ts = set(h)
    for idx in range(L):
        j = L - 1 - idx
        if n >= 2 and j < n - 1:
            seq[idx] = h[n - 2 - j]
            pos[idx] = h[n - 1 - j]
            if pos[idx] != 0:
                k = 0
                while True:
                    r = philox.philox4x32_10(np.uint32([user]), np.uint32(idx), np.uint32(k >> 2), np.uint32(epoch),
                                             np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF))
                    w = int(r[k & 3][0])
                    t = 1 + _bounded(w, itemnum)
                    if t not in ts or k > 4096:
                        break
                    k += 1
                neg[idx] = t
    dec[1:] = seq[:-1]
    return seq, dec, pos, neg


def eval_sample(hist, seen, answer, last_item, user, L, itemnum, S, alias_prob, alias_idx, seed, epoch):
    """-> seq [L], item_idx [1 + S]"""
    seq = np.zeros(L, np.int32)
    h = list(hist)
    if last_item:
        h = h + [last_item]
    tail = h[-L:]
    if tail:
        seq[L - len(tail):] = tail
    out = [int(answer)]
    seen = set(int(x) for x in seen)
    k = 0
    while len(out) < 1 + S and k < 4096 * 32:
        r = philox.philox4x32_10(np.uint32([user]), np.uint32(k), np.uint32(0x5eed), np.uint32(epoch),
                                 np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF))
        col = _bounded(int(r[0][0]), itemnum)
        coin = np.float32(int(r[1][0]) >> 8) * np.float32(1.0 / 16777216.0)
        cand = col if coin < alias_prob[col] else int(alias_idx[col])
        if cand not in seen and cand not in out[1:]:
            out.append(cand)
        k += 1
    return seq, np.asarray(out, np.int32)


def cloze_sample(hist, user, start, length, dup, L, itemnum, mask_prob, seed, epoch):
    """BertTrainDataset.sample_data / _mask_last (/root/reference/bert4rec/datasets/dataset.py:99-158) for one instance, with the
    reference's random.Random draws replaced by the kernel's Philox counters -> tokens, dec_tokens, labels (int32 [L])"""
    mask_token = itemnum + 1
    tokens, dec, labels = (np.zeros(L, np.int32) for _ in range(3))
    win = list(hist[start:start + length])
    for j, s in enumerate(win):
        idx = L - length + j
        tok = dtok = s
        lab = 0
        if dup < 0:
            if j == length - 1:
                tok = dtok = mask_token
                lab = s
        else:
            r = philox.philox4x32_10(np.uint32([user]), np.uint32(start + j), np.uint32(dup), np.uint32(epoch),
                                     np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF))
            p = np.float32(int(r[0][0]) >> 8) * np.float32(1.0 / 16777216.0)
            if p < np.float32(mask_prob):
                p = np.float32(p / np.float32(mask_prob))
                if p < np.float32(0.8):
                    tok = mask_token
                elif p < np.float32(0.9):
                    tok = 1 + _bounded(int(r[1][0]), itemnum)
                dtok = tok
                lab = s
            if j == length - 1:
                dtok = mask_token
        tokens[idx], dec[idx], labels[idx] = tok, dtok, lab
    return tokens, dec, labels
