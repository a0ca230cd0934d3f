"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/bert_*.npz by running the UNMODIFIED reference
(/root/reference/bert4rec/model/bert.py + the loss/optimiser lines of bert4rec/trainer.py:100-132) on CPU here.

    python -m oracle.make_golden_bert
"""
import os
import sys
import types
import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/bert4rec"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
from . import philox  # noqa: E402


class Inj:
    """k-th active F.dropout call -> Philox site k; every site of the Bert path has the natural [B,L,H] / [B,nh,L,L] layout."""

    def __init__(self, seed, step):
        self.seed, self.step, self.k = seed, step, 0

    def __call__(self, x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        site = self.k
        self.k += 1
        if x.dim() == 4:     # attention weights [B,nh,L,L]
            B, nh, Lq, Lk = x.shape
            keep = philox.keep_mask_attn(B * nh * Lq, Lk, p, self.seed, self.step, site).reshape(tuple(x.shape))
        else:
            keep = philox.keep_mask(x.numel(), p, self.seed, self.step, site).reshape(tuple(x.shape))
        return x * torch.from_numpy(keep).to(x.dtype) * torch.tensor(1.0 / (1.0 - p), dtype=torch.float32).to(x.dtype)


def cloze_batch(rng, B, L, I, mask_token, mask_prob=0.3):
    """right-aligned sequences; ~mask_prob of the real positions replaced by the mask token with label = original item
    (bert4rec/datasets/dataset.py:131-149, simplified to the 'replace by [MASK]' branch); dec = unmasked sequence."""
    seq = np.zeros((B, L), np.int64); dec = np.zeros((B, L), np.int64); lab = np.zeros((B, L), np.int64)
    for b in range(B):
        n = int(rng.integers(2, L + 1))
        items = rng.integers(1, I + 1, size=n)
        dec[b, L - n:] = items
        m = rng.random(n) < mask_prob
        m[-1] = True
        seq[b, L - n:] = np.where(m, mask_token, items)
        lab[b, L - n:] = np.where(m, items, 0)
    return seq, dec, lab


def run(name, B, L, H, nh, nl, inner, I, p, pa, lambda1, lambda2, wd, seed=23):
    sys.path.insert(0, REF)
    from model import bert as refbert  # noqa
    torch.manual_seed(seed)
    args = types.SimpleNamespace(device="cpu", num_heads=nh, maxlen=L, num_layers=nl, hidden_units=H, dropout=p, attention_dropout=pa,
                                 inner_units=inner, type_vocab_size=2)
    m = refbert.BertModel(100, I, args)
    g = torch.Generator().manual_seed(seed + 1)
    for _, prm in m.named_parameters():     # trainer.py:29-37 style init, plus noise on 1-D params so they are exercised
        if prm.dim() >= 2:
            prm.data.normal_(mean=0.01, std=0.1, generator=g)
        else:
            prm.data.add_(0.1 * torch.randn(prm.shape, generator=g))
    rng = np.random.default_rng(seed)
    seq, dec, lab = cloze_batch(rng, B, L, I, mask_token=I + 1)
    sd0 = {k: v.detach().clone().numpy() for k, v in m.state_dict().items()}
    t = lambda a: torch.from_numpy(a)
    pos_ids = torch.arange(L).repeat(B, 1)
    sent = torch.zeros(B, L, dtype=torch.long)
    inj = Inj(1234, 7)
    orig = F.dropout
    F.dropout = inj
    try:
        m.train()
        logits, enc_in, dec_out, ind = m(t(seq), t(dec), pos_ids, sent, pos_ids, sent)
    finally:
        F.dropout = orig
    ce = torch.nn.CrossEntropyLoss(ignore_index=0)
    loss = ce(logits.view(-1, logits.size(-1)), t(lab).view(-1))
    for i in range(len(enc_in)):
        if lambda1[i] != 0:
            loss = loss + lambda1[i] * F.mse_loss(enc_in[i], dec_out[i])
    label = torch.tile(torch.arange(nh), [B * L, 1])
    for l in range(len(ind)):
        if lambda2[l] != 0:
            loss = loss + lambda2[l] * F.nll_loss(ind[l].view(B * L, nh, nh), label)
    opt = torch.optim.Adam(m.parameters(), lr=0.001, betas=(0.9, 0.999), weight_decay=wd)
    opt.zero_grad()
    loss.backward()
    gnorm = torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
    grads = {k: prm.grad.detach().clone().numpy() for k, prm in m.named_parameters() if prm.grad is not None}
    opt.step()
    sd1 = {k: v.detach().clone().numpy() for k, v in m.state_dict().items()}
    m.eval()
    with torch.no_grad():
        cand = rng.integers(1, I + 1, size=(B, 11))
        pred = m.predict(None, t(seq), pos_ids, sent, t(cand)).numpy()
    out = {"seq": seq, "dec": dec, "labels": lab, "cand": cand, "cfg": np.array([B, L, H, nh, nl, I, inner]), "p": np.array(p),
           "pa": np.array(pa), "drop_seed": np.array(1234), "drop_step": np.array(7), "lambda1": np.array(lambda1),
           "lambda2": np.array(lambda2), "wd": np.array(wd), "logits": logits.detach().numpy(), "loss": loss.detach().numpy(),
           "gnorm": gnorm.numpy(), "pred": pred, "sites": np.array(inj.k)}
    for i in range(nl):
        out[f"enc_in{i}"] = enc_in[i].detach().numpy()
        out[f"dec_out{i}"] = dec_out[i].detach().numpy()
        out[f"ind{i}"] = ind[i].detach().numpy()
    for k, v in sd0.items():
        out["sd0/" + k] = v
    for k, v in sd1.items():
        out["sd1/" + k] = v
    for k, v in grads.items():
        out["grad/" + k] = v
    np.savez_compressed(os.path.join(OUT, f"bert_{name}.npz"), **out)
    print("bert", name, "loss", float(loss.detach()), "gnorm", float(gnorm), "sites", inj.k, "grads", len(grads), "/", len(sd0))


if __name__ == "__main__":
    run("tiny_p0", B=3, L=8, H=16, nh=2, nl=2, inner=32, I=30, p=0.0, pa=0.0, lambda1=[0.05, 0.1], lambda2=[0.02, 0.07], wd=1e-4)
    run("tiny_p3", B=3, L=8, H=16, nh=2, nl=2, inner=32, I=30, p=0.3, pa=0.2, lambda1=[0.05, 0.1], lambda2=[0.02, 0.07], wd=1e-4)
    run("mid_p5", B=4, L=20, H=64, nh=4, nl=1, inner=256, I=150, p=0.5, pa=0.5, lambda1=[0.005], lambda2=[0.0019], wd=1e-4)
