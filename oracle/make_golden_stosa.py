"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/stosa_*.npz by running the UNMODIFIED reference
(/root/reference/stosa/models.py `DisenDistSAModel`, stosa/trainer.py `DistSAModelTrainer.bpr_optimization`,
`.iteration(train=True)` for the optimiser step and `.dist_predict_full` for the evaluation scores) on CPU here.

    python -m oracle.make_golden_stosa
"""
import io
import os
import sys
import types
import contextlib
import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/stosa"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
from .make_golden_bert import Inj  # noqa: E402  (k-th active F.dropout call -> Philox site k)


def batch(rng, B, L, I):
    """right-aligned sequences; dec = seq shifted right; pos = next item; neg = random item not in the sequence."""
    seq = np.zeros((B, L), np.int64); pos = np.zeros((B, L), np.int64); neg = np.zeros((B, L), np.int64)
    for b in range(B):
        n = int(rng.integers(2, L + 2))
        items = rng.integers(1, I + 1, size=n)
        hist, nxt = items[:-1][-L:], items[1:][-L:]
        seq[b, L - len(hist):] = hist
        pos[b, L - len(nxt):] = nxt
        neg[b, L - len(nxt):] = rng.integers(1, I + 1, size=len(nxt))
    dec = np.zeros_like(seq)
    dec[:, 1:] = seq[:, :-1]
    return seq, dec, pos, neg


def run(name, B, L, H, nh, nl, I, p, pa, lambda1, lambda2, pvn=0.005, wd=0.0, seed=23):
    sys.path.insert(0, REF)
    import models as refmodels   # noqa
    import trainer as reftrainer  # noqa
    torch.manual_seed(seed)
    args = types.SimpleNamespace(item_size=I + 2, num_users=B, maxlen=L, hidden_units=H, num_heads=nh, num_layers=nl, dropout=p,
                                 attention_dropout=pa, distance_metric="wasserstein", kernel_param=1.0, initializer_range=0.02,
                                 cuda_condition=False, no_cuda=True, pvn_weight=pvn, lr=0.001, weight_decay=wd, adam_beta1=0.9,
                                 adam_beta2=0.999, log_freq=1000)
    m = refmodels.DisenDistSAModel(args)
    g = torch.Generator().manual_seed(seed + 1)
    for _, prm in m.named_parameters():      # wider than initializer_range=0.02 so that every term is exercised
        if prm.dim() >= 2:
            prm.data.normal_(mean=0.01, std=0.15, generator=g)
        else:
            prm.data.add_(0.1 * torch.randn(prm.shape, generator=g))
    rng = np.random.default_rng(seed)
    seq, dec, pos, neg = batch(rng, B, L, I)
    t = lambda a: torch.from_numpy(a)
    users = torch.arange(B)
    sd0 = {k: v.detach().clone().numpy() for k, v in m.state_dict().items()}
    tr = reftrainer.DistSAModelTrainer(m, None, None, None, args, lambda1, lambda2)

    # (1) forward + loss with the reference's own methods; the loss lines are trainer.py:517-533 restated
    inj = Inj(1234, 7)
    orig = F.dropout
    F.dropout = inj
    try:
        m.train()
        om, oc, _, _, enc_in, recs, dec_out = m.finetune(t(seq), t(dec), users)
    finally:
        F.dropout = orig
    bpr, auc, pvn_loss = tr.bpr_optimization(om, oc, t(pos), t(neg))
    loss = bpr
    dec_rev = list(reversed(dec_out))
    for l in range(nl):
        loss = loss + lambda1[l] * F.mse_loss(enc_in[l][0], dec_rev[l][0]) + lambda1[l] * F.mse_loss(enc_in[l][1], dec_rev[l][1])
    label = torch.tile(torch.arange(nh), [B * L, 1])
    for l in range(nl):
        loss = loss + lambda2[l] * F.nll_loss(recs[l][0].view(B * L, nh, nh), label) + lambda2[l] * F.nll_loss(recs[l][1].view(B * L, nh, nh), label)
    loss = loss + pvn_loss
    m.zero_grad()
    loss.backward()
    grads = {k: prm.grad.detach().clone().numpy() for k, prm in m.named_parameters() if prm.grad is not None}

    # (2) the reference's own train iteration on the same batch / same masks -> parameters after one Adam step
    inj2 = Inj(1234, 7)
    F.dropout = inj2
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            tr.iteration(0, [(users, t(seq), t(dec), t(pos), t(neg), t(pos[:, -1:]))], train=True)
    finally:
        F.dropout = orig
    for k, prm in m.named_parameters():     # the restated loss lines above must reproduce the reference's gradients
        if prm.grad is not None:
            assert np.allclose(prm.grad.numpy(), grads[k], rtol=1e-5, atol=1e-7), k
    sd1 = {k: v.detach().clone().numpy() for k, v in m.state_dict().items()}

    # (3) evaluation scores of the updated model
    m.eval()
    with torch.no_grad():
        um, uc, *_ = m.finetune(t(seq), t(dec), users)
        dist = tr.dist_predict_full(um[:, -1, :], uc[:, -1, :]).numpy()
    out = {"seq": seq, "dec": dec, "pos": pos, "neg": neg, "cfg": np.array([B, L, H, nh, nl, I]), "p": np.array(p), "pa": np.array(pa),
           "pvn": np.array(pvn), "drop_seed": np.array(1234), "drop_step": np.array(7), "lambda1": np.array(lambda1),
           "lambda2": np.array(lambda2), "wd": np.array(wd), "loss": loss.detach().numpy(), "bpr": bpr.detach().numpy(),
           "pvn_loss": pvn_loss.detach().numpy(), "auc": auc.detach().numpy(), "mean": om.detach().numpy(), "cov": oc.detach().numpy(),
           "dist": dist, "sites": np.array(inj.k)}
    for l in range(nl):
        out[f"enc_in_mean{l}"], out[f"enc_in_cov{l}"] = enc_in[l][0].detach().numpy(), enc_in[l][1].detach().numpy()
        out[f"dec_out_mean{l}"], out[f"dec_out_cov{l}"] = dec_rev[l][0].detach().numpy(), dec_rev[l][1].detach().numpy()
        out[f"rec_mean{l}"], out[f"rec_cov{l}"] = recs[l][0].detach().numpy(), recs[l][1].detach().numpy()
    for k, v in sd0.items():
        out["sd0/" + k] = v
    for k, v in sd1.items():
        out["sd1/" + k] = v
    for k, v in grads.items():
        out["grad/" + k] = v
    np.savez_compressed(os.path.join(OUT, f"stosa_{name}.npz"), **out)
    print("stosa", name, "loss", float(loss.detach()), "bpr", float(bpr.detach()), "pvn", float(pvn_loss.detach()), "sites", inj.k, "grads",
          len(grads), "/", len(sd0))


if __name__ == "__main__":
    run("tiny_p0", B=3, L=8, H=16, nh=2, nl=2, I=30, p=0.0, pa=0.0, lambda1=[0.05, 0.1], lambda2=[0.02, 0.07], pvn=0.05)
    run("tiny_p3", B=3, L=8, H=16, nh=2, nl=2, I=30, p=0.3, pa=0.2, lambda1=[0.05, 0.1], lambda2=[0.02, 0.07], pvn=0.05)
    run("beauty_p3", B=4, L=20, H=64, nh=4, nl=1, I=150, p=0.3, pa=0.3, lambda1=[0.0021], lambda2=[0.0009], pvn=0.005)
