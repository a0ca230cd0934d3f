"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/sasrec_*.npz by running the
UNMODIFIED reference (/root/reference/sasrec/model.py + the loss/optimiser lines
of sasrec/main.py:146-173) on CPU in the build container.  /root/reference does
not exist on the GPU box, so the fixtures (not this script) travel.

    python -m oracle.make_golden            # writes tests/golden/*.npz

Dropout: torch.nn.functional.dropout is monkey-patched so the k-th call of one
forward draws its mask from oracle/philox.py (site k), permuted into the layout
the reference has at that call.
"""
import os
import sys
import types
import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/sasrec"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

from . import philox  # noqa: E402


class DropInjector:
    def __init__(self, B, L, H, nh, layers, p, seed, step):
        self.B, self.L, self.H, self.nh, self.nl, self.p, self.seed, self.step = B, L, H, nh, layers, p, seed, step
        self.k = 0
        kinds = ["blh"]
        for _ in range(layers):
            kinds += ["attn", "bhl", "bhl"]
        kinds += ["blh"]
        for _ in range(layers):
            kinds += ["attn", "attn", "lhb", "lhb"]
        self.kinds = kinds

    def __call__(self, x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        site, kind = self.k, self.kinds[self.k]
        self.k += 1
        B, L, H, nh = self.B, self.L, self.H, self.nh
        nat = (B, nh, L, L) if kind == "attn" else (B, L, H)
        if kind == "attn":
            keep = philox.keep_mask_attn(B * nh * L, L, p, self.seed, self.step, site).reshape(nat)
        else:
            keep = philox.keep_mask(int(np.prod(nat)), p, self.seed, self.step, site).reshape(nat)
        m = torch.from_numpy(keep).to(x.dtype)
        if kind == "attn":
            m = m.reshape(B * nh, L, L)
        elif kind == "bhl":
            m = m.permute(0, 2, 1)
        elif kind == "lhb":
            m = m.permute(1, 2, 0)
        assert m.shape == x.shape, (site, kind, m.shape, x.shape)
        return x * m * torch.tensor(1.0 / (1.0 - p), dtype=torch.float32).to(x.dtype)


def synth_batch(rng, B, L, I, fill=0.6):
    """Right-aligned sequences, left padded with 0 (utils.py:288-307): dec = seq shifted right."""
    seq = np.zeros((B, L), np.int64)
    dec = np.zeros((B, L), np.int64)
    pos = np.zeros((B, L), np.int64)
    neg = np.zeros((B, L), np.int64)
    for b in range(B):
        n = int(np.clip(rng.geometric(1.0 / max(2.0, fill * L)), 2, L + 1))
        items = rng.integers(1, I + 1, size=n)
        hist, nxt = items[:-1][-L:], items[1:][-L:]
        m = len(hist)
        seq[b, L - m:] = hist
        pos[b, L - m:] = nxt
        neg[b, L - m:] = rng.integers(1, I + 1, size=m)
        dec[b, L - m + 1:] = hist[:-1]
    return seq, dec, pos, neg


def run(name, B, L, H, nh, nl, I, p, lambdas1, lambdas2, wd, seed=23, dtype=torch.float32):
    sys.path.insert(0, REF)
    import model as refmodel  # noqa
    torch.manual_seed(seed)
    args = types.SimpleNamespace(device="cpu", num_heads=nh, maxlen=L, num_layers=nl, hidden_units=H, dropout=p)
    m = refmodel.SASRecADT(100, I, args)
    for _, prm in m.named_parameters():  # main.py:95-99
        try:
            torch.nn.init.xavier_normal_(prm.data)
        except Exception:
            pass
    # randomise the 1-D params too so biases/LN affine are actually exercised
    g = torch.Generator().manual_seed(seed + 1)
    for n_, prm in m.named_parameters():
        if prm.dim() == 1:
            prm.data.add_(0.1 * torch.randn(prm.shape, generator=g))
    m = m.to(dtype)
    rng = np.random.default_rng(seed)
    seq, dec, pos, neg = synth_batch(rng, B, L, I)
    sd0 = {k: v.detach().clone().numpy() for k, v in m.state_dict().items()}

    inj = DropInjector(B, L, H, nh, nl, p, seed=1234, step=7)
    orig = F.dropout
    F.dropout = inj
    try:
        m.train()
        pl, nlg, enc_in, dec_out, rec = m(None, seq, dec, pos, neg)
    finally:
        F.dropout = orig
    # loss lines of main.py:147-170, verbatim semantics
    bce = torch.nn.BCEWithLogitsLoss()
    idx = np.where(pos != 0)
    loss = bce(pl[idx], torch.ones_like(pl)[idx]) + bce(nlg[idx], torch.zeros_like(nlg)[idx])
    for i in range(len(enc_in)):
        loss = loss + lambdas1[i] * F.mse_loss(enc_in[i], dec_out[i])
    if nh > 1:
        label = torch.tile(torch.arange(nh), [B * L, 1])
        for l in range(len(rec)):
            loss = loss + lambdas2[i] * F.nll_loss(rec[l].view(B * L, nh, nh), label)  # stale i (quirk B1)
    for prm in m.item_emb.parameters():
        loss = loss + wd * torch.norm(prm)
    opt = torch.optim.Adam(m.parameters(), lr=0.001, betas=(0.9, 0.98))
    opt.zero_grad()
    loss.backward()
    gnorm = torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
    grads = {k: prm.grad.detach().clone().numpy() for k, prm in m.named_parameters() if prm.grad is not None}
    opt.step()
    sd1 = {k: v.detach().clone().numpy() for k, v in m.state_dict().items()}

    m.eval()
    with torch.no_grad():
        cand = rng.integers(1, I + 1, size=(B, 11))
        pred_c = m.predict(None, seq, cand).numpy()
        pred_f = m.predict(None, seq, None, True).numpy()

    out = {"seq": seq, "dec": dec, "pos": pos, "neg": neg, "cand": cand,
           "cfg": np.array([B, L, H, nh, nl, I]), "p": np.array(p), "drop_seed": np.array(1234), "drop_step": np.array(7),
           "lambdas1": np.array(lambdas1), "lambdas2": np.array(lambdas2), "wd": np.array(wd),
           "pos_logits": pl.detach().numpy(), "neg_logits": nlg.detach().numpy(), "loss": loss.detach().numpy(),
           "gnorm": gnorm.numpy(), "pred_cand": pred_c, "pred_full": pred_f}
    for i in range(nl):
        out[f"enc_in{i}"] = enc_in[i].detach().numpy()
        out[f"dec_out{i}"] = dec_out[i].detach().numpy()
        out[f"rec_ind{i}"] = rec[i].detach().numpy()
    for k, v in sd0.items():
        out["sd0/" + k] = v
    for k, v in sd1.items():
        out["sd1/" + k] = v
    for k, v in grads.items():
        out["grad/" + k] = v
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"sasrec_{name}.npz"), **out)
    print(name, "loss", float(loss), "gnorm", float(gnorm), "sites", inj.k)


def run_super(name, B, L, H, nh, nl, I, p, cand, wd, seed=23):
    """SuperSASRecModel (sasrec/supersasrec.py) + the warm-up step of sasrec/evolution.py:282-316."""
    sys.path.insert(0, REF)
    import supersasrec as refsuper  # noqa
    rec_choice = [0, 0.0001, 0.0005, 0.001, 0.005, 0.01]     # evolution.py:95-96
    ind_choice = [0, 0.0001, 0.0005, 0.001, 0.0015, 0.002]
    torch.manual_seed(seed)
    args = types.SimpleNamespace(device="cpu", num_heads=nh, maxlen=L, num_layers=nl, hidden_units=H, dropout=p)
    m = refsuper.SuperSASRecModel(100, I, rec_choice, ind_choice, args)
    for _, prm in m.named_parameters():
        try:
            torch.nn.init.xavier_normal_(prm.data)
        except Exception:
            pass
    g = torch.Generator().manual_seed(seed + 1)
    for _, prm in m.named_parameters():
        if prm.dim() == 1:
            prm.data.add_(0.1 * torch.randn(prm.shape, generator=g))
    rng = np.random.default_rng(seed)
    seq, dec, pos, neg = synth_batch(rng, B, L, I)
    cand = np.array(cand)
    m.set_choice(cand)
    rec_w, ind_w = [cand[2 * i] for i in range(nl)], [cand[2 * i + 1] for i in range(nl)]
    sd0 = {k: v.detach().clone().numpy() for k, v in m.state_dict().items()}
    kinds = ["blh"] + ["attn", "bhl", "bhl"] * (4 * nl) + ["blh"] + ["attn", "attn", "lhb", "lhb"] * (4 * nl)
    inj = DropInjector(B, L, H, nh, nl, p, seed=1234, step=7)
    inj.kinds = kinds
    orig = F.dropout
    F.dropout = inj
    try:
        m.train()
        pl, nlg, enc_in, dec_out, rec = m(None, seq, dec, pos, neg)
    finally:
        F.dropout = orig
    bce = torch.nn.BCEWithLogitsLoss()
    idx = np.where(pos != 0)
    loss = bce(pl[idx], torch.ones_like(pl)[idx]) + bce(nlg[idx], torch.zeros_like(nlg)[idx])
    for i in range(len(enc_in)):
        loss = loss + rec_w[i] * F.mse_loss(enc_in[i], dec_out[i])
    label = torch.tile(torch.arange(nh), [B * L, 1])
    for l in range(len(rec)):
        loss = loss + ind_w[i] * F.nll_loss(rec[l].view(B * L, nh, nh), label)   # stale i (evolution.py:313)
    opt = torch.optim.Adam(m.parameters(), lr=0.001, betas=(0.9, 0.999), weight_decay=wd)
    opt.zero_grad()
    loss.backward()
    gnorm = torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
    grads = {k: prm.grad.detach().clone().numpy() for k, prm in m.named_parameters() if prm.grad is not None}
    opt.step()
    m.eval()
    with torch.no_grad():
        cnd = rng.integers(1, I + 1, size=(B, 11))
        pred_c = m.predict(None, seq, cnd).numpy()
    out = {"seq": seq, "dec": dec, "pos": pos, "neg": neg, "cand_items": cnd, "cfg": np.array([B, L, H, nh, nl, I]), "p": np.array(p),
           "drop_seed": np.array(1234), "drop_step": np.array(7), "choice": cand, "wd": np.array(wd),
           "shared_idx": np.array(m.encoder.shared_idx), "shared_weights": np.array(m.encoder.shared_weights),
           "pos_logits": pl.detach().numpy(), "neg_logits": nlg.detach().numpy(), "loss": loss.detach().numpy(), "gnorm": gnorm.numpy(),
           "pred_cand": pred_c, "n_grads": np.array(len(grads))}
    for i in range(nl):
        out[f"enc_in{i}"] = enc_in[i].detach().numpy()
        out[f"dec_out{i}"] = dec_out[i].detach().numpy()
        out[f"rec_ind{i}"] = rec[i].detach().numpy()
    for k, v in sd0.items():
        out["sd0/" + k] = v.astype(np.float16) if k not in grads and v.ndim > 1 else v   # inactive blocks: names/shapes only matter
    for k, v in grads.items():
        out["grad/" + k] = v
        out["sd1/" + k] = m.state_dict()[k].detach().numpy()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"super_{name}.npz"), **out)
    print("super", name, "loss", float(loss), "gnorm", float(gnorm), "sites", inj.k, "grads", len(grads), "idx", m.encoder.shared_idx)


def lambdas_known_answer():
    """candidates_to_lambdas.py:11-24 run as __main__ -> its printed output is the only
    golden vector the reference itself carries for this path (pins _get_weight)."""
    import subprocess
    txt = subprocess.run([sys.executable, "/root/reference/candidates_to_lambdas.py"], capture_output=True, text=True).stdout
    with open(os.path.join(OUT, "candidates_to_lambdas.txt"), "w") as f:
        f.write(txt)
    print(txt)


if __name__ == "__main__":
    run("tiny_p0", B=4, L=8, H=16, nh=2, nl=2, I=30, p=0.0, lambdas1=[0.0124, 0.122], lambdas2=[0.0001, 0.05], wd=1e-4)
    run("tiny_p5", B=4, L=8, H=16, nh=2, nl=2, I=30, p=0.5, lambdas1=[0.0124, 0.122], lambdas2=[0.0001, 0.05], wd=1e-4)
    run("c2mini_p5", B=6, L=50, H=64, nh=2, nl=2, I=200, p=0.5, lambdas1=[0.0124, 0.122], lambdas2=[0.0001, 0.0], wd=1e-4)
    run("h128_p2", B=3, L=20, H=128, nh=4, nl=1, I=60, p=0.2, lambdas1=[0.104292], lambdas2=[0.100833], wd=1e-3)
    run_super("tiny_p5", B=4, L=8, H=16, nh=2, nl=2, I=30, p=0.5, cand=[3e-4, 1.2e-3, 2e-3, 5e-5], wd=1e-4)
    lambdas_known_answer()
