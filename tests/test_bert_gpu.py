"""GPU parity of Bert4Rec-ADT against fixtures of the UNMODIFIED reference BertModel +
trainer.py:100-132 step: compat forward (full logits), caller-side loss, backward, clip, Adam; fused masked-CE loss;
predict."""
import glob
import os
import types
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def grad_close(a, b):
    """|a-b| <= 1e-3 * max|b| + 1e-7: the key-bias gradients are identically zero in exact arithmetic (softmax is invariant
    to a constant added to every key score), so both sides only hold ~1e-9 rounding noise there."""
    a = a.detach().cpu().numpy().astype(np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() <= 1e-3 * np.abs(b).max() + 1e-7
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "bert_*.npz")))


def _load(name):
    z = np.load(os.path.join(GOLDEN, f"bert_{name}.npz"))
    return {k: z[k] for k in z.files}


def _model(g, prefix="sd0/"):
    from adt_b200.bert4rec import BertModel
    B, L, H, nh, nl, I, inner = [int(v) for v in g["cfg"]]
    args = types.SimpleNamespace(device="cuda", num_heads=nh, maxlen=L, num_layers=nl, hidden_units=H, dropout=float(g["p"]),
                                 attention_dropout=float(g["pa"]), inner_units=inner, type_vocab_size=2)
    m = BertModel(100, I, args)
    sd = {k[len(prefix):]: torch.from_numpy(np.array(v)) for k, v in g.items() if k.startswith(prefix)}
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m = m.cuda()
    m.drop_seed, m.drop_step = int(g["drop_seed"]), int(g["drop_step"])
    return m


@pytest.mark.parametrize("name", NAMES)
def test_bert_compat_step(name):
    import torch.nn.functional as F
    from adt_b200.testing import rel_err
    g = _load(name)
    B, L, H, nh, nl, I, inner = [int(v) for v in g["cfg"]]
    m = _model(g).train()
    pos_ids = torch.arange(L).repeat(B, 1)
    sent = torch.zeros(B, L, dtype=torch.long)
    logits, enc_in, dec_out, ind = m(torch.from_numpy(g["seq"]), torch.from_numpy(g["dec"]), pos_ids, sent, pos_ids, sent)
    assert rel_err(logits, g["logits"]) < 5e-5
    for i in range(nl):
        assert rel_err(enc_in[i], g[f"enc_in{i}"]) < 5e-5
        assert rel_err(dec_out[i], g[f"dec_out{i}"]) < 5e-5
        assert rel_err(ind[i], g[f"ind{i}"]) < 5e-5
    lab = torch.from_numpy(g["labels"]).cuda()
    loss = torch.nn.CrossEntropyLoss(ignore_index=0)(logits.view(-1, logits.size(-1)), lab.view(-1))   # trainer.py:113-115
    l1, l2 = list(g["lambda1"]), list(g["lambda2"])
    for i in range(nl):
        if l1[i] != 0:
            loss = loss + l1[i] * F.mse_loss(enc_in[i], dec_out[i])
    label = torch.tile(torch.arange(nh), [B * L, 1]).cuda()
    for l in range(nl):
        if l2[l] != 0:
            loss = loss + l2[l] * F.nll_loss(ind[l].view(B * L, nh, nh), label)
    assert abs(float(loss) - float(g["loss"])) / abs(float(g["loss"])) < 1e-5
    opt = torch.optim.Adam(m.parameters(), lr=0.001, betas=(0.9, 0.999), weight_decay=float(g["wd"]))
    opt.zero_grad()
    loss.backward()
    gn = torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
    assert abs(float(gn) - float(g["gnorm"])) / float(g["gnorm"]) < 1e-4
    for k, p in m.named_parameters():
        assert grad_close(p.grad, g["grad/" + k]), k
    opt.step()
    for k, p in m.named_parameters():
        big = np.abs(g["grad/" + k]) > 1e-5
        assert np.abs(p.detach().cpu().numpy() - g["sd1/" + k])[big].max(initial=0.0) < 5e-6, k


@pytest.mark.parametrize("name", NAMES)
def test_bert_fused_masked_loss_and_predict(name):
    from adt_b200.testing import rel_err
    g = _load(name)
    m = _model(g).train()
    loss = m.fused_loss(g["seq"], g["dec"], g["labels"], list(g["lambda1"]), list(g["lambda2"]))
    assert abs(float(loss) - float(g["loss"])) / abs(float(g["loss"])) < 1e-5
    loss.backward()
    gn = torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
    assert abs(float(gn) - float(g["gnorm"])) / float(g["gnorm"]) < 1e-4
    for k, p in m.named_parameters():
        assert grad_close(p.grad, g["grad/" + k]), k
    B, L = g["seq"].shape
    m2 = _model(g, prefix="sd1/").eval()
    pos_ids = torch.arange(L).repeat(B, 1)
    sent = torch.zeros(B, L, dtype=torch.long)
    pred = m2.predict(None, torch.from_numpy(g["seq"]), pos_ids, sent, torch.from_numpy(g["cand"]))
    assert rel_err(pred, g["pred"]) < 5e-5


@pytest.mark.parametrize("precision", [0, 1])
def test_bert_ml20m_like_shape_vs_oracle(precision):
    """template shape of bert4rec/templates/ml-1m.json (maxlen 200, hidden 256, 4 heads, inner 1024) with a vocabulary wide
    enough (V = 1600) to need two column blocks of the tied head; fused masked-CE loss and all gradients vs the oracle.
    precision 1 = the bf16 mode: linear layers on adt_gemm_tc, attention (200 positions, key-padding mask) as strided-batch tcgen05
    GEMMs + the softmax row kernel -- loss within 2e-2, every gradient's direction (cosine) against the fp32 oracle."""
    from adt_b200.bert4rec import BertModel
    from oracle import bert_oracle as BO
    from oracle.sasrec_oracle import Drop
    B, L, H, nh, nl, I, inner, p, pa = 2, 200, 256, 4, 1, 1500, 1024, 0.5, 0.5
    torch.manual_seed(0)
    args = types.SimpleNamespace(device="cuda", num_heads=nh, maxlen=L, num_layers=nl, hidden_units=H, dropout=p, attention_dropout=pa,
                                 inner_units=inner, type_vocab_size=2)
    m = BertModel(10, I, args)
    for prm in m.parameters():
        prm.data.normal_(0.01, 0.05) if prm.dim() >= 2 else prm.data.add_(0.1 * torch.randn(prm.shape))
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    rng = np.random.default_rng(4)
    seq = np.zeros((B, L), np.int64); dec = seq.copy(); lab = seq.copy()
    for b, n in enumerate([200, 61]):
        items = rng.integers(1, I + 1, size=n)
        msk = rng.random(n) < 0.2
        msk[-1] = True
        dec[b, L - n:] = items
        seq[b, L - n:] = np.where(msk, I + 1, items)
        lab[b, L - n:] = np.where(msk, items, 0)
    cfg = BO.BCfg(I, L, H, nh, nl, inner, p, pa)
    out = BO.forward(sd, cfg, torch.from_numpy(seq), torch.from_numpy(dec), Drop(0.5, 77, 5, train=True))
    l1, l2 = [0.005], [0.0019]
    ref = BO.loss(cfg, out, torch.from_numpy(lab), l1, l2)
    ref.backward()
    m = m.cuda().train()
    m.drop_seed, m.drop_step = 77, 5
    m.precision = precision
    loss = m.fused_loss(seq, dec, lab, l1, l2)
    assert abs(float(loss) - float(ref)) / abs(float(ref)) < (2e-2 if precision else 1e-5)
    loss.backward()
    for k, prm in m.named_parameters():
        if precision:
            a, b = prm.grad.detach().cpu().numpy().ravel().astype(np.float64), sd[k].grad.numpy().ravel().astype(np.float64)
            if np.linalg.norm(b) > 1e-7:
                assert a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30) > 0.98, k
        else:
            assert grad_close(prm.grad, sd[k].grad.numpy()), k


@pytest.mark.parametrize("name", NAMES[:2])
def test_bert_flat_optimizer_step(name):
    """fused masked loss + FlatOptimizer (clip 5.0 + Adam in libadt_b200.so) reproduces the reference's optimiser step."""
    from adt_b200.dp import FlatOptimizer
    g = _load(name)
    m = _model(g).train()
    opt = FlatOptimizer(m, lr=0.001, betas=(0.9, 0.999), weight_decay=float(g["wd"]), clip=5.0)
    opt.zero_grad()
    loss = m.fused_loss(g["seq"], g["dec"], g["labels"], list(g["lambda1"]), list(g["lambda2"]))
    loss.backward()
    opt.step()
    assert abs(opt.grad_norm() - float(g["gnorm"])) / float(g["gnorm"]) < 1e-4
    for k, p in m.named_parameters():
        big = np.abs(g["grad/" + k]) > 1e-5
        assert np.abs(p.detach().cpu().numpy() - g["sd1/" + k])[big].max(initial=0.0) < 5e-6, k
