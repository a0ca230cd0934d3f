"""GPU parity of STOSA-ADT against fixtures of the UNMODIFIED reference DisenDistSAModel +
DistSAModelTrainer (oracle/make_golden_stosa.py): fused loss / gradients / Adam step, the compat `finetune` 7-tuple,
the distance matrix of `dist_predict_full` and the full-sort top-K."""
import glob
import os
import types
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLDEN, "stosa_*.npz")))


def grad_close(a, b):
    a = a.detach().cpu().numpy().astype(np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() <= 1e-3 * np.abs(b).max() + 1e-7


def _load(name):
    z = np.load(os.path.join(GOLDEN, f"stosa_{name}.npz"))
    return {k: z[k] for k in z.files}


def _model(g, prefix="sd0/"):
    from adt_b200.stosa import DisenDistSAModel
    B, L, H, nh, nl, I = [int(v) for v in g["cfg"]]
    args = types.SimpleNamespace(item_size=I + 2, num_users=B, maxlen=L, hidden_units=H, num_heads=nh, num_layers=nl,
                                 dropout=float(g["p"]), attention_dropout=float(g["pa"]), initializer_range=0.02,
                                 pvn_weight=float(g["pvn"]), cuda_condition=True)
    m = DisenDistSAModel(args)
    sd = {k[len(prefix):]: torch.from_numpy(np.array(v)) for k, v in g.items() if k.startswith(prefix)}
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m = m.cuda()
    m.drop_seed, m.drop_step = int(g["drop_seed"]), int(g["drop_step"])
    return m


@pytest.mark.parametrize("name", NAMES)
def test_stosa_fused_step(name):
    g = _load(name)
    B, L, H, nh, nl, I = [int(v) for v in g["cfg"]]
    m = _model(g).train()
    loss, bpr, pvn, auc = m.fused_loss(g["seq"], g["dec"], g["pos"], g["neg"], list(g["lambda1"]), list(g["lambda2"]))
    assert abs(float(loss) - float(g["loss"])) / abs(float(g["loss"])) < 1e-5
    assert abs(float(bpr) - float(g["bpr"])) / abs(float(g["bpr"])) < 1e-5
    assert abs(float(pvn) - float(g["pvn_loss"])) / abs(float(g["pvn_loss"])) < 1e-5
    assert abs(float(auc) - float(g["auc"])) < 1e-6
    opt = torch.optim.Adam(m.parameters(), lr=0.001, betas=(0.9, 0.999), weight_decay=float(g["wd"]))
    opt.zero_grad()
    loss.backward()
    n = 0
    for k, p in m.named_parameters():
        if "grad/" + k in g:
            assert p.grad is not None, k
            assert grad_close(p.grad, g["grad/" + k]), k
            n += 1
        else:      # user margins, the discarded decoder self attention, decLayerNorm
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
    assert n > 50
    opt.step()
    for k, p in m.named_parameters():
        if "grad/" + k not in g:
            continue
        big = np.abs(g["grad/" + k]) > 1e-5
        assert np.abs(p.detach().cpu().numpy() - g["sd1/" + k])[big].max(initial=0.0) < 5e-6, k


@pytest.mark.parametrize("name", NAMES)
def test_stosa_finetune_tuple(name):
    from adt_b200.testing import rel_err
    g = _load(name)
    B, L, H, nh, nl, I = [int(v) for v in g["cfg"]]
    m = _model(g).train()
    mean, cov, _, margins, enc_in, recs, dec_out = m.finetune(g["seq"], g["dec"], np.arange(B))
    assert margins.shape == (B, 1)
    assert rel_err(mean, g["mean"]) < 5e-5 and rel_err(cov, g["cov"]) < 5e-5
    dec_out.reverse()                                   # trainer.py:515
    for l in range(nl):
        for j, s in enumerate(("mean", "cov")):
            assert rel_err(enc_in[l][j], g[f"enc_in_{s}{l}"]) < 5e-5
            assert rel_err(dec_out[l][j], g[f"dec_out_{s}{l}"]) < 5e-5
            assert rel_err(recs[l][j], g[f"rec_{s}{l}"]) < 5e-5
    assert float(cov.min()) > 0.0                       # covariance stream stays positive (ELU + 1)


@pytest.mark.parametrize("name", NAMES)
def test_stosa_full_sort(name):
    from adt_b200.testing import rel_err
    g = _load(name)
    B, L, H, nh, nl, I = [int(v) for v in g["cfg"]]
    m = _model(g, "sd1/").eval()
    um, uc = m._last_states(g["seq"])
    dist = m.dist_predict_full(um, uc)
    assert rel_err(dist, g["dist"]) < 5e-5
    # top-K protocol of trainer.py:604-614 on the reference's own distance matrix (seen -> 1e24, ascending, ties by id)
    seen = [sorted(set(int(v) for v in row if v > 0)) for row in g["seq"]]
    indptr = np.concatenate([[0], np.cumsum([len(s) for s in seen])]).astype(np.int32)
    idx = np.concatenate(seen).astype(np.int32)
    K = 10
    ids = m.full_sort_topk(g["seq"], indptr, idx, K=K).cpu().numpy()
    d = np.array(g["dist"], dtype=np.float64)
    for u, s in enumerate(seen):
        d[u, s] = 1e24
    order = np.lexsort((np.broadcast_to(np.arange(d.shape[1]), d.shape), d), axis=1)[:, :K]
    for u in range(B):
        if (ids[u] == order[u]).all():
            continue
        # allow swaps only between near-tied distances (fp32 rounding of the matmul form)
        du = d[u]
        assert np.abs(np.sort(du[ids[u]]) - du[order[u]]).max() < 1e-4 * max(1.0, np.abs(du[order[u]]).max()), u
        assert not set(ids[u]) & set(s for s in seen[u])


def test_stosa_beauty_shape_properties():
    """C4-like shape (L=100, H=64, nh=4): runs, loss finite, gradient reaches every live parameter, cov streams positive."""
    from adt_b200.stosa import DisenDistSAModel
    rng = np.random.default_rng(5)
    B, L, H, nh, I = 16, 100, 64, 4, 3000
    args = types.SimpleNamespace(item_size=I + 2, num_users=B, maxlen=L, hidden_units=H, num_heads=nh, num_layers=1, dropout=0.3,
                                 attention_dropout=0.3, initializer_range=0.02, pvn_weight=0.005)
    torch.manual_seed(0)
    m = DisenDistSAModel(args).cuda().train()
    seq = np.zeros((B, L), np.int64); pos = np.zeros_like(seq); neg = np.zeros_like(seq)
    for b in range(B):
        n = int(rng.integers(3, L + 1))
        it = rng.integers(1, I + 1, size=n + 1)
        seq[b, L - n:], pos[b, L - n:], neg[b, L - n:] = it[:-1], it[1:], rng.integers(1, I + 1, size=n)
    dec = np.zeros_like(seq); dec[:, 1:] = seq[:, :-1]
    loss, bpr, pvn, auc = m.fused_loss(seq, dec, pos, neg, [0.0021], [0.0009])
    assert torch.isfinite(loss) and 0.0 <= float(auc) <= 1.0
    loss.backward()
    for k, p in m.named_parameters():
        if "dec_attention" in k or "user_margins" in k or "decLayerNorm" in k:
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().max()) > 0, k
    ids = m.full_sort_topk(seq, K=40).cpu().numpy()
    assert ids.shape == (B, 40) and (ids >= 0).all() and all(len(set(r)) == 40 for r in ids)


def test_stosa_flat_optimizer_step():
    """FlatOptimizer (Adam in libadt_b200.so, no clipping: stosa/trainer.py:535-537) reproduces the reference's step."""
    from adt_b200.dp import FlatOptimizer
    g = _load("beauty_p3")
    m = _model(g).train()
    opt = FlatOptimizer(m, lr=0.001, betas=(0.9, 0.999), weight_decay=float(g["wd"]))
    opt.zero_grad()
    loss, _, _, _ = m.fused_loss(g["seq"], g["dec"], g["pos"], g["neg"], list(g["lambda1"]), list(g["lambda2"]))
    loss.backward()
    opt.step()
    for k, p in m.named_parameters():
        if "grad/" + k not in g:
            assert np.abs(p.detach().cpu().numpy() - g["sd0/" + k]).max() == 0.0, k     # untouched parameters stay put
            continue
        big = np.abs(g["grad/" + k]) > 1e-5
        assert np.abs(p.detach().cpu().numpy() - g["sd1/" + k])[big].max(initial=0.0) < 5e-6, k


def test_stosa_graphed_step_matches_eager():
    """GraphedStep (whole step = one CUDA graph, device-side dropout/Adam counters) reproduces the eager loop step by step."""
    from adt_b200.dp import FlatOptimizer, GraphedStep
    g = _load("beauty_p3")
    l1, l2 = list(g["lambda1"]), list(g["lambda2"])
    batch = (g["seq"], g["dec"], g["pos"], g["neg"])
    # eager reference loop
    m0 = _model(g).train()
    o0 = FlatOptimizer(m0, lr=0.001)
    losses0 = []
    for _ in range(3):
        o0.zero_grad()
        loss = m0.fused_loss(*batch, l1, l2)[0]
        loss.backward()
        o0.step()
        losses0.append(float(loss))
    # graphed loop
    m1 = _model(g).train()
    o1 = FlatOptimizer(m1, lr=0.001)
    gs = GraphedStep(m1, o1, lambda s, d, p, n: m1.fused_loss(s, d, p, n, l1, l2)[0])
    losses1 = [float(gs.step(*batch)) for _ in range(3)]
    assert abs(losses0[0] - float(g["loss"])) / abs(float(g["loss"])) < 1e-5
    assert np.allclose(losses0, losses1, rtol=2e-5), (losses0, losses1)
    assert torch.allclose(o0.pflat, o1.pflat, rtol=1e-3, atol=2e-5)
