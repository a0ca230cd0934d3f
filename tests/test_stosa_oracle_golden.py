"""CPU: the STOSA oracle restatement (oracle/stosa_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden_stosa.py: DisenDistSAModel.finetune + DistSAModelTrainer.{bpr_optimization,iteration,dist_predict_full})."""
import glob
import os
import numpy as np
import pytest
import torch

from oracle import stosa_oracle as SO
from oracle.sasrec_oracle import Drop
from helpers import rel_err, GOLDEN

NAMES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLDEN, "stosa_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, f"stosa_{name}.npz"))
    return {k: z[k] for k in z.files}


def test_fixtures_present():
    assert len(NAMES) >= 3


@pytest.mark.parametrize("name", NAMES)
def test_stosa_forward_loss_grads(name):
    g = load(name)
    B, L, H, nh, nl, I = [int(v) for v in g["cfg"]]
    cfg = SO.SCfg(I + 2, B, L, H, nh, nl, float(g["p"]), float(g["pa"]), float(g["pvn"]))
    sd = {k[4:]: torch.from_numpy(np.array(v)).requires_grad_(True) for k, v in g.items() if k.startswith("sd0/")}
    seq, dec, pos, neg = (torch.from_numpy(g[k]) for k in ("seq", "dec", "pos", "neg"))
    drop = Drop(0.5, int(g["drop_seed"]), int(g["drop_step"]), train=True)
    out = SO.forward(sd, cfg, seq, dec, drop)
    assert rel_err(out["mean"].detach(), g["mean"]) < 2e-5
    assert rel_err(out["cov"].detach(), g["cov"]) < 2e-5
    dec_rev = list(reversed(out["dec_outputs"]))
    for l in range(nl):
        for j, s in enumerate(("mean", "cov")):
            assert rel_err(out["enc_inputs"][l][j].detach(), g[f"enc_in_{s}{l}"]) < 2e-5
            assert rel_err(dec_rev[l][j].detach(), g[f"dec_out_{s}{l}"]) < 2e-5
            assert rel_err(out["recs"][l][j].detach(), g[f"rec_{s}{l}"]) < 2e-5
    total, bpr, pvn, auc = SO.loss(sd, cfg, out, pos, neg, list(g["lambda1"]), list(g["lambda2"]))
    assert abs(float(total) - float(g["loss"])) / abs(float(g["loss"])) < 1e-5
    assert abs(float(bpr) - float(g["bpr"])) / abs(float(g["bpr"])) < 1e-5
    assert abs(float(pvn) - float(g["pvn_loss"])) / abs(float(g["pvn_loss"])) < 1e-5
    assert abs(float(auc) - float(g["auc"])) < 1e-6
    total.backward()
    n = 0
    for k, p in sd.items():
        if "grad/" + k in g:
            assert rel_err(p.grad, g["grad/" + k]) < 5e-4, k
            n += 1
        else:   # user margins, the discarded decoder self attention, decLayerNorm: never reached by the loss
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
    assert n > 50
    # evaluation scores of the updated model + the reference's top-40 protocol
    sd1 = {k[4:]: torch.from_numpy(np.array(v)) for k, v in g.items() if k.startswith("sd1/")}
    dist = SO.predict_full(sd1, cfg, seq).numpy()
    assert rel_err(dist, g["dist"]) < 2e-5
    seen = [set(int(v) for v in row if v > 0) for row in g["seq"]]
    top = SO.full_sort_topk(dist, seen, K=10)
    ref = SO.full_sort_topk(g["dist"], seen, K=10)
    assert (top == ref).mean() > 0.95
