"""CPU: the Bert4Rec oracle restatement (oracle/bert_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden_bert.py)."""
import glob
import os
import numpy as np
import pytest
import torch

from oracle import bert_oracle as BO
from oracle.sasrec_oracle import Drop
from helpers import rel_err, GOLDEN

NAMES = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "bert_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, f"bert_{name}.npz"))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", NAMES)
def test_bert_forward_loss_grads(name):
    g = load(name)
    B, L, H, nh, nl, I, inner = [int(v) for v in g["cfg"]]
    cfg = BO.BCfg(I, L, H, nh, nl, inner, float(g["p"]), float(g["pa"]))
    sd = {k[4:]: torch.from_numpy(np.array(v)).requires_grad_(True) for k, v in g.items() if k.startswith("sd0/")}
    src, dec, lab = (torch.from_numpy(g[k]) for k in ("seq", "dec", "labels"))
    drop = Drop(0.5, int(g["drop_seed"]), int(g["drop_step"]), train=True)
    out = BO.forward(sd, cfg, src, dec, drop)
    assert rel_err(out["logits"].detach(), g["logits"]) < 2e-5
    for i in range(nl):
        assert rel_err(out["enc_inputs"][i].detach(), g[f"enc_in{i}"]) < 2e-5
        assert rel_err(out["dec_outputs"][i].detach(), g[f"dec_out{i}"]) < 2e-5
        assert rel_err(out["ind_outputs"][i].detach(), g[f"ind{i}"]) < 2e-5
    total = BO.loss(cfg, out, lab, list(g["lambda1"]), list(g["lambda2"]))
    assert abs(float(total) - float(g["loss"])) / abs(float(g["loss"])) < 1e-5
    total.backward()
    gn = torch.nn.utils.clip_grad_norm_(list(sd.values()), 5.0)
    assert abs(float(gn) - float(g["gnorm"])) / float(g["gnorm"]) < 1e-4
    for k, p in sd.items():
        assert rel_err(p.grad, g["grad/" + k]) < 5e-4, k
    cand = torch.from_numpy(g["cand"])
    sd1 = {k[4:]: torch.from_numpy(np.array(v)) for k, v in g.items() if k.startswith("sd1/")}
    assert rel_err(BO.predict(sd1, cfg, src, cand), g["pred"]) < 2e-5
