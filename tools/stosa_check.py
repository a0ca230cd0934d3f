"""debug: STOSA-ADT forward pieces against the reference fixtures (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_stosa_gpu import _load, _model, NAMES
from adt_b200.testing import rel_err

for name in NAMES:
    g = _load(name)
    B, L, H, nh, nl, I = [int(v) for v in g["cfg"]]
    m = _model(g).train()
    mean, cov, _, _, enc_in, recs, dec_out = m.finetune(g["seq"], g["dec"], np.arange(B))
    dec_out.reverse()
    print(name, "mean", rel_err(mean, g["mean"]), "cov", rel_err(cov, g["cov"]))
    for l in range(nl):
        for j, s in enumerate(("mean", "cov")):
            print("  l", l, s, "enc_in", rel_err(enc_in[l][j], g[f"enc_in_{s}{l}"]), "dec_out", rel_err(dec_out[l][j], g[f"dec_out_{s}{l}"]),
                  "rec", rel_err(recs[l][j], g[f"rec_{s}{l}"]))
    m = _model(g).train()
    loss, bpr, pvn, auc = m.fused_loss(g["seq"], g["dec"], g["pos"], g["neg"], list(g["lambda1"]), list(g["lambda2"]))
    print("  loss", float(loss), float(g["loss"]), "bpr", float(bpr), float(g["bpr"]), "pvn", float(pvn), float(g["pvn_loss"]), "auc", float(auc), float(g["auc"]))
