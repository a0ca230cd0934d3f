"""CPU: the oracle restatement (oracle/sasrec_oracle.py) against fixtures produced by the
UNMODIFIED reference (oracle/make_golden.py).  This is the pin the oracle's header cites."""
import os
import numpy as np
import pytest
import torch

from oracle import sasrec_oracle as O
from oracle import philox
from helpers import golden_names, load_golden, sd_from, ids, rel_err, GOLDEN


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    out = philox.philox4x32_10(np.uint32([0]), 0, 0, 0, 0, 0)
    assert [int(x[0]) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xFFFFFFFF
    out = philox.philox4x32_10(np.uint32([f]), f, f, f, f, f)
    assert [int(x[0]) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = philox.philox4x32_10(np.uint32([0x243f6a88]), 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(x[0]) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_get_weight_matches_reference_known_answer():
    # candidates_to_lambdas.py:11-24 run as __main__ (the reference's only golden vector for this path)
    txt = open(os.path.join(GOLDEN, "candidates_to_lambdas.txt")).read().strip()
    choice = [0, 0.0001, 0.0005, 0.001, 0.005, 0.01]
    cand = [0.7053411308078107, 0.9542592593410837, 0.9296478828883573, 0.28425047269448145, 0.1600125621449342, 0.47495464861462977]
    rec = [O.get_weight(choice, cand[i]) for i in range(0, 6, 2)]
    ind = [O.get_weight(choice, cand[i + 1]) for i in range(0, 6, 2)]
    assert f"{rec} {ind}" == txt


@pytest.mark.parametrize("name", golden_names())
def test_forward_loss_grads_and_step(name):
    g = load_golden(name)
    d = g["dims"]
    cfg = O.Cfg(d["I"], d["L"], d["H"], d["nh"], d["nl"], float(g["p"]))
    sd = sd_from(g, requires_grad=True)
    seq, dec, pos, neg = ids(g)
    drop = O.Drop(float(g["p"]), int(g["drop_seed"]), int(g["drop_step"]))
    l1, l2, wd = list(g["lambdas1"]), list(g["lambdas2"]), float(g["wd"])
    total, grads, gnorm, _, out = O.train_step(sd, cfg, (seq, dec, pos, neg), l1, l2, wd, drop=drop)
    assert rel_err(out["pos_logits"].detach(), g["pos_logits"]) < 2e-5
    assert rel_err(out["neg_logits"].detach(), g["neg_logits"]) < 2e-5
    for i in range(d["nl"]):
        # embedding gathers (enc_in0) are bit exact
        if i == 0:
            assert np.array_equal(out["enc_inputs"][0].detach().numpy(), g["enc_in0"])
        assert rel_err(out["enc_inputs"][i].detach(), g[f"enc_in{i}"]) < 2e-5
        assert rel_err(out["dec_outputs"][i].detach(), g[f"dec_out{i}"]) < 2e-5
        assert rel_err(out["rec_ind"][i].detach(), g[f"rec_ind{i}"]) < 2e-5
    assert abs(float(total) - float(g["loss"])) / abs(float(g["loss"])) < 1e-5
    assert abs(float(gnorm) - float(g["gnorm"])) / float(g["gnorm"]) < 1e-4
    for k, v in grads.items():
        if v is None:
            assert "grad/" + k not in g  # unused params have no grad in the reference either
        else:
            assert rel_err(v, g["grad/" + k]) < 5e-4, k
    for k, p in sd.items():
        # Adam's first step is lr*g/(|g|+eps): only well-conditioned where |g| >> eps=1e-8
        diff = np.abs(p.detach().numpy() - g["sd1/" + k])
        assert diff.max() <= 2.1e-3, k  # never more than 2*lr apart
        if "grad/" + k in g:
            big = np.abs(g["grad/" + k]) > 1e-5
            assert diff[big].max(initial=0.0) < 2e-6, k


@pytest.mark.parametrize("name", golden_names())
def test_predict(name):
    g = load_golden(name)
    d = g["dims"]
    cfg = O.Cfg(d["I"], d["L"], d["H"], d["nh"], d["nl"])
    sd = sd_from(g, prefix="sd1/")
    seq = torch.from_numpy(g["seq"]).long()
    pc = O.predict(sd, cfg, seq, torch.from_numpy(g["cand"]).long())
    pf = O.predict(sd, cfg, seq, full=True)
    assert rel_err(pc, g["pred_cand"]) < 2e-5
    assert rel_err(pf, g["pred_full"]) < 2e-5
    assert np.array_equal(np.argsort(-pf.numpy(), axis=1)[:, :5], np.argsort(-g["pred_full"], axis=1)[:, :5])
