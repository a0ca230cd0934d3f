"""Parity harness shared by tests/, __graft_entry__.smoke() and tools/gpu_check.py.

This is the only module of the package that touches oracle/ -- and only as the CHECKER of results produced by
the CUDA path (never as a compute path)."""
import glob
import os
import types
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_names():
    return sorted(os.path.basename(p)[len("sasrec_"):-4] for p in glob.glob(os.path.join(GOLDEN, "sasrec_*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"sasrec_{name}.npz"))
    g = {k: z[k] for k in z.files}
    B, L, H, nh, nl, I = [int(v) for v in g["cfg"]]
    g["dims"] = dict(B=B, L=L, H=H, nh=nh, nl=nl, I=I)
    return g


def make_args(L, H, nh, nl, p, device="cuda"):
    return types.SimpleNamespace(device=device, num_heads=nh, maxlen=L, num_layers=nl, hidden_units=H, dropout=p)


def model_from_golden(g, prefix="sd0/", device="cuda"):
    from .model import SASRecADT
    d = g["dims"]
    m = SASRecADT(100, d["I"], make_args(d["L"], d["H"], d["nh"], d["nl"], float(g["p"]), device))
    sd = {k[len(prefix):]: torch.from_numpy(np.array(v)) for k, v in g.items() if k.startswith(prefix)}
    m.load_state_dict(sd)
    return m.to(device)


def rel_err(a, b):
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def compat_loss(model, outs, pos, l1, l2, wd):
    """the caller-side loss lines of sasrec/main.py:147-170, written with torch ops on our module's outputs."""
    import torch.nn.functional as F
    pl, nlg, enc_in, dec_out, rec = outs
    dev = pl.device
    idx = np.where(pos != 0)
    bce = torch.nn.BCEWithLogitsLoss()
    loss = bce(pl[idx], torch.ones_like(pl)[idx]) + bce(nlg[idx], torch.zeros_like(nlg)[idx])
    for i in range(len(enc_in)):
        loss = loss + l1[i] * F.mse_loss(enc_in[i], dec_out[i])
    nh = model.num_heads
    if nh > 1:
        B, Lq = pos.shape
        label = torch.tile(torch.arange(nh), [B * Lq, 1]).to(dev)
        for l in range(len(rec)):
            loss = loss + l2[i] * F.nll_loss(rec[l].view(B * Lq, nh, nh), label)
    for prm in model.item_emb.parameters():
        loss = loss + wd * torch.norm(prm)
    return loss


def check_golden(name, verbose=False, precision="fp32"):
    """Run compat forward/backward and the fused step on fixture `name`; return {quantity: relative error}."""
    from .trainer import FusedTrainer
    g = load_golden(name)
    d = g["dims"]
    l1, l2, wd = [float(x) for x in g["lambdas1"]], [float(x) for x in g["lambdas2"]], float(g["wd"])
    seq, dec, pos, neg = g["seq"], g["dec"], g["pos"], g["neg"]
    errs = {}
    # ---- compat path (autograd bridge)
    m = model_from_golden(g)
    m.train()
    m.engine.drop_seed, m.engine.drop_step = int(g["drop_seed"]), int(g["drop_step"])
    m.engine.precision = {"fp32": 0, "bf16": 1}[precision]
    outs = m(None, seq, dec, pos, neg)
    errs["pos_logits"] = rel_err(outs[0], g["pos_logits"])
    errs["neg_logits"] = rel_err(outs[1], g["neg_logits"])
    for i in range(d["nl"]):
        errs[f"enc_in{i}"] = rel_err(outs[2][i], g[f"enc_in{i}"])
        errs[f"dec_out{i}"] = rel_err(outs[3][i], g[f"dec_out{i}"])
        errs[f"rec_ind{i}"] = rel_err(outs[4][i], g[f"rec_ind{i}"])
    errs["enc_in0_bitexact"] = 0.0 if np.array_equal(outs[2][0].detach().cpu().numpy(), g["enc_in0"]) else 1.0
    loss = compat_loss(m, outs, pos, l1, l2, wd)
    errs["compat_loss"] = abs(float(loss) - float(g["loss"])) / abs(float(g["loss"]))
    loss.backward()
    gn = torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
    errs["compat_gnorm"] = abs(float(gn) - float(g["gnorm"])) / float(g["gnorm"])
    for k, p in m.named_parameters():
        if "grad/" + k in g:
            errs["compat_grad/" + k] = rel_err(p.grad, g["grad/" + k])
    # ---- fused path
    m2 = model_from_golden(g)
    m2.train()
    tr = FusedTrainer(m2, l1, l2, weight_decay=wd, lr=1e-3, betas=(0.9, 0.98), clip=5.0, seed=int(g["drop_seed"]), precision=precision)
    tr.t = int(g["drop_step"])
    tr.step(seq, dec, pos, neg)
    errs["fused_loss"] = abs(tr.loss() - float(g["loss"])) / abs(float(g["loss"]))
    errs["fused_gnorm"] = abs(tr.grad_norm() - float(g["gnorm"])) / float(g["gnorm"])
    eng = m2.engine
    for k, _ in eng.order:
        if "grad/" + k in g:
            errs["fused_grad/" + k] = rel_err(eng.grad_view(k), g["grad/" + k])
    worst = 0.0
    for k, p in m2.named_parameters():
        diff = np.abs(p.detach().cpu().numpy() - g["sd1/" + k])
        if "grad/" + k in g:
            big = np.abs(g["grad/" + k]) > 1e-5
            worst = max(worst, float(diff[big].max(initial=0.0)))
        errs["fused_step_maxabs_any"] = max(errs.get("fused_step_maxabs_any", 0.0), float(diff.max()))
    errs["fused_step_maxabs_wellcond"] = worst
    # ---- predict
    m3 = model_from_golden(g, prefix="sd1/")
    m3.eval()
    m3.engine.precision = {"fp32": 0, "bf16": 1}[precision]
    errs["pred_cand"] = rel_err(m3.predict(None, seq, g["cand"]), g["pred_cand"])
    errs["pred_full"] = rel_err(m3.predict(None, seq, None, True), g["pred_full"])
    if verbose:
        for k, v in errs.items():
            print(f"  {name:12s} {k:70s} {v:.3e}")
    return errs


def tolerance(key):
    if key.endswith("bitexact"):
        return 0.5
    if "grad/" in key:
        return 1e-3
    if key.startswith("fused_step_maxabs_wellcond"):
        return 5e-6
    if key.startswith("fused_step_maxabs_any"):
        return 2.1e-3
    if "gnorm" in key:
        return 1e-4
    if "loss" in key:
        return 1e-5
    return 5e-5


def smoke_check(verbose=False):
    """one tiny training step + predict on cuda:0 against the golden fixture of the unmodified reference."""
    errs = check_golden("tiny_p5", verbose=verbose)
    bad = {k: v for k, v in errs.items() if not (v <= tolerance(k))}
    if bad:
        raise AssertionError(f"smoke parity failed: {bad}")
    torch.cuda.synchronize()
    if verbose:
        print("smoke ok:", len(errs), "quantities within tolerance")
