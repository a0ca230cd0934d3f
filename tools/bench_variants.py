"""Step time of the Bert4Rec-ADT (C3 shape) and STOSA-ADT (C4 shape) training paths at BASELINE.json's full sizes: fused loss +
backward + FlatOptimizer step, synthetic data, CUDA events, 3 warm-up + 10 timed steps.  Prints one JSON line per model.
(These configs are parity-test cases, not the bench.py headline; this tool records that the full-size paths run and how fast.)"""
import json, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from adt_b200.dp import FlatOptimizer


def timed(step, warm=3, iters=10):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def seqs(rng, B, L, I, mean_len):
    s = np.zeros((B, L), np.int64)
    for b in range(B):
        n = int(np.clip(rng.geometric(1.0 / mean_len) + 2, 3, L))
        s[b, L - n:] = rng.integers(1, I + 1, size=n)
    return s


def bert():
    from adt_b200.bert4rec import BertModel
    B, L, H, nh, nl, I, inner = 256, 200, 256, 4, 2, 26744, 1024
    args = types.SimpleNamespace(device="cuda", num_heads=nh, maxlen=L, num_layers=nl, hidden_units=H, dropout=0.1, attention_dropout=0.1,
                                 inner_units=inner, type_vocab_size=2)
    torch.manual_seed(0)
    m = BertModel(100, I, args).cuda().train()
    m.precision = int(os.environ.get("BERT_PREC", "1"))     # 1: bf16 GEMM cores -- the linear layers run on tcgen05 (adt_gemm_tc)
    rng = np.random.default_rng(23)
    dec = seqs(rng, B, L, I, 144)
    mask = (rng.random((B, L)) < 0.2) & (dec > 0)
    src = np.where(mask, I + 1, dec)
    lab = np.where(mask, dec, 0)
    opt = FlatOptimizer(m, lr=1e-3, clip=5.0)
    out = {}

    def step():
        opt.zero_grad()
        loss = m.fused_loss(src, dec, lab, [0.01, 0.01], [0.001, 0.001])
        loss.backward()
        opt.step()
        out["loss"] = loss
    ms = timed(step)
    print(json.dumps({"model": "Bert4Rec-ADT C3 (B=256, L=200, H=256, nh=4, inner=1024, items=26744, mask_prob=0.2)", "precision": m.precision,
                      "ms_per_step": ms,
                      "seqs_per_sec": B / ms * 1e3, "loss": float(out["loss"]), "labelled_positions": int(mask.sum())}), flush=True)


def stosa():
    from adt_b200.stosa import DisenDistSAModel
    B, L, H, nh, nl, I = 256, 100, 64, 4, 1, 12101
    args = types.SimpleNamespace(item_size=I + 2, num_users=22363, maxlen=L, hidden_units=H, num_heads=nh, num_layers=nl, dropout=0.3,
                                 attention_dropout=0.3, initializer_range=0.02, pvn_weight=0.005)
    torch.manual_seed(0)
    m = DisenDistSAModel(args).cuda().train()
    rng = np.random.default_rng(23)
    full = seqs(rng, B, L + 1, I, 9)
    seq, pos = full[:, :-1], full[:, 1:] * (full[:, :-1] > 0)
    neg = rng.integers(1, I + 1, size=(B, L)) * (pos > 0)
    dec = np.zeros_like(seq); dec[:, 1:] = seq[:, :-1]
    opt = FlatOptimizer(m, lr=1e-3)
    out = {}

    def step():
        opt.zero_grad()
        loss, bpr, pvn, auc = m.fused_loss(seq, dec, pos, neg, [0.0021], [0.0009])
        loss.backward()
        opt.step()
        out["loss"] = loss
    ms = timed(step)
    from adt_b200.dp import GraphedStep
    # fresh parameters: autograd binds a leaf's gradient accumulator to the stream of its FIRST forward, and the eager loop above
    # ran on the default stream, which cannot be captured (see GraphedStep's docstring)
    torch.manual_seed(0)
    m = DisenDistSAModel(args).cuda().train()
    opt = FlatOptimizer(m, lr=1e-3)
    gs = GraphedStep(m, opt, lambda s, d, p, n: m.fused_loss(s, d, p, n, [0.0021], [0.0009])[0])
    dev_batch = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).cuda() for a in (seq, dec, pos, neg)]
    ms_graph = timed(lambda: gs.step(*dev_batch))
    ev = timed(lambda: m.full_sort_topk(seq, K=40), warm=2, iters=5)
    print(json.dumps({"model": "STOSA-ADT C4 (B=256, L=100, H=64, nh=4, nl=1, items=12101)", "ms_per_step": ms, "seqs_per_sec": B / ms * 1e3,
                      "ms_per_step_graphed": ms_graph, "seqs_per_sec_graphed": B / ms_graph * 1e3,
                      "loss": float(out["loss"]), "full_sort_users_per_sec": B / ev * 1e3}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "bert"):
        bert()
    if which in ("all", "stosa"):
        stosa()
