import sys, types, numpy as np, torch
sys.path.insert(0, "/root/repo")
from adt_b200 import synth
from adt_b200.model import SASRecADT
from adt_b200.evaluate import CatalogScorer
cfg = synth.CONFIGS["C2"]; dev = torch.device("cuda", 0)
margs = types.SimpleNamespace(device=dev, num_heads=cfg["nh"], maxlen=cfg["L"], num_layers=cfg["nl"], hidden_units=cfg["H"], dropout=cfg["p"])
m = SASRecADT(1, cfg["items"], margs).to(dev).eval(); m.engine.precision = 1
U = 512; rng = np.random.default_rng(99)
seq, ans, ip, ix = synth.make_eval_batch(rng, cfg, U)
d = [torch.from_numpy(a).to(dev) for a in (seq, ip, ix)]
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
feats = m.final_feats(d[0])
for tc, mi in ((False, 32768), (True, 0)):
    sc = CatalogScorer(m, K=10, use_tensor_cores=tc, tc_min_items=mi)
    print("tc" if tc else "exact", "score us", t(lambda: sc.topk_from_feats(feats, d[1], d[2])))
print("final_feats us", t(lambda: m.final_feats(d[0])))
eng = m.engine; w = eng.workspace(U, cfg["L"])
print("encode us", t(lambda: eng.encode(d[0], False, w)))
