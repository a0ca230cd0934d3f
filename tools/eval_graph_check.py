"""Full-catalog evaluation (encoder forward + K7 scoring + top-10) launched eagerly vs replayed as one CUDA graph, at the
bench.py workload (C2, 512 users per batch).  Checks that both give identical ids and prints users/s for each.

    python tools/eval_graph_check.py [config]
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from adt_b200 import synth  # noqa: E402
from adt_b200.model import SASRecADT  # noqa: E402
from adt_b200.evaluate import CatalogScorer, GraphedScorer  # noqa: E402


def main():
    cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
    dev = torch.device("cuda", 0)
    margs = types.SimpleNamespace(device=dev, num_heads=cfg["nh"], maxlen=cfg["L"], num_layers=cfg["nl"], hidden_units=cfg["H"],
                                  dropout=cfg["p"])
    torch.manual_seed(23)
    model = SASRecADT(1, cfg["items"], margs).to(dev).eval()
    model.engine.precision = 1
    U = 512
    rng = np.random.default_rng(99)
    batches = [synth.make_eval_batch(rng, cfg, U) for _ in range(4)]
    scorer = CatalogScorer(model, K=10)
    gs = GraphedScorer(scorer, U, cfg["L"], max_seen=max(len(b[3]) for b in batches))

    def run(fn, iters=50):
        for b in batches[:3]:
            fn(b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(batches[i % len(batches)])
        e1.record()
        torch.cuda.synchronize()
        return U * iters / (e0.elapsed_time(e1) * 1e-3)

    dbat = [tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (b[0], b[2], b[3])) for b in batches]
    for b, d in zip(batches, dbat):
        _, ids_e = scorer.topk(*d)
        ids_e = ids_e.clone()
        _, ids_g = gs.topk(b[0], b[2], b[3])
        assert torch.equal(ids_e, ids_g), "graphed eval differs from eager eval"
    idx = {id(b): d for b, d in zip(batches, dbat)}
    print("eager  (ids resident): %.0f users/s" % run(lambda b: scorer.topk(*idx[id(b)])))
    print("eager  (host ids)    : %.0f users/s" % run(lambda b: scorer.topk(b[0], b[2], b[3])))
    print("graphed(host ids)    : %.0f users/s" % run(lambda b: gs.topk(b[0], b[2], b[3])))


if __name__ == "__main__":
    main()
