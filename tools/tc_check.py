"""GPU diagnostic for K7: tensor-core (tcgen05) scoring path vs the exact fp32 path, plus throughput on big catalogs."""
import sys, os, types, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from adt_b200.evaluate import CatalogScorer

def run(U, H, I, K, seen=True, seed=0):
    rng = np.random.default_rng(seed)
    feats = torch.from_numpy(rng.standard_normal((U, H)).astype(np.float32)).cuda()
    E = torch.from_numpy((rng.standard_normal((I + 1, H)) * 0.1).astype(np.float32)).cuda()
    ip = ix = None
    if seen:
        s = [np.unique(rng.integers(1, I + 1, size=rng.integers(0, 30))) for _ in range(U)]
        ip = np.zeros(U + 1, np.int32); ip[1:] = np.cumsum([len(x) for x in s]); ix = np.concatenate(s).astype(np.int32)
    fake = types.SimpleNamespace(item_emb=types.SimpleNamespace(weight=E))
    ex = CatalogScorer(fake, K=K, use_tensor_cores=False)
    tc = CatalogScorer(fake, K=K, use_tensor_cores=True, tc_min_items=0)
    s0, i0 = ex.topk_from_feats(feats, ip, ix)
    s1, i1 = tc.topk_from_feats(feats, ip, ix)
    torch.cuda.synchronize()
    same = (i0 == i1).all(dim=1).float().mean().item()
    print(f"U={U} H={H} I={I} K={K}: rows identical {same:.4f}  max|ds|={float((s0 - s1).abs().max()):.3e}  fallback={tc.fallback_users}")
    if same < 1.0:
        bad = torch.nonzero(~(i0 == i1).all(dim=1)).flatten()[:3]
        for b in bad.tolist():
            print("  row", b, i0[b].tolist(), i1[b].tolist(), s0[b].tolist()[:4], s1[b].tolist()[:4])
    return tc, feats, ip, ix

def bench(U, H, I, K, iters=10):
    tc, feats, ip, ix = run(U, H, I, K, seen=False)
    for _ in range(2):
        tc.topk_from_feats(feats)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        tc.topk_from_feats(feats)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    fl = 2.0 * U * (I + 1) * H
    print(f"  bench U={U} H={H} I={I}: {ms:.3f} ms  {U / ms * 1e3:.0f} users/s  {fl / ms / 1e9:.1f} TFLOP/s (incl. convert+rescore)")

if __name__ == "__main__":
    run(70, 64, 3000, 10)
    run(300, 128, 20000, 10)
    run(130, 256, 999, 40)
    run(512, 64, 12101, 10)
    if len(sys.argv) > 1 and sys.argv[1] == "dbg":
        for d in ("1", "2", "0"):
            os.environ["ADT_TC_DEBUG"] = d
            print("ADT_TC_DEBUG", d)
            bench(512, 64, 1_000_000, 10, iters=5)
            bench(4096, 256, 1_000_000, 10, iters=3)
    elif len(sys.argv) > 1 and sys.argv[1] == "one":
        bench(512, 64, 1_000_000, 10, iters=2)
    elif len(sys.argv) > 1:
        bench(512, 64, 1_000_000, 10)
        bench(512, 256, 1_000_000, 10)
        bench(4096, 256, 1_000_000, 10, iters=3)
