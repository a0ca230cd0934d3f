"""small workload for compute-sanitizer: tensor-core catalog scoring (two-pass), exact scorer, device samplers"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from adt_b200.evaluate import CatalogScorer
from adt_b200.sampler import DeviceSampler, ClozeSampler
g = torch.Generator().manual_seed(0)
I, H, U = 100_000, 64, 256
E = (torch.randn(I + 1, H, generator=g) * 0.1).cuda()
feats = torch.randn(U, H, generator=g).cuda()
fake = types.SimpleNamespace(item_emb=types.SimpleNamespace(weight=E), hidden=H)
acc = torch.zeros(6, dtype=torch.float64, device="cuda")
ans = torch.randint(1, I, (U,), generator=g).int().cuda()
s1, i1 = CatalogScorer(fake, K=10).topk_from_feats(feats, answers=ans, metric_acc=acc)
s0, i0 = CatalogScorer(fake, K=10, use_tensor_cores=False).topk_from_feats(feats)
assert torch.equal(i0, i1)
rng = np.random.default_rng(0)
train = {u: [int(x) for x in rng.integers(1, 300, size=int(rng.integers(3, 60)))] for u in range(1, 101)}
valid = {u: [train[u].pop()] for u in train}
test = {u: [train[u].pop()] for u in train}
ds = DeviceSampler(train, valid, test, 100, 300, 50)
ds.train_batch(np.arange(1, 101), epoch=1)
ds.eval_batch(np.arange(1, 101), mode="test", n_candidates=100)
cs = ClozeSampler(train, 100, 300, 20, 0.2, dupe_factor=2)
cs.batch(np.arange(len(cs)))
torch.cuda.synchronize()
print("ok")
