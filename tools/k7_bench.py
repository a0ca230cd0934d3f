"""K7 throughput at 1M items: full path, and the kernel's debug modes (ADT_TC_DEBUG=1 pipeline only, =2 filter without inserts).
   python tools/k7_bench.py            (set ADT_TC_DEBUG in the environment for the experiments)"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from adt_b200.evaluate import CatalogScorer

def main():
    I = 1_000_000
    only = os.environ.get("K7_ONLY")           # e.g. "512,64"
    for H in (64, 256):
        torch.manual_seed(H)
        E = (torch.randn(I + 1, H, device="cuda") * 0.1)
        fake = types.SimpleNamespace(item_emb=types.SimpleNamespace(weight=E))
        for U in (512, 4096):
            if only and only != f"{U},{H}":
                continue
            feats = torch.randn(U, H, device="cuda")
            sc = CatalogScorer(fake, K=10, use_tensor_cores=True, tc_min_items=0)
            sc.refresh_table()
            for _ in range(2):
                sc.topk_from_feats(feats)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                sc.topk_from_feats(feats)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            for k, b in sc._buf.items():
                if len(b) == 6:
                    print(f"  candidates/user mean {(b[1] >= 0).sum().item() / U:.1f}  max {(b[1] >= 0).sum(dim=(0, 2)).max().item()}  slots {b[1].shape[0] * b[1].shape[2]}")
            print(f"DEBUG={os.environ.get('ADT_TC_DEBUG','0')} U={U} H={H}: {ms:.3f} ms  {2.0*U*(I+1)*H/ms/1e9:.1f} TFLOP/s  {U/ms*1e3:.0f} users/s  fallback={sc.fallback_users}", flush=True)
        del E
main()
