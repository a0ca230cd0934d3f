"""Per-kernel roofline microbenchmark of the memory-bound kernels (K1 gather, K2 sort + segmented scatter-add, K6 clip+Adam)
and of catalog scoring (K7) at shapes large enough to leave the launch-latency regime.  Each kernel is timed ALONE with CUDA
events on its launch stream after warm-up, with an L2 flush (256 MB write) between iterations; achieved = ALGORITHMIC bytes
 / time, peak = MEASURED_PEAKS.json.  Prints one JSON line per kernel and writes profiles/<tag>_micro.json.

    python tools/micro_bench.py [tag]
"""
import ctypes
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from adt_b200 import _lib as L  # noqa: E402
from adt_b200.model import SASRecADT  # noqa: E402
from adt_b200.evaluate import CatalogScorer  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 7700.0, 2250.0, "fallback (B200_PROFILING.md nominal)"


def timeit(fn, iters=10, warm=3):
    if os.environ.get("MICRO_ONCE"):     # profiling runs (ncu): one launch of everything
        iters, warm = 1, 0
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e-3


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    hbm, tf, src = peaks()
    lib = L.lib()
    dev = torch.device("cuda")
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = []

    def rec(name, shape, seconds, alg_bytes=None, flops=None):
        r = {"kernel": name, "shape": shape, "us": seconds * 1e6}
        if alg_bytes is not None:
            r.update(bound="hbm", alg_bytes=alg_bytes, achieved=alg_bytes / seconds / 1e9, peak=hbm, unit="GB/s")
        else:
            r.update(bound="tensor", flops=flops, achieved=flops / seconds / 1e12, peak=tf, unit="TFLOP/s")
        r["frac"] = r["achieved"] / r["peak"]
        r["peak_source"] = src
        out.append(r)
        print(json.dumps(r), flush=True)

    rng = np.random.default_rng(0)
    for (B, Lq, H, I) in ((2048, 200, 64, 1_000_000), (1024, 200, 256, 1_000_000)):
        M = B * Lq
        args = types.SimpleNamespace(device="cuda", num_heads=2, maxlen=Lq, num_layers=1, hidden_units=H, dropout=0.5)
        m = SASRecADT(1, I, args).cuda().train()
        eng = m.engine
        ids = torch.from_numpy(rng.integers(1, I + 1, size=(4, B, Lq)).astype(np.int32)).to(dev)
        seq, dec, pos, neg = ids[0], ids[1], ids[2], ids[3]
        x = torch.empty(M, H, device=dev)
        shape = f"B={B} L={Lq} H={H} items={I}"
        # K1: gather fwd. per looked-up row: H*4 read + H*4 write + 4 B index
        d = eng._drop(0, "row", True, B, Lq)
        a = L.fill(L.adt_embed_fwd_args(), ids=seq, item_emb=m.item_emb.weight, pos_emb=m.pos_emb.weight, x=x, B=B, L=Lq, H=H, drop=d)
        t = timeit(lambda: L.check(lib.adt_embed_fwd(ctypes.byref(a), st()), "embed_fwd"))
        rec("K1 embed_fwd (gather, x sqrt(H), +pos, dropout, pad mask)", shape, t, alg_bytes=M * (2 * H * 4 + 4))
        # K2: sort + scatter.  per looked-up row: H*4 read (grad) + H*4 RMW write + 8 B (key+perm); rows = 4*M
        w = eng.workspace(B, Lq)
        t = timeit(lambda: eng.sort_ids(seq, dec, pos, neg, w))
        rec("K2a embed_sort (stable LSD radix sort of 4*M (id, element) pairs)", shape, t, alg_bytes=4 * M * 8 * 2 * 4)
        dE = torch.zeros_like(m.item_emb.weight)
        dP = torch.zeros_like(m.pos_emb.weight)
        dx = torch.randn(M, H, device=dev)
        cp = torch.randn(M, device=dev)
        nodrop = L.adt_dropout()
        b = L.fill(L.adt_embed_bwd_args(), keys=w["keys"], vals=w["vals"], seq=seq, dec=dec, B=B, L=Lq, H=H, dx_enc=dx, dx_dec=dx, feats=dx,
                   cpos=cp, cneg=cp, drop_enc=nodrop, drop_dec=nodrop, d_item_emb=dE, d_pos_emb=dP, head=w["head"], tail=w["tail"],
                   has_tail=w["has_tail"], emb_scale=1.0)
        t = timeit(lambda: L.check(lib.adt_embed_bwd(ctypes.byref(b), st()), "embed_bwd"))
        rec("K2b embed_bwd (segmented scatter-add of 4*M rows + pos_emb reduction)", shape, t, alg_bytes=4 * M * (2 * H * 4 + 8))
        # K6: clip + Adam over the flat buffer: 7 * n * 4 bytes
        eng.ensure_flat()
        n = eng.pflat.numel()
        mbuf, vbuf = torch.zeros_like(eng.pflat), torch.zeros_like(eng.pflat)
        eng.gflat.normal_()
        gn = torch.ones(1, dtype=torch.float64, device=dev)
        aa = L.fill(L.adt_adam_args(), p=eng.pflat, g=eng.gflat, m=mbuf, v=vbuf, n=n, lr=1e-3, beta1=0.9, beta2=0.98, eps=1e-8,
                    weight_decay=0.0, step=1, max_norm=5.0, gnormsq=gn, step_dev=None)
        t = timeit(lambda: L.check(lib.adt_adam(ctypes.byref(aa), st()), "adam"))
        rec("K6 adam (clip + Adam on the flat parameter buffer)", f"n={n}", t, alg_bytes=7 * n * 4)
        acc = torch.zeros(1, dtype=torch.float64, device=dev)
        t = timeit(lambda: L.check(lib.adt_sumsq(L.ptr(eng.gflat), ctypes.c_int64(n), L.ptr(acc), st()), "sumsq"))
        rec("K6 sumsq (global gradient norm)", f"n={n}", t, alg_bytes=n * 4)
        # K7: catalog scoring + top-10
        if H % 64 == 0:
            for U in (512, 4096):
                feats = torch.randn(U, H, device=dev)
                sc = CatalogScorer(m, K=10)
                sc.refresh_table()
                t = timeit(lambda: sc.topk_from_feats(feats), iters=5, warm=2)
                rec("K7 score_topk_tc (bf16 tcgen05 GEMM + fused top-K + fp32 re-score)", f"U={U} items={I + 1} H={H}", t,
                    flops=2.0 * U * (I + 1) * H)
        del m, eng, w, dE, dx, mbuf, vbuf
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"{tag}_micro.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
