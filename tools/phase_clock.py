"""debug: per-phase clock64() deltas of CTA 0 in pre_fwd / attn_fwd during a C2 step."""
import ctypes, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from adt_b200 import synth, _lib as L
from adt_b200.model import SASRecADT
from adt_b200.trainer import FusedTrainer
from adt_b200.lambdas import get_lambdas
cfg = synth.CONFIGS["C2"]
args = types.SimpleNamespace(device="cuda", num_heads=cfg["nh"], maxlen=cfg["L"], num_layers=cfg["nl"], hidden_units=cfg["H"], dropout=cfg["p"])
m = SASRecADT(1, cfg["items"], args).cuda().train()
l1, l2 = get_lambdas("beauty")
for prec in ("fp32", "bf16"):
    tr = FusedTrainer(m, l1, l2, weight_decay=1e-4, precision=prec)
    b = synth.make_batch(np.random.default_rng(0), cfg)
    for _ in range(3):
        tr.step(*b)
    buf = (ctypes.c_longlong * 64)()
    L.lib().adt_debug_read(buf, 64)
    v = list(buf)
    print(prec, "pre_fwd  phases (cycles): setup %d, load %d, LN %d, store+gemm_q %d, gemm_kv %d | total %d" % (v[1]-v[0], v[2]-v[1], v[3]-v[2], v[4]-v[3], v[5]-v[4], v[5]-v[0]))
    print(prec, "  last gemm_stream(gi=0,t=0) seen: issue %d, wait+sync %d, mma %d, epilogue %d, sync %d" % (v[17]-v[16], v[18]-v[17], v[19]-v[18], v[20]-v[19], v[21]-v[20]))
    print(prec, "attn_fwd phases (cycles): setup %d, loadQ %d, S gemm %d, softmax %d, PV gemm %d | total %d" % (v[9]-v[8], v[10]-v[9], v[11]-v[10], v[12]-v[11], v[13]-v[12], v[13]-v[8]))
    print(prec, "pre_bwd  phases: load %d, LN %d, wgrad_q %d, colsum %d, gemm_dq %d, lnbwd(enc) %d, k/v loop %d, lnbwd(dec) %d, store %d | total %d" % (
        v[25]-v[24], v[26]-v[25], v[27]-v[26], v[28]-v[27], v[29]-v[28], v[30]-v[29], v[31]-v[30], v[32]-v[31], v[33]-v[32], v[33]-v[24]))
    print(prec, "attn_bwd phases: load %d, S gemm %d, dP gemm %d, softmax-bwd %d, dq gemm %d, dk/dv wgrad %d | total %d" % (
        v[37]-v[36], v[38]-v[37], v[39]-v[38], v[40]-v[39], v[41]-v[40], v[42]-v[41], v[42]-v[36]))
