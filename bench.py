#!/usr/bin/env python
"""bench.py -- headline benchmark of the ADT hot path (BASELINE.json: SASRec train seqs/sec + full-catalog eval
users/sec) on N B200s of one node.

    python bench.py [--gpus N --steps K --warmup W]                 our CUDA path
    python bench.py --impl reference [...]                           the reference algorithm on the host CPU cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1, one rank per GPU)

A "step" = one SASRec-ADT optimisation step (forward, fused losses, backward, sort+segmented embedding
backward, clip + Adam) over one synthetic batch of the C2 shape (configs[1]: ~12k items, maxlen 50, hidden 64,
2 heads, 2 blocks, 256 sequences per GPU, dropout 0.5).  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from adt_b200 import synth  # noqa: E402
from adt_b200.lambdas import get_lambdas  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"],
                    help="GEMM cores of the block kernels: fp32 FFMA (reference precision) or bf16 tensor cores with fp32 accumulate")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=15.0)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.samples, self.stop, self.index = [], False, index
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}


# ------------------------------------------------------------------------------------------------ reference arm
def oracle_setup(cfg, seed=23):
    from oracle import sasrec_oracle as O
    import types
    torch.manual_seed(seed)
    sd = init_state_dict(cfg, seed)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ocfg = O.Cfg(cfg["items"], cfg["L"], cfg["H"], cfg["nh"], cfg["nl"], cfg["p"])
    return O, sd, ocfg


def init_state_dict(cfg, seed=23):
    """random-init weights of the architecture, the way sasrec/main.py:93-99 does it (xavier_normal_ on >=2-D)."""
    import types
    from adt_b200.model import SASRecADT
    torch.manual_seed(seed)
    args = types.SimpleNamespace(device="cpu", num_heads=cfg["nh"], maxlen=cfg["L"], num_layers=cfg["nl"], hidden_units=cfg["H"],
                                 dropout=cfg["p"])
    m = SASRecADT(1, cfg["items"], args)
    for _, prm in m.named_parameters():
        try:
            torch.nn.init.xavier_normal_(prm.data)
        except Exception:
            pass
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def time_cpu_reference(cfg, budget_s, steps=None, warmup=1):
    """the reference algorithm (oracle port: reference model code restated, main.py:146-173 loss/clip/Adam) on the
    host cores.  Returns (seqs_per_sec, ms_per_step, n_steps, cores)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O, sd, ocfg = oracle_setup(cfg)
    l1, l2 = get_lambdas(cfg["dataset"])
    rng = np.random.default_rng(23)
    batch = [torch.from_numpy(a).long() for a in synth.make_batch(rng, cfg)]
    opt = None
    ts = []
    t_begin = time.time()
    i = 0
    while True:
        t0 = time.time()
        drop = O.Drop(cfg["p"], 1234, i)
        _, _, _, opt, _ = O.train_step(sd, ocfg, batch, l1, l2, cfg["wd"], drop=drop, adam_state=opt)
        dt = time.time() - t0
        if i >= warmup:
            ts.append(dt)
        i += 1
        if steps is not None and len(ts) >= steps:
            break
        if steps is None and (time.time() - t_begin > budget_s and len(ts) >= 2):
            break
        if time.time() - t_begin > 8 * budget_s:
            break
    ms = 1e3 * float(np.median(ts))
    return cfg["B"] / (ms / 1e3), ms, len(ts), cores


def workload_name(args, cfg):
    """the SAME workload string for both arms (the driver compares the arms on it)"""
    return (f"SASRec-ADT {args.config}: train step (items={cfg['items']}, maxlen={cfg['L']}, hidden={cfg['H']}, heads={cfg['nh']}, "
            f"blocks={cfg['nl']}, batch={cfg['B']}/GPU, dropout={cfg['p']})")


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    v, ms, n, cores = time_cpu_reference(cfg, args.cpu_budget_s * 2, steps=max(2, min(args.steps, 10)), warmup=min(args.warmup, 1))
    line = {"metric": "train_seqs_per_sec", "value": v, "unit": "seqs/s", "n_gpus": args.gpus, "steps": n, "warmup": 1,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": workload_name(args, cfg), "parallelism": "host cores of rank 0 (all torch threads)",
                       "global_batch": cfg["B"]},
            "cpu_baseline": {"value": v, "unit": "seqs/s", "cores": cores, "kind": "port",
                             "sample": f"{n} full optimisation steps of one {cfg['B']}-sequence batch (median)"},
            "e2e": {"value": v, "unit": "seqs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def main():
    args = parse()
    cfg = synth.CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return
    import types
    from adt_b200 import _lib as L
    from adt_b200.model import SASRecADT
    from adt_b200.trainer import FusedTrainer
    from adt_b200.evaluate import CatalogScorer, hit_ndcg_mrr
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    lib = L.lib()

    margs = types.SimpleNamespace(device=dev, num_heads=cfg["nh"], maxlen=cfg["L"], num_layers=cfg["nl"], hidden_units=cfg["H"],
                                  dropout=cfg["p"])
    model = SASRecADT(1, cfg["items"], margs)
    model.load_state_dict(init_state_dict(cfg))
    model = model.to(dev).train()
    l1, l2 = get_lambdas(cfg["dataset"])
    tr = FusedTrainer(model, l1, l2, weight_decay=cfg["wd"], lr=1e-3, betas=(0.9, 0.98), clip=5.0, seed=23, use_graph=True, precision=args.precision)
    B, Lq, H = cfg["B"], cfg["L"], cfg["H"]

    rng = np.random.default_rng(23 + rank)
    POOL = 8
    host = [[torch.from_numpy(a).pin_memory() for a in synth.make_batch(rng, cfg)] for _ in range(POOL)]
    resident = [[a.to(dev) for a in b] for b in host]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def trace(msg):
        if os.environ.get("ADT_BENCH_TRACE"):
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    trace("warmup")
    for i in range(max(args.warmup, 3)):
        tr.step(*resident[i % POOL])
    barrier()
    trace("timed region")

    # ---- device-resident timing: per-step CUDA events, L2 flushed between steps (outside the event pairs)
    K = args.steps
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    with ClockSampler(local) as clk:
        barrier()
        for k in range(K):
            flush.zero_()
            evs[k][0].record()
            tr.step(*resident[k % POOL])
            evs[k][1].record()
        barrier()
        step_ms = [a.elapsed_time(b) for a, b in evs]
        total_ms = float(sum(step_ms))
        trace("e2e region")
        # ---- end to end through the public API: pinned host ids in, loss scalar out, every step.  Every step synchronises with
        # the host (loss read-back), so one scheduler hiccup on the box moves a K-step sum by >10 %: the K-step region is repeated
        # E2E_REPS times and the MEDIAN repetition is reported.
        E2E_REPS = 5
        reps = []
        last_loss = None
        for _ in range(E2E_REPS):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(K):
                tr.step(*host[k % POOL])
                last_loss = tr.loss()
            e1.record()
            barrier()
            reps.append(e0.elapsed_time(e1))
        e2e_ms = float(np.median(reps))
    clocks = clk.summary()
    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    total_ms, e2e_ms = t.tolist()
    value = world * B * K / (total_ms / 1e3)
    e2e_value = world * B * K / (e2e_ms / 1e3)

    trace("per-kernel timing")
    # ---- the other precision mode of the block GEMM cores, for the record (short device-resident run)
    other = "fp32" if args.precision == "bf16" else "bf16"
    tr2 = FusedTrainer(model, l1, l2, weight_decay=cfg["wd"], lr=1e-3, betas=(0.9, 0.98), clip=5.0, seed=23, use_graph=True,
                       precision=other)
    tr2.t, tr2._counter_t = tr.t, None
    for i in range(3):
        tr2.step(*resident[i % POOL])
    barrier()
    K2 = min(K, 20)
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K2)]
    for k in range(K2):
        flush.zero_()
        ev2[k][0].record()
        tr2.step(*resident[k % POOL])
        ev2[k][1].record()
    barrier()
    t2 = torch.tensor([sum(a.elapsed_time(b) for a, b in ev2)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t2, op=torch.distributed.ReduceOp.MAX)
    other_mode = {"dtype": other, "value": world * B * K2 / (t2.item() / 1e3), "ms_per_step": t2.item() / K2, "steps": K2}
    tr.eng.precision = {"fp32": 0, "bf16": 1}[args.precision]
    tr.t, tr._counter_t = tr2.t, None

    # ---- per-kernel live timing (separate pass, events inside the library) for the roofline object
    names_buf = ctypes.create_string_buffer(4096)
    tot = (ctypes.c_float * 64)()
    cnt = (ctypes.c_int * 64)()
    tr.use_graph = False     # the library's event scopes only exist on the eager launch path
    tr.step(*resident[0])
    torch.cuda.synchronize()
    lib.adt_timing_enable(1)
    KT = min(K, 20)
    for k in range(KT):
        flush.zero_()
        tr.step(*resident[k % POOL])
    n = lib.adt_timing_collect(names_buf, 4096, tot, cnt, 64)
    lib.adt_timing_enable(0)
    tr.use_graph = True
    knames = names_buf.value.decode().split("\n")[:n]
    kern = {knames[i]: {"ms_total": tot[i], "launches": cnt[i], "avg_us": 1e3 * tot[i] / max(cnt[i], 1)} for i in range(n)}
    step_kernel_ms = sum(v["ms_total"] for v in kern.values()) / KT
    top = max(kern, key=lambda k_: kern[k_]["ms_total"])
    M = B * Lq
    nh, nl = cfg["nh"], cfg["nl"]
    e = 4
    alg_bytes = {  # algorithmic bytes per launch (SURVEY.md section 8d per-unit figures x rows per launch)
        "enc_post_bwd": M * (3 * H * e + nh * nh * e) // 1,      # x? no: dout, y/ctx -> dctx, dy   (3 LH e + L nh^2 e per seq)
        "dec_post_bwd": M * 5 * H * e // 2,                      # half of the decoder block's 5 LH e backward bytes
        "attn_bwd": M * 7 * H * e,                               # q,k,v,dctx in ; dq,dk,dv out
        "attn_fwd": M * 4 * H * e,
        "pre_bwd": M * 5 * H * e,
        "mid_bwd": M * 8 * H * e,
        "enc_post_fwd": M * (2 * H * e + nh * nh * e),
        "dec_post_fwd": M * 3 * H * e,
        "pre_fwd": M * 4 * H * e,
        "mid_fwd": M * 6 * H * e,
    }
    peak, peak_src = peaks()
    ach = alg_bytes.get(top, 0) / (kern[top]["avg_us"] * 1e-6) / 1e9 if top in alg_bytes else None
    traffic = None      # measured DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (C2 shape only)
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")
    if args.config == "C2" and os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)["bytes_per_launch"].get(top)
    roofline = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                "traffic": traffic, "alg_bytes": alg_bytes.get(top), "peak_source": peak_src, "avg_us": kern[top]["avg_us"],
                "share_of_step": kern[top]["ms_total"] / max(sum(v["ms_total"] for v in kern.values()), 1e-9),
                "note": "at B=256 x L=50 x H=64 one launch is 0.45 waves of row-tile CTAs: latency/issue bound, not HBM bound (DESIGN.md section 4)"}

    trace("eval")
    # ---- full-catalog evaluation users/sec (encoder forward + K7 scoring + fused top-10), 512 users per batch
    model.eval()
    U = 512
    erng = np.random.default_rng(99)   # item-sharded eval: every rank scores the SAME users against its catalog shard
    eseq, eans, eip, eix = synth.make_eval_batch(erng, cfg, U)
    scorer = CatalogScorer(model, K=10, process_group=None) if world == 1 else CatalogScorer(model, K=10)
    d_seq = torch.from_numpy(eseq).to(dev)
    d_ip, d_ix = torch.from_numpy(eip).to(dev), torch.from_numpy(eix).to(dev)
    eval_launch = "eager"
    topk = scorer.topk
    if world == 1:
        try:   # single GPU: the whole evaluation batch (encoder forward + scoring + top-K) replayed as one CUDA graph
            from adt_b200.evaluate import GraphedScorer
            topk = GraphedScorer(scorer, U, cfg["L"], max_seen=len(eix)).topk
            eval_launch = "one CUDA graph per batch"
        except Exception as e:   # noqa: BLE001 -- report, then measure the eager path instead
            print(f"[bench] graphed evaluation unavailable ({e}); using eager launches", file=sys.stderr)
            topk = scorer.topk
    for _ in range(3):
        topk(d_seq, d_ip, d_ix)
    barrier()
    KE = max(10, min(K, 50))
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(KE):
        _, ids = topk(d_seq, d_ip, d_ix)
    a1.record()
    barrier()
    ev_ms = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(ev_ms, op=torch.distributed.ReduceOp.MAX)
    eval_users = U * KE / (ev_ms.item() / 1e3)   # item-sharded: all ranks score the SAME users against their catalog shard
    metrics = hit_ndcg_mrr(eans, ids)
    model.train()

    trace("cpu baseline / print")
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms, ns, cores = time_cpu_reference(cfg, args.cpu_budget_s)
        cpu = {"value": v, "unit": "seqs/s", "cores": cores, "kind": "port", "ms_per_step": ms,
               "sample": f"{ns} full optimisation steps of one {B}-sequence batch of the same workload (median, 1 warm-up)"}

    if rank == 0:
        launches_per_step = sum(v["launches"] for v in kern.values()) / KT
        # embed_sort / embed_bwd scopes contain several launches each
        passes = max(1, (int(cfg["items"]).bit_length() + 7) // 8)
        launches_per_step += (3 * passes - 1) + 3
        line = {
            "metric": "train_seqs_per_sec", "value": value, "unit": "seqs/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
            "data": "synthetic",
            "config": {"workload": workload_name(args, cfg),
                       "parallelism": f"dp{world}", "global_batch": world * B, "l2": "flushed between timed steps (256 MB write)",
                       "timing": "per-step CUDA events on the launch stream, max over ranks", "launch": "whole step replayed as one CUDA graph",
                       "e2e_timing": "median of 5 repetitions of the K-step region (each step: pinned H2D of the ids + loss read-back)"},
            "e2e": {"value": e2e_value, "unit": "seqs/s", "h2d_bytes_per_step": 4 * B * Lq * 4, "d2h_bytes_per_step": 8 * (8 + 2 * nl),
                    "ms_per_step": e2e_ms / K},
            "gpu_launches": int(round(launches_per_step * K)),
            "eval_users_per_sec": eval_users, "eval": {"users_per_batch": U, "K": 10, "items": cfg["items"] + 1, "launch": eval_launch, **metrics},
            "loss": last_loss, "other_precision": other_mode,
            "roofline": roofline, "kernels_us": {k_: round(v["avg_us"], 2) for k_, v in sorted(kern.items())},
            "kernel_ms_per_step": step_kernel_ms,
            "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
