"""N>1 on real GPUs (skipped when fewer than 2 are visible): data-parallel training == single-GPU training on the
global batch, item-sharded evaluation == single-GPU evaluation.  Spawns one process per GPU over NCCL."""
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from adt_b200 import testing as T
    from adt_b200.trainer import FusedTrainer
    from adt_b200.evaluate import CatalogScorer
    g = T.load_golden("c2mini_p5")
    l1, l2, wd = [float(x) for x in g["lambdas1"]], [float(x) for x in g["lambdas2"]], float(g["wd"])
    seq, dec, pos, neg = g["seq"], g["dec"], g["pos"], g["neg"]
    B = seq.shape[0]
    sl = slice(rank * B // world, (rank + 1) * B // world)
    m = T.model_from_golden(g, device=f"cuda:{rank}").train()
    tr = FusedTrainer(m, l1, l2, weight_decay=wd, seed=int(g["drop_seed"]))
    tr.t = int(g["drop_step"])
    tr.step(seq[sl], dec[sl], pos[sl], neg[sl])
    loss = tr.loss()
    # the DP step must reproduce the single-process reference fixture (same dropout stream thanks to the batch offset)
    ok_loss = abs(loss - float(g["loss"])) / abs(float(g["loss"])) < 1e-5
    ok_gn = abs(tr.grad_norm() - float(g["gnorm"])) / float(g["gnorm"]) < 1e-4
    worst = 0.0
    for k, p in m.named_parameters():
        if "grad/" + k in g:
            big = np.abs(g["grad/" + k]) > 1e-5
            worst = max(worst, float(np.abs(p.detach().cpu().numpy() - g["sd1/" + k])[big].max(initial=0.0)))
    # replicas identical across ranks
    flat = m.engine.pflat.clone()
    other = flat.clone()
    dist.broadcast(other, src=0)
    same = bool(torch.equal(flat, other))
    # item-sharded eval == unsharded exact eval
    m.eval()
    sc = CatalogScorer(m, K=5, use_tensor_cores=True, tc_min_items=0, shard=True)
    assert sc.sharded
    s, i = sc.topk(seq)
    ref = m.predict(None, seq, None, True)
    rs, ri = torch.topk(ref, 5, dim=1)
    ok_eval = bool(torch.equal(i.long(), ri)) and bool(torch.allclose(s, rs, rtol=1e-5, atol=1e-6))
    ret[rank] = (ok_loss, ok_gn, worst < 5e-6, same, ok_eval)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_dp_training_and_sharded_eval_match_single_gpu():
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29400 + os.getpid() % 500, ret), nprocs=world, join=True)
    assert dict(ret) == {r: (True, True, True, True, True) for r in range(world)}, dict(ret)


def _worker_dp_generic(rank, world, port, ret):
    """Bert4Rec-ADT and STOSA-ADT data parallel: every rank takes half of the fixture batch; the all-reduced (averaged) flat
    gradient must equal the mean of the two half-batch gradients computed on one device, and replicas must stay identical."""
    import glob
    import types
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from adt_b200.dp import FlatOptimizer
    import test_bert_gpu as TB
    import test_stosa_gpu as TS
    res = []
    for kind in ("bert", "stosa"):
        if kind == "bert":
            g = TB._load("mid_p5")
            keys = ("seq", "dec", "labels")
            run = lambda m, sl: m.fused_loss(*(g[k][sl] for k in keys), list(g["lambda1"]), list(g["lambda2"]))
            mk = lambda: TB._model(g)
        else:
            g = TS._load("beauty_p3")
            keys = ("seq", "dec", "pos", "neg")
            run = lambda m, sl: m.fused_loss(*(g[k][sl] for k in keys), list(g["lambda1"]), list(g["lambda2"]))[0]
            mk = lambda: TS._model(g)
        B = g["seq"].shape[0]
        halves = [slice(r * B // world, (r + 1) * B // world) for r in range(world)]
        # expectation on one device: mean of the per-shard gradients (same dropout step for both shards, like the DP run)
        exp = None
        for sl in halves:
            m = mk().train()
            o = FlatOptimizer(m, process_group=None)
            o.world = 1
            o.zero_grad()
            run(m, sl).backward()
            exp = o.gflat.clone() if exp is None else exp + o.gflat
        exp /= world
        m = mk().train()
        opt = FlatOptimizer(m, lr=1e-3, clip=5.0)
        opt.zero_grad()
        run(m, halves[rank]).backward()
        opt.step()                       # all-reduce(avg) + clip + Adam
        gn_exp = float(exp.double().pow(2).sum().sqrt())
        coef = min(5.0 / (gn_exp + 1e-6), 1.0)
        ok_g = bool(torch.allclose(opt.gflat, exp * coef, rtol=1e-4, atol=1e-7))
        other = opt.pflat.clone()
        dist.broadcast(other, src=0)
        res.append((ok_g, bool(torch.equal(other, opt.pflat)), abs(opt.grad_norm() - gn_exp) / gn_exp < 1e-4))
    ret[rank] = tuple(res)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_dp_bert_and_stosa_flat_optimizer():
    import sys
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_dp_generic, args=(world, 29900 + os.getpid() % 90, ret), nprocs=world, join=True)
    ok = ((True, True, True), (True, True, True))
    assert dict(ret) == {r: ok for r in range(world)}, dict(ret)
