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
    sc = CatalogScorer(m, K=5, use_tensor_cores=True, tc_min_items=0)
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
