"""Flat-buffer optimiser step + data-parallel gradient exchange for the autograd-composed models (Bert4Rec-ADT, STOSA-ADT,
supernet): the reference's `clip_grad_norm_` + `torch.optim.Adam` (bert4rec/trainer.py:27,129-131; stosa/trainer.py:36,
535-537; sasrec/evolution.py:111,316-318) as three launches of libadt_b200.so over ONE flat parameter / gradient / state
buffer, with a single NCCL all-reduce(avg) of the flat gradient when a process group is active (training is
data parallel, model replicated, replicas stay bit-identical)."""
import ctypes
import torch

from . import _lib as L


class FlatOptimizer:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip=None, process_group=None):
        self.params = [p for p in model.parameters() if p.requires_grad]
        if not self.params or self.params[0].device.type != "cuda":
            raise L.AdtError("adt_b200.FlatOptimizer needs the model on a CUDA device (no CPU fallback)")
        dev = self.params[0].device
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4            # 16-byte aligned segments
        self.pflat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.gflat = torch.zeros_like(self.pflat)
        self.m, self.v = torch.zeros_like(self.pflat), torch.zeros_like(self.pflat)
        for p, o in zip(self.params, offs):
            n = p.numel()
            self.pflat[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.pflat[o:o + n].view(p.shape)   # parameters become views of the flat buffer
            p.grad = self.gflat[o:o + n].view(p.shape)   # autograd accumulates straight into the flat gradient
        self.lr, self.betas, self.eps, self.wd = float(lr), betas, float(eps), float(weight_decay)
        self.clip = float(clip) if clip else 0.0
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.t = 0
        self.gn = torch.zeros(1, dtype=torch.float64, device=dev)
        self.lib = L.lib()

    def zero_grad(self):
        self.gflat.zero_()

    def grad_norm(self):
        """global gradient norm seen by the last step (before clipping)."""
        return float(self.gn.sqrt())

    def step(self, step_dev=None):
        """step_dev: optional int32 device tensor holding Adam's t (already incremented) -- used under CUDA-graph replay."""
        st = ctypes.c_void_p(torch.cuda.current_stream(self.pflat.device).cuda_stream)
        if self.world > 1:
            torch.distributed.all_reduce(self.gflat, op=torch.distributed.ReduceOp.AVG, group=self.pg)
        n = self.pflat.numel()
        self.gn.zero_()
        L.check(self.lib.adt_sumsq(L.ptr(self.gflat), ctypes.c_int64(n), L.ptr(self.gn), st), "adt_sumsq")
        self.t += 1
        a = L.fill(L.adt_adam_args(), p=self.pflat, g=self.gflat, m=self.m, v=self.v, n=n, lr=self.lr, beta1=self.betas[0],
                   beta2=self.betas[1], eps=self.eps, weight_decay=self.wd, step=self.t, max_norm=self.clip, gnormsq=self.gn,
                   step_dev=step_dev)
        L.check(self.lib.adt_adam(ctypes.byref(a), st), "adt_adam")


class GraphedStep:
    """One optimisation step of an autograd-composed model (Bert4Rec-ADT, STOSA-ADT) captured ONCE as a CUDA graph and replayed:
    zero_grad -> loss_fn(*static inputs) -> backward -> FlatOptimizer.step.  Dropout's step index and Adam's t advance through
    device-side counters, so replay k reproduces eager step k.  Single-process only (the NCCL all-reduce of the data-parallel
    path is issued eagerly by FlatOptimizer; use the eager loop there).  Build it BEFORE any eager forward/backward of the same
    parameters on the default stream: autograd binds each leaf's gradient accumulator to the stream of its first forward, and
    work on the legacy default stream cannot be captured (the warm-up here runs on a side stream for that reason).

        gs = GraphedStep(model, opt, lambda seq, dec, pos, neg: model.fused_loss(seq, dec, pos, neg, l1, l2)[0])
        loss = gs.step(seq, dec, pos, neg)          # int arrays [B, L]; returns a device scalar (no sync)
    """

    def __init__(self, model, opt, loss_fn, warmup=2):
        if opt.world > 1:
            raise L.AdtError("GraphedStep is single-process; run the eager FlatOptimizer loop under data parallelism")
        self.model, self.opt, self.loss_fn, self.warmup = model, opt, loss_fn, warmup
        self.graph, self.static, self.loss = None, None, None
        dev = opt.pflat.device
        self.counters = torch.zeros(2, dtype=torch.int32, device=dev)     # [0] dropout step offset, [1] Adam t
        self.steps = 0

    def _one(self):
        opt = self.opt
        opt.zero_grad()
        loss = self.loss_fn(*self.static)
        loss.backward()
        self.counters[1:2].add_(1)
        opt.step(step_dev=self.counters[1:2])
        self.counters[0:1].add_(1)
        return loss.detach()

    def _capture(self, arrays):
        m, opt, dev = self.model, self.opt, self.opt.pflat.device
        self.static = [torch.zeros(tuple(a.shape), dtype=torch.int32, device=dev) for a in arrays]
        for d, a in zip(self.static, arrays):
            d.copy_(_to_i32(a, dev))
        m.step_dev = self.counters[0:1]
        snap = (opt.pflat.clone(), opt.m.clone(), opt.v.clone(), self.counters.clone(), m.drop_step, opt.t)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up outside capture (lazy allocations, function attributes)
            for _ in range(self.warmup):
                self._one()
        torch.cuda.current_stream(dev).wait_stream(side)

        def restore():
            opt.pflat.copy_(snap[0]); opt.m.copy_(snap[1]); opt.v.copy_(snap[2]); self.counters.copy_(snap[3])
            m.drop_step, opt.t = snap[4], snap[5]
        restore()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._one()
        restore()                                          # capture does not execute; keep host-side counters explicit

    def step(self, *arrays):
        if self.graph is None:
            self._capture(arrays)
        dev = self.opt.pflat.device
        for d, a in zip(self.static, arrays):
            d.copy_(_to_i32(a, dev), non_blocking=True)
        self.graph.replay()
        self.steps += 1
        return self.loss


def _to_i32(a, dev):
    import numpy as np
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=torch.int32, non_blocking=True)
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev, non_blocking=True)
