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

    def step(self):
        st = ctypes.c_void_p(torch.cuda.current_stream(self.pflat.device).cuda_stream)
        if self.world > 1:
            torch.distributed.all_reduce(self.gflat, op=torch.distributed.ReduceOp.AVG, group=self.pg)
        n = self.pflat.numel()
        self.gn.zero_()
        L.check(self.lib.adt_sumsq(L.ptr(self.gflat), ctypes.c_int64(n), L.ptr(self.gn), st), "adt_sumsq")
        self.t += 1
        a = L.fill(L.adt_adam_args(), p=self.pflat, g=self.gflat, m=self.m, v=self.v, n=n, lr=self.lr, beta1=self.betas[0],
                   beta2=self.betas[1], eps=self.eps, weight_decay=self.wd, step=self.t, max_norm=self.clip, gnormsq=self.gn,
                   step_dev=None)
        L.check(self.lib.adt_adam(ctypes.byref(a), st), "adt_adam")
