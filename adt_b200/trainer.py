"""Fused SASRec-ADT training step (the reference's sasrec/main.py:142-175 inner loop) on one or more B200s.

    forward (CUDA blocks, losses fused into the epilogues)
 -> [DP] allreduce of the 8+2*nl loss accumulators (global BCE count: main.py:151-153 takes the mean over the
         valid positions of the WHOLE batch, so ranks must agree on n before the backward)
 -> backward (CUDA blocks; weight grads by vector atomics into ONE flat buffer)
 -> sort-then-segmented embedding backward into the flat buffer's table segment
 -> [DP] ONE NCCL allreduce(sum) of the flat gradient buffer (dense params || item table)
 -> + wd * E/||E||   (main.py:170; added after the allreduce so it is counted once)
 -> global grad-norm, clip (main.py:172) and Adam(b1=.9, b2=.98) (main.py:122,173) in one pass over the flat buffer.

Every rank applies the identical update to its replica, so parameters stay bit-identical across ranks.
"""
import ctypes
import numpy as np
import torch

from . import _lib as L
from .model import _as_ids


class FusedTrainer:
    def __init__(self, model, lambdas1, lambdas2, weight_decay=0.0, lr=1e-3, betas=(0.9, 0.98), eps=1e-8, clip=5.0,
                 adam_weight_decay=0.0, seed=0, process_group=None, use_norm_decay=True):
        self.model = model
        self.eng = model.engine
        self.l1, self.l2 = [float(x) for x in lambdas1], [float(x) for x in lambdas2]
        assert len(self.l1) == model.num_layers and len(self.l2) == model.num_layers
        self.wd, self.lr, self.betas, self.eps, self.clip = float(weight_decay), lr, betas, eps, clip
        self.adam_wd = float(adam_weight_decay)       # evolution.py:111 uses Adam(weight_decay=...) instead of ||E||
        self.use_norm_decay = use_norm_decay
        self.pg = process_group
        self.world = 1
        self.rank = 0
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
            self.rank = torch.distributed.get_rank(process_group)
        self.eng.drop_seed = int(seed)
        self.t = 0
        self._w = None
        self.lib = L.lib()

    def _stream(self):
        return self.eng._stream()

    def step(self, seq, dec, pos, neg):
        """One optimisation step on this rank's shard of the batch.  ids: [B_local, L] host arrays or device tensors.
        Asynchronous; call loss() to read the step's (global) loss."""
        eng, m = self.eng, self.model
        dev = eng.dev()
        seq, dec, pos, neg = (_as_ids(a, dev) for a in (seq, dec, pos, neg))
        B, Lq = seq.shape
        eng.ensure_flat()
        grads = {n: eng.grad_view(n) for n, _ in eng.order}
        eng.batch_offset = self.rank * B
        eng.global_rows = self.world * B * Lq
        eng.drop_step = self.t
        self.t += 1
        eng.gflat.zero_()
        w = eng.forward(seq, dec, pos, neg, training=True, fused_loss=True)
        eng.sort_ids(seq, dec, pos, neg, w)
        if self.world > 1:
            torch.distributed.all_reduce(w["acc"], group=self.pg)
        eng.backward(seq, dec, pos, neg, w, grads, lambdas1=self.l1, lambdas2=self.l2)
        if self.world > 1:
            torch.distributed.all_reduce(eng.gflat, group=self.pg)
        nl = m.num_layers
        acc = w["acc"]
        s = self._stream()
        E = m.item_emb.weight
        if self.use_norm_decay and self.wd != 0.0:
            normsq = acc[3 + 2 * nl:]
            L.check(self.lib.adt_sumsq(L.ptr(E), ctypes.c_int64(E.numel()), L.ptr(normsq), s), "adt_sumsq")
            L.check(self.lib.adt_norm_decay_grad(L.ptr(grads["item_emb.weight"]), L.ptr(E), ctypes.c_int64(E.numel()),
                                                 ctypes.c_float(self.wd), L.ptr(normsq), s), "adt_norm_decay_grad")
        gn = acc[4 + 2 * nl:]
        n = eng.gflat.numel()
        L.check(self.lib.adt_sumsq(L.ptr(eng.gflat), ctypes.c_int64(n), L.ptr(gn), s), "adt_sumsq")
        eng.adam_t += 1
        a = L.fill(L.adt_adam_args(), p=eng.pflat, g=eng.gflat, m=eng.adam_m, v=eng.adam_v, n=n, lr=self.lr, beta1=self.betas[0],
                   beta2=self.betas[1], eps=self.eps, weight_decay=self.adam_wd, step=eng.adam_t, max_norm=self.clip, gnormsq=gn,
                   step_dev=None)
        L.check(self.lib.adt_adam(ctypes.byref(a), s), "adt_adam")
        self._w = w
        return w

    def loss(self):
        """Loss of the last step as main.py:174 would print it (synchronises)."""
        w, nl = self._w, self.model.num_layers
        acc = w["acc"].tolist()
        emb_norm = float(np.sqrt(acc[3 + 2 * nl])) if (self.use_norm_decay and self.wd != 0.0) else 0.0
        return self.eng.loss_from_acc(w, self.l1, self.l2, self.wd if self.use_norm_decay else 0.0, emb_norm)

    def grad_norm(self):
        return float(np.sqrt(self._w["acc"][4 + 2 * self.model.num_layers].item()))
