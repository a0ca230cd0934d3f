"""Fused SASRec-ADT training step (the reference's sasrec/main.py:142-175 inner loop) on one or more B200s.

    forward (CUDA blocks, losses fused into the epilogues)
    [DP] beside it: count(pos != 0) of the local batch -> all-reduce of that ONE double (main.py:151-153 takes the BCE mean
         over the valid positions of the WHOLE batch; the count depends on the batch only, so the exchange overlaps the
         forward pass instead of sitting between forward and backward)
 -> backward (CUDA blocks; weight grads by vector atomics into ONE flat buffer)
 -> sort-then-segmented embedding backward into the flat buffer's table segment
 -> [DP] ONE NCCL allreduce(sum) of the flat gradient buffer (dense params || item table)
 -> + wd * E/||E||   (main.py:170; added after the allreduce so it is counted once)
 -> global grad-norm, clip (main.py:172) and Adam(b1=.9, b2=.98) (main.py:122,173) in one pass over the flat buffer.

Every rank applies the identical update to its replica, so parameters stay bit-identical across ranks.
With world > 1 the WHOLE step, both NCCL collectives included, is captured into one CUDA graph (capture_collectives=True;
if the NCCL build refuses capture the trainer falls back to three graph segments with eager collectives between them).
The loss accumulators are all-reduced lazily, only when loss() is called.
"""
import ctypes
import numpy as np
import torch

from . import _lib as L
from .model import _as_ids


class FusedTrainer:
    def __init__(self, model, lambdas1, lambdas2, weight_decay=0.0, lr=1e-3, betas=(0.9, 0.98), eps=1e-8, clip=5.0,
                 adam_weight_decay=0.0, seed=0, process_group=None, use_norm_decay=True, use_graph=False, precision="fp32",
                 overlap=True, capture_collectives=True):
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
        self.eng.precision = {"fp32": 0, "bf16": 1}[precision]
        self.t = 0
        self._w = None
        self.use_graph = use_graph
        self.overlap = overlap        # run the encoder-independent decoder kernels on a second stream (fork/join inside the step)
        self._graphs = {}
        self._mirror_on = False
        self.capture_collectives = capture_collectives
        self.launch_mode = "eager"     # how the last step was issued (reported by bench.py)
        self.step_dev = None
        self._counter_t = None
        self.lib = L.lib()

    def _stream(self):
        return self.eng._stream()

    # ------------------------------------------------------------------------------------------
    # the step is three stream-ordered segments; with world > 1 the two NCCL all-reduces sit BETWEEN them (issued
    # eagerly), so each segment can be captured into its own CUDA graph without capturing a collective
    def _seg_forward(self, seq, dec, pos, neg):
        eng = self.eng
        B, Lq = seq.shape
        self.step_dev.add_(1)
        eng.gflat.zero_()
        cur = torch.cuda.current_stream()
        w = eng.workspace(B, Lq)
        # the radix sort of the lookup ids only depends on the batch: run it beside the forward pass
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            if self.world > 1:    # global BCE normaliser: one double, exchanged beside the forward pass
                L.check(self.lib.adt_count_nonzero(L.ptr(pos), ctypes.c_int32(pos.numel()), L.ptr(self._nvalid), self._stream()),
                        "adt_count_nonzero")
                torch.distributed.all_reduce(self._nvalid, group=self.pg)
            eng.sort_ids(seq, dec, pos, neg, w)
            if self.use_norm_decay and self.wd != 0.0:
                # ||E||^2 of the weight-decay term (main.py:170) only depends on the parameters: also beside the forward pass
                E = self.model.item_emb.weight
                self._normsq.zero_()
                L.check(self.lib.adt_sumsq(L.ptr(E), ctypes.c_int64(E.numel()), L.ptr(self._normsq), self._stream()), "adt_sumsq")
        w = eng.forward(seq, dec, pos, neg, training=True, fused_loss=True)
        cur.wait_stream(self.side)
        return w

    def _seg_backward(self, seq, dec, pos, neg, w):
        eng = self.eng
        grads = {n: eng.grad_view(n) for n, _ in eng.order}
        eng.backward(seq, dec, pos, neg, w, grads, lambdas1=self.l1, lambdas2=self.l2,
                     n_valid=self._nvalid if self.world > 1 else None)

    def _seg_optimizer(self, w):
        eng, m = self.eng, self.model
        nl = m.num_layers
        acc = w["acc"]
        s = self._stream()
        E = m.item_emb.weight
        gn = acc[4 + 2 * nl:]
        n = eng.gflat.numel()
        if self.use_norm_decay and self.wd != 0.0:
            # wd * E / ||E|| into the table segment of the flat gradient and the global gradient norm, in ONE pass
            L.check(self.lib.adt_sumsq_decay(L.ptr(eng.gflat), L.ptr(eng.pflat), ctypes.c_int64(n), ctypes.c_int64(eng.table_off),
                                             ctypes.c_float(self.wd), L.ptr(self._normsq), L.ptr(gn), s), "adt_sumsq_decay")
        else:
            L.check(self.lib.adt_sumsq(L.ptr(eng.gflat), ctypes.c_int64(n), L.ptr(gn), s), "adt_sumsq")
        a = L.fill(L.adt_adam_args(), p=eng.pflat, g=eng.gflat, m=eng.adam_m, v=eng.adam_v, n=n, lr=self.lr, beta1=self.betas[0],
                   beta2=self.betas[1], eps=self.eps, weight_decay=self.adam_wd, step=0, max_norm=self.clip, gnormsq=gn,
                   step_dev=self.step_dev[1:], mirror=(eng.mirror if self._mirror_on else None), mirror_n=eng.table_off)
        L.check(self.lib.adt_adam(ctypes.byref(a), s), "adt_adam")

    def _step_impl(self, seq, dec, pos, neg):
        """eager step on int32 device ids; every launch goes to the current stream."""
        w = self._seg_forward(seq, dec, pos, neg)
        self._seg_backward(seq, dec, pos, neg, w)
        if self.world > 1:
            torch.distributed.all_reduce(self.eng.gflat, group=self.pg)
        self._seg_optimizer(w)
        return w

    def _prepare(self, B, Lq):
        eng = self.eng
        # fast path of the replay loop: the flat buffers were validated when the graph for this shape was captured (the graph
        # replays baked pointers anyway); only probe that the first / last parameters are still views of the flat buffer
        if self.use_graph and (B, Lq) in self._graphs and eng.pflat is not None:
            (n0, p0), (n1, p1) = eng.order[0], eng.order[-1]
            base = eng.pflat.data_ptr()
            if p0.data_ptr() == base + 4 * eng.offs[n0] and p1.data_ptr() == base + 4 * eng.offs[n1]:
                return
            self._graphs.clear()        # parameters were re-homed (e.g. .to() / re-init): rebuild buffers and re-capture
        dev = eng.dev()
        eng.ensure_flat()
        # bf16 weight mirror for the sequence-resident block kernels: refreshed here, then rewritten by every Adam launch
        self._mirror_on = bool(eng.use_mirror and eng.precision and eng.mirror_kernels(Lq))
        if self._mirror_on:
            eng.refresh_mirror()
        if self.step_dev is None or self.step_dev.device != dev:
            # [0] dropout stream counter (step index of the NEXT step minus one), [1] Adam step count
            self.step_dev = torch.zeros(2, dtype=torch.int32, device=dev)
            self.side = torch.cuda.Stream(device=dev)
            self._normsq = torch.zeros(1, dtype=torch.float64, device=dev)
            self._nvalid = torch.zeros(1, dtype=torch.float64, device=dev)
            if self.overlap:
                eng.side_stream = torch.cuda.Stream(device=dev)
        eng.batch_offset = self.rank * B
        eng.global_rows = self.world * B * Lq
        eng.drop_step = 0
        eng.step_dev = self.step_dev

    def step(self, seq, dec, pos, neg):
        """One optimisation step on this rank's shard of the batch.  ids: [B_local, L] host arrays or device tensors.
        Asynchronous; call loss() to read the step's (global) loss."""
        eng = self.eng
        B, Lq = seq.shape
        self._prepare(B, Lq)
        if self._mirror_on and eng._mirror_key != eng._param_key():
            eng.refresh_mirror()       # somebody changed the parameters behind the trainer's back (load_state_dict, manual edits)
        w = self._step_body(seq, dec, pos, neg)
        # the Adam kernel rewrites the flat buffer (and the bf16 mirror) through raw pointers (torch's tensor versions do not move):
        # tell anything that caches derived copies of the parameters (CatalogScorer's bf16 table) that they are stale now
        self.model._adt_param_version = getattr(self.model, "_adt_param_version", 0) + 1
        if self._mirror_on:
            eng.mirror_marked_current()
        return w

    def _step_body(self, seq, dec, pos, neg):
        eng = self.eng
        dev = eng.dev()
        B, Lq = seq.shape
        self._prepare(B, Lq)
        # device counters: dropout step = self.t (pre-increment inside _step_impl), Adam t = number of steps taken + 1
        if self._counter_t != self.t:
            self.step_dev.copy_(torch.tensor([self.t - 1, eng.adam_t], dtype=torch.int32), non_blocking=False)
        self.t += 1
        eng.adam_t += 1
        self._counter_t = self.t
        if not self.use_graph:
            ids = [_as_ids(a, dev) for a in (seq, dec, pos, neg)]
            self._w = self._step_impl(*ids)
            return self._w
        key = (B, Lq)
        if key not in self._graphs:
            static = [torch.zeros(B, Lq, dtype=torch.int32, device=dev) for _ in range(4)]
            for dst, src in zip(static, (seq, dec, pos, neg)):
                dst.copy_(_as_ids(src, dev))
            saved = self.step_dev.clone()
            # warm-up outside capture (lazy allocations, cudaFuncSetAttribute), on a side stream as torch requires
            cap = torch.cuda.Stream(device=dev)
            cap.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap):
                snap = (eng.pflat.clone(), eng.adam_m.clone(), eng.adam_v.clone())
                self._step_impl(*static)
                eng.pflat.copy_(snap[0]); eng.adam_m.copy_(snap[1]); eng.adam_v.copy_(snap[2])
                self.step_dev.copy_(saved)
                if self._mirror_on:
                    eng.refresh_mirror()     # the warm-up step's Adam launch advanced the mirror together with the weights
            torch.cuda.current_stream().wait_stream(cap)
            w = eng.workspace(B, Lq)
            graphs = None
            if self.world == 1 or self.capture_collectives:
                try:
                    g = torch.cuda.CUDAGraph(keep_graph=True)
                    with torch.cuda.graph(g):
                        self._step_impl(*static)
                    g.instantiate()
                    graphs = [g]
                    self.launch_mode = "whole step replayed as one CUDA graph" + (" (NCCL all-reduces captured inside)" if self.world > 1 else "")
                except Exception as e:   # noqa: BLE001 -- an NCCL build that refuses stream capture: keep the collectives eager
                    if self.world == 1:
                        raise
                    import warnings
                    warnings.warn(f"adt_b200: capturing the NCCL all-reduces failed ({e}); using three graph segments")
                    torch.cuda.synchronize(dev)
            if graphs is None:
                graphs = []
                for seg in (lambda: self._seg_forward(*static), lambda: self._seg_backward(*static, w), lambda: self._seg_optimizer(w)):
                    g = torch.cuda.CUDAGraph(keep_graph=True)
                    with torch.cuda.graph(g):
                        seg()
                    g.instantiate()
                    graphs.append(g)
                self.launch_mode = "three CUDA graph segments, the gradient all-reduce issued eagerly between them"
            self.step_dev.copy_(saved)   # capture does not execute, but keep the counters explicit
            self._graphs[key] = (graphs, static, w)
        graphs, static, w = self._graphs[key]
        for dst, src in zip(static, (seq, dec, pos, neg)):
            if isinstance(src, torch.Tensor):
                dst.copy_(src, non_blocking=True)
            else:
                dst.copy_(torch.from_numpy(np.ascontiguousarray(src, dtype=np.int32)), non_blocking=True)
        if len(graphs) == 1:
            graphs[0].replay()
        else:
            graphs[0].replay()
            graphs[1].replay()
            torch.distributed.all_reduce(eng.gflat, group=self.pg)
            graphs[2].replay()
        self._w = w
        return w

    def loss(self):
        """Loss of the last step as main.py:174 would print it (synchronises)."""
        w, nl = self._w, self.model.num_layers
        acc = self._read_acc(w)
        emb_norm = float(np.sqrt(acc[3 + 2 * nl])) if (self.use_norm_decay and self.wd != 0.0) else 0.0
        return self.eng.loss_from_acc(w, self.l1, self.l2, self.wd if self.use_norm_decay else 0.0, emb_norm, acc=acc)

    def _read_acc(self, w):
        """device accumulators -> host list through a persistent pinned buffer (one async copy + one stream sync)."""
        a = w["acc"]
        h = getattr(self, "_acc_host", None)
        if h is None or h.numel() != a.numel():
            h = self._acc_host = torch.empty(a.numel(), dtype=a.dtype).pin_memory()
        h.copy_(a, non_blocking=True)
        hn = getattr(self, "_normsq_host", None)
        if hn is None:
            hn = self._normsq_host = torch.empty(1, dtype=torch.float64).pin_memory()
        hn.copy_(self._normsq, non_blocking=True)       # ||E||^2 lives in its own buffer (computed beside the forward pass)
        torch.cuda.current_stream(a.device).synchronize()
        if self.world > 1:        # loss sums of the global batch: exchanged only when somebody asks for the loss
            n_loss = 3 + 2 * self.model.num_layers          # BCE sums, valid count, MSE / NLL sums are per-rank partial sums;
            t = h[:n_loss].clone().to(a.device)              # ||E||^2 and the gradient norm behind them are already global
            torch.distributed.all_reduce(t, group=self.pg)
            h = h.clone()
            h[:n_loss] = t.cpu()
        out = h.tolist()
        out[3 + 2 * self.model.num_layers] = float(hn[0])
        return out

    def kernel_nodes(self):
        """number of kernel nodes of the captured step graph(s) of the last shape = kernels launched per replayed step."""
        from .graphs import count_kernel_nodes
        if not self._graphs or self._w is None:
            return None
        graphs = self._graphs[(self._w["B"], self._w["L"])][0]
        return sum(count_kernel_nodes(g) for g in graphs)

    def grad_norm(self):
        return float(np.sqrt(self._read_acc(self._w)[4 + 2 * self.model.num_layers]))
