"""Per-block torch.autograd bridges over the C-ABI kernels (embedding, encoder block, decoder block, logits).

The monolithic `_CompatForward` in model.py is the fast compat path of the plain SASRecADT; these finer-grained
functions let the host compose blocks the way the supernet does (4 weighted candidate blocks per layer on the SAME
input, /root/reference/sasrec/super_modules.py:35-50, :74-85) while every device op still runs in libadt_b200.so.
"""
import ctypes
import math
import torch

from . import _lib as L

ENC_KEYS = ("attention_layernorm.weight", "attention_layernorm.bias", "attention_layer.in_proj_weight", "attention_layer.in_proj_bias",
            "attention_layer.out_proj.weight", "attention_layer.out_proj.bias", "forward_layernorm.weight", "forward_layernorm.bias",
            "forward_layer.conv1.weight", "forward_layer.conv1.bias", "forward_layer.conv2.weight", "forward_layer.conv2.bias",
            "sparse.weight", "sparse.bias")
DEC_KEYS = ("layer_norm.weight", "layer_norm.bias", "slf_attn.in_proj_weight", "slf_attn.in_proj_bias", "slf_attn.out_proj.weight",
            "slf_attn.out_proj.bias", "enc_attn.in_proj_weight", "enc_attn.in_proj_bias", "enc_attn.out_proj.weight",
            "enc_attn.out_proj.bias", "pos_ffn.conv1.weight", "pos_ffn.conv1.bias", "pos_ffn.conv2.weight", "pos_ffn.conv2.bias")


def layer_params(layer, keys):
    sd = dict(layer.named_parameters())
    return [sd[k] for k in keys]


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class DropCfg:
    """dropout context of one forward: p, seed, step, training, batch offset, and the running site counter that follows
    the reference's F.dropout call order."""

    def __init__(self, p, seed, step, training, b0=0, step_dev=None):
        self.p, self.seed, self.step, self.training, self.b0 = float(p), int(seed), int(step), bool(training), int(b0)
        self.step_dev = step_dev      # optional int32 device tensor added to `step` at run time (CUDA-graph replay)
        self.site = 0

    def next(self, kind, nh, Lq, H):
        d = L.adt_dropout()
        d.enabled = 1 if (self.training and self.p > 0.0) else 0
        d.p, d.seed, d.step, d.site = self.p, self.seed, self.step, self.site
        d.base = self.b0 * (nh * Lq if kind == "attn" else Lq * H)   # attention sites: ROW offset ; row sites: element offset
        d.step_dev = self.step_dev.data_ptr() if self.step_dev is not None else None
        if self.training and self.p > 0.0:
            self.site += 1
        return d


def _f(ref, *shape):
    return torch.empty(*shape, dtype=torch.float32, device=ref.device)


def _ws_sizes(B, Lq, H, nh):
    q = L.fill(L.adt_workspace_query(), B=B, L=Lq, H=H, nh=nh, nl=1, K=1, n_splits=1)
    sz = L.adt_workspace_sizes()
    L.check(L.lib().adt_workspace_bytes(ctypes.byref(q), ctypes.byref(sz)), "adt_workspace_bytes")
    return sz


def _fs(ref, B, Lq, H, nh, prec):
    """bf16 operand scratch of the tcgen05 forward path (adt_workspace_sizes.fwd_scratch), wide bf16 models only"""
    n = int(_ws_sizes(B, Lq, H, nh).fwd_scratch) if (prec and H >= 128) else 0
    return torch.empty(n, dtype=torch.uint8, device=ref.device) if n else None


def _wg(ref, B, Lq, H, nh, prec):
    """bf16 operand scratch of the tcgen05 backward path / hoisted weight gradients (adt_workspace_sizes.wgrad_scratch), wide bf16 models only"""
    n = int(_ws_sizes(B, Lq, H, nh).wgrad_scratch) if (prec and H >= 128) else 0
    return torch.empty(n, dtype=torch.uint8, device=ref.device) if n else None


class EmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, E, P, drop):
        B, Lq = ids.shape
        H = E.shape[1]
        x = _f(E, B * Lq, H)
        a = L.fill(L.adt_embed_fwd_args(), ids=ids, item_emb=E, pos_emb=P, x=x, B=B, L=Lq, H=H, drop=drop)
        L.check(L.lib().adt_embed_fwd(ctypes.byref(a), _stream(E.device)), "adt_embed_fwd")
        ctx.ids, ctx.drop, ctx.shapes = ids, drop, (E.shape, P.shape)
        return x

    @staticmethod
    def backward(ctx, dx):
        ids, (es, ps) = ctx.ids, ctx.shapes
        B, Lq = ids.shape
        H = es[1]
        dev = dx.device
        M = B * Lq
        dx = dx.contiguous()
        dE, dP = torch.zeros(es, device=dev), torch.zeros(ps, device=dev)
        _scatter(dev, M, B, Lq, H, es[0] - 1, seq=ids, dx_enc=dx, drop_enc=ctx.drop, dE=dE, dP=dP)
        return None, dE, dP, None


def _scatter(dev, M, B, Lq, H, max_id, seq=None, pos=None, neg=None, dx_enc=None, feats=None, cpos=None, cneg=None, drop_enc=None,
             dE=None, dP=None, emb_scale=0.0, dec=None, dx_dec=None):
    """sorted segmented scatter-add for a subset of the four lookup sources (missing sources are all-padding)."""
    lib = L.lib()
    zeros = torch.zeros(B, Lq, dtype=torch.int32, device=dev)
    i32 = lambda n: torch.empty(n, dtype=torch.int32, device=dev)
    N = 4 * M
    keys, vals, kt, vt, hist = i32(N), i32(N), i32(N), i32(N), i32(256 * ((N + 255) // 256))
    a = L.fill(L.adt_embed_sort_args(), seq=seq if seq is not None else zeros, dec=dec if dec is not None else zeros,
               pos=pos if pos is not None else zeros,
               neg=neg if neg is not None else zeros, M=M, max_id=max_id, keys=keys, vals=vals, keys_tmp=kt, vals_tmp=vt, hist=hist)
    L.check(lib.adt_embed_sort(ctypes.byref(a), _stream(dev)), "adt_embed_sort")
    nb = (N + 31) // 32
    nodrop = L.adt_dropout()
    b = L.fill(L.adt_embed_bwd_args(), keys=keys, vals=vals, seq=seq if seq is not None else zeros, dec=dec if dec is not None else zeros,
               B=B, L=Lq, H=H, dx_enc=dx_enc, dx_dec=dx_dec, feats=feats, cpos=cpos, cneg=cneg, drop_enc=drop_enc if drop_enc is not None else nodrop,
               drop_dec=nodrop, d_item_emb=dE, d_pos_emb=dP, head=_f(dE, nb, H), tail=_f(dE, nb, H), has_tail=i32(nb), emb_scale=emb_scale)
    L.check(lib.adt_embed_bwd(ctypes.byref(b), _stream(dev)), "adt_embed_bwd")


def _mha_w(in_w, in_b, out_w, out_b):
    return L.fill(L.adt_mha_w(), in_w=in_w, in_b=in_b, out_w=out_w, out_b=out_b)


def _mha_g(in_w, in_b, out_w, out_b):
    return L.fill(L.adt_mha_g(), in_w=in_w, in_b=in_b, out_w=out_w, out_b=out_b)


class EncBlockFn(torch.autograd.Function):
    """EncoderLayer.forward (sasrec/modules.py:644-655) -> (out [M,H], rec_true [M,nh,nh])."""

    @staticmethod
    def forward(ctx, x, ids, cfg, *p):
        nh, training, d_attn, d_f1, d_f2, prec = cfg
        B, Lq = ids.shape
        M, H = x.shape
        x = x.contiguous()
        sv = {k: _f(x, M, H) for k in ("q", "k", "v", "ctx", "y", "h1", "out")}
        sv["lse"], sv["rec"] = _f(x, B, nh, Lq), _f(x, M, nh, nh)
        fs = _fs(x, B, Lq, H, nh, prec)
        a = L.fill(L.adt_enc_block_fwd_args(), x=x, ids=ids, ln1_w=p[0], ln1_b=p[1], attn=_mha_w(p[2], p[3], p[4], p[5]), ln2_w=p[6],
                   ln2_b=p[7], ffn=L.fill(L.adt_ffn_w(), w1=p[8], b1=p[9], w2=p[10], b2=p[11]), sparse_w=p[12], sparse_b=p[13],
                   q=sv["q"], k=sv["k"], v=sv["v"], ctx=sv["ctx"], lse=sv["lse"], y=sv["y"], h1=sv["h1"], out=sv["out"], rec=sv["rec"],
                   nll_acc=None, B=B, L=Lq, H=H, nh=nh, training=int(training), mask_mode=0, drop_attn=d_attn, drop_ffn1=d_f1,
                   drop_ffn2=d_f2, precision=int(prec), tc_scratch=fs)
        L.check(L.lib().adt_enc_block_fwd(ctypes.byref(a), _stream(x.device)), "adt_enc_block_fwd")
        ctx.sv, ctx.x, ctx.ids, ctx.cfg, ctx.p = sv, x, ids, cfg, p
        return sv["out"], sv["rec"]

    @staticmethod
    def backward(ctx, dout, drec):
        sv, x, ids, p = ctx.sv, ctx.x, ctx.ids, ctx.p
        nh, training, d_attn, d_f1, d_f2, prec = ctx.cfg
        B, Lq = ids.shape
        M, H = x.shape
        g = [torch.zeros_like(t) for t in p]
        sc = {k: _f(x, M, H) for k in ("dq", "dk", "dv", "dctx", "dy", "dx")}
        wg = _wg(x, B, Lq, H, nh, prec)
        a = L.fill(L.adt_enc_block_bwd_args(), x=x, ids=ids, ln1_w=p[0], ln1_b=p[1], attn=_mha_w(p[2], p[3], p[4], p[5]), ln2_w=p[6],
                   ln2_b=p[7], ffn=L.fill(L.adt_ffn_w(), w1=p[8], b1=p[9], w2=p[10], b2=p[11]), sparse_w=p[12], sparse_b=p[13],
                   q=sv["q"], k=sv["k"], v=sv["v"], ctx=sv["ctx"], lse=sv["lse"], y=sv["y"], h1=sv["h1"],
                   dout=dout.contiguous() if dout is not None else None, dx_extra=None,
                   drec=drec.contiguous() if drec is not None else None, nll_coef=0.0,
                   dq=sc["dq"], dk=sc["dk"], dv=sc["dv"], dctx=sc["dctx"], dy=sc["dy"], dx=sc["dx"], wgrad_scratch=wg,
                   g_ln1_w=g[0], g_ln1_b=g[1], g_attn=_mha_g(g[2], g[3], g[4], g[5]), g_ln2_w=g[6], g_ln2_b=g[7],
                   g_ffn=L.fill(L.adt_ffn_g(), w1=g[8], b1=g[9], w2=g[10], b2=g[11]), g_sparse_w=g[12], g_sparse_b=g[13],
                   B=B, L=Lq, H=H, nh=nh, mask_mode=0, drop_attn=d_attn, drop_ffn1=d_f1, drop_ffn2=d_f2, precision=int(prec))
        L.check(L.lib().adt_enc_block_bwd(ctypes.byref(a), _stream(x.device)), "adt_enc_block_bwd")
        return (sc["dx"], None, None) + tuple(g)


class DecBlockFn(torch.autograd.Function):
    """DecoderLayer.forward (sasrec/modules.py:666-677) -> out [M,H]."""

    SAVED = ("d", "q1", "k1", "v1", "ctx1", "a", "q2", "k2", "v2", "ctx2", "c", "h1")

    @staticmethod
    def forward(ctx, x, feats, ids, cfg, *p):
        nh, training, d_s, d_e, d_f1, d_f2, prec = cfg
        B, Lq = ids.shape
        M, H = x.shape
        x, feats = x.contiguous(), feats.contiguous()
        sv = {k: _f(x, M, H) for k in DecBlockFn.SAVED + ("out",)}
        sv["lse1"], sv["lse2"] = _f(x, B, nh, Lq), _f(x, B, nh, Lq)
        fs = _fs(x, B, Lq, H, nh, prec)
        a = L.fill(L.adt_dec_block_fwd_args(), x=x, feats=feats, ids=ids, ln_w=p[0], ln_b=p[1], slf=_mha_w(p[2], p[3], p[4], p[5]),
                   enc=_mha_w(p[6], p[7], p[8], p[9]), ffn=L.fill(L.adt_ffn_w(), w1=p[10], b1=p[11], w2=p[12], b2=p[13]), enc_in=None,
                   out=sv["out"], mse_acc=None, B=B, L=Lq, H=H, nh=nh, training=int(training), mask_mode=0, drop_slf=d_s, drop_enc=d_e,
                   drop_ffn1=d_f1, drop_ffn2=d_f2, precision=int(prec), tc_scratch=fs, lse1=sv["lse1"], lse2=sv["lse2"], **{k: sv[k] for k in DecBlockFn.SAVED})
        L.check(L.lib().adt_dec_block_fwd(ctypes.byref(a), _stream(x.device)), "adt_dec_block_fwd")
        ctx.sv, ctx.x, ctx.feats, ctx.ids, ctx.cfg, ctx.p = sv, x, feats, ids, cfg, p
        return sv["out"]

    @staticmethod
    def backward(ctx, dout):
        sv, x, feats, ids, p = ctx.sv, ctx.x, ctx.feats, ctx.ids, ctx.p
        nh, training, d_s, d_e, d_f1, d_f2, prec = ctx.cfg
        B, Lq = ids.shape
        M, H = x.shape
        g = [torch.zeros_like(t) for t in p]
        sc = {k: _f(x, M, H) for k in ("dq", "dk", "dv", "dctx", "dd", "dq2", "dk2", "dv2", "dctx2", "dx")}
        wg = _wg(x, B, Lq, H, nh, prec)
        dfeats = torch.zeros_like(feats)
        a = L.fill(L.adt_dec_block_bwd_args(), x=x, feats=feats, ids=ids, ln_w=p[0], ln_b=p[1], slf=_mha_w(p[2], p[3], p[4], p[5]),
                   enc=_mha_w(p[6], p[7], p[8], p[9]), ffn=L.fill(L.adt_ffn_w(), w1=p[10], b1=p[11], w2=p[12], b2=p[13]),
                   out=sv["out"], enc_in=None, mse_coef=0.0, dout=dout.contiguous(), denc=None,
                   dq=sc["dq"], dk=sc["dk"], dv=sc["dv"], dctx=sc["dctx"], dd=sc["dd"], dq2=sc["dq2"], dk2=sc["dk2"], dv2=sc["dv2"],
                   dctx2=sc["dctx2"], dfeats=dfeats, dx=sc["dx"], wgrad_scratch=wg, g_ln_w=g[0], g_ln_b=g[1], g_slf=_mha_g(g[2], g[3], g[4], g[5]),
                   g_enc=_mha_g(g[6], g[7], g[8], g[9]), g_ffn=L.fill(L.adt_ffn_g(), w1=g[10], b1=g[11], w2=g[12], b2=g[13]),
                   B=B, L=Lq, H=H, nh=nh, mask_mode=0, drop_slf=d_s, drop_enc=d_e, drop_ffn1=d_f1, drop_ffn2=d_f2, precision=int(prec),
                   lse1=sv["lse1"], lse2=sv["lse2"], **{k: sv[k] for k in DecBlockFn.SAVED})
        L.check(L.lib().adt_dec_block_bwd(ctypes.byref(a), _stream(x.device)), "adt_dec_block_bwd")
        return (sc["dx"], dfeats, None, None) + tuple(g)


class LogitsFn(torch.autograd.Function):
    """pos/neg logits <feats, E[pos|neg]> (sasrec/model.py:72-76), optionally behind a LayerNorm (ln_w/ln_b or None)."""

    @staticmethod
    def forward(ctx, x, E, pos, neg, ln_w, ln_b):
        M, H = x.shape
        x = x.contiguous()
        feats, pl, nl = _f(x, M, H), _f(x, M), _f(x, M)
        a = L.fill(L.adt_final_fwd_args(), x=x, ln_w=ln_w, ln_b=ln_b, item_emb=E, pos=pos, neg=neg, feats=feats, pos_logits=pl,
                   neg_logits=nl, acc=None, M=M, H=H)
        L.check(L.lib().adt_final_logits_loss_fwd(ctypes.byref(a), _stream(x.device)), "adt_final_logits_loss_fwd")
        ctx.saved = (x, E, pos, neg, ln_w, feats, pl, nl)
        return feats, pl, nl

    @staticmethod
    def backward(ctx, dfeats, dpl, dnl):
        x, E, pos, neg, ln_w, feats, pl, nl = ctx.saved
        M, H = x.shape
        B, Lq = pos.shape
        dev = x.device
        dx, cpos, cneg = _f(x, M, H), _f(x, M), _f(x, M)
        gw = torch.zeros_like(ln_w) if ln_w is not None else None
        gb = torch.zeros_like(ln_w) if ln_w is not None else None
        one = torch.ones(1, dtype=torch.float64, device=dev)
        z = lambda t: t.contiguous().view(-1) if t is not None else torch.zeros(M, device=dev)
        a = L.fill(L.adt_final_bwd_args(), x=x, ln_w=ln_w, item_emb=E, pos=pos, neg=neg, pos_logits=pl, neg_logits=nl,
                   dfeats_in=dfeats.contiguous() if dfeats is not None else None, n_valid=one, bce_weight=0.0, dpl_ext=z(dpl),
                   dnl_ext=z(dnl), dx=dx, cpos=cpos, cneg=cneg, g_ln_w=gw, g_ln_b=gb, M=M, H=H)
        L.check(L.lib().adt_final_logits_loss_bwd(ctypes.byref(a), _stream(dev)), "adt_final_logits_loss_bwd")
        dE = torch.zeros_like(E)
        _scatter(dev, M, B, Lq, H, E.shape[0] - 1, pos=pos, neg=neg, feats=feats, cpos=cpos, cneg=cneg, dE=dE, dP=None)
        return dx, dE, None, None, gw, gb
