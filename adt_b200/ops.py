"""torch.autograd bridges over the generic C-ABI ops (include/adt_b200.h "generic ops"): linear, dropout+residual+
LayerNorm, attention core, three-table embedding sum.  Used to compose the post-LN backbones on the host
(bert4rec.py); every device op is a libadt_b200.so kernel."""
import ctypes
import torch

from . import _lib as L
from .blocks import _scatter


def _st(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


NBLK = 1024   # column block of wide linear layers


def no_drop():
    return L.adt_dropout()


SM_COUNT = 148
TC_MIN_WORK = 1 << 22     # M*N*K below this: the row-tile kernel's single launch beats convert + GEMM


def _ceil8(n):
    return (n + 7) // 8 * 8


def _bf16(x, transpose=False):
    """bf16 operand copy of a contiguous fp32 matrix [R, C] (leading dimension padded to a multiple of 8 for TMA), or of its transpose"""
    lib = L.lib()
    R, C = x.shape
    st = _st(x.device)
    if transpose:
        ld = _ceil8(R)
        y = torch.empty(C, ld, dtype=torch.bfloat16, device=x.device)
        if ld != R:
            y[:, R:].zero_()
        L.check(lib.adt_to_bf16_t(L.ptr(x), ctypes.c_int64(x.stride(0)), L.ptr(y), ctypes.c_int64(ld), ctypes.c_int32(R), ctypes.c_int32(C), st),
                "adt_to_bf16_t")
        return y, ld
    ld = _ceil8(C)
    y = torch.empty(R, ld, dtype=torch.bfloat16, device=x.device)
    if ld != C:
        y[:, C:].zero_()
    L.check(lib.adt_to_bf16_ld(L.ptr(x), ctypes.c_int64(x.stride(0)), L.ptr(y), ctypes.c_int64(ld), ctypes.c_int64(R), ctypes.c_int32(C), st),
            "adt_to_bf16_ld")
    return y, ld


def _gemm_tc(a16, lda, b16, ldb, c, M, N, K, bias=None, act=0, scale=1.0, pre=None, accumulate=False, a_mn=False, b_mn=False, split_k=1):
    a = L.fill(L.adt_gemm_tc_args(), a_bf16=a16, b_bf16=b16, lda=lda, ldb=ldb, c=c, pre=pre, bias=bias, ldc=c.stride(0), M=M, N=N, K=K,
               act=act, accumulate=int(accumulate), scale=scale, a_mn=int(a_mn), b_mn=int(b_mn), split_k=split_k)
    L.check(L.lib().adt_gemm_tc(ctypes.byref(a), _st(c.device)), "adt_gemm_tc")


def _use_tc(precision, M, N, K):
    return bool(precision) and M >= 128 and M * N * K >= TC_MIN_WORK and torch.cuda.get_device_capability()[0] >= 10


def _attn_scratch(ref, B, Lq, H, nh, precision, backward):
    """scratch of the tcgen05 attention path (bf16 mode, 65..256 positions): packed bf16 operands, scores and probabilities in HBM"""
    if not precision or torch.cuda.get_device_capability()[0] < 10:
        return None
    f = L.lib().adt_attention_scratch_bytes
    f.restype = ctypes.c_int64
    n = int(f(ctypes.c_int32(B), ctypes.c_int32(Lq), ctypes.c_int32(H), ctypes.c_int32(nh), ctypes.c_int32(backward)))
    return torch.empty(n, dtype=torch.uint8, device=ref.device) if n else None


class LinearFn(torch.autograd.Function):
    """y = act((x W^T + b) * scale), x [M,K], W [N,K]  (act: 0 none, 1 relu, 2 gelu)."""

    @staticmethod
    def forward(ctx, x, W, b, act, scale, precision):
        x = x.contiguous()
        M, K = x.shape
        N = W.shape[0]
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        pre = torch.empty_like(y) if act else None
        if _use_tc(precision, M, N, K):   # tcgen05 GEMM on bf16 operand copies (fp32 accumulate, fp32 bias / activation epilogue)
            x16, lda = _bf16(x)
            w16, ldb = _bf16(W)
            _gemm_tc(x16, lda, w16, ldb, y, M, N, K, bias=b, act=act, scale=scale, pre=pre)
            ctx.save_for_backward(x, W, pre if pre is not None else y)
            ctx.cfg = (act, scale, precision, b is not None)
            return y
        for n0 in range(0, N, NBLK):      # wide outputs (the vocabulary head) are produced in column blocks
            n1 = min(N, n0 + NBLK)
            a = L.fill(L.adt_linear_fwd_args(), x=x, w=W[n0:n1], b=b[n0:n1] if b is not None else None, y=y[:, n0:n1],
                       pre=pre[:, n0:n1] if pre is not None else None, M=M, K=K, N=n1 - n0, act=act, scale=scale, precision=precision, ldy=N)
            L.check(L.lib().adt_linear_fwd(ctypes.byref(a), _st(x.device)), "adt_linear_fwd")
        ctx.save_for_backward(x, W, pre if pre is not None else y)
        ctx.cfg = (act, scale, precision, b is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, pre = ctx.saved_tensors
        act, scale, precision, has_b = ctx.cfg
        M, K = x.shape
        N = W.shape[0]
        dy = dy.contiguous()
        lib = L.lib()
        if act:
            dpre = torch.empty_like(dy)
            L.check(lib.adt_act_bwd(L.ptr(dy), L.ptr(pre), L.ptr(dpre), ctypes.c_int64(dy.numel()), ctypes.c_int32(act), _st(x.device)),
                    "adt_act_bwd")
            dy = dpre
        dx = torch.empty_like(x)
        gW = torch.zeros_like(W)
        gb = torch.zeros(N, dtype=torch.float32, device=x.device) if has_b else None
        if _use_tc(precision, M, N, K):
            # dx = scale * dy W ; dW = scale * dy^T x ; db = scale * colsum(dy): three more products on the tensor cores
            # (W, dy and x are read MN-major where the product needs their transpose: no transposed copies)
            dy16, ldd = _bf16(dy)
            w16, ldw = _bf16(W)
            x16, ldx = _bf16(x)
            _gemm_tc(dy16, ldd, w16, ldw, dx, M, K, N, scale=scale, b_mn=True)
            tiles = ((N + 127) // 128) * ((K + 127) // 128)
            _gemm_tc(dy16, ldd, x16, ldx, gW, N, K, M, scale=scale, a_mn=True, b_mn=True, accumulate=True,
                     split_k=max(1, min((M + 63) // 64, SM_COUNT // tiles)))
            if has_b:
                L.check(lib.adt_colsum(L.ptr(dy), ctypes.c_int64(dy.stride(0)), ctypes.c_int32(M), ctypes.c_int32(N), L.ptr(gb), _st(x.device)),
                        "adt_colsum")
                if scale != 1.0:
                    gb.mul_(scale)
            return dx, gW, gb, None, None, None
        for n0 in range(0, N, NBLK):
            n1 = min(N, n0 + NBLK)
            a = L.fill(L.adt_linear_bwd_args(), x=x, w=W[n0:n1], dy=dy[:, n0:n1], dx=dx, g_w=gW[n0:n1], g_b=gb[n0:n1] if has_b else None,
                       M=M, K=K, N=n1 - n0, accumulate_dx=1 if n0 else 0, scale=scale, precision=precision, lddy=N)
            L.check(lib.adt_linear_bwd(ctypes.byref(a), _st(x.device)), "adt_linear_bwd")
        return dx, gW, gb, None, None, None


def linear(x, W, b=None, act=0, scale=1.0, precision=0):
    return LinearFn.apply(x, W, b, act, scale, precision)


class DrlFn(torch.autograd.Function):
    """mode 0: LN(dropout(a) + r) ; mode 1: dropout(LN(a + r)).  r may be None."""

    @staticmethod
    def forward(ctx, a, r, gamma, beta, mode, eps, drop):
        a = a.contiguous()
        r = r.contiguous() if r is not None else None
        M, H = a.shape
        y = torch.empty_like(a)
        args = L.fill(L.adt_drl_args(), a=a, r=r, gamma=gamma, beta=beta, y=y, dy=None, da=None, dr=None, g_gamma=None, g_beta=None, M=M,
                      H=H, mode=mode, eps=eps, drop=drop)
        L.check(L.lib().adt_drop_res_ln_fwd(ctypes.byref(args), _st(a.device)), "adt_drop_res_ln_fwd")
        ctx.save_for_backward(a, r if r is not None else a, gamma, beta)
        ctx.cfg = (mode, eps, drop, r is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        a, r, gamma, beta = ctx.saved_tensors
        mode, eps, drop, has_r = ctx.cfg
        M, H = a.shape
        dy = dy.contiguous()
        da = torch.empty_like(a)
        dr = torch.empty_like(a) if has_r else None
        gg, gb = torch.zeros_like(gamma), torch.zeros_like(beta)
        args = L.fill(L.adt_drl_args(), a=a, r=r if has_r else None, gamma=gamma, beta=beta, y=None, dy=dy, da=da, dr=dr, g_gamma=gg,
                      g_beta=gb, M=M, H=H, mode=mode, eps=eps, drop=drop)
        L.check(L.lib().adt_drop_res_ln_bwd(ctypes.byref(args), _st(a.device)), "adt_drop_res_ln_bwd")
        return da, dr, gg, gb, None, None, None


class AttnFn(torch.autograd.Function):
    """attention core on projected q (pre-scaled), k, v [B*L, H] -> ctx [B*L, H]."""

    @staticmethod
    def forward(ctx, q, k, v, key_ids, dims, drop, training, precision):
        B, Lq, nh, mask_mode = dims
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        H = q.shape[1]
        out = torch.empty_like(q)
        lse = torch.empty(B, nh, Lq, dtype=torch.float32, device=q.device)
        ws = _attn_scratch(q, B, Lq, H, nh, precision, 0)
        a = L.fill(L.adt_attention_args(), q=q, k=k, v=v, ctx=out, lse=lse, key_ids=key_ids, dctx=None, dq=None, dk=None, dv=None, B=B,
                   L=Lq, H=H, nh=nh, mask_mode=mask_mode, training=int(training), drop=drop, precision=precision, tc_scratch=ws)
        L.check(L.lib().adt_attention_fwd(ctypes.byref(a), _st(q.device)), "adt_attention_fwd")
        ctx.save_for_backward(q, k, v, lse)
        ctx.cfg = (key_ids, dims, drop, training, precision)
        return out

    @staticmethod
    def backward(ctx, dctx):
        q, k, v, lse = ctx.saved_tensors
        key_ids, (B, Lq, nh, mask_mode), drop, training, precision = ctx.cfg
        H = q.shape[1]
        dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
        d = drop
        if not training:
            d = no_drop()
        ws = _attn_scratch(q, B, Lq, H, nh, precision, 1)
        a = L.fill(L.adt_attention_args(), q=q, k=k, v=v, ctx=None, lse=lse, key_ids=key_ids, dctx=dctx.contiguous(), dq=dq, dk=dk, dv=dv,
                   B=B, L=Lq, H=H, nh=nh, mask_mode=mask_mode, training=int(training), drop=d, precision=precision, tc_scratch=ws)
        L.check(L.lib().adt_attention_bwd(ctypes.byref(a), _st(q.device)), "adt_attention_bwd")
        return dq, dk, dv, None, None, None, None, None


class Gather3Fn(torch.autograd.Function):
    """s = word[ids] + pos[pos_ids] + sent[sent_ids]; all three tables have padding_idx=0 (no lookup gradient for row 0)."""

    @staticmethod
    def forward(ctx, ids, pos_ids, sent_ids, word, pos, sent):
        B, Lq = ids.shape
        H = word.shape[1]
        s = torch.empty(B * Lq, H, dtype=torch.float32, device=word.device)
        L.check(L.lib().adt_gather3(L.ptr(ids), L.ptr(word), L.ptr(pos_ids), L.ptr(pos), L.ptr(sent_ids), L.ptr(sent), L.ptr(s),
                                    ctypes.c_int32(B * Lq), ctypes.c_int32(H), _st(word.device)), "adt_gather3")
        ctx.ids = (ids, pos_ids, sent_ids)
        ctx.shapes = (word.shape, pos.shape, sent.shape)
        return s

    @staticmethod
    def backward(ctx, ds):
        ids, pos_ids, sent_ids = ctx.ids
        ws, ps, ss = ctx.shapes
        B, Lq = ids.shape
        H = ws[1]
        dev = ds.device
        ds = ds.contiguous()
        dW = torch.zeros(ws, device=dev)
        _scatter(dev, B * Lq, B, Lq, H, ws[0] - 1, seq=ids, dx_enc=ds, dE=dW, dP=None, emb_scale=1.0)
        dP, dS = torch.zeros(ps, device=dev), torch.zeros(ss, device=dev)
        lib = L.lib()
        for idx, g in ((pos_ids, dP), (sent_ids, dS)):
            L.check(lib.adt_small_table_grad(L.ptr(idx), L.ptr(ds), L.ptr(g), ctypes.c_int32(B * Lq), ctypes.c_int32(H), ctypes.c_int32(0),
                                             _st(dev)), "adt_small_table_grad")
        return None, None, None, dW, dP, dS
