"""SuperSASRecModel on B200: the weight-sharing supernet of ADT's evolutionary lambda search.

Mirrors /root/reference/sasrec/supersasrec.py:10-108 (constructor, forward/predict/set_choice, NO last LayerNorm),
/root/reference/sasrec/super_modules.py:11-85 (rec_size*ind_size candidate blocks per layer, 4 active per layer,
outputs blended with bilinear weights; rec head = log_softmax of the blended log-softmaxes; decoder blends outputs)
and /root/reference/sasrec/base_super_modules.py:15-55 (choice -> 4 block indices + weights, including the
`rec_size` stride quirk B9).  Parameter names match the reference (`encoder.encoder_layers.{l}.{c}.…`), so
`./checkpoint/super.pth` files interchange.  Every device op is a libadt_b200.so kernel (blocks.py); the blending
of the 4 candidate outputs is plain elementwise torch on the results.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from .blocks import DEC_KEYS, ENC_KEYS, DecBlockFn, DropCfg, EmbedFn, EncBlockFn, LogitsFn, layer_params
from .model import _as_ids, _DecoderLayer, _EncoderLayer, score_candidates


def get_position(weight, choice):
    """base_super_modules.py:15-19: bracket `weight` on the grid -> (i0, i1, p0, 1-p0).  Raises IndexError when
    weight >= max(choice), like the reference."""
    choice = np.asarray(choice, dtype=np.float64)
    i1 = int(np.where(choice > weight)[0][0])
    i0 = i1 - 1
    p0 = (weight - choice[i0]) / (choice[i1] - choice[i0])
    return i0, i1, p0, 1 - p0


def get_shared(cand, rec_choice, ind_choice):
    """base_super_modules.py:21-40 -> ([4 block indices per layer], [4 blend weights per layer])."""
    rec_size = len(rec_choice)
    idxs, weights = [], []
    for i in range(len(cand) // 2):
        i0, i1, p0, p1 = get_position(cand[2 * i], rec_choice)
        i2, i3, p2, p3 = get_position(cand[2 * i + 1], ind_choice)
        idxs.append((i0 * rec_size + i2, i1 * rec_size + i2, i0 * rec_size + i3, i1 * rec_size + i3))
        weights.append((p1 * p3, p0 * p3, p1 * p2, p0 * p2))
    return idxs, weights


class _SuperStack(nn.Module):
    def __init__(self, kind, num_layers, H, nh, rec_choice, ind_choice):
        super().__init__()
        self.rec_choice, self.ind_choice = rec_choice, ind_choice
        self.rec_size, self.ind_size = len(rec_choice), len(ind_choice)
        self.choice_block_size = self.rec_size * self.ind_size
        mk = _EncoderLayer if kind == "enc" else _DecoderLayer
        layers = nn.ModuleList(nn.ModuleList(mk(H, nh) for _ in range(self.choice_block_size)) for _ in range(num_layers))
        setattr(self, "encoder_layers" if kind == "enc" else "decoder_layers", layers)
        self.shared_idx = [[0, 0, 0, 0] for _ in range(num_layers)]
        self.shared_weights = [[0, 0, 0, 0] for _ in range(num_layers)]

    def set_choice(self, cand):
        self.shared_idx, self.shared_weights = get_shared(cand, self.rec_choice, self.ind_choice)


class SuperSASRecModel(nn.Module):
    def __init__(self, usernum, itemnum, rec_choice, ind_choice, args):
        super().__init__()
        self.usernum, self.itemnum = usernum, itemnum
        self.item_num = itemnum
        self.dev = args.device
        self.num_heads, self.maxlen, self.num_layers = args.num_heads, args.maxlen, args.num_layers
        self.hidden, self.dropout_p = args.hidden_units, float(args.dropout)
        self.choice = np.zeros(args.num_layers)
        H = args.hidden_units
        self.item_emb = nn.Embedding(itemnum + 1, H, padding_idx=0)
        self.pos_emb = nn.Embedding(args.maxlen, H)
        self.emb_dropout = nn.Dropout(p=args.dropout)
        self.encoder = _SuperStack("enc", args.num_layers, H, args.num_heads, rec_choice, ind_choice)
        self.decoder = _SuperStack("dec", args.num_layers, H, args.num_heads, rec_choice, ind_choice)
        self.drop_seed, self.drop_step = 0, 0
        self.precision = 0     # GEMM cores of the blocks: 0 fp32 (parity mode), 1 bf16 tensor cores (evaluation / fitness passes)
        L.lib()

    def set_choice(self, cand):
        """supersasrec.py:106-108.  cand = [rec_0, ind_0, rec_1, ind_1, ...] (lambda values, not [0,1] candidates)."""
        self.encoder.set_choice(cand)
        self.decoder.set_choice(cand)

    def _check(self):
        if self.item_emb.weight.device.type != "cuda":
            raise L.AdtError("adt_b200.SuperSASRecModel runs on CUDA only (no CPU fallback): call .to('cuda') first")

    def log2feats(self, seq, drop):
        """supersasrec.py:43-60 on int32 device ids -> (feats [M,H], enc_inputs, rec_true list)."""
        B, Lq = seq.shape
        H, nh = self.hidden, self.num_heads
        x = EmbedFn.apply(seq, self.item_emb.weight, self.pos_emb.weight, drop.next("row", nh, Lq, H))
        enc_inputs, recs = [], []
        for layer, idxs, weights in zip(self.encoder.encoder_layers, self.encoder.shared_idx, self.encoder.shared_weights):
            enc_inputs.append(x)
            out = rec = None
            for idx, w in zip(idxs, weights):
                cfg = (nh, drop.training, drop.next("attn", nh, Lq, H), drop.next("row", nh, Lq, H), drop.next("row", nh, Lq, H), self.precision)
                o, r = EncBlockFn.apply(x, seq, cfg, *layer_params(layer[idx], ENC_KEYS))
                out = o * w if out is None else out + o * w
                rec = r * w if rec is None else rec + r * w
            x = out
            recs.append(torch.log_softmax(rec, dim=-1))   # super_modules.py:49 (log_softmax applied a second time, quirk B10)
        return x, enc_inputs, recs

    def forward(self, user_ids, log_seqs, dec_seqs, pos_seqs, neg_seqs):
        """supersasrec.py:78-90 -> (pos_logits, neg_logits, enc_inputs, dec_outputs reversed, rec_ind)."""
        self._check()
        dev = self.item_emb.weight.device
        seq, dec, pos, neg = (_as_ids(a, dev) for a in (log_seqs, dec_seqs, pos_seqs, neg_seqs))
        B, Lq = seq.shape
        H, nh = self.hidden, self.num_heads
        drop = DropCfg(self.dropout_p, self.drop_seed, self.drop_step, self.training)
        x, enc_inputs, recs = self.log2feats(seq, drop)
        feats, pl, nl = LogitsFn.apply(x, self.item_emb.weight, pos, neg, None, None)
        xd = EmbedFn.apply(dec, self.item_emb.weight, self.pos_emb.weight, drop.next("row", nh, Lq, H))
        dec_outs = []
        for layer, idxs, weights in zip(self.decoder.decoder_layers, self.decoder.shared_idx, self.decoder.shared_weights):
            out = None
            for idx, w in zip(idxs, weights):
                cfg = (nh, drop.training, drop.next("attn", nh, Lq, H), drop.next("attn", nh, Lq, H), drop.next("row", nh, Lq, H),
                       drop.next("row", nh, Lq, H), self.precision)
                o = DecBlockFn.apply(xd, feats, dec, cfg, *layer_params(layer[idx], DEC_KEYS))
                out = o * w if out is None else out + o * w
            xd = out
            dec_outs.append(xd.view(B, Lq, H))
        dec_outs.reverse()
        self.drop_step += 1 if self.training else 0
        # modules.py:517-518 mis-view: a pure row permutation of the true [B,L] rows
        rec_ind = [r.view(B, Lq, nh, nh).transpose(0, 1).contiguous().view(B, Lq, nh, nh) for r in recs]
        return pl.view(B, Lq), nl.view(B, Lq), [e.view(B, Lq, H) for e in enc_inputs], dec_outs, rec_ind

    @torch.no_grad()
    def predict(self, user_ids, log_seqs, item_indices, full=False):
        """supersasrec.py:92-104."""
        self._check()
        dev = self.item_emb.weight.device
        seq = _as_ids(log_seqs, dev)
        B, Lq = seq.shape
        x, _, _ = self.log2feats(seq, DropCfg(0.0, 0, 0, False))
        final = x.view(B, Lq, self.hidden)[:, -1, :].contiguous()
        return score_candidates(self.item_emb.weight, final, item_indices, full)
