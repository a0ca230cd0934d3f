"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path.

CPU restatement (plain PyTorch) of the reference's Bert4Rec-ADT hot path.  Pinned against the UNMODIFIED reference
(/root/reference/bert4rec/model/bert.py run by oracle/make_golden_bert.py) through tests/golden/bert_*.npz.

Reference map (file:line under /root/reference/bert4rec):
  embed()          model/modules.py:42-48   (word + pos + sent -> LayerNorm(1e-5) -> dropout)
  mha()            model/modules.py:74-102  (separate q/k/v/out Linear, key-padding mask -> -1e9, dropout on weights)
  drl()            model/modules.py:112-117 (LayerNorm(dropout(out) + prev))
  encoder_layer()  model/modules.py:166-184 ; decoder_layer() :298-325
  downstream()     model/bert.py:80-90      (Linear -> GELU -> LayerNorm -> tied word-embedding logits + bias)
  forward()        model/bert.py:92-108 ; predict() :110-116
  loss()           trainer.py:112-128       (CE ignore_index=0 + lambda1*MSE + lambda2[l]*NLL)
Dropout sites follow the reference's call order: emb ; per encoder layer (attn weights, after-multi, final) ; decoder
emb ; per decoder layer (self weights, after-multi, cross weights, after-src-dec, final).
"""
import math
import torch
import torch.nn.functional as F

from .sasrec_oracle import Drop


class BCfg:
    def __init__(self, item_num, maxlen, hidden, heads, layers, inner, dropout=0.0, attention_dropout=0.0, vocab_extra=100):
        self.item_num, self.maxlen, self.hidden, self.heads, self.layers, self.inner = item_num, maxlen, hidden, heads, layers, inner
        self.dropout, self.attention_dropout, self.vocab = dropout, attention_dropout, item_num + vocab_extra


class _Sites:
    def __init__(self):
        self.k = 0

    def next(self, active):
        s = self.k
        if active:
            self.k += 1
        return s


def _drop(x, drop, p, sites):
    """dropout with probability p at the next site (site counter only advances when the call is active, like the injector)."""
    active = drop.train and p > 0.0
    site = sites.next(active)
    if not active:
        return x
    d = Drop(p, drop.seed, drop.step, drop.b0, True)
    return d.apply(x, site)


def embed(sd, ids, pos_ids, sent_ids, cfg, drop, sites):
    # all three tables are nn.Embedding(padding_idx=0) (modules.py:14-32): row 0 is looked up but receives no lookup gradient
    s = F.embedding(ids, sd["item_emb.word_emb.weight"], padding_idx=0) + F.embedding(pos_ids, sd["item_emb.pos_emb.weight"], padding_idx=0) \
        + F.embedding(sent_ids, sd["item_emb.sent_emb.weight"], padding_idx=0)
    s = F.layer_norm(s, (cfg.hidden,), sd["item_emb.layer_norm.weight"], sd["item_emb.layer_norm.bias"], 1e-5)
    return _drop(s, drop, cfg.dropout, sites)


def mha(sd, pre, q_in, kv_in, key_mask, cfg, drop, sites):
    B, L, H = q_in.shape
    nh, dk = cfg.heads, H // cfg.heads
    q = F.linear(q_in, sd[pre + "query_transfer.weight"], sd[pre + "query_transfer.bias"]).view(B, L, nh, dk).transpose(1, 2)
    k = F.linear(kv_in, sd[pre + "key_transfer.weight"], sd[pre + "key_transfer.bias"]).view(B, L, nh, dk).transpose(1, 2)
    v = F.linear(kv_in, sd[pre + "value_transfer.weight"], sd[pre + "value_transfer.bias"]).view(B, L, nh, dk).transpose(1, 2)
    w = q @ k.transpose(-2, -1) / math.sqrt(dk)
    w = w.masked_fill(key_mask[:, None, None, :] == 0, -1e9)
    w = torch.softmax(w, dim=-1)
    w = _drop(w, drop, cfg.attention_dropout, sites)
    ctx = (w @ v).transpose(1, 2).reshape(B, L, H)
    return F.linear(ctx, sd[pre + "out_transfer.weight"], sd[pre + "out_transfer.bias"]), ctx.view(B, L, nh, dk)


def drl(sd, pre, out, prev, cfg, drop, sites):
    y = _drop(out, drop, cfg.attention_dropout, sites) + prev
    return F.layer_norm(y, (cfg.hidden,), sd[pre + "layer_norm.weight"], sd[pre + "layer_norm.bias"], 1e-5)


def ffn(sd, pre, x):
    return F.linear(F.gelu(F.linear(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"])), sd[pre + "fc2.weight"], sd[pre + "fc2.bias"])


def forward(sd, cfg, src, dec, drop=None, with_logits=True):
    """bert.py:92-108 with pos ids 0..L-1 and sentence ids 0 (trainer.py:106-109)."""
    drop = drop or Drop(train=False)
    sites = _Sites()
    B, L = src.shape
    pos_ids = torch.arange(L).repeat(B, 1)
    sent_ids = torch.zeros_like(src)
    x = embed(sd, src, pos_ids, sent_ids, cfg, drop, sites)
    src_mask = (src > 0)
    enc_inputs, inds = [], []
    for l in range(cfg.layers):
        pre = f"encoder.encoder_layers.{l}."
        enc_inputs.append(x)
        m, ctx = mha(sd, pre + "multi_head_attention.", x, x, src_mask, cfg, drop, sites)
        h = drl(sd, pre + "drop_residual_normalize_layer_after_multi.", m, x, cfg, drop, sites)
        x = drl(sd, pre + "drop_residual_normalize_layer_final.", ffn(sd, pre + "ffn.", h), h, cfg, drop, sites)
        inds.append(F.log_softmax(F.linear(ctx, sd[pre + "head_classifier.weight"], sd[pre + "head_classifier.bias"]), dim=3))
    feats = x
    y = embed(sd, dec, pos_ids, sent_ids, cfg, drop, sites)
    dec_mask = (dec > 0)
    dec_outs = []
    for l in range(cfg.layers):
        pre = f"decoder.decoder_layers.{l}."
        m, _ = mha(sd, pre + "dec_multi_head_attention.", y, y, dec_mask, cfg, drop, sites)
        y = drl(sd, pre + "drop_residual_normalize_layer_after_multi.", m, y, cfg, drop, sites)
        m, _ = mha(sd, pre + "src_dec_attention.", y, feats, src_mask, cfg, drop, sites)
        y = drl(sd, pre + "drop_residual_normalize_layer_after_src_dec.", m, y, cfg, drop, sites)
        y = drl(sd, pre + "drop_residual_normalize_layer_final.", ffn(sd, pre + "ffn.", y), y, cfg, drop, sites)
        dec_outs.append(y)
    dec_outs.reverse()
    out = {"feats": feats, "enc_inputs": enc_inputs, "dec_outputs": dec_outs, "ind_outputs": inds}
    if with_logits:
        out["logits"] = downstream(sd, cfg, feats)
    return out


def downstream(sd, cfg, feats):
    h = F.gelu(F.linear(feats, sd["mask_trans_feat.weight"], sd["mask_trans_feat.bias"]))
    h = F.layer_norm(h, (cfg.hidden,), sd["mask_layer_norm.weight"], sd["mask_layer_norm.bias"], 1e-5)
    return h @ sd["item_emb.word_emb.weight"].t() + sd["mask_bias"]


def loss(cfg, out, labels, lambda1, lambda2):
    """trainer.py:112-128."""
    logits = out["logits"]
    total = F.cross_entropy(logits.view(-1, logits.size(-1)), labels.view(-1), ignore_index=0)
    for i in range(cfg.layers):
        if lambda1[i] != 0:
            total = total + lambda1[i] * F.mse_loss(out["enc_inputs"][i], out["dec_outputs"][i])
    if cfg.heads > 1:
        B, L = labels.shape
        label = torch.arange(cfg.heads).repeat(B * L, 1)
        for l in range(cfg.layers):
            if lambda2[l] != 0:
                total = total + lambda2[l] * F.nll_loss(out["ind_outputs"][l].reshape(B * L, cfg.heads, cfg.heads), label)
    return total


@torch.no_grad()
def predict(sd, cfg, seqs, candidates):
    """bert.py:110-116: head on all positions, last position, gather candidates."""
    out = forward(sd, cfg, seqs, seqs, Drop(train=False))
    return out["logits"][:, -1, :].gather(1, candidates)
