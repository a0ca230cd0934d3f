"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path.

Counter-based dropout mask generator shared (by construction) with the CUDA
kernels (adt_b200/csrc/philox.cuh).  The reference draws its dropout masks from
torch's global RNG stream (`F.dropout`, /root/reference/sasrec/modules.py:61,
`nn.Dropout` in /root/reference/sasrec/model.py:20 and modules.py:626,628),
which cannot be reproduced on a GPU.  For parity both sides therefore draw the
mask from Philox4x32-10 keyed by (seed, step) with counter (element>>2, site):

    keep(site, idx) = philox(ctr=(lo32(idx>>2), hi32(idx>>2), site, step),
                             key=(lo32(seed), hi32(seed)))[idx & 3] >= thr
    thr = round(p * 2**32)          (so P[keep] = 1-p)

`idx` is the element's linear index in OUR natural layout
  * [B,L,H] sites        : ((b0+b)*L + t)*H + c          (32-bit lanes, formula above)
  * attention-prob sites : row r = ((b0+b)*nh + h)*L + i, key j, padded row stride Lp = ceil8(L):
        e = r*Lp + j ; 16-bit lane (e & 7) of philox(ctr=(e>>3, site, step)) ; keep iff rnd16 >= round(p*2**16)
    (one Philox call per 8 probabilities, always aligned -> cheap inside the softmax loop)
The oracle permutes the mask into whatever layout the reference has at that
call site.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. All inputs uint32 arrays (broadcastable)."""
    c0 = np.asarray(c0, dtype=np.uint32)
    shape = np.broadcast(c0, c1, c2, c3).shape
    c0 = np.broadcast_to(c0, shape).astype(np.uint32)
    c1 = np.broadcast_to(np.asarray(c1, dtype=np.uint32), shape).astype(np.uint32)
    c2 = np.broadcast_to(np.asarray(c2, dtype=np.uint32), shape).astype(np.uint32)
    c3 = np.broadcast_to(np.asarray(c3, dtype=np.uint32), shape).astype(np.uint32)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def threshold(p):
    """uint32 threshold: element is KEPT iff random >= threshold."""
    t = int(round(float(p) * 4294967296.0))
    return min(max(t, 0), 0xFFFFFFFF)


def keep_mask(n, p, seed, step, site, offset=0):
    """Boolean keep-mask for linear element indices offset..offset+n-1."""
    idx = np.arange(offset, offset + n, dtype=np.uint64)
    q = idx >> np.uint64(2)
    lane = (idx & np.uint64(3)).astype(np.int64)
    r = philox4x32_10((q & _MASK32).astype(np.uint32), (q >> np.uint64(32)).astype(np.uint32),
                      np.uint32(site), np.uint32(step),
                      np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF))
    r = np.stack(r, axis=-1)
    rnd = np.take_along_axis(r, lane[:, None], axis=1)[:, 0]
    return rnd >= np.uint32(threshold(p))


def keep_mask_attn(rows, L, p, seed, step, site, row_offset=0):
    """Boolean keep-mask [rows, L] of an attention-probability site (16-bit lanes, padded row stride)."""
    Lp = (L + 7) & ~7
    r = np.arange(row_offset, row_offset + rows, dtype=np.uint64)[:, None]
    e = r * np.uint64(Lp) + np.arange(L, dtype=np.uint64)[None, :]
    q = e >> np.uint64(3)
    lane = (e & np.uint64(7)).astype(np.int64)
    out = philox4x32_10((q & _MASK32).astype(np.uint32), (q >> np.uint64(32)).astype(np.uint32), np.uint32(site), np.uint32(step),
                        np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF))
    words = np.stack(out, axis=-1)                                    # [rows, L, 4]
    w = np.take_along_axis(words, (lane >> 1)[..., None], axis=-1)[..., 0]
    rnd16 = (w >> ((lane & 1) * 16).astype(np.uint32)) & np.uint32(0xFFFF)
    thr = min(max(int(round(float(p) * 65536.0)), 0), 0xFFFF)
    return rnd16 >= np.uint32(thr)


if __name__ == "__main__":
    # Known-answer test from the Random123 distribution (philox4x32-10):
    # ctr = key = 0  ->  6627e8d5 e169c58d bc57ac4c 9b00dbd8
    out = philox4x32_10(np.uint32([0]), 0, 0, 0, 0, 0)
    print([hex(int(x[0])) for x in out])
    out = philox4x32_10(np.uint32([0xFFFFFFFF]), 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
    print([hex(int(x[0])) for x in out])
