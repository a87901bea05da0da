"""TEST INFRASTRUCTURE (oracle) -- CPU statement of the counter-based random streams.

The reference draws its randomness from torch's global generator:
  * dropout masks          nn.Dropout in the head towers   (probabilistic_retinanet.py:422-424)
  * logit noise            Normal.rsample((S,))            (probabilistic_inference.py:291-294)
  * box-delta noise        MultivariateNormal.rsample((1000,))  (probabilistic_inference.py:351-356)
A GPU path cannot reproduce torch's CPU generator, so parity is defined on a
counter-based stream (Philox4x32-10, Salmon et al. SC'11) that the oracle injects
into the reference (oracle/ref_runner.py) and that the CUDA kernels evaluate
in-register (pod_compare_b200/csrc/philox.cuh).  This file is the CPU side of
that contract; only tests/, bench.py's cpu_baseline leg and smoke() import it.

Stream layout (key = (seed_lo, seed_hi ^ STREAM), counter = (c0, c1, c2, c3)):
  STREAM_DROPOUT : c0 = (pixel*C + channel)//8, c1 = level | layer<<8 | tower<<16 | pass<<24,
                   c2 = sample, c3 = image; 16 random bits per decision: lane i (0..7) = bits [16*(i%2), +16) of word
                   i//2 -> element 8*c0 + i; keep iff lane >= floor(p*2^16)  (P(keep) within 1.5e-5 of 1 - p)
  STREAM_LOGIT   : c0 = (anchor*K + k)//4 (row-major over (HWA, K) of one level), c1 = level | run<<8,
                   c2 = draw j, c3 = image; 4 words -> 4 normals (two Box-Muller pairs)
  STREAM_BOX     : c0 = global anchor id (level offset + index in level), c1 = run,
                   c2 = draw j, c3 = image; 4 words -> eps[j, m, 0..3]
`run` is 0 except in the post-NMS merge modes, where every MC sample / ensemble member is its own
inference with its own draws (probabilistic_inference.py:444-461,506-511).
Uniforms use the top 23 bits, u = ((w >> 9) + 0.5) * 2^-23, which is exact in fp32.
Normals: r = sqrt(-2 ln u_a), n_a = r cos(2 pi u_b), n_b = r sin(2 pi u_b); the CPU
side evaluates this in float64 and rounds once to fp32 (the GPU evaluates in fp32,
a few ulp away -- covered by the stated tolerances).
"""
import numpy as np

STREAM_DROPOUT = 0x0D120F01
STREAM_LOGIT = 0x0D120F02
STREAM_BOX = 0x0D120F03

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32-`rounds`. Counters broadcast; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(
        *[np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3)])
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def _key(seed, stream):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, ((seed >> 32) ^ stream) & 0xFFFFFFFF


def u23(w):
    """uint32 word -> uniform in (0,1), exact in fp32."""
    return ((w >> np.uint32(9)).astype(np.float64) + 0.5) * (2.0 ** -23)


def box_muller(wa, wb):
    ua, ub = u23(wa), u23(wb)
    r = np.sqrt(-2.0 * np.log(ua))
    t = 2.0 * np.pi * ub
    return (r * np.cos(t)).astype(np.float32), (r * np.sin(t)).astype(np.float32)


def dropout_threshold(p):
    """16-bit lane threshold: keep iff lane >= floor(p * 2^16)."""
    return min(int(np.floor(float(p) * 65536.0)), 65535)


def dropout_keep_mask(seed, image, sample, pass_, tower, layer, level, H, W, C, p):
    """Boolean keep-mask of shape (H, W, C) (channels-last element order)."""
    assert C % 8 == 0
    k0, k1 = _key(seed, STREAM_DROPOUT)
    n = H * W * C // 8
    c1 = (level & 0xFF) | ((layer & 0xFF) << 8) | ((tower & 0xFF) << 16) | ((pass_ & 0xFF) << 24)
    w = philox4x32(np.arange(n, dtype=np.uint64), c1, sample, image, k0, k1)
    words = np.stack(w, axis=1)                                              # (n, 4) uint32
    lanes = np.stack([words & np.uint32(0xFFFF), words >> np.uint32(16)], axis=2)   # (n, 4, 2): low half first
    return lanes.reshape(H, W, C) >= np.uint32(dropout_threshold(p))


def logit_normals(seed, image, level, draws, n_anchor, K, run=0):
    """fp32 array (draws, n_anchor, K) -- the eps of Normal.rsample((draws,))."""
    k0, k1 = _key(seed, STREAM_LOGIT)
    n = n_anchor * K
    nq = (n + 3) // 4
    q = np.arange(nq, dtype=np.uint64)[None, :]
    j = np.arange(draws, dtype=np.uint64)[:, None]
    w0, w1, w2, w3 = philox4x32(q, (level & 0xFF) | ((run & 0xFFFFFF) << 8), j, image, k0, k1)
    n0, n1 = box_muller(w0, w1)
    n2, n3 = box_muller(w2, w3)
    out = np.stack([n0, n1, n2, n3], axis=2).reshape(draws, nq * 4)[:, :n]
    return np.ascontiguousarray(out.reshape(draws, n_anchor, K))


def box_normals(seed, image, anchor_ids, draws, run=0):
    """fp32 array (draws, M, 4) -- the eps of MultivariateNormal.rsample((draws,))."""
    k0, k1 = _key(seed, STREAM_BOX)
    a = np.asarray(anchor_ids, dtype=np.uint64)[None, :]
    j = np.arange(draws, dtype=np.uint64)[:, None]
    w0, w1, w2, w3 = philox4x32(a, run, j, image, k0, k1)
    n0, n1 = box_muller(w0, w1)
    n2, n3 = box_muller(w2, w3)
    return np.ascontiguousarray(np.stack([n0, n1, n2, n3], axis=2))


# Known-answer vectors of Philox4x32-10 from the Random123 distribution (kat_vectors):
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]
