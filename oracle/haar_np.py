"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity unpinned.

numpy restatement of the 2-D Haar analysis/synthesis the reference reaches through
``pytorch_wavelets`` (call sites: diff_cifar/model.py:263-267 and :310-321,
diff_cifar/diffusion.py:63-70, pdearena/pdearena/modules/twod_unetbase.py:169-193,
wmh/model.py:68-95), written with explicit 2x2 butterflies instead of
convolutions so that it is an *independent* second statement next to
``pytorch_wavelets_restated``:

    block [[a, b], [c, d]]   (a, b on the upper row)
    lo_t = s*a + s*b   hi_t = s*a - s*b        (W axis first, as upstream does)
    lo_b = s*c + s*d   hi_b = s*c - s*d
    LL = s*lo_t + s*lo_b      LH = s*lo_t - s*lo_b   (W-low , H-high)
    HL = s*hi_t + s*hi_b      HH = s*hi_t - s*hi_b   (W-high, H-low / H-high)

All arithmetic is float32 with ``s = float32(1/sqrt 2)``; an odd extent gets one zero
appended at its END (``mode='zero'``), so output extents are ceil(n/2).
"""
from __future__ import annotations

import numpy as np

S = np.float32(1.0 / np.sqrt(2.0))

# The ONE unpinned convention of this path: where mode='zero' puts the single zero that makes an odd
# extent even.  Upstream pytorch_wavelets appends it at the END; the reference only pins the output extent ceil(n/2)
# (wmh/model.py:146-155: 25 -> 13).  Flip here and set UB200_HAAR_PAD_AT_START=1 for the kernels (csrc/haar.cu
# `pad_at_start()`) to move it to the start; only odd extents (config 5's 25 -> 13, odd sweep entries) change.
PAD_AT_END = True


def _pad_even(x: np.ndarray) -> np.ndarray:
    h, w = x.shape[-2:]
    if h % 2 or w % 2:
        pad_h, pad_w = ((0, h % 2), (0, w % 2)) if PAD_AT_END else ((h % 2, 0), (w % 2, 0))
        x = np.pad(x, [(0, 0)] * (x.ndim - 2) + [pad_h, pad_w])
    return x


def _crop(x: np.ndarray, h: int, w: int) -> np.ndarray:
    """Undo the zero extension: drop the row / column that was added (end or start)."""
    if PAD_AT_END:
        return x[..., :h, :w]
    return x[..., x.shape[-2] - h:, x.shape[-1] - w:]


def dwt2_level(x: np.ndarray):
    """One analysis level.  x [..., H, W] float32 -> (LL, LH, HL, HH) each [..., ceil(H/2), ceil(W/2)]."""
    x = _pad_even(np.asarray(x, dtype=np.float32))
    a, b = x[..., 0::2, 0::2], x[..., 0::2, 1::2]
    c, d = x[..., 1::2, 0::2], x[..., 1::2, 1::2]
    lo_t, hi_t = S * a + S * b, S * a - S * b
    lo_b, hi_b = S * c + S * d, S * c - S * d
    return (S * lo_t + S * lo_b, S * lo_t - S * lo_b, S * hi_t + S * hi_b, S * hi_t - S * hi_b)


def dwt2(x: np.ndarray, J: int):
    """J-level analysis.  Returns (Yl, [Yh_1..Yh_J]) with Yh_j [N, C, 3, h_j, w_j], finest first."""
    ll = np.asarray(x, dtype=np.float32)
    highs = []
    for _ in range(J):
        ll, lh, hl, hh = dwt2_level(ll)
        highs.append(np.stack([lh, hl, hh], axis=2))
    return ll, highs


def idwt2_level(ll, lh, hl, hh) -> np.ndarray:
    """One synthesis level; exact inverse of `dwt2_level` on even extents."""
    ll, lh, hl, hh = (np.asarray(t, dtype=np.float32) for t in (ll, lh, hl, hh))
    lo_t, lo_b = S * ll + S * lh, S * ll - S * lh      # undo H
    hi_t, hi_b = S * hl + S * hh, S * hl - S * hh
    out = np.empty(ll.shape[:-2] + (2 * ll.shape[-2], 2 * ll.shape[-1]), dtype=np.float32)
    out[..., 0::2, 0::2] = S * lo_t + S * hi_t         # undo W
    out[..., 0::2, 1::2] = S * lo_t - S * hi_t
    out[..., 1::2, 0::2] = S * lo_b + S * hi_b
    out[..., 1::2, 1::2] = S * lo_b - S * hi_b
    return out


def idwt2(yl: np.ndarray, highs) -> np.ndarray:
    """Synthesis from (Yl, [Yh_1..Yh_J]); an empty list returns Yl (the reference's only use)."""
    ll = np.asarray(yl, dtype=np.float32)
    for band in highs[::-1]:
        ll = _crop(ll, min(ll.shape[-2], band.shape[-2]), min(ll.shape[-1], band.shape[-1]))
        ll = idwt2_level(ll, band[:, :, 0], band[:, :, 1], band[:, :, 2])
    return ll


def channel_tile(x: np.ndarray, out_channels: int) -> np.ndarray:
    """`x.repeat(1, out//C + 1, 1, 1)[:, :out]`  (diff_cifar/model.py:275, :319)."""
    c = x.shape[1]
    return np.ascontiguousarray(x[:, np.arange(out_channels) % c])


def dwtblock(x: np.ndarray, J: int, out_channels: int | None, tile_if_equal: bool = True) -> np.ndarray:
    """DTWBlock / DWTBlock forward (diff_cifar/model.py:270-323; twod_unetbase.py:173-193).

    J == 0: channel tile only.  J > 0: LL_J / 2**J then channel tile.  pdearena / wmh skip the
    tile when C == out (`tile_if_equal=False`).
    """
    x = np.asarray(x, dtype=np.float32)
    if J > 0:
        x, _ = dwt2(x, J)
        x = x / np.float32(2.0 ** J)
    if out_channels is None or (not tile_if_equal and out_channels == x.shape[1]):
        return x
    return channel_tile(x, out_channels)


def dwtblock_bwd(grad_out: np.ndarray, in_shape, J: int) -> np.ndarray:
    """Adjoint of `dwtblock`: fold the tiled channels back (sum), then spread LL^T / 2**J."""
    n, c, h, w = in_shape
    g = np.asarray(grad_out, dtype=np.float32)
    folded = np.zeros((n, c) + g.shape[2:], dtype=np.float32)
    for k in range(g.shape[1]):
        folded[:, k % c] += g[:, k]
    if J == 0:
        return folded
    cur = folded / np.float32(2.0 ** J)
    # extents of every intermediate level, finest first
    ext = [(h, w)]
    for _ in range(J):
        ext.append(((ext[-1][0] + 1) // 2, (ext[-1][1] + 1) // 2))
    for lvl in range(J, 0, -1):
        z = np.zeros_like(cur)
        cur = _crop(idwt2_level(cur, z, z, z), ext[lvl - 1][0], ext[lvl - 1][1])
    return cur
