"""Deterministic synthetic corpora: S0 "photo", S1 "graphic", S2 "noise".
Image i of a corpus uses seed 0x5EED0000 + i.  Test / bench infrastructure."""
from __future__ import annotations

import zlib

import numpy as np

SEED0 = 0x5EED0000


def photo(w: int, h: int, seed: int, sixteen: bool = False) -> np.ndarray:
    """S0: smooth sinusoids + gaussian noise per colour channel, opaque alpha.
    returns uint8 array (h, w, 4) or, for 16-bit, (h, w, 8) big-endian sample bytes."""
    rng = np.random.Generator(np.random.PCG64(seed))
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.empty((h, w, 4), dtype=np.float32)
    for c in range(3):
        fx, fy, gx, gy = rng.integers(1, 8, size=4)
        v = 128 + 64 * np.sin(2 * np.pi * (x * fx + y * fy) / w) + 48 * np.sin(2 * np.pi * (x * gx - y * gy) / h)
        v += rng.normal(0, 6, size=(h, w)).astype(np.float32)
        out[..., c] = v
    out[..., 3] = 255
    u8 = np.clip(np.rint(out), 0, 255).astype(np.uint8)
    if not sixteen:
        return u8
    lo = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    lo[..., 3] = 255
    be = np.empty((h, w, 8), dtype=np.uint8)
    be[..., 0::2] = u8  # x257 + low byte: high byte = value, low byte = value + noise (mod 256)
    be[..., 1::2] = (u8.astype(np.uint16) + lo).astype(np.uint8)
    be[..., 7] = 255
    return be


def graphic(w: int, h: int, seed: int) -> np.ndarray:
    """S1: 256 random flat rectangles, painter's order."""
    rng = np.random.Generator(np.random.PCG64(seed))
    img = np.zeros((h, w, 4), dtype=np.uint8)
    img[...] = rng.integers(0, 256, size=4, dtype=np.uint8)
    for _ in range(256):
        x0, x1 = sorted(rng.integers(0, w + 1, size=2))
        y0, y1 = sorted(rng.integers(0, h + 1, size=2))
        img[y0:y1, x0:x1] = rng.integers(0, 256, size=4, dtype=np.uint8)
    return img


def noise(w: int, h: int, seed: int) -> np.ndarray:
    """S2: incompressible."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)


def make(kind: str, w: int, h: int, index: int, sixteen: bool = False) -> np.ndarray:
    seed = SEED0 + index
    if kind == "photo":
        return photo(w, h, seed, sixteen)
    if kind == "graphic":
        return graphic(w, h, seed)
    if kind == "noise":
        return noise(w, h, seed)
    raise ValueError(kind)


def filter_rows_numpy(storage: np.ndarray, bpp: int) -> bytes:
    """The reference's filter rule (min sum|int8|, first minimum in order 0..4) in numpy:
    an independent cross-check of the oracle's orc_png_filter, fast enough for big images.
    storage: (h, pitch) uint8."""
    h, pitch = storage.shape
    cur = storage.astype(np.int16)
    prev = np.zeros_like(cur)
    prev[1:] = cur[:-1]
    a = np.zeros_like(cur)
    a[:, bpp:] = cur[:, :-bpp]
    c = np.zeros_like(cur)
    c[:, bpp:] = prev[:, :-bpp]
    pa, pb, pc = np.abs(prev - c), np.abs(a - c), np.abs(a + prev - 2 * c)
    paeth = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
    cands = [cur, cur - a, cur - prev, cur - ((a + prev) >> 1), cur - paeth]
    cands = [(x & 0xff).astype(np.uint8) for x in cands]
    scores = np.stack([np.abs(x.view(np.int8).astype(np.int32)).sum(axis=1) for x in cands], axis=1)
    best = scores.argmin(axis=1)  # first minimum
    out = np.empty((h, pitch + 1), dtype=np.uint8)
    out[:, 0] = best
    stack = np.stack(cands, axis=0)
    out[:, 1:] = stack[best, np.arange(h)]
    return out.tobytes()


def zlib_png_stream(storage: np.ndarray, bpp: int, level: int = 6) -> tuple[bytes, bytes]:
    """(filtered, zlib-compressed) for a non-interlaced >= 8-bit image, libpng-like."""
    h = storage.shape[0]
    flat = storage.reshape(h, -1)
    filtered = filter_rows_numpy(flat, bpp)
    return filtered, zlib.compress(filtered, level)
