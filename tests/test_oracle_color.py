"""Pins the oracle's colour targets (unpack / pack / premultiply / straighten; SURVEY section 8f
row N1) against the reference's own vectors: the 161 PngSuite RGBA<UInt16> goldens, the 32 CgBI
inputs against the premultiplied(as: UInt8) goldens, and the Premultiplication suite's identities."""
import hashlib
import json
import os

import numpy as np
import pytest

import pngio
from conftest import GOLDEN

PNGSUITE = sorted(f for f in os.listdir(os.path.join(GOLDEN, "pngsuite")) if f.endswith(".png"))
IOS = sorted(f for f in os.listdir(os.path.join(GOLDEN, "ios")) if f.endswith(".png"))
DIGESTS = json.load(open(os.path.join(GOLDEN, "pngsuite_rgba.json")))
IOS_DIGESTS = json.load(open(os.path.join(GOLDEN, "ios_rgba.json")))


def load(orc, sub, name):
    png = pngio.parse(open(os.path.join(GOLDEN, sub, name), "rb").read())
    st, storage, _ = orc.png_decode(png.idat, png.width, png.height, png.volume, png.depth,
                                    png.interlaced, fmt=png.fmt)
    assert st == orc.OK
    return png, storage, orc.make_format(**pngio.format_fields(png))


@pytest.mark.parametrize("name", PNGSUITE)
def test_unpack_rgba16_golden(orc, name):
    """Roundtripping.decode (Roundtripping.swift:166-232): unpack(as: RGBA<UInt16>) == golden;
    plus the other targets against numpy restatements of the scaling rules"""
    png, storage, fmt = load(orc, "pngsuite", name)
    st, rgba16 = orc.unpack(storage, fmt, orc.TARGET_RGBA16)
    assert st == 0 and len(rgba16) == DIGESTS[name]["bytes"]
    assert hashlib.sha256(rgba16).hexdigest() == DIGESTS[name]["sha256"]
    wide = np.frombuffer(rgba16, dtype="<u2").reshape(-1, 4)
    # RGBA<UInt8>: 16-bit samples >> 8; narrower samples x (255 / (2^d - 1)); palettes as they are.
    # All of those equal the high byte of the RGBA<UInt16> value (x 257 replicates the byte).
    st, rgba8 = orc.unpack(storage, fmt, orc.TARGET_RGBA8)
    assert st == 0 and rgba8 == (wide >> 8).astype(np.uint8).tobytes()
    st, va16 = orc.unpack(storage, fmt, orc.TARGET_VA16)
    assert st == 0 and va16 == np.ascontiguousarray(wide[:, [0, 3]]).tobytes()
    st, va8 = orc.unpack(storage, fmt, orc.TARGET_VA8)
    assert st == 0 and va8 == (wide[:, [0, 3]] >> 8).astype(np.uint8).tobytes()
    # pack is the inverse of unpack on everything unpack can produce
    for target, px in ((orc.TARGET_RGBA16, rgba16), (orc.TARGET_RGBA8, rgba8)):
        back = orc.pack(px, fmt, target)
        if png.color == 3 or (png.depth == 16 and target == orc.TARGET_RGBA8):
            assert orc.unpack(back, fmt, target)[1] == px  # duplicate palette entries / lost low bytes
        else:
            assert back == storage


@pytest.mark.parametrize("name", IOS)
def test_unpack_ios_golden(orc, name):
    """iOS inputs (CgBI: raw deflate, BGR(A) sample order, premultiplied 8-bit alpha) compare
    against golden.premultiplied(as: UInt8.self) (Roundtripping.swift:206-211)"""
    png, storage, fmt = load(orc, "ios", name)
    assert png.cgbi
    st, rgba16 = orc.unpack(storage, fmt, orc.TARGET_RGBA16)
    assert st == 0 and hashlib.sha256(rgba16).hexdigest() == IOS_DIGESTS[name]["sha256"]


def test_premultiplied_as8_matches_ios_goldens(orc):
    """the common decode + unpack + premultiplied(as: UInt8) reproduces the iOS golden too (same
    picture, so the alpha modes are pinned by the fixtures and not only by identities)"""
    for name in IOS:
        if name not in DIGESTS:
            continue
        png, storage, fmt = load(orc, "pngsuite", name)
        st, px = orc.unpack(storage, fmt, orc.TARGET_RGBA16, orc.ALPHA_PREMULTIPLIED_AS8)
        assert st == 0 and hashlib.sha256(px).hexdigest() == IOS_DIGESTS[name]["sha256"], name


def test_premultiplication_identities(orc):
    """Premultiplication.VA8 / VA16 (Sources/PNGTests/Premultiplication.swift:7-47): premultiplied ==
    round(alpha * color / T.max), and premultiply . straighten . premultiply == premultiply"""
    for color in range(256):
        for alpha in range(256):
            p = orc.premultiply(color, alpha, 8)
            assert p == int(alpha * color / 255 + 0.5)
            assert orc.premultiply(orc.straighten(p, alpha, 8), alpha, 8) == p
    rng = np.random.default_rng(16)
    for color, alpha in rng.integers(0, 65536, size=(4096, 2)):
        color, alpha = int(color), int(alpha)
        p = orc.premultiply(color, alpha, 16)
        assert p == int(np.float64(alpha) * np.float64(color) / 65535.0 + 0.5)
        assert orc.premultiply(orc.straighten(p, alpha, 16), alpha, 16) == p


def test_unpack_edge_cases(orc):
    assert orc.unpack(b"", orc.make_format(6, 8), orc.TARGET_RGBA8) == (0, b"")
    # palette index out of range: the reference traps, the restatement reports it
    fmt = orc.make_format(3, 8, palette=bytes([1, 2, 3, 255, 4, 5, 6, 128]))
    assert orc.unpack(bytes([0, 1, 2]), fmt, orc.TARGET_RGBA8)[0] == orc.ERR_PALETTE_INDEX
    assert orc.unpack(bytes([1, 0]), fmt, orc.TARGET_VA16) == (0, np.array([4 * 257, 128 * 257, 257, 65535], dtype="<u2").tobytes())
    # colours missing from the palette pack to entry 0 (PNG.Color.swift default indexer)
    assert orc.pack(bytes([4, 5, 6, 128, 9, 9, 9, 9]), fmt, orc.TARGET_RGBA8) == bytes([1, 0])
    # chroma keys compare the raw samples; bgr formats store (and key) b, g, r
    fmt = orc.make_format(2, 8, bgr=True, key=(30, 20, 10))
    st, px = orc.unpack(bytes([30, 20, 10, 10, 20, 30]), fmt, orc.TARGET_RGBA8)
    assert st == 0 and px == bytes([10, 20, 30, 0, 30, 20, 10, 255])
