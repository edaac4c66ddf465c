"""Parity of the GPU encode path (filter select + LZ77.Deflator) against the CPU oracle and the
reference's committed encoder outputs.  Needs a B200."""
import gzip
import os
import zlib

import numpy as np
import pytest

import corpus
import pngio
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

KEPT = sorted(f[4:] for f in os.listdir(os.path.join(GOLDEN, "encode")) if f.startswith("out-"))


def _inputs():
    rng = np.random.default_rng(42)
    text = b"the quick brown fox jumps over the lazy dog. " * 300
    return [b"", b"a", b"ab", b"abc", b"abcd", bytes(10), bytes(5000), text, text[:3000] + bytes(rng.integers(0, 256, 3000, dtype=np.uint8)),
            rng.integers(0, 4, 20000, dtype=np.uint8).tobytes(), rng.integers(0, 256, 9000, dtype=np.uint8).tobytes(),
            b"ab" * 30000, corpus.make("photo", 96, 64, 5).tobytes(), corpus.make("graphic", 160, 120, 2).tobytes()]


@pytest.mark.parametrize("level", [0, 1, 3, 4, 6, 7, 8, 9, 10, 13])
def test_deflate_matches_oracle_every_mode(pngb200, ctx, orc, level):
    """greedy / lazy / full: compressed bytes identical to the CPU restatement (which is pinned to
    the reference's own outputs), for zlib, gzip and raw (.ios) wrappers"""
    data = _inputs()
    for fmt, ofmt in ((pngb200.FORMAT_ZLIB, orc.ZLIB), (pngb200.FORMAT_GZIP, orc.GZIP), (pngb200.FORMAT_IOS, orc.IOS)):
        got = pngb200.deflate_batch(ctx, data, level, fmt)
        for d, (st, comp) in zip(data, got):
            assert st == 0
            assert comp == orc.deflate(d, level, ofmt), (level, fmt, len(d))


@pytest.mark.parametrize("name", KEPT)
def test_level9_reference_outputs_byte_exact(pngb200, ctx, orc, name):
    """Tests/Outputs/<name>: pixels of Tests/Baselines/<name> -> pngb200_encode_batch(level 9) ==
    the reference encoder's committed IDAT payload"""
    out = pngio.parse(open(os.path.join(GOLDEN, "encode", "out-" + name), "rb").read())
    base = pngio.parse(open(os.path.join(GOLDEN, "encode", "in-" + name), "rb").read())
    storage = pngb200.decode_batch(ctx, [dict(idat=base.idat, width=base.width, height=base.height,
                                              volume=base.volume, depth=base.depth, interlaced=base.interlaced)])[0].pixels
    (st, idat), = pngb200.encode_batch(ctx, [dict(pixels=storage, width=out.width, height=out.height,
                                                  volume=out.volume, depth=out.depth, interlaced=out.interlaced)], level=9)
    assert st == 0 and idat == out.idat


def test_gzip_fixtures_byte_exact(pngb200, ctx):
    g = os.path.join(GOLDEN, "gzip")
    assert pngb200.gzip_archive(ctx, b"", 10) == open(os.path.join(g, "empty.gz"), "rb").read()
    one = open(os.path.join(g, "single-byte.gz"), "rb").read()
    assert pngb200.gzip_archive(ctx, gzip.decompress(one), 10) == one
    txt = open(os.path.join(g, "GzipCompression.txt.gz"), "rb").read()
    assert pngb200.gzip_archive(ctx, gzip.decompress(txt), 13) == txt


def test_encode_decode_roundtrip_on_device(pngb200, ctx, orc):
    """Roundtripping.Encode* (Sources/PNGIntegrationTests/Roundtripping.swift:80-150): decode ->
    compress(level in {4, 7, 10}) -> decode gives the same pixels; all on the GPU path"""
    names = ["basn2c08.png", "basi6a16.png", "basn3p04.png", "f04n2c08.png", "s33i3p04.png", "basn0g16.png"]
    for level in (4, 7, 10):
        pngs = [pngio.parse(open(os.path.join(GOLDEN, "pngsuite", n), "rb").read()) for n in names]
        first = pngb200.decode_batch(ctx, [dict(idat=p.idat, width=p.width, height=p.height, volume=p.volume,
                                                depth=p.depth, interlaced=p.interlaced) for p in pngs])
        enc = pngb200.encode_batch(ctx, [dict(pixels=f.pixels, width=p.width, height=p.height, volume=p.volume,
                                              depth=p.depth, interlaced=p.interlaced) for f, p in zip(first, pngs)], level=level)
        again = pngb200.decode_batch(ctx, [dict(idat=e[1], width=p.width, height=p.height, volume=p.volume,
                                                depth=p.depth, interlaced=p.interlaced) for e, p in zip(enc, pngs)])
        for f, a, e, p in zip(first, again, enc, pngs):
            assert e[0] == 0 and a.status == 0 and a.pixels == f.pixels
            st, storage, _ = orc.png_decode(e[1], p.width, p.height, p.volume, p.depth, p.interlaced)
            assert st == 0 and storage == f.pixels


def test_streaming_deflator_in_png_encoder_call_order(pngb200, ctx, orc):
    """LZ77.Deflator.push(_:last:) / pop() / pull() used exactly as PNG.Encoder.pull uses them
    (Sources/PNG/Encoding/PNG.Encoder.swift:68-128: one push per filtered scanline, pop() after every push,
    push([], last: true) at the end, pull() until nil): the blocks are the reference's -- every complete block
    2 * capacity = 65544 bytes, the concatenation the level-9 stream of the committed output"""
    name = KEPT[0]
    out = pngio.parse(open(os.path.join(GOLDEN, "encode", "out-" + name), "rb").read())
    base = pngio.parse(open(os.path.join(GOLDEN, "encode", "in-" + name), "rb").read())
    storage = pngb200.decode_batch(ctx, [dict(idat=base.idat, width=base.width, height=base.height, volume=base.volume,
                                              depth=base.depth, interlaced=base.interlaced)])[0].pixels
    (filtered,) = pngb200.filter_batch(ctx, [dict(pixels=storage, width=out.width, height=out.height, volume=out.volume,
                                                  depth=out.depth, interlaced=out.interlaced)])
    pitch1 = len(filtered) // out.height
    z = pngb200.Deflator(ctx, pngb200.FORMAT_ZLIB, level=9)
    blocks = []
    for y in range(out.height):
        z.push(filtered[y * pitch1:(y + 1) * pitch1])
        while (b := z.pop()) is not None:
            blocks.append(b)
    z.push(b"", last=True)
    while (b := z.pull()) is not None:
        blocks.append(b)
    assert z.pop() is None and z.pull() is None
    z.close()
    assert b"".join(blocks) == out.idat == orc.deflate(filtered, 9)
    assert all(len(b) == 65544 for b in blocks[:-1]) and 0 < len(blocks[-1]) <= 65544
    # the IDAT chunks of the committed file are these very blocks
    assert [len(b) for b in blocks] == [len(c) for c in pngio.idat_chunks(open(os.path.join(GOLDEN, "encode", "out-" + name), "rb").read())]
    # a small chunk size and the gzip wrapper
    data = corpus.make("photo", 96, 64, 5).tobytes()
    z = pngb200.Deflator(ctx, pngb200.FORMAT_GZIP, level=10, chunk_bytes=1000)
    z.push(data[:5000])
    assert z.pop() is None
    z.push(data[5000:], last=True)
    parts = []
    while (b := z.pull()) is not None:
        parts.append(b)
    assert b"".join(parts) == orc.deflate(data, 10, orc.GZIP) and all(len(p) == 1000 for p in parts[:-1])
