"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Needs a B200."""
import gzip
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

import corpus
import pngio
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

PNGSUITE = sorted(f for f in os.listdir(os.path.join(GOLDEN, "pngsuite")) if f.endswith(".png"))
DIGESTS = json.load(open(os.path.join(GOLDEN, "pngsuite_rgba.json")))


def job_of(png):
    return dict(idat=png.idat, width=png.width, height=png.height, volume=png.volume,
                depth=png.depth, interlaced=png.interlaced, fmt=png.fmt)


def test_pngsuite_all_formats(pngb200, ctx, orc):
    """all 161 PngSuite goldens through pngb200_decode_batch in ONE batch: storage == oracle
    storage, and unpack(as: RGBA16) == the reference's .rgba golden"""
    pngs = [pngio.parse(open(os.path.join(GOLDEN, "pngsuite", n), "rb").read()) for n in PNGSUITE]
    got = pngb200.decode_batch(ctx, [job_of(p) for p in pngs])
    for name, png, g in zip(PNGSUITE, pngs, got):
        st, storage, res = orc.png_decode(png.idat, png.width, png.height, png.volume, png.depth, png.interlaced)
        assert g.status == st == 0, name
        assert g.pixels == storage, name
        assert g.checksum == res.checksum and g.produced == res.produced, name
        rgba = pngio.unpack_rgba16(png, g.pixels).astype("<u2").tobytes()
        assert hashlib.sha256(rgba).hexdigest() == DIGESTS[name]["sha256"], name


@pytest.mark.parametrize("mode", [1, 2])
def test_pngsuite_both_inflate_kernels(pngb200, ctx, orc, mode):
    ctx.set_inflate_mode(mode)
    try:
        pngs = [pngio.parse(open(os.path.join(GOLDEN, "pngsuite", n), "rb").read()) for n in PNGSUITE]
        got = pngb200.decode_batch(ctx, [job_of(p) for p in pngs])
        for name, png, g in zip(PNGSUITE, pngs, got):
            st, storage, _ = orc.png_decode(png.idat, png.width, png.height, png.volume, png.depth, png.interlaced)
            assert g.status == st == 0 and g.pixels == storage, name
    finally:
        ctx.set_inflate_mode(0)


SHAPES = [(1, 1), (3, 2), (4, 33), (17, 5), (33, 64), (64, 33), (100, 100), (257, 65), (640, 97)]


@pytest.mark.parametrize("bpp,depth", [(1, 8), (2, 8), (3, 8), (4, 8), (6, 16), (8, 16), (2, 16)])
def test_unfilter_random_filters(pngb200, ctx, orc, bpp, depth):
    """every filter type on every row position / width residue, wavefront kernel vs oracle"""
    rng = np.random.default_rng(bpp * 100 + depth)
    jobs, want = [], []
    for (w, h) in SHAPES:
        pitch = w * bpp
        rows = rng.integers(0, 256, size=(h, pitch + 1), dtype=np.uint8)
        rows[:, 0] = rng.integers(0, 5, size=h)
        if h > 3:
            rows[3, 0] = 9  # invalid filter byte: passthrough quirk
        filtered = rows.tobytes()
        jobs.append(dict(filtered=filtered, width=w, height=h, volume=8 * bpp, depth=depth))
        st, px = orc.png_unfilter(filtered, w, h, 8 * bpp, depth)
        assert st == 0
        want.append(px)
    got = pngb200.unfilter_batch(ctx, jobs)
    for (w, h), (st, px), ref in zip(SHAPES, got, want):
        assert st == 0 and px == ref, (w, h, bpp)


@pytest.mark.parametrize("ftype", [0, 1, 2, 3, 4])
def test_unfilter_single_type_tall(pngb200, ctx, orc, ftype):
    """one filter type for a whole tall image: exercises the band-to-band pipeline (many bands)"""
    rng = np.random.default_rng(ftype)
    w, h, bpp = 300, 700, 4
    rows = rng.integers(0, 256, size=(h, w * bpp + 1), dtype=np.uint8)
    rows[:, 0] = ftype
    st, ref = orc.png_unfilter(rows.tobytes(), w, h, 32, 8)
    (gst, px), = pngb200.unfilter_batch(ctx, [dict(filtered=rows.tobytes(), width=w, height=h, volume=32, depth=8)])
    assert gst == st == 0 and px == ref


def test_unfilter_interlaced_and_subbyte(pngb200, ctx, orc):
    rng = np.random.default_rng(7)
    jobs, want = [], []
    for (w, h, vol, depth, il) in [(9, 9, 1, 1, True), (33, 17, 2, 2, False), (40, 40, 4, 4, True),
                                   (31, 31, 32, 8, True), (5, 3, 64, 16, True), (12, 7, 24, 8, True),
                                   (1, 1, 8, 8, True), (2, 3, 16, 16, True)]:
        n = orc.filtered_size(w, h, vol, il)
        filtered = bytearray(rng.integers(0, 256, size=n, dtype=np.uint8).tobytes())
        # put valid filter bytes at row starts
        st, _ = orc.png_unfilter(bytes(filtered), w, h, vol, depth, il)
        jobs.append(dict(filtered=bytes(filtered), width=w, height=h, volume=vol, depth=depth, interlaced=il))
        want.append(orc.png_unfilter(bytes(filtered), w, h, vol, depth, il))
    got = pngb200.unfilter_batch(ctx, jobs)
    for j, (st, px), (rst, ref) in zip(jobs, got, want):
        assert st == rst and px == ref, j["width"]


def _streams():
    rng = np.random.default_rng(11)
    text = (b"the quick brown fox jumps over the lazy dog. " * 400)
    out = {
        "empty": b"",
        "one": b"x",
        "text9": text,
        "zeros": bytes(100000),
        "noise": rng.integers(0, 256, size=70000, dtype=np.uint8).tobytes(),
        "rle258": b"ab" * 40000,
        "mixed": text[:5000] + rng.integers(0, 256, size=5000, dtype=np.uint8).tobytes() + bytes(9000) + text,
    }
    return out


@pytest.mark.parametrize("level,strategy", [(0, 0), (1, 0), (6, 0), (9, 0), (6, zlib.Z_FIXED), (9, zlib.Z_HUFFMAN_ONLY), (9, zlib.Z_RLE)])
def test_inflate_zlib_streams(pngb200, ctx, orc, level, strategy):
    """stored / fixed / dynamic blocks, literals-only, long RLE matches"""
    items = _streams()
    comp = []
    for k, v in items.items():
        c = zlib.compressobj(level, zlib.DEFLATED, 15, 9, strategy)
        comp.append(c.compress(v) + c.flush())
    got = pngb200.inflate_batch(ctx, comp, pngb200.FORMAT_ZLIB, caps=[len(v) for v in items.values()])
    for (k, v), c, (st, out, d) in zip(items.items(), comp, got):
        ost, oout, ores = orc.inflate(c)
        assert st == ost == 0, (k, st, d.err_a, d.err_b)
        assert out == oout == v, k
        assert d.checksum == ores.checksum == zlib.adler32(v), k
        assert d.blocks == ores.blocks and d.consumed_bits == ores.consumed_bits, k


def test_inflate_formats(pngb200, ctx, orc):
    data = b"gzip and raw deflate " * 1000
    raw = zlib.compressobj(9, zlib.DEFLATED, -15)
    rawc = raw.compress(data) + raw.flush()
    gz = gzip.compress(data, 9)
    (st, out, d), = pngb200.inflate_batch(ctx, [rawc], pngb200.FORMAT_IOS, caps=[len(data)])
    assert st == 0 and out == data and d.checksum == zlib.adler32(data)
    (st, out, d), = pngb200.inflate_batch(ctx, [gz], pngb200.FORMAT_GZIP, caps=[len(data)])
    assert st == 0 and out == data and d.checksum == zlib.crc32(data)
    for name in ["empty.gz", "single-byte.gz", "GzipCompression.txt.gz", "GzipCompression.gz"]:
        blob = open(os.path.join(GOLDEN, "gzip", name), "rb").read()
        assert pngb200.gzip_extract(ctx, blob, cap=1 << 16) == gzip.decompress(blob)


def test_inflate_errors_match_oracle(pngb200, ctx, orc):
    payload = b"hello hello hello hello" * 10
    good = zlib.compress(payload, 9)
    bad_adler = bytearray(good); bad_adler[-1] ^= 1
    cases = [good[:-5], good[:7], good[:2], b"", bytes(bad_adler), b"\x79\x9c" + good[2:], b"\x88\x1c" + good[2:],
             b"\x78\x9d" + good[2:], b"\x78\xbb" + good[2:], b"\x78\x9c\x07", b"\x78\x9c\x01\x01\x00\x00\x00",
             b"\x78\x9c\x05\xc0\x81\x00\x00\x00\x00\x00",  # dynamic block, bad code-length code
             ]
    rng = np.random.default_rng(5)
    for i in range(40):  # random corruption of a dynamic stream
        c = bytearray(zlib.compress(bytes(rng.integers(0, 64, size=3000, dtype=np.uint8)), 9))
        for _ in range(3):
            c[int(rng.integers(2, len(c)))] ^= 1 << int(rng.integers(0, 8))
        cases.append(bytes(c))
    got = pngb200.inflate_batch(ctx, cases, pngb200.FORMAT_ZLIB, caps=[1 << 16] * len(cases))
    for c, (st, out, d) in zip(cases, got):
        ost, oout, ores = orc.inflate(c, cap=1 << 16)
        assert st == ost, (c[:8], st, ost)
        if st == pngb200.ERR_STREAM_CHECKSUM:
            assert (d.err_a, d.err_b) == (ores.a, ores.b)
        if st >= 0:
            assert out == oout
    (st, out, d), = pngb200.inflate_batch(ctx, [good], pngb200.FORMAT_ZLIB, caps=[10])
    assert st == pngb200.ERR_OUTPUT_CAPACITY


def test_decode_synthetic_corpus(pngb200, ctx, orc):
    """S0/S1/S2 RGBA8 + RGBA16 images, zlib level 6 with the reference's filter rule"""
    jobs, want = [], []
    for kind, w, h, sixteen in [("photo", 256, 192, False), ("graphic", 256, 192, False), ("noise", 128, 64, False),
                                ("photo", 160, 120, True), ("photo", 1920, 64, False), ("graphic", 1000, 333, False)]:
        img = corpus.make(kind, w, h, 0, sixteen)
        bpp = 8 if sixteen else 4
        filtered, comp = corpus.zlib_png_stream(img, bpp, 6)
        assert filtered == orc.png_filter(img.tobytes(), w, h, 8 * bpp, 16 if sixteen else 8)
        jobs.append(dict(idat=comp, width=w, height=h, volume=8 * bpp, depth=16 if sixteen else 8))
        want.append(img.tobytes())
    got = pngb200.decode_batch(ctx, jobs)
    for g, ref, j in zip(got, want, jobs):
        assert g.status == 0 and g.pixels == ref, (j["width"], j["height"])
        assert g.checksum == orc.adler32(orc.png_filter(ref, j["width"], j["height"], j["volume"], j["depth"]))


def test_decode_error_mapping(pngb200, ctx, orc):
    """PNG.Decoder / PNG.Context error cases: truncated stream -> incompleteImageData...,
    too many bytes -> extraneousImageData, short-but-complete stream -> ok"""
    w, h = 16, 8
    rng = np.random.default_rng(3)
    rows = rng.integers(0, 256, size=(h, w * 4 + 1), dtype=np.uint8)
    rows[:, 0] = rng.integers(0, 5, size=h)
    f = rows.tobytes()
    cases = [zlib.compress(f)[:-9], zlib.compress(f + b"\x00" * 3), zlib.compress(f[:-65]),
             zlib.compress(f + bytes(1000))]
    got = pngb200.decode_batch(ctx, [dict(idat=c, width=w, height=h, volume=32, depth=8) for c in cases])
    for c, g in zip(cases, got):
        st, storage, _ = orc.png_decode(c, w, h, 32, 8)
        assert g.status == st, (g.status, st)
    assert [g.status for g in got] == [pngb200.ERR_PNG_INCOMPLETE_DATASTREAM, pngb200.ERR_PNG_EXTRANEOUS_IMAGE_DATA,
                                       pngb200.OK, pngb200.ERR_PNG_EXTRANEOUS_IMAGE_DATA]
    # the rows that were available are decoded, the missing one stays as PNG.Image.storage is initialised: zero
    # (not the bytes of whatever image used the arena before -- decode a batch of noise first to dirty it)
    st, storage, _ = orc.png_decode(cases[2], w, h, 32, 8)
    assert got[2].pixels[: (h - 1) * w * 4] == storage[: (h - 1) * w * 4]
    assert got[2].pixels[(h - 1) * w * 4:] == bytes(w * 4)
    assert len(got[2].pixels) == h * w * 4


def test_filter_batch_matches_oracle(pngb200, ctx, orc):
    rng = np.random.default_rng(21)
    jobs, want = [], []
    for (w, h, vol, depth, il) in [(24, 16, 8, 8, False), (24, 16, 24, 8, False), (100, 37, 32, 8, False),
                                   (33, 9, 64, 16, False), (256, 128, 32, 8, False), (9, 9, 1, 1, True),
                                   (17, 5, 4, 4, False), (31, 33, 32, 8, True), (64, 64, 48, 16, False)]:
        if depth >= 8:
            img = corpus.make("photo", w, h, 1)[..., : max(1, vol // 8)] if vol <= 32 else corpus.make("photo", w, h, 1, True)[..., : vol // 8]
            storage = np.ascontiguousarray(img).tobytes()
        else:
            storage = rng.integers(0, 1 << depth, size=w * h, dtype=np.uint8).tobytes()
        jobs.append(dict(pixels=storage, width=w, height=h, volume=vol, depth=depth, interlaced=il))
        want.append(orc.png_filter(storage, w, h, vol, depth, il))
    got = pngb200.filter_batch(ctx, jobs)
    for j, g, ref in zip(jobs, got, want):
        assert g == ref, (j["width"], j["height"], j["volume"])


def test_streaming_inflator(pngb200, ctx, orc):
    """LZ77.Inflator push/pull semantics: slices of arbitrary size, pull(n) exact-or-None"""
    rng = np.random.default_rng(2)
    data = corpus.make("photo", 200, 150, 3).tobytes()
    comp = zlib.compress(data, 6)
    z = pngb200.Inflator(ctx, pngb200.FORMAT_ZLIB)
    out, at = b"", 0
    sizes = [1, 1, 3, 100, 4000, 1, 65536, 10 ** 9]
    status = None
    for s in sizes:
        if at >= len(comp):
            break
        status = z.push(comp[at:at + s])
        at += s
        while True:
            row = z.pull(801)
            if row is None:
                break
            out += row
    assert status == pngb200.OK
    out += z.pull_all()
    assert out == data
    z.close()
    # error surfaces as an exception with the Swift enum's payload
    bad = bytearray(comp); bad[-2] ^= 0x55
    z = pngb200.Inflator(ctx)
    with pytest.raises(pngb200.PNGB200Error) as e:
        z.push(bytes(bad))
    assert e.value.status == pngb200.ERR_STREAM_CHECKSUM and e.value.payload[1] == zlib.adler32(data)
    z.close()


def test_streaming_inflator_idat_sized_pushes(pngb200, ctx, orc):
    """a 6 MB stream pushed in 65544-byte slices (the reference's IDAT chunk size), zlib and gzip: big pushes run through the
    intra-stream parallel kernel from the last block boundary, the checksum pass runs once, when the trailer has been read"""
    import gzip as gz
    data = (corpus.make("photo", 1024, 768, 5).tobytes() * 2)[: 6_000_000]
    for fmt, comp, check in ((pngb200.FORMAT_ZLIB, zlib.compress(data, 6), zlib.adler32(data)),
                             (pngb200.FORMAT_GZIP, gz.compress(data, 6), zlib.crc32(data))):
        z = pngb200.Inflator(ctx, fmt)
        launches0 = ctx.launches
        out, pushes, status = b"", 0, None
        for at in range(0, len(comp), 65544):
            status = z.push(comp[at:at + 65544])
            pushes += 1
            while True:
                row = z.pull(4097)
                if row is None:
                    break
                out += row
        assert status == pngb200.OK
        out += z.pull_all()
        assert out == data
        assert ctx.launches - launches0 <= pushes * 3 // 2 + 6, (ctx.launches - launches0, pushes)   # ~one decode launch per push + one checksum pass (round 1: three per push)
        z.close()
    # a corrupted byte in the middle: the status is the oracle's
    bad = bytearray(zlib.compress(data, 6))
    bad[len(bad) // 2] ^= 0x10
    z = pngb200.Inflator(ctx, pngb200.FORMAT_ZLIB)
    ost = orc.inflate(bytes(bad), orc.ZLIB, len(data))[0]
    got = None
    try:
        for at in range(0, len(bad), 65544):
            got = z.push(bytes(bad[at:at + 65544]))
    except pngb200.PNGB200Error as e:
        got = e.status
    assert got == ost
    z.close()


def test_decode_reference_encoded_streams_large_blocks(pngb200, ctx, orc):
    """streams exactly as the reference's encoder emits them (level 9: dynamic blocks of 2047, 4095,
    ... bytes; level 4: <= 2047 terms per block) produced by our bit-exact GPU encoder, decoded by
    the block-parallel kernel: multi-wave blocks, waves without an end-of-block symbol"""
    jobs, want = [], []
    for kind, w, h, level in [("photo", 640, 480, 9), ("graphic", 800, 600, 9), ("photo", 512, 512, 4),
                              ("noise", 256, 256, 9)]:
        img = corpus.make(kind, w, h, 7)
        (st, idat), = pngb200.encode_batch(ctx, [dict(pixels=img.tobytes(), width=w, height=h, volume=32, depth=8)],
                                           level=level)
        assert st == 0
        filtered = orc.png_filter(img.tobytes(), w, h, 32, 8)
        assert idat == orc.deflate(filtered, level)
        jobs.append(dict(idat=idat, width=w, height=h, volume=32, depth=8))
        want.append(img.tobytes())
    for mode in (1, 2):
        ctx.set_inflate_mode(mode)
        try:
            got = pngb200.decode_batch(ctx, jobs)
        finally:
            ctx.set_inflate_mode(0)
        for g, ref in zip(got, want):
            assert g.status == 0 and g.pixels == ref


def test_host_batch_pipelined_over_lanes(pngb200, ctx, orc):
    """big host-memory batches are cut into chunks worked through by helper lanes (copy/compute
    overlap); results must be the same as the single-lane path, image by image"""
    w, h = 1024, 1024
    imgs = [corpus.make(kind, w, h, i) for i, kind in enumerate(["photo", "graphic", "noise", "photo"])]
    streams = [corpus.zlib_png_stream(im, 4, 6)[1] for im in imgs]
    jobs = [dict(idat=streams[i % 4], width=w, height=h, volume=32, depth=8) for i in range(72)]
    jobs[5] = dict(jobs[5], idat=jobs[5]["idat"][:-7])          # one truncated stream in the middle
    got = pngb200.decode_batch(ctx, jobs)
    for i, g in enumerate(got):
        if i == 5:
            assert g.status == pngb200.ERR_PNG_INCOMPLETE_DATASTREAM
        else:
            assert g.status == 0 and g.pixels == imgs[i % 4].tobytes(), i
    assert ctx.launches > 0


def test_inflate_large_gzip_and_zlib_streams(pngb200, ctx, orc):
    """config 5 shape: standalone multi-MB streams (many blocks, thousands of waves, CRC-32 folded over
    ~1000 chunks); the 48 MB stream is also compared block count for block count with the oracle"""
    import gzip as gz
    rng = np.random.default_rng(9)
    base = corpus.make("photo", 1024, 1024, 11).tobytes()
    big = (base * 12)[: 48 * 1024 * 1024 + 12345]
    streams = [gz.compress(big, 6), zlib.compress(big[: 5_000_001], 9), gz.compress(bytes(rng.integers(0, 256, 3_000_000, dtype=np.uint8)), 1)]
    plain = [big, big[: 5_000_001], None]
    fmts = [pngb200.FORMAT_GZIP, pngb200.FORMAT_ZLIB, pngb200.FORMAT_GZIP]
    got = pngb200.inflate_batch(ctx, streams, fmts, caps=[len(big), 5_000_001, 3_000_000])
    for i, (st, out, d) in enumerate(got):
        assert st == 0, (i, st)
        ref = plain[i] if plain[i] is not None else gz.decompress(streams[i])
        assert out == ref
        assert d.checksum == (zlib.crc32(ref) if fmts[i] == pngb200.FORMAT_GZIP else zlib.adler32(ref))
    ost, oout, ores = orc.inflate(streams[1])
    assert ost == 0 and got[1][2].blocks == ores.blocks and got[1][2].consumed_bits == ores.consumed_bits


# ---- more than one CTA per stream (csrc/inflate_segments.cuh): few big streams are cut at block boundaries ----
def test_segmented_decode_matches_whole_stream_decode(pngb200, ctx, orc):
    """two 2048x1536 RGBA8 and one RGBA16 image (BASELINE config 4's per-GPU shape in small): cut into segments
    by the automatic path, decoded whole with mode 2, compared with the source pixels and with each other"""
    imgs = [corpus.make("photo", 2048, 1536, 21), corpus.make("photo", 2048, 1536, 22), corpus.make("photo", 1536, 1024, 23, True)]
    jobs = []
    for im, (bpp, depth) in zip(imgs, [(4, 8), (4, 8), (8, 16)]):
        filt, z = corpus.zlib_png_stream(im, bpp, 6)
        jobs.append(dict(idat=z, width=im.shape[1], height=im.shape[0], volume=8 * bpp, depth=depth, interlaced=0, fmt=0))
    got = pngb200.decode_batch(ctx, jobs)
    stats = ctx.segment_stats()
    assert stats["streams"] == 3 and stats["segments"] > 6 and stats["fallbacks"] == 0, stats
    ctx.set_inflate_mode(2)
    try:
        whole = pngb200.decode_batch(ctx, jobs)
        assert ctx.segment_stats()["streams"] == 0
    finally:
        ctx.set_inflate_mode(0)
    for g, w, im in zip(got, whole, imgs):
        assert g.status == w.status == 0
        assert g.pixels == w.pixels == np.ascontiguousarray(im).tobytes()
        assert (g.checksum, g.produced, g.blocks) == (w.checksum, w.produced, w.blocks)


def test_segmented_inflate_reference_streams_stored_blocks_and_lookalikes(pngb200, ctx, orc):
    """streams as the reference's encoder writes them (few, growing blocks), stored + fixed blocks, and a complete
    dynamic-block header hidden in stored data: whatever the split-point search finds, the bytes are zlib's"""
    rng = np.random.default_rng(31)
    filt, _ = corpus.zlib_png_stream(corpus.make("photo", 1024, 700, 24), 4, 6)
    inner = zlib.compress(filt[:200_000], 6)[2:-4]
    noise = rng.integers(0, 256, 1_500_000, dtype=np.uint8).tobytes()
    cases = {
        "reference level 9": (orc.deflate(filt, 9), filt),
        "stored + lookalike": (zlib.compress(noise[:700_000] + inner + noise[700_000:] + filt, 6), noise[:700_000] + inner + noise[700_000:] + filt),
        "fixed blocks": ((lambda c: c.compress(filt) + c.flush())(zlib.compressobj(6, zlib.DEFLATED, 15, 9, zlib.Z_FIXED)), filt),
    }
    streams = [v[0] for v in cases.values()]
    got = pngb200.inflate_batch(ctx, streams, pngb200.FORMAT_ZLIB, caps=[len(v[1]) for v in cases.values()])
    for (name, (z, plain)), (st, out, d) in zip(cases.items(), got):
        ost, oout, ores = orc.inflate(z)
        assert st == ost == 0, name
        assert out == plain == oout, name
        assert d.checksum == zlib.adler32(plain) and d.blocks == ores.blocks and d.consumed_bits == ores.consumed_bits, name


def test_segmented_stream_errors_fall_back(pngb200, ctx, orc):
    """a truncated and a corrupted big stream: the segments do not line up, the whole-stream path reports
    exactly what the oracle reports"""
    filt, z = corpus.zlib_png_stream(corpus.make("photo", 1600, 1200, 25), 4, 6)
    bad = bytearray(z)
    bad[len(z) // 2] ^= 0x10
    streams = [z[: 2 * len(z) // 3], bytes(bad)]
    got = pngb200.inflate_batch(ctx, streams, pngb200.FORMAT_ZLIB, caps=[len(filt)] * 2)
    for s, (st, out, d) in zip(streams, got):
        ost, oout, ores = orc.inflate(s, orc.ZLIB, len(filt))
        assert st == ost, (st, ost)
        if st == 0:
            assert out == oout


def test_filter_type_histogram_counter(pngb200, ctx, orc):
    """the device-side form of the reference's -DDUMP_FILTERED_SCANLINES dump: scanlines per filter type of a batch"""
    rng = np.random.default_rng(77)
    w, h = 200, 150
    jobs, want = [], np.zeros(6, dtype=np.int64)
    for k in range(3):
        rows = rng.integers(0, 256, size=(h, w * 4 + 1), dtype=np.uint8)
        rows[:, 0] = rng.integers(0, 5, size=h)
        rows[7 * k, 0] = 9          # an invalid filter byte: the row is left as it is (PNG.Decoder.swift:193-194)
        for v in rows[:, 0]:
            want[min(int(v), 5)] += 1
        jobs.append(dict(idat=zlib.compress(rows.tobytes(), 6), width=w, height=h, volume=32, depth=8, interlaced=0, fmt=0))
    got = pngb200.decode_batch(ctx, jobs)
    assert all(g.status == 0 for g in got)
    assert ctx.filter_histogram() == [int(x) for x in want]
    for j, g in zip(jobs, got):
        st, storage, _ = orc.png_decode(j["idat"], w, h, 32, 8)
        assert st == 0 and storage == g.pixels

