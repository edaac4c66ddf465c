"""The cell engine (csrc/inflate_cells.cuh, inflate_mode 6) on the GPU against the oracle.  Runs last in the GPU suite
(file name): the engine is opt-in, the suite in front of it covers the default path."""
import zlib

import numpy as np
import pytest

import corpus

pytestmark = pytest.mark.gpu


# ---- the cell engine (csrc/inflate_cells.cuh): pointer cells + pointer jumping, window gathered from HBM ----
def test_cells_engine_matches_oracle_on_every_corpus(pngb200, ctx, orc):
    """inflate_cells_kernel forced (mode 6) over photographic, flat and noisy images, RGBA8 and RGBA16, zlib levels
    1/6/9 and the reference's own level-9 streams, enough streams to exercise the persistent-CTA ticket: pixels,
    Adler-32, produced and block counts against the oracle"""
    jobs, want = [], []
    for i, (kind, w, h, level, wide) in enumerate([("photo", 1024, 768, 6, False), ("graphic", 1280, 720, 6, False),
                                                   ("noise", 300, 200, 1, False), ("photo", 777, 555, 9, True),
                                                   ("graphic", 640, 480, 9, False), ("photo", 1920, 1080, 6, False)]):
        im = corpus.make(kind, w, h, 40 + i, wide)
        bpp = 8 if wide else 4
        filt, z = corpus.zlib_png_stream(im, bpp, level)
        jobs.append(dict(idat=z, width=w, height=h, volume=8 * bpp, depth=16 if wide else 8, interlaced=0, fmt=0))
        want.append(np.ascontiguousarray(im).tobytes())
    filt = orc.png_filter(want[0], 1024, 768, 32, 8)
    jobs.append(dict(idat=orc.deflate(filt[: 300 * 4097], 9), width=1024, height=300, volume=32, depth=8, interlaced=0, fmt=0))
    want.append(want[0][: 300 * 4096])
    jobs, want = jobs * 3, want * 3
    ctx.set_inflate_mode(6)
    try:
        got = pngb200.decode_batch(ctx, jobs)
        counters = ctx.inflate_counters(len(jobs))
    finally:
        ctx.set_inflate_mode(0)
    for k, (g, ref, job) in enumerate(zip(got, want, jobs)):
        st, storage, res = orc.png_decode(job["idat"], job["width"], job["height"], job["volume"], job["depth"])
        assert g.status == st == 0, k
        assert g.pixels == storage == ref, k
        assert (g.checksum, g.produced, g.blocks) == (res.checksum, res.produced, res.blocks), k
    assert counters["fallbacks"] == 0 and counters["waves"] > 0, counters


def test_cells_engine_errors_and_large_streams(pngb200, ctx, orc):
    """mode 6 on standalone streams: a 24 MB gzip stream (CRC-32 by the checksum kernels), a truncated stream, a bad
    Adler-32, a too-small output buffer -- statuses are the oracle's"""
    import gzip as gz
    base = corpus.make("photo", 1024, 1024, 12).tobytes()
    big = (base * 6)[: 24 * 1024 * 1024 + 777]
    z = zlib.compress(big[:3_000_000], 6)
    bad = bytearray(z)
    bad[-2] ^= 8
    streams = [gz.compress(big, 6), z[: len(z) // 2], bytes(bad), z]
    fmts = [pngb200.FORMAT_GZIP, pngb200.FORMAT_ZLIB, pngb200.FORMAT_ZLIB, pngb200.FORMAT_ZLIB]
    caps = [len(big), 3_000_000, 3_000_000, 1_000_000]
    ctx.set_inflate_mode(6)
    try:
        got = pngb200.inflate_batch(ctx, streams, fmts, caps=caps)
    finally:
        ctx.set_inflate_mode(0)
    assert got[0][0] == 0 and got[0][1] == big and got[0][2].checksum == zlib.crc32(big)
    for i in (1, 2, 3):
        ost, oout, ores = orc.inflate(streams[i], orc.ZLIB, caps[i])
        assert got[i][0] == ost != 0, (i, got[i][0], ost)
