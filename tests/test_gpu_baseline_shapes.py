"""GPU parity at the shapes BASELINE.json's configs name (not only at test-sized images): 8K RGBA16 decode,
256 MiB and 1 GiB gzip streams, full 1080p level-9 encodes, RGBA16 through the host lane pipeline.  Slow-ish
(a few minutes together); the size-independent checks are checksums + sampled windows against zlib / the oracle."""
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import corpus

pytestmark = pytest.mark.gpu


def test_8k_rgba16_decode_matches_oracle(pngb200, ctx, orc):
    """BASELINE configs[3]'s image: 7680x4320 RGBA16 (265 MB of pixels, one stream cut into ~290 segments)"""
    w, h = 7680, 4320
    img = corpus.make("photo", w, h, 0, True)
    filtered, z = corpus.zlib_png_stream(img, 8, 6)
    (g,) = pngb200.decode_batch(ctx, [dict(idat=z, width=w, height=h, volume=64, depth=16, interlaced=0, fmt=0)])
    stats = ctx.segment_stats()
    assert g.status == 0 and g.produced == len(filtered) and g.checksum == zlib.adler32(filtered)
    assert stats["streams"] == 1 and stats["segments"] > 100 and stats["fallbacks"] == 0, stats
    assert g.pixels == np.ascontiguousarray(img).tobytes()
    st, storage, res = orc.png_decode(z, w, h, 64, 16)
    assert st == 0 and storage == g.pixels and res.checksum == g.checksum and res.blocks == g.blocks


@pytest.mark.parametrize("mib", [256, 1024])
def test_big_gzip_stream(pngb200, ctx, mib):
    """BASELINE configs[4]: one gzip stream of 256 MiB / 1 GiB: CRC-32 of the output, length, and sampled
    windows against the plain text (which zlib produced the stream from)"""
    base = b"".join(corpus.zlib_png_stream(corpus.make("photo", 2048, 1024, 0x5EED + k), 4, 6)[0] for k in range(2))
    n = mib << 20
    plain = (base * (n // len(base) + 1))[:n]
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    z = co.compress(plain) + co.flush()
    ((st, out, d),) = pngb200.inflate_batch(ctx, [z], pngb200.FORMAT_GZIP, caps=[n])
    assert st == 0 and d.produced == n and d.checksum == zlib.crc32(plain)
    assert ctx.segment_stats()["fallbacks"] == 0
    rng = np.random.default_rng(mib)
    for o in [0, n - 65536] + [int(x) for x in rng.integers(0, n - 65536, 64)]:
        assert out[o:o + 65536] == plain[o:o + 65536], o
    assert len(out) == n
    ctx.trim()


def test_1080p_level9_encode_four_images(pngb200, ctx, orc):
    """BASELINE configs[2]'s unit of work: whole 1920x1080 RGBA8 images through filter select + level-9 deflate,
    IDAT payload identical to the CPU restatement of the reference's encoder, and it decodes back"""
    w, h = 1920, 1080
    imgs = [np.ascontiguousarray(corpus.make("photo", w, h, 40 + i)).tobytes() for i in range(3)]
    imgs.append(np.ascontiguousarray(corpus.make("graphic", w, h, 44)).tobytes())
    got = pngb200.encode_batch(ctx, [dict(pixels=p, width=w, height=h, volume=32, depth=8, interlaced=0) for p in imgs], level=9)

    def ref(p):
        return orc.deflate(orc.png_filter(p, w, h, 32, 8), 9)

    with ThreadPoolExecutor(4) as ex:
        want = list(ex.map(ref, imgs))
    for (st, idat), wnt, p in zip(got, want, imgs):
        assert st == 0 and idat == wnt
    back = pngb200.decode_batch(ctx, [dict(idat=g[1], width=w, height=h, volume=32, depth=8, interlaced=0, fmt=0) for g in got])
    assert all(b.status == 0 and b.pixels == p for b, p in zip(back, imgs))


def test_rgba16_host_batch_over_the_lanes(pngb200, ctx, orc):
    """80 RGBA16 images in host memory (>= 256 MB to move): the batch is cut into chunks that four lanes decode
    while the others copy; same pixels as the source, same checksums as zlib"""
    w, h = 1024, 512
    uniq = [corpus.make("photo", w, h, 60 + i, True) for i in range(5)]
    streams = [corpus.zlib_png_stream(u, 8, 6) for u in uniq]
    jobs = [dict(idat=streams[i % 5][1], width=w, height=h, volume=64, depth=16, interlaced=0, fmt=0) for i in range(80)]
    got = pngb200.decode_batch(ctx, jobs)
    for i, g in enumerate(got):
        assert g.status == 0 and g.checksum == zlib.adler32(streams[i % 5][0]), i
        assert g.pixels == np.ascontiguousarray(uniq[i % 5]).tobytes(), i
