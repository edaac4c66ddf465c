"""Pins the CPU oracle (oracle/) against the reference's own golden vectors for the decode path
and against zlib as an independent inflate.  No GPU needed."""
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

import pngio
from conftest import GOLDEN

PNGSUITE = sorted(f for f in os.listdir(os.path.join(GOLDEN, "pngsuite")) if f.endswith(".png"))
DIGESTS = json.load(open(os.path.join(GOLDEN, "pngsuite_rgba.json")))


def test_all_goldens_present():
    assert len(PNGSUITE) == 161 and set(PNGSUITE) == set(DIGESTS)


@pytest.mark.parametrize("name", PNGSUITE)
def test_pngsuite_golden(orc, name):
    """Roundtripping.Decode* (Sources/PNGIntegrationTests/Roundtripping.swift:7-77): decode ->
    unpack(as: RGBA<UInt16>) == RGBA/<name>.png.rgba"""
    png = pngio.parse(open(os.path.join(GOLDEN, "pngsuite", name), "rb").read())
    st, storage, res = orc.png_decode(png.idat, png.width, png.height, png.volume, png.depth, png.interlaced)
    assert st == orc.OK
    rgba = pngio.unpack_rgba16(png, storage).astype("<u2").tobytes()
    assert len(rgba) == DIGESTS[name]["bytes"]
    assert hashlib.sha256(rgba).hexdigest() == DIGESTS[name]["sha256"]
    # independent inflate
    st, filtered, res = orc.inflate(png.idat)
    assert st == orc.OK and filtered == zlib.decompress(png.idat)
    assert res.checksum == zlib.adler32(filtered)


def test_crc32_kats(orc):
    """ErrorHandling.swift:30,42 known-answer CRC-32s are plain zlib.crc32"""
    for name, computed in (("xhdn0g08.png", 1443964200), ("xcsn0g01.png", 3492746441)):
        data = open(os.path.join(GOLDEN, "invalid", name), "rb").read()
        at, found = 8, False
        while at < len(data):
            n = int.from_bytes(data[at:at + 4], "big")
            body = data[at + 4:at + 8 + n]
            if orc.crc32(body) == computed:
                found = True
            assert orc.crc32(body) == zlib.crc32(body)
            at += 12 + n
        assert found, name


@pytest.mark.parametrize("name", ["empty.gz", "single-byte.gz", "GzipCompression.txt.gz", "GzipCompression.gz"])
def test_gzip_fixtures(orc, name):
    import gzip
    data = open(os.path.join(GOLDEN, "gzip", name), "rb").read()
    st, out, res = orc.inflate(data, orc.GZIP)
    assert st == orc.OK
    assert out == gzip.decompress(data)
    assert res.checksum == zlib.crc32(out)
    assert res.consumed_bits == 8 * len(data)


def test_inflate_errors(orc):
    good = zlib.compress(b"hello hello hello hello" * 10, 9)
    st, out, res = orc.inflate(good)
    assert st == orc.OK and out == b"hello hello hello hello" * 10
    # truncation -> need more input, prefix of the output is available
    st, out, res = orc.inflate(good[:-5])
    assert st == orc.NEED_MORE_INPUT
    # bad adler
    bad = bytearray(good); bad[-1] ^= 1
    st, out, res = orc.inflate(bytes(bad))
    assert st == orc.ERR_STREAM_CHECKSUM and res.b == zlib.adler32(b"hello hello hello hello" * 10)
    # header errors (LZ77.StreamHeaderError)
    assert orc.inflate(b"\x79\x9c" + good[2:])[0] == orc.ERR_ZLIB_METHOD
    assert orc.inflate(b"\x88\x1c" + good[2:])[0] == orc.ERR_ZLIB_WINDOW
    assert orc.inflate(b"\x78\x9d" + good[2:])[0] == orc.ERR_ZLIB_CHECK_BITS
    assert orc.inflate(b"\x78\xbb" + good[2:])[0] == orc.ERR_ZLIB_DICTIONARY
    # block type 3
    assert orc.inflate(b"\x78\x9c\x07")[0] == orc.ERR_BLOCK_TYPE
    # stored block with bad NLEN
    assert orc.inflate(b"\x78\x9c\x01\x01\x00\x00\x00")[0] == orc.ERR_BLOCK_COUNT_PARITY
    # distance too far back: fixed block, literal 'a', match len 3 dist 2
    raw = zlib.compressobj(9, zlib.DEFLATED, -15)
    # hand-assembled fixed block: BFINAL=1 BTYPE=01, lit 'a' (0x61 -> code 0x91, 8 bits),
    # length 3 (sym 257 -> 7-bit 0000001), distance code 1 (5 bits 00001) => distance 2 > 1 byte
    bits = []
    def put(v, n, msb=False):
        for i in (range(n - 1, -1, -1) if msb else range(n)):
            bits.append((v >> i) & 1)
    put(1, 1); put(1, 2)
    put(0x30 + 0x61, 8, msb=True)
    put(1, 7, msb=True)
    put(1, 5, msb=True)
    put(0, 7, msb=True)
    while len(bits) % 8: bits.append(0)
    body = bytes(sum(bits[i + k] << k for k in range(8)) for i in range(0, len(bits), 8))
    assert orc.inflate(body, orc.IOS)[0] == orc.ERR_STRING_REFERENCE


def test_filter_defilter_roundtrip(orc):
    """Filtering.Delay (Sources/PNGTests/Filtering.swift:9-64): filter . defilter = id for
    delay 1...8 -- here through the whole-image entry points and against the numpy restatement"""
    import corpus
    rng = np.random.default_rng(1)
    for bpp, depth in ((1, 8), (2, 8), (3, 8), (4, 8), (6, 16), (8, 16)):
        w, h = 24, 16
        storage = rng.integers(0, 256, size=(h, w * bpp), dtype=np.uint8)
        storage[5] = storage[4]          # make Up attractive
        storage[7, bpp:] = storage[7, :-bpp]  # and Sub
        filtered = orc.png_filter(storage.tobytes(), w, h, 8 * bpp, depth)
        assert filtered == corpus.filter_rows_numpy(storage, bpp)
        st, back = orc.png_unfilter(filtered, w, h, 8 * bpp, depth)
        assert st == orc.OK and back == storage.tobytes()


def test_invalid_filter_byte_passthrough(orc):
    """PNG.Decoder.swift:193-194: an unknown filter byte leaves the row unchanged"""
    row0 = bytes([0]) + bytes(range(8))
    row1 = bytes([7]) + bytes(range(8, 16))
    st, px = orc.png_unfilter(row0 + row1, 2, 2, 32, 8)
    assert st == orc.OK and px == bytes(range(16))
