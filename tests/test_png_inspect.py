"""Host logic of the file-level entry points (no GPU): pngb200_png_inspect_batch walks chunk headers and
parses IHDR / PLTE / tRNS exactly like the oracle's restatement of PNG.Image.decompress(stream:), minus
the CRC check (that runs on the device in pngb200_png_decode_batch)."""
import os

import container_cases as cc
import pngio
from conftest import GOLDEN


def same(im, info):
    return (im.status, im.err_a, im.err_b) == (info.status, info.a, info.b)


def test_inspect_matches_oracle_on_every_fixture(pngb200, orc):
    files = []
    for sub in ("pngsuite", "ios", "invalid"):
        for f in sorted(os.listdir(os.path.join(GOLDEN, sub))):
            if f.endswith(".png"):
                files.append((sub, f, open(os.path.join(GOLDEN, sub, f), "rb").read()))
    got = pngb200.png_inspect([d for _, _, d in files])
    for (sub, name, data), im in zip(files, got):
        info = orc.png_inspect(data)
        if info.status == orc.ERR_LEX_INVALID_CHUNK_CHECKSUM:
            assert im.status == 0, name  # CRC errors are the device's to find
            continue
        assert same(im, info), (name, im.status, info.status)
        if info.status == 0:
            assert (im.width, im.height, im.depth, im.color, im.interlaced, im.standard) == \
                   (info.width, info.height, info.depth, info.color, bool(info.interlaced), info.standard)
            assert im.fields == info.fields() and im.idat_bytes == info.idat_bytes
            assert (im.idat_chunks, im.chunks) == (info.idat_chunks, info.chunks)


def test_inspect_structural_errors(pngb200, orc):
    cases = cc.structural_cases(orc)
    got = pngb200.png_inspect([d for d, _ in cases])
    for (data, want), im in zip(cases, got):
        assert (im.status, im.err_a, im.err_b)[: len(want)] == want, (data[:40], im.status, want)
        assert same(im, orc.png_inspect(data))


def test_status_codes_agree_with_oracle(pngb200, orc):
    import re
    hdr = open(os.path.join(os.path.dirname(GOLDEN), "..", "include", "pngb200.h")).read()
    ours = {m.group(1): int(m.group(2)) for m in re.finditer(r"PNGB200_ERR_(\w+)\s*=\s*(-\d+)", hdr)}
    orh = open(os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "oracle.h")).read()
    theirs = {m.group(1): int(m.group(2)) for m in re.finditer(r"ORC_ERR_(\w+)\s*=\s*(-\d+)", orh)}
    for name, value in theirs.items():
        key = name if name in ours else "PNG_" + name
        assert ours.get(key) == value, name


def test_inspect_differential_random_chunk_sequences(pngb200, orc):
    """random (mostly invalid) chunk sequences with correct CRCs: the product's header walk and the
    oracle's restatement of decompress(stream:) must stop at the same place for the same reason"""
    import random
    import struct
    import zlib
    rnd = random.Random(20260923)
    types = [b"IHDR", b"PLTE", b"tRNS", b"bKGD", b"IDAT", b"IEND", b"gAMA", b"cHRM", b"sRGB", b"iCCP", b"sBIT",
             b"hIST", b"pHYs", b"sPLT", b"tIME", b"tEXt", b"zTXt", b"iTXt", b"CgBI", b"prVt", b"PUBl", b"aBcD", b"ab\x00d"]
    files = []
    for _ in range(600):
        color = rnd.choice([0, 2, 3, 4, 6, 6, 3, 5])
        depth = rnd.choice([1, 2, 4, 8, 8, 16, 3])
        ihdr = struct.pack(">IIBBBBB", rnd.choice([0, 1, 5, 300]), rnd.choice([1, 2, 77]), depth, color,
                           rnd.choice([0, 0, 0, 1]), rnd.choice([0, 0, 0, 3]), rnd.choice([0, 1, 1, 2]))
        seq = []
        if rnd.random() < 0.15:
            seq.append((b"CgBI", bytes(4)))
        if rnd.random() < 0.93:
            seq.append((b"IHDR", ihdr if rnd.random() < 0.95 else ihdr[:-1]))
        for _ in range(rnd.randrange(0, 9)):
            t = rnd.choice(types)
            if t == b"PLTE":
                body = bytes(rnd.randrange(256) for _ in range(rnd.choice([3, 6, 12, 48, 768, 771, 7, 0])))
            elif t == b"tRNS":
                body = bytes(rnd.randrange(4) for _ in range(rnd.choice([1, 2, 2, 6, 6, 3, 0, 17])))
            elif t == b"IDAT":
                body = zlib.compress(bytes(rnd.randrange(64)))[: rnd.choice([100, 100, 3])]
            elif t == b"IHDR":
                body = ihdr
            else:
                body = bytes(rnd.randrange(8))
            seq.append((t, body))
        if rnd.random() < 0.7:
            seq += [(b"IDAT", b"\x78\x9c\x03\x00\x00\x00\x00\x01")] * rnd.choice([1, 1, 2])
            for _ in range(rnd.randrange(0, 3)):
                seq.append((rnd.choice(types), bytes(rnd.randrange(5))))
        if rnd.random() < 0.8:
            seq.append((b"IEND", b""))
        data = pngio.SIGNATURE + b"".join(cc.chunk(t, b) for t, b in seq)
        if rnd.random() < 0.1:
            data = data[: rnd.randrange(len(data) + 1)]
        files.append(data)
    got = pngb200.png_inspect(files)
    kinds = set()
    for data, im in zip(files, got):
        info = orc.png_inspect(data)
        assert same(im, info), (data[8:80], (im.status, im.err_a, im.err_b), (info.status, info.a, info.b))
        kinds.add(info.status)
        if info.status == 0:
            assert im.fields == info.fields() and (im.idat_bytes, im.idat_chunks, im.chunks) == (info.idat_bytes, info.idat_chunks, info.chunks)
    assert len(kinds) >= 15 and 0 in kinds  # the generator reaches most of the error space and some valid files
