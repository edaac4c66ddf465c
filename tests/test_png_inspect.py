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
