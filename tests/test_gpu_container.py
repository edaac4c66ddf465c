"""GPU parity for the file-level entry points: whole PNG files through
pngb200_png_decode_batch / pngb200_png_encode_batch against the reference's fixtures and the oracle --
the 14 malformed inputs with the errors ErrorHandling.swift expects (CRC-32 computed on the device),
every PngSuite / CgBI golden from the file bytes, the committed level-9 outputs byte for byte as whole
files, and files with thousands of IDAT chunks."""
import hashlib
import json
import os
import struct
import zlib

import numpy as np
import pytest

import container_cases as cc
import corpus
import pngio
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

PNGSUITE = sorted(f for f in os.listdir(os.path.join(GOLDEN, "pngsuite")) if f.endswith(".png"))
IOS = sorted(f for f in os.listdir(os.path.join(GOLDEN, "ios")) if f.endswith(".png"))
INVALID = sorted(f for f in os.listdir(os.path.join(GOLDEN, "invalid")) if f.endswith(".png"))
DIGESTS = json.load(open(os.path.join(GOLDEN, "pngsuite_rgba.json")))
IOS_DIGESTS = json.load(open(os.path.join(GOLDEN, "ios_rgba.json")))
ENC = json.load(open(os.path.join(GOLDEN, "encode.json")))
KEPT = sorted(f[4:] for f in os.listdir(os.path.join(GOLDEN, "encode")) if f.startswith("out-"))


def read(sub, name):
    return open(os.path.join(GOLDEN, sub, name), "rb").read()


def test_reference_error_cases(pngb200, ctx, orc):
    """ErrorHandling.swift:7-76, with the chunk CRCs computed by crc_regions_kernel"""
    want = {
        "xhdn0g08.png": (pngb200.ERR_LEX_INVALID_CHUNK_CHECKSUM, 1129534797, 1443964200),
        "xcsn0g01.png": (pngb200.ERR_LEX_INVALID_CHUNK_CHECKSUM, 1129534797, 3492746441),
        "xc1n0g08.png": (pngb200.ERR_PARSE_HEADER_PIXEL_FORMAT_CODE, 8, 1),
        "xc9n2c08.png": (pngb200.ERR_PARSE_HEADER_PIXEL_FORMAT_CODE, 8, 9),
        "xd0n2c08.png": (pngb200.ERR_PARSE_HEADER_PIXEL_FORMAT_CODE, 0, 2),
        "xd3n2c08.png": (pngb200.ERR_PARSE_HEADER_PIXEL_FORMAT_CODE, 3, 2),
        "xd9n2c08.png": (pngb200.ERR_PARSE_HEADER_PIXEL_FORMAT_CODE, 99, 2),
        "xdtn0g01.png": (pngb200.ERR_DECODE_REQUIRED_CHUNK, cc.fourcc("IDAT"), cc.fourcc("IEND")),
    }
    files = [read("invalid", n) for n in INVALID]
    for name, data, im in zip(INVALID, files, pngb200.png_decode_batch(ctx, files)):
        info, _ = orc.png_decompress(data)
        assert (im.status, im.err_a, im.err_b) == (info.status, info.a, info.b), name
        if name in want:
            assert (im.status, im.err_a, im.err_b) == want[name], name
        else:
            assert im.status == pngb200.ERR_LEX_INVALID_SIGNATURE, name


def test_goldens_from_file_bytes(pngb200, ctx):
    """PNG.Image.decompress(path:) + unpack(as: RGBA<UInt16>) == golden, one batch for all 193 files"""
    names = [("pngsuite", n) for n in PNGSUITE] + [("ios", n) for n in IOS]
    images = pngb200.png_decode_batch(ctx, [read(s, n) for s, n in names])
    assert all(im.status == 0 for im in images)
    px = pngb200.unpack_batch(ctx, [dict(storage=im.storage, **im.fields) for im in images], pngb200.TARGET_RGBA16)
    for (sub, name), (st, rgba) in zip(names, px):
        want = (DIGESTS if sub == "pngsuite" else IOS_DIGESTS)[name]["sha256"]
        assert st == 0 and hashlib.sha256(rgba).hexdigest() == want, (sub, name)


def test_structural_and_ordering_errors_match_oracle(pngb200, ctx, orc):
    cases = [d for d, _ in cc.structural_cases(orc)]
    ihdr, plte, iend, idat = cc.IHDR, cc.PLTE, cc.IEND, cc.IDAT
    bad = cc.chunk(b"IDAT", b"\x78\x9c\x07")
    cases += [
        cc.png([ihdr, plte, bad, cc.chunk(b"tEXt", b"k\0v", crc=1), iend]),     # decoder error before a later CRC error
        cc.png([ihdr, plte, cc.chunk(b"IDAT", b"\x78\x9c\x07", crc=5), iend]),  # the IDAT's own CRC first
        cc.png([ihdr, plte, cc.chunk(b"IDAT", zlib.compress(bytes([0, 0, 1, 0, 1, 0]))[:-6]), iend]),  # incomplete
        cc.png([ihdr, cc.chunk(b"PLTE", bytes(range(6)), crc=7), idat, iend]),   # CRC error in the preamble
        cc.png([cc.chunk(b"IHDR", bytes(12), crc=9)]),                           # CRC check precedes parsing
        cc.png([ihdr, plte, idat, cc.chunk(b"IDAT", b"", crc=3), iend]),         # second IDAT bad: first is complete
        cc.png([ihdr, plte, cc.chunk(b"IDAT", zlib.compress(bytes(6))[:5]), cc.chunk(b"IDAT", zlib.compress(bytes(6))[5:]), iend]),
        cc.png([ihdr, plte, idat, cc.chunk(b"tEXt", b"a\0b"), iend]),
        cc.png([ihdr, plte, idat, iend]) + b"trailing bytes are never lexed",
        cc.png([ihdr, plte, cc.chunk(b"IDAT", zlib.compress(bytes([0, 0, 1, 0, 1, 0, 0, 1, 1])))  # extra row
                , iend]),
    ]
    for data, im in zip(cases, pngb200.png_decode_batch(ctx, cases)):
        info, storage = orc.png_decompress(data)
        assert (im.status, im.err_a, im.err_b) == (info.status, info.a, info.b), (data[8:60], im.status, info.status)
        assert im.storage == storage


@pytest.mark.parametrize("name", KEPT)
def test_level9_outputs_whole_file(pngb200, ctx, name):
    """Tests/Outputs/<name> == image.compress(level: 9) of Tests/Baselines/<name>: decode the baseline on
    the GPU, encode on the GPU, compare every byte of the file (signature, IHDR, PLTE, 65544-byte IDAT
    chunks with device-computed CRCs, IEND)"""
    (im,) = pngb200.png_decode_batch(ctx, [read("encode", "in-" + name)])
    assert im.status == 0
    ((st, got),) = pngb200.png_encode_batch(ctx, [dict(storage=im.storage, width=im.width, height=im.height,
                                                       interlaced=im.interlaced, **im.fields)], level=9)
    assert st == 0 and hashlib.sha256(got).hexdigest() == ENC[name]["file_sha256"]
    assert got == read("encode", "out-" + name)


def test_encode_every_format_matches_oracle(pngb200, ctx, orc):
    rng = np.random.default_rng(3)
    images, want = [], []
    for fields, w, h in ((dict(color=6, depth=8, bgr=True), 5, 4), (dict(color=2, depth=8, bgr=True, key=(3, 2, 1)), 4, 4),
                         (dict(color=0, depth=4, key=(9,)), 7, 3), (dict(color=2, depth=16, key=(1, 2, 3)), 3, 3),
                         (dict(color=3, depth=2, palette=bytes([1, 2, 3, 255, 4, 5, 6, 7, 8, 9, 10, 255])), 9, 2),
                         (dict(color=6, depth=16), 64, 48), (dict(color=4, depth=8), 33, 17)):
        ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[fields["color"]]
        n = w * h * ch * (2 if fields["depth"] == 16 else 1)
        top = 3 if fields["color"] == 3 else (1 << min(fields["depth"], 8))
        storage = rng.integers(0, top, n, dtype=np.uint8).tobytes()
        for interlaced in (False, True):
            images.append(dict(storage=storage, width=w, height=h, interlaced=interlaced, **fields))
            want.append(orc.png_compress(storage, w, h, orc.make_format(**fields), interlaced, 6, idat_chunk=16))
    got = pngb200.png_encode_batch(ctx, images, level=6, idat_chunk=16)
    for (st, data), ref, im in zip(got, want, images):
        assert st == 0 and data == ref, (im["color"], im["depth"], im["interlaced"])
    # and back through the decoder
    for im, back in zip(images, pngb200.png_decode_batch(ctx, [d for _, d in got])):
        assert back.status == 0 and back.storage == im["storage"]


def test_many_idat_chunks_and_big_files(pngb200, ctx, orc):
    """IDAT framing as other encoders write it (8 KiB chunks -> hundreds of segments to gather, chunk
    bodies straddling CRC pieces), a single 2 MB IDAT decoded in place, odd chunk sizes"""
    files, want = [], []
    for k, (w, h, piece) in enumerate(((1024, 768, 8192), (1920, 1080, 1 << 30), (640, 480, 1), (800, 600, 65537), (333, 777, 3001))):
        px = corpus.make("photo", w, h, 40 + k)
        _, z = corpus.zlib_png_stream(px, 4, 6)
        files.append(pngio.write(w, h, 8, 6, z, idat_chunk=piece))
        want.append(px.tobytes())
    files[2] = files[2][:20000 * 13 + 33 + 5]  # 20 000 one-byte IDAT chunks, then truncated inside a chunk header
    got = pngb200.png_decode_batch(ctx, files)
    for i, (im, data) in enumerate(zip(got, files)):
        info, storage = orc.png_decompress(data)
        assert (im.status, im.err_a, im.err_b) == (info.status, info.a, info.b), i
        if i != 2:
            assert im.status == 0 and im.storage == want[i] and im.idat_chunks == info.idat_chunks
    # one flipped bit deep inside a big IDAT: the device CRC must say exactly what zlib.crc32 says
    hurt = bytearray(files[1])
    hurt[len(hurt) // 2] ^= 0x10
    (im,) = pngb200.png_decode_batch(ctx, [bytes(hurt)])
    at = files[1].index(b"IDAT")
    n = struct.unpack(">I", files[1][at - 4:at])[0]
    declared = struct.unpack(">I", files[1][at + 4 + n:at + 8 + n])[0]
    assert (im.status, im.err_a, im.err_b) == (pngb200.ERR_LEX_INVALID_CHUNK_CHECKSUM, declared, zlib.crc32(bytes(hurt[at:at + 4 + n])))


def test_device_pixels_and_lanes(pngb200, ctx):
    """pixels in DEVICE memory; and a batch big enough for the lane pipeline (host pixels)"""
    import ctypes as C
    import torch
    px = corpus.make("photo", 1920, 1080, 9)
    data = pngio.write(1920, 1080, 8, 6, corpus.zlib_png_stream(px, 4, 6)[1], idat_chunk=32768)
    n = 72
    descs = (pngb200.PngDesc * n)()
    src = C.create_string_buffer(data, len(data))
    out = torch.zeros((n, px.nbytes), dtype=torch.uint8, device="cuda")
    for i in range(n):
        descs[i].file, descs[i].file_len = C.addressof(src), len(data)
    assert ctx._lib.pngb200_png_inspect_batch(descs, n) == 0
    for i in range(n):
        descs[i].pixels, descs[i].pixels_cap = out[i].data_ptr(), px.nbytes
    torch.cuda.synchronize()
    ctx.check(ctx._lib.pngb200_png_decode_batch(ctx.handle, descs, n, pngb200.MEM_DEVICE))
    assert all(descs[i].status == 0 for i in range(n))
    assert bytes(out[n - 1].cpu().numpy().tobytes()) == px.tobytes() and bool((out == out[0]).all())
    images = pngb200.png_decode_batch(ctx, [data] * n)  # 72 x (3.4 + 8.3) MB > 256 MB: goes over the lanes
    assert all(im.status == 0 and im.storage == px.tobytes() for im in images)
