"""GPU parity for the colour targets: pngb200_unpack_batch /
pngb200_pack_batch against the reference's RGBA<UInt16> goldens (PngSuite + CgBI inputs, decoded on
the GPU as well) and against the oracle for every target, alpha mode and format."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import pngio
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

PNGSUITE = sorted(f for f in os.listdir(os.path.join(GOLDEN, "pngsuite")) if f.endswith(".png"))
IOS = sorted(f for f in os.listdir(os.path.join(GOLDEN, "ios")) if f.endswith(".png"))
DIGESTS = json.load(open(os.path.join(GOLDEN, "pngsuite_rgba.json")))
IOS_DIGESTS = json.load(open(os.path.join(GOLDEN, "ios_rgba.json")))


def decoded(pngb200, ctx, sub, names):
    pngs = [pngio.parse(open(os.path.join(GOLDEN, sub, n), "rb").read()) for n in names]
    got = pngb200.decode_batch(ctx, [dict(idat=p.idat, width=p.width, height=p.height, volume=p.volume, depth=p.depth,
                                          interlaced=p.interlaced, fmt=p.fmt) for p in pngs])
    assert all(g.status == 0 for g in got)
    return pngs, [dict(storage=g.pixels, **pngio.format_fields(p)) for p, g in zip(pngs, got)]


def test_goldens_through_gpu_decode_and_unpack(pngb200, ctx, orc):
    """Roundtripping.decode, all on the device: IDAT -> inflate -> unfilter -> unpack(RGBA<UInt16>)"""
    pngs, images = decoded(pngb200, ctx, "pngsuite", PNGSUITE)
    for (st, px), name in zip(pngb200.unpack_batch(ctx, images, pngb200.TARGET_RGBA16), PNGSUITE):
        assert st == 0 and hashlib.sha256(px).hexdigest() == DIGESTS[name]["sha256"], name
    _, ios = decoded(pngb200, ctx, "ios", IOS)
    for (st, px), name in zip(pngb200.unpack_batch(ctx, ios, pngb200.TARGET_RGBA16), IOS):
        assert st == 0 and hashlib.sha256(px).hexdigest() == IOS_DIGESTS[name]["sha256"], name
    # premultiplied(as: UInt8.self) of the common decode is the CgBI golden (Roundtripping.swift:206-211)
    common = [images[PNGSUITE.index(n)] for n in IOS if n in DIGESTS]
    for (st, px), name in zip(pngb200.unpack_batch(ctx, common, pngb200.TARGET_RGBA16, pngb200.ALPHA_PREMULTIPLIED_AS8),
                              [n for n in IOS if n in DIGESTS]):
        assert st == 0 and hashlib.sha256(px).hexdigest() == IOS_DIGESTS[name]["sha256"], name


@pytest.mark.parametrize("target", [0, 1, 2, 3])
def test_every_target_and_alpha_mode_matches_oracle(pngb200, ctx, orc, target):
    _, images = decoded(pngb200, ctx, "pngsuite", PNGSUITE)
    modes = [0, 1, 2] + ([3, 4] if target in (1, 3) else [])
    for mode in modes:
        got = pngb200.unpack_batch(ctx, images, target, mode)
        for g, im, name in zip(got, images, PNGSUITE):
            fmt = orc.make_format(im["color"], im["depth"], im["bgr"], im["key"], im["palette"])
            assert g == orc.unpack(im["storage"], fmt, target, mode), (name, mode)
    # pack: the oracle's bytes for every format from the unpacked pixels
    unpacked = pngb200.unpack_batch(ctx, images, target)
    back = pngb200.pack_batch(ctx, [dict(pixels=px, **{k: v for k, v in im.items() if k != "storage"})
                                    for (_, px), im in zip(unpacked, images)], target)
    for b, (_, px), im, name in zip(back, unpacked, images, PNGSUITE):
        fmt = orc.make_format(im["color"], im["depth"], im["bgr"], im["key"], im["palette"])
        assert b == orc.pack(px, fmt, target), name


def test_random_storages_all_formats(pngb200, ctx, orc):
    """every (colour type, depth, bgr, key) the format enum has, random samples, odd pixel counts"""
    rng = np.random.default_rng(77)
    images = []
    for color, depths in ((0, (1, 2, 4, 8, 16)), (2, (8, 16)), (3, (1, 2, 4, 8)), (4, (8, 16)), (6, (8, 16))):
        for depth in depths:
            for variant in range(3):
                n = int(rng.integers(1, 70000))
                ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[color]
                if depth < 8 or color == 3:
                    top = min(1 << depth, 200) if color == 3 else 1 << depth
                    st = rng.integers(0, top, n * ch, dtype=np.uint8).tobytes()
                else:
                    st = rng.integers(0, 256, n * ch * (depth // 8), dtype=np.uint8).tobytes()
                im = dict(storage=st, color=color, depth=depth, bgr=False, key=None, palette=None)
                if color == 3:
                    im["palette"] = rng.integers(0, 256, 4 * 200, dtype=np.uint8).tobytes()
                if color in (0, 2) and variant == 1:
                    im["key"] = tuple(int(x) for x in rng.integers(0, min(1 << depth, 4), 3 if color == 2 else 1))
                    if depth >= 8:  # make the key actually occur
                        arr = np.frombuffer(st, dtype=np.uint8).copy().reshape(n, -1)
                        keyb = b"".join(int(k).to_bytes(depth // 8, "big") for k in im["key"])
                        arr[::3] = np.frombuffer(keyb, dtype=np.uint8)
                        im["storage"] = arr.tobytes()
                if color in (2, 6) and depth == 8 and variant == 2:
                    im["bgr"] = True
                images.append(im)
    for target in range(4):
        for mode in (0, 1):
            got = pngb200.unpack_batch(ctx, images, target, mode)
            for g, im in zip(got, images):
                fmt = orc.make_format(im["color"], im["depth"], im["bgr"], im["key"], im["palette"])
                assert g == orc.unpack(im["storage"], fmt, target, mode), (im["color"], im["depth"], target, mode)
        px = [dict(pixels=rng.integers(0, 256, (len(im["storage"]) // ({0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[im["color"]] * (2 if im["depth"] == 16 else 1)))
                                       * pngb200._TARGET_BYTES[target], dtype=np.uint8).tobytes(),
                   **{k: v for k, v in im.items() if k != "storage"}) for im in images]
        for b, p in zip(pngb200.pack_batch(ctx, px, target), px):
            fmt = orc.make_format(p["color"], p["depth"], p["bgr"], p["key"], p["palette"])
            assert b == orc.pack(p["pixels"], fmt, target), (p["color"], p["depth"], target)


def test_palette_index_out_of_range_and_bad_arguments(pngb200, ctx):
    pal = bytes([1, 2, 3, 255, 4, 5, 6, 128])
    got = pngb200.unpack_batch(ctx, [dict(storage=bytes([0, 1, 2]), color=3, depth=8, palette=pal),
                                     dict(storage=bytes([1, 0]), color=3, depth=8, palette=pal)])
    assert got[0][0] == pngb200.ERR_PNG_PALETTE_INDEX
    assert got[1] == (0, bytes([4, 5, 6, 128, 1, 2, 3, 255]))
    with pytest.raises(pngb200.PNGB200Error):  # premultiplied(as: UInt8) of an 8-bit target
        pngb200.unpack_batch(ctx, [dict(storage=b"\0" * 4, color=6, depth=8)], pngb200.TARGET_RGBA8, pngb200.ALPHA_PREMULTIPLIED_AS8)
    with pytest.raises(pngb200.PNGB200Error):
        pngb200.unpack_batch(ctx, [dict(storage=b"\0" * 4, color=6, depth=4)])
    assert pngb200.unpack_batch(ctx, [dict(storage=b"", color=6, depth=8)]) == [(0, b"")]


def test_device_memspace_misaligned_storage(pngb200, ctx, orc):
    """DEVICE pointers: an rgba16 storage that starts at an odd address (byte path) and an aligned one
    (vector path) give the oracle's pixels; 8K-row-sized so several tiles per CTA run"""
    import torch
    rng = np.random.default_rng(5)
    n = 7680 * 64 + 13
    st = rng.integers(0, 256, n * 8, dtype=np.uint8)
    fmt = orc.make_format(6, 16)
    want16 = orc.unpack(st.tobytes(), fmt, orc.TARGET_RGBA16, orc.ALPHA_PREMULTIPLIED)[1]
    want8 = orc.unpack(st.tobytes(), fmt, orc.TARGET_RGBA8)[1]
    for shift in (0, 1, 4):
        dev = torch.zeros(n * 8 + 64, dtype=torch.uint8, device="cuda")
        dev[shift:shift + n * 8] = torch.from_numpy(st).cuda()
        for target, mode, want in ((pngb200.TARGET_RGBA16, pngb200.ALPHA_PREMULTIPLIED, want16), (pngb200.TARGET_RGBA8, 0, want8)):
            out = torch.zeros(n * pngb200._TARGET_BYTES[target], dtype=torch.uint8, device="cuda")
            d = (pngb200.ColorDesc * 1)()
            d[0].storage, d[0].storage_len = dev.data_ptr() + shift, n * 8
            d[0].pixels, d[0].pixels_len = out.data_ptr(), out.numel()
            d[0].count = n
            d[0].format.color, d[0].format.depth = 6, 16
            torch.cuda.synchronize()
            ctx.check(ctx._lib.pngb200_unpack_batch(ctx.handle, d, 1, target, mode, pngb200.MEM_DEVICE))
            assert d[0].status == 0 and out.cpu().numpy().tobytes() == want, (shift, target)
