"""inflate_cells_kernel (cells + pointer jumping, csrc/inflate_cells.cuh) under the host SIMT emulator: the same cases
as test_emu_inflate.py runs through inflate_wave_kernel, plus the ones that are specific to this engine -- waves cut
at the cell capacity (compressible data), pointer chains across every chunk, window cells at every alignment."""
from __future__ import annotations

import ctypes as C
import os
import sys
import zlib

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))
import corpus  # noqa: E402
import emu  # noqa: E402
from oracle import oracle  # noqa: E402

ZLIB, RAW, GZIP = 0, 1, 2


@pytest.fixture(scope="module")
def lib():
    L = emu.load("emu_inflate_cells")
    assert L.emu_result_size() == C.sizeof(emu.Result)
    L.emu_shared_size.restype = C.c_size_t
    assert L.emu_shared_size() <= 76800          # three CTAs per SM
    L.emu_inflate_cells.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(emu.Result), C.c_int]
    return L


def run(L, z: bytes, cap: int, fmt: int = ZLIB, order: int = 0, misalign: int = 0):
    src = (C.c_uint8 * (len(z) + 8)).from_buffer_copy(z + b"\0" * 8)
    out = (C.c_uint8 * (cap + 64 + misalign))()
    r = emu.Result()
    st = L.emu_inflate_cells(C.addressof(src), len(z), C.addressof(out) + misalign, cap, fmt, C.byref(r), order)
    return st, bytes(out)[misalign:misalign + r.produced], r


def photo_stream(w, h, level=6, seed=0):
    img = corpus.make("photo", w, h, seed)
    return corpus.zlib_png_stream(img, 4, level)


@pytest.mark.parametrize("order", [0, 1, 5])
def test_photo_all_orders(lib, order):
    filt, z = photo_stream(320, 200)
    st, got, r = run(lib, z, len(filt), order=order)
    assert st == 0 and got == filt
    assert r.ck_done == 1 and r.checksum == zlib.adler32(filt) and r.stat[3] == 0


@pytest.mark.parametrize("kind,w,h", [("graphic", 640, 480), ("noise", 128, 64), ("photo", 97, 33)])
def test_corpora_every_alignment_class(lib, kind, w, h):
    img = corpus.make(kind, w, h, 1)
    filt, z = corpus.zlib_png_stream(img, 4, 6)
    for mis in (0, 5, 15):
        st, got, r = run(lib, z, len(filt), misalign=mis)
        assert st == 0 and got == filt and r.checksum == zlib.adler32(filt) and r.stat[3] == 0


def test_cut_waves_and_shrinking_speculation(lib):
    """flat data: a wave of 8 KiB of compressed bits would expand to megabytes; the engine cuts it at the cell capacity,
    restarts at the token that did not fit and speculates on fewer subsequences (r.deferred counts the cuts)"""
    rng = np.random.default_rng(3)
    plain = bytes(300_000) + b"abcd" * 60_000 + rng.integers(0, 4, 100_000, dtype=np.uint8).tobytes() + bytes(range(256)) * 400
    z = zlib.compress(plain, 9)
    for order in (0, 2):
        st, got, r = run(lib, z, len(plain), order=order)
        assert st == 0 and got == plain and r.checksum == zlib.adler32(plain) and r.stat[3] == 0
        assert r.deferred > 10
    # long pointer chains: distance-1 runs that cross every chunk of the sweep
    plain = b"".join(bytes([k]) * 15_000 for k in range(7))
    z = zlib.compress(plain, 6)
    st, got, r = run(lib, z, len(plain))
    assert st == 0 and got == plain and r.stat[3] == 0


def test_levels_and_strategies(lib):
    filt, _ = photo_stream(256, 96)
    for level in (1, 4, 9):
        z = zlib.compress(filt, level)
        st, got, r = run(lib, z, len(filt))
        assert st == 0 and got == filt
    for strategy in (zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
        co = zlib.compressobj(6, zlib.DEFLATED, 15, 9, strategy)
        z = co.compress(filt) + co.flush()
        st, got, r = run(lib, z, len(filt))
        assert st == 0 and got == filt and r.checksum == zlib.adler32(filt), strategy


def test_mixed_blocks(lib):
    """stored blocks, flat data and ordinary waves in one stream: window cells must find bytes that stored blocks and
    cut waves wrote to HBM"""
    rng = np.random.default_rng(7)
    filt, _ = photo_stream(256, 64)
    parts = [filt, bytes(200_000), rng.integers(0, 256, 70_000, dtype=np.uint8).tobytes(), filt[::-1],
             b"ab" * 150_000, filt]
    co = zlib.compressobj(6)
    z = b""
    for i, p in enumerate(parts):
        z += co.compress(p)
        z += co.flush(zlib.Z_FULL_FLUSH if i % 2 else zlib.Z_SYNC_FLUSH)
    z += co.flush()
    plain = b"".join(parts)
    for order in (0, 3):
        st, got, r = run(lib, z, len(plain), order=order)
        assert st == 0 and got == plain and r.checksum == zlib.adler32(plain) and r.stat[3] == 0


def test_reference_encoder_stream(lib):
    filt, _ = photo_stream(200, 120)
    z = oracle.deflate(filt, 9)
    st, got, r = run(lib, z, len(filt))
    assert st == 0 and got == filt and r.checksum == zlib.adler32(filt)


def test_gzip_and_raw(lib):
    filt, _ = photo_stream(160, 100)
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    z = co.compress(filt) + co.flush()
    st, got, r = run(lib, z, len(filt), fmt=GZIP)
    assert st == 0 and got == filt and r.ck_done == 0 and r.declared == zlib.crc32(filt)
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    z = co.compress(filt) + co.flush()
    st, got, r = run(lib, z, len(filt), fmt=RAW)
    assert st == 0 and got == filt


def test_errors_fall_back_to_the_serial_decoder(lib):
    filt, z = photo_stream(256, 128)
    st, got, r = run(lib, z[: len(z) // 2], len(filt))
    ost, oout, ores = oracle.inflate(z[: len(z) // 2], oracle.ZLIB, len(filt))
    assert st == ost != 0 and r.stat[3] == 1
    st, got, r = run(lib, z, len(filt) // 3)
    ost, _, _ = oracle.inflate(z, oracle.ZLIB, len(filt) // 3)
    assert st == ost != 0
    bad = bytearray(z)
    for k in range(20):
        bad[len(z) // 2 + 37 * k] ^= 0x5A
    st, got, r = run(lib, bytes(bad), len(filt))
    ost, oout, ores = oracle.inflate(bytes(bad), oracle.ZLIB, len(filt))
    assert st == ost and (st != 0 or got == oout)
    bad = bytearray(z)
    bad[-1] ^= 1
    st, got, r = run(lib, bytes(bad), len(filt))
    assert st == oracle.inflate(bytes(bad), oracle.ZLIB, len(filt))[0] != 0 and got == filt


def test_tiny_and_empty(lib):
    for plain in (b"", b"a", b"abc" * 5, bytes(range(256)) * 3):
        z = zlib.compress(plain, 6)
        st, got, r = run(lib, z, len(plain) + 8)
        assert st == 0 and got == plain and r.checksum == zlib.adler32(plain)
