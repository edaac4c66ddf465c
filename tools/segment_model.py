"""Runs the CPU model of the segment-parallel inflate (tools/segment_model.c) on the benchmark's streams:
bit-exact against zlib, and how many output bytes of each segment stay symbolic until the segment in
front has been resolved."""
import ctypes as C
import os
import subprocess
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import corpus
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def lib():
    so = os.path.join(HERE, "libsegmentmodel.so")
    srcs = [os.path.join(HERE, "segment_model.c"), os.path.join(HERE, "block_probe.c")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", so] + srcs, check=True)
    L = C.CDLL(so)
    L.segment_model.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64),
                                C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    L.segment_model.restype = C.c_size_t
    return L


def run(stream: bytes, segments: int):
    """(output bytes, [(start bit, symbols, markers, markers past the first 32 KiB)], bit offsets scanned)"""
    L = lib()
    want = zlib.decompress(stream)
    out = C.create_string_buffer(len(want) + 1)
    stats = (C.c_uint64 * (4 * 257))()
    nseg, scanned = C.c_int(0), C.c_uint64(0)
    n = L.segment_model(stream, len(stream), segments, out, len(want), stats, C.byref(nseg), C.byref(scanned))
    rows = [tuple(stats[4 * i + j] for j in range(4)) for i in range(nseg.value)]
    return out.raw[:n], rows, scanned.value


def report(name, stream, segments):
    got, rows, scanned = run(stream, segments)
    ok = got == zlib.decompress(stream)
    sym = sum(r[1] for r in rows[1:]) or 1
    print(f"{name}: {segments} wanted -> {len(rows)} segments joined, bit-exact {ok}, search scanned {scanned} bit offsets "
          f"({scanned / (len(stream) * 8):.2%} of the stream)")
    print(f"   symbolic bytes in segments 1..: {sum(r[2] for r in rows[1:]) / sym:.2%} of their output, "
          f"{sum(r[3] for r in rows[1:]) / sym:.2%} beyond the first 32 KiB of a segment "
          f"(worst segment {max((r[2] / max(r[1], 1) for r in rows[1:]), default=0):.2%})")
    return ok


def main():
    px = corpus.make("photo", 1920, 1080, 0)
    filtered, z6 = corpus.zlib_png_stream(px, 4, 6)
    ok = report("photo 1080p, zlib level 6", z6, 16)
    ok &= report("photo 1080p, zlib level 6", z6, 64)
    ok &= report("photo 1080p (2 MiB), reference level 9", oracle.deflate(filtered[: 2 << 20], 9), 8)
    g = corpus.make("graphic", 1920, 1080, 1)
    ok &= report("graphic 1080p, zlib level 6", corpus.zlib_png_stream(g, 4, 6)[1], 4)
    ok &= report("graphic 1080p, zlib level 1", corpus.zlib_png_stream(g, 4, 1)[1], 16)
    nz = corpus.make("noise", 1024, 1024, 2)
    ok &= report("noise 1024x1024 (stored blocks: nothing to split at)", corpus.zlib_png_stream(nz, 4, 6)[1], 8)
    assert ok


if __name__ == "__main__":
    main()
