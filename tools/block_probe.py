"""Measures the speculative DEFLATE block-boundary search on real streams:
how many bit offsets look like a dynamic-block header, at which check the others are rejected, and how
large the blocks between true boundaries are.  CPU only; uses the oracle's block-start trace hook."""
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
    so = os.path.join(HERE, "libblockprobe.so")
    src = os.path.join(HERE, "block_probe.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src], check=True)
    L = C.CDLL(so)
    L.probe_scan.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_size_t]
    L.probe_scan.restype = C.c_size_t
    return L


def block_starts(stream: bytes):
    cap = 1 << 16
    trace = (C.c_uint64 * (3 * cap))()
    L = oracle.lib()
    L.orc_debug_block_starts.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64), C.c_size_t]
    L.orc_debug_block_starts.restype = C.c_size_t
    n = L.orc_debug_block_starts(oracle.ZLIB, stream, len(stream), trace, cap)
    return [(trace[3 * i], trace[3 * i + 1], trace[3 * i + 2]) for i in range(min(n, cap))]


def report(name, stream):
    L = lib()
    starts = block_starts(stream)
    true_bits = {b for b, _, t in starts if t == 2}
    hist = (C.c_uint64 * 8)()
    hits = (C.c_uint64 * 4096)()
    limit = min(len(stream) * 8, 64 << 20)  # every bit offset: the worst case of the search
    found = L.probe_scan(stream, len(stream), 16, limit, hist, hits, 4096)
    got = [hits[i] for i in range(min(found, 4096))]
    false_pos = [p for p in got if p not in true_bits]
    missed = [b for b in true_bits if 16 <= b < limit and b not in set(got)]
    gaps = np.diff(sorted(b for b, _, _ in starts)) if len(starts) > 1 else np.array([0])
    scanned = limit - 16
    print(f"{name}: {len(stream)} bytes, {len(starts)} blocks ({sum(1 for s in starts if s[2] == 2)} dynamic), "
          f"mean block {gaps.mean() / 8 / 1024:.1f} KiB compressed (max {gaps.max() / 8 / 1024:.1f})")
    print(f"   scanned {scanned} bit offsets: plausible {found} (true {len(got) - len(false_pos)}, false {len(false_pos)}, missed {len(missed)})")
    names = ["plausible", "BTYPE", "HLIT/HDIST", "code-length code", "code-length sequence", "literal code", "distance code", "truncated"]
    print("   rejected at: " + ", ".join(f"{names[i]} {hist[i] / scanned:.4%}" for i in range(1, 8)))
    return len(false_pos), len(missed)


def main():
    px = corpus.make("photo", 1920, 1080, 0)
    filtered, z6 = corpus.zlib_png_stream(px, 4, 6)
    report("photo 1080p, zlib level 6", z6)
    report("photo 1080p, zlib level 9", zlib.compress(filtered, 9))
    report("photo 1080p (first 2 MiB), reference level 9 (oracle deflate)", oracle.deflate(filtered[: 2 << 20], 9))
    g = corpus.make("graphic", 1920, 1080, 1)
    report("graphic 1080p, zlib level 6", corpus.zlib_png_stream(g, 4, 6)[1])
    nz = corpus.make("noise", 1024, 1024, 2)
    report("noise 1024x1024, zlib level 6", corpus.zlib_png_stream(nz, 4, 6)[1])


if __name__ == "__main__":
    main()
