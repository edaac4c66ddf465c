"""The CPU model of the speculative block-boundary search (tools/block_probe.*)
finds exactly the dynamic-block headers the oracle's decode passes through -- no false positives, no
misses -- when every bit offset of a stream is tested."""
import os
import sys
import zlib

import numpy as np

import corpus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_probe_finds_exactly_the_true_boundaries(orc):
    import block_probe
    import ctypes as C
    L = block_probe.lib()
    rng = np.random.default_rng(4)
    streams = [corpus.zlib_png_stream(corpus.make("photo", 640, 360, 5), 4, 6)[1],
               zlib.compress(bytes(rng.integers(0, 256, 200_000, dtype=np.uint8)), 6),      # stored blocks: raw noise
               zlib.compress(bytes(rng.integers(0, 4, 400_000, dtype=np.uint8)), 1),        # many short dynamic blocks
               orc.deflate(corpus.zlib_png_stream(corpus.make("graphic", 320, 200, 6), 4, 6)[0], 9)]
    for z in streams:
        starts = block_probe.block_starts(z)
        assert starts and starts[0][0] == 16  # first block right behind the 2-byte zlib header
        true_bits = sorted(b for b, _, t in starts if t == 2)
        hist, hits = (C.c_uint64 * 8)(), (C.c_uint64 * 4096)()
        found = L.probe_scan(z, len(z), 16, len(z) * 8, hist, hits, 4096)
        assert sorted(hits[i] for i in range(found)) == true_bits
        assert sum(hist) == len(z) * 8 - 16 and hist[1] > 0.7 * sum(hist)  # three quarters die on BTYPE alone


def test_segment_model_is_bit_exact(orc):
    """the whole planned pipeline on the CPU: search, independent segment decodes with a symbolic window,
    chaining by exact arrival, marker resolution -- equals zlib's output whatever the split count"""
    import segment_model
    rng = np.random.default_rng(8)
    streams = [corpus.zlib_png_stream(corpus.make("photo", 800, 600, 5), 4, 6)[1],
               corpus.zlib_png_stream(corpus.make("graphic", 640, 480, 6), 4, 1)[1],
               zlib.compress(bytes(rng.integers(0, 3, 600_000, dtype=np.uint8)), 2),
               zlib.compress(bytes(rng.integers(0, 256, 100_000, dtype=np.uint8)), 6),
               orc.deflate(corpus.zlib_png_stream(corpus.make("photo", 512, 384, 7), 4, 6)[0], 7)]
    for z in streams:
        for want in (1, 2, 7, 40):
            got, rows, _ = segment_model.run(z, want)
            assert got == zlib.decompress(z)
            assert 1 <= len(rows) <= want and rows[0][2] == 0  # segment 0 knows its (empty) window: no markers
