"""unfilter_wave_kernel (csrc/unfilter.cuh) under the host SIMT emulator, in the round-2 configuration (one 16-byte chunk
per lane and step) and with burst staging (a 128-byte line per row and refill, output collected per line): every filter
type, ragged widths, misaligned rows, several bands talking through the progress words, against the oracle."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))
import emu  # noqa: E402
from oracle import oracle  # noqa: E402

HERE = os.path.join(os.path.dirname(__file__), "emu")


def build(burst: int, warps: int):
    lib = os.path.join(HERE, f"libemu_unfilter_b{burst}w{warps}.so")
    src = os.path.join(HERE, "emu_unfilter.cpp")
    dep = os.path.join(HERE, "..", "..", "swift-png_b200", "csrc", "unfilter.cuh")
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in (src, dep, os.path.join(HERE, "simt.h"))):
        subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-shared", "-fPIC", "-I" + HERE, "-Wno-attributes",
                        f"-DPNGB200_WAVE_BURST={burst}", f"-DPNGB200_WAVE_WARPS={warps}", "-o", lib, src], check=True)
    L = C.CDLL(lib)
    L.emu_unfilter.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint, C.c_int]
    return L


@pytest.mark.parametrize("burst,warps", [(1, 8), (8, 4), (4, 8)])
def test_wave_kernel_matches_oracle(burst, warps):
    L = build(burst, warps)
    rng = np.random.default_rng(burst * 10 + warps)
    for bpp, depth, w, h in [(4, 8, 300, 100), (4, 8, 37, 70), (8, 16, 129, 67), (3, 8, 211, 40), (1, 8, 1000, 33),
                             (4, 8, 3, 200), (2, 8, 64, 64), (6, 16, 50, 97), (4, 8, 1024, 64)]:
        pitch = w * bpp
        rows = rng.integers(0, 256, size=(h, pitch + 1), dtype=np.uint8)
        rows[:, 0] = rng.integers(0, 5, size=h)
        if h > 40:
            rows[37, 0] = 7     # invalid filter byte: the row passes through unchanged
        filtered = rows.tobytes()
        st, want = oracle.png_unfilter(filtered, w, h, 8 * bpp, depth)
        assert st == 0
        for order, grid in ((0, 2), (3, 3)):
            src = (C.c_uint8 * (len(filtered) + 64)).from_buffer_copy(filtered + bytes(64))
            out = (C.c_uint8 * (h * pitch + 64))()
            L.emu_unfilter(C.addressof(src), len(filtered), C.addressof(out), w, h, bpp, depth, grid, order)
            assert bytes(out)[: h * pitch] == want, (bpp, depth, w, h, order)
            assert bytes(out)[h * pitch:] == bytes(64)      # nothing written past the image
