"""Times pngb200_png_decode_batch on a batch of synthetic 1080p PNG files (wall clock, pixels left on the
device) -- run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import ctypes as C
import importlib
import os
import struct
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch

import corpus

pkg = importlib.import_module("swift-png_b200")


def chunk(t, body):
    return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
    w, h = 1920, 1080
    px = corpus.make("photo", w, h, 3)
    z = corpus.zlib_png_stream(px, 4, 6)[1]
    data = (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) +
            b"".join(chunk(b"IDAT", z[o:o + 65544]) for o in range(0, len(z), 65544)) + chunk(b"IEND", b""))
    ctx = pkg.Context(0)
    host = torch.frombuffer(bytearray(data), dtype=torch.uint8).repeat(n).pin_memory()
    out = torch.zeros((n, px.nbytes), dtype=torch.uint8, device="cuda")
    descs = (pkg.PngDesc * n)()
    for i in range(n):
        descs[i].file, descs[i].file_len = host.data_ptr() + i * len(data), len(data)
        descs[i].pixels, descs[i].pixels_cap = out[i].data_ptr(), px.nbytes
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.check(ctx._lib.pngb200_png_decode_batch(ctx.handle, descs, n, pkg.MEM_DEVICE))
        dt = time.perf_counter() - t0
        assert all(descs[i].status == 0 for i in range(n))
        print(f"rep {rep}: {dt * 1e3:.2f} ms, {n * w * h / dt / 1e6:.0f} MPixels/s ({n} files of {len(data)} bytes)")
    assert bytes(out[n - 1].cpu().numpy().tobytes()) == px.tobytes()


if __name__ == "__main__":
    main()
