"""PNG container helpers for the tests (chunk walk, IHDR, PLTE/tRNS, RGBA16 unpack, writer).

Test infrastructure: the container layer (Sources/PNG/Lexing, Parsing, ColorTargets in the
reference) is outside the hot path (rows N1/N2 are "next"); the tests need
just enough of it to feed the reference's golden files through the hot path.
"""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass, field

import numpy as np

SIGNATURE = b"\x89PNG\r\n\x1a\n"
CHANNELS = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}


@dataclass
class Png:
    width: int
    height: int
    depth: int
    color: int
    interlaced: bool
    idat: bytes
    palette: bytes | None = None
    trns: bytes | None = None
    cgbi: bool = False
    chunks: list = field(default_factory=list)

    @property
    def volume(self) -> int:  # PNG.Format.Pixel.volume: bits per pixel
        return self.depth * CHANNELS[self.color]

    @property
    def bpp(self) -> int:  # storage bytes per pixel == filter delay
        return (self.volume + 7) >> 3

    @property
    def fmt(self) -> int:  # 0 = zlib, 1 = ios (raw deflate)
        return 1 if self.cgbi else 0


def parse(data: bytes, check_crc: bool = True) -> Png:
    if data[:8] != SIGNATURE:
        raise ValueError("bad signature")
    at, idat, hdr, plte, trns, cgbi, chunks = 8, [], None, None, None, False, []
    while at < len(data):
        (n,) = struct.unpack(">I", data[at:at + 4])
        typ = data[at + 4:at + 8]
        body = data[at + 8:at + 8 + n]
        (crc,) = struct.unpack(">I", data[at + 8 + n:at + 12 + n])
        if check_crc and zlib.crc32(typ + body) != crc:
            raise ValueError(f"bad crc in {typ!r}: declared {crc} computed {zlib.crc32(typ + body)}")
        chunks.append(typ)
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"PLTE":
            plte = body
        elif typ == b"tRNS":
            trns = body
        elif typ == b"CgBI":
            cgbi = True
        elif typ == b"IDAT":
            idat.append(body)
        elif typ == b"IEND":
            break
        at += 12 + n
    w, h, depth, color, _, _, il = hdr
    return Png(w, h, depth, color, bool(il), b"".join(idat), plte, trns, cgbi, chunks)


def chunk(typ: bytes, body: bytes) -> bytes:
    return struct.pack(">I", len(body)) + typ + body + struct.pack(">I", zlib.crc32(typ + body))


def write(width: int, height: int, depth: int, color: int, idat: bytes, interlaced: bool = False,
          palette: bytes | None = None, trns: bytes | None = None, idat_chunk: int = 1 << 30) -> bytes:
    out = [SIGNATURE, chunk(b"IHDR", struct.pack(">IIBBBBB", width, height, depth, color, 0, 0,
                                                 int(interlaced)))]
    if palette is not None:
        out.append(chunk(b"PLTE", palette))
    if trns is not None:
        out.append(chunk(b"tRNS", trns))
    for i in range(0, max(len(idat), 1), idat_chunk):
        out.append(chunk(b"IDAT", idat[i:i + idat_chunk]))
    out.append(chunk(b"IEND", b""))
    return b"".join(out)


def unpack_rgba16(png: Png, storage: bytes) -> np.ndarray:
    """image.unpack(as: PNG.RGBA<UInt16>.self) for the common (non-iOS) standard:
    storage (PNG.Image.storage layout) -> (h*w, 4) uint16."""
    w, h, d, c = png.width, png.height, png.depth, png.color
    n = w * h
    s = np.frombuffer(storage, dtype=np.uint8)
    if d == 16:
        samples = s.reshape(n, -1, 2).astype(np.uint32)
        v = (samples[..., 0] << 8 | samples[..., 1]).astype(np.uint32)  # big-endian samples
        scale = 1
    else:
        v = s.reshape(n, -1).astype(np.uint32)
        scale = 65535 // ((1 << d) - 1)
    out = np.empty((n, 4), dtype=np.uint32)
    if c == 3:
        pal = np.frombuffer(png.palette, dtype=np.uint8).reshape(-1, 3).astype(np.uint32)
        alpha = np.full(len(pal), 255, dtype=np.uint32)
        if png.trns is not None:
            t = np.frombuffer(png.trns, dtype=np.uint8)
            alpha[: len(t)] = t
        idx = v[:, 0]
        out[:, :3] = pal[idx] * 257
        out[:, 3] = alpha[idx] * 257
        return out.astype(np.uint16)
    if c in (0, 4):
        g = v[:, 0] * scale
        out[:, 0] = out[:, 1] = out[:, 2] = g
        if c == 4:
            out[:, 3] = v[:, 1] * scale
        else:
            out[:, 3] = 65535
            if png.trns is not None:
                (key,) = struct.unpack(">H", png.trns[:2])
                out[v[:, 0] == key, 3] = 0
    else:
        out[:, :3] = v[:, :3] * scale
        if c == 6:
            out[:, 3] = v[:, 3] * scale
        else:
            out[:, 3] = 65535
            if png.trns is not None:
                key = np.array(struct.unpack(">HHH", png.trns[:6]), dtype=np.uint32)
                out[(v[:, :3] == key).all(axis=1), 3] = 0
    return out.astype(np.uint16)


def format_fields(png: Png) -> dict:
    """PNG.Format.recognize (Sources/PNG/Formats/PNG.Format.swift:161-330) reduced to what the
    colour-target kernels need: sample order, chroma key in storage order, palette with tRNS merged."""
    out = dict(color=png.color, depth=png.depth, bgr=png.cgbi and png.color in (2, 6), key=None, palette=None)
    if png.color == 3:
        pal = bytearray()
        entries = [png.palette[i:i + 3] for i in range(0, len(png.palette), 3)]
        alpha = list(png.trns or b"")
        for i, e in enumerate(entries):
            pal += bytes(e) + bytes([alpha[i] if i < len(alpha) else 255])
        out["palette"] = bytes(pal)
    elif png.trns is not None and png.color == 0:
        out["key"] = struct.unpack(">H", png.trns[:2])
    elif png.trns is not None and png.color == 2:
        r, g, b = struct.unpack(">HHH", png.trns[:6])
        out["key"] = (b, g, r) if png.cgbi else (r, g, b)
    return out


def idat_chunks(data: bytes) -> list:
    """the IDAT chunk bodies of a PNG file, one by one (the blocks the reference's Deflator handed to its encoder)"""
    at, out = 8, []
    while at < len(data):
        (n,) = struct.unpack(">I", data[at:at + 4])
        typ = data[at + 4:at + 8]
        if typ == b"IDAT":
            out.append(data[at + 8:at + 8 + n])
        if typ == b"IEND":
            break
        at += 12 + n
    return out
