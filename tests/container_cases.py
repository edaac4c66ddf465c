"""Synthetic PNG files that exercise the container layer's lexing / parsing / ordering rules, shared by
the oracle tests (CPU) and the GPU parity tests.  `o` is any module exposing the status constants
(the oracle binding or the product package) and fourcc()."""
import struct
import zlib

import pngio


def chunk(typ, body, crc=None):
    return struct.pack(">I", len(body)) + typ + body + struct.pack(">I", zlib.crc32(typ + body) if crc is None else crc)


def png(chunks):
    return pngio.SIGNATURE + b"".join(chunks)


def fourcc(name: str) -> int:
    return int.from_bytes(name.encode("ascii"), "big")


IHDR = chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, 3, 0, 0, 0))
GRAY = chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, 0, 0, 0, 0))
IDAT = chunk(b"IDAT", zlib.compress(bytes([0, 0, 1, 0, 1, 0])))
PLTE = chunk(b"PLTE", bytes(range(6)))
IEND = chunk(b"IEND", b"")


def structural_cases(o):
    """[(file bytes, (status, a, b)[:k])]: errors that do not depend on any chunk's CRC"""
    _chunk, _png = chunk, png
    ihdr, gray, idat, plte, iend = IHDR, GRAY, IDAT, PLTE, IEND
    o_fourcc = fourcc
    ihdr = _chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, 3, 0, 0, 0))
    gray = _chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, 0, 0, 0, 0))
    idat = _chunk(b"IDAT", zlib.compress(bytes([0, 0, 1, 0, 1, 0])))
    plte = _chunk(b"PLTE", bytes(range(6)))
    iend = _chunk(b"IEND", b"")
    ok = _png([ihdr, plte, _chunk(b"tRNS", b"\x80"), idat, iend])
    info, storage = o.png_decompress(ok)
    assert info.status == 0 and storage == bytes([0, 1, 1, 0])
    assert info.fields()["palette"] == bytes([0, 1, 2, 0x80, 3, 4, 5, 255])
    cases = [
        (b"\x89PNG", (o.ERR_LEX_TRUNCATED_SIGNATURE,)),
        (pngio.SIGNATURE + b"\0\0\0", (o.ERR_LEX_TRUNCATED_CHUNK_HEADER,)),
        (_png([ihdr])[:-3], (o.ERR_LEX_TRUNCATED_CHUNK_BODY, 17)),
        (_png([_chunk(b"IH\x7fR", b"")]), (o.ERR_LEX_INVALID_CHUNK_TYPE, int.from_bytes(b"IH\x7fR", "big"))),
        (_png([_chunk(b"abCd", b""), ihdr]), (o.ERR_DECODE_REQUIRED_CHUNK, o_fourcc("IHDR"), o_fourcc("abCd"))),
        (_png([_chunk(b"aBcD", b"")]), (o.ERR_LEX_INVALID_CHUNK_TYPE,)),  # reserved bit set
        (_png([plte, ihdr]), (o.ERR_DECODE_REQUIRED_CHUNK, o_fourcc("IHDR"), o_fourcc("PLTE"))),
        (_png([ihdr, ihdr]), (o.ERR_DECODE_DUPLICATE_CHUNK, o_fourcc("IHDR"))),
        (_png([ihdr, plte, plte]), (o.ERR_DECODE_DUPLICATE_CHUNK, o_fourcc("PLTE"))),
        (_png([ihdr, idat, iend]), (o.ERR_DECODE_REQUIRED_CHUNK, o_fourcc("PLTE"), o_fourcc("IDAT"))),
        (_png([ihdr, _chunk(b"tRNS", b"\1"), plte]), (o.ERR_DECODE_REQUIRED_CHUNK, o_fourcc("PLTE"), o_fourcc("tRNS"))),
        (_png([ihdr, plte, _chunk(b"tRNS", b"\1\2\3")]), (o.ERR_PARSE_TRANSPARENCY_COUNT, 3, 2)),
        (_png([ihdr, _chunk(b"PLTE", bytes(7))]), (o.ERR_PARSE_PALETTE_CHUNK_LENGTH, 7)),
        (_png([ihdr, _chunk(b"PLTE", bytes(3 * 257))]), (o.ERR_PARSE_PALETTE_COUNT, 257, 256)),
        (_png([gray, plte]), (o.ERR_PARSE_UNEXPECTED_PALETTE,)),
        (_png([gray, _chunk(b"tRNS", b"\1")]), (o.ERR_PARSE_TRANSPARENCY_CHUNK_LENGTH, 1, 2)),
        (_png([gray, _chunk(b"tRNS", b"\1\0")]), (o.ERR_PARSE_TRANSPARENCY_SAMPLE, 256, 255)),
        (_png([ihdr, plte, _chunk(b"gAMA", bytes(4))]), (o.ERR_DECODE_UNEXPECTED_CHUNK, o_fourcc("gAMA"), o_fourcc("PLTE"))),
        (_png([ihdr, plte, idat, plte]), (o.ERR_DECODE_UNEXPECTED_CHUNK, o_fourcc("PLTE"), o_fourcc("IDAT"))),
        (_png([ihdr, plte, idat, _chunk(b"tEXt", b"k\0v"), idat, iend]), (o.ERR_DECODE_UNEXPECTED_CHUNK, o_fourcc("IDAT"), o_fourcc("IDAT"))),
        (_png([ihdr, plte, idat]), (o.ERR_LEX_TRUNCATED_CHUNK_HEADER,)),
        (_png([_chunk(b"IHDR", bytes(12))]), (o.ERR_PARSE_HEADER_CHUNK_LENGTH, 12)),
        (_png([_chunk(b"IHDR", struct.pack(">IIBBBBB", 0, 2, 8, 0, 0, 0, 0))]), (o.ERR_PARSE_HEADER_SIZE, 0, 2)),
        (_png([_chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, 0, 1, 0, 0))]), (o.ERR_PARSE_HEADER_COMPRESSION_CODE, 1)),
        (_png([_chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, 0, 0, 2, 0))]), (o.ERR_PARSE_HEADER_FILTER_CODE, 2)),
        (_png([_chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, 0, 0, 0, 2))]), (o.ERR_PARSE_HEADER_INTERLACING_CODE, 2)),
        (_png([_chunk(b"CgBI", bytes(4)), gray]), (o.ERR_PARSE_HEADER_PIXEL_FORMAT,)),
    ]
    return cases
