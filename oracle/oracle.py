"""ctypes binding of oracle/liboracle.so -- the CPU restatement of swift-png's hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

ZLIB, IOS, GZIP = 0, 1, 2

OK = 0
NEED_MORE_INPUT = 1
ERR_STREAM_CHECKSUM = -1
ERR_BLOCK_TYPE = -2
ERR_BLOCK_COUNT_PARITY = -3
ERR_RUNLITERAL_SYMBOL_COUNT = -4
ERR_CODELENGTH_HUFFMAN_TABLE = -5
ERR_CODELENGTH_SEQUENCE = -6
ERR_HUFFMAN_TABLE = -7
ERR_STRING_REFERENCE = -8
ERR_INVALID_SYMBOL = -9
ERR_ZLIB_METHOD = -16
ERR_ZLIB_WINDOW = -17
ERR_ZLIB_CHECK_BITS = -18
ERR_ZLIB_DICTIONARY = -19
ERR_GZIP_SIGIL = -32
ERR_GZIP_METHOD = -33
ERR_GZIP_FLAG_BITS = -34
ERR_GZIP_HEADER_CHECKSUM_UNSUPPORTED = -35
ERR_PNG_EXTRANEOUS_IMAGE_DATA = -48
ERR_PNG_EXTRANEOUS_COMPRESSED_DATA = -49
ERR_PNG_INCOMPLETE_DATASTREAM = -50
ERR_OUTPUT_CAPACITY = -64


TARGET_RGBA8, TARGET_RGBA16, TARGET_VA8, TARGET_VA16 = 0, 1, 2, 3
ALPHA_ASIS, ALPHA_PREMULTIPLIED, ALPHA_STRAIGHTENED, ALPHA_PREMULTIPLIED_AS8, ALPHA_STRAIGHTENED_AS8 = range(5)
ERR_PALETTE_INDEX = -51


class Format(C.Structure):
    """orc_format: PNG.Format as the colour-target kernels see it."""
    _fields_ = [
        ("color", C.c_uint8),
        ("depth", C.c_uint8),
        ("bgr", C.c_uint8),
        ("has_key", C.c_uint8),
        ("key", C.c_uint16 * 3),
        ("palette_count", C.c_uint16),
        ("palette", C.c_char_p),
    ]


def make_format(color, depth, bgr=False, key=None, palette=None) -> Format:
    f = Format(color=color, depth=depth, bgr=int(bgr), has_key=int(key is not None))
    if key is not None:
        for i, k in enumerate(key):
            f.key[i] = k
    if palette is not None:
        f._keep = bytes(palette)
        f.palette = f._keep
        f.palette_count = len(f._keep) // 4
    return f


ERR_LEX_TRUNCATED_SIGNATURE, ERR_LEX_INVALID_SIGNATURE, ERR_LEX_TRUNCATED_CHUNK_HEADER = -80, -81, -82
ERR_LEX_TRUNCATED_CHUNK_BODY, ERR_LEX_INVALID_CHUNK_TYPE, ERR_LEX_INVALID_CHUNK_CHECKSUM = -83, -84, -85
ERR_PARSE_HEADER_CHUNK_LENGTH, ERR_PARSE_HEADER_PIXEL_FORMAT_CODE, ERR_PARSE_HEADER_PIXEL_FORMAT = -96, -97, -98
ERR_PARSE_HEADER_COMPRESSION_CODE, ERR_PARSE_HEADER_FILTER_CODE, ERR_PARSE_HEADER_INTERLACING_CODE = -99, -100, -101
ERR_PARSE_HEADER_SIZE, ERR_PARSE_UNEXPECTED_PALETTE, ERR_PARSE_PALETTE_CHUNK_LENGTH = -102, -103, -104
ERR_PARSE_PALETTE_COUNT, ERR_PARSE_UNEXPECTED_TRANSPARENCY, ERR_PARSE_TRANSPARENCY_CHUNK_LENGTH = -105, -106, -107
ERR_PARSE_TRANSPARENCY_SAMPLE, ERR_PARSE_TRANSPARENCY_COUNT = -108, -109
ERR_DECODE_REQUIRED_CHUNK, ERR_DECODE_DUPLICATE_CHUNK, ERR_DECODE_UNEXPECTED_CHUNK = -112, -113, -114


def fourcc(name: str) -> int:
    return int.from_bytes(name.encode("ascii"), "big")


class InflateResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("a", C.c_uint32),
        ("b", C.c_uint32),
        ("consumed_bits", C.c_uint64),
        ("produced", C.c_uint64),
        ("checksum", C.c_uint32),
        ("blocks", C.c_uint32),
    ]


class PngInfo(C.Structure):
    """orc_png_info"""
    _fields_ = [
        ("status", C.c_int32), ("a", C.c_uint32), ("b", C.c_uint32),
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("depth", C.c_uint8), ("color", C.c_uint8), ("interlaced", C.c_uint8), ("standard", C.c_uint8),
        ("format", Format), ("palette_rgba", C.c_uint8 * 1024),
        ("idat_bytes", C.c_uint64), ("idat_chunks", C.c_uint32), ("chunks", C.c_uint32),
        ("inflate", InflateResult),
    ]

    def fields(self) -> dict:
        """the PNG.Format fields as keyword arguments for make_format / the product's colour calls"""
        f = self.format
        return dict(color=f.color, depth=f.depth, bgr=bool(f.bgr),
                    key=tuple(f.key[: 1 if f.color == 0 else 3]) if f.has_key else None,
                    palette=bytes(self.palette_rgba[: 4 * f.palette_count]) if f.color == 3 else None)


def build(force: bool = False) -> str:
    """Compile oracle/*.c into oracle/liboracle.so (gcc; a second or two)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p = C.POINTER(C.c_uint8)
        L.orc_inflate.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                  C.POINTER(InflateResult)]
        L.orc_inflate.restype = None
        L.orc_adler32.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
        L.orc_adler32.restype = C.c_uint32
        L.orc_crc32.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
        L.orc_crc32.restype = C.c_uint32
        L.orc_paeth.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8]
        L.orc_paeth.restype = C.c_uint8
        L.orc_defilter.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int]
        L.orc_defilter.restype = None
        L.orc_png_unfilter.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int,
                                       C.c_int, C.c_int, C.c_void_p]
        L.orc_png_unfilter.restype = C.c_int
        L.orc_png_decode.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_uint32, C.c_uint32,
                                     C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.POINTER(InflateResult)]
        L.orc_png_decode.restype = C.c_int
        L.orc_filter_row.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_int, C.c_void_p]
        L.orc_filter_row.restype = None
        L.orc_png_filter.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_size_t]
        L.orc_png_filter.restype = C.c_size_t
        L.orc_png_filtered_size.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        L.orc_png_filtered_size.restype = C.c_size_t
        if hasattr(L, "orc_deflate"):
            L.orc_deflate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t,
                                      C.c_void_p, C.c_size_t]
            L.orc_deflate.restype = C.c_size_t
            L.orc_deflate_bound.argtypes = [C.c_size_t]
            L.orc_deflate_bound.restype = C.c_size_t
            L.orc_debug_greedy_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_long, C.c_int,
                                                 C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_size_t]
            L.orc_debug_greedy_parse.restype = C.c_size_t
        L.orc_premultiply.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
        L.orc_premultiply.restype = C.c_uint32
        L.orc_straighten.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
        L.orc_straighten.restype = C.c_uint32
        L.orc_unpack.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(Format), C.c_int, C.c_int, C.c_void_p]
        L.orc_unpack.restype = C.c_int
        L.orc_pack.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(Format), C.c_int, C.c_void_p]
        L.orc_pack.restype = C.c_int
        L.orc_png_inspect.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(PngInfo)]
        L.orc_png_inspect.restype = C.c_int
        L.orc_png_decompress.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(PngInfo), C.c_void_p, C.c_size_t]
        L.orc_png_decompress.restype = C.c_int
        L.orc_png_compress.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(Format), C.c_int, C.c_int,
                                       C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_png_compress.restype = C.c_size_t
        L.orc_png_compress_bound.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(Format), C.c_int, C.c_size_t]
        L.orc_png_compress_bound.restype = C.c_size_t
        _lib = L
    return _lib


def inflate(data: bytes, fmt: int = ZLIB, cap: int | None = None):
    """LZ77.Inflator / Gzip.Inflator one-shot.  Returns (status, bytes, InflateResult)."""
    L = lib()
    res = InflateResult()
    if cap is None:
        L.orc_inflate(fmt, data, len(data), None, 0, C.byref(res))  # measure
        cap = int(res.produced)
        res = InflateResult()
    buf = C.create_string_buffer(max(cap, 1))
    L.orc_inflate(fmt, data, len(data), buf, cap, C.byref(res))
    return res.status, buf.raw[: res.produced], res


def adler32(data: bytes, start: int = 1) -> int:
    return lib().orc_adler32(start, data, len(data))


def crc32(data: bytes, start: int = 0) -> int:
    return lib().orc_crc32(start, data, len(data))


def filtered_size(w: int, h: int, volume: int, interlaced: bool = False) -> int:
    return lib().orc_png_filtered_size(w, h, volume, int(interlaced))


def storage_size(w: int, h: int, volume: int) -> int:
    return w * h * ((volume + 7) >> 3)


def png_unfilter(filtered: bytes, w: int, h: int, volume: int, depth: int,
                 interlaced: bool = False):
    buf = C.create_string_buffer(max(storage_size(w, h, volume), 1))
    st = lib().orc_png_unfilter(filtered, len(filtered), w, h, volume, depth, int(interlaced), buf)
    return st, buf.raw[: storage_size(w, h, volume)]


def png_decode(idat: bytes, w: int, h: int, volume: int, depth: int, interlaced: bool = False,
               fmt: int = ZLIB):
    """PNG.Decoder path: inflate + unfilter + assign.  Returns (status, storage, InflateResult)."""
    buf = C.create_string_buffer(max(storage_size(w, h, volume), 1))
    res = InflateResult()
    st = lib().orc_png_decode(fmt, idat, len(idat), w, h, volume, depth, int(interlaced), buf,
                              C.byref(res))
    return st, buf.raw[: storage_size(w, h, volume)], res


def png_filter(storage: bytes, w: int, h: int, volume: int, depth: int,
               interlaced: bool = False) -> bytes:
    """PNG.Encoder filter-select over an image: storage -> filtered stream."""
    n = filtered_size(w, h, volume, interlaced)
    buf = C.create_string_buffer(max(n, 1))
    got = lib().orc_png_filter(storage, w, h, volume, depth, int(interlaced), buf, n)
    assert got == n, (got, n)
    return buf.raw[:n]


def deflate(data: bytes, level: int = 9, fmt: int = ZLIB, exponent: int = 15) -> bytes:
    """LZ77.Deflator / Gzip.Deflator one-shot."""
    L = lib()
    cap = L.orc_deflate_bound(len(data))
    buf = C.create_string_buffer(cap)
    n = L.orc_deflate(fmt, level, exponent, data, len(data), buf, cap)
    assert n != C.c_size_t(-1).value
    return buf.raw[:n]


def greedy_parse(data: bytes, exponent: int, attempts: int = 2 ** 62, goal: int = 2 ** 30):
    """[(run, distance)] of the greedy segmentation (test hook for the reference's Matching KAT)."""
    n = len(data)
    runs, dists = (C.c_int * (n + 8))(), (C.c_int * (n + 8))()
    k = lib().orc_debug_greedy_parse(data, n, exponent, attempts, goal, runs, dists, n + 8)
    return [(runs[i], dists[i]) for i in range(k)]


_TARGET_BYTES = {TARGET_RGBA8: 4, TARGET_RGBA16: 8, TARGET_VA8: 2, TARGET_VA16: 4}
_CHANNELS = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}


def unpack(storage: bytes, fmt: Format, target: int, alpha_mode: int = ALPHA_ASIS):
    """image.unpack(as: PNG.RGBA<T> / PNG.VA<T>) [.premultiplied / .straightened].
    Returns (status, bytes of native little-endian T components)."""
    pixels = len(storage) // (_CHANNELS[fmt.color] * (2 if fmt.depth == 16 else 1))
    buf = C.create_string_buffer(max(pixels * _TARGET_BYTES[target], 1))
    st = lib().orc_unpack(storage, pixels, C.byref(fmt), target, alpha_mode, buf)
    return st, buf.raw[: pixels * _TARGET_BYTES[target]]


def pack(pixels: bytes, fmt: Format, target: int) -> bytes:
    """PNG.Image(packing:size:layout:) storage of an RGBA<T> / VA<T> array."""
    n = len(pixels) // _TARGET_BYTES[target]
    size = n * _CHANNELS[fmt.color] * (2 if fmt.depth == 16 else 1)
    buf = C.create_string_buffer(max(size, 1))
    st = lib().orc_pack(pixels, n, C.byref(fmt), target, buf)
    assert st == 0, st
    return buf.raw[:size]


def premultiply(color: int, alpha: int, bits: int) -> int:
    return lib().orc_premultiply(color, alpha, bits)


def straighten(color: int, alpha: int, bits: int) -> int:
    return lib().orc_straighten(color, alpha, bits)


def png_inspect(data: bytes) -> PngInfo:
    """lex + parse a PNG file (every chunk CRC-checked), no image data decoded"""
    info = PngInfo()
    lib().orc_png_inspect(data, len(data), C.byref(info))
    return info


def png_decompress(data: bytes):
    """PNG.Image.decompress(stream:).  Returns (PngInfo, storage bytes or None on error)."""
    probe = png_inspect(data)
    size = probe.width * probe.height * ((probe.depth * _CHANNELS.get(probe.color, 0) + 7) >> 3)
    buf = C.create_string_buffer(max(size, 1))
    info = PngInfo()
    st = lib().orc_png_decompress(data, len(data), C.byref(info), buf, size)
    return info, (buf.raw[:size] if st == 0 else None)


def png_compress(storage: bytes, w: int, h: int, fmt: Format, interlaced: bool = False, level: int = 9,
                 idat_chunk: int = 65544) -> bytes:
    """PNG.Image.compress(stream:level:) of an image without metadata -> the PNG file"""
    L = lib()
    cap = L.orc_png_compress_bound(w, h, C.byref(fmt), int(interlaced), idat_chunk)
    buf = C.create_string_buffer(cap)
    n = L.orc_png_compress(storage, w, h, C.byref(fmt), int(interlaced), level, idat_chunk, buf, cap)
    assert n != C.c_size_t(-1).value
    return buf.raw[:n]
