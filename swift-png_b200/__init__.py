"""swift-png_b200: B200-native PNG hot path (DEFLATE inflate + scanline unfilter, filter-select)
behind swift-png's `PNG.Decoder` / `PNG.Encoder` / `LZ77.Inflator` interface.

This module is the thin host-side binding of the C ABI in include/pngb200.h (ctypes; the
library itself has no Python or torch dependency).  It mirrors the reference's names:

    LZ77.Inflator(format:).push/pull      -> Inflator(ctx, format).push / .pull / .pull_all
    Gzip.extract(from:)                   -> gzip_extract(ctx, data)
    PNG.Decoder.push + PNG.Image.assign   -> decode_batch(ctx, [ImageJob...])
    PNG.Encoder.filter (+ collect)        -> filter_batch(ctx, [...])

There is NO CPU fallback: importing works anywhere (so the symbol table can be checked without a
GPU) but creating a Context raises unless an sm_100 GPU and the in-tree libpngb200.so exist.
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
from dataclasses import dataclass

from . import shard  # noqa: F401  (multi-GPU partitioning helpers)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PNGB200_LIB") or os.path.join(_HERE, "libpngb200.so")  # override: tuning experiments only
HEADER_PATH = os.path.join(_HERE, "..", "include", "pngb200.h")

# pngb200_status
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
ERR_LEX_TRUNCATED_SIGNATURE, ERR_LEX_INVALID_SIGNATURE, ERR_LEX_TRUNCATED_CHUNK_HEADER = -80, -81, -82
ERR_LEX_TRUNCATED_CHUNK_BODY, ERR_LEX_INVALID_CHUNK_TYPE, ERR_LEX_INVALID_CHUNK_CHECKSUM = -83, -84, -85
ERR_PARSE_HEADER_CHUNK_LENGTH, ERR_PARSE_HEADER_PIXEL_FORMAT_CODE, ERR_PARSE_HEADER_PIXEL_FORMAT = -96, -97, -98
ERR_PARSE_HEADER_COMPRESSION_CODE, ERR_PARSE_HEADER_FILTER_CODE, ERR_PARSE_HEADER_INTERLACING_CODE = -99, -100, -101
ERR_PARSE_HEADER_SIZE, ERR_PARSE_UNEXPECTED_PALETTE, ERR_PARSE_PALETTE_CHUNK_LENGTH = -102, -103, -104
ERR_PARSE_PALETTE_COUNT, ERR_PARSE_UNEXPECTED_TRANSPARENCY, ERR_PARSE_TRANSPARENCY_CHUNK_LENGTH = -105, -106, -107
ERR_PARSE_TRANSPARENCY_SAMPLE, ERR_PARSE_TRANSPARENCY_COUNT = -108, -109
ERR_DECODE_REQUIRED_CHUNK, ERR_DECODE_DUPLICATE_CHUNK, ERR_DECODE_UNEXPECTED_CHUNK = -112, -113, -114
ERR_OUTPUT_CAPACITY = -64
ERR_BAD_ARGUMENT = -65
ERR_CUDA = -66
ERR_INTERNAL = -67

FORMAT_ZLIB, FORMAT_IOS, FORMAT_GZIP = 0, 1, 2
MEM_HOST, MEM_DEVICE = 0, 1


class StreamDesc(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("src_len", C.c_size_t),
        ("dst", C.c_void_p), ("dst_cap", C.c_size_t),
        ("format", C.c_int32),
        ("status", C.c_int32), ("err_a", C.c_uint32), ("err_b", C.c_uint32),
        ("checksum", C.c_uint32), ("blocks", C.c_uint32),
        ("produced", C.c_uint64), ("consumed_bits", C.c_uint64),
    ]


class ImageDesc(C.Structure):
    _fields_ = [
        ("idat", C.c_void_p), ("idat_len", C.c_size_t),
        ("pixels", C.c_void_p), ("pixels_cap", C.c_size_t),
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("volume", C.c_uint8), ("depth", C.c_uint8), ("interlaced", C.c_uint8), ("format", C.c_uint8),
        ("status", C.c_int32), ("err_a", C.c_uint32), ("err_b", C.c_uint32),
        ("checksum", C.c_uint32), ("blocks", C.c_uint32),
        ("produced", C.c_uint64),
    ]


class FilterDesc(C.Structure):
    _fields_ = [
        ("pixels", C.c_void_p), ("pixels_len", C.c_size_t),
        ("filtered", C.c_void_p), ("filtered_cap", C.c_size_t),
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("volume", C.c_uint8), ("depth", C.c_uint8), ("interlaced", C.c_uint8), ("reserved", C.c_uint8),
        ("status", C.c_int32), ("produced", C.c_uint64),
    ]


class DeflateDesc(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("src_len", C.c_size_t),
        ("dst", C.c_void_p), ("dst_cap", C.c_size_t),
        ("format", C.c_int32), ("level", C.c_int32), ("exponent", C.c_int32),
        ("status", C.c_int32), ("checksum", C.c_uint32), ("blocks", C.c_uint32),
        ("produced", C.c_uint64),
    ]


class EncodeDesc(C.Structure):
    _fields_ = [
        ("pixels", C.c_void_p), ("pixels_len", C.c_size_t),
        ("idat", C.c_void_p), ("idat_cap", C.c_size_t),
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("volume", C.c_uint8), ("depth", C.c_uint8), ("interlaced", C.c_uint8), ("format", C.c_uint8),
        ("level", C.c_int32), ("status", C.c_int32), ("checksum", C.c_uint32), ("blocks", C.c_uint32),
        ("produced", C.c_uint64),
    ]


TARGET_RGBA8, TARGET_RGBA16, TARGET_VA8, TARGET_VA16 = 0, 1, 2, 3
ALPHA_ASIS, ALPHA_PREMULTIPLIED, ALPHA_STRAIGHTENED, ALPHA_PREMULTIPLIED_AS8, ALPHA_STRAIGHTENED_AS8 = range(5)
ERR_PNG_PALETTE_INDEX = -51
_TARGET_BYTES = {TARGET_RGBA8: 4, TARGET_RGBA16: 8, TARGET_VA8: 2, TARGET_VA16: 4}
_CHANNELS = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}


class PixelFormat(C.Structure):
    """pngb200_pixel_format: PNG.Format as the colour-target kernels see it."""
    _fields_ = [
        ("color", C.c_uint8), ("depth", C.c_uint8), ("bgr", C.c_uint8), ("has_key", C.c_uint8),
        ("key", C.c_uint16 * 3), ("palette_count", C.c_uint16), ("palette", C.c_void_p),
    ]


class ColorDesc(C.Structure):
    _fields_ = [
        ("storage", C.c_void_p), ("storage_len", C.c_size_t),
        ("pixels", C.c_void_p), ("pixels_len", C.c_size_t),
        ("count", C.c_uint64), ("format", PixelFormat), ("status", C.c_int32),
    ]


class PngDesc(C.Structure):
    """pngb200_png_desc"""
    _fields_ = [
        ("file", C.c_void_p), ("file_len", C.c_size_t), ("pixels", C.c_void_p), ("pixels_cap", C.c_size_t),
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("depth", C.c_uint8), ("color", C.c_uint8), ("interlaced", C.c_uint8), ("standard", C.c_uint8),
        ("format", PixelFormat), ("palette_rgba", C.c_uint8 * 1024),
        ("storage_size", C.c_uint64), ("idat_bytes", C.c_uint64), ("idat_chunks", C.c_uint32), ("chunks", C.c_uint32),
        ("status", C.c_int32), ("err_a", C.c_uint32), ("err_b", C.c_uint32),
        ("checksum", C.c_uint32), ("blocks", C.c_uint32), ("produced", C.c_uint64),
    ]


class PngEncodeDesc(C.Structure):
    """pngb200_png_encode_desc"""
    _fields_ = [
        ("pixels", C.c_void_p), ("pixels_len", C.c_size_t), ("width", C.c_uint32), ("height", C.c_uint32),
        ("format", PixelFormat), ("interlaced", C.c_uint8), ("level", C.c_int32), ("idat_chunk", C.c_uint32),
        ("file", C.c_void_p), ("file_cap", C.c_size_t),
        ("status", C.c_int32), ("checksum", C.c_uint32), ("blocks", C.c_uint32), ("produced", C.c_uint64),
    ]


class PNGB200Error(RuntimeError):
    def __init__(self, status: int, message: str = ""):
        super().__init__(f"pngb200 status {status}: {message}")
        self.status = status


_lib = None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree (nvcc, sm_100a).  Works without a GPU."""
    spec = importlib.util.spec_from_file_location("_pngb200_build", os.path.join(_HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force, verbose=verbose)


def lib():
    """Load libpngb200.so (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PNGB200Error(ERR_CUDA, f"{LIB_PATH} is missing: run `python __graft_entry__.py build`; "
                           "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.pngb200_ctx_create.argtypes = [C.c_int]
    L.pngb200_ctx_create.restype = C.c_void_p
    L.pngb200_ctx_destroy.argtypes = [C.c_void_p]
    L.pngb200_ctx_destroy.restype = None
    L.pngb200_unpack_batch.argtypes = [C.c_void_p, C.POINTER(ColorDesc), C.c_size_t, C.c_int, C.c_int, C.c_int]
    L.pngb200_unpack_batch.restype = C.c_int
    L.pngb200_pack_batch.argtypes = [C.c_void_p, C.POINTER(ColorDesc), C.c_size_t, C.c_int, C.c_int]
    L.pngb200_pack_batch.restype = C.c_int
    L.pngb200_png_inspect_batch.argtypes = [C.POINTER(PngDesc), C.c_size_t]
    L.pngb200_png_inspect_batch.restype = C.c_int
    L.pngb200_png_decode_batch.argtypes = [C.c_void_p, C.POINTER(PngDesc), C.c_size_t, C.c_int]
    L.pngb200_png_decode_batch.restype = C.c_int
    L.pngb200_png_encode_bound.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(PixelFormat), C.c_int, C.c_uint32]
    L.pngb200_png_encode_bound.restype = C.c_size_t
    L.pngb200_png_encode_batch.argtypes = [C.c_void_p, C.POINTER(PngEncodeDesc), C.c_size_t, C.c_int]
    L.pngb200_png_encode_batch.restype = C.c_int
    L.pngb200_ctx_trim.argtypes = [C.c_void_p]
    L.pngb200_ctx_trim.restype = C.c_int
    L.pngb200_last_error.argtypes = [C.c_void_p]
    L.pngb200_last_error.restype = C.c_char_p
    L.pngb200_ctx_stream.argtypes = [C.c_void_p]
    L.pngb200_ctx_stream.restype = C.c_void_p
    L.pngb200_ctx_device.argtypes = [C.c_void_p]
    L.pngb200_ctx_device.restype = C.c_int
    L.pngb200_ctx_launch_count.argtypes = [C.c_void_p]
    L.pngb200_ctx_launch_count.restype = C.c_uint64
    L.pngb200_ctx_set_inflate_mode.argtypes = [C.c_void_p, C.c_int]
    L.pngb200_ctx_set_inflate_mode.restype = None
    L.pngb200_ctx_stage_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.pngb200_ctx_stage_ms.restype = C.c_int
    L.pngb200_ctx_inflate_stats.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
    L.pngb200_ctx_inflate_stats.restype = C.c_int
    L.pngb200_deflator_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t]
    L.pngb200_deflator_create.restype = C.c_void_p
    L.pngb200_deflator_destroy.argtypes = [C.c_void_p]
    L.pngb200_deflator_destroy.restype = None
    L.pngb200_deflator_push.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int]
    L.pngb200_deflator_push.restype = C.c_int
    for fn in (L.pngb200_deflator_pop, L.pngb200_deflator_pull):
        fn.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t)]
        fn.restype = C.c_int
    L.pngb200_ctx_filter_histogram.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.pngb200_ctx_filter_histogram.restype = C.c_int
    L.pngb200_ctx_segment_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.pngb200_ctx_segment_stats.restype = C.c_int
    if os.environ.get("PNGB200_LIB") is None or hasattr(L, "pngb200_ctx_last_inflate_engine"):   # (older tuning builds lack it)
        L.pngb200_ctx_last_inflate_engine.argtypes = [C.c_void_p]
        L.pngb200_ctx_last_inflate_engine.restype = C.c_int
    L.pngb200_ctx_inflate_counters.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
    L.pngb200_ctx_inflate_counters.restype = C.c_int
    L.pngb200_inflate_batch.argtypes = [C.c_void_p, C.POINTER(StreamDesc), C.c_size_t, C.c_int]
    L.pngb200_inflate_batch.restype = C.c_int
    for name in ("pngb200_decode_batch", "pngb200_decode_batch_enqueue", "pngb200_unfilter_batch"):
        getattr(L, name).argtypes = [C.c_void_p, C.POINTER(ImageDesc), C.c_size_t, C.c_int]
        getattr(L, name).restype = C.c_int
    L.pngb200_decode_batch_finish.argtypes = [C.c_void_p, C.POINTER(ImageDesc), C.c_size_t]
    L.pngb200_decode_batch_finish.restype = C.c_int
    L.pngb200_filter_batch.argtypes = [C.c_void_p, C.POINTER(FilterDesc), C.c_size_t, C.c_int]
    L.pngb200_filter_batch.restype = C.c_int
    L.pngb200_deflate_batch.argtypes = [C.c_void_p, C.POINTER(DeflateDesc), C.c_size_t, C.c_int]
    L.pngb200_deflate_batch.restype = C.c_int
    L.pngb200_deflate_bound.argtypes = [C.c_size_t]
    L.pngb200_deflate_bound.restype = C.c_size_t
    L.pngb200_encode_batch.argtypes = [C.c_void_p, C.POINTER(EncodeDesc), C.c_size_t, C.c_int]
    L.pngb200_encode_batch.restype = C.c_int
    L.pngb200_filtered_size.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int]
    L.pngb200_filtered_size.restype = C.c_size_t
    L.pngb200_storage_size.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
    L.pngb200_storage_size.restype = C.c_size_t
    L.pngb200_inflator_create.argtypes = [C.c_void_p, C.c_int]
    L.pngb200_inflator_create.restype = C.c_void_p
    L.pngb200_inflator_destroy.argtypes = [C.c_void_p]
    L.pngb200_inflator_destroy.restype = None
    L.pngb200_inflator_push.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.pngb200_inflator_push.restype = C.c_int
    L.pngb200_inflator_pull.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.pngb200_inflator_pull.restype = C.c_int
    L.pngb200_inflator_pull_all.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.pngb200_inflator_pull_all.restype = C.c_size_t
    L.pngb200_inflator_available.argtypes = [C.c_void_p]
    L.pngb200_inflator_available.restype = C.c_size_t
    L.pngb200_inflator_error.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_uint32),
                                         C.POINTER(C.c_uint32)]
    L.pngb200_inflator_error.restype = None
    _lib = L
    return L


def filtered_size(w: int, h: int, volume: int, interlaced: bool = False) -> int:
    return lib().pngb200_filtered_size(w, h, volume, int(interlaced))


def storage_size(w: int, h: int, volume: int) -> int:
    return lib().pngb200_storage_size(w, h, volume)


class Context:
    """One GPU: a CUDA stream plus grow-only workspaces (pngb200_ctx)."""

    def __init__(self, device: int = -1):
        self._lib = lib()
        self.handle = self._lib.pngb200_ctx_create(device)
        if not self.handle:
            raise PNGB200Error(ERR_CUDA, self._lib.pngb200_last_error(None).decode())

    def close(self):
        if getattr(self, "handle", None):
            self._lib.pngb200_ctx_destroy(self.handle)
            self.handle = None

    __del__ = close

    def trim(self):
        """Give the grow-only device arenas back to the driver (pngb200_ctx_trim)."""
        self.check(self._lib.pngb200_ctx_trim(self.handle))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def stream(self) -> int:
        return self._lib.pngb200_ctx_stream(self.handle) or 0

    @property
    def device(self) -> int:
        return self._lib.pngb200_ctx_device(self.handle)

    @property
    def launches(self) -> int:
        return self._lib.pngb200_ctx_launch_count(self.handle)

    def stage_ms(self):
        """(inflate, checksum, unfilter) device milliseconds of the last finished decode batch"""
        ms = (C.c_float * 3)()
        self.check(self._lib.pngb200_ctx_stage_ms(self.handle, ms))
        return tuple(ms)

    def inflate_stats(self, count: int):
        """dict of device-side counters summed over the last batch of `count` items"""
        out = (C.c_uint64 * 4)()
        self.check(self._lib.pngb200_ctx_inflate_stats(self.handle, count, out))
        return dict(waves=out[0], sync_rounds=out[1], resolve_rounds=out[2], fallbacks=out[3])

    def inflate_counters(self, count: int):
        """all device-side counters of the last batch (see pngb200_ctx_inflate_counters)"""
        out = (C.c_uint64 * 24)()
        self.check(self._lib.pngb200_ctx_inflate_counters(self.handle, count, out))
        names = ["header_tables", "stage", "speculate", "walk", "chain", "count_scan", "emit", "resolve", "store",
                 "stored_blocks"]
        return dict(waves=out[0], walk_tokens=out[1], resolve_rounds=out[2], fallbacks=out[3], tokens=out[4],
                    matches=out[5], deferred_matches=out[6], blocks=out[7],
                    cycles={n: out[8 + i] for i, n in enumerate(names)})

    def filter_histogram(self):
        """scanlines per filter type (None, Sub, Up, Average, Paeth, invalid) of the last wavefront-unfilter batch"""
        out = (C.c_uint64 * 6)()
        self.check(self._lib.pngb200_ctx_filter_histogram(self.handle, out))
        return list(out)

    def segment_stats(self):
        """(streams cut into segments, segments, streams decoded whole after all) of the last batch"""
        out = (C.c_uint64 * 3)()
        self.check(self._lib.pngb200_ctx_segment_stats(self.handle, out))
        return dict(streams=out[0], segments=out[1], fallbacks=out[2])

    def last_inflate_engine(self) -> str:
        """kernel name of the whole-stream inflate engine the last batch used ('' when none ran)"""
        if not hasattr(self._lib, "pngb200_ctx_last_inflate_engine"):
            return ""
        return {0: "inflate_parallel_kernel", 1: "inflate_wave_kernel", 2: "inflate_cells_kernel"}.get(
            self._lib.pngb200_ctx_last_inflate_engine(self.handle), "")

    def set_inflate_mode(self, mode: int):
        self._lib.pngb200_ctx_set_inflate_mode(self.handle, mode)

    def check(self, rc: int):
        if rc != OK:
            raise PNGB200Error(rc, self._lib.pngb200_last_error(self.handle).decode())


@dataclass
class DecodedImage:
    status: int
    pixels: bytes
    checksum: int
    produced: int
    blocks: int
    err_a: int = 0
    err_b: int = 0


def _buf_addr(b) -> int:
    return C.addressof(b)


def inflate_batch(ctx: Context, streams, fmt: int = FORMAT_ZLIB, caps=None):
    """LZ77.Inflator one-shot over a batch of host byte strings.
    Returns a list of (status, bytes, StreamDesc)."""
    n = len(streams)
    descs = (StreamDesc * n)()
    keep = []
    for i, s in enumerate(streams):
        cap = caps[i] if caps is not None else max(len(s) * 1100 + 1024, 1024)
        src = C.create_string_buffer(bytes(s), len(s)) if len(s) else C.create_string_buffer(1)
        dst = C.create_string_buffer(max(cap, 1))
        keep.append((src, dst))
        descs[i].src = _buf_addr(src)
        descs[i].src_len = len(s)
        descs[i].dst = _buf_addr(dst)
        descs[i].dst_cap = cap
        descs[i].format = fmt if isinstance(fmt, int) else fmt[i]
    ctx.check(ctx._lib.pngb200_inflate_batch(ctx.handle, descs, n, MEM_HOST))
    return [(descs[i].status, keep[i][1].raw[: descs[i].produced], descs[i]) for i in range(n)]


def gzip_extract(ctx: Context, data: bytes, cap: int | None = None) -> bytes:
    """Gzip.extract(from:)"""
    (st, out, d), = inflate_batch(ctx, [data], FORMAT_GZIP, None if cap is None else [cap])
    if st != OK:
        raise PNGB200Error(st, "gzip_extract")
    return out


def decode_batch(ctx: Context, images, memspace: int = MEM_HOST):
    """PNG.Decoder over a batch.  `images`: iterable of dicts/objects with
    idat (bytes), width, height, volume, depth, interlaced, fmt.  Host memory path."""
    images = list(images)
    n = len(images)
    descs = (ImageDesc * n)()
    keep = []
    for i, im in enumerate(images):
        g = im if isinstance(im, dict) else im.__dict__
        idat = bytes(g["idat"])
        w, h, vol = g["width"], g["height"], g["volume"]
        size = storage_size(w, h, vol)
        src = C.create_string_buffer(idat, len(idat)) if len(idat) else C.create_string_buffer(1)
        dst = C.create_string_buffer(max(size, 1))
        keep.append((src, dst, size))
        descs[i].idat = _buf_addr(src)
        descs[i].idat_len = len(idat)
        descs[i].pixels = _buf_addr(dst)
        descs[i].pixels_cap = size
        descs[i].width, descs[i].height = w, h
        descs[i].volume, descs[i].depth = vol, g["depth"]
        descs[i].interlaced = int(bool(g.get("interlaced", False)))
        descs[i].format = g.get("fmt", FORMAT_ZLIB)
    ctx.check(ctx._lib.pngb200_decode_batch(ctx.handle, descs, n, memspace))
    return [DecodedImage(descs[i].status, keep[i][1].raw[: keep[i][2]], descs[i].checksum,
                         descs[i].produced, descs[i].blocks, descs[i].err_a, descs[i].err_b)
            for i in range(n)]


def unfilter_batch(ctx: Context, images):
    """PNG.Decoder.defilter + PNG.Image.assign over already inflated streams
    (`filtered` bytes per image).  Returns list of (status, pixels)."""
    images = list(images)
    n = len(images)
    descs = (ImageDesc * n)()
    keep = []
    for i, g in enumerate(images):
        f = bytes(g["filtered"])
        w, h, vol = g["width"], g["height"], g["volume"]
        size = storage_size(w, h, vol)
        src = C.create_string_buffer(f, len(f)) if len(f) else C.create_string_buffer(1)
        dst = C.create_string_buffer(max(size, 1))
        keep.append((src, dst, size))
        descs[i].idat = _buf_addr(src)
        descs[i].idat_len = len(f)
        descs[i].pixels = _buf_addr(dst)
        descs[i].pixels_cap = size
        descs[i].width, descs[i].height = w, h
        descs[i].volume, descs[i].depth = vol, g["depth"]
        descs[i].interlaced = int(bool(g.get("interlaced", False)))
    ctx.check(ctx._lib.pngb200_unfilter_batch(ctx.handle, descs, n, MEM_HOST))
    return [(descs[i].status, keep[i][1].raw[: keep[i][2]]) for i in range(n)]


def filter_batch(ctx: Context, images):
    """PNG.Image.collect + PNG.Encoder.filter over a batch of storages.  Returns list of bytes."""
    images = list(images)
    n = len(images)
    descs = (FilterDesc * n)()
    keep = []
    for i, g in enumerate(images):
        px = bytes(g["pixels"])
        w, h, vol = g["width"], g["height"], g["volume"]
        il = bool(g.get("interlaced", False))
        fsz = filtered_size(w, h, vol, il)
        src = C.create_string_buffer(px, len(px))
        dst = C.create_string_buffer(max(fsz, 1))
        keep.append((src, dst, fsz))
        descs[i].pixels = _buf_addr(src)
        descs[i].pixels_len = len(px)
        descs[i].filtered = _buf_addr(dst)
        descs[i].filtered_cap = fsz
        descs[i].width, descs[i].height = w, h
        descs[i].volume, descs[i].depth, descs[i].interlaced = vol, g["depth"], int(il)
    ctx.check(ctx._lib.pngb200_filter_batch(ctx.handle, descs, n, MEM_HOST))
    return [keep[i][1].raw[: keep[i][2]] for i in range(n)]


def _fill_format(f: PixelFormat, keep: list, color, depth, bgr=False, key=None, palette=None):
    f.color, f.depth, f.bgr, f.has_key = color, depth, int(bool(bgr)), int(key is not None)
    for k, v in enumerate(key or ()):
        f.key[k] = v
    if palette is not None:
        pal = C.create_string_buffer(bytes(palette), len(palette))
        keep.append(pal)
        f.palette = _buf_addr(pal)
        f.palette_count = len(palette) // 4


def unpack_batch(ctx: Context, images, target: int = TARGET_RGBA8, alpha_mode: int = ALPHA_ASIS):
    """image.unpack(as: PNG.RGBA<T>.self / PNG.VA<T>.self) [.premultiplied / .straightened] over a
    batch.  images: dicts with storage (bytes) and the PNG.Format fields color, depth[, bgr, key,
    palette (r, g, b, a bytes)].  Returns [(status, bytes of native-endian T components)]."""
    images = list(images)
    n = len(images)
    descs = (ColorDesc * n)()
    keep = []
    for i, g in enumerate(images):
        st = bytes(g["storage"])
        count = len(st) // (_CHANNELS[g["color"]] * (2 if g["depth"] == 16 else 1))
        src = C.create_string_buffer(st, len(st)) if st else C.create_string_buffer(1)
        dst = C.create_string_buffer(max(count * _TARGET_BYTES[target], 1))
        keep.append((src, dst, count))
        descs[i].storage, descs[i].storage_len = _buf_addr(src), len(st)
        descs[i].pixels, descs[i].pixels_len = _buf_addr(dst), count * _TARGET_BYTES[target]
        descs[i].count = count
        _fill_format(descs[i].format, keep, g["color"], g["depth"], g.get("bgr"), g.get("key"), g.get("palette"))
    ctx.check(ctx._lib.pngb200_unpack_batch(ctx.handle, descs, n, target, alpha_mode, MEM_HOST))
    return [(descs[i].status, keep_i[1].raw[: keep_i[2] * _TARGET_BYTES[target]])
            for i, keep_i in enumerate(k for k in keep if isinstance(k, tuple))]


def pack_batch(ctx: Context, images, target: int = TARGET_RGBA8):
    """PNG.Image(packing:size:layout:) storage of [RGBA<T>] / [VA<T>] arrays.  images: dicts with
    pixels (bytes) and the PNG.Format fields.  Returns [bytes] (PNG.Image.storage)."""
    images = list(images)
    n = len(images)
    descs = (ColorDesc * n)()
    keep = []
    for i, g in enumerate(images):
        px = bytes(g["pixels"])
        count = len(px) // _TARGET_BYTES[target]
        size = count * _CHANNELS[g["color"]] * (2 if g["depth"] == 16 else 1)
        src = C.create_string_buffer(px, len(px)) if px else C.create_string_buffer(1)
        dst = C.create_string_buffer(max(size, 1))
        keep.append((src, dst, size))
        descs[i].pixels, descs[i].pixels_len = _buf_addr(src), len(px)
        descs[i].storage, descs[i].storage_len = _buf_addr(dst), size
        descs[i].count = count
        _fill_format(descs[i].format, keep, g["color"], g["depth"], g.get("bgr"), g.get("key"), g.get("palette"))
    ctx.check(ctx._lib.pngb200_pack_batch(ctx.handle, descs, n, target, MEM_HOST))
    return [k[1].raw[: k[2]] for k in keep if isinstance(k, tuple)]


class PngImage:
    """What PNG.Image.decompress(stream:) returns, as far as the hot path goes: status (+ the Swift
    error's associated values), header, PNG.Format fields (`fields`, ready for unpack_batch) and storage."""

    def __init__(self, d: PngDesc, storage):
        self.status, self.err_a, self.err_b = d.status, d.err_a, d.err_b
        self.width, self.height, self.depth, self.color = d.width, d.height, d.depth, d.color
        self.interlaced, self.standard = bool(d.interlaced), d.standard
        self.idat_bytes, self.idat_chunks, self.chunks = d.idat_bytes, d.idat_chunks, d.chunks
        self.checksum, self.blocks, self.produced = d.checksum, d.blocks, d.produced
        f = d.format
        self.fields = dict(color=f.color, depth=f.depth, bgr=bool(f.bgr),
                           key=tuple(f.key[: 1 if f.color == 0 else 3]) if f.has_key else None,
                           palette=bytes(d.palette_rgba[: 4 * f.palette_count]) if f.color == 3 else None)
        self.storage = storage


def _png_descs(files):
    files = [bytes(f) for f in files]
    descs = (PngDesc * max(len(files), 1))()
    keep = []
    for i, f in enumerate(files):
        src = C.create_string_buffer(f, len(f)) if f else C.create_string_buffer(1)
        keep.append(src)
        descs[i].file, descs[i].file_len = _buf_addr(src), len(f)
    return files, descs, keep


def png_inspect(files):
    """pngb200_png_inspect_batch: header walk only (no CRC check, no GPU).  Returns [PngImage]."""
    files, descs, keep = _png_descs(files)
    rc = lib().pngb200_png_inspect_batch(descs, len(files))
    if rc != OK:
        raise PNGB200Error(rc, "png_inspect")
    return [PngImage(descs[i], None) for i in range(len(files))]


def png_decode_batch(ctx: Context, files):
    """PNG.Image.decompress(stream:) over a batch of PNG files (bytes).  Returns [PngImage]."""
    files, descs, keep = _png_descs(files)
    n = len(files)
    rc = ctx._lib.pngb200_png_inspect_batch(descs, n)
    if rc != OK:
        raise PNGB200Error(rc, "png_inspect")
    outs = []
    for i in range(n):
        dst = C.create_string_buffer(max(int(descs[i].storage_size), 1))
        outs.append(dst)
        descs[i].pixels, descs[i].pixels_cap = _buf_addr(dst), int(descs[i].storage_size)
    ctx.check(ctx._lib.pngb200_png_decode_batch(ctx.handle, descs, n, MEM_HOST))
    return [PngImage(descs[i], outs[i].raw[: descs[i].storage_size] if descs[i].status == OK else None) for i in range(n)]


def png_encode_batch(ctx: Context, images, level: int = 9, idat_chunk: int = 0):
    """PNG.Image.compress(stream:level:) over a batch.  images: dicts with storage, width, height,
    [interlaced] and the PNG.Format fields.  Returns [(status, file bytes)]."""
    images = list(images)
    n = len(images)
    descs = (PngEncodeDesc * max(n, 1))()
    keep = []
    for i, g in enumerate(images):
        st = bytes(g["storage"])
        src = C.create_string_buffer(st, len(st))
        _fill_format(descs[i].format, keep, g["color"], g["depth"], g.get("bgr"), g.get("key"), g.get("palette"))
        il = int(bool(g.get("interlaced", False)))
        cap = ctx._lib.pngb200_png_encode_bound(g["width"], g["height"], C.byref(descs[i].format), il, idat_chunk)
        dst = C.create_string_buffer(cap)
        keep.append((src, dst))
        descs[i].pixels, descs[i].pixels_len = _buf_addr(src), len(st)
        descs[i].width, descs[i].height, descs[i].interlaced = g["width"], g["height"], il
        descs[i].level = level if isinstance(level, int) else level[i]
        descs[i].idat_chunk = idat_chunk
        descs[i].file, descs[i].file_cap = _buf_addr(dst), cap
    ctx.check(ctx._lib.pngb200_png_encode_batch(ctx.handle, descs, n, MEM_HOST))
    pairs = [k for k in keep if isinstance(k, tuple)]
    return [(descs[i].status, pairs[i][1].raw[: descs[i].produced]) for i in range(n)]


def deflate_batch(ctx: Context, streams, level: int = 9, fmt: int = FORMAT_ZLIB, exponent: int = 15):
    """LZ77.Deflator(format:level:exponent:).push(data, last: true) + all pull()s, per stream.
    `level` / `fmt` may be ints or per-stream sequences.  Returns list of (status, bytes)."""
    streams = list(streams)
    n = len(streams)
    descs = (DeflateDesc * n)()
    keep = []
    for i, data in enumerate(streams):
        data = bytes(data)
        cap = ctx._lib.pngb200_deflate_bound(len(data))
        src = C.create_string_buffer(data, len(data)) if len(data) else C.create_string_buffer(1)
        dst = C.create_string_buffer(cap)
        keep.append((src, dst))
        descs[i].src = _buf_addr(src)
        descs[i].src_len = len(data)
        descs[i].dst = _buf_addr(dst)
        descs[i].dst_cap = cap
        descs[i].format = fmt if isinstance(fmt, int) else fmt[i]
        descs[i].level = level if isinstance(level, int) else level[i]
        descs[i].exponent = exponent if isinstance(exponent, int) else exponent[i]
    ctx.check(ctx._lib.pngb200_deflate_batch(ctx.handle, descs, n, MEM_HOST))
    return [(descs[i].status, keep[i][1].raw[: descs[i].produced]) for i in range(n)]


def gzip_archive(ctx: Context, data: bytes, level: int = 7) -> bytes:
    """Gzip.archive(bytes:level:)"""
    (st, out), = deflate_batch(ctx, [data], level, FORMAT_GZIP)
    if st != OK:
        raise PNGB200Error(st, "gzip_archive")
    return out


def encode_batch(ctx: Context, images, level: int = 9):
    """PNG.Encoder over a batch: storage -> concatenated IDAT payload (filter select + deflate).
    `images`: dicts with pixels, width, height, volume, depth, interlaced, fmt."""
    images = list(images)
    n = len(images)
    descs = (EncodeDesc * n)()
    keep = []
    for i, g in enumerate(images):
        px = bytes(g["pixels"])
        w, h, vol = g["width"], g["height"], g["volume"]
        il = bool(g.get("interlaced", False))
        cap = ctx._lib.pngb200_deflate_bound(filtered_size(w, h, vol, il))
        src = C.create_string_buffer(px, len(px))
        dst = C.create_string_buffer(cap)
        keep.append((src, dst))
        descs[i].pixels = _buf_addr(src)
        descs[i].pixels_len = len(px)
        descs[i].idat = _buf_addr(dst)
        descs[i].idat_cap = cap
        descs[i].width, descs[i].height = w, h
        descs[i].volume, descs[i].depth, descs[i].interlaced = vol, g["depth"], int(il)
        descs[i].format = g.get("fmt", FORMAT_ZLIB)
        descs[i].level = g.get("level", level)
    ctx.check(ctx._lib.pngb200_encode_batch(ctx.handle, descs, n, MEM_HOST))
    return [(descs[i].status, keep[i][1].raw[: descs[i].produced]) for i in range(n)]


class Deflator:
    """LZ77.Deflator value semantics (push(_:last:), pop(), pull()) over the GPU encoder."""

    def __init__(self, ctx: Context, fmt: int = FORMAT_ZLIB, level: int = 9, exponent: int = 15, chunk_bytes: int = 0):
        self.ctx = ctx
        L = ctx._lib
        self.handle = L.pngb200_deflator_create(ctx.handle, fmt, level, exponent, chunk_bytes)
        if not self.handle:
            raise PNGB200Error(ERR_BAD_ARGUMENT, "deflator_create")

    def close(self):
        if getattr(self, "handle", None):
            self.ctx._lib.pngb200_deflator_destroy(self.handle)
            self.handle = None

    __del__ = close

    def push(self, data: bytes, last: bool = False):
        self.ctx.check(self.ctx._lib.pngb200_deflator_push(self.handle, bytes(data), len(data), int(last)))

    def _take(self, fn):
        p, n = C.POINTER(C.c_uint8)(), C.c_size_t()
        got = fn(self.handle, C.byref(p), C.byref(n))
        if got < 0:
            raise PNGB200Error(got, "deflator")
        return C.string_at(p, n.value) if got else None

    def pop(self):
        """a complete block or None (Swift: pop() -> [UInt8]?)"""
        return self._take(self.ctx._lib.pngb200_deflator_pop)

    def pull(self):
        """a complete block, else the flushed rest, else None (Swift: pull() -> [UInt8]?)"""
        return self._take(self.ctx._lib.pngb200_deflator_pull)


class Inflator:
    """LZ77.Inflator / Gzip.Inflator value semantics over the GPU path (streaming push/pull)."""

    def __init__(self, ctx: Context, fmt: int = FORMAT_ZLIB):
        self.ctx = ctx
        self.handle = ctx._lib.pngb200_inflator_create(ctx.handle, fmt)
        if not self.handle:
            raise PNGB200Error(ERR_BAD_ARGUMENT, "inflator_create")

    def close(self):
        if getattr(self, "handle", None):
            self.ctx._lib.pngb200_inflator_destroy(self.handle)
            self.handle = None

    __del__ = close

    def push(self, data: bytes) -> int:
        """Returns OK when the stream is complete (Swift: nil), NEED_MORE_INPUT otherwise;
        raises PNGB200Error on a decompression error (Swift: throws)."""
        st = self.ctx._lib.pngb200_inflator_push(self.handle, bytes(data), len(data))
        if st < 0:
            a, b, s = C.c_uint32(), C.c_uint32(), C.c_int()
            self.ctx._lib.pngb200_inflator_error(self.handle, C.byref(s), C.byref(a), C.byref(b))
            e = PNGB200Error(st, self.ctx._lib.pngb200_last_error(self.ctx.handle).decode())
            e.payload = (a.value, b.value)
            raise e
        return st

    def pull(self, count: int):
        """Exactly `count` bytes or None (Swift: pull(_:) -> [UInt8]?)."""
        buf = C.create_string_buffer(max(count, 1))
        st = self.ctx._lib.pngb200_inflator_pull(self.handle, buf, count)
        return buf.raw[:count] if st == OK else None

    def pull_all(self) -> bytes:
        n = self.ctx._lib.pngb200_inflator_available(self.handle)
        buf = C.create_string_buffer(max(n, 1))
        got = self.ctx._lib.pngb200_inflator_pull_all(self.handle, buf, n)
        return buf.raw[:got]
