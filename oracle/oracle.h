/*
 * oracle/oracle.h -- CPU restatement of swift-png's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or call this.  The product (swift-png_b200/) never does.
 *
 * Parity status: PINNED.  Decode is pinned by the reference's 161 PngSuite goldens
 * (Sources/PNGIntegrationTests/Inputs/Common/ *.png <-> RGBA/ *.png.rgba), its gzip fixtures
 * and zlib cross-checks; encode is pinned byte-for-byte by the reference's committed level-9
 * outputs (Tests/Outputs/ *.png) and its gzip fixtures (levels 10 and 13); see
 * tests/test_oracle_*.py.  The reference itself is Swift and no Swift toolchain exists in the
 * image, so it cannot be compiled into oracle/_ref.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference checkout).
 */
#ifndef PNGB200_ORACLE_H
#define PNGB200_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes shared with include/pngb200.h (same numbering) */
enum {
    ORC_OK                                   = 0,
    ORC_NEED_MORE_INPUT                      = 1,
    /* LZ77.DecompressionError (Sources/LZ77/Inflator/LZ77.DecompressionError.swift:19-60) */
    ORC_ERR_STREAM_CHECKSUM                  = -1,
    ORC_ERR_BLOCK_TYPE                       = -2,
    ORC_ERR_BLOCK_COUNT_PARITY               = -3,
    ORC_ERR_RUNLITERAL_SYMBOL_COUNT          = -4,
    ORC_ERR_CODELENGTH_HUFFMAN_TABLE         = -5,
    ORC_ERR_CODELENGTH_SEQUENCE              = -6,
    ORC_ERR_HUFFMAN_TABLE                    = -7,
    ORC_ERR_STRING_REFERENCE                 = -8,
    /* stricter than the reference: symbols the reference maps to padding rows */
    ORC_ERR_INVALID_SYMBOL                   = -9,
    /* LZ77.StreamHeaderError (Sources/LZ77/Inflator/LZ77.StreamHeaderError.swift:5-28) */
    ORC_ERR_ZLIB_METHOD                      = -16,
    ORC_ERR_ZLIB_WINDOW                      = -17,
    ORC_ERR_ZLIB_CHECK_BITS                  = -18,
    ORC_ERR_ZLIB_DICTIONARY                  = -19,
    /* Gzip.StreamHeaderError (Sources/LZ77/Gzip/Gzip.StreamHeaderError.swift) */
    ORC_ERR_GZIP_SIGIL                       = -32,
    ORC_ERR_GZIP_METHOD                      = -33,
    ORC_ERR_GZIP_FLAG_BITS                   = -34,
    ORC_ERR_GZIP_HEADER_CHECKSUM_UNSUPPORTED = -35,
    /* PNG.DecodingError cases raised by PNG.Decoder / PNG.Context */
    ORC_ERR_PNG_EXTRANEOUS_IMAGE_DATA        = -48,
    ORC_ERR_PNG_EXTRANEOUS_COMPRESSED_DATA   = -49,
    ORC_ERR_PNG_INCOMPLETE_DATASTREAM        = -50,
    /* API-level */
    ORC_ERR_OUTPUT_CAPACITY                  = -64,
    ORC_ERR_BAD_ARGUMENT                     = -65
};

enum { ORC_FORMAT_ZLIB = 0, ORC_FORMAT_IOS = 1, ORC_FORMAT_GZIP = 2 };

typedef struct {
    int32_t  status;
    uint32_t a, b;          /* error payload (declared/computed checksum, bad code, ...) */
    uint64_t consumed_bits; /* bit cursor when the call returned */
    uint64_t produced;      /* bytes written to `out` */
    uint32_t checksum;      /* Adler-32 (zlib/ios) or CRC-32 (gzip) of the output */
    uint32_t blocks;        /* number of DEFLATE blocks seen */
} orc_inflate_result;

/* One-shot LZ77.Inflator / Gzip.Inflator: push(all of `in`) then pull().
 * `out` may be NULL with cap 0 to measure; otherwise ORC_ERR_OUTPUT_CAPACITY if it overflows. */
void orc_inflate(int format, const uint8_t* in, size_t n, uint8_t* out, size_t cap,
                 orc_inflate_result* res);

uint32_t orc_adler32(uint32_t adler, const uint8_t* p, size_t n);
uint32_t orc_crc32(uint32_t crc, const uint8_t* p, size_t n);

/* PNG.Decoder.defilter (Sources/PNG/Decoding/PNG.Decoder.swift:152-196).  `line` and `last`
 * are pitch+1 bytes (filter byte first); in place. */
void orc_defilter(uint8_t* line, const uint8_t* last, size_t count, int delay);
uint8_t orc_paeth(uint8_t a, uint8_t b, uint8_t c);

/* PNG.Decoder.push over a complete filtered stream + PNG.Image.assign
 * (PNG.Decoder.swift:47-149, PNG.Image.swift:186-285).  `filtered` is the inflated IDAT stream.
 * volume = bits per pixel (PNG.Format.Pixel.volume); depth = bits per sample (only used to pick
 * the sub-byte expansion).  storage = w*h*((volume+7)>>3) bytes.  Returns ORC_OK,
 * ORC_ERR_PNG_EXTRANEOUS_IMAGE_DATA or ORC_NEED_MORE_INPUT (too few rows). */
int orc_png_unfilter(const uint8_t* filtered, size_t n, uint32_t w, uint32_t h, int volume,
                     int depth, int interlaced, uint8_t* storage);

/* whole PNG.Decoder path: inflate + unfilter + assign */
int orc_png_decode(int format, const uint8_t* idat, size_t n, uint32_t w, uint32_t h, int volume,
                   int depth, int interlaced, uint8_t* storage, orc_inflate_result* res);

/* PNG.Encoder.filter (Sources/PNG/Encoding/PNG.Encoder.swift:132-204): choose + apply; writes
 * pitch+1 bytes to `out`; `line`/`last` are pitch+1 bytes with a dummy byte 0. */
void orc_filter_row(const uint8_t* line, const uint8_t* last, size_t count, int delay,
                    uint8_t* out);
/* PNG.Image.collect + Encoder.filter over a whole image: storage -> filtered stream
 * (h*(pitch+1) bytes for non-interlaced; sum over passes for Adam7). returns bytes written */
size_t orc_png_filter(const uint8_t* storage, uint32_t w, uint32_t h, int volume, int depth,
                      int interlaced, uint8_t* filtered, size_t cap);
size_t orc_png_filtered_size(uint32_t w, uint32_t h, int volume, int interlaced);

/* LZ77.Deflator / Gzip.Deflator one-shot: push(in, last: true) then concatenated pull()s.
 * Returns bytes written or (size_t)-1 on capacity overflow. */
size_t orc_deflate(int format, int level, int exponent, const uint8_t* in, size_t n,
                   uint8_t* out, size_t cap);
size_t orc_deflate_bound(size_t n);
size_t orc_debug_block_starts(int format, const uint8_t* in, size_t n, uint64_t* trace, size_t cap);
size_t orc_debug_greedy_parse(const uint8_t* in, size_t n, int exponent, long attempts, int goal,
                              int* runs, int* dists, size_t cap);

/* ---- colour targets -------------------------------------------
 * PNG.Format as the unpack/pack kernels see it (Sources/PNG/Formats/PNG.Format.swift:6-43):
 * colour type + sample depth, sample order (bgr = 1 for the ios standard's bgr8/bgra8), the
 * chroma key in STORAGE sample order (Format.recognize, PNG.Format.swift:161-330, stores the key of
 * a bgr8 image as (b, g, r)), and for indexed formats the (r, g, b, a) palette with tRNS merged. */
typedef struct {
    uint8_t        color;   /* 0 v, 2 rgb, 3 indexed, 4 va, 6 rgba */
    uint8_t        depth;   /* 1, 2, 4, 8, 16 */
    uint8_t        bgr;
    uint8_t        has_key;
    uint16_t       key[3];
    uint16_t       palette_count;
    const uint8_t* palette; /* palette_count x 4 bytes */
} orc_format;

enum { ORC_TARGET_RGBA8 = 0, ORC_TARGET_RGBA16 = 1, ORC_TARGET_VA8 = 2, ORC_TARGET_VA16 = 3 };
/* 3 / 4: premultiplied(as: UInt8.self) / straightened(as: UInt8.self) of a 16-bit target
 * (PNG.RGBA.swift:141-155, 187-201): the arithmetic runs on the high bytes, results x 257 */
enum { ORC_ALPHA_ASIS = 0, ORC_ALPHA_PREMULTIPLIED = 1, ORC_ALPHA_STRAIGHTENED = 2,
       ORC_ALPHA_PREMULTIPLIED_AS8 = 3, ORC_ALPHA_STRAIGHTENED_AS8 = 4 };
enum { ORC_ERR_PALETTE_INDEX = -51 };  /* the reference traps (Swift array bounds) */

/* PNG.premultiply / PNG.straighten (Sources/PNG/PNG.swift:54-120) for T = UInt8 (bits 8) or UInt16 */
uint32_t orc_premultiply(uint32_t color, uint32_t alpha, int bits);
uint32_t orc_straighten(uint32_t premultiplied, uint32_t alpha, int bits);

/* PNG.RGBA<T>.unpack / PNG.VA<T>.unpack with the default deindexer
 * (ColorTargets/PNG.RGBA.swift:262-365, PNG.VA.swift, PNG.Color.swift), then optionally
 * `.premultiplied` / `.straightened` per pixel.  `storage` is PNG.Image.storage (pixels x bpp
 * bytes, 16-bit samples big-endian, sub-byte depths one byte per sample); `out` receives native
 * (little-endian) T components, 4 (RGBA) or 2 (VA) per pixel. */
int orc_unpack(const uint8_t* storage, size_t pixels, const orc_format* f, int target,
               int alpha_mode, void* out);
/* PNG.RGBA<T>.pack / PNG.VA<T>.pack with the default indexer (PNG.RGBA.swift:405-478,
 * PNG.Color.swift: colours missing from the palette map to entry 0) */
int orc_pack(const void* pixels, size_t n, const orc_format* f, int target, uint8_t* storage);

/* ---- PNG container ------------------------------------------------
 * PNG.Image.decompress(stream:) / compress(stream:level:hint:) restated at file level
 * (Sources/PNG/PNG.Image.swift:298-401, 576-670): signature, chunk framing and per-chunk CRC-32
 * (Lexing/PNG.BytestreamSource.swift:17-83, PNG.BytestreamDestination.swift:66-95), IHDR / PLTE /
 * tRNS parsing (Parsing/PNG.Header.swift:40-98, PNG.Palette.swift:27-55, PNG.Transparency.swift:68-122),
 * the chunk-order rules that involve those chunks, IDAT concatenation.  Ancillary chunks other than
 * PLTE / tRNS / bKGD are CRC-checked and otherwise ignored (metadata is outside the hot path). */
enum {
    /* PNG.LexingError (Lexing/PNG.LexingError.swift) */
    ORC_ERR_LEX_TRUNCATED_SIGNATURE    = -80,
    ORC_ERR_LEX_INVALID_SIGNATURE      = -81,
    ORC_ERR_LEX_TRUNCATED_CHUNK_HEADER = -82,
    ORC_ERR_LEX_TRUNCATED_CHUNK_BODY   = -83, /* a = expected bytes */
    ORC_ERR_LEX_INVALID_CHUNK_TYPE     = -84, /* a = type code */
    ORC_ERR_LEX_INVALID_CHUNK_CHECKSUM = -85, /* a = declared, b = computed */
    /* PNG.ParsingError (Parsing/PNG.ParsingError.swift), the cases IHDR / PLTE / tRNS can raise */
    ORC_ERR_PARSE_HEADER_CHUNK_LENGTH      = -96,  /* a = length */
    ORC_ERR_PARSE_HEADER_PIXEL_FORMAT_CODE = -97,  /* a = depth code, b = colour code */
    ORC_ERR_PARSE_HEADER_PIXEL_FORMAT      = -98,  /* pixel format not allowed by the ios standard */
    ORC_ERR_PARSE_HEADER_COMPRESSION_CODE  = -99,  /* a = code */
    ORC_ERR_PARSE_HEADER_FILTER_CODE       = -100, /* a = code */
    ORC_ERR_PARSE_HEADER_INTERLACING_CODE  = -101, /* a = code */
    ORC_ERR_PARSE_HEADER_SIZE              = -102, /* a = x, b = y */
    ORC_ERR_PARSE_UNEXPECTED_PALETTE       = -103,
    ORC_ERR_PARSE_PALETTE_CHUNK_LENGTH     = -104, /* a = length */
    ORC_ERR_PARSE_PALETTE_COUNT            = -105, /* a = count, b = max */
    ORC_ERR_PARSE_UNEXPECTED_TRANSPARENCY  = -106,
    ORC_ERR_PARSE_TRANSPARENCY_CHUNK_LENGTH = -107, /* a = length, b = expected */
    ORC_ERR_PARSE_TRANSPARENCY_SAMPLE      = -108, /* a = sample, b = max */
    ORC_ERR_PARSE_TRANSPARENCY_COUNT       = -109, /* a = count, b = max */
    /* PNG.DecodingError (Decoding/PNG.DecodingError.swift): a = chunk, b = the other chunk */
    ORC_ERR_DECODE_REQUIRED_CHUNK   = -112,
    ORC_ERR_DECODE_DUPLICATE_CHUNK  = -113,
    ORC_ERR_DECODE_UNEXPECTED_CHUNK = -114
};

typedef struct {
    int32_t    status;
    uint32_t   a, b;
    uint32_t   width, height;
    uint8_t    depth, color, interlaced, standard; /* standard: 0 common, 1 ios (CgBI) */
    orc_format format;                             /* format.palette points at palette_rgba */
    uint8_t    palette_rgba[1024];
    uint64_t   idat_bytes;                         /* concatenated IDAT payload */
    uint32_t   idat_chunks, chunks;
    orc_inflate_result inflate;                    /* orc_png_decompress only */
} orc_png_info;

/* lex + parse the whole file (every chunk's CRC included), no image data decoded */
int orc_png_inspect(const uint8_t* file, size_t n, orc_png_info* info);
/* PNG.Image.decompress(stream:): storage gets w*h*bpp bytes; errors in the order the reference's
 * streaming loop would meet them */
int orc_png_decompress(const uint8_t* file, size_t n, orc_png_info* info, uint8_t* storage, size_t cap);
/* PNG.Image.compress(stream:level:hint:) for an image without metadata: signature, [CgBI], IHDR,
 * [PLTE], [tRNS], IDAT chunks of `idat_chunk` bytes (65544 in the reference's committed outputs:
 * 2 x the ManagedBuffer capacity DeflatorOut gets for hint 1 << 15), IEND.  Returns bytes written. */
size_t orc_png_compress(const uint8_t* storage, uint32_t w, uint32_t h, const orc_format* f,
                        int interlaced, int level, size_t idat_chunk, uint8_t* out, size_t cap);
size_t orc_png_compress_bound(uint32_t w, uint32_t h, const orc_format* f, int interlaced, size_t idat_chunk);

#ifdef __cplusplus
}
#endif
#endif
