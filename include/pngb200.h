/*
 * pngb200.h -- C ABI of the B200-native PNG hot path (DEFLATE inflate + scanline unfilter, and
 * the encode-side mirrors), the drop-in boundary for swift-png's `PNG.Decoder` / `PNG.Encoder`
 * / `LZ77.Inflator` / `LZ77.Deflator` call sites.
 *
 * swift-png has no FFI of its own: its hot path sits behind internal Swift value types.  The
 * entry points below are what a Swift `CPNGB200` system-library target would bind (see
 * INTEGRATION.md for the module map and the replacement bodies).  Each one cites the reference
 * interface it replaces (paths relative to the swift-png checkout).
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every function returns a
 * pngb200_status; nothing throws or aborts.  Per-item results (status + payload mirroring the
 * Swift error enums' associated values) are written into the descriptor arrays.
 * A pngb200_ctx is bound to one CUDA device and owns one CUDA stream plus grow-only device
 * workspaces; it must be used by one thread at a time (same rule as the Swift structs).
 * There is NO CPU fallback: if CUDA is unavailable pngb200_ctx_create fails.
 */
#ifndef PNGB200_H
#define PNGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNGB200_VERSION 1

/* ---- status codes (negative = error).  The numbering is shared with oracle/oracle.h ---- */
typedef enum pngb200_status {
    PNGB200_OK                                   = 0,  /* stream complete (push returned nil) */
    PNGB200_NEED_MORE_INPUT                      = 1,  /* push returned () : truncated stream */
    /* LZ77.DecompressionError, Sources/LZ77/Inflator/LZ77.DecompressionError.swift:19-60 */
    PNGB200_ERR_STREAM_CHECKSUM                  = -1, /* payload a=declared b=computed */
    PNGB200_ERR_BLOCK_TYPE                       = -2, /* a=code */
    PNGB200_ERR_BLOCK_COUNT_PARITY               = -3, /* a=LEN b=NLEN */
    PNGB200_ERR_RUNLITERAL_SYMBOL_COUNT          = -4, /* a=count */
    PNGB200_ERR_CODELENGTH_HUFFMAN_TABLE         = -5,
    PNGB200_ERR_CODELENGTH_SEQUENCE              = -6,
    PNGB200_ERR_HUFFMAN_TABLE                    = -7,
    PNGB200_ERR_STRING_REFERENCE                 = -8,
    PNGB200_ERR_INVALID_SYMBOL                   = -9, /* stricter than the reference, see DESIGN.md */
    /* LZ77.StreamHeaderError, Sources/LZ77/Inflator/LZ77.StreamHeaderError.swift:5-28 */
    PNGB200_ERR_ZLIB_METHOD                      = -16, /* a=code */
    PNGB200_ERR_ZLIB_WINDOW                      = -17, /* a=exponent */
    PNGB200_ERR_ZLIB_CHECK_BITS                  = -18,
    PNGB200_ERR_ZLIB_DICTIONARY                  = -19,
    /* Gzip.StreamHeaderError, Sources/LZ77/Gzip/Gzip.StreamHeaderError.swift */
    PNGB200_ERR_GZIP_SIGIL                       = -32,
    PNGB200_ERR_GZIP_METHOD                      = -33, /* a=code */
    PNGB200_ERR_GZIP_FLAG_BITS                   = -34, /* a=flags */
    PNGB200_ERR_GZIP_HEADER_CHECKSUM_UNSUPPORTED = -35,
    /* PNG.DecodingError cases raised on this path, Sources/PNG/Decoding/PNG.Decoder.swift:51-55,
     * :142-147 and PNG.Context.swift:134-141 */
    PNGB200_ERR_PNG_EXTRANEOUS_IMAGE_DATA        = -48,
    PNGB200_ERR_PNG_EXTRANEOUS_COMPRESSED_DATA   = -49,
    PNGB200_ERR_PNG_INCOMPLETE_DATASTREAM        = -50,
    PNGB200_ERR_PNG_PALETTE_INDEX                = -51, /* indexed pixel beyond the palette: the reference traps */
    /* PNG.LexingError (Sources/PNG/Lexing/PNG.LexingError.swift), file-level entry points */
    PNGB200_ERR_LEX_TRUNCATED_SIGNATURE          = -80,
    PNGB200_ERR_LEX_INVALID_SIGNATURE            = -81, /* a,b = the eight bytes found */
    PNGB200_ERR_LEX_TRUNCATED_CHUNK_HEADER       = -82,
    PNGB200_ERR_LEX_TRUNCATED_CHUNK_BODY         = -83, /* a = expected bytes */
    PNGB200_ERR_LEX_INVALID_CHUNK_TYPE           = -84, /* a = type code */
    PNGB200_ERR_LEX_INVALID_CHUNK_CHECKSUM       = -85, /* a = declared, b = computed */
    /* PNG.ParsingError (Sources/PNG/Parsing/PNG.ParsingError.swift): what IHDR / PLTE / tRNS can raise */
    PNGB200_ERR_PARSE_HEADER_CHUNK_LENGTH        = -96,  /* a = length */
    PNGB200_ERR_PARSE_HEADER_PIXEL_FORMAT_CODE   = -97,  /* a = depth code, b = colour code */
    PNGB200_ERR_PARSE_HEADER_PIXEL_FORMAT        = -98,  /* not allowed by the ios standard */
    PNGB200_ERR_PARSE_HEADER_COMPRESSION_CODE    = -99,  /* a = code */
    PNGB200_ERR_PARSE_HEADER_FILTER_CODE         = -100, /* a = code */
    PNGB200_ERR_PARSE_HEADER_INTERLACING_CODE    = -101, /* a = code */
    PNGB200_ERR_PARSE_HEADER_SIZE                = -102, /* a = x, b = y */
    PNGB200_ERR_PARSE_UNEXPECTED_PALETTE         = -103,
    PNGB200_ERR_PARSE_PALETTE_CHUNK_LENGTH       = -104, /* a = length */
    PNGB200_ERR_PARSE_PALETTE_COUNT              = -105, /* a = count, b = max */
    PNGB200_ERR_PARSE_UNEXPECTED_TRANSPARENCY    = -106,
    PNGB200_ERR_PARSE_TRANSPARENCY_CHUNK_LENGTH  = -107, /* a = length, b = expected */
    PNGB200_ERR_PARSE_TRANSPARENCY_SAMPLE        = -108, /* a = sample, b = max */
    PNGB200_ERR_PARSE_TRANSPARENCY_COUNT         = -109, /* a = count, b = max */
    /* PNG.DecodingError (Sources/PNG/Decoding/PNG.DecodingError.swift): a = chunk, b = the other chunk */
    PNGB200_ERR_DECODE_REQUIRED_CHUNK            = -112,
    PNGB200_ERR_DECODE_DUPLICATE_CHUNK           = -113,
    PNGB200_ERR_DECODE_UNEXPECTED_CHUNK          = -114,
    /* API-level */
    PNGB200_ERR_OUTPUT_CAPACITY                  = -64,
    PNGB200_ERR_BAD_ARGUMENT                     = -65,
    PNGB200_ERR_CUDA                             = -66, /* see pngb200_last_error */
    PNGB200_ERR_INTERNAL                         = -67
} pngb200_status;

/* LZ77.Format (.zlib, .ios) + Gzip.Format, Sources/LZ77/Wrappers/LZ77.Format.swift:8-12 */
typedef enum pngb200_format { PNGB200_FORMAT_ZLIB = 0, PNGB200_FORMAT_IOS = 1, PNGB200_FORMAT_GZIP = 2 } pngb200_format;

/* where the data pointers of a batch live */
typedef enum pngb200_memspace {
    PNGB200_MEM_HOST   = 0, /* host pointers (pinned or pageable); copies are part of the call */
    PNGB200_MEM_DEVICE = 1  /* device pointers on the context's GPU; nothing is copied */
} pngb200_memspace;

/* ---- context ---- */
typedef struct pngb200_ctx pngb200_ctx;

/* device < 0: the calling thread's current CUDA device.  Returns NULL on failure (no GPU, no
 * sm_100 device, out of memory); pngb200_last_error(NULL) then describes why. */
pngb200_ctx* pngb200_ctx_create(int device);
void         pngb200_ctx_destroy(pngb200_ctx* ctx);
/* Return the context's grow-only device arenas (and its pipeline lanes') to the driver; the next
 * batch call allocates again.  Fails with BAD_ARGUMENT while a decode batch is pending. */
int          pngb200_ctx_trim(pngb200_ctx* ctx);
const char*  pngb200_last_error(const pngb200_ctx* ctx);
/* the cudaStream_t all of this context's work is enqueued on (for event timing / interop) */
void*        pngb200_ctx_stream(pngb200_ctx* ctx);
int          pngb200_ctx_device(const pngb200_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t     pngb200_ctx_launch_count(const pngb200_ctx* ctx);
/* tuning knob: 0 = automatic, 1 = force the one-warp-per-stream inflate kernel,
 * 2 = force the block-parallel inflate kernel */
void         pngb200_ctx_set_inflate_mode(pngb200_ctx* ctx, int mode);
/* device time (CUDA events on the context's stream) of the last completed decode batch's stages:
 * ms[0] inflate kernels, ms[1] checksum kernels, ms[2] unfilter kernels */
int          pngb200_ctx_stage_ms(pngb200_ctx* ctx, float ms[3]);
/* device-side counters of the last finished batch of `count` streams/images, summed:
 * out[0] waves, out[1] sync rounds, out[2] copy-resolve rounds, out[3] streams that fell back to
 * the serial decoder (the analogue of the reference's -DDUMP_LZ77_BLOCKS statistics) */
int          pngb200_ctx_inflate_stats(pngb200_ctx* ctx, size_t count, uint64_t out[4]);
/* the full counter set of the last finished batch, summed over its `count` streams (the reference prints the
 * same kind of numbers under -DDUMP_LZ77_BLOCKS / -DDUMP_LZ77_BLOCKS_STATISTICS,
 * Sources/LZ77/Inflator/LZ77.InflatorBuffers.Stream.swift:11,292-297,364-374,472-486):
 * out[0..3] as pngb200_ctx_inflate_stats (out[1] = tokens decoded by chain walks), out[4] tokens (literals +
 * matches), out[5] matches, out[6] matches whose source was produced in the same 8 KiB wave by another
 * thread (they wait in the deferred-copy list), out[7] DEFLATE blocks, out[8..19] SM cycles per phase of
 * inflate_wave_kernel as seen by thread 0 of each CTA: header+tables, stage, speculate, walk, chain,
 * count+scan, emit, resolve, store, stored blocks, (2 spare); out[20..23] reserved */
int          pngb200_ctx_inflate_counters(pngb200_ctx* ctx, size_t count, uint64_t out[24]);
/* more than one CTA per stream: a batch with fewer big streams than half the CTA slots has its streams cut at
 * DEFLATE block boundaries and decoded segment by segment (csrc/inflate_segments.cuh).  Of the last inflate /
 * decode batch on this context: out[0] streams that were cut, out[1] segments they were cut into, out[2] streams
 * whose segments did not line up and that were decoded whole after all (results are identical either way). */
int          pngb200_ctx_segment_stats(pngb200_ctx* ctx, uint64_t out[3]);
/* which whole-stream engine the last inflate / decode batch on this context launched: 0 inflate_parallel_kernel
 * (round 1), 1 inflate_wave_kernel (ring window), 2 inflate_cells_kernel, -1 none (tiny streams, segments only) */
int          pngb200_ctx_last_inflate_engine(pngb200_ctx* ctx);
/* scanlines per filter type of the last decode / unfilter batch that went through the wavefront kernel
 * (non-interlaced, >= 8 bits per sample): out[0..4] = None, Sub, Up, Average, Paeth, out[5] = rows with an
 * invalid filter byte (left unchanged, as the reference does).  The device-side form of the reference's
 * -DDUMP_FILTERED_SCANLINES output (Sources/PNG/Decoding/PNG.Decoder.swift:96-98,128). */
int          pngb200_ctx_filter_histogram(pngb200_ctx* ctx, uint64_t out[6]);
/* inflate_mode: 0 automatic; 1 one warp per stream; 2 a whole CTA per stream, never cut; 3 / 4 force the
 * ring-window / the round-1 intra-stream kernel; 5 as 0; 6 force the cell kernel (inflate_cells.cuh) */

/* ---- batched one-shot entry points (the throughput path) ---- */

/* One standalone DEFLATE stream: LZ77.Inflator(format:).push(all); pull()  /  Gzip.extract.
 * Replaces Sources/LZ77/Inflator/LZ77.Inflator.swift:30-61 and Sources/LZ77/Gzip/Gzip.swift:6-11. */
typedef struct pngb200_stream_desc {
    const uint8_t* src;       /* compressed stream */
    size_t         src_len;
    uint8_t*       dst;       /* inflated bytes */
    size_t         dst_cap;
    int32_t        format;    /* pngb200_format */
    /* results */
    int32_t        status;    /* pngb200_status */
    uint32_t       err_a, err_b;
    uint32_t       checksum;  /* Adler-32 (zlib, ios) or CRC-32 (gzip) of the output */
    uint32_t       blocks;    /* DEFLATE blocks decoded */
    uint64_t       produced;  /* bytes written to dst */
    uint64_t       consumed_bits;
} pngb200_stream_desc;

int pngb200_inflate_batch(pngb200_ctx* ctx, pngb200_stream_desc* streams, size_t count, int memspace);

/* One PNG image's IDAT stream: PNG.Decoder.push(data, size:, pixel:, delegate: image.assign)
 * over the concatenated IDAT payload.  Replaces Sources/PNG/Decoding/PNG.Decoder.swift:47-149
 * (+ defilter :152-196, PNG.paeth PNG.swift:124-147) and PNG.Image.assign
 * (Sources/PNG/PNG.Image.swift:186-285).  `pixels` receives PNG.Image.storage exactly:
 * width*height*((volume+7)>>3) bytes, row-major, 16-bit samples big-endian, sub-byte depths
 * expanded to one byte per pixel. */
typedef struct pngb200_image_desc {
    const uint8_t* idat;      /* concatenated IDAT payloads (a zlib stream; raw DEFLATE for .ios) */
    size_t         idat_len;
    uint8_t*       pixels;
    size_t         pixels_cap;
    uint32_t       width, height;
    uint8_t        volume;    /* bits per pixel, PNG.Format.Pixel.volume */
    uint8_t        depth;     /* bits per sample */
    uint8_t        interlaced;
    uint8_t        format;    /* PNGB200_FORMAT_ZLIB (PNG.Standard.common) or _IOS */
    /* results */
    int32_t        status;
    uint32_t       err_a, err_b;
    uint32_t       checksum;  /* Adler-32 of the filtered stream */
    uint32_t       blocks;
    uint64_t       produced;  /* inflated (filtered) bytes */
} pngb200_image_desc;

int pngb200_decode_batch(pngb200_ctx* ctx, pngb200_image_desc* images, size_t count, int memspace);
/* same, split so that device time can be bracketed with events: enqueue returns once all work is
 * queued on pngb200_ctx_stream; finish synchronises and fills in the result fields. */
int pngb200_decode_batch_enqueue(pngb200_ctx* ctx, pngb200_image_desc* images, size_t count, int memspace);
int pngb200_decode_batch_finish(pngb200_ctx* ctx, pngb200_image_desc* images, size_t count);

/* The unfilter stage alone (PNG.Decoder.defilter + PNG.Image.assign over an already inflated
 * stream); `idat`/`idat_len` hold the FILTERED bytes.  Exposed for kernel-level tests/benchmarks. */
int pngb200_unfilter_batch(pngb200_ctx* ctx, pngb200_image_desc* images, size_t count, int memspace);

/* Encode-side stage 1: PNG.Image.collect + PNG.Encoder.filter for every row
 * (Sources/PNG/Encoding/PNG.Encoder.swift:132-204,230-234; PNG.Image.swift:431-544).
 * `pixels` is the INPUT (PNG.Image.storage), `idat`... see pngb200_filter_desc. */
typedef struct pngb200_filter_desc {
    const uint8_t* pixels;    /* PNG.Image.storage */
    size_t         pixels_len;
    uint8_t*       filtered;  /* out: height*(pitch+1) bytes (sum over Adam7 passes if interlaced) */
    size_t         filtered_cap;
    uint32_t       width, height;
    uint8_t        volume, depth, interlaced, reserved;
    int32_t        status;
    uint64_t       produced;
} pngb200_filter_desc;

int pngb200_filter_batch(pngb200_ctx* ctx, pngb200_filter_desc* images, size_t count, int memspace);

/* ---- colour targets ------------------------------------------------
 * image.unpack(as: PNG.RGBA<T>.self) / PNG.VA<T> and PNG.Image.init(packing:size:layout:) with the
 * default deindexer / indexer, T = UInt8 or UInt16: Sources/PNG/ColorTargets/PNG.RGBA.swift:262-478,
 * PNG.VA.swift, PNG.Color.swift, over the convolve / deconvolve closures of Sources/PNG/PNG.swift:149-1285.
 * The unpack is inside the reference's own timed decode loop (Benchmarks/Decompression/Swift/Main.swift:105-106).
 * For rgba8 -> RGBA<UInt8> and va8 -> VA<UInt8> the unpacked array IS PNG.Image.storage, byte for
 * byte, so pngb200_decode_batch's output needs no second pass. */
typedef enum pngb200_target {
    PNGB200_TARGET_RGBA8 = 0, PNGB200_TARGET_RGBA16 = 1, PNGB200_TARGET_VA8 = 2, PNGB200_TARGET_VA16 = 3
} pngb200_target;
/* applied per pixel after unpacking: .premultiplied / .straightened (PNG.RGBA.swift:115-121, 163-169),
 * or premultiplied(as: UInt8.self) / straightened(as: UInt8.self) of a 16-bit target (:141-155, 187-201) */
typedef enum pngb200_alpha_mode {
    PNGB200_ALPHA_ASIS = 0, PNGB200_ALPHA_PREMULTIPLIED = 1, PNGB200_ALPHA_STRAIGHTENED = 2,
    PNGB200_ALPHA_PREMULTIPLIED_AS8 = 3, PNGB200_ALPHA_STRAIGHTENED_AS8 = 4
} pngb200_alpha_mode;
/* PNG.Format (Sources/PNG/Formats/PNG.Format.swift:6-43) as these kernels see it */
typedef struct pngb200_pixel_format {
    uint8_t        color;         /* PNG colour type: 0 v, 2 rgb, 3 indexed, 4 va, 6 rgba */
    uint8_t        depth;         /* bits per sample: 1, 2, 4, 8, 16 */
    uint8_t        bgr;           /* 1: .bgr8 / .bgra8 (the ios standard's sample order) */
    uint8_t        has_key;       /* chroma key present (v / rgb / bgr formats) */
    uint16_t       key[3];        /* raw sample values in STORAGE order (Format.recognize, :161-330) */
    uint16_t       palette_count; /* indexed formats: entries in `palette` */
    const uint8_t* palette;       /* palette_count x (r, g, b, a), tRNS merged; always HOST memory */
} pngb200_pixel_format;
typedef struct pngb200_color_desc {
    void*                storage;      /* PNG.Image.storage: unpack reads it, pack writes it */
    size_t               storage_len;  /* bytes (capacity for pack) */
    void*                pixels;       /* [RGBA<T>] / [VA<T>]: native-endian T components, aligned to the pixel size */
    size_t               pixels_len;   /* bytes (capacity for unpack) */
    uint64_t             count;        /* number of pixels */
    pngb200_pixel_format format;
    int32_t              status;       /* out: PNGB200_OK or PNGB200_ERR_PNG_PALETTE_INDEX */
} pngb200_color_desc;
int pngb200_unpack_batch(pngb200_ctx* ctx, pngb200_color_desc* images, size_t count, int target,
                         int alpha_mode, int memspace);
int pngb200_pack_batch(pngb200_ctx* ctx, pngb200_color_desc* images, size_t count, int target, int memspace);

/* Encode-side stage 2: LZ77.Deflator(format:level:exponent:hint:).push(src, last: true) and the
 * concatenation of every pull() -- the reference's compressed bytes, bit for bit (its output does
 * not depend on push granularity or on `hint`).  Replaces Sources/LZ77/Deflator/LZ77.Deflator.swift:8-44,
 * Sources/LZ77/Gzip/Gzip.swift:34-46 (Gzip.archive) and everything behind them.  level 0...13 as in
 * LZ77.DeflatorSearch (0-3 greedy, 4-7 lazy, 8-13 full); exponent 8...15. */
typedef struct pngb200_deflate_desc {
    const uint8_t* src;
    size_t         src_len;
    uint8_t*       dst;
    size_t         dst_cap;   /* pngb200_deflate_bound(src_len) is always enough */
    int32_t        format;    /* pngb200_format */
    int32_t        level;
    int32_t        exponent;
    /* results */
    int32_t        status;
    uint32_t       checksum;  /* Adler-32 / CRC-32 of src as written into the trailer */
    uint32_t       blocks;
    uint64_t       produced;
} pngb200_deflate_desc;

int    pngb200_deflate_batch(pngb200_ctx* ctx, pngb200_deflate_desc* streams, size_t count, int memspace);
size_t pngb200_deflate_bound(size_t src_len);

/* PNG.Encoder.pull over a whole image: collect + filter + deflate (format .zlib / .ios) -> the
 * concatenated IDAT payload.  Replaces Sources/PNG/Encoding/PNG.Encoder.swift:33-129. */
typedef struct pngb200_encode_desc {
    const uint8_t* pixels;    /* PNG.Image.storage */
    size_t         pixels_len;
    uint8_t*       idat;      /* out: concatenated IDAT payload */
    size_t         idat_cap;
    uint32_t       width, height;
    uint8_t        volume, depth, interlaced, format;
    int32_t        level;
    int32_t        status;
    uint32_t       checksum, blocks;
    uint64_t       produced;
} pngb200_encode_desc;

int pngb200_encode_batch(pngb200_ctx* ctx, pngb200_encode_desc* images, size_t count, int memspace);

/* ---- whole PNG files ------------------------------------------------
 * PNG.Image.decompress(stream:) and PNG.Image.compress(stream:level:hint:) at file level
 * (Sources/PNG/PNG.Image.swift:298-401, 576-670): signature, chunk framing, per-chunk CRC-32
 * (Lexing/PNG.BytestreamSource.swift:17-83, PNG.BytestreamDestination.swift:66-95), IHDR / PLTE / tRNS,
 * the ordering rules that involve them, IDAT concatenation and framing.  The host reads chunk HEADERS
 * only; CRC-32 of every chunk, IDAT gather / scatter and the codec run on the device.  Ancillary chunks
 * other than PLTE / tRNS / bKGD are CRC-checked and otherwise ignored (metadata is not on the hot path).
 * Errors come back in the order the reference's streaming loop meets them. */
typedef struct pngb200_png_desc {
    const uint8_t*       file;          /* in: the PNG file, HOST memory */
    size_t               file_len;
    void*                pixels;        /* out: PNG.Image.storage (memspace of the call) */
    size_t               pixels_cap;    /* >= storage_size (pngb200_png_inspect_batch reports it) */
    /* out: PNG.Header + PNG.Layout.format */
    uint32_t             width, height;
    uint8_t              depth, color, interlaced, standard; /* standard: 0 common, 1 ios (CgBI) */
    pngb200_pixel_format format;        /* ready for pngb200_unpack_batch; palette -> palette_rgba */
    uint8_t              palette_rgba[1024];
    uint64_t             storage_size;  /* width * height * bytes per pixel */
    uint64_t             idat_bytes;    /* concatenated IDAT payload */
    uint32_t             idat_chunks, chunks;
    int32_t              status;        /* pngb200_status */
    uint32_t             err_a, err_b;
    uint32_t             checksum, blocks;
    uint64_t             produced;
} pngb200_png_desc;
/* host only: walk the chunk headers, parse IHDR / PLTE / tRNS, fill the out fields (no CRC check, no
 * GPU) -- what a caller needs to size `pixels` */
int pngb200_png_inspect_batch(pngb200_png_desc* files, size_t count);
/* `memspace` is where `pixels` live; files are always host memory */
int pngb200_png_decode_batch(pngb200_ctx* ctx, pngb200_png_desc* files, size_t count, int memspace);

typedef struct pngb200_png_encode_desc {
    const void*          pixels;        /* in: PNG.Image.storage (memspace of the call) */
    size_t               pixels_len;
    uint32_t             width, height;
    pngb200_pixel_format format;        /* bgr = 1 writes the ios standard (CgBI chunk, raw deflate) */
    uint8_t              interlaced;
    int32_t              level;         /* 0...13 */
    uint32_t             idat_chunk;    /* bytes per IDAT chunk; 0 = 65544, what the reference emits for its
                                           default hint (2 x the capacity malloc gives DeflatorOut's buffer) */
    uint8_t*             file;          /* out: the PNG file, HOST memory */
    size_t               file_cap;      /* >= pngb200_png_encode_bound(...) */
    int32_t              status;
    uint32_t             checksum, blocks;
    uint64_t             produced;      /* file bytes */
} pngb200_png_encode_desc;
size_t pngb200_png_encode_bound(uint32_t width, uint32_t height, const pngb200_pixel_format* format, int interlaced,
                                uint32_t idat_chunk);
int    pngb200_png_encode_batch(pngb200_ctx* ctx, pngb200_png_encode_desc* images, size_t count, int memspace);

/* size helpers (host arithmetic only) */
size_t pngb200_filtered_size(uint32_t width, uint32_t height, int volume, int interlaced);
size_t pngb200_storage_size(uint32_t width, uint32_t height, int volume);

/* ---- streaming handles (LZ77.Inflator value-type semantics, layered on the batch path) ---- */
typedef struct pngb200_inflator pngb200_inflator;
/* LZ77.Inflator.init(format:) / Gzip.Inflator.init(), LZ77.Inflator.swift:18-23 */
pngb200_inflator* pngb200_inflator_create(pngb200_ctx* ctx, int format);
void              pngb200_inflator_destroy(pngb200_inflator* z);
/* push(_:) : copies `data`; returns PNGB200_OK when the stream is complete (Swift: nil),
 * PNGB200_NEED_MORE_INPUT when it wants more (Swift: ()), <0 on error (Swift: throws). */
int    pngb200_inflator_push(pngb200_inflator* z, const uint8_t* data, size_t n);
/* pull(_ count:) : exactly `count` bytes or PNGB200_NEED_MORE_INPUT (Swift: nil) */
int    pngb200_inflator_pull(pngb200_inflator* z, uint8_t* dst, size_t count);
/* pull() : everything available; returns the number of bytes copied (<= cap) */
size_t pngb200_inflator_pull_all(pngb200_inflator* z, uint8_t* dst, size_t cap);
size_t pngb200_inflator_available(const pngb200_inflator* z);
void   pngb200_inflator_error(const pngb200_inflator* z, int* status, uint32_t* a, uint32_t* b);

/* LZ77.Deflator value-type semantics (Sources/LZ77/Deflator/LZ77.Deflator.swift:8-44; the call sites are
 * PNG.Encoder.pull, Sources/PNG/Encoding/PNG.Encoder.swift:68,85,101,117,121,128).
 * init(format:level:exponent:hint:) -- `chunk_bytes` is the size of a complete output block: the reference hands out
 * 2 * capacity bytes, capacity being whatever malloc grants for `hint` UInt16 atoms (LZ77.DeflatorOut.swift:15-27,
 * 109-135); 0 = 65544, the value behind the reference's committed outputs (hint 1 << 15).
 * The device compresses a stream in one launch, so compressed blocks become available when push(last:) arrives;
 * pop() before that returns "nil" where the reference might already have a block.  The sequence of blocks a
 * caller sees -- sizes and bytes -- is the reference's. */
typedef struct pngb200_deflator pngb200_deflator;
pngb200_deflator* pngb200_deflator_create(pngb200_ctx* ctx, int format, int level, int exponent, size_t chunk_bytes);
void              pngb200_deflator_destroy(pngb200_deflator* z);
/* push(_:last:) : copies `data`.  Returns PNGB200_OK, or < 0 when compressing (on last) failed */
int    pngb200_deflator_push(pngb200_deflator* z, const uint8_t* data, size_t n, int last);
/* pop() : a complete block (exactly chunk_bytes) -> returns 1 and sets *block / *n (valid until the next call on
 * this handle); 0 = nil */
int    pngb200_deflator_pop(pngb200_deflator* z, const uint8_t** block, size_t* n);
/* pull() : a complete block if there is one, else the flushed incomplete block (non-empty); 0 = nil */
int    pngb200_deflator_pull(pngb200_deflator* z, const uint8_t** block, size_t* n);

#ifdef __cplusplus
}
#endif
#endif /* PNGB200_H */
