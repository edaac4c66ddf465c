/*
 * oracle/png_container.c -- CPU restatement of swift-png's container layer at file level:
 * PNG.Image.decompress(stream:) and PNG.Image.compress(stream:level:hint:).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Follows Sources/PNG/PNG.Image.swift:298-401 (decode loop) and :576-670 (encode), with the
 * lexer of Lexing/PNG.BytestreamSource.swift:17-83, the chunk-type rule of Lexing/PNG.Chunk.swift:39-58,
 * and the IHDR / PLTE / tRNS parsers.  Ancillary chunks other than PLTE / tRNS / bKGD are lexed
 * (CRC included) and ignored.
 */
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define FOURCC(a, b, c, d) ((uint32_t)(a) << 24 | (uint32_t)(b) << 16 | (uint32_t)(c) << 8 | (uint32_t)(d))
enum {
    T_CgBI = FOURCC('C', 'g', 'B', 'I'), T_IHDR = FOURCC('I', 'H', 'D', 'R'), T_PLTE = FOURCC('P', 'L', 'T', 'E'),
    T_IDAT = FOURCC('I', 'D', 'A', 'T'), T_IEND = FOURCC('I', 'E', 'N', 'D'), T_tRNS = FOURCC('t', 'R', 'N', 'S'),
    T_bKGD = FOURCC('b', 'K', 'G', 'D')
};
static const uint8_t SIGNATURE[8] = {137, 80, 78, 71, 13, 10, 26, 10};
static const uint32_t PUBLIC_CHUNKS[] = {
    FOURCC('C', 'g', 'B', 'I'), FOURCC('I', 'H', 'D', 'R'), FOURCC('P', 'L', 'T', 'E'), FOURCC('I', 'D', 'A', 'T'),
    FOURCC('I', 'E', 'N', 'D'), FOURCC('c', 'H', 'R', 'M'), FOURCC('g', 'A', 'M', 'A'), FOURCC('i', 'C', 'C', 'P'),
    FOURCC('s', 'B', 'I', 'T'), FOURCC('s', 'R', 'G', 'B'), FOURCC('b', 'K', 'G', 'D'), FOURCC('h', 'I', 'S', 'T'),
    FOURCC('t', 'R', 'N', 'S'), FOURCC('p', 'H', 'Y', 's'), FOURCC('s', 'P', 'L', 'T'), FOURCC('t', 'I', 'M', 'E'),
    FOURCC('i', 'T', 'X', 't'), FOURCC('t', 'E', 'X', 't'), FOURCC('z', 'T', 'X', 't')};

static uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
static uint32_t be16(const uint8_t* p) { return (uint32_t)p[0] << 8 | p[1]; }
static void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24), p[1] = (uint8_t)(v >> 16), p[2] = (uint8_t)(v >> 8), p[3] = (uint8_t)v; }

/* PNG.Chunk.init(validating:) (Lexing/PNG.Chunk.swift:39-58): public chunks by name; anything else
 * must be a private (bit 5 of byte 1 set ... the reference tests name & 0x20002000 == 0x20000000:
 * ancillary bit set, reserved bit clear) */
static int chunk_type_valid(uint32_t name)
{
    for (size_t i = 0; i < sizeof PUBLIC_CHUNKS / sizeof *PUBLIC_CHUNKS; ++i)
        if (PUBLIC_CHUNKS[i] == name) return 1;
    return (name & 0x20002000u) == 0x20000000u;
}

static int fail(orc_png_info* info, int status, uint32_t a, uint32_t b)
{
    info->status = status, info->a = a, info->b = b;
    return status;
}

typedef struct {
    const uint8_t* file;
    size_t         n, at;
} lexer;

/* BytestreamSource.chunk() (PNG.BytestreamSource.swift:33-83) */
static int lex_chunk(lexer* lx, orc_png_info* info, uint32_t* type, const uint8_t** data, uint32_t* len)
{
    if (lx->n - lx->at < 8) return fail(info, ORC_ERR_LEX_TRUNCATED_CHUNK_HEADER, 0, 0);
    const uint32_t length = be32(lx->file + lx->at), name = be32(lx->file + lx->at + 4);
    if (!chunk_type_valid(name)) return fail(info, ORC_ERR_LEX_INVALID_CHUNK_TYPE, name, 0);
    if ((uint64_t)(lx->n - lx->at - 8) < (uint64_t)length + 4)
        return fail(info, ORC_ERR_LEX_TRUNCATED_CHUNK_BODY, length + 4, 0);
    const uint8_t* body = lx->file + lx->at + 8;
    const uint32_t declared = be32(body + length), computed = orc_crc32(0, lx->file + lx->at + 4, (size_t)length + 4);
    if (declared != computed) return fail(info, ORC_ERR_LEX_INVALID_CHUNK_CHECKSUM, declared, computed);
    *type = name, *data = body, *len = length;
    lx->at += 12 + (size_t)length;
    info->chunks++;
    return ORC_OK;
}

static int channels_of(int color) { return color == 0 || color == 3 ? 1 : color == 2 ? 3 : color == 4 ? 2 : 4; }

/* PNG.Header.init(parsing:standard:) (Parsing/PNG.Header.swift:40-98) */
static int parse_header(orc_png_info* info, const uint8_t* d, uint32_t len)
{
    if (len != 13) return fail(info, ORC_ERR_PARSE_HEADER_CHUNK_LENGTH, len, 0);
    const int depth = d[8], color = d[9];
    int ok;  /* PNG.Format.Pixel.recognize(code:) */
    switch (color) {
    case 0: ok = depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16; break;
    case 3: ok = depth == 1 || depth == 2 || depth == 4 || depth == 8; break;
    case 2: case 4: case 6: ok = depth == 8 || depth == 16; break;
    default: ok = 0;
    }
    if (!ok) return fail(info, ORC_ERR_PARSE_HEADER_PIXEL_FORMAT_CODE, (uint32_t)depth, (uint32_t)color);
    if (info->standard == 1 && !(depth == 8 && (color == 2 || color == 6)))
        return fail(info, ORC_ERR_PARSE_HEADER_PIXEL_FORMAT, (uint32_t)depth, (uint32_t)color);
    if (d[10] != 0) return fail(info, ORC_ERR_PARSE_HEADER_COMPRESSION_CODE, d[10], 0);
    if (d[11] != 0) return fail(info, ORC_ERR_PARSE_HEADER_FILTER_CODE, d[11], 0);
    if (d[12] > 1) return fail(info, ORC_ERR_PARSE_HEADER_INTERLACING_CODE, d[12], 0);
    info->width = be32(d), info->height = be32(d + 4);
    if (info->width == 0 || info->height == 0) return fail(info, ORC_ERR_PARSE_HEADER_SIZE, info->width, info->height);
    info->depth = (uint8_t)depth, info->color = (uint8_t)color, info->interlaced = d[12];
    info->format.color = (uint8_t)color, info->format.depth = (uint8_t)depth;
    info->format.bgr = info->standard == 1;
    return ORC_OK;
}

/* the part of decompress(stream:) before the first IDAT (PNG.Image.swift:301-383) + Format.recognize */
static int lex_preamble(lexer* lx, orc_png_info* info, uint32_t* type, const uint8_t** data, uint32_t* len)
{
    memset(info, 0, sizeof *info);
    info->format.palette = info->palette_rgba;
    if (lx->n < 8) return fail(info, ORC_ERR_LEX_TRUNCATED_SIGNATURE, 0, 0);
    if (memcmp(lx->file, SIGNATURE, 8)) return fail(info, ORC_ERR_LEX_INVALID_SIGNATURE, be32(lx->file), be32(lx->file + 4));
    lx->at = 8;
    int st = lex_chunk(lx, info, type, data, len);
    if (st) return st;
    if (*type == T_CgBI) {
        info->standard = 1;
        if ((st = lex_chunk(lx, info, type, data, len))) return st;
    }
    if (*type != T_IHDR) return fail(info, ORC_ERR_DECODE_REQUIRED_CHUNK, T_IHDR, *type);
    if ((st = parse_header(info, *data, *len))) return st;
    int have_palette = 0, have_background = 0, have_transparency = 0;
    uint32_t npal = 0, nalpha = 0;
    uint8_t  alpha[256];
    for (;;) {
        if ((st = lex_chunk(lx, info, type, data, len))) return st;
        const uint8_t* d = *data;
        const uint32_t n = *len;
        if (*type == T_IHDR) return fail(info, ORC_ERR_DECODE_DUPLICATE_CHUNK, T_IHDR, 0);
        if (*type == T_PLTE) {
            if (have_palette) return fail(info, ORC_ERR_DECODE_DUPLICATE_CHUNK, T_PLTE, 0);
            if (have_background) return fail(info, ORC_ERR_DECODE_UNEXPECTED_CHUNK, T_PLTE, T_bKGD);
            if (have_transparency) return fail(info, ORC_ERR_DECODE_UNEXPECTED_CHUNK, T_PLTE, T_tRNS);
            /* PNG.Palette.init(parsing:pixel:) (Parsing/PNG.Palette.swift:27-55) */
            if (info->color == 0 || info->color == 4) return fail(info, ORC_ERR_PARSE_UNEXPECTED_PALETTE, 0, 0);
            if (n % 3) return fail(info, ORC_ERR_PARSE_PALETTE_CHUNK_LENGTH, n, 0);
            const uint32_t max = 1u << (info->depth < 8 ? info->depth : 8);
            if (n / 3 < 1 || n / 3 > max) return fail(info, ORC_ERR_PARSE_PALETTE_COUNT, n / 3, max);
            have_palette = 1, npal = n / 3;
            if (info->color == 3)
                for (uint32_t i = 0; i < npal; ++i) {
                    memcpy(info->palette_rgba + 4 * i, d + 3 * i, 3);
                    info->palette_rgba[4 * i + 3] = 255;
                }
        } else if (*type == T_tRNS) {
            /* Metadata.push(ancillary:) unique-assigns tRNS; PNG.Transparency.init(parsing:pixel:palette:)
             * (Parsing/PNG.Transparency.swift:68-122) */
            if (have_transparency) return fail(info, ORC_ERR_DECODE_DUPLICATE_CHUNK, T_tRNS, 0);
            const uint32_t max = 0xffffu >> (16 - info->depth);
            if (info->color == 0) {
                if (n != 2) return fail(info, ORC_ERR_PARSE_TRANSPARENCY_CHUNK_LENGTH, n, 2);
                if (be16(d) > max) return fail(info, ORC_ERR_PARSE_TRANSPARENCY_SAMPLE, be16(d), max);
                info->format.has_key = 1, info->format.key[0] = (uint16_t)be16(d);
            } else if (info->color == 2) {
                if (n != 6) return fail(info, ORC_ERR_PARSE_TRANSPARENCY_CHUNK_LENGTH, n, 6);
                const uint32_t r = be16(d), g = be16(d + 2), b = be16(d + 4);
                const uint32_t top = r > g ? (r > b ? r : b) : (g > b ? g : b);
                if (top > max) return fail(info, ORC_ERR_PARSE_TRANSPARENCY_SAMPLE, top, max);
                info->format.has_key = 1;  /* Format.recognize stores a bgr8 key as (b, g, r) */
                info->format.key[0] = (uint16_t)(info->standard ? b : r), info->format.key[1] = (uint16_t)g;
                info->format.key[2] = (uint16_t)(info->standard ? r : b);
            } else if (info->color == 3) {
                if (!have_palette) return fail(info, ORC_ERR_DECODE_REQUIRED_CHUNK, T_PLTE, T_tRNS);
                if (n > npal) return fail(info, ORC_ERR_PARSE_TRANSPARENCY_COUNT, n, npal);
                memcpy(alpha, d, n), nalpha = n;
            } else
                return fail(info, ORC_ERR_PARSE_UNEXPECTED_TRANSPARENCY, 0, 0);
            have_transparency = 1;
        } else if (*type == T_bKGD) {
            if (have_background) return fail(info, ORC_ERR_DECODE_DUPLICATE_CHUNK, T_bKGD, 0);
            if (info->color == 3 && !have_palette) return fail(info, ORC_ERR_DECODE_REQUIRED_CHUNK, T_PLTE, T_bKGD);
            have_background = 1;
        } else if (*type == FOURCC('c', 'H', 'R', 'M') || *type == FOURCC('g', 'A', 'M', 'A') || *type == FOURCC('s', 'R', 'G', 'B') ||
                   *type == FOURCC('i', 'C', 'C', 'P') || *type == FOURCC('s', 'B', 'I', 'T')) {
            /* Metadata.push(ancillary:): before-palette chunk ordering (Decoding/PNG.Metadata.swift:77-83) */
            if (have_palette) return fail(info, ORC_ERR_DECODE_UNEXPECTED_CHUNK, *type, T_PLTE);
        } else if (*type == FOURCC('h', 'I', 'S', 'T')) {
            if (!have_palette) return fail(info, ORC_ERR_DECODE_REQUIRED_CHUNK, T_PLTE, *type);
        } else if (*type == T_IDAT) {
            /* PNG.Context.init fails only for an indexed image without a palette */
            if (info->color == 3 && !have_palette) return fail(info, ORC_ERR_DECODE_REQUIRED_CHUNK, T_PLTE, T_IDAT);
            for (uint32_t i = 0; i < nalpha; ++i) info->palette_rgba[4 * i + 3] = alpha[i];
            info->format.palette_count = info->color == 3 ? (uint16_t)npal : 0;
            return ORC_OK;
        } else if (*type == T_IEND) {
            return fail(info, ORC_ERR_DECODE_REQUIRED_CHUNK, T_IDAT, T_IEND);
        }
    }
}

/* after the IDAT run: Context.push(ancillary:) (Decoding/PNG.Context.swift:51-81) */
static int lex_trailer(lexer* lx, orc_png_info* info, uint32_t type, const uint8_t* data, uint32_t len)
{
    for (;;) {
        switch (type) {
        case T_IEND: return ORC_OK;
        case T_CgBI: case T_IHDR: case T_PLTE: case T_bKGD: case T_tRNS: case T_IDAT:
        case FOURCC('h', 'I', 'S', 'T'): case FOURCC('c', 'H', 'R', 'M'): case FOURCC('g', 'A', 'M', 'A'):
        case FOURCC('s', 'R', 'G', 'B'): case FOURCC('i', 'C', 'C', 'P'): case FOURCC('s', 'B', 'I', 'T'):
        case FOURCC('p', 'H', 'Y', 's'): case FOURCC('s', 'P', 'L', 'T'):
            return fail(info, ORC_ERR_DECODE_UNEXPECTED_CHUNK, type, T_IDAT);
        default: break;
        }
        int st = lex_chunk(lx, info, &type, &data, &len);
        if (st) return st;
    }
}

int orc_png_inspect(const uint8_t* file, size_t n, orc_png_info* info)
{
    lexer lx = {file, n, 0};
    uint32_t type, len;
    const uint8_t* data;
    int st = lex_preamble(&lx, info, &type, &data, &len);
    if (st) return st;
    while (type == T_IDAT) {
        info->idat_bytes += len, info->idat_chunks++;
        if ((st = lex_chunk(&lx, info, &type, &data, &len))) return st;
    }
    return lex_trailer(&lx, info, type, data, len);
}

int orc_png_decompress(const uint8_t* file, size_t n, orc_png_info* info, uint8_t* storage, size_t cap)
{
    lexer lx = {file, n, 0};
    uint32_t type, len;
    const uint8_t* data;
    int st = lex_preamble(&lx, info, &type, &data, &len);
    if (st) return st;
    const int    volume = info->depth * channels_of(info->color);
    const size_t need   = (size_t)info->width * info->height * (size_t)((volume + 7) >> 3);
    if (cap < need) return fail(info, ORC_ERR_OUTPUT_CAPACITY, 0, 0);
    /* The reference pushes IDAT chunks into the decoder as it lexes them, so a decoder error in
     * chunk k surfaces before a lexing error in chunk k+1.  Restated by decoding the payload
     * gathered so far whenever lexing stops, and taking the decoder's verdict if it is an error
     * other than "needs more input". */
    size_t   idat_cap = 0, lexed = lx.at;
    uint8_t* idat = NULL;
    int      lex_status = ORC_OK;
    orc_png_info lex_info;
    while (type == T_IDAT) {
        if (info->idat_bytes + len > idat_cap) {
            idat_cap = (size_t)((info->idat_bytes + len) * 2 + 4096);
            idat = (uint8_t*)realloc(idat, idat_cap);
        }
        memcpy(idat + info->idat_bytes, data, len);
        info->idat_bytes += len, info->idat_chunks++;
        lexed = lx.at;
        if ((lex_status = lex_chunk(&lx, info, &type, &data, &len))) break;
    }
    (void)lexed;
    lex_info = *info;
    st = orc_png_decode(info->standard ? ORC_FORMAT_IOS : ORC_FORMAT_ZLIB, idat ? idat : (const uint8_t*)"", (size_t)info->idat_bytes,
                        info->width, info->height, volume, info->depth, info->interlaced, storage, &info->inflate);
    free(idat);
    if (lex_status) {
        /* decoder errors that do not depend on data still to come win over the later lexing error */
        if (st < 0 && st != ORC_ERR_PNG_INCOMPLETE_DATASTREAM) return fail(info, st, info->inflate.a, info->inflate.b);
        return fail(info, lex_status, lex_info.a, lex_info.b);
    }
    if (st < 0 && st != ORC_ERR_PNG_INCOMPLETE_DATASTREAM) return fail(info, st, info->inflate.a, info->inflate.b);
    const int decode_status = st;
    if ((st = lex_trailer(&lx, info, type, data, len))) return st;
    /* IEND with the decoder still expecting data: DecodingError.incompleteImageDataCompressedDatastream */
    if (decode_status) return fail(info, decode_status, info->inflate.a, info->inflate.b);
    return ORC_OK;
}

static size_t put_chunk(uint8_t* out, size_t at, uint32_t type, const uint8_t* data, size_t n)
{
    /* BytestreamDestination.format(type:data:) (PNG.BytestreamDestination.swift:66-95) */
    put32(out + at, (uint32_t)n), put32(out + at + 4, type);
    if (n) memcpy(out + at + 8, data, n);
    put32(out + at + 8 + n, orc_crc32(0, out + at + 4, n + 4));
    return at + 12 + n;
}

size_t orc_png_compress_bound(uint32_t w, uint32_t h, const orc_format* f, int interlaced, size_t idat_chunk)
{
    const int    volume   = f->depth * channels_of(f->color);
    const size_t filtered = orc_png_filtered_size(w, h, volume, interlaced);
    const size_t z        = orc_deflate_bound(filtered);
    return 8 + 16 + 25 + (12 + 768) + (12 + 256) + z + 12 * (z / (idat_chunk ? idat_chunk : 1) + 2) + 12;
}

size_t orc_png_compress(const uint8_t* storage, uint32_t w, uint32_t h, const orc_format* f,
                        int interlaced, int level, size_t idat_chunk, uint8_t* out, size_t cap)
{
    if (cap < orc_png_compress_bound(w, h, f, interlaced, idat_chunk)) return (size_t)-1;
    const int volume = f->depth * channels_of(f->color);
    size_t    at = 8;
    memcpy(out, SIGNATURE, 8);
    if (f->bgr) {  /* PNG.Image.encode (PNG.Image.swift:416-423) */
        const uint8_t cgbi[4] = {48, 0, 32, (uint8_t)(f->color == 2 ? 6 : 2)};
        at = put_chunk(out, at, T_CgBI, cgbi, 4);
    }
    uint8_t hdr[13];  /* PNG.Header.serialized (PNG.Header.swift:100-113) */
    put32(hdr, w), put32(hdr + 4, h);
    hdr[8] = f->depth, hdr[9] = f->color, hdr[10] = 0, hdr[11] = 0, hdr[12] = interlaced ? 1 : 0;
    at = put_chunk(out, at, T_IHDR, hdr, 13);
    if (f->color == 3) {  /* Layout.palette / Layout.transparency (Formats/PNG.Layout.swift:43-135) */
        uint8_t rgb[768], alpha[256];
        int     last = -1;
        for (int i = 0; i < f->palette_count; ++i) {
            memcpy(rgb + 3 * i, f->palette + 4 * i, 3);
            alpha[i] = f->palette[4 * i + 3];
            if (alpha[i] != 255) last = i;
        }
        at = put_chunk(out, at, T_PLTE, rgb, 3 * (size_t)f->palette_count);
        if (last >= 0) at = put_chunk(out, at, T_tRNS, alpha, (size_t)last + 1);
    } else if (f->has_key) {
        uint8_t k[6];
        if (f->color == 0) {
            k[0] = (uint8_t)(f->key[0] >> 8), k[1] = (uint8_t)f->key[0];
            at = put_chunk(out, at, T_tRNS, k, 2);
        } else {
            const uint16_t r = f->bgr ? f->key[2] : f->key[0], g = f->key[1], b = f->bgr ? f->key[0] : f->key[2];
            k[0] = (uint8_t)(r >> 8), k[1] = (uint8_t)r, k[2] = (uint8_t)(g >> 8), k[3] = (uint8_t)g;
            k[4] = (uint8_t)(b >> 8), k[5] = (uint8_t)b;
            at = put_chunk(out, at, T_tRNS, k, 6);
        }
    }
    const size_t fsz = orc_png_filtered_size(w, h, volume, interlaced);
    uint8_t*     filtered = (uint8_t*)malloc(fsz ? fsz : 1);
    orc_png_filter(storage, w, h, volume, f->depth, interlaced, filtered, fsz);
    const size_t zcap = orc_deflate_bound(fsz);
    uint8_t*     z = (uint8_t*)malloc(zcap);
    const size_t zn = orc_deflate(f->bgr ? ORC_FORMAT_IOS : ORC_FORMAT_ZLIB, level, 15, filtered, fsz, z, zcap);
    /* Encoder.pull hands out DeflatorOut's queued buffers of 2 x capacity bytes, then the rest
     * (Encoding/PNG.Encoder.swift:33-129, Deflator/LZ77.DeflatorOut.swift:73-125) */
    for (size_t o = 0; o < zn; o += idat_chunk)
        at = put_chunk(out, at, T_IDAT, z + o, zn - o < idat_chunk ? zn - o : idat_chunk);
    free(filtered), free(z);
    return put_chunk(out, at, T_IEND, NULL, 0);
}
