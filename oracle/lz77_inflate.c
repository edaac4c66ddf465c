/*
 * oracle/lz77_inflate.c -- CPU restatement of LZ77.Inflator / Gzip.Inflator.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Parity: pinned (tests/test_oracle_decode.py).
 *
 * Follows, function by function (paths relative to the reference checkout):
 *   Sources/LZ77/Inflator/LZ77.InflatorIn.swift:156-198        bit addressing (in_bits, in_peek16)
 *   Sources/LZ77/Inflator/LZ77.StreamHeader.swift:16-54        zlib header            (read_zlib_header)
 *   Sources/LZ77/Gzip/Gzip.StreamHeader.swift:19-83            gzip header            (read_gzip_header)
 *   Sources/LZ77/Inflator/LZ77.InflatorBuffers.swift:25-230    state machine          (orc_inflate)
 *   Sources/LZ77/Inflator/LZ77.InflatorBuffers.Stream.swift:59-141   readBlockMetadata
 *   ...Stream.swift:144-263                                    readBlockTables
 *   ...Stream.swift:266-381                                    readBlock(with:)  token loop
 *   ...Stream.swift:384-399                                    readBlock(upTo:)  stored blocks
 *   Sources/LZ77/HuffmanCoding/LZ77.HuffmanTree.swift:49-202   validate / size / table
 *   Sources/LZ77/Inflator/LZ77.InflatorTables.swift:103-119    two-level lookup (fence)
 *   Sources/LZ77/LZ77.Composites.swift:19-111                  length/distance base+extra
 *   Sources/LZ77/Inflator/LZ77.InflatorOut.swift:114-140       append / expand (forward byte copy)
 *   Sources/LZ77/Wrappers/LZ77.MRC32.swift:26-47               Adler-32
 *
 * Deliberate differences from the reference (all on INVALID streams only):
 *   - length symbols 286/287, distance symbols 30/31 and unassigned stub-tree codes return
 *     ORC_ERR_INVALID_SYMBOL instead of copying 0 bytes / reading uninitialised memory;
 *   - the stored-block rebase bug is not reproduced (RFC-correct behaviour);
 *   - one-shot: startIndex is 0, so invalidStringReference fires iff distance > bytes produced.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---- LZ77.Reversed (bit reversal of a byte), computed rather than tabulated ---- */
static uint8_t REV8[256];
static int     rev8_ready = 0;
static void rev8_init(void)
{
    if (rev8_ready) return;
    for (int i = 0; i < 256; ++i) {
        int r = 0;
        for (int k = 0; k < 8; ++k)
            if (i & (1 << k)) r |= 0x80 >> k;
        REV8[i] = (uint8_t)r;
    }
    rev8_ready = 1;
}

/* ---- LZ77.Composites: RFC 1951 base/extra tables, index 0 = front padding ---- */
static const uint16_t RUN_EXTRA[32] = {0, 0,0,0,0,0, 0,0,0,1,1, 1,1,2,2,2, 2,3,3,3,3,
                                       4,4,4,4,5, 5,5,5,0, 0,0};
static const uint16_t RUN_BASE[32]  = {0, 3,4,5,6,7, 8,9,10,11,13, 15,17,19,23,27, 31,35,43,51,59,
                                       67,83,99,115,131, 163,195,227,258, 0,0};
static const uint16_t DIST_EXTRA[32] = {0,0,0,0,1, 1,2,2,3,3, 4,4,5,5,6, 6,7,7,8,8,
                                        9,9,10,10,11, 11,12,12,13,13, 0,0};
static const uint16_t DIST_BASE[32]  = {1,2,3,4,5, 7,9,13,17,25, 33,49,65,97,129,
                                        193,257,385,513,769, 1025,1537,2049,3073,4097,
                                        6145,8193,12289,16385,24577, 0,0};

/* ---- LZ77.InflatorIn: LSB-first bit addressing, zero padding past the end ---- */
typedef struct { const uint8_t* p; size_t bytes; } bitin;

static inline size_t   in_count(const bitin* s) { return s->bytes << 3; }
static inline uint32_t in_byte(const bitin* s, size_t i) { return i < s->bytes ? s->p[i] : 0u; }
/* InflatorIn[i] (16 bits at bit i), InflatorIn.swift:181-198 */
static inline uint32_t in_peek16(const bitin* s, size_t b)
{
    size_t   a = b >> 3;
    uint32_t v = in_byte(s, a) | in_byte(s, a + 1) << 8 | in_byte(s, a + 2) << 16;
    return (v >> (b & 7)) & 0xffffu;
}
/* InflatorIn[i, count:, as:], InflatorIn.swift:156-179 (count <= 16) */
static inline uint32_t in_bits(const bitin* s, size_t b, int count)
{
    if (count <= 0) return 0;
    return in_peek16(s, b) & ~(0xffffffffu << count);
}

/* ---- LZ77.HuffmanTree (decoder half) ---- */
typedef struct {
    uint16_t symbols[320];
    int      lo[15], hi[15]; /* levels[l-1] = lo ..< hi */
    int      n, z;           /* size.n (fence), size.z (table entries) */
} htree;

/* HuffmanTree.size(_:), HuffmanTree.swift:80-108 */
static int tree_size(htree* t)
{
    long interior = 1;
    for (int i = 0; i < 8; ++i) interior = 2 * interior - (t->hi[i] - t->lo[i]);
    long n = 256 - interior, z = 256;
    for (int i = 0; i < 7; ++i) {
        int leaves = t->hi[8 + i] - t->lo[8 + i];
        z += (long)leaves << (6 - i);
        interior = 2 * interior - leaves;
    }
    if (interior != 0) return 0;
    t->n = (int)n;
    t->z = (int)z;
    return 1;
}

/* HuffmanTree.validate(symbols:lengths:), HuffmanTree.swift:137-174.  symbols are 0,1,2,... */
static int tree_validate(htree* t, const int* lengths, int count)
{
    int counts[16] = {0};
    for (int i = 0; i < count; ++i) counts[lengths[i]]++;
    int base = 0;
    for (int l = 1; l <= 15; ++l) {
        t->lo[l - 1] = base;
        t->hi[l - 1] = base + counts[l];
        base += counts[l];
    }
    if (!tree_size(t)) return 0;
    for (int s = 0; s < count; ++s) {
        int l = lengths[s];
        if (l > 0) {
            t->symbols[t->hi[l - 1] - counts[l]] = (uint16_t)s;
            counts[l]--;
        }
    }
    return 1;
}

/* HuffmanTree.init(stub:), HuffmanTree.swift:52-65 */
static void tree_stub(htree* t, int stub)
{
    int k = stub >= 0 ? 1 : 0;
    if (k) t->symbols[0] = (uint16_t)stub;
    t->lo[0] = 0;
    t->hi[0] = k;
    for (int i = 1; i < 15; ++i) t->lo[i] = t->hi[i] = k;
    t->n = 256;
    t->z = 256;
}

/* HuffmanTree.validate(symbols:normalizing:), HuffmanTree.swift:112-135 */
static int tree_validate_normalizing(htree* t, const int* lengths, int count)
{
    int first = -1;
    for (int s = 0; s < count; ++s) {
        if (lengths[s] <= 0) continue;
        if (first < 0 && lengths[s] == 1) first = s;
        else return tree_validate(t, lengths, count);
    }
    tree_stub(t, first);
    return 1;
}

/* decode-table entry: length << 16 | symbol; 0 = never initialised by table() */
typedef uint32_t hentry;
#define H_SYM(e) ((e) & 0xffffu)
#define H_LEN(e) ((int)((e) >> 16))

/* HuffmanTree.table(initializing:), HuffmanTree.swift:176-201 */
static void tree_table(const htree* t, hentry* dst)
{
    memset(dst, 0, sizeof(hentry) * (size_t)t->z);
    hentry* cur = dst;
    for (int l = 1; l <= 8; ++l) {
        int clones = 256 >> l;
        for (int i = t->lo[l - 1]; i < t->hi[l - 1]; ++i) {
            hentry e = (hentry)l << 16 | t->symbols[i];
            for (int c = 0; c < clones; ++c) *cur++ = e;
        }
    }
    cur = dst + 256;
    for (int l = 9; l <= 15; ++l) {
        int clones = 32768 >> l;
        for (int i = t->lo[l - 1]; i < t->hi[l - 1]; ++i) {
            hentry e = (hentry)l << 16 | t->symbols[i];
            for (int c = 0; c < clones; ++c) *cur++ = e;
        }
    }
}

/* InflatorTables.index(_:fence:), InflatorTables.swift:113-119 */
static inline int table_index(uint32_t codeword, int fence)
{
    int first = REV8[codeword & 0xff];
    return first < fence ? first : (((first - fence + 2) << 8) | REV8[(codeword >> 8) & 0xff]) >> 1;
}

/* LZ77.InflatorTables: run-literal + distance decode tables */
typedef struct {
    hentry runliteral[256 + 128 * 144];
    hentry distance[256 + 128 * 16];
    int    fence_rl, fence_d;
} itables;

static void tables_init(itables* tb, const htree* lit, const htree* dist)
{
    tree_table(lit, tb->runliteral);
    tree_table(dist, tb->distance);
    tb->fence_rl = lit->n;
    tb->fence_d  = dist->n;
}

/* HuffmanTree.runliteral / .distance (fixed trees), HuffmanTree.swift:23-47 */
static void tables_fixed(itables* tb)
{
    htree lit, dist;
    int   k = 0;
    for (int s = 256; s <= 279; ++s) lit.symbols[k++] = (uint16_t)s;
    for (int s = 0; s <= 143; ++s) lit.symbols[k++] = (uint16_t)s;
    for (int s = 280; s <= 287; ++s) lit.symbols[k++] = (uint16_t)s;
    for (int s = 144; s <= 255; ++s) lit.symbols[k++] = (uint16_t)s;
    for (int i = 0; i < 15; ++i) lit.lo[i] = lit.hi[i] = (i < 6 ? 0 : 288);
    lit.lo[6] = 0;   lit.hi[6] = 24;
    lit.lo[7] = 24;  lit.hi[7] = 176;
    lit.lo[8] = 176; lit.hi[8] = 288;
    tree_size(&lit);
    for (int s = 0; s < 32; ++s) dist.symbols[s] = (uint16_t)s;
    for (int i = 0; i < 15; ++i) dist.lo[i] = dist.hi[i] = (i < 4 ? 0 : 32);
    dist.lo[4] = 0;  dist.hi[4] = 32;
    tree_size(&dist);
    tables_init(tb, &lit, &dist);
}

/* ---- checksums ---- */
/* LZ77.MRC32.update, MRC32.swift:26-47 */
uint32_t orc_adler32(uint32_t adler, const uint8_t* p, size_t n)
{
    uint32_t single = adler & 0xffff, dbl = adler >> 16;
    while (n) {
        size_t k = n < 5552 ? n : 5552;
        for (size_t j = 0; j < k; ++j) {
            single += p[j];
            dbl += single;
        }
        single %= 65521;
        dbl %= 65521;
        p += k;
        n -= k;
    }
    return dbl << 16 | single;
}

/* CRC-32 as used by swift-hash 0.7.1 `CRC32` (reflected 0xEDB88320, init/xorout ~0);
 * the dependency is not vendored; pinned by the reference's KATs
 * (Sources/PNGIntegrationTests/ErrorHandling.swift:30,42) and the gzip fixtures' trailers. */
uint32_t orc_crc32(uint32_t crc, const uint8_t* p, size_t n)
{
    static uint32_t table[256];
    static int      ready = 0;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = c & 1 ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = 1;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}

/* ---- LZ77.InflatorOut (one-shot: nothing is ever released while decoding) ---- */
typedef struct {
    uint8_t* p;
    size_t   end, cap;
    int      owned, overflow;
} outbuf;

static int out_reserve(outbuf* o, size_t count)
{
    if (o->end + count <= o->cap) return 1;
    if (!o->owned) { o->overflow = 1; return 0; }
    size_t cap = o->cap ? o->cap : 1 << 16;
    while (cap < o->end + count) cap <<= 1;
    uint8_t* q = (uint8_t*)realloc(o->p, cap);
    if (!q) { o->overflow = 1; return 0; }
    o->p = q;
    o->cap = cap;
    return 1;
}

/* ---- stream state ---- */
typedef struct {
    bitin  in;
    size_t b;
    outbuf out;
} stream;

static int fail(orc_inflate_result* r, int code, uint32_t a, uint32_t b)
{
    r->status = code;
    r->a = a;
    r->b = b;
    return code;
}

/* LZ77.StreamHeader.read, StreamHeader.swift:16-54.  returns 0 ok, 1 need more, <0 error */
static int read_zlib_header(stream* s, orc_inflate_result* r)
{
    if (s->b + 16 > in_count(&s->in)) return ORC_NEED_MORE_INPUT;
    uint32_t method = in_bits(&s->in, s->b, 4);
    if (method != 8) return fail(r, ORC_ERR_ZLIB_METHOD, method, 0);
    uint32_t e = in_bits(&s->in, s->b + 4, 4);
    if (e >= 8) return fail(r, ORC_ERR_ZLIB_WINDOW, e + 8, 0);
    uint32_t flags = in_bits(&s->in, s->b + 8, 8);
    /* Swift precedence: `e << 12 | 8 << 8 + flags` parses as (e << 12) | (8 << (8 + flags))?  No:
     * in Swift `<<` binds tighter than `+` (BitwiseShiftPrecedence > AdditionPrecedence) and
     * `|` has AdditionPrecedence, left-assoc: ((e << 12) | (8 << 8)) + flags. */
    if ((((e << 12) | (8u << 8)) + flags) % 31 != 0) return fail(r, ORC_ERR_ZLIB_CHECK_BITS, 0, 0);
    if (flags & 0x20) return fail(r, ORC_ERR_ZLIB_DICTIONARY, 0, 0);
    s->b += 16;
    return ORC_OK;
}

/* readByte, Stream.swift:456-469 */
static int read_byte(stream* s, uint32_t* byte)
{
    if (s->b + 8 > in_count(&s->in)) return 0;
    *byte = in_bits(&s->in, s->b, 8);
    s->b += 8;
    return 1;
}

/* Gzip.StreamHeader.read + the .strings state, Gzip.StreamHeader.swift:19-83,
 * InflatorBuffers.swift:153-197 */
static int read_gzip_header(stream* s, orc_inflate_result* r)
{
    if (s->b + 80 > in_count(&s->in)) return ORC_NEED_MORE_INPUT;
    if (in_peek16(&s->in, s->b) != 0x8b1f) return fail(r, ORC_ERR_GZIP_SIGIL, 0, 0);
    uint32_t method = in_bits(&s->in, s->b + 16, 8);
    if (method != 8) return fail(r, ORC_ERR_GZIP_METHOD, method, 0);
    uint32_t flags = in_bits(&s->in, s->b + 24, 8);
    if (flags & 0xe0) return fail(r, ORC_ERR_GZIP_FLAG_BITS, flags, 0);
    if (flags & 0x02) return fail(r, ORC_ERR_GZIP_HEADER_CHECKSUM_UNSUPPORTED, 0, 0);
    size_t xlen = 0;
    if (flags & 0x04) {
        if (s->b + 96 > in_count(&s->in)) return ORC_NEED_MORE_INPUT;
        xlen = in_peek16(&s->in, s->b + 80);
        s->b += 96;
    } else {
        s->b += 80;
    }
    int count = ((flags & 0x08) ? 1 : 0) + ((flags & 0x10) ? 1 : 0);
    if (xlen) {
        if (s->b + 8 * xlen > in_count(&s->in)) return ORC_NEED_MORE_INPUT;
        s->b += 8 * xlen;
    }
    while (count > 0) { /* readString, Stream.swift:441-453 */
        uint32_t byte;
        do {
            if (!read_byte(s, &byte)) return ORC_NEED_MORE_INPUT;
        } while (byte != 0);
        count--;
    }
    return ORC_OK;
}

/* readBigEndianUInt32, Stream.swift:401-431 */
static int read_be32_aligned(stream* s, uint32_t* v)
{
    size_t boundary = (s->b + 7) & ~(size_t)7;
    if (boundary + 32 > in_count(&s->in)) return 0;
    s->b = boundary + 32;
    *v = in_bits(&s->in, boundary, 8) << 24 | in_bits(&s->in, boundary + 8, 8) << 16 |
         in_bits(&s->in, boundary + 16, 8) << 8 | in_bits(&s->in, boundary + 24, 8);
    return 1;
}

static const int CODELENGTH_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

/* test hook: where each block header starts (bit offset, output offset, block type) */
static __thread uint64_t* g_trace = NULL;
static __thread size_t    g_trace_cap = 0, g_trace_n = 0;

/* one DEFLATE block.  returns ORC_OK (block done), ORC_NEED_MORE_INPUT, or error; *final set */
static int read_block(stream* s, itables* tb, int* final, orc_inflate_result* r)
{
    const bitin* in = &s->in;
    if (g_trace && s->b + 3 <= in_count(in)) {
        if (g_trace_n < g_trace_cap) {
            g_trace[3 * g_trace_n] = s->b, g_trace[3 * g_trace_n + 1] = s->out.end;
            g_trace[3 * g_trace_n + 2] = in_bits(in, s->b + 1, 2);
        }
        g_trace_n++;
    }
    /* ---- readBlockMetadata, Stream.swift:59-141 ---- */
    if (s->b + 3 > in_count(in)) return ORC_NEED_MORE_INPUT;
    *final = in_bits(in, s->b, 1) != 0;
    uint32_t type = in_bits(in, s->b + 1, 2);
    if (type == 0) {
        size_t boundary = (s->b + 3 + 7) & ~(size_t)7;
        if (boundary + 32 > in_count(in)) return ORC_NEED_MORE_INPUT;
        uint32_t l = in_bits(in, boundary, 16), m = in_bits(in, boundary + 16, 16);
        if (l != (~m & 0xffffu)) return fail(r, ORC_ERR_BLOCK_COUNT_PARITY, l, m);
        s->b = boundary + 32;
        /* readBlock(upTo:), Stream.swift:384-399 */
        for (uint32_t i = 0; i < l; ++i) {
            uint32_t byte;
            if (!read_byte(s, &byte)) return ORC_NEED_MORE_INPUT;
            if (!out_reserve(&s->out, 1)) return fail(r, ORC_ERR_OUTPUT_CAPACITY, 0, 0);
            s->out.p[s->out.end++] = (uint8_t)byte;
        }
        return ORC_OK;
    } else if (type == 1) {
        s->b += 3;
        tables_fixed(tb);
    } else if (type == 2) {
        if (s->b + 17 > in_count(in)) return ORC_NEED_MORE_INPUT;
        int codelengths = 4 + (int)in_bits(in, s->b + 13, 4);
        if (s->b + 17 + 3 * (size_t)codelengths > in_count(in)) return ORC_NEED_MORE_INPUT;
        int literals  = 257 + (int)in_bits(in, s->b + 3, 5);
        int distances = 1 + (int)in_bits(in, s->b + 8, 5);
        if (literals < 257 || literals > 286)
            return fail(r, ORC_ERR_RUNLITERAL_SYMBOL_COUNT, (uint32_t)literals, 0);
        int metalengths[19] = {0};
        for (int i = 0; i < codelengths; ++i)
            metalengths[CODELENGTH_ORDER[i]] = (int)in_bits(in, s->b + 17 + 3 * (size_t)i, 3);
        htree meta;
        if (!tree_validate(&meta, metalengths, 19))
            return fail(r, ORC_ERR_CODELENGTH_HUFFMAN_TABLE, 0, 0);
        hentry metatable[256]; /* BlockMetadata.replace(tree:), BlockMetadata.swift:49-61 */
        tree_table(&meta, metatable);
        s->b += 17 + 3 * (size_t)codelengths;

        /* ---- readBlockTables, Stream.swift:144-263 ---- */
        int total = literals + distances, have = 0;
        int lengths[320 + 138];
        while (have < total) {
            if (s->b >= in_count(in)) return ORC_NEED_MORE_INPUT;
            hentry mw = metatable[REV8[in_peek16(in, s->b) & 0xff]];
            if (s->b + (size_t)H_LEN(mw) > in_count(in)) return ORC_NEED_MORE_INPUT;
            int sym = (int)H_SYM(mw), element, extra, base;
            if (sym < 16) {
                lengths[have++] = sym;
                s->b += (size_t)H_LEN(mw);
                continue;
            } else if (sym == 16) {
                if (have == 0) return fail(r, ORC_ERR_CODELENGTH_SEQUENCE, 0, 0);
                element = lengths[have - 1]; extra = 2; base = 3;
            } else if (sym == 17) {
                element = 0; extra = 3; base = 3;
            } else {
                element = 0; extra = 7; base = 11;
            }
            if (s->b + (size_t)H_LEN(mw) + (size_t)extra > in_count(in)) return ORC_NEED_MORE_INPUT;
            int reps = base + (int)in_bits(in, s->b + (size_t)H_LEN(mw), extra);
            for (int k = 0; k < reps; ++k) lengths[have++] = element;
            s->b += (size_t)H_LEN(mw) + (size_t)extra;
        }
        if (have != total) return fail(r, ORC_ERR_CODELENGTH_SEQUENCE, 0, 0);
        htree lit, dist;
        if (!tree_validate(&lit, lengths, literals) ||
            !tree_validate_normalizing(&dist, lengths + literals, distances))
            return fail(r, ORC_ERR_HUFFMAN_TABLE, 0, 0);
        tables_init(tb, &lit, &dist);
    } else {
        return fail(r, ORC_ERR_BLOCK_TYPE, type, 0);
    }

    /* ---- readBlock(with:), Stream.swift:266-381 ---- */
    while (s->b < in_count(in)) {
        uint32_t first = in_peek16(in, s->b);
        hentry   rl    = tb->runliteral[table_index(first, tb->fence_rl)];
        int      rlen  = H_LEN(rl);
        uint32_t sym   = H_SYM(rl);
        if (rlen == 0) return fail(r, ORC_ERR_INVALID_SYMBOL, 0, 0);
        if (sym < 256) {
            if (s->b + (size_t)rlen > in_count(in)) return ORC_NEED_MORE_INPUT;
            s->b += (size_t)rlen;
            if (!out_reserve(&s->out, 1)) return fail(r, ORC_ERR_OUTPUT_CAPACITY, 0, 0);
            s->out.p[s->out.end++] = (uint8_t)sym; /* InflatorOut.append */
        } else if (sym == 256) {
            if (s->b + (size_t)rlen > in_count(in)) return ORC_NEED_MORE_INPUT;
            s->b += (size_t)rlen;
            return ORC_OK;
        } else {
            uint64_t slug = (uint64_t)in_peek16(in, s->b + 32) << 32 |
                            (uint64_t)in_peek16(in, s->b + 16) << 16 | first;
            slug >>= rlen;
            int decade = (int)(sym & 0xff); /* RunLiteral.decade: 257 -> 1 */
            if (decade > 29) return fail(r, ORC_ERR_INVALID_SYMBOL, sym, 0);
            int    cextra = RUN_EXTRA[decade];
            size_t count  = RUN_BASE[decade] + (size_t)(slug & ~(~(uint64_t)0 << cextra));
            slug >>= cextra;
            hentry de   = tb->distance[table_index((uint32_t)(slug & 0xffff), tb->fence_d)];
            int    dlen = H_LEN(de);
            if (dlen == 0) return fail(r, ORC_ERR_INVALID_SYMBOL, 0, 1);
            slug >>= dlen;
            int ddec = (int)H_SYM(de);
            if (ddec > 29) return fail(r, ORC_ERR_INVALID_SYMBOL, (uint32_t)ddec, 1);
            int    oextra = DIST_EXTRA[ddec];
            size_t offset = DIST_BASE[ddec] + (size_t)(slug & ~(~(uint64_t)0 << oextra));
            size_t b      = s->b + (size_t)(rlen + cextra + dlen + oextra);
            if (b > in_count(in)) return ORC_NEED_MORE_INPUT;
            if (offset > s->out.end) return fail(r, ORC_ERR_STRING_REFERENCE, 0, 0);
            if (!out_reserve(&s->out, count)) return fail(r, ORC_ERR_OUTPUT_CAPACITY, 0, 0);
            /* InflatorOut.expand: forward byte-by-byte copy (overlap = RLE), InflatorOut.swift:124-140 */
            uint8_t* dst = s->out.p + s->out.end;
            for (size_t k = 0; k < count; ++k) dst[k] = dst[k - offset];
            s->out.end += count;
            s->b = b;
        }
    }
    return ORC_NEED_MORE_INPUT;
}

void orc_inflate(int format, const uint8_t* in, size_t n, uint8_t* out, size_t cap,
                 orc_inflate_result* res)
{
    rev8_init();
    memset(res, 0, sizeof *res);
    stream s;
    s.in.p = in;
    s.in.bytes = n;
    s.b = 0;
    s.out.p = out;
    s.out.end = 0;
    s.out.cap = out ? cap : 0;
    s.out.owned = out == NULL;
    s.out.overflow = 0;
    itables* tb = (itables*)malloc(sizeof(itables));
    int      st = ORC_OK;

    /* .initial, InflatorBuffers.swift:91-107 / :153-178 */
    if (format == ORC_FORMAT_ZLIB) st = read_zlib_header(&s, res);
    else if (format == ORC_FORMAT_GZIP) st = read_gzip_header(&s, res);
    else if (format != ORC_FORMAT_IOS) st = fail(res, ORC_ERR_BAD_ARGUMENT, 0, 0);

    /* .block(...) */
    while (st == ORC_OK) {
        int final = 0;
        st = read_block(&s, tb, &final, res);
        if (st == ORC_OK) res->blocks++;
        if (st != ORC_OK || final) break;
    }

    /* .checksum, InflatorBuffers.swift:109-130 / :206-223 */
    uint32_t computed = format == ORC_FORMAT_GZIP ? orc_crc32(0, s.out.p, s.out.end)
                                                  : orc_adler32(1, s.out.p, s.out.end);
    if (st == ORC_OK && format != ORC_FORMAT_IOS) {
        uint32_t declared;
        if (!read_be32_aligned(&s, &declared)) st = ORC_NEED_MORE_INPUT;
        else {
            if (format == ORC_FORMAT_GZIP) /* readLittleEndianUInt32 */
                declared = declared >> 24 | (declared >> 8 & 0xff00) | (declared << 8 & 0xff0000) |
                           declared << 24;
            if (declared != computed) st = fail(res, ORC_ERR_STREAM_CHECKSUM, declared, computed);
            else if (format == ORC_FORMAT_GZIP) { /* .epilogue: ISIZE read, not validated */
                uint32_t isize;
                if (!read_be32_aligned(&s, &isize)) st = ORC_NEED_MORE_INPUT;
            }
        }
    }
    if (res->status == 0) res->status = st;
    res->consumed_bits = s.b;
    res->produced = s.out.end;
    res->checksum = computed;
    if (s.out.owned) free(s.out.p);
    free(tb);
}

/* test hook for tools/block_probe.py: inflate while recording (bit offset, output offset, BTYPE) of
 * every block header; returns the number of blocks (entries beyond cap are counted, not stored) */
size_t orc_debug_block_starts(int format, const uint8_t* in, size_t n, uint64_t* trace, size_t cap)
{
    orc_inflate_result res;
    g_trace = trace, g_trace_cap = cap, g_trace_n = 0;
    orc_inflate(format, in, n, NULL, 0, &res);
    g_trace = NULL;
    return res.status == ORC_OK ? g_trace_n : 0;
}
