/*
 * oracle/png_scanline.c -- CPU restatement of swift-png's scanline codec (layer L1).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Parity: pinned (tests/test_oracle_decode.py,
 * tests/test_oracle_encode.py).
 *
 * Follows (paths relative to the reference checkout):
 *   Sources/PNG/PNG.swift:124-147                     PNG.paeth                 (orc_paeth)
 *   Sources/PNG/Decoding/PNG.Decoder.swift:152-196    PNG.Decoder.defilter      (orc_defilter)
 *   Sources/PNG/Decoding/PNG.Decoder.swift:6-15,47-149 Adam7 geometry + row loop (orc_png_unfilter)
 *   Sources/PNG/PNG.Image.swift:186-285               PNG.Image.assign          (assign_row)
 *   Sources/PNG/PNG.Image.swift:431-544               PNG.Image.collect         (collect_row)
 *   Sources/PNG/Encoding/PNG.Encoder.swift:132-204,230-234  Encoder.filter + score (orc_filter_row)
 *   Sources/PNG/Encoding/PNG.Encoder.swift:33-129     row loop                  (orc_png_filter)
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

/* PNG.paeth, PNG.swift:124-147 (tie order a, b, c) */
uint8_t orc_paeth(uint8_t a, uint8_t b, uint8_t c)
{
    int d0 = (int)b - (int)c, d1 = (int)a - (int)c;
    int f0 = d0 < 0 ? -d0 : d0, f1 = d1 < 0 ? -d1 : d1, f2 = d0 + d1 < 0 ? -(d0 + d1) : d0 + d1;
    if (f0 <= f1 && f0 <= f2) return a;
    return f1 <= f2 ? b : c;
}

/* PNG.Decoder.defilter, PNG.Decoder.swift:152-196.  count = pitch + 1 */
void orc_defilter(uint8_t* line, const uint8_t* last, size_t count, int delay)
{
    size_t d = (size_t)delay;
    switch (line[0]) {
    case 1:
        for (size_t i = 1 + d; i < count; ++i) line[i] = (uint8_t)(line[i] + line[i - d]);
        break;
    case 2:
        for (size_t i = 1; i < count; ++i) line[i] = (uint8_t)(line[i] + last[i]);
        break;
    case 3:
        for (size_t i = 1; i < count && i < 1 + d; ++i) line[i] = (uint8_t)(line[i] + (last[i] >> 1));
        for (size_t i = 1 + d; i < count; ++i)
            line[i] = (uint8_t)(line[i] + (((unsigned)line[i - d] + (unsigned)last[i]) >> 1));
        break;
    case 4:
        for (size_t i = 1; i < count && i < 1 + d; ++i) line[i] = (uint8_t)(line[i] + orc_paeth(0, last[i], 0));
        for (size_t i = 1 + d; i < count; ++i)
            line[i] = (uint8_t)(line[i] + orc_paeth(line[i - d], last[i], last[i - d]));
        break;
    default: /* 0, and any invalid filter byte: unchanged (PNG.Decoder.swift:193-194) */
        break;
    }
}

/* PNG.adam7, PNG.Decoder.swift:6-15 */
static const int ADAM7[7][4] = {/* base.x, base.y, exponent.x, exponent.y */
                                {0, 0, 3, 3}, {4, 0, 3, 3}, {0, 4, 2, 3}, {2, 0, 2, 2},
                                {0, 2, 1, 2}, {1, 0, 1, 1}, {0, 1, 0, 1}};

/* PNG.Image.assign, PNG.Image.swift:186-285.  scanline excludes the filter byte */
static void assign_row(uint8_t* storage, uint32_t w, const uint8_t* scanline, int bx, int by,
                       int stride, int volume, int depth)
{
    size_t i = 0;
    if (depth < 8) { /* .v1/.v2/.v4/.indexed1/2/4: one storage byte per pixel, MSB-first */
        int per = 8 / depth, mask = (1 << depth) - 1;
        for (uint32_t x = (uint32_t)bx; x < w; x += (uint32_t)stride, ++i) {
            size_t a = i / (size_t)per;
            int    b = (int)((~i) & (size_t)(per - 1)) * depth;
            storage[(size_t)by * w + x] = (uint8_t)((scanline[a] >> b) & mask);
        }
    } else {
        size_t bpp = (size_t)volume >> 3;
        for (uint32_t x = (uint32_t)bx; x < w; x += (uint32_t)stride, ++i)
            memcpy(storage + bpp * ((size_t)by * w + x), scanline + bpp * i, bpp);
    }
}

/* PNG.Image.collect, PNG.Image.swift:431-544 */
static void collect_row(const uint8_t* storage, uint32_t w, uint8_t* scanline, size_t pitch, int bx,
                        int by, int stride, int volume, int depth)
{
    size_t i = 0;
    if (depth < 8) {
        int per = 8 / depth, mask = (1 << depth) - 1;
        memset(scanline, 0, pitch);
        for (uint32_t x = (uint32_t)bx; x < w; x += (uint32_t)stride, ++i) {
            size_t a = i / (size_t)per;
            int    b = (int)((~i) & (size_t)(per - 1)) * depth;
            scanline[a] |= (uint8_t)((storage[(size_t)by * w + x] & mask) << b);
        }
    } else {
        size_t bpp = (size_t)volume >> 3;
        for (uint32_t x = (uint32_t)bx; x < w; x += (uint32_t)stride, ++i)
            memcpy(scanline + bpp * i, storage + bpp * ((size_t)by * w + x), bpp);
    }
}

typedef struct { int bx, by, sx, sy; uint32_t w, h; size_t pitch; } pass_t;

static int passes(uint32_t w, uint32_t h, int volume, int interlaced, pass_t out[7])
{
    if (!interlaced) {
        out[0] = (pass_t){0, 0, 1, 1, w, h, ((size_t)w * (size_t)volume + 7) >> 3};
        return 1;
    }
    int n = 0;
    for (int z = 0; z < 7; ++z) {
        int      sx = 1 << ADAM7[z][2], sy = 1 << ADAM7[z][3];
        uint32_t subx = (uint32_t)(((long)w + sx - ADAM7[z][0] - 1) >> ADAM7[z][2]);
        uint32_t suby = (uint32_t)(((long)h + sy - ADAM7[z][1] - 1) >> ADAM7[z][3]);
        if (subx == 0 || suby == 0) continue;
        out[n++] = (pass_t){ADAM7[z][0], ADAM7[z][1], sx, sy, subx, suby,
                            ((size_t)subx * (size_t)volume + 7) >> 3};
    }
    return n;
}

size_t orc_png_filtered_size(uint32_t w, uint32_t h, int volume, int interlaced)
{
    pass_t p[7];
    int    n = passes(w, h, volume, interlaced, p);
    size_t total = 0;
    for (int z = 0; z < n; ++z) total += (size_t)p[z].h * (p[z].pitch + 1);
    return total;
}

/* PNG.Decoder.push row loop, PNG.Decoder.swift:58-147 */
int orc_png_unfilter(const uint8_t* filtered, size_t n, uint32_t w, uint32_t h, int volume,
                     int depth, int interlaced, uint8_t* storage)
{
    pass_t p[7];
    int    np    = passes(w, h, volume, interlaced, p);
    int    delay = (volume + 7) >> 3;
    size_t at    = 0;
    for (int z = 0; z < np; ++z) {
        size_t   count = p[z].pitch + 1;
        uint8_t* last  = (uint8_t*)calloc(count, 1);
        uint8_t* line  = (uint8_t*)malloc(count);
        for (uint32_t y = 0; y < p[z].h; ++y) {
            if (at + count > n) { /* inflator.pull(count) == nil */
                free(last);
                free(line);
                return ORC_NEED_MORE_INPUT;
            }
            memcpy(line, filtered + at, count);
            at += count;
            orc_defilter(line, last, count, delay);
            assign_row(storage, w, line + 1, p[z].bx, p[z].by + (int)y * p[z].sy, p[z].sx, volume,
                       depth);
            uint8_t* t = last; last = line; line = t;
        }
        free(last);
        free(line);
    }
    return at == n ? ORC_OK : ORC_ERR_PNG_EXTRANEOUS_IMAGE_DATA; /* PNG.Decoder.swift:142-147 */
}

int orc_png_decode(int format, const uint8_t* idat, size_t n, uint32_t w, uint32_t h, int volume,
                   int depth, int interlaced, uint8_t* storage, orc_inflate_result* res)
{
    size_t   need = orc_png_filtered_size(w, h, volume, interlaced);
    size_t   cap  = need + 65536; /* room to detect extraneous image data */
    /* the inflated stream lives in a per-thread scratch that is kept between calls: a fresh 100+ MB
     * malloc per image means a page-fault storm when 64 host threads decode 8K images side by side,
     * which says nothing about the algorithm (the reference's Inflator keeps its window buffer too) */
    static __thread uint8_t* scratch = NULL;
    static __thread size_t   scratch_cap = 0;
    if (scratch_cap < cap) {
        free(scratch);
        scratch = (uint8_t*)malloc(cap ? cap : 1);
        scratch_cap = scratch ? cap : 0;
    }
    uint8_t* filtered = scratch;
    orc_inflate(format, idat, n, filtered, cap, res);
    int st = res->status;
    if (st == ORC_ERR_OUTPUT_CAPACITY) st = ORC_ERR_PNG_EXTRANEOUS_IMAGE_DATA;
    if (st == ORC_OK || st == ORC_NEED_MORE_INPUT) {
        int u = orc_png_unfilter(filtered, (size_t)res->produced, w, h, volume, depth, interlaced,
                                 storage);
        /* a complete stream with too few rows is NOT an error in the reference: Decoder.push
         * returns `continue` == nil and the IEND check (PNG.Context.swift:134-141) passes; the
         * missing rows of `storage` are simply never assigned. */
        if (u == ORC_ERR_PNG_EXTRANEOUS_IMAGE_DATA) st = u;
        else if (st == ORC_NEED_MORE_INPUT)
            st = ORC_ERR_PNG_INCOMPLETE_DATASTREAM; /* PNG.Context.swift:134-141 at IEND */
    }
    return st;
}

/* Encoder.score, PNG.Encoder.swift:230-234 */
static long score(const uint8_t* p, size_t n)
{
    long s = 0;
    for (size_t i = 0; i < n; ++i) {
        int v = (int8_t)p[i];
        s += v < 0 ? -v : v;
    }
    return s;
}

/* Encoder.filter, PNG.Encoder.swift:132-204.  line/last: count bytes, byte 0 ignored */
void orc_filter_row(const uint8_t* line, const uint8_t* last, size_t count, int delay, uint8_t* out)
{
    size_t   d = (size_t)delay, pitch = count - 1;
    uint8_t* cand = (uint8_t*)malloc(5 * count);
    uint8_t *c0 = cand, *c1 = cand + count, *c2 = cand + 2 * count, *c3 = cand + 3 * count,
            *c4 = cand + 4 * count;
    memcpy(c0, line, count);
    c0[0] = 0; c1[0] = 1; c2[0] = 2; c3[0] = 3; c4[0] = 4;
    for (size_t i = 1; i < count; ++i) {
        uint8_t x = line[i], b = last[i];
        uint8_t a = i > d ? line[i - d] : 0, c = i > d ? last[i - d] : 0;
        c1[i] = (uint8_t)(x - a);
        c2[i] = (uint8_t)(x - b);
        c3[i] = (uint8_t)(x - (uint8_t)(((unsigned)a + (unsigned)b) >> 1));
        c4[i] = (uint8_t)(x - orc_paeth(a, b, c));
    }
    int  best = 0;
    long minimum = -1;
    for (int f = 0; f < 5; ++f) {
        long s = score(cand + (size_t)f * count + 1, pitch);
        if (minimum < 0 || s < minimum) { /* strict <: first minimum wins */
            minimum = s;
            best = f;
        }
    }
    memcpy(out, cand + (size_t)best * count, count);
    free(cand);
}

/* PNG.Encoder.pull row loop (filtering half), PNG.Encoder.swift:33-129 */
size_t orc_png_filter(const uint8_t* storage, uint32_t w, uint32_t h, int volume, int depth,
                      int interlaced, uint8_t* filtered, size_t cap)
{
    pass_t p[7];
    int    np    = passes(w, h, volume, interlaced, p);
    int    delay = (volume + 7) >> 3;
    size_t at    = 0;
    for (int z = 0; z < np; ++z) {
        size_t   count = p[z].pitch + 1;
        uint8_t* last  = (uint8_t*)calloc(count, 1);
        uint8_t* line  = (uint8_t*)calloc(count, 1);
        for (uint32_t y = 0; y < p[z].h; ++y) {
            if (at + count > cap) { free(last); free(line); return (size_t)-1; }
            line[0] = 0;
            collect_row(storage, w, line + 1, p[z].pitch, p[z].bx, p[z].by + (int)y * p[z].sy,
                        p[z].sx, volume, depth);
            orc_filter_row(line, last, count, delay, filtered + at);
            at += count;
            uint8_t* t = last; last = line; line = t;
        }
        free(last);
        free(line);
    }
    return at;
}
