/*
 * oracle/png_color.c -- CPU restatement of swift-png's colour targets (unpack / pack /
 * premultiply / straighten).  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restates, scalar and per pixel, what the reference expresses as generic convolve / deconvolve
 * closures: Sources/PNG/PNG.swift:149-1285, ColorTargets/PNG.RGBA.swift:262-478,
 * ColorTargets/PNG.VA.swift, ColorTargets/PNG.Color.swift.
 */
#include "oracle.h"

/* PNG.quantum (PNG.swift:255-261): T.max >> (T.bitWidth - destination) / T.max >> (T.bitWidth - source) */
static uint32_t quantum(int source, int destination)
{
    return ((1u << destination) - 1u) / ((1u << source) - 1u);
}

/* the `transform` closures of convolve(_:of:depth:kernel:) (PNG.swift:494-523): scale a sample of
 * `depth` bits to the range of a T of `tbits` bits */
static uint32_t widen(uint32_t v, int depth, int tbits)
{
    if (tbits == depth) return v;
    if (tbits > depth) return (quantum(depth, tbits) * v) & ((1u << tbits) - 1u);
    return v >> (depth - tbits);
}

/* the `transform` closures of deconvolve(_:as:depth:kernel:) (PNG.swift:1063-1095) */
static uint32_t narrow(uint32_t v, int tbits, int depth)
{
    if (tbits == depth) return v;
    if (tbits < depth) return (quantum(tbits, depth) * v) & ((1u << depth) - 1u);
    return v >> (tbits - depth);
}

/* PNG.premultiply (PNG.swift:54-66): (color * alpha + T.max >> 1) / T.max */
uint32_t orc_premultiply(uint32_t color, uint32_t alpha, int bits)
{
    const uint64_t max = (1ull << bits) - 1;
    return (uint32_t)(((uint64_t)color * alpha + (max >> 1)) / max);
}

/* PNG.straighten (PNG.swift:100-120): alpha == 0 ? premultiplied : (T.max * p + alpha >> 1) / alpha.
 * The reference's dividingFullWidth traps when the quotient overflows T (p > alpha, which no
 * premultiplied pixel has); the restatement saturates there. */
uint32_t orc_straighten(uint32_t p, uint32_t alpha, int bits)
{
    const uint64_t max = (1ull << bits) - 1;
    if (alpha == 0) return p;
    uint64_t q = (max * p + (alpha >> 1)) / alpha;
    return (uint32_t)(q > max ? max : q);
}

static int channels_of(int color)
{
    switch (color) {
    case 0: return 1;
    case 2: return 3;
    case 3: return 1;
    case 4: return 2;
    case 6: return 4;
    }
    return 0;
}

static void store(void* out, size_t i, int tbits, uint32_t v)
{
    if (tbits == 8) ((uint8_t*)out)[i] = (uint8_t)v;
    else ((uint16_t*)out)[i] = (uint16_t)v;
}
static uint32_t load(const void* in, size_t i, int tbits)
{
    return tbits == 8 ? ((const uint8_t*)in)[i] : ((const uint16_t*)in)[i];
}

int orc_unpack(const uint8_t* storage, size_t pixels, const orc_format* f, int target,
               int alpha_mode, void* out)
{
    const int tbits = (target == ORC_TARGET_RGBA8 || target == ORC_TARGET_VA8) ? 8 : 16;
    const int va    = target == ORC_TARGET_VA8 || target == ORC_TARGET_VA16;
    const int ch    = channels_of(f->color);
    const int wide  = f->depth == 16;
    const uint32_t tmax = (1u << tbits) - 1u;
    if (!ch) return ORC_ERR_BAD_ARGUMENT;
    for (size_t i = 0; i < pixels; ++i) {
        uint32_t raw[4] = {0, 0, 0, 0}, r, g, b, a;
        for (int c = 0; c < ch; ++c)  /* A(bigEndian:) (PNG.swift:152-204) */
            raw[c] = wide ? (uint32_t)storage[(i * ch + c) * 2] << 8 | storage[(i * ch + c) * 2 + 1]
                          : storage[i * ch + c];
        if (f->color == 3) {
            /* convolve(_:dereference:kernel:) over palette aggregates (PNG.swift:284-315); the
             * default deindexer is palette[i] (PNG.Color.swift) which traps out of range */
            if (raw[0] >= f->palette_count) return ORC_ERR_PALETTE_INDEX;
            const uint8_t* e = f->palette + 4 * raw[0];
            r = widen(e[0], 8, tbits), g = widen(e[1], 8, tbits), b = widen(e[2], 8, tbits);
            a = widen(e[3], 8, tbits);
        } else if (f->color == 0 || f->color == 4) {
            r = g = b = widen(raw[0], f->depth, tbits);
            if (f->color == 4) a = widen(raw[1], f->depth, tbits);
            else a = (f->has_key && raw[0] == f->key[0]) ? 0 : tmax;  /* k == key ? .min : .max */
        } else {
            uint32_t c0 = widen(raw[0], f->depth, tbits), c1 = widen(raw[1], f->depth, tbits),
                     c2 = widen(raw[2], f->depth, tbits);
            r = f->bgr ? c2 : c0, g = c1, b = f->bgr ? c0 : c2;  /* .bgr8 / .bgra8: (c.2, c.1, c.0) */
            if (f->color == 6) a = widen(raw[3], f->depth, tbits);
            else a = (f->has_key && raw[0] == f->key[0] && raw[1] == f->key[1] && raw[2] == f->key[2])
                         ? 0 : tmax;
        }
        if (alpha_mode == ORC_ALPHA_PREMULTIPLIED) {  /* RGBA.premultiplied (PNG.RGBA.swift:115-121) */
            r = orc_premultiply(r, a, tbits), g = orc_premultiply(g, a, tbits);
            b = orc_premultiply(b, a, tbits);
        } else if (alpha_mode == ORC_ALPHA_STRAIGHTENED) {  /* RGBA.straightened (:163-169) */
            r = orc_straighten(r, a, tbits), g = orc_straighten(g, a, tbits);
            b = orc_straighten(b, a, tbits);
        }
        else if (alpha_mode == ORC_ALPHA_PREMULTIPLIED_AS8 || alpha_mode == ORC_ALPHA_STRAIGHTENED_AS8) {
            /* premultiplied(as: U) / straightened(as: U), U = UInt8 (PNG.RGBA.swift:141-155, 187-201):
             * shift = T.bitWidth - 8, q = T.max / (T.max >> shift); alpha is requantised too */
            if (tbits != 16) return ORC_ERR_BAD_ARGUMENT;
            uint32_t (*op)(uint32_t, uint32_t, int) =
                alpha_mode == ORC_ALPHA_PREMULTIPLIED_AS8 ? orc_premultiply : orc_straighten;
            const uint32_t a8 = a >> 8;
            r = op(r >> 8, a8, 8) * 257u, g = op(g >> 8, a8, 8) * 257u, b = op(b >> 8, a8, 8) * 257u;
            a = a8 * 257u;
        }
        if (va) {  /* PNG.VA.unpack keeps the red sample (c.0, or c.2 of a bgr format) */
            store(out, 2 * i, tbits, r), store(out, 2 * i + 1, tbits, a);
        } else {
            store(out, 4 * i, tbits, r), store(out, 4 * i + 1, tbits, g);
            store(out, 4 * i + 2, tbits, b), store(out, 4 * i + 3, tbits, a);
        }
    }
    return ORC_OK;
}

int orc_pack(const void* pixels, size_t n, const orc_format* f, int target, uint8_t* storage)
{
    const int tbits = (target == ORC_TARGET_RGBA8 || target == ORC_TARGET_VA8) ? 8 : 16;
    const int va    = target == ORC_TARGET_VA8 || target == ORC_TARGET_VA16;
    const int ch    = channels_of(f->color);
    const int wide  = f->depth == 16;
    if (!ch) return ORC_ERR_BAD_ARGUMENT;
    for (size_t i = 0; i < n; ++i) {
        uint32_t r, g, b, a;
        if (va) r = g = b = load(pixels, 2 * i, tbits), a = load(pixels, 2 * i + 1, tbits);
        else r = load(pixels, 4 * i, tbits), g = load(pixels, 4 * i + 1, tbits),
             b = load(pixels, 4 * i + 2, tbits), a = load(pixels, 4 * i + 3, tbits);
        uint32_t s[4];
        if (f->color == 3) {
            /* deconvolve(_:reference:kernel:) (PNG.swift:818-850) + default indexer: a hash table
             * palette -> index, missing colours -> entry 0 (PNG.Color.swift).  Duplicate palette
             * entries trap in the reference (Dictionary(uniqueKeysWithValues:)); first match here. */
            uint32_t q[4] = {narrow(r, tbits, 8), narrow(g, tbits, 8), narrow(b, tbits, 8),
                             narrow(a, tbits, 8)};
            uint32_t idx = 0;
            for (uint32_t k = 0; k < f->palette_count; ++k) {
                const uint8_t* e = f->palette + 4 * k;
                if (e[0] == q[0] && e[1] == q[1] && e[2] == q[2] && e[3] == q[3]) { idx = k; break; }
            }
            storage[i] = (uint8_t)idx;
            continue;
        }
        switch (f->color) {
        case 0: s[0] = r; break;
        case 4: s[0] = r, s[1] = a; break;
        case 2: s[0] = f->bgr ? b : r, s[1] = g, s[2] = f->bgr ? r : b; break;
        default: s[0] = f->bgr ? b : r, s[1] = g, s[2] = f->bgr ? r : b, s[3] = a; break;
        }
        for (int c = 0; c < ch; ++c) {
            uint32_t v = narrow(s[c], tbits, f->depth);
            if (wide) storage[(i * ch + c) * 2] = (uint8_t)(v >> 8), storage[(i * ch + c) * 2 + 1] = (uint8_t)v;
            else storage[i * ch + c] = (uint8_t)v;
        }
    }
    return ORC_OK;
}
