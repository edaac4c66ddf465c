/*
 * oracle/lz77_deflate.c -- CPU restatement of LZ77.Deflator / Gzip.Deflator (all 14 levels).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Parity: pinned byte-for-byte by the reference's
 * committed level-9 outputs (Tests/Outputs/ *.png, 28 files) and its gzip fixtures (levels 10, 13);
 * levels 0-8, 11, 12 have no byte-level goldens in the reference ("parity unpinned" at the
 * compressed-byte level; round-trip only) -- see tests/test_oracle_encode.py.
 *
 * Follows (paths relative to the reference checkout, Sources/LZ77/):
 *   Deflator/LZ77.DeflatorSearch.swift:13-34          level -> (mode, attempts, goal, iterations)
 *   Deflator/LZ77.DeflatorWindow.swift:59-212         window update (exact 4-byte-key dictionary with
 *                                                     expiry) and hash-chain match search
 *   F14/F14.HashTable.swift                           only its dictionary semantics are contractual
 *                                                     (Sources/LZ77Tests/HardwareAcceleration.swift:9-49)
 *   Deflator/LZ77.DeflatorBuffers.Stream.swift:30-404 compressBlocks / compress (greedy, lazy, full)
 *   Deflator/LZ77.DeflatorMatches.swift:55-379        term vector / graph, set(edge:), trees, minimize, explore
 *   Deflator/LZ77.DeflatorMatches.Depths.swift:32-99  quarter-bit cost table, update, generalize
 *   HuffmanCoding/LZ77.HuffmanTree.swift:206-404      codewords, init(frequencies:limit:), limitHeight
 *   HuffmanCoding/LZ77.Heap.swift:7-183               the binary heap (tie-breaking matters)
 *   Deflator/LZ77.DeflatorBuffers.Stream.swift:417-709 writeBlock (stored / dynamic), code-length RLE,
 *                                                     writeBlockMetadata / Tables / terms
 *   Deflator/LZ77.DeflatorOut.swift:105-145           LSB-first bit writer
 *   Inflator/LZ77.StreamHeader.swift:56-62, Gzip/Gzip.StreamHeader.swift:86-97   headers
 *   Deflator/LZ77.DeflatorBuffers.swift:68-137        trailers (Adler-32 BE / CRC-32 LE + ISIZE LE)
 *
 * The reference's output does not depend on push granularity, so this is a
 * one-shot push(all, last: true).
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---------------- static tables (RFC 1951; LZ77.Composites / LZ77.Decades) ---------------- */
static const uint16_t RUN_EXTRA[30] = {0, 0,0,0,0,0, 0,0,0,1,1, 1,1,2,2,2, 2,3,3,3,3, 4,4,4,4,5, 5,5,5,0};
static const uint16_t RUN_BASE[30]  = {0, 3,4,5,6,7, 8,9,10,11,13, 15,17,19,23,27, 31,35,43,51,59,
                                       67,83,99,115,131, 163,195,227,258};
static const uint16_t DIST_EXTRA[30] = {0,0,0,0,1, 1,2,2,3,3, 4,4,5,5,6, 6,7,7,8,8, 9,9,10,10,11, 11,12,12,13,13};
static const uint16_t DIST_BASE[30]  = {1,2,3,4,5, 7,9,13,17,25, 33,49,65,97,129, 193,257,385,513,769,
                                        1025,1537,2049,3073,4097, 6145,8193,12289,16385,24577};
static uint8_t RUN_DECADE[259];    /* LZ77.Decades[run:]      : run 3...258 -> 1...29 */
static uint8_t DIST_DECADE_LO[257], DIST_DECADE_HI[256]; /* LZ77.Decades[distance:] */
static int     tables_ready = 0;

static void init_tables(void)
{
    if (tables_ready) return;
    for (int d = 1; d <= 28; ++d)
        for (int r = RUN_BASE[d]; r < RUN_BASE[d] + (1 << RUN_EXTRA[d]) && r <= 257; ++r) RUN_DECADE[r] = (uint8_t)d;
    RUN_DECADE[258] = 29; /* "there is an overlapping composite for length = 258" (Decades.swift) */
    for (int d = 0; d < 30; ++d)
        for (int x = DIST_BASE[d]; x < DIST_BASE[d] + (1 << DIST_EXTRA[d]); ++x) {
            if (x <= 256) DIST_DECADE_LO[x] = (uint8_t)d;
            else DIST_DECADE_HI[(x - 1) >> 7] = (uint8_t)d;
        }
    tables_ready = 1;
}
static inline int dist_decade(int distance)
{
    return distance <= 256 ? DIST_DECADE_LO[distance] : DIST_DECADE_HI[(distance - 1) >> 7];
}

/* ---------------- LZ77.DeflatorOut: LSB-first bit writer ---------------- */
typedef struct { uint8_t* p; size_t cap; uint64_t bits; int overflow; } bitout;

static void put_bits(bitout* o, uint32_t v, int count)
{
    for (int i = 0; i < count; ++i) {
        size_t byte = (size_t)(o->bits >> 3);
        if (byte >= o->cap) { o->overflow = 1; return; }
        if ((o->bits & 7) == 0) o->p[byte] = 0;
        o->p[byte] |= (uint8_t)(((v >> i) & 1u) << (o->bits & 7));
        o->bits++;
    }
}
static void pad_to_byte(bitout* o) { put_bits(o, 0, (int)(-(int64_t)o->bits & 7)); }
/* Stream.writeBigEndianUInt32, Stream.swift:795-802 */
static void put_be32(bitout* o, uint32_t v)
{
    pad_to_byte(o);
    put_bits(o, v >> 24, 8); put_bits(o, (v >> 16) & 0xff, 8); put_bits(o, (v >> 8) & 0xff, 8); put_bits(o, v & 0xff, 8);
}
static void put_le32(bitout* o, uint32_t v)
{
    pad_to_byte(o);
    put_bits(o, v & 0xff, 8); put_bits(o, (v >> 8) & 0xff, 8); put_bits(o, (v >> 16) & 0xff, 8); put_bits(o, v >> 24, 8);
}

/* ---------------- LZ77.HuffmanTree (encoder half) ---------------- */
typedef struct {
    int symbols[320];    /* ordered by (length, symbol) */
    int lo[15], hi[15];  /* levels[l-1] */
    int nsym;
} etree;

/* LZ77.Heap<Int, [Int]>: key = frequency, value = leaf counts per level (deepest first, root last) */
typedef struct { long key; int n; int v[96]; } hnode;
typedef struct { hnode* a; int count; } heap_t; /* 1-based via a[i-1] */

static int heap_lowest(heap_t* h, int parent)
{
    int r = (parent << 1) + 1, l = parent << 1, end = 1 + h->count;
    if (l >= end) return 0;
    if (r >= end) return h->a[l - 1].key < h->a[parent - 1].key ? l : 0;
    int c = h->a[r - 1].key < h->a[l - 1].key ? r : l;
    return h->a[c - 1].key < h->a[parent - 1].key ? c : 0;
}
static void heap_swap(heap_t* h, int i, int j) { hnode t = h->a[i - 1]; h->a[i - 1] = h->a[j - 1]; h->a[j - 1] = t; }
static void heap_sift_down(heap_t* h, int i)
{
    int c;
    while ((c = heap_lowest(h, i)) != 0) { heap_swap(h, i, c); i = c; }
}
static void heap_sift_up(heap_t* h, int i)
{
    for (;;) {
        int p = i >> 1;
        if (p < 1 || !(h->a[i - 1].key < h->a[p - 1].key)) return;
        heap_swap(h, i, p);
        i = p;
    }
}
static int heap_dequeue(heap_t* h, hnode* out)
{
    if (h->count == 0) return 0;
    if (h->count == 1) { *out = h->a[0]; h->count = 0; return 1; }
    heap_swap(h, 1, h->count);
    *out = h->a[h->count - 1];
    h->count--;
    heap_sift_down(h, 1);
    return 1;
}
static void heap_enqueue(heap_t* h, const hnode* n)
{
    h->a[h->count++] = *n;
    heap_sift_up(h, h->count);
}

/* HuffmanTree.limitHeight, HuffmanTree.swift:348-404.  levels[0] = leaves at depth 1 */
static int limit_height(int* levels, int count, int height)
{
    if (count <= height) return count;
    long unhoused = 0;
    for (int l = count - 1; l >= height; --l) {
        int pairs = levels[l] >> 1;
        unhoused += pairs;
        levels[l - 1] += pairs;
    }
    int split = height - 2;
    while (unhoused > 0) {
        if (levels[split] <= 0) { split -= 1; continue; }
        long resettled = levels[split] < unhoused ? levels[split] : unhoused;
        unhoused -= resettled;
        levels[split] -= (int)resettled;
        levels[split + 1] += (int)(2 * resettled);
        if (split < height - 2) split += 1;
    }
    return height;
}

static int cmp_int(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }

/* HuffmanTree.init(frequencies:limit:), HuffmanTree.swift:247-344; init(stub:) :52-65 */
static void etree_build(etree* t, const long* freq, int n, int limit)
{
    int syms[320], ns = 0;
    for (int i = 0; i < n; ++i) if (freq[i] > 0) syms[ns++] = i;
    /* stable sort by decreasing frequency (insertion sort keeps ascending symbol order on ties) */
    for (int i = 1; i < ns; ++i) {
        int s = syms[i], j = i;
        while (j > 0 && freq[syms[j - 1]] < freq[s]) { syms[j] = syms[j - 1]; --j; }
        syms[j] = s;
    }
    t->nsym = ns;
    for (int i = 0; i < 15; ++i) t->lo[i] = t->hi[i] = 0;
    if (ns <= 1) { /* stub */
        if (ns == 1) t->symbols[0] = syms[0];
        t->lo[0] = 0; t->hi[0] = ns;
        for (int i = 1; i < 15; ++i) t->lo[i] = t->hi[i] = ns;
        return;
    }
    hnode  nodes[320];
    heap_t h = {nodes, 0};
    for (int i = ns - 1; i >= 0; --i) { /* symbols.reversed() */
        hnode* x = &nodes[h.count++];
        x->key = freq[syms[i]]; x->n = 1; x->v[0] = 1;
    }
    for (int i = h.count >> 1; i >= 1; --i) heap_sift_down(&h, i); /* heapify */
    hnode first, second;
    int   leaves[96], nl = 0;
    for (;;) {
        heap_dequeue(&h, &first);
        if (!heap_dequeue(&h, &second)) {
            /* first.value.dropLast().reversed() */
            nl = first.n - 1;
            for (int i = 0; i < nl; ++i) leaves[i] = first.v[nl - 1 - i];
            break;
        }
        hnode *merged = first.n > second.n ? &first : &second, *mergee = first.n > second.n ? &second : &first;
        hnode m = *merged;
        for (int i = 0; i < mergee->n; ++i) m.v[m.n - 1 - i] += mergee->v[mergee->n - 1 - i];
        m.v[m.n++] = 0;
        m.key = first.key + second.key;
        heap_enqueue(&h, &m);
    }
    nl = limit_height(leaves, nl, limit);
    int base = 0;
    for (int i = 0; i < 15; ++i) {
        int c = i < nl ? leaves[i] : 0;
        t->lo[i] = base; t->hi[i] = base + c; base += c;
    }
    memcpy(t->symbols, syms, sizeof(int) * (size_t)ns);
    for (int i = 0; i < 15; ++i) qsort(t->symbols + t->lo[i], (size_t)(t->hi[i] - t->lo[i]), sizeof(int), cmp_int);
}

/* HuffmanTree.codewords + LZ77.Codeword, HuffmanTree.swift:206-230, Codeword.swift:22-38 */
typedef struct { uint16_t bits; uint8_t length; } codeword;
static void etree_codewords(const etree* t, codeword* dst, int count)
{
    memset(dst, 0, sizeof(codeword) * (size_t)count);
    uint32_t counter = 0;
    for (int l = 1; l <= 15; ++l) {
        for (int i = t->lo[l - 1]; i < t->hi[l - 1]; ++i) {
            uint32_t rev = 0;
            for (int k = 0; k < l; ++k) if (counter & (1u << k)) rev |= 1u << (l - 1 - k);
            dst[t->symbols[i]].bits = (uint16_t)rev;
            dst[t->symbols[i]].length = (uint8_t)l;
            counter += 1;
        }
        counter <<= 1;
    }
}
static void etree_lengths(const etree* t, uint8_t* lengths, int offset)
{
    for (int l = 1; l <= 15; ++l)
        for (int i = t->lo[l - 1]; i < t->hi[l - 1]; ++i) lengths[offset + t->symbols[i]] = (uint8_t)l;
}

/* ---------------- LZ77.DeflatorMatches.Depths ---------------- */
typedef struct { uint8_t storage[542]; uint8_t dflt[542]; int generic; } depths_t;
static void depths_init(depths_t* d)
{
    for (int i = 0; i < 256; ++i) d->dflt[i] = 33;
    for (int run = 3; run <= 258; ++run) d->dflt[253 + run] = (uint8_t)(30 + (RUN_EXTRA[RUN_DECADE[run]] << 2));
    for (int k = 0; k < 30; ++k) d->dflt[512 + k] = (uint8_t)(19 + (DIST_EXTRA[k] << 2));
    memcpy(d->storage, d->dflt, 542);
    d->generic = 1;
}
/* Depths.update, Depths.swift:53-86 (iteration order = (length, symbol): later writes win) */
static void depths_update(depths_t* d, const etree* rl, const etree* dist)
{
    for (int l = 1; l <= 15; ++l)
        for (int i = rl->lo[l - 1]; i < rl->hi[l - 1]; ++i) {
            int sym = rl->symbols[i];
            if (sym < 256) d->storage[sym] = (uint8_t)(l << 2);
            else if (sym > 256) {
                int dec = sym & 0xff, len = l + RUN_EXTRA[dec], base = 253 + RUN_BASE[dec], count = 1 << RUN_EXTRA[dec];
                for (int k = base; k < base + count; ++k) d->storage[k] = (uint8_t)(len << 2);
            }
        }
    for (int l = 1; l <= 15; ++l)
        for (int i = dist->lo[l - 1]; i < dist->hi[l - 1]; ++i) {
            int sym = dist->symbols[i];
            d->storage[512 + sym] = (uint8_t)((l + DIST_EXTRA[sym]) << 2);
        }
    d->generic = 0;
}
static void depths_generalize(depths_t* d)
{
    for (int i = 0; i < 542; ++i) {
        uint8_t s = d->storage[i], g = d->dflt[i];
        d->storage[i] = (uint8_t)((s & g) + ((s ^ g) >> 1));
    }
}

/* ---------------- the deflator ---------------- */
enum { MODE_GREEDY, MODE_LAZY, MODE_FULL };
typedef struct { int mode; long attempts; int goal, iterations; } search_t;

/* LZ77.DeflatorSearch.init(level:), DeflatorSearch.swift:13-34 */
static search_t search_for(int level)
{
    static const search_t T[13] = {
        {MODE_GREEDY, 1, 6, 0}, {MODE_GREEDY, 2, 8, 0}, {MODE_GREEDY, 4, 10, 0}, {MODE_GREEDY, 40, 24, 0},
        {MODE_LAZY, 20, 32, 0}, {MODE_LAZY, 40, 54, 0}, {MODE_LAZY, 64, 80, 0}, {MODE_LAZY, 100, 160, 0},
        {MODE_FULL, 14, 20, 1}, {MODE_FULL, 20, 32, 2}, {MODE_FULL, 30, 50, 3}, {MODE_FULL, 60, 80, 4},
        {MODE_FULL, 100, 133, 5}};
    if (level <= 0) return T[0];
    if (level <= 12) return T[level];
    search_t s = {MODE_FULL, 0x7fffffffffffffffL, 258, 6};
    return s;
}

typedef struct { uint32_t upstream, depth; uint32_t edge[30]; } vertex_t; /* 32 x u32, DeflatorMatches.swift:55-58 */
typedef struct { uint16_t run, dist; uint8_t lit; } term_t;              /* run == 0: literal */

typedef struct {
    const uint8_t* x;        /* input */
    int64_t        n;
    search_t       search;
    int64_t        mask;     /* window size - 1 */
    /* LZ77.DeflatorWindow: endIndex = position of the next literal (starts at -3) */
    int64_t        end_index;
    int64_t        dequeued; /* bytes taken from the input queue */
    int32_t*       head;     /* hash -> most recent position with that hash */
    int32_t*       prevh;    /* [pos & mask] previous position with the same hash */
    int32_t*       next;     /* [pos & mask] previous position with the same 4-byte key, or -1 */
    /* LZ77.DeflatorMatches */
    int64_t        limit, capacity, count;
    term_t*        terms;
    vertex_t*      graph;
    depths_t       depths;
    bitout         out;
} deflator;

#define HASH_BITS 18
static inline uint32_t key_at(const deflator* z, int64_t p)
{
    uint32_t k = 0;
    for (int i = 0; i < 4; ++i) k = k << 8 | (uint32_t)(p + i >= 0 && p + i < z->n ? z->x[p + i] : 0);
    return k;
}
static inline uint32_t hash_key(uint32_t k) { return (k * 2654435761u) >> (32 - HASH_BITS); }

static inline int64_t input_count(const deflator* z) { return z->n - z->dequeued; }

/* DeflatorWindow.update, DeflatorWindow.swift:78-113: returns the position `a` just entered and
 * sets *next to the most recent earlier position with the same 4-byte key inside the window (-1) */
static int64_t window_update(deflator* z, int64_t* next)
{
    int64_t  a = z->end_index;
    uint32_t k = key_at(z, a), h = hash_key(k);
    int64_t  p = z->head[h], found = -1;
    while (p >= 0 && a - p <= z->mask) {
        if (key_at(z, p) == k) { found = p; break; }
        p = z->prevh[p & z->mask];
    }
    z->next[a & z->mask]  = (int32_t)found;
    z->prevh[a & z->mask] = z->head[h];
    z->head[h]            = (int32_t)a;
    z->end_index += 1;
    z->dequeued += 1;
    if (next) *next = found;
    return a;
}

typedef void (*match_fn)(void* ctx, int run, int distance);

/* DeflatorWindow.match(from:lookahead:attempts:goal:delegate:), DeflatorWindow.swift:132-212 */
static void window_match(const deflator* z, int64_t a, int64_t next, match_fn delegate, void* ctx)
{
    if (next < 0) return;
    int64_t lookahead = input_count(z);
    int     limit = (int)(lookahead + 4 < 258 ? lookahead + 4 : 258);
    int64_t current = next, distance = a - current;
    long    remaining = z->search.attempts;
    const uint8_t* v = z->x + a;
    for (;;) {
        int run = 4;
        int amax = (int)(distance < limit ? distance : limit);
        int broke = 0;
        while (run < amax) {
            if (z->x[current + run] != v[run]) { broke = 1; break; }
            run += 1;
        }
        if (!broke) {
            int i = 4 - distance > 0 ? (int)(4 - distance) : 0;
            while (run < limit && v[i] == v[run]) { i += 1; run += 1; }
        }
        delegate(ctx, run, (int)distance);
        remaining -= 1;
        if (!(remaining > 0 && z->search.goal > run)) break;
        int64_t nx = z->next[current & z->mask];
        if (nx < 0) break;
        distance += current - nx;
        current = nx;
        if (!(distance < z->mask)) break;
    }
}

typedef struct { int run, distance; } best_t;
static void best_delegate(void* ctx, int run, int distance)
{
    best_t* b = (best_t*)ctx;
    if (b->run < run) { b->run = run; b->distance = distance; }
}
/* DeflatorWindow.match(...) -> (run, distance)?, DeflatorWindow.swift:115-130: accepts run > 5 only */
static int window_best(const deflator* z, int64_t a, int64_t next, best_t* best)
{
    best->run = 5; best->distance = 1;
    window_match(z, a, next, best_delegate, best);
    return best->run > 5;
}

static inline int64_t unfilled(const deflator* z) { return z->limit - 1 - z->count; }
static inline uint8_t literal_at(const deflator* z, int64_t a) { return a >= 0 && a < z->n ? z->x[a] : 0; }

static void store_literal(deflator* z, uint8_t lit) { term_t t = {0, 0, lit}; z->terms[z->count++] = t; }
static void store_match(deflator* z, int run, int distance) { term_t t = {(uint16_t)run, (uint16_t)(distance - 1), 0}; z->terms[z->count++] = t; }
/* DeflatorMatches.store(vertex:), DeflatorMatches.swift:161-179 */
static int64_t store_vertex(deflator* z, uint8_t lit)
{
    vertex_t* v = &z->graph[z->count];
    v->upstream = lit;
    v->depth = 0xffffffffu;
    memset(v->edge, 0, sizeof v->edge);
    return z->count++;
}
typedef struct { deflator* z; int64_t index; int extent; } edge_ctx;
/* DeflatorMatches.set(edge:at:), DeflatorMatches.swift:180-194: longest run per distance decade, first wins */
static void edge_delegate(void* ctx, int run, int distance)
{
    edge_ctx* e = (edge_ctx*)ctx;
    if (run > e->extent) e->extent = run;
    uint32_t* slot = &e->z->graph[e->index].edge[dist_decade(distance)];
    if ((uint32_t)run > (*slot & 0xffffu)) *slot = (uint32_t)distance << 16 | (uint32_t)run;
}

/* Stream.compress(all: true), Stream.swift:64-404.  returns 1 when the match buffer is full */
static int compress(deflator* z)
{
    while (z->end_index < 0 && input_count(z) > 0) { /* DeflatorWindow.initialize */
        z->end_index += 1;
        z->dequeued += 1;
    }
    int64_t next;
    if (z->search.mode == MODE_GREEDY) {
        while (input_count(z) > 0) {
            if (unfilled(z) <= 0) return 1;
            int64_t a = window_update(z, &next);
            best_t  m;
            if (window_best(z, a, next, &m)) {
                for (int k = 1; k < m.run; ++k) window_update(z, NULL);
                store_match(z, m.run, m.distance);
            } else {
                store_literal(z, literal_at(z, a));
            }
        }
    } else if (z->search.mode == MODE_LAZY) {
        while (input_count(z) > 0) {
            if (unfilled(z) <= 1) return 1;
            int64_t a = window_update(z, &next);
            uint8_t first = literal_at(z, a);
            best_t  eager, lazy;
            if (window_best(z, a, next, &eager)) {
                int64_t a1 = window_update(z, &next);
                if (window_best(z, a1, next, &lazy) && eager.run < lazy.run) {
                    store_literal(z, first);
                    store_match(z, lazy.run, lazy.distance);
                    for (int k = 1; k < lazy.run; ++k) window_update(z, NULL);
                } else {
                    store_match(z, eager.run, eager.distance);
                    for (int k = 2; k < eager.run; ++k) window_update(z, NULL);
                }
            } else {
                store_literal(z, first);
            }
        }
    } else {
        while (input_count(z) > 0) {
            if (unfilled(z) <= 0) return 1;
            int64_t  a = window_update(z, &next);
            edge_ctx e = {z, store_vertex(z, literal_at(z, a)), 1};
            window_match(z, a, next, edge_delegate, &e);
            int64_t skip = e.extent - 100 < unfilled(z) ? e.extent - 100 : unfilled(z);
            for (int64_t k = 0; k < skip; ++k) {
                int64_t b = window_update(z, NULL);
                store_vertex(z, literal_at(z, b));
            }
        }
    }
    /* epilogue: the literals still sitting in the 3-byte pipeline */
    int64_t epilogue = -3 - (z->end_index < 0 ? z->end_index : 0);
    while (input_count(z) > epilogue) {
        if (unfilled(z) <= 0) return 1;
        int64_t a = window_update(z, NULL);
        if (z->search.mode == MODE_FULL) store_vertex(z, literal_at(z, a));
        else store_literal(z, literal_at(z, a));
    }
    return 0;
}

/* DeflatorMatches.explore, DeflatorMatches.swift:324-379 */
static void explore(deflator* z, int64_t index)
{
    vertex_t* g = z->graph;
    const uint32_t cur_up = g[index].upstream, cur_depth = g[index].depth;
    uint32_t ld = cur_depth + z->depths.storage[cur_up & 0xff];
    if (ld < g[index + 1].depth) {
        g[index + 1].upstream = 0x0001ff00u | (g[index + 1].upstream & 0xffu);
        g[index + 1].depth = ld;
    }
    int64_t remaining = z->count - index;
    if (remaining < 3) return;
    for (int decade = 0; decade < 30; ++decade) {
        int64_t maxlength = g[index].edge[decade] & 0xffffu;
        if (maxlength > remaining) maxlength = remaining;
        if (maxlength <= 0) continue;
        uint32_t depth = cur_depth + z->depths.storage[512 + decade];
        for (int64_t length = 3; length <= maxlength; ++length) {
            uint32_t d = depth + z->depths.storage[253 + length];
            vertex_t* nx = &g[index + length];
            if (!(d < nx->depth)) continue;
            nx->upstream = (uint32_t)length << 16 | (uint32_t)decade << 8 | (nx->upstream & 0xffu);
            nx->depth = d;
        }
    }
}

/* DeflatorMatches.minimize, DeflatorMatches.swift:265-321 */
static void minimize(deflator* z, long* freq /* [320] */)
{
    vertex_t* g = z->graph;
    g[0].depth = 0;
    g[z->count].depth = 0xffffffffu;
    for (int64_t node = 0; node < z->count; ++node) explore(z, node);
    memset(freq, 0, sizeof(long) * 320);
    int64_t  ci = z->count;
    uint32_t cu = g[ci].upstream;
    if (z->count > 0) do {
        int64_t  length = cu >> 16;
        int64_t  ni = ci - length;
        uint32_t nu = g[ni].upstream;
        g[ni].upstream = (cu & 0xffffff00u) | (nu & 0xffu);
        if (length == 1) freq[nu & 0xff] += 1;
        else {
            freq[256 | RUN_DECADE[length]] += 1;
            freq[288 + ((cu >> 8) & 0xff)] += 1;
        }
        ci = ni;
        cu = nu;
    } while (ci > 0);
    freq[256] = 1;
}

static void write_codeword(bitout* o, codeword c) { put_bits(o, c.bits, c.length); }

/* Stream.writeBlock(final:), Stream.swift:440-571 (+ writeBlockMetadata :577-612, Tables :615-623,
 * terms :626-709) */
static void write_block(deflator* z, int final)
{
    etree rl, dist, meta;
    long  freq[320];
    if (z->search.mode != MODE_FULL) { /* DeflatorMatches.trees(), :138-159 */
        memset(freq, 0, sizeof freq);
        for (int64_t i = 0; i < z->count; ++i) {
            term_t t = z->terms[i];
            if (t.run == 0) { freq[t.lit] += 1; freq[288 + 31] += 1; }
            else { freq[256 | RUN_DECADE[t.run]] += 1; freq[288 + dist_decade(t.dist + 1)] += 1; }
        }
        freq[256] = 1;
        etree_build(&rl, freq, 286, 15);
        etree_build(&dist, freq + 288, 30, 15);
    } else { /* DeflatorMatches.trees(iterations:), :225-260 */
        z->limit = 2 * z->limit < z->capacity ? 2 * z->limit : z->capacity;
        int i = z->depths.generic ? -z->search.iterations : 0;
        for (;;) {
            minimize(z, freq);
            etree_build(&rl, freq, 286, 15);
            etree_build(&dist, freq + 288, 30, 15);
            i += 1;
            if (!(i < z->search.iterations)) break;
            depths_update(&z->depths, &rl, &dist);
            for (int64_t k = 0; k < z->count; ++k) z->graph[k].depth = 0xffffffffu;
        }
    }
    uint8_t lengths[318];
    memset(lengths, 0, sizeof lengths);
    etree_lengths(&rl, lengths, 0);
    int r = 286;
    while (r > 0 && lengths[r - 1] == 0) --r;
    if (r < 257) r = 257;
    etree_lengths(&dist, lengths, r);
    int d = 32;
    while (d > 0 && lengths[r + d - 1] == 0) --d;
    if (d < 1) d = 1;
    /* code-length run-length terms, Stream.swift:482-543 */
    struct { uint8_t symbol, bits; } mt[320];
    int nmt = 0, repetitions = 1;
    uint8_t last = lengths[0];
    for (int at = 1;; ++at) {
        int have = at < r + d;
        if (have && lengths[at] == last) { repetitions += 1; continue; }
        if (last == 0) {
            while (repetitions > 138) { mt[nmt].symbol = 18; mt[nmt++].bits = 138 - 11; repetitions -= 138; }
            if (repetitions > 2) {
                if (repetitions < 11) { mt[nmt].symbol = 17; mt[nmt++].bits = (uint8_t)(repetitions - 3); }
                else { mt[nmt].symbol = 18; mt[nmt++].bits = (uint8_t)(repetitions - 11); }
            } else for (int k = 0; k < repetitions; ++k) { mt[nmt].symbol = 0; mt[nmt++].bits = 0; }
        } else {
            mt[nmt].symbol = last; mt[nmt++].bits = 0;
            repetitions -= 1;
            while (repetitions > 6) { mt[nmt].symbol = 16; mt[nmt++].bits = 6 - 3; repetitions -= 6; }
            if (repetitions > 2) { mt[nmt].symbol = 16; mt[nmt++].bits = (uint8_t)(repetitions - 3); }
            else for (int k = 0; k < repetitions; ++k) { mt[nmt].symbol = last; mt[nmt++].bits = 0; }
        }
        if (!have) break;
        last = lengths[at];
        repetitions = 1;
    }
    long mfreq[19] = {0};
    for (int i = 0; i < nmt; ++i) mfreq[mt[i].symbol] += 1;
    etree_build(&meta, mfreq, 19, 7);
    /* writeBlockMetadata */
    static const int ZPOS[19] = {3, 17, 15, 13, 11, 9, 7, 5, 4, 6, 8, 10, 12, 14, 16, 18, 0, 1, 2};
    int cl[19] = {0};
    for (int l = 1; l <= 8; ++l)
        for (int i = meta.lo[l - 1]; i < meta.hi[l - 1]; ++i) cl[ZPOS[meta.symbols[i]]] = l;
    int ncl = 19;
    while (ncl > 0 && cl[ncl - 1] == 0) --ncl;
    if (ncl < 4) ncl = 4;
    put_bits(&z->out, final ? 5 : 4, 3); /* 0b10_1 / 0b10_0 */
    put_bits(&z->out, (uint32_t)(r - 257), 5);
    put_bits(&z->out, (uint32_t)(d - 1), 5);
    put_bits(&z->out, (uint32_t)(ncl - 4), 4);
    for (int i = 0; i < ncl; ++i) put_bits(&z->out, (uint32_t)cl[i], 3);
    /* writeBlockTables */
    codeword cw_rl[288], cw_d[32], cw_m[19];
    etree_codewords(&rl, cw_rl, 288);
    etree_codewords(&dist, cw_d, 32);
    etree_codewords(&meta, cw_m, 19);
    for (int i = 0; i < nmt; ++i) {
        write_codeword(&z->out, cw_m[mt[i].symbol]);
        int extra = mt[i].symbol == 18 ? 7 : mt[i].symbol == 17 ? 3 : mt[i].symbol == 16 ? 2 : 0;
        put_bits(&z->out, mt[i].bits, extra);
    }
    /* writeBlock(with:) */
    if (z->search.mode != MODE_FULL) {
        for (int64_t i = 0; i < z->count; ++i) {
            term_t t = z->terms[i];
            if (t.run == 0) write_codeword(&z->out, cw_rl[t.lit]);
            else {
                int rd = RUN_DECADE[t.run], dd = dist_decade(t.dist + 1);
                write_codeword(&z->out, cw_rl[256 | rd]);
                put_bits(&z->out, (uint32_t)(t.run - RUN_BASE[rd]), RUN_EXTRA[rd]);
                write_codeword(&z->out, cw_d[dd]);
                put_bits(&z->out, (uint32_t)(t.dist + 1 - DIST_BASE[dd]), DIST_EXTRA[dd]);
            }
        }
        write_codeword(&z->out, cw_rl[256]);
        z->count = 0; /* resetTerms */
    } else {
        int64_t index = 0;
        while (index < z->count) {
            uint32_t up = z->graph[index].upstream;
            int64_t  count = up >> 16;
            if (count == 1) write_codeword(&z->out, cw_rl[up & 0xff]);
            else {
                int rd = RUN_DECADE[count], dd = (int)((up >> 8) & 0xff);
                uint32_t offset = z->graph[index].edge[dd] >> 16;
                write_codeword(&z->out, cw_rl[256 | rd]);
                put_bits(&z->out, (uint32_t)(count - RUN_BASE[rd]), RUN_EXTRA[rd]);
                write_codeword(&z->out, cw_d[dd]);
                put_bits(&z->out, offset - DIST_BASE[dd], DIST_EXTRA[dd]);
            }
            index += count;
        }
        write_codeword(&z->out, cw_rl[256]);
        z->count = 0; /* resetGraph */
        depths_generalize(&z->depths);
    }
}

size_t orc_deflate_bound(size_t n) { return n + n / 2 + 4096; }

/* Test hook for the reference's `Matching` KAT (Sources/LZ77Tests/Bitstreams.swift:96-185): the
 * greedy segmentation of `in` with an explicit window exponent (the KAT uses 4), attempts and goal.
 * Writes one (run, distance) pair per term (run 1, distance 0 for a literal); returns the count. */
size_t orc_debug_greedy_parse(const uint8_t* in, size_t n, int exponent, long attempts, int goal,
                              int* runs, int* dists, size_t cap)
{
    init_tables();
    deflator z;
    memset(&z, 0, sizeof z);
    z.x = in; z.n = (int64_t)n;
    z.search.mode = MODE_GREEDY; z.search.attempts = attempts; z.search.goal = goal;
    z.mask = ((int64_t)1 << exponent) - 1;
    z.end_index = -3;
    z.limit = (int64_t)n + 8; z.capacity = z.limit;
    z.terms = (term_t*)malloc(sizeof(term_t) * (n + 8));
    z.head = (int32_t*)malloc(sizeof(int32_t) << HASH_BITS);
    for (size_t i = 0; i < ((size_t)1 << HASH_BITS); ++i) z.head[i] = -1;
    z.prevh = (int32_t*)malloc(sizeof(int32_t) * (size_t)(z.mask + 1));
    z.next = (int32_t*)malloc(sizeof(int32_t) * (size_t)(z.mask + 1));
    compress(&z);
    size_t k = 0;
    for (; k < (size_t)z.count && k < cap; ++k) {
        runs[k] = z.terms[k].run ? z.terms[k].run : 1;
        dists[k] = z.terms[k].run ? z.terms[k].dist + 1 : 0;
    }
    free(z.head); free(z.prevh); free(z.next); free(z.terms);
    return k;
}

size_t orc_deflate(int format, int level, int exponent, const uint8_t* in, size_t n, uint8_t* out, size_t cap)
{
    init_tables();
    if (exponent < 8 || exponent > 15) return (size_t)-1;
    if (format == ORC_FORMAT_IOS) exponent = 15;
    deflator z;
    memset(&z, 0, sizeof z);
    z.x = in; z.n = (int64_t)n;
    z.search = search_for(level);
    z.mask = ((int64_t)1 << exponent) - 1;
    z.end_index = -3;
    z.out.p = out; z.out.cap = cap;
    z.limit = 2048; /* DeflatorMatches.init ignores its `limit:` argument (:70) */
    if (z.search.mode == MODE_FULL) {
        z.capacity = (int64_t)1 << 21;
        size_t want = n + 16 < (size_t)z.capacity ? n + 16 : (size_t)z.capacity;
        z.graph = (vertex_t*)malloc(sizeof(vertex_t) * (want + 1));
        depths_init(&z.depths);
    } else {
        z.capacity = 1 << 15;
        z.terms = (term_t*)malloc(sizeof(term_t) * 2048);
    }
    z.head = (int32_t*)malloc(sizeof(int32_t) << HASH_BITS);
    for (size_t i = 0; i < ((size_t)1 << HASH_BITS); ++i) z.head[i] = -1;
    z.prevh = (int32_t*)malloc(sizeof(int32_t) * (size_t)(z.mask + 1));
    z.next = (int32_t*)malloc(sizeof(int32_t) * (size_t)(z.mask + 1));

    /* stream header */
    if (format == ORC_FORMAT_ZLIB) { /* StreamHeader.write */
        uint32_t unpaired = (uint32_t)(exponent - 8) << 4 | 8;
        uint32_t check = ~(((unpaired << 8) | (unpaired >> 8)) % 31) & 31;
        put_bits(&z.out, check << 8 | unpaired, 16);
    } else if (format == ORC_FORMAT_GZIP) { /* Gzip.StreamHeader.write */
        put_bits(&z.out, 0x8b1f, 16); put_bits(&z.out, 0x0008, 16);
        put_bits(&z.out, 0, 16); put_bits(&z.out, 0, 16); put_bits(&z.out, 0xff00, 16);
    }
    /* Stream.compressBlocks(final: true), Stream.swift:30-60 */
    if (n >= 3) {
        while (compress(&z)) write_block(&z, 0);
        write_block(&z, 1);
    } else { /* stored final block, Stream.swift:417-435 */
        put_bits(&z.out, 1, 3);
        pad_to_byte(&z.out);
        put_bits(&z.out, (uint32_t)n, 16);
        put_bits(&z.out, ~(uint32_t)n & 0xffff, 16);
        for (size_t i = 0; i < n; ++i) put_bits(&z.out, in[i], 8);
    }
    if (format == ORC_FORMAT_ZLIB) put_be32(&z.out, orc_adler32(1, in, n));
    else if (format == ORC_FORMAT_GZIP) {
        put_le32(&z.out, orc_crc32(0, in, n));
        put_le32(&z.out, (uint32_t)n);
    }
    pad_to_byte(&z.out);
    free(z.head); free(z.prevh); free(z.next); free(z.terms); free(z.graph);
    return z.out.overflow ? (size_t)-1 : (size_t)(z.out.bits >> 3);
}
