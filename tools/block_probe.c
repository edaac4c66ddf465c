/*
 * tools/block_probe.c -- CPU model of the speculative DEFLATE block-boundary search planned for the
 * segment-parallel inflate.  Not product code and not the oracle: a
 * measurement tool.  For every bit offset it asks "could a dynamic-Huffman block header start here?"
 * exactly as a GPU thread would (a pure function of the bit offset), and reports at which stage each
 * offset is rejected, so that the cost of the search and its false-positive rate can be measured on
 * real streams against the true boundaries (oracle trace hook).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static inline uint32_t bits_at(const uint8_t* p, size_t nbytes, uint64_t bit, int count)
{
    uint64_t v = 0;
    size_t   byte = (size_t)(bit >> 3);
    for (int k = 0; k < 8 && byte + (size_t)k < nbytes; ++k) v |= (uint64_t)p[byte + k] << (8 * k);
    return (uint32_t)((v >> (bit & 7)) & ((1ull << count) - 1));
}

static const uint8_t CLEN_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

/* stage at which the candidate is rejected: 1 BTYPE, 2 HLIT/HDIST, 3 code-length code incomplete,
 * 4 code-length sequence (repeat without predecessor / overrun), 5 literal code incomplete or no
 * end-of-block symbol, 6 distance code invalid, 7 truncated; 0 = plausible header.
 * *header_bits receives the header's length when plausible. */
int probe_dynamic_header(const uint8_t* in, size_t n, uint64_t p, uint64_t* header_bits)
{
    const uint64_t total = (uint64_t)n * 8;
    if (p + 17 > total) return 7;
    if (bits_at(in, n, p + 1, 2) != 2) return 1;
    const uint32_t hlit = bits_at(in, n, p + 3, 5), hdist = bits_at(in, n, p + 8, 5), hclen = bits_at(in, n, p + 13, 4);
    if (hlit > 29 || hdist > 29) return 2;
    uint64_t at = p + 17;
    uint8_t  cl[19];
    memset(cl, 0, sizeof cl);
    if (at + 3 * (uint64_t)(hclen + 4) > total) return 7;
    uint32_t kraft = 0;
    for (uint32_t i = 0; i < hclen + 4; ++i, at += 3) {
        cl[CLEN_ORDER[i]] = (uint8_t)bits_at(in, n, at, 3);
        if (cl[CLEN_ORDER[i]]) kraft += 128u >> cl[CLEN_ORDER[i]];
    }
    if (kraft != 128) return 3;
    /* canonical code of the code-length alphabet (max 7 bits), decoded by linear search per length */
    uint32_t count[8] = {0}, first[8] = {0}, offs[8] = {0};
    uint8_t  sorted[19];
    for (int s = 0; s < 19; ++s) count[cl[s]]++;
    count[0] = 0;
    for (int l = 1, code = 0, off = 0; l <= 7; ++l) {
        code = (code + (int)count[l - 1]) << 1;
        first[l] = (uint32_t)code, offs[l] = (uint32_t)off;
        off += (int)count[l];
    }
    {
        uint32_t fill[8] = {0};
        for (int s = 0; s < 19; ++s)
            if (cl[s]) sorted[offs[cl[s]] + fill[cl[s]]++] = (uint8_t)s;
    }
    const uint32_t nsym = hlit + 257 + hdist + 1;
    uint8_t  lens[320];
    uint32_t i = 0;
    while (i < nsym) {
        uint32_t code = 0, sym = 99;
        for (int l = 1; l <= 7; ++l) {
            if (at >= total) return 7;
            code = code << 1 | bits_at(in, n, at++, 1);
            if (count[l] && code - first[l] < count[l]) { sym = sorted[offs[l] + code - first[l]]; break; }
        }
        if (sym == 99) return 4;
        if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
        uint32_t rep, val = 0;
        if (sym == 16) {
            if (i == 0) return 4;
            val = lens[i - 1];
            rep = 3 + bits_at(in, n, at, 2), at += 2;
        } else if (sym == 17) rep = 3 + bits_at(in, n, at, 3), at += 3;
        else rep = 11 + bits_at(in, n, at, 7), at += 7;
        if (at > total) return 7;
        if (i + rep > nsym) return 4;
        while (rep--) lens[i++] = (uint8_t)val;
    }
    uint32_t k = 0;
    for (uint32_t s = 0; s < hlit + 257; ++s)
        if (lens[s]) k += 32768u >> lens[s];
    if (k != 32768 || lens[256] == 0) return 5;
    uint32_t kd = 0, nz = 0;
    for (uint32_t s = 0; s < hdist + 1; ++s)
        if (lens[hlit + 257 + s]) kd += 32768u >> lens[hlit + 257 + s], nz++;
    if (!(kd == 32768 || nz == 0 || (nz == 1 && kd == 16384))) return 6;
    *header_bits = at - p;
    return 0;
}

/* scan [from, to) bit offsets; stage_hist[8] counts rejections per stage (index 0 = plausible);
 * hits[] receives up to cap plausible offsets.  returns number of plausible offsets */
size_t probe_scan(const uint8_t* in, size_t n, uint64_t from, uint64_t to, uint64_t* stage_hist, uint64_t* hits, size_t cap)
{
    size_t found = 0;
    for (uint64_t p = from; p < to; ++p) {
        uint64_t hb;
        int st = probe_dynamic_header(in, n, p, &hb);
        stage_hist[st]++;
        if (st == 0) {
            if (found < cap) hits[found] = p;
            found++;
        }
    }
    return found;
}
