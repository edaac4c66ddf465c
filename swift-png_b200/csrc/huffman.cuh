// huffman.cuh -- DEFLATE Huffman decode tables built cooperatively in shared memory.
//
// Replaces the reference's heap-allocated two-level tables (Sources/LZ77/HuffmanCoding/
// LZ77.HuffmanTree.swift:80-201, Sources/LZ77/Inflator/LZ77.InflatorTables.swift:67-119) with
// root+subtable tables indexed directly by the LSB-first bit buffer (no byte-reversal LUT), each
// entry carrying the RFC 1951 base value and extra-bit count so the token loop needs one lookup.
// Validation rules are the reference's: the literal/length code must be complete
// (HuffmanTree.validate(symbols:lengths:)), the distance code may also have zero symbols or a
// single 1-bit symbol (validate(symbols:normalizing:)).
#pragma once

#include "common.cuh"

namespace pngb200 {

// entry: [3:0] code length, [8:4] bits to skip (code length + extra bits; root + subtable bits for a
// pointer), [9] copy (length / distance base), [10] special, [12:11] special kind (0 end of block,
// 1 subtable pointer, 2 invalid), [31:16] value.  Laid out so that the decode loops test single bits
// and pull the extra-bit field with one BFE (width skip - length, 0 for literals).
enum : uint32_t { K_LIT = 0, K_BASE = 1, K_EOB = 2, K_PTR = 3, K_INVALID = 4 };
enum : uint32_t { E_COPY = 1u << 9, E_SPECIAL = 1u << 10, E_PTR = 1u << 11, E_INVALID = 2u << 11 };
__device__ __forceinline__ uint32_t mk_entry(uint32_t kind, uint32_t len, uint32_t extra, uint32_t value)
{
    uint32_t e = len | (len + extra) << 4 | value << 16;
    if (kind == K_BASE) e |= E_COPY;
    else if (kind == K_EOB) e |= E_SPECIAL;
    else if (kind == K_PTR) e |= E_SPECIAL | E_PTR;
    else if (kind == K_INVALID) e |= E_SPECIAL | E_INVALID;
    return e;
}
__device__ __forceinline__ uint32_t e_len(uint32_t e) { return e & 15u; }
__device__ __forceinline__ uint32_t e_skip(uint32_t e) { return (e >> 4) & 31u; }
__device__ __forceinline__ uint32_t e_extra(uint32_t e) { return e_skip(e) - e_len(e); }
__device__ __forceinline__ uint32_t e_kind(uint32_t e)
{
    return (e & E_SPECIAL) ? (((e >> 11) & 3u) == 0 ? K_EOB : ((e >> 11) & 3u) == 1 ? K_PTR : K_INVALID)
                           : ((e >> 9) & 1u);
}
__device__ __forceinline__ uint32_t e_value(uint32_t e) { return e >> 16; }

constexpr int LIT_ROOT  = 10;
constexpr int DIST_ROOT = 8;
constexpr int META_ROOT = 7;
constexpr int LIT_CAP   = (1 << LIT_ROOT) + 512;   // zlib's ENOUGH bound for (286, 10, 15) is 1332
constexpr int DIST_CAP  = (1 << DIST_ROOT) + 512;
constexpr int META_CAP  = 1 << META_ROOT;

// RFC 1951 section 3.2.5 (the reference tabulates the same numbers in LZ77.Composites.swift:19-111)
__constant__ uint16_t c_len_base[29]  = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27,
                                         31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t  c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2,
                                         2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193,
                                         257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
                                         8193, 12289, 16385, 24577};
__constant__ uint8_t  c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6,
                                          7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t  c_clen_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

enum : int { ALPHA_LITLEN = 0, ALPHA_DIST = 1, ALPHA_META = 2 };

// scratch for one table build
struct HuffScratch {
    uint32_t count[16];
    uint32_t first[16];   // canonical first code of each length
    uint32_t offs[16];    // offset of each length's symbols in sorted[]
    uint32_t running[16];
    uint16_t sorted[320];
    uint32_t sub_alloc;
    int32_t  status;      // 0 ok, else pngb200_status
    int32_t  stub;        // distance code: -2 normal, -1 empty stub, >=0 single-symbol stub
};

__device__ __forceinline__ uint32_t symbol_entry(int alphabet, uint32_t sym, uint32_t len)
{
    if (alphabet == ALPHA_LITLEN) {
        if (sym < 256) return mk_entry(K_LIT, len, 0, sym);
        if (sym == 256) return mk_entry(K_EOB, len, 0, 0);
        if (sym < 286) return mk_entry(K_BASE, len, c_len_extra[sym - 257], c_len_base[sym - 257]);
        return mk_entry(K_INVALID, len, 0, sym);
    } else if (alphabet == ALPHA_DIST) {
        if (sym < 30) return mk_entry(K_BASE, len, c_dist_extra[sym], c_dist_base[sym]);
        return mk_entry(K_INVALID, len, 0, sym);
    }
    return mk_entry(K_LIT, len, 0, sym);
}

// Cooperative build by `nt` threads (tid in [0, nt)); nt == 32 -> warp-synchronous, otherwise the
// whole CTA must call it and CTA barriers are used.  `lens[0..nsym)` are the code lengths.
// On return (after the final barrier) scratch->status tells whether the code was valid.
template <int ROOT, int CAP>
__device__ void build_table(uint32_t* table, const uint8_t* lens, int nsym, int alphabet,
                            HuffScratch* S, int tid, int nt)
{
    auto sync = [&]() {
        if (nt == 32) __syncwarp();
        else __syncthreads();
    };
    if (tid < 16) { S->count[tid] = 0; S->running[tid] = 0; }
    if (tid == 0) { S->sub_alloc = 1u << ROOT; S->status = 0; S->stub = -2; }
    sync();
    for (int s = tid; s < nsym; s += nt) atomicAdd(&S->count[lens[s]], 1u);
    sync();
    if (tid == 0) {
        // HuffmanTree.size: interior-node bookkeeping == Kraft equality for a complete code
        int nonzero = nsym - (int)S->count[0];
        bool stub = false;
        if (alphabet == ALPHA_DIST) {  // validate(symbols:normalizing:)
            if (nonzero == 0) { S->stub = -1; stub = true; }
            else if (nonzero == 1 && S->count[1] == 1) { S->stub = 0; stub = true; }
        }
        long interior = 1;
        uint32_t code = 0, off = 0;
        S->count[0] = 0;
        for (int l = 1; l <= 15; ++l) {
            code = (code + S->count[l - 1]) << 1;
            S->first[l] = code;
            S->offs[l] = off;
            off += S->count[l];
            interior = 2 * interior - (long)S->count[l];
        }
        if (!stub && interior != 0)
            S->status = alphabet == ALPHA_META ? PNGB200_ERR_CODELENGTH_HUFFMAN_TABLE
                                               : PNGB200_ERR_HUFFMAN_TABLE;
    }
    sync();
    if (S->status != 0) return;
    // sorted[]: symbols ordered by (length, symbol) -- one warp, ballot-ranked, chunk by chunk
    if (tid < 32) {
        for (int base = 0; base < nsym; base += 32) {
            int      s = base + tid;
            uint32_t l = s < nsym ? lens[s] : 0;
            unsigned peers = __match_any_sync(0xffffffffu, l);
            if (l) {
                uint32_t rank = __popc(peers & ((1u << tid) - 1u));
                uint32_t r0   = S->running[l];
                S->sorted[S->offs[l] + r0 + rank] = (uint16_t)s;
                __syncwarp(peers);
                if (rank == 0) S->running[l] = r0 + __popc(peers);
            }
            __syncwarp();
        }
    }
    sync();
    // roots: one table index per thread-iteration, canonical decode of its MSB-first prefix.  A prefix
    // shared by longer codes gets a subtable (allocated here, pre-filled with "invalid" for incomplete
    // codes); the long codes themselves are entered in the next step, one symbol per thread, so that no
    // thread fills a whole subtable alone
    for (int i = tid; i < (1 << ROOT); i += nt) {
        uint32_t v = __brev((uint32_t)i) >> (32 - ROOT);
        uint32_t e = 0;
        bool     found = false;
#pragma unroll
        for (int l = 1; l <= ROOT; ++l) {
            uint32_t d = (v >> (ROOT - l)) - S->first[l];
            if (!found && d < S->count[l]) {
                e = symbol_entry(alphabet, S->sorted[S->offs[l] + d], (uint32_t)l);
                found = true;
            }
        }
        if (!found) {
            int maxl = 0;
            for (int l = ROOT + 1; l <= 15; ++l) {
                uint32_t lo = v << (l - ROOT), hi = (v + 1) << (l - ROOT);
                uint32_t a = S->first[l], b = a + S->count[l];
                if (S->count[l] && lo < b && a < hi) maxl = l;
            }
            if (maxl == 0) {
                e = mk_entry(K_INVALID, 0, 0, 0);
            } else {
                int      k    = maxl - ROOT;
                uint32_t slot = atomicAdd(&S->sub_alloc, 1u << k);
                if (slot + (1u << k) > (uint32_t)CAP) {
                    S->status = PNGB200_ERR_INTERNAL;
                    e = mk_entry(K_INVALID, 0, 0, 0);
                } else {
                    e = mk_entry(K_PTR, ROOT, (uint32_t)k, slot);
                    const uint32_t inv = mk_entry(K_INVALID, 0, 0, 0);
                    for (uint32_t j = 0; j < (1u << k); ++j) table[slot + j] = inv;
                }
            }
        }
        table[i] = e;
    }
    sync();
    // long codes (length > ROOT): the symbol at position idx of sorted[] has the canonical code
    // first[l] + idx - offs[l]; its subtable index is the reversed tail of the code, replicated over
    // the subtable's unused high bits
    if (S->status == 0 && ROOT < 15) {
        int nlong_from = (int)S->offs[ROOT] + (int)S->count[ROOT];  // first position with length > ROOT
        int total = (int)S->offs[15] + (int)S->count[15];
        for (int idx = nlong_from + tid; idx < total; idx += nt) {
            const uint32_t sym = S->sorted[idx];
            const uint32_t l   = lens[sym];
            const uint32_t c   = S->first[l] + (uint32_t)idx - S->offs[l];
            const uint32_t t2  = l - ROOT;                                    // tail bits
            const uint32_t v   = c >> t2;                                     // MSB-first root prefix
            const uint32_t e   = table[__brev(v) >> (32 - ROOT)];
            if (e_kind(e) != K_PTR) continue;                                 // (allocation failed)
            const uint32_t k = e_extra(e), slot = e_value(e);
            const uint32_t jl = __brev(c & ((1u << t2) - 1u)) >> (32 - t2);   // LSB-first tail
            const uint32_t se = symbol_entry(alphabet, sym, l);
            for (uint32_t h = 0; h < (1u << (k - t2)); ++h) table[slot + jl + (h << t2)] = se;
        }
    }
    sync();
}

// decode one symbol's table entry from the low bits of `bits` (needs >= 15 valid bits)
template <int ROOT>
__device__ __forceinline__ uint32_t lookup(const uint32_t* table, uint32_t bits)
{
    uint32_t e = table[bits & ((1u << ROOT) - 1u)];
    if (e_kind(e) == K_PTR) e = table[e_value(e) + ((bits >> ROOT) & ((1u << e_extra(e)) - 1u))];
    return e;
}

}  // namespace pngb200
