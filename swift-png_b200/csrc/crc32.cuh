// crc32.cuh -- PNG container work on the device: CRC-32 of every chunk
// of every file in a batch, and the scatter / gather between IDAT chunk bodies and the contiguous
// zlib stream.
//
// Replaces the per-chunk CRC the reference computes on the host while lexing and formatting
// (Sources/PNG/Lexing/PNG.BytestreamSource.swift:62-74, PNG.BytestreamDestination.swift:66-95; CRC-32
// itself comes from swift-hash 0.7.1, module CRC -- the standard reflected 0xEDB88320 polynomial,
// pinned by ErrorHandling.swift:30,42) and the [UInt8] appends that concatenate IDAT payloads.
// CRC-32 is linear over GF(2): crc(A || B) = shift(crc(A), |B|) ^ crc(B), where shift multiplies by
// x^(8|B|).  Every thread takes a 256-byte slice, slices combine inside the CTA, 64 KiB pieces combine
// with one atomicXor per piece into the chunk's accumulator.  HBM-bound in principle (each byte read
// once); in practice bound by the shared-memory table look-up per byte.
#pragma once

#include "common.cuh"

namespace pngb200 {

constexpr uint32_t CRC_PIECE   = 1u << 16;
constexpr uint32_t CRC_THREADS = 256;
constexpr uint32_t CRC_SLICE   = CRC_PIECE / CRC_THREADS;
constexpr uint32_t CRC_TABLE_WORDS = 256 + 32 * 32;  // byte table, then shift operators for 2^k bytes

struct CrcRegion {
    const uint8_t* ptr;
    uint64_t       len;
    uint32_t       prefix;      // big-endian fourcc that virtually precedes ptr (encode: the chunk type)
    uint32_t       has_prefix;
};

struct CrcParams {
    const CrcRegion* regions;
    const uint32_t*  piece_base;  // [count + 1] exclusive prefix of max(1, ceil(len / CRC_PIECE))
    uint32_t*        acc;         // [count], zeroed; receives the CRC-32 of each region
    const uint32_t*  tables;      // CRC_TABLE_WORDS
    uint32_t         count;
    uint32_t         total_pieces;
};

// host side: the byte table and the GF(2) operators "append 2^k zero bytes", k = 0..31
inline void crc_build_tables(uint32_t* t)
{
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        t[i] = c;
    }
    auto times = [](const uint32_t* m, uint32_t v) { uint32_t s = 0; for (int i = 0; v; v >>= 1, ++i) if (v & 1) s ^= m[i]; return s; };
    uint32_t op[32], sq[32];
    op[0] = 0xEDB88320u;
    for (int n = 1; n < 32; ++n) op[n] = 1u << (n - 1);  // one zero bit
    for (int r = 0; r < 3; ++r) {                         // -> one zero byte
        for (int n = 0; n < 32; ++n) sq[n] = times(op, op[n]);
        for (int n = 0; n < 32; ++n) op[n] = sq[n];
    }
    for (int k = 0; k < 32; ++k) {
        for (int n = 0; n < 32; ++n) t[256 + 32 * k + n] = op[n];
        for (int n = 0; n < 32; ++n) sq[n] = times(op, op[n]);
        for (int n = 0; n < 32; ++n) op[n] = sq[n];
    }
}

__device__ __forceinline__ uint32_t crc_times(const uint32_t* m, uint32_t v)
{
    uint32_t s = 0;
    while (v) {
        const int i = __ffs(v) - 1;
        s ^= m[i];
        v &= v - 1;
    }
    return s;
}
// crc of (M || n zero-effect bytes): multiply by x^(8n)
__device__ __forceinline__ uint32_t crc_shift(const uint32_t* ops, uint32_t crc, uint64_t n)
{
    for (int k = 0; n && crc; n >>= 1, ++k)
        if (n & 1) crc = crc_times(ops + 32 * k, crc);
    return crc;
}

// CRC-32 of `L` <= CRC_PIECE bytes at `ptr`, computed by a whole CTA of CRC_THREADS threads; valid on
// thread 0.  `table` (256 words) and `ops` (16 x 32 words: shift operators for 2^0 .. 2^15 bytes) live in
// shared memory, `red` is CRC_THREADS / 32 words of shared scratch.
__device__ __forceinline__ uint32_t crc_piece_cta(const uint8_t* ptr, uint32_t L, const uint32_t* table,
                                                  const uint32_t* ops, uint32_t* red)
{
    const uint32_t begin = min(L, threadIdx.x * CRC_SLICE), end = min(L, begin + CRC_SLICE);
    uint32_t crc = 0;
    if (end > begin) {
        const uint8_t* s = ptr + begin;
        uint32_t n = end - begin, c = 0xffffffffu;
        while (n && (((uintptr_t)s) & 15)) { c = table[(c ^ *s++) & 0xff] ^ (c >> 8); --n; }
        for (; n >= 16; n -= 16, s += 16) {
            const uint4 v = *(const uint4*)s;
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                c ^= w[q];  // slicing by one over a word: four dependent look-ups
#pragma unroll
                for (int b = 0; b < 4; ++b) c = table[c & 0xff] ^ (c >> 8);
            }
        }
        while (n) { c = table[(c ^ *s++) & 0xff] ^ (c >> 8); --n; }
        crc = crc_shift(ops, ~c, L - end);  // bytes of this piece that follow the slice
    }
    for (int o = 16; o; o >>= 1) crc ^= __shfl_xor_sync(0xffffffffu, crc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = crc;
    __syncthreads();
    crc = 0;
    if (threadIdx.x == 0)
        for (uint32_t w = 0; w < CRC_THREADS / 32; ++w) crc ^= red[w];
    return crc;
}

__global__ void __launch_bounds__(CRC_THREADS) crc_regions_kernel(CrcParams p)
{
    __shared__ uint32_t table[256];
    __shared__ uint32_t ops[16 * 32];
    __shared__ uint32_t red[CRC_THREADS / 32];
    const uint32_t piece = blockIdx.x;
    if (piece >= p.total_pieces) return;
    table[threadIdx.x] = p.tables[threadIdx.x];
    ops[threadIdx.x] = p.tables[256 + threadIdx.x];
    ops[threadIdx.x + 256] = p.tables[512 + threadIdx.x];
    uint32_t lo = 0, hi = p.count;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (p.piece_base[mid] <= piece) lo = mid;
        else hi = mid;
    }
    const CrcRegion reg = p.regions[lo];
    const uint64_t  off = (uint64_t)(piece - p.piece_base[lo]) * CRC_PIECE;
    const uint32_t  L   = off < reg.len ? (uint32_t)min((uint64_t)CRC_PIECE, reg.len - off) : 0;
    __syncthreads();
    uint32_t crc = crc_piece_cta(reg.ptr + off, L, table, ops, red);
    if (threadIdx.x == 0) {
        const uint32_t* all = p.tables + 256;
        crc = crc_shift(all, crc, reg.len - off - L);  // the rest of the region
        if (off == 0 && reg.has_prefix) {
            uint32_t c = 0xffffffffu;
            for (int b = 3; b >= 0; --b) c = table[(c ^ (reg.prefix >> (8 * b))) & 0xff] ^ (c >> 8);
            crc ^= crc_shift(all, ~c, reg.len);
        }
        if (crc) atomicXor(p.acc + lo, crc);
    }
}

// ---- IDAT scatter / gather: byte-granular segment copies with word-wide, coalesced traffic ----
struct CopySegment {
    const uint8_t* src;
    uint8_t*       dst;
    uint64_t       len;
};
struct CopyParams {
    const CopySegment* segments;
    const uint32_t*    piece_base;  // [count + 1], pieces of CRC_PIECE bytes
    uint32_t           count;
    uint32_t           total_pieces;
};

__global__ void __launch_bounds__(CRC_THREADS) segment_copy_kernel(CopyParams p)
{
    const uint32_t piece = blockIdx.x;
    if (piece >= p.total_pieces) return;
    uint32_t lo = 0, hi = p.count;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (p.piece_base[mid] <= piece) lo = mid;
        else hi = mid;
    }
    const CopySegment seg = p.segments[lo];
    const uint64_t    off = (uint64_t)(piece - p.piece_base[lo]) * CRC_PIECE;
    if (off >= seg.len) return;
    const uint32_t L = (uint32_t)min((uint64_t)CRC_PIECE, seg.len - off);
    const uint8_t* src = seg.src + off;
    uint8_t*       dst = seg.dst + off;
    const uint32_t head = min(L, (uint32_t)((4 - (((uintptr_t)dst) & 3)) & 3));
    if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
    const uint32_t words = (L - head) >> 2;
    const uint8_t* s0 = src + head;
    const uint32_t mis = (uint32_t)(((uintptr_t)s0) & 3);
    const uint32_t* sw = (const uint32_t*)(s0 - mis);  // the aligned word holding s0's first byte
    uint32_t*       dw = (uint32_t*)(dst + head);
    if (mis == 0) {
        for (uint32_t w = threadIdx.x; w < words; w += CRC_THREADS) dw[w] = sw[w];
    } else {
        // bytes beyond the segment's end inside the last aligned word are read but never stored;
        // every arena these pointers come from is padded past its last byte
        for (uint32_t w = threadIdx.x; w < words; w += CRC_THREADS) dw[w] = __funnelshift_r(sw[w], sw[w + 1], 8 * mis);
    }
    const uint32_t done = head + 4 * words;
    if (threadIdx.x < L - done) dst[done + threadIdx.x] = src[done + threadIdx.x];
}

// encode: write each chunk's length + type in front of its body and the finished CRC behind it
struct FrameItem {
    uint8_t* chunk;   // first byte of the chunk (its length field)
    uint32_t len;     // body bytes
    uint32_t type;    // big-endian fourcc
};
__global__ void frame_chunks_kernel(const FrameItem* items, const uint32_t* crc, uint32_t count, int write_crc)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const FrameItem it = items[i];
    if (!write_crc) {
        for (int b = 0; b < 4; ++b) {
            it.chunk[b] = (uint8_t)(it.len >> (24 - 8 * b));
            it.chunk[4 + b] = (uint8_t)(it.type >> (24 - 8 * b));
        }
    } else {
        for (int b = 0; b < 4; ++b) it.chunk[8 + (size_t)it.len + b] = (uint8_t)(crc[i] >> (24 - 8 * b));
    }
}

}  // namespace pngb200
