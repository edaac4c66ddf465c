// block_search.cuh -- speculative DEFLATE block-boundary search (the CPU model in
// tools/block_probe.c is the validated reference for every check below).
//
// "Could a dynamic-Huffman block header start at bit p?" is a pure function of p, so a CTA tests
// CTA-width consecutive bit offsets per step and reports the first plausible one at or after each
// requested split point.  99.9 % of the offsets die within 74 bits (BTYPE 75 %, HLIT/HDIST 3 %, Kraft
// sum of the code-length code 22 %); the survivors decode the code-length sequence and check that the
// literal/length code is complete with an end-of-block symbol and the distance code is complete, empty
// or a single one-bit code (the reference's HuffmanTree.validate rules,
// Sources/LZ77/HuffmanCoding/LZ77.HuffmanTree.swift:80-201).
#pragma once

#include "common.cuh"

namespace pngb200 {

// 57 or more bits of the stream starting at bit `bit` (zero beyond the end): three aligned words when they lie
// inside the buffer, bytes at its edges
__device__ __forceinline__ uint64_t bs_bits64(const uint8_t* in, uint64_t nbytes, uint64_t bit)
{
    const uint64_t  byte = bit >> 3;
    const uintptr_t a = (uintptr_t)in + byte, a4 = a & ~(uintptr_t)3;
    if (a4 >= (uintptr_t)in && a4 + 12 <= (uintptr_t)in + nbytes) {
        const uint32_t* w = (const uint32_t*)a4;
        const uint32_t  w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
        const uint32_t  sh = (uint32_t)(a & 3) * 8 + (uint32_t)(bit & 7);   // < 32
        const uint32_t  lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
        return (uint64_t)hi << 32 | lo;
    }
    uint64_t v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (byte + k < nbytes) v |= (uint64_t)in[byte + k] << (8 * k);
    return v >> (bit & 7);
}
__device__ __forceinline__ uint32_t bs_bits(const uint8_t* in, uint64_t nbytes, uint64_t bit, int count)
{
    return (uint32_t)(bs_bits64(in, nbytes, bit) & ((1ull << count) - 1));
}

// 0 = plausible dynamic-block header at bit p; otherwise the stage that rejected it (as in block_probe.c)
__device__ int bs_probe_dynamic_header(const uint8_t* in, uint64_t n, uint64_t p)
{
    const uint64_t total = n * 8;
    if (p + 17 > total) return 7;
    const uint64_t head = bs_bits64(in, n, p);                    // BFINAL, BTYPE, HLIT, HDIST, HCLEN: 17 bits
    if (((head >> 1) & 3) != 2) return 1;
    const uint32_t hlit = (uint32_t)(head >> 3) & 31, hdist = (uint32_t)(head >> 8) & 31, hclen = (uint32_t)(head >> 13) & 15;
    if (hlit > 29 || hdist > 29) return 2;
    uint64_t at = p + 17;
    if (at + 3 * (uint64_t)(hclen + 4) > total) return 7;
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t  cl[19];
#pragma unroll
    for (int i = 0; i < 19; ++i) cl[i] = 0;
    uint32_t kraft = 0;
    {
        const uint64_t v = bs_bits64(in, n, at);                  // up to 19 x 3 = 57 bits in one fetch
        for (uint32_t i = 0; i < hclen + 4; ++i) {
            const uint32_t l = (uint32_t)(v >> (3 * i)) & 7;
            cl[order[i]] = (uint8_t)l;
            if (l) kraft += 128u >> l;
        }
        at += 3 * (uint64_t)(hclen + 4);
    }
    if (kraft != 128) return 3;
    uint32_t count[8], first[8], offs[8];
    uint8_t  sorted[19];
    for (int l = 0; l < 8; ++l) count[l] = 0;
    for (int s = 0; s < 19; ++s) count[cl[s]]++;
    count[0] = 0;
    {
        uint32_t code = 0, off = 0;
        first[0] = offs[0] = 0;
        for (int l = 1; l <= 7; ++l) {
            code = (code + count[l - 1]) << 1;
            first[l] = code, offs[l] = off;
            off += count[l];
        }
        uint32_t fill[8];
        for (int l = 0; l < 8; ++l) fill[l] = 0;
        for (int s = 0; s < 19; ++s)
            if (cl[s]) sorted[offs[cl[s]] + fill[cl[s]]++] = (uint8_t)s;
    }
    const uint32_t nsym = hlit + 257 + hdist + 1;
    // Kraft sums are accumulated on the fly: no per-symbol array is needed (registers, not local memory)
    uint32_t i = 0, prev = 0, klit = 0, kdist = 0, ndist = 0, eob = 0;
    auto account = [&](uint32_t index, uint32_t len) {
        if (index < hlit + 257) {
            if (len) klit += 32768u >> len;
            if (index == 256) eob = len;
        } else if (len) {
            kdist += 32768u >> len;
            ++ndist;
        }
    };
    while (i < nsym) {
        uint32_t code = 0, sym = 99;
        for (int l = 1; l <= 7; ++l) {
            if (at >= total) return 7;
            code = code << 1 | bs_bits(in, n, at++, 1);
            if (count[l] && code - first[l] < count[l]) { sym = sorted[offs[l] + code - first[l]]; break; }
        }
        if (sym == 99) return 4;
        if (sym < 16) {
            account(i++, sym);
            prev = sym;
            continue;
        }
        uint32_t rep, val = 0;
        if (sym == 16) {
            if (i == 0) return 4;
            val = prev;
            rep = 3 + bs_bits(in, n, at, 2), at += 2;
        } else if (sym == 17) rep = 3 + bs_bits(in, n, at, 3), at += 3;
        else rep = 11 + bs_bits(in, n, at, 7), at += 7;
        if (at > total) return 7;
        if (i + rep > nsym) return 4;
        while (rep--) account(i++, val);
        prev = val;
    }
    if (klit != 32768 || eob == 0) return 5;
    if (!(kdist == 32768 || ndist == 0 || (ndist == 1 && kdist == 16384))) return 6;
    return 0;
}

struct SearchJob {
    const uint8_t* src;      // compressed stream
    uint64_t       src_len;  // bytes
    uint64_t       from_bit; // search starts here
    uint64_t       limit_bit;// and gives up here (exclusive)
    uint64_t       found;    // out: first plausible offset, or ~0 if none before limit_bit
};

// BS_CTAS CTAs per job (blockIdx.y), each testing every BS_CTAS-th chunk of 256 consecutive offsets; the
// smallest hit wins (atomicMin on job.found, which the host presets to ~0) and a CTA stops as soon as its next
// chunk lies behind the best hit so far
constexpr uint32_t BS_CTAS = 8;
__global__ void __launch_bounds__(256) block_search_kernel(SearchJob* jobs, uint32_t count)
{
    if (blockIdx.x >= count) return;
    SearchJob& job = jobs[blockIdx.x];
    unsigned long long* const found = reinterpret_cast<unsigned long long*>(&job.found);
    __shared__ unsigned long long best;
    for (uint64_t base = job.from_bit + (uint64_t)blockIdx.y * blockDim.x; base < job.limit_bit; base += (uint64_t)gridDim.y * blockDim.x) {
        if (threadIdx.x == 0) best = *reinterpret_cast<volatile unsigned long long*>(found);
        __syncthreads();
        if (best < base) break;
        const uint64_t p = base + threadIdx.x;
        if (p < job.limit_bit && bs_probe_dynamic_header(job.src, job.src_len, p) == 0) atomicMin(found, (unsigned long long)p);
        __syncthreads();
    }
}

}  // namespace pngb200
