// checksum.cuh -- Adler-32 / CRC-32 of the inflated stream, computed after the fact in parallel.
//
// The reference folds the checksum into the output window as bytes are evicted
// (LZ77.InflatorOut.shift, Sources/LZ77/Inflator/LZ77.InflatorOut.swift:153-207; LZ77.MRC32,
// Sources/LZ77/Wrappers/LZ77.MRC32.swift:26-47; CRC-32 via swift-hash for gzip,
// Sources/LZ77/Gzip/Gzip.Format.Integral.swift:6-32).  Here the sums are reassociated:
//   s1 = 1 + sum b_i,   s2 = n + sum (n - i) b_i   (mod 65521),
// so 64 KiB chunks are reduced independently (dp4a for the weighted term) and one warp per stream
// folds the chunk partials and compares against the trailer the inflate kernel recorded.
#pragma once

#include "common.cuh"
#include "crc32.cuh"

namespace pngb200 {

constexpr uint32_t CK_CHUNK   = 1u << 16;
constexpr uint32_t CK_THREADS = 256;
constexpr uint32_t ADLER_MOD  = 65521;
static_assert(CK_CHUNK == CRC_PIECE && CK_THREADS == CRC_THREADS, "the gzip branch reuses crc_piece_cta");

struct ChecksumParams {
    const StreamJob* jobs;
    StreamResult*    results;
    const uint32_t*  chunk_base;  // [count + 1] exclusive prefix of ceil(dst_cap / CK_CHUNK)
    uint64_t*        partial;     // [total_chunks][2]: sum b, sum (L - k) b  (or crc, length)
    uint32_t         count;
    uint32_t         total_chunks;
    const uint32_t*  crc_tables;  // CRC_TABLE_WORDS (crc32.cuh); needed when the batch holds gzip streams
};

__device__ __forceinline__ uint32_t crc32_byte_table(uint32_t i)
{
    uint32_t c = i;
#pragma unroll
    for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    return c;
}

__global__ void __launch_bounds__(CK_THREADS) checksum_chunk_kernel(ChecksumParams p)
{
    const uint32_t chunk = blockIdx.x;
    if (chunk >= p.total_chunks) return;
    uint32_t lo = 0, hi = p.count;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (p.chunk_base[mid] <= chunk) lo = mid;
        else hi = mid;
    }
    if (p.results[lo].ck_done) return;  // zlib stream inflated by the wave kernel: Adler-32 came with the store
    const StreamJob job = p.jobs[lo];
    const uint64_t  n   = p.results[lo].produced;
    const uint64_t  off = (uint64_t)(chunk - p.chunk_base[lo]) * CK_CHUNK;
    __shared__ uint64_t red[2][CK_THREADS / 32];
    uint64_t A = 0, B = 0;
    const uint32_t L = off < n ? (uint32_t)min((uint64_t)CK_CHUNK, n - off) : 0;
    const uint8_t* src = job.dst + off;
    if (job.format == PNGB200_FORMAT_GZIP) {
        // CRC-32 of the chunk by the whole CTA (256-byte slices combined with GF(2) shift operators, the
        // same device code as crc_regions_kernel); the fold below combines the chunk CRCs
        __shared__ uint32_t table[256];
        __shared__ uint32_t ops[16 * 32];
        __shared__ uint32_t cred[CRC_THREADS / 32];
        table[threadIdx.x] = p.crc_tables[threadIdx.x];
        ops[threadIdx.x] = p.crc_tables[256 + threadIdx.x];
        ops[threadIdx.x + 256] = p.crc_tables[512 + threadIdx.x];
        __syncthreads();
        const uint32_t crc = crc_piece_cta(src, L, table, ops, cred);
        if (threadIdx.x == 0) {
            p.partial[2 * (uint64_t)chunk]     = crc;
            p.partial[2 * (uint64_t)chunk + 1] = L;
        }
        return;
    }
    if ((((uintptr_t)src) & 15) == 0) {
        const uint4* q = (const uint4*)src;
        for (uint32_t v = threadIdx.x; v * 16 + 16 <= L; v += CK_THREADS) {
            uint4    x = q[v];
            uint32_t k = v * 16;
            uint32_t s = __vsadu4(x.x, 0) + __vsadu4(x.y, 0) + __vsadu4(x.z, 0) + __vsadu4(x.w, 0);
            // sum (L - k - i) b_i = (L - k) * s - sum i * b_i
            uint32_t wsum = __dp4a(x.x, 0x03020100u, 0u) + __dp4a(x.y, 0x07060504u, 0u) +
                            __dp4a(x.z, 0x0b0a0908u, 0u) + __dp4a(x.w, 0x0f0e0d0cu, 0u);
            A += s;
            B += (uint64_t)(L - k) * s - wsum;
        }
        for (uint32_t k = (L & ~15u) + threadIdx.x; k < L; k += CK_THREADS) {
            A += src[k];
            B += (uint64_t)(L - k) * src[k];
        }
    } else {
        for (uint32_t k = threadIdx.x; k < L; k += CK_THREADS) {
            A += src[k];
            B += (uint64_t)(L - k) * src[k];
        }
    }
    for (int o = 16; o; o >>= 1) {
        A += __shfl_down_sync(0xffffffffu, A, o);
        B += __shfl_down_sync(0xffffffffu, B, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = A;
        red[1][threadIdx.x >> 5] = B;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        A = B = 0;
        for (uint32_t w = 0; w < CK_THREADS / 32; ++w) {
            A += red[0][w];
            B += red[1][w];
        }
        p.partial[2 * (uint64_t)chunk]     = A;
        p.partial[2 * (uint64_t)chunk + 1] = B;
    }
}

// ---- CRC-32 combine (GF(2) 32x32 operator for "append len zero bytes"), zlib's construction ----
__device__ inline uint32_t gf2_times(const uint32_t* mat, uint32_t vec)
{
    uint32_t sum = 0;
    for (int i = 0; vec; vec >>= 1, ++i)
        if (vec & 1) sum ^= mat[i];
    return sum;
}
__device__ inline void gf2_square(uint32_t* sq, const uint32_t* mat)
{
    for (int n = 0; n < 32; ++n) sq[n] = gf2_times(mat, mat[n]);
}
__device__ inline uint32_t crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2)
{
    if (len2 == 0) return crc1;
    uint32_t even[32], odd[32];
    odd[0] = 0xEDB88320u;
    uint32_t row = 1;
    for (int n = 1; n < 32; ++n) { odd[n] = row; row <<= 1; }
    gf2_square(even, odd);
    gf2_square(odd, even);
    do {
        gf2_square(even, odd);
        if (len2 & 1) crc1 = gf2_times(even, crc1);
        len2 >>= 1;
        if (len2 == 0) break;
        gf2_square(odd, even);
        if (len2 & 1) crc1 = gf2_times(odd, crc1);
        len2 >>= 1;
    } while (len2 != 0);
    return crc1 ^ crc2;
}

// one warp per stream: fold chunk partials, compare with the declared trailer
__global__ void __launch_bounds__(32) checksum_fold_kernel(ChecksumParams p)
{
    const uint32_t j = blockIdx.x;
    if (j >= p.count) return;
    const unsigned  lane = lane_id();
    StreamResult*   r    = p.results + j;
    if (r->ck_done) return;
    const StreamJob job  = p.jobs[j];
    const uint64_t  n    = r->produced;
    const uint32_t  c0   = p.chunk_base[j];
    const uint32_t  nch  = (uint32_t)((n + CK_CHUNK - 1) / CK_CHUNK);
    uint32_t computed;
    if (job.format == PNGB200_FORMAT_GZIP) {
        // crc(A || B) = shift(crc(A), |B|) ^ crc(B).  All chunks but the last are CK_CHUNK long, so
        // the shift operator for CK_CHUNK bytes is built once (squarings) and applied per chunk
        computed = 0;
        if (lane == 0) {
            uint32_t op[32], tmp[32];  // GF(2) operator for "append CK_CHUNK zero bytes"
            op[0] = 0xEDB88320u;
            for (int n = 1; n < 32; ++n) op[n] = 1u << (n - 1);       // one zero BIT
            for (int k = 0; k < 3; ++k) { gf2_square(tmp, op); for (int n = 0; n < 32; ++n) op[n] = tmp[n]; }  // one byte
            for (uint32_t len = 1; len < CK_CHUNK; len <<= 1) { gf2_square(tmp, op); for (int n = 0; n < 32; ++n) op[n] = tmp[n]; }
            for (uint32_t c = 0; c < nch; ++c) {
                const uint32_t crc2 = (uint32_t)p.partial[2 * (uint64_t)(c0 + c)];
                const uint64_t len2 = p.partial[2 * (uint64_t)(c0 + c) + 1];
                computed = len2 == CK_CHUNK ? (gf2_times(op, computed) ^ crc2) : crc32_combine(computed, crc2, len2);
            }
        }
        computed = __shfl_sync(0xffffffffu, computed, 0);
    } else {
        uint64_t s1 = 0, s2 = 0;
        for (uint32_t c = lane; c < nch; c += 32) {
            uint64_t A = p.partial[2 * (uint64_t)(c0 + c)], B = p.partial[2 * (uint64_t)(c0 + c) + 1];
            uint64_t off = (uint64_t)c * CK_CHUNK;
            uint64_t L   = min((uint64_t)CK_CHUNK, n - off);
            uint64_t after = (n - off - L) % ADLER_MOD;  // bytes that follow this chunk
            s1 = (s1 + A) % ADLER_MOD;
            s2 = (s2 + B % ADLER_MOD + after * (A % ADLER_MOD)) % ADLER_MOD;
        }
        for (int o = 16; o; o >>= 1) {
            s1 += __shfl_down_sync(0xffffffffu, s1, o);
            s2 += __shfl_down_sync(0xffffffffu, s2, o);
        }
        s1 = (s1 + 1) % ADLER_MOD;
        s2 = (s2 + n % ADLER_MOD) % ADLER_MOD;
        computed = (uint32_t)(s2 << 16 | s1);
        computed = __shfl_sync(0xffffffffu, computed, 0);
    }
    if (lane == 0) {
        r->checksum = computed;
        if (r->trailer_seen && job.format != PNGB200_FORMAT_IOS && r->status >= 0 &&
            r->declared != computed) {
            r->status = PNGB200_ERR_STREAM_CHECKSUM;
            r->err_a  = r->declared;
            r->err_b  = computed;
        }
    }
}

}  // namespace pngb200
