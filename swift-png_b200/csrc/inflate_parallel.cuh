// inflate_parallel.cuh -- intra-stream parallel DEFLATE inflate: one CTA (1024 threads) per stream.
//
// DEFLATE has no sync markers, but Huffman codes self-synchronise: a decoder started at a wrong
// bit offset falls into step with the true symbol sequence after a few symbols.  Per block the
// CTA parses the header and builds the tables once (shared memory), then eats the block in WAVES
// of 1024 subsequences x 256 bits (32 KiB of compressed data staged in shared memory, padded
// 9/8 so that lane-strided word reads are bank-conflict free):
//
//   1. sync   every thread decodes its subsequence from a guessed start (thread 0's start is
//             exact) and publishes the bit position where it crossed into the next subsequence;
//             threads whose predecessor's exit differs from their start re-decode from it.  The
//             verified prefix grows every round.  The first end-of-block symbol on the verified
//             chain ends the wave (and the block).
//   2. scan   a CTA-wide exclusive scan of per-thread output byte counts gives every thread its
//             output offset.
//   3. emit   every thread decodes its subsequence once more: literals go straight to HBM, and so
//             do LZ77 copies whose source lies before the wave (final already -- the common case
//             in PNG, where the distance is about one scanline).  A copy whose source is inside
//             the wave is DEFERRED: its destination bytes are flagged in an "unresolved" bitmap
//             (1 bit per output byte, shared memory for waves up to 256 KiB of output, HBM above)
//             and the copy goes on a CTA-wide work list.
//   4. resolve the work list is swept by all 1024 threads in rounds; a copy runs as soon as its
//             source bytes carry no unresolved flag, then clears its own flags.  The number of
//             rounds is the depth of the copy->copy dependency chain, not the number of copies.
//
// Anything unusual (invalid symbol on the verified chain, truncation, output overflow, distance
// before the start of the output) is not handled here: warp 0 re-runs the block with the serial
// decoder (inflate_serial.cuh), which owns the exact error semantics of the reference.
//
// CTAs are persistent: each takes streams from an atomic ticket (the host orders streams longest
// first), so per-CTA scratch in HBM is bounded by the number of resident CTAs.
//
// Replaces the reference's serial token loop Stream.readBlock(with:) and InflatorOut.expand
// (Sources/LZ77/Inflator/LZ77.InflatorBuffers.Stream.swift:266-381, LZ77.InflatorOut.swift:124-140).
#pragma once

#include "inflate_serial.cuh"

namespace pngb200 {

constexpr int      PAR_THREADS      = 1024;
constexpr uint32_t PAR_SUB_BITS     = 256;
constexpr uint32_t PAR_SUB_WORDS    = PAR_SUB_BITS / 32;
constexpr uint32_t PAR_WAVE_WORDS   = PAR_THREADS * PAR_SUB_WORDS + 8;
constexpr uint32_t PAR_SMEM_WORDS   = PAR_WAVE_WORDS + PAR_WAVE_WORDS / 8 + 1;
constexpr uint32_t PAR_BITMAP_WORDS = 8192;                        // 256 Ki output bytes per wave in smem
constexpr uint32_t PAR_LIST_CAP     = PAR_THREADS * (PAR_SUB_BITS / 2);  // >= copies per wave (2 bits min each)
constexpr uint64_t PAR_MAX_WAVE_OUT = (uint64_t)PAR_LIST_CAP * 258;

enum : uint32_t { PF_EOB = 1, PF_BAD = 2 };

struct ParHeader {  // block header as parsed by warp 0, broadcast to the CTA
    int32_t  status, type, final, nlit, ndist;
    uint32_t stored;
    uint64_t pos;     // reader position after the header
};

struct ParShared {
    SerialShared ser;
    uint32_t     words[PAR_SMEM_WORDS];
    uint32_t     start_[PAR_THREADS];   // per subsequence: decode start (wave-relative bit)
    uint32_t     exit_[PAR_THREADS];    //                  first symbol boundary past its end
    uint32_t     nout_[PAR_THREADS];    //                  output bytes
    uint8_t      flag_[PAR_THREADS];    //                  PF_EOB / PF_BAD
    uint16_t     list_[PAR_THREADS];    // compacted ids of subsequences that must be re-decoded
    uint32_t     bitmap[PAR_BITMAP_WORDS];
    uint32_t     warp_sums[32];
    uint32_t     npend[2];
    uint32_t     first_need[2], first_stop[2], nlist[2];
    uint32_t     anomaly, ticket;
    ParHeader    hdr;
};

struct ParParams {
    const StreamJob* jobs;
    StreamResult*    results;
    const uint32_t*  order;
    uint32_t*        ticket;       // global work counter (zeroed before launch)
    uint8_t*         scratch;      // per-CTA: two copy lists + unresolved bitmap
    uint64_t         scratch_stride;
    uint64_t         bitmap_words; // size of the HBM bitmap of each CTA
    int              count;
};

struct CopyItem { uint32_t o; uint32_t run_dist; };  // run | dist << 16 ... dist 32768 -> stored as dist-1

struct SmemBits {
    const uint32_t* w;
    uint32_t        wi;
    uint64_t        buf;
    int             cnt;
    uint32_t        pos;
    __device__ __forceinline__ void init(const uint32_t* words, uint32_t start)
    {
        w   = words;
        wi  = start >> 5;
        buf = 0;
        cnt = 0;
        pos = start;
        refill();
        refill();
        int skip = (int)(start & 31);
        buf >>= skip;
        cnt -= skip;
    }
    // one word is always enough: a token takes <= 20 bits before the next refill and <= 28 after
    __device__ __forceinline__ void refill()
    {
        if (cnt <= 32) {
            buf |= (uint64_t)w[wi + (wi >> 3)] << cnt;
            cnt += 32;
            ++wi;
        }
    }
    __device__ __forceinline__ void consume(uint32_t n)
    {
        buf >>= n;
        cnt -= (int)n;
        pos += n;
    }
    __device__ __forceinline__ uint32_t take(uint32_t n)
    {
        uint32_t v = (uint32_t)buf & ((1u << n) - 1u);
        consume(n);
        return v;
    }
};

// sync-phase decode: symbol boundaries and output byte count only
__device__ __forceinline__ void par_decode_count(const ParShared& sh, uint32_t start, uint32_t limit,
                                                 uint32_t& exit_bit, uint32_t& nout, uint32_t& flags)
{
    SmemBits b;
    b.init(sh.words, start);
    nout  = 0;
    flags = 0;
    while (b.pos < limit) {
        b.refill();
        uint32_t e = lookup<LIT_ROOT>(sh.ser.lit, (uint32_t)b.buf);
        uint32_t kind = e_kind(e);
        if (kind == K_LIT) {
            b.consume(e_len(e));
            ++nout;
        } else if (kind == K_BASE) {
            b.consume(e_len(e));
            uint32_t run = e_value(e) + b.take(e_extra(e));
            b.refill();
            uint32_t d = lookup<DIST_ROOT>(sh.ser.dist, (uint32_t)b.buf);
            if (e_kind(d) != K_BASE) { flags = PF_BAD; break; }
            b.consume(e_len(d) + e_extra(d));
            nout += run;
        } else if (kind == K_EOB) {
            b.consume(e_len(e));
            flags = PF_EOB;
            break;
        } else {
            flags = PF_BAD;
            break;
        }
    }
    exit_bit = (flags & PF_BAD) ? max(b.pos, limit) : b.pos;
}

// ---- unresolved-byte bitmap (bit i = output byte i of the wave is not final yet) ----
__device__ __forceinline__ uint32_t bit_mask(uint32_t lo, uint32_t hi)  // bits [lo, hi) of a word, hi <= 32
{
    return (hi >= 32 ? ~0u : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
}
__device__ __forceinline__ void bits_set(uint32_t* U, uint32_t a, uint32_t b)
{
    for (uint32_t w = a >> 5; w <= (b - 1) >> 5; ++w) {
        uint32_t lo = w == (a >> 5) ? (a & 31) : 0, hi = w == ((b - 1) >> 5) ? ((b - 1) & 31) + 1 : 32;
        atomicOr(U + w, bit_mask(lo, hi));
    }
}
__device__ __forceinline__ void bits_clear(uint32_t* U, uint32_t a, uint32_t b)
{
    for (uint32_t w = a >> 5; w <= (b - 1) >> 5; ++w) {
        uint32_t lo = w == (a >> 5) ? (a & 31) : 0, hi = w == ((b - 1) >> 5) ? ((b - 1) & 31) + 1 : 32;
        atomicAnd(U + w, ~bit_mask(lo, hi));
    }
}
__device__ __forceinline__ bool bits_all_clear(const uint32_t* U, uint32_t a, uint32_t b)
{
    const volatile uint32_t* V = U;
    uint32_t any = 0;
    for (uint32_t w = a >> 5; w <= (b - 1) >> 5; ++w) {
        uint32_t lo = w == (a >> 5) ? (a & 31) : 0, hi = w == ((b - 1) >> 5) ? ((b - 1) & 31) + 1 : 32;
        any |= V[w] & bit_mask(lo, hi);
    }
    __threadfence_block();
    return any == 0;
}

// LZ77 copy of `run` bytes to `to` from `dist` bytes back; all needed source bytes are final
__device__ __forceinline__ void lz_copy(uint8_t* to, uint32_t run, uint32_t dist)
{
    const uint8_t* from = to - dist;
    if (dist >= 4) {  // byte k+3 reads k+3-dist < k: four independent loads per step even when overlapping
        uint32_t k = 0;
        for (; k + 4 <= run; k += 4) {
            uint8_t b0 = from[k], b1 = from[k + 1], b2 = from[k + 2], b3 = from[k + 3];
            to[k] = b0; to[k + 1] = b1; to[k + 2] = b2; to[k + 3] = b3;
        }
        for (; k < run; ++k) to[k] = from[k];
    } else {
        uint32_t q = 0;
        for (uint32_t k = 0; k < run; ++k) {
            to[k] = from[q];
            if (++q == dist) q = 0;
        }
    }
}

__global__ void __launch_bounds__(PAR_THREADS, 1) inflate_parallel_kernel(ParParams P)
{
    extern __shared__ __align__(16) unsigned char par_smem[];
    ParShared& sh = *reinterpret_cast<ParShared*>(par_smem);
    const uint32_t t    = threadIdx.x;
    const unsigned lane = lane_id(), warp = t >> 5;
    CopyItem* const  lists[2] = {reinterpret_cast<CopyItem*>(P.scratch + blockIdx.x * P.scratch_stride),
                                 reinterpret_cast<CopyItem*>(P.scratch + blockIdx.x * P.scratch_stride) + PAR_LIST_CAP};
    uint32_t* const  gbitmap  = reinterpret_cast<uint32_t*>(P.scratch + blockIdx.x * P.scratch_stride +
                                                           2 * sizeof(CopyItem) * PAR_LIST_CAP);
    for (uint32_t k = t; k < PAR_BITMAP_WORDS; k += PAR_THREADS) sh.bitmap[k] = 0;

    for (;;) {
        __syncthreads();
        if (t == 0) {
            sh.ticket = atomicAdd(P.ticket, 1u);
            sh.first_need[0] = sh.first_need[1] = PAR_THREADS;
            sh.first_stop[0] = sh.first_stop[1] = PAR_THREADS;
            sh.nlist[0] = sh.nlist[1] = 0;
            sh.anomaly = 0;
        }
        __syncthreads();
        if (sh.ticket >= (uint32_t)P.count) return;
        const int       j   = P.order ? (int)P.order[sh.ticket] : (int)sh.ticket;
        const StreamJob job = P.jobs[j];
        StreamResult*   r   = P.results + j;

        BitReader br;
        br.init(job.src, job.src_len, job.start_bit);
        uint64_t out    = job.start_out;
        uint32_t blocks = 0, waves = 0, sync_rounds = 0, resolve_rounds = 0;
        int      st     = PNGB200_OK;
        uint32_t phase  = (uint32_t)job.phase;
        uint64_t resume_bit = job.start_bit, resume_out = job.start_out;
        uint8_t* const dst = job.dst;
        bool fallback = false;

        if (phase == 0) {
            st = read_stream_header(br, job.format, r);
            if (st == PNGB200_OK) {
                resume_bit = br.at();
                phase = 1;
            }
        }
        if (st == PNGB200_OK && phase == 2) st = read_trailer(br, job.format, r);

        while (st == PNGB200_OK && phase == 1) {
            // warp 0 walks the header bits alone; the CTA then builds the tables together
            __syncthreads();
            if (warp == 0) {
                int      type0 = 0, final0 = 0, nlit0 = 0, ndist0 = 0;
                uint32_t stored0 = 0;
                int st0 = parse_block_header(br, &sh.ser, r, (int)lane, &type0, &final0, &stored0, &nlit0, &ndist0);
                if (lane == 0) sh.hdr = ParHeader{st0, type0, final0, nlit0, ndist0, stored0, br.pos};
            }
            __syncthreads();
            const ParHeader hdr = sh.hdr;
            st = hdr.status;
            if (st != PNGB200_OK) break;
            const int      type = hdr.type, final = hdr.final;
            const uint32_t stored = hdr.stored;
            if (warp != 0) br.seek(hdr.pos);
            if (type != 0) {
                st = build_block_tables(&sh.ser, r, hdr.nlit, hdr.ndist, (int)t, PAR_THREADS);
                if (st != PNGB200_OK) break;
            }
            if (type == 0) {
                if (!br.have(8 * (uint64_t)stored)) { st = PNGB200_NEED_MORE_INPUT; break; }
                if (out + stored > job.dst_cap) { st = fail(r, PNGB200_ERR_OUTPUT_CAPACITY); break; }
                const uint8_t* s = job.src + (br.at() >> 3);
                for (uint32_t k = t; k < stored; k += PAR_THREADS) dst[out + k] = s[k];
                out += stored;
                br.seek(br.pos + 8 * (uint64_t)stored);
                __syncthreads();
            } else {
                bool block_done = false;
                while (!block_done) {
                    ++waves;
                    // ---- stage the wave's bits in shared memory ----
                    const uint64_t wstart = br.pos;                       // absolute bit (reader space)
                    const uint64_t wbase  = (wstart >> 5) & ~(uint64_t)7; // first staged word
                    __syncthreads();
                    for (uint32_t k = t; k < PAR_WAVE_WORDS; k += PAR_THREADS)
                        sh.words[k + (k >> 3)] = br.load_word(wbase + k);
                    if (t == 0) {
                        sh.npend[0] = 0;
                        sh.first_need[0] = sh.first_need[1] = PAR_THREADS;
                        sh.first_stop[0] = sh.first_stop[1] = PAR_THREADS;
                        sh.nlist[0] = sh.nlist[1] = 0;
                    }
                    __syncthreads();
                    const uint32_t rel0  = (uint32_t)(wstart - (wbase << 5));  // < 256
                    const uint32_t limit = (t + 1) * PAR_SUB_BITS;
                    {
                        uint32_t s0 = t == 0 ? rel0 : t * PAR_SUB_BITS, ex0, n0, fl0;
                        par_decode_count(sh, s0, limit, ex0, n0, fl0);
                        sh.start_[t] = s0;
                        sh.exit_[t]  = ex0;
                        sh.nout_[t]  = n0;
                        sh.flag_[t]  = (uint8_t)fl0;
                    }
                    // ---- sync rounds: only subsequences whose start moved are re-decoded, compacted
                    //      onto the lowest threads so that a round costs what it re-decodes ----
                    uint32_t nvalid = PAR_THREADS;
                    bool     stop_found = false;
                    for (uint32_t round = 0;; ++round) {
                        const uint32_t p = round & 1;
                        ++sync_rounds;
                        __syncthreads();
                        if (t == 0) {  // the other parity's counters are idle during this round
                            sh.first_need[p ^ 1] = PAR_THREADS;
                            sh.first_stop[p ^ 1] = PAR_THREADS;
                            sh.nlist[p ^ 1] = 0;
                        }
                        const uint32_t prev_exit = t > 0 ? sh.exit_[t - 1] : 0;
                        const bool need = t > 0 && prev_exit != sh.start_[t];
                        if (need) atomicMin(&sh.first_need[p], t);
                        __syncthreads();
                        const uint32_t fn = sh.first_need[p];
                        if (t < fn && sh.flag_[t] != 0) atomicMin(&sh.first_stop[p], t);
                        __syncthreads();
                        const uint32_t fs = sh.first_stop[p];
                        if (fs < PAR_THREADS) { nvalid = fs + 1; stop_found = true; break; }
                        if (fn == PAR_THREADS) break;
                        const unsigned ballot = __ballot_sync(0xffffffffu, need);
                        if (need) {
                            uint32_t base = 0;
                            const int leader = __ffs(ballot) - 1;
                            if ((int)lane == leader) base = atomicAdd(&sh.nlist[p], (uint32_t)__popc(ballot));
                            base = __shfl_sync(ballot, base, leader);
                            sh.list_[base + __popc(ballot & ((1u << lane) - 1u))] = (uint16_t)t;
                            sh.start_[t] = prev_exit;
                        }
                        __syncthreads();
                        const uint32_t cnt = sh.nlist[p];
                        if (t < cnt) {
                            const uint32_t u = sh.list_[t];
                            uint32_t ex0, n0, fl0;
                            par_decode_count(sh, sh.start_[u], (u + 1) * PAR_SUB_BITS, ex0, n0, fl0);
                            sh.exit_[u] = ex0;
                            sh.nout_[u] = n0;
                            sh.flag_[u] = (uint8_t)fl0;
                        }
                    }
                    __syncthreads();
                    const uint32_t my_start = sh.start_[t], ex = sh.exit_[t], n = sh.nout_[t], fl = sh.flag_[t];
                    // ---- anomalies on the verified chain -> serial decoder ----
                    if (t == nvalid - 1 && ((fl & PF_BAD) || (wbase << 5) + ex > br.total_bits)) sh.anomaly = 1;
                    // ---- scan of output counts ----
                    uint32_t mine = t < nvalid ? n : 0;
                    uint32_t incl = mine;
                    for (int o = 1; o < 32; o <<= 1) {
                        uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                        if ((int)lane >= o) incl += v;
                    }
                    if (lane == 31) sh.warp_sums[warp] = incl;
                    __syncthreads();
                    if (warp == 0) {
                        uint32_t ws = sh.warp_sums[lane], wi = ws;
                        for (int o = 1; o < 32; o <<= 1) {
                            uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
                            if ((int)lane >= o) wi += v;
                        }
                        sh.warp_sums[lane] = wi - ws;  // exclusive
                    }
                    __syncthreads();
                    const uint32_t o_start = sh.warp_sums[warp] + incl - mine;
                    __syncthreads();
                    if (t == PAR_THREADS - 1) sh.warp_sums[0] = o_start + mine;  // wave total
                    __syncthreads();
                    const uint32_t total = sh.warp_sums[0];
                    if (sh.anomaly || out + total > job.dst_cap || total > P.bitmap_words * 32) {
                        fallback = true;
                        break;
                    }
                    // ---- emit: literals + copies from behind the wave; defer in-wave copies ----
                    uint8_t* const  wdst = dst + out;
                    uint32_t* const U    = total <= PAR_BITMAP_WORDS * 32 ? sh.bitmap : gbitmap;
                    if (t < nvalid) {
                        SmemBits b;
                        b.init(sh.words, my_start);
                        uint32_t o = o_start;
                        while (b.pos < limit) {
                            b.refill();
                            uint32_t e = lookup<LIT_ROOT>(sh.ser.lit, (uint32_t)b.buf);
                            uint32_t kind = e_kind(e);
                            if (kind == K_LIT) {
                                b.consume(e_len(e));
                                wdst[o++] = (uint8_t)e_value(e);
                            } else if (kind == K_BASE) {
                                b.consume(e_len(e));
                                uint32_t run = e_value(e) + b.take(e_extra(e));
                                b.refill();
                                uint32_t d = lookup<DIST_ROOT>(sh.ser.dist, (uint32_t)b.buf);
                                b.consume(e_len(d));
                                uint32_t dist = e_value(d) + b.take(e_extra(d));
                                if ((uint64_t)dist > out + o) {  // invalidStringReference
                                    sh.anomaly = 1;
                                    break;
                                }
                                if ((int64_t)o - (int64_t)dist + (int64_t)min(run, dist) <= 0) {
                                    lz_copy(wdst + o, run, dist);  // source entirely behind the wave
                                } else {
                                    bits_set(U, o, o + run);
                                    uint32_t slot = atomicAdd(&sh.npend[0], 1u);
                                    lists[0][slot] = CopyItem{o, run | (dist - 1) << 16};
                                }
                                o += run;
                            } else {
                                break;  // end of block
                            }
                        }
                    }
                    __threadfence_block();
                    __syncthreads();
                    // ---- resolve deferred copies in dependency order ----
                    uint32_t cur = 0;
                    while (!sh.anomaly) {
                        const uint32_t np = sh.npend[cur];
                        if (np == 0) break;
                        ++resolve_rounds;
                        __syncthreads();
                        if (t == 0) sh.npend[cur ^ 1] = 0;
                        __syncthreads();
                        for (uint32_t i = t; i < np; i += PAR_THREADS) {
                            const CopyItem it = lists[cur][i];
                            const uint32_t run = it.run_dist & 0xffff, dist = (it.run_dist >> 16) + 1;
                            const int64_t  src = (int64_t)it.o - (int64_t)dist;
                            const uint32_t a = (uint32_t)max(src, (int64_t)0);
                            const uint32_t c = (uint32_t)(src + (int64_t)min(run, dist));
                            if (bits_all_clear(U, a, c)) {
                                lz_copy(wdst + it.o, run, dist);
                                __threadfence_block();
                                bits_clear(U, it.o, it.o + run);
                            } else {
                                lists[cur ^ 1][atomicAdd(&sh.npend[cur ^ 1], 1u)] = it;
                            }
                        }
                        __threadfence_block();
                        __syncthreads();
                        cur ^= 1;
                    }
                    if (sh.anomaly) {
                        // leave the bitmap clean for whoever uses it next
                        for (uint32_t k = t; k < (total + 31) / 32; k += PAR_THREADS) U[k] = 0;
                        fallback = true;
                        break;
                    }
                    out += total;
                    br.seek((wbase << 5) + sh.exit_[nvalid - 1]);
                    if (stop_found) block_done = true;
                }
                if (fallback) break;
            }
            ++blocks;
            resume_bit = br.at();
            resume_out = out;
            if (final) {
                phase = 2;
                st = read_trailer(br, job.format, r);
                break;
            }
        }
        if (fallback) {
            // the serial decoder redoes this block (and whatever follows) and owns the result record
            __syncthreads();
            if (warp == 0) serial_inflate(sh.ser, job, r, resume_bit, resume_out, 1, blocks);
        } else if (t == 0) {
            if (r->status == 0) r->status = st;
            r->produced      = out;
            r->consumed_bits = br.at();
            r->blocks        = blocks;
            r->resume_bit    = resume_bit;
            r->resume_out    = resume_out;
            r->phase         = phase;
        }
        if (t == 0) {
            r->stat_waves          = waves;
            r->stat_sync_rounds    = sync_rounds;
            r->stat_resolve_rounds = resolve_rounds;
            r->stat_fallback       = fallback ? 1u : 0u;
        }
    }
}

// host side: opt in to the large dynamic shared memory on the current device (once per context)
inline int configure_inflate_parallel()
{
    return (int)cudaFuncSetAttribute(inflate_parallel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(ParShared));
}

inline uint64_t par_bitmap_words(uint64_t max_dst_cap)
{
    uint64_t bytes = max_dst_cap < PAR_MAX_WAVE_OUT ? max_dst_cap : PAR_MAX_WAVE_OUT;
    return (bytes + 31) / 32 + 8;
}
inline uint64_t par_scratch_stride(uint64_t bitmap_words)
{
    uint64_t s = 2 * sizeof(CopyItem) * (uint64_t)PAR_LIST_CAP + 4 * bitmap_words;
    return (s + 255) / 256 * 256;
}

}  // namespace pngb200
