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
//             verified prefix grows every round; typically 2-3 rounds settle all 1024 threads.
//             The first end-of-block symbol on the verified chain ends the wave (and the block).
//   2. scan   a CTA-wide exclusive scan of per-thread output byte counts gives every thread its
//             output offset.
//   3. emit   every thread decodes its subsequence once more and writes literals and LZ77 copies
//             straight to HBM.  A copy whose source lies in bytes another thread of the wave has
//             not produced yet blocks that thread until the producer's published "resolved up to"
//             cursor covers the source (rounds separated by CTA barriers; copies that reach behind
//             the wave -- the common case in PNG, where the distance is about one scanline --
//             never block).
//
// Anything unusual (invalid symbol on the verified chain, truncation, output overflow, distance
// before the start of the output) is not handled here: warp 0 re-runs the block with the serial
// decoder (inflate_serial.cuh), which owns the exact error semantics of the reference.
//
// Replaces the reference's serial token loop Stream.readBlock(with:) and InflatorOut.expand
// (Sources/LZ77/Inflator/LZ77.InflatorBuffers.Stream.swift:266-381, LZ77.InflatorOut.swift:124-140).
#pragma once

#include "inflate_serial.cuh"

namespace pngb200 {

constexpr int      PAR_THREADS    = 1024;
constexpr uint32_t PAR_SUB_BITS   = 256;
constexpr uint32_t PAR_SUB_WORDS  = PAR_SUB_BITS / 32;
constexpr uint32_t PAR_WAVE_WORDS = PAR_THREADS * PAR_SUB_WORDS + 8;
constexpr uint32_t PAR_SMEM_WORDS = PAR_WAVE_WORDS + PAR_WAVE_WORDS / 8 + 1;

enum : uint32_t { PF_EOB = 1, PF_BAD = 2 };

struct ParShared {
    SerialShared ser;
    uint32_t     words[PAR_SMEM_WORDS];
    uint32_t     exit_[PAR_THREADS];
    uint32_t     ostart[PAR_THREADS + 1];
    uint32_t     resolved[PAR_THREADS];
    uint32_t     warp_sums[32];
    uint32_t     first_need, first_stop, anomaly, pad;
};

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
        int skip = (int)(start & 31);
        buf >>= skip;
        cnt -= skip;
    }
    __device__ __forceinline__ void refill()
    {
        while (cnt <= 32) {
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

// are wave-relative output bytes [a, b) final?  (a < b <= my own range start)
__device__ __forceinline__ bool par_resolved(const ParShared& sh, uint32_t a, uint32_t b, uint32_t me)
{
    // owner of byte a: last thread u < me with ostart[u] <= a
    uint32_t lo = 0, hi = me;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (sh.ostart[mid] <= a) lo = mid;
        else hi = mid;
    }
    const volatile uint32_t* R = sh.resolved;
    bool ok = true;
    for (uint32_t u = lo; u < me; ++u) {
        uint32_t end = sh.ostart[u + 1];
        if (R[u] < min(b, end)) { ok = false; break; }
        if (end >= b) break;
    }
    __threadfence_block();
    return ok;
}

__global__ void __launch_bounds__(PAR_THREADS, 1)
inflate_parallel_kernel(const StreamJob* jobs, StreamResult* results, const uint32_t* order, int count)
{
    extern __shared__ __align__(16) unsigned char par_smem[];
    ParShared& sh = *reinterpret_cast<ParShared*>(par_smem);
    if ((int)blockIdx.x >= count) return;
    const int       j   = order ? (int)order[blockIdx.x] : (int)blockIdx.x;
    const StreamJob job = jobs[j];
    StreamResult*   r   = results + j;
    const uint32_t  t   = threadIdx.x;
    const unsigned  lane = lane_id(), warp = t >> 5;

    BitReader br;
    br.init(job.src, job.src_len, job.start_bit);
    uint64_t out    = job.start_out;
    uint32_t blocks = 0;
    int      st     = PNGB200_OK;
    uint32_t phase  = (uint32_t)job.phase;
    uint64_t resume_bit = job.start_bit, resume_out = job.start_out;
    uint8_t* const dst = job.dst;
    bool fallback = false;

    if (t == 0) { sh.first_need = PAR_THREADS; sh.first_stop = PAR_THREADS; sh.anomaly = 0; }
    __syncthreads();

    if (phase == 0) {
        st = read_stream_header(br, job.format, r);
        if (st == PNGB200_OK) {
            resume_bit = br.at();
            phase = 1;
        }
    }
    if (st == PNGB200_OK && phase == 2) st = read_trailer(br, job.format, r);

    while (st == PNGB200_OK && phase == 1) {
        int      type, final;
        uint32_t stored = 0;
        st = read_block_header(br, &sh.ser, r, (int)t, PAR_THREADS, &type, &final, &stored);
        if (st != PNGB200_OK) break;
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
                // ---- stage the wave's bits in shared memory ----
                const uint64_t wstart = br.pos;                       // absolute bit (reader space)
                const uint64_t wbase  = (wstart >> 5) & ~(uint64_t)7; // first staged word
                __syncthreads();
                for (uint32_t k = t; k < PAR_WAVE_WORDS; k += PAR_THREADS)
                    sh.words[k + (k >> 3)] = br.load_word(wbase + k);
                __syncthreads();
                const uint32_t rel0  = (uint32_t)(wstart - (wbase << 5));
                const uint32_t limit = (t + 1) * PAR_SUB_BITS;
                uint32_t my_start = t == 0 ? rel0 : t * PAR_SUB_BITS;
                if (t == 0 && rel0 >= limit) my_start = rel0;  // (rel0 < 256 always)
                uint32_t ex, n, fl;
                par_decode_count(sh, my_start, limit, ex, n, fl);
                // ---- sync rounds ----
                uint32_t nvalid = PAR_THREADS;
                bool     stop_found = false;
                for (;;) {
                    sh.exit_[t] = ex;
                    __syncthreads();
                    const bool need = t > 0 && sh.exit_[t - 1] != my_start;
                    if (need) atomicMin(&sh.first_need, t);
                    __syncthreads();
                    const uint32_t fn = sh.first_need;
                    if (t < fn && fl != 0) atomicMin(&sh.first_stop, t);
                    __syncthreads();
                    const uint32_t fs = sh.first_stop;
                    __syncthreads();
                    if (t == 0) { sh.first_need = PAR_THREADS; sh.first_stop = PAR_THREADS; }
                    if (fs < PAR_THREADS) { nvalid = fs + 1; stop_found = true; break; }
                    if (fn == PAR_THREADS) break;
                    if (need) {
                        my_start = sh.exit_[t - 1];
                        par_decode_count(sh, my_start, limit, ex, n, fl);
                    }
                }
                // ---- anomalies on the verified chain -> serial decoder ----
                // (uniform decisions: every thread reads the same shared values)
                bool bad = false;
                if (t == nvalid - 1) {
                    if ((fl & PF_BAD) || wbase * 32 + ex > br.total_bits) bad = true;
                    if (bad) sh.anomaly = 1;
                }
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
                sh.ostart[t] = o_start;
                if (t == PAR_THREADS - 1) sh.ostart[PAR_THREADS] = o_start + mine;
                sh.resolved[t] = o_start;
                __syncthreads();
                const uint32_t total = sh.ostart[PAR_THREADS];
                if (sh.anomaly || out + total > job.dst_cap) { fallback = true; break; }
                // ---- emit ----
                uint8_t* const wdst = dst + out;
                SmemBits b;
                b.init(sh.words, my_start);
                uint32_t o    = o_start;
                bool     done = t >= nvalid;
                for (;;) {
                    bool blocked = false;
                    while (!done) {
                        if (b.pos >= limit) { done = true; break; }
                        b.refill();
                        uint32_t e = lookup<LIT_ROOT>(sh.ser.lit, (uint32_t)b.buf);
                        uint32_t kind = e_kind(e);
                        if (kind == K_LIT) {
                            b.consume(e_len(e));
                            wdst[o++] = (uint8_t)e_value(e);
                        } else if (kind == K_BASE) {
                            SmemBits save = b;
                            b.consume(e_len(e));
                            uint32_t run = e_value(e) + b.take(e_extra(e));
                            b.refill();
                            uint32_t d = lookup<DIST_ROOT>(sh.ser.dist, (uint32_t)b.buf);
                            b.consume(e_len(d));
                            uint32_t dist = e_value(d) + b.take(e_extra(d));
                            if ((uint64_t)dist > out + o) {  // invalidStringReference
                                sh.anomaly = 1;
                                done = true;
                                break;
                            }
                            // source bytes that must already be final: the first min(run, dist)
                            const int64_t src = (int64_t)o - (int64_t)dist;
                            const int64_t hi  = src + (int64_t)min(run, dist);
                            bool ready = true;
                            if (hi > 0 && src < (int64_t)o_start) {
                                uint32_t a = (uint32_t)max(src, (int64_t)0);
                                uint32_t c = (uint32_t)min(hi, (int64_t)o_start);
                                if (a < c) ready = par_resolved(sh, a, c, t);
                            }
                            if (!ready) {
                                b = save;
                                blocked = true;
                                break;
                            }
                            const uint8_t* from = wdst + src;  // may point before the wave
                            uint8_t*       to   = wdst + o;
                            if (dist >= run) {
                                for (uint32_t k = 0; k < run; ++k) to[k] = from[k];
                            } else {
                                uint32_t q = 0;
                                for (uint32_t k = 0; k < run; ++k) {
                                    to[k] = from[q];
                                    if (++q == dist) q = 0;
                                }
                            }
                            o += run;
                            __threadfence_block();
                            ((volatile uint32_t*)sh.resolved)[t] = o;
                        } else {  // end of block (invalid entries cannot be on a verified chain)
                            done = true;
                        }
                    }
                    __threadfence_block();
                    ((volatile uint32_t*)sh.resolved)[t] = o;
                    if (!__syncthreads_or(blocked ? 1 : 0)) break;
                }
                if (sh.anomaly) { fallback = true; break; }
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
        return;
    }
    if (t == 0) {
        if (r->status == 0) r->status = st;
        r->produced      = out;
        r->consumed_bits = br.at();
        r->blocks        = blocks;
        r->resume_bit    = resume_bit;
        r->resume_out    = resume_out;
        r->phase         = phase;
    }
}

// host side: opt in to the large dynamic shared memory on the current device (once per context)
inline int configure_inflate_parallel()
{
    return (int)cudaFuncSetAttribute(inflate_parallel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(ParShared));
}

// returns a cudaError_t as int
inline int launch_inflate_parallel(cudaStream_t stream, const StreamJob* jobs, StreamResult* results,
                                   const uint32_t* order, int count, uint64_t* launches)
{
    inflate_parallel_kernel<<<(unsigned)count, PAR_THREADS, sizeof(ParShared), stream>>>(jobs, results, order, count);
    ++*launches;
    return (int)cudaGetLastError();
}

}  // namespace pngb200
