// common.cuh -- device-side job/result records and small helpers shared by all kernels.
#pragma once

#ifdef PNGB200_EMU
// host-side SIMT emulation of the kernels (tests/emu/simt.h): test infrastructure, never the product
#include "simt.h"
#else
#include <cuda_runtime.h>
#define PNGB200_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
#include <stdint.h>

#include "../../include/pngb200.h"

namespace pngb200 {

// One DEFLATE stream to inflate (device-resident record).
struct StreamJob {
    const uint8_t* src;
    uint64_t       src_len;
    uint8_t*       dst;
    uint64_t       dst_cap;
    uint64_t       start_bit;   // resume point (bit offset of a block header), 0 = from the top
    uint64_t       start_out;   // bytes already produced before start_bit
    int32_t        format;      // pngb200_format
    int32_t        phase;       // where to resume: 0 stream header, 1 block header, 2 trailer
    // segments of a stream that several CTAs decode side by side (inflate_wave_kernel only):
    uint64_t       stop_bit;    // 0 = to the end of the stream; else stop at the first block boundary >= stop_bit
    uint32_t       symbolic;    // 1: dst is uint16_t[dst_cap]; a byte copied from in front of the segment becomes
                                // the marker 0x8000 | index into the 32 KiB window that precedes the segment
    uint32_t       pad_;
};

struct StreamResult {
    int32_t  status;
    uint32_t err_a, err_b;
    uint32_t checksum;
    uint32_t blocks;
    uint32_t declared;        // trailer checksum as read from the stream
    uint64_t produced;
    uint64_t consumed_bits;
    uint64_t resume_bit;      // start of the last block header not yet completed
    uint64_t resume_out;      // output size at resume_bit
    uint32_t trailer_seen;    // 1 when the final block and trailer were parsed
    uint32_t phase;           // phase to resume in at resume_bit (see StreamJob.phase)
    // device-side counters (the reference's -DDUMP_LZ77_BLOCKS style statistics)
    uint32_t stat_waves, stat_sync_rounds, stat_resolve_rounds, stat_fallback;
    uint64_t stat_tokens, stat_matches, stat_deferred;  // tokens / LZ77 matches on the stream, matches that had to wait
    uint32_t ck_done;         // 1: `checksum` was computed (and compared) by the inflate kernel itself
    uint32_t pad_;
    // SM cycles thread 0 of the stream's CTA spent per phase of the wave kernel (each ends at a barrier):
    // 0 header+tables 1 stage 2 speculate 3 walk 4 chain 5 count+scan 6 emit 7 resolve 8 store 9 stored blocks
    uint64_t stat_cycles[12];
};

// One image for the unfilter stage.
struct ImageJob {
    const uint8_t*      filtered;   // inflated IDAT stream
    uint8_t*            pixels;     // PNG.Image.storage
    const StreamResult* inflated;   // producer's result (rows available = produced / (pitch+1)); may be null
    uint64_t            filtered_len;  // used when `inflated` is null
    uint32_t            width, height;
    uint32_t            pitch;      // bytes per row (non-interlaced)
    uint8_t             volume, depth, interlaced, bpp;
};

// bytes of filtered stream that may be consumed: nothing after an inflate error (the reference
// throws out of Inflator.push before any row is pulled), everything produced otherwise
__device__ __forceinline__ uint64_t usable_bytes(const StreamResult* r, uint64_t fallback)
{
    if (r == nullptr) return fallback;
    return r->status < 0 ? 0 : r->produced;
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p)
{
#ifdef PNGB200_EMU
    return *(const volatile uint32_t*)p;
#else
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v)
{
#ifdef PNGB200_EMU
    *(volatile uint32_t*)p = v;
#else
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}

}  // namespace pngb200
