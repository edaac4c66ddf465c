// inflate_segments.cuh -- more than one CTA per DEFLATE stream (SURVEY section 7 "Stage B").
//
// The reference's token loop is serial per stream (Stream.readBlock(with:),
// Sources/LZ77/Inflator/LZ77.InflatorBuffers.Stream.swift:266-381) and so is one CTA of
// inflate_wave_kernel.  A batch with fewer streams than CTA slots (BASELINE configs 4 and 5: 8 images
// per GPU, one 1 GiB gzip stream) would leave the GPU idle, so such streams are cut into SEGMENTS:
//
//   1. block_search_kernel (block_search.cuh) finds, for every wanted split point, the next bit offset
//      that holds a plausible dynamic-block header (a pure function of the offset; false positives are
//      caught in step 3).
//   2. inflate_wave_kernel decodes every segment with its own CTA into 16-bit symbols: bytes copied from
//      in front of the segment are markers 0x8000 | window index (see emit_token_sym).
//   3. the host accepts a stream only if every segment ended exactly where the next one starts (on a
//      block boundary) -- otherwise the stream is simply decoded whole; correct by construction.
//   4. window_propagate_kernel resolves, segment by segment, the 32 KiB window in front of each segment
//      (a serial chain per stream, 32 KiB per link), marker_resolve_kernel then replaces the markers of
//      all segments in parallel and packs the symbols into the byte stream at their final offsets.
#pragma once

#include "common.cuh"

namespace pngb200 {

constexpr uint32_t SEG_WINDOW = 32768;

struct SegmentRecord {         // one per segment, in stream order; segments of one stream are adjacent
    const uint16_t* sym;       // the segment's symbols
    uint8_t*        out;       // final position of the segment's first byte
    uint64_t        produced;  // symbols in the segment
    uint32_t        stream;    // index of the stream (window chain) it belongs to
    uint32_t        first;     // 1: first segment of its stream (no window in front, no markers)
};

// window[k] (32 KiB of bytes) = the output in front of segment k.  One CTA per stream walks the chain:
// window[k+1] = last 32 KiB of (window[k] ++ resolved segment k).
__global__ void __launch_bounds__(256) window_propagate_kernel(const SegmentRecord* segs, const uint32_t* stream_first,
                                                               uint32_t nstreams, uint8_t* windows)
{
    if (blockIdx.x >= nstreams) return;
    const uint32_t lo = stream_first[blockIdx.x], hi = stream_first[blockIdx.x + 1];
    for (uint32_t k = lo; k + 1 < hi; ++k) {
        const SegmentRecord s  = segs[k];
        const uint8_t*      w  = windows + (size_t)k * SEG_WINDOW;        // window in front of segment k
        uint8_t*            wn = windows + (size_t)(k + 1) * SEG_WINDOW;  // window in front of segment k + 1
        const uint64_t      n  = s.produced;
        for (uint32_t i = threadIdx.x; i < SEG_WINDOW; i += blockDim.x) {
            // byte i of the next window is byte (n - 32768 + i) of this segment, or, if the segment is shorter
            // than the window, byte (i + n) of this segment's own window
            uint8_t v;
            if (n + i >= SEG_WINDOW) {
                const uint16_t x = s.sym[n + i - SEG_WINDOW];
                v = (x & 0x8000u) ? (s.first ? 0 : w[x & 0x7fffu]) : (uint8_t)x;
            } else {
                v = s.first ? 0 : w[i + n];
            }
            wn[i] = v;
        }
        __threadfence();
        __syncthreads();
    }
}

// every symbol of every segment -> its byte at its final place; 16 bytes per thread-iteration
__global__ void __launch_bounds__(256) marker_resolve_kernel(const SegmentRecord* segs, uint32_t nsegs, const uint8_t* windows,
                                                             const uint64_t* chunk_base)
{
    // chunk_base[k]: exclusive prefix of ceil(produced / 4096) over the segments; one CTA per 4096-symbol chunk
    const uint64_t chunk = blockIdx.x;
    uint32_t lo = 0, hi = nsegs;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (chunk_base[mid] <= chunk) lo = mid;
        else hi = mid;
    }
    const SegmentRecord s = segs[lo];
    const uint8_t*      w = windows + (size_t)lo * SEG_WINDOW;
    const uint64_t base = (chunk - chunk_base[lo]) * 4096;
    for (uint32_t j = threadIdx.x; j < 4096; j += blockDim.x) {
        const uint64_t i = base + j;
        if (i >= s.produced) break;
        const uint16_t x = s.sym[i];
        s.out[i] = (x & 0x8000u) ? (s.first ? 0 : w[x & 0x7fffu]) : (uint8_t)x;
    }
}

}  // namespace pngb200
