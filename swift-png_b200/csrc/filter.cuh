// filter.cuh -- encode-side stage 1: PNG.Image.collect + PNG.Encoder.filter for every row.
//
// Replaces PNG.Encoder.filter / score (Sources/PNG/Encoding/PNG.Encoder.swift:132-204,230-234)
// and the gather of PNG.Image.collect (Sources/PNG/PNG.Image.swift:431-544).  All five candidate
// rows are functions of the UNFILTERED current and previous rows only, so every row of every
// image is independent: one warp per row computes the five sum|int8| scores in a single pass
// (warp-shuffle reduction), picks the first minimum in the order None, Sub, Up, Average, Paeth
// (strict <), and writes the winning candidate behind its filter-type byte.
#pragma once

#include "common.cuh"
#include "unfilter.cuh"

namespace pngb200 {

struct FilterJob {
    const uint8_t* pixels;    // PNG.Image.storage
    uint8_t*       filtered;  // out
    uint32_t       width, height;
    uint8_t        volume, depth, interlaced, bpp;
};

constexpr int FILTER_WARPS = 8;

inline uint64_t filter_rows(uint32_t w, uint32_t h, int interlaced)
{
    if (!interlaced) return h;
    static const int A7[7][4] = {{0, 0, 3, 3}, {4, 0, 3, 3}, {0, 4, 2, 3}, {2, 0, 2, 2},
                                 {0, 2, 1, 2}, {1, 0, 1, 1}, {0, 1, 0, 1}};
    uint64_t rows = 0;
    for (int z = 0; z < 7; ++z) {
        uint64_t sx = ((uint64_t)w + (1u << A7[z][2]) - A7[z][0] - 1) >> A7[z][2];
        uint64_t sy = ((uint64_t)h + (1u << A7[z][3]) - A7[z][1] - 1) >> A7[z][3];
        if (sx && sy) rows += sy;
    }
    return rows;
}

// a scanline of one (sub)image, addressed bytewise as PNG.Image.collect would have packed it
struct RowView {
    const uint8_t* storage;
    uint32_t       width;     // full image width
    uint32_t       oy;        // storage row
    uint32_t       bx, ex;    // first column, log2 column stride
    uint32_t       sw;        // pixels in this scanline
    uint32_t       depth, bpp;
    bool           valid;     // false: the all-zero reference row above the first row of a pass

    __device__ __forceinline__ uint32_t byte(uint32_t i) const
    {
        if (!valid) return 0;
        if (depth >= 8) {
            uint32_t px = i / bpp, c = i - px * bpp;
            return storage[((uint64_t)oy * width + bx + ((uint64_t)px << ex)) * bpp + c];
        }
        uint32_t per = 8 / depth, mask = (1u << depth) - 1, v = 0;
        for (uint32_t k = 0; k < per; ++k) {
            uint32_t px = i * per + k;
            if (px < sw) {
                uint32_t s = storage[(uint64_t)oy * width + bx + ((uint64_t)px << ex)] & mask;
                v |= s << (((~px) & (per - 1)) * depth);
            }
        }
        return v;
    }
};

__device__ __forceinline__ uint32_t abs_i8(uint32_t b) { return b & 0x80 ? 256 - b : b; }

__global__ void __launch_bounds__(FILTER_WARPS * 32)
filter_rows_kernel(const FilterJob* jobs, const uint32_t* row_base, uint32_t njobs, uint32_t total_rows)
{
    const unsigned lane = lane_id();
    const uint32_t t    = blockIdx.x * FILTER_WARPS + (threadIdx.x >> 5);
    if (t >= total_rows) return;
    uint32_t lo = 0, hi = njobs;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (row_base[mid] <= t) lo = mid;
        else hi = mid;
    }
    const FilterJob job = jobs[lo];
    uint32_t        r   = t - row_base[lo];  // row index within the image's filtered stream
    // locate (pass, y) and the output offset
    uint32_t bx = 0, by = 0, ex = 0, ey = 0, sw = job.width, y = r;
    uint64_t out_off = 0;
    uint32_t pitch = (uint32_t)(((uint64_t)job.width * job.volume + 7) >> 3);
    if (job.interlaced) {
        for (int z = 0; z < 7; ++z) {
            uint32_t px = (job.width + (1u << c_adam7[z][2]) - c_adam7[z][0] - 1) >> c_adam7[z][2];
            uint32_t py = (job.height + (1u << c_adam7[z][3]) - c_adam7[z][1] - 1) >> c_adam7[z][3];
            if (px == 0 || py == 0) continue;
            uint32_t pp = (uint32_t)(((uint64_t)px * job.volume + 7) >> 3);
            if (y < py) {
                bx = c_adam7[z][0]; by = c_adam7[z][1]; ex = c_adam7[z][2]; ey = c_adam7[z][3];
                sw = px;
                pitch = pp;
                break;
            }
            y -= py;
            out_off += (uint64_t)py * (pp + 1);
        }
    }
    out_off += (uint64_t)y * (pitch + 1);
    RowView cur{job.pixels, job.width, by + (y << ey), bx, ex, sw, job.depth, job.bpp, true};
    RowView prev = cur;
    prev.valid = y > 0;
    prev.oy = by + ((y ? y - 1 : 0) << ey);
    const uint32_t d = job.bpp;

    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    for (uint32_t i = lane; i < pitch; i += 32) {
        uint32_t x = cur.byte(i), b = prev.byte(i);
        uint32_t a = i >= d ? cur.byte(i - d) : 0, c = i >= d ? prev.byte(i - d) : 0;
        s0 += abs_i8(x);
        s1 += abs_i8((x - a) & 0xff);
        s2 += abs_i8((x - b) & 0xff);
        s3 += abs_i8((x - ((a + b) >> 1)) & 0xff);
        s4 += abs_i8((x - paeth1(a, b, c)) & 0xff);
    }
    for (int o = 16; o; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        s3 += __shfl_xor_sync(0xffffffffu, s3, o);
        s4 += __shfl_xor_sync(0xffffffffu, s4, o);
    }
    uint32_t best = 0, minimum = s0;
    if (s1 < minimum) { minimum = s1; best = 1; }
    if (s2 < minimum) { minimum = s2; best = 2; }
    if (s3 < minimum) { minimum = s3; best = 3; }
    if (s4 < minimum) { minimum = s4; best = 4; }
    uint8_t* out = job.filtered + out_off;
    if (lane == 0) out[0] = (uint8_t)best;
    for (uint32_t i = lane; i < pitch; i += 32) {
        uint32_t x = cur.byte(i), p = 0;
        if (best) {
            uint32_t b = prev.byte(i);
            uint32_t a = i >= d ? cur.byte(i - d) : 0, c = i >= d ? prev.byte(i - d) : 0;
            p = best == 1 ? a : best == 2 ? b : best == 3 ? (a + b) >> 1 : paeth1(a, b, c);
        }
        out[1 + i] = (uint8_t)(x - p);
    }
}

}  // namespace pngb200
