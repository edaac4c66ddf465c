// color.cuh -- colour targets on the device: PNG.Image.storage <-> [PNG.RGBA<T>] / [PNG.VA<T>]
// (the unpack that sits inside the reference's own timed decode loop,
// Benchmarks/Decompression/Swift/Main.swift:105-106, and its inverse used by PNG.Image.init(packing:)).
//
// The reference expresses these as generic closures over convolve / deconvolve
// (Sources/PNG/PNG.swift:149-1285, ColorTargets/PNG.RGBA.swift:262-478, ColorTargets/PNG.VA.swift,
// ColorTargets/PNG.Color.swift).  Here each is one elementwise, HBM-bound map: one thread per pixel
// per iteration, the format switch uniform across the CTA, palettes staged in shared memory.
// Algorithmic bytes per pixel: storage bytes read + target bytes written (or the reverse).
#pragma once

#include "common.cuh"

namespace pngb200 {

constexpr int COLOR_THREADS = 256;
constexpr int COLOR_TILE    = 8 * COLOR_THREADS;  // pixels per CTA iteration

struct ColorJob {
    uint8_t*  storage;      // PNG.Image.storage (unpack: read, pack: written)
    uint8_t*  pixels;       // target array, native-endian T components
    uint64_t  count;        // pixels
    uint32_t  palette_off;  // first entry of this job's palette in ColorParams::palettes
    uint16_t  palette_count;
    uint16_t  key[3];
    uint8_t   color, depth, bgr, has_key;
    int32_t   status;
};

struct ColorParams {
    ColorJob*       jobs;
    const uint32_t* palettes;  // r | g << 8 | b << 16 | a << 24
    uint32_t        count;
    int             target;      // PNGB200_TARGET_*
    int             alpha_mode;  // PNGB200_ALPHA_*
};

__device__ __forceinline__ int color_channels(int color) { return color == 0 || color == 3 ? 1 : color == 2 ? 3 : color == 4 ? 2 : 4; }

// PNG.quantum + the transform closures of convolve(_:of:depth:kernel:) (PNG.swift:255-261, 494-523)
__device__ __forceinline__ uint32_t color_widen(uint32_t v, int depth, int tbits)
{
    if (tbits == depth) return v;
    if (tbits > depth) return v * (((1u << tbits) - 1u) / ((1u << depth) - 1u));
    return v >> (depth - tbits);
}
// the transform closures of deconvolve(_:as:depth:kernel:) (PNG.swift:1063-1095)
__device__ __forceinline__ uint32_t color_narrow(uint32_t v, int tbits, int depth)
{
    if (tbits == depth) return v;
    if (tbits < depth) return v * (((1u << depth) - 1u) / ((1u << tbits) - 1u));
    return v >> (tbits - depth);
}
// PNG.premultiply (PNG.swift:54-66)
template <int BITS> __device__ __forceinline__ uint32_t color_premultiply(uint32_t c, uint32_t a)
{
    constexpr uint32_t MAX = (1u << BITS) - 1u;
    return (c * a + (MAX >> 1)) / MAX;
}
// PNG.straighten (PNG.swift:100-120); saturates where the reference's dividingFullWidth traps
template <int BITS> __device__ __forceinline__ uint32_t color_straighten(uint32_t p, uint32_t a)
{
    constexpr uint32_t MAX = (1u << BITS) - 1u;
    if (a == 0) return p;
    return min((MAX * p + (a >> 1)) / a, MAX);
}

template <int TBITS>
__device__ __forceinline__ void color_alpha(uint32_t& r, uint32_t& g, uint32_t& b, uint32_t& a, int mode)
{
    if (mode == PNGB200_ALPHA_PREMULTIPLIED) {
        r = color_premultiply<TBITS>(r, a), g = color_premultiply<TBITS>(g, a), b = color_premultiply<TBITS>(b, a);
    } else if (mode == PNGB200_ALPHA_STRAIGHTENED) {
        r = color_straighten<TBITS>(r, a), g = color_straighten<TBITS>(g, a), b = color_straighten<TBITS>(b, a);
    } else if (TBITS == 16 && (mode == PNGB200_ALPHA_PREMULTIPLIED_AS8 || mode == PNGB200_ALPHA_STRAIGHTENED_AS8)) {
        // premultiplied(as: UInt8.self) / straightened(as: UInt8.self) (PNG.RGBA.swift:141-155, 187-201)
        const uint32_t a8 = a >> 8;
        if (mode == PNGB200_ALPHA_PREMULTIPLIED_AS8) {
            r = color_premultiply<8>(r >> 8, a8) * 257u, g = color_premultiply<8>(g >> 8, a8) * 257u;
            b = color_premultiply<8>(b >> 8, a8) * 257u;
        } else {
            r = color_straighten<8>(r >> 8, a8) * 257u, g = color_straighten<8>(g >> 8, a8) * 257u;
            b = color_straighten<8>(b >> 8, a8) * 257u;
        }
        a = a8 * 257u;
    }
}

// the bytes of pixel i (bpp = 1..8), first sample in the low bits
__device__ __forceinline__ uint64_t color_load_pixel(const uint8_t* storage, uint64_t i, int bpp, bool aligned)
{
    const uint8_t* p = storage + i * bpp;
    if (aligned) {
        if (bpp == 8) { uint2 v = *(const uint2*)p; return (uint64_t)v.y << 32 | v.x; }
        if (bpp == 4) return *(const uint32_t*)p;
        if (bpp == 2) return *(const uint16_t*)p;
    }
    uint64_t v = 0;
    for (int k = 0; k < bpp; ++k) v |= (uint64_t)p[k] << (8 * k);
    return v;
}
__device__ __forceinline__ void color_store_pixel(uint8_t* storage, uint64_t i, int bpp, bool aligned, uint64_t v)
{
    uint8_t* p = storage + i * bpp;
    if (aligned) {
        if (bpp == 8) { *(uint2*)p = make_uint2((uint32_t)v, (uint32_t)(v >> 32)); return; }
        if (bpp == 4) { *(uint32_t*)p = (uint32_t)v; return; }
        if (bpp == 2) { *(uint16_t*)p = (uint16_t)v; return; }
    }
    for (int k = 0; k < bpp; ++k) p[k] = (uint8_t)(v >> (8 * k));
}

template <int TBITS, bool VA>
__device__ void unpack_image(const ColorJob& job, ColorJob* slot, const uint32_t* palette, int alpha_mode)
{
    const int  ch      = color_channels(job.color);
    const int  bpp     = ch * (job.depth == 16 ? 2 : 1);
    const bool wide    = job.depth == 16;
    const bool aligned = (((uintptr_t)job.storage) & 7) == 0;
    constexpr uint32_t TMAX = (1u << TBITS) - 1u;
    for (uint64_t base = (uint64_t)blockIdx.x * COLOR_TILE; base < job.count; base += (uint64_t)gridDim.x * COLOR_TILE) {
#pragma unroll 2
        for (uint64_t i = base + threadIdx.x; i < min(base + COLOR_TILE, job.count); i += COLOR_THREADS) {
            const uint64_t bits = color_load_pixel(job.storage, i, bpp, aligned);
            uint32_t raw[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                // A(bigEndian:) (PNG.swift:152-204)
                raw[c] = wide ? __byte_perm((uint32_t)(bits >> (16 * c)), 0, 0x4401) & 0xffffu
                              : (uint32_t)(bits >> (8 * c)) & 0xffu;
            }
            uint32_t r, g, b, a;
            if (job.color == 3) {
                if (raw[0] >= job.palette_count) {  // palette[i] traps in the reference
                    atomicMin(&slot->status, (int32_t)PNGB200_ERR_PNG_PALETTE_INDEX);
                    continue;
                }
                const uint32_t e = palette[raw[0]];
                r = color_widen(e & 0xff, 8, TBITS), g = color_widen((e >> 8) & 0xff, 8, TBITS);
                b = color_widen((e >> 16) & 0xff, 8, TBITS), a = color_widen(e >> 24, 8, TBITS);
            } else if (job.color == 0 || job.color == 4) {
                r = g = b = color_widen(raw[0], job.depth, TBITS);
                a = job.color == 4 ? color_widen(raw[1], job.depth, TBITS)
                                   : (job.has_key && raw[0] == job.key[0]) ? 0u : TMAX;
            } else {
                const uint32_t c0 = color_widen(raw[0], job.depth, TBITS), c1 = color_widen(raw[1], job.depth, TBITS),
                               c2 = color_widen(raw[2], job.depth, TBITS);
                r = job.bgr ? c2 : c0, g = c1, b = job.bgr ? c0 : c2;
                a = job.color == 6 ? color_widen(raw[3], job.depth, TBITS)
                                   : (job.has_key && raw[0] == job.key[0] && raw[1] == job.key[1] && raw[2] == job.key[2]) ? 0u : TMAX;
            }
            color_alpha<TBITS>(r, g, b, a, alpha_mode);
            if (VA) {
                if (TBITS == 8) ((uint16_t*)job.pixels)[i] = (uint16_t)(r | a << 8);
                else ((uint32_t*)job.pixels)[i] = r | a << 16;
            } else {
                if (TBITS == 8) ((uint32_t*)job.pixels)[i] = r | g << 8 | b << 16 | a << 24;
                else ((uint2*)job.pixels)[i] = make_uint2(r | g << 16, b | a << 16);
            }
        }
    }
}

__global__ void __launch_bounds__(COLOR_THREADS) unpack_kernel(ColorParams p)
{
    __shared__ uint32_t palette[256];
    for (uint32_t j = blockIdx.y; j < p.count; j += gridDim.y) {
        const ColorJob job = p.jobs[j];
        if (job.color == 3) {
            __syncthreads();
            if (threadIdx.x < job.palette_count) palette[threadIdx.x] = p.palettes[job.palette_off + threadIdx.x];
            __syncthreads();
        }
        switch (p.target) {
        case PNGB200_TARGET_RGBA8:  unpack_image<8, false>(job, p.jobs + j, palette, p.alpha_mode); break;
        case PNGB200_TARGET_RGBA16: unpack_image<16, false>(job, p.jobs + j, palette, p.alpha_mode); break;
        case PNGB200_TARGET_VA8:    unpack_image<8, true>(job, p.jobs + j, palette, p.alpha_mode); break;
        default:                    unpack_image<16, true>(job, p.jobs + j, palette, p.alpha_mode); break;
        }
    }
}

template <int TBITS, bool VA>
__device__ void pack_image(const ColorJob& job, const uint32_t* palette)
{
    const int  ch      = color_channels(job.color);
    const int  bpp     = ch * (job.depth == 16 ? 2 : 1);
    const bool wide    = job.depth == 16;
    const bool aligned = (((uintptr_t)job.storage) & 7) == 0;
    for (uint64_t base = (uint64_t)blockIdx.x * COLOR_TILE; base < job.count; base += (uint64_t)gridDim.x * COLOR_TILE) {
#pragma unroll 2
        for (uint64_t i = base + threadIdx.x; i < min(base + COLOR_TILE, job.count); i += COLOR_THREADS) {
            uint32_t r, g, b, a;
            if (VA) {
                if (TBITS == 8) { uint32_t v = ((const uint16_t*)job.pixels)[i]; r = v & 0xff, a = v >> 8; }
                else { uint32_t v = ((const uint32_t*)job.pixels)[i]; r = v & 0xffff, a = v >> 16; }
                g = b = r;
            } else {
                if (TBITS == 8) { uint32_t v = ((const uint32_t*)job.pixels)[i]; r = v & 0xff, g = (v >> 8) & 0xff, b = (v >> 16) & 0xff, a = v >> 24; }
                else { uint2 v = ((const uint2*)job.pixels)[i]; r = v.x & 0xffff, g = v.x >> 16, b = v.y & 0xffff, a = v.y >> 16; }
            }
            if (job.color == 3) {
                // default indexer (PNG.Color.swift): palette -> index hash table, missing colours -> 0
                const uint32_t q = color_narrow(r, TBITS, 8) | color_narrow(g, TBITS, 8) << 8 |
                                   color_narrow(b, TBITS, 8) << 16 | color_narrow(a, TBITS, 8) << 24;
                uint32_t idx = 0;
                for (uint32_t k = 0; k < job.palette_count; ++k)
                    if (palette[k] == q) { idx = k; break; }
                job.storage[i] = (uint8_t)idx;
                continue;
            }
            uint32_t s[4] = {0, 0, 0, 0};
            if (job.color == 0) s[0] = r;
            else if (job.color == 4) s[0] = r, s[1] = a;
            else s[0] = job.bgr ? b : r, s[1] = g, s[2] = job.bgr ? r : b, s[3] = a;
            uint64_t bits = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t v = color_narrow(s[c], TBITS, job.depth);
                if (c < ch) bits |= wide ? (uint64_t)(__byte_perm(v, 0, 0x4401) & 0xffffu) << (16 * c) : (uint64_t)(v & 0xffu) << (8 * c);
            }
            color_store_pixel(job.storage, i, bpp, aligned, bits);
        }
    }
}

__global__ void __launch_bounds__(COLOR_THREADS) pack_kernel(ColorParams p)
{
    __shared__ uint32_t palette[256];
    for (uint32_t j = blockIdx.y; j < p.count; j += gridDim.y) {
        const ColorJob job = p.jobs[j];
        if (job.color == 3) {
            __syncthreads();
            if (threadIdx.x < job.palette_count) palette[threadIdx.x] = p.palettes[job.palette_off + threadIdx.x];
            __syncthreads();
        }
        switch (p.target) {
        case PNGB200_TARGET_RGBA8:  pack_image<8, false>(job, palette); break;
        case PNGB200_TARGET_RGBA16: pack_image<16, false>(job, palette); break;
        case PNGB200_TARGET_VA8:    pack_image<8, true>(job, palette); break;
        default:                    pack_image<16, true>(job, palette); break;
        }
    }
}

}  // namespace pngb200
