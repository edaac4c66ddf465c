// simt.h -- a tiny host-side SIMT emulator: runs a CUDA kernel body on the CPU, one fiber (ucontext)
// per CUDA thread, cooperative round-robin scheduling in ONE OS thread.  TEST INFRASTRUCTURE ONLY:
// it exists so that the control flow of the hand-written kernels (barrier protocol, warp collectives,
// shared-memory indexing, speculative decode bookkeeping) can be checked against zlib / the oracle in
// this container, which has no GPU.  Nothing in the product links or includes this file; the kernels
// are compiled for it by defining PNGB200_EMU (see csrc/common.cuh).
//
// What is modelled: threadIdx/blockIdx/blockDim/gridDim, __syncthreads (exited threads count as
// arrived), full- or sub-mask warp collectives (shfl*, ballot, any, all, match_any, syncwarp), atomics
// (plain operations: one OS thread), dynamic shared memory, __nanosleep as a yield (spin loops on
// other threads' progress must contain one, or a collective).  CTAs of a grid run either one after
// the other or interleaved (for kernels whose CTAs talk through global memory).  The scheduling order
// can be reversed / shuffled per run to shake out code that depends on thread order between barriers.
#pragma once

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <type_traits>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __constant__ static const
#define __align__(n) alignas(n)
#define __restrict__

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x = 1, y = 1, z = 1; };
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

namespace simt {

struct Coll {                      // a warp collective that is gathering its lanes (keyed by its mask:
    uint32_t mask = 0, arrived = 0;  // independent thread scheduling allows several at once in one warp)
    uint64_t vals[32];
    bool     active = false;
};

struct Warp {
    Coll     colls[8];
    uint64_t snap[32][32];         // per participating lane: the exchanged values
    uint32_t snap_mask[32];
    uint32_t released = 0;         // lanes whose snapshot is ready to be read
    uint32_t alive = 0;
};

struct Cta {
    unsigned id = 0;
    unsigned nthreads = 0, exited = 0, at_barrier = 0;
    uint64_t barrier_gen = 0;
    std::vector<Warp> warps;
    unsigned char* smem = nullptr;
};

struct Fiber {
    ucontext_t ctx;
    void*      stack = nullptr;
    Cta*       cta = nullptr;
    unsigned   tid = 0;
    bool       done = false;
    const char* wait = "";   // what the fiber is blocked on (stall dumps)
    uint32_t   wait_arg = 0;
};

struct Scheduler {
    ucontext_t          main_ctx;
    std::vector<Fiber*> fibers;
    Fiber*              cur = nullptr;
    std::function<void()> body;
    uint64_t            switches = 0;
};

inline Scheduler*& sched() { static Scheduler* s = nullptr; return s; }

}  // namespace simt

// CUDA built-in variables (single OS thread: plain globals refreshed at every fiber switch)
inline uint3& simt_tid() { static uint3 v; return v; }
inline uint3& simt_bid() { static uint3 v; return v; }
inline dim3&  simt_bdim() { static dim3 v; return v; }
inline dim3&  simt_gdim() { static dim3 v; return v; }
#define threadIdx simt_tid()
#define blockIdx simt_bid()
#define blockDim simt_bdim()
#define gridDim simt_gdim()

namespace simt {

inline void yield()
{
    Scheduler* S = sched();
    Fiber* f = S->cur;
    swapcontext(&f->ctx, &S->main_ctx);
}

inline Fiber* self() { return sched()->cur; }
inline unsigned char* dyn_smem() { return self()->cta->smem; }

inline void fiber_entry()
{
    Scheduler* S = sched();
    S->body();
    Fiber* f = S->cur;
    f->done = true;
    Cta* c = f->cta;
    c->exited++;
    c->warps[f->tid >> 5].alive &= ~(1u << (f->tid & 31));
    swapcontext(&f->ctx, &S->main_ctx);
}

// order: 0 ascending, 1 descending, >= 2 pseudo-random with that seed
// interleave: all CTAs resident at once (needed when CTAs wait for each other through global memory)
inline void launch(unsigned grid, unsigned block, size_t smem_bytes, std::function<void()> body, int order = 0,
                   bool interleave = false, size_t stack_bytes = 256 << 10, unsigned grid_y = 1)
{
  for (unsigned by = 0; by < grid_y; ++by) {
    simt_bid().y = by;
    simt_gdim().y = grid_y;
    Scheduler S;
    sched() = &S;
    S.body = body;
    simt_bdim().x = block;
    simt_gdim().x = grid;
    unsigned step = interleave ? grid : 1;
    for (unsigned b0 = 0; b0 < grid; b0 += step) {
        std::vector<Cta> ctas(step);
        std::vector<Fiber> fibers((size_t)step * block);
        for (unsigned c = 0; c < step; ++c) {
            Cta& cta = ctas[c];
            cta.id = b0 + c;
            cta.nthreads = block;
            cta.warps.resize((block + 31) / 32);
            cta.smem = (unsigned char*)aligned_alloc(128, (smem_bytes + 127) / 128 * 128 + 128);
            memset(cta.smem, 0xCD, smem_bytes);  // poison: shared memory is not zero on the device either
            for (unsigned t = 0; t < block; ++t) {
                cta.warps[t >> 5].alive |= 1u << (t & 31);
                Fiber& f = fibers[(size_t)c * block + t];
                f.cta = &cta;
                f.tid = t;
                f.stack = malloc(stack_bytes);
                getcontext(&f.ctx);
                f.ctx.uc_stack.ss_sp = f.stack;
                f.ctx.uc_stack.ss_size = stack_bytes;
                f.ctx.uc_link = &S.main_ctx;
                makecontext(&f.ctx, (void (*)())fiber_entry, 0);
            }
        }
        std::vector<size_t> idx(fibers.size());
        for (size_t i = 0; i < idx.size(); ++i) idx[i] = i;
        if (order == 1) std::reverse(idx.begin(), idx.end());
        uint64_t rng = 0x9E3779B97F4A7C15ull * (uint64_t)(order + 1);
        size_t live = fibers.size();
        uint64_t idle_rounds = 0;
        while (live) {
            if (order >= 2) {  // reshuffle every round
                for (size_t i = idx.size(); i > 1; --i) {
                    rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
                    std::swap(idx[i - 1], idx[rng % i]);
                }
            }
            uint64_t before = S.switches;
            size_t   finished = 0;
            for (size_t i : idx) {
                Fiber& f = fibers[i];
                if (f.done) continue;
                S.cur = &f;
                simt_tid().x = f.tid;
                simt_bid().x = f.cta->id;
                ++S.switches;
                swapcontext(&S.main_ctx, &f.ctx);
                if (f.done) ++finished;
            }
            live -= finished;
            (void)before;
            ++idle_rounds;
            if (getenv("SIMT_DUMP_ROUNDS") && idle_rounds == strtoull(getenv("SIMT_DUMP_ROUNDS"), nullptr, 0)) {
                for (Fiber& f : fibers)
                    if (!f.done) fprintf(stderr, "simt: cta %u tid %u waits on %s (%08x)\n", f.cta->id, f.tid, f.wait, f.wait_arg);
                for (Cta& c : ctas)
                    for (size_t wi = 0; wi < c.warps.size(); ++wi)
                        for (Coll& k : c.warps[wi].colls)
                            if (k.active)
                                fprintf(stderr, "simt: cta %u warp %zu collective mask %08x arrived %08x alive %08x\n", c.id, wi,
                                        k.mask, k.arrived, c.warps[wi].alive);
                abort();
            }
        }
        for (Fiber& f : fibers) free(f.stack);
        for (Cta& c : ctas) free(c.smem);
    }
    sched() = nullptr;
  }
}

inline void cta_barrier()
{
    Fiber* f = self();
    Cta*   c = f->cta;
    uint64_t gen = c->barrier_gen;
    f->wait = "__syncthreads";
    c->at_barrier++;
    for (;;) {
        if (c->barrier_gen != gen) return;
        if (c->at_barrier + c->exited >= c->nthreads) {
            c->at_barrier = 0;
            c->barrier_gen++;
            return;
        }
        yield();
    }
}

// every lane of `mask` deposits v; returns this lane's private copy of all deposited values
inline const uint64_t* warp_exchange(uint32_t mask, uint64_t v, uint32_t* got_mask = nullptr)
{
    Fiber*   f = self();
    Warp&    w = f->cta->warps[f->tid >> 5];
    unsigned lane = f->tid & 31;
    f->wait = "warp collective";
    f->wait_arg = mask;
    if (!(mask >> lane & 1)) { fprintf(stderr, "simt: lane %u not in its own mask %08x\n", lane, mask); abort(); }
    Coll* c = nullptr;
    for (;;) {
        for (Coll& k : w.colls)
            if (k.active && k.mask == mask) { c = &k; break; }
        if (!c)
            for (Coll& k : w.colls)
                if (!k.active) { c = &k; k.active = true; k.mask = mask; k.arrived = 0; break; }
        if (c) break;
        yield();
    }
    c->vals[lane] = v;
    c->arrived |= 1u << lane;
    for (;;) {
        if (w.released >> lane & 1) break;
        uint32_t need = mask & w.alive;
        if (c->active && c->mask == mask && (c->arrived >> lane & 1) && (c->arrived & need) == need) {
            uint32_t part = c->arrived;
            for (unsigned l = 0; l < 32; ++l)
                if (part >> l & 1) {
                    memcpy(w.snap[l], c->vals, sizeof c->vals);
                    w.snap_mask[l] = part;
                }
            w.released |= part;
            c->active = false;
            break;
        }
        yield();
    }
    w.released &= ~(1u << lane);
    if (got_mask) *got_mask = w.snap_mask[lane];
    return w.snap[lane];
}

template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

}  // namespace simt

#define PNGB200_DYN_SMEM(name) unsigned char* name = simt::dyn_smem()
#define __shared__ static

static inline void __syncthreads() { simt::cta_barrier(); }
static inline void __syncwarp(uint32_t mask = 0xffffffffu) { simt::warp_exchange(mask, 0); }
static inline void __threadfence_block() {}
static inline void __threadfence() {}
static inline long long clock64() { return 0; }
static inline void __nanosleep(unsigned) { simt::self()->wait = "nanosleep"; simt::yield(); }

template <class T> static inline T __shfl_sync(uint32_t mask, T v, int src)
{
    static_assert(sizeof(T) <= 8, "shfl width");
    const uint64_t* s = simt::warp_exchange(mask, simt::to_bits(v));
    return simt::from_bits<T>(s[src & 31]);
}
template <class T> static inline T __shfl_up_sync(uint32_t mask, T v, unsigned d)
{
    unsigned lane = threadIdx.x & 31;
    const uint64_t* s = simt::warp_exchange(mask, simt::to_bits(v));
    return lane >= d ? simt::from_bits<T>(s[lane - d]) : v;
}
template <class T> static inline T __shfl_down_sync(uint32_t mask, T v, unsigned d)
{
    unsigned lane = threadIdx.x & 31;
    const uint64_t* s = simt::warp_exchange(mask, simt::to_bits(v));
    return lane + d < 32 ? simt::from_bits<T>(s[lane + d]) : v;
}
template <class T> static inline T __shfl_xor_sync(uint32_t mask, T v, unsigned x)
{
    unsigned lane = threadIdx.x & 31;
    const uint64_t* s = simt::warp_exchange(mask, simt::to_bits(v));
    return simt::from_bits<T>(s[(lane ^ x) & 31]);
}
static inline unsigned __ballot_sync(uint32_t mask, int pred)
{
    uint32_t part;
    const uint64_t* s = simt::warp_exchange(mask, pred ? 1 : 0, &part);
    unsigned r = 0;
    for (unsigned l = 0; l < 32; ++l)
        if ((part >> l & 1) && s[l]) r |= 1u << l;
    return r;
}
static inline int __any_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(uint32_t mask, int pred)
{
    uint32_t part;
    const uint64_t* s = simt::warp_exchange(mask, pred ? 1 : 0, &part);
    for (unsigned l = 0; l < 32; ++l)
        if ((part >> l & 1) && !s[l]) return 0;
    return 1;
}
template <class T> static inline unsigned __match_any_sync(uint32_t mask, T v)
{
    uint32_t part;
    const uint64_t  mine = simt::to_bits(v);
    const uint64_t* s = simt::warp_exchange(mask, mine, &part);
    unsigned r = 0;
    for (unsigned l = 0; l < 32; ++l)
        if ((part >> l & 1) && s[l] == mine) r |= 1u << l;
    return r;
}
static inline int __syncthreads_or(int pred)
{
    // two barriers around a CTA-wide flag in a static (one CTA at a time uses it between the barriers)
    static int flag;
    simt::cta_barrier();
    flag = 0;
    simt::cta_barrier();
    if (pred) flag = 1;
    simt::cta_barrier();
    int r = flag;
    simt::cta_barrier();
    return r;
}

// ---- atomics (one OS thread: plain read-modify-write) ----
template <class T, class U> static inline T atomicAdd(T* p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicOr(T* p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> static inline T atomicAnd(T* p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class U> static inline T atomicXor(T* p, U v) { T o = *p; *p = (T)(o ^ (T)v); return o; }
template <class T, class U> static inline T atomicMin(T* p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMax(T* p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicExch(T* p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U, class V> static inline T atomicCAS(T* p, U c, V v) { T o = *p; if (o == (T)c) *p = (T)v; return o; }

// ---- loads ----
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }

// ---- integer intrinsics ----
static inline int      __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int      __popcll(uint64_t x) { return __builtin_popcountll(x); }
static inline int      __ffs(int x) { return __builtin_ffs(x); }
static inline int      __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int      __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int      __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline uint32_t __brev(uint32_t x)
{
    x = (x >> 16) | (x << 16);
    x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
    x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
    x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
    return x;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s)
{
    uint64_t v = (uint64_t)hi << 32 | lo;
    return (uint32_t)(v >> (s & 31));
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s)
{
    uint64_t v = (uint64_t)hi << 32 | lo;
    return (uint32_t)((v << (s & 31)) >> 32);
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
    uint64_t v = (uint64_t)b << 32 | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        uint32_t s = (sel >> (4 * i)) & 0xf;
        uint32_t byte = (uint32_t)(v >> (8 * (s & 7))) & 0xff;
        if (s & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
#define SIMT_PERBYTE(expr)                                            \
    uint32_t r = 0;                                                   \
    for (int i = 0; i < 4; ++i) {                                     \
        uint32_t x = (a >> (8 * i)) & 0xff, y = (b >> (8 * i)) & 0xff; \
        (void)x; (void)y;                                             \
        r |= ((uint32_t)(expr) & 0xff) << (8 * i);                    \
    }                                                                 \
    return r;
static inline uint32_t __vadd4(uint32_t a, uint32_t b) { SIMT_PERBYTE(x + y) }
static inline uint32_t __vsub4(uint32_t a, uint32_t b) { SIMT_PERBYTE(x - y) }
static inline uint32_t __vaddus4(uint32_t a, uint32_t b) { SIMT_PERBYTE(x + y > 255 ? 255 : x + y) }
static inline uint32_t __vabsdiffu4(uint32_t a, uint32_t b) { SIMT_PERBYTE(x > y ? x - y : y - x) }
static inline uint32_t __vcmpgeu4(uint32_t a, uint32_t b) { SIMT_PERBYTE(x >= y ? 0xff : 0) }
static inline uint32_t __vcmpleu4(uint32_t a, uint32_t b) { SIMT_PERBYTE(x <= y ? 0xff : 0) }
static inline uint32_t __vcmpeq4(uint32_t a, uint32_t b) { SIMT_PERBYTE(x == y ? 0xff : 0) }
static inline uint32_t __vavgu4(uint32_t a, uint32_t b) { SIMT_PERBYTE((x + y + 1) >> 1) }
static inline uint32_t __vsadu4(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        uint32_t x = (a >> (8 * i)) & 0xff, y = (b >> (8 * i)) & 0xff;
        r += x > y ? x - y : y - x;
    }
    return r;
}
static inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c)
{
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
    return c;
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

// CUDA's min/max accept mixed integer types
template <class A, class B> static inline constexpr std::common_type_t<A, B> min(A a, B b)
{
    using T = std::common_type_t<A, B>;
    return (T)a < (T)b ? (T)a : (T)b;
}
template <class A, class B> static inline constexpr std::common_type_t<A, B> max(A a, B b)
{
    using T = std::common_type_t<A, B>;
    return (T)a > (T)b ? (T)a : (T)b;
}
