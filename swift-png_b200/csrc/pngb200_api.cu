// pngb200_api.cu -- the C ABI declared in include/pngb200.h: contexts, batching, host glue.
//
// The library talks to the CUDA runtime directly (no torch types anywhere); callers that live in
// a PyTorch process pass raw device pointers (tensor.data_ptr()) with PNGB200_MEM_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "checksum.cuh"
#include "common.cuh"
#include "deflate.cuh"
#include "filter.cuh"
#include "color.cuh"
#include "crc32.cuh"
#include "inflate_wave.cuh"
#include "inflate_parallel.cuh"
#include "inflate_cells.cuh"
#include "block_search.cuh"
#include "inflate_segments.cuh"
#include "inflate_serial.cuh"
#include "unfilter.cuh"

using namespace pngb200;

namespace {

thread_local std::string g_last_error;

struct DevBuf {
    void*  p   = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max(n + std::min(n / 4, (size_t)1 << 30), (size_t)1 << 16);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            want = n;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return (T*)p; }
};

struct PinBuf {
    void*  p   = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max(n + n / 4, (size_t)1 << 12);
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return (T*)p; }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

const int ADAM7[7][4] = {{0, 0, 3, 3}, {4, 0, 3, 3}, {0, 4, 2, 3}, {2, 0, 2, 2},
                         {0, 2, 1, 2}, {1, 0, 1, 1}, {0, 1, 0, 1}};

}  // namespace

struct pngb200_ctx {
    int          device = 0;
    cudaStream_t stream = nullptr;
    uint64_t     launches = 0;
    int          inflate_mode = 0;
    int          last_engine = -1;       // whole-stream engine of the last batch: 0 round-1, 1 ring, 2 cells, -1 none
    bool         cells_auto = false;     // automatic mode may pick inflate_cells_kernel (set once measured faster)
    bool         cells_segments = true;  // segments (several CTAs per stream) are decoded by inflate_cells_kernel: 3 CTA slots per SM instead
                                         // of 2 (r02b: 8 x 8K RGBA8 6 143 -> 7 089 MPixels/s; PNGB200_CELLS_SEGMENTS=0 restores the ring kernel)
    int          sm_count = 148;
    std::string  error;
    bool         pending = false;
    int          pending_memspace = 0;
    // device workspaces (grow-only)
    DevBuf d_jobs, d_results, d_imgjobs, d_genjobs, d_misc, d_partial, d_filtered, d_in, d_out, d_order, d_scratch, d_dfscratch, d_dfjobs, d_dfres, d_enc,
           d_file, d_crc, d_seg, d_crctab, d_sgjobs, d_sgres, d_sgsym, d_sgsearch, d_sgrec, d_sgwin;
    // pinned host tables
    PinBuf h_jobs, h_results, h_imgjobs, h_genjobs, h_misc, h_order, h_crc, h_seg, h_sgsearch, h_sgjobs, h_sgres, h_sgrec;
    uint64_t seg_streams = 0, seg_segments = 0, seg_fallbacks = 0;  // last batch: streams cut into segments, segments, rejected
    uint64_t scratch_stride = 0;       // layout of d_scratch the last inflate launch used
    size_t parallel_threshold = 8192;  // streams at least this long use the block-parallel kernel
    unsigned long long* d_hist = nullptr;   // filter-type histogram of the last wavefront-unfilter launch (in d_imgjobs)
    size_t peer_streams = 0;           // lanes: streams of the whole host batch (its chunks run side by side on this GPU)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // decode stage boundaries
    // geometry of the pending decode batch
    std::vector<uint64_t> expected;   // filtered bytes expected per image
    std::vector<size_t>   out_offset; // staging offsets (HOST memspace)
    std::vector<size_t>   out_bytes;
    // helper contexts (own stream + workspaces) that pipeline big host-memory batches: while one
    // lane's PCIe copies run, another lane's kernels do
    std::vector<pngb200_ctx*> lanes;
    // Bulk H2D copies of the call in flight, issued by run_inflate after its own small table uploads and
    // right before its first launch: copies of one direction are served in issue order across all
    // streams, so a table queued behind another lane's gigabyte would hold this lane's kernels back.
    std::function<int()> bulk_h2d;
};

namespace {

int set_error(pngb200_ctx* ctx, int code, const char* fmt, ...)
{
    char    buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf;
    g_last_error = buf;
    return code;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return set_error(ctx, PNGB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call,               \
                             cudaGetErrorString(e_), __FILE__, __LINE__);                         \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// the CRC-32 byte table and shift operators (crc32.cuh), uploaded once per context
int ensure_crc_tables(pngb200_ctx* ctx)
{
    if (ctx->d_crctab.p) return PNGB200_OK;
    std::vector<uint32_t> t(CRC_TABLE_WORDS);
    crc_build_tables(t.data());
    CU(ctx->d_crctab.reserve(sizeof(uint32_t) * CRC_TABLE_WORDS));
    CU(cudaMemcpyAsync(ctx->d_crctab.p, t.data(), sizeof(uint32_t) * CRC_TABLE_WORDS, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PNGB200_OK;
}

// ---- more than one CTA per stream (inflate_segments.cuh) ----
// Streams of `par` that are worth cutting are decoded here, segment by segment; on return `par` holds the
// streams that still have to go through the whole-stream kernels (not cut, or a segment did not line up).
// Two host round trips (split points, segment results): the path exists for batches that would otherwise
// leave most of the GPU idle.
int run_segments(pngb200_ctx* ctx, const StreamJob* h_jobs, std::vector<uint32_t>& par)
{
    ctx->seg_streams = ctx->seg_segments = ctx->seg_fallbacks = 0;
    const size_t slots = (size_t)ctx->sm_count * (ctx->cells_segments ? CL_CTAS_PER_SM : WV_CTAS_PER_SM);
    constexpr uint64_t kMinSegment = 256u << 10;  // compressed bytes per segment, at least
    if (par.empty() || par.size() * 2 > slots) return PNGB200_OK;
    const size_t per_stream = std::max<size_t>(1, slots / par.size());
    struct Cut { uint32_t stream; uint32_t nseg; size_t first_search; };
    std::vector<Cut> cuts;
    size_t nsearch = 0;
    for (uint32_t i : par) {
        const size_t nseg = std::min<size_t>(per_stream, h_jobs[i].src_len / kMinSegment);
        if (nseg < 4 || h_jobs[i].start_bit != 0 || h_jobs[i].phase != 0 || h_jobs[i].dst_cap < (1u << 20)) continue;
        cuts.push_back(Cut{i, (uint32_t)nseg, nsearch});
        nsearch += nseg - 1;
    }
    if (cuts.empty()) return PNGB200_OK;
    // 1. split points: the first plausible dynamic-block header at or after k / nseg of the stream
    CU(ctx->h_sgsearch.reserve(sizeof(SearchJob) * nsearch));
    CU(ctx->d_sgsearch.reserve(sizeof(SearchJob) * nsearch));
    SearchJob* sj = ctx->h_sgsearch.as<SearchJob>();
    for (const Cut& c : cuts) {
        const StreamJob& j = h_jobs[c.stream];
        const uint64_t bits = 8 * j.src_len, step = bits / c.nseg;
        for (uint32_t k = 1; k < c.nseg; ++k) {
            SearchJob& q = sj[c.first_search + k - 1];
            q.src = j.src;
            q.src_len = j.src_len;
            q.from_bit = k * step;
            q.limit_bit = k + 1 < c.nseg ? (k + 1) * step : bits;
            q.found = ~0ull;
        }
    }
    CU(cudaMemcpyAsync(ctx->d_sgsearch.p, sj, sizeof(SearchJob) * nsearch, cudaMemcpyHostToDevice, ctx->stream));
    block_search_kernel<<<dim3((unsigned)nsearch, BS_CTAS), 256, 0, ctx->stream>>>(ctx->d_sgsearch.as<SearchJob>(), (uint32_t)nsearch);
    ctx->launches++;
    CU(cudaMemcpyAsync(sj, ctx->d_sgsearch.p, sizeof(SearchJob) * nsearch, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    // 2. one job per segment; symbols go to a scratch sized 1.5 x the segment's share of the output bound
    std::vector<StreamJob> sg;
    std::vector<uint32_t>  sg_stream, first_of;   // stream of each segment; first segment of each cut (+ end)
    size_t sym_total = 0;
    uint64_t max_cap = 0;
    for (const Cut& c : cuts) {
        const StreamJob& j = h_jobs[c.stream];
        std::vector<uint64_t> at{0};
        for (uint32_t k = 1; k < c.nseg; ++k)
            if (sj[c.first_search + k - 1].found != ~0ull) at.push_back(sj[c.first_search + k - 1].found);
        first_of.push_back((uint32_t)sg.size());
        for (size_t k = 0; k < at.size(); ++k) {
            const uint64_t end = k + 1 < at.size() ? at[k + 1] : 8 * j.src_len;
            StreamJob s = j;
            s.start_bit = at[k];
            s.start_out = 0;
            s.phase = k == 0 ? 0 : 1;
            s.stop_bit = k + 1 < at.size() ? at[k + 1] : 0;
            s.symbolic = 1;
            s.dst_cap = (uint64_t)((double)j.dst_cap * 1.5 * (double)(end - at[k]) / (double)(8 * j.src_len)) + (64u << 10);
            s.dst = (uint8_t*)(uintptr_t)sym_total;  // offset for now (symbols)
            sym_total += align_up(s.dst_cap + 8, 128);
            max_cap = std::max(max_cap, s.dst_cap);
            sg.push_back(s);
            sg_stream.push_back(c.stream);
        }
    }
    first_of.push_back((uint32_t)sg.size());
    const size_t n = sg.size();
    CU(ctx->d_sgsym.reserve(sizeof(uint16_t) * sym_total));
    for (StreamJob& s : sg) s.dst = (uint8_t*)(ctx->d_sgsym.as<uint16_t>() + (size_t)(uintptr_t)s.dst);
    CU(ctx->h_sgjobs.reserve(sizeof(StreamJob) * n));
    CU(ctx->d_sgjobs.reserve(sizeof(StreamJob) * n));
    CU(ctx->d_sgres.reserve(sizeof(StreamResult) * n));
    CU(ctx->h_sgres.reserve(sizeof(StreamResult) * n));
    memcpy(ctx->h_sgjobs.p, sg.data(), sizeof(StreamJob) * n);
    CU(cudaMemcpyAsync(ctx->d_sgjobs.p, ctx->h_sgjobs.p, sizeof(StreamJob) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_sgres.p, 0, sizeof(StreamResult) * n, ctx->stream));
    {
        WvParams pp;
        pp.bitmap_words = ctx->cells_segments ? 8 : wv_bitmap_words(max_cap);
        pp.scratch_stride = ctx->cells_segments ? CL_SCRATCH : wv_scratch_stride(pp.bitmap_words);
        unsigned grid = (unsigned)std::min<size_t>(n, slots);
        size_t need = (size_t)pp.scratch_stride * grid + 256;
        if (need > ctx->d_scratch.cap || pp.scratch_stride != ctx->scratch_stride) {
            CU(ctx->d_scratch.reserve(need));
            CU(cudaMemsetAsync(ctx->d_scratch.p, 0, ctx->d_scratch.cap, ctx->stream));
            ctx->scratch_stride = pp.scratch_stride;
        }
        pp.ticket = (uint32_t*)((char*)ctx->d_scratch.p + (size_t)pp.scratch_stride * grid);
        CU(cudaMemsetAsync(pp.ticket, 0, sizeof(uint32_t), ctx->stream));
        pp.jobs = ctx->d_sgjobs.as<StreamJob>();
        pp.results = ctx->d_sgres.as<StreamResult>();
        pp.order = nullptr;
        pp.scratch = ctx->d_scratch.as<uint8_t>();
        pp.count = (int)n;
        if (ctx->cells_segments) inflate_cells_kernel<<<grid, WV_THREADS, sizeof(ClShared), ctx->stream>>>(pp);
        else inflate_wave_kernel<<<grid, WV_THREADS, sizeof(WvShared), ctx->stream>>>(pp);
        ctx->launches++;
    }
    CU(cudaMemcpyAsync(ctx->h_sgres.p, ctx->d_sgres.p, sizeof(StreamResult) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    // 3. a stream is accepted when every segment ended exactly where the next one starts
    const StreamResult* sr = ctx->h_sgres.as<StreamResult>();
    std::vector<SegmentRecord> recs;
    std::vector<uint32_t> stream_first{0};
    std::vector<uint64_t> chunk_base;
    std::vector<uint32_t> accepted;
    std::vector<StreamResult> finals;
    uint64_t chunks = 0;
    for (size_t c = 0; c < cuts.size(); ++c) {
        const uint32_t lo = first_of[c], hi = first_of[c + 1];
        const StreamJob& j = h_jobs[cuts[c].stream];
        bool ok = true;
        uint64_t total = 0, blocks = 0;
        for (uint32_t k = lo; k < hi && ok; ++k) {
            const bool last = k + 1 == hi;
            ok = sr[k].status == PNGB200_OK && (last ? sr[k].phase == 2 : (sr[k].phase == 1 && sr[k].consumed_bits == sg[k].stop_bit));
            total += sr[k].produced;
            blocks += sr[k].blocks;
        }
        if (ok && total > j.dst_cap) ok = false;
        ctx->seg_streams++;
        ctx->seg_segments += hi - lo;
        if (!ok) { ctx->seg_fallbacks++; continue; }
        uint64_t off = 0;
        for (uint32_t k = lo; k < hi; ++k) {
            SegmentRecord rec;
            rec.sym = (const uint16_t*)sg[k].dst;
            rec.out = j.dst + off;
            rec.produced = sr[k].produced;
            rec.stream = cuts[c].stream;
            rec.first = k == lo;
            chunk_base.push_back(chunks);
            chunks += (sr[k].produced + 4095) / 4096;
            off += sr[k].produced;
            recs.push_back(rec);
        }
        stream_first.push_back((uint32_t)recs.size());
        accepted.push_back(cuts[c].stream);
        StreamResult f = sr[hi - 1];   // trailer fields come from the last segment
        f.produced = total;
        f.blocks = (uint32_t)blocks;
        f.resume_out = total;
        f.ck_done = 0;
        f.stat_waves = 0;
        f.stat_tokens = f.stat_matches = f.stat_deferred = 0;
        for (int q = 0; q < 12; ++q) f.stat_cycles[q] = 0;
        for (uint32_t k = lo; k < hi; ++k) {
            f.stat_waves += sr[k].stat_waves;
            f.stat_tokens += sr[k].stat_tokens;
            f.stat_matches += sr[k].stat_matches;
            f.stat_deferred += sr[k].stat_deferred;
            for (int q = 0; q < 12; ++q) f.stat_cycles[q] += sr[k].stat_cycles[q];
        }
        finals.push_back(f);
    }
    if (!recs.empty()) {
        // 4. windows in front of the segments, then markers -> bytes at their final places
        const size_t nr = recs.size(), ns = stream_first.size() - 1;
        const size_t off_first = align_up(sizeof(SegmentRecord) * nr, 256);
        const size_t off_chunk = align_up(off_first + sizeof(uint32_t) * (ns + 1), 256);
        const size_t table = off_chunk + sizeof(uint64_t) * nr;
        CU(ctx->h_sgrec.reserve(table));
        CU(ctx->d_sgrec.reserve(table));
        memcpy(ctx->h_sgrec.p, recs.data(), sizeof(SegmentRecord) * nr);
        memcpy((char*)ctx->h_sgrec.p + off_first, stream_first.data(), sizeof(uint32_t) * (ns + 1));
        memcpy((char*)ctx->h_sgrec.p + off_chunk, chunk_base.data(), sizeof(uint64_t) * nr);
        CU(cudaMemcpyAsync(ctx->d_sgrec.p, ctx->h_sgrec.p, table, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx->d_sgwin.reserve((size_t)SEG_WINDOW * nr));
        const SegmentRecord* d_recs = ctx->d_sgrec.as<SegmentRecord>();
        window_propagate_kernel<<<(unsigned)ns, 256, 0, ctx->stream>>>(d_recs, (const uint32_t*)((char*)ctx->d_sgrec.p + off_first),
                                                                       (uint32_t)ns, ctx->d_sgwin.as<uint8_t>());
        if (chunks)
            marker_resolve_kernel<<<(unsigned)chunks, 256, 0, ctx->stream>>>(d_recs, (uint32_t)nr, ctx->d_sgwin.as<uint8_t>(),
                                                                             (const uint64_t*)((char*)ctx->d_sgrec.p + off_chunk));
        ctx->launches += 2;
        CU(cudaGetLastError());
        // the streams' result records (the whole-stream kernels will not touch them)
        StreamResult* d_results = ctx->d_results.as<StreamResult>();
        StreamResult* hf = ctx->h_sgres.as<StreamResult>();  // reuse: the segment results have been consumed
        for (size_t a = 0; a < accepted.size(); ++a) {
            hf[a] = finals[a];
            CU(cudaMemcpyAsync(d_results + accepted[a], hf + a, sizeof(StreamResult), cudaMemcpyHostToDevice, ctx->stream));
        }
        std::vector<uint32_t> rest;
        for (uint32_t i : par)
            if (std::find(accepted.begin(), accepted.end(), i) == accepted.end()) rest.push_back(i);
        par.swap(rest);
    }
    return PNGB200_OK;
}

// ---- inflate (+ checksum) over a device-resident job table ----
// h_jobs: host copy (for dst_cap based chunk layout); d_jobs/d_results device arrays of `count`.
int run_inflate(pngb200_ctx* ctx, const StreamJob* h_jobs, size_t count)
{
    ctx->last_engine = -1;
    StreamJob*    d_jobs    = ctx->d_jobs.as<StreamJob>();
    StreamResult* d_results = ctx->d_results.as<StreamResult>();
    ctx->seg_streams = ctx->seg_segments = ctx->seg_fallbacks = 0;
    CU(cudaMemsetAsync(d_results, 0, sizeof(StreamResult) * count, ctx->stream));
    auto before_first_launch = [&]() -> int {
        int rc = PNGB200_OK;
        if (ctx->bulk_h2d) {
            rc = ctx->bulk_h2d();
            ctx->bulk_h2d = nullptr;
        }
        if (rc == PNGB200_OK) CU(cudaEventRecord(ctx->ev[0], ctx->stream));
        return rc;
    };
    // big streams get a whole CTA each (block-parallel kernel); tiny ones a warp each
    {
        std::vector<uint32_t> par, ser;
        for (size_t i = 0; i < count; ++i) {
            bool big = h_jobs[i].src_len >= ctx->parallel_threshold;
            if (ctx->inflate_mode == 1) big = false;
            if (ctx->inflate_mode == 2) big = true;
            (big ? par : ser).push_back((uint32_t)i);
        }
        // longest first: the persistent CTAs pull streams from a ticket, so this is LPT scheduling
        std::stable_sort(par.begin(), par.end(),
                         [&](uint32_t a, uint32_t b) { return h_jobs[a].src_len > h_jobs[b].src_len; });
        CU(ctx->h_order.reserve(sizeof(uint32_t) * count));
        CU(ctx->d_order.reserve(sizeof(uint32_t) * count));
        uint32_t* ho = ctx->h_order.as<uint32_t>();
        std::copy(par.begin(), par.end(), ho);
        std::copy(ser.begin(), ser.end(), ho + par.size());
        CU(cudaMemcpyAsync(ctx->d_order.p, ho, sizeof(uint32_t) * count, cudaMemcpyHostToDevice, ctx->stream));
        const uint32_t* d_order = ctx->d_order.as<uint32_t>();
        bool hooked = false;
        if (ctx->inflate_mode == 0 || ctx->inflate_mode == 5) {
            // few big streams: cut them so that every CTA slot has something to decode
            const size_t before = par.size();
            if (!par.empty() && std::max(par.size(), ctx->peer_streams) * 2 <= (size_t)ctx->sm_count * WV_CTAS_PER_SM) {
                if (int rc = before_first_launch()) return rc;
                hooked = true;
                if (int rc = run_segments(ctx, h_jobs, par)) return rc;
            }
            if (par.size() != before) {   // the order table lists what is left for the whole-stream kernels
                std::copy(par.begin(), par.end(), ho);
                std::copy(ser.begin(), ser.end(), ho + par.size());
                CU(cudaMemcpyAsync(ctx->d_order.p, ho, sizeof(uint32_t) * (par.size() + ser.size()), cudaMemcpyHostToDevice, ctx->stream));
            }
        }
        if (!par.empty()) {
            uint64_t max_cap = 0;
            for (uint32_t i : par) max_cap = std::max<uint64_t>(max_cap, h_jobs[i].dst_cap);
            // Two engines for the big streams, same decomposition (8 KiB waves of 256 subsequences), chosen by
            // how many streams there are to keep the SMs busy:
            //  * inflate_wave_kernel: LZ77 window in a 64 KiB shared-memory ring, Adler-32 folded into the store;
            //    2 CTAs per SM.  Fastest per stream (146 K cycles per wave against 234 K), so it takes every batch
            //    that fits its 2 x SMs CTA slots -- and all multi-CTA-per-stream work.
            //  * inflate_parallel_kernel (round 1): 16 KiB output image, window read back from HBM/L2, 4 CTAs per
            //    SM.  Slower per stream, but twice the streams in flight hide its barrier phases: measured r02 on
            //    8K RGBA8 1.96 ms per image against 2.51 ms once a batch exceeds the wave kernel's slots.
            //  * inflate_cells_kernel (round 2, second half): the same waves, but the LZ77 half works on 16-bit cells in
            //    shared memory resolved by pointer jumping; no window in shared memory, 3 CTAs per SM.  Waves that
            //    expand beyond ~16 KB are cut, so it is meant for streams that expand less than ~2x per wave
            //    (photographic PNG data); flat graphics stay with the ring kernel.
            const size_t wave_slots = (size_t)ctx->sm_count * WV_CTAS_PER_SM;
            const size_t streams_in_flight = std::max(par.size(), ctx->peer_streams);   // (lanes: the chunks of a host batch run side by side)
            enum { ENG_PARALLEL = 0, ENG_WAVE = 1, ENG_CELLS = 2 };
            int engine = streams_in_flight <= wave_slots ? ENG_WAVE : ENG_PARALLEL;
            if (ctx->inflate_mode == 0 || ctx->inflate_mode == 5) {
                uint64_t in_bytes = 0, out_bytes = 0;
                for (uint32_t i : par) { in_bytes += h_jobs[i].src_len; out_bytes += h_jobs[i].dst_cap; }
                if (ctx->cells_auto && out_bytes <= in_bytes * 9 / 4) engine = ENG_CELLS;
            }
            if (ctx->inflate_mode == 3) engine = ENG_WAVE;
            if (ctx->inflate_mode == 4) engine = ENG_PARALLEL;
            if (ctx->inflate_mode == 6) engine = ENG_CELLS;
            const bool use_wave = engine == ENG_WAVE;
            ctx->last_engine = engine;
            const uint64_t bitmap_words = engine == ENG_CELLS ? 8 : use_wave ? wv_bitmap_words(max_cap) : par_bitmap_words(max_cap);
            const uint64_t stride = engine == ENG_CELLS ? CL_SCRATCH : use_wave ? wv_scratch_stride(bitmap_words) : par_scratch_stride(bitmap_words);
            unsigned grid = (unsigned)std::min<size_t>(par.size(), engine == ENG_CELLS ? (size_t)ctx->sm_count * CL_CTAS_PER_SM
                                                                   : use_wave         ? wave_slots
                                                                                      : (size_t)ctx->sm_count * PAR_CTAS_PER_SM);
            size_t need = (size_t)stride * grid + 256;
            // The per-CTA "unresolved" bitmaps must be all-zero when a launch starts; the kernels
            // leave them clean.  A different stride moves the bitmaps onto bytes that held copy
            // lists before, and a fresh allocation is garbage: zero the whole arena in both cases.
            if (need > ctx->d_scratch.cap || stride != ctx->scratch_stride) {
                CU(ctx->d_scratch.reserve(need));
                CU(cudaMemsetAsync(ctx->d_scratch.p, 0, ctx->d_scratch.cap, ctx->stream));
                ctx->scratch_stride = stride;
            }
            uint32_t* ticket = (uint32_t*)((char*)ctx->d_scratch.p + (size_t)stride * grid);
            CU(cudaMemsetAsync(ticket, 0, sizeof(uint32_t), ctx->stream));
            if (!hooked)
                if (int rc = before_first_launch()) return rc;
            hooked = true;
            if (engine == ENG_CELLS) {
                WvParams pp{};
                pp.scratch = ctx->d_scratch.as<uint8_t>();
                pp.scratch_stride = stride;
                pp.ticket = ticket;
                pp.jobs = d_jobs;
                pp.results = d_results;
                pp.order = d_order;
                pp.count = (int)par.size();
                inflate_cells_kernel<<<grid, WV_THREADS, sizeof(ClShared), ctx->stream>>>(pp);
            } else if (use_wave) {
                WvParams pp;
                pp.bitmap_words = bitmap_words;
                pp.scratch_stride = stride;
                pp.ticket = ticket;
                pp.jobs = d_jobs;
                pp.results = d_results;
                pp.order = d_order;
                pp.scratch = ctx->d_scratch.as<uint8_t>();
                pp.count = (int)par.size();
                inflate_wave_kernel<<<grid, WV_THREADS, sizeof(WvShared), ctx->stream>>>(pp);
            } else {
                ParParams pp;
                pp.bitmap_words = bitmap_words;
                pp.scratch_stride = stride;
                pp.ticket = ticket;
                pp.jobs = d_jobs;
                pp.results = d_results;
                pp.order = d_order;
                pp.scratch = ctx->d_scratch.as<uint8_t>();
                pp.count = (int)par.size();
                if (par.size() <= (size_t)ctx->sm_count * 3)
                    inflate_parallel_kernel3<<<grid, PAR_THREADS, sizeof(ParShared), ctx->stream>>>(pp);
                else
                    inflate_parallel_kernel<<<grid, PAR_THREADS, sizeof(ParShared), ctx->stream>>>(pp);
            }
            ctx->launches++;
        }
        if (!hooked)
            if (int rc = before_first_launch()) return rc;
        if (!ser.empty()) {
            inflate_serial_kernel<<<(unsigned)ser.size(), 32, 0, ctx->stream>>>(d_jobs, d_results, d_order + par.size(),
                                                                                 (int)ser.size());
            ctx->launches++;
        }
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    // checksum: chunk layout from dst_cap (an upper bound of `produced`)
    CU(ctx->h_misc.reserve(sizeof(uint32_t) * (count + 1)));
    uint32_t* base = ctx->h_misc.as<uint32_t>();
    uint64_t  total = 0;
    for (size_t i = 0; i < count; ++i) {
        base[i] = (uint32_t)total;
        total += (h_jobs[i].dst_cap + CK_CHUNK - 1) / CK_CHUNK;
    }
    base[count] = (uint32_t)total;
    if (total >= (1ull << 31)) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "batch too large");
    CU(ctx->d_misc.reserve(sizeof(uint32_t) * (count + 1)));
    CU(ctx->d_partial.reserve(sizeof(uint64_t) * 2 * std::max<uint64_t>(total, 1)));
    CU(cudaMemcpyAsync(ctx->d_misc.p, base, sizeof(uint32_t) * (count + 1), cudaMemcpyHostToDevice,
                       ctx->stream));
    ChecksumParams cp;
    cp.jobs = d_jobs;
    cp.results = d_results;
    cp.chunk_base = ctx->d_misc.as<uint32_t>();
    cp.partial = ctx->d_partial.as<uint64_t>();
    cp.count = (uint32_t)count;
    cp.total_chunks = (uint32_t)total;
    cp.crc_tables = nullptr;
    for (size_t i = 0; i < count; ++i)
        if (h_jobs[i].format == PNGB200_FORMAT_GZIP) {
            if (int rc = ensure_crc_tables(ctx)) return rc;
            cp.crc_tables = ctx->d_crctab.as<uint32_t>();
            break;
        }
    if (total) {
        checksum_chunk_kernel<<<(unsigned)total, CK_THREADS, 0, ctx->stream>>>(cp);
        ctx->launches++;
    }
    checksum_fold_kernel<<<(unsigned)count, 32, 0, ctx->stream>>>(cp);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    return PNGB200_OK;
}

struct Geometry {
    uint64_t filtered;  // total filtered bytes
    uint64_t storage;
    uint32_t pitch;     // non-interlaced pitch
    uint8_t  bpp;
    bool     fast;      // eligible for the wavefront kernel
};

bool geometry(uint32_t w, uint32_t h, int volume, int depth, int interlaced, Geometry* g)
{
    if (w == 0 || h == 0 || volume <= 0 || volume > 64 || depth <= 0 || depth > 16) return false;
    // Dimensions come from untrusted files: the reference traps when w * h * bpp overflows (PNG.Image.swift:84);
    // here an image whose sizes do not fit is refused before any buffer is sized from a wrapped product.
    {
        if (w > 0x7fffffffu || h > 0x7fffffffu) return false;
        const uint64_t pitch64 = ((uint64_t)w * (uint64_t)volume + 7) >> 3;
        uint64_t prod;
        if (pitch64 > 0xfffffff0ull) return false;
        if (__builtin_mul_overflow((uint64_t)h, pitch64 + 1, &prod) || prod > (1ull << 46)) return false;
        if (__builtin_mul_overflow((uint64_t)w * (uint64_t)h, (uint64_t)((volume + 7) >> 3), &prod) || prod > (1ull << 46)) return false;
    }
    g->filtered = pngb200_filtered_size(w, h, volume, interlaced);
    g->storage  = pngb200_storage_size(w, h, volume);
    g->pitch    = (uint32_t)(((uint64_t)w * volume + 7) >> 3);
    g->bpp      = (uint8_t)((volume + 7) >> 3);
    g->fast     = !interlaced && depth >= 8 &&
              (g->bpp == 1 || g->bpp == 2 || g->bpp == 3 || g->bpp == 4 || g->bpp == 6 || g->bpp == 8);
    return true;
}

// unfilter stage over device-resident filtered streams
struct UnfilterItem {
    const uint8_t*      filtered;
    uint8_t*            filtered_mut;  // same buffer when the library owns it, else null
    uint8_t*            pixels;
    const StreamResult* inflated;
    uint64_t            filtered_len;
    uint32_t            w, h;
    uint8_t             volume, depth, interlaced;
    Geometry            g;
};

int run_unfilter(pngb200_ctx* ctx, const std::vector<UnfilterItem>& items)
{
    ctx->d_hist = nullptr;

    std::vector<ImageJob>   fast;
    std::vector<GenericJob> slow;
    std::vector<uint32_t>   band_base;
    uint64_t                bands = 0;
    for (const UnfilterItem& it : items) {
        if (it.g.fast) {
            ImageJob j;
            j.filtered = it.filtered;
            j.pixels = it.pixels;
            j.inflated = it.inflated;
            j.filtered_len = it.filtered_len;
            j.width = it.w;
            j.height = it.h;
            j.pitch = it.g.pitch;
            j.volume = it.volume;
            j.depth = it.depth;
            j.interlaced = 0;
            j.bpp = it.g.bpp;
            fast.push_back(j);
            band_base.push_back((uint32_t)bands);
            bands += (it.h + 31) / 32;
        } else {
            GenericJob j;
            j.filtered = it.filtered_mut;
            j.pixels = it.pixels;
            j.inflated = it.inflated;
            j.filtered_len = it.filtered_len;
            j.width = it.w;
            j.height = it.h;
            j.volume = it.volume;
            j.depth = it.depth;
            j.interlaced = it.interlaced;
            j.bpp = it.g.bpp;
            slow.push_back(j);
        }
    }
    band_base.push_back((uint32_t)bands);
    if (bands >= (1ull << 31)) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "batch too large");
    if (!fast.empty()) {
        // Tickets go out band level by band level (see WaveParams): jobs sorted by band count, descending; level_start[b] =
        // tickets in front of level b.  (Images of more than 4096 bands -- 131072 rows -- keep the image-major order.)
        std::vector<uint32_t> level_start;
        {
            std::vector<uint32_t> idx(fast.size());
            for (size_t i = 0; i < idx.size(); ++i) idx[i] = (uint32_t)i;
            auto nb = [&](uint32_t i) { return (fast[i].height + 31) / 32; };
            std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return nb(a) > nb(b); });
            const uint32_t maxb = nb(idx[0]);
#ifndef PNGB200_WAVE_IMAGE_MAJOR   // (tuning builds: the round-2a order)
            if (maxb <= 4096) {
                std::vector<ImageJob> sorted(fast.size());
                for (size_t i = 0; i < idx.size(); ++i) sorted[i] = fast[idx[i]];
                fast.swap(sorted);
                uint64_t at = 0;
                for (size_t i = 0; i < fast.size(); ++i) {
                    band_base[i] = (uint32_t)at;
                    at += (fast[i].height + 31) / 32;
                }
                level_start.assign(maxb + 1, 0);
                size_t alive = fast.size();          // images with more than b bands: a prefix of the sorted list
                for (uint32_t b = 0; b < maxb; ++b) {
                    while (alive && (fast[alive - 1].height + 31) / 32 <= b) --alive;
                    level_start[b + 1] = level_start[b] + (uint32_t)alive;
                }
            }
#else
            (void)maxb;
#endif
        }
        size_t jb = sizeof(ImageJob) * fast.size(), bb = sizeof(uint32_t) * band_base.size();
        size_t lb = sizeof(uint32_t) * level_start.size();
        size_t off_bb = align_up(jb, 256), off_ls = align_up(off_bb + bb, 256), off_pr = align_up(off_ls + lb, 256);
        size_t total = align_up(off_pr + sizeof(uint32_t) * (bands + 1), 8) + 8 * sizeof(unsigned long long);
        CU(ctx->h_imgjobs.reserve(off_pr));
        CU(ctx->d_imgjobs.reserve(total));
        memcpy(ctx->h_imgjobs.p, fast.data(), jb);
        memcpy((char*)ctx->h_imgjobs.p + off_bb, band_base.data(), bb);
        if (lb) memcpy((char*)ctx->h_imgjobs.p + off_ls, level_start.data(), lb);
        CU(cudaMemcpyAsync(ctx->d_imgjobs.p, ctx->h_imgjobs.p, off_pr, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync((char*)ctx->d_imgjobs.p + off_pr, 0, sizeof(uint32_t) * (bands + 1), ctx->stream));
        WaveParams p;
        p.jobs = ctx->d_imgjobs.as<ImageJob>();
        p.band_base = (const uint32_t*)((char*)ctx->d_imgjobs.p + off_bb);
        p.progress = (uint32_t*)((char*)ctx->d_imgjobs.p + off_pr);
        p.ticket = p.progress + bands;
        p.hist = (unsigned long long*)((char*)ctx->d_imgjobs.p + align_up(off_pr + sizeof(uint32_t) * (bands + 1), 8));
        CU(cudaMemsetAsync(p.hist, 0, 8 * sizeof(unsigned long long), ctx->stream));
        ctx->d_hist = p.hist;
        p.njobs = (uint32_t)fast.size();
        p.total_bands = (uint32_t)bands;
        p.level_start = (const uint32_t*)((char*)ctx->d_imgjobs.p + off_ls);
        p.levels = level_start.empty() ? 0u : (uint32_t)level_start.size() - 1;
        unsigned grid = (unsigned)std::min<uint64_t>((bands + WAVE_WARPS - 1) / WAVE_WARPS,
                                                     (uint64_t)ctx->sm_count * 8);
        unfilter_wave_kernel<<<grid, WAVE_WARPS * 32, WAVE_SMEM, ctx->stream>>>(p);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    if (!slow.empty()) {
        size_t jb = sizeof(GenericJob) * slow.size();
        CU(ctx->h_genjobs.reserve(jb));
        CU(ctx->d_genjobs.reserve(jb));
        memcpy(ctx->h_genjobs.p, slow.data(), jb);
        CU(cudaMemcpyAsync(ctx->d_genjobs.p, ctx->h_genjobs.p, jb, cudaMemcpyHostToDevice, ctx->stream));
        unfilter_generic_kernel<<<(unsigned)slow.size(), 128, 0, ctx->stream>>>(
            ctx->d_genjobs.as<GenericJob>(), (int)slow.size());
        ctx->launches++;
        CU(cudaGetLastError());
    }
    return PNGB200_OK;
}


// Host-memory batches with a lot of bytes to move are cut into chunks that four lanes (helper
// contexts on the same GPU, one host thread each) work through round-robin, so that one chunk's
// H2D / D2H copies overlap another chunk's kernels.  Results are identical: images are
// independent units.  Chunk size is a trade: the inflate kernel runs one CTA per stream, so small
// chunks leave SMs idle, while few chunks leave nothing to overlap.  Measured on B200 (r01,
// GPixels/s end to end, lanes x chunks per lane): 444 x 8K RGBA8 4x1 7.1, 3x1 6.9, 2x1 6.4, 4x2 5.9,
// 4x4 4.1; 1184 x 1080p 4x2 8.8, 3x2 8.0, 4x3 8.2, 4x4 7.2, 4x1 5.7, 1x1 5.6 -- i.e. about one stream
// per SM in every chunk, chunk count a multiple of the lane count.
// `bytes_of(i)`: bytes item i moves over PCIe; `work(lane, lo, n)`: process items [lo, lo + n) on `lane`.
template <typename BytesOf, typename Work>
int run_over_lanes(pngb200_ctx* ctx, size_t count, int memspace, BytesOf bytes_of, Work work)
{
    for (pngb200_ctx* lane : ctx->lanes) lane->d_hist = nullptr;   // counters describe the batch that starts now
    ctx->d_hist = nullptr;
    size_t bytes = 0;
    for (size_t i = 0; i < count; ++i) bytes += bytes_of(i);
    // tunable for experiments: PNGB200_LANES, PNGB200_CHUNKS_PER_LANE (0 / unset = the rule above); read once
    static const size_t kLanes = getenv("PNGB200_LANES") ? std::max(1, atoi(getenv("PNGB200_LANES"))) : 4;
    static const size_t kPerLane = getenv("PNGB200_CHUNKS_PER_LANE") ? std::max(0, atoi(getenv("PNGB200_CHUNKS_PER_LANE"))) : 0;
    constexpr size_t kMinChunk = 32;
    if (memspace != PNGB200_MEM_HOST || count < 2 * kMinChunk || bytes < ((size_t)256 << 20)) return work(ctx, 0, count);
    while (ctx->lanes.size() < kLanes) {
        pngb200_ctx* lane = pngb200_ctx_create(ctx->device);
        if (!lane) return set_error(ctx, PNGB200_ERR_CUDA, "cannot create a pipeline lane: %s", g_last_error.c_str());
        ctx->lanes.push_back(lane);
    }
    size_t nchunks;
    if (kPerLane) {
        nchunks = kPerLane * kLanes;
    } else {
        nchunks = std::max<size_t>(1, count / (size_t)ctx->sm_count);
        nchunks = (nchunks + kLanes - 1) / kLanes * kLanes;
    }
    nchunks = std::max<size_t>(1, std::min(nchunks, count / kMinChunk));
    std::vector<size_t> cut(nchunks + 1);  // chunk boundaries balanced by bytes
    {
        size_t acc = 0, k = 1;
        cut[0] = 0;
        for (size_t i = 0; i < count && k < nchunks; ++i) {
            acc += bytes_of(i);
            if (acc * nchunks >= bytes * k) cut[k++] = i + 1;
        }
        while (k <= nchunks) cut[k++] = count;
    }
    std::vector<int> rcs(kLanes, PNGB200_OK);
    std::vector<std::thread> workers;
    for (size_t l = 0; l < kLanes; ++l)
        workers.emplace_back([&, l]() {
            pngb200_ctx* lane = ctx->lanes[l];
            lane->inflate_mode = ctx->inflate_mode;
            lane->cells_auto = ctx->cells_auto;
            lane->cells_segments = ctx->cells_segments;
            lane->parallel_threshold = ctx->parallel_threshold;
            lane->peer_streams = count;
            for (size_t c = l; c < nchunks; c += kLanes) {
                size_t lo = cut[c], n = cut[c + 1] - cut[c];
                if (n == 0) continue;
                int rc = work(lane, lo, n);
                if (rc != PNGB200_OK) { rcs[l] = rc; return; }   // the lane keeps its own error text
            }
        });
    for (std::thread& t : workers) t.join();
    for (size_t l = 0; l < kLanes; ++l)
        if (rcs[l] != PNGB200_OK) {
            ctx->error = ctx->lanes[l]->error;   // after the join: one writer
            return rcs[l];
        }
    return PNGB200_OK;
}
}  // namespace

// ================================ C ABI ================================

extern "C" {

size_t pngb200_filtered_size(uint32_t w, uint32_t h, int volume, int interlaced)
{
    if (!interlaced) return (size_t)h * ((((size_t)w * (size_t)volume + 7) >> 3) + 1);
    size_t total = 0;
    for (int z = 0; z < 7; ++z) {
        size_t sx = ((size_t)w + (1u << ADAM7[z][2]) - ADAM7[z][0] - 1) >> ADAM7[z][2];
        size_t sy = ((size_t)h + (1u << ADAM7[z][3]) - ADAM7[z][1] - 1) >> ADAM7[z][3];
        if (sx == 0 || sy == 0) continue;
        total += sy * (((sx * (size_t)volume + 7) >> 3) + 1);
    }
    return total;
}

size_t pngb200_storage_size(uint32_t w, uint32_t h, int volume)
{
    return (size_t)w * (size_t)h * (size_t)((volume + 7) >> 3);
}

pngb200_ctx* pngb200_ctx_create(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error(nullptr, PNGB200_ERR_CUDA, "no CUDA device: %s (there is no CPU fallback)",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return nullptr;
    }
    if (device < 0) cudaGetDevice(&device);
    if (device >= n) {
        set_error(nullptr, PNGB200_ERR_BAD_ARGUMENT, "device %d out of range (%d devices)", device, n);
        return nullptr;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
        set_error(nullptr, PNGB200_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
                  device, prop.major, prop.minor);
        return nullptr;
    }
    pngb200_ctx* ctx = new pngb200_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (const char* v = getenv("PNGB200_CELLS_AUTO")) ctx->cells_auto = atoi(v) != 0;   // tuning overrides, read once per context
    if (const char* v = getenv("PNGB200_CELLS_SEGMENTS")) ctx->cells_segments = atoi(v) != 0;
    DeviceGuard guard(device);
    if (cudaFuncSetAttribute(deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DfShared)) != cudaSuccess) {
        set_error(nullptr, PNGB200_ERR_CUDA, "cannot opt in to %zu bytes of shared memory", sizeof(DfShared));
        delete ctx;
        return nullptr;
    }
    if (configure_inflate_parallel() != 0) {
        set_error(nullptr, PNGB200_ERR_CUDA, "cannot opt in to %zu bytes of shared memory", sizeof(ParShared));
        delete ctx;
        return nullptr;
    }
    if (configure_inflate_wave() != 0) {
        set_error(nullptr, PNGB200_ERR_CUDA, "cannot opt in to %zu bytes of shared memory", sizeof(WvShared));
        delete ctx;
        return nullptr;
    }
    if (cudaFuncSetAttribute(unfilter_wave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WAVE_SMEM) != cudaSuccess) {
        set_error(nullptr, PNGB200_ERR_CUDA, "cannot opt in to %zu bytes of shared memory", (size_t)WAVE_SMEM);
        delete ctx;
        return nullptr;
    }
    if (configure_inflate_cells() != 0) {
        set_error(nullptr, PNGB200_ERR_CUDA, "cannot opt in to %zu bytes of shared memory", sizeof(ClShared));
        delete ctx;
        return nullptr;
    }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error(nullptr, PNGB200_ERR_CUDA, "cudaStreamCreate failed");
        delete ctx;
        return nullptr;
    }
    for (cudaEvent_t& e : ctx->ev) cudaEventCreate(&e);
    return ctx;
}

void pngb200_ctx_destroy(pngb200_ctx* ctx)
{
    if (!ctx) return;
    for (pngb200_ctx* lane : ctx->lanes) pngb200_ctx_destroy(lane);
    ctx->lanes.clear();
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (DevBuf* b : {&ctx->d_jobs, &ctx->d_results, &ctx->d_imgjobs, &ctx->d_genjobs, &ctx->d_misc,
                      &ctx->d_partial, &ctx->d_filtered, &ctx->d_in, &ctx->d_out, &ctx->d_order, &ctx->d_scratch, &ctx->d_dfscratch,
                      &ctx->d_dfjobs, &ctx->d_dfres, &ctx->d_enc, &ctx->d_file, &ctx->d_crc, &ctx->d_seg, &ctx->d_crctab,
                      &ctx->d_sgjobs, &ctx->d_sgres, &ctx->d_sgsym, &ctx->d_sgsearch, &ctx->d_sgrec, &ctx->d_sgwin})
        b->release();
    for (PinBuf* b : {&ctx->h_jobs, &ctx->h_results, &ctx->h_imgjobs, &ctx->h_genjobs, &ctx->h_misc, &ctx->h_order, &ctx->h_crc, &ctx->h_seg,
                      &ctx->h_sgsearch, &ctx->h_sgjobs, &ctx->h_sgres, &ctx->h_sgrec})
        b->release();
    for (cudaEvent_t e : ctx->ev) if (e) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int pngb200_ctx_trim(pngb200_ctx* ctx)
{
    if (!ctx) return PNGB200_ERR_BAD_ARGUMENT;
    if (ctx->pending) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "a decode batch is pending");
    for (pngb200_ctx* lane : ctx->lanes) pngb200_ctx_trim(lane);
    DeviceGuard guard(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    for (DevBuf* b : {&ctx->d_partial, &ctx->d_filtered, &ctx->d_in, &ctx->d_out, &ctx->d_scratch, &ctx->d_dfscratch, &ctx->d_enc, &ctx->d_file, &ctx->d_sgsym, &ctx->d_sgwin})
        b->release();
    ctx->scratch_stride = 0;
    return PNGB200_OK;
}

const char* pngb200_last_error(const pngb200_ctx* ctx) { return ctx ? ctx->error.c_str() : g_last_error.c_str(); }
void*       pngb200_ctx_stream(pngb200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int         pngb200_ctx_device(const pngb200_ctx* ctx) { return ctx ? ctx->device : -1; }
uint64_t    pngb200_ctx_launch_count(const pngb200_ctx* ctx)
{
    if (!ctx) return 0;
    uint64_t n = ctx->launches;
    for (const pngb200_ctx* lane : ctx->lanes) n += lane->launches;
    return n;
}
void        pngb200_ctx_set_inflate_mode(pngb200_ctx* ctx, int mode) { if (ctx) ctx->inflate_mode = mode; }

int pngb200_ctx_inflate_stats(pngb200_ctx* ctx, size_t count, uint64_t out[4])
{
    if (!ctx || !out || ctx->h_results.cap < sizeof(StreamResult) * count) return PNGB200_ERR_BAD_ARGUMENT;
    const StreamResult* r = ctx->h_results.as<StreamResult>();
    out[0] = out[1] = out[2] = out[3] = 0;
    for (size_t i = 0; i < count; ++i) {
        out[0] += r[i].stat_waves;
        out[1] += r[i].stat_sync_rounds;
        out[2] += r[i].stat_resolve_rounds;
        out[3] += r[i].stat_fallback;
    }
    return PNGB200_OK;
}

int pngb200_ctx_inflate_counters(pngb200_ctx* ctx, size_t count, uint64_t out[24])
{
    if (!ctx || !out || ctx->h_results.cap < sizeof(StreamResult) * count) return PNGB200_ERR_BAD_ARGUMENT;
    const StreamResult* r = ctx->h_results.as<StreamResult>();
    for (int k = 0; k < 24; ++k) out[k] = 0;
    for (size_t i = 0; i < count; ++i) {
        out[0] += r[i].stat_waves;
        out[1] += r[i].stat_sync_rounds;
        out[2] += r[i].stat_resolve_rounds;
        out[3] += r[i].stat_fallback;
        out[4] += r[i].stat_tokens;
        out[5] += r[i].stat_matches;
        out[6] += r[i].stat_deferred;
        out[7] += r[i].blocks;
        for (int k = 0; k < 12; ++k) out[8 + k] += r[i].stat_cycles[k];
    }
    return PNGB200_OK;
}

int pngb200_ctx_filter_histogram(pngb200_ctx* ctx, uint64_t out[6])
{
    if (!ctx || !out) return PNGB200_ERR_BAD_ARGUMENT;
    for (int k = 0; k < 6; ++k) out[k] = 0;
    DeviceGuard guard(ctx->device);
    std::vector<pngb200_ctx*> all{ctx};
    all.insert(all.end(), ctx->lanes.begin(), ctx->lanes.end());
    for (pngb200_ctx* c : all) {
        if (!c->d_hist) continue;
        unsigned long long h[6];
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaMemcpy(h, c->d_hist, sizeof h, cudaMemcpyDeviceToHost));
        for (int k = 0; k < 6; ++k) out[k] += h[k];
    }
    return PNGB200_OK;
}

int pngb200_ctx_last_inflate_engine(pngb200_ctx* ctx) { return ctx ? ctx->last_engine : -1; }

int pngb200_ctx_segment_stats(pngb200_ctx* ctx, uint64_t out[3])
{
    if (!ctx || !out) return PNGB200_ERR_BAD_ARGUMENT;
    out[0] = ctx->seg_streams;
    out[1] = ctx->seg_segments;
    out[2] = ctx->seg_fallbacks;
    return PNGB200_OK;
}

int pngb200_ctx_stage_ms(pngb200_ctx* ctx, float ms[3])
{
    if (!ctx || !ms) return PNGB200_ERR_BAD_ARGUMENT;
    DeviceGuard guard(ctx->device);
    for (int i = 0; i < 3; ++i)
        if (cudaEventElapsedTime(&ms[i], ctx->ev[i], ctx->ev[i + 1]) != cudaSuccess) {
            cudaGetLastError();
            return set_error(ctx, PNGB200_ERR_CUDA, "stage events not recorded yet");
        }
    return PNGB200_OK;
}

// ---------------- standalone inflate ----------------
int pngb200_inflate_batch(pngb200_ctx* ctx, pngb200_stream_desc* s, size_t count, int memspace)
{
    if (!ctx || (!s && count)) return PNGB200_ERR_BAD_ARGUMENT;
    if (ctx->pending) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "a decode batch is pending");
    if (count == 0) return PNGB200_OK;
    DeviceGuard guard(ctx->device);
    CU(ctx->h_jobs.reserve(sizeof(StreamJob) * count));
    CU(ctx->d_jobs.reserve(sizeof(StreamJob) * count));
    CU(ctx->d_results.reserve(sizeof(StreamResult) * count));
    CU(ctx->h_results.reserve(sizeof(StreamResult) * count));
    StreamJob* jobs = ctx->h_jobs.as<StreamJob>();
    std::vector<size_t> in_off(count), out_off(count);
    size_t in_total = 0, out_total = 0;
    for (size_t i = 0; i < count; ++i) {
        if ((!s[i].src && s[i].src_len) || (!s[i].dst && s[i].dst_cap) || s[i].format < 0 || s[i].format > 2)
            return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "stream %zu: bad descriptor", i);
        in_off[i] = in_total;
        out_off[i] = out_total;
        in_total += align_up(s[i].src_len + 16, 256);
        out_total += align_up(s[i].dst_cap + 16, 256);
    }
    if (memspace == PNGB200_MEM_HOST) {
        CU(ctx->d_in.reserve(in_total));
        CU(ctx->d_out.reserve(out_total));
        for (size_t i = 0; i < count; ++i)
            if (s[i].src_len)
                CU(cudaMemcpyAsync(ctx->d_in.as<uint8_t>() + in_off[i], s[i].src, s[i].src_len,
                                   cudaMemcpyHostToDevice, ctx->stream));
    }
    for (size_t i = 0; i < count; ++i) {
        bool host = memspace == PNGB200_MEM_HOST;
        jobs[i].src = host ? ctx->d_in.as<uint8_t>() + in_off[i] : s[i].src;
        jobs[i].src_len = s[i].src_len;
        jobs[i].dst = host ? ctx->d_out.as<uint8_t>() + out_off[i] : s[i].dst;
        jobs[i].dst_cap = s[i].dst_cap;
        jobs[i].start_bit = 0;
        jobs[i].start_out = 0;
        jobs[i].format = s[i].format;
        jobs[i].phase = 0;
        jobs[i].stop_bit = 0;
        jobs[i].symbolic = 0;
        jobs[i].pad_ = 0;
    }
    CU(cudaMemcpyAsync(ctx->d_jobs.p, jobs, sizeof(StreamJob) * count, cudaMemcpyHostToDevice, ctx->stream));
    int rc = run_inflate(ctx, jobs, count);
    if (rc != PNGB200_OK) return rc;
    CU(cudaMemcpyAsync(ctx->h_results.p, ctx->d_results.p, sizeof(StreamResult) * count,
                       cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const StreamResult* r = ctx->h_results.as<StreamResult>();
    for (size_t i = 0; i < count; ++i) {
        s[i].status = r[i].status;
        s[i].err_a = r[i].err_a;
        s[i].err_b = r[i].err_b;
        s[i].checksum = r[i].checksum;
        s[i].blocks = r[i].blocks;
        s[i].produced = r[i].produced;
        s[i].consumed_bits = r[i].consumed_bits;
        if (memspace == PNGB200_MEM_HOST && r[i].produced)
            CU(cudaMemcpyAsync(s[i].dst, ctx->d_out.as<uint8_t>() + out_off[i], r[i].produced,
                               cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return PNGB200_OK;
}

// ---------------- PNG decode: inflate + unfilter ----------------
int pngb200_decode_batch_enqueue(pngb200_ctx* ctx, pngb200_image_desc* im, size_t count, int memspace)
{
    if (!ctx || (!im && count)) return PNGB200_ERR_BAD_ARGUMENT;
    if (ctx->pending) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "a decode batch is already pending");
    if (count == 0) return PNGB200_OK;
    DeviceGuard guard(ctx->device);
    const bool host = memspace == PNGB200_MEM_HOST;
    std::vector<Geometry> geo(count);
    std::vector<size_t>   f_off(count), in_off(count);
    ctx->expected.assign(count, 0);
    ctx->out_offset.assign(count, 0);
    ctx->out_bytes.assign(count, 0);
    size_t f_total = 0, in_total = 0, out_total = 0;
    for (size_t i = 0; i < count; ++i) {
        if (!geometry(im[i].width, im[i].height, im[i].volume, im[i].depth, im[i].interlaced, &geo[i]) ||
            (!im[i].idat && im[i].idat_len) || !im[i].pixels || im[i].format > 1)
            return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "image %zu: bad descriptor", i);
        if (im[i].pixels_cap < geo[i].storage)
            return set_error(ctx, PNGB200_ERR_OUTPUT_CAPACITY, "image %zu: pixels_cap %zu < %llu", i,
                             im[i].pixels_cap, (unsigned long long)geo[i].storage);
        ctx->expected[i] = geo[i].filtered;
        f_off[i] = f_total;
        f_total += align_up(geo[i].filtered + 64, 256);
        in_off[i] = in_total;
        in_total += align_up(im[i].idat_len + 16, 256);
        ctx->out_offset[i] = out_total;
        ctx->out_bytes[i] = geo[i].storage;
        out_total += align_up(geo[i].storage + 16, 256);
    }
    CU(ctx->d_filtered.reserve(f_total));
    CU(ctx->h_jobs.reserve(sizeof(StreamJob) * count));
    CU(ctx->d_jobs.reserve(sizeof(StreamJob) * count));
    CU(ctx->d_results.reserve(sizeof(StreamResult) * count));
    CU(ctx->h_results.reserve(sizeof(StreamResult) * count));
    if (host) {
        CU(ctx->d_in.reserve(in_total));
        CU(ctx->d_out.reserve(out_total));
        ctx->bulk_h2d = [ctx, im, count, &in_off]() -> int {  // runs inside run_inflate below, after its tables
            for (size_t i = 0; i < count; ++i)
                if (im[i].idat_len)
                    CU(cudaMemcpyAsync(ctx->d_in.as<uint8_t>() + in_off[i], im[i].idat, im[i].idat_len,
                                       cudaMemcpyHostToDevice, ctx->stream));
            return PNGB200_OK;
        };
    }
    StreamJob* jobs = ctx->h_jobs.as<StreamJob>();
    for (size_t i = 0; i < count; ++i) {
        jobs[i].src = host ? ctx->d_in.as<uint8_t>() + in_off[i] : im[i].idat;
        jobs[i].src_len = im[i].idat_len;
        jobs[i].dst = ctx->d_filtered.as<uint8_t>() + f_off[i];
        jobs[i].dst_cap = geo[i].filtered + 16;  // room to notice extraneous image data
        jobs[i].start_bit = 0;
        jobs[i].start_out = 0;
        jobs[i].format = im[i].format;
        jobs[i].phase = 0;
        jobs[i].stop_bit = 0;
        jobs[i].symbolic = 0;
        jobs[i].pad_ = 0;
    }
    CU(cudaMemcpyAsync(ctx->d_jobs.p, jobs, sizeof(StreamJob) * count, cudaMemcpyHostToDevice, ctx->stream));
    int rc = run_inflate(ctx, jobs, count);
    ctx->bulk_h2d = nullptr;  // captured this frame's locals
    if (rc != PNGB200_OK) return rc;
    std::vector<UnfilterItem> items(count);
    for (size_t i = 0; i < count; ++i) {
        UnfilterItem& it = items[i];
        it.filtered = jobs[i].dst;
        it.filtered_mut = jobs[i].dst;
        it.pixels = host ? ctx->d_out.as<uint8_t>() + ctx->out_offset[i] : im[i].pixels;
        it.inflated = ctx->d_results.as<StreamResult>() + i;
        it.filtered_len = 0;
        it.w = im[i].width;
        it.h = im[i].height;
        it.volume = im[i].volume;
        it.depth = im[i].depth;
        it.interlaced = im[i].interlaced;
        it.g = geo[i];
    }
    rc = run_unfilter(ctx, items);
    if (rc != PNGB200_OK) return rc;
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    CU(cudaMemcpyAsync(ctx->h_results.p, ctx->d_results.p, sizeof(StreamResult) * count,
                       cudaMemcpyDeviceToHost, ctx->stream));
    if (host)
        for (size_t i = 0; i < count; ++i)
            CU(cudaMemcpyAsync(im[i].pixels, ctx->d_out.as<uint8_t>() + ctx->out_offset[i], ctx->out_bytes[i],
                               cudaMemcpyDeviceToHost, ctx->stream));
    ctx->pending = true;
    ctx->pending_memspace = memspace;
    return PNGB200_OK;
}

int pngb200_decode_batch_finish(pngb200_ctx* ctx, pngb200_image_desc* im, size_t count)
{
    if (!ctx || (!im && count)) return PNGB200_ERR_BAD_ARGUMENT;
    if (count == 0) return PNGB200_OK;
    if (!ctx->pending || ctx->expected.size() != count)
        return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "no matching pending decode batch");
    DeviceGuard guard(ctx->device);
    ctx->pending = false;
    CU(cudaStreamSynchronize(ctx->stream));
    const StreamResult* r = ctx->h_results.as<StreamResult>();
    for (size_t i = 0; i < count; ++i) {
        int st = r[i].status;
        // PNG.Decoder.push / PNG.Context.push(ancillary: IEND) error mapping
        if (st == PNGB200_ERR_OUTPUT_CAPACITY) st = PNGB200_ERR_PNG_EXTRANEOUS_IMAGE_DATA;
        else if (st == PNGB200_NEED_MORE_INPUT) st = PNGB200_ERR_PNG_INCOMPLETE_DATASTREAM;
        else if (st == PNGB200_OK && r[i].produced > ctx->expected[i]) st = PNGB200_ERR_PNG_EXTRANEOUS_IMAGE_DATA;
        im[i].status = st;
        im[i].err_a = r[i].err_a;
        im[i].err_b = r[i].err_b;
        im[i].checksum = r[i].checksum;
        im[i].blocks = r[i].blocks;
        im[i].produced = r[i].produced;
    }
    return PNGB200_OK;
}

int pngb200_decode_batch(pngb200_ctx* ctx, pngb200_image_desc* im, size_t count, int memspace)
{
    if (!ctx || (!im && count)) return PNGB200_ERR_BAD_ARGUMENT;
    return run_over_lanes(ctx, count, memspace,
                          [&](size_t i) { return im[i].idat_len + pngb200_storage_size(im[i].width, im[i].height, im[i].volume); },
                          [&](pngb200_ctx* lane, size_t lo, size_t n) {
                              int rc = pngb200_decode_batch_enqueue(lane, im + lo, n, memspace);
                              return rc != PNGB200_OK ? rc : pngb200_decode_batch_finish(lane, im + lo, n);
                          });
}

int pngb200_unfilter_batch(pngb200_ctx* ctx, pngb200_image_desc* im, size_t count, int memspace)
{
    if (!ctx || (!im && count)) return PNGB200_ERR_BAD_ARGUMENT;
    if (ctx->pending) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "a decode batch is pending");
    if (count == 0) return PNGB200_OK;
    DeviceGuard guard(ctx->device);
    const bool host = memspace == PNGB200_MEM_HOST;
    std::vector<UnfilterItem> items(count);
    std::vector<size_t>       f_off(count), o_off(count);
    size_t f_total = 0, o_total = 0;
    for (size_t i = 0; i < count; ++i) {
        UnfilterItem& it = items[i];
        if (!geometry(im[i].width, im[i].height, im[i].volume, im[i].depth, im[i].interlaced, &it.g) ||
            !im[i].idat || !im[i].pixels)
            return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "image %zu: bad descriptor", i);
        if (im[i].pixels_cap < it.g.storage) return set_error(ctx, PNGB200_ERR_OUTPUT_CAPACITY, "image %zu: pixels_cap", i);
        f_off[i] = f_total;
        // the generic kernel reconstructs in place, so it always works on a private copy
        if (host || !it.g.fast) f_total += align_up(im[i].idat_len + 64, 256);
        o_off[i] = o_total;
        o_total += align_up(it.g.storage + 16, 256);
    }
    CU(ctx->d_filtered.reserve(std::max<size_t>(f_total, 256)));
    if (host) CU(ctx->d_out.reserve(o_total));
    for (size_t i = 0; i < count; ++i) {
        UnfilterItem& it = items[i];
        uint8_t* priv = ctx->d_filtered.as<uint8_t>() + f_off[i];
        if (host || !it.g.fast) {
            CU(cudaMemcpyAsync(priv, im[i].idat, im[i].idat_len,
                               host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
            it.filtered = priv;
            it.filtered_mut = priv;
        } else {
            it.filtered = im[i].idat;
            it.filtered_mut = nullptr;
        }
        it.pixels = host ? ctx->d_out.as<uint8_t>() + o_off[i] : im[i].pixels;
        it.inflated = nullptr;
        it.filtered_len = im[i].idat_len;
        it.w = im[i].width;
        it.h = im[i].height;
        it.volume = im[i].volume;
        it.depth = im[i].depth;
        it.interlaced = im[i].interlaced;
    }
    int rc = run_unfilter(ctx, items);
    if (rc != PNGB200_OK) return rc;
    if (host)
        for (size_t i = 0; i < count; ++i)
            CU(cudaMemcpyAsync(im[i].pixels, ctx->d_out.as<uint8_t>() + o_off[i], items[i].g.storage,
                               cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < count; ++i) {
        uint64_t expect = items[i].g.filtered;
        im[i].produced = im[i].idat_len;
        im[i].status = im[i].idat_len > expect ? PNGB200_ERR_PNG_EXTRANEOUS_IMAGE_DATA : PNGB200_OK;
    }
    return PNGB200_OK;
}

// ---------------- encode stage 1: filter select + apply ----------------
int pngb200_filter_batch(pngb200_ctx* ctx, pngb200_filter_desc* im, size_t count, int memspace)
{
    if (!ctx || (!im && count)) return PNGB200_ERR_BAD_ARGUMENT;
    if (ctx->pending) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "a decode batch is pending");
    if (count == 0) return PNGB200_OK;
    DeviceGuard guard(ctx->device);
    const bool host = memspace == PNGB200_MEM_HOST;
    std::vector<FilterJob> jobs(count);
    std::vector<size_t>    p_off(count), f_off(count);
    size_t p_total = 0, f_total = 0;
    uint64_t rows = 0;
    std::vector<uint32_t> row_base(count + 1);
    for (size_t i = 0; i < count; ++i) {
        Geometry g;
        if (!geometry(im[i].width, im[i].height, im[i].volume, im[i].depth, im[i].interlaced, &g) ||
            !im[i].pixels || !im[i].filtered)
            return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "image %zu: bad descriptor", i);
        if (im[i].pixels_len < g.storage || im[i].filtered_cap < g.filtered)
            return set_error(ctx, PNGB200_ERR_OUTPUT_CAPACITY, "image %zu: buffer too small", i);
        p_off[i] = p_total;
        f_off[i] = f_total;
        p_total += align_up(g.storage + 16, 256);
        f_total += align_up(g.filtered + 16, 256);
        im[i].produced = g.filtered;
        row_base[i] = (uint32_t)rows;
        rows += filter_rows(im[i].width, im[i].height, im[i].interlaced);
    }
    row_base[count] = (uint32_t)rows;
    if (rows >= (1ull << 31)) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "batch too large");
    if (host) {
        CU(ctx->d_in.reserve(p_total));
        CU(ctx->d_out.reserve(f_total));
        for (size_t i = 0; i < count; ++i)
            CU(cudaMemcpyAsync(ctx->d_in.as<uint8_t>() + p_off[i], im[i].pixels,
                               pngb200_storage_size(im[i].width, im[i].height, im[i].volume),
                               cudaMemcpyHostToDevice, ctx->stream));
    }
    for (size_t i = 0; i < count; ++i) {
        jobs[i].pixels = host ? ctx->d_in.as<uint8_t>() + p_off[i] : im[i].pixels;
        jobs[i].filtered = host ? ctx->d_out.as<uint8_t>() + f_off[i] : im[i].filtered;
        jobs[i].width = im[i].width;
        jobs[i].height = im[i].height;
        jobs[i].volume = im[i].volume;
        jobs[i].depth = im[i].depth;
        jobs[i].interlaced = im[i].interlaced;
        jobs[i].bpp = (uint8_t)((im[i].volume + 7) >> 3);
    }
    size_t jb = sizeof(FilterJob) * count, rb = sizeof(uint32_t) * (count + 1);
    size_t off_rb = align_up(jb, 256);
    CU(ctx->h_genjobs.reserve(off_rb + rb));
    CU(ctx->d_genjobs.reserve(off_rb + rb));
    memcpy(ctx->h_genjobs.p, jobs.data(), jb);
    memcpy((char*)ctx->h_genjobs.p + off_rb, row_base.data(), rb);
    CU(cudaMemcpyAsync(ctx->d_genjobs.p, ctx->h_genjobs.p, off_rb + rb, cudaMemcpyHostToDevice, ctx->stream));
    filter_rows_kernel<<<(unsigned)std::max<uint64_t>(1, (rows + FILTER_WARPS - 1) / FILTER_WARPS),
                         FILTER_WARPS * 32, 0, ctx->stream>>>(
        ctx->d_genjobs.as<FilterJob>(), (const uint32_t*)((char*)ctx->d_genjobs.p + off_rb),
        (uint32_t)count, (uint32_t)rows);
    ctx->launches++;
    CU(cudaGetLastError());
    if (host)
        for (size_t i = 0; i < count; ++i)
            CU(cudaMemcpyAsync(im[i].filtered, ctx->d_out.as<uint8_t>() + f_off[i], im[i].produced,
                               cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < count; ++i) im[i].status = PNGB200_OK;
    return PNGB200_OK;
}

}  // extern "C"

// ---------------- streaming LZ77.Deflator handle ----------------
struct pngb200_deflator {
    pngb200_ctx*         ctx = nullptr;
    int                  format = 0, level = 9, exponent = 15;
    size_t               chunk = 65544;
    std::vector<uint8_t> input, output;
    size_t               at = 0;       // next output byte to hand out
    bool                 finished = false;
};

extern "C" {

pngb200_deflator* pngb200_deflator_create(pngb200_ctx* ctx, int format, int level, int exponent, size_t chunk_bytes)
{
    if (!ctx || format < 0 || format > 2 || level < 0 || level > 13 || exponent < 8 || exponent > 15) {
        set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "deflator_create: bad format / level / exponent");
        return nullptr;
    }
    pngb200_deflator* z = new pngb200_deflator();
    z->ctx = ctx;
    z->format = format;
    z->level = level;
    z->exponent = exponent;
    z->chunk = chunk_bytes ? chunk_bytes : 65544;
    return z;
}

void pngb200_deflator_destroy(pngb200_deflator* z) { delete z; }

int pngb200_deflator_push(pngb200_deflator* z, const uint8_t* data, size_t n, int last)
{
    if (!z || (!data && n)) return PNGB200_ERR_BAD_ARGUMENT;
    if (z->finished) return set_error(z->ctx, PNGB200_ERR_BAD_ARGUMENT, "deflator: push after push(last: true)");
    z->input.insert(z->input.end(), data, data + n);
    if (!last) return PNGB200_OK;
    z->output.resize(pngb200_deflate_bound(z->input.size()));
    pngb200_deflate_desc d;
    memset(&d, 0, sizeof d);
    d.src = z->input.data();
    d.src_len = z->input.size();
    d.dst = z->output.data();
    d.dst_cap = z->output.size();
    d.format = z->format;
    d.level = z->level;
    d.exponent = z->exponent;
    int rc = pngb200_deflate_batch(z->ctx, &d, 1, PNGB200_MEM_HOST);
    if (rc != PNGB200_OK) return rc;
    if (d.status != PNGB200_OK) return d.status;
    z->output.resize((size_t)d.produced);
    z->finished = true;
    std::vector<uint8_t>().swap(z->input);
    return PNGB200_OK;
}

int pngb200_deflator_pop(pngb200_deflator* z, const uint8_t** block, size_t* n)
{
    if (!z || !block || !n) return PNGB200_ERR_BAD_ARGUMENT;
    // DeflatorOut queues a block the moment its buffer is full (LZ77.DeflatorOut.swift:109-135): complete blocks only
    if (!z->finished || z->output.size() - z->at < z->chunk) return 0;
    *block = z->output.data() + z->at;
    *n = z->chunk;
    z->at += z->chunk;
    return 1;
}

int pngb200_deflator_pull(pngb200_deflator* z, const uint8_t** block, size_t* n)
{
    if (!z || !block || !n) return PNGB200_ERR_BAD_ARGUMENT;
    if (int got = pngb200_deflator_pop(z, block, n)) return got;
    if (!z->finished || z->at >= z->output.size()) return 0;   // pull(): flushed.isEmpty ? nil : flushed
    *block = z->output.data() + z->at;
    *n = z->output.size() - z->at;
    z->at = z->output.size();
    return 1;
}

}  // extern "C"

// ---------------- colour targets: unpack / pack ----------------
namespace {
int run_color(pngb200_ctx* ctx, pngb200_color_desc* im, size_t count, int target, int alpha_mode, int memspace, bool unpack)
{
    if (!ctx || (!im && count)) return PNGB200_ERR_BAD_ARGUMENT;
    if (ctx->pending) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "a decode batch is pending");
    if (target < PNGB200_TARGET_RGBA8 || target > PNGB200_TARGET_VA16 || alpha_mode < PNGB200_ALPHA_ASIS ||
        alpha_mode > PNGB200_ALPHA_STRAIGHTENED_AS8)
        return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "bad colour target / alpha mode");
    const bool wide_target = target == PNGB200_TARGET_RGBA16 || target == PNGB200_TARGET_VA16;
    if (alpha_mode >= PNGB200_ALPHA_PREMULTIPLIED_AS8 && !wide_target)
        return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "premultiplied(as: UInt8) needs a 16-bit target");
    if (count == 0) return PNGB200_OK;
    DeviceGuard guard(ctx->device);
    const bool   host = memspace == PNGB200_MEM_HOST;
    const size_t tpx  = target == PNGB200_TARGET_RGBA8 ? 4 : target == PNGB200_TARGET_RGBA16 ? 8 : target == PNGB200_TARGET_VA8 ? 2 : 4;
    std::vector<ColorJob> jobs(count);
    std::vector<size_t>   s_off(count), p_off(count), s_len(count);
    std::vector<uint32_t> palettes;
    size_t s_total = 0, p_total = 0;
    uint64_t most = 0;
    for (size_t i = 0; i < count; ++i) {
        const pngb200_pixel_format& f = im[i].format;
        const int ch = f.color == 0 || f.color == 3 ? 1 : f.color == 2 ? 3 : f.color == 4 ? 2 : f.color == 6 ? 4 : 0;
        const bool depth_ok = f.color == 3 ? (f.depth == 1 || f.depth == 2 || f.depth == 4 || f.depth == 8)
                            : f.color == 0 ? (f.depth == 1 || f.depth == 2 || f.depth == 4 || f.depth == 8 || f.depth == 16)
                                           : (f.depth == 8 || f.depth == 16);
        if (!ch || !depth_ok || (f.bgr && (f.depth != 8 || (f.color != 2 && f.color != 6))) ||
            (f.color == 3 && (!f.palette || f.palette_count == 0 || f.palette_count > 256)) ||
            (im[i].count && (!im[i].storage || !im[i].pixels)))
            return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "image %zu: bad colour descriptor", i);
        s_len[i] = (size_t)im[i].count * ch * (f.depth == 16 ? 2 : 1);
        if (im[i].storage_len < s_len[i] || im[i].pixels_len < im[i].count * tpx)
            return set_error(ctx, PNGB200_ERR_OUTPUT_CAPACITY, "image %zu: buffer too small", i);
        if (!host && (((uintptr_t)im[i].pixels) & (tpx - 1)))
            return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "image %zu: pixel array is not aligned to its element size", i);
        s_off[i] = s_total, p_off[i] = p_total;
        s_total += align_up(s_len[i] + 16, 256);
        p_total += align_up(im[i].count * tpx + 16, 256);
        ColorJob& j = jobs[i];
        j.count = im[i].count;
        j.color = f.color, j.depth = f.depth, j.bgr = f.bgr, j.has_key = f.has_key;
        j.key[0] = f.key[0], j.key[1] = f.key[1], j.key[2] = f.key[2];
        j.palette_off = (uint32_t)palettes.size();
        j.palette_count = f.color == 3 ? f.palette_count : 0;
        for (uint32_t k = 0; k < j.palette_count; ++k)
            palettes.push_back(f.palette[4 * k] | f.palette[4 * k + 1] << 8 | f.palette[4 * k + 2] << 16 | (uint32_t)f.palette[4 * k + 3] << 24);
        j.status = PNGB200_OK;
        most = std::max<uint64_t>(most, im[i].count);
    }
    if (host) {
        CU(ctx->d_in.reserve(unpack ? s_total : p_total));
        CU(ctx->d_out.reserve(unpack ? p_total : s_total));
        for (size_t i = 0; i < count; ++i) {
            const size_t n = unpack ? s_len[i] : im[i].count * tpx;
            if (n)
                CU(cudaMemcpyAsync(ctx->d_in.as<uint8_t>() + (unpack ? s_off[i] : p_off[i]), unpack ? im[i].storage : im[i].pixels,
                                   n, cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    for (size_t i = 0; i < count; ++i) {
        uint8_t* dev_s = host ? (unpack ? ctx->d_in : ctx->d_out).as<uint8_t>() + s_off[i] : (uint8_t*)im[i].storage;
        uint8_t* dev_p = host ? (unpack ? ctx->d_out : ctx->d_in).as<uint8_t>() + p_off[i] : (uint8_t*)im[i].pixels;
        jobs[i].storage = dev_s, jobs[i].pixels = dev_p;
    }
    const size_t jb = sizeof(ColorJob) * count, off_pal = align_up(jb, 256), pb = sizeof(uint32_t) * std::max<size_t>(palettes.size(), 1);
    CU(ctx->h_genjobs.reserve(off_pal + pb));
    CU(ctx->d_genjobs.reserve(off_pal + pb));
    memcpy(ctx->h_genjobs.p, jobs.data(), jb);
    if (!palettes.empty()) memcpy((char*)ctx->h_genjobs.p + off_pal, palettes.data(), sizeof(uint32_t) * palettes.size());
    CU(cudaMemcpyAsync(ctx->d_genjobs.p, ctx->h_genjobs.p, off_pal + pb, cudaMemcpyHostToDevice, ctx->stream));
    ColorParams p;
    p.jobs = ctx->d_genjobs.as<ColorJob>();
    p.palettes = (const uint32_t*)((char*)ctx->d_genjobs.p + off_pal);
    p.count = (uint32_t)count;
    p.target = target;
    p.alpha_mode = alpha_mode;
    // x: tiles of the largest image, capped so that x * y stays near 8 CTAs per SM; y: images
    const unsigned gy = (unsigned)std::min<size_t>(count, 65535);
    const uint64_t tiles = std::max<uint64_t>(1, (most + COLOR_TILE - 1) / COLOR_TILE);
    const unsigned gx = (unsigned)std::min<uint64_t>(tiles, std::max<uint64_t>(1, (uint64_t)ctx->sm_count * 8 / gy));
    if (unpack) unpack_kernel<<<dim3(gx, gy), COLOR_THREADS, 0, ctx->stream>>>(p);
    else pack_kernel<<<dim3(gx, gy), COLOR_THREADS, 0, ctx->stream>>>(p);
    ctx->launches++;
    CU(cudaGetLastError());
    if (host)
        for (size_t i = 0; i < count; ++i) {
            const size_t n = unpack ? im[i].count * tpx : s_len[i];
            if (n)
                CU(cudaMemcpyAsync(unpack ? im[i].pixels : im[i].storage, ctx->d_out.as<uint8_t>() + (unpack ? p_off[i] : s_off[i]), n,
                                   cudaMemcpyDeviceToHost, ctx->stream));
        }
    CU(cudaMemcpyAsync(ctx->h_genjobs.p, ctx->d_genjobs.p, jb, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const ColorJob* done = ctx->h_genjobs.as<ColorJob>();
    for (size_t i = 0; i < count; ++i) im[i].status = done[i].status;
    return PNGB200_OK;
}
}  // namespace

extern "C" {
int pngb200_unpack_batch(pngb200_ctx* ctx, pngb200_color_desc* im, size_t count, int target, int alpha_mode, int memspace)
{
    return run_color(ctx, im, count, target, alpha_mode, memspace, true);
}
int pngb200_pack_batch(pngb200_ctx* ctx, pngb200_color_desc* im, size_t count, int target, int memspace)
{
    return run_color(ctx, im, count, target, PNGB200_ALPHA_ASIS, memspace, false);
}
}  // extern "C"

// ---------------- encode stage 2: deflate ----------------
namespace {
// device-resident jobs -> compressed streams (results stay in d_dfres / are copied to `hres`)
int run_deflate(pngb200_ctx* ctx, const std::vector<DeflateJob>& jobs, DeflateResult* hres)
{
    size_t count = jobs.size();
    uint64_t verts = 2;
    for (const DeflateJob& j : jobs)
        if (j.level >= 8) verts = std::max<uint64_t>(verts, std::min<uint64_t>(j.n, DF_GRAPH_CAP) + 2);
    uint64_t stride = df_scratch_stride(verts);
    uint64_t budget = 48ull << 30;
    size_t slots = std::min<size_t>({count, (size_t)std::max<uint64_t>(1, budget / stride), (size_t)ctx->sm_count * 8});
    CU(ctx->d_dfscratch.reserve(stride * slots + 256));
    CU(ctx->d_dfjobs.reserve(sizeof(DeflateJob) * count));
    CU(ctx->d_dfres.reserve(sizeof(DeflateResult) * count));
    CU(cudaMemcpyAsync(ctx->d_dfjobs.p, jobs.data(), sizeof(DeflateJob) * count, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_dfres.p, 0, sizeof(DeflateResult) * count, ctx->stream));
    DfParams P;
    P.jobs = ctx->d_dfjobs.as<DeflateJob>();
    P.results = ctx->d_dfres.as<DeflateResult>();
    P.scratch = ctx->d_dfscratch.as<uint8_t>();
    P.scratch_stride = stride;
    P.graph_vertices = verts;
    P.ticket = (uint32_t*)(ctx->d_dfscratch.as<uint8_t>() + stride * slots);
    P.count = (int)count;
    CU(cudaMemsetAsync(P.ticket, 0, sizeof(uint32_t), ctx->stream));
    deflate_kernel<<<(unsigned)slots, 32, sizeof(DfShared), ctx->stream>>>(P);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(hres, ctx->d_dfres.p, sizeof(DeflateResult) * count, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PNGB200_OK;
}
}  // namespace

extern "C" size_t pngb200_deflate_bound(size_t n) { return n + n / 2 + 4096; }

extern "C" int pngb200_deflate_batch(pngb200_ctx* ctx, pngb200_deflate_desc* s, size_t count, int memspace)
{
    if (!ctx || (!s && count)) return PNGB200_ERR_BAD_ARGUMENT;
    if (ctx->pending) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "a decode batch is pending");
    if (count == 0) return PNGB200_OK;
    DeviceGuard guard(ctx->device);
    const bool host = memspace == PNGB200_MEM_HOST;
    std::vector<DeflateJob> jobs(count);
    std::vector<size_t> in_off(count), out_off(count);
    size_t in_total = 0, out_total = 0;
    for (size_t i = 0; i < count; ++i) {
        if ((!s[i].src && s[i].src_len) || !s[i].dst || s[i].format < 0 || s[i].format > 2 || s[i].exponent < 8 ||
            s[i].exponent > 15)
            return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "stream %zu: bad descriptor", i);
        in_off[i] = in_total;
        out_off[i] = out_total;
        in_total += align_up(s[i].src_len + 16, 256);
        out_total += align_up(s[i].dst_cap + 16, 256);
    }
    if (host) {
        CU(ctx->d_in.reserve(in_total));
        CU(ctx->d_out.reserve(out_total));
        for (size_t i = 0; i < count; ++i)
            if (s[i].src_len)
                CU(cudaMemcpyAsync(ctx->d_in.as<uint8_t>() + in_off[i], s[i].src, s[i].src_len,
                                   cudaMemcpyHostToDevice, ctx->stream));
    }
    for (size_t i = 0; i < count; ++i) {
        jobs[i].src = host ? ctx->d_in.as<uint8_t>() + in_off[i] : s[i].src;
        jobs[i].n = s[i].src_len;
        jobs[i].dst = host ? ctx->d_out.as<uint8_t>() + out_off[i] : s[i].dst;
        jobs[i].cap = s[i].dst_cap;
        jobs[i].format = s[i].format;
        jobs[i].level = s[i].level;
        jobs[i].exponent = s[i].exponent;
        jobs[i].pad = 0;
    }
    std::vector<DeflateResult> res(count);
    int rc = run_deflate(ctx, jobs, res.data());
    if (rc != PNGB200_OK) return rc;
    for (size_t i = 0; i < count; ++i) {
        s[i].status = res[i].status;
        s[i].checksum = res[i].checksum;
        s[i].blocks = res[i].blocks;
        s[i].produced = res[i].produced;
        if (host && res[i].status == PNGB200_OK && res[i].produced)
            CU(cudaMemcpyAsync(s[i].dst, ctx->d_out.as<uint8_t>() + out_off[i], res[i].produced,
                               cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return PNGB200_OK;
}

extern "C" int pngb200_encode_batch(pngb200_ctx* ctx, pngb200_encode_desc* im, size_t count, int memspace)
{
    if (!ctx || (!im && count)) return PNGB200_ERR_BAD_ARGUMENT;
    if (ctx->pending) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "a decode batch is pending");
    if (count == 0) return PNGB200_OK;
    for (size_t i = 0; i < count; ++i)
        if (!im[i].pixels || !im[i].idat)
            return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "image %zu: bad descriptor", i);
    DeviceGuard guard(ctx->device);
    const bool host = memspace == PNGB200_MEM_HOST;
    // stage 1: filter into a private device workspace (device memspace of the filter entry point)
    std::vector<pngb200_filter_desc> fd(count);
    std::vector<size_t> f_off(count), p_off(count), o_off(count);
    size_t f_total = 0, p_total = 0, o_total = 0;
    for (size_t i = 0; i < count; ++i) {
        size_t fsz = pngb200_filtered_size(im[i].width, im[i].height, im[i].volume, im[i].interlaced);
        f_off[i] = f_total;
        f_total += align_up(fsz + 16, 256);
        p_off[i] = p_total;
        p_total += align_up(im[i].pixels_len + 16, 256);
        o_off[i] = o_total;
        o_total += align_up(im[i].idat_cap + 16, 256);
    }
    CU(ctx->d_enc.reserve(f_total + (host ? p_total + o_total : 0) + 256));
    uint8_t* d_f = ctx->d_enc.as<uint8_t>();
    uint8_t* d_p = d_f + f_total;
    uint8_t* d_o = d_p + (host ? p_total : 0);
    for (size_t i = 0; i < count; ++i) {
        if (host) CU(cudaMemcpyAsync(d_p + p_off[i], im[i].pixels, im[i].pixels_len, cudaMemcpyHostToDevice, ctx->stream));
        fd[i].pixels = host ? d_p + p_off[i] : im[i].pixels;
        fd[i].pixels_len = im[i].pixels_len;
        fd[i].filtered = d_f + f_off[i];
        fd[i].filtered_cap = pngb200_filtered_size(im[i].width, im[i].height, im[i].volume, im[i].interlaced);
        fd[i].width = im[i].width;
        fd[i].height = im[i].height;
        fd[i].volume = im[i].volume;
        fd[i].depth = im[i].depth;
        fd[i].interlaced = im[i].interlaced;
    }
    int rc = pngb200_filter_batch(ctx, fd.data(), count, PNGB200_MEM_DEVICE);
    if (rc != PNGB200_OK) return rc;
    std::vector<DeflateJob> jobs(count);
    for (size_t i = 0; i < count; ++i) {
        jobs[i].src = fd[i].filtered;
        jobs[i].n = fd[i].filtered_cap;
        jobs[i].dst = host ? d_o + o_off[i] : im[i].idat;
        jobs[i].cap = im[i].idat_cap;
        jobs[i].format = im[i].format;
        jobs[i].level = im[i].level;
        jobs[i].exponent = 15;
        jobs[i].pad = 0;
    }
    std::vector<DeflateResult> res(count);
    rc = run_deflate(ctx, jobs, res.data());
    if (rc != PNGB200_OK) return rc;
    for (size_t i = 0; i < count; ++i) {
        im[i].status = res[i].status;
        im[i].checksum = res[i].checksum;
        im[i].blocks = res[i].blocks;
        im[i].produced = res[i].produced;
        if (host && res[i].status == PNGB200_OK)
            CU(cudaMemcpyAsync(im[i].idat, d_o + o_off[i], res[i].produced, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return PNGB200_OK;
}

extern "C" {
// ---------------- streaming inflator handle ----------------
struct pngb200_inflator {
    pngb200_ctx*         ctx;
    int                  format;
    std::vector<uint8_t> input;      // everything pushed so far
    DevBuf               d_in, d_out, d_job, d_res, d_misc, d_partial;
    PinBuf               h_res;
    size_t               uploaded = 0;
    uint64_t             resume_bit = 0, resume_out = 0, produced = 0, current = 0;
    uint32_t             phase = 0;
    bool                 terminal = false;
    int                  status = PNGB200_NEED_MORE_INPUT;
    uint32_t             err_a = 0, err_b = 0;
};

pngb200_inflator* pngb200_inflator_create(pngb200_ctx* ctx, int format)
{
    if (!ctx || format < 0 || format > 2) return nullptr;
    pngb200_inflator* z = new pngb200_inflator();
    z->ctx = ctx;
    z->format = format;
    return z;
}

void pngb200_inflator_destroy(pngb200_inflator* z)
{
    if (!z) return;
    DeviceGuard guard(z->ctx->device);
    cudaStreamSynchronize(z->ctx->stream);
    for (DevBuf* b : {&z->d_in, &z->d_out, &z->d_job, &z->d_res, &z->d_misc, &z->d_partial}) b->release();
    z->h_res.release();
    delete z;
}

static int inflator_grow_out(pngb200_inflator* z, size_t need)
{
    pngb200_ctx* ctx = z->ctx;
    if (need <= z->d_out.cap) return PNGB200_OK;
    DevBuf bigger;
    CU(bigger.reserve(std::max(need, z->d_out.cap * 2)));
    cudaError_t e = cudaSuccess;
    if (z->produced) e = cudaMemcpyAsync(bigger.p, z->d_out.p, z->produced, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        bigger.release();   // not adopted: give it back
        return set_error(ctx, PNGB200_ERR_CUDA, "inflator: growing the output failed: %s", cudaGetErrorString(e));
    }
    z->d_out.release();
    z->d_out = bigger;
    return PNGB200_OK;
}

int pngb200_inflator_push(pngb200_inflator* z, const uint8_t* data, size_t n)
{
    if (!z || (!data && n)) return PNGB200_ERR_BAD_ARGUMENT;
    pngb200_ctx* ctx = z->ctx;
    if (z->terminal) return PNGB200_OK;  // LZ77.Inflator ignores input after the terminal state
    if (z->status < 0) return z->status;
    DeviceGuard guard(ctx->device);
    z->input.insert(z->input.end(), data, data + n);
    // device copy of the whole input (grow-only; new bytes appended)
    if (z->input.size() + 16 > z->d_in.cap) {
        DevBuf bigger;
        CU(bigger.reserve(z->input.size() * 2 + 4096));
        z->d_in.release();
        z->d_in = bigger;
        z->uploaded = 0;
    }
    if (z->input.size() > z->uploaded)
        CU(cudaMemcpyAsync(z->d_in.as<uint8_t>() + z->uploaded, z->input.data() + z->uploaded,
                           z->input.size() - z->uploaded, cudaMemcpyHostToDevice, ctx->stream));
    z->uploaded = z->input.size();
    CU(z->d_job.reserve(sizeof(StreamJob)));
    CU(z->d_res.reserve(sizeof(StreamResult)));
    CU(z->h_res.reserve(sizeof(StreamResult) + sizeof(StreamJob)));
    int rc = inflator_grow_out(z, std::max<size_t>(1 << 16, z->produced + 4 * n + 1024));
    if (rc != PNGB200_OK) return rc;
    for (;;) {
        StreamJob* job = (StreamJob*)((char*)z->h_res.p + sizeof(StreamResult));
        job->src = z->d_in.as<uint8_t>();
        job->src_len = z->input.size();
        job->dst = z->d_out.as<uint8_t>();
        job->dst_cap = z->d_out.cap;
        job->start_bit = z->resume_bit;
        job->start_out = z->resume_out;
        job->format = z->format;
        job->phase = (int32_t)z->phase;
        job->stop_bit = 0;
        job->symbolic = 0;
        job->pad_ = 0;
        CU(cudaMemcpyAsync(z->d_job.p, job, sizeof(StreamJob), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(z->d_res.p, 0, sizeof(StreamResult), ctx->stream));
        // A push with a lot of undecoded input goes through the intra-stream parallel kernel (one CTA: ~25 x the
        // lock-step warp); short ones, and everything irregular inside it, through the serial decoder.  Both resume
        // at the last completed block boundary.
        const uint64_t pending = z->input.size() - std::min<uint64_t>(z->input.size(), z->resume_bit >> 3);
        if (pending >= (64u << 10)) {
            const uint64_t bitmap_words = wv_bitmap_words(job->dst_cap);
            const uint64_t stride = wv_scratch_stride(bitmap_words);
            const size_t   need = (size_t)stride + 256;
            if (need > ctx->d_scratch.cap || stride != ctx->scratch_stride) {   // (same convention as run_inflate: bitmaps start zeroed)
                CU(ctx->d_scratch.reserve(need));
                CU(cudaMemsetAsync(ctx->d_scratch.p, 0, ctx->d_scratch.cap, ctx->stream));
                ctx->scratch_stride = stride;
            }
            uint32_t* ticket = (uint32_t*)((char*)ctx->d_scratch.p + (size_t)stride);
            CU(cudaMemsetAsync(ticket, 0, sizeof(uint32_t), ctx->stream));
            WvParams pp;
            pp.bitmap_words = bitmap_words;
            pp.scratch_stride = stride;
            pp.ticket = ticket;
            pp.jobs = z->d_job.as<StreamJob>();
            pp.results = z->d_res.as<StreamResult>();
            pp.order = nullptr;
            pp.scratch = ctx->d_scratch.as<uint8_t>();
            pp.count = 1;
            inflate_wave_kernel<<<1, WV_THREADS, sizeof(WvShared), ctx->stream>>>(pp);
        } else {
            inflate_serial_kernel<<<1, 32, 0, ctx->stream>>>(z->d_job.as<StreamJob>(), z->d_res.as<StreamResult>(), nullptr, 1);
        }
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(z->h_res.p, z->d_res.p, sizeof(StreamResult), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        // The stream checksum is only due when the trailer has been read: ONE pass over the output at the end of the
        // stream, not one per push (round 1 re-checksummed everything produced so far on every push).
        if (z->h_res.as<StreamResult>()->trailer_seen && !z->h_res.as<StreamResult>()->ck_done) {
            uint32_t base[2] = {0, (uint32_t)((z->d_out.cap + CK_CHUNK - 1) / CK_CHUNK)};
            CU(z->d_misc.reserve(sizeof base));
            CU(z->d_partial.reserve(sizeof(uint64_t) * 2 * std::max<uint32_t>(base[1], 1)));
            CU(cudaMemcpyAsync(z->d_misc.p, base, sizeof base, cudaMemcpyHostToDevice, ctx->stream));
            ChecksumParams cp;
            cp.jobs = z->d_job.as<StreamJob>();
            cp.results = z->d_res.as<StreamResult>();
            cp.chunk_base = z->d_misc.as<uint32_t>();
            cp.partial = z->d_partial.as<uint64_t>();
            cp.count = 1;
            cp.total_chunks = base[1];
            cp.crc_tables = nullptr;
            if (z->format == PNGB200_FORMAT_GZIP) {
                if (int rc = ensure_crc_tables(ctx)) return rc;
                cp.crc_tables = ctx->d_crctab.as<uint32_t>();
            }
            checksum_chunk_kernel<<<base[1], CK_THREADS, 0, ctx->stream>>>(cp);
            checksum_fold_kernel<<<1, 32, 0, ctx->stream>>>(cp);
            ctx->launches += 2;
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(z->h_res.p, z->d_res.p, sizeof(StreamResult), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        }
        const StreamResult r = *z->h_res.as<StreamResult>();
        // remember the last completed block boundary: the next run resumes there
        z->phase = r.phase;
        z->resume_bit = r.resume_bit;
        z->resume_out = r.resume_out;
        if (r.status == PNGB200_ERR_OUTPUT_CAPACITY) {
            z->produced = r.resume_out;
            rc = inflator_grow_out(z, z->d_out.cap * 2);
            if (rc != PNGB200_OK) return rc;
            continue;
        }
        z->produced = r.produced;
        z->status = r.status;
        z->err_a = r.err_a;
        z->err_b = r.err_b;
        if (r.status == PNGB200_OK) z->terminal = true;
        return r.status;
    }
}

size_t pngb200_inflator_available(const pngb200_inflator* z) { return z ? (size_t)(z->produced - z->current) : 0; }

int pngb200_inflator_pull(pngb200_inflator* z, uint8_t* dst, size_t count)
{
    if (!z || (!dst && count)) return PNGB200_ERR_BAD_ARGUMENT;
    if (z->produced - z->current < count) return PNGB200_NEED_MORE_INPUT;
    pngb200_ctx* ctx = z->ctx;
    DeviceGuard guard(ctx->device);
    if (count) {
        CU(cudaMemcpyAsync(dst, z->d_out.as<uint8_t>() + z->current, count, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    z->current += count;
    return PNGB200_OK;
}

size_t pngb200_inflator_pull_all(pngb200_inflator* z, uint8_t* dst, size_t cap)
{
    if (!z) return 0;
    size_t n = std::min<size_t>(cap, (size_t)(z->produced - z->current));
    if (pngb200_inflator_pull(z, dst, n) != PNGB200_OK) return 0;
    return n;
}

void pngb200_inflator_error(const pngb200_inflator* z, int* status, uint32_t* a, uint32_t* b)
{
    if (!z) return;
    if (status) *status = z->status;
    if (a) *a = z->err_a;
    if (b) *b = z->err_b;
}

}  // extern "C"

#include "png_file.cuh"
