// Host build of inflate_parallel_kernel under the SIMT emulator (tests/emu/simt.h): test infrastructure.
#define PNGB200_EMU 1
#include "../../swift-png_b200/csrc/inflate_parallel.cuh"

using namespace pngb200;

extern "C" int emu_inflate_parallel(const uint8_t* src, uint64_t len, uint8_t* dst, uint64_t cap, int format,
                                    StreamResult* res, int order)
{
    StreamJob job{};
    job.src = src; job.src_len = len; job.dst = dst; job.dst_cap = cap; job.format = format;
    memset(res, 0, sizeof *res);
    uint32_t ticket = 0;
    ParParams P{};
    P.jobs = &job; P.results = res; P.order = nullptr; P.ticket = &ticket; P.count = 1;
    P.bitmap_words = par_bitmap_words(cap);
    P.scratch_stride = par_scratch_stride(P.bitmap_words);
    std::vector<uint8_t> scratch(P.scratch_stride + 256, 0);
    P.scratch = scratch.data();
    simt::launch(1, PAR_THREADS, sizeof(ParShared), [&]() { inflate_parallel_kernel(P); }, order);
    return res->status;
}
