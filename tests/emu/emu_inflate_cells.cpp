// Host build of inflate_cells_kernel under the SIMT emulator (tests/emu/simt.h): test infrastructure.
#define PNGB200_EMU 1
#define WV_PROFILE 1
#include "../../swift-png_b200/csrc/inflate_cells.cuh"

using namespace pngb200;

extern "C" int emu_inflate_cells(const uint8_t* src, uint64_t len, uint8_t* dst, uint64_t cap, int format,
                                 StreamResult* res, int order)
{
    StreamJob job{};
    job.src = src; job.src_len = len; job.dst = dst; job.dst_cap = cap; job.format = format;
    memset(res, 0, sizeof *res);
    uint32_t ticket = 0;
    WvParams P{};
    P.jobs = &job; P.results = res; P.order = nullptr; P.ticket = &ticket; P.count = 1;
    std::vector<uint8_t> scratch(CL_SCRATCH + 256, 0);
    P.scratch = scratch.data();
    P.scratch_stride = CL_SCRATCH;
    simt::launch(1, WV_THREADS, sizeof(ClShared), [&]() { inflate_cells_kernel(P); }, order);
    return res->status;
}
extern "C" size_t emu_result_size() { return sizeof(StreamResult); }
extern "C" size_t emu_shared_size() { return sizeof(ClShared); }
extern "C" void emu_profile(uint64_t* thread_iters, uint64_t* warp_iters, int reset)
{
    for (int i = 0; i < 8; ++i) { thread_iters[i] = wv_profile().thread_iters[i]; warp_iters[i] = wv_profile().warp_iters[i]; }
    if (reset) memset(&wv_profile(), 0, sizeof(WvProfile));
}
