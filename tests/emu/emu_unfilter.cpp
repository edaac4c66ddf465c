// Host build of unfilter_wave_kernel under the SIMT emulator (tests/emu/simt.h): test infrastructure.
#define PNGB200_EMU 1
#include "../../swift-png_b200/csrc/unfilter.cuh"

using namespace pngb200;

// one image, `grid` CTAs resident at once (bands wait for each other through the progress words)
extern "C" int emu_unfilter(const uint8_t* filtered, uint64_t filtered_len, uint8_t* pixels, uint32_t w, uint32_t h,
                            uint32_t bpp, uint32_t depth, unsigned grid, int order)
{
    ImageJob job{};
    job.filtered = filtered; job.pixels = pixels; job.inflated = nullptr; job.filtered_len = filtered_len;
    job.width = w; job.height = h; job.pitch = w * bpp; job.volume = (uint8_t)(8 * bpp); job.depth = (uint8_t)depth;
    job.interlaced = 0; job.bpp = (uint8_t)bpp;
    const uint32_t bands = (h + 31) / 32;
    uint32_t band_base[2] = {0, bands};
    std::vector<uint32_t> progress(bands + 1, 0);
    uint32_t ticket = 0;
    WaveParams p{};
    p.jobs = &job; p.band_base = band_base; p.progress = progress.data(); p.ticket = &ticket;
    p.njobs = 1; p.total_bands = bands; p.hist = nullptr;
    simt::launch(grid, WAVE_WARPS * 32, WAVE_SMEM, [&]() { unfilter_wave_kernel(p); }, order, true);
    return 0;
}
extern "C" int emu_unfilter_config(int* burst, int* depth, int* warps) { *burst = WAVE_BURST; *depth = WAVE_DEPTH; *warps = WAVE_WARPS; return (int)WAVE_SMEM; }
