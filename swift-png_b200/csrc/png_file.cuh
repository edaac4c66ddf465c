// png_file.cuh -- file-level entry points: whole PNG files in, pixels out,
// and back.  Included at the end of pngb200_api.cu (uses the context internals defined there).
//
// The reference's PNG.Image.decompress(stream:) (Sources/PNG/PNG.Image.swift:298-401) interleaves
// lexing, per-chunk CRC-32, chunk parsing and decoding, one chunk at a time, on the host.  Here the
// host only walks chunk HEADERS (8 bytes per chunk); every payload byte is touched on the device:
// the file goes to HBM once, crc_regions_kernel checks all chunks of all files in one launch,
// segment_copy_kernel concatenates IDAT bodies when a file has more than one, and the batch decode
// path runs on the result.  Errors are reported in the order the reference's streaming loop would
// meet them (see resolve order below).  Encode mirrors compress(stream:level:hint:) (:576-670).
#pragma once

namespace {

constexpr uint32_t fourcc(char a, char b, char c, char d)
{
    return (uint32_t)(uint8_t)a << 24 | (uint32_t)(uint8_t)b << 16 | (uint32_t)(uint8_t)c << 8 | (uint32_t)(uint8_t)d;
}
constexpr uint32_t CK_CgBI = fourcc('C', 'g', 'B', 'I'), CK_IHDR = fourcc('I', 'H', 'D', 'R'), CK_PLTE = fourcc('P', 'L', 'T', 'E'),
                   CK_IDAT = fourcc('I', 'D', 'A', 'T'), CK_IEND = fourcc('I', 'E', 'N', 'D'), CK_tRNS = fourcc('t', 'R', 'N', 'S'),
                   CK_bKGD = fourcc('b', 'K', 'G', 'D'), CK_hIST = fourcc('h', 'I', 'S', 'T'), CK_cHRM = fourcc('c', 'H', 'R', 'M'),
                   CK_gAMA = fourcc('g', 'A', 'M', 'A'), CK_sRGB = fourcc('s', 'R', 'G', 'B'), CK_iCCP = fourcc('i', 'C', 'C', 'P'),
                   CK_sBIT = fourcc('s', 'B', 'I', 'T'), CK_pHYs = fourcc('p', 'H', 'Y', 's'), CK_sPLT = fourcc('s', 'P', 'L', 'T'),
                   CK_tIME = fourcc('t', 'I', 'M', 'E'), CK_iTXt = fourcc('i', 'T', 'X', 't'), CK_tEXt = fourcc('t', 'E', 'X', 't'),
                   CK_zTXt = fourcc('z', 'T', 'X', 't');
const uint8_t PNG_SIGNATURE[8] = {137, 80, 78, 71, 13, 10, 26, 10};

inline uint32_t load_be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
inline uint32_t load_be16(const uint8_t* p) { return (uint32_t)p[0] << 8 | p[1]; }
inline void     store_be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24), p[1] = (uint8_t)(v >> 16), p[2] = (uint8_t)(v >> 8), p[3] = (uint8_t)v; }

// PNG.Chunk.init(validating:) (Lexing/PNG.Chunk.swift:39-58)
inline bool chunk_type_ok(uint32_t name)
{
    switch (name) {
    case CK_CgBI: case CK_IHDR: case CK_PLTE: case CK_IDAT: case CK_IEND: case CK_cHRM: case CK_gAMA: case CK_iCCP:
    case CK_sBIT: case CK_sRGB: case CK_bKGD: case CK_hIST: case CK_tRNS: case CK_pHYs: case CK_sPLT: case CK_tIME:
    case CK_iTXt: case CK_tEXt: case CK_zTXt:
        return true;
    default:
        return (name & 0x20002000u) == 0x20000000u;
    }
}

struct ChunkRec {
    uint64_t off;       // offset of the chunk's length field in the file
    uint32_t len;       // body bytes
    uint32_t type;
    uint32_t declared;  // CRC-32 stored behind the body
};

// What the header walk learned about one file.  `stop` is the index of the chunk at which the
// reference would have thrown for a structural reason (chunks.size() if none); `stop_before_crc`
// tells whether that happens before the chunk's own CRC check (lexing) or after it (parsing /
// ordering).
struct FileWalk {
    std::vector<ChunkRec> chunks;
    int      status = PNGB200_OK;
    uint32_t a = 0, b = 0;
    size_t   stop = (size_t)-1;
    bool     stop_before_crc = false;
    size_t   first_idat = (size_t)-1, idat_end = 0;  // [first_idat, idat_end): the contiguous IDAT run
};

inline int channels_of_color(int color) { return color == 0 || color == 3 ? 1 : color == 2 ? 3 : color == 4 ? 2 : 4; }

// Walks the chunk headers of one file and parses IHDR / PLTE / tRNS into `d`
// (PNG.Image.decompress(stream:), PNG.Image.swift:298-401; PNG.Header.init(parsing:standard:),
// Parsing/PNG.Header.swift:40-98; PNG.Palette.init(parsing:pixel:), PNG.Palette.swift:27-55;
// PNG.Transparency.init(parsing:pixel:palette:), PNG.Transparency.swift:68-122; ordering rules of
// Decoding/PNG.Metadata.swift:70-92 and PNG.Context.swift:51-81).  No payload byte other than those
// three chunks' is read.
void walk_file(pngb200_png_desc& d, FileWalk& w)
{
    const uint8_t* f = d.file;
    const size_t   n = d.file_len;
    auto stop = [&](int status, uint32_t a, uint32_t b, bool before_crc) {
        w.status = status, w.a = a, w.b = b;
        w.stop = w.chunks.size() - (before_crc ? 0 : 1);
        w.stop_before_crc = before_crc;
    };
    d.width = d.height = 0;
    d.depth = d.color = d.interlaced = d.standard = 0;
    memset(&d.format, 0, sizeof d.format);
    d.format.palette = d.palette_rgba;
    d.storage_size = d.idat_bytes = 0;
    d.idat_chunks = d.chunks = 0;
    if (n < 8) { w.status = PNGB200_ERR_LEX_TRUNCATED_SIGNATURE, w.stop = 0, w.stop_before_crc = true; return; }
    if (memcmp(f, PNG_SIGNATURE, 8)) {
        w.status = PNGB200_ERR_LEX_INVALID_SIGNATURE, w.a = load_be32(f), w.b = load_be32(f + 4), w.stop = 0, w.stop_before_crc = true;
        return;
    }
    size_t at = 8;
    // lexes one chunk header; false = stopped
    auto lex = [&]() -> bool {
        if (n - at < 8) { stop(PNGB200_ERR_LEX_TRUNCATED_CHUNK_HEADER, 0, 0, true); return false; }
        const uint32_t len = load_be32(f + at), name = load_be32(f + at + 4);
        if (!chunk_type_ok(name)) { stop(PNGB200_ERR_LEX_INVALID_CHUNK_TYPE, name, 0, true); return false; }
        if ((uint64_t)(n - at - 8) < (uint64_t)len + 4) { stop(PNGB200_ERR_LEX_TRUNCATED_CHUNK_BODY, len + 4, 0, true); return false; }
        w.chunks.push_back({at, len, name, load_be32(f + at + 8 + len)});
        at += 12 + (size_t)len;
        return true;
    };
    if (!lex()) return;
    if (w.chunks.back().type == CK_CgBI) {
        d.standard = 1;
        if (!lex()) return;
    }
    {
        const ChunkRec& c = w.chunks.back();
        if (c.type != CK_IHDR) return stop(PNGB200_ERR_DECODE_REQUIRED_CHUNK, CK_IHDR, c.type, false);
        const uint8_t* h = f + c.off + 8;
        if (c.len != 13) return stop(PNGB200_ERR_PARSE_HEADER_CHUNK_LENGTH, c.len, 0, false);
        const int depth = h[8], color = h[9];
        bool ok;  // PNG.Format.Pixel.recognize(code:)
        switch (color) {
        case 0: ok = depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16; break;
        case 3: ok = depth == 1 || depth == 2 || depth == 4 || depth == 8; break;
        case 2: case 4: case 6: ok = depth == 8 || depth == 16; break;
        default: ok = false;
        }
        if (!ok) return stop(PNGB200_ERR_PARSE_HEADER_PIXEL_FORMAT_CODE, (uint32_t)depth, (uint32_t)color, false);
        if (d.standard == 1 && !(depth == 8 && (color == 2 || color == 6)))
            return stop(PNGB200_ERR_PARSE_HEADER_PIXEL_FORMAT, (uint32_t)depth, (uint32_t)color, false);
        if (h[10]) return stop(PNGB200_ERR_PARSE_HEADER_COMPRESSION_CODE, h[10], 0, false);
        if (h[11]) return stop(PNGB200_ERR_PARSE_HEADER_FILTER_CODE, h[11], 0, false);
        if (h[12] > 1) return stop(PNGB200_ERR_PARSE_HEADER_INTERLACING_CODE, h[12], 0, false);
        d.width = load_be32(h), d.height = load_be32(h + 4);
        if (!d.width || !d.height) return stop(PNGB200_ERR_PARSE_HEADER_SIZE, d.width, d.height, false);
        {
            // the reference traps when the storage size overflows (PNG.Image.swift:84); refuse such a file here
            const uint64_t bpp = (uint64_t)((depth * channels_of_color(color) + 7) >> 3);
            uint64_t prod;
            if (d.width > 0x7fffffffu || d.height > 0x7fffffffu ||
                __builtin_mul_overflow((uint64_t)d.width * d.height, bpp ? bpp : 1, &prod) || prod > (1ull << 46))
                return stop(PNGB200_ERR_PARSE_HEADER_SIZE, d.width, d.height, false);
        }
        d.depth = (uint8_t)depth, d.color = (uint8_t)color, d.interlaced = h[12];
        d.format.color = d.color, d.format.depth = d.depth, d.format.bgr = d.standard;
        d.storage_size = (uint64_t)d.width * d.height * (uint64_t)((depth * channels_of_color(color) + 7) >> 3);
    }
    bool     have_palette = false, have_background = false, have_transparency = false;
    uint32_t npal = 0, nalpha = 0;
    uint8_t  alpha[256];
    for (;;) {  // up to the first IDAT
        if (!lex()) return;
        const ChunkRec& c = w.chunks.back();
        const uint8_t*  body = f + c.off + 8;
        if (c.type == CK_IHDR) return stop(PNGB200_ERR_DECODE_DUPLICATE_CHUNK, CK_IHDR, 0, false);
        if (c.type == CK_PLTE) {
            if (have_palette) return stop(PNGB200_ERR_DECODE_DUPLICATE_CHUNK, CK_PLTE, 0, false);
            if (have_background) return stop(PNGB200_ERR_DECODE_UNEXPECTED_CHUNK, CK_PLTE, CK_bKGD, false);
            if (have_transparency) return stop(PNGB200_ERR_DECODE_UNEXPECTED_CHUNK, CK_PLTE, CK_tRNS, false);
            if (d.color == 0 || d.color == 4) return stop(PNGB200_ERR_PARSE_UNEXPECTED_PALETTE, 0, 0, false);
            if (c.len % 3) return stop(PNGB200_ERR_PARSE_PALETTE_CHUNK_LENGTH, c.len, 0, false);
            const uint32_t max = 1u << std::min<int>(d.depth, 8);
            if (c.len / 3 < 1 || c.len / 3 > max) return stop(PNGB200_ERR_PARSE_PALETTE_COUNT, c.len / 3, max, false);
            have_palette = true, npal = c.len / 3;
            if (d.color == 3)
                for (uint32_t i = 0; i < npal; ++i) {
                    memcpy(d.palette_rgba + 4 * i, body + 3 * i, 3);
                    d.palette_rgba[4 * i + 3] = 255;
                }
        } else if (c.type == CK_tRNS) {
            if (have_transparency) return stop(PNGB200_ERR_DECODE_DUPLICATE_CHUNK, CK_tRNS, 0, false);
            const uint32_t max = 0xffffu >> (16 - d.depth);
            if (d.color == 0) {
                if (c.len != 2) return stop(PNGB200_ERR_PARSE_TRANSPARENCY_CHUNK_LENGTH, c.len, 2, false);
                if (load_be16(body) > max) return stop(PNGB200_ERR_PARSE_TRANSPARENCY_SAMPLE, load_be16(body), max, false);
                d.format.has_key = 1, d.format.key[0] = (uint16_t)load_be16(body);
            } else if (d.color == 2) {
                if (c.len != 6) return stop(PNGB200_ERR_PARSE_TRANSPARENCY_CHUNK_LENGTH, c.len, 6, false);
                const uint32_t r = load_be16(body), g = load_be16(body + 2), b = load_be16(body + 4);
                if (std::max({r, g, b}) > max) return stop(PNGB200_ERR_PARSE_TRANSPARENCY_SAMPLE, std::max({r, g, b}), max, false);
                d.format.has_key = 1;  // Format.recognize keeps a bgr8 key in (b, g, r) order (PNG.Format.swift:228-240)
                d.format.key[0] = (uint16_t)(d.standard ? b : r), d.format.key[1] = (uint16_t)g, d.format.key[2] = (uint16_t)(d.standard ? r : b);
            } else if (d.color == 3) {
                if (!have_palette) return stop(PNGB200_ERR_DECODE_REQUIRED_CHUNK, CK_PLTE, CK_tRNS, false);
                if (c.len > npal) return stop(PNGB200_ERR_PARSE_TRANSPARENCY_COUNT, c.len, npal, false);
                memcpy(alpha, body, c.len), nalpha = c.len;
            } else
                return stop(PNGB200_ERR_PARSE_UNEXPECTED_TRANSPARENCY, 0, 0, false);
            have_transparency = true;
        } else if (c.type == CK_bKGD) {
            if (have_background) return stop(PNGB200_ERR_DECODE_DUPLICATE_CHUNK, CK_bKGD, 0, false);
            if (d.color == 3 && !have_palette) return stop(PNGB200_ERR_DECODE_REQUIRED_CHUNK, CK_PLTE, CK_bKGD, false);
            have_background = true;
        } else if (c.type == CK_cHRM || c.type == CK_gAMA || c.type == CK_sRGB || c.type == CK_iCCP || c.type == CK_sBIT) {
            if (have_palette) return stop(PNGB200_ERR_DECODE_UNEXPECTED_CHUNK, c.type, CK_PLTE, false);
        } else if (c.type == CK_hIST) {
            if (!have_palette) return stop(PNGB200_ERR_DECODE_REQUIRED_CHUNK, CK_PLTE, CK_hIST, false);
        } else if (c.type == CK_IDAT) {
            if (d.color == 3 && !have_palette) return stop(PNGB200_ERR_DECODE_REQUIRED_CHUNK, CK_PLTE, CK_IDAT, false);
            for (uint32_t i = 0; i < nalpha; ++i) d.palette_rgba[4 * i + 3] = alpha[i];
            d.format.palette_count = d.color == 3 ? (uint16_t)npal : 0;
            break;
        } else if (c.type == CK_IEND) {
            return stop(PNGB200_ERR_DECODE_REQUIRED_CHUNK, CK_IDAT, CK_IEND, false);
        }
    }
    w.first_idat = w.chunks.size() - 1;
    while (w.chunks.back().type == CK_IDAT) {
        d.idat_bytes += w.chunks.back().len, d.idat_chunks++;
        w.idat_end = w.chunks.size();
        if (!lex()) return;
    }
    for (;;) {  // Context.push(ancillary:) until IEND
        const uint32_t t = w.chunks.back().type;
        if (t == CK_IEND) return;
        switch (t) {
        case CK_CgBI: case CK_IHDR: case CK_PLTE: case CK_bKGD: case CK_tRNS: case CK_IDAT: case CK_hIST: case CK_cHRM:
        case CK_gAMA: case CK_sRGB: case CK_iCCP: case CK_sBIT: case CK_pHYs: case CK_sPLT:
            return stop(PNGB200_ERR_DECODE_UNEXPECTED_CHUNK, t, CK_IDAT, false);
        default: break;
        }
        if (!lex()) return;
    }
}

// CRC-32 of `regions` (device pointers) in three enqueue steps, so that a caller can put the small
// table upload in front of its bulk H2D copies and the small result download behind its kernels (copies
// of one direction are served in issue order across all streams: a few KB queued behind another lane's
// gigabyte would stall this lane for its whole duration).
struct CrcPlan {
    CrcParams p;
    size_t    acc_bytes = 0, off_acc = 0;
    bool      any = false;
};
int crc_upload(pngb200_ctx* ctx, const std::vector<CrcRegion>& regions, uint32_t* d_acc_out, CrcPlan* plan)
{
    const size_t count = regions.size();
    plan->any = count != 0;
    if (count == 0) return PNGB200_OK;
    int rc = ensure_crc_tables(ctx);
    if (rc != PNGB200_OK) return rc;
    std::vector<uint32_t> base(count + 1);
    uint64_t pieces = 0;
    for (size_t i = 0; i < count; ++i) {
        base[i] = (uint32_t)pieces;
        pieces += std::max<uint64_t>(1, (regions[i].len + CRC_PIECE - 1) / CRC_PIECE);
    }
    base[count] = (uint32_t)pieces;
    if (pieces >= (1ull << 31)) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "batch too large");
    const size_t rb = sizeof(CrcRegion) * count, off_base = align_up(rb, 256), bb = sizeof(uint32_t) * (count + 1),
                 off_acc = align_up(off_base + bb, 256), ab = sizeof(uint32_t) * count;
    CU(ctx->h_crc.reserve(off_acc + ab));
    CU(ctx->d_crc.reserve(off_acc + ab));
    memcpy(ctx->h_crc.p, regions.data(), rb);
    memcpy((char*)ctx->h_crc.p + off_base, base.data(), bb);
    CU(cudaMemcpyAsync(ctx->d_crc.p, ctx->h_crc.p, off_base + bb, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t* acc = d_acc_out ? d_acc_out : (uint32_t*)((char*)ctx->d_crc.p + off_acc);
    CU(cudaMemsetAsync(acc, 0, ab, ctx->stream));
    plan->p.regions = ctx->d_crc.as<CrcRegion>();
    plan->p.piece_base = (const uint32_t*)((char*)ctx->d_crc.p + off_base);
    plan->p.acc = acc;
    plan->p.tables = ctx->d_crctab.as<uint32_t>();
    plan->p.count = (uint32_t)count;
    plan->p.total_pieces = (uint32_t)pieces;
    plan->acc_bytes = ab, plan->off_acc = off_acc;
    return PNGB200_OK;
}
int crc_launch(pngb200_ctx* ctx, const CrcPlan& plan)
{
    if (!plan.any) return PNGB200_OK;
    crc_regions_kernel<<<plan.p.total_pieces, CRC_THREADS, 0, ctx->stream>>>(plan.p);
    ctx->launches++;
    CU(cudaGetLastError());
    return PNGB200_OK;
}
// enqueue the download of the results into pinned memory; valid after the stream's next synchronisation
int crc_fetch(pngb200_ctx* ctx, const CrcPlan& plan, const uint32_t** pinned_out)
{
    *pinned_out = nullptr;
    if (!plan.any) return PNGB200_OK;
    CU(cudaMemcpyAsync((char*)ctx->h_crc.p + plan.off_acc, plan.p.acc, plan.acc_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    *pinned_out = (const uint32_t*)((char*)ctx->h_crc.p + plan.off_acc);
    return PNGB200_OK;
}
// all three at once; host_out: wait for the results
int run_crc(pngb200_ctx* ctx, const std::vector<CrcRegion>& regions, uint32_t* d_acc_out, std::vector<uint32_t>* host_out)
{
    CrcPlan plan;
    int rc = crc_upload(ctx, regions, d_acc_out, &plan);
    if (rc == PNGB200_OK) rc = crc_launch(ctx, plan);
    if (rc != PNGB200_OK || !host_out || !plan.any) return rc;
    const uint32_t* pinned = nullptr;
    rc = crc_fetch(ctx, plan, &pinned);
    if (rc != PNGB200_OK) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    host_out->assign(pinned, pinned + regions.size());
    return PNGB200_OK;
}

struct CopyPlan {
    CopyParams p;
    bool       any = false;
};
int copy_upload(pngb200_ctx* ctx, const std::vector<CopySegment>& segs, CopyPlan* plan)
{
    const size_t count = segs.size();
    plan->any = count != 0;
    if (count == 0) return PNGB200_OK;
    std::vector<uint32_t> base(count + 1);
    uint64_t pieces = 0;
    for (size_t i = 0; i < count; ++i) {
        base[i] = (uint32_t)pieces;
        pieces += std::max<uint64_t>(1, (segs[i].len + CRC_PIECE - 1) / CRC_PIECE);
    }
    base[count] = (uint32_t)pieces;
    if (pieces >= (1ull << 31)) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "batch too large");
    const size_t sb = sizeof(CopySegment) * count, off_base = align_up(sb, 256), bb = sizeof(uint32_t) * (count + 1);
    CU(ctx->h_seg.reserve(off_base + bb));
    CU(ctx->d_seg.reserve(off_base + bb));
    memcpy(ctx->h_seg.p, segs.data(), sb);
    memcpy((char*)ctx->h_seg.p + off_base, base.data(), bb);
    CU(cudaMemcpyAsync(ctx->d_seg.p, ctx->h_seg.p, off_base + bb, cudaMemcpyHostToDevice, ctx->stream));
    plan->p.segments = ctx->d_seg.as<CopySegment>();
    plan->p.piece_base = (const uint32_t*)((char*)ctx->d_seg.p + off_base);
    plan->p.count = (uint32_t)count;
    plan->p.total_pieces = (uint32_t)pieces;
    return PNGB200_OK;
}
int copy_launch(pngb200_ctx* ctx, const CopyPlan& plan)
{
    if (!plan.any) return PNGB200_OK;
    segment_copy_kernel<<<plan.p.total_pieces, CRC_THREADS, 0, ctx->stream>>>(plan.p);
    ctx->launches++;
    CU(cudaGetLastError());
    return PNGB200_OK;
}
int run_segment_copy(pngb200_ctx* ctx, const std::vector<CopySegment>& segs)
{
    CopyPlan plan;
    int rc = copy_upload(ctx, segs, &plan);
    return rc != PNGB200_OK ? rc : copy_launch(ctx, plan);
}

// One chunk of a png_decode batch on one context.  `optimistic`: the chunk CRCs are computed on the
// device while the decode is already running on the assumption that they all match (no host round trip
// between the two); the rare file whose CRC failure changes what the decoder may see -- a bad chunk
// before the end of its IDAT run -- is put through the exact order once more (`optimistic` = false:
// CRCs first, then the decode of just the IDAT chunks lexed before the failure).
int png_decode_some(pngb200_ctx* ctx, pngb200_png_desc* d, size_t count, int memspace, bool optimistic = true)
{
    DeviceGuard guard(ctx->device);
    const bool host_pixels = memspace == PNGB200_MEM_HOST;
    std::vector<FileWalk> walks(count);
    // Device image of a file: [bytes in front of the first IDAT | bytes behind the IDAT run] in the file
    // arena, and the bodies of the IDAT run back to back in the payload arena.  The H2D copy itself
    // does the concatenation: a run of equal-sized IDAT chunks (what every encoder writes, the reference
    // included) is one pitched copy (cudaMemcpy2DAsync: row = body, source pitch = body + 12).
    std::vector<size_t> f_off(count), g_off(count), pre_len(count), post_at(count), post_len(count);
    size_t f_total = 0, g_total = 0;
    for (size_t i = 0; i < count; ++i) {
        if (!d[i].file && d[i].file_len) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "file %zu: null pointer", i);
        FileWalk& w = walks[i];
        walk_file(d[i], w);
        d[i].chunks = (uint32_t)w.chunks.size();
        d[i].status = PNGB200_OK, d[i].err_a = d[i].err_b = 0;
        d[i].checksum = d[i].blocks = 0, d[i].produced = 0;
        if (w.first_idat != (size_t)-1 && (!d[i].pixels || d[i].pixels_cap < d[i].storage_size))
            return set_error(ctx, PNGB200_ERR_OUTPUT_CAPACITY, "file %zu: pixels_cap %zu < %llu", i, d[i].pixels_cap,
                             (unsigned long long)d[i].storage_size);
        const size_t lexed = w.chunks.empty() ? 0 : (size_t)(w.chunks.back().off + 12 + w.chunks.back().len);
        if (w.first_idat == (size_t)-1) {
            pre_len[i] = lexed, post_at[i] = lexed, post_len[i] = 0;
        } else {
            pre_len[i] = (size_t)w.chunks[w.first_idat].off;
            post_at[i] = w.idat_end < w.chunks.size() ? (size_t)w.chunks[w.idat_end].off : lexed;
            post_len[i] = lexed - post_at[i];
        }
        f_off[i] = f_total;
        f_total += align_up(pre_len[i] + post_len[i] + 16, 256);
        g_off[i] = g_total;
        g_total += align_up(d[i].idat_bytes + 16, 256);
    }
    CU(ctx->d_file.reserve(std::max<size_t>(f_total, 256)));
    CU(ctx->d_in.reserve(std::max<size_t>(g_total, 256)));
    // every lexed chunk's CRC region (type + body): in the file arena, or -- IDAT run -- the body in the
    // payload arena with the type folded in as a prefix
    std::vector<CrcRegion> regions;
    std::vector<size_t>    region_base(count + 1);
    for (size_t i = 0; i < count; ++i) {
        region_base[i] = regions.size();
        const FileWalk& w = walks[i];
        uint8_t* const  meta = ctx->d_file.as<uint8_t>() + f_off[i];
        uint8_t*        body = ctx->d_in.as<uint8_t>() + g_off[i];
        for (size_t k = 0; k < w.chunks.size(); ++k) {
            const ChunkRec& c = w.chunks[k];
            if (w.first_idat != (size_t)-1 && k >= w.first_idat && k < w.idat_end) {
                regions.push_back({body, (uint64_t)c.len, CK_IDAT, 1});
                body += c.len;
            } else {
                const size_t at = c.off < pre_len[i] ? (size_t)c.off : pre_len[i] + ((size_t)c.off - post_at[i]);
                regions.push_back({meta + at + 4, (uint64_t)c.len + 4, 0, 0});
            }
        }
    }
    region_base[count] = regions.size();
    auto upload_files = [&]() -> int {
        for (size_t i = 0; i < count; ++i) {
            const FileWalk& w = walks[i];
            uint8_t* const  meta = ctx->d_file.as<uint8_t>() + f_off[i];
            if (pre_len[i]) CU(cudaMemcpyAsync(meta, d[i].file, pre_len[i], cudaMemcpyHostToDevice, ctx->stream));
            if (post_len[i])
                CU(cudaMemcpyAsync(meta + pre_len[i], d[i].file + post_at[i], post_len[i], cudaMemcpyHostToDevice, ctx->stream));
            if (w.first_idat == (size_t)-1) continue;
            uint8_t* body = ctx->d_in.as<uint8_t>() + g_off[i];
            for (size_t k = w.first_idat; k < w.idat_end;) {
                const uint32_t len = w.chunks[k].len;
                size_t run = 1;
                while (k + run < w.idat_end && w.chunks[k + run].len == len) ++run;
                const uint8_t* src = d[i].file + w.chunks[k].off + 8;
                if (len == 0) {
                } else if (run == 1) {
                    CU(cudaMemcpyAsync(body, src, len, cudaMemcpyHostToDevice, ctx->stream));
                } else {
                    CU(cudaMemcpy2DAsync(body, len, src, (size_t)len + 12, len, run, cudaMemcpyHostToDevice, ctx->stream));
                }
                body += (size_t)len * run;
                k += run;
            }
        }
        return PNGB200_OK;
    };
    std::vector<uint32_t> crc;
    const uint32_t*       crc_late = nullptr;
    int rc = PNGB200_OK;
    if (!optimistic) {
        if ((rc = upload_files()) != PNGB200_OK) return rc;
        if ((rc = run_crc(ctx, regions, nullptr, &crc)) != PNGB200_OK) return rc;
    }
    // Resolve where the reference's loop would have stopped lexing: the first chunk, in file order, with
    // a bad CRC or a structural error (a lexing error precedes that chunk's CRC check, a parsing /
    // ordering error follows it).
    struct Plan { size_t stop; int status; uint32_t a, b; size_t idat_lo, idat_hi; };
    std::vector<Plan> plans(count);
    std::vector<pngb200_image_desc> images;
    std::vector<size_t>             owner;
    std::vector<size_t>             o_off(count);
    size_t o_total = 0;
    for (size_t i = 0; i < count; ++i) {
        const FileWalk& w = walks[i];
        Plan& p = plans[i];
        p.stop = w.stop, p.status = w.status, p.a = w.a, p.b = w.b;
        for (size_t k = 0; k < w.chunks.size(); ++k) {
            if (optimistic || k > w.stop || (k == w.stop && w.stop_before_crc)) break;
            if (crc[region_base[i] + k] != w.chunks[k].declared) {
                p.stop = k, p.status = PNGB200_ERR_LEX_INVALID_CHUNK_CHECKSUM, p.a = w.chunks[k].declared, p.b = crc[region_base[i] + k];
                break;
            }
        }
        // IDAT chunks pushed into the decoder before lexing stopped
        p.idat_lo = p.idat_hi = 0;
        if (w.first_idat != (size_t)-1 && p.stop > w.first_idat) {
            p.idat_lo = w.first_idat;
            p.idat_hi = std::min(w.idat_end, p.stop);
        }
        if (p.idat_hi == p.idat_lo) continue;
        uint64_t payload = 0;
        for (size_t k = p.idat_lo; k < p.idat_hi; ++k) payload += w.chunks[k].len;
        pngb200_image_desc im;
        memset(&im, 0, sizeof im);
        im.idat = ctx->d_in.as<uint8_t>() + g_off[i];  // the chunks pushed so far are a prefix of the gathered run
        im.idat_len = payload;
        if (host_pixels) {
            o_off[i] = o_total;
            o_total += align_up(d[i].storage_size + 16, 256);
        } else {
            im.pixels = (uint8_t*)d[i].pixels;
        }
        im.pixels_cap = d[i].storage_size;
        im.width = d[i].width, im.height = d[i].height;
        im.volume = (uint8_t)(d[i].depth * channels_of_color(d[i].color)), im.depth = d[i].depth;
        im.interlaced = d[i].interlaced;
        im.format = d[i].standard ? PNGB200_FORMAT_IOS : PNGB200_FORMAT_ZLIB;
        images.push_back(im);
        owner.push_back(i);
    }
    if (o_total) CU(ctx->d_out.reserve(o_total));
    if (host_pixels)
        for (size_t j = 0; j < images.size(); ++j) images[j].pixels = ctx->d_out.as<uint8_t>() + o_off[owner[j]];
    CrcPlan crc_plan;
    if (optimistic) {
        // issue order: small table, bulk files, kernels, decode, and only then the small CRC download
        if ((rc = crc_upload(ctx, regions, nullptr, &crc_plan)) != PNGB200_OK) return rc;
        if ((rc = upload_files()) != PNGB200_OK) return rc;
        if ((rc = crc_launch(ctx, crc_plan)) != PNGB200_OK) return rc;
    }
    if (!images.empty()) {
        rc = pngb200_decode_batch_enqueue(ctx, images.data(), images.size(), PNGB200_MEM_DEVICE);
        if (rc != PNGB200_OK) return rc;
    }
    if (optimistic && (rc = crc_fetch(ctx, crc_plan, &crc_late)) != PNGB200_OK) return rc;
    if (!images.empty()) {
        rc = pngb200_decode_batch_finish(ctx, images.data(), images.size());
        if (rc != PNGB200_OK) return rc;
    }
    std::vector<size_t> redo;
    std::vector<char>   skip(count, 0);
    if (optimistic) {
        CU(cudaStreamSynchronize(ctx->stream));  // the CRCs have landed in pinned memory
        for (size_t i = 0; i < count; ++i) {
            const FileWalk& w = walks[i];
            for (size_t k = 0; k < w.chunks.size(); ++k) {
                if (k > w.stop || (k == w.stop && w.stop_before_crc)) break;
                const uint32_t computed = crc_late[region_base[i] + k];
                if (computed == w.chunks[k].declared) continue;
                if (k < plans[i].idat_hi && k > plans[i].idat_lo) {
                    redo.push_back(i), skip[i] = 1;  // the decoder was shown IDAT chunks it should not have seen
                } else {
                    plans[i].stop = k, plans[i].status = PNGB200_ERR_LEX_INVALID_CHUNK_CHECKSUM;
                    plans[i].a = w.chunks[k].declared, plans[i].b = computed;
                    if (k <= plans[i].idat_lo && plans[i].idat_hi > plans[i].idat_lo) skip[i] = 2;  // failed before any IDAT
                }
                break;
            }
        }
    }
    for (size_t j = 0; j < images.size(); ++j) {
        const size_t i = owner[j];
        if (skip[i]) continue;
        d[i].checksum = images[j].checksum, d[i].blocks = images[j].blocks, d[i].produced = images[j].produced;
        const bool hard = images[j].status < 0 && images[j].status != PNGB200_ERR_PNG_INCOMPLETE_DATASTREAM;
        if (hard) {
            // the decoder throws while the offending IDAT is pushed, before any later chunk is lexed
            plans[i].status = images[j].status, plans[i].a = images[j].err_a, plans[i].b = images[j].err_b;
        } else if (plans[i].status == PNGB200_OK && images[j].status != PNGB200_OK) {
            plans[i].status = images[j].status;  // IEND while the decoder still expects data
        }
        if (plans[i].status == PNGB200_OK && host_pixels)
            CU(cudaMemcpyAsync(d[i].pixels, images[j].pixels, d[i].storage_size, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < count; ++i) {
        d[i].status = plans[i].status, d[i].err_a = plans[i].a, d[i].err_b = plans[i].b;
        if (skip[i] == 2) d[i].checksum = d[i].blocks = 0, d[i].produced = 0;
    }
    if (!redo.empty()) {
        std::vector<pngb200_png_desc> again(redo.size());
        for (size_t r = 0; r < redo.size(); ++r) again[r] = d[redo[r]];
        rc = png_decode_some(ctx, again.data(), again.size(), memspace, false);
        if (rc != PNGB200_OK) return rc;
        for (size_t r = 0; r < redo.size(); ++r) {
            d[redo[r]] = again[r];
            d[redo[r]].format.palette = d[redo[r]].palette_rgba;
        }
    }
    return PNGB200_OK;
}

}  // namespace

extern "C" {

int pngb200_png_inspect_batch(pngb200_png_desc* d, size_t count)
{
    if (!d && count) return PNGB200_ERR_BAD_ARGUMENT;
    for (size_t i = 0; i < count; ++i) {
        if (!d[i].file && d[i].file_len) return PNGB200_ERR_BAD_ARGUMENT;
        FileWalk w;
        walk_file(d[i], w);
        d[i].chunks = (uint32_t)w.chunks.size();
        d[i].status = w.status, d[i].err_a = w.a, d[i].err_b = w.b;
        d[i].checksum = d[i].blocks = 0, d[i].produced = 0;
    }
    return PNGB200_OK;
}

int pngb200_png_decode_batch(pngb200_ctx* ctx, pngb200_png_desc* d, size_t count, int memspace)
{
    if (!ctx || (!d && count)) return PNGB200_ERR_BAD_ARGUMENT;
    if (ctx->pending) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "a decode batch is pending");
    if (count == 0) return PNGB200_OK;
    return run_over_lanes(ctx, count, memspace,
                          [&](size_t i) { return d[i].file_len + (size_t)0; },
                          [&](pngb200_ctx* lane, size_t lo, size_t n) { return png_decode_some(lane, d + lo, n, memspace); });
}

size_t pngb200_png_encode_bound(uint32_t width, uint32_t height, const pngb200_pixel_format* f, int interlaced, uint32_t idat_chunk)
{
    if (!f) return 0;
    const size_t chunk = idat_chunk ? idat_chunk : 65544;
    const int    volume = f->depth * channels_of_color(f->color);
    const size_t z = pngb200_deflate_bound(pngb200_filtered_size(width, height, volume, interlaced));
    return 8 + 16 + 25 + (12 + 768) + (12 + 256) + z + 12 * (z / chunk + 2) + 12 + 64;
}

int pngb200_png_encode_batch(pngb200_ctx* ctx, pngb200_png_encode_desc* d, size_t count, int memspace)
{
    if (!ctx || (!d && count)) return PNGB200_ERR_BAD_ARGUMENT;
    if (ctx->pending) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "a decode batch is pending");
    if (count == 0) return PNGB200_OK;
    DeviceGuard guard(ctx->device);
    const bool host_pixels = memspace == PNGB200_MEM_HOST;
    // stage 1: filter + deflate on the device, payload left in HBM (pngb200_encode_batch, DEVICE memspace)
    std::vector<pngb200_encode_desc> enc(count);
    std::vector<size_t> p_off(count), z_off(count), head_len(count);
    std::vector<std::vector<uint8_t>> heads(count);
    size_t p_total = 0, z_total = 0;
    for (size_t i = 0; i < count; ++i) {
        const pngb200_pixel_format& f = d[i].format;
        const int ch = f.color == 0 || f.color == 3 ? 1 : f.color == 2 ? 3 : f.color == 4 ? 2 : f.color == 6 ? 4 : 0;
        const bool depth_ok = f.color == 3 ? (f.depth == 1 || f.depth == 2 || f.depth == 4 || f.depth == 8)
                            : f.color == 0 ? (f.depth == 1 || f.depth == 2 || f.depth == 4 || f.depth == 8 || f.depth == 16)
                                           : (f.depth == 8 || f.depth == 16);
        if (!ch || !depth_ok || !d[i].width || !d[i].height || !d[i].pixels || !d[i].file ||
            (f.bgr && (f.depth != 8 || (f.color != 2 && f.color != 6))) ||
            (f.color == 3 && (!f.palette || !f.palette_count || f.palette_count > (1u << f.depth))))
            return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "image %zu: bad descriptor", i);
        const int    volume = f.depth * ch;
        const size_t storage = pngb200_storage_size(d[i].width, d[i].height, volume);
        if (d[i].pixels_len < storage) return set_error(ctx, PNGB200_ERR_BAD_ARGUMENT, "image %zu: pixels_len", i);
        if (d[i].file_cap < pngb200_png_encode_bound(d[i].width, d[i].height, &f, d[i].interlaced, d[i].idat_chunk))
            return set_error(ctx, PNGB200_ERR_OUTPUT_CAPACITY, "image %zu: file_cap below pngb200_png_encode_bound", i);
        p_off[i] = p_total;
        p_total += align_up(storage + 16, 256);
        z_off[i] = z_total;
        z_total += align_up(pngb200_deflate_bound(pngb200_filtered_size(d[i].width, d[i].height, volume, d[i].interlaced)) + 16, 256);
    }
    if (host_pixels) CU(ctx->d_in.reserve(p_total));
    CU(ctx->d_out.reserve(z_total));
    for (size_t i = 0; i < count; ++i) {
        const pngb200_pixel_format& f = d[i].format;
        const int    volume = f.depth * channels_of_color(f.color);
        const size_t storage = pngb200_storage_size(d[i].width, d[i].height, volume);
        pngb200_encode_desc& e = enc[i];
        memset(&e, 0, sizeof e);
        if (host_pixels) {
            CU(cudaMemcpyAsync(ctx->d_in.as<uint8_t>() + p_off[i], d[i].pixels, storage, cudaMemcpyHostToDevice, ctx->stream));
            e.pixels = ctx->d_in.as<uint8_t>() + p_off[i];
        } else {
            e.pixels = (const uint8_t*)d[i].pixels;
        }
        e.pixels_len = storage;
        e.idat = ctx->d_out.as<uint8_t>() + z_off[i];
        e.idat_cap = pngb200_deflate_bound(pngb200_filtered_size(d[i].width, d[i].height, volume, d[i].interlaced));
        e.width = d[i].width, e.height = d[i].height;
        e.volume = (uint8_t)volume, e.depth = f.depth, e.interlaced = d[i].interlaced;
        e.format = f.bgr ? PNGB200_FORMAT_IOS : PNGB200_FORMAT_ZLIB;
        e.level = d[i].level;
        // everything in front of the first IDAT, built on the host (a few hundred bytes):
        // signature, [CgBI], IHDR, [PLTE], [tRNS]  (PNG.Image.swift:580-629, PNG.Image.encode :416-423,
        // Layout.palette / Layout.transparency, Formats/PNG.Layout.swift:43-135)
        std::vector<uint8_t>& h = heads[i];
        auto put = [&](uint32_t type, const uint8_t* body, size_t n) {
            const size_t at = h.size();
            h.resize(at + 12 + n);
            store_be32(h.data() + at, (uint32_t)n), store_be32(h.data() + at + 4, type);
            if (n) memcpy(h.data() + at + 8, body, n);
            uint32_t c = 0xffffffffu;  // these few bytes are CRC'd where they are built
            for (size_t k = at + 4; k < at + 8 + n; ++k) {
                c ^= h[k];
                for (int b = 0; b < 8; ++b) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            }
            store_be32(h.data() + at + 8 + n, ~c);
        };
        h.assign(PNG_SIGNATURE, PNG_SIGNATURE + 8);
        if (f.bgr) {
            const uint8_t cgbi[4] = {48, 0, 32, (uint8_t)(f.color == 2 ? 6 : 2)};
            put(CK_CgBI, cgbi, 4);
        }
        uint8_t ihdr[13];
        store_be32(ihdr, d[i].width), store_be32(ihdr + 4, d[i].height);
        ihdr[8] = f.depth, ihdr[9] = f.color, ihdr[10] = 0, ihdr[11] = 0, ihdr[12] = d[i].interlaced ? 1 : 0;
        put(CK_IHDR, ihdr, 13);
        if (f.color == 3) {
            uint8_t rgb[768], alpha[256];
            int last = -1;
            for (int k = 0; k < f.palette_count; ++k) {
                memcpy(rgb + 3 * k, f.palette + 4 * k, 3);
                alpha[k] = f.palette[4 * k + 3];
                if (alpha[k] != 255) last = k;
            }
            put(CK_PLTE, rgb, 3 * (size_t)f.palette_count);
            if (last >= 0) put(CK_tRNS, alpha, (size_t)last + 1);
        } else if (f.has_key && (f.color == 0 || f.color == 2)) {
            uint8_t k[6];
            if (f.color == 0) {
                k[0] = (uint8_t)(f.key[0] >> 8), k[1] = (uint8_t)f.key[0];
                put(CK_tRNS, k, 2);
            } else {
                const uint16_t r = f.bgr ? f.key[2] : f.key[0], g = f.key[1], b = f.bgr ? f.key[0] : f.key[2];
                k[0] = (uint8_t)(r >> 8), k[1] = (uint8_t)r, k[2] = (uint8_t)(g >> 8), k[3] = (uint8_t)g, k[4] = (uint8_t)(b >> 8), k[5] = (uint8_t)b;
                put(CK_tRNS, k, 6);
            }
        }
        head_len[i] = h.size();
    }
    int rc = pngb200_encode_batch(ctx, enc.data(), count, PNGB200_MEM_DEVICE);
    if (rc != PNGB200_OK) return rc;
    // stage 2: frame the payload into IDAT chunks inside a device image of the file, CRC them there
    std::vector<size_t> file_off(count), file_len(count);
    size_t file_total = 0;
    std::vector<FrameItem>   frames;
    std::vector<CrcRegion>   regions;
    std::vector<CopySegment> segs;
    for (size_t i = 0; i < count; ++i) {
        d[i].status = enc[i].status, d[i].checksum = enc[i].checksum, d[i].blocks = enc[i].blocks, d[i].produced = 0;
        file_off[i] = file_total;
        if (enc[i].status != PNGB200_OK) { file_len[i] = 0; continue; }
        const size_t chunk = d[i].idat_chunk ? d[i].idat_chunk : 65544;
        const size_t z = enc[i].produced, nchunks = (z + chunk - 1) / chunk;
        file_len[i] = head_len[i] + z + 12 * nchunks + 12;
        file_total += align_up(file_len[i] + 16, 256);
    }
    CU(ctx->d_file.reserve(std::max<size_t>(file_total, 256)));
    for (size_t i = 0; i < count; ++i) {
        if (enc[i].status != PNGB200_OK) continue;
        uint8_t* base = ctx->d_file.as<uint8_t>() + file_off[i];
        CU(cudaMemcpyAsync(base, heads[i].data(), head_len[i], cudaMemcpyHostToDevice, ctx->stream));
        const size_t chunk = d[i].idat_chunk ? d[i].idat_chunk : 65544;
        // Encoder.pull hands out DeflatorOut's queued buffers (2 x capacity bytes each), then the rest
        // (Encoding/PNG.Encoder.swift:33-129, Deflator/LZ77.DeflatorOut.swift:73-125)
        size_t at = head_len[i];
        for (size_t o = 0; o < enc[i].produced; o += chunk) {
            const size_t n = std::min(chunk, (size_t)enc[i].produced - o);
            frames.push_back({base + at, (uint32_t)n, CK_IDAT});
            segs.push_back({enc[i].idat + o, base + at + 8, n});
            regions.push_back({base + at + 8, n, CK_IDAT, 1});
            at += 12 + n;
        }
        frames.push_back({base + at, 0, CK_IEND});
        regions.push_back({base + at + 8, 0, CK_IEND, 1});
    }
    CU(cudaStreamSynchronize(ctx->stream));  // `heads` is pageable host memory read by the copies above
    rc = run_segment_copy(ctx, segs);
    if (rc != PNGB200_OK) return rc;
    if (!frames.empty()) {
        const size_t fb = sizeof(FrameItem) * frames.size(), off_crc = align_up(fb, 256), cb = sizeof(uint32_t) * frames.size();
        CU(ctx->h_genjobs.reserve(fb));
        CU(ctx->d_genjobs.reserve(off_crc + cb));
        memcpy(ctx->h_genjobs.p, frames.data(), fb);
        CU(cudaMemcpyAsync(ctx->d_genjobs.p, ctx->h_genjobs.p, fb, cudaMemcpyHostToDevice, ctx->stream));
        uint32_t* d_crc = (uint32_t*)((char*)ctx->d_genjobs.p + off_crc);
        const unsigned blocks = (unsigned)((frames.size() + 255) / 256);
        frame_chunks_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_genjobs.as<FrameItem>(), d_crc, (uint32_t)frames.size(), 0);
        ctx->launches++;
        rc = run_crc(ctx, regions, d_crc, nullptr);  // the chunk type is folded in as a 4-byte prefix
        if (rc != PNGB200_OK) return rc;
        frame_chunks_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_genjobs.as<FrameItem>(), d_crc, (uint32_t)frames.size(), 1);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    for (size_t i = 0; i < count; ++i) {
        if (enc[i].status != PNGB200_OK) continue;
        CU(cudaMemcpyAsync(d[i].file, ctx->d_file.as<uint8_t>() + file_off[i], file_len[i], cudaMemcpyDeviceToHost, ctx->stream));
        d[i].produced = file_len[i];
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return PNGB200_OK;
}

}  // extern "C"
