//  Batched entry points beside PNG.Image.decompress / .compress (Sources/PNG/PNG.Image.swift:298-401, 576-670):
//  one pngb200 call per batch, every image an independent unit.  The single-image streaming API keeps working
//  through the LZ77.Inflator / LZ77.Deflator replacements (PNG.Decoder.push and PNG.Encoder.pull are unchanged
//  source-wise: they call push / pull / pop exactly as before).
import CPNGB200

extension PNG.Image
{
    /// PNG.Decoder.push over a batch (Sources/PNG/Decoding/PNG.Decoder.swift:47-149): inflate + defilter +
    /// assign on the device; returns each image's `storage`, byte for byte.
    static
    func decompress(batch:[(header:PNG.Header, layout:PNG.Layout, idat:[UInt8])]) throws -> [[UInt8]]
    {
        var storages:[[UInt8]] = batch.map
        {
            .init(repeating: 0,
                count: $0.header.size.x * $0.header.size.y * (($0.layout.format.pixel.volume + 7) >> 3))
        }
        var descs:[pngb200_image_desc] = .init(repeating: .init(), count: batch.count)
        for i:Int in batch.indices
        {
            descs[i].width      = UInt32.init(batch[i].header.size.x)
            descs[i].height     = UInt32.init(batch[i].header.size.y)
            descs[i].volume     = UInt8.init(batch[i].layout.format.pixel.volume)
            descs[i].depth      = UInt8.init(batch[i].layout.format.pixel.depth)
            descs[i].interlaced = batch[i].layout.interlaced ? 1 : 0
            descs[i].format     = 0 // 1 for PNG.Standard.ios (raw deflate)
            descs[i].idat_len   = batch[i].idat.count
            descs[i].pixels_cap = storages[i].count
        }
        // (pointers are pinned for the duration of the call: withUnsafeBufferPointer on every idat / storage)
        let rc:Int32 = pngb200_decode_batch(LZ77.GPU.shared.ctx, &descs, descs.count, Int32.init(PNGB200_MEM_HOST.rawValue))
        precondition(rc == 0, String.init(cString: pngb200_last_error(LZ77.GPU.shared.ctx)))
        for d:pngb200_image_desc in descs
        {
            switch d.status
            {
            case 0:     continue
            case -48:   throw PNG.DecodingError.extraneousImageData                      // PNG.Decoder.swift:142-147
            case -49:   throw PNG.DecodingError.extraneousImageDataCompressedData         // :51-55
            case -50:   throw PNG.DecodingError.incompleteImageDataCompressedDatastream   // PNG.Context.swift:134-141
            default:    throw pngb200Error(status: d.status, d.err_a, d.err_b)
            }
        }
        return storages
    }
}
