//  Replacement bodies for swift-png's LZ77 module on top of libpngb200 (include/pngb200.h).
//  Written against the C ABI; NOT compiled in the pngb200 repository (its build image has no Swift toolchain).
//  The same entry points, in the same order, are exercised by the ctypes mirror swift-png_b200/__init__.py and
//  by tests/test_gpu_*.py.
import CPNGB200

extension LZ77
{
    /// One process-wide context per GPU.  Handles are used by one thread at a time, exactly like the value
    /// types they replace.
    final
    class GPU
    {
        static let shared:GPU = .init()
        let ctx:OpaquePointer

        init()
        {
            guard let ctx:OpaquePointer = pngb200_ctx_create(-1)
            else
            {
                fatalError(String.init(cString: pngb200_last_error(nil)))
            }
            self.ctx = ctx
        }
        deinit
        {
            pngb200_ctx_destroy(self.ctx)
        }
    }
}

extension LZ77.Format
{
    /// pngb200_format
    var code:Int32
    {
        switch self
        {
        case .zlib: 0
        case .ios:  1
        }
    }
}

/// status + payload -> the reference's typed errors (LZ77.DecompressionError.swift:19-60,
/// LZ77.StreamHeaderError.swift:5-28, Gzip.StreamHeaderError)
func pngb200Error(status:Int32, _ a:UInt32, _ b:UInt32) -> any Error
{
    switch status
    {
    case -1:  LZ77.DecompressionError.invalidStreamChecksum(declared: a, computed: b)
    case -2:  LZ77.DecompressionError.invalidBlockTypeCode(UInt8.init(a))
    case -3:  LZ77.DecompressionError.invalidBlockElementCountParity(UInt16.init(a), UInt16.init(b))
    case -4:  LZ77.DecompressionError.invalidHuffmanRunLiteralSymbolCount(Int.init(a))
    case -5:  LZ77.DecompressionError.invalidHuffmanCodelengthHuffmanTable
    case -6:  LZ77.DecompressionError.invalidHuffmanCodelengthSequence
    case -7, -9: LZ77.DecompressionError.invalidHuffmanTable
    case -8:  LZ77.DecompressionError.invalidStringReference
    case -16: LZ77.StreamHeaderError.invalidCompressionMethod(UInt8.init(a))
    case -17: LZ77.StreamHeaderError.invalidWindowSize(exponent: Int.init(a))
    case -18: LZ77.StreamHeaderError.invalidCheckBits
    case -19: LZ77.StreamHeaderError.unexpectedDictionary
    case -32: Gzip.StreamHeaderError.invalidSigil
    case -33: Gzip.StreamHeaderError.invalidCompressionMethod(UInt8.init(a))
    case -34: Gzip.StreamHeaderError.invalidFlagBits(UInt8.init(a))
    case -35: Gzip.StreamHeaderError._headerChecksumUnsupported
    default:  fatalError("pngb200 status \(status)")
    }
}
