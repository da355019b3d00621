//
//  AirwaveCUDA.swift — Swift shim over libairwave_cuda.so (module CAirwaveCUDA, include/module.modulemap).
//
//  Keeps the reference's type names and signatures for the binaural path so the reference's call sites and XCTest
//  files compile unchanged against the B200 back end:
//      ConvolutionEngine(hrirSamples:blockSize:)           Airwave/ConvolutionEngine.swift:68
//      ConvolutionEngine.process(input:output:) / reset()  Airwave/ConvolutionEngine.swift:232, 397
//      RealtimeAudioProcessor(renderers:blockSize:maxFramesPerCallback:) / process / reset
//                                                          Airwave/RealtimeAudioProcessor.swift:30, 77, 121
//      StereoAudioProcessing.process(inputLeft:inputRight:outputLeft:outputRight:frameCount:)
//                                                          Airwave/AudioPipeline.swift:3-11
//  NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Swift toolchain.  The C ABI underneath is what the
//  tests exercise (tests/test_gpu_*.py through ctypes, include/airwave.hpp through C++).
//
import CAirwaveCUDA

public protocol StereoAudioProcessing: AnyObject {
    func process(inputLeft: UnsafePointer<Float>, inputRight: UnsafePointer<Float>?,
                 outputLeft: UnsafeMutablePointer<Float>, outputRight: UnsafeMutablePointer<Float>, frameCount: Int)
}

/// Mono partitioned convolver: one engine with one stream, one speaker whose two "ears" are the same impulse response.
public final class ConvolutionEngine {
    public let blockSize: Int
    let hrirSamples: [Float]
    private var bank: OpaquePointer?
    private var engine: OpaquePointer?
    private var scratch: [Float]

    public init?(hrirSamples: [Float], blockSize: Int = 512, device: Int32 = 0) {
        self.blockSize = blockSize
        self.hrirSamples = hrirSamples
        self.scratch = [Float](repeating: 0, count: 2 * blockSize)
        var idx: Int32 = 0
        let rc = hrirSamples.withUnsafeBufferPointer { pcm in
            aw_bank_create(device, pcm.baseAddress, 1, Int32(hrirSamples.count), 48_000, 48_000, &idx, &idx, 1,
                           Int32(blockSize), &bank)
        }
        guard rc == AW_OK.rawValue else { return nil }          // init? -> nil (ConvolutionEngine.swift:82-85)
        var partitions: Int32 = 0
        aw_bank_info(bank, nil, nil, &partitions, nil)
        var cfg = aw_engine_config(device: device, n_streams: 1, n_speakers: 1, block: Int32(blockSize), sample_rate: 48_000,
                                   max_frames_per_call: Int32(blockSize), max_partitions: partitions, flags: 0)
        guard aw_engine_create(&cfg, &engine) == AW_OK.rawValue, aw_engine_set_bank(engine, 0, 1, bank) == AW_OK.rawValue
        else { return nil }
    }

    deinit { aw_engine_destroy(engine); aw_bank_destroy(bank) }

    /// Allocation-free: all device and staging memory was reserved in init.
    public func process(input: UnsafePointer<Float>, output: UnsafeMutablePointer<Float>) {
        scratch.withUnsafeMutableBufferPointer { out in
            _ = aw_engine_process(engine, input, out.baseAddress, Int32(blockSize))
            output.update(from: out.baseAddress!, count: blockSize)          // left ear == right ear
        }
    }

    public func process(input: [Float], output: inout [Float], frameCount: Int? = nil) {
        guard (frameCount ?? blockSize) == blockSize else { return }          // ConvolutionEngine.swift:370-372
        input.withUnsafeBufferPointer { i in output.withUnsafeMutableBufferPointer { o in
            process(input: i.baseAddress!, output: o.baseAddress!) } }
    }

    public func reset() { _ = aw_engine_reset(engine, 0, 1, Int32(AW_RESET_SPATIAL)) }
}

public struct VirtualSpeakerRenderer {
    public let convolverLeftEar: ConvolutionEngine
    public let convolverRightEar: ConvolutionEngine
}

/// Frame adapter + renderers for one stereo stream; the renderers' impulse responses are gathered into one filter bank.
public final class RealtimeAudioProcessor: StereoAudioProcessing {
    public let blockSize: Int
    public let maxFramesPerCallback: Int
    private var bank: OpaquePointer?
    private var engine: OpaquePointer?

    public init(renderers: [VirtualSpeakerRenderer], blockSize: Int = 512, maxFramesPerCallback: Int = 4096, device: Int32 = 0) {
        precondition(blockSize > 0 && maxFramesPerCallback > 0)
        self.blockSize = blockSize
        self.maxFramesPerCallback = maxFramesPerCallback
        var cfg = aw_engine_config(device: device, n_streams: 1, n_speakers: 2, block: Int32(blockSize), sample_rate: 48_000,
                                   max_frames_per_call: Int32(maxFramesPerCallback), max_partitions: 0,
                                   flags: UInt32(AW_ENGINE_LITERAL_STEREO))
        precondition(aw_engine_create(&cfg, &engine) == AW_OK.rawValue, String(cString: aw_last_error()))
        let used = min(renderers.count, 2)                                     // RealtimeAudioProcessor.swift:145
        guard used > 0 else { return }
        let taps = renderers.prefix(used).map { max($0.convolverLeftEar.hrirSamples.count, $0.convolverRightEar.hrirSamples.count) }.max()!
        var pcm = [Float](repeating: 0, count: 2 * used * taps)
        for (i, r) in renderers.prefix(used).enumerated() {
            pcm.replaceSubrange((2 * i) * taps ..< (2 * i) * taps + r.convolverLeftEar.hrirSamples.count, with: r.convolverLeftEar.hrirSamples)
            pcm.replaceSubrange((2 * i + 1) * taps ..< (2 * i + 1) * taps + r.convolverRightEar.hrirSamples.count, with: r.convolverRightEar.hrirSamples)
        }
        var left = (0..<used).map { Int32(2 * $0) }, right = (0..<used).map { Int32(2 * $0 + 1) }
        precondition(aw_bank_create(device, pcm, Int32(2 * used), Int32(taps), 48_000, 48_000, &left, &right, Int32(used),
                                    Int32(blockSize), &bank) == AW_OK.rawValue, String(cString: aw_last_error()))
        precondition(aw_engine_set_bank(engine, 0, 1, bank) == AW_OK.rawValue)
    }

    deinit { aw_engine_destroy(engine); aw_bank_destroy(bank) }

    public func process(inputLeft: UnsafePointer<Float>, inputRight: UnsafePointer<Float>?,
                        outputLeft: UnsafeMutablePointer<Float>, outputRight: UnsafeMutablePointer<Float>, frameCount: Int) {
        guard frameCount > 0 else { return }
        precondition(frameCount <= maxFramesPerCallback)                       // RealtimeAudioProcessor.swift:85
        _ = aw_engine_process_stereo(engine, inputLeft, inputRight, outputLeft, outputRight, Int32(frameCount))
    }

    public func reset() { _ = aw_engine_reset(engine, 0, 1, Int32(AW_RESET_SPATIAL)) }
}
