//
//  AirwaveCUDAEffects.swift — the two effects AudioEffectGraph composes, backed by libairwave_cuda.so.
//
//  Add this file to the Airwave app target next to AudioEffectGraph.swift (it uses the app's own declarations:
//  AudioSpatialEffect / AudioEqualizerEffect  Airwave/AudioEffectGraph.swift:47-54,
//  EqualizerDefinition / EqualizerFilter      Airwave/EqualizerPreset.swift:9-27,
//  EqualizerAudioEffectError                  Airwave/AudioEffectGraph.swift:28-45) and construct the graph with
//
//      AudioEffectGraph(spatial: CUDASpatialEffect(), equalizer: CUDAEqualizerEffect())
//
//  instead of  AudioEffectGraph(spatial: HRIRManager.shared, equalizer: EqualizerManager.shared.runtimeEffect)
//  (Airwave/AudioRuntimeController.swift:62).  Render calls are allocation-free and never block on a lock: everything
//  the engines need was reserved by activatePreset / prepare on the control thread.
//  NOT COMPILED IN THIS REPOSITORY'S CI (no Swift toolchain in the build image); the C ABI underneath is what the tests
//  exercise.  INTEGRATION.md has the entry-point table.
//
import CAirwaveCUDA
import Foundation

/// Stands in for HRIRManager as the graph's spatial effect (HRIRManager.swift:517-568): stereo in, binaural out.
final class CUDASpatialEffect: AudioSpatialEffect {
    private var engine: OpaquePointer?
    private var bank: OpaquePointer?
    private let device: Int32
    private let blockSize: Int32
    private(set) var isReady = false                       // hasPublishedRendererForAudioCallback

    init(device: Int32 = 0, blockSize: Int32 = 512, maxFramesPerCallback: Int32 = 4096) {   // processingBlockSize, HRIRManager.swift:149
        self.device = device
        self.blockSize = blockSize
        var cfg = aw_engine_config(device: device, n_streams: 1, n_speakers: 2, block: blockSize, sample_rate: 48_000,
                                   max_frames_per_call: maxFramesPerCallback, max_partitions: 0,
                                   flags: UInt32(AW_ENGINE_LITERAL_STEREO))
        precondition(aw_engine_create(&cfg, &engine) == AW_OK.rawValue, String(cString: aw_last_error()))
    }

    deinit { aw_engine_destroy(engine); aw_bank_destroy(bank) }

    /// HRIRManager.activatePreset(_:targetSampleRate:inputLayout:hrirMap:) (HRIRManager.swift:316-449) for the stereo layout the
    /// app always uses (DeviceProfileRuntimeCoordinator.swift:104-108).  Control thread only.
    func activatePreset(fileURL: URL, targetSampleRate: Double) throws {
        var wav: OpaquePointer?
        guard aw_wav_load(fileURL.path, &wav) == AW_OK.rawValue else { throw HRIRError.convolutionSetupFailed(String(cString: aw_last_error())) }
        defer { aw_wav_destroy(wav) }
        var newBank: OpaquePointer?
        guard aw_bank_create_from_wav(device, wav, targetSampleRate, Int32(AW_LAYOUT_STEREO.rawValue), blockSize, &newBank) == AW_OK.rawValue
        else { throw HRIRError.convolutionSetupFailed(String(cString: aw_last_error())) }
        guard aw_engine_set_bank(engine, 0, 1, newBank) == AW_OK.rawValue else {
            aw_bank_destroy(newBank)
            throw HRIRError.convolutionSetupFailed(String(cString: aw_last_error()))
        }
        aw_bank_destroy(bank)                               // the previous renderer state is retired whole
        bank = newBank
        isReady = true
    }

    func deactivatePreset() {
        _ = aw_engine_set_bank(engine, 0, 1, nil)           // passthrough (HRIRManager.swift:555-564)
        aw_bank_destroy(bank)
        bank = nil
        isReady = false
    }

    func process(inputLeft: UnsafePointer<Float>, inputRight: UnsafePointer<Float>?,
                 outputLeft: UnsafeMutablePointer<Float>, outputRight: UnsafeMutablePointer<Float>, frameCount: Int) {
        guard frameCount > 0 else { return }
        _ = aw_engine_process_stereo(engine, inputLeft, inputRight, outputLeft, outputRight, Int32(frameCount))
    }
}

/// Stands in for EqualizerRuntimeEffect (EqualizerRuntimeEffect.swift:5-101): an engine without renderers copies its input and
/// runs the parametric EQ — Double TDF-II cascade, 20 ms crossfade to new targets, deferred retirement — in place.
final class CUDAEqualizerEffect: AudioEqualizerEffect {
    private var engine: OpaquePointer?
    private var preparedRate = 0.0
    private let device: Int32
    private let maxFramesPerCallback: Int32

    init(device: Int32 = 0, maxFramesPerCallback: Int32 = 4096) {
        self.device = device
        self.maxFramesPerCallback = maxFramesPerCallback
    }

    deinit { aw_engine_destroy(engine) }

    private static func pack(_ definition: EqualizerDefinition?) -> (Double, [aw_eq_filter], Int32) {
        guard let definition else { return (0, [], -1) }   // nil definition: unity state (ParametricEqualizerProcessor.swift:182,191)
        let filters = definition.filters.map { f -> aw_eq_filter in
            let type: aw_filter_type
            switch f.type {
            case .peaking: type = AW_FILTER_PEAKING
            case .lowShelf: type = AW_FILTER_LOW_SHELF
            case .highShelf: type = AW_FILTER_HIGH_SHELF
            }
            return aw_eq_filter(type: Int32(type.rawValue), enabled: f.isEnabled ? 1 : 0, frequency_hz: f.frequencyHz, gain_db: f.gainDB,
                                q: f.q, source_line: Int32(f.sourceLine), source_number: Int32(f.sourceNumber ?? -1))
        }
        return (definition.preampDB, filters, Int32(filters.count))
    }

    private func mapError(_ status: Int32, _ definition: EqualizerDefinition?, _ badIndex: Int32) -> EqualizerAudioEffectError {
        switch status {
        case Int32(AW_ERR_EQ_INVALID_SAMPLE_RATE.rawValue): return .invalidSampleRate
        case Int32(AW_ERR_NOT_READY.rawValue): return .unavailable(String(cString: aw_last_error()))
        default:                                            // EqualizerRuntimeEffect.swift:69-100
            let enabled = definition?.filters.filter(\.isEnabled) ?? []
            let line = badIndex >= 0 && Int(badIndex) < enabled.count ? enabled[Int(badIndex)].sourceLine : nil
            return .invalidFilter(line: line, reason: String(cString: aw_last_error()))
        }
    }

    /// EqualizerRuntimeEffect.prepare (:10-34): a processor per output sample rate.  Control thread only.
    func prepare(definition: EqualizerDefinition?, sampleRate: Double) throws {
        if engine == nil || preparedRate != sampleRate {
            aw_engine_destroy(engine)
            engine = nil
            var cfg = aw_engine_config(device: device, n_streams: 1, n_speakers: 2, block: 512, sample_rate: sampleRate,
                                       max_frames_per_call: maxFramesPerCallback, max_partitions: 0, flags: 0)
            guard aw_engine_create(&cfg, &engine) == AW_OK.rawValue else { throw EqualizerAudioEffectError.invalidSampleRate }
            preparedRate = sampleRate
        }
        var (preamp, filters, n) = Self.pack(definition)
        var bad: Int32 = -1, why: Int32 = 0
        let rc = aw_engine_eq_prepare(engine, 0, 1, preamp, &filters, n, &bad, &why)
        guard rc == AW_OK.rawValue else { throw mapError(rc, definition, bad) }
    }

    /// EqualizerRuntimeEffect.setTarget (:36-48): publishes a new target; the render thread crossfades to it over 20 ms.
    func setTarget(definition: EqualizerDefinition?) throws {
        guard engine != nil else { throw EqualizerAudioEffectError.unavailable("Equalizer has not been prepared for an output.") }
        var (preamp, filters, n) = Self.pack(definition)
        var bad: Int32 = -1, why: Int32 = 0
        let rc = aw_engine_eq_update(engine, 0, 1, preamp, &filters, n, &bad, &why)
        guard rc == AW_OK.rawValue else { throw mapError(rc, definition, bad) }
    }

    func process(inputLeft: UnsafePointer<Float>, inputRight: UnsafePointer<Float>?,
                 outputLeft: UnsafeMutablePointer<Float>, outputRight: UnsafeMutablePointer<Float>, frameCount: Int) {
        guard frameCount > 0, engine != nil else { return }
        _ = aw_engine_process_stereo(engine, inputLeft, inputRight, outputLeft, outputRight, Int32(frameCount))
    }
}
