/*
 * airwave_cuda.h — C ABI of the B200 batched binaural renderer (libairwave_cuda.so).
 *
 * This is the drop-in boundary for ONE path of sallliisa/Airwave: HRIR convolution
 * (ConvolutionEngine / RealtimeAudioProcessor / HRIRManager.processAudio) -> L/R downmix ->
 * parametric EQ (ParametricEqualizerProcessor), behind AudioEffectGraph.  The reference has no
 * FFI today (it calls Apple vDSP in-process); a Swift shim keeps the reference's type names and
 * calls these entry points through `module CAirwaveCUDA` (include/module.modulemap), see
 * INTEGRATION.md.  Each entry point cites the reference interface it replaces (paths relative
 * to the reference checkout).
 *
 * Conventions (mirroring SURVEY.md 8(b)):
 *  - plain pointers and sizes only; opaque handles; every function returns an aw_status
 *    (0 = AW_OK) unless stated otherwise; nothing aborts or throws across the boundary;
 *  - audio is planar float32, caller-owned; a handle is thread-compatible (one render thread
 *    calls aw_engine_process*, control calls must not overlap it), not thread-safe;
 *  - all device and pinned memory is reserved by the *_create functions: aw_engine_process*
 *    performs no allocation, no logging and no blocking lock (the reference's real-time rule,
 *    scripts/check-audio-safety-invariants.sh:23-41);
 *  - there is NO CPU fallback: without a CUDA device every *_create fails with AW_ERR_CUDA.
 */
#ifndef AIRWAVE_CUDA_H
#define AIRWAVE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define AW_API __attribute__((visibility("default")))
#else
#define AW_API
#endif

typedef enum aw_status {
    AW_OK = 0,
    AW_ERR_INVALID_ARGUMENT = 1,
    AW_ERR_CUDA = 2,               /* CUDA runtime/driver error or no device; see aw_last_error() */
    AW_ERR_OUT_OF_MEMORY = 3,
    AW_ERR_INVALID_BLOCK_SIZE = 4, /* ConvolutionEngine.init? -> nil: block must be a power of two (Q10) */
    AW_ERR_FRAME_COUNT = 5,        /* precondition(frameCount <= maxFramesPerCallback), RealtimeAudioProcessor.swift:85 */
    AW_ERR_CHANNEL_MAPPING = 6,    /* HRIRError.invalidChannelMapping, HRIRManager.swift:375-379 */
    AW_ERR_NO_RENDERERS = 7,       /* HRIRError.convolutionSetupFailed("No valid renderers created"), :420-422 */
    AW_ERR_RANGE = 8,              /* stream range outside [0, n_streams) or not tile-aligned */
    AW_ERR_MISMATCH = 9,           /* bank/engine disagree on device, block size or speaker count */
    AW_ERR_UNSUPPORTED = 10,
    AW_ERR_RESAMPLE_DOWN = 11,     /* down-sampling: the reference reads past its control vector (Q7) */
    AW_ERR_NOT_READY = 12,         /* EqualizerAudioEffectError.unavailable("Equalizer has not been prepared for an output.") */
    /* ParametricEqualizerPreparationError, ParametricEqualizerProcessor.swift:100-118 */
    AW_ERR_EQ_INVALID_SAMPLE_RATE = 20,
    AW_ERR_EQ_NON_FINITE_PREAMP = 21,
    AW_ERR_EQ_TOO_MANY_FILTERS = 22,
    AW_ERR_EQ_INVALID_FILTER = 23, /* *bad_index = index among enabled filters, *bad_reason = aw_biquad_error */
    /* WAVError, WAVLoader.swift:132-147 */
    AW_ERR_WAV_READ = 30,
    AW_ERR_WAV_CHANNEL_COUNT = 31,
    AW_ERR_WAV_EMPTY = 32,
    AW_ERR_WAV_UNSUPPORTED_FORMAT = 33,
    /* EqualizerParseError, EqualizerAPOParser.swift:8-21 */
    AW_ERR_EQ_PARSE = 40
} aw_status;

/* BiquadCoefficientError, BiquadCoefficientBuilder.swift:11-27 */
typedef enum aw_biquad_error {
    AW_BIQUAD_OK = 0,
    AW_BIQUAD_INVALID_SAMPLE_RATE = 1,
    AW_BIQUAD_INVALID_FREQUENCY = 2,
    AW_BIQUAD_INVALID_Q = 3,
    AW_BIQUAD_NON_FINITE_INPUT = 4,
    AW_BIQUAD_NON_FINITE_COEFFICIENTS = 5
} aw_biquad_error;

/* EqualizerFilterType, EqualizerPreset.swift:3-7 */
typedef enum aw_filter_type { AW_FILTER_PEAKING = 0, AW_FILTER_LOW_SHELF = 1, AW_FILTER_HIGH_SHELF = 2 } aw_filter_type;

/* VirtualSpeaker, VirtualSpeaker.swift:11-35 (custom speakers are not representable: they never map) */
typedef enum aw_speaker {
    AW_SPK_FL = 0, AW_SPK_FR, AW_SPK_FC, AW_SPK_LFE, AW_SPK_BL, AW_SPK_BR, AW_SPK_SL, AW_SPK_SR,
    AW_SPK_TFL, AW_SPK_TFR, AW_SPK_TBL, AW_SPK_TBR, AW_SPK_FLC, AW_SPK_FRC, AW_SPK_BC, AW_SPK_COUNT
} aw_speaker;

/* InputLayout, VirtualSpeaker.swift:59-100 */
typedef enum aw_input_layout { AW_LAYOUT_STEREO = 2, AW_LAYOUT_SURROUND51 = 6, AW_LAYOUT_SURROUND71 = 8, AW_LAYOUT_ATMOS714 = 12 } aw_input_layout;

/* EqualizerFilter, EqualizerPreset.swift:9-17 */
typedef struct aw_eq_filter {
    int32_t type;          /* aw_filter_type */
    int32_t enabled;       /* isEnabled */
    double frequency_hz;
    double gain_db;
    double q;
    int32_t source_line;   /* sourceLine (1-based), 0 when unknown */
    int32_t source_number; /* sourceNumber, -1 when absent */
} aw_eq_filter;

typedef struct aw_wav aw_wav;       /* WAVData, WAVLoader.swift:12-17 */
typedef struct aw_bank aw_bank;     /* frequency-domain HRIR filter bank resident in HBM (replaces per-engine hrirReal/Imag) */
typedef struct aw_engine aw_engine; /* batched renderer: n_streams x (RendererState + RealtimeAudioProcessor + EQ) */

typedef struct aw_engine_config {
    int32_t device;              /* CUDA ordinal */
    int32_t n_streams;           /* independent audio streams rendered in lock-step */
    int32_t n_speakers;          /* S: input channels per stream (InputLayout.channels.count) */
    int32_t block;               /* B: processingBlockSize (HRIRManager.swift:149 uses 512); power of two, 4..8192 */
    double sample_rate;          /* output.nominalSampleRate: EQ coefficients and the 20 ms crossfade use it */
    int32_t max_frames_per_call; /* maxFramesPerCallback (default 4096 when 0) */
    int32_t max_partitions;      /* FDL capacity per (stream, speaker); 0 = size on first aw_engine_set_bank */
    uint32_t flags;              /* AW_ENGINE_* */
} aw_engine_config;

#define AW_ENGINE_DEFAULT 0u
#define AW_ENGINE_LITERAL_STEREO 1u /* reference-literal RealtimeAudioProcessor: min(renderers, 2), inputs = (left, right) (Q1) */
#define AW_ENGINE_PIPELINED 2u      /* reserve a second staging set so aw_engine_submit can overlap copies and kernels */
/* The equalizer of a call (AudioEffectGraph runs it after the spatial effect, AudioEffectGraph.swift:195-210) is launched on an
 * internal stream and runs next to the convolution of the NEXT call.  Nothing changes for aw_engine_process,
 * aw_engine_process_stereo, aw_engine_submit, aw_engine_submit_device and aw_engine_wait.  For aw_engine_process_device the output of call j is complete,
 * in the order of aw_engine_stream(), once the work of call j+1 is, or after aw_engine_flush(): consecutive calls must therefore
 * write to different output buffers (two, alternating, are enough). */
#define AW_ENGINE_OVERLAP_EQ 4u

/* ---- library / device ------------------------------------------------------------------- */
AW_API const char *aw_version(void);
AW_API const char *aw_status_string(int status);
/* Message of the last failing call on this thread (empty string when none). */
AW_API const char *aw_last_error(void);
AW_API int aw_device_count(void);

/* ---- FFT plan cache: replaces FFTSetupManager.getSetup / getCacheStats (FFTSetupManager.swift:41-69)
 * A plan = the twiddle table for real length 2^log2n, resident on `device`, shared by all engines. */
AW_API int aw_plan_prepare(int device, int log2n);
AW_API int aw_plan_cache_stats(int device, int *count, int *sizes, int capacity);

/* ---- WAVLoader.load (WAVLoader.swift:26-99): RIFF/WAVE incl. EXTENSIBLE; float32/float64, int16/24/32 */
AW_API int aw_wav_load(const char *path, aw_wav **out);
AW_API int aw_wav_load_memory(const void *bytes, size_t size, aw_wav **out);
AW_API int aw_wav_info(const aw_wav *wav, double *sample_rate, int *channels, int *frames);
AW_API const float *aw_wav_channel(const aw_wav *wav, int channel); /* audioData[channel], NULL if out of range */
AW_API void aw_wav_destroy(aw_wav *wav);

/* ---- InputLayout / HRIRChannelMap (VirtualSpeaker.swift:59-100, 224-297) ------------------ */
/* Writes the layout's speakers (aw_speaker) and returns their count (0 for an unknown layout). */
AW_API int aw_layout_speakers(int layout, int *speakers, int capacity);
/* hesuvi7Channel when wav_channels == 7, else hesuvi14Channel (HRIRManager.swift:355-360).
 * left_idx/right_idx[i] = -1 for a speaker without mapping. */
AW_API int aw_hesuvi_map(int wav_channels, const int *speakers, int n_speakers, int *left_idx, int *right_idx);
/* HRIRChannelMap.parseHeSuViFormat (VirtualSpeaker.swift:301-346): fills left/right[AW_SPK_COUNT] (-1 = unmapped). */
AW_API int aw_hesuvi_parse(const char *text, int *left_idx, int *right_idx);

/* ---- Resampler.resampleHighQuality (Resampler.swift:31-68), computed on `device` ---------- */
AW_API int aw_resample_output_count(int count, double from_rate, double to_rate);
AW_API int aw_resample(int device, const float *input, int count, double from_rate, double to_rate, float *output, int capacity, int *written);
/* Resampling modes.  AW_RESAMPLE_REFERENCE (default everywhere) reproduces Resampler.swift:31-68 literally: the vDSP_vgenp
 * breakpoints compress a 44.1 kHz response when the target is 48 kHz, the last sample is held, and down-sampling
 * (where the reference reads past its control vector) is refused with AW_ERR_RESAMPLE_DOWN.
 * AW_RESAMPLE_CORRECT is the conversion the reference's doc comment describes: out[n] = linear interpolation of the input at
 * source position n * from_rate / to_rate (float64 position and blend), same output count, last sample held; up- and
 * down-sampling allowed (no anti-alias filter: meant for impulse responses with little energy above the target Nyquist). */
#define AW_RESAMPLE_REFERENCE 0
#define AW_RESAMPLE_CORRECT 1
AW_API int aw_resample_ex(int device, const float *input, int count, double from_rate, double to_rate, int mode, float *output,
                          int capacity, int *written);

/* ---- HRIR filter bank: HRIRManager.activatePreset build loop (HRIRManager.swift:347-423) +
 *      ConvolutionEngine.init partitioning/FFT (ConvolutionEngine.swift:68-197), done once per
 *      (preset, rate, block) instead of once per engine.
 * pcm: planar [channels][frames].  Speaker i uses channels left_idx[i] / right_idx[i]; a speaker with
 * index -1 is skipped (no renderer), an index >= channels fails with AW_ERR_CHANNEL_MAPPING.
 * If |src_rate - dst_rate| > 0.01 the IRs are resampled with the reference's vgenp semantics. */
AW_API int aw_bank_create(int device, const float *pcm, int channels, int frames, double src_rate, double dst_rate,
                          const int *left_idx, const int *right_idx, int n_speakers, int block, aw_bank **out);
/* Same with an explicit resampling mode (AW_RESAMPLE_*). */
AW_API int aw_bank_create_ex(int device, const float *pcm, int channels, int frames, double src_rate, double dst_rate,
                             const int *left_idx, const int *right_idx, int n_speakers, int block, int resample_mode, aw_bank **out);
/* Convenience: load -> layout -> HeSuVi map -> bank, i.e. activatePreset(preset, targetSampleRate, inputLayout). */
AW_API int aw_bank_create_from_wav(int device, const aw_wav *wav, double dst_rate, int layout, int block, aw_bank **out);
AW_API int aw_bank_info(const aw_bank *bank, int *n_speakers, int *block, int *partitions, int *taps);
/* Distinct (left, right) channel pairs among the bank's speakers = frequency-domain delay lines the block kernel keeps per stream.
 * Speakers that share a pair — FC and LFE in both HeSuVi maps (VirtualSpeaker.swift:235-236, 281-283) — are summed before the
 * forward transform and filtered once (linearity); 7 for 7.1, 5 for 5.1, 2 for stereo. */
AW_API int aw_bank_rows(const aw_bank *bank);
/* Copies the bank to the host as [speaker][partition][bin]{L.re, L.im, R.re, R.im}; bin 0 = DC; nyquist[speaker][partition]{L,R}. */
AW_API int aw_bank_read(const aw_bank *bank, float *spectrum, float *nyquist);
AW_API void aw_bank_destroy(aw_bank *bank);

/* ---- Engine ------------------------------------------------------------------------------- */
/* RealtimeAudioProcessor.init + RendererState (RealtimeAudioProcessor.swift:30-62, HRIRManager.swift:123-131) */
AW_API int aw_engine_create(const aw_engine_config *config, aw_engine **out);
AW_API void aw_engine_destroy(aw_engine *engine);
/* Publishes a bank for streams [first, first+count) and clears their convolution state, like a fresh
 * RendererState (HRIRManager.swift:480-501).  bank == NULL = no renderers: passthrough (HRIRManager.swift:555-564). */
AW_API int aw_engine_set_bank(aw_engine *engine, int first, int count, const aw_bank *bank);
/* ---- EQ control for streams [first, first+count).  A definition is (preamp_db, filters, n_filters);
 *      n_filters < 0 means `definition == nil`.  On AW_ERR_EQ_INVALID_FILTER, *bad_index is the index among the
 *      ENABLED filters (EqualizerRuntimeEffect.map looks up its sourceLine) and *bad_reason an aw_biquad_error;
 *      the target then falls back to unity exactly as the reference does (`try? setTarget(nil)`). */
/* AudioEffectGraph.prepare(for:equalizerDefinition:) -> EqualizerRuntimeEffect.prepare
 * (AudioEffectGraph.swift:94-138, EqualizerRuntimeEffect.swift:10-34): creates the processor on first use,
 * setTarget + drainRetiredStates, equalizer-active flag = (definition != nil), false on error. */
AW_API int aw_engine_eq_prepare(aw_engine *engine, int first, int count, double preamp_db, const aw_eq_filter *filters,
                                int n_filters, int *bad_index, int *bad_reason);
/* AudioEffectGraph.updateEqualizer -> EqualizerRuntimeEffect.setTarget (AudioEffectGraph.swift:140-176,
 * EqualizerRuntimeEffect.swift:36-48): AW_ERR_NOT_READY before the first prepare; the flag becomes true either way. */
AW_API int aw_engine_eq_update(aw_engine *engine, int first, int count, double preamp_db, const aw_eq_filter *filters,
                               int n_filters, int *bad_index, int *bad_reason);
/* ParametricEqualizerProcessor.setTarget alone (ParametricEqualizerProcessor.swift:236-238), optionally followed by
 * drainRetiredStates; does not touch the equalizer-active flag.  Creates the processor if needed. */
AW_API int aw_engine_eq_set_target(aw_engine *engine, int first, int count, double preamp_db, const aw_eq_filter *filters,
                                   int n_filters, int drain_retired, int *bad_index, int *bad_reason);
/* Installs a freshly prepared ParametricEqualizerState (zero history) as the ACTIVE state with no crossfade and cancels any
 * transition: the equivalent of calling ParametricEqualizerState.process directly (ParametricEqualizerProcessor.swift:16-98),
 * which the reference's state-level tests do.  Sets the equalizer-active flag. */
AW_API int aw_engine_eq_install_state(aw_engine *engine, int first, int count, double preamp_db, const aw_eq_filter *filters,
                                      int n_filters, int *bad_index, int *bad_reason);
/* ParametricEqualizerProcessor.drainRetiredStates (:247-251) */
AW_API int aw_engine_eq_drain_retired(aw_engine *engine, int first, int count);
/* Forces AudioEffectGraph's equalizer-active flag (tests that drive the processor directly). */
AW_API int aw_engine_eq_active(aw_engine *engine, int first, int count, int active);
/* Models a contended publication lock (withPublicationLockForTesting, :229-234): while held, process keeps the prior target. */
AW_API int aw_engine_eq_hold_publication(aw_engine *engine, int first, int count, int held);
/* RealtimeAudioProcessor.process / AudioEffectGraph.process for every stream
 * (RealtimeAudioProcessor.swift:77-119, AudioEffectGraph.swift:179-246).
 * in : planar [stream][n_speakers][frames]   out: planar [stream][2][frames]   (HOST pointers).
 * Any 0 < frames <= max_frames_per_call; output is delayed by the adapter latency of the reference. */
AW_API int aw_engine_process(aw_engine *engine, const float *in, float *out, int frames);
/* Same with DEVICE pointers and explicit strides (in elements): in[stream*in_stream_stride + ch*in_channel_stride + i]. */
AW_API int aw_engine_process_device(aw_engine *engine, const float *in, long long in_stream_stride, long long in_channel_stride,
                                    float *out, long long out_stream_stride, long long out_channel_stride, int frames);
/* Pipelined host path for offline batch rendering: copies and kernels of consecutive submits overlap;
 * `out` is valid after aw_engine_wait().  in/out must stay alive (and should be pinned) until then. */
AW_API int aw_engine_submit(aw_engine *engine, const float *in, float *out, int frames);
/* Same pipeline with the input already resident on the DEVICE (strides in elements): offline batch rendering whose input is
 * produced or decoded on the GPU — BASELINE.json configs[4] (16,384 streams x 60 s: 188.7 GB of input per GPU would not cross
 * the host link).  `in` may be rewritten once work submitted later on aw_engine_stream() runs; `out`
 * (HOST) is valid after aw_engine_wait().  No reference counterpart: the reference renders one stream on the CPU. */
AW_API int aw_engine_submit_device(aw_engine *engine, const float *in, long long in_stream_stride, long long in_channel_stride,
                                   float *out, int frames);
AW_API int aw_engine_wait(aw_engine *engine);
/* Orders aw_engine_stream() behind everything the engine still has in flight on internal streams (AW_ENGINE_OVERLAP_EQ);
 * does not block the host. */
AW_API int aw_engine_flush(aw_engine *engine);
/* Single-stream mirror of StereoAudioProcessing.process (AudioPipeline.swift:3-11) for an engine with n_streams == 1
 * and n_speakers <= 2: input_right may be NULL (mono duplicated), output_left may alias output_right. HOST pointers. */
AW_API int aw_engine_process_stereo(aw_engine *engine, const float *input_left, const float *input_right, float *output_left,
                                    float *output_right, int frames);
/* what: AW_RESET_SPATIAL = RealtimeAudioProcessor.reset (RealtimeAudioProcessor.swift:121-127),
 *       AW_RESET_EQ = ParametricEqualizerProcessor.reset (:240-244, applied at the next process call). */
#define AW_RESET_SPATIAL 1
#define AW_RESET_EQ 2
AW_API int aw_engine_reset(aw_engine *engine, int first, int count, int what);
/* Counters since creation: kernels launched by this engine, blocks rendered, bytes copied H2D / D2H. */
AW_API int aw_engine_counters(const aw_engine *engine, unsigned long long *kernel_launches, unsigned long long *blocks,
                              unsigned long long *h2d_bytes, unsigned long long *d2h_bytes);
/* Execution plan chosen for the engine: fused_tile = streams per tile of the block kernel that does K2+K3+K4 in one launch
 * (the persistent kernel KP, or the single-wave KF when AW_PERSISTENT=0; 0 = split kernels K2, K3, K4), mac_tile = streams per
 * thread of the stand-alone K3, partitions_cap = FDL slots per (stream, speaker). */
AW_API int aw_engine_plan(const aw_engine *engine, int *fused_tile, int *mac_tile, int *partitions_cap);
/* Per-kernel device timing for benchmarks: between begin and end, CUDA events bracket the per-block kernels of up to `max_blocks`
 * blocks on the engine's stream.  end() synchronises and returns the summed milliseconds and launch counts of 4 slots: [0..2] the
 * kernels aw_engine_kernels() names, in that order (one-kernel plans use slot 0 only), [3] the equalizer launches of a call.
 * Profiling serialises work that otherwise overlaps (side streams).  Not for the real-time path. */
AW_API int aw_engine_profile_begin(aw_engine *engine, int max_blocks);
AW_API int aw_engine_profile_end(aw_engine *engine, double *kernel_ms, unsigned long long *kernel_launches);
/* Semicolon-separated names of the kernels the engine launches per block, in launch order (e.g. "k_persistent<8,4>"). */
AW_API int aw_engine_kernels(const aw_engine *engine, char *names, int capacity);
/* 1 when the block kernel stages the FDL with tensor-map TMA copies (cp.async.bulk.tensor), 0 when it uses 1-D bulk copies
 * (driver without cuTensorMapEncodeTiled, or AW_KP_TENSOR_TMA=0).  Diagnostics only. */
AW_API int aw_engine_uses_tensor_maps(const aw_engine *engine);
/* Raw CUDA stream (cudaStream_t) the engine launches on, for event timing by the caller. */
AW_API void *aw_engine_stream(const aw_engine *engine);
/* Pinned host memory helpers for callers that cannot call cudaHostAlloc themselves. */
AW_API void *aw_host_alloc(size_t bytes);
AW_API void aw_host_free(void *ptr);

/* ---- EQ building blocks (host, float64) --------------------------------------------------- */
/* BiquadCoefficientBuilder.make (BiquadCoefficientBuilder.swift:30-107): out5 = {b0,b1,b2,a1,a2}; returns aw_biquad_error. */
AW_API int aw_biquad_make(int type, double gain_db, double frequency_hz, double q, double sample_rate, double *out5);
/* EqualizerAPOParser.parse (EqualizerAPOParser.swift:36-151).  Returns AW_OK or AW_ERR_EQ_PARSE;
 * on error `issues` receives "line N: reason; ..." (the reference's errorDescription without the filename prefix). */
AW_API int aw_eq_parse(const void *bytes, size_t size, double *preamp_db, aw_eq_filter *filters, int capacity, int *n_filters,
                       char *issues, size_t issues_capacity);

/* ---- Deterministic synthetic input (bench / tests): uniform [-0.25, 0.25], keyed by (seed, stream, speaker, frame) */
AW_API int aw_synth_fill_device(int device, float *d_out, int first_stream, int n_streams, int n_speakers, long long frame0,
                                int frames, uint32_t seed, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* AIRWAVE_CUDA_H */
