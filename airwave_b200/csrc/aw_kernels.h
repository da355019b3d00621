// aw_kernels.h — host-callable launchers of the sm_100a kernels (internal to libairwave_cuda.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace aw {

// Geometry shared by the per-block kernels of one engine segment (streams that share a bank).
struct BlockGeom {
    int first_stream;   // first stream of the segment
    int n_streams;      // streams in the segment
    int S;              // FDL rows of the segment's bank (distinct renderers: speakers that share a filter pair share a row)
    int Se;             // speakers per stream the engine's state arrays are laid out for (S <= Se)
    int B;              // block size = complex bins per spectrum
    int log2m;          // log2(B)
    int P;              // partitions of the segment's bank (ring modulus, ConvolutionEngine.swift:256-259)
    int Pm;             // ring modulus when it differs from P (KP keeps one spare slot, see KpSegment); 0 = P
    int P_cap;          // FDL slots allocated per (stream, speaker)
    int head;           // fdlIndex after the decrement for this block
    int prev_is_rows;   // 1: the `prev` operand of the forward kernels is the engine's overlap buffer (one summed block per FDL row)
    const struct KpRowTable *rows;   // which input channels an FDL row sums (KpRowTable); nullptr: row s = input channel s
};

struct StridedIn {      // planar input: ptr[stream*ss + channel*cs + i]
    const float *ptr;
    long long ss, cs;
};

struct StridedOut {     // planar output; when ring_cap > 0 sample i lands at (ring_start + i) % ring_cap
    float *ptr;
    long long ss, cs;
    int ring_cap, ring_start;
};

// K2: [prev | cur] -> real FFT -> FDL slot `head` (+ Nyquist side array); optionally saves cur as next overlap.
cudaError_t launch_input_rfft(const BlockGeom &g, StridedIn cur, StridedIn prev, float *overlap_save, float2 *fdl,
                              float *fdl_ny, const float2 *tw, cudaStream_t st);
// Filter bank layout: row (s, p) = two planes of B/2 float4 {L.re, L.im, R.re, R.im}: even bins, then odd bins.
// K3: acc[stream][ear][bin] = sum_{s,p} FDL[stream][s][(head+p)%P][bin] * bank[s][p][bin].ear
cudaError_t launch_fdl_cmac(const BlockGeom &g, const float2 *fdl, const float4 *bank, float2 *acc, int tile, cudaStream_t st);
// K4: Nyquist reduction + inverse real FFT + overlap-save discard -> out (planar or FIFO ring)
cudaError_t launch_irfft_out(const BlockGeom &g, const float2 *acc, const float *fdl_ny, const float *bank_ny,
                             StridedOut out, const float2 *tw, cudaStream_t st);
// K1: partition + zero-pad + real FFT of every (speaker, ear) impulse response -> filter bank
cudaError_t launch_bank_build(const float *ir, int S, int taps, int B, int log2m, int P, float4 *bank, float *bank_ny,
                              const float2 *tw, cudaStream_t st);
// K6: Resampler.resampleHighQuality (vDSP_vramp + vDSP_vgenp semantics), rows x count -> rows x out_count
cudaError_t launch_resample_vgenp(const float *in, int rows, int count, float step, float *out, int out_count, cudaStream_t st);
// K6b: time-correct linear interpolation (AW_RESAMPLE_CORRECT): out[n] = lerp(in, n * step), float64 position and blend
cudaError_t launch_resample_linear(const float *in, int rows, int count, double step, float *out, int out_count, cudaStream_t st);
// K7: frame adapter pieces (RealtimeAudioProcessor.swift:88-116, 166-171, 174-190)
cudaError_t launch_gather_pending(StridedIn in, int in_offset, int copy_count, float *pending, int pending_count, int n_streams,
                                  int S, int B, int dup_mono, cudaStream_t st);
cudaError_t launch_drain_fifo(const float *fifo, int fifo_cap, int fifo_read, int fifo_count, StridedOut out, int out_offset,
                              int frames, int n_streams, cudaStream_t st);
// passthrough (HRIRManager.swift:555-564): L = channel 0, R = channel 1 (or channel 0 when mono)
cudaError_t launch_passthrough(StridedIn in, StridedOut out, int first_stream, int n_streams, int S, int frames, cudaStream_t st);
cudaError_t launch_synth_fill(float *out, int first_stream, int n_streams, int S, long long frame0, int frames, uint32_t seed,
                              cudaStream_t st);

// K5: stereo float64 TDF-II biquad cascade (ParametricEqualizerState.process, ParametricEqualizerProcessor.swift:58-91)
// with the 20 ms crossfade of ParametricEqualizerProcessor.process (:254-314).
struct EqProgram {          // one ParametricEqualizerState's immutable part, resident in HBM
    double preamp_linear;
    int n_filters;
    int pad;
    double coef[64][5];     // b0 b1 b2 a1 a2
};
struct EqLaunch {
    int first_stream, n_streams;
    const EqProgram *from;  // program producing the "old" signal (the active state when no transition)
    const EqProgram *to;    // transition target or nullptr
    int from_voice, to_voice;   // which z-state voice each program uses
    int seg_start, seg_len; // frames [seg_start, seg_start+seg_len) of this call
    int transition_frame;   // transitionFrame at seg_start (only when to != nullptr)
    int transition_length;
};
// z: [stream][voice(2)][ear(2)][filter(64)][2] doubles; io: planar stereo in place.
cudaError_t launch_eq(const EqLaunch &l, int max_filters, double *z, StridedOut io, cudaStream_t st);
// Steady-state cascades (no crossfade, 1..32 filters) of up to kEqMaxSegments stream ranges in ONE launch: per-device
// profiles give every range its own equalizer (DeviceProfileManager.swift:4-12).
struct EqSegment {
    int first_stream, n_streams;
    int n_filters, voice;
    int warp0;              // filled in by the launcher: first warp of this range
    int pad;
    const EqProgram *prog;
};
constexpr int kEqMaxSegments = 64;
cudaError_t launch_eq_steady(const EqSegment *segs, int n_segs, int seg_start, int seg_len, double *z, StridedOut io, cudaStream_t st);
cudaError_t launch_eq_reset(double *z, int first_stream, int n_streams, int voice_mask, cudaStream_t st);

// KF: K2 + K3 + K4 fused for a tile of `tile` (1, 2 or 4) streams per CTA; 64 <= B <= 512.
bool fused_supported(int log2m);
int fused_blocks_per_sm(int log2m, int tile);   // resident CTAs per SM of that variant (occupancy API)
cudaError_t launch_fused(const BlockGeom &g, StridedIn cur, StridedIn prev, float *overlap_save, float2 *fdl, float *fdl_ny,
                         const float4 *bank, const float *bank_ny, StridedOut out, const float2 *tw, int tile, cudaStream_t st);

// KP: persistent warp-specialised block kernel (aw_persistent.cu): one CTA per SM, TMA bulk-copy ring, FFT warps one tile
// ahead of the MAC warps; 64 <= B <= 2048.  `tile` = streams per tile (4 or 2).
struct EqProgram;
struct EqFuse {             // equalizer fused into KP's epilogue (steady state only); n_filters == 0: none
    const EqProgram *prog;
    double *z;              // [stream][voice(2)][ear(2)][filter(64)][2]
    int voice;
    int n_filters;          // 1..32
};
bool persistent_can_fuse_eq(int log2m, int tile, int n_filters);
int persistent_tiles(int log2m);                // bit mask of the tiles available for that transform size (0 = unsupported)
// FDL rows of a bank.  Speakers that share one (left, right) impulse-response pair — FC and LFE in both HeSuVi maps
// (VirtualSpeaker.swift:235-236, 281-283) — share ONE row: convolution is linear, so their input channels are
// added before the forward transform and filtered once (conv(a, h) + conv(b, h) = conv(a + b, h), rounding aside).  A row names
// the input channels it sums (up to 4, -1 = none) and the bank speaker whose filter rows it uses.
constexpr int kKpMaxRows = 64, kKpMaxRowSources = 4;
struct KpRowTable {
    signed char src[kKpMaxRows][kKpMaxRowSources];
    signed char spk[kKpMaxRows];
};
struct KpSegment {          // a stream range bound to one bank (tile0 and n_big are filled in by the launcher)
    int first_stream, n_streams;
    int S, P;               // FDL rows (distinct renderers, see KpRowTable), partitions
    int Pm;                 // ring modulus of the range's FDL rows: P + 1.  The reference's modulus is partitionCount
                            // (ConvolutionEngine.swift:256-259); the spare slot lets the forward transform of block b+1 be
                            // written while block b's multiply-accumulate still reads its P slots (one launch walks the k
                            // blocks of a call).  Which slot holds which partition is internal state: outputs do not change.
    int head;               // fdlIndex of the call's FIRST block, after its decrement; block b uses (head - b) mod Pm
    int tile0, n_big;       // first tile of the range; its first n_big tiles hold T streams, the rest `small` streams
    const float4 *bank;
    const float *bank_ny;
    const KpRowTable *rows; // device memory (part of the bank)
};
constexpr int kKpMaxSegments = 64;   // per launch (64 x 56 B of the 4 KB kernel parameter space)
struct KpCall {
    int nb;                 // blocks of this call (frames = nb * B); input block b at cur + b*B, output block b at out + b*B
    int order;              // walk order of the (tile, block) items of a CTA: 0 = block-major (all tiles of block 0, then block
                            // 1, ...), 1 = tile-major (all blocks of a tile back to back: its FDL rows are re-read from L2)
    int keep_pct;           // tile-major: percentage of the history rows loaded with L2 evict_last in all but the last block
    int debug;              // timing experiments; ignored unless the library is built with -DAW_TIMING_EXPERIMENTS
    int prev_is_rows;       // 1: `prev` is the engine's overlap buffer (one summed block per FDL row); 0: it is laid out like the input
};
// Tensor maps of the engine's FDL for KP's stage loads: 128-byte CUtensorMap objects in global memory, indexed
// (rows - 1) * 3 + {0: 4 streams, 1: 2 streams, 2: 1 stream} for rows = 1..persistent_stage_rows(log2m); nullptr = bulk copies.
// Tensor: {bin within a 256-bin chunk, chunk, stream, ring slot, speaker}; box: {min(B,256), chunks of a stage column, streams,
// rows, 1} -> shared memory [row][stream][bins of the stage column].
int persistent_stage_rows(int log2m);           // RS: consecutive partitions of one speaker per stage
int persistent_stage_bins(int log2m);           // bins per stage column (2C)
cudaError_t launch_persistent(const KpSegment *segs, int n_segs, int Se, int P_cap, int log2m, StridedIn cur, StridedIn prev,
                              float *overlap_save, float2 *fdl, float *fdl_ny, const void *tmaps, StridedOut out, const float2 *tw,
                              int tile, int max_ctas, const KpCall &call, const EqFuse &eq, cudaStream_t st);

// Plan layout in global memory: M half-circle twiddles exp(-2*pi*i*k/(2M)) followed by plan_pt_entries(log2m) per-pass
// twiddles (RegFft::pt_entry layout), M = 2^log2m complex points.
int plan_pt_entries(int log2m);
cudaError_t launch_build_pt(int log2m, const float2 *hc, float2 *pt, cudaStream_t st);
size_t fft_smem_bytes(int log2m);
cudaError_t configure_kernels(int log2m);   // opt in to > 48 KB dynamic shared memory for that transform size

}  // namespace aw
