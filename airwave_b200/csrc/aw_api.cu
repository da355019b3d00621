// aw_api.cu — C ABI of libairwave_cuda.so: plan cache, HRIR filter banks, the batched engine
// (frame adapter + convolution + EQ state machines) and the host<->device staging.
//
// Host-side control mirrors, per stream range, the objects the reference keeps per process:
//   Segment   = RendererState (HRIRManager.swift:123-131): streams sharing one bank + FDL head
//   EqMachine = EqualizerRuntimeEffect + ParametricEqualizerProcessor render-thread state
//               (EqualizerRuntimeEffect.swift:5-78, ParametricEqualizerProcessor.swift:121-408)
// All streams of an engine advance in lock-step, so adapter counters are per engine.
// aw_engine_process* allocates nothing: every buffer, stream, event and table is created up front.
#include <cuda.h>   // CUtensorMap and the cuTensorMapEncodeTiled prototype only: the entry point is fetched at run time, libcuda is not linked
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/airwave_cuda.h"
#include "aw_internal.h"
#include "aw_kernels.h"

namespace aw {
static thread_local std::string g_last_error;
int set_error(int status, const std::string &message)
{
    g_last_error = message;
    return status;
}
}  // namespace aw

using namespace aw;

#define AW_CUDA(call)                                                                                     \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return set_error(AW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// ---- plan cache (FFTSetupManager.swift:41-69) ----------------------------------------------------
std::mutex g_plan_mutex;
std::map<std::pair<int, int>, float2 *> g_plans;   // (device, log2n) -> twiddles exp(-2*pi*i*k/N), k < N/2

int get_plan(int device, int log2n, const float2 **out)
{
    if (log2n < 3 || log2n > 14) return set_error(AW_ERR_INVALID_ARGUMENT, "plan: log2n out of range [3, 14]");
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    auto it = g_plans.find({device, log2n});
    if (it != g_plans.end()) { if (out) *out = it->second; return AW_OK; }
    DeviceGuard guard(device);
    if (!guard.ok) return set_error(AW_ERR_CUDA, "plan: cudaSetDevice failed");
    const int N = 1 << log2n, half = N / 2;
    std::vector<float2> tw(half);
    for (int k = 0; k < half; ++k) {
        const double a = -2.0 * M_PI * (double)k / (double)N;
        tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    float2 *d = nullptr;
    AW_CUDA(cudaMalloc(&d, sizeof(float2) * (half + plan_pt_entries(log2n - 1))));
    cudaError_t e = cudaMemcpy(d, tw.data(), sizeof(float2) * half, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_build_pt(log2n - 1, d, d + half, 0);   // per-pass tables derived from the same values
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(d); return set_error(AW_ERR_CUDA, std::string("plan upload: ") + cudaGetErrorString(e)); }
    e = configure_kernels(log2n - 1);
    if (e != cudaSuccess) { cudaFree(d); return set_error(AW_ERR_CUDA, std::string("configure_kernels: ") + cudaGetErrorString(e)); }
    g_plans[{device, log2n}] = d;
    if (out) *out = d;
    return AW_OK;
}

}  // namespace

// ---- opaque types ----------------------------------------------------------------------------------
struct aw_bank {
    int device = 0, S = 0, B = 0, log2m = 0, P = 0, taps = 0;
    float4 *d_bank = nullptr;   // [S][P][2 planes: even bins, odd bins][B/2] {L.re, L.im, R.re, R.im}
    float *d_ny = nullptr;      // [S][P][2]
    // FDL rows for the block kernel: speakers with the same (left, right) channel pair share one (KpRowTable); [1] = one row per
    // speaker (literal-stereo engines, AW_KP_MERGE_ROWS=0)
    int R = 0;
    KpRowTable *d_rows = nullptr;   // [2]
    float4 *d_bank_rows = nullptr;  // filter rows of the R distinct pairs, [R][P][2 planes][B/2] (== d_bank when R == S)
    float *d_ny_rows = nullptr;     // [R][P][2]
};

namespace {

struct Segment {
    int first, count;
    const aw_bank *bank;   // nullptr = passthrough
    int head;              // fdlIndex (ConvolutionEngine.swift:39)
};

constexpr size_t kZeroCopyBytes = 2u << 20;   // engines whose staging fits in this use mapped host buffers (no copies)
constexpr int kEqPool = 1024;   // ParametricEqualizerState objects alive per engine at creation (slot 0 = shared unity); the pool
                                // doubles on demand in the control path (eq_pool_grow), never in aw_engine_process*

struct EqMachine {
    int first = 0, count = 0;
    bool hasProcessor = false;     // EqualizerRuntimeEffect.controlProcessor != nil
    bool eqActive = false;         // AudioEffectGraph.equalizerActiveLock state
    bool lockHeld = false;         // publication lock contended (testing)
    bool resetRequested = false;
    int published = -1, audioThreadTarget = -1, activeState = 0, transitionFrom = -1, transitionTo = -1;
    int pendingTarget = -1, observedTarget = -1, pendingRetirement = -1, retired = -1;
    int transitionFrame = 0;
    int activeVoice = 0;           // z-state voice holding activeState's history
};

struct Staging {
    float *d_in = nullptr, *d_out = nullptr;
    cudaEvent_t in_ready = nullptr, compute_done = nullptr, out_done = nullptr;
    bool busy = false;
};

}  // namespace

struct aw_engine {
    aw_engine_config cfg{};
    int n = 0, S = 0, B = 0, log2m = 0, maxFrames = 0, P_cap = 0, fifoCap = 0;
    int macTile = 0;
    int fusedTile = 0;             // 0 = split path (K2, K3, K4), else streams per CTA of the fused kernel
    int numSMs = 148;
    bool persistent = false;       // use the persistent warp-specialised kernel KP (aw_persistent.cu) instead of KF
    int persistentTile = 0;        // streams per tile of KP (4 or 2)
    int persistentCtas = 0;        // CTAs of KP (default: one per SM)
    int persistentDebug = 0;       // timing experiments only (AW_PERSISTENT_DEBUG; ignored unless built with -DAW_TIMING_EXPERIMENTS)
    int ringExtra = 0;             // spare FDL ring slots per (stream, speaker): 1 with KP (see KpSegment::Pm), else 0
    int kpOrder = 1;               // walk order of a multi-block call's (tile, block) items: 1 = tile-major, 0 = block-major
    int kpKeepPct = 0;             // tile-major: share of a tile's history rows loaded evict_last in all but the last block (AW_KP_KEEP;
                                   // measured: 0..25 % is best at C2 — plain LRU keeps what fits, pinning more evicts the filter bank)
    bool kpMultiBlock = true;      // one launch per call (AW_KP_MULTIBLOCK=0: one launch per block)
    bool kpMergeRows = true;       // speakers that share a filter pair share an FDL row (AW_KP_MERGE_ROWS=0: one row per speaker)
    bool eqFusion = false;         // AW_EQ_FUSION=1: steady-state EQ rides in KP's epilogue.  Off by default: the bit-exact float64
                                   // recurrence needs ~220 cycles per sample, 29 us per 256-frame block on the few FFT warps of a
                                   // CTA — longer than a tile lasts — whereas the separate K5 pass hides it behind 37 warps per SM
                                   // (C4: 0.74 ms fused vs 0.54 ms separate)
    const float2 *d_tw = nullptr;
    float2 *d_fdl = nullptr;
    float *d_fdl_ny = nullptr, *d_overlap = nullptr, *d_pending = nullptr, *d_fifo = nullptr;
    float2 *d_acc = nullptr;
    void *d_tmaps = nullptr;       // CUtensorMap[RS][3] over d_fdl for KP's stage loads (aw_kernels.h); nullptr = 1-D bulk copies
    double *d_eq_z = nullptr;
    EqProgram *d_eq_prog = nullptr;
    Staging stage[2];
    int nextStage = 0;
    // Small engines (the reference's own case: one stereo stream per callback, AudioPipeline.swift:3-11): the synchronous host entry
    // points stage the caller's buffers in page-locked memory the kernels read and write DIRECTLY over PCIe (zero-copy) — one
    // launch and one stream synchronisation per call instead of four driver copies around it.
    float *h_in = nullptr, *h_out = nullptr;
    cudaStream_t stream = nullptr, h2d = nullptr, d2h = nullptr;
    // engines whose stream ranges are bound to different banks (per-device profiles, DeviceProfileManager.swift:4-12) launch one
    // grid per range; small ranges leave most SMs idle, so their grids are spread over side streams and run concurrently
    static constexpr int kSideStreams = 8;
    cudaStream_t side[kSideStreams] = {};
    cudaEvent_t forkEvent = nullptr, joinEvent[kSideStreams] = {};
    cudaStream_t eqStream = nullptr;   // where the equalizer of the machine being processed launches (the main stream, or a side stream)
    // AW_ENGINE_OVERLAP_EQ: every equalizer launch goes to eqSide, ordered behind the call's convolution by evKp; the main stream
    // picks a call's equalizer up (evEq) one call later, so the float64 cascade of call j runs next to the convolution of call j+1
    cudaStream_t eqSide = nullptr;
    cudaEvent_t evKp = nullptr, evEq[2] = {nullptr, nullptr};
    bool eqPending[2] = {false, false};
    int eqParity = 0;
    bool eqRanThisCall = false;
    std::vector<Segment> segments;
    std::vector<EqMachine> machines;
    // EQ state-object pool (host mirror of the programs resident on the device)
    std::vector<int> eqRef;          // refcounts; 0 = free
    std::vector<int> eqFilters;      // filter count per slot
    std::vector<double> eqPreamp;    // preampLinear per slot
    int transitionLength = 1;
    // adapter counters (RealtimeAudioProcessor.swift:26-28)
    int pendingCount = 0, fifoReadIndex = 0, fifoCount = 0;
    unsigned long long launches = 0, blocks = 0, h2dBytes = 0, d2hBytes = 0;
    // The host-side state of a call (FDL heads, adapter counters, equalizer state machines) advances while its launches are
    // enqueued.  If one of them fails, host and device no longer agree: the engine refuses to render until it has been reset.
    bool poisoned = false;
    // optional per-kernel event timing (benchmarks only)
    std::vector<cudaEvent_t> profEvents, profEqEvents;
    size_t profUsed = 0, profEqUsed = 0;
    bool profOn = false;
};

namespace {

#define AW_LAUNCH(e, call)                                                                                 \
    do {                                                                                                   \
        cudaError_t le_ = (call);                                                                          \
        ++(e)->launches;                                                                                   \
        if (le_ != cudaSuccess) return set_error(AW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(le_)); \
    } while (0)

int check_range(const aw_engine *e, int first, int count)
{
    if (!e) return set_error(AW_ERR_INVALID_ARGUMENT, "null engine");
    if (first < 0 || count <= 0 || first + count > e->n) return set_error(AW_ERR_RANGE, "stream range outside [0, n_streams)");
    return AW_OK;
}

// Splits the segment list so that [first, first+count) is covered by whole segments.
void split_segments(aw_engine *e, int at)
{
    for (size_t i = 0; i < e->segments.size(); ++i) {
        Segment &s = e->segments[i];
        if (at > s.first && at < s.first + s.count) {
            Segment tail = s;
            tail.first = at;
            tail.count = s.first + s.count - at;
            s.count = at - s.first;
            e->segments.insert(e->segments.begin() + i + 1, tail);
            return;
        }
    }
}

void split_machines(aw_engine *e, int at);
int eq_pool_grow(aw_engine *e);

// ---- EQ state pool ---------------------------------------------------------------------------------
void eq_retain(aw_engine *e, int slot) { if (slot > 0) ++e->eqRef[slot]; }
void eq_release(aw_engine *e, int slot) { if (slot > 0 && e->eqRef[slot] > 0) --e->eqRef[slot]; }
void eq_assign(aw_engine *e, int &dst, int slot)
{
    eq_retain(e, slot);
    eq_release(e, dst);
    dst = slot;
}

void split_machines(aw_engine *e, int at)
{
    for (size_t i = 0; i < e->machines.size(); ++i) {
        EqMachine &m = e->machines[i];
        if (at > m.first && at < m.first + m.count) {
            EqMachine tail = m;
            tail.first = at;
            tail.count = m.first + m.count - at;
            m.count = at - m.first;
            int *refs[] = {&tail.published, &tail.audioThreadTarget, &tail.activeState, &tail.transitionFrom, &tail.transitionTo,
                           &tail.pendingTarget, &tail.observedTarget, &tail.pendingRetirement, &tail.retired};
            for (int *r : refs) eq_retain(e, *r);
            e->machines.insert(e->machines.begin() + i + 1, tail);
            return;
        }
    }
}

// Doubles the device array of EqProgram slots (control path only; the render path never allocates).  Bounded by the number of
// state objects the engine's streams can reference at once (9 references per range, at most n_streams ranges).
int eq_pool_grow(aw_engine *e)
{
    const size_t cap = e->eqRef.size(), limit = (size_t)e->n * 9 + kEqPool;
    if (cap >= limit) return set_error(AW_ERR_OUT_OF_MEMORY, "equalizer state pool exhausted");
    const size_t grown = std::min(limit, cap * 2);
    EqProgram *d = nullptr;
    cudaError_t ce = cudaMalloc(&d, sizeof(EqProgram) * grown);
    if (ce != cudaSuccess) { cudaGetLastError(); return set_error(AW_ERR_OUT_OF_MEMORY, "equalizer state pool: cudaMalloc failed"); }
    // nothing may still be reading the old array: the engine's streams are idle after this
    ce = cudaStreamSynchronize(e->stream);
    for (int k = 0; k < aw_engine::kSideStreams && ce == cudaSuccess; ++k) ce = cudaStreamSynchronize(e->side[k]);
    if (ce == cudaSuccess && e->eqSide) ce = cudaStreamSynchronize(e->eqSide);
    if (ce == cudaSuccess) ce = cudaMemcpy(d, e->d_eq_prog, sizeof(EqProgram) * cap, cudaMemcpyDeviceToDevice);
    if (ce == cudaSuccess) ce = cudaMemset(d + cap, 0, sizeof(EqProgram) * (grown - cap));
    if (ce != cudaSuccess) { cudaFree(d); return set_error(AW_ERR_CUDA, std::string("equalizer state pool: ") + cudaGetErrorString(ce)); }
    cudaFree(e->d_eq_prog);
    e->d_eq_prog = d;
    e->eqRef.resize(grown, 0);
    e->eqFilters.resize(grown, 0);
    e->eqPreamp.resize(grown, 1.0);
    return AW_OK;
}

// ParametricEqualizerProcessor.prepare (ParametricEqualizerProcessor.swift:174-217): validates, designs the biquads and
// uploads one immutable program.  Returns the pool slot (refcount 1) in *slot.
int eq_prepare_state(aw_engine *e, double preampDB, const aw_eq_filter *filters, int n_filters, int *slot, int *bad_index,
                     int *bad_reason)
{
    if (bad_index) *bad_index = -1;
    if (bad_reason) *bad_reason = 0;
    const double sampleRate = e->cfg.sample_rate;
    if (!(std::isfinite(sampleRate) && sampleRate > 0)) return set_error(AW_ERR_EQ_INVALID_SAMPLE_RATE, "Sample rate must be finite and positive.");
    if (n_filters < 0) { preampDB = 0; n_filters = 0; }   // definition == nil (:182,191)
    if (!std::isfinite(preampDB)) return set_error(AW_ERR_EQ_NON_FINITE_PREAMP, "Preamp must produce a finite linear gain.");
    const double preampLinear = std::pow(10.0, preampDB / 20.0);
    if (!std::isfinite(preampLinear)) return set_error(AW_ERR_EQ_NON_FINITE_PREAMP, "Preamp must produce a finite linear gain.");
    int enabled = 0;
    for (int i = 0; i < n_filters; ++i) if (filters[i].enabled) ++enabled;
    if (enabled > 64) {
        if (bad_index) *bad_index = enabled;
        return set_error(AW_ERR_EQ_TOO_MANY_FILTERS, "Equalizer supports at most 64 filters; received " + std::to_string(enabled) + ".");
    }
    EqProgram prog;
    memset(&prog, 0, sizeof(prog));
    prog.preamp_linear = preampLinear;
    int k = 0;
    for (int i = 0; i < n_filters; ++i) {
        if (!filters[i].enabled) continue;
        const int rc = aw_biquad_make(filters[i].type, filters[i].gain_db, filters[i].frequency_hz, filters[i].q, sampleRate, prog.coef[k]);
        if (rc != AW_BIQUAD_OK) {
            if (bad_index) *bad_index = k;
            if (bad_reason) *bad_reason = rc;
            static const char *reasons[] = {"", "Sample rate must be finite and positive.", "Frequency must be finite, positive, and below Nyquist.",
                                            "Q must be finite and positive.", "Filter parameters must be finite.", "Filter coefficients must be finite."};
            return set_error(AW_ERR_EQ_INVALID_FILTER, "Filter " + std::to_string(k + 1) + " is invalid: " + reasons[rc]);
        }
        ++k;
    }
    prog.n_filters = k;
    int s = -1;
    for (int i = 1; i < (int)e->eqRef.size(); ++i) if (e->eqRef[i] == 0) { s = i; break; }
    if (s < 0) {
        // every stream range may hold several live state objects (active, transition target, pending, retired): grow
        s = (int)e->eqRef.size();
        const int rc = eq_pool_grow(e);
        if (rc != AW_OK) return rc;
    }
    AW_CUDA(cudaMemcpyAsync(e->d_eq_prog + s, &prog, sizeof(prog), cudaMemcpyHostToDevice, e->stream));
    AW_CUDA(cudaStreamSynchronize(e->stream));
    e->eqRef[s] = 1;
    e->eqFilters[s] = k;
    e->eqPreamp[s] = preampLinear;
    *slot = s;
    return AW_OK;
}

// BEGIN REALTIME PATH — everything from here to the END marker runs inside aw_engine_process*: no allocation, no container growth,
// no locks, no logging, no device-wide synchronisation (the analogue of scripts/check-audio-safety-invariants.sh:23-41 is
// tests/test_abi.py::test_realtime_path_gate, which greps this region).  Error paths (set_error) may build a message.
// ---- render-thread half of ParametricEqualizerProcessor (:317-407), one instance per EqMachine --------
int eq_begin_transition(aw_engine *e, EqMachine &m, int target)   // :354-359
{
    if (target == m.activeState) return AW_OK;
    eq_assign(e, m.transitionFrom, m.activeState);
    eq_assign(e, m.transitionTo, target);
    m.transitionFrame = 0;
    // the target state object starts with zero history: clear the voice it will run on
    AW_LAUNCH(e, launch_eq_reset(e->d_eq_z, m.first, m.count, 1 << (1 - m.activeVoice), e->eqStream));
    return AW_OK;
}

int eq_start_pending(aw_engine *e, EqMachine &m)
{
    if (m.pendingTarget >= 0) {
        const int pending = m.pendingTarget;
        eq_retain(e, pending);
        eq_assign(e, m.pendingTarget, -1);
        int rc = AW_OK;
        if (pending != m.activeState) rc = eq_begin_transition(e, m, pending);
        eq_release(e, pending);
        return rc;
    }
    return AW_OK;
}

bool eq_retire(aw_engine *e, EqMachine &m, int state)   // :377-389
{
    if (m.pendingRetirement >= 0) return false;
    if (m.retired < 0) { eq_assign(e, m.retired, state); return true; }
    eq_assign(e, m.pendingRetirement, state);
    return false;
}

int eq_finish_transition(aw_engine *e, EqMachine &m)   // :361-375
{
    if (m.transitionFrom < 0 || m.transitionTo < 0) return AW_OK;
    const int from = m.transitionFrom;
    eq_retain(e, from);
    eq_assign(e, m.activeState, m.transitionTo);
    m.activeVoice = 1 - m.activeVoice;
    eq_assign(e, m.transitionFrom, -1);
    eq_assign(e, m.transitionTo, -1);
    m.transitionFrame = 0;
    const bool ok = eq_retire(e, m, from);
    eq_release(e, from);
    if (!ok) return AW_OK;
    return eq_start_pending(e, m);
}

int eq_observe_published_target(aw_engine *e, EqMachine &m)   // :317-339
{
    if (!m.lockHeld && m.published >= 0) eq_assign(e, m.audioThreadTarget, m.published);
    const int target = m.audioThreadTarget;
    if (target < 0 || target == m.observedTarget) return AW_OK;
    eq_assign(e, m.observedTarget, target);
    if (m.transitionTo >= 0) {
        if (target != m.transitionTo) eq_assign(e, m.pendingTarget, target);
    } else if (m.pendingRetirement >= 0) {
        eq_assign(e, m.pendingTarget, target);
    } else if (target != m.activeState) {
        return eq_begin_transition(e, m, target);
    }
    return AW_OK;
}

int eq_flush_pending_retirement(aw_engine *e, EqMachine &m)   // :391-407
{
    if (m.pendingRetirement < 0 || m.retired >= 0) return AW_OK;
    eq_assign(e, m.retired, m.pendingRetirement);
    eq_assign(e, m.pendingRetirement, -1);
    return eq_start_pending(e, m);
}

int eq_apply_pending_reset(aw_engine *e, EqMachine &m)   // :341-352
{
    if (!m.resetRequested) return AW_OK;
    m.resetRequested = false;
    AW_LAUNCH(e, launch_eq_reset(e->d_eq_z, m.first, m.count, 3, e->eqStream));
    return AW_OK;
}

// First half of ParametricEqualizerProcessor.process (:254-267): what the render thread does before it touches samples.
int eq_begin_call(aw_engine *e, EqMachine &m)
{
    if (!m.eqActive || !m.hasProcessor) return AW_OK;   // graph bypass / EqualizerRuntimeEffect passthrough
    int rc;
    if ((rc = eq_observe_published_target(e, m)) != AW_OK) return rc;
    if ((rc = eq_flush_pending_retirement(e, m)) != AW_OK) return rc;
    return eq_apply_pending_reset(e, m);
}

// Steady state (no crossfade) with a cascade the block kernel can fuse into its epilogue?
bool eq_fusable(const aw_engine *e, const EqMachine &m, EqFuse *f)
{
    if (!m.eqActive || !m.hasProcessor || m.transitionFrom >= 0 || m.transitionTo >= 0) return false;
    const int a = m.activeState;
    if (!e->persistent || !persistent_can_fuse_eq(e->log2m, e->persistentTile, e->eqFilters[a])) return false;
    f->prog = e->d_eq_prog + a;
    f->z = e->d_eq_z;
    f->voice = m.activeVoice;
    f->n_filters = e->eqFilters[a];
    return true;
}

// Second half of ParametricEqualizerProcessor.process (:269-314) for one machine, in place on `io`.
int eq_process_machine(aw_engine *e, EqMachine &m, StridedOut io, int frames)
{
    if (!m.eqActive || !m.hasProcessor) return AW_OK;
    int rc;
    int offset = 0;
    while (offset < frames) {
        EqLaunch l;
        memset(&l, 0, sizeof(l));
        l.first_stream = m.first;
        l.n_streams = m.count;
        l.transition_length = e->transitionLength;
        if (m.transitionFrom < 0 || m.transitionTo < 0) {
            const int a = m.activeState;
            if (e->eqFilters[a] == 0 && e->eqPreamp[a] == 1.0) return AW_OK;   // unity: Float(Double(x) * 1) == x
            l.from = e->d_eq_prog + a;
            l.to = nullptr;
            l.from_voice = m.activeVoice;
            l.to_voice = 1 - m.activeVoice;
            l.seg_start = offset;
            l.seg_len = frames - offset;
            AW_LAUNCH(e, launch_eq(l, e->eqFilters[a], e->d_eq_z, io, e->eqStream));
            return AW_OK;
        }
        const int remaining = e->transitionLength - m.transitionFrame;
        const int segment = std::min(remaining, frames - offset);
        l.from = e->d_eq_prog + m.transitionFrom;
        l.to = e->d_eq_prog + m.transitionTo;
        l.from_voice = m.activeVoice;
        l.to_voice = 1 - m.activeVoice;
        l.seg_start = offset;
        l.seg_len = segment;
        l.transition_frame = m.transitionFrame;
        AW_LAUNCH(e, launch_eq(l, std::max(e->eqFilters[m.transitionFrom], e->eqFilters[m.transitionTo]), e->d_eq_z, io, e->eqStream));
        m.transitionFrame += segment;
        offset += segment;
        if (m.transitionFrame == e->transitionLength && (rc = eq_finish_transition(e, m)) != AW_OK) return rc;
    }
    return AW_OK;
}

// ---- overlapped equalizer (AW_ENGINE_OVERLAP_EQ) ------------------------------------------------------------------------
bool eq_overlap(const aw_engine *e) { return e->eqSide != nullptr && !e->profOn; }
cudaStream_t eq_base_stream(const aw_engine *e) { return eq_overlap(e) ? e->eqSide : e->stream; }

// Orders the engine's main stream behind every equalizer still in flight on the internal stream.
int join_deferred_eq(aw_engine *e)
{
    for (int p = 0; p < 2; ++p) {
        if (!e->eqPending[p]) continue;
        AW_CUDA(cudaStreamWaitEvent(e->stream, e->evEq[p], 0));
        e->eqPending[p] = false;
    }
    return AW_OK;
}

// ---- nb blocks of UPOLS for every rendering segment (nb > 1: KP only) ------------------------------------------------
int ring_modulus(const aw_engine *e, const aw_bank *b) { return b->P + e->ringExtra; }

int process_block(aw_engine *e, StridedIn cur, StridedIn prev, bool save_overlap, StridedOut out, const EqFuse &eq, int nb = 1)
{
    // fdlIndex of the first block of this call (ConvolutionEngine.swift:256-259); it moves down once per block
    for (Segment &seg : e->segments) {
        if (!seg.bank) continue;
        seg.head -= 1;
        if (seg.head < 0) seg.head += ring_modulus(e, seg.bank);
    }
    const bool literal = (e->cfg.flags & AW_ENGINE_LITERAL_STEREO) != 0;   // RealtimeAudioProcessor.swift:145
    if (e->persistent) {
        // KP: ONE launch walks the tiles of every range (up to kKpMaxSegments per launch), whatever bank each is bound to, and
        // the nb blocks of the call
        KpSegment segs[kKpMaxSegments];
        const KpCall call{nb, e->kpOrder, e->kpKeepPct, e->persistentDebug, prev.ptr == e->d_overlap ? 1 : 0};
        size_t i = 0;
        while (i < e->segments.size()) {
            int n = 0;
            for (; i < e->segments.size() && n < kKpMaxSegments; ++i) {
                const Segment &seg = e->segments[i];
                if (!seg.bank) continue;
                KpSegment &k = segs[n++];
                k.first_stream = seg.first;
                k.n_streams = seg.count;
                // rows: merged (FC + LFE share one) unless the engine is reference-literal stereo or merging is switched off
                const bool merged = e->kpMergeRows && !literal && seg.bank->d_rows != nullptr && seg.bank->R > 0;
                k.S = literal ? std::min(seg.bank->S, 2) : (merged ? seg.bank->R : seg.bank->S);
                k.rows = seg.bank->d_rows + (merged ? 0 : 1);
                k.P = seg.bank->P;
                k.Pm = ring_modulus(e, seg.bank);
                k.head = seg.head;
                k.tile0 = 0;
                k.n_big = 0;
                const bool compact = merged && seg.bank->d_bank_rows != nullptr;   // (R == S: the rows are the speakers)
                k.bank = compact ? seg.bank->d_bank_rows : seg.bank->d_bank;
                k.bank_ny = compact ? seg.bank->d_ny_rows : seg.bank->d_ny;
            }
            if (n == 0) break;
            const bool prof = e->profOn && e->profUsed + 4 <= e->profEvents.size();
            cudaEvent_t *ev = prof ? &e->profEvents[e->profUsed] : nullptr;
            if (prof) { e->profUsed += 4; cudaEventRecord(ev[0], e->stream); }
            AW_LAUNCH(e, launch_persistent(segs, n, e->S, e->P_cap, e->log2m, cur, prev, save_overlap ? e->d_overlap : nullptr, e->d_fdl,
                                           e->d_fdl_ny, e->d_tmaps, out, e->d_tw, e->persistentTile, e->persistentCtas, call, eq, e->stream));
            if (prof) { cudaEventRecord(ev[1], e->stream); cudaEventRecord(ev[2], e->stream); cudaEventRecord(ev[3], e->stream); }
        }
        // the remaining nb - 1 decrements of fdlIndex
        for (Segment &seg : e->segments) {
            if (!seg.bank || nb <= 1) continue;
            const int pm = ring_modulus(e, seg.bank);
            seg.head -= (nb - 1) % pm;
            if (seg.head < 0) seg.head += pm;
        }
        e->blocks += (unsigned long long)nb;
        return AW_OK;
    }
    if (nb != 1) return set_error(AW_ERR_UNSUPPORTED, "multi-block launches need the persistent kernel");
    int rendering = 0;
    for (const Segment &seg : e->segments) rendering += seg.bank != nullptr;
    // KF / split kernels, several ranges: fork the block onto the side streams (round-robin), join before anything else touches
    // the output
    const bool fork = rendering >= 2 && !e->profOn && e->forkEvent != nullptr;
    const int lanes = fork ? std::min(rendering, (int)aw_engine::kSideStreams) : 0;
    if (fork) {
        AW_CUDA(cudaEventRecord(e->forkEvent, e->stream));
        for (int k = 0; k < lanes; ++k) AW_CUDA(cudaStreamWaitEvent(e->side[k], e->forkEvent, 0));
    }
    int ordinal = 0;
    for (Segment &seg : e->segments) {
        if (!seg.bank) continue;
        const aw_bank *b = seg.bank;
        const cudaStream_t st = fork ? e->side[ordinal % lanes] : e->stream;
        ++ordinal;
        BlockGeom g;
        g.first_stream = seg.first;
        g.n_streams = seg.count;
        const bool merged = e->kpMergeRows && !literal && b->d_rows != nullptr && b->R > 0;
        const bool compact = merged && b->d_bank_rows != nullptr;
        const float4 *bank_rows = compact ? b->d_bank_rows : b->d_bank;
        const float *ny_rows = compact ? b->d_ny_rows : b->d_ny;
        g.S = literal ? std::min(b->S, 2) : (merged ? b->R : b->S);
        g.rows = merged ? b->d_rows : nullptr;
        g.prev_is_rows = prev.ptr == e->d_overlap ? 1 : 0;
        g.Se = e->S;
        g.B = e->B;
        g.log2m = e->log2m;
        g.P = b->P;
        g.Pm = 0;
        g.P_cap = e->P_cap;
        g.head = seg.head;
        const bool prof = e->profOn && e->profUsed + 4 <= e->profEvents.size();
        cudaEvent_t *ev = prof ? &e->profEvents[e->profUsed] : nullptr;
        if (prof) { e->profUsed += 4; cudaEventRecord(ev[0], e->stream); }
        if (e->fusedTile > 0) {
            // K2 + K3 + K4 in one kernel; events 0..1 bracket it, 1..3 collapse to zero-length intervals
            AW_LAUNCH(e, launch_fused(g, cur, prev, save_overlap ? e->d_overlap : nullptr, e->d_fdl, e->d_fdl_ny, bank_rows, ny_rows, out,
                                      e->d_tw, e->fusedTile, st));
            if (prof) { cudaEventRecord(ev[1], e->stream); cudaEventRecord(ev[2], e->stream); cudaEventRecord(ev[3], e->stream); }
        } else {
            AW_LAUNCH(e, launch_input_rfft(g, cur, prev, save_overlap ? e->d_overlap : nullptr, e->d_fdl, e->d_fdl_ny, e->d_tw, st));
            if (prof) cudaEventRecord(ev[1], e->stream);
            AW_LAUNCH(e, launch_fdl_cmac(g, e->d_fdl, bank_rows, e->d_acc, e->macTile, st));
            if (prof) cudaEventRecord(ev[2], e->stream);
            AW_LAUNCH(e, launch_irfft_out(g, e->d_acc, e->d_fdl_ny, ny_rows, out, e->d_tw, st));
            if (prof) cudaEventRecord(ev[3], e->stream);
        }
    }
    if (fork) {
        for (int k = 0; k < lanes; ++k) {
            AW_CUDA(cudaEventRecord(e->joinEvent[k], e->side[k]));
            AW_CUDA(cudaStreamWaitEvent(e->stream, e->joinEvent[k], 0));
        }
    }
    ++e->blocks;
    return AW_OK;
}

bool any_rendering(const aw_engine *e)
{
    for (const Segment &s : e->segments) if (s.bank) return true;
    return false;
}

// RealtimeAudioProcessor.process (RealtimeAudioProcessor.swift:77-119) + AudioEffectGraph routing (:179-246), device side.
int process_device_body(aw_engine *e, StridedIn in, StridedOut out, int frames, bool dup_mono);

int process_device_impl(aw_engine *e, StridedIn in, StridedOut out, int frames, bool dup_mono, bool join = false)
{
    if (frames <= 0) return AW_OK;                                              // :84
    if (frames > e->maxFrames) return set_error(AW_ERR_FRAME_COUNT, "frameCount exceeds maxFramesPerCallback");   // :85
    if (e->poisoned)
        return set_error(AW_ERR_NOT_READY, "a launch of an earlier call failed, engine state is undefined: call aw_engine_reset on the "
                                           "whole engine (AW_RESET_SPATIAL | AW_RESET_EQ) first");
    int rc = process_device_body(e, in, out, frames, dup_mono);
    if (rc == AW_OK && join) rc = join_deferred_eq(e);   // synchronous entry points: the call's equalizer is part of the call
    if (rc != AW_OK) e->poisoned = true;   // heads / counters / EQ transitions may have advanced past what the device executed
    return rc;
}

int process_device_body(aw_engine *e, StridedIn in, StridedOut out, int frames, bool dup_mono)
{
    const int B = e->B;
    const EqFuse no_eq{nullptr, nullptr, 0, 0};
    const bool overlap = eq_overlap(e);
    const cudaStream_t eq_base = eq_base_stream(e);
    e->eqStream = eq_base;
    e->eqRanThisCall = false;
    // the equalizer's per-call bookkeeping (target observation, retirement, reset) does not depend on the samples: do it first,
    // so that a steady-state cascade can ride in the block kernel's epilogue instead of a separate pass over the output
    for (EqMachine &m : e->machines) {
        const int rc = eq_begin_call(e, m);
        if (rc != AW_OK) return rc;
    }
    EqFuse fused = no_eq;
    // streams without a published renderer: passthrough copy (HRIRManager.swift:555-564)
    for (const Segment &seg : e->segments)
        if (!seg.bank) AW_LAUNCH(e, launch_passthrough(in, out, seg.first, seg.count, dup_mono ? 1 : e->S, frames, e->stream));
    if (any_rendering(e)) {
        const bool aligned = ((reinterpret_cast<uintptr_t>(in.ptr) & 7u) == 0) && (in.ss % 2 == 0) && (in.cs % 2 == 0);
        const bool fast = e->pendingCount == 0 && e->fifoCount == 0 && frames % B == 0 && aligned && !dup_mono;
        if (fast) {
            // block-aligned fast path: same arithmetic, no pending/FIFO traffic (latency of one block is zero here
            // exactly as in the reference: B frames in -> processPendingBlock -> B frames drained in the same call)
            const int nb = frames / B;
            if (e->eqFusion && e->segments.size() == 1 && e->machines.size() == 1) eq_fusable(e, e->machines[0], &fused);
            StridedIn ov{e->d_overlap, (long long)e->S * B, (long long)B};
            if (e->persistent && e->kpMultiBlock && fused.n_filters == 0) {
                // KP walks the nb blocks of the call in one launch: block b's frame is [block b-1 | block b] of `in` (the engine's
                // overlap buffer before block 0), and only the call's last block is parked for the next call
                const int rc = process_block(e, in, ov, true, out, no_eq, nb);
                if (rc != AW_OK) return rc;
            } else {
                for (int b = 0; b < nb; ++b) {
                    StridedIn cur{in.ptr + (size_t)b * B, in.ss, in.cs};
                    StridedIn prev = b == 0 ? ov : StridedIn{in.ptr + (size_t)(b - 1) * B, in.ss, in.cs};
                    StridedOut o{out.ptr + (size_t)b * B, out.ss, out.cs, 0, 0};
                    const int rc = process_block(e, cur, prev, b == nb - 1, o, fused);
                    if (rc != AW_OK) return rc;
                }
            }
        } else {
            int inputOffset = 0;
            StridedIn pend{e->d_pending, (long long)e->S * B, (long long)B};
            StridedIn ov{e->d_overlap, (long long)e->S * B, (long long)B};
            while (inputOffset < frames) {                                       // :88-116
                const int copyCount = std::min(B - e->pendingCount, frames - inputOffset);
                AW_LAUNCH(e, launch_gather_pending(in, inputOffset, copyCount, e->d_pending, e->pendingCount, e->n, e->S, B,
                                                   dup_mono ? 1 : 0, e->stream));
                e->pendingCount += copyCount;
                inputOffset += copyCount;
                if (e->pendingCount == B) {
                    const int writeIndex = (e->fifoReadIndex + e->fifoCount) % e->fifoCap;   // :167
                    StridedOut o{e->d_fifo, (long long)2 * e->fifoCap, (long long)e->fifoCap, e->fifoCap, writeIndex};
                    const int rc = process_block(e, pend, ov, true, o, no_eq);
                    if (rc != AW_OK) return rc;
                    e->fifoCount += B;
                    e->pendingCount = 0;
                }
            }
            // drain (:174-190); passthrough streams were already written, so drain only rendering segments
            for (const Segment &seg : e->segments) {
                if (!seg.bank) continue;
                StridedOut o{out.ptr + (size_t)seg.first * out.ss, out.ss, out.cs, 0, 0};
                AW_LAUNCH(e, launch_drain_fifo(e->d_fifo + (size_t)seg.first * 2 * e->fifoCap, e->fifoCap, e->fifoReadIndex, e->fifoCount, o, 0,
                                               frames, seg.count, e->stream));
            }
            const int drained = std::min(e->fifoCount, frames);
            e->fifoReadIndex = (e->fifoReadIndex + drained) % e->fifoCap;
            e->fifoCount -= drained;
        }
    }
    // equalizer after spatial (AudioEffectGraph.swift:195-210), in place on the output
    const bool prof = e->profOn && e->profEqUsed + 2 <= e->profEqEvents.size();
    if (prof) cudaEventRecord(e->profEqEvents[e->profEqUsed], e->stream);
    // several stream ranges with their own equalizers (per-device profiles): their cascades are independent, so they run
    // concurrently on the side streams and join before the call's output is handed back
    int active_machines = 0;
    for (const EqMachine &m : e->machines) active_machines += m.eqActive && m.hasProcessor;
    const bool fork_eq = fused.n_filters == 0 && active_machines >= 2 && !e->profOn;
    const int lanes = fork_eq ? std::min(active_machines, (int)aw_engine::kSideStreams) : 0;
    const bool deferred = overlap && fused.n_filters == 0 && active_machines >= 1;
    if (deferred) {
        // the call's convolution (and adapter / passthrough kernels) are all enqueued on the main stream: the equalizer follows them on
        // its own stream, and the main stream moves on to the next call without waiting for it
        AW_CUDA(cudaEventRecord(e->evKp, e->stream));
        AW_CUDA(cudaStreamWaitEvent(eq_base, e->evKp, 0));
    }
    if (fork_eq) {
        AW_CUDA(cudaEventRecord(e->forkEvent, eq_base));
        for (int k = 0; k < lanes; ++k) AW_CUDA(cudaStreamWaitEvent(e->side[k], e->forkEvent, 0));
    }
    int ordinal = 0, eq_rc = AW_OK;
    // ranges in steady state (no crossfade running) share ONE launch of the systolic kernel, whatever their equalizers
    EqSegment steady[kEqMaxSegments];
    int n_steady = 0;
    for (EqMachine &m : e->machines) {
        if (fused.n_filters != 0) continue;              // already applied by the block kernel
        const int act = m.activeState;
        const bool is_steady = active_machines >= 2 && m.eqActive && m.hasProcessor && m.transitionFrom < 0 && m.transitionTo < 0 &&
                               e->eqFilters[act] >= 1 && e->eqFilters[act] <= 32 && n_steady < kEqMaxSegments;
        if (is_steady) {
            steady[n_steady++] = EqSegment{m.first, m.count, e->eqFilters[act], m.activeVoice, 0, 0, e->d_eq_prog + act};
            continue;
        }
        if (fork_eq && m.eqActive && m.hasProcessor) e->eqStream = e->side[ordinal++ % lanes];
        eq_rc = eq_process_machine(e, m, out, frames);
        e->eqStream = eq_base;
        if (eq_rc != AW_OK) break;
    }
    if (eq_rc == AW_OK && n_steady > 0) {
        const cudaStream_t st = fork_eq ? e->side[ordinal++ % lanes] : eq_base;
        cudaError_t le = launch_eq_steady(steady, n_steady, 0, frames, e->d_eq_z, out, st);
        ++e->launches;
        if (le != cudaSuccess) eq_rc = set_error(AW_ERR_CUDA, std::string("launch_eq_steady: ") + cudaGetErrorString(le));
    }
    if (fork_eq) {
        for (int k = 0; k < lanes; ++k) {
            AW_CUDA(cudaEventRecord(e->joinEvent[k], e->side[k]));
            AW_CUDA(cudaStreamWaitEvent(eq_base, e->joinEvent[k], 0));
        }
    }
    if (eq_rc != AW_OK) return eq_rc;
    if (deferred) {
        // this call's equalizer is complete at evEq[parity]; the main stream waits for the PREVIOUS call's now — after this call's
        // convolution has been enqueued, so the two ran side by side — and for this one at the end of the next call (or in
        // aw_engine_flush / aw_engine_wait / any synchronous entry point)
        const int p = e->eqParity;
        AW_CUDA(cudaEventRecord(e->evEq[p], eq_base));
        e->eqPending[p] = true;
        e->eqRanThisCall = true;
        if (e->eqPending[p ^ 1]) {
            AW_CUDA(cudaStreamWaitEvent(e->stream, e->evEq[p ^ 1], 0));
            e->eqPending[p ^ 1] = false;
        }
        e->eqParity = p ^ 1;
    }
    if (prof) { cudaEventRecord(e->profEqEvents[e->profEqUsed + 1], e->stream); e->profEqUsed += 2; }
    return AW_OK;
}

// END REALTIME PATH

void free_engine(aw_engine *e)
{
    if (!e) return;
    DeviceGuard guard(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->d_fdl); cudaFree(e->d_fdl_ny); cudaFree(e->d_overlap); cudaFree(e->d_pending); cudaFree(e->d_fifo);
    cudaFree(e->d_acc); cudaFree(e->d_eq_z); cudaFree(e->d_eq_prog); cudaFree(e->d_tmaps);
    for (Staging &s : e->stage) {
        cudaFree(s.d_in); cudaFree(s.d_out);
        if (s.in_ready) cudaEventDestroy(s.in_ready);
        if (s.compute_done) cudaEventDestroy(s.compute_done);
        if (s.out_done) cudaEventDestroy(s.out_done);
    }
    if (e->h_in) cudaFreeHost(e->h_in);
    if (e->h_out) cudaFreeHost(e->h_out);
    for (cudaEvent_t ev : e->profEvents) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->profEqEvents) cudaEventDestroy(ev);
    for (int k = 0; k < aw_engine::kSideStreams; ++k) {
        if (e->side[k]) cudaStreamDestroy(e->side[k]);
        if (e->joinEvent[k]) cudaEventDestroy(e->joinEvent[k]);
    }
    if (e->forkEvent) cudaEventDestroy(e->forkEvent);
    if (e->eqSide) { cudaStreamSynchronize(e->eqSide); cudaStreamDestroy(e->eqSide); }
    if (e->evKp) cudaEventDestroy(e->evKp);
    for (cudaEvent_t ev : e->evEq) if (ev) cudaEventDestroy(ev);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->h2d) cudaStreamDestroy(e->h2d);
    if (e->d2h) cudaStreamDestroy(e->d2h);
    delete e;
}

// Tensor maps over the FDL for KP's stage loads (cp.async.bulk.tensor): one per (rows of a stage part, streams of a tile).
// Failure is not an error: the kernel then issues one 1-D bulk copy per stream instead.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void build_fdl_tensor_maps(aw_engine *e)
{
    const char *env = getenv("AW_KP_TENSOR_TMA");
    if (!e->persistent || (env && atoi(env) == 0)) return;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return;
    }
    const int RS = persistent_stage_rows(e->log2m), bins = persistent_stage_bins(e->log2m);
    if (RS <= 0 || bins <= 0) return;
    const cuuint64_t B = (cuuint64_t)e->B, e0 = B < 256 ? B : 256, row_bytes = B * sizeof(float2);
    // element = one bin (float2, moved as an opaque 64-bit value)
    const cuuint64_t gdim[5] = {e0, B / e0, (cuuint64_t)e->n, (cuuint64_t)e->P_cap, (cuuint64_t)e->S};
    const cuuint64_t gstride[4] = {e0 * sizeof(float2), (cuuint64_t)e->S * e->P_cap * row_bytes, row_bytes, (cuuint64_t)e->P_cap * row_bytes};
    const cuuint32_t estride[5] = {1, 1, 1, 1, 1};
    std::vector<CUtensorMap> maps((size_t)RS * 3);
    const int sizes[3] = {4, 2, 1};
    for (int rows = 1; rows <= RS; ++rows) {
        for (int k = 0; k < 3; ++k) {
            const cuuint32_t box[5] = {(cuuint32_t)((cuuint64_t)bins < e0 ? bins : e0), (cuuint32_t)(bins > 256 ? bins / 256 : 1), (cuuint32_t)sizes[k],
                                       (cuuint32_t)rows, 1};
            const CUresult r = reinterpret_cast<EncodeTiledFn>(fn)(&maps[(size_t)(rows - 1) * 3 + k], CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, e->d_fdl, gdim,
                                                                    gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return;
        }
    }
    void *d = nullptr;
    if (cudaMalloc(&d, maps.size() * sizeof(CUtensorMap)) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaMemcpy(d, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); cudaFree(d); return; }
    e->d_tmaps = d;
}

int alloc_fdl(aw_engine *e, int P_cap)
{
    const size_t rows = (size_t)e->n * e->S * P_cap;
    AW_CUDA(cudaMalloc(&e->d_fdl, rows * e->B * sizeof(float2)));
    AW_CUDA(cudaMalloc(&e->d_fdl_ny, rows * sizeof(float)));
    AW_CUDA(cudaMemsetAsync(e->d_fdl, 0, rows * e->B * sizeof(float2), e->stream));
    AW_CUDA(cudaMemsetAsync(e->d_fdl_ny, 0, rows * sizeof(float), e->stream));
    AW_CUDA(cudaStreamSynchronize(e->stream));
    e->P_cap = P_cap;
    build_fdl_tensor_maps(e);
    return AW_OK;
}

int clear_spatial_state(aw_engine *e, int first, int count)
{
    if (e->d_fdl) {
        const size_t per = (size_t)e->S * e->P_cap;
        AW_CUDA(cudaMemsetAsync(e->d_fdl + (size_t)first * per * e->B, 0, (size_t)count * per * e->B * sizeof(float2), e->stream));
        AW_CUDA(cudaMemsetAsync(e->d_fdl_ny + (size_t)first * per, 0, (size_t)count * per * sizeof(float), e->stream));
    }
    AW_CUDA(cudaMemsetAsync(e->d_overlap + (size_t)first * e->S * e->B, 0, (size_t)count * e->S * e->B * sizeof(float), e->stream));
    return AW_OK;
}

}  // namespace

// ---- library / device ----------------------------------------------------------------------------------
extern "C" const char *aw_version(void) { return "airwave-b200 0.1 (sm_100a)"; }

extern "C" const char *aw_status_string(int status)
{
    switch (status) {
    case AW_OK: return "ok";
    case AW_ERR_INVALID_ARGUMENT: return "invalid argument";
    case AW_ERR_CUDA: return "CUDA error";
    case AW_ERR_OUT_OF_MEMORY: return "out of memory";
    case AW_ERR_INVALID_BLOCK_SIZE: return "block size must be a power of two in [4, 8192]";
    case AW_ERR_FRAME_COUNT: return "frameCount exceeds maxFramesPerCallback";
    case AW_ERR_CHANNEL_MAPPING: return "HRIR channel mapping out of range";
    case AW_ERR_NO_RENDERERS: return "no valid renderers created";
    case AW_ERR_RANGE: return "stream range out of bounds";
    case AW_ERR_MISMATCH: return "bank and engine do not match";
    case AW_ERR_UNSUPPORTED: return "unsupported";
    case AW_ERR_RESAMPLE_DOWN: return "down-sampling is undefined in the reference resampler";
    case AW_ERR_NOT_READY: return "not ready";
    case AW_ERR_EQ_INVALID_SAMPLE_RATE: return "invalid sample rate";
    case AW_ERR_EQ_NON_FINITE_PREAMP: return "non-finite preamp";
    case AW_ERR_EQ_TOO_MANY_FILTERS: return "too many filters";
    case AW_ERR_EQ_INVALID_FILTER: return "invalid filter";
    case AW_ERR_WAV_READ: return "WAV read error";
    case AW_ERR_WAV_CHANNEL_COUNT: return "invalid WAV channel count";
    case AW_ERR_WAV_EMPTY: return "empty WAV file";
    case AW_ERR_WAV_UNSUPPORTED_FORMAT: return "unsupported WAV format";
    case AW_ERR_EQ_PARSE: return "equalizer parse error";
    default: return "unknown status";
    }
}

extern "C" const char *aw_last_error(void) { return g_last_error.c_str(); }

extern "C" int aw_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int aw_plan_prepare(int device, int log2n) { return get_plan(device, log2n, nullptr); }

extern "C" int aw_plan_cache_stats(int device, int *count, int *sizes, int capacity)
{
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    int c = 0;
    for (auto &kv : g_plans) {
        if (kv.first.first != device) continue;
        if (sizes && c < capacity) sizes[c] = 1 << kv.first.second;
        ++c;
    }
    if (count) *count = c;
    return AW_OK;
}

extern "C" void *aw_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

extern "C" void aw_host_free(void *ptr) { if (ptr) cudaFreeHost(ptr); }

// ---- resampler -------------------------------------------------------------------------------------------
extern "C" int aw_resample_output_count(int count, double from_rate, double to_rate)
{
    if (std::fabs(from_rate - to_rate) < 0.01) return count;     // Resampler.swift:33
    const double stride = from_rate / to_rate;                   // :38
    return (int)((double)count / stride);                        // :39
}

extern "C" int aw_resample_ex(int device, const float *input, int count, double from_rate, double to_rate, int mode, float *output,
                              int capacity, int *written)
{
    if (!input || !output || count <= 0) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_resample: bad argument");
    if (mode != AW_RESAMPLE_REFERENCE && mode != AW_RESAMPLE_CORRECT) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_resample: unknown mode");
    if (!(from_rate > 0) || !(to_rate > 0) || !std::isfinite(from_rate) || !std::isfinite(to_rate))
        return set_error(AW_ERR_INVALID_ARGUMENT, "aw_resample: sample rates must be finite and positive");
    const int outCount = aw_resample_output_count(count, from_rate, to_rate);
    if (written) *written = outCount > 0 ? outCount : 0;
    if (std::fabs(from_rate - to_rate) < 0.01) {
        if (capacity < count) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_resample: capacity too small");
        memcpy(output, input, sizeof(float) * count);
        return AW_OK;
    }
    if (outCount <= 0) return AW_OK;                             // :41
    if (mode == AW_RESAMPLE_REFERENCE && outCount < count)
        return set_error(AW_ERR_RESAMPLE_DOWN, "down-sampling reads past the control vector in the reference");
    if (capacity < outCount) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_resample: capacity too small");
    DeviceGuard guard(device);
    if (!guard.ok) return set_error(AW_ERR_CUDA, "cudaSetDevice failed");
    float *d_in = nullptr, *d_out = nullptr;
    AW_CUDA(cudaMalloc(&d_in, sizeof(float) * count));
    cudaError_t e = cudaMalloc(&d_out, sizeof(float) * outCount);
    if (e == cudaSuccess) e = cudaMemcpy(d_in, input, sizeof(float) * count, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = mode == AW_RESAMPLE_CORRECT ? launch_resample_linear(d_in, 1, count, from_rate / to_rate, d_out, outCount, 0)
                                        : launch_resample_vgenp(d_in, 1, count, (float)(from_rate / to_rate), d_out, outCount, 0);
    if (e == cudaSuccess) e = cudaMemcpy(output, d_out, sizeof(float) * outCount, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return set_error(AW_ERR_CUDA, std::string("aw_resample: ") + cudaGetErrorString(e));
    return AW_OK;
}

extern "C" int aw_resample(int device, const float *input, int count, double from_rate, double to_rate, float *output, int capacity,
                           int *written)
{
    return aw_resample_ex(device, input, count, from_rate, to_rate, AW_RESAMPLE_REFERENCE, output, capacity, written);
}

// ---- bank -------------------------------------------------------------------------------------------------
extern "C" int aw_bank_create(int device, const float *pcm, int channels, int frames, double src_rate, double dst_rate,
                              const int *left_idx, const int *right_idx, int n_speakers, int block, aw_bank **out)
{
    return aw_bank_create_ex(device, pcm, channels, frames, src_rate, dst_rate, left_idx, right_idx, n_speakers, block,
                             AW_RESAMPLE_REFERENCE, out);
}

extern "C" int aw_bank_create_ex(int device, const float *pcm, int channels, int frames, double src_rate, double dst_rate,
                                 const int *left_idx, const int *right_idx, int n_speakers, int block, int resample_mode, aw_bank **out)
{
    if (!out) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_bank_create: null out");
    *out = nullptr;
    if (!pcm || !left_idx || !right_idx || channels <= 0 || frames <= 0 || n_speakers <= 0)
        return set_error(AW_ERR_INVALID_ARGUMENT, "aw_bank_create: bad argument");
    if (!is_pow2(block) || block < 4 || block > 8192) return set_error(AW_ERR_INVALID_BLOCK_SIZE, "ConvolutionEngine: block size must be a power of two in [4, 8192]");
    if (resample_mode != AW_RESAMPLE_REFERENCE && resample_mode != AW_RESAMPLE_CORRECT)
        return set_error(AW_ERR_INVALID_ARGUMENT, "aw_bank_create: unknown resampling mode");
    // build loop of HRIRManager.activatePreset (:366-418): skip unmapped speakers, validate indices
    std::vector<int> l, r;
    for (int i = 0; i < n_speakers; ++i) {
        if (left_idx[i] < 0 || right_idx[i] < 0) continue;                      // getIndices == nil -> continue (:370-372)
        if (!(left_idx[i] < channels && right_idx[i] < channels))               // :375-379
            return set_error(AW_ERR_CHANNEL_MAPPING, "HRIR indices (" + std::to_string(left_idx[i]) + ", " + std::to_string(right_idx[i]) +
                                                         ") out of range for " + std::to_string(channels) + " channels");
        l.push_back(left_idx[i]);
        r.push_back(right_idx[i]);
    }
    const int S = (int)l.size();
    if (S == 0) return set_error(AW_ERR_NO_RENDERERS, "No valid renderers created");   // :420-422
    const bool resample = std::fabs(src_rate - dst_rate) > 0.01;                // :389
    int taps = frames;
    if (resample) {
        taps = aw_resample_output_count(frames, src_rate, dst_rate);
        if (taps <= 0) return set_error(AW_ERR_INVALID_ARGUMENT, "resampled impulse response is empty");
        if (taps < frames && resample_mode == AW_RESAMPLE_REFERENCE)
            return set_error(AW_ERR_RESAMPLE_DOWN, "down-sampling reads past the control vector in the reference");
    }
    DeviceGuard guard(device);
    if (!guard.ok) return set_error(AW_ERR_CUDA, "cudaSetDevice failed (no CUDA device? there is no CPU fallback)");
    const int log2m = ilog2(block);
    const float2 *tw = nullptr;
    int rc = get_plan(device, log2m + 1, &tw);
    if (rc != AW_OK) return rc;
    std::vector<float> ir((size_t)S * 2 * frames);
    for (int s = 0; s < S; ++s) {
        memcpy(&ir[((size_t)s * 2 + 0) * frames], pcm + (size_t)l[s] * frames, sizeof(float) * frames);
        memcpy(&ir[((size_t)s * 2 + 1) * frames], pcm + (size_t)r[s] * frames, sizeof(float) * frames);
    }
    aw_bank *b = new aw_bank();
    b->device = device; b->S = S; b->B = block; b->log2m = log2m; b->taps = taps;
    b->P = (taps + block - 1) / block;                                          // ConvolutionEngine.swift:93
    float *d_ir = nullptr, *d_rs = nullptr;
    cudaError_t e = cudaMalloc(&d_ir, sizeof(float) * ir.size());
    if (e == cudaSuccess) e = cudaMemcpy(d_ir, ir.data(), sizeof(float) * ir.size(), cudaMemcpyHostToDevice);
    const float *d_src = d_ir;
    if (e == cudaSuccess && resample) {
        e = cudaMalloc(&d_rs, sizeof(float) * (size_t)S * 2 * taps);
        if (e == cudaSuccess)
            e = resample_mode == AW_RESAMPLE_CORRECT ? launch_resample_linear(d_ir, S * 2, frames, src_rate / dst_rate, d_rs, taps, 0)
                                                     : launch_resample_vgenp(d_ir, S * 2, frames, (float)(src_rate / dst_rate), d_rs, taps, 0);
        d_src = d_rs;
    }
    int row_rep[kKpMaxRows] = {};
    if (e == cudaSuccess && S <= kKpMaxRows) {
        KpRowTable t[2];
        memset(t, -1, sizeof(t));
        int rows = 0;
        for (int sp = 0; sp < S; ++sp) {
            int row = -1;
            for (int k = 0; k < rows && row < 0; ++k)
                if (l[t[0].spk[k]] == l[sp] && r[t[0].spk[k]] == r[sp] && t[0].src[k][kKpMaxRowSources - 1] < 0) row = k;
            if (row < 0) { row = rows++; t[0].spk[row] = (signed char)sp; }
            for (int q = 0; q < kKpMaxRowSources; ++q)
                if (t[0].src[row][q] < 0) { t[0].src[row][q] = (signed char)sp; break; }
            t[1].spk[sp] = (signed char)sp;
            t[1].src[sp][0] = (signed char)sp;
        }
        b->R = rows;
        for (int k = 0; k < rows; ++k) { row_rep[k] = t[0].spk[k]; t[0].spk[k] = (signed char)k; }   // rows index the compacted bank below
        e = cudaMalloc(&b->d_rows, sizeof(t));
        if (e == cudaSuccess) e = cudaMemcpy(b->d_rows, t, sizeof(t), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMalloc(&b->d_bank, sizeof(float4) * (size_t)S * b->P * block);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_ny, sizeof(float) * (size_t)S * b->P * 2);
    if (e == cudaSuccess) e = launch_bank_build(d_src, S, taps, block, log2m, b->P, b->d_bank, b->d_ny, tw, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess && b->R > 0 && b->R < S) {
        // the filter rows of the distinct pairs, contiguous in row order: what the kernels walk when rows are shared
        const size_t row_f4 = (size_t)b->P * block, row_ny = (size_t)b->P * 2;
        e = cudaMalloc(&b->d_bank_rows, sizeof(float4) * row_f4 * b->R);
        if (e == cudaSuccess) e = cudaMalloc(&b->d_ny_rows, sizeof(float) * row_ny * b->R);
        for (int k = 0; k < b->R && e == cudaSuccess; ++k) {
            e = cudaMemcpy(b->d_bank_rows + k * row_f4, b->d_bank + row_rep[k] * row_f4, sizeof(float4) * row_f4, cudaMemcpyDeviceToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(b->d_ny_rows + k * row_ny, b->d_ny + row_rep[k] * row_ny, sizeof(float) * row_ny, cudaMemcpyDeviceToDevice);
        }
    }
    cudaFree(d_ir);
    cudaFree(d_rs);
    if (e != cudaSuccess) {
        cudaFree(b->d_bank); cudaFree(b->d_ny); cudaFree(b->d_rows); cudaFree(b->d_bank_rows); cudaFree(b->d_ny_rows);
        delete b;
        return set_error(e == cudaErrorMemoryAllocation ? AW_ERR_OUT_OF_MEMORY : AW_ERR_CUDA, std::string("aw_bank_create: ") + cudaGetErrorString(e));
    }
    *out = b;
    return AW_OK;
}

extern "C" int aw_bank_create_from_wav(int device, const aw_wav *wav, double dst_rate, int layout, int block, aw_bank **out)
{
    if (!wav || !out) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_bank_create_from_wav: null argument");
    int speakers[16], l[16], r[16];
    const int n = aw_layout_speakers(layout, speakers, 16);
    if (n == 0) return set_error(AW_ERR_INVALID_ARGUMENT, "unknown input layout");
    aw_hesuvi_map(wav->channels, speakers, n, l, r);
    return aw_bank_create(device, wav->data.data(), wav->channels, wav->frames, wav->sample_rate, dst_rate, l, r, n, block, out);
}

extern "C" int aw_bank_info(const aw_bank *bank, int *n_speakers, int *block, int *partitions, int *taps)
{
    if (!bank) return set_error(AW_ERR_INVALID_ARGUMENT, "null bank");
    if (n_speakers) *n_speakers = bank->S;
    if (block) *block = bank->B;
    if (partitions) *partitions = bank->P;
    if (taps) *taps = bank->taps;
    return AW_OK;
}

extern "C" int aw_bank_rows(const aw_bank *bank) { return bank ? bank->R : 0; }

extern "C" int aw_bank_read(const aw_bank *bank, float *spectrum, float *nyquist)
{
    if (!bank) return set_error(AW_ERR_INVALID_ARGUMENT, "null bank");
    DeviceGuard guard(bank->device);
    if (spectrum) {
        // device rows are two planes (even bins, odd bins) of B/2 float4; the host view is bin-major
        const size_t rows = (size_t)bank->S * bank->P, B = (size_t)bank->B, half = B / 2;
        std::vector<float4> tmp(rows * B);
        AW_CUDA(cudaMemcpy(tmp.data(), bank->d_bank, sizeof(float4) * rows * B, cudaMemcpyDeviceToHost));
        float4 *dst = reinterpret_cast<float4 *>(spectrum);
        for (size_t r = 0; r < rows; ++r)
            for (size_t k = 0; k < B; ++k) dst[r * B + k] = tmp[r * B + (k & 1) * half + (k >> 1)];
    }
    if (nyquist) AW_CUDA(cudaMemcpy(nyquist, bank->d_ny, sizeof(float) * (size_t)bank->S * bank->P * 2, cudaMemcpyDeviceToHost));
    return AW_OK;
}

extern "C" void aw_bank_destroy(aw_bank *bank)
{
    if (!bank) return;
    DeviceGuard guard(bank->device);
    cudaFree(bank->d_bank);
    cudaFree(bank->d_ny);
    cudaFree(bank->d_rows);
    cudaFree(bank->d_bank_rows);
    cudaFree(bank->d_ny_rows);
    delete bank;
}

// ---- engine -------------------------------------------------------------------------------------------------
extern "C" int aw_engine_create(const aw_engine_config *config, aw_engine **out)
{
    if (!out) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_engine_create: null out");
    *out = nullptr;
    if (!config) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_engine_create: null config");
    if (config->n_streams <= 0 || config->n_speakers <= 0 || config->n_speakers > 64)
        return set_error(AW_ERR_INVALID_ARGUMENT, "aw_engine_create: n_streams and n_speakers must be positive");
    if (!is_pow2(config->block) || config->block < 4 || config->block > 8192)
        return set_error(AW_ERR_INVALID_BLOCK_SIZE, "block size must be a power of two in [4, 8192]");
    const int maxFrames = config->max_frames_per_call > 0 ? config->max_frames_per_call : 4096;
    DeviceGuard guard(config->device);
    if (!guard.ok) return set_error(AW_ERR_CUDA, "cudaSetDevice failed (no CUDA device? there is no CPU fallback)");
    aw_engine *e = new aw_engine();
    e->cfg = *config;
    e->n = config->n_streams; e->S = config->n_speakers; e->B = config->block; e->log2m = ilog2(config->block);
    e->maxFrames = maxFrames;
    e->fifoCap = maxFrames + e->B;                                              // RealtimeAudioProcessor.swift:41
    int rc = get_plan(config->device, e->log2m + 1, &e->d_tw);
    if (rc != AW_OK) { delete e; return rc; }
    auto fail = [&](int status) { free_engine(e); return status; };
#define AW_TRY(call)                                                                                       \
    do {                                                                                                   \
        cudaError_t te_ = (call);                                                                          \
        if (te_ != cudaSuccess)                                                                            \
            return fail(set_error(te_ == cudaErrorMemoryAllocation ? AW_ERR_OUT_OF_MEMORY : AW_ERR_CUDA,   \
                                  std::string(#call) + ": " + cudaGetErrorString(te_)));                  \
    } while (0)
    AW_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    AW_TRY(cudaStreamCreateWithFlags(&e->h2d, cudaStreamNonBlocking));
    AW_TRY(cudaStreamCreateWithFlags(&e->d2h, cudaStreamNonBlocking));
    for (int k = 0; k < aw_engine::kSideStreams; ++k) {
        AW_TRY(cudaStreamCreateWithFlags(&e->side[k], cudaStreamNonBlocking));
        AW_TRY(cudaEventCreateWithFlags(&e->joinEvent[k], cudaEventDisableTiming));
    }
    AW_TRY(cudaEventCreateWithFlags(&e->forkEvent, cudaEventDisableTiming));
    e->eqStream = e->stream;
    if (config->flags & AW_ENGINE_OVERLAP_EQ) {
        AW_TRY(cudaStreamCreateWithFlags(&e->eqSide, cudaStreamNonBlocking));
        AW_TRY(cudaEventCreateWithFlags(&e->evKp, cudaEventDisableTiming));
        for (cudaEvent_t &ev : e->evEq) AW_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    const size_t n = e->n, S = e->S, B = e->B;
    AW_TRY(cudaMalloc(&e->d_overlap, n * S * B * sizeof(float)));
    AW_TRY(cudaMalloc(&e->d_pending, n * S * B * sizeof(float)));
    AW_TRY(cudaMalloc(&e->d_fifo, n * 2 * e->fifoCap * sizeof(float)));
    AW_TRY(cudaMalloc(&e->d_acc, n * 2 * B * sizeof(float2)));
    AW_TRY(cudaMalloc(&e->d_eq_z, n * 2 * 2 * 64 * 2 * sizeof(double)));
    AW_TRY(cudaMalloc(&e->d_eq_prog, sizeof(EqProgram) * kEqPool));
    AW_TRY(cudaMemsetAsync(e->d_overlap, 0, n * S * B * sizeof(float), e->stream));
    AW_TRY(cudaMemsetAsync(e->d_pending, 0, n * S * B * sizeof(float), e->stream));
    AW_TRY(cudaMemsetAsync(e->d_fifo, 0, n * 2 * e->fifoCap * sizeof(float), e->stream));
    AW_TRY(cudaMemsetAsync(e->d_eq_z, 0, n * 2 * 2 * 64 * 2 * sizeof(double), e->stream));
    AW_TRY(cudaMemsetAsync(e->d_eq_prog, 0, sizeof(EqProgram) * kEqPool, e->stream));
    const int sets = (config->flags & AW_ENGINE_PIPELINED) ? 2 : 1;
    for (int i = 0; i < sets; ++i) {
        AW_TRY(cudaMalloc(&e->stage[i].d_in, n * S * maxFrames * sizeof(float)));
        AW_TRY(cudaMalloc(&e->stage[i].d_out, n * 2 * maxFrames * sizeof(float)));
        AW_TRY(cudaEventCreateWithFlags(&e->stage[i].in_ready, cudaEventDisableTiming));
        AW_TRY(cudaEventCreateWithFlags(&e->stage[i].compute_done, cudaEventDisableTiming));
        AW_TRY(cudaEventCreateWithFlags(&e->stage[i].out_done, cudaEventDisableTiming));
    }
    {
        const char *zc_env = getenv("AW_ZERO_COPY");   // 0 = always stage through device memory with copies (comparisons)
        const size_t in_bytes = n * S * maxFrames * sizeof(float), out_bytes = n * 2 * maxFrames * sizeof(float);
        if (in_bytes + out_bytes <= kZeroCopyBytes && !(zc_env && atoi(zc_env) == 0)) {
            AW_TRY(cudaHostAlloc(&e->h_in, in_bytes, cudaHostAllocMapped));
            AW_TRY(cudaHostAlloc(&e->h_out, out_bytes, cudaHostAllocMapped));
        }
    }
    // unity program in slot 0 (ParametricEqualizerProcessor.unityState, :128,158)
    EqProgram unity;
    memset(&unity, 0, sizeof(unity));
    unity.preamp_linear = 1.0;
    AW_TRY(cudaMemcpyAsync(e->d_eq_prog, &unity, sizeof(unity), cudaMemcpyHostToDevice, e->stream));
    AW_TRY(cudaStreamSynchronize(e->stream));
    e->eqRef.assign(kEqPool, 0);
    e->eqFilters.assign(kEqPool, 0);
    e->eqPreamp.assign(kEqPool, 1.0);
    e->eqRef[0] = 1 << 30;
    const long tl = std::lround(config->sample_rate * 0.020);                   // ParametricEqualizerProcessor.swift:160
    e->transitionLength = tl < 1 ? 1 : (int)tl;
    e->segments.reserve(256);
    e->machines.reserve(256);
    e->segments.push_back(Segment{0, e->n, nullptr, 0});
    EqMachine m;
    m.first = 0; m.count = e->n;
    e->machines.push_back(m);
    const char *tile_env = getenv("AW_MAC_TILE");
    e->macTile = tile_env ? atoi(tile_env) : 0;
    if (!(e->macTile == 1 || e->macTile == 2 || e->macTile == 4 || e->macTile == 8))
        e->macTile = e->n >= 1024 ? 2 : 1;
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, config->device) == cudaSuccess) e->numSMs = prop.multiProcessorCount;
    }
    // Plan.  KP (persistent, aw_persistent.cu) is the default wherever it exists (64 <= B <= 2048); AW_PERSISTENT=0 selects
    // the single-wave fused kernel KF (64 <= B <= 512) and AW_FUSED_TILE=0 the split kernels K2/K3/K4 for comparisons.
    // None of these choices changes a stream's arithmetic with the number of streams: the tile size only sets how many
    // streams share one pass over the filter rows.
    e->fusedTile = 0;
    const char *fused_env = getenv("AW_FUSED_TILE");   // 0 forces the split path, 1/2/4 force a KF tile
    const bool force_split = fused_env && atoi(fused_env) == 0;
    {
        const char *p_env = getenv("AW_PERSISTENT");
        const int mask = persistent_tiles(e->log2m);
        e->persistent = mask != 0 && !force_split && !(p_env && atoi(p_env) == 0);
        const char *c_env = getenv("AW_PERSISTENT_CTAS");
        e->persistentCtas = c_env && atoi(c_env) > 0 ? atoi(c_env) : e->numSMs;
        const char *d_env = getenv("AW_PERSISTENT_DEBUG");
        e->persistentDebug = d_env ? atoi(d_env) : 0;
        const char *q_env = getenv("AW_EQ_FUSION");
        e->eqFusion = q_env && atoi(q_env) != 0;
        e->ringExtra = e->persistent ? 1 : 0;
        if (const char *v = getenv("AW_KP_RING_EXTRA")) {   // 0: the reference's modulus P; then a call is one launch per block
            if (atoi(v) == 0) { e->ringExtra = 0; e->kpMultiBlock = false; }
        }
        if (const char *v = getenv("AW_KP_MERGE_ROWS")) e->kpMergeRows = atoi(v) != 0;
        if (const char *v = getenv("AW_KP_ORDER")) e->kpOrder = atoi(v) != 0;
        if (const char *v = getenv("AW_KP_KEEP")) e->kpKeepPct = std::max(0, std::min(100, atoi(v)));
        if (const char *v = getenv("AW_KP_MULTIBLOCK")) e->kpMultiBlock = atoi(v) != 0 && e->ringExtra > 0;
        if (e->persistent) {
            // largest tile (most filter reuse) unless the smaller one loses clearly less to round quantisation
            auto eff = [&](int T) {
                const double tiles = (double)((e->n + T - 1) / T);
                return (double)e->n / ((double)T * std::ceil(tiles / e->persistentCtas) * e->persistentCtas);
            };
            // measured (C5 sweep, 2,048 streams): T = 4 wins up to B = 512 even with a ragged last round, T = 2 from B = 1024
            // (its ring is deeper there); tiny batches take T = 2 so that more SMs get a tile
            e->persistentTile = (mask & 4) && e->log2m <= 9 ? 4 : 2;
            if ((mask & 2) && (e->n + 3) / 4 < e->persistentCtas && eff(2) > eff(4)) e->persistentTile = 2;
            const char *t_env = getenv("AW_PERSISTENT_TILE");
            if (t_env && (atoi(t_env) & mask) && (atoi(t_env) == 2 || atoi(t_env) == 4)) e->persistentTile = atoi(t_env);
        }
    }
    if (!e->persistent && fused_supported(e->log2m) && !force_split) {
        // KF: the CTA owns whole streams, so the grid is n/T CTAs; prefer the largest tile that still fits in ONE wave of
        // resident CTAs, else the tile with the least wave-quantisation loss.
        int forced = fused_env ? atoi(fused_env) : -1;
        if (forced == 1 || forced == 2 || forced == 4) e->fusedTile = forced;
        else {
            double best = 0;
            const int tiles[3] = {4, 2, 1};
            for (int T : tiles) {
                const int bps = fused_blocks_per_sm(e->log2m, T);
                if (bps <= 0) continue;
                const double slots = (double)bps * e->numSMs;
                const double ctas = (double)((e->n + T - 1) / T);
                const double waves = ctas / slots;
                const double eff = waves / std::ceil(waves);            // wave-quantisation efficiency
                const double reuse = 1.0 / (1.0 + 0.5 / T);             // filter traffic through L2 costs ~0.5/T of the FDL stream
                const double score = eff * reuse * (waves <= 1.0 ? 1.0 : 0.9);
                if (score > best) { best = score; e->fusedTile = T; }
            }
        }
    }
    if (config->max_partitions > 0) {
        rc = alloc_fdl(e, config->max_partitions + e->ringExtra);
        if (rc != AW_OK) return fail(rc);
    }
#undef AW_TRY
    *out = e;
    return AW_OK;
}

extern "C" void aw_engine_destroy(aw_engine *engine) { free_engine(engine); }

extern "C" int aw_engine_set_bank(aw_engine *e, int first, int count, const aw_bank *bank)
{
    int rc = check_range(e, first, count);
    if (rc != AW_OK) return rc;
    DeviceGuard guard(e->cfg.device);
    if (bank) {
        if (bank->device != e->cfg.device || bank->B != e->B) return set_error(AW_ERR_MISMATCH, "bank device/block size differ from the engine's");
        if (bank->S > e->S) return set_error(AW_ERR_MISMATCH, "bank has more speakers than the engine has input channels");
        if (!e->d_fdl) { if ((rc = alloc_fdl(e, bank->P + e->ringExtra)) != AW_OK) return rc; }
        if (bank->P + e->ringExtra > e->P_cap) return set_error(AW_ERR_MISMATCH, "bank has more partitions than the engine's max_partitions");
    }
    if ((rc = join_deferred_eq(e)) != AW_OK) return rc;
    AW_CUDA(cudaStreamSynchronize(e->stream));
    split_segments(e, first);
    split_segments(e, first + count);
    for (Segment &s : e->segments) {
        if (s.first >= first && s.first + s.count <= first + count) { s.bank = bank; s.head = 0; }
    }
    // merge neighbours that ended up identical (same bank, same head)
    for (size_t i = 0; i + 1 < e->segments.size();) {
        Segment &a = e->segments[i], &b = e->segments[i + 1];
        if (a.bank == b.bank && a.head == b.head) { a.count += b.count; e->segments.erase(e->segments.begin() + i + 1); }
        else ++i;
    }
    if ((rc = clear_spatial_state(e, first, count)) != AW_OK) return rc;        // fresh engines: zero overlap + FDL
    // a new RealtimeAudioProcessor starts with empty pending/FIFO buffers (RealtimeAudioProcessor.swift:41-61): the rows of the
    // re-bound range are cleared, so nothing rendered with the previous bank (or never written, for a passthrough range) is
    // drained.  The adapter COUNTERS are engine-wide (all streams advance in lock-step): a range re-bound in
    // the middle of a block shares the engine's phase and emits silence for what would have been buffered samples.
    AW_CUDA(cudaMemsetAsync(e->d_pending + (size_t)first * e->S * e->B, 0, (size_t)count * e->S * e->B * sizeof(float), e->stream));
    AW_CUDA(cudaMemsetAsync(e->d_fifo + (size_t)first * 2 * e->fifoCap, 0, (size_t)count * 2 * e->fifoCap * sizeof(float), e->stream));
    if (first == 0 && count == e->n) { e->pendingCount = 0; e->fifoReadIndex = 0; e->fifoCount = 0; }
    AW_CUDA(cudaStreamSynchronize(e->stream));
    return AW_OK;
}

namespace {

enum EqMode { kEqPrepare, kEqUpdate, kEqSetTarget, kEqInstall };

int eq_control(aw_engine *e, int first, int count, double preamp_db, const aw_eq_filter *filters, int n_filters, EqMode mode,
               bool drain, int *bad_index, int *bad_reason)
{
    int rc = check_range(e, first, count);
    if (rc != AW_OK) return rc;
    if (n_filters > 0 && !filters) return set_error(AW_ERR_INVALID_ARGUMENT, "null filters");
    DeviceGuard guard(e->cfg.device);
    split_machines(e, first);
    split_machines(e, first + count);
    if (mode == kEqUpdate) {
        for (EqMachine &m : e->machines)
            if (m.first >= first && m.first + m.count <= first + count && !m.hasProcessor)
                return set_error(AW_ERR_NOT_READY, "Equalizer has not been prepared for an output.");   // EqualizerRuntimeEffect.swift:37-39
    }
    int slot = -1;
    std::string message;
    int status = eq_prepare_state(e, preamp_db, filters, n_filters, &slot, bad_index, bad_reason);
    if (status != AW_OK) {
        message = g_last_error;
        if (status == AW_ERR_OUT_OF_MEMORY || status == AW_ERR_CUDA) return status;
        if (mode == kEqSetTarget || mode == kEqInstall) return status;          // bare setTarget / prepare just throws
        int unused_i, unused_r;
        rc = eq_prepare_state(e, 0, nullptr, -1, &slot, &unused_i, &unused_r);  // try? processor.setTarget(definition: nil)
        if (rc != AW_OK) return rc;
    }
    for (EqMachine &m : e->machines) {
        if (!(m.first >= first && m.first + m.count <= first + count)) continue;
        m.hasProcessor = true;
        if (mode == kEqInstall) {   // use the state object directly: active, zero history, no transition
            eq_assign(e, m.activeState, slot);
            eq_assign(e, m.transitionFrom, -1);
            eq_assign(e, m.transitionTo, -1);
            eq_assign(e, m.pendingTarget, -1);
            eq_assign(e, m.published, slot);
            eq_assign(e, m.audioThreadTarget, slot);
            eq_assign(e, m.observedTarget, slot);
            m.transitionFrame = 0;
            m.eqActive = true;
            cudaError_t ze = launch_eq_reset(e->d_eq_z, m.first, m.count, 3, eq_base_stream(e));
            ++e->launches;
            if (ze != cudaSuccess) return set_error(AW_ERR_CUDA, std::string("launch_eq_reset: ") + cudaGetErrorString(ze));
            continue;
        }
        eq_assign(e, m.published, slot);                                        // publish (:219-226)
        if (drain) eq_assign(e, m.retired, -1);                                 // drainRetiredStates (:247-251)
        if (mode == kEqPrepare) m.eqActive = (status == AW_OK) && n_filters >= 0;   // AudioEffectGraph.swift:104-106,115-117
        else if (mode == kEqUpdate) m.eqActive = true;                          // :150-152,159-161
    }
    eq_release(e, slot);   // the machines now hold the references
    if (status != AW_OK) return set_error(status, message);
    return AW_OK;
}

}  // namespace

extern "C" int aw_engine_eq_prepare(aw_engine *e, int first, int count, double preamp_db, const aw_eq_filter *filters, int n_filters,
                                    int *bad_index, int *bad_reason)
{
    return eq_control(e, first, count, preamp_db, filters, n_filters, kEqPrepare, true, bad_index, bad_reason);
}

extern "C" int aw_engine_eq_update(aw_engine *e, int first, int count, double preamp_db, const aw_eq_filter *filters, int n_filters,
                                   int *bad_index, int *bad_reason)
{
    return eq_control(e, first, count, preamp_db, filters, n_filters, kEqUpdate, true, bad_index, bad_reason);
}

extern "C" int aw_engine_eq_set_target(aw_engine *e, int first, int count, double preamp_db, const aw_eq_filter *filters, int n_filters,
                                       int drain_retired, int *bad_index, int *bad_reason)
{
    return eq_control(e, first, count, preamp_db, filters, n_filters, kEqSetTarget, drain_retired != 0, bad_index, bad_reason);
}

extern "C" int aw_engine_eq_install_state(aw_engine *e, int first, int count, double preamp_db, const aw_eq_filter *filters, int n_filters,
                                          int *bad_index, int *bad_reason)
{
    return eq_control(e, first, count, preamp_db, filters, n_filters, kEqInstall, false, bad_index, bad_reason);
}

extern "C" int aw_engine_eq_drain_retired(aw_engine *e, int first, int count)
{
    int rc = check_range(e, first, count);
    if (rc != AW_OK) return rc;
    split_machines(e, first);
    split_machines(e, first + count);
    for (EqMachine &m : e->machines)
        if (m.first >= first && m.first + m.count <= first + count) eq_assign(e, m.retired, -1);
    return AW_OK;
}

extern "C" int aw_engine_eq_active(aw_engine *e, int first, int count, int active)
{
    int rc = check_range(e, first, count);
    if (rc != AW_OK) return rc;
    split_machines(e, first);
    split_machines(e, first + count);
    for (EqMachine &m : e->machines)
        if (m.first >= first && m.first + m.count <= first + count) m.eqActive = active != 0;
    return AW_OK;
}

extern "C" int aw_engine_eq_hold_publication(aw_engine *e, int first, int count, int held)
{
    int rc = check_range(e, first, count);
    if (rc != AW_OK) return rc;
    split_machines(e, first);
    split_machines(e, first + count);
    for (EqMachine &m : e->machines)
        if (m.first >= first && m.first + m.count <= first + count) m.lockHeld = held != 0;
    return AW_OK;
}

// BEGIN REALTIME PATH — the process entry points
extern "C" int aw_engine_process_device(aw_engine *e, const float *in, long long in_stream_stride, long long in_channel_stride,
                                        float *out, long long out_stream_stride, long long out_channel_stride, int frames)
{
    if (!e || !in || !out) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_engine_process_device: null argument");
    DeviceGuard guard(e->cfg.device);
    return process_device_impl(e, StridedIn{in, in_stream_stride, in_channel_stride},
                               StridedOut{out, out_stream_stride, out_channel_stride, 0, 0}, frames, false);
}

extern "C" int aw_engine_process(aw_engine *e, const float *in, float *out, int frames)
{
    if (!e || !in || !out) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_engine_process: null argument");
    if (frames <= 0) return AW_OK;
    if (frames > e->maxFrames) return set_error(AW_ERR_FRAME_COUNT, "frameCount exceeds maxFramesPerCallback");
    DeviceGuard guard(e->cfg.device);
    Staging &s = e->stage[0];
    const size_t inBytes = (size_t)e->n * e->S * frames * sizeof(float), outBytes = (size_t)e->n * 2 * frames * sizeof(float);
    if (e->h_in) {   // zero-copy: the kernels read the staged input and write the output over PCIe themselves
        memcpy(e->h_in, in, inBytes);
        const int zrc = process_device_impl(e, StridedIn{e->h_in, (long long)e->S * frames, (long long)frames},
                                            StridedOut{e->h_out, (long long)2 * frames, (long long)frames, 0, 0}, frames, false, true);
        if (zrc != AW_OK) return zrc;
        AW_CUDA(cudaStreamSynchronize(e->stream));
        memcpy(out, e->h_out, outBytes);
        e->h2dBytes += inBytes;
        e->d2hBytes += outBytes;
        return AW_OK;
    }
    AW_CUDA(cudaMemcpyAsync(s.d_in, in, inBytes, cudaMemcpyHostToDevice, e->stream));
    const int rc = process_device_impl(e, StridedIn{s.d_in, (long long)e->S * frames, (long long)frames},
                                       StridedOut{s.d_out, (long long)2 * frames, (long long)frames, 0, 0}, frames, false, true);
    if (rc != AW_OK) return rc;
    AW_CUDA(cudaMemcpyAsync(out, s.d_out, outBytes, cudaMemcpyDeviceToHost, e->stream));
    AW_CUDA(cudaStreamSynchronize(e->stream));
    e->h2dBytes += inBytes;
    e->d2hBytes += outBytes;
    return AW_OK;
}

extern "C" int aw_engine_process_stereo(aw_engine *e, const float *input_left, const float *input_right, float *output_left,
                                        float *output_right, int frames)
{
    if (!e || !input_left || !output_left || !output_right) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_engine_process_stereo: null argument");
    if (e->n != 1 || e->S > 2) return set_error(AW_ERR_UNSUPPORTED, "aw_engine_process_stereo needs n_streams == 1 and n_speakers <= 2");
    if (frames <= 0) return AW_OK;
    if (frames > e->maxFrames) return set_error(AW_ERR_FRAME_COUNT, "frameCount exceeds maxFramesPerCallback");
    DeviceGuard guard(e->cfg.device);
    Staging &s = e->stage[0];
    const size_t bytes = (size_t)frames * sizeof(float);
    if (e->h_in) {   // zero-copy (see aw_engine::h_in): one launch + one synchronisation per callback
        const bool zdup = input_right == nullptr;
        memcpy(e->h_in, input_left, bytes);
        if (e->S == 2 && !zdup) memcpy(e->h_in + frames, input_right, bytes);
        const int zrc = process_device_impl(e, StridedIn{e->h_in, (long long)e->S * frames, (long long)frames},
                                            StridedOut{e->h_out, (long long)2 * frames, (long long)frames, 0, 0}, frames, zdup && e->S == 2, true);
        if (zrc != AW_OK) return zrc;
        AW_CUDA(cudaStreamSynchronize(e->stream));
        // left first, right second: with aliased outputs the right channel wins, as in RealtimeAudioProcessor.swift:181-182
        memcpy(output_left, e->h_out, bytes);
        memcpy(output_right, e->h_out + frames, bytes);
        e->h2dBytes += bytes * ((e->S == 2 && !zdup) ? 2 : 1);
        e->d2hBytes += 2 * bytes;
        return AW_OK;
    }
    AW_CUDA(cudaMemcpyAsync(s.d_in, input_left, bytes, cudaMemcpyHostToDevice, e->stream));
    const bool dup = input_right == nullptr;
    if (e->S == 2 && !dup) AW_CUDA(cudaMemcpyAsync(s.d_in + frames, input_right, bytes, cudaMemcpyHostToDevice, e->stream));
    const int rc = process_device_impl(e, StridedIn{s.d_in, (long long)e->S * frames, (long long)frames},
                                       StridedOut{s.d_out, (long long)2 * frames, (long long)frames, 0, 0}, frames, dup && e->S == 2, true);
    if (rc != AW_OK) return rc;
    // left first, right second: with aliased outputs the right channel wins, as in RealtimeAudioProcessor.swift:181-182
    AW_CUDA(cudaMemcpyAsync(output_left, s.d_out, bytes, cudaMemcpyDeviceToHost, e->stream));
    AW_CUDA(cudaMemcpyAsync(output_right, s.d_out + frames, bytes, cudaMemcpyDeviceToHost, e->stream));
    AW_CUDA(cudaStreamSynchronize(e->stream));
    e->h2dBytes += bytes * ((e->S == 2 && !dup) ? 2 : 1);
    e->d2hBytes += 2 * bytes;
    return AW_OK;
}

extern "C" int aw_engine_submit(aw_engine *e, const float *in, float *out, int frames)
{
    if (!e || !in || !out) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_engine_submit: null argument");
    if (!(e->cfg.flags & AW_ENGINE_PIPELINED)) return set_error(AW_ERR_UNSUPPORTED, "engine was not created with AW_ENGINE_PIPELINED");
    if (frames <= 0) return AW_OK;
    if (frames > e->maxFrames) return set_error(AW_ERR_FRAME_COUNT, "frameCount exceeds maxFramesPerCallback");
    DeviceGuard guard(e->cfg.device);
    Staging &s = e->stage[e->nextStage];
    e->nextStage ^= 1;
    const size_t inBytes = (size_t)e->n * e->S * frames * sizeof(float), outBytes = (size_t)e->n * 2 * frames * sizeof(float);
    // staging set reuse: its previous input must have been consumed, its previous output copied out
    if (s.busy) {
        AW_CUDA(cudaStreamWaitEvent(e->h2d, s.compute_done, 0));
        AW_CUDA(cudaStreamWaitEvent(e->stream, s.out_done, 0));
    }
    AW_CUDA(cudaMemcpyAsync(s.d_in, in, inBytes, cudaMemcpyHostToDevice, e->h2d));
    AW_CUDA(cudaEventRecord(s.in_ready, e->h2d));
    AW_CUDA(cudaStreamWaitEvent(e->stream, s.in_ready, 0));
    const int rc = process_device_impl(e, StridedIn{s.d_in, (long long)e->S * frames, (long long)frames},
                                       StridedOut{s.d_out, (long long)2 * frames, (long long)frames, 0, 0}, frames, false);
    if (rc != AW_OK) return rc;
    // the staged output is complete where the call's last kernel ran: the equalizer's stream when it is overlapped
    AW_CUDA(cudaEventRecord(s.compute_done, e->eqRanThisCall ? e->eqSide : e->stream));
    AW_CUDA(cudaStreamWaitEvent(e->d2h, s.compute_done, 0));
    AW_CUDA(cudaMemcpyAsync(out, s.d_out, outBytes, cudaMemcpyDeviceToHost, e->d2h));
    AW_CUDA(cudaEventRecord(s.out_done, e->d2h));
    s.busy = true;
    e->h2dBytes += inBytes;
    e->d2hBytes += outBytes;
    return AW_OK;
}

extern "C" int aw_engine_submit_device(aw_engine *e, const float *in, long long in_stream_stride, long long in_channel_stride, float *out,
                                       int frames)
{
    if (!e || !in || !out) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_engine_submit_device: null argument");
    if (!(e->cfg.flags & AW_ENGINE_PIPELINED)) return set_error(AW_ERR_UNSUPPORTED, "engine was not created with AW_ENGINE_PIPELINED");
    if (frames <= 0) return AW_OK;
    if (frames > e->maxFrames) return set_error(AW_ERR_FRAME_COUNT, "frameCount exceeds maxFramesPerCallback");
    DeviceGuard guard(e->cfg.device);
    Staging &s = e->stage[e->nextStage];
    e->nextStage ^= 1;
    const size_t outBytes = (size_t)e->n * 2 * frames * sizeof(float);
    if (s.busy) AW_CUDA(cudaStreamWaitEvent(e->stream, s.out_done, 0));   // this set's previous output has left the device
    const int rc = process_device_impl(e, StridedIn{in, in_stream_stride, in_channel_stride},
                                       StridedOut{s.d_out, (long long)2 * frames, (long long)frames, 0, 0}, frames, false);
    if (rc != AW_OK) return rc;
    // the staged output is complete where the call's last kernel ran: the equalizer's stream when it is overlapped
    AW_CUDA(cudaEventRecord(s.compute_done, e->eqRanThisCall ? e->eqSide : e->stream));
    AW_CUDA(cudaStreamWaitEvent(e->d2h, s.compute_done, 0));
    AW_CUDA(cudaMemcpyAsync(out, s.d_out, outBytes, cudaMemcpyDeviceToHost, e->d2h));
    AW_CUDA(cudaEventRecord(s.out_done, e->d2h));
    s.busy = true;
    e->d2hBytes += outBytes;
    return AW_OK;
}

extern "C" int aw_engine_wait(aw_engine *e)
{
    if (!e) return set_error(AW_ERR_INVALID_ARGUMENT, "null engine");
    DeviceGuard guard(e->cfg.device);
    const int jrc = join_deferred_eq(e);
    if (jrc != AW_OK) return jrc;
    AW_CUDA(cudaStreamSynchronize(e->h2d));
    AW_CUDA(cudaStreamSynchronize(e->stream));
    AW_CUDA(cudaStreamSynchronize(e->d2h));
    return AW_OK;
}

extern "C" int aw_engine_flush(aw_engine *e)
{
    if (!e) return set_error(AW_ERR_INVALID_ARGUMENT, "null engine");
    DeviceGuard guard(e->cfg.device);
    return join_deferred_eq(e);
}

// END REALTIME PATH

extern "C" int aw_engine_reset(aw_engine *e, int first, int count, int what)
{
    int rc = check_range(e, first, count);
    if (rc != AW_OK) return rc;
    DeviceGuard guard(e->cfg.device);
    if (what & AW_RESET_SPATIAL) {
        // RealtimeAudioProcessor.reset (:121-139): engines reset (overlap, FDL, fdlIndex) + pending/FIFO cleared
        split_segments(e, first);
        split_segments(e, first + count);
        for (Segment &s : e->segments)
            if (s.first >= first && s.first + s.count <= first + count) s.head = 0;
        if ((rc = clear_spatial_state(e, first, count)) != AW_OK) return rc;
        AW_CUDA(cudaMemsetAsync(e->d_pending + (size_t)first * e->S * e->B, 0, (size_t)count * e->S * e->B * sizeof(float), e->stream));
        AW_CUDA(cudaMemsetAsync(e->d_fifo + (size_t)first * 2 * e->fifoCap, 0, (size_t)count * 2 * e->fifoCap * sizeof(float), e->stream));
        if (first == 0 && count == e->n) { e->pendingCount = 0; e->fifoReadIndex = 0; e->fifoCount = 0; }
        AW_CUDA(cudaStreamSynchronize(e->stream));
    }
    if (first == 0 && count == e->n && (what & AW_RESET_SPATIAL) && (what & AW_RESET_EQ)) e->poisoned = false;
    if (what & AW_RESET_EQ) {
        split_machines(e, first);
        split_machines(e, first + count);
        for (EqMachine &m : e->machines)
            if (m.first >= first && m.first + m.count <= first + count) m.resetRequested = true;   // :240-244
    }
    return AW_OK;
}

extern "C" int aw_engine_counters(const aw_engine *e, unsigned long long *kernel_launches, unsigned long long *blocks,
                                  unsigned long long *h2d_bytes, unsigned long long *d2h_bytes)
{
    if (!e) return set_error(AW_ERR_INVALID_ARGUMENT, "null engine");
    if (kernel_launches) *kernel_launches = e->launches;
    if (blocks) *blocks = e->blocks;
    if (h2d_bytes) *h2d_bytes = e->h2dBytes;
    if (d2h_bytes) *d2h_bytes = e->d2hBytes;
    return AW_OK;
}

extern "C" int aw_engine_plan(const aw_engine *e, int *fused_tile, int *mac_tile, int *partitions_cap)
{
    if (!e) return set_error(AW_ERR_INVALID_ARGUMENT, "null engine");
    if (fused_tile) *fused_tile = e->persistent ? e->persistentTile : e->fusedTile;
    if (mac_tile) *mac_tile = e->macTile;
    if (partitions_cap) *partitions_cap = e->P_cap > 0 ? e->P_cap - e->ringExtra : 0;
    return AW_OK;
}

extern "C" int aw_engine_profile_begin(aw_engine *e, int max_blocks)
{
    if (!e || max_blocks <= 0) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_engine_profile_begin: bad argument");
    DeviceGuard guard(e->cfg.device);
    const size_t want = (size_t)max_blocks * 4 * std::max<size_t>(1, e->segments.size());
    while (e->profEvents.size() < want) {
        cudaEvent_t ev;
        AW_CUDA(cudaEventCreate(&ev));
        e->profEvents.push_back(ev);
    }
    while (e->profEqEvents.size() < (size_t)max_blocks * 2) {
        cudaEvent_t ev;
        AW_CUDA(cudaEventCreate(&ev));
        e->profEqEvents.push_back(ev);
    }
    e->profUsed = 0;
    e->profEqUsed = 0;
    e->profOn = true;
    return AW_OK;
}

extern "C" int aw_engine_profile_end(aw_engine *e, double *kernel_ms, unsigned long long *kernel_launches)
{
    if (!e) return set_error(AW_ERR_INVALID_ARGUMENT, "null engine");
    DeviceGuard guard(e->cfg.device);
    e->profOn = false;
    AW_CUDA(cudaStreamSynchronize(e->stream));
    double ms[4] = {0, 0, 0, 0};
    unsigned long long cnt[4] = {0, 0, 0, 0};
    for (size_t i = 0; i + 4 <= e->profUsed; i += 4) {
        for (int k = 0; k < 3; ++k) {
            float t = 0.f;
            AW_CUDA(cudaEventElapsedTime(&t, e->profEvents[i + k], e->profEvents[i + k + 1]));
            ms[k] += t;
            ++cnt[k];
        }
    }
    for (size_t i = 0; i + 2 <= e->profEqUsed; i += 2) {   // slot 3: the equalizer launches of one process call
        float t = 0.f;
        AW_CUDA(cudaEventElapsedTime(&t, e->profEqEvents[i], e->profEqEvents[i + 1]));
        ms[3] += t;
        ++cnt[3];
    }
    for (int k = 0; k < 4; ++k) {
        if (kernel_ms) kernel_ms[k] = ms[k];
        if (kernel_launches) kernel_launches[k] = cnt[k];
    }
    e->profUsed = 0;
    e->profEqUsed = 0;
    return AW_OK;
}

extern "C" int aw_engine_kernels(const aw_engine *e, char *names, int capacity)
{
    if (!e || !names || capacity <= 0) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_engine_kernels: bad argument");
    const int lb = e->log2m;
    std::string n;
    if (e->persistent) n = "k_persistent<" + std::to_string(lb) + "," + std::to_string(e->persistentTile) + ">";
    else if (e->fusedTile > 0) n = "k_fused<" + std::to_string(lb) + "," + std::to_string(e->fusedTile) + ">";
    else n = "k_input_rfft<" + std::to_string(lb) + ">;k_fdl_cmac<" + std::to_string(e->macTile) + ">;k_irfft_out<" + std::to_string(lb) + ">";
    if ((int)n.size() + 1 > capacity) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_engine_kernels: capacity too small");
    memcpy(names, n.c_str(), n.size() + 1);
    return AW_OK;
}

extern "C" int aw_engine_uses_tensor_maps(const aw_engine *e) { return e && e->d_tmaps ? 1 : 0; }

extern "C" void *aw_engine_stream(const aw_engine *e) { return e ? (void *)e->stream : nullptr; }

extern "C" int aw_synth_fill_device(int device, float *d_out, int first_stream, int n_streams, int n_speakers, long long frame0,
                                    int frames, uint32_t seed, void *cuda_stream)
{
    if (!d_out) return set_error(AW_ERR_INVALID_ARGUMENT, "null output");
    DeviceGuard guard(device);
    AW_CUDA(launch_synth_fill(d_out, first_stream, n_streams, n_speakers, frame0, frames, seed, (cudaStream_t)cuda_stream));
    return AW_OK;
}
