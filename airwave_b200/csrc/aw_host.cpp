// aw_host.cpp — setup-time host logic behind the C ABI (no CUDA here): WAV loading, speaker
// layouts and HeSuVi channel maps, EqualizerAPO text parsing, biquad coefficient design.
// Mirrors WAVLoader.swift, VirtualSpeaker.swift, EqualizerAPOParser.swift and
// BiquadCoefficientBuilder.swift of the reference (file:line cited per function).
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/airwave_cuda.h"
#include "aw_internal.h"

using namespace aw;

// ---------------------------------------------------------------------------------------------
// WAVLoader.load  (WAVLoader.swift:26-99) — AVAudioFile replaced by a RIFF/WAVE chunk parser (Q13)
// ---------------------------------------------------------------------------------------------
static uint32_t rd32(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t rd16(const unsigned char *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static const long long kWavMaxFrames = 1ll << 28;   // frames fit an int; channels <= 65535 by the field width

extern "C" int aw_wav_load_memory(const void *bytes, size_t size, aw_wav **out)
{
    if (!bytes || !out) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_wav_load_memory: null argument");
    *out = nullptr;
    const unsigned char *d = (const unsigned char *)bytes;
    if (size < 12 || memcmp(d, "RIFF", 4) != 0 || memcmp(d + 8, "WAVE", 4) != 0)
        return set_error(AW_ERR_WAV_READ, "WAV file read error: not a RIFF/WAVE file");
    const unsigned char *fmt = nullptr, *pcm = nullptr;
    size_t fmt_size = 0, pcm_size = 0, pos = 12;
    while (pos + 8 <= size) {
        const uint32_t csize = rd32(d + pos + 4);
        const size_t avail = size - (pos + 8);
        const size_t body = csize < avail ? csize : avail;
        if (memcmp(d + pos, "fmt ", 4) == 0) { fmt = d + pos + 8; fmt_size = body; }
        else if (memcmp(d + pos, "data", 4) == 0) { pcm = d + pos + 8; pcm_size = body; }
        pos += 8 + (size_t)csize + (csize & 1u);
    }
    if (!fmt || !pcm || fmt_size < 16) return set_error(AW_ERR_WAV_READ, "WAV file read error: missing fmt or data chunk");
    uint16_t tag = rd16(fmt);
    const int channels = rd16(fmt + 2);
    const uint32_t rate = rd32(fmt + 4);
    const int block_align = rd16(fmt + 12);
    const int bits = rd16(fmt + 14);
    if (tag == 0xFFFE && fmt_size >= 26) tag = rd16(fmt + 24);   // WAVE_FORMAT_EXTENSIBLE sub-format
    if (channels <= 0) return set_error(AW_ERR_WAV_CHANNEL_COUNT, "Invalid channel count: 0. WAV file must have at least 1 channel.");
    const size_t frames = block_align > 0 ? pcm_size / (size_t)block_align : 0;
    if (frames == 0) return set_error(AW_ERR_WAV_EMPTY, "WAV file is empty (0 frames)");
    enum { F32, F64, I16, I24, I32 } kind;
    if (tag == 3 && bits == 32) kind = F32;
    else if (tag == 3 && bits == 64) kind = F64;
    else if (tag == 1 && bits == 16) kind = I16;
    else if (tag == 1 && bits == 24) kind = I24;
    else if (tag == 1 && bits == 32) kind = I32;
    else return set_error(AW_ERR_WAV_UNSUPPORTED_FORMAT, "Unsupported WAV format");
    // The header is untrusted (the reference gets this validation from AVAudioFile): a frame must hold all of its
    // samples, or the reads below would run past the data chunk; and the planar copy must be allocatable.
    if (block_align < channels * (bits / 8))
        return set_error(AW_ERR_WAV_UNSUPPORTED_FORMAT, "Unsupported WAV format: block alignment smaller than channels x sample size");
    if (frames > (size_t)kWavMaxFrames) return set_error(AW_ERR_WAV_UNSUPPORTED_FORMAT, "Unsupported WAV format: more than 2^28 frames");
    aw_wav *w = nullptr;
    try {
        w = new aw_wav();
        w->data.resize((size_t)channels * frames);
    } catch (const std::bad_alloc &) {
        delete w;
        return set_error(AW_ERR_OUT_OF_MEMORY, "WAV file too large to load");
    }
    w->sample_rate = (double)rate;
    w->channels = channels;
    w->frames = (int)frames;
    const int bps = bits / 8;
    for (size_t f = 0; f < frames; ++f) {
        for (int c = 0; c < channels; ++c) {
            const unsigned char *p = pcm + f * (size_t)block_align + (size_t)c * bps;
            float v;
            switch (kind) {
            case F32: { uint32_t u = rd32(p); memcpy(&v, &u, 4); break; }
            case F64: { double dv; memcpy(&dv, p, 8); v = (float)dv; break; }
            case I16: v = (float)(int16_t)rd16(p) / 32768.0f; break;                       // WAVLoader.swift:78
            case I24: { int32_t s = (int32_t)(p[0] | (p[1] << 8) | (p[2] << 16)); if (s & 0x800000) s -= (1 << 24); v = (float)s / 8388608.0f; break; }
            default: v = (float)(int32_t)rd32(p) / 2147483648.0f; break;                   // WAVLoader.swift:86
            }
            w->data[(size_t)c * frames + f] = v;
        }
    }
    *out = w;
    return AW_OK;
}

extern "C" int aw_wav_load(const char *path, aw_wav **out)
{
    if (!path || !out) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_wav_load: null argument");
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return set_error(AW_ERR_WAV_READ, std::string("WAV file read error: Failed to open WAV file: ") + path);
    std::vector<unsigned char> buf;
    unsigned char tmp[65536];
    size_t n;
    try {
        while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    } catch (const std::bad_alloc &) {
        fclose(f);
        return set_error(AW_ERR_OUT_OF_MEMORY, "WAV file too large to load");
    }
    fclose(f);
    return aw_wav_load_memory(buf.data(), buf.size(), out);
}

extern "C" int aw_wav_info(const aw_wav *wav, double *sample_rate, int *channels, int *frames)
{
    if (!wav) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_wav_info: null wav");
    if (sample_rate) *sample_rate = wav->sample_rate;
    if (channels) *channels = wav->channels;
    if (frames) *frames = wav->frames;
    return AW_OK;
}

extern "C" const float *aw_wav_channel(const aw_wav *wav, int channel)
{
    if (!wav || channel < 0 || channel >= wav->channels) return nullptr;
    return wav->data.data() + (size_t)channel * wav->frames;
}

extern "C" void aw_wav_destroy(aw_wav *wav) { delete wav; }

// ---------------------------------------------------------------------------------------------
// InputLayout (VirtualSpeaker.swift:59-100), HRIRChannelMap (VirtualSpeaker.swift:224-346)
// ---------------------------------------------------------------------------------------------
extern "C" int aw_layout_speakers(int layout, int *speakers, int capacity)
{
    static const int order[12] = {AW_SPK_FL, AW_SPK_FR, AW_SPK_FC, AW_SPK_LFE, AW_SPK_BL, AW_SPK_BR,
                                  AW_SPK_SL, AW_SPK_SR, AW_SPK_TFL, AW_SPK_TFR, AW_SPK_TBL, AW_SPK_TBR};
    int n;
    switch (layout) {
    case AW_LAYOUT_STEREO: n = 2; break;        // :64-67
    case AW_LAYOUT_SURROUND51: n = 6; break;    // :70-73
    case AW_LAYOUT_SURROUND71: n = 8; break;    // :76-79
    case AW_LAYOUT_ATMOS714: n = 12; break;     // :82-85
    default: return 0;
    }
    for (int i = 0; i < n && i < capacity; ++i) speakers[i] = order[i];
    return n;
}

extern "C" int aw_hesuvi_map(int wav_channels, const int *speakers, int n_speakers, int *left_idx, int *right_idx)
{
    if (!speakers || !left_idx || !right_idx || n_speakers < 0) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_hesuvi_map: null argument");
    //                         FL  FR  FC LFE  BL  BR  SL  SR
    static const int l14[8] = {0, 8, 6, 6, 4, 12, 2, 10}, r14[8] = {1, 7, 13, 13, 5, 11, 3, 9};   // :270-297
    static const int l7[8] = {0, 1, 2, 2, 3, 4, 5, 6}, r7[8] = {1, 0, 2, 2, 4, 3, 6, 5};          // :224-250
    const bool seven = wav_channels == 7;                                                         // HRIRManager.swift:355-360
    for (int i = 0; i < n_speakers; ++i) {
        const int sp = speakers[i];
        if (sp >= AW_SPK_FL && sp <= AW_SPK_SR) {
            left_idx[i] = seven ? l7[sp] : l14[sp];
            right_idx[i] = seven ? r7[sp] : r14[sp];
        } else { left_idx[i] = -1; right_idx[i] = -1; }   // `default: break` — no mapping, the speaker is skipped
    }
    return AW_OK;
}

static std::string trim_ws(const std::string &s, const char *set = " \t")
{
    const size_t a = s.find_first_not_of(set);
    if (a == std::string::npos) return "";
    const size_t b = s.find_last_not_of(set);
    return s.substr(a, b - a + 1);
}

static std::string upper(std::string s) { for (auto &c : s) c = (char)toupper((unsigned char)c); return s; }

static bool parse_int_strict(const std::string &s, int *out)
{
    if (s.empty()) return false;
    size_t i = (s[0] == '+' || s[0] == '-') ? 1 : 0;
    if (i >= s.size()) return false;
    for (size_t k = i; k < s.size(); ++k) if (!isdigit((unsigned char)s[k])) return false;
    *out = atoi(s.c_str());
    return true;
}

extern "C" int aw_hesuvi_parse(const char *text, int *left_idx, int *right_idx)   // :301-346
{
    if (!text || !left_idx || !right_idx) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_hesuvi_parse: null argument");
    for (int i = 0; i < AW_SPK_COUNT; ++i) { left_idx[i] = -1; right_idx[i] = -1; }
    static const struct { const char *name; int sp; } names[] = {
        {"FL", AW_SPK_FL}, {"L", AW_SPK_FL}, {"FR", AW_SPK_FR}, {"R", AW_SPK_FR}, {"FC", AW_SPK_FC}, {"C", AW_SPK_FC},
        {"LFE", AW_SPK_LFE}, {"SUB", AW_SPK_LFE}, {"BL", AW_SPK_BL}, {"RL", AW_SPK_BL}, {"BR", AW_SPK_BR}, {"RR", AW_SPK_BR},
        {"SL", AW_SPK_SL}, {"SR", AW_SPK_SR}, {"TFL", AW_SPK_TFL}, {"TFR", AW_SPK_TFR}, {"TBL", AW_SPK_TBL}, {"TBR", AW_SPK_TBR}};
    const std::string src(text);
    size_t pos = 0;
    while (pos <= src.size()) {
        size_t end = src.find_first_of("\r\n", pos);
        if (end == std::string::npos) end = src.size();
        const std::string line = trim_ws(src.substr(pos, end - pos));
        pos = end + 1;
        if (line.empty() || line[0] == '#' || line[0] == ';') continue;
        const size_t eq = line.find('=');
        if (eq == std::string::npos || line.find('=', eq + 1) != std::string::npos) continue;   // parts.count == 2
        const std::string name = upper(trim_ws(line.substr(0, eq)));
        std::vector<int> idx;
        std::string rest = trim_ws(line.substr(eq + 1));
        size_t p = 0;
        while (p <= rest.size()) {
            size_t c = rest.find(',', p);
            if (c == std::string::npos) c = rest.size();
            int v;
            if (parse_int_strict(trim_ws(rest.substr(p, c - p)), &v)) idx.push_back(v);   // compactMap { Int(...) }
            p = c + 1;
        }
        if (idx.size() != 2) continue;
        for (const auto &n : names)
            if (name == n.name) { left_idx[n.sp] = idx[0]; right_idx[n.sp] = idx[1]; break; }
        // unknown names become .custom(name) in the reference: not representable here, ignored
    }
    return AW_OK;
}

// ---------------------------------------------------------------------------------------------
// BiquadCoefficientBuilder.make  (BiquadCoefficientBuilder.swift:30-107)
// ---------------------------------------------------------------------------------------------
extern "C" int aw_biquad_make(int type, double gainDB, double frequencyHz, double q, double sampleRate, double *out5)
{
    if (!out5) return AW_BIQUAD_NON_FINITE_INPUT;
    if (!(std::isfinite(sampleRate) && sampleRate > 0)) return AW_BIQUAD_INVALID_SAMPLE_RATE;                   // :37
    if (!(std::isfinite(gainDB) && std::isfinite(frequencyHz) && std::isfinite(q))) return AW_BIQUAD_NON_FINITE_INPUT;   // :40
    if (!(frequencyHz > 0 && frequencyHz < sampleRate / 2)) return AW_BIQUAD_INVALID_FREQUENCY;                 // :43
    if (!(q > 0)) return AW_BIQUAD_INVALID_Q;                                                                   // :46
    const double A = std::pow(10.0, gainDB / 40.0);                                                             // :50
    const double omega = 2.0 * M_PI * frequencyHz / sampleRate;
    const double sn = std::sin(omega), cs = std::cos(omega);
    const double alpha = sn / (2.0 * q);
    const double beta = 2.0 * std::sqrt(A) * alpha;                                                             // :55
    double b0, b1, b2, a0, a1, a2;
    switch (type) {
    case AW_FILTER_PEAKING:                                                                                     // :59-67
        b0 = 1 + alpha * A; b1 = -2 * cs; b2 = 1 - alpha * A;
        a0 = 1 + alpha / A; a1 = -2 * cs; a2 = 1 - alpha / A;
        break;
    case AW_FILTER_LOW_SHELF:                                                                                   // :68-76
        b0 = A * ((A + 1) - (A - 1) * cs + beta);
        b1 = 2 * A * ((A - 1) - (A + 1) * cs);
        b2 = A * ((A + 1) - (A - 1) * cs - beta);
        a0 = (A + 1) + (A - 1) * cs + beta;
        a1 = -2 * ((A - 1) + (A + 1) * cs);
        a2 = (A + 1) + (A - 1) * cs - beta;
        break;
    case AW_FILTER_HIGH_SHELF:                                                                                  // :77-85
        b0 = A * ((A + 1) + (A - 1) * cs + beta);
        b1 = -2 * A * ((A - 1) + (A + 1) * cs);
        b2 = A * ((A + 1) + (A - 1) * cs - beta);
        a0 = (A + 1) - (A - 1) * cs + beta;
        a1 = 2 * ((A - 1) - (A + 1) * cs);
        a2 = (A + 1) - (A - 1) * cs - beta;
        break;
    default: return AW_BIQUAD_NON_FINITE_INPUT;
    }
    if (!(std::isfinite(a0) && a0 != 0)) return AW_BIQUAD_NON_FINITE_COEFFICIENTS;                              // :88
    out5[0] = b0 / a0; out5[1] = b1 / a0; out5[2] = b2 / a0; out5[3] = a1 / a0; out5[4] = a2 / a0;
    for (int i = 0; i < 5; ++i) if (!std::isfinite(out5[i])) return AW_BIQUAD_NON_FINITE_COEFFICIENTS;          // :99-105
    return AW_BIQUAD_OK;
}

// ---------------------------------------------------------------------------------------------
// EqualizerAPOParser.parse  (EqualizerAPOParser.swift:36-151).  The two NSRegularExpressions
// (:27-34) are matched by a hand-written tokenizer with the same grammar.
// ---------------------------------------------------------------------------------------------
namespace {

struct Cursor {
    const std::string &s;
    size_t i;
    bool ws0() { while (i < s.size() && isspace((unsigned char)s[i])) ++i; return true; }        // \s*
    bool ws1() { const size_t a = i; ws0(); return i > a; }                                        // \s+
    bool lit(const char *w) {                                                                       // case-insensitive literal
        const size_t n = strlen(w);
        if (i + n > s.size()) return false;
        for (size_t k = 0; k < n; ++k) if (tolower((unsigned char)s[i + k]) != tolower((unsigned char)w[k])) return false;
        i += n;
        return true;
    }
    bool token(std::string *out) {                                                                  // (\S+)
        const size_t a = i;
        while (i < s.size() && !isspace((unsigned char)s[i])) ++i;
        if (i == a) return false;
        *out = s.substr(a, i - a);
        return true;
    }
    bool end() const { return i == s.size(); }
};

// Swift Double(String): whole-string decimal/hex float, "inf"/"infinity"/"nan" accepted, no whitespace.
bool finite_double(const std::string &t, double *out)
{
    if (t.empty() || isspace((unsigned char)t[0])) return false;
    const std::string body = upper((t[0] == '+' || t[0] == '-') ? t.substr(1) : t);
    if (body.empty()) return false;
    if (!(isdigit((unsigned char)body[0]) || body[0] == '.' || body == "INF" || body == "INFINITY" || body == "NAN")) return false;
    char *endp = nullptr;
    const double v = strtod(t.c_str(), &endp);
    if (endp == t.c_str() || *endp != '\0') return false;
    if (!std::isfinite(v)) return false;
    *out = v;
    return true;
}

bool match_preamp(const std::string &line, std::string *value)   // ^Preamp\s*:\s*(\S+)\s+dB$
{
    Cursor c{line, 0};
    if (!c.lit("preamp")) return false;
    c.ws0();
    if (!c.lit(":")) return false;
    c.ws0();
    // (\S+)\s+dB$ with backtracking: the value token is everything up to the last whitespace run before a final "dB"
    const size_t start = c.i;
    if (line.size() < start + 4) return false;
    const std::string tail = line.substr(line.size() - 2);
    if (tolower((unsigned char)tail[0]) != 'd' || tolower((unsigned char)tail[1]) != 'b') return false;
    size_t e = line.size() - 2;
    if (e == start || !isspace((unsigned char)line[e - 1])) return false;
    while (e > start && isspace((unsigned char)line[e - 1])) --e;
    if (e == start) return false;
    const std::string tok = line.substr(start, e - start);
    for (char ch : tok) if (isspace((unsigned char)ch)) return false;
    *value = tok;
    return true;
}

// ^Filter(?:\s+([0-9]+))?\s*:\s+(ON|OFF)\s+(PK|LSC|HSC)\s+Fc\s+(\S+)\s+Hz\s+Gain\s+(\S+)\s+dB\s+Q\s+(\S+)$
bool match_filter(const std::string &line, std::string cap[6])
{
    Cursor c{line, 0};
    if (!c.lit("filter")) return false;
    cap[0].clear();
    {
        const size_t save = c.i;
        if (c.ws1()) {
            const size_t a = c.i;
            while (c.i < line.size() && isdigit((unsigned char)line[c.i])) ++c.i;
            if (c.i > a) cap[0] = line.substr(a, c.i - a);
            else c.i = save;
        }
    }
    c.ws0();
    if (!c.lit(":")) return false;
    if (!c.ws1()) return false;
    if (c.lit("on")) cap[1] = "ON"; else if (c.lit("off")) cap[1] = "OFF"; else return false;
    if (!c.ws1()) return false;
    if (c.lit("pk")) cap[2] = "PK"; else if (c.lit("lsc")) cap[2] = "LSC"; else if (c.lit("hsc")) cap[2] = "HSC"; else return false;
    if (!c.ws1() || !c.lit("fc") || !c.ws1() || !c.token(&cap[3]) || !c.ws1() || !c.lit("hz")) return false;
    if (!c.ws1() || !c.lit("gain") || !c.ws1() || !c.token(&cap[4]) || !c.ws1() || !c.lit("db")) return false;
    if (!c.ws1() || !c.lit("q") || !c.ws1() || !c.token(&cap[5])) return false;
    return c.end();
}

bool starts_with_ci(const std::string &s, const char *prefix)
{
    const size_t n = strlen(prefix);
    if (s.size() < n) return false;
    for (size_t k = 0; k < n; ++k) if (tolower((unsigned char)s[k]) != prefix[k]) return false;
    return true;
}

bool valid_utf8(const unsigned char *p, size_t n)
{
    size_t i = 0;
    while (i < n) {
        const unsigned char c = p[i];
        int extra;
        if (c < 0x80) extra = 0;
        else if ((c & 0xE0) == 0xC0 && c >= 0xC2) extra = 1;
        else if ((c & 0xF0) == 0xE0) extra = 2;
        else if ((c & 0xF8) == 0xF0 && c <= 0xF4) extra = 3;
        else return false;
        for (int k = 1; k <= extra; ++k) { if (i + k >= n || (p[i + k] & 0xC0) != 0x80) return false; }
        i += (size_t)extra + 1;
    }
    return true;
}

}  // namespace

extern "C" int aw_eq_parse(const void *bytes, size_t size, double *preamp_db, aw_eq_filter *filters, int capacity, int *n_filters,
                           char *issues_out, size_t issues_capacity)
{
    if (n_filters) *n_filters = 0;
    if (issues_out && issues_capacity) issues_out[0] = '\0';
    std::vector<std::string> issues;
    auto issue = [&](int line, const std::string &reason) {
        issues.push_back(line > 0 ? ("line " + std::to_string(line) + ": " + reason) : reason);
    };
    auto fail = [&]() {
        std::string joined;
        for (size_t i = 0; i < issues.size(); ++i) { if (i) joined += "; "; joined += issues[i]; }
        if (issues_out && issues_capacity) { snprintf(issues_out, issues_capacity, "%s", joined.c_str()); }
        return set_error(AW_ERR_EQ_PARSE, joined);
    };
    if (!bytes && size) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_eq_parse: null data");
    if (size > 1048576) { issue(0, "file exceeds the 1 MiB limit"); return fail(); }                 // :37-42
    const unsigned char *d = (const unsigned char *)bytes;
    if (!valid_utf8(d, size)) { issue(0, "file is not valid UTF-8"); return fail(); }                // :43-48
    std::string source((const char *)d, size);
    if (source.size() >= 3 && (unsigned char)source[0] == 0xEF && (unsigned char)source[1] == 0xBB && (unsigned char)source[2] == 0xBF)
        source.erase(0, 3);                                                                          // :49-51
    double preamp = 0.0;
    bool hasPreamp = false;
    int declCount = 0;
    std::vector<aw_eq_filter> parsed;
    int lineNumber = 0;
    size_t pos = 0;
    while (pos <= source.size()) {
        // components(separatedBy: .newlines): every newline scalar splits (LF, VT, FF, CR, NEL, LS, PS)
        size_t end = pos;
        size_t next = std::string::npos;
        for (; end < source.size(); ++end) {
            const unsigned char c = (unsigned char)source[end];
            if (c == '\n' || c == '\r' || c == 0x0B || c == 0x0C) { next = end + 1; break; }
            if (c == 0xC2 && end + 1 < source.size() && (unsigned char)source[end + 1] == 0x85) { next = end + 2; break; }
            if (c == 0xE2 && end + 2 < source.size() && (unsigned char)source[end + 1] == 0x80 &&
                ((unsigned char)source[end + 2] == 0xA8 || (unsigned char)source[end + 2] == 0xA9)) { next = end + 3; break; }
        }
        ++lineNumber;
        std::string line = trim_ws(source.substr(pos, end - pos), " \t\n\r\x0b\x0c");
        pos = (next == std::string::npos) ? source.size() + 1 : next;
        if (line.empty() || line[0] == '#') continue;                                                // :62
        std::string value;
        if (match_preamp(line, &value)) {                                                            // :64-76
            if (hasPreamp) { issue(lineNumber, "duplicate Preamp directive"); continue; }
            double v;
            if (!finite_double(value, &v)) { issue(lineNumber, "Preamp must be a finite number"); continue; }
            preamp = v; hasPreamp = true;
            continue;
        }
        if (starts_with_ci(line, "filter")) {                                                        // :78-136
            ++declCount;
            if (declCount > 64) { issue(lineNumber, "more than 64 filter declarations are not allowed"); continue; }
            std::string cap[6];
            if (!match_filter(line, cap)) { issue(lineNumber, "malformed Filter directive"); continue; }
            aw_eq_filter f;
            memset(&f, 0, sizeof(f));
            f.source_line = lineNumber;
            f.source_number = cap[0].empty() ? -1 : atoi(cap[0].c_str());
            f.enabled = cap[1] == "ON";
            f.type = cap[2] == "PK" ? AW_FILTER_PEAKING : (cap[2] == "LSC" ? AW_FILTER_LOW_SHELF : AW_FILTER_HIGH_SHELF);
            double fc = 0, gain = 0, q = 0;
            const bool okF = finite_double(cap[3], &fc), okG = finite_double(cap[4], &gain), okQ = finite_double(cap[5], &q);
            std::vector<std::string> numeric;
            if (okF) { if (fc <= 0) numeric.push_back("frequency must be positive"); }
            else numeric.push_back("frequency must be a finite number");
            if (!okG) numeric.push_back("gain must be a finite number");
            if (okQ) { if (q <= 0) numeric.push_back("Q must be positive"); }
            else numeric.push_back("Q must be a finite number");
            if (!numeric.empty()) { for (auto &r : numeric) issue(lineNumber, r); continue; }
            f.frequency_hz = fc; f.gain_db = gain; f.q = q;
            parsed.push_back(f);
            continue;
        }
        if (starts_with_ci(line, "preamp")) issue(lineNumber, "malformed Preamp directive");         // :138-142
        else issue(lineNumber, "unsupported directive");
    }
    bool anyEnabled = false;
    for (auto &f : parsed) anyEnabled = anyEnabled || f.enabled;
    if (issues.empty() && preamp == 0 && !anyEnabled)                                                // :145-147
        issue(0, "effective configuration must contain a non-zero preamp or an enabled supported filter");
    if (!issues.empty()) return fail();
    if (preamp_db) *preamp_db = preamp;
    if (n_filters) *n_filters = (int)parsed.size();
    if ((int)parsed.size() > capacity) return set_error(AW_ERR_INVALID_ARGUMENT, "aw_eq_parse: filter capacity too small");
    for (size_t i = 0; i < parsed.size(); ++i) filters[i] = parsed[i];
    return AW_OK;
}
