// hb_audio_out.cu -- the writing half of the reference's AudioFile component and raw frame reads: host code only.
//
// hb_audio_writer_* produce, byte for byte, the files the reference's OAudioFile writes (AudioFile/OAudioFile.cpp): a 44-byte
// RIFF / RIFX WAVE header (:374-402) or an AIFC header with FVER, COMM (compression tag and Pascal string) and SSND chunks
// (:431-475; a request for AIFF becomes AIFC, :58), sizes patched after every write that extends the file (:478-523), a pad byte
// behind an odd number of data bytes, and the reference's sample conversions (:566-585) -- including that integer samples are
// rounded half away from zero and NOT clipped (the std::min / std::max result is discarded at :574, so out-of-range input wraps),
// that 8-bit WAVE is unsigned and clipped, and that little-endian AIFC still carries the tag "NONE" (:404-417).
//
// The design differs from the reference's (one seek and one 1-8 byte write per sample through an ofstream): a call converts its
// samples into one byte buffer and writes it once; a single-channel write into a multichannel file reads the frames that already
// exist, patches the channel's bytes, zero-fills what lies beyond the end (the reference's resize, :525-547) and writes the range back.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "hb_common.cuh"

using namespace hb;

struct hb_audio_writer
{
    FILE *f = nullptr;
    int file_type = 0;          // BaseAudioFile::FileType
    int pcm = 0;                // BaseAudioFile::PCMFormat
    bool header_be = false, audio_be = false;
    double rate = 0;
    uint32_t channels = 0, frames = 0;
    uint64_t pcm_offset = 0;
    int flags = 0;
};

namespace
{
constexpr int TYPE_AIFF = 1, TYPE_AIFC = 2, TYPE_WAVE = 3;
constexpr int ERR_OPEN = 1 << 2, ERR_WRITE = 1 << 9;
constexpr uint32_t AIFC_VERSION = 0xA2805140u;          // AIFC_CURRENT_SPECIFICATION (BaseAudioFile.h)

const int BITS[6] = {8, 16, 24, 32, 32, 64};
inline uint32_t byte_depth(const hb_audio_writer *w) { return BITS[w->pcm] / 8; }
inline uint64_t frame_bytes(const hb_audio_writer *w) { return uint64_t(w->channels) * byte_depth(w); }

// integer of `n` bytes into p, in the given byte order
inline void put(unsigned char *p, uint64_t v, int n, bool big)
{
    for (int k = 0; k < n; k++) p[big ? n - 1 - k : k] = (unsigned char) (v >> (8 * k));
}

struct Bytes
{
    std::vector<unsigned char> b;
    void tag(const char *t) { b.insert(b.end(), t, t + 4); }
    void u(uint64_t v, int n, bool big) { b.resize(b.size() + n); put(b.data() + b.size() - n, v, n, big); }
};

// 80-bit IEEE extended of a sampling rate (OAudioFile.cpp:271-331: frexp / ldexp construction, truncated mantissa)
void extended80(double x, unsigned char *out)
{
    int sign = 0, expo = 0;
    uint32_t hi = 0, lo = 0;
    if (x < 0) { sign = 0x8000; x = -x; }
    if (x != 0)
    {
        double m = std::frexp(x, &expo);
        if (expo > 16384 || !(m < 1)) { expo = sign | 0x7FFF; }
        else
        {
            expo += 16382;
            if (expo < 0) { m = std::ldexp(m, expo); expo = 0; }
            expo |= sign;
            m = std::ldexp(m, 32);
            double whole = std::floor(m);
            hi = (uint32_t) (uint64_t) whole;
            m = std::ldexp(m - whole, 32);
            lo = (uint32_t) (uint64_t) std::floor(m);
        }
    }
    put(out, (uint64_t) expo, 2, true);
    put(out + 2, hi, 4, true);
    put(out + 6, lo, 4, true);
}

const char *compression_tag(int pcm) { return pcm == 4 ? "fl32" : (pcm == 5 ? "fl64" : "NONE"); }
const char *compression_name(int pcm) { return pcm == 4 ? "32-bit floating point" : (pcm == 5 ? "64-bit floating point" : "not compressed"); }
// bytes a Pascal string takes: count byte + characters, padded to an even total
uint32_t pstring_bytes(const char *s) { const uint32_t n = (uint32_t) strlen(s) + 1; return n + (n & 1); }

bool write_at(hb_audio_writer *w, uint64_t pos, const void *p, size_t n)
{
    if (fseeko(w->f, (off_t) pos, SEEK_SET)) return false;
    return n == 0 || fwrite(p, 1, n, w->f) == n;
}

bool write_header(hb_audio_writer *w)
{
    Bytes h;
    const bool be = w->header_be;
    if (w->file_type == TYPE_WAVE)
    {
        h.tag(be ? "RIFX" : "RIFF"); h.u(36, 4, be); h.tag("WAVE");
        h.tag("fmt "); h.u(16, 4, be);
        h.u(w->pcm >= 4 ? 3 : 1, 2, be);
        h.u(w->channels, 2, be);
        h.u((uint32_t) w->rate, 4, be);
        h.u((uint32_t) (w->rate * double(frame_bytes(w))), 4, be);
        h.u((uint16_t) frame_bytes(w), 2, be);
        h.u(BITS[w->pcm], 2, be);
        h.tag("data"); h.u(0, 4, be);
    }
    else
    {
        const char *name = compression_name(w->pcm);
        const uint32_t ps = pstring_bytes(name);
        h.tag("FORM"); h.u(62 + ps, 4, true); h.tag("AIFC");
        h.tag("FVER"); h.u(4, 4, true); h.u(AIFC_VERSION, 4, true);
        h.tag("COMM"); h.u(22 + ps, 4, true);
        h.u(w->channels, 2, true); h.u(0, 4, true); h.u(BITS[w->pcm], 2, true);
        unsigned char ext[10];
        extended80(w->rate, ext);
        h.b.insert(h.b.end(), ext, ext + 10);
        h.tag(compression_tag(w->pcm));
        const size_t len = strlen(name);
        h.b.push_back((unsigned char) len);
        h.b.insert(h.b.end(), name, name + len);
        if ((len + 1) & 1) h.b.push_back(0);
        h.tag("SSND"); h.u(8, 4, true); h.u(0, 4, true); h.u(0, 4, true);
    }
    w->pcm_offset = h.b.size();
    return write_at(w, 0, h.b.data(), h.b.size());
}

// the size fields after the file has grown to `frames` frames whose data ends at byte `data_end` (OAudioFile.cpp:478-523)
bool patch_sizes(hb_audio_writer *w, uint64_t data_end)
{
    const uint64_t data_bytes = frame_bytes(w) * w->frames;
    const uint64_t padded = data_bytes + (data_bytes & 1);
    const uint64_t header = w->pcm_offset - 8;
    bool ok = true;
    unsigned char v[4];
    if (data_bytes & 1) { const unsigned char z = 0; ok = write_at(w, data_end, &z, 1) && ok; }
    put(v, (uint32_t) (header + padded), 4, w->header_be);
    ok = write_at(w, 4, v, 4) && ok;
    if (w->file_type == TYPE_WAVE)
    {
        put(v, (uint32_t) data_bytes, 4, w->header_be);
        ok = write_at(w, w->pcm_offset - 4, v, 4) && ok;
    }
    else
    {
        put(v, w->frames, 4, true);
        ok = write_at(w, 34, v, 4) && ok;
        put(v, (uint32_t) data_bytes + 8, 4, true);
        ok = write_at(w, w->pcm_offset - 12, v, 4) && ok;
    }
    return ok;
}

uint64_t position_bytes(hb_audio_writer *w) { return (uint64_t) ftello(w->f); }

// one sample in the file's representation (OAudioFile.cpp:566-585, 606-676)
inline void encode(const hb_audio_writer *w, double x, unsigned char *p)
{
    switch (w->pcm)
    {
        case 0:
            if (w->file_type == TYPE_WAVE) { *p = (unsigned char) std::min(std::max(std::round(x * 128.0 + 128.0), 0.0), 255.0); }
            else *p = (unsigned char) (uint32_t) (int64_t) std::round(x * 128.0);
            break;
        case 1: put(p, (uint32_t) (int64_t) std::round(x * 32768.0), 2, w->audio_be); break;
        case 2: put(p, (uint32_t) (int64_t) std::round(x * 8388608.0), 3, w->audio_be); break;
        case 3: put(p, (uint32_t) (int64_t) std::round(x * 2147483648.0), 4, w->audio_be); break;
        case 4: { const float v = (float) x; uint32_t u; memcpy(&u, &v, 4); put(p, u, 4, w->audio_be); break; }
        default: { uint64_t u; memcpy(&u, &x, 8); put(p, u, 8, w->audio_be); break; }
    }
}

template <class T>
bool write_samples(hb_audio_writer *w, const T *in, uint32_t frames, int32_t channel)
{
    const uint32_t bd = byte_depth(w);
    const uint64_t fb = frame_bytes(w);
    const uint64_t start = position_bytes(w);
    std::vector<unsigned char> buf;
    bool ok = true;
    if (channel < 0 || w->channels == 1)
    {
        // every byte of the range is new
        const size_t n = size_t(frames) * (channel < 0 ? w->channels : 1);
        buf.resize(n * bd);
        for (size_t k = 0; k < n; k++) encode(w, (double) in[k], buf.data() + k * bd);
    }
    else
    {
        // one channel of several: keep what the other channels hold, silence where the file ends earlier
        buf.assign(size_t(frames) * fb, 0);
        const uint64_t have_end = w->pcm_offset + fb * w->frames;
        if (start < have_end)
        {
            const size_t n = (size_t) std::min<uint64_t>(have_end - start, buf.size());
            if (fseeko(w->f, (off_t) start, SEEK_SET) || fread(buf.data(), 1, n, w->f) != n) ok = false;
        }
        for (size_t k = 0; k < frames; k++) encode(w, (double) in[k], buf.data() + k * fb + size_t(channel) * bd);
    }
    ok = write_at(w, start, buf.data(), buf.size()) && ok;
    return ok;
}

// after a write that ended at the current position: extend the frame count and the header if the file grew
bool after_write(hb_audio_writer *w)
{
    const uint64_t end = position_bytes(w);
    const uint64_t fb = frame_bytes(w);
    const uint32_t end_frame = fb ? (uint32_t) ((end - w->pcm_offset) / fb) : 0;
    bool ok = true;
    if (end_frame > w->frames)
    {
        w->frames = end_frame;
        ok = patch_sizes(w, end);
        ok = (fseeko(w->f, (off_t) end, SEEK_SET) == 0) && ok;
    }
    return ok;
}
} // namespace

extern "C" int hb_audio_writer_open(hb_audio_writer **out, const char *path, int file_type, int pcm_format, uint32_t channels, double rate, int big_endian)
{
    if (!out || !path) { set_error("hb_audio_writer_open: null argument"); return HB_ERR_BAD_ARG; }
    *out = nullptr;
    if (file_type < TYPE_AIFF || file_type > TYPE_WAVE || pcm_format < 0 || pcm_format > 5 || channels > 65535)
    {
        set_error("hb_audio_writer_open: file type 1..3 (AIFF, AIFC, WAVE), PCM format 0..5, at most 65535 channels");
        return HB_ERR_BAD_ARG;
    }
    hb_audio_writer *w = new hb_audio_writer;
    *out = w;
    w->f = fopen(path, "w+b");
    if (!w->f) { w->flags |= ERR_OPEN; return HB_OK; }              // as the reference: an object that is not open, flag set
    const bool be = big_endian < 0 ? file_type != TYPE_WAVE : big_endian != 0;
    w->file_type = file_type == TYPE_AIFF ? TYPE_AIFC : file_type;
    w->pcm = pcm_format;
    w->header_be = w->file_type == TYPE_WAVE ? be : true;
    w->audio_be = be;
    w->rate = rate;
    w->channels = channels;
    if (!write_header(w)) w->flags |= ERR_WRITE;
    return HB_OK;
}

extern "C" int hb_audio_writer_write(hb_audio_writer *w, const void *in, int in_dtype, uint32_t frames, int32_t channel)
{
    if (!w) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    if (!w->f || !frames) return HB_OK;
    if (!in || (in_dtype != HB_F32 && in_dtype != HB_F64) || channel >= (int32_t) w->channels)
    {
        set_error("hb_audio_writer_write: null input, unknown element type or channel out of range");
        return HB_ERR_BAD_ARG;
    }
    bool ok = in_dtype == HB_F64 ? write_samples(w, (const double *) in, frames, channel) : write_samples(w, (const float *) in, frames, channel);
    ok = after_write(w) && ok;
    (void) ok;                                                        // (the reference drops the result as well, OAudioFile.cpp:679-682)
    return HB_OK;
}

extern "C" int hb_audio_writer_write_raw(hb_audio_writer *w, const void *raw, uint32_t frames)
{
    if (!w) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    if (!w->f || !frames) return HB_OK;
    if (!raw) { set_error("hb_audio_writer_write_raw: null input"); return HB_ERR_BAD_ARG; }
    write_at(w, position_bytes(w), raw, size_t(frames) * frame_bytes(w));
    after_write(w);
    return HB_OK;
}

extern "C" int hb_audio_writer_seek(hb_audio_writer *w, uint32_t frame)
{
    if (!w) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    if (w->f && w->pcm_offset) fseeko(w->f, (off_t) (w->pcm_offset + frame_bytes(w) * frame), SEEK_SET);
    return HB_OK;
}

extern "C" uint32_t hb_audio_writer_position(hb_audio_writer *w)
{
    if (!w || !w->f || !w->pcm_offset || !frame_bytes(w)) return 0;
    return (uint32_t) ((position_bytes(w) - w->pcm_offset) / frame_bytes(w));
}

extern "C" int hb_audio_writer_info(const hb_audio_writer *w, hb_audio_info *info, int *is_open)
{
    if (!w || !info) { set_error("null argument"); return HB_ERR_BAD_ARG; }
    memset(info, 0, sizeof(*info));
    info->file_type = w->file_type;
    info->pcm_format = w->pcm;
    info->header_big_endian = w->header_be;
    info->audio_big_endian = w->audio_be;
    info->channels = w->channels;
    info->frames = w->frames;
    info->sampling_rate = w->rate;
    info->pcm_offset = w->pcm_offset;
    info->error_flags = w->flags;
    if (is_open) *is_open = w->f != nullptr;
    return HB_OK;
}

extern "C" void hb_audio_writer_close(hb_audio_writer *w)
{
    if (!w) return;
    if (w->f) fclose(w->f);
    delete w;
}

// IAudioFile::seek(first_frame) then readRaw(out, frames) (IAudioFile.cpp:74-88): the frames as stored in the file
extern "C" int hb_audio_read_raw(const char *path, uint32_t first_frame, uint32_t frames, void *out)
{
    hb_audio_info info;
    int rc = hb_audio_probe(path, &info);
    if (rc) return rc;
    if (info.error_flags) { set_error("hb_audio_read_raw: %s is not readable (error flags %d)", path, info.error_flags); return HB_ERR_BAD_ARG; }
    if (!frames) return HB_OK;
    if (!out) { set_error("hb_audio_read_raw: null output"); return HB_ERR_BAD_ARG; }
    const uint64_t fb = uint64_t(info.channels) * (BITS[info.pcm_format] / 8);
    FILE *f = fopen(path, "rb");
    if (!f) { set_error("hb_audio_read_raw: cannot open %s", path); return HB_ERR_BAD_ARG; }
    size_t got = 0;
    if (!fseeko(f, (off_t) (info.pcm_offset + fb * first_frame), SEEK_SET)) got = fread(out, 1, size_t(fb) * frames, f);
    fclose(f);
    // a read past the end leaves the rest of `out` as it was (the reference's ifstream read fails the same way)
    (void) got;
    return HB_OK;
}
