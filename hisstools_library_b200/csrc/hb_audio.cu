// hb_audio.cu -- impulse-response ingestion from WAV / AIFF / AIFC files: the hb_audio_* entry points of
// include/hisstools_b200.h (the step before Convolver::set in real use).
//
// Replaces the reading half of the reference's AudioFile component: header parsing (IAudioFile.cpp:375-609
// parseHeader / parseAIFFHeader / parseWaveHeader with the chunk helpers :242-287) runs on the host -- it is a few
// dozen bytes of control data -- and sample decoding (readAudio :613-689 with the conversions :137-239: 8 / 16 / 24 /
// 32-bit integers of either byte order, 32 / 64-bit floats, de-interleaving of one channel) runs on the GPU: the raw
// PCM bytes are uploaded once and k_audio_decode writes planar rows that hb_conv_set_ir_dev consumes in place, so a
// multichannel IR file never exists as a host float array.
#include "hb_common.cuh"

#include <cmath>
#include <cstdio>
#include <mutex>
#include <vector>

using namespace hb;

namespace
{
// BaseAudioFile.h:19-60
enum { TYPE_NONE = 0, TYPE_AIFF = 1, TYPE_AIFC = 2, TYPE_WAVE = 3 };
enum { PCM_INT8 = 0, PCM_INT16, PCM_INT24, PCM_INT32, PCM_FLOAT32, PCM_FLOAT64 };
enum { ERR_FILE_COULDNT_OPEN = 1 << 2, ERR_FILE_BAD_FORMAT = 1 << 3, ERR_FILE_UNKNOWN_FORMAT = 1 << 4, ERR_FILE_UNSUPPORTED_PCM_FORMAT = 1 << 5,
       ERR_AIFC_WRONG_VERSION = 1 << 6, ERR_AIFC_UNSUPPORTED_FORMAT = 1 << 7, ERR_WAVE_UNSUPPORTED_FORMAT = 1 << 8 };

const uint32_t kBitDepth[6] = {8, 16, 24, 32, 32, 64};

struct Reader
{
    FILE *f = nullptr;
    bool big = false;                                   // header byte order
    ~Reader() { if (f) fclose(f); }
    bool read(void *dst, size_t n) { return fread(dst, 1, n, f) == n; }
    long pos() { return ftell(f); }
    bool seek(long p) { return fseek(f, p, SEEK_SET) == 0 && ftell(f) == p; }
    bool advance(long off) { return seek(pos() + off); }
    uint32_t u32(const unsigned char *b) const { return big ? (uint32_t(b[0]) << 24) | (b[1] << 16) | (b[2] << 8) | b[3] : (uint32_t(b[3]) << 24) | (b[2] << 16) | (b[1] << 8) | b[0]; }
    uint32_t u16(const unsigned char *b) const { return big ? (b[0] << 8) | b[1] : (b[1] << 8) | b[0]; }
    // chunk header: 4-byte tag + 32-bit size (IAudioFile.cpp:247-257)
    bool chunk_header(char *tag, uint32_t &size)
    {
        unsigned char h[8];
        if (!read(h, 8)) return false;
        memcpy(tag, h, 4);
        size = u32(h + 4);
        return true;
    }
    // read the first `want` bytes of a chunk of `size` bytes and skip the rest, padded to even (:273-284)
    bool chunk(void *dst, uint32_t want, uint32_t size)
    {
        if (want && (want > size || !read(dst, want))) return false;
        return advance(long(size + (size & 1u)) - long(want));
    }
    // search forward for a chunk (:259-271)
    bool find(const char *tag, uint32_t &size)
    {
        char t[4];
        while (chunk_header(t, size))
        {
            if (!memcmp(t, tag, 4)) return true;
            if (!advance(long(size + (size & 1u)))) return false;
        }
        return false;
    }
};

// 80-bit IEEE extended, big-endian (IAudioFile.cpp:188-216)
double extended_to_double(const unsigned char *b)
{
    const uint32_t se = (b[0] << 8) | b[1];
    const bool sign = se & 0x8000;
    int32_t exponent = se & 0x7FFF;
    const uint32_t hi = (uint32_t(b[2]) << 24) | (b[3] << 16) | (b[4] << 8) | b[5];
    const uint32_t lo = (uint32_t(b[6]) << 24) | (b[7] << 16) | (b[8] << 8) | b[9];
    if (!exponent && !hi && !lo) return 0.0;
    if (exponent == 0x7FFF) return HUGE_VAL;
    exponent -= 16383;
    double v = ldexp(double(hi), exponent - 31) + ldexp(double(lo), exponent - 63);
    return sign ? -v : v;
}

// bit depth + number format -> PCM format (IAudioFile.cpp:289-322); false = unsupported
bool pcm_format(uint32_t bits, bool is_float, int32_t &fmt)
{
    if (!is_float)
    {
        if (bits == 8) { fmt = PCM_INT8; return true; }
        if (bits == 16) { fmt = PCM_INT16; return true; }
        if (bits == 24) { fmt = PCM_INT24; return true; }
        if (bits == 32) { fmt = PCM_INT32; return true; }
    }
    else
    {
        if (bits == 32) { fmt = PCM_FLOAT32; return true; }
        if (bits == 64) { fmt = PCM_FLOAT64; return true; }
    }
    return false;
}

void parse_aiff(Reader &r, bool aifc, hb_audio_info *in)
{
    // IAudioFile.cpp:407-548
    const uint32_t TAG_VERSION = 1, TAG_COMMON = 2, TAG_AUDIO = 4;
    uint32_t valid = TAG_COMMON | TAG_AUDIO, seen = 0;
    r.big = true;
    in->header_big_endian = 1;
    if (aifc) { in->file_type = TYPE_AIFC; valid |= TAG_VERSION; }
    char tag[4];
    uint32_t size;
    while (r.chunk_header(tag, size))
    {
        unsigned char c[22] = {};
        if (!memcmp(tag, "FVER", 4))
        {
            seen |= TAG_VERSION;
            if (!r.chunk(c, 4, size)) { in->error_flags |= ERR_FILE_BAD_FORMAT; return; }
            if (r.u32(c) != 0xA2805140u) { in->error_flags |= ERR_AIFC_WRONG_VERSION; return; }
        }
        else if (!memcmp(tag, "COMM", 4))
        {
            seen |= TAG_COMMON;
            if (!r.chunk(c, size > 22 ? 22 : (size < 18 ? 18 : size), size)) { in->error_flags |= ERR_FILE_BAD_FORMAT; return; }
            in->channels = r.u16(c);
            in->frames = r.u32(c + 2);
            uint32_t bits = r.u16(c + 6);
            in->sampling_rate = extended_to_double(c + 8);
            bool is_float = false;
            in->audio_big_endian = 1;
            if (!in->frames) seen |= TAG_AUDIO;
            if (aifc)
            {
                // compression tag (:354-389).  "fl64" is given a bit depth of 32 there, i.e. read as 32-bit floats: kept.
                const char *t = (const char *) c + 18;
                if (!memcmp(t, "NONE", 4)) {}
                else if (!memcmp(t, "twos", 4)) bits = 16;
                else if (!memcmp(t, "sowt", 4)) { bits = 16; in->audio_big_endian = 0; }
                else if (!memcmp(t, "fl32", 4) || !memcmp(t, "FL32", 4) || !memcmp(t, "fl64", 4) || !memcmp(t, "FL64", 4)) { bits = 32; is_float = true; }
                else { in->error_flags |= ERR_AIFC_UNSUPPORTED_FORMAT; return; }
            }
            else
                in->file_type = TYPE_AIFF;
            if (!pcm_format(bits, is_float, in->pcm_format)) { in->error_flags |= ERR_FILE_UNSUPPORTED_PCM_FORMAT; return; }
        }
        else if (!memcmp(tag, "SSND", 4))
        {
            seen |= TAG_AUDIO;
            in->pcm_offset = uint64_t(r.pos()) + 8;
            if (!r.chunk(c, 4, size)) { in->error_flags |= ERR_FILE_BAD_FORMAT; return; }
            in->pcm_offset += r.u32(c);
        }
        else if (!r.chunk(nullptr, 0, size)) { in->error_flags |= ERR_FILE_BAD_FORMAT; return; }
    }
    if (~seen & valid) in->error_flags |= ERR_FILE_BAD_FORMAT;
}

void parse_wave(Reader &r, bool rifx, hb_audio_info *in)
{
    // IAudioFile.cpp:550-609
    r.big = rifx;
    in->header_big_endian = in->audio_big_endian = rifx ? 1 : 0;
    unsigned char c[16];
    uint32_t size;
    if (!(r.find("fmt ", size) && r.chunk(c, 16, size))) { in->error_flags |= ERR_FILE_BAD_FORMAT; return; }
    const uint32_t tag = r.u16(c);
    if (tag != 1 && tag != 3) { in->error_flags |= ERR_WAVE_UNSUPPORTED_FORMAT; return; }
    in->channels = r.u16(c + 2);
    in->sampling_rate = r.u32(c + 4);
    if (!pcm_format(r.u16(c + 14), tag == 3, in->pcm_format)) { in->error_flags |= ERR_FILE_UNSUPPORTED_PCM_FORMAT; return; }
    if (!r.find("data", size)) { in->error_flags |= ERR_FILE_BAD_FORMAT; return; }
    const uint32_t frame_bytes = in->channels * (kBitDepth[in->pcm_format] / 8);
    in->frames = frame_bytes ? size / frame_bytes : 0;
    in->pcm_offset = (uint64_t) r.pos();
    in->file_type = TYPE_WAVE;
}

// ---- device side ----------------------------------------------------------------------------------------
// One sample: bytes -> value (IAudioFile.cpp:137-239, 640-682).  Integers go through a 32-bit word with the sample in
// its top bits, times 2^-31; WAVE 8-bit is unsigned, (v - 128) / 128.
template <class T>
__device__ __forceinline__ T decode_sample(const unsigned char *__restrict__ p, int fmt, int big, int wave)
{
    switch (fmt)
    {
        case PCM_INT8:
            if (wave) return (T(p[0]) - T(128)) / T(128);
            return T(int32_t(uint32_t(p[0]) << 24)) * T(4.656612873077392578125e-10);
        case PCM_INT16:
        {
            const uint32_t v = big ? (uint32_t(p[0]) << 8) | p[1] : (uint32_t(p[1]) << 8) | p[0];
            return T(int32_t(v << 16)) * T(4.656612873077392578125e-10);
        }
        case PCM_INT24:
        {
            const uint32_t v = big ? (uint32_t(p[0]) << 16) | (uint32_t(p[1]) << 8) | p[2] : (uint32_t(p[2]) << 16) | (uint32_t(p[1]) << 8) | p[0];
            return T(int32_t(v << 8)) * T(4.656612873077392578125e-10);
        }
        case PCM_INT32:
        case PCM_FLOAT32:
        {
            const uint32_t v = big ? (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]
                                   : (uint32_t(p[3]) << 24) | (uint32_t(p[2]) << 16) | (uint32_t(p[1]) << 8) | p[0];
            if (fmt == PCM_INT32) return T(int32_t(v)) * T(4.656612873077392578125e-10);
            return T(__uint_as_float(v));
        }
        default:
        {
            unsigned long long v = 0;
            for (int k = 0; k < 8; k++) v = (v << 8) | p[big ? k : 7 - k];
            return T(__longlong_as_double((long long) v));
        }
    }
}

// raw interleaved frames -> rows.  planar = 1: out[c][ld] for channels c0 .. c0 + nch - 1 (readChannel per row);
// planar = 0: out[frame * nch + c] (readInterleaved).  grid.y = channel of the request.
template <class T>
__global__ void k_audio_decode(const unsigned char *__restrict__ raw, uint32_t file_channels, uint32_t byte_depth, int fmt, int big, int wave,
                               uint32_t c0, uint32_t nch, uint64_t frames, T *__restrict__ out, uint64_t ld, int planar)
{
    const uint32_t c = blockIdx.y;
    const uint64_t frame_bytes = uint64_t(file_channels) * byte_depth;
    for (uint64_t f = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; f < frames; f += uint64_t(gridDim.x) * blockDim.x)
    {
        const T v = decode_sample<T>(raw + f * frame_bytes + uint64_t(c0 + c) * byte_depth, fmt, big, wave);
        if (planar) out[uint64_t(c) * ld + f] = v;
        else out[f * nch + c] = v;
    }
}

int check_info(const hb_audio_info *in)
{
    if (!in || in->error_flags || in->file_type == TYPE_NONE || in->pcm_format < 0 || in->pcm_format > PCM_FLOAT64 || !in->channels)
    {
        set_error("hb_audio: the file description is not that of a readable file (error flags 0x%x)", in ? in->error_flags : -1);
        return HB_ERR_BAD_ARG;
    }
    return HB_OK;
}

template <class T>
int decode_launch(const hb_audio_info *in, const void *d_raw, uint64_t frames, int32_t channel, uint32_t nch, T *d_out, uint64_t ld, int planar, cudaStream_t st)
{
    if (!frames || !nch) return HB_OK;
    const dim3 grid((unsigned) std::min<uint64_t>((frames + 255) / 256, 2048), nch);
    k_audio_decode<T><<<grid, 256, 0, st>>>((const unsigned char *) d_raw, in->channels, kBitDepth[in->pcm_format] / 8, in->pcm_format, in->audio_big_endian,
                                            in->file_type == TYPE_WAVE, channel < 0 ? 0u : (uint32_t) channel, nch, frames, d_out, ld, planar);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

// raw PCM bytes of frames [first, first + frames) into device memory (clamped reads leave the remainder zero, like a
// short fread into the reference's work buffer would leave stale data: callers are expected to stay inside the file)
int upload_frames(const char *path, const hb_audio_info *in, uint64_t first, uint64_t frames, DevBuf &d_raw, cudaStream_t st)
{
    const uint64_t frame_bytes = uint64_t(in->channels) * (kBitDepth[in->pcm_format] / 8);
    const uint64_t bytes = frames * frame_bytes;
    int rc = d_raw.ensure(std::max<uint64_t>(bytes, 16));
    if (rc) return rc;
    PinnedBuf h;
    if ((rc = h.ensure(std::max<uint64_t>(bytes, 16)))) return rc;
    memset(h.p, 0, bytes);
    FILE *f = fopen(path, "rb");
    if (!f) { h.release(); set_error("hb_audio: cannot open %s", path); return HB_ERR_BAD_ARG; }
    if (fseek(f, (long) (in->pcm_offset + first * frame_bytes), SEEK_SET) == 0) (void) !fread(h.p, 1, bytes, f);
    fclose(f);
    cudaError_t e = cudaMemcpyAsync(d_raw.p, h.p, bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    h.release();
    if (e != cudaSuccess) { set_error("hb_audio: upload -> %s", cudaGetErrorString(e)); return HB_ERR_CUDA; }
    return HB_OK;
}
} // namespace

extern "C" int hb_audio_probe(const char *path, hb_audio_info *info)
{
    if (!path || !info) { set_error("hb_audio_probe: null argument"); return HB_ERR_BAD_ARG; }
    memset(info, 0, sizeof(*info));                     // BaseAudioFile::close() defaults (BaseAudioFile.cpp:17-28)
    Reader r;
    r.f = fopen(path, "rb");
    if (!r.f) { info->error_flags = ERR_FILE_COULDNT_OPEN; return HB_OK; }
    unsigned char head[12] = {};
    if (!r.read(head, 12)) { info->error_flags |= ERR_FILE_BAD_FORMAT; return HB_OK; }        // IAudioFile.cpp:381-385
    const bool form = !memcmp(head, "FORM", 4), riff = !memcmp(head, "RIFF", 4), rifx = !memcmp(head, "RIFX", 4);
    if (form && (!memcmp(head + 8, "AIFF", 4) || !memcmp(head + 8, "AIFC", 4))) parse_aiff(r, !memcmp(head + 8, "AIFC", 4), info);
    else if ((riff || rifx) && !memcmp(head + 8, "WAVE", 4)) parse_wave(r, rifx, info);
    else info->error_flags |= ERR_FILE_UNKNOWN_FORMAT;
    return HB_OK;
}

extern "C" int hb_audio_decode_dev(const hb_audio_info *info, const void *d_raw, uint64_t frames, int32_t channel, void *d_out, uint64_t ld,
                                   int out_dtype, int device, void *stream)
{
    int rc = check_info(info);
    if (rc) return rc;
    if ((out_dtype != HB_F32 && out_dtype != HB_F64) || (frames && (!d_raw || !d_out)) || channel >= (int32_t) info->channels)
    {
        set_error("hb_audio_decode_dev: bad argument");
        return HB_ERR_BAD_ARG;
    }
    if ((rc = use_device(device))) return rc;
    const uint32_t nch = channel < 0 ? info->channels : 1;
    cudaStream_t st = (cudaStream_t) stream;
    return out_dtype == HB_F64 ? decode_launch<double>(info, d_raw, frames, channel, nch, (double *) d_out, ld, 1, st)
                               : decode_launch<float>(info, d_raw, frames, channel, nch, (float *) d_out, ld, 1, st);
}

extern "C" int hb_audio_read(const char *path, uint32_t first_frame, uint32_t frames, int32_t channel, void *out, int out_dtype, int device)
{
    if (!path || (frames && !out) || (out_dtype != HB_F32 && out_dtype != HB_F64)) { set_error("hb_audio_read: bad argument"); return HB_ERR_BAD_ARG; }
    hb_audio_info info;
    int rc = hb_audio_probe(path, &info);
    if (rc) return rc;
    if ((rc = check_info(&info))) return rc;
    if (channel >= (int32_t) info.channels) { set_error("hb_audio_read: channel %d of %u", channel, info.channels); return HB_ERR_BAD_ARG; }
    if (!frames) return HB_OK;
    if ((rc = use_device(device))) return rc;
    const uint32_t nch = channel < 0 ? info.channels : 1;
    const size_t es = dtype_size(out_dtype);
    DevBuf d_raw, d_out;
    cudaStream_t st = nullptr;                          // the legacy default stream: the call is synchronous anyway
    if ((rc = upload_frames(path, &info, first_frame, frames, d_raw, st)) || (rc = d_out.ensure(size_t(frames) * nch * es)))
    {
        d_raw.release(); d_out.release();
        return rc;
    }
    // readInterleaved keeps the file's frame order; readChannel picks one channel (IAudioFile.cpp:96-115)
    rc = out_dtype == HB_F64 ? decode_launch<double>(&info, d_raw.p, frames, channel, nch, (double *) d_out.p, frames, 0, st)
                             : decode_launch<float>(&info, d_raw.p, frames, channel, nch, (float *) d_out.p, frames, 0, st);
    if (rc == HB_OK)
    {
        cudaError_t e = cudaMemcpy(out, d_out.p, size_t(frames) * nch * es, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("hb_audio_read: download -> %s", cudaGetErrorString(e)); rc = HB_ERR_CUDA; }
    }
    d_raw.release(); d_out.release();
    return rc;
}

extern "C" int hb_conv_set_ir_file(hb_conv *c, uint32_t group, uint32_t in, uint32_t out, const char *path, uint32_t channel, int device)
{
    if (!c || !path) { set_error("hb_conv_set_ir_file: null argument"); return HB_ERR_BAD_ARG; }
    hb_audio_info info;
    int rc = hb_audio_probe(path, &info);
    if (rc) return rc;
    if ((rc = check_info(&info))) return rc;
    if (channel >= info.channels) { set_error("hb_conv_set_ir_file: channel %u of %u", channel, info.channels); return HB_ERR_BAD_ARG; }
    if ((rc = use_device(device))) return rc;
    // the engine's element type decides the decode target; 0 = float engine, 1 = double engine
    const int dtype = hb_conv_dtype(c);
    const size_t es = dtype_size(dtype);
    DevBuf d_raw, d_ir;
    cudaStream_t st = nullptr;
    if ((rc = upload_frames(path, &info, 0, info.frames, d_raw, st)) || (rc = d_ir.ensure(std::max<size_t>(size_t(info.frames) * es, 16))))
    {
        d_raw.release(); d_ir.release();
        return rc;
    }
    rc = dtype == HB_F64 ? decode_launch<double>(&info, d_raw.p, info.frames, (int32_t) channel, 1, (double *) d_ir.p, info.frames, 1, st)
                         : decode_launch<float>(&info, d_raw.p, info.frames, (int32_t) channel, 1, (float *) d_ir.p, info.frames, 1, st);
    if (rc == HB_OK && cudaStreamSynchronize(st) != cudaSuccess) { set_error("hb_conv_set_ir_file: decode failed"); rc = HB_ERR_CUDA; }
    if (rc == HB_OK) rc = hb_conv_set_ir_dev(c, group, in, out, d_ir.p, info.frames);
    cudaDeviceSynchronize();                            // the spectra are written from d_ir on the engine's stream
    d_raw.release(); d_ir.release();
    return rc;
}
