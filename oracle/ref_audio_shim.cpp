// oracle/ref_audio_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" window onto the reference's audio-file reader and writer (AudioFile/IAudioFile.h:30-54,
// AudioFile/OAudioFile.h:14-31), compiled in place from $REF_ROOT by oracle/Makefile into
// oracle/_ref/libhisstools_ref_audio.so.  The writer only serves to make test files (tests/golden/make_golden.py);
// the reader is the parity target of hb_audio_* (IR ingestion).

#include "AudioFile/IAudioFile.h"
#include "AudioFile/OAudioFile.h"

#include <cstdint>

#define SHIM extern "C" __attribute__((visibility("default")))

using namespace HISSTools;

// type: 1 AIFF, 2 AIFC, 3 WAVE; pcm: 0 int8 .. 3 int32, 4 float32, 5 float64; big_endian: -1 = the format's default
SHIM int ref_audio_write(const char *path, int type, int pcm, int channels, double rate, int big_endian, const double *interleaved, uint32_t frames)
{
    OAudioFile f;
    if (big_endian < 0)
        f.open(path, static_cast<BaseAudioFile::FileType>(type), static_cast<BaseAudioFile::PCMFormat>(pcm), (uint16_t) channels, rate);
    else
        f.open(path, static_cast<BaseAudioFile::FileType>(type), static_cast<BaseAudioFile::PCMFormat>(pcm), (uint16_t) channels, rate,
               big_endian ? BaseAudioFile::kAudioFileBigEndian : BaseAudioFile::kAudioFileLittleEndian);
    if (!f.isOpen()) return -1;
    f.writeInterleaved(interleaved, frames);
    int flags = f.getErrorFlags();
    f.close();
    return flags;
}

struct ref_audio_info
{
    int32_t file_type, pcm_format, header_big_endian, audio_big_endian;
    uint32_t channels, frames;
    double sampling_rate;
    int32_t error_flags, is_open;
};

SHIM void ref_audio_probe(const char *path, ref_audio_info *info)
{
    IAudioFile f(path);
    info->file_type = f.getFileType();
    info->pcm_format = f.getPCMFormat();
    info->header_big_endian = f.getHeaderEndianness() == BaseAudioFile::kAudioFileBigEndian;
    info->audio_big_endian = f.getAudioEndianness() == BaseAudioFile::kAudioFileBigEndian;
    info->channels = f.getChannels();
    info->frames = f.getFrames();
    info->sampling_rate = f.getSamplingRate();
    info->error_flags = f.getErrorFlags();
    info->is_open = f.isOpen();
}

// seek(first) then readChannel (channel >= 0) or readInterleaved (channel < 0)
template <class T>
static int audio_read(const char *path, uint32_t first, uint32_t frames, int channel, T *out)
{
    IAudioFile f(path);
    if (!f.isOpen() || f.getIsError()) return f.getErrorFlags() ? f.getErrorFlags() : -1;
    f.seek(first);
    if (channel < 0) f.readInterleaved(out, frames);
    else f.readChannel(out, frames, (uint16_t) channel);
    return 0;
}

SHIM int ref_audio_read_f32(const char *path, uint32_t first, uint32_t frames, int channel, float *out) { return audio_read<float>(path, first, frames, channel, out); }
SHIM int ref_audio_read_f64(const char *path, uint32_t first, uint32_t frames, int channel, double *out) { return audio_read<double>(path, first, frames, channel, out); }

// ---- the writer as a session (tests/test_audio_writer.py: files written call by call by the reference and by hb_audio_writer_*) ----
SHIM void *ref_oaudio_open(const char *path, int type, int pcm, int channels, double rate, int big_endian)
{
    OAudioFile *f = new OAudioFile;
    if (big_endian < 0)
        f->open(path, static_cast<BaseAudioFile::FileType>(type), static_cast<BaseAudioFile::PCMFormat>(pcm), (uint16_t) channels, rate);
    else
        f->open(path, static_cast<BaseAudioFile::FileType>(type), static_cast<BaseAudioFile::PCMFormat>(pcm), (uint16_t) channels, rate,
                big_endian ? BaseAudioFile::kAudioFileBigEndian : BaseAudioFile::kAudioFileLittleEndian);
    return f;
}
SHIM void ref_oaudio_write_f64(void *h, const double *in, uint32_t frames, int channel)
{
    OAudioFile *f = static_cast<OAudioFile *>(h);
    if (channel < 0) f->writeInterleaved(in, frames); else f->writeChannel(in, frames, (uint16_t) channel);
}
SHIM void ref_oaudio_write_f32(void *h, const float *in, uint32_t frames, int channel)
{
    OAudioFile *f = static_cast<OAudioFile *>(h);
    if (channel < 0) f->writeInterleaved(in, frames); else f->writeChannel(in, frames, (uint16_t) channel);
}
SHIM void ref_oaudio_write_raw(void *h, const char *in, uint32_t frames) { static_cast<OAudioFile *>(h)->writeRaw(in, frames); }
SHIM void ref_oaudio_seek(void *h, uint32_t frame) { static_cast<OAudioFile *>(h)->seek(frame); }
SHIM uint32_t ref_oaudio_position(void *h) { return static_cast<OAudioFile *>(h)->getPosition(); }
SHIM uint32_t ref_oaudio_frames(void *h) { return static_cast<OAudioFile *>(h)->getFrames(); }
SHIM int ref_oaudio_flags(void *h) { return static_cast<OAudioFile *>(h)->getErrorFlags(); }
SHIM int ref_oaudio_is_open(void *h) { return static_cast<OAudioFile *>(h)->isOpen(); }
SHIM int ref_oaudio_file_type(void *h) { return static_cast<OAudioFile *>(h)->getFileType(); }
SHIM void ref_oaudio_close(void *h) { delete static_cast<OAudioFile *>(h); }

// seek(first) then readRaw: the frames as stored
SHIM int ref_audio_read_raw(const char *path, uint32_t first, uint32_t frames, void *out)
{
    IAudioFile f(path);
    if (!f.isOpen() || f.getIsError()) return f.getErrorFlags() ? f.getErrorFlags() : -1;
    f.seek(first);
    f.readRaw(out, frames);
    return 0;
}
