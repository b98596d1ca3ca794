// A caller of the reference's writer API (AudioFile/OAudioFile.h:9-31) and of IAudioFile::readRaw, compiled against this repo's
// include/ and linked to libhisstools_b200.so (tests/test_audio_writer.py compiles and runs it; no GPU involved).
//   audio_writer_test <out.wav> <out.aifc>
#include "AudioFile/OAudioFile.h"

#include <cmath>
#include <cstdio>
#include <vector>

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    const uint32_t frames = 301;
    std::vector<double> left(frames), right(frames);
    std::vector<float> inter(2 * frames);
    for (uint32_t k = 0; k < frames; k++)
    {
        left[k] = std::sin(0.05 * k) * 0.9;
        right[k] = std::cos(0.031 * k) * 1.2;             // past full scale: wraps in integer formats, as in the reference
        inter[2 * k] = float(left[k]);
        inter[2 * k + 1] = float(right[k]);
    }
    {
        HISSTools::OAudioFile f(argv[1], HISSTools::IAudioFile::kAudioFileWAVE, HISSTools::IAudioFile::kAudioFileInt24, 2, 48000.0);
        if (!f.isOpen() || f.getIsError()) return 3;
        f.writeChannel(right.data(), frames, 1);
        f.seek(0);
        f.writeChannel(left.data(), frames, 0);
        if (f.getFrames() != frames || f.getPosition() != frames || f.getFrameByteCount() != 6) return 4;
    }
    {
        HISSTools::OAudioFile f;
        f.open(argv[2], HISSTools::IAudioFile::kAudioFileAIFF, HISSTools::IAudioFile::kAudioFileFloat32, 2, 44100.0);
        if (!f.isOpen() || f.getFileType() != HISSTools::IAudioFile::kAudioFileAIFC) return 5;
        f.writeInterleaved(inter.data(), frames);
        f.close();
        if (f.isOpen()) return 6;
    }
    // raw frames back: 24-bit little-endian, frame 7, channel 0
    HISSTools::IAudioFile r(argv[1]);
    if (!r.isOpen() || r.getFrames() != frames) return 7;
    std::vector<unsigned char> raw(6 * 10);
    r.seek(7);
    r.readRaw(raw.data(), 10);
    const int32_t v = int32_t(uint32_t(raw[0]) << 8 | uint32_t(raw[1]) << 16 | uint32_t(raw[2]) << 24) >> 8;
    if (v != int32_t(std::round(left[7] * 8388608.0)) || r.getPosition() != 17) return 8;
    printf("ok\n");
    return 0;
}
