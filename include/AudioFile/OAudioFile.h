// AudioFile/OAudioFile.h -- B200 drop-in for the writing half of the reference's AudioFile component
// (AudioFile/OAudioFile.h:9-31 + the BaseAudioFile getters, BaseAudioFile.h:64-90): WAVE (RIFF / RIFX) and AIFC files, byte for
// byte what the reference writes.  Header-only; forwards to hb_audio_writer_* of hisstools_b200.h (host code of the library).
#ifndef HISSTOOLS_B200_OAUDIOFILE_H
#define HISSTOOLS_B200_OAUDIOFILE_H

#include <cstdint>
#include <stdexcept>
#include <string>

#include "IAudioFile.h"

namespace HISSTools
{
    class OAudioFile
    {
    public:

        typedef IAudioFile::FrameCount FrameCount;
        typedef IAudioFile::ByteCount ByteCount;
        typedef IAudioFile::FileType FileType;
        typedef IAudioFile::PCMFormat PCMFormat;
        typedef IAudioFile::Endianness Endianness;
        typedef IAudioFile::NumberFormat NumberFormat;
        typedef IAudioFile::Error Error;

        OAudioFile() : mHandle(nullptr) {}
        OAudioFile(const std::string& path, FileType type, PCMFormat format, uint16_t channels, double sr) : mHandle(nullptr) { open(path, type, format, channels, sr); }
        OAudioFile(const std::string& path, FileType type, PCMFormat format, uint16_t channels, double sr, Endianness e) : mHandle(nullptr) { open(path, type, format, channels, sr, e); }
        ~OAudioFile() { close(); }
        OAudioFile(const OAudioFile&) = delete;
        OAudioFile& operator=(const OAudioFile&) = delete;

        void open(const std::string& path, FileType type, PCMFormat format, uint16_t channels, double sr) { openInternal(path, type, format, channels, sr, -1); }
        void open(const std::string& path, FileType type, PCMFormat format, uint16_t channels, double sr, Endianness e)
        {
            openInternal(path, type, format, channels, sr, e == IAudioFile::kAudioFileBigEndian ? 1 : 0);
        }
        void close() { if (mHandle) hb_audio_writer_close(mHandle); mHandle = nullptr; }
        bool isOpen() { int open = 0; hb_audio_info i; return mHandle && hb_audio_writer_info(mHandle, &i, &open) == HB_OK && open; }
        void seek(FrameCount position = 0) { if (mHandle) hb_audio_writer_seek(mHandle, position); }
        FrameCount getPosition() { return mHandle ? hb_audio_writer_position(mHandle) : 0; }

        void writeInterleaved(const double* input, FrameCount numFrames) { write(input, HB_F64, numFrames, -1); }
        void writeInterleaved(const float* input, FrameCount numFrames) { write(input, HB_F32, numFrames, -1); }
        void writeChannel(const double* input, FrameCount numFrames, uint16_t channel) { write(input, HB_F64, numFrames, channel); }
        void writeChannel(const float* input, FrameCount numFrames, uint16_t channel) { write(input, HB_F32, numFrames, channel); }
        void writeRaw(const char *input, FrameCount numFrames) { if (mHandle && hb_audio_writer_write_raw(mHandle, input, numFrames) < 0) throw std::runtime_error(hb_last_error()); }

        FileType getFileType() const { return static_cast<FileType>(info().file_type); }
        PCMFormat getPCMFormat() const { return static_cast<PCMFormat>(info().pcm_format); }
        Endianness getHeaderEndianness() const { return info().header_big_endian ? IAudioFile::kAudioFileBigEndian : IAudioFile::kAudioFileLittleEndian; }
        Endianness getAudioEndianness() const { return info().audio_big_endian ? IAudioFile::kAudioFileBigEndian : IAudioFile::kAudioFileLittleEndian; }
        double getSamplingRate() const { return info().sampling_rate; }
        uint16_t getChannels() const { return static_cast<uint16_t>(info().channels); }
        FrameCount getFrames() const { return info().frames; }
        uint16_t getBitDepth() const { static const uint16_t bits[6] = {8, 16, 24, 32, 32, 64}; return bits[info().pcm_format]; }
        uint16_t getByteDepth() const { return getBitDepth() / 8; }
        ByteCount getFrameByteCount() const { return ByteCount(getChannels()) * getByteDepth(); }
        NumberFormat getNumberFormat() const { return info().pcm_format >= IAudioFile::kAudioFileFloat32 ? IAudioFile::kAudioFileFloat : IAudioFile::kAudioFileInt; }
        int getErrorFlags() const { return info().error_flags; }
        bool getIsError() const { return info().error_flags != IAudioFile::ERR_NONE; }

    private:

        void openInternal(const std::string& path, FileType type, PCMFormat format, uint16_t channels, double sr, int bigEndian)
        {
            close();
            if (hb_audio_writer_open(&mHandle, path.c_str(), type, format, channels, sr, bigEndian) < 0) throw std::runtime_error(hb_last_error());
        }
        void write(const void *input, int dtype, FrameCount numFrames, int32_t channel)
        {
            if (mHandle && hb_audio_writer_write(mHandle, input, dtype, numFrames, channel) < 0) throw std::runtime_error(hb_last_error());
        }
        hb_audio_info info() const
        {
            hb_audio_info i = hb_audio_info();
            if (mHandle) hb_audio_writer_info(mHandle, &i, nullptr);
            return i;
        }

        hb_audio_writer *mHandle;
    };
}

#endif
