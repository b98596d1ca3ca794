// AudioFile/IAudioFile.h -- B200 drop-in for the reading half of the reference's AudioFile component
// (AudioFile/IAudioFile.h:30-54 + the BaseAudioFile getters, BaseAudioFile.h:64-90): WAV / AIFF / AIFC files
// as the source of impulse responses for Convolver::set.  Header-only; forwards to hb_audio_* of hisstools_b200.h
// (header parsing on the host, PCM decoding on the GPU).  The writer is AudioFile/OAudioFile.h.
#ifndef HISSTOOLS_B200_IAUDIOFILE_H
#define HISSTOOLS_B200_IAUDIOFILE_H

#include <cstdint>
#include <stdexcept>
#include <string>

#include "../hisstools_b200.h"

namespace HISSTools
{
    class IAudioFile
    {
    public:

        typedef uint32_t FrameCount;
        typedef uintptr_t ByteCount;

        enum FileType { kAudioFileNone, kAudioFileAIFF, kAudioFileAIFC, kAudioFileWAVE };
        enum PCMFormat { kAudioFileInt8, kAudioFileInt16, kAudioFileInt24, kAudioFileInt32, kAudioFileFloat32, kAudioFileFloat64 };
        enum Endianness { kAudioFileLittleEndian, kAudioFileBigEndian };
        enum NumberFormat { kAudioFileInt, kAudioFileFloat };
        enum Error
        {
            ERR_NONE = 0, ERR_MEM_COULD_NOT_ALLOCATE = 1 << 0, ERR_FILE_ERROR = 1 << 1, ERR_FILE_COULDNT_OPEN = 1 << 2,
            ERR_FILE_BAD_FORMAT = 1 << 3, ERR_FILE_UNKNOWN_FORMAT = 1 << 4, ERR_FILE_UNSUPPORTED_PCM_FORMAT = 1 << 5,
            ERR_AIFC_WRONG_VERSION = 1 << 6, ERR_AIFC_UNSUPPORTED_FORMAT = 1 << 7, ERR_WAVE_UNSUPPORTED_FORMAT = 1 << 8,
            ERR_FILE_COULDNT_WRITE = 1 << 9
        };

        IAudioFile(const std::string& path = std::string(), int device = 0) : mOpen(false), mPosition(0), mDevice(device), mInfo() { open(path); }

        void open(const std::string& path)
        {
            close();
            if (path.empty()) return;
            if (hb_audio_probe(path.c_str(), &mInfo) < 0) throw std::runtime_error(hb_last_error());
            mOpen = !(mInfo.error_flags & ERR_FILE_COULDNT_OPEN);
            mPath = path;
        }
        void close() { mOpen = false; mPosition = 0; mInfo = hb_audio_info(); mPath.clear(); }
        bool isOpen() { return mOpen; }

        FileType getFileType() const { return static_cast<FileType>(mInfo.file_type); }
        PCMFormat getPCMFormat() const { return static_cast<PCMFormat>(mInfo.pcm_format); }
        Endianness getHeaderEndianness() const { return mInfo.header_big_endian ? kAudioFileBigEndian : kAudioFileLittleEndian; }
        Endianness getAudioEndianness() const { return mInfo.audio_big_endian ? kAudioFileBigEndian : kAudioFileLittleEndian; }
        double getSamplingRate() const { return mInfo.sampling_rate; }
        uint16_t getChannels() const { return static_cast<uint16_t>(mInfo.channels); }
        FrameCount getFrames() const { return mInfo.frames; }
        uint16_t getBitDepth() const { static const uint16_t bits[6] = {8, 16, 24, 32, 32, 64}; return bits[mInfo.pcm_format]; }
        uint16_t getByteDepth() const { return getBitDepth() / 8; }
        ByteCount getFrameByteCount() const { return ByteCount(getChannels()) * getByteDepth(); }
        NumberFormat getNumberFormat() const { return mInfo.pcm_format >= kAudioFileFloat32 ? kAudioFileFloat : kAudioFileInt; }
        int getErrorFlags() const { return mInfo.error_flags; }
        bool getIsError() const { return mInfo.error_flags != ERR_NONE; }
        void clearErrorFlags() { mInfo.error_flags = ERR_NONE; }

        void seek(FrameCount position = 0) { mPosition = position; }
        FrameCount getPosition() { return mPosition; }

        void readRaw(void* output, FrameCount numFrames)
        {
            if (!mOpen || !numFrames) return;
            if (hb_audio_read_raw(mPath.c_str(), mPosition, numFrames, output) < 0) throw std::runtime_error(hb_last_error());
            mPosition += numFrames;
        }
        void readInterleaved(double* output, FrameCount numFrames) { read(output, numFrames, -1, HB_F64); }
        void readInterleaved(float* output, FrameCount numFrames) { read(output, numFrames, -1, HB_F32); }
        void readChannel(double* output, FrameCount numFrames, uint16_t channel) { read(output, numFrames, channel, HB_F64); }
        void readChannel(float* output, FrameCount numFrames, uint16_t channel) { read(output, numFrames, channel, HB_F32); }

    private:

        void read(void *output, FrameCount numFrames, int32_t channel, int dtype)
        {
            if (!mOpen || !numFrames) return;
            if (hb_audio_read(mPath.c_str(), mPosition, numFrames, channel, output, dtype, mDevice) < 0) throw std::runtime_error(hb_last_error());
            mPosition += numFrames;
        }

        bool mOpen;
        FrameCount mPosition;
        int mDevice;
        hb_audio_info mInfo;
        std::string mPath;
    };
}

#endif
