"""IAudioFile / OAudioFile -- WAV / AIFF / AIFC reader feeding impulse responses to the convolvers, and the writer.

Mirror of the reading half of the reference's AudioFile component (AudioFile/IAudioFile.h:30-54 with the
BaseAudioFile getters, BaseAudioFile.h:64-90): same method names and meanings.  The header is parsed by the
library's host code, the PCM samples are decoded on the GPU (hb_audio_* of include/hisstools_b200.h).
"""
import ctypes as C
import enum

import numpy as np

from . import _abi


class AudioInfo(C.Structure):
    """hb_audio_info"""
    _fields_ = [("file_type", C.c_int32), ("pcm_format", C.c_int32), ("header_big_endian", C.c_int32), ("audio_big_endian", C.c_int32),
                ("channels", C.c_uint32), ("frames", C.c_uint32), ("sampling_rate", C.c_double), ("pcm_offset", C.c_uint64),
                ("error_flags", C.c_int32), ("reserved", C.c_int32)]


class FileType(enum.IntEnum):
    kAudioFileNone = 0
    kAudioFileAIFF = 1
    kAudioFileAIFC = 2
    kAudioFileWAVE = 3


class PCMFormat(enum.IntEnum):
    kAudioFileInt8 = 0
    kAudioFileInt16 = 1
    kAudioFileInt24 = 2
    kAudioFileInt32 = 3
    kAudioFileFloat32 = 4
    kAudioFileFloat64 = 5


_BITS = (8, 16, 24, 32, 32, 64)


class IAudioFile:
    FileType = FileType
    PCMFormat = PCMFormat

    def __init__(self, path="", device=0):
        self._path = None
        self._pos = 0
        self._device = int(device)
        self._info = AudioInfo()
        if path:
            self.open(path)

    # File Open / Close (IAudioFile.cpp:36-68)
    def open(self, path):
        self.close()
        info = AudioInfo()
        _abi.check(_abi.lib().hb_audio_probe(str(path).encode(), C.byref(info)))
        self._info = info
        self._pos = 0
        if not (info.error_flags & 4):                      # ERR_FILE_COULDNT_OPEN
            self._path = str(path)

    def close(self):
        self._path = None
        self._info = AudioInfo()
        self._pos = 0

    def isOpen(self):
        return self._path is not None

    # BaseAudioFile getters
    def getFileType(self):
        return FileType(self._info.file_type)

    def getPCMFormat(self):
        return PCMFormat(self._info.pcm_format)

    def getHeaderEndianness(self):
        return int(self._info.header_big_endian)

    def getAudioEndianness(self):
        return int(self._info.audio_big_endian)

    def getSamplingRate(self):
        return float(self._info.sampling_rate)

    def getChannels(self):
        return int(self._info.channels)

    def getFrames(self):
        return int(self._info.frames)

    def getBitDepth(self):
        return _BITS[self._info.pcm_format]

    def getByteDepth(self):
        return self.getBitDepth() // 8

    def getFrameByteCount(self):
        return self.getChannels() * self.getByteDepth()

    def getErrorFlags(self):
        return int(self._info.error_flags)

    def getIsError(self):
        return self._info.error_flags != 0

    # File Position (IAudioFile.cpp:74-88)
    def seek(self, position=0):
        self._pos = int(position)

    def getPosition(self):
        return self._pos

    def _read(self, output, num_frames, channel):
        n = int(num_frames)
        per = self.getChannels() if channel < 0 else 1
        if output.dtype not in (np.float32, np.float64) or not output.flags["C_CONTIGUOUS"] or output.size < n * per:
            raise ValueError("output must be a contiguous float32 / float64 array of at least %d elements" % (n * per))
        if not self.isOpen():
            raise _abi.HissError(_abi.HB_ERR_BAD_ARG, "no file is open")
        _abi.check(_abi.lib().hb_audio_read(self._path.encode(), self._pos, n, int(channel), output.ctypes.data_as(C.c_void_p),
                                            _abi.HB_F64 if output.dtype == np.float64 else _abi.HB_F32, self._device))
        self._pos += n

    # File Reading (IAudioFile.cpp:96-115)
    def readInterleaved(self, output, numFrames):
        self._read(output, numFrames, -1)

    def readChannel(self, output, numFrames, channel):
        self._read(output, numFrames, int(channel))

    def readRaw(self, numFrames):
        """the next numFrames frames as the file stores them (IAudioFile.cpp:85-88): bytes"""
        n = int(numFrames)
        if not self.isOpen():
            raise _abi.HissError(_abi.HB_ERR_BAD_ARG, "no file is open")
        buf = (C.c_ubyte * (n * self.getFrameByteCount()))()
        _abi.check(_abi.lib().hb_audio_read_raw(self._path.encode(), self._pos, n, buf))
        self._pos += n
        return bytes(buf)


class OAudioFile:
    """Mirror of the reference's writer (AudioFile/OAudioFile.h:18-30): same method names and meanings; the files are byte for
    byte what the reference writes (host code of the library: hb_audio_writer_*)."""
    FileType = FileType
    PCMFormat = PCMFormat

    def __init__(self, path=None, type=None, format=None, channels=0, sr=0.0, endianness=None):
        self._h = C.c_void_p()
        if path is not None:
            self.open(path, type, format, channels, sr, endianness)

    def open(self, path, type, format, channels, sr, endianness=None):
        """endianness: None = the type's default, 0 little, 1 big (BaseAudioFile::Endianness)"""
        self.close()
        _abi.check(_abi.lib().hb_audio_writer_open(C.byref(self._h), str(path).encode(), int(type), int(format), int(channels), float(sr),
                                                   -1 if endianness is None else int(endianness)))

    def close(self):
        if self._h:
            _abi.lib().hb_audio_writer_close(self._h)
            self._h = C.c_void_p()

    def _info(self):
        info, is_open = AudioInfo(), C.c_int(0)
        if self._h:
            _abi.check(_abi.lib().hb_audio_writer_info(self._h, C.byref(info), C.byref(is_open)))
        return info, bool(is_open.value)

    def isOpen(self):
        return self._info()[1]

    def getFileType(self):
        return FileType(self._info()[0].file_type)

    def getPCMFormat(self):
        return PCMFormat(self._info()[0].pcm_format)

    def getChannels(self):
        return int(self._info()[0].channels)

    def getFrames(self):
        return int(self._info()[0].frames)

    def getSamplingRate(self):
        return float(self._info()[0].sampling_rate)

    def getErrorFlags(self):
        return int(self._info()[0].error_flags)

    def seek(self, position=0):
        if self._h:
            _abi.check(_abi.lib().hb_audio_writer_seek(self._h, int(position)))

    def getPosition(self):
        return int(_abi.lib().hb_audio_writer_position(self._h)) if self._h else 0

    def _write(self, input, num_frames, channel):
        a = np.ascontiguousarray(input)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        n = int(num_frames)
        per = self.getChannels() if channel < 0 else 1
        if a.size < n * per:
            raise ValueError("input holds fewer than %d samples" % (n * per))
        if self._h:
            _abi.check(_abi.lib().hb_audio_writer_write(self._h, a.ctypes.data_as(C.c_void_p), _abi.HB_F64 if a.dtype == np.float64 else _abi.HB_F32,
                                                        n, int(channel)))

    def writeInterleaved(self, input, numFrames):
        self._write(input, numFrames, -1)

    def writeChannel(self, input, numFrames, channel):
        self._write(input, numFrames, int(channel))

    def writeRaw(self, input, numFrames):
        buf = bytes(input)
        if self._h:
            _abi.check(_abi.lib().hb_audio_writer_write_raw(self._h, buf, int(numFrames)))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
