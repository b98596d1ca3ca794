"""oracle/audio_oracle.py -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the reference's audio-file reader for the byte / integer work of IR ingestion:
header parsing (AudioFile/IAudioFile.cpp:375-609) and sample decoding (readAudio :613-689 with the conversions
:137-239).  Pinned bit for bit against the unmodified reference (oracle/_ref/libhisstools_ref_audio.so) by
tests/test_oracle_vs_reference.py and against the fixtures under tests/golden/audio/ (files written by the reference's
OAudioFile, samples read back by its IAudioFile; tests/golden/make_golden.py).  Only tests/ may import it.
"""
import struct

import numpy as np

ERR_FILE_COULDNT_OPEN, ERR_FILE_BAD_FORMAT, ERR_FILE_UNKNOWN_FORMAT, ERR_FILE_UNSUPPORTED_PCM_FORMAT = 1 << 2, 1 << 3, 1 << 4, 1 << 5
ERR_AIFC_WRONG_VERSION, ERR_AIFC_UNSUPPORTED_FORMAT, ERR_WAVE_UNSUPPORTED_FORMAT = 1 << 6, 1 << 7, 1 << 8
BITS = (8, 16, 24, 32, 32, 64)


def _pcm(bits, is_float):                                   # IAudioFile.cpp:289-322
    table = {(8, 0): 0, (16, 0): 1, (24, 0): 2, (32, 0): 3, (32, 1): 4, (64, 1): 5}
    return table.get((bits, int(is_float)))


def _extended(b):                                           # IAudioFile.cpp:188-216
    se, hi, lo = struct.unpack(">HII", b)
    exp = se & 0x7FFF
    if not exp and not hi and not lo:
        return 0.0
    if exp == 0x7FFF:
        return float("inf")
    exp -= 16383
    v = float(np.ldexp(float(hi), exp - 31) + np.ldexp(float(lo), exp - 63))
    return -v if se & 0x8000 else v


def probe(path):
    """dict with the BaseAudioFile state after IAudioFile::open (error_flags as getErrorFlags())."""
    info = dict(file_type=0, pcm_format=0, header_big_endian=0, audio_big_endian=0, channels=0, frames=0, sampling_rate=0.0, pcm_offset=0, error_flags=0)
    try:
        data = open(path, "rb").read()
    except OSError:
        info["error_flags"] = ERR_FILE_COULDNT_OPEN
        return info
    if len(data) < 12:
        info["error_flags"] |= ERR_FILE_BAD_FORMAT
        return info
    head, sub = data[:4], data[8:12]
    pos = 12

    def pad(n):
        return n + (n & 1)

    if head == b"FORM" and sub in (b"AIFF", b"AIFC"):
        aifc = sub == b"AIFC"
        info["header_big_endian"] = 1
        valid, seen = (2 | 4 | (1 if aifc else 0)), 0
        if aifc:
            info["file_type"] = 2
        while pos + 8 <= len(data):
            tag, size = data[pos:pos + 4], struct.unpack(">I", data[pos + 4:pos + 8])[0]
            pos += 8
            if tag == b"FVER":
                seen |= 1
                if size < 4 or pos + 4 > len(data):
                    info["error_flags"] |= ERR_FILE_BAD_FORMAT
                    return info
                if struct.unpack(">I", data[pos:pos + 4])[0] != 0xA2805140:
                    info["error_flags"] |= ERR_AIFC_WRONG_VERSION
                    return info
            elif tag == b"COMM":
                seen |= 2
                want = 22 if size > 22 else (18 if size < 18 else size)
                if want > size or pos + want > len(data):
                    info["error_flags"] |= ERR_FILE_BAD_FORMAT
                    return info
                c = data[pos:pos + want] + b"\0" * (22 - want)
                info["channels"], info["frames"], bits = struct.unpack(">HIH", c[:8])
                info["sampling_rate"] = _extended(c[8:18])
                is_float = False
                info["audio_big_endian"] = 1
                if not info["frames"]:
                    seen |= 4
                if aifc:
                    t = c[18:22]
                    if t == b"NONE":
                        pass
                    elif t == b"twos":
                        bits = 16
                    elif t == b"sowt":
                        bits, info["audio_big_endian"] = 16, 0
                    elif t in (b"fl32", b"FL32", b"fl64", b"FL64"):      # fl64 gets 32 bits in the reference (:380-384)
                        bits, is_float = 32, True
                    else:
                        info["error_flags"] |= ERR_AIFC_UNSUPPORTED_FORMAT
                        return info
                else:
                    info["file_type"] = 1
                fmt = _pcm(bits, is_float)
                if fmt is None:
                    info["error_flags"] |= ERR_FILE_UNSUPPORTED_PCM_FORMAT
                    return info
                info["pcm_format"] = fmt
            elif tag == b"SSND":
                seen |= 4
                if size < 4 or pos + 4 > len(data):
                    info["error_flags"] |= ERR_FILE_BAD_FORMAT
                    return info
                info["pcm_offset"] = pos + 8 + struct.unpack(">I", data[pos:pos + 4])[0]
            pos += pad(size)
            if pos > len(data):
                # the reference's seek past the end succeeds (ifstream), the next header read then fails: loop ends
                break
        if ~seen & valid:
            info["error_flags"] |= ERR_FILE_BAD_FORMAT
        return info
    if head in (b"RIFF", b"RIFX") and sub == b"WAVE":
        big = head == b"RIFX"
        e = ">" if big else "<"
        info["header_big_endian"] = info["audio_big_endian"] = int(big)

        def find(tag, pos):
            while pos + 8 <= len(data):
                t, size = data[pos:pos + 4], struct.unpack(e + "I", data[pos + 4:pos + 8])[0]
                pos += 8
                if t == tag:
                    return pos, size
                pos += pad(size)
            return None, 0

        pos, size = find(b"fmt ", pos)
        if pos is None or size < 16 or pos + 16 > len(data):
            info["error_flags"] |= ERR_FILE_BAD_FORMAT
            return info
        tag, ch, rate, _, _, bits = struct.unpack(e + "HHIIHH", data[pos:pos + 16])
        pos += pad(size)
        if tag not in (1, 3):
            info["error_flags"] |= ERR_WAVE_UNSUPPORTED_FORMAT
            return info
        info["channels"], info["sampling_rate"] = ch, float(rate)
        fmt = _pcm(bits, tag == 3)
        if fmt is None:
            info["error_flags"] |= ERR_FILE_UNSUPPORTED_PCM_FORMAT
            return info
        info["pcm_format"] = fmt
        pos, size = find(b"data", pos)
        if pos is None:
            info["error_flags"] |= ERR_FILE_BAD_FORMAT
            return info
        fb = ch * (BITS[fmt] // 8)
        info["frames"] = size // fb if fb else 0
        info["pcm_offset"] = pos
        info["file_type"] = 3
        return info
    info["error_flags"] |= ERR_FILE_UNKNOWN_FORMAT
    return info


def read(path, first, frames, channel, dtype):
    """seek(first) + readChannel(channel >= 0) / readInterleaved(channel < 0) as an array of dtype."""
    info = probe(path)
    assert not info["error_flags"]
    ch, fmt, big = info["channels"], info["pcm_format"], info["audio_big_endian"]
    bd = BITS[fmt] // 8
    data = open(path, "rb").read()
    raw = np.frombuffer(data, np.uint8, count=frames * ch * bd, offset=info["pcm_offset"] + first * ch * bd).reshape(frames, ch, bd)
    if channel >= 0:
        raw = raw[:, channel:channel + 1, :]
    b = raw.astype(np.uint32) if bd <= 4 else raw.astype(np.uint64)
    order = range(bd) if big else range(bd - 1, -1, -1)                  # most significant byte first
    word = np.zeros(raw.shape[:2], b.dtype)
    for k in order:
        word = (word << 8) | b[:, :, k]
    dtype = np.dtype(dtype)
    scale = dtype.type(4.656612873077392578125e-10)
    if fmt == 0 and info["file_type"] == 3:                              # WAVE 8-bit: unsigned (IAudioFile.cpp:225-228, 643-647)
        out = (word.astype(dtype) - dtype.type(128)) / dtype.type(128)
    elif fmt <= 3:
        i32 = (word << np.uint32(32 - 8 * bd)).astype(np.uint32).view(np.int32)
        out = i32.astype(dtype) * scale                                  # int32 -> T, times 2^-31 (:218-223)
    elif fmt == 4:
        with np.errstate(invalid="ignore"):                               # signalling NaNs of a misread file (fl64 as 32-bit floats)
            out = word.astype(np.uint32).view(np.float32).astype(dtype)
    else:
        out = word.view(np.float64).astype(dtype)
    return np.ascontiguousarray(out.reshape(-1))
