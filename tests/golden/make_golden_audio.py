"""Makes tests/golden/audio/*.{wav,aif,aifc} with the reference's own writer (OAudioFile) and golden_audio.npz with what the
reference's reader (IAudioFile) returns for them -- through oracle/_ref/libhisstools_ref_audio.so (oracle/ref_audio_shim.cpp).
Run in the build container (needs /root/reference); the fixtures are committed."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import checkers as ck  # noqa: E402

CASES = []          # (name, type, pcm, big_endian or -1)
for pcm, tag in enumerate(("i8", "i16", "i24", "i32", "f32", "f64")):
    CASES.append(("wave_%s.wav" % tag, 3, pcm, -1))
    CASES.append(("rifx_%s.wav" % tag, 3, pcm, 1))
    CASES.append(("aifc_%s.aifc" % tag, 2, pcm, -1))
    if pcm < 4:
        CASES.append(("aiff_%s.aif" % tag, 1, pcm, -1))
CASES.append(("aifc_sowt.aifc", 2, 1, 0))


def main():
    ra = ck.ref_audio()
    rng = np.random.default_rng(4242)
    out = {}
    os.makedirs(os.path.join(HERE, "audio"), exist_ok=True)
    for name, ftype, pcm, big in CASES:
        channels, frames = 3, 257
        x = rng.uniform(-1, 1, (frames, channels))
        x[0, 0], x[1, 0], x[2, 0] = 1.0, -1.0, 0.0
        path = os.path.join(HERE, "audio", name)
        flags = ra.ref_audio_write(path.encode(), ftype, pcm, channels, 48000.0 if "i24" in name else 44100.0, big, ck.fptr(np.ascontiguousarray(x)), frames)
        info = ck.RefAudioInfo()
        ra.ref_audio_probe(path.encode(), C.byref(info))
        key = name.replace(".", "_")
        out[key + "_meta"] = np.array([flags, info.file_type, info.pcm_format, info.header_big_endian, info.audio_big_endian, info.channels, info.frames,
                                       info.error_flags, info.is_open], np.int64)
        out[key + "_rate"] = np.array([info.sampling_rate])
        if info.error_flags or not info.frames:
            print(name, "write flags", flags, "read flags", info.error_flags, "frames", info.frames)
            continue
        for suf, dt in (("f32", np.float32), ("f64", np.float64)):
            inter = np.zeros(info.frames * info.channels, dt)
            assert getattr(ra, "ref_audio_read_" + suf)(path.encode(), 0, info.frames, -1, ck.fptr(inter)) == 0
            out[key + "_inter_" + suf] = inter
            ch1 = np.zeros(100, dt)
            assert getattr(ra, "ref_audio_read_" + suf)(path.encode(), 17, 100, 1, ck.fptr(ch1)) == 0
            out[key + "_ch1_from17_" + suf] = ch1
    # malformed / unsupported headers: what the reference reports
    bad = {"empty.wav": b"", "short.wav": b"RIFF\x00\x00", "noise.bin": bytes(range(64)),
           "wave_nofmt.wav": b"RIFF\x24\x00\x00\x00WAVEdata\x04\x00\x00\x00\x00\x00\x00\x00",
           "wave_adpcm.wav": b"RIFF\x24\x00\x00\x00WAVEfmt \x10\x00\x00\x00\x02\x00\x01\x00\x44\xac\x00\x00\x88\x58\x01\x00\x02\x00\x10\x00data\x00\x00\x00\x00",
           "wave_12bit.wav": b"RIFF\x24\x00\x00\x00WAVEfmt \x10\x00\x00\x00\x01\x00\x01\x00\x44\xac\x00\x00\x88\x58\x01\x00\x02\x00\x0c\x00data\x00\x00\x00\x00",
           "aiff_nocomm.aif": b"FORM\x00\x00\x00\x10AIFFSSND\x00\x00\x00\x08\x00\x00\x00\x00\x00\x00\x00\x00"}
    for name, data in bad.items():
        path = os.path.join(HERE, "audio", name)
        open(path, "wb").write(data)
        info = ck.RefAudioInfo()
        ra.ref_audio_probe(path.encode(), C.byref(info))
        out[name.replace(".", "_") + "_meta"] = np.array([0, info.file_type, info.pcm_format, info.header_big_endian, info.audio_big_endian, info.channels,
                                                         info.frames, info.error_flags, info.is_open], np.int64)
    path = os.path.join(HERE, "golden_audio.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays;", len(os.listdir(os.path.join(HERE, "audio"))), "files")


if __name__ == "__main__":
    main()
