"""The writer (hb_audio_writer_* / OAudioFile) and raw frame reads against the unmodified reference (oracle/_ref/libhisstools_ref_audio.so,
AudioFile/OAudioFile.cpp): the same sequence of calls must leave byte-identical files.  Host code only -- no GPU needed -- and
against committed fixtures where the reference library is not present (tests/golden/audio: files its OAudioFile wrote)."""
import ctypes as C
import os

import sys

import numpy as np
import pytest

import checkers as ck

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import audio_oracle as ao  # noqa: E402

AIFF, AIFC, WAVE = 1, 2, 3
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "audio")


@pytest.fixture(scope="module")
def hb():
    import hisstools_library_b200 as hb
    return hb


def signal(frames, channels, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1.0, 1.0, frames * channels)
    # values on and past the ends of the range, exact halves for the rounding rule, zeros and denormal-sized values
    special = np.array([1.0, -1.0, 0.999999, -0.999999, 1.5, -1.5, 2.75, -3.25, 0.5 / 32768, -0.5 / 32768, 1.5 / 128, -1.5 / 128,
                        0.0, -0.0, 1e-30, -1e-30, 0.5, -0.5, 127.5 / 128, -127.5 / 128])
    x[:min(len(special), x.size)] = special[:x.size]
    return x.astype(dtype)


def ref_or_skip():
    ref = ck.ref_audio()
    if ref is None or not hasattr(ref, "ref_oaudio_open"):
        pytest.skip("reference writer shim not built (oracle/_ref)")
    return ref


class Both:
    """the same calls on the reference's OAudioFile and on ours"""

    def __init__(self, hb, ref, tmp, type_, pcm, channels, rate, big):
        self.ref = ref
        self.pa, self.pb = os.path.join(tmp, "ref.bin"), os.path.join(tmp, "ours.bin")
        self.a = ref.ref_oaudio_open(self.pa.encode(), type_, pcm, channels, rate, big)
        self.b = hb.OAudioFile(self.pb, type_, pcm, channels, rate, None if big < 0 else big)

    def write(self, x, frames, channel=-1):
        x = np.ascontiguousarray(x)
        if x.dtype == np.float64:
            self.ref.ref_oaudio_write_f64(self.a, x.ctypes.data_as(ck.c_f64p), frames, channel)
        else:
            self.ref.ref_oaudio_write_f32(self.a, x.ctypes.data_as(ck.c_f32p), frames, channel)
        if channel < 0:
            self.b.writeInterleaved(x, frames)
        else:
            self.b.writeChannel(x, frames, channel)
        self.check_state()

    def raw(self, data, frames):
        self.ref.ref_oaudio_write_raw(self.a, data, frames)
        self.b.writeRaw(data, frames)
        self.check_state()

    def seek(self, frame):
        self.ref.ref_oaudio_seek(self.a, frame)
        self.b.seek(frame)
        self.check_state()

    def check_state(self):
        assert self.b.getPosition() == self.ref.ref_oaudio_position(self.a)
        assert self.b.getFrames() == self.ref.ref_oaudio_frames(self.a)
        assert self.b.getErrorFlags() == self.ref.ref_oaudio_flags(self.a)
        assert int(self.b.getFileType()) == self.ref.ref_oaudio_file_type(self.a)

    def close_and_compare(self):
        self.ref.ref_oaudio_close(self.a)
        self.b.close()
        a, b = open(self.pa, "rb").read(), open(self.pb, "rb").read()
        assert len(a) == len(b)
        assert a == b
        return a


@pytest.mark.parametrize("type_,big", [(WAVE, -1), (WAVE, 1), (WAVE, 0), (AIFC, -1), (AIFF, -1), (AIFC, 0)])
@pytest.mark.parametrize("pcm", range(6))
@pytest.mark.parametrize("channels,frames", [(1, 1001), (2, 500), (3, 333)])
def test_interleaved_files_byte_identical(hb, tmp_path, type_, big, pcm, channels, frames):
    """every file type x sample format x byte order, mono / stereo / three channels (odd byte counts: the pad byte), written in
    three calls of double and float input"""
    ref = ref_or_skip()
    x = signal(frames, channels, 10 * pcm + channels)
    w = Both(hb, ref, str(tmp_path), type_, pcm, channels, 44100.0, big)
    n1, n2 = frames // 3, frames // 2
    w.write(x[:n1 * channels], n1)
    w.write(x[n1 * channels:n2 * channels].astype(np.float32), n2 - n1)
    w.write(x[n2 * channels:], frames - n2)
    data = w.close_and_compare()
    assert len(data) > frames * channels * (8, 16, 24, 32, 32, 64)[pcm] // 8


@pytest.mark.parametrize("type_", [WAVE, AIFC])
@pytest.mark.parametrize("pcm", [0, 1, 2, 4, 5])
def test_channel_writes_seeks_and_raw_frames(hb, tmp_path, type_, pcm):
    """a three-channel file built one channel at a time (silence in the channels not yet written, later channels patched into
    existing frames), overwrites after seek, a write that starts past the end, raw frames, and the position / frame count after
    every call"""
    ref = ref_or_skip()
    ch, n = 3, 257
    xs = [signal(n, 1, 50 + c) for c in range(ch)]
    w = Both(hb, ref, str(tmp_path), type_, pcm, ch, 48000.0, -1)
    w.write(xs[1][:100], 100, 1)                 # channel 1 first: channels 0 and 2 are silence
    w.seek(0)
    w.write(xs[0], n, 0)                         # longer than what exists: extends the file
    w.seek(50)
    w.write(xs[2][50:].astype(np.float32), n - 50, 2)
    w.seek(10)
    w.write(signal(5, ch, 77), 5)                # interleaved overwrite in the middle
    w.seek(n + 7)
    w.write(xs[1][:9], 9, 1)                     # starts past the end: the gap is silence
    bd = (1, 2, 3, 4, 4, 8)[pcm]
    w.seek(3)
    w.raw(bytes(range(1, 1 + 2 * ch * bd)), 2)
    w.seek(n + 16)
    w.raw(bytes([0xAA] * (ch * bd)), 1)
    w.close_and_compare()


def test_sampling_rates_and_empty_file(hb, tmp_path):
    """the 80-bit extended sampling rate of AIFC for ordinary and odd rates, and files closed without any audio"""
    ref = ref_or_skip()
    for k, rate in enumerate([44100.0, 48000.0, 96000.0, 22050.5, 1.0, 192000.0, 8000.0, 11025.0, 0.0, 3.5e9]):
        for type_ in (WAVE, AIFC):
            d = tmp_path / ("r%d_%d" % (k, type_))
            d.mkdir()
            w = Both(hb, ref, str(d), type_, 1, 2, rate, -1)
            if k % 2:
                w.write(signal(10, 2, k), 10)
            w.close_and_compare()


def test_unwritable_path_and_closed_object(hb, tmp_path):
    ref = ref_or_skip()
    bad = os.path.join(str(tmp_path), "no_such_dir", "x.wav")
    a = ref.ref_oaudio_open(bad.encode(), WAVE, 1, 1, 44100.0, -1)
    b = hb.OAudioFile(bad, WAVE, 1, 1, 44100.0)
    assert not b.isOpen() and not ref.ref_oaudio_is_open(a)
    assert b.getErrorFlags() == ref.ref_oaudio_flags(a) == 4
    b.writeInterleaved(np.zeros(4), 4)           # ignored, as by the reference
    ref.ref_oaudio_close(a)
    b.close()
    c = hb.OAudioFile()
    assert not c.isOpen() and c.getPosition() == 0 and c.getFrames() == 0


FIXTURES = []          # (name, type, pcm, big_endian or -1): how tests/golden/make_golden_audio.py had the reference write them
for _pcm, _tag in enumerate(("i8", "i16", "i24", "i32", "f32", "f64")):
    FIXTURES.append(("wave_%s.wav" % _tag, WAVE, _pcm, -1))
    FIXTURES.append(("rifx_%s.wav" % _tag, WAVE, _pcm, 1))
    FIXTURES.append(("aifc_%s.aifc" % _tag, AIFC, _pcm, -1))
    if _pcm < 4:
        FIXTURES.append(("aiff_%s.aif" % _tag, AIFF, _pcm, -1))
FIXTURES.append(("aifc_sowt.aifc", AIFC, 1, 0))


@pytest.mark.parametrize("name,type_,pcm,big", FIXTURES)
def test_fixture_files_rewritten_from_their_samples(hb, tmp_path, name, type_, pcm, big):
    """without the reference library: decode a fixture the reference's OAudioFile wrote (numpy oracle), write the samples again with
    our writer and the parameters the fixture was made with, and get the fixture back byte for byte (every format decodes to a
    double that encodes to the same bytes); readRaw returns the stored frames"""
    path = os.path.join(GOLD, name)
    raw = open(path, "rb").read()
    info = ao.probe(path)
    assert info["error_flags"] == 0
    frames, ch = info["frames"], info["channels"]
    x = ao.read(path, 0, frames, -1, np.float64)
    if info["pcm_format"] != pcm:
        # the reference's READER takes an AIFC "fl64" file for 32-bit floats (twice the frames); the file itself is what its writer
        # made of 257 frames of doubles, so take the doubles straight from the bytes
        frames = 257
        x = np.frombuffer(raw, ">f8", count=frames * ch, offset=info["pcm_offset"]).astype(np.float64)
    if name == "aifc_sowt.aifc":
        # little-endian samples under the tag "NONE" (OAudioFile.cpp:404-417): a reader takes them for big-endian ones
        x = np.frombuffer(raw, "<i2", count=frames * ch, offset=info["pcm_offset"]).astype(np.float64) / 32768.0
    out = os.path.join(str(tmp_path), "again.bin")
    w = hb.OAudioFile(out, type_, pcm, ch, info["sampling_rate"], None if big < 0 else big)
    assert w.isOpen() and int(w.getFileType()) == (AIFC if type_ == AIFF else type_)
    w.writeInterleaved(x, frames)
    assert w.getFrames() == frames and w.getPosition() == frames
    w.close()
    assert open(out, "rb").read() == raw
    if info["pcm_format"] != pcm:
        return
    r = hb.IAudioFile(path)
    r.seek(1)
    got = r.readRaw(frames - 1)
    fb = ch * (8, 16, 24, 32, 32, 64)[pcm] // 8
    assert got == raw[info["pcm_offset"] + fb:info["pcm_offset"] + fb * frames]
    assert r.getPosition() == frames


def test_read_raw_against_reference(hb, tmp_path):
    ref = ref_or_skip()
    path = os.path.join(str(tmp_path), "x.aifc")
    x = signal(300, 2, 5)
    assert ref.ref_audio_write(path.encode(), AIFC, 2, 2, 44100.0, -1, x.ctypes.data_as(ck.c_f64p), 300) == 0
    want = (C.c_ubyte * (100 * 6))()
    assert ref.ref_audio_read_raw(path.encode(), 37, 100, want) == 0
    f = hb.IAudioFile(path)
    f.seek(37)
    assert f.readRaw(100) == bytes(want)


def test_cpp_caller_of_the_writer(hb, tmp_path):
    """tests/cpp/audio_writer_test.cpp, written against the reference's OAudioFile / IAudioFile::readRaw API, compiled against
    include/ and linked to the library: the files it writes equal what the same calls give through the reference."""
    import shutil
    import subprocess
    from hisstools_library_b200 import build
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(build.lib_path())
    exe = os.path.join(str(tmp_path), "audio_writer_test")
    subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "cpp", "audio_writer_test.cpp"),
                    "-o", exe, "-L" + libdir, "-lhisstools_b200", "-Wl,-rpath," + libdir], check=True)
    wav, aifc = os.path.join(str(tmp_path), "a.wav"), os.path.join(str(tmp_path), "a.aifc")
    res = subprocess.run([exe, wav, aifc], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.strip() == "ok", (res.returncode, res.stdout, res.stderr)
    frames = 301
    k = np.arange(frames)
    left, right = np.sin(0.05 * k) * 0.9, np.cos(0.031 * k) * 1.2
    inter = np.stack([left.astype(np.float32), right.astype(np.float32)], axis=1).reshape(-1)
    ref = ck.ref_audio()
    if ref is not None and hasattr(ref, "ref_oaudio_open"):
        p1, p2 = os.path.join(str(tmp_path), "r.wav"), os.path.join(str(tmp_path), "r.aifc")
        h = ref.ref_oaudio_open(p1.encode(), WAVE, 2, 2, 48000.0, -1)
        ref.ref_oaudio_write_f64(h, right.ctypes.data_as(ck.c_f64p), frames, 1)
        ref.ref_oaudio_seek(h, 0)
        ref.ref_oaudio_write_f64(h, left.ctypes.data_as(ck.c_f64p), frames, 0)
        ref.ref_oaudio_close(h)
        h = ref.ref_oaudio_open(p2.encode(), AIFF, 4, 2, 44100.0, -1)
        ref.ref_oaudio_write_f32(h, inter.ctypes.data_as(ck.c_f32p), frames, -1)
        ref.ref_oaudio_close(h)
        assert open(wav, "rb").read() == open(p1, "rb").read()
        assert open(aifc, "rb").read() == open(p2, "rb").read()
    else:
        w = hb.OAudioFile(os.path.join(str(tmp_path), "p.aifc"), AIFF, 4, 2, 44100.0)
        w.writeInterleaved(inter, frames)
        w.close()
        assert open(aifc, "rb").read() == open(os.path.join(str(tmp_path), "p.aifc"), "rb").read()
