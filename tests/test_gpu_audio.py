"""IR-file ingestion on the GPU: hb_audio_read / hb_audio_decode_dev / hb_conv_set_ir_file and the IAudioFile
mirror against fixtures written and read back by the unmodified reference (bit-exact), the numpy oracle on fresh files,
and a convolver fed from a file against one fed from the decoded array."""
import os
import struct
import sys

import numpy as np
import pytest

import checkers as ck

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import audio_oracle as ao  # noqa: E402

AUDIO = os.path.join(HERE, "golden", "audio")
G = np.load(os.path.join(HERE, "golden", "golden_audio.npz"))
GOOD = sorted(f for f in os.listdir(AUDIO) if (f.replace(".", "_") + "_inter_f32") in G.files)
BAD = sorted(f for f in os.listdir(AUDIO) if f not in GOOD)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


@pytest.fixture(scope="module")
def hb():
    import hisstools_library_b200 as h
    return h


@pytest.mark.parametrize("name", GOOD)
def test_reference_fixtures_bit_exact(hb, name):
    key = name.replace(".", "_")
    f = hb.IAudioFile(os.path.join(AUDIO, name))
    meta = G[key + "_meta"]
    assert f.isOpen() and not f.getIsError()
    assert [int(f.getFileType()), int(f.getPCMFormat()), f.getHeaderEndianness(), f.getAudioEndianness(), f.getChannels(), f.getFrames()] == list(meta[1:7])
    assert f.getSamplingRate() == G[key + "_rate"][0]
    for suf, dt in (("f32", np.float32), ("f64", np.float64)):
        inter = np.zeros(f.getFrames() * f.getChannels(), dt)
        f.seek(0)
        f.readInterleaved(inter, f.getFrames())
        assert f.getPosition() == f.getFrames()
        assert np.array_equal(bits(inter), bits(G[key + "_inter_" + suf]))
        ch1 = np.zeros(100, dt)
        f.seek(17)
        f.readChannel(ch1, 100, 1)
        assert f.getPosition() == 117
        assert np.array_equal(bits(ch1), bits(G[key + "_ch1_from17_" + suf]))


@pytest.mark.parametrize("name", BAD)
def test_unreadable_files_report_the_reference_flags(hb, name):
    f = hb.IAudioFile(os.path.join(AUDIO, name))
    assert f.getErrorFlags() == G[name.replace(".", "_") + "_meta"][7] != 0
    with pytest.raises(hb.HissError):
        f.readChannel(np.zeros(4, np.float32), 4, 0)


def _write_wav(path, x, bits_per_sample=24):
    """little-endian integer WAVE written by hand (not by the library under test)"""
    frames, ch = x.shape
    q = np.clip(np.round(x * (2 ** (bits_per_sample - 1))), -(2 ** (bits_per_sample - 1)), 2 ** (bits_per_sample - 1) - 1).astype(np.int64)
    bd = bits_per_sample // 8
    raw = bytearray()
    for v in q.reshape(-1):
        raw += int(v).to_bytes(bd, "little", signed=True)
    hdr = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, ch, 48000, 48000 * ch * bd, ch * bd, bits_per_sample)
    open(path, "wb").write(hdr + b"data" + struct.pack("<I", len(raw)) + bytes(raw))


def test_large_file_against_the_numpy_oracle_and_planar_device_decode(hb, tmp_path):
    import ctypes as C
    import torch
    from hisstools_library_b200 import _abi
    from hisstools_library_b200.audiofile import AudioInfo
    rng = np.random.default_rng(31)
    frames, ch = 70001, 5
    x = rng.uniform(-1, 1, (frames, ch))
    path = str(tmp_path / "big24.wav")
    _write_wav(path, x, 24)
    f = hb.IAudioFile(path)
    assert (f.getChannels(), f.getFrames(), f.getBitDepth()) == (ch, frames, 24)
    for dt in (np.float32, np.float64):
        got = np.zeros(frames * ch, dt)
        f.seek(0)
        f.readInterleaved(got, frames)
        assert np.array_equal(bits(got), bits(ao.read(path, 0, frames, -1, dt)))
        assert np.max(np.abs(got.reshape(frames, ch) - x)) <= 2.0 ** -23
        one = np.zeros(1234, dt)
        f.seek(60000)
        f.readChannel(one, 1234, 3)
        assert np.array_equal(bits(one), bits(ao.read(path, 60000, 1234, 3, dt)))
    # decode step alone, planar rows on the device
    info = AudioInfo()
    _abi.check(_abi.lib().hb_audio_probe(path.encode(), C.byref(info)))
    raw = np.frombuffer(open(path, "rb").read(), np.uint8, offset=info.pcm_offset)
    d_raw = torch.from_numpy(raw.copy()).cuda()
    d_out = torch.zeros(ch, frames + 7, device="cuda")
    _abi.check(_abi.lib().hb_audio_decode_dev(C.byref(info), C.c_void_p(d_raw.data_ptr()), frames, -1, C.c_void_p(d_out.data_ptr()), frames + 7, _abi.HB_F32, 0, None))
    torch.cuda.synchronize()
    want = ao.read(path, 0, frames, -1, np.float32).reshape(frames, ch).T
    assert np.array_equal(bits(d_out[:, :frames].cpu().numpy()), bits(want))
    assert float(d_out[:, frames:].abs().max()) == 0.0


def test_convolver_fed_from_a_file_equals_one_fed_from_the_array(hb, tmp_path):
    from hisstools_library_b200 import _abi
    from hisstools_library_b200.convolve import _Engine
    rng = np.random.default_rng(32)
    L, B, ch = 3000, 256, 4                                  # a 2-in x 2-out matrix from one 4-channel IR file
    ir = rng.standard_normal((L, ch)) * np.exp(-6.9 * np.arange(L) / L)[:, None] * 0.2
    path = str(tmp_path / "ir.wav")
    _write_wav(path, ir, 24)
    decoded = ao.read(path, 0, L, -1, np.float32).reshape(L, ch)
    x = np.stack([ck.synth_audio(B * 20, 40 + i) for i in range(2)])
    outs = []
    for from_file in (True, False):
        e = _Engine(np.float32, 1, 2, 2, 2 * B, L, 0, 0, 0)
        e.set_reset_offset(0)
        for o in range(2):
            for i in range(2):
                if from_file:
                    assert _abi.check(_abi.lib().hb_conv_set_ir_file(e._h, 0, i, o, path.encode(), o * 2 + i, 0)) == 0
                else:
                    assert e.set_ir(0, i, o, np.ascontiguousarray(decoded[:, o * 2 + i]), L) == 0
        y = [np.zeros(x.shape[1], np.float32) for _ in range(2)]
        e.process([x[0], x[1]], y, x.shape[1])
        outs.append(np.stack(y))
        e.close()
    assert np.array_equal(outs[0], outs[1])
    for o in range(2):
        truth = sum(ck.direct_convolve_delayed(decoded[:, o * 2 + i], x[i], B) for i in range(2))
        assert ck.rel_rms(outs[0][o], truth) <= 1e-5
