"""IR-file ingestion, CPU side: the numpy oracle (oracle/audio_oracle.py) and the library's host-side header
parser (hb_audio_probe: no GPU needed) against fixtures written and read back by the unmodified reference
(tests/golden/audio/, tests/golden/golden_audio.npz, made by tests/golden/make_golden_audio.py), and the oracle against
the compiled reference on fresh files where it is available.  Byte / integer work: everything is bit-exact."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import checkers as ck

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import audio_oracle as ao  # noqa: E402

AUDIO = os.path.join(HERE, "golden", "audio")
G = np.load(os.path.join(HERE, "golden", "golden_audio.npz"))
FILES = sorted(f for f in os.listdir(AUDIO))
GOOD = [f for f in FILES if (f.replace(".", "_") + "_inter_f32") in G.files]


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def meta_of(info):
    return [info["file_type"], info["pcm_format"], info["header_big_endian"], info["audio_big_endian"], info["channels"], info["frames"], info["error_flags"]]


@pytest.mark.parametrize("name", FILES)
def test_oracle_probe_matches_reference_fixture(name):
    want = G[name.replace(".", "_") + "_meta"]
    info = ao.probe(os.path.join(AUDIO, name))
    if want[7]:                                          # unreadable: the error flags are what counts
        assert info["error_flags"] == want[7]
        return
    assert meta_of(info) == list(want[1:8])
    assert info["sampling_rate"] == G[name.replace(".", "_") + "_rate"][0]


@pytest.mark.parametrize("name", FILES)
def test_library_probe_matches_reference_fixture(name):
    """hb_audio_probe is host code: it runs without a GPU."""
    from hisstools_library_b200 import _abi
    from hisstools_library_b200.audiofile import AudioInfo
    info = AudioInfo()
    assert _abi.lib().hb_audio_probe(os.path.join(AUDIO, name).encode(), C.byref(info)) == 0
    want = G[name.replace(".", "_") + "_meta"]
    assert info.error_flags == want[7]
    if not want[7]:
        assert [info.file_type, info.pcm_format, info.header_big_endian, info.audio_big_endian, info.channels, info.frames] == list(want[1:7])
        assert info.sampling_rate == G[name.replace(".", "_") + "_rate"][0]
        assert info.pcm_offset == ao.probe(os.path.join(AUDIO, name))["pcm_offset"]


def test_library_probe_missing_file():
    from hisstools_library_b200 import _abi
    from hisstools_library_b200.audiofile import AudioInfo
    info = AudioInfo()
    assert _abi.lib().hb_audio_probe(b"/nonexistent/file.wav", C.byref(info)) == 0
    assert info.error_flags == ao.ERR_FILE_COULDNT_OPEN == ao.probe("/nonexistent/file.wav")["error_flags"]


@pytest.mark.parametrize("name", GOOD)
@pytest.mark.parametrize("suf,dt", [("f32", np.float32), ("f64", np.float64)])
def test_oracle_decode_is_bit_exact(name, suf, dt):
    key = name.replace(".", "_")
    path = os.path.join(AUDIO, name)
    info = ao.probe(path)
    got = ao.read(path, 0, info["frames"], -1, dt)
    assert np.array_equal(bits(got), bits(G[key + "_inter_" + suf]))
    got = ao.read(path, 17, 100, 1, dt)
    assert np.array_equal(bits(got), bits(G[key + "_ch1_from17_" + suf]))


def test_oracle_against_compiled_reference_on_fresh_files(tmp_path):
    ra = ck.ref_audio()
    if ra is None:
        pytest.skip("compiled reference not available")
    rng = np.random.default_rng(9)
    for ftype, ext in ((1, "aif"), (2, "aifc"), (3, "wav")):
        for pcm in range(6 if ftype != 1 else 4):
            for big in ((-1,) if ftype != 3 else (-1, 1)):
                channels, frames = int(rng.integers(1, 6)), int(rng.integers(1, 3000))
                x = np.ascontiguousarray(rng.uniform(-1.2, 1.2, (frames, channels)))
                path = str(tmp_path / ("t_%d_%d_%d.%s" % (ftype, pcm, big, ext)))
                ra.ref_audio_write(path.encode(), ftype, pcm, channels, 96000.0, big, ck.fptr(x), frames)
                info = ao.probe(path)
                rinfo = ck.RefAudioInfo()
                ra.ref_audio_probe(path.encode(), C.byref(rinfo))
                assert meta_of(info) == [rinfo.file_type, rinfo.pcm_format, rinfo.header_big_endian, rinfo.audio_big_endian, rinfo.channels, rinfo.frames,
                                         rinfo.error_flags]
                assert info["sampling_rate"] == rinfo.sampling_rate
                first = int(rng.integers(0, frames))
                n = frames - first
                for suf, dt in (("f32", np.float32), ("f64", np.float64)):
                    for channel in (-1, channels - 1):
                        want = np.zeros(n * (channels if channel < 0 else 1), dt)
                        assert getattr(ra, "ref_audio_read_" + suf)(path.encode(), first, n, channel, ck.fptr(want)) == 0
                        assert np.array_equal(bits(ao.read(path, first, n, channel, dt)), bits(want)), (ftype, pcm, big, suf, channel)
