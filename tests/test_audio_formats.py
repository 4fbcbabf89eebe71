"""WAV decoding on the host side of the file-level seams (ser_b200/audio.py; the reference reads through
librosa.load -> soundfile, ser/_internal/utils/audio_utils.py:104-113): every sample format lands as
float32 in soundfile's convention, then the reference's preparation (mono mix, peak normalisation)."""

from __future__ import annotations

import struct
import wave

import numpy as np
import pytest

from ser_b200 import audio


def _riff(fmt_body: bytes, data: bytes, extra_chunks: bytes = b"") -> bytes:
    chunks = b"fmt " + struct.pack("<I", len(fmt_body)) + fmt_body + (b"\0" if len(fmt_body) & 1 else b"")
    chunks += extra_chunks
    chunks += b"data" + struct.pack("<I", len(data)) + data + (b"\0" if len(data) & 1 else b"")
    return b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks


def _fmt(tag, channels, rate, bits, extensible_sub=None):
    align = channels * bits // 8
    body = struct.pack("<HHIIHH", tag, channels, rate, rate * align, align, bits)
    if extensible_sub is not None:
        guid_tail = bytes.fromhex("000000001000800000aa00389b71")
        body += struct.pack("<HHI", 22, bits, 0) + struct.pack("<H", extensible_sub) + guid_tail
    return body


def _samples(n, channels, seed):
    rng = np.random.default_rng(seed)
    x = 0.7 * rng.uniform(-1.0, 1.0, size=(n, channels))
    return x if channels > 1 else x[:, 0]


@pytest.mark.parametrize("dtype,bits", [("<f4", 32), ("<f8", 64)])
@pytest.mark.parametrize("channels", [1, 2])
def test_ieee_float_wav(tmp_path, dtype, bits, channels):
    x = _samples(1000, channels, bits + channels).astype(dtype)
    path = tmp_path / "f.wav"
    path.write_bytes(_riff(_fmt(3, channels, 22050, bits), x.tobytes(),
                           extra_chunks=b"fact" + struct.pack("<II", 4, 1000)))
    data, sr = audio.decode_wav(path)
    assert sr == 22050 and data.dtype == np.float32
    assert np.array_equal(data, x.astype(np.float32))
    prepared, sr2 = audio.read_audio_file(str(path))
    assert sr2 == 22050 and np.array_equal(prepared, audio.prepare_audio_buffer(x.astype(np.float32)))
    assert np.max(np.abs(prepared)) == 1.0
    assert audio.read_pcm16_file(str(path)) is None         # not the device PCM16 path


@pytest.mark.parametrize("sub,bits", [(1, 16), (1, 24), (3, 32)])
def test_extensible_header(tmp_path, sub, bits):
    rng = np.random.default_rng(bits)
    path = tmp_path / "e.wav"
    if sub == 3:
        x = rng.uniform(-1, 1, size=(500, 2)).astype("<f4")
        raw, want = x.tobytes(), x
    elif bits == 16:
        q = rng.integers(-32768, 32767, size=(500, 2), dtype=np.int16)
        raw, want = q.astype("<i2").tobytes(), q.astype(np.float32) / np.float32(32768.0)
    else:
        q = rng.integers(-(1 << 23), (1 << 23) - 1, size=(500, 2), dtype=np.int32)
        b = q.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]
        raw, want = b.tobytes(), (q.astype(np.float64) / 8388608.0).astype(np.float32)
    path.write_bytes(_riff(_fmt(0xFFFE, 2, 48000, bits, extensible_sub=sub), raw))
    data, sr = audio.decode_wav(path)
    assert sr == 48000 and data.shape == (500, 2)
    assert np.array_equal(data, want)


@pytest.mark.parametrize("width", [1, 2, 3, 4])
def test_integer_pcm_widths_follow_soundfile_scaling(tmp_path, width):
    rng = np.random.default_rng(width)
    n = 777
    if width == 1:
        q = rng.integers(0, 256, size=n, dtype=np.uint8)
        raw, want = q.tobytes(), (q.astype(np.float32) - 128.0) / 128.0
    elif width == 2:
        q = rng.integers(-32768, 32768, size=n).astype("<i2")
        raw, want = q.tobytes(), q.astype(np.float32) / 32768.0
    elif width == 3:
        q = rng.integers(-(1 << 23), 1 << 23, size=n).astype("<i4")
        raw, want = q.view(np.uint8).reshape(-1, 4)[:, :3].tobytes(), (q.astype(np.float64) / 8388608.0).astype(np.float32)
    else:
        q = rng.integers(-(1 << 31), 1 << 31, size=n).astype("<i4")
        raw, want = q.tobytes(), (q.astype(np.float64) / 2147483648.0).astype(np.float32)
    path = tmp_path / "p.wav"
    with wave.open(str(path), "wb") as handle:
        handle.setnchannels(1)
        handle.setsampwidth(width)
        handle.setframerate(16000)
        handle.writeframes(raw)
    data, sr = audio.decode_wav(path)
    assert sr == 16000 and np.array_equal(data, want.astype(np.float32))


def test_undecodable_files_raise_the_decode_error(tmp_path):
    junk = tmp_path / "junk.wav"
    junk.write_bytes(b"ID3\x04" + bytes(64))
    with pytest.raises(audio.AudioDecodeError, match="Could not decode audio file"):
        audio.read_audio_file(str(junk))
    adpcm = tmp_path / "adpcm.wav"
    adpcm.write_bytes(_riff(_fmt(0x0002, 1, 8000, 8), bytes(100)))
    with pytest.raises(audio.AudioDecodeError, match="unsupported WAVE format tag 0x0002"):
        audio.read_audio_file(str(adpcm))
    nodata = tmp_path / "nodata.wav"
    nodata.write_bytes(b"RIFF" + struct.pack("<I", 4 + 8 + 16) + b"WAVE" + b"fmt " + struct.pack("<I", 16) + _fmt(3, 1, 8000, 32))
    with pytest.raises(audio.AudioDecodeError, match="missing fmt or data chunk"):
        audio.read_audio_file(str(nodata))
    silent = tmp_path / "empty.wav"
    silent.write_bytes(_riff(_fmt(3, 1, 8000, 32), b""))
    with pytest.raises(OSError, match="Audio file contains no samples."):
        audio.read_audio_file(str(silent))
    lfs = tmp_path / "lfs.wav"
    lfs.write_bytes(b"version https://git-lfs.github.com/spec/v1\noid sha256:0\n")
    with pytest.raises(audio.AudioIntegrityError):
        audio.read_audio_file(str(lfs))


def test_float_wav_segment_bounds_follow_librosa_load(tmp_path):
    x = _samples(22050 * 2, 1, 5).astype("<f4")
    path = tmp_path / "seg.wav"
    path.write_bytes(_riff(_fmt(3, 1, 22050, 32), x.tobytes()))
    seg, sr = audio.read_audio_file(str(path), start_seconds=0.25, duration_seconds=0.5)
    first, count = int(0.25 * sr), int(0.5 * sr)
    assert np.array_equal(seg, audio.prepare_audio_buffer(x[first:first + count]))
