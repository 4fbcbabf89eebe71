"""Deterministic synthetic RAVDESS-shape audio (no dataset can be downloaded here).

Follows the reference generator scripts/build_synthetic_ravdess_dataset.py:76-141 (mono
PCM16 sines, amplitude 0.15, ``f = 180 + 22*emotion + 7*actor`` Hz, file names
``Actor_{aa}/03-01-{ee}-{ii}-{ss}-{rr}-{aa}.wav``) and extends it as SURVEY.md section 8(d)
specifies so that tuning estimation and the noise floor are not degenerate: an 8-partial
harmonic stack, a 4 Hz raised-cosine amplitude modulation and white noise at -30 dBFS.

Everything returns what ``read_audio_file`` would hand the feature path
(ser/_internal/utils/audio_utils.py:53-60,104-113): PCM16 decoded as ``x / 32768`` in
float32, then peak-normalised over the whole clip.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

RAVDESS_EMOTIONS = {
    1: "neutral", 2: "calm", 3: "happy", 4: "sad",
    5: "angry", 6: "fearful", 7: "disgust", 8: "surprised",
}


@dataclass(frozen=True)
class ClipSpec:
    """One synthetic utterance (RAVDESS file-name fields)."""

    index: int
    actor: int
    emotion: int
    intensity: int = 1
    statement: int = 1
    repetition: int = 1

    @property
    def f0(self) -> float:
        return (
            180.0 + 22.0 * self.emotion + 7.0 * self.actor
            + 3.0 * (self.intensity + 2 * self.statement + 4 * self.repetition)
        )

    @property
    def label(self) -> str:
        return RAVDESS_EMOTIONS[self.emotion]

    @property
    def file_name(self) -> str:
        return (
            f"Actor_{self.actor:02d}/03-01-{self.emotion:02d}-{self.intensity:02d}-"
            f"{self.statement:02d}-{self.repetition:02d}-{self.actor:02d}.wav"
        )


def ravdess_specs(n_clips: int) -> list[ClipSpec]:
    """First ``n_clips`` of the 24-actor x 60-utterance RAVDESS grid, cycled when n_clips > 1440."""
    grid: list[tuple[int, int, int, int, int]] = []
    for actor in range(1, 25):
        for emotion in range(1, 9):
            for intensity in (1, 2):
                if emotion == 1 and intensity == 2:
                    continue  # neutral has no strong intensity
                for statement in (1, 2):
                    for repetition in (1, 2):
                        grid.append((actor, emotion, intensity, statement, repetition))
    specs = []
    for index in range(n_clips):
        actor, emotion, intensity, statement, repetition = grid[index % len(grid)]
        specs.append(ClipSpec(index, actor, emotion, intensity, statement, repetition))
    return specs


def pure_sine_pcm16(sample_rate: int, duration_seconds: float, frequency_hz: float,
                    amplitude: float = 0.15) -> np.ndarray:
    """The reference generator's clip, sample for sample (build_synthetic_ravdess_dataset.py:76-101)."""
    total = int(round(sample_rate * duration_seconds))
    n = np.arange(total, dtype=np.float64)
    value = amplitude * np.sin(2.0 * math.pi * frequency_hz * (n / float(sample_rate)))
    value = np.clip(value, -1.0, 1.0)
    return np.rint(value * 32767).astype(np.int16)  # Python round() is also half-to-even


def clip_pcm16(spec: ClipSpec, sample_rate: int, n_samples: int, *, seed: int = 1234) -> np.ndarray:
    """Harmonic stack x 4 Hz AM + -30 dBFS noise, peak 0.15, as int16 PCM."""
    t = np.arange(n_samples, dtype=np.float64) / float(sample_rate)
    phase = 2.0 * math.pi * spec.f0 * t
    x = np.zeros(n_samples, dtype=np.float64)
    for h in range(1, 9):
        if h * spec.f0 < 0.5 * sample_rate:
            x += (0.6**h) * np.sin(h * phase)
    x *= 0.5 * (1.0 - 0.8 * np.cos(2.0 * math.pi * 4.0 * t))
    rng = np.random.default_rng(seed + spec.index)
    x += (10.0 ** (-30.0 / 20.0)) * rng.standard_normal(n_samples)
    peak = float(np.max(np.abs(x)))
    if peak > 0:
        x *= 0.15 / peak
    return np.rint(x * 32767).astype(np.int16)


def decode_pcm16(pcm: np.ndarray) -> np.ndarray:
    """soundfile's float32 read of PCM16 followed by the reference's peak normalisation."""
    audio = pcm.astype(np.float32) / np.float32(32768.0)
    peak = float(np.max(np.abs(audio))) if audio.size else 0.0
    if peak == 0:
        return np.zeros_like(audio)
    return audio / peak  # float32 / python float stays float32 (audio_utils.py:28-35)


def clip_audio(spec: ClipSpec, sample_rate: int, n_samples: int, *, seed: int = 1234) -> np.ndarray:
    return decode_pcm16(clip_pcm16(spec, sample_rate, n_samples, seed=seed))


def long_recording(sample_rate: int, n_samples: int, *, seed: int = 99, section_seconds: float = 20.0) -> np.ndarray:
    """Config c4: piecewise sections drawn from the clip family, peak-normalised as one file."""
    rng = np.random.default_rng(seed)
    section = int(round(section_seconds * sample_rate))
    pcm = np.empty(n_samples, dtype=np.int16)
    pos = 0
    k = 0
    while pos < n_samples:
        length = min(section, n_samples - pos)
        spec = ClipSpec(index=10_000 + k, actor=int(rng.integers(1, 25)), emotion=int(rng.integers(1, 9)),
                        intensity=int(rng.integers(1, 3)), statement=int(rng.integers(1, 3)),
                        repetition=int(rng.integers(1, 3)))
        pcm[pos : pos + length] = clip_pcm16(spec, sample_rate, length, seed=seed)
        pos += length
        k += 1
    return decode_pcm16(pcm)


def batch_audio_torch(n_clips: int, sample_rate: int, n_samples: int, *, device, seed: int = 1234,
                      first_index: int = 0, chunk: int = 64, return_pcm: bool = False):
    """Same clip family generated on ``device`` with torch, as a (n_clips, n_samples) float32 tensor.

    Used by bench.py for the large configurations (c2/c3), where host-side generation would
    dominate the run.  Noise comes from torch's generator, so values differ from
    ``clip_audio`` (statistics are the same); both bench arms read the same tensor.
    With ``return_pcm`` the int16 PCM the audio was decoded from comes back as well
    (``audio == decode_pcm16(pcm)`` per clip): what the files of the data set hold.
    """
    import torch

    specs = ravdess_specs(first_index + n_clips)[first_index:]
    out = torch.empty((n_clips, n_samples), dtype=torch.float32, device=device)
    out_pcm = torch.empty((n_clips, n_samples), dtype=torch.int16, device=device) if return_pcm else None
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + first_index)
    t = torch.arange(n_samples, dtype=torch.float64, device=device) / float(sample_rate)
    am = 0.5 * (1.0 - 0.8 * torch.cos(2.0 * math.pi * 4.0 * t))
    for lo in range(0, n_clips, chunk):
        hi = min(lo + chunk, n_clips)
        f0 = torch.tensor([s.f0 for s in specs[lo:hi]], dtype=torch.float64, device=device)
        phase = 2.0 * math.pi * f0[:, None] * t[None, :]
        x = torch.zeros((hi - lo, n_samples), dtype=torch.float64, device=device)
        for h in range(1, 9):
            keep = (h * f0 < 0.5 * sample_rate).to(torch.float64)[:, None]
            x += keep * (0.6**h) * torch.sin(h * phase)
        x *= am[None, :]
        x += (10.0 ** (-30.0 / 20.0)) * torch.randn((hi - lo, n_samples), dtype=torch.float64,
                                                    device=device, generator=gen)
        x *= 0.15 / x.abs().amax(dim=1, keepdim=True)
        pcm = torch.round(x * 32767)
        audio = (pcm / 32768.0).to(torch.float32)
        audio = audio / audio.abs().amax(dim=1, keepdim=True)
        out[lo:hi] = audio
        if out_pcm is not None:
            out_pcm[lo:hi] = pcm.to(torch.int16)
    return (out, out_pcm) if return_pcm else out
