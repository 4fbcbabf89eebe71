"""Pageable caller memory (plain numpy, what every caller of the Python mirror passes) travels
through the pinned staging ring of ``csrc/host_stage.h``; pinned memory goes to the copy engine as
it is; ``SERB_STAGE_THREADS=0`` hands pageable pointers to ``cudaMemcpyAsync`` unchanged.  All three
must give bit-identical rows -- staging moves bytes, nothing else -- including files that span several
32 MiB slots, files whose sizes leave alignment gaps in the device layout, and more files than slots.

Reference side: numpy arrays from soundfile / librosa.load (ser/_internal/utils/audio_utils.py:63-113).
"""

from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SR = 16000
BITS = 15          # mfcc + chroma + mel + contrast: the staging path is the same for every flag set


def _pinned(array: np.ndarray) -> np.ndarray:
    import torch

    t = torch.from_numpy(np.ascontiguousarray(array)).pin_memory()
    return t.numpy()


@pytest.fixture()
def unstaged_ctx(monkeypatch):
    from ser_b200 import _native

    monkeypatch.setenv("SERB_STAGE_THREADS", "0")
    ctx = _native.Context(0)
    yield ctx
    ctx.close()


def _pcm_files(rng):
    sizes = [40001, 7, 2049, 123457, 5000, 99999, 31, 64000] * 3
    return [rng.integers(-20000, 20000, size=n).astype(np.int16) for n in sizes]


def test_pcm16_files_pageable_pinned_and_unstaged_rows_are_identical(gpu_ctx, unstaged_ctx):
    rng = np.random.default_rng(5)
    files = _pcm_files(rng)
    usable = [i for i, f in enumerate(files) if f.size >= 2048]
    clip_file = np.asarray(usable, dtype=np.int64)
    starts = np.zeros(len(usable), dtype=np.int64)
    lengths = np.asarray([files[i].size for i in usable], dtype=np.int64)
    staged = gpu_ctx.features_host_pcm16(files, 1, clip_file, starts, lengths, SR, BITS)
    pinned = gpu_ctx.features_host_pcm16([_pinned(f) for f in files], 1, clip_file, starts, lengths, SR, BITS)
    plain = unstaged_ctx.features_host_pcm16(files, 1, clip_file, starts, lengths, SR, BITS)
    assert np.all(np.isfinite(staged))
    np.testing.assert_array_equal(staged, pinned)
    np.testing.assert_array_equal(staged, plain)


def test_one_file_larger_than_a_staging_slot(gpu_ctx, unstaged_ctx):
    # 20 M int16 samples = 40 MB: two slots; windows at both ends and across the slot boundary
    rng = np.random.default_rng(6)
    big = rng.integers(-12000, 12000, size=20_000_000).astype(np.int16)
    boundary = (32 << 20) // 2
    starts = np.asarray([0, boundary - 24000, boundary - 1, big.size - 48000], dtype=np.int64)
    lengths = np.full(starts.size, 48000, dtype=np.int64)
    clip_file = np.zeros(starts.size, dtype=np.int64)
    staged = gpu_ctx.features_host_pcm16([big], 1, clip_file, starts, lengths, SR, BITS)
    plain = unstaged_ctx.features_host_pcm16([big], 1, clip_file, starts, lengths, SR, BITS)
    np.testing.assert_array_equal(staged, plain)


def test_float32_buffer_and_clip_list_entries(gpu_ctx, unstaged_ctx):
    rng = np.random.default_rng(7)
    # one buffer of 10 M samples (40 MB: two pieces, two slots) with windows over the piece boundary
    wave = (0.3 * rng.standard_normal(10_000_000)).astype(np.float32)
    piece = 8 << 20
    starts = np.asarray([0, piece - 30000, piece - 2048, wave.size - 64000], dtype=np.int64)
    lengths = np.asarray([64000, 60000, 4096, 64000], dtype=np.int64)
    staged = gpu_ctx.features_host(wave, starts, lengths, SR, BITS)
    pinned = gpu_ctx.features_host(_pinned(wave), starts, lengths, SR, BITS)
    plain = unstaged_ctx.features_host(wave, starts, lengths, SR, BITS)
    np.testing.assert_array_equal(staged, pinned)
    np.testing.assert_array_equal(staged, plain)
    # a list of separately allocated clips of ragged lengths (the training loader's shape)
    clips = [(0.2 * rng.standard_normal(n)).astype(np.float32) for n in (2048, 2051, 48000, 16001, 100003, 4097) * 4]
    staged = gpu_ctx.features_host_clips(clips, SR, BITS)
    plain = unstaged_ctx.features_host_clips(clips, SR, BITS)
    np.testing.assert_array_equal(staged, plain)


def test_non_finite_audio_is_still_reported_through_the_ring(gpu_ctx):
    wave = np.zeros(9_000_000, dtype=np.float32)
    wave[8_500_000] = np.nan                      # in the second piece
    starts = np.asarray([0, 8_400_000], dtype=np.int64)
    lengths = np.asarray([48000, 200_000], dtype=np.int64)
    with pytest.raises(ValueError, match="not finite"):
        gpu_ctx.features_host(wave, starts, lengths, SR, BITS)


def test_two_contexts_on_two_threads_give_the_single_caller_rows(gpu_ctx):
    """A server overlaps calls by giving every host thread its own context (own scratch, own streams);
    contexts share nothing but the device, so rows are bit-identical to one caller's
    (bench.py: e2e.two_callers; INTEGRATION.md section on threads)."""
    import threading

    from ser_b200 import _native

    rng = np.random.default_rng(9)
    files = [rng.integers(-15000, 15000, size=n).astype(np.int16) for n in (48000, 80000, 36001, 64000) * 6]
    clip_file = np.arange(len(files), dtype=np.int64)
    starts = np.zeros(len(files), dtype=np.int64)
    lengths = np.asarray([f.size for f in files], dtype=np.int64)
    bits = 0x1F
    expected = gpu_ctx.features_host_pcm16(files, 1, clip_file, starts, lengths, SR, bits)
    other = _native.Context(0)
    out = {}

    def run(name, ctx):
        out[name] = [ctx.features_host_pcm16(files, 1, clip_file, starts, lengths, SR, bits) for _ in range(4)]

    try:
        threads = [threading.Thread(target=run, args=("a", gpu_ctx)), threading.Thread(target=run, args=("b", other))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    finally:
        other.close()
    for rows in out["a"] + out["b"]:
        np.testing.assert_array_equal(rows, expected)


@pytest.mark.parametrize("sr", [16000, 48000])
def test_device_entry_writes_its_rows_and_nothing_else(gpu_ctx, sr):
    """compute-sanitizer is not available on the pool, so the caller-owned buffers get guard regions: the
    device entry must fill exactly rows [0, n) of ``d_out`` and leave the waveform and the bytes around the
    output untouched, for a ragged batch that exercises partial tiles, partial constant-Q blocks (both
    kernel variants) and short clips."""
    import torch

    rng = np.random.default_rng(21)
    sizes = [2048, 2049, 4095, 7777, sr, 3 * sr + 123, 513, 100, 5 * sr, 2 * sr - 1]
    clips = [(0.25 * rng.standard_normal(n)).astype(np.float32) for n in sizes]
    starts = np.cumsum([0] + sizes[:-1]).astype(np.int64)
    lengths = np.asarray(sizes, dtype=np.int64)
    flat = np.concatenate(clips)
    guard = 1024
    wave = torch.full((flat.size + 2 * guard,), 7.5, dtype=torch.float32, device="cuda")
    wave[guard:guard + flat.size] = torch.from_numpy(flat).cuda()
    before = wave.clone()
    n, dim, pad_rows = len(sizes), 193, 3
    out = torch.full(((n + 2 * pad_rows) * dim,), -123.0, dtype=torch.float32, device="cuda")
    body = out[pad_rows * dim:(pad_rows + n) * dim]
    torch.cuda.synchronize()
    gpu_ctx.features_device(wave[guard:].data_ptr(), flat.size, starts, lengths, sr, 0x1F, body.data_ptr(), 0)
    gpu_ctx.features_device_check(0)
    torch.cuda.synchronize()
    assert torch.equal(wave, before)
    assert bool((out[: pad_rows * dim] == -123.0).all()) and bool((out[(pad_rows + n) * dim:] == -123.0).all())
    rows = body.reshape(n, dim).cpu().numpy()
    assert np.all(np.isfinite(rows))
    np.testing.assert_array_equal(rows, gpu_ctx.features_host(flat, starts, lengths, sr, 0x1F))
