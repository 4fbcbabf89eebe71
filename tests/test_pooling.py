"""Temporal pooling operators against fixtures produced by the reference's own
ser/_internal/pool modules (tests/golden/make_pooling_golden.py).  Window generation is host
logic (CPU test); the segmented mean/std reduction runs on the GPU and must be bit-identical."""

from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def pooling_golden():
    with np.load(REPO / "tests" / "golden" / "pooling_golden.npz", allow_pickle=False) as data:
        return {key: data[key] for key in data.files}


def _encoded(g, name):
    from ser_b200.backend import EncodedSequence

    return EncodedSequence(embeddings=g[f"{name}/embeddings"], frame_start_seconds=g[f"{name}/starts"],
                           frame_end_seconds=g[f"{name}/ends"], backend_id="handcrafted")


def _windows(g, name):
    from ser_b200.pooling import temporal_pooling_windows

    size, stride = g[f"{name}/config"]
    return temporal_pooling_windows(_encoded(g, name), window_size_seconds=float(size),
                                    window_stride_seconds=float(stride))


def test_temporal_pooling_windows_match_reference(pooling_golden):
    for name in pooling_golden["names"].tolist():
        windows = _windows(pooling_golden, name)
        np.testing.assert_array_equal([w.start_seconds for w in windows], pooling_golden[f"{name}/win_starts"])
        np.testing.assert_array_equal([w.end_seconds for w in windows], pooling_golden[f"{name}/win_ends"])


def test_frame_ranges_select_the_overlap_mask(pooling_golden):
    from ser_b200.backend import PoolingWindow, overlap_frame_mask
    from ser_b200.pooling import frame_ranges

    for name in pooling_golden["names"].tolist():
        encoded = _encoded(pooling_golden, name)
        windows = _windows(pooling_golden, name)
        lo, hi = frame_ranges(encoded, windows)
        for a, b, w in zip(lo, hi, windows):
            mask = overlap_frame_mask(encoded, w)
            assert mask[a:b].all() and mask.sum() == b - a
    encoded = _encoded(pooling_golden, "fast_like")
    with pytest.raises(ValueError, match="outside encoded sequence range"):
        frame_ranges(encoded, [PoolingWindow(start_seconds=1.0, end_seconds=1.0e6)])


def test_window_argument_validation(pooling_golden):
    from ser_b200.pooling import temporal_pooling_windows

    encoded = _encoded(pooling_golden, "fast_like")
    with pytest.raises(ValueError, match="window_size_seconds"):
        temporal_pooling_windows(encoded, window_size_seconds=0.0, window_stride_seconds=1.0)
    with pytest.raises(ValueError, match="window_stride_seconds"):
        temporal_pooling_windows(encoded, window_size_seconds=1.0, window_stride_seconds=float("nan"))


@pytest.mark.gpu
def test_mean_std_pool_is_bit_identical_to_reference(pooling_golden):
    from ser_b200.handcrafted import HandcraftedBackend
    from ser_b200.pooling import mean_std_pool

    for name in pooling_golden["names"].tolist():
        encoded = _encoded(pooling_golden, name)
        windows = _windows(pooling_golden, name)
        got = mean_std_pool(encoded, windows)
        assert got.dtype == np.float64
        np.testing.assert_array_equal(got, pooling_golden[f"{name}/mean_std"])
        np.testing.assert_array_equal(HandcraftedBackend().pool(encoded, windows), pooling_golden[f"{name}/mean"])
    assert mean_std_pool(_encoded(pooling_golden, "dense"), []).shape == (0, 128)
