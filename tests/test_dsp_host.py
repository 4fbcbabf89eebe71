"""Boundary B1 on the host (ser_b200/dsp.py): the reference's argument checks keep their order and texts
(ser/_internal/utils/dsp.py:85-95) and fire before any device is looked for, so they can be shown on
a machine without a GPU; with every group switched off nothing is computed at all."""

from __future__ import annotations

import numpy as np
import pytest

from ser_b200 import dsp
from ser_b200.config import FeatureFlags, feature_dim, flag_bits

OK = np.zeros(4096, dtype=np.float32)


@pytest.mark.parametrize("audio,sr,text", [
    (OK, 0, "Sample rate must be a positive integer."),
    (OK, -16000, "Sample rate must be a positive integer."),
    (np.zeros((2, 100), dtype=np.float32), 16000, r"Audio must be mono \(1D array\)."),
    (np.zeros((2, 100), dtype=np.float32), 0, "Sample rate must be a positive integer."),       # rate is checked first
    (np.zeros(0, dtype=np.float32), 16000, "Audio contains no samples."),
    (np.array([0.0, np.nan, 0.5], dtype=np.float32), 16000, "Audio buffer is not finite everywhere."),
    (np.array([0.0, np.inf], dtype=np.float64), 16000, "Audio buffer is not finite everywhere."),
])
def test_single_clip_argument_errors(audio, sr, text):
    with pytest.raises(ValueError, match=text):
        dsp.extract_feature_from_signal(audio, sr)


def test_batch_and_ragged_argument_errors():
    with pytest.raises(ValueError, match="Sample rate must be a positive integer."):
        dsp.extract_features_batch([OK], 0)
    with pytest.raises(ValueError, match="Audio contains no samples."):
        dsp.extract_features_batch([OK, np.zeros(0, dtype=np.float32)], 16000)
    with pytest.raises(ValueError, match=r"Audio must be mono \(1D array\)."):
        dsp.extract_features_batch([np.zeros((3, 3), dtype=np.float32)], 16000)
    with pytest.raises(ValueError, match=r"Audio must be mono \(1D array\)."):
        dsp.extract_features_ragged(np.zeros((3, 3), dtype=np.float32), np.zeros(1, np.int64), np.ones(1, np.int64), 16000)
    with pytest.raises(ValueError, match="Sample rate must be a positive integer."):
        dsp.extract_features_pcm16([np.zeros(10, np.int16)], 1, np.zeros(1, np.int64), np.zeros(1, np.int64),
                                   np.full(1, 10, np.int64), 0)


def test_no_groups_selected_returns_empty_without_a_device():
    none = FeatureFlags(mfcc=False, chroma=False, mel=False, contrast=False, tonnetz=False)
    assert feature_dim(none) == 0 and flag_bits(none) == 0
    out = dsp.extract_feature_from_signal(OK, 16000, feature_flags=none)
    assert out.shape == (0,) and out.dtype == np.float64
    assert dsp.extract_features_batch([], 16000).shape == (0, 193)


def test_flag_bits_follow_the_output_order():
    # mfcc 1 | chroma 2 | mel 4 | contrast 8 | tonnetz 16 (config/schema.py:219-227)
    assert flag_bits(FeatureFlags()) == 0x1F and feature_dim(FeatureFlags()) == 193
    for name, bit, dim in (("mfcc", 1, 40), ("chroma", 2, 12), ("mel", 4, 128), ("contrast", 8, 7), ("tonnetz", 16, 6)):
        only = FeatureFlags(**{k: k == name for k in ("mfcc", "chroma", "mel", "contrast", "tonnetz")})
        assert flag_bits(only) == bit and feature_dim(only) == dim
