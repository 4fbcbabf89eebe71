"""Argument preparation of the ctypes wrapper (ser_b200/_native.py): everything the C entries read through
a bare pointer is sized on the Python side, so a mismatch there must be a ValueError / TypeError here and
never an out-of-bounds read in the library.  Pure host code, no device."""

from __future__ import annotations

import numpy as np
import pytest

from ser_b200 import _native


def test_clip_vectors_must_pair_up():
    starts, lengths = _native._clip_arrays([0, 10, 20], np.asarray([5, 5, 5], dtype=np.int32))
    assert starts.dtype == lengths.dtype == np.int64 and starts.flags.c_contiguous and starts.tolist() == [0, 10, 20]
    assert _native._clip_arrays([], [])[0].shape == (0,)
    assert _native._clip_arrays(np.arange(10)[::2], np.ones(5))[0].flags.c_contiguous       # strided input is packed
    for bad in (([0, 1], [5]), (np.zeros((2, 2)), np.zeros((2, 2))), ([0], [])):
        with pytest.raises(ValueError, match="one entry per clip"):
            _native._clip_arrays(*bad)


def test_classifier_arrays_must_be_consistent():
    rng = np.random.default_rng(0)
    good = dict(mean=rng.random(193), scale=rng.random(193), w1=rng.random((193, 300)).astype(np.float32),
                b1=rng.random(300), w2=np.asfortranarray(rng.random((300, 8))), b2=rng.random(8))
    arrays = _native._mlp_arrays(**good)
    assert all(a.dtype == np.float64 and a.flags.c_contiguous for a in arrays)
    assert np.array_equal(arrays[4], good["w2"])
    binary = dict(good, w2=rng.random((300, 1)), b2=rng.random(1))
    assert _native._mlp_arrays(**binary)[4].shape == (300, 1)
    for key, value in (("mean", rng.random(192)), ("scale", rng.random((193, 1))), ("b1", rng.random(299)),
                       ("w2", rng.random((299, 8))), ("b2", rng.random(7))):
        with pytest.raises(ValueError, match="inconsistent classifier shapes"):
            _native._mlp_arrays(**dict(good, **{key: value}))
    with pytest.raises(ValueError, match="must be 2-D"):
        _native._mlp_arrays(**dict(good, w1=rng.random(193)))


def test_pcm16_file_lists():
    prepare = _native.Context._pcm16_args
    mono = np.arange(10, dtype=np.int16)
    stereo = np.arange(12, dtype=np.int16)
    keep, _ptrs, frames, ch, cf, cs, cl = prepare([mono, stereo], [1, 2], [0, 1, 1], [0, 0, 3], [10, 3, 3])
    assert frames.tolist() == [10, 6] and ch.tolist() == [1, 2] and frames.dtype == np.int64 and ch.dtype == np.int32
    assert cf.dtype == cs.dtype == cl.dtype == np.int64 and len(keep) == 2
    assert prepare([mono, mono], 1, [], [], [])[2].tolist() == [10, 10]                     # one channel count for all
    with pytest.raises(TypeError, match="int16"):
        prepare([mono.astype(np.float32)], 1, [0], [0], [10])
    with pytest.raises(ValueError, match="multiple of its channel count"):
        prepare([np.arange(11, dtype=np.int16)], 2, [0], [0], [5])
    with pytest.raises(ValueError, match="multiple of its channel count"):
        prepare([mono], 0, [0], [0], [5])
    with pytest.raises(ValueError, match="one channel count per file"):
        prepare([mono, mono], [1], [0], [0], [5])
    with pytest.raises(ValueError, match="one entry per clip"):
        prepare([mono], 1, [0, 0], [0], [5])
    # the 2-D fast path: equal-length files in one block, pointers computed vectorised
    block = np.arange(24, dtype=np.int16).reshape(3, 8)
    keep, _ptrs, frames, ch, cf, cs, cl = prepare(block, 2, [0, 1, 2], [0, 0, 0], [4, 4, 4])
    assert frames.tolist() == [4, 4, 4] and ch.tolist() == [2, 2, 2]
    pointers = keep[1]
    assert pointers.tolist() == [block.ctypes.data + 16 * i for i in range(3)]
    with pytest.raises(ValueError, match="multiple of its channel count"):
        prepare(block, 3, [0], [0], [1])
    with pytest.raises(TypeError, match="int16"):
        prepare(block.astype(np.int32), 1, [0], [0], [1])


def test_status_codes_map_to_the_documented_exceptions():
    class FakeLib:
        @staticmethod
        def serb_last_error(_ctx):
            return b"text from the library"

    for code, kind in ((-1, ValueError), (-2, ValueError), (-3, ValueError), (-4, ValueError), (-6, ValueError),
                       (-5, _native.ParameterError), (-7, _native.UnsupportedConfigurationError),
                       (1, RuntimeError), (2, RuntimeError)):
        with pytest.raises(kind, match="text from the library") as err:
            _native._raise(FakeLib, None, code)
        assert type(err.value) is kind
    assert issubclass(_native.UnsupportedConfigurationError, NotImplementedError)
    assert not issubclass(_native.UnsupportedConfigurationError, ValueError)
