"""Restatement of the librosa.util helpers the fast-profile path reaches (librosa 0.11.0).

TEST INFRASTRUCTURE (oracle).  Each function names the upstream function it follows;
the reference call sites that reach it are in ser/_internal/utils/dsp.py:100-141.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse

MAX_MEM_BLOCK = 2**8 * 2**10


class LibrosaError(Exception):
    """Root of the shim's exception tree (librosa.util.exceptions.LibrosaError)."""


class ParameterError(LibrosaError):
    """Invalid-parameter error (librosa.util.exceptions.ParameterError)."""


def dtype_r2c(d, *, default=np.complex64):
    """librosa.util.dtype_r2c: real dtype -> complex dtype of the same precision."""
    mapping = {
        np.dtype(np.float32): np.complex64,
        np.dtype(np.float64): np.complex128,
    }
    dt = np.dtype(d)
    if dt.kind == "c":
        return dt
    return np.dtype(mapping.get(dt, default))


def dtype_c2r(d, *, default=np.float32):
    """librosa.util.dtype_c2r."""
    mapping = {
        np.dtype(np.complex64): np.float32,
        np.dtype(np.complex128): np.float64,
    }
    dt = np.dtype(d)
    if dt.kind == "f":
        return dt
    return np.dtype(mapping.get(dt, default))


def tiny(x):
    """librosa.util.tiny: smallest positive normal number of x's float type."""
    x = np.asarray(x)
    if np.issubdtype(x.dtype, np.floating) or np.issubdtype(x.dtype, np.complexfloating):
        dtype = x.dtype
    else:
        dtype = np.dtype(np.float32)
    return np.finfo(dtype).tiny


def pad_center(data, *, size, axis=-1, **kwargs):
    """librosa.util.pad_center."""
    kwargs.setdefault("mode", "constant")
    n = data.shape[axis]
    lpad = int((size - n) // 2)
    lengths = [(0, 0)] * data.ndim
    lengths[axis] = (lpad, int(size - n - lpad))
    if lpad < 0:
        raise ParameterError(f"Target size ({size:d}) must be at least input size ({n:d})")
    return np.pad(data, lengths, **kwargs)


def fix_length(data, *, size, axis=-1, **kwargs):
    """librosa.util.fix_length: trim or zero-pad to exactly ``size``."""
    kwargs.setdefault("mode", "constant")
    n = data.shape[axis]
    if n > size:
        slices = [slice(None)] * data.ndim
        slices[axis] = slice(0, size)
        return data[tuple(slices)]
    if n < size:
        lengths = [(0, 0)] * data.ndim
        lengths[axis] = (0, size - n)
        return np.pad(data, lengths, **kwargs)
    return data


def frame(x, *, frame_length, hop_length):
    """librosa.util.frame for 1-D input, frames along the last axis -> (frame_length, n_frames)."""
    x = np.asarray(x)
    if x.shape[-1] < frame_length:
        raise ParameterError(
            f"Input is too short (n={x.shape[-1]:d}) for frame_length={frame_length:d}"
        )
    n_frames = 1 + (x.shape[-1] - frame_length) // hop_length
    idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n_frames)[None, :]
    return x[idx]


def normalize(S, *, norm=np.inf, axis=0, threshold=None, fill=None):
    """librosa.util.normalize (norm in {inf, 1, 2, None}; fill=None)."""
    if threshold is None:
        threshold = tiny(S)
    elif threshold <= 0:
        raise ParameterError(f"threshold={threshold} must be strictly positive")
    if fill not in [None, False, True]:
        raise ParameterError(f"fill={fill} must be None or boolean")
    if not np.all(np.isfinite(S)):
        raise ParameterError("Input must be finite")

    mag = np.abs(S).astype(float)
    fill_norm = 1
    if norm is None:
        return S
    if norm == np.inf:
        length = np.max(mag, axis=axis, keepdims=True)
    elif norm == -np.inf:
        length = np.min(mag, axis=axis, keepdims=True)
    elif norm == 0:
        if fill is True:
            raise ParameterError("Cannot normalize with norm=0 and fill=True")
        length = np.sum(mag > 0, axis=axis, keepdims=True, dtype=mag.dtype)
    elif np.issubdtype(type(norm), np.number) and norm > 0:
        length = np.sum(mag**norm, axis=axis, keepdims=True) ** (1.0 / norm)
        if axis is None:
            fill_norm = mag.size ** (-1.0 / norm)
        else:
            fill_norm = mag.shape[axis] ** (-1.0 / norm)
    else:
        raise ParameterError(f"Unsupported norm: {repr(norm)}")

    small_idx = length < threshold
    Snorm = np.empty_like(S)
    if fill is None:
        length[small_idx] = 1.0
        Snorm[:] = S / length
    elif fill:
        length[small_idx] = np.nan
        Snorm[:] = S / length
        Snorm[np.isnan(Snorm)] = fill_norm
    else:
        length[small_idx] = np.inf
        Snorm[:] = S / length
    return Snorm


def localmax(x, *, axis=0):
    """librosa.util.localmax: x[k] > x[k-1] and x[k] >= x[k+1]; first False, last x[-1] > x[-2]."""
    xi = np.moveaxis(np.asarray(x), axis, -1)
    out = np.zeros(xi.shape, dtype=bool)
    if xi.shape[-1] >= 3:
        out[..., 1:-1] = (xi[..., 1:-1] > xi[..., :-2]) & (xi[..., 1:-1] >= xi[..., 2:])
    if xi.shape[-1] >= 2:
        out[..., -1] = xi[..., -1] > xi[..., -2]
    return np.moveaxis(out, -1, axis)


def softmask(X, X_ref, *, power=1, split_zeros=False):
    """librosa.util.softmask."""
    if X.shape != X_ref.shape:
        raise ParameterError(f"Shape mismatch: {X.shape}!={X_ref.shape}")
    if np.any(X < 0) or np.any(X_ref < 0):
        raise ParameterError("X and X_ref must be non-negative")
    if power <= 0:
        raise ParameterError("power must be strictly positive")
    dtype = X.dtype
    if not np.issubdtype(dtype, np.floating):
        dtype = np.float32
    Z = np.maximum(X, X_ref).astype(dtype)
    bad_idx = Z < np.finfo(dtype).tiny
    Z[bad_idx] = 1
    if np.isfinite(power):
        mask = (X / Z) ** power
        ref_mask = (X_ref / Z) ** power
        good_idx = ~bad_idx
        mask[good_idx] /= mask[good_idx] + ref_mask[good_idx]
        if split_zeros:
            mask[bad_idx] = 0.5
        else:
            mask[bad_idx] = 0.0
    else:
        mask = X > X_ref
    return mask


def sparsify_rows(x, *, quantile=0.01, dtype=None):
    """librosa.util.sparsify_rows: drop the smallest entries holding < quantile of each row's L1 mass."""
    if x.ndim == 1:
        x = x.reshape((1, -1))
    elif x.ndim > 2:
        raise ParameterError(f"Input must have 2 or fewer dimensions. Provided x.shape={x.shape}.")
    if not 0.0 <= quantile < 1:
        raise ParameterError(f"Invalid quantile {quantile:.2f}")
    if dtype is None:
        dtype = x.dtype
    x_sparse = scipy.sparse.lil_matrix(x.shape, dtype=dtype)
    mags = np.abs(x)
    norms = np.sum(mags, axis=1, keepdims=True)
    mag_sort = np.sort(mags, axis=1)
    cumulative_mag = np.cumsum(mag_sort / norms, axis=1)
    threshold_idx = np.argmin(cumulative_mag < quantile, axis=1)
    for i, j in enumerate(threshold_idx):
        idx = np.where(mags[i] >= mag_sort[i, j])
        x_sparse[i, idx] = x[i, idx]
    return x_sparse.tocsr()


def phasor(angles):
    """librosa.util.phasor: cos(angles) + 1j sin(angles)."""
    angles = np.asarray(angles, dtype=float)
    return np.cos(angles) + 1j * np.sin(angles)


def valid_audio(y):
    """librosa.util.valid_audio (the checks that can fire on this path)."""
    if not isinstance(y, np.ndarray):
        raise ParameterError("Audio data must be of type numpy.ndarray")
    if not np.issubdtype(y.dtype, np.floating):
        raise ParameterError("Audio data must be floating-point")
    if y.ndim == 0 or y.shape[-1] == 0:
        raise ParameterError("Audio data must be at least one-dimensional and non-empty")
    if not np.isfinite(y).all():
        raise ParameterError("Audio buffer is not finite everywhere")
    return True
