"""ctypes binding of libser_b200.so (the C ABI declared in include/ser_b200.h).

This is the only place the package touches native code.  There is no CPU fallback: if the
library is missing or no CUDA device is visible every compute call raises.
"""

from __future__ import annotations

import ctypes
import threading
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint32, c_void_p
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "libser_b200.so"

FLAG_MFCC, FLAG_CHROMA, FLAG_MEL, FLAG_CONTRAST, FLAG_TONNETZ = 1, 2, 4, 8, 16
FLAG_ALL = 31
HAS_TONNETZ = True   # the tonnetz chain (harmonic + chroma_cqt) is implemented by this build

OUT_SOFTMAX, OUT_LOGISTIC = 0, 1

# every symbol include/ser_b200.h declares: (restype, argtypes)
_P = c_void_p
SIGNATURES: dict[str, tuple] = {
    "serb_version": (c_char_p, []),
    "serb_device_count": (c_int, []),
    "serb_ctx_create": (c_int, [c_int, POINTER(_P)]),
    "serb_ctx_destroy": (None, [_P]),
    "serb_last_error": (c_char_p, [_P]),
    "serb_feature_dim": (c_int, [c_uint32]),
    "serb_features_device": (c_int, [_P, _P, c_int64, _P, _P, c_int64, c_int32, c_uint32, _P, _P]),
    "serb_features_host": (c_int, [_P, _P, c_int64, _P, _P, c_int64, c_int32, c_uint32, _P]),
    "serb_features_host_clips": (c_int, [_P, _P, _P, c_int64, c_int32, c_uint32, _P]),
    "serb_features_device_check": (c_int, [_P, _P]),
    "serb_features_host_pcm16": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, c_int64, c_int32, c_uint32, _P]),
    "serb_infer_host_pcm16": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, c_int64, c_int32, c_uint32, _P, _P, _P]),
    "serb_mlp_load": (c_int, [_P, c_int32, c_int32, c_int32, _P, _P, _P, _P, _P, _P, c_int32]),
    "serb_mlp_n_classes": (c_int, [_P]),
    "serb_mlp_predict_host": (c_int, [_P, _P, c_int64, _P, _P]),
    "serb_mlp_predict_device": (c_int, [_P, _P, c_int64, _P, _P, _P]),
    "serb_infer_host": (c_int, [_P, _P, c_int64, _P, _P, c_int64, c_int32, c_uint32, _P, _P, _P]),
    "serb_pool_frames_host": (c_int, [_P, _P, c_int64, c_int32, _P, _P, c_int64, c_int32, _P]),
    "serb_prepare_pcm16_host": (c_int, [_P, _P, c_int64, _P]),
    "serb_prepare_pcm16_device": (c_int, [_P, _P, c_int64, _P, _P]),
    "serb_prepare_pcm16_files_host": (c_int, [_P, _P, _P, _P, c_int64, _P]),
    "serb_debug_filterbank": (c_int, [c_int32, c_int32, c_int32, c_int32, _P]),
    "serb_debug_stft_host": (c_int, [_P, _P, c_int64, _P, c_int64]),
    "serb_debug_last_tuning": (c_int, [_P, _P, c_int64]),
    "serb_debug_tonnetz_stages": (c_int, [_P, _P, c_int64, c_int32, _P, _P, _P, c_int64, _P, _P]),
    "serb_debug_cqt_plan": (c_int, [c_int32, _P]),
    "serb_debug_cqt_basis": (c_int, [c_int32, c_int32, c_int32, _P, _P]),
    "serb_debug_cqt_set_basis": (c_int, [c_int32, c_int32, c_int32, _P, _P]),
    "serb_debug_decimation_taps": (c_int, [c_int32, _P, c_int32]),
    "serb_debug_launch_count": (c_int64, [_P]),
    "serb_debug_fp32_peak": (c_int, [_P, _P]),
    "serb_debug_last_compute_ms": (c_float, [_P]),
    "serb_debug_set_profile": (c_int, [_P, c_int32]),
    "serb_debug_kernel_ms": (c_int, [_P, c_int32, _P, _P]),
}

_lib = None
_lib_lock = threading.Lock()


class ParameterError(Exception):
    """Counterpart of librosa.util.exceptions.ParameterError raised by the reference path
    (e.g. spectral_contrast's Nyquist check, ser/_internal/utils/dsp.py:127-136)."""


class UnsupportedConfigurationError(NotImplementedError):
    """SERB_ERR_UNSUPPORTED: the input is valid for the reference but outside what the CUDA path
    implements -- today the ~5.5 % of sample rates (e.g. 20.5-20.8, 33.1-33.6, 40.9-41.6,
    66.1-67.2 kHz; none of 8 / 11.025 / 16 / 22.05 / 24 / 32 / 44.1 / 48 / 96 kHz) whose constant-Q
    plan (early-downsampling count or FFT size) would depend on the per-clip tuning estimate.
    Deliberately NOT a ValueError: callers that treat ValueError as "bad audio" (the reference's
    run_fast_inference does) must not mistake it for one; route such files to the reference's own
    CPU path or resample them."""


def load_library() -> ctypes.CDLL:
    """Loads libser_b200.so once; raises RuntimeError with build instructions if it is absent."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m ser_b200.build` "
                "(ser_b200 has no CPU fallback)"
            )
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here means header and library disagree
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return lib


def _ptr(array: np.ndarray | None):
    return None if array is None else array.ctypes.data_as(c_void_p)


def _raise(lib, ctx, code: int) -> None:
    message = lib.serb_last_error(ctx)
    text = message.decode("utf-8", "replace") if message else f"ser_b200 error {code}"
    if code == -5:
        raise ParameterError(text)
    if code == -7:
        raise UnsupportedConfigurationError(text)
    if code < 0:
        raise ValueError(text)
    raise RuntimeError(text)


def _clip_arrays(starts, lengths) -> tuple[np.ndarray, np.ndarray]:
    """``starts`` / ``lengths`` as the contiguous int64 vectors the C entries read ``n_clips`` items of;
    a length mismatch would be an out-of-bounds read on the other side, so it is refused here."""
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    if starts.ndim != 1 or lengths.ndim != 1 or starts.size != lengths.size:
        raise ValueError(f"starts and lengths must be 1-D with one entry per clip, got {starts.shape} and {lengths.shape}")
    return starts, lengths


def _mlp_arrays(mean, scale, w1, b1, w2, b2) -> list[np.ndarray]:
    """The six weight arrays as contiguous float64 with consistent shapes (serb_mlp_load reads
    n_in, n_in, n_in x n_hidden, n_hidden, n_hidden x n_out and n_out doubles from them)."""
    arrays = [np.ascontiguousarray(a, dtype=np.float64) for a in (mean, scale, w1, b1, w2, b2)]
    mean, scale, w1, b1, w2, b2 = arrays
    if w1.ndim != 2 or w2.ndim != 2:
        raise ValueError("classifier weight matrices must be 2-D (n_in x n_hidden, n_hidden x n_out)")
    n_in, n_hidden = w1.shape
    if mean.shape != (n_in,) or scale.shape != (n_in,) or b1.shape != (n_hidden,) or \
            w2.shape[0] != n_hidden or b2.shape != (w2.shape[1],):
        raise ValueError(
            f"inconsistent classifier shapes: mean {mean.shape}, scale {scale.shape}, w1 {w1.shape}, b1 {b1.shape}, "
            f"w2 {w2.shape}, b2 {b2.shape}")
    return arrays


class Context:
    """One libser_b200 context (one CUDA device).  Thread-safe; calls are serialised natively."""

    def __init__(self, device: int = 0) -> None:
        self._lib = load_library()
        handle = c_void_p()
        code = self._lib.serb_ctx_create(int(device), ctypes.byref(handle))
        if code != 0:
            _raise(self._lib, None, code)
        self._handle = handle
        self.device = int(device)
        self._mlp_classes: int = 0
        self._mlp_n_in: int = 0

    def close(self) -> None:
        if getattr(self, "_handle", None):
            self._lib.serb_ctx_destroy(self._handle)
            self._handle = None

    def __del__(self) -> None:  # pragma: no cover - interpreter shutdown ordering
        try:
            self.close()
        except Exception:
            pass

    def _check(self, code: int) -> None:
        if code != 0:
            _raise(self._lib, self._handle, code)

    # ---- features ------------------------------------------------------------------------
    def features_host(self, wave: np.ndarray, starts: np.ndarray, lengths: np.ndarray,
                      sample_rate: int, flag_bits: int) -> np.ndarray:
        """Ragged batch over a host waveform -> (n_clips, dim) float32."""
        wave = np.ascontiguousarray(wave, dtype=np.float32)
        starts, lengths = _clip_arrays(starts, lengths)
        dim = self._lib.serb_feature_dim(flag_bits)
        out = np.empty((starts.size, dim), dtype=np.float32)
        self._check(self._lib.serb_features_host(
            self._handle, _ptr(wave), wave.size, _ptr(starts), _ptr(lengths), starts.size,
            int(sample_rate), int(flag_bits), _ptr(out)))
        return out

    def features_host_clips(self, clips: list[np.ndarray], sample_rate: int, flag_bits: int) -> np.ndarray:
        """One row per clip; every clip is its own contiguous float32 array (no packing copy)."""
        arrays = [np.ascontiguousarray(c, dtype=np.float32) for c in clips]
        n = len(arrays)
        pointers = (c_void_p * max(n, 1))(*[a.ctypes.data for a in arrays])
        lengths = np.asarray([a.size for a in arrays], dtype=np.int64)
        dim = self._lib.serb_feature_dim(flag_bits)
        out = np.empty((n, dim), dtype=np.float32)
        self._check(self._lib.serb_features_host_clips(
            self._handle, ctypes.cast(pointers, c_void_p), _ptr(lengths), n, int(sample_rate), int(flag_bits), _ptr(out)))
        return out

    def features_device(self, d_wave_ptr: int, n_wave: int, starts: np.ndarray, lengths: np.ndarray,
                        sample_rate: int, flag_bits: int, d_out_ptr: int, stream: int = 0) -> None:
        """Device pointers in, device pointer out; enqueues on ``stream`` (0 = the context's)."""
        starts, lengths = _clip_arrays(starts, lengths)
        self._check(self._lib.serb_features_device(
            self._handle, c_void_p(d_wave_ptr), int(n_wave), _ptr(starts), _ptr(lengths), starts.size,
            int(sample_rate), int(flag_bits), c_void_p(d_out_ptr), c_void_p(stream)))

    def features_device_check(self, stream: int = 0) -> None:
        """Synchronises ``stream`` and raises the reference's ValueError if the last device-entry
        chain staged a non-finite sample."""
        self._check(self._lib.serb_features_device_check(self._handle, c_void_p(stream)))

    # ---- PCM16 files (N1) ------------------------------------------------------------------
    @staticmethod
    def _pcm16_args(files, channels, clip_file, clip_starts, clip_lengths):
        if isinstance(files, np.ndarray) and files.ndim == 2:
            # a batch of equal-length files in one C-contiguous (n_files, frames * channels) int16 array:
            # pointers and sizes are computed vectorised (no per-file Python work in the caller's loop)
            if files.dtype != np.int16:
                raise TypeError(f"PCM16 files must be int16 arrays, got {files.dtype}")
            block = np.ascontiguousarray(files)
            n, row = block.shape
            ch = np.full(n, int(channels), dtype=np.int32) if isinstance(channels, int) else \
                np.asarray(channels, dtype=np.int32).reshape(-1)
            if ch.size != n or np.any(ch < 1) or np.any(row % ch != 0):
                raise ValueError("PCM16 file size is not a multiple of its channel count")
            frames = (row // ch).astype(np.int64)
            pointers = (block.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(row * 2)).astype(np.uint64)
            clip_file = np.ascontiguousarray(clip_file, dtype=np.int64)
            clip_starts = np.ascontiguousarray(clip_starts, dtype=np.int64)
            clip_lengths = np.ascontiguousarray(clip_lengths, dtype=np.int64)
            if not (clip_file.size == clip_starts.size == clip_lengths.size):
                raise ValueError("clip_file, clip_starts and clip_lengths must have one entry per clip")
            return [block, pointers], pointers.ctypes.data_as(c_void_p), frames, ch, clip_file, clip_starts, clip_lengths
        arrays = []
        for f in files:
            a = np.asarray(f)
            if a.dtype != np.int16:
                raise TypeError(f"PCM16 files must be int16 arrays, got {a.dtype}")
            arrays.append(np.ascontiguousarray(a))
        n = len(arrays)
        if isinstance(channels, int):
            channels = [channels] * n
        ch = np.asarray(channels, dtype=np.int32).reshape(-1)
        if ch.size != n:
            raise ValueError("one channel count per file")
        frames = np.asarray([a.size // max(int(c), 1) for a, c in zip(arrays, ch)], dtype=np.int64)
        for a, c, fr in zip(arrays, ch, frames):
            if c < 1 or a.size != fr * c:
                raise ValueError("PCM16 file size is not a multiple of its channel count")
        pointers = (c_void_p * max(n, 1))(*[a.ctypes.data for a in arrays])
        clip_file = np.ascontiguousarray(clip_file, dtype=np.int64)
        clip_starts = np.ascontiguousarray(clip_starts, dtype=np.int64)
        clip_lengths = np.ascontiguousarray(clip_lengths, dtype=np.int64)
        if not (clip_file.size == clip_starts.size == clip_lengths.size):
            raise ValueError("clip_file, clip_starts and clip_lengths must have one entry per clip")
        return arrays, pointers, frames, ch, clip_file, clip_starts, clip_lengths

    def features_host_pcm16(self, files, channels, clip_file, clip_starts, clip_lengths,
                            sample_rate: int, flag_bits: int) -> np.ndarray:
        """Ragged batch over int16 PCM files prepared on the device (x / 32768, channel mean,
        per-file peak normalisation) -> (n_clips, dim) float32.  ``files[f]`` is a 1-D int16 array of
        ``frames * channels[f]`` interleaved samples; clip i is frames
        ``[clip_starts[i], clip_starts[i] + clip_lengths[i])`` of file ``clip_file[i]``."""
        keep, pointers, frames, ch, cf, cs, cl = self._pcm16_args(files, channels, clip_file, clip_starts, clip_lengths)
        dim = self._lib.serb_feature_dim(flag_bits)
        out = np.empty((cf.size, dim), dtype=np.float32)
        self._check(self._lib.serb_features_host_pcm16(
            self._handle, ctypes.cast(pointers, c_void_p), _ptr(frames), _ptr(ch), frames.size, _ptr(cf), _ptr(cs), _ptr(cl),
            cf.size, int(sample_rate), int(flag_bits), _ptr(out)))
        del keep
        return out

    def infer_host_pcm16(self, files, channels, clip_file, clip_starts, clip_lengths, sample_rate: int,
                         flag_bits: int, *, want_features: bool = True):
        keep, pointers, frames, ch, cf, cs, cl = self._pcm16_args(files, channels, clip_file, clip_starts, clip_lengths)
        n = cf.size
        dim = self._lib.serb_feature_dim(flag_bits)
        feats = np.empty((n, dim), dtype=np.float32) if want_features else None
        proba = np.empty((n, max(self._mlp_classes, 1)), dtype=np.float64)
        labels = np.empty(n, dtype=np.int32)
        self._check(self._lib.serb_infer_host_pcm16(
            self._handle, ctypes.cast(pointers, c_void_p), _ptr(frames), _ptr(ch), frames.size, _ptr(cf), _ptr(cs), _ptr(cl),
            n, int(sample_rate), int(flag_bits), _ptr(feats), _ptr(proba), _ptr(labels)))
        del keep
        return feats, proba, labels

    # ---- classifier ----------------------------------------------------------------------
    def mlp_load(self, mean, scale, w1, b1, w2, b2, out_activation: int) -> None:
        arrays = _mlp_arrays(mean, scale, w1, b1, w2, b2)
        n_in, n_hidden = arrays[2].shape
        n_out = arrays[4].shape[1]
        self._check(self._lib.serb_mlp_load(self._handle, n_in, n_hidden, n_out,
                                            *[_ptr(a) for a in arrays], int(out_activation)))
        self._mlp_classes = self._lib.serb_mlp_n_classes(self._handle)
        self._mlp_n_in = int(n_in)

    @property
    def mlp_n_classes(self) -> int:
        return self._mlp_classes

    def mlp_predict_host(self, x: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim != 2 or (self._mlp_n_in and x.shape[1] != self._mlp_n_in):
            raise ValueError(f"X has shape {x.shape}, but the loaded classifier expects (n, {self._mlp_n_in}).")
        n = x.shape[0]
        proba = np.empty((n, max(self._mlp_classes, 1)), dtype=np.float64)
        labels = np.empty(n, dtype=np.int32)
        self._check(self._lib.serb_mlp_predict_host(self._handle, _ptr(x), n, _ptr(proba), _ptr(labels)))
        return proba, labels

    def mlp_predict_device(self, d_x_ptr: int, n: int, d_proba_ptr: int, d_label_ptr: int, stream: int = 0) -> None:
        self._check(self._lib.serb_mlp_predict_device(self._handle, c_void_p(d_x_ptr), int(n),
                                                      c_void_p(d_proba_ptr), c_void_p(d_label_ptr),
                                                      c_void_p(stream)))

    def infer_host(self, wave: np.ndarray, starts: np.ndarray, lengths: np.ndarray, sample_rate: int,
                   flag_bits: int, *, want_features: bool = True):
        wave = np.ascontiguousarray(wave, dtype=np.float32)
        starts, lengths = _clip_arrays(starts, lengths)
        n = starts.size
        dim = self._lib.serb_feature_dim(flag_bits)
        feats = np.empty((n, dim), dtype=np.float32) if want_features else None
        proba = np.empty((n, max(self._mlp_classes, 1)), dtype=np.float64)
        labels = np.empty(n, dtype=np.int32)
        self._check(self._lib.serb_infer_host(
            self._handle, _ptr(wave), wave.size, _ptr(starts), _ptr(lengths), n, int(sample_rate),
            int(flag_bits), _ptr(feats), _ptr(proba), _ptr(labels)))
        return feats, proba, labels

    # ---- pooling -------------------------------------------------------------------------
    def pool_frames_host(self, embeddings: np.ndarray, lo: np.ndarray, hi: np.ndarray, mode: int) -> np.ndarray:
        """Frame ranges [lo, hi) -> float64 rows: mode 0 mean, 1 mean+std, 2 float32 mean (widened)."""
        emb = np.ascontiguousarray(embeddings, dtype=np.float32)
        lo = np.ascontiguousarray(lo, dtype=np.int32)
        hi = np.ascontiguousarray(hi, dtype=np.int32)
        if emb.ndim != 2 or lo.ndim != 1 or lo.shape != hi.shape:
            raise ValueError(f"embeddings must be 2-D and lo / hi 1-D of one length, got {emb.shape}, {lo.shape}, {hi.shape}")
        n_frames, dim = emb.shape
        out = np.empty((lo.size, 2 * dim if mode == 1 else dim), dtype=np.float64)
        self._check(self._lib.serb_pool_frames_host(self._handle, _ptr(emb), n_frames, dim, _ptr(lo), _ptr(hi),
                                                    lo.size, int(mode), _ptr(out)))
        return out

    # ---- audio prep ----------------------------------------------------------------------
    def prepare_pcm16_host(self, pcm: np.ndarray) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        out = np.empty(pcm.size, dtype=np.float32)
        self._check(self._lib.serb_prepare_pcm16_host(self._handle, _ptr(pcm), pcm.size, _ptr(out)))
        return out

    def prepare_pcm16_files_host(self, files, channels) -> list[np.ndarray]:
        """``_prepare_audio_buffer`` of every int16 file on the device -> list of float32 mono arrays."""
        keep, pointers, frames, ch, _, _, _ = self._pcm16_args(files, channels, [], [], [])
        outs = [np.empty(int(fr), dtype=np.float32) for fr in frames]
        out_ptrs = (c_void_p * max(len(outs), 1))(*[o.ctypes.data for o in outs])
        self._check(self._lib.serb_prepare_pcm16_files_host(
            self._handle, ctypes.cast(pointers, c_void_p), _ptr(frames), _ptr(ch), frames.size, ctypes.cast(out_ptrs, c_void_p)))
        del keep
        return outs

    # ---- introspection -------------------------------------------------------------------
    def debug_stft_host(self, wave: np.ndarray) -> np.ndarray:
        wave = np.ascontiguousarray(wave, dtype=np.float32)
        n_cols = 1 + wave.size // 512
        out = np.empty((n_cols, 1025), dtype=np.float32)
        self._check(self._lib.serb_debug_stft_host(self._handle, _ptr(wave), wave.size, _ptr(out), n_cols))
        return out

    def debug_last_tuning(self, n_clips: int) -> np.ndarray:
        out = np.empty(n_clips, dtype=np.int32)
        self._check(self._lib.serb_debug_last_tuning(self._handle, _ptr(out), n_clips))
        return out

    def debug_tonnetz_stages(self, wave: np.ndarray, sample_rate: int) -> dict:
        """Intermediates of the tonnetz chain for one clip: harmonic signal, tuning bin (36 bins per
        octave), scaled constant-Q magnitudes [columns][252], tonnetz means."""
        wave = np.ascontiguousarray(wave, dtype=np.float32)
        plen = max(wave.size, 512)
        yharm = np.empty(plen, dtype=np.float32)
        cap = 4 + plen // 256
        cqmag = np.empty((cap, 252), dtype=np.float32)
        tuning = c_int32(-1)
        cq_cols = c_int32(0)
        ton = np.empty(6, dtype=np.float32)
        self._check(self._lib.serb_debug_tonnetz_stages(
            self._handle, _ptr(wave), wave.size, int(sample_rate), _ptr(yharm), ctypes.byref(tuning),
            _ptr(cqmag), cap, ctypes.byref(cq_cols), _ptr(ton)))
        return {"yharm": yharm, "tuning_index": int(tuning.value), "cqmag": cqmag[: cq_cols.value].copy(),
                "tonnetz": ton}

    @property
    def launch_count(self) -> int:
        return int(self._lib.serb_debug_launch_count(self._handle))

    def set_profile(self, enabled: bool) -> None:
        self._check(self._lib.serb_debug_set_profile(self._handle, 1 if enabled else 0))

    def kernel_ms(self) -> dict[str, tuple[float, int]]:
        """{kernel: (total device ms, launches)} since set_profile(True)."""
        out = {}
        for kind, name in enumerate(("stft", "tuning", "proj", "pool", "short", "mlp", "hpss_harm", "hpss_perc", "istft", "ola",
                                     "decimate", "cqt", "tonnetz", "pcm_prepare")):
            ms = c_double(0.0)
            n = c_int64(0)
            self._check(self._lib.serb_debug_kernel_ms(self._handle, kind, ctypes.byref(ms), ctypes.byref(n)))
            out[name] = (float(ms.value), int(n.value))
        return out

    def fp32_peak_tflops(self) -> float:
        """FFMA ceiling of this GPU measured now (dependent-free chains on every SM)."""
        out = c_double(0.0)
        self._check(self._lib.serb_debug_fp32_peak(self._handle, ctypes.byref(out)))
        return float(out.value)

    def last_compute_ms(self) -> float:
        return float(self._lib.serb_debug_last_compute_ms(self._handle))


def debug_filterbank(kind: int, sample_rate: int, n_fft: int, tuning_index: int = 50) -> np.ndarray:
    """Host-side tables of the library (no GPU needed): 0 mel, 1 chroma, 2 DCT, 3 Hann."""
    lib = load_library()
    n_bins = 1 + n_fft // 2
    shape = {0: (128, n_bins), 1: (12, n_bins), 2: (40, 128), 3: (n_fft,)}[kind]
    out = np.empty(shape, dtype=np.float32)
    code = lib.serb_debug_filterbank(kind, int(sample_rate), int(n_fft), int(tuning_index), _ptr(out))
    if code != 0:
        raise ValueError(f"serb_debug_filterbank failed with {code}")
    return out


def debug_cqt_plan(sample_rate: int) -> dict:
    """Constant-Q plan of the tonnetz chain at one sample rate (host only)."""
    out = np.zeros(10, dtype=np.int32)
    code = load_library().serb_debug_cqt_plan(int(sample_rate), _ptr(out))
    if code != 0:
        raise ValueError(f"serb_debug_cqt_plan failed with {code}")
    return {"status": int(out[0]), "early_factor": int(out[1]), "hop0": int(out[2]), "n_fft": [int(v) for v in out[3:]]}


def debug_cqt_basis(sample_rate: int, tuning_index: int, octave: int) -> tuple[np.ndarray, np.ndarray]:
    """(sparsified FFT-domain basis [36][1 + n_fft/2] complex64, 1/sqrt(length) [36]) of one octave."""
    plan = debug_cqt_plan(sample_rate)
    n_bins = 1 + plan["n_fft"][octave] // 2
    basis = np.zeros((36, n_bins, 2), dtype=np.float32)
    scale = np.zeros(36, dtype=np.float32)
    code = load_library().serb_debug_cqt_basis(int(sample_rate), int(tuning_index), int(octave), _ptr(basis), _ptr(scale))
    if code != 0:
        raise ValueError(f"serb_debug_cqt_basis failed with {code}")
    return basis[..., 0] + 1j * basis[..., 1], scale


def debug_cqt_set_basis(sample_rate: int, tuning_index: int, octave: int) -> tuple[np.ndarray, np.ndarray] | None:
    """The same basis expanded from the column-mapped layout cqtc_kernel reads (None if it does not fit)."""
    plan = debug_cqt_plan(sample_rate)
    n_bins = 1 + plan["n_fft"][octave] // 2
    basis = np.zeros((36, n_bins, 2), dtype=np.float32)
    scale = np.zeros(36, dtype=np.float32)
    code = load_library().serb_debug_cqt_set_basis(int(sample_rate), int(tuning_index), int(octave), _ptr(basis), _ptr(scale))
    if code == -7:
        return None
    if code != 0:
        raise ValueError(f"serb_debug_cqt_set_basis failed with {code}")
    return basis[..., 0] + 1j * basis[..., 1], scale


def debug_decimation_taps(factor: int) -> np.ndarray:
    lib = load_library()
    n = lib.serb_debug_decimation_taps(int(factor), None, 0)
    if n <= 0:
        raise ValueError(f"serb_debug_decimation_taps failed with {n}")
    out = np.zeros(n, dtype=np.float64)
    lib.serb_debug_decimation_taps(int(factor), _ptr(out), n)
    return out


def device_count() -> int:
    return int(load_library().serb_device_count())


_contexts: dict[int, Context] = {}
_contexts_lock = threading.Lock()


def get_context(device: int = 0) -> Context:
    """Lazily created per-device context (CUDA is never touched at import time, so fork-based
    callers such as the reference's mp.Pool path stay safe: ser/_internal/data/data_loader.py:378)."""
    with _contexts_lock:
        ctx = _contexts.get(device)
        if ctx is None:
            ctx = Context(device)
            _contexts[device] = ctx
        return ctx
