"""How much do the tonnetz features (and the predicted labels) depend on the exact soxr "HQ" filter?

TEST INFRASTRUCTURE (uses oracle/; CPU only).  Run in any container:

    python scripts/soxr_sensitivity_study.py [--clips 192] [--workers 8] > profiles/r02_soxr_sensitivity.txt

Why: ``librosa.cqt`` halves the signal between octaves with ``res_type="soxr_hq"``
(ser/_internal/utils/dsp.py:140-143 -> librosa.feature.tonnetz -> chroma_cqt -> vqt -> resample).
libsoxr 1.0.0 (uv.lock:2175-2176) is an un-vendored C dependency that cannot be installed here, so
``oracle/shim/librosa/core.py:_soxr_hq_decimation_filter`` and the CUDA decimator restate its
*published* quality recipe (20-bit precision -> 126.4 dB rejection, pass-band end 0.9136 of the
output Nyquist, stop-band at the Nyquist, linear phase, Kaiser-windowed sinc).  The exact tap
count / Kaiser beta libsoxr derives from that recipe are not reproducible bit-for-bit without its
source, so this script measures the spread of the 6 tonnetz features -- and of the labels of a
fitted Pipeline(StandardScaler, MLPClassifier(300)) -- over a family of decimators that all meet
the published spec, plus deliberately out-of-spec ones to show where the tolerance breaks.

Variants (all zero-latency linear-phase FIR, applied as the oracle applies its own):
  default          the oracle's / CUDA kernel's filter since round 2: libsoxr's design procedure as
                   recollected from its filter.c (lsx_design_lpf + lsx_kaiser_params + lsx_make_lpf):
                   Fc = Fs - tr_bw, beta from its cubic fit (13.04), tap count from its attenuation
                   polynomial rounded up to 1 mod 4 (389), window argument 1 / (m/2 + 0.5), no DC
                   renormalisation
  lsx_389_f32      the same taps rounded to float32 and a float32 FFT convolution (libsoxr runs its
                   <= 20-bit recipes in single precision)
  firwin_381       round 1's stand-in: scipy kaiserord / firwin at pass-band end 0.913 (381 taps)
  lsx_385          the same procedure with beta 0.025 lower (the recollected fit's digits are the
                   uncertain part): the tap-count formula then lands on 385
  kaiser_std_385   385 taps, textbook beta = 0.1102 (A - 8.7)
  taps_361/409     -5 % / +7 % length at the default beta (rejection 120 / 135 dB)
  pass_0.900/0.925 pass-band end moved by -1.5 % / +1.2 % of the Nyquist (transition re-centred)
  short_191        OUT of spec: half the length (63 dB) -- what a "cheaper" decimator would do
"""

from __future__ import annotations

import argparse
import math
import os
import sys
import time
import warnings
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

ATT_DB = 21 * 20.0 * math.log10(2.0)            # (bits + 1) * 6.02 dB, bits = 20 for "HQ"
TO_3DB = (1.6e-6 * (20 * 20.0 * math.log10(2.0)) - 7.5e-4) * (20 * 20.0 * math.log10(2.0)) + 0.646
PASS_END = 1.0 - 0.05 / TO_3DB                   # 0.91363 of the output Nyquist


def _bessel_i0(x):
    return np.i0(x)


# libsoxr's cubic fits of the Kaiser beta against attenuation, one row per octave of transition width
# (0.0005 x 2^row), interpolated linearly in log2(width).  RECOLLECTED from libsoxr's filter.c, not
# copied from a source available here: the digits cannot be verified offline, which is exactly why
# the study also runs `lsx_385` with a beta 0.025 lower (that alone moves the tap count 389 -> 385).
_LSX_BETA_ROWS = (
    (-6.784957e-10, 1.02856e-05, 0.1087556, -0.8988365 + 0.001),
    (-6.897885e-10, 1.027433e-05, 0.10876, -0.8994658 + 0.002),
    (-1.000683e-09, 1.030092e-05, 0.1087677, -0.9007898 + 0.003),
    (-3.654474e-10, 1.040631e-05, 0.1087085, -0.8977766 + 0.006),
    (8.106988e-09, 6.983091e-06, 0.1091387, -0.9172048 + 0.015),
    (9.519571e-09, 7.272678e-06, 0.1090068, -0.9140768 + 0.025),
    (-5.626821e-09, 1.342186e-05, 0.1083999, -0.9065452 + 0.05),
    (-9.965946e-08, 5.073548e-05, 0.1040967, -0.7672778 + 0.085),
    (1.604808e-07, -5.856462e-05, 0.1185998, -1.34824 + 0.1),
    (-1.511964e-07, 6.363034e-05, 0.1064627, -0.9876665 + 0.18),
)


def lsx_kaiser_beta(att, tr_bw):
    realm = math.log(tr_bw / 0.0005) / math.log(2.0)
    i0 = min(max(int(realm), 0), len(_LSX_BETA_ROWS) - 1)
    i1 = min(max(1 + int(realm), 0), len(_LSX_BETA_ROWS) - 1)
    b0, b1 = (((c[0] * att + c[1]) * att + c[2]) * att + c[3] for c in (_LSX_BETA_ROWS[i0], _LSX_BETA_ROWS[i1]))
    return b0 + (b1 - b0) * (realm - int(realm))


def lsx_design(pass_end=PASS_END, stop_begin=1.0, att=ATT_DB, beta=None, num_taps=None):
    """Kaiser-windowed sinc the way libsoxr designs its DFT-stage low-pass for a 2:1 decimation.
    Frequencies are fractions of the OUTPUT Nyquist; the filter runs at the input rate (Fn = 2)."""
    fp, fs = pass_end / 2.0, stop_begin / 2.0           # normalised to the input Nyquist
    tr_bw = 0.5 * (fs - fp)
    tr_bw = min(tr_bw, 0.5 * fs)
    fc = fs - tr_bw
    if beta is None:
        beta = lsx_kaiser_beta(att, tr_bw * 0.5 / fc)
    if num_taps is None:
        a = ((0.0007528358 - 1.577737e-05 * beta) * beta + 0.6248022) * beta + 0.06186902
        n = int(math.ceil(a / tr_bw + 1))
        num_taps = (n + 4 - 2) // 4 * 4 + 1            # 1 mod 4
    m = num_taps - 1
    i = np.arange(num_taps, dtype=np.float64)
    z = i - 0.5 * m
    x = z * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        h = np.where(x != 0, np.sin(fc * x) / x, fc)
    y = z / (0.5 * m + 0.5)
    h = h * _bessel_i0(beta * np.sqrt(np.maximum(0.0, 1.0 - y * y))) / _bessel_i0(beta)
    return h


def firwin_design(numtaps=None, pass_end=0.913, stop_begin=1.0, att=21 * 6.0206, beta=None):
    import scipy.signal

    fp, fs = pass_end / 2.0, stop_begin / 2.0
    n, b = scipy.signal.kaiserord(att, fs - fp)
    if numtaps is None:
        numtaps = n + (1 - n % 2)
    if beta is None:
        beta = b
    return scipy.signal.firwin(numtaps, 0.5 * (fp + fs), window=("kaiser", beta), scale=True)


def variants():
    out = {
        "default": (lsx_design(), False),
        "lsx_389_f32": (lsx_design(), True),
        "firwin_381": (firwin_design(), False),
        "lsx_385": (lsx_design(beta=lsx_kaiser_beta(ATT_DB, 0.0225677) - 0.025), False),
        "kaiser_std_385": (lsx_design(beta=0.1102 * (ATT_DB - 8.7), num_taps=385), False),
        "taps_361": (firwin_design(361), False),
        "taps_409": (firwin_design(409), False),
        "pass_0.900": (firwin_design(pass_end=0.900), False),
        "pass_0.925": (firwin_design(pass_end=0.925), False),
        "short_191": (firwin_design(191), False),
    }
    return out


def stopband_db(taps):
    spec = np.abs(np.fft.rfft(taps, 1 << 16))
    freqs = np.arange(spec.size) / (spec.size - 1)       # fraction of the input Nyquist
    return 20 * math.log10(max(spec[freqs >= 0.5].max(), 1e-300)), float(np.max(np.abs(spec[freqs <= 0.33] - 1.0)))


# ---- workers --------------------------------------------------------------------------------
_VARIANTS = None


def _init():
    global _VARIANTS
    _VARIANTS = variants()
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=1)
    except Exception:
        pass


def _decimate_with(taps, f32):
    def decimate(y, factor):
        assert factor == 2
        n_out = int(np.ceil(len(y) / 2))
        half = (len(taps) - 1) // 2
        if f32:
            import scipy.signal

            full = scipy.signal.fftconvolve(np.asarray(y, np.float32), taps.astype(np.float32), mode="full")
        else:
            import scipy.signal

            full = scipy.signal.fftconvolve(np.asarray(y, np.float64), taps, mode="full")
        return full[half + 2 * np.arange(n_out)]

    return decimate


def _window_job(args):
    """One inference window: 187 base dims once, harmonic once, tonnetz per decimator variant."""
    audio, sr = args
    from oracle import ser_oracle
    from oracle.shim import librosa
    from oracle.shim.librosa import core

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        base = ser_oracle.extract_feature_from_signal(audio, sr, feature_flags=ser_oracle.FeatureFlags(tonnetz=False))
        prepared = ser_oracle.pad_audio_for_fft(np.asarray(audio, dtype=np.float32))
        harmonic = librosa.effects.harmonic(prepared)
        rows = {}
        original = core._soxr_hq_decimate
        try:
            for name, (taps, f32) in _VARIANTS.items():
                core._soxr_hq_decimate = _decimate_with(taps, f32)
                rows[name] = np.mean(librosa.feature.tonnetz(y=harmonic, sr=sr), axis=1).astype(np.float64)
        finally:
            core._soxr_hq_decimate = original
    return base, rows


def stress_signals():
    """Non-tonal / non-stationary inputs at other sample rates (3 s windows)."""
    rng = np.random.default_rng(5)
    out = []
    for sr in (16000, 22050, 44100):
        n = 3 * sr
        t = np.arange(n) / sr
        white = rng.standard_normal(n)
        out.append((f"white_noise@{sr}", white, sr))
        out.append((f"am_noise@{sr}", white * (0.5 - 0.5 * np.cos(2 * np.pi * 3 * t)), sr))
        out.append((f"brown_noise@{sr}", np.cumsum(white), sr))
        out.append((f"impulses@{sr}", (np.arange(n) % (sr // 7) == 0).astype(float) + 1e-3 * white, sr))
        out.append((f"square_220@{sr}", np.sign(np.sin(2 * np.pi * 220 * t)) + 1e-3 * white, sr))
        out.append((f"vibrato_voice@{sr}", sum((0.7**h) * np.sin(2 * np.pi * h * (140 * t + 0.6 * np.sin(2 * np.pi * 5.5 * t)))
                                                 for h in range(1, 12)) * (0.6 + 0.4 * np.sin(2 * np.pi * 2 * t)) + 0.01 * white, sr))
    res = []
    for name, x, sr in out:
        x = x - np.mean(x)
        x = (x / np.max(np.abs(x))).astype(np.float32)
        res.append((name, x, sr))
    return res


def scaled_err(a, b):
    floor = np.maximum(np.abs(b), 1e-3 * np.max(np.abs(b), axis=-1, keepdims=True))
    return np.abs(a - b) / np.where(floor == 0, 1.0, floor)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=192)
    ap.add_argument("--workers", type=int, default=min(8, os.cpu_count() or 1))
    args = ap.parse_args()
    import multiprocessing as mp

    from ser_b200 import synth
    from ser_b200.handcrafted import frame_bounds

    t0 = time.time()
    var = variants()
    print("# soxr_hq stand-in sensitivity study (scripts/soxr_sensitivity_study.py)")
    print(f"# published HQ recipe: rejection {ATT_DB:.2f} dB, pass-band end {PASS_END:.5f} x Nyquist, stop-band 1.0 x Nyquist")
    print("\n## decimators")
    print("| variant | taps | stop-band peak (dB) | max pass-band deviation below 0.66 x out-Nyquist |")
    print("|---|---|---|---|")
    for name, (taps, f32) in var.items():
        sb, ripple = stopband_db(taps)
        print(f"| {name}{' (float32 arithmetic)' if f32 else ''} | {len(taps)} | {sb:.1f} | {ripple:.2e} |")

    sr, n = 48000, 168000
    # every 1440 // clips-th clip of the c2 grid so that all 8 emotions and many actors appear
    step = max(1, 1440 // args.clips)
    specs = synth.ravdess_specs(1440)[::step][: args.clips]
    starts, ends = frame_bounds(n, sr, 3, 1)
    jobs, labels = [], []
    for spec in specs:
        audio = synth.clip_audio(spec, sr, n)
        for a, b in zip(starts, ends):
            jobs.append((audio[a:b], sr))
            labels.append(spec.label)
    stress = stress_signals()
    for name, x, s in stress:
        jobs.append((x, s))
    with mp.get_context("fork").Pool(args.workers, initializer=_init) as pool:
        results = pool.map(_window_job, jobs, chunksize=4)
    n_c2 = len(labels)
    base = np.stack([r[0] for r in results[:n_c2]])
    ton = {name: np.stack([r[1][name] for r in results]) for name in var}

    print(f"\n## config c2 windows: {len(specs)} clips x {len(starts)} windows = {n_c2} rows (3 s / 1 s @ 48 kHz)")
    print("tonnetz error of each variant against `default`; scaled = |a-b| / max(|b|, 1e-3 * max|b| over the 6 dims)")
    print("| variant | max abs | max scaled | p99 scaled | median scaled | rows above 1e-4 |")
    print("|---|---|---|---|---|---|")
    ref = ton["default"]
    for name in var:
        if name == "default":
            continue
        se = scaled_err(ton[name][:n_c2], ref[:n_c2])
        ab = np.abs(ton[name][:n_c2] - ref[:n_c2])
        print(f"| {name} | {ab.max():.2e} | {se.max():.2e} | {np.quantile(se, 0.99):.2e} | {np.median(se):.2e} | "
              f"{int(np.sum(se.max(axis=1) > 1e-4))} / {n_c2} |")

    family = [k for k in var if k.startswith(("default", "lsx_", "kaiser_std"))]      # libsoxr's procedure, uncertain digits
    worst_abs = worst_scaled = 0.0
    for i, a in enumerate(family):
        for b in family[i + 1:]:
            worst_abs = max(worst_abs, float(np.abs(ton[a][:n_c2] - ton[b][:n_c2]).max()))
            worst_scaled = max(worst_scaled, float(scaled_err(ton[a][:n_c2], ton[b][:n_c2]).max()))
    print(f"\nin-spec family {family}: worst pairwise max abs {worst_abs:.2e}, max scaled {worst_scaled:.2e}")

    print("\n## stress signals (one 3 s window each): max ABSOLUTE tonnetz difference against `default` "
          "(max |tonnetz| of the signal in the second column; noise-like inputs have near-zero means, so a "
          "relative figure is ill-conditioned there)")
    print("| signal | max abs tonnetz | " + " | ".join(k for k in var if k != "default") + " |")
    print("|---|---|" + "---|" * (len(var) - 1))
    for i, (name, _x, _s) in enumerate(stress):
        row = n_c2 + i
        cells = [f"{np.abs(ton[k][row] - ref[row]).max():.1e}" for k in var if k != "default"]
        print(f"| {name} | {np.abs(ref[row]).max():.2e} | " + " | ".join(cells) + " |")

    # labels: a pipeline of the reference's shape fitted on the default-filter features
    from sklearn.neural_network import MLPClassifier
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler

    def full(name):
        return np.concatenate([base, ton[name][:n_c2]], axis=1).astype(np.float32).astype(np.float64)

    x_ref = full("default")
    y = np.asarray(labels)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # ser/_internal/models/training_support.py:87-106 hyper-parameters
        model = Pipeline([("scaler", StandardScaler()),
                          ("classifier", MLPClassifier(alpha=0.01, batch_size=256, epsilon=1e-8, hidden_layer_sizes=(300,),
                                                       learning_rate="adaptive", max_iter=500, random_state=42))])
        model.fit(x_ref, y)
    p_ref = model.predict_proba(x_ref)
    l_ref = model.predict(x_ref)
    top2 = np.sort(p_ref, axis=1)[:, -2:]
    margin = top2[:, 1] - top2[:, 0]
    print(f"\n## labels: Pipeline(StandardScaler, MLPClassifier(300)) fitted on the {n_c2} default rows "
          f"(train accuracy {np.mean(l_ref == y):.3f}, smallest top-1/top-2 probability margin {margin.min():.3e})")
    print("| variant | labels flipped | max |delta proba| |")
    print("|---|---|---|")
    for name in var:
        if name == "default":
            continue
        xv = full(name)
        lv = model.predict(xv)
        pv = model.predict_proba(xv)
        print(f"| {name} | {int(np.sum(lv != l_ref))} / {n_c2} | {np.abs(pv - p_ref).max():.2e} |")
    print(f"\n# {time.time() - t0:.0f} s on {args.workers} workers")


if __name__ == "__main__":
    main()
