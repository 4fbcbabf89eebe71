"""Generates tests/golden/pooling_golden.npz from the REFERENCE's pooling operators.

Run in the build container only (needs /root/reference):

    python tests/golden/make_pooling_golden.py

Runs unchanged: ser/_internal/pool/windowing.py (temporal_pooling_windows),
ser/_internal/pool/stats_pool.py (mean_std_pool), ser/_internal/repr/handcrafted.py
(HandcraftedBackend.pool), ser/_internal/repr/backend.py (EncodedSequence, PoolingWindow).
The oracle shim only satisfies the package's `import librosa`; no librosa arithmetic is involved.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "oracle" / "shim"))
sys.path.insert(0, "/root/reference")

from ser._internal.pool import mean_std_pool, temporal_pooling_windows  # noqa: E402
from ser._internal.repr import EncodedSequence  # noqa: E402
from ser._internal.repr.handcrafted import HandcraftedBackend  # noqa: E402

OUT = Path(__file__).resolve().parent


def main() -> None:
    rng = np.random.default_rng(2024)
    payload: dict[str, np.ndarray] = {}
    cases = [
        # (name, n_frames, dim, frame seconds, stride seconds, window size, window stride)
        ("fast_like", 12, 193, 3.0, 1.0, 4.0, 2.0),
        ("dense", 250, 64, 0.025, 0.02, 1.0, 0.25),
        ("single_window", 5, 33, 1.0, 1.0, 30.0, 5.0),
        ("ragged_tail", 41, 1024, 0.5, 0.37, 3.3, 1.7),
    ]
    names = []
    for name, n, dim, fsec, fstride, wsize, wstride in cases:
        starts = np.arange(n, dtype=np.float64) * fstride
        ends = starts + fsec
        ends[-3:] = ends[-1] - np.asarray([0.2, 0.1, 0.0]) * fstride if n > 3 else ends[-3:]
        ends = np.maximum.accumulate(ends)
        emb = (rng.standard_normal((n, dim)) * rng.uniform(0.1, 50.0, size=dim)).astype(np.float32)
        encoded = EncodedSequence(embeddings=emb, frame_start_seconds=starts, frame_end_seconds=ends,
                                  backend_id="handcrafted")
        windows = temporal_pooling_windows(encoded, window_size_seconds=wsize, window_stride_seconds=wstride)
        payload[f"{name}/embeddings"] = emb
        payload[f"{name}/starts"] = starts
        payload[f"{name}/ends"] = ends
        payload[f"{name}/config"] = np.asarray([wsize, wstride])
        payload[f"{name}/win_starts"] = np.asarray([w.start_seconds for w in windows])
        payload[f"{name}/win_ends"] = np.asarray([w.end_seconds for w in windows])
        payload[f"{name}/mean_std"] = mean_std_pool(encoded, windows)
        payload[f"{name}/mean"] = HandcraftedBackend().pool(encoded, windows)
        names.append(name)
        print(name, emb.shape, len(windows), payload[f"{name}/mean_std"].shape)
    payload["names"] = np.asarray(names)
    # window generation only: 400 random (timeline, size, stride) configurations -> flattened windows
    cfg, flat, counts = [], [], []
    for _ in range(400):
        n = int(rng.integers(1, 40))
        t0 = float(rng.choice([0.0, rng.random() * 5]))
        step = float(rng.choice([1.0, 0.02, rng.random() + 0.01]))
        flen = float(rng.choice([step, 3 * step, 0.025]))
        size = float(rng.choice([0.1, 0.5, 1.0, 3.0, rng.random() * 4 + 1e-3, (n - 1) * step + flen]))
        stride = float(rng.choice([0.1, 0.25, 1.0, rng.random() * 2 + 1e-3]))
        starts = t0 + np.arange(n) * step
        encoded = EncodedSequence(embeddings=np.zeros((n, 1), np.float32), frame_start_seconds=starts,
                                  frame_end_seconds=starts + flen, backend_id="handcrafted")
        windows = temporal_pooling_windows(encoded, window_size_seconds=size, window_stride_seconds=stride)
        cfg.append([n, t0, step, flen, size, stride])
        counts.append(len(windows))
        flat.extend((w.start_seconds, w.end_seconds) for w in windows)
    payload["random_windows/config"] = np.asarray(cfg, dtype=np.float64)
    payload["random_windows/counts"] = np.asarray(counts, dtype=np.int64)
    payload["random_windows/bounds"] = np.asarray(flat, dtype=np.float64)
    np.savez_compressed(OUT / "pooling_golden.npz", **payload)


if __name__ == "__main__":
    main()
