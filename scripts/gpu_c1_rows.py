"""Per-row, per-group scaled error of the five sample.wav windows (config c1) against the oracle rows,
with the stage outputs of each window: python scripts/gpu_c1_rows.py   (SERB_DECIMATE=ffma for the FFMA2 decimator)"""
import os
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
from conftest import group_errors  # noqa: E402

from ser_b200.feature_extractor import extract_feature_frames  # noqa: E402

with np.load(REPO / "tests" / "golden" / "config_golden.npz") as data:
    golden = {k: data[k] for k in data.files}
frames = extract_feature_frames(str(REPO / "tests" / "golden" / "sample.wav"))
rows = np.vstack([f.features for f in frames])
print("decimator:", os.environ.get("SERB_DECIMATE", "tcgen05"))
for i in range(rows.shape[0]):
    rep = group_errors(rows[i], golden["c1/rows"][i], groups=("mfcc", "chroma", "mel", "tonnetz"))
    print(i, {k: f"{v[0]:.2e}" for k, v in rep.items()}, "margins", golden["c1/margins"][i],
          "tonnetz got", np.array2string(rows[i, 187:193], precision=6), "want", np.array2string(golden["c1/rows"][i, 187:193], precision=6))
