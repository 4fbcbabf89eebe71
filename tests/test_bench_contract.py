"""bench.py's output contract on the CPU-runnable arm: exactly one JSON line on stdout, whatever
libraries print, with the keys the driver reads (`--impl reference` times the CPU restatement)."""
import json
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_json_line():
    proc = subprocess.run(
        [sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-clips", "1"],
        capture_output=True, text=True, timeout=600, cwd=REPO)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, proc.stdout[:2000]
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    assert line["unit"] == "audio-s/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
