"""Headline benchmark: audio-seconds/sec of fast-profile features + predict (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU path on the host cores

A step is one pass of the hot path over one batch of synthetic RAVDESS-shape audio
(config c2: 1 440 mono clips x 168 000 samples @ 48 kHz per GPU), run the way
``ser.api.infer`` runs it: every clip is cut into 3 s / 1 s sliding windows
(ser/_internal/repr/handcrafted.py:78-97), each window yields one feature row, and the
scaler+MLP classifier labels every row.  Clips are independent, so with N GPUs every rank
processes its own 1 440 clips (weak scaling) and no data-path collective exists; NCCL is
used only for the timing barrier and the max-over-ranks reduction.

One JSON line is printed by rank 0 (keys: see the task contract).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

METRIC = "audio_seconds_per_second_fast_profile_features_predict"
UNIT = "audio-s/s"
FRAME_SECONDS, STRIDE_SECONDS = 3, 1
FP32_PEAK_TFLOPS_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12   # CUDA-core FFMA peak at max clock
FP32_PEAK_TFLOPS_MEASURED = 72.2                             # profiles/r01_fp32_microbench.txt (dependent-free FFMA)
FLOP_PER_COLUMN = 97_000                                     # SURVEY.md section 8(d), 187-d slice
FLOP_PER_COLUMN_193 = 1_045_000                              # itemised ops of the 193-d chain
KERNEL_NAMES = {
    "stft": "stft_kernel", "tuning": "tuning_kernel", "proj": "proj_kernel", "pool": "pool_kernel",
    "short": "short_kernel", "mlp": "mlp_kernel", "hpss_harm": "hpss_harm_kernel", "hpss_perc": "hpss_perc_kernel",
    "istft": "istft_kernel", "ola": "ola_kernel", "decimate": "decimate2_kernel", "cqt": "cqt_kernel",
    "tonnetz": "tonnetz_kernel",
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=1440, help="clips per GPU per step")
    ap.add_argument("--clip-samples", type=int, default=168000)
    ap.add_argument("--sample-rate", type=int, default=48000)
    ap.add_argument("--cpu-clips", type=int, default=0, help="clips in the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def supported_flags():
    """Feature groups this build computes: everything the library implements."""
    from ser_b200 import _native
    from ser_b200.config import FeatureFlags

    lib = _native.load_library()
    del lib
    tonnetz = os.environ.get("SERB_BENCH_TONNETZ", "auto")
    if tonnetz == "auto":
        tonnetz = "1" if getattr(_native, "HAS_TONNETZ", False) else "0"
    return FeatureFlags(tonnetz=(tonnetz == "1"))


def window_plan(n_clips: int, clip_samples: int, sr: int):
    """(starts, lengths) of every sliding window of every clip inside one packed buffer."""
    from ser_b200.handcrafted import frame_bounds

    w_starts, w_ends = frame_bounds(clip_samples, sr, FRAME_SECONDS, STRIDE_SECONDS)
    base = (np.arange(n_clips, dtype=np.int64) * clip_samples)[:, None]
    starts = (base + w_starts[None, :]).reshape(-1)
    lengths = np.tile(w_ends - w_starts, n_clips).astype(np.int64)
    return starts, lengths, int(w_starts.size)


def peaks():
    path = REPO / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            data = json.loads(path.read_text())
            return float(data["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe)."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows: list[list[str]] = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "", 1).isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "", 1).isdigit()]
        reasons = set()
        for r in self.rows:
            for name, cell in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ----------------------------------------------------------------------------------------
# CPU arm: the oracle (numpy/scipy restatement of the reference's librosa + sklearn path)
# ----------------------------------------------------------------------------------------
def _cpu_worker(args):
    clip, sr, flag_tuple, weights = args
    import warnings

    from oracle import ser_oracle

    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        flags = ser_oracle.FeatureFlags(*flag_tuple)
        emb, starts, ends = ser_oracle.encode_sequence(clip, sr, feature_flags=flags)
        frames, segments = ser_oracle.predict_frames(weights, emb, starts, ends)
    del limiter
    return len(frames)


def cpu_pass(clips: np.ndarray, sr: int, flags, weights, workers: int) -> float:
    """Seconds to run the CPU path over ``clips`` (rows) with ``workers`` processes."""
    flag_tuple = (flags.mfcc, flags.chroma, flags.mel, flags.contrast, flags.tonnetz)
    jobs = [(clips[i], sr, flag_tuple, weights) for i in range(clips.shape[0])]
    t0 = time.perf_counter()
    if workers <= 1:
        for job in jobs:
            _cpu_worker(job)
    else:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(workers) as pool:
            list(pool.imap_unordered(_cpu_worker, jobs, chunksize=max(1, len(jobs) // (4 * workers))))
    return time.perf_counter() - t0


def oracle_weights(dim: int, seed: int = 0):
    from oracle import ser_oracle
    from ser_b200 import synth

    rng = np.random.default_rng(seed)
    return ser_oracle.MlpWeights(
        mean=rng.standard_normal(dim), scale=1.0 + rng.random(dim),
        coefs=(rng.standard_normal((dim, 300)) * 0.1, rng.standard_normal((300, 8)) * 0.1),
        intercepts=(rng.standard_normal(300) * 0.1, rng.standard_normal(8) * 0.1),
        classes=tuple(sorted(synth.RAVDESS_EMOTIONS.values())), out_activation="softmax")


def run_reference(args) -> None:
    """--impl reference: the CPU implementation of the path on the host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ser_b200 import synth
    from ser_b200.config import feature_dim

    flags = supported_flags()
    dim = feature_dim(flags)
    sr, n = args.sample_rate, args.clip_samples
    cores = os.cpu_count() or 1
    per_step = args.cpu_clips or max(cores, 8)
    specs = synth.ravdess_specs(per_step)
    clips = np.stack([synth.clip_audio(s, sr, n) for s in specs])
    weights = oracle_weights(dim)
    for _ in range(args.warmup):
        cpu_pass(clips, sr, flags, weights, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pass(clips, sr, flags, weights, cores)
    elapsed = time.perf_counter() - t0
    audio_seconds = per_step * n / sr * args.steps
    value = audio_seconds / elapsed
    sample = f"{per_step} clips x {n} samples @ {sr} Hz per step ({dim}-d features + MLP), numpy/scipy oracle"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"c2 sample: {sample}", "feature_dim": dim},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "librosa 0.11.0 is not installable here (SURVEY.md F2): this is the CPU restatement "
                "(oracle/) of the reference's librosa+sklearn path, one process per host core",
    }
    emit(line)


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import torch

    from ser_b200 import _native, mlp, synth
    from ser_b200.config import feature_dim, flag_bits

    from ser_b200 import multi_gpu

    # stdout carries exactly one JSON line: NCCL's version / debug banner goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    info = multi_gpu.rank_info()
    rank, local_rank, world, distributed = info.rank, info.local_rank, info.world, info.distributed
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: ser_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    multi_gpu.init_process_group(info, "nccl", device=torch.device("cuda", local_rank))

    def barrier():
        multi_gpu.barrier(info)

    def max_over_ranks(x: float) -> float:
        return multi_gpu.max_over_ranks(info, x, device="cuda")

    flags = supported_flags()
    bits = flag_bits(flags)
    dim = feature_dim(flags)
    sr, n_samples, n_clips = args.sample_rate, args.clip_samples, args.clips
    ctx = _native.get_context(local_rank)

    wave = synth.batch_audio_torch(n_clips, sr, n_samples, device="cuda", first_index=rank * n_clips)
    wave = wave.reshape(-1).contiguous()
    starts, lengths, windows_per_clip = window_plan(n_clips, n_samples, sr)
    n_rows = int(starts.size)
    audio_seconds_step = n_clips * n_samples / sr

    rng = np.random.default_rng(0)
    weights = mlp.MlpWeights(
        mean=rng.standard_normal(dim), scale=1.0 + rng.random(dim),
        w1=rng.standard_normal((dim, 300)) * 0.1, b1=rng.standard_normal(300) * 0.1,
        w2=rng.standard_normal((300, 8)) * 0.1, b2=rng.standard_normal(8) * 0.1,
        classes=tuple(sorted(synth.RAVDESS_EMOTIONS.values())), out_activation=_native.OUT_SOFTMAX)
    mlp.ensure_loaded(weights, local_rank)

    feats = torch.empty((n_rows, dim), dtype=torch.float32, device="cuda")
    proba = torch.empty((n_rows, 8), dtype=torch.float64, device="cuda")
    labels = torch.empty((n_rows,), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    side = torch.cuda.Stream()          # an explicit stream: torch events and the library share it
    torch.cuda.set_stream(side)
    stream = side.cuda_stream

    def step():
        ctx.features_device(wave.data_ptr(), wave.numel(), starts, lengths, sr, bits, feats.data_ptr(), stream)
        ctx.mlp_predict_device(feats.data_ptr(), n_rows, proba.data_ptr(), labels.data_ptr(), stream)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count
    barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(float(ev0.elapsed_time(ev1)))
    launches = ctx.launch_count - launches0
    ms_per_step = ms_total / args.steps
    value = world * audio_seconds_step * args.steps / (ms_total / 1e3)

    # ---- end to end: pinned host buffers in, labels + probabilities back, every step ----
    e2e = None
    if not args.no_e2e:
        host_wave = torch.empty(wave.numel(), dtype=torch.float32, pin_memory=True)
        host_wave.copy_(wave)
        torch.cuda.synchronize()
        hw = host_wave.numpy()
        e2e_steps = max(2, min(args.steps, 5))
        ctx.infer_host(hw, starts, lengths, sr, bits, want_features=False)   # warm the staging buffers
        barrier()
        t0 = time.perf_counter()
        chain_ms = []
        for _ in range(e2e_steps):
            _f, p_host, l_host = ctx.infer_host(hw, starts, lengths, sr, bits, want_features=False)
            chain_ms.append(ctx.last_compute_ms())
        elapsed = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e = {"value": world * audio_seconds_step * e2e_steps / elapsed, "unit": UNIT,
               "h2d_bytes_per_step": int(hw.nbytes), "d2h_bytes_per_step": int(p_host.nbytes + l_host.nbytes),
               "steps": e2e_steps, "ms_per_step": 1e3 * elapsed / e2e_steps,
               "device_chain_ms": float(np.median(chain_ms))}
        assert np.array_equal(l_host, labels.cpu().numpy()), "host-entry labels differ from the device path"

    # ---- the 187-d slice (tonnetz off) timed the same way, for continuity with earlier rounds ----
    slice187 = None
    if flags.tonnetz:
        from ser_b200.config import FeatureFlags

        bits187 = flag_bits(FeatureFlags(tonnetz=False))
        feats187 = torch.empty((n_rows, 187), dtype=torch.float32, device="cuda")
        for _ in range(3):
            ctx.features_device(wave.data_ptr(), wave.numel(), starts, lengths, sr, bits187, feats187.data_ptr(), stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            ctx.features_device(wave.data_ptr(), wave.numel(), starts, lengths, sr, bits187, feats187.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        ms187 = max_over_ranks(float(e0.elapsed_time(e1))) / args.steps
        slice187 = {"ms_per_step": ms187, "value": world * audio_seconds_step / (ms187 / 1e3), "unit": UNIT,
                    "note": "features only, FeatureFlags(tonnetz=False): mfcc + chroma + mel + contrast"}
        del feats187

    # ---- roofline of the dominant kernel: CUDA events around every launch, one extra pass ----
    ctx.set_profile(True)
    step()
    torch.cuda.synchronize()
    kms = ctx.kernel_ms()
    ctx.set_profile(False)
    dom = max(kms, key=lambda k: kms[k][0])
    dom_ms, dom_n = kms[dom]
    # the library brackets the decimation launches of a chunk (and the two tonnetz kernels) with one
    # event pair: count kernel launches, not brackets (the constant-Q kernel is one launch per chunk)
    dom_n *= {"decimate": 7 if sr >= 33400 else 6, "tonnetz": 2}.get(dom, 1)
    total_cols = int(np.sum(1 + lengths // 512))
    # algorithmic bytes: every input sample once + every output row once
    alg_bytes_step = 4 * n_clips * n_samples + 4 * dim * n_rows
    peak, peak_src = peaks()
    achieved = alg_bytes_step / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
    # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture
    # (bytes per STFT column per launch x the columns an average launch of this run covers)
    traffic = None
    traffic_file = REPO / "profiles" / "dominant_kernel_traffic.json"
    if traffic_file.exists():
        try:
            entry = json.loads(traffic_file.read_text()).get(KERNEL_NAMES[dom])
            if entry:
                traffic = entry["dram_bytes_per_stft_column_per_launch"] * total_cols / max(dom_n, 1)
        except Exception:
            traffic = None
    flop_per_column = FLOP_PER_COLUMN_193 if flags.tonnetz else FLOP_PER_COLUMN
    roofline = {
        "kernel": KERNEL_NAMES[dom], "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes_step / max(dom_n, 1), "launches_per_step": dom_n,
        "avg_launch_ms": dom_ms / max(dom_n, 1),
        "kernel_ms_per_step": {k: v[0] for k, v in kms.items()},
        "share_of_step": dom_ms / max(sum(v[0] for v in kms.values()), 1e-9),
        "fp32": {"achieved_tflops": total_cols * flop_per_column / (ms_per_step / 1e3) / 1e12,
                 "peak_tflops_nominal": FP32_PEAK_TFLOPS_NOMINAL, "peak_tflops_measured": FP32_PEAK_TFLOPS_MEASURED,
                 "note": f"whole step, {flop_per_column // 1000} kop/column (DESIGN.md section 4); the path is "
                         "ALU/FP32/shared-memory bound, not HBM bound"},
        "how": "CUDA events around every launch of the dominant kernel in one extra pass of the same step",
    }

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        per = args.cpu_clips or (8 if flags.tonnetz else 24)
        sample_clips = wave[: per * n_samples].reshape(per, n_samples).cpu().numpy()
        secs = cpu_pass(sample_clips, sr, flags, oracle_weights(dim), workers=1)
        cpu_baseline = {"value": per * n_samples / sr / secs, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": f"first {per} clips of the step ({per * windows_per_clip} windows), one process, "
                                  f"BLAS threads limited to 1, {secs:.1f} s",
                        "host_cores_available": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"c2: {n_clips} clips x {n_samples} samples @ {sr} Hz per GPU, "
                            f"{FRAME_SECONDS}s/{STRIDE_SECONDS}s sliding windows ({n_rows} rows/GPU), "
                            f"{dim}-d features + MLP(300) predict",
                "feature_dim": dim, "rows_per_gpu": n_rows, "stft_columns_per_gpu": total_cols,
                "weights": "random-init Pipeline(StandardScaler, MLPClassifier(300)) shape",
                "l2": f"inputs {wave.numel() * 4 / 1e6:.0f} MB per GPU exceed the 126 MB L2; no explicit flush",
                "parallelism": f"clips sharded over {world} GPU(s), no collective",
            },
            "e2e": e2e, "gpu_launches": int(launches) * world, "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "slice_187d": slice187,
        }
        emit(line)
    multi_gpu.destroy_process_group(info)


_RESULT_FD = None


def claim_stdout() -> None:
    """stdout must carry exactly one JSON line.  Libraries write banners straight to file
    descriptor 1 (NCCL's version line; NCCL_DEBUG_FILE=/dev/stderr does not hold when stderr is a
    redirected file or a socket), so the real stdout is set aside for the result and descriptor 1
    is pointed at stderr for everything else."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
