"""Headline benchmark: audio-seconds/sec of fast-profile features + predict (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W                 # config c2, this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...       # the CPU path on the host cores
    python bench.py --config c3 --gpus N ...                      # config c3: 20 000 clips, strong scaling

Config c2 (default): a step is one pass of the hot path over one batch of synthetic RAVDESS-shape
audio (1 440 mono clips x 168 000 samples @ 48 kHz per GPU), run the way ``ser.api.infer`` runs
it: every clip is cut into 3 s / 1 s sliding windows (ser/_internal/repr/handcrafted.py:78-97),
each window yields one 193-d feature row, and the scaler + MLP classifier labels every row.
Clips are independent, so with N GPUs every rank processes its own 1 440 clips (weak scaling)
and no data-path collective exists; NCCL carries only the timing barrier, the max-over-ranks
reduction and the gather of the small result rows to rank 0.

Config c3: 20 000 whole clips (the ``ser --train`` extraction, one row per file,
ser/_internal/data/data_loader.py:485-529) split over the ranks by ``sharding.shard_bounds``;
total work is fixed, rows are gathered to rank 0 (``"scaling": "strong"``).

The classifier is a Pipeline(StandardScaler, MLPClassifier(300)) FITTED by scikit-learn on oracle
rows of the same synthetic family (tests/golden/c2_oracle_rows.npz, made by
tests/golden/make_c2_oracle_rows.py); both arms load the same weights.

One JSON line is printed by rank 0 (keys: the task contract; DESIGN.md sections 4 and 8).
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
FP32_PEAK_TFLOPS_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12   # CUDA-core FFMA peak at the maximum SM clock
OPS_PER_COLUMN_SURVEY = {187: 97_000, 193: 750_000}
OPS_PER_COLUMN_ITEMISED = {187: 97_000, 193: 1_045_000}      # DESIGN.md section 4 (counts min/max/compare as ops)
PARITY_TOLERANCE = 1e-4                                      # north_star: pooled features within 1e-4 (scaled, conftest.group_errors)
GROUPS = {"mfcc": (0, 40), "chroma": (40, 52), "mel": (52, 180), "contrast": (180, 187), "tonnetz": (187, 193)}
KERNEL_NAMES = {
    "stft": "stft_kernel", "tuning": "tuning_kernel", "proj": "proj_kernel", "pool": "pool_kernel",
    "short": "short_kernel", "mlp": "mlp_kernel", "hpss_harm": "hpss_harm_kernel", "hpss_perc": "hpss_perc_kernel",
    "istft": "istft_ola_kernel", "ola": "ola_kernel", "decimate": "decimate2_mma_kernel", "cqt": "cqtc_kernel",
    "tonnetz": "tonnetz_kernel", "pcm_prepare": "pcm_file_scale_kernel",
}
TONNETZ = os.environ.get("SERB_BENCH_TONNETZ", "1") == "1"   # the build implements all five groups (193-d)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3"])
    ap.add_argument("--clips", type=int, default=0, help="clips per step (c2: per GPU, default 1440; c3: total, default 20000)")
    ap.add_argument("--clip-samples", type=int, default=168000)
    ap.add_argument("--sample-rate", type=int, default=48000)
    ap.add_argument("--cpu-clips", type=int, default=0, help="clips in the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def window_plan(n_clips: int, clip_samples: int, sr: int):
    """(clip index, start inside the clip, length) of every sliding window of every clip."""
    from ser_b200.handcrafted import frame_bounds

    w_starts, w_ends = frame_bounds(clip_samples, sr, FRAME_SECONDS, STRIDE_SECONDS)
    clip_of = np.repeat(np.arange(n_clips, dtype=np.int64), w_starts.size)
    starts = np.tile(w_starts, n_clips).astype(np.int64)
    lengths = np.tile(w_ends - w_starts, n_clips).astype(np.int64)
    return clip_of, starts, lengths, int(w_starts.size)


def hbm_peak():
    path = REPO / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_fitted_model():
    """The fitted classifier both arms use: arrays of a scikit-learn Pipeline(StandardScaler,
    MLPClassifier(300)) trained by tests/golden/make_c2_oracle_rows.py."""
    with np.load(REPO / "tests" / "golden" / "c2_oracle_rows.npz", allow_pickle=False) as data:
        return {k.split("/", 1)[1]: data[k] for k in data.files if k.startswith("model/")}


def scaled_group_errors(actual: np.ndarray, expected: np.ndarray) -> dict[str, float]:
    """|a - b| / max(|b|, 1e-3 * max|b| over the group): the parity metric of tests/conftest.py."""
    out = {}
    for name, (lo, hi) in GROUPS.items():
        if hi > expected.shape[1]:
            continue
        a, b = actual[:, lo:hi].astype(np.float64), expected[:, lo:hi].astype(np.float64)
        floor = np.maximum(np.abs(b), 1e-3 * np.max(np.abs(b), axis=1, keepdims=True))
        out[name] = float(np.max(np.abs(a - b) / np.where(floor == 0.0, 1.0, floor))) if a.size else 0.0
    return out


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
# CPU arm: the oracle (numpy/scipy restatement of the reference's librosa + sklearn path).
# Nothing in this section imports ser_b200._native or loads libser_b200.so.
# ----------------------------------------------------------------------------------------
def _cpu_worker(args):
    index, clip, sr, flag_tuple, weights, whole_clip = args
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
        if whole_clip:
            rows = ser_oracle.extract_feature_from_signal(clip, sr, feature_flags=flags)[None, :].astype(np.float32)
            labels = []
        else:
            rows, starts, ends = ser_oracle.encode_sequence(clip, sr, feature_flags=flags)
            frames, _segments = ser_oracle.predict_frames(weights, rows, starts, ends)
            labels = [f.emotion for f in frames]
    del limiter
    return index, rows, labels


def _cpu_warm(_):
    """Imports the oracle in a pool worker (numpy / scipy pages, FFT plans) before anything is timed."""
    import warnings

    from oracle import ser_oracle

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")      # a silent clip: "empty frequency set" from the tuning estimate
        ser_oracle.extract_feature_from_signal(np.zeros(4096, dtype=np.float32), 16000,
                                               feature_flags=ser_oracle.FeatureFlags(True, True, True, False, False))
    time.sleep(0.05)        # keeps the worker busy long enough for every process of the pool to take one
    return os.getpid()


class CpuPool:
    """One process per host core, forked once and warmed, kept for every timed pass of the CPU arm (a pool
    forked per pass and cold workers cost the CPU arm ~25 % on an 8-core box: 12.4 -> 16.7 audio-s/s)."""

    def __init__(self, workers: int):
        import multiprocessing as mp

        self.workers = max(int(workers), 1)
        self.pool = mp.get_context("fork").Pool(self.workers) if self.workers > 1 else None
        if self.pool is not None:
            self.pool.map(_cpu_warm, range(4 * self.workers), chunksize=1)

    def map_unordered(self, fn, jobs):
        if self.pool is None:
            return [fn(job) for job in jobs]
        return list(self.pool.imap_unordered(fn, jobs, chunksize=1))

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def cpu_pass(clips: np.ndarray, sr: int, flag_tuple, weights, pool: CpuPool, whole_clip: bool = False):
    """(seconds, rows, labels) of the CPU path over ``clips`` (one clip per row) on ``pool``'s processes;
    longest-first order is moot (equal clips), one clip per task so the cores drain evenly."""
    jobs = [(i, clips[i], sr, flag_tuple, weights, whole_clip) for i in range(clips.shape[0])]
    t0 = time.perf_counter()
    results = pool.map_unordered(_cpu_worker, jobs)
    seconds = time.perf_counter() - t0
    results.sort(key=lambda r: r[0])
    rows = np.concatenate([r[1] for r in results], axis=0)
    labels = [label for r in results for label in r[2]]
    return seconds, rows, labels


def oracle_weights(model: dict):
    from oracle import ser_oracle

    return ser_oracle.MlpWeights(mean=model["mean"], scale=model["scale"], coefs=(model["w1"], model["w2"]),
                                 intercepts=(model["b1"], model["b2"]), classes=tuple(model["classes"].tolist()),
                                 out_activation=str(model["out_activation"]))


def flag_tuple_and_dim():
    flags = (True, True, True, True, TONNETZ)
    return flags, 193 if TONNETZ else 187


def run_reference(args) -> None:
    """--impl reference: the CPU implementation of the path on the host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ser_b200 import synth        # pure numpy; the native library is never loaded by this arm

    flag_tuple, dim = flag_tuple_and_dim()
    sr, n = args.sample_rate, args.clip_samples
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_step = args.cpu_clips or max(4 * cores, 8)     # four clips per core: the cores drain evenly
    whole_clip = args.config == "c3"
    specs = synth.ravdess_specs(per_step)
    clips = np.stack([synth.clip_audio(s, sr, n) for s in specs])
    model = load_fitted_model()
    if dim != model["w1"].shape[0]:
        raise RuntimeError("the committed classifier expects 193-d rows; run without SERB_BENCH_TONNETZ=0")
    weights = oracle_weights(model)
    warm = max(args.warmup, 1)
    with CpuPool(cores) as pool:
        for _ in range(warm):
            cpu_pass(clips, sr, flag_tuple, weights, pool, whole_clip)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_pass(clips, sr, flag_tuple, weights, pool, whole_clip)
        elapsed = time.perf_counter() - t0
    audio_seconds = per_step * n / sr * args.steps
    value = audio_seconds / elapsed
    what = "whole-clip rows" if whole_clip else f"{FRAME_SECONDS}s/{STRIDE_SECONDS}s windows + MLP"
    sample = (f"{per_step} clips x {n} samples @ {sr} Hz per step ({dim}-d features, {what}), "
              f"numpy/scipy oracle, one warmed process per host core kept across steps, BLAS threads limited to 1")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * elapsed / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong" if whole_clip else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.config} sample: {sample}", "feature_dim": dim,
                   "weights": "fitted scikit-learn Pipeline(StandardScaler, MLPClassifier(300)) (tests/golden/c2_oracle_rows.npz)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "librosa 0.11.0 is not installable here (SURVEY.md F2): this is the CPU restatement "
                "(oracle/) of the reference's librosa+sklearn path",
    }
    emit(line)


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
def kernel_bounds():
    """Binding unit of each kernel as ncu reports it (profiles/kernel_bounds.json, with sources)."""
    path = REPO / "profiles" / "kernel_bounds.json"
    try:
        return json.loads(path.read_text())
    except Exception:
        return {}


def run_b200(args) -> None:
    import torch

    from ser_b200 import _native, mlp, multi_gpu, sharding, synth

    # stdout carries exactly one JSON line: NCCL's version / debug banner goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    info = multi_gpu.rank_info()
    rank, local_rank, world = info.rank, info.local_rank, info.world
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: ser_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # host side of this rank next to its GPU (pinned buffers, staging threads); the CPU baseline leg
    # below gets the whole machine back
    numa = multi_gpu.bind_host_to_gpu(local_rank)
    full_affinity = numa.pop("restore", None) if numa else None
    multi_gpu.init_process_group(info, "nccl", device=torch.device("cuda", local_rank))

    def barrier():
        multi_gpu.barrier(info)

    def max_over_ranks(x: float) -> float:
        return multi_gpu.max_over_ranks(info, x, device="cuda")

    flag_tuple, dim = flag_tuple_and_dim()
    bits = sum(bit for bit, on in zip((1, 2, 4, 8, 16), flag_tuple) if on)
    sr, n_samples = args.sample_rate, args.clip_samples
    c3 = args.config == "c3"
    if c3:
        total_clips = args.clips or 20000
        lo, hi = sharding.shard_bounds(np.full(total_clips, n_samples, dtype=np.int64), world)[rank]
        n_clips, first_index = hi - lo, lo
        audio_seconds_step = total_clips * n_samples / sr          # total work is fixed
    else:
        n_clips, first_index = args.clips or 1440, rank * (args.clips or 1440)
        total_clips = n_clips * world
        audio_seconds_step = total_clips * n_samples / sr
    ctx = _native.get_context(local_rank)

    # synthetic audio on the device: float32 as read_audio_file returns it, and the int16 PCM it was decoded from
    wave = torch.empty((n_clips, n_samples), dtype=torch.float32, device="cuda")
    pcm = torch.empty((n_clips, n_samples), dtype=torch.int16, device="cuda")
    for a in range(0, n_clips, 1440):
        b = min(a + 1440, n_clips)
        wave[a:b], pcm[a:b] = synth.batch_audio_torch(b - a, sr, n_samples, device="cuda", first_index=first_index + a,
                                                      return_pcm=True)
    wave = wave.reshape(-1).contiguous()
    if c3:
        clip_of = np.arange(n_clips, dtype=np.int64)
        w_starts = np.zeros(n_clips, dtype=np.int64)
        lengths = np.full(n_clips, n_samples, dtype=np.int64)
        windows_per_clip = 1
    else:
        clip_of, w_starts, lengths, windows_per_clip = window_plan(n_clips, n_samples, sr)
    starts = clip_of * n_samples + w_starts
    n_rows = int(starts.size)
    total_rows = total_clips * windows_per_clip        # rows of all ranks: what rank 0 holds after the gather

    model = load_fitted_model()
    if dim != model["w1"].shape[0]:
        raise RuntimeError("the committed classifier expects 193-d rows; run without SERB_BENCH_TONNETZ=0")
    weights = mlp.MlpWeights(mean=model["mean"], scale=model["scale"], w1=model["w1"], b1=model["b1"], w2=model["w2"],
                             b2=model["b2"], classes=tuple(model["classes"].tolist()),
                             out_activation=_native.OUT_SOFTMAX if str(model["out_activation"]) == "softmax" else _native.OUT_LOGISTIC)
    mlp.ensure_loaded(weights, local_rank)
    n_classes = len(weights.classes)

    feats = torch.empty((n_rows, dim), dtype=torch.float32, device="cuda")
    proba = torch.empty((n_rows, n_classes), dtype=torch.float64, device="cuda")
    labels = torch.empty((n_rows,), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    side = torch.cuda.Stream()          # an explicit stream: torch events and the library share it
    torch.cuda.set_stream(side)
    stream = side.cuda_stream

    def step():
        ctx.features_device(wave.data_ptr(), wave.numel(), starts, lengths, sr, bits, feats.data_ptr(), stream)
        if not c3:
            ctx.mlp_predict_device(feats.data_ptr(), n_rows, proba.data_ptr(), labels.data_ptr(), stream)

    warm = max(args.warmup, 3)          # the timing rules ask for at least three warm-up steps
    for _ in range(warm):
        step()
    ctx.features_device_check(stream)   # synchronises; raises if any staged sample was not finite

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
    value = audio_seconds_step * args.steps / (ms_total / 1e3)
    dev_labels = labels.cpu().numpy()
    dev_feats = feats.cpu().numpy()

    # ---- end to end: host PCM16 in (what the files hold), labels + probabilities back, every step ----
    e2e = None
    if not args.no_e2e:
        files_pinned = torch.empty((n_clips, n_samples), dtype=torch.int16, pin_memory=True)
        files_pinned.copy_(pcm)
        torch.cuda.synchronize()
        pinned_np = files_pinned.numpy()
        file_list = pinned_np                                       # (n_files, frames) int16: one contiguous pinned buffer

        def e2e_call(files):
            if c3:
                rows = ctx.features_host_pcm16(files, 1, clip_of, w_starts, lengths, sr, bits)
                return rows, None, None
            return ctx.infer_host_pcm16(files, 1, clip_of, w_starts, lengths, sr, bits, want_features=False)

        rank_rows = [0] * world
        if info.distributed:
            import torch.distributed as dist

            dist.all_gather_object(rank_rows, n_rows)
        else:
            rank_rows = [n_rows]

        def local_block(f_host, p_host, l_host):
            # the small per-row results that go to rank 0 (north_star: "only the small per-clip feature vectors
            # are gathered to the host")
            return f_host if c3 else np.concatenate([p_host, l_host[:, None].astype(np.float64)], axis=1)

        # one padded tensor gather per step, GPU to GPU under NCCL, double-buffered: step i's rows travel (and
        # rank 0 waits for the slowest rank) while step i + 1 computes; the last step's rows are collected
        # before the clock stops, so every row is on rank 0 inside the timed region
        gatherer = multi_gpu.RowGatherer(info, rank_rows, (dim,) if c3 else (n_classes + 1,),
                                         np.float32 if c3 else np.float64)

        def e2e_timed(files, steps):
            for _ in range(warm):  # warm-up: staging buffers, the gather's channels, and the clocks after the idle gap
                gatherer.collect(gatherer.submit(local_block(*e2e_call(files))))
            barrier()
            t0 = time.perf_counter()
            chain = []
            gathered = None
            pending = None
            trace = os.environ.get("SERB_BENCH_TRACE") == "1"
            for _ in range(steps):
                ta = time.perf_counter()
                f_host, p_host, l_host = e2e_call(files)
                tb = time.perf_counter()
                chain.append(ctx.last_compute_ms())
                ticket = gatherer.submit(local_block(f_host, p_host, l_host))
                if pending is not None:
                    gathered = gatherer.collect(pending)
                pending = ticket
                if trace:
                    print(f"[trace] rank {rank}: call {1e3 * (tb - ta):.2f} ms (device chain {chain[-1]:.2f}), "
                          f"gather {1e3 * (time.perf_counter() - tb):.2f} ms", file=sys.stderr, flush=True)
            gathered = gatherer.collect(pending)                    # inside the timed region
            elapsed = max_over_ranks(time.perf_counter() - t0)
            barrier()
            if rank == 0:
                assert gathered.shape[0] == total_rows, "rank 0 does not hold every rank's rows"
                assert np.array_equal(gathered[:n_rows], local_block(f_host, p_host, l_host)), "gathered rows differ"
            return elapsed, float(np.median(chain)), (f_host, p_host, l_host), gathered

        e2e_steps = max(2, args.steps)          # the same K steps as the device-resident region
        elapsed, chain_ms, last, gathered = e2e_timed(file_list, e2e_steps)
        f_host, p_host, l_host = last
        d2h = int(f_host.nbytes) if c3 else int(p_host.nbytes + l_host.nbytes)
        e2e = {"value": audio_seconds_step * e2e_steps / elapsed, "unit": UNIT,
               "h2d_bytes_per_step": int(pinned_np.nbytes), "d2h_bytes_per_step": d2h,
               "steps": e2e_steps, "ms_per_step": 1e3 * elapsed / e2e_steps, "device_chain_ms": chain_ms,
               "input": "int16 PCM in pinned host memory (serb_infer_host_pcm16: decode scaling, peak normalisation, "
                        "features and classifier on the device)" if not c3 else
                        "int16 PCM in pinned host memory (serb_features_host_pcm16)",
               "gathered_rows_on_rank0": None if gathered is None else int(gathered.shape[0]),
               "host_binding": numa}
        if c3:
            assert np.array_equal(f_host, dev_feats), "PCM16 host-entry rows differ from the device path"
        else:
            assert np.array_equal(l_host, dev_labels), "PCM16 host-entry labels differ from the device path"
        # the same call from pageable memory (what a numpy caller of the Python API passes)
        pageable = [np.array(pinned_np[i]) for i in range(n_clips)]
        elapsed_p, chain_p, last_p, _ = e2e_timed(pageable, 2)
        if c3:
            assert np.array_equal(last_p[0], dev_feats), "rows from pageable memory differ from the device path"
        else:
            assert np.array_equal(last_p[2], dev_labels), "labels from pageable memory differ from the device path"
        e2e["pageable"] = {"value": audio_seconds_step * 2 / elapsed_p, "unit": UNIT, "ms_per_step": 1e3 * elapsed_p / 2,
                           "device_chain_ms": chain_p,
                           "input": "one pageable int16 numpy array per file, gathered into pinned slots by the "
                                    "library's staging threads (csrc/host_stage.h)"}
        del pageable
        # Two callers: two host threads, each with its own context (its own scratch and streams), take the
        # steps alternately through the same public call.  Every step still copies its 484 MB in and reads
        # its rows back inside the timed region; what overlaps is one caller's planning, first pieces and
        # small first chunks with the other caller's long chunks -- how a server keeps one GPU busy.
        if not c3:
            try:
                import threading

                ctx_b = _native.Context(local_rank)
                ctx_b.mlp_load(weights.mean, weights.scale, weights.w1, weights.b1, weights.w2, weights.b2,
                               weights.out_activation)
                callers = (ctx, ctx_b)
                n_two = max(4, e2e_steps - e2e_steps % 2)
                results = [None] * n_two

                def caller(which, first, count):
                    c = callers[which]
                    for k in range(count):
                        results[first + 2 * k] = c.infer_host_pcm16(file_list, 1, clip_of, w_starts, lengths, sr, bits,
                                                                    want_features=False)

                for which in (0, 1):                                   # warm the second context
                    caller(which, which, 1)
                barrier()
                t0 = time.perf_counter()
                threads = [threading.Thread(target=caller, args=(w, w, n_two // 2)) for w in (0, 1)]
                for t in threads:
                    t.start()
                pending = None
                for k in range(n_two):                                 # rows go to rank 0 in step order, as they complete
                    while results[k] is None:
                        time.sleep(0.0002)
                    _, p_k, l_k = results[k]
                    ticket = gatherer.submit(local_block(None, p_k, l_k))
                    if pending is not None:
                        gatherer.collect(pending)
                    pending = ticket
                for t in threads:
                    t.join()
                gathered_two = gatherer.collect(pending)
                elapsed_two = max_over_ranks(time.perf_counter() - t0)
                barrier()
                assert all(np.array_equal(r[2], dev_labels) for r in results), "two-caller labels differ from the device path"
                if rank == 0:
                    assert gathered_two.shape[0] == total_rows
                e2e["two_callers"] = {"value": audio_seconds_step * n_two / elapsed_two, "unit": UNIT, "steps": n_two,
                                      "ms_per_step": 1e3 * elapsed_two / n_two,
                                      "input": "the same call from two host threads with one context each, steps taken "
                                               "alternately; every step's H2D, D2H and gather inside the timed region"}
                ctx_b.close()
            except Exception as exc:  # an auxiliary number must not take the line down
                e2e["two_callers"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        if not c3:
            # round 1's float32 entry, for continuity: 4 bytes per sample over PCIe
            host_wave = torch.empty(wave.numel(), dtype=torch.float32, pin_memory=True)
            host_wave.copy_(wave)
            torch.cuda.synchronize()
            hw = host_wave.numpy()
            ctx.infer_host(hw, starts, lengths, sr, bits, want_features=False)
            barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                ctx.infer_host(hw, starts, lengths, sr, bits, want_features=False)
            elapsed_f = max_over_ranks(time.perf_counter() - t0)
            barrier()
            e2e["float32_pinned"] = {"value": audio_seconds_step * 2 / elapsed_f, "unit": UNIT,
                                     "ms_per_step": 1e3 * elapsed_f / 2, "h2d_bytes_per_step": int(hw.nbytes)}
            del host_wave, hw

    # ---- the 187-d slice (tonnetz off) timed the same way, for continuity with earlier rounds ----
    slice187 = None
    if TONNETZ and not c3:
        bits187 = 15
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
        slice187 = {"ms_per_step": ms187, "value": audio_seconds_step / (ms187 / 1e3), "unit": UNIT,
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
    # event pair: count kernel launches, not brackets.  The constant-Q bracket holds the kernel's two
    # instantiations (per-column octaves, shared-stage octaves) and is reported as ONE launch of the
    # kernel: together they pass over the step's algorithmic bytes once.
    dom_n *= {"decimate": 7 if sr >= 33400 else 6, "tonnetz": 2}.get(dom, 1)
    total_cols = int(np.sum(1 + lengths // 512))
    # algorithmic bytes: every input sample once (float32 at boundary B1/B2) + every output row once
    alg_bytes_step = 4 * n_clips * n_samples + 4 * dim * n_rows
    peak, peak_src = hbm_peak()
    achieved = alg_bytes_step / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
    bounds = kernel_bounds().get(KERNEL_NAMES[dom], {})
    traffic = None
    if bounds.get("dram_bytes_per_stft_column_per_launch") is not None:
        traffic = bounds["dram_bytes_per_stft_column_per_launch"] * total_cols / max(dom_n, 1)
    fp32_measured = ctx.fp32_peak_tflops()
    step_ops_survey = total_cols * OPS_PER_COLUMN_SURVEY[dim]
    step_ops_itemised = total_cols * OPS_PER_COLUMN_ITEMISED[dim]
    sum_kernel_ms = max(sum(v[0] for v in kms.values()), 1e-9)
    roofline = {
        "kernel": KERNEL_NAMES[dom], "bound": bounds.get("bound", "unprofiled"),
        "binding_pipe": bounds.get("binding_pipe"), "binding_pipe_frac": bounds.get("binding_pipe_frac"),
        "bound_source": bounds.get("source"),
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "traffic_source": bounds.get("traffic_source"), "peak_source": peak_src,
        "chain_dram_bytes_per_stft_column": kernel_bounds().get("_chain", {}).get("dram_bytes_per_stft_column"),
        "chain_dram_source": kernel_bounds().get("_chain", {}).get("source"),
        "algorithmic_bytes_per_launch": alg_bytes_step / max(dom_n, 1), "launches_per_step": dom_n,
        "avg_launch_ms": dom_ms / max(dom_n, 1),
        "kernel_ms_per_step": {k: v[0] for k, v in kms.items()},
        "share_of_step": dom_ms / sum_kernel_ms,
        "whole_step_hbm_frac": alg_bytes_step / (ms_per_step / 1e3) / 1e9 / peak,
        "fp32": {"peak_tflops_measured": fp32_measured, "peak_tflops_nominal": FP32_PEAK_TFLOPS_NOMINAL,
                 "peak_how": "serb_debug_fp32_peak: dependent-free FFMA chains on every SM, this run",
                 "achieved_tflops_survey": step_ops_survey / (ms_per_step / 1e3) / 1e12,
                 "achieved_tflops_itemised": step_ops_itemised / (ms_per_step / 1e3) / 1e12,
                 "frac_survey": step_ops_survey / (ms_per_step / 1e3) / 1e12 / max(fp32_measured, 1e-9),
                 "frac_itemised": step_ops_itemised / (ms_per_step / 1e3) / 1e12 / max(fp32_measured, 1e-9),
                 "note": f"whole step; SURVEY 8(d) {OPS_PER_COLUMN_SURVEY[dim] // 1000} kop/column next to this repo's "
                         f"itemised {OPS_PER_COLUMN_ITEMISED[dim] // 1000} kop/column (DESIGN.md section 4, counts "
                         "min/max/compare as ops)"},
        "how": "CUDA events around every launch of the dominant kernel in one extra pass of the same step",
    }

    # ---- CPU baseline on the box's host cores + parity of the GPU rows against it, rank 0 at N=1 ----
    cpu_baseline = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if full_affinity:
            os.sched_setaffinity(0, full_affinity)      # the CPU arm runs on every host core
        cores = min(os.cpu_count() or 1, 32)
        per = args.cpu_clips or min(n_clips, max(4 * cores, 8))
        sample_clips = wave[: per * n_samples].reshape(per, n_samples).cpu().numpy()
        pool_note = "one warmed process per core"
        try:
            with CpuPool(cores) as pool:    # forked and warmed outside the timed pass; the children never touch CUDA
                secs, cpu_rows, cpu_labels = cpu_pass(sample_clips, sr, flag_tuple, oracle_weights(model), pool,
                                                      whole_clip=c3)
        except Exception as exc:            # noqa: BLE001 - a broken process pool must not cost the GPU line its parity check
            print(f"bench: CPU pool failed ({type(exc).__name__}: {exc}); CPU sample on one core instead", file=sys.stderr)
            cores, per = 1, min(per, 8)
            sample_clips = sample_clips[:per]
            pool_note = "ONE process (the per-core pool failed to start)"
            secs, cpu_rows, cpu_labels = cpu_pass(sample_clips, sr, flag_tuple, oracle_weights(model), CpuPool(1),
                                                  whole_clip=c3)
        cpu_baseline = {"value": per * n_samples / sr / secs, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {per} clips of the step ({per * windows_per_clip} rows), {pool_note}, "
                                  f"BLAS threads limited to 1, {secs:.1f} s wall",
                        "host_cores_available": os.cpu_count()}
        gpu_rows = dev_feats[: per * windows_per_clip]
        errors = scaled_group_errors(gpu_rows, cpu_rows)
        parity = {"rows": int(gpu_rows.shape[0]), "max_scaled_err": errors, "tolerance": PARITY_TOLERANCE,
                  "against": "oracle rows of the same clips (the cpu_baseline pass)"}
        if not c3:
            gpu_labels = [weights.classes[i] for i in dev_labels[: per * windows_per_clip]]
            parity["labels_equal"] = gpu_labels == cpu_labels
            parity["labels_compared"] = len(cpu_labels)
        parity["ok"] = bool(max(errors.values()) <= PARITY_TOLERANCE and parity.get("labels_equal", True))

    if rank == 0:
        what = "whole-clip rows (training extraction)" if c3 else \
            f"{FRAME_SECONDS}s/{STRIDE_SECONDS}s sliding windows ({n_rows} rows/GPU), {dim}-d features + MLP(300) predict"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "warmup_requested": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if c3 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": (f"c3: {total_clips} clips x {n_samples} samples @ {sr} Hz split over {world} GPU(s), {what}" if c3 else
                             f"c2: {n_clips} clips x {n_samples} samples @ {sr} Hz per GPU, {what}"),
                "feature_dim": dim, "rows_per_gpu": n_rows, "stft_columns_per_gpu": total_cols,
                "weights": "fitted scikit-learn Pipeline(StandardScaler, MLPClassifier(300)) (tests/golden/c2_oracle_rows.npz)",
                "l2": f"inputs {wave.numel() * 4 / 1e6:.0f} MB per GPU exceed the 126 MB L2; no explicit flush",
                "parallelism": f"clips sharded over {world} GPU(s), no data-path collective; result rows gathered to rank 0 in e2e",
            },
            "e2e": e2e, "gpu_launches": int(launches) * world, "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity, "slice_187d": slice187,
        }
        emit(line)
    multi_gpu.destroy_process_group(info)
    if parity is not None and not parity["ok"]:
        print(f"bench.py: parity check failed: {parity}", file=sys.stderr)
        sys.exit(3)


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
