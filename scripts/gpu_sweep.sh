mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err || tail -3 gpurun_out/sw_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/sw_$tag.json"))
k=d["roofline"]["kernel_ms_per_step"]
print("$tag", "ms/step %.2f" % d["ms_per_step"], {n: round(v,2) for n,v in k.items() if n in ("stft","hpss_harm","hpss_perc","istft","ola","decimate","cqt","tonnetz")})
PY
}
run base A=1
run chunk131k SERB_TON_CHUNK_COLS=131072
run chunk262k SERB_TON_CHUNK_COLS=262144
run chunk262k_seg256 SERB_TON_CHUNK_COLS=262144 SERB_HARM_SEG=256
run chunk262k_seg64 SERB_TON_CHUNK_COLS=262144 SERB_HARM_SEG=64
run chunk262k_runs8 SERB_TON_CHUNK_COLS=262144 SERB_PERC_RUNS=8
run chunk262k_runs4 SERB_TON_CHUNK_COLS=262144 SERB_PERC_RUNS=4
run chunk32k SERB_TON_CHUNK_COLS=32768
