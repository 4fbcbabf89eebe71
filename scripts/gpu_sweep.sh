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
run seg256 SERB_HARM_SEG=256
run seg512 SERB_HARM_SEG=512
run runs8 SERB_PERC_RUNS=8
run runs4 SERB_PERC_RUNS=4
run chunk524k SERB_CHUNK_COLS=524288
