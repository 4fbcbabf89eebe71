# usage: bash scripts/gpu_bench_quick.sh  (prints kernel_ms_per_step for a few chunk sizes)
mkdir -p gpurun_out
for c in ${CHUNKS:-131072}; do
  SERB_CHUNK_COLS=$c python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bq_$c.json 2> gpurun_out/bq_$c.err || tail -3 gpurun_out/bq_$c.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bq_$c.json"))
print("chunk $c ms/step %.3f value %.0f" % (d["ms_per_step"], d["value"]), {k: round(v,3) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, "launches", d["gpu_launches"])
PY
done
