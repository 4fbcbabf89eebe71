# usage: bash scripts/gpu_16k_ab.sh   (RAVDESS-shape batch at 16 kHz: column-mapped constant-Q rows against the lane = row kernels)
mkdir -p gpurun_out
for m in cols rows; do
  SERB_CQT=$m python bench.py --sample-rate 16000 --clip-samples 56000 --clips 4320 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/k16_$m.json 2> gpurun_out/k16_$m.err || tail -3 gpurun_out/k16_$m.err
  python -c "
import json
d=json.load(open('gpurun_out/k16_$m.json')); print('$m', 'ms/step', round(d['ms_per_step'],3), 'audio-s/s', round(d['value']), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items() if v})"
done
