# usage: gpurun --gpus 8 -- bash scripts/gpu_n8_trace.sh   (c2 and c3 under torchrun on 8 GPUs, per-step e2e trace)
mkdir -p gpurun_out
for cfg in c2 c3; do
  SERB_BENCH_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 8 --steps 5 --warmup 3 --config $cfg > gpurun_out/n8_$cfg.json 2> gpurun_out/n8_$cfg.err || tail -5 gpurun_out/n8_$cfg.err
  grep "\[trace\] rank [07]" gpurun_out/n8_$cfg.err | head -16
  python - <<PY
import json
d=json.load(open("gpurun_out/n8_$cfg.json"))
e=d["e2e"]
print("$cfg: value %.0f ms/step %.2f | e2e %.0f ms %.2f chain %.2f | pageable ms %.2f" % (d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], e["device_chain_ms"], e["pageable"]["ms_per_step"]))
PY
done
