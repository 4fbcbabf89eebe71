# usage: gpurun --gpus 8 -- bash scripts/gpu_n8_numa_ab.sh   (c2 under torchrun on 8 GPUs, host binding on / off)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/n8_topo.txt 2>&1
lscpu | grep -i "numa\|socket\|^CPU(s)" >> gpurun_out/n8_topo.txt
for bind in 1 0; do
  SERB_NUMA_BIND=$bind python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/n8_bind$bind.json 2> gpurun_out/n8_bind$bind.err || tail -5 gpurun_out/n8_bind$bind.err
  python - <<PY
import json
d=json.load(open("gpurun_out/n8_bind$bind.json"))
e=d["e2e"]
print("bind $bind: value %.0f ms/step %.2f | e2e %.0f ms %.2f chain %.2f | pageable ms %.2f | binding %s" % (d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], e["device_chain_ms"], e["pageable"]["ms_per_step"], e.get("host_binding")))
PY
done
