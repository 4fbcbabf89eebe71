#!/bin/bash
# e2e sensitivity of the PCM16 host entry to the chunk ramp (first chunk size x growth factor)
for cfg in "32768 30" "65536 30" "65536 40" "131072 30" "131072 40" "262144 30" "262144 20" "16384 40"; do
  set -- $cfg
  SERB_RAMP_START=$1 SERB_RAMP_FACTOR_X10=$2 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ramp16_$1_$2.json 2>/dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/ramp16_$1_$2.json"))
e = d["e2e"]
print("start $1 factor $2: e2e %.2f ms (chain %.2f), resident %.2f ms" % (e["ms_per_step"], e["device_chain_ms"], d["ms_per_step"]))
PY
done
