#!/bin/bash
# e2e sensitivity of the PCM16 host entry to the chunk ramp with the first piece prefetched (final round-2 build)
for cfg in "65536 40" "49152 40" "81920 40" "98304 40" "65536 50" "65536 30" "81920 50" "98304 30" "65536 40"; do
  set -- $cfg
  SERB_RAMP_START=$1 SERB_RAMP_FACTOR_X10=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/rampp_$1_$2.json 2>/dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/rampp_$1_$2.json"))
e = d["e2e"]
print("start $1 factor $2: e2e %.2f ms (chain %.2f), resident %.2f ms" % (e["ms_per_step"], e["device_chain_ms"], d["ms_per_step"]))
PY
done
