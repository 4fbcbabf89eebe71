# e2e sensitivity to the host-buffer chunk ramp: bash scripts/gpu_ramp_sweep.sh (on the GPU box)
# prints tag, device-resident ms per step, e2e ms per step, device chain ms inside the e2e call
run() { tag=$1; shift; env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ramp_$tag.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/ramp_$tag.json')); print('$tag', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), round(d['e2e']['device_chain_ms'],2))"; }
run s32_f30 SERB_RAMP_START=32768 SERB_RAMP_FACTOR_X10=30
run s32_f40 SERB_RAMP_START=32768 SERB_RAMP_FACTOR_X10=40
run s48_f30 SERB_RAMP_START=49152 SERB_RAMP_FACTOR_X10=30
run s24_f30 SERB_RAMP_START=24576 SERB_RAMP_FACTOR_X10=30
run s32_f30b SERB_RAMP_START=32768 SERB_RAMP_FACTOR_X10=30
run s16_f30 SERB_RAMP_START=16384 SERB_RAMP_FACTOR_X10=30
