# usage: bash scripts/gpu_launches_and_traffic.sh
# plain run (must exit 0), then the launch list of the DEFAULT bench command (profiles/*_launches*.md) and one
# pass of duration + DRAM bytes over every launch of a 288-clip step (chain DRAM bytes per STFT column)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/lt_plain.json 2> gpurun_out/lt_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_default.csv $CMD > gpurun_out/lt_list.log 2>&1
echo "list exit $?"
CMD2="python bench.py --clips 288 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD2 > gpurun_out/lt_plain2.json 2> gpurun_out/lt_plain2.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:serb -c 400 --csv --log-file gpurun_out/chain_traffic.csv $CMD2 > gpurun_out/lt_traffic.log 2>&1
echo "traffic exit $?"
tail -n 2 gpurun_out/lt_list.log gpurun_out/lt_traffic.log
