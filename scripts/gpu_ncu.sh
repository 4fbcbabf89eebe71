# usage: bash scripts/gpu_ncu.sh <kernel-regex> [extra bench args]
# plain run first (must exit 0), then the launch list, then one --set full capture of the named kernel
mkdir -p gpurun_out
KREGEX=${1:-cqt_kernel}
shift
CMD="python bench.py --clips 288 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline $*"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 8 -c 3 -o gpurun_out/prof_$KREGEX -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "exit $?"
cat gpurun_out/ncu_plain.json
tail -n 3 gpurun_out/ncu_list.log; tail -n 3 gpurun_out/ncu_full.log
