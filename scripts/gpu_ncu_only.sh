# usage: bash scripts/gpu_ncu_only.sh <kernel-regex> [skip] [extra bench args]
# plain run first (must exit 0), then one --set full capture of the named kernel
mkdir -p gpurun_out
KREGEX=${1:-cqt_kernel}
SKIP=${2:-4}
shift; shift
CMD="python bench.py --clips 288 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline $*"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c 1 -o gpurun_out/prof_$KREGEX -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "exit $?"
tail -n 3 gpurun_out/ncu_full.log
