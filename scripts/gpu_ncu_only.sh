# usage: bash scripts/gpu_ncu_only.sh <kernel-regex> <skip> <count> <tag> [bench args]
mkdir -p gpurun_out
KREGEX=$1; SKIP=$2; COUNT=$3; TAG=$4; shift 4
CMD="python bench.py --clips 288 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline $*"
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c $COUNT -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu_full_$TAG.log
