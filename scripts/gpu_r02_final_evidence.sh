# usage: gpurun --timeout 1200 -- bash scripts/gpu_r02_final_evidence.sh
# Evidence of the final round-2 build: plain runs first (each must exit 0), then
#  1. the launch list of the DEFAULT bench command,
#  2. duration + DRAM bytes of every launch of one 288-clip step (chain DRAM bytes per STFT column),
#  3. ncu --set full of the two constant-Q launches and of the first decimation of the timed chain.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/lt_plain.json 2> gpurun_out/lt_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_default.csv $CMD > gpurun_out/lt_list.log 2>&1
echo "list exit $?"
CMD2="python bench.py --clips 288 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD2 > gpurun_out/lt_plain2.json 2> gpurun_out/lt_plain2.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:^(cqtc_kernel|cqt_kernel|decimate|stft_kernel|hpss_|istft|ola_kernel|proj_kernel|tuning|tune_long|tonnetz|pool_kernel|mlp_kernel|expand_tiles)' -c 600 --csv --log-file gpurun_out/chain_traffic.csv $CMD2 > gpurun_out/lt_traffic.log 2>&1
echo "traffic exit $?"
ncu --set full --clock-control none --import-source on -k regex:cqtc_kernel -s 6 -c 2 -o gpurun_out/prof_cqt16 -f $CMD2 > gpurun_out/ncu_full_cqt16.log 2>&1
echo "cqt16 exit $?"
ncu --set full --clock-control none --import-source on -k regex:decimate2_mma -s 21 -c 1 -o gpurun_out/prof_dec_mma -f $CMD2 > gpurun_out/ncu_full_dec.log 2>&1
echo "dec exit $?"
tail -n 2 gpurun_out/lt_list.log gpurun_out/lt_traffic.log gpurun_out/ncu_full_cqt16.log gpurun_out/ncu_full_dec.log
