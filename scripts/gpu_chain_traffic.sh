# usage: bash scripts/gpu_chain_traffic.sh
# duration + DRAM bytes of every libser_b200 launch of a 288-clip step (chain DRAM bytes per STFT column)
mkdir -p gpurun_out
CMD2="python bench.py --clips 288 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD2 > gpurun_out/lt_plain2.json 2> gpurun_out/lt_plain2.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k 'regex:^(cqtc_kernel|cqt_kernel|decimate|stft_kernel|hpss_|istft|ola_kernel|proj_kernel|tuning|tune_long|tonnetz|pool_kernel|mlp_kernel|expand_tiles)' \
    -c 600 --csv --log-file gpurun_out/chain_traffic.csv $CMD2 > gpurun_out/lt_traffic.log 2>&1
echo "traffic exit $?"
